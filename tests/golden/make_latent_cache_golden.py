"""Write tests/golden/latent_cache/ with the REFERENCE's own CacheManager.save_latents
(src/data/preprocessing/cache_manager.py:289-402) and record what its load_tensors returns
(tests/golden/latent_cache_expected.pt).  Build container only (needs /root/reference):
    python tests/golden/make_latent_cache_golden.py
Tiny tensors (latent 4x8x8 / 4x8x12, 77x16 embeddings) keep the fixture at a few tens of KB."""
import os
import shutil
import sys
import tempfile
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_schedule_golden import REF, _stub_modules  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    from src.data.preprocessing.cache_manager import CacheManager
    os.chdir(cwd)
    out = os.path.join(HERE, "latent_cache")
    shutil.rmtree(out, ignore_errors=True)
    cm = CacheManager(out, device=torch.device("cpu"))
    g = torch.Generator().manual_seed(0)
    expected = {}
    items = [("/data/img_000.png", (8, 8), "a cat"), ("/data/img_001.png", (8, 8), "a dog, masterpiece"),
             ("/data/img_002.png", (8, 12), "wide landscape"), ("/data/img_003.png", (8, 12), "wide city"),
             ("/data/img_004.png", (8, 8), "a bird")]
    for path, (h, w), text in items:
        tensors = {"vae_latents": torch.randn(4, h, w, generator=g), "time_ids": torch.tensor([[8. * w, 8. * h, 0, 0, 8. * w, 8. * h]]),
                   "prompt_embeds": torch.randn(77, 16, generator=g), "pooled_prompt_embeds": torch.randn(12, generator=g)}
        bucket = types.SimpleNamespace(
            dimensions=types.SimpleNamespace(width=8 * w, height=8 * h, width_latent=w, height_latent=h, aspect_ratio=w / h,
                                             aspect_ratio_inverse=h / w, total_pixels=64 * w * h, total_latents=w * h),
            pixel_dims=(8 * w, 8 * h), latent_dims=(w, h), bucket_index=0 if w == h else 1, size_class="tiny",
            aspect_class="square" if w == h else "landscape")
        tag_info = {"tags": {"subject": [{"tag": text.split()[1], "weight": 1.25}], "style": [], "quality": [], "technical": [], "meta": []}} \
            if "dog" in text else None
        assert cm.save_latents(tensors, path, {"text": text}, bucket_info=bucket, tag_info=tag_info)
        key = cm.get_cache_key(path)
        expected[path] = {"key": key, "loaded": cm.load_tensors(key)}
    # drop files the reader does not need (absolute build-container paths inside are irrelevant to the format)
    torch.save(expected, os.path.join(HERE, "latent_cache_expected.pt"))
    print("wrote", out, sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(out) for f in fs), "bytes")


if __name__ == "__main__":
    main()
