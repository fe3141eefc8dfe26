"""the latent-cache reader and the rank-aware bucket sampler (CPU).

tests/golden/latent_cache/ was WRITTEN by the reference's own `CacheManager.save_latents`
(src/data/preprocessing/cache_manager.py:289-402) and tests/golden/latent_cache_expected.pt holds what the reference's
`load_tensors` (:404-509) returned for every entry (tests/golden/make_latent_cache_golden.py): the reader here must
return the same tensors and metadata, bit for bit."""
import os

import pytest
import torch

from sdxl_training_improvements_b200.data import BucketBatchSampler, LatentCacheDataset, LatentCacheReader, get_cache_key

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(HERE, "golden", "latent_cache")
EXPECTED = os.path.join(HERE, "golden", "latent_cache_expected.pt")


def test_reader_matches_reference_load_tensors():
    exp = torch.load(EXPECTED, weights_only=False)
    rd = LatentCacheReader(CACHE)
    assert set(rd.keys()) == {v["key"] for v in exp.values()}
    for path, rec in exp.items():
        assert get_cache_key(path) == rec["key"]                 # md5 of the path string (cache_manager.py:726-729)
        for handle in (rec["key"], path):                         # index key, or the image path (defect B18)
            got = rd.load_tensors(handle)
            ref = rec["loaded"]
            assert set(got) == set(ref)
            for k in ("vae_latents", "prompt_embeds", "pooled_prompt_embeds", "time_ids"):
                assert torch.equal(got[k], ref[k]), (path, k)
            assert got["metadata"] == ref["metadata"], path
    with pytest.raises(RuntimeError, match="Cache entry not found"):
        rd.load_tensors("/data/not_cached.png")


def test_dataset_batches_follow_the_plugin_batch_schema():
    ds = LatentCacheDataset(CACHE)
    assert sorted(ds.bucket_indices) == [(8, 8), (8, 12)]
    assert sorted(len(v) for v in ds.bucket_indices.values()) == [2, 3]
    sampler = BucketBatchSampler(ds.bucket_indices, batch_size=2, drop_last=True, shuffle=False)
    batches = list(sampler)
    assert len(batches) == 2                                      # 3 -> one batch of 2 (incomplete dropped), 2 -> one
    for idx in batches:
        b = ds.collate([ds[i] for i in idx])
        h, w = b["vae_latents"].shape[-2:]
        assert b["vae_latents"].shape == (2, 4, h, w) and b["prompt_embeds"].shape == (2, 77, 16)
        assert b["pooled_prompt_embeds"].shape == (2, 12) and b["time_ids"].shape == (2, 1, 6)
        assert b["time_ids"][0, 0].tolist() == [8. * w, 8. * h, 0, 0, 8. * w, 8. * h]
        assert isinstance(b["metadata"], list) and {"text", "bucket_info", "tag_info"} <= set(b["metadata"][0])
        assert len({tuple(ds[i]["vae_latents"].shape) for i in idx}) == 1      # one bucket per batch


def test_rank_sharded_sampler():
    buckets = {(8, 8): list(range(0, 40)), (8, 12): list(range(40, 58)), (12, 8): list(range(58, 61))}
    world, bs = 4, 2
    per_rank = []
    for r in range(world):
        s = BucketBatchSampler(buckets, bs, drop_last=True, shuffle=True, rank=r, world_size=world, seed=3)
        s.set_epoch(1)
        per_rank.append(list(s))
    n = len(per_rank[0])
    assert all(len(p) == n for p in per_rank) and n == 40 // (bs * world) + 18 // (bs * world)   # 5 + 2 global steps
    inv = {i: k for k, v in buckets.items() for i in v}
    seen = set()
    for step in range(n):
        shapes = {inv[i] for r in range(world) for i in per_rank[r][step]}
        assert len(shapes) == 1, "every rank must run the same latent shape in a given optimizer step"
        for r in range(world):
            assert not (set(per_rank[r][step]) & seen), "ranks must see disjoint samples"
            seen |= set(per_rank[r][step])
    s0 = BucketBatchSampler(buckets, bs, rank=0, world_size=world, seed=3)
    s0.set_epoch(2)
    assert list(s0) != per_rank[0]                                 # reshuffled per epoch
    # world_size 1 == the reference sampler's batch set (samplers.py:28-58)
    ref_batches = []
    for shape, idx in buckets.items():
        if len(idx) < bs:
            continue
        ch = [idx[i:i + bs] for i in range(0, len(idx), bs)]
        ref_batches += ch[:-1] if len(ch[-1]) < bs else ch
    one = BucketBatchSampler(buckets, bs, drop_last=True, shuffle=False)
    assert sorted(map(tuple, one)) == sorted(map(tuple, ref_batches))
    with pytest.raises(ValueError):
        BucketBatchSampler({(8, 8): [0]}, 2)
