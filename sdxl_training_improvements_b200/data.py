"""Latent-cache reader and rank-aware bucket sampler — the step either side of the hot path.

The reference pre-encodes every image once (VAE latents + dual-CLIP embeddings) and stores the result on disk
(src/data/preprocessing/cache_manager.py:289-402); training then only has to turn cache entries into the batch dict the
method plugins consume (src/data/dataset.py:209-229).  This module reads that on-disk format as the reference writes it
and builds the batches — nothing here touches the GPU kernels:

  <cache_dir>/cache_index.json                      zlib-compressed JSON (plain JSON accepted, :652-668), "entries":
                                                    {md5(image path): {vae_latent_path, clip_latent_path, metadata_path,
                                                     is_valid, bucket_info, tag_info, ...}} with paths relative to latents/
  <cache_dir>/latents/vae/<key>.pt                  torch.save({"vae_latents": [4,h,w], "time_ids": [1,6]})
  <cache_dir>/latents/clip/<key>.pt                 torch.save({"prompt_embeds": [77,2048], "pooled_prompt_embeds": [1280]})
  <cache_dir>/latents/metadata/<key>.json           {"text", "bucket_info", ...}

`LatentCacheReader.load_tensors` returns exactly what `CacheManager.load_tensors` returns (:404-509; pinned by
tests/test_data_cache.py against a cache written by the reference's own `save_latents`).  Two reference defects are
handled as SURVEY Appendix B decides: B18 (the dataset passes the image PATH where the index is keyed by md5 — both are
accepted here) and the missing rank sharding (samplers.py:28-58 is rank-agnostic, every rank would see the same
batches): `BucketBatchSampler` shards whole same-bucket global steps across ranks from a rank-shared RNG, so that all
ranks run the same latent shape in the same optimizer step (no straggler skew under the single all-reduce).
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import zlib
from pathlib import Path
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import torch

_DEFAULT_TAG_INFO = {"tags": {"subject": [], "style": [], "quality": [], "technical": [], "meta": []}}


def get_cache_key(path: Union[str, Path]) -> str:
    """cache_manager.py:726-729 — md5 of the (POSIX) path string."""
    return hashlib.md5(str(Path(path)).encode()).hexdigest()


class LatentCacheReader:
    def __init__(self, cache_dir: Union[str, Path], device: Union[str, torch.device] = "cpu"):
        self.cache_dir = Path(cache_dir)
        self.latents_dir = self.cache_dir / "latents"
        self.index_path = self.cache_dir / "cache_index.json"
        self.device = device
        self.cache_index = self._load_cache_index()

    def _load_cache_index(self) -> Dict[str, Any]:
        if not self.index_path.exists():
            raise FileNotFoundError(f"no cache index at {self.index_path}")
        raw = self.index_path.read_bytes()
        try:
            return json.loads(zlib.decompress(raw))
        except zlib.error:  # old uncompressed format (cache_manager.py:660-666)
            return json.loads(raw.decode("utf-8"))

    @property
    def entries(self) -> Dict[str, Dict[str, Any]]:
        return self.cache_index.get("entries", {})

    def keys(self, valid_only: bool = True) -> List[str]:
        return [k for k, e in self.entries.items() if e.get("is_valid", True) or not valid_only]

    def resolve_key(self, key_or_path: Union[str, Path]) -> str:
        k = str(key_or_path)
        if k in self.entries:
            return k
        h = get_cache_key(k)  # B18: callers hand over the image path
        if h in self.entries:
            return h
        raise RuntimeError(f"Cache entry not found for key: {k}")

    def load_tensors(self, key_or_path: Union[str, Path]) -> Dict[str, Any]:
        key = self.resolve_key(key_or_path)
        entry = self.entries[key]
        paths = {}
        for name, field in (("vae", "vae_latent_path"), ("clip", "clip_latent_path"), ("metadata", "metadata_path")):
            p = self.latents_dir / entry[field]
            if not p.exists():
                raise RuntimeError(f"File does not exist: {p}")
            if p.stat().st_size == 0:
                raise RuntimeError(f"File is empty: {p}")
            paths[name] = p
        vae = torch.load(paths["vae"], map_location=self.device, weights_only=True)
        clip = torch.load(paths["clip"], map_location=self.device, weights_only=True)
        for d, req, what in ((vae, ("vae_latents", "time_ids"), "VAE"), (clip, ("prompt_embeds", "pooled_prompt_embeds"), "CLIP")):
            missing = [k for k in req if k not in d]
            if missing:
                raise RuntimeError(f"Invalid {what} data structure. Missing keys: {missing}")
        meta = json.loads(paths["metadata"].read_text(encoding="utf-8"))
        missing = [k for k in ("text", "bucket_info") if k not in meta]
        if missing:
            raise RuntimeError(f"Invalid metadata structure. Missing keys: {missing}")
        return {
            "vae_latents": vae["vae_latents"],
            "prompt_embeds": clip["prompt_embeds"],
            "pooled_prompt_embeds": clip["pooled_prompt_embeds"],
            "time_ids": vae["time_ids"],
            "metadata": {"text": meta.get("text"), "bucket_info": entry.get("bucket_info"),
                         # as written (:480-489): the default applies only when the key is ABSENT; a stored null stays None
                         "tag_info": entry.get("tag_info", _DEFAULT_TAG_INFO)},
        }


class LatentCacheDataset(torch.utils.data.Dataset):
    """Cache entries -> samples; `bucket_indices` groups dataset indices by latent shape (h, w) for the sampler."""

    def __init__(self, cache_dir: Union[str, Path], pin_memory: bool = False):
        self.reader = LatentCacheReader(cache_dir)
        self.keys = sorted(self.reader.keys())
        self.pin_memory = pin_memory
        self.bucket_indices: Dict[Tuple[int, ...], List[int]] = {}
        for i, k in enumerate(self.keys):
            bi = self.reader.entries[k].get("bucket_info") or {}
            dims = bi.get("latent_dims")
            if dims is None:  # no bucket record: read the tensor header once
                dims = tuple(self.reader.load_tensors(k)["vae_latents"].shape[-2:])[::-1]
            w, h = int(dims[0]), int(dims[1])  # bucket_info stores (width, height)
            self.bucket_indices.setdefault((h, w), []).append(i)

    def __len__(self) -> int:
        return len(self.keys)

    def __getitem__(self, i: int) -> Dict[str, Any]:
        return self.reader.load_tensors(self.keys[i])

    def collate(self, samples: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
        """Batch dict of the method plugins (dataset.py:209-229): vae_latents [B,4,h,w], prompt_embeds [B,77,2048],
        pooled_prompt_embeds [B,1280], time_ids [B,1,6], metadata (list)."""
        def stack(key, ndim):
            ts = [s[key] for s in samples]
            ts = [t[0] if t.dim() == ndim + 1 and t.shape[0] == 1 else t for t in ts]  # tolerate a leading 1
            out = torch.stack(ts)
            return out.pin_memory() if self.pin_memory and torch.cuda.is_available() else out
        tid = torch.stack([s["time_ids"].reshape(1, -1) for s in samples])
        return {"vae_latents": stack("vae_latents", 3), "prompt_embeds": stack("prompt_embeds", 2),
                "pooled_prompt_embeds": stack("pooled_prompt_embeds", 1),
                "time_ids": tid.pin_memory() if self.pin_memory and torch.cuda.is_available() else tid,
                "metadata": [s["metadata"] for s in samples]}


class BucketBatchSampler:
    """Same batching rule as the reference sampler (samplers.py:8-61: per-bucket consecutive chunks of `batch_size`,
    incomplete chunks dropped when `drop_last`), plus what data parallelism needs: one GLOBAL step = `world_size` batches
    of the SAME bucket, shuffled identically on every rank (seed + epoch), rank r taking the r-th batch of each step."""

    def __init__(self, bucket_indices: Dict[Tuple[int, ...], List[int]], batch_size: int, drop_last: bool = True,
                 shuffle: bool = True, rank: int = 0, world_size: int = 1, seed: int = 0):
        if not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside world of {world_size}")
        self.bucket_indices, self.batch_size, self.drop_last, self.shuffle = bucket_indices, batch_size, drop_last, shuffle
        self.rank, self.world_size, self.seed, self.epoch = rank, world_size, seed, 0
        self.steps: List[List[List[int]]] = []
        for shape, indices in bucket_indices.items():
            if len(indices) < batch_size and drop_last:
                continue
            chunks = [indices[i:i + batch_size] for i in range(0, len(indices), batch_size)]
            if drop_last and len(chunks[-1]) < batch_size:
                chunks = chunks[:-1]
            # whole global steps only: every rank gets a batch of this bucket
            for j in range(0, len(chunks) - world_size + 1, world_size):
                self.steps.append(chunks[j:j + world_size])
        if not self.steps:
            raise ValueError("No valid batches created - check bucket sizes and batch size")

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __iter__(self) -> Iterator[List[int]]:
        order = list(range(len(self.steps)))
        if self.shuffle:
            random.Random(self.seed * 1000003 + self.epoch).shuffle(order)  # identical on every rank
        return iter([self.steps[i][self.rank] for i in order])

    def __len__(self) -> int:
        return len(self.steps)
