"""The data-parallel chunk plan derived from the REAL backward tape of the SDXL-base UNet, on CPU.

The engine's forward / backward are run with every kernel call replaced by a shape-only stand-in on the `meta` device (no
memory, no arithmetic): what is exercised is the tape itself — which closure writes which gradient view, in which order —
i.e. exactly the log `dp.plan_chunks` consumes on the GPU.  Pins:
  * every one of the 1,680 parameters gets its gradient written by some tape entry (a gradient view taken at FORWARD time
    would be invisible to the log — the bug the first 2-GPU run caught — and makes plan_chunks refuse);
  * the plan the 8-GPU run printed (profiles/r1_dpx_transports_n4.txt): 873 tape entries, first cut at 167, 11 chunks, the
    stacked cross-attention K/V weight gradients final well before the end, tail chunk under total/128;
  * the plan does not depend on the latent shape (one plan serves every aspect-ratio bucket).
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sdxl_training_improvements_b200 import dp, unet as unet_mod  # noqa: E402
from _shape_only_ops import ShapeOnlyOps  # noqa: E402
from sdxl_training_improvements_b200.params import SDXL_BASE, ParamStore  # noqa: E402

bf16 = torch.bfloat16


def _real_tape_log(H, W, B=1):
    store = ParamStore(SDXL_BASE, device="meta")
    store.flush_small_grads = lambda: None
    eng = unet_mod.UNetEngine(store)
    x = torch.empty(B * H * W, 8, device="meta", dtype=bf16)
    t = torch.empty(B, device="meta", dtype=torch.float32)
    ctx = torch.empty(B * 77, SDXL_BASE["cross_attention_dim"], device="meta", dtype=bf16)
    pooled = torch.empty(B, 1280, device="meta", dtype=bf16)
    tid = torch.empty(B, 6, device="meta", dtype=torch.float32)
    eng.forward(x, t, ctx, pooled, tid, B, H, W)
    saved = eng.detach_tape()
    n_tape = len(saved[0])
    store.touch_log = []
    eng.backward(torch.empty(B * H * W, 64, device="meta", dtype=bf16), saved)
    log, store.touch_log = store.touch_log, None
    return store, log, n_tape


@pytest.fixture()
def shape_only(monkeypatch):
    monkeypatch.setattr(unet_mod, "ops", ShapeOnlyOps())


def test_sdxl_plan_from_the_real_tape(shape_only):
    store, log, n_tape = _real_tape_log(128, 128)
    assert n_tape == 873
    last = dp.last_touch_positions(store, log)
    assert len(last) == 1680 and min(last.values()) >= 0, "a parameter gradient is never written on the tape"
    plan = dp.plan_chunks(store, log, n_tape)
    assert plan.n_chunks == 11 and plan.cuts[0] == 167 and plan.cuts[-1] == n_tape - 1
    sizes = [sum(n for _, n in rg) for rg in plan.ranges]
    assert sum(sizes) == store.total
    assert sizes[-1] <= store.total // 128, "tail chunk (the only exchange that cannot hide) too large"
    assert max(sizes) <= store.total // 6
    # the stacked cross-attention K/V weight gradients (0.73 GB) no longer wait for the end of the backward pass
    for Cc, pfxs in store.kv_groups.items():
        k = plan.param_chunk[pfxs[0] + ".to_k.weight"]
        assert all(plan.param_chunk[p + s] == k for p in pfxs for s in (".to_k.weight", ".to_v.weight"))
        assert k < plan.n_chunks - 1
    # the time / added-condition embeddings collect gradient from every resnet: they are final last
    assert plan.param_chunk["time_embedding.linear_1.weight"] == plan.n_chunks - 1
    # cuts are ascending tape positions and every parameter's last write is at or before its chunk's cut
    assert plan.cuts == sorted(set(plan.cuts))
    assert all(plan.cuts[plan.param_chunk[n]] >= p for n, p in last.items())
    # pieces per chunk stay few (copy-engine transfers are per piece per peer)
    assert max(len(rg) for rg in plan.ranges) <= 6


def test_plan_is_the_same_for_every_bucket(shape_only):
    store_a, log_a, n_a = _real_tape_log(128, 128)
    store_b, log_b, n_b = _real_tape_log(96, 96)
    store_c, log_c, n_c = _real_tape_log(120, 160, B=2)
    pa, pb, pc = (dp.plan_chunks(s, lg, n) for s, lg, n in ((store_a, log_a, n_a), (store_b, log_b, n_b), (store_c, log_c, n_c)))
    assert n_a == n_b == n_c
    assert pa.cuts == pb.cuts == pc.cuts and pa.ranges == pb.ranges == pc.ranges and pa.small_segs == pb.small_segs
