"""Fused attention kernels (b2_attn_fwd / b2_attn_bwd) vs torch fp32 attention + autograd on the same bf16 inputs.

Tolerances (stated): O is one bf16 rounding of an fp32 result whose P operand was rounded to bf16 before the PV MMA
(as every flash kernel does): rel-L2 <= 1e-2, max-abs <= 3e-2 on O(1) outputs.  Gradients: rel-L2 <= 2e-2.
Shapes cover SDXL's (n=1024/4096 keys self, 77 keys cross) and the ragged bucket sizes (576, 1200 tokens).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rel(a, b):
    a = a.float().flatten(); b = b.float().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _ref(q, k, v, scale):
    s = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    return torch.einsum("bhij,bhjd->bhid", torch.softmax(s, -1), v)


def _heads(x, B, n, H):  # [B*n, H*64] -> [B,H,n,64] fp32
    return x.float().view(B, n, H, 64).permute(0, 2, 1, 3)


CASES = [
    # B, H, n_q, n_k, fused-qkv?, logit gain
    (2, 5, 256, 256, True, 1.0),
    (1, 3, 1024, 1024, True, 1.0),
    (1, 2, 4096, 4096, True, 1.0),
    (2, 4, 1024, 77, False, 1.0),     # cross-attention
    (1, 2, 576, 576, True, 1.0),      # 768^2 bucket: 4.5 query tiles, ragged key block
    (1, 2, 1200, 1200, True, 1.0),    # 1280x960 bucket at 64x: ragged
    (1, 2, 200, 77, False, 1.0),
    (1, 2, 512, 512, True, 6.0),      # peaky logits: exercises the lazy rescale path
    (3, 20, 1024, 1024, True, 1.0),   # 480 tiles on 148 SMs: 3-4 tiles per persistent CTA, tile boundaries inside a head
    (2, 3, 1536, 1536, True, 5.0),    # 12 key blocks, peaky logits, several tiles per CTA: rescale across the deferred merge
    (1, 1, 300, 300, True, 1.0),      # 3 tiles, ragged query AND key tails, fewer tiles than SMs
]


def _select_variant(variant, monkeypatch, n_k=1024):
    """p2: the production forward kernel (attn_pfwd2_kernel: persistent CTA per SM, deferred two-group merge).
    v3: the round-1 one-CTA-per-query-tile forward kernel, kept selectable (B2_ATTN_FWD3=1) for A/B timing.
    The backward kernels and the cross-attention kernel (n_k <= 96) have no variants."""
    monkeypatch.delenv("B2_ATTN_FWD3", raising=False)
    if variant == "v3":
        if n_k <= 96:
            pytest.skip("the cross-attention kernel has no variants")
        monkeypatch.setenv("B2_ATTN_FWD3", "1")


@pytest.mark.parametrize("variant", ["p2", "v3"])
@pytest.mark.parametrize("B,H,n_q,n_k,fused,gain", CASES)
def test_attention_fwd_bwd(B, H, n_q, n_k, fused, gain, variant, monkeypatch):
    _select_variant(variant, monkeypatch, n_k)
    from sdxl_training_improvements_b200 import ops
    Cc = H * 64
    g = torch.Generator(device="cuda").manual_seed(n_q * 7 + n_k)
    scale = 1.0 / math.sqrt(64)
    if fused:
        qkv = (torch.randn(B * n_q, 3 * Cc, device="cuda", generator=g) * gain).to(bf16)
        q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
        dqkv = torch.full_like(qkv, float("nan"))
        dq, dk, dv = dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:]
    else:
        q = (torch.randn(B * n_q, Cc, device="cuda", generator=g) * gain).to(bf16)
        kv = (torch.randn(B * n_k, 2 * Cc, device="cuda", generator=g) * gain).to(bf16)
        k, v = kv[:, :Cc], kv[:, Cc:]
        dq = torch.full_like(q, float("nan"))
        dkv = torch.full_like(kv, float("nan"))
        dk, dv = dkv[:, :Cc], dkv[:, Cc:]
    o, lse = ops.attn_fwd(q, k, v, B, H, n_q, n_k, scale)
    qf = _heads(q, B, n_q, H).requires_grad_(True)
    kf = _heads(k, B, n_k, H).requires_grad_(True)
    vf = _heads(v, B, n_k, H).requires_grad_(True)
    ro = _ref(qf, kf, vf, scale)
    assert torch.isfinite(o.float()).all()
    e = _rel(_heads(o, B, n_q, H), ro)
    assert e <= 1e-2, f"fwd rel-L2 {e}"
    # LSE (log2 domain) vs reference logsumexp
    rl = torch.logsumexp(torch.einsum("bhid,bhjd->bhij", qf, kf) * scale, -1) / math.log(2.0)
    assert float((lse[:, :, :n_q] - rl).abs().max()) <= 2e-2 * max(1.0, float(rl.abs().max()) * 0.05)
    assert torch.isinf(lse[:, :, n_q:]).all()

    do = torch.randn(B * n_q, Cc, device="cuda", generator=g).to(bf16)
    ops.attn_bwd(q, k, v, o, lse, do, dq, dk, dv, B, H, n_q, n_k, scale)
    ro.backward(_heads(do, B, n_q, H))
    for name, got, ref, n in (("dQ", dq, qf.grad, n_q), ("dK", dk, kf.grad, n_k), ("dV", dv, vf.grad, n_k)):
        assert torch.isfinite(got.float()).all(), name
        e = _rel(_heads(got, B, n, H), ref)
        assert e <= 2e-2, f"{name} rel-L2 {e}"
    # determinism: the kernels have no atomics and a fixed block order -> a second run is bit-identical
    dq2, dk2, dv2 = torch.empty_like(dq.contiguous()), torch.empty_like(dk.contiguous()), torch.empty_like(dv.contiguous())
    ops.attn_bwd(q, k, v, o, lse, do, dq2, dk2, dv2, B, H, n_q, n_k, scale)
    assert torch.equal(dq2, dq.contiguous()) and torch.equal(dk2, dk.contiguous()) and torch.equal(dv2, dv.contiguous())
    torch.cuda.synchronize()


@pytest.mark.parametrize("variant", ["p2", "v3"])
def test_attention_speed_report(variant, monkeypatch):
    """Not an assertion on speed — prints achieved TFLOP/s of the three kernels for the bench log."""
    from sdxl_training_improvements_b200 import ops
    _select_variant(variant, monkeypatch)
    poly = variant
    if variant == "p2":  # cross-attention core (77 keys), forward + backward, once
        for (B, H, n, nk) in ((4, 20, 1024, 77), (4, 10, 4096, 77)):
            Cc = H * 64
            q = torch.randn(B * n, Cc, device="cuda").to(bf16)
            kv = torch.randn(B * nk, 2 * Cc, device="cuda").to(bf16)
            k, v = kv[:, :Cc], kv[:, Cc:]
            do = torch.randn(B * n, Cc, device="cuda").to(bf16)
            dq, dkv = torch.empty_like(q), torch.empty_like(kv)
            o, lse = ops.attn_fwd(q, k, v, B, H, n, nk, 0.125)
            ops.attn_bwd(q, k, v, o, lse, do, dq, dkv[:, :Cc], dkv[:, Cc:], B, H, n, nk, 0.125)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            it = 20
            e[0].record()
            for _ in range(it):
                ops.attn_fwd(q, k, v, B, H, n, nk, 0.125, out=o)
            e[1].record()
            for _ in range(it):
                ops.attn_bwd(q, k, v, o, lse, do, dq, dkv[:, :Cc], dkv[:, Cc:], B, H, n, nk, 0.125)
            e[2].record()
            torch.cuda.synchronize()
            print(f"\ncross attn B={B} H={H} n={n} nk={nk}: fwd {e[0].elapsed_time(e[1]) / it * 1e3:.1f} us, "
                  f"bwd {e[1].elapsed_time(e[2]) / it * 1e3:.1f} us (stream launches, not graph replay)")
    for (B, H, n) in ((4, 20, 1024), (4, 10, 4096)):
        Cc = H * 64
        qkv = torch.randn(B * n, 3 * Cc, device="cuda").to(bf16)
        q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
        dqkv = torch.empty_like(qkv)
        do = torch.randn(B * n, Cc, device="cuda").to(bf16)
        scale = 0.125
        o, lse = ops.attn_fwd(q, k, v, B, H, n, n, scale)
        ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:], B, H, n, n, scale)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        it = 10
        e[0].record()
        for _ in range(it):
            ops.attn_fwd(q, k, v, B, H, n, n, scale, out=o)
        e[1].record()
        for _ in range(it):
            ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:], B, H, n, n, scale)
        e[2].record()
        torch.cuda.synchronize()
        f = 4.0 * B * H * n * n * 64
        tf, tb = e[0].elapsed_time(e[1]) / it, e[1].elapsed_time(e[2]) / it
        print(f"\nattn fwd variant={poly} B={B} H={H} n={n}: fwd {tf * 1e3:.0f} us = {f / tf / 1e9:.0f} TFLOP/s; "
              f"bwd {tb * 1e3:.0f} us = {2.5 * f / tb / 1e9:.0f} TFLOP/s (5-GEMM algorithmic flops)")
