"""b2_xattn_q_core (csrc/xattn.cu): cross-attention query projection fused with the 77-key attention core, against
  (a) the un-fused kernels it replaces (b2_gemm for to_q + b2_attn_fwd), and
  (b) torch fp32: diffusers Attention.forward of attn2 = to_q Linear -> F.scaled_dot_product_attention over the text tokens.
Tolerances (stated): Q is the same CTA-pair GEMM -> bit-identical to b2_gemm; O / LSE as in tests/test_gpu_attention.py
(O rel-L2 <= 1e-2 vs fp32, LSE |d| <= 2e-2); fused vs un-fused O rel-L2 <= 2e-3 (same math, P rounded to bf16 in both)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rel(a, b):
    a = a.float().flatten(); b = b.float().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("B,n_q,n_k,Cc,L", [
    (4, 1024, 77, 1280, 3),   # SDXL 1024 px, level 2: 64 tiles = one per CTA pair
    (2, 256, 77, 640, 1),     # 4 tiles
    (4, 4096, 77, 640, 2),    # SDXL 1024 px, level 1: 128 tiles on 74 pairs -> two tiles per pair (stage / TMEM hand-back)
    (1, 512, 40, 320, 1),     # fewer keys than one CTA's half; one head group
    (3, 768, 80, 960, 1),     # three 320-column tiles, odd batch, n_k at the limit
])
def test_xattn_q_core_matches_unfused_and_torch(B, n_q, n_k, Cc, L):
    from sdxl_training_improvements_b200 import ops
    assert ops.xattn_q_core_ok(B, n_q, n_k, Cc)
    H = Cc // 64
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + n_q + Cc)
    xn = torch.randn(B * n_q, Cc, device="cuda", generator=g).to(bf16)
    Wq = (torch.randn(Cc, Cc, device="cuda", generator=g) * Cc ** -0.5).to(bf16)
    # K / V as column slices of a grouped projection buffer [B*n_k, L*2C] (what UNetEngine._project_context produces)
    kv_all = torch.randn(B * n_k, L * 2 * Cc, device="cuda", generator=g).to(bf16)
    l = L - 1
    k, v = kv_all[:, l * 2 * Cc:l * 2 * Cc + Cc], kv_all[:, l * 2 * Cc + Cc:(l + 1) * 2 * Cc]
    scale = 1.0 / math.sqrt(64)
    q, o, lse = ops.xattn_q_core(xn, Wq, k, v, B, n_q, n_k, scale)
    torch.cuda.synchronize()
    # (a) the kernels it replaces
    q2 = ops.linear_fwd(xn, Wq)
    o2, lse2 = ops.attn_fwd(q2, k, v, B, H, n_q, n_k, scale)
    assert torch.equal(q, q2), f"Q differs from b2_gemm in {int((q != q2).sum())} of {q.numel()} elements"
    assert _rel(o, o2) <= 2e-3, _rel(o, o2)
    assert float((lse[:, :, :n_q] - lse2[:, :, :n_q]).abs().max()) <= 1e-3
    # (b) torch fp32 on the same bf16 inputs (Q rounded to bf16 as both kernel paths do)
    qf = q2.float().view(B, n_q, H, 64).permute(0, 2, 1, 3)
    kf = k.float().reshape(B, n_k, H, 64).permute(0, 2, 1, 3)
    vf = v.float().reshape(B, n_k, H, 64).permute(0, 2, 1, 3)
    s = torch.einsum("bhid,bhjd->bhij", qf, kf) * scale
    ro = torch.einsum("bhij,bhjd->bhid", torch.softmax(s, -1), vf).permute(0, 2, 1, 3).reshape(B * n_q, Cc)
    assert _rel(o, ro) <= 1e-2, _rel(o, ro)
    rl = torch.logsumexp(s, -1) / math.log(2.0)
    assert float((lse[:, :, :n_q] - rl).abs().max()) <= 2e-2 * max(1.0, float(rl.abs().max()) * 0.05)
    # determinism
    q3, o3, lse3 = ops.xattn_q_core(xn, Wq, k, v, B, n_q, n_k, scale)
    assert torch.equal(o, o3) and torch.equal(lse, lse3)


def test_xattn_q_core_speed_report():
    """Prints the north-star figures: fused Q-projection + core vs the two launches it replaces, and the whole block
    LN -> [Wq + core] -> Wo + bias + residual (28.46 GFLOP at C = 1280, n = 1024, B = 4; target 0.6 of burst = 28.6 us)."""
    from sdxl_training_improvements_b200 import ops
    for (B, n, Cc) in ((4, 1024, 1280), (4, 4096, 640)):
        H, nk = Cc // 64, 77
        x = torch.randn(B * n, Cc, device="cuda").to(bf16)
        gamma, beta = torch.ones(Cc, device="cuda", dtype=bf16), torch.zeros(Cc, device="cuda", dtype=bf16)
        Wq = (torch.randn(Cc, Cc, device="cuda") * 0.02).to(bf16)
        Wo = (torch.randn(Cc, Cc, device="cuda") * 0.02).to(bf16)
        bo = torch.zeros(Cc, device="cuda", dtype=bf16)
        kv = torch.randn(B * nk, 2 * Cc, device="cuda").to(bf16)
        k, v = kv[:, :Cc], kv[:, Cc:]
        yb = torch.empty_like(x)

        def fused():
            xn, _, _ = ops.ln_fwd(x, gamma, beta, 1e-5)
            q, o, _ = ops.xattn_q_core(xn, Wq, k, v, B, n, nk, 0.125)
            ops.linear_fwd(o, Wo, bias=bo, residual=x, out=yb)

        def unfused():
            xn, _, _ = ops.ln_fwd(x, gamma, beta, 1e-5)
            q = ops.linear_fwd(xn, Wq)
            o, _ = ops.attn_fwd(q, k, v, B, H, n, nk, 0.125)
            ops.linear_fwd(o, Wo, bias=bo, residual=x, out=yb)

        def time_graph(fn, iters=20):
            fn(); fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(iters):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / iters

        xn, _, _ = ops.ln_fwd(x, gamma, beta, 1e-5)
        t_k1 = time_graph(lambda: ops.xattn_q_core(xn, Wq, k, v, B, n, nk, 0.125))
        t_f, t_u = time_graph(fused), time_graph(unfused)
        flops = 2 * (2.0 * B * n * Cc * Cc) + 4.0 * n * nk * 64 * B * H
        print(f"\nxattn C={Cc} n={n} B={B}: [Wq + core] fused {t_k1:.1f} us; block LN->Wq->core->Wo+res: fused {t_f:.1f} us = "
              f"{flops / t_f / 1e6:.0f} TFLOP/s, un-fused {t_u:.1f} us = {flops / t_u / 1e6:.0f} TFLOP/s ({flops / 1e9:.2f} GFLOP)")
