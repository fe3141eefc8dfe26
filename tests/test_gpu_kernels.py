"""Per-kernel parity (T2 of SURVEY.md §4.1): every C-ABI entry vs the equivalent torch op on the same inputs.

Tolerances: inputs/outputs are bf16, math fp32.  A GEMM output element is compared against the fp32 matmul of
the same bf16-rounded inputs: |err| <= 2^-8 * |ref| + small absolute slack (one bf16 rounding of the result).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from sdxl_training_improvements_b200 import ops as o
    return o


def _close(got, ref, rtol=2 ** -7, atol=1e-2, what=""):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} bad, max err {float(err.max()):.4g}, " \
                          f"ref absmax {float(ref.abs().max()):.4g}"


def _rand(*shape, scale=1.0, seed=None):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed if seed is not None else (hash(shape) & 0xffff))
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(bf16)


GEMM_SHAPES = [
    # M, N, K
    (128, 128, 64),
    (128, 128, 256),
    (256, 384, 320),
    (4096, 1280, 1280),
    (300, 200, 136),      # ragged everything (K%64 != 0, M,N tails)
    (4, 1280, 320),       # time-embedding MLP: M = batch
    (308, 640, 2048),     # cross-attn K/V projection (B*77 rows)
    (1024, 64, 1024),     # N = 64 tile
    (16384, 8, 2880),     # conv_out (Cout padded to 8)
    (4000, 1280, 192),    # wide 256x320 single-round tiles, ragged M, short K (3 k-blocks < ring depth)
    (4096, 1280, 5120),   # wide tiles, long K (ff2 forward)
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_linear_fwd(ops, M, N, K):
    x = _rand(M, K, seed=1)
    W = _rand(N, K, scale=1 / math.sqrt(K), seed=2)
    b = _rand(N, seed=3)
    r = _rand(M, N, seed=4)
    ref = x.float() @ W.float().t()
    _close(ops.linear_fwd(x, W), ref, what="plain")
    _close(ops.linear_fwd(x, W, bias=b, residual=r), ref + b.float() + r.float(), what="bias+res")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES[:7])
def test_gemm_dgrad_wgrad(ops, M, N, K):
    x = _rand(M, K, seed=5)
    W = _rand(N, K, scale=1 / math.sqrt(K), seed=6)
    dy = _rand(M, N, seed=7)
    _close(ops.linear_dgrad(dy, W), dy.float() @ W.float(), atol=2e-2, what="dgrad (A K-major, B MN-major)")
    dW = torch.zeros(N, K, device="cuda", dtype=bf16)
    ops.linear_wgrad(dy, x, dW, accumulate=False)
    ref = dy.float().t() @ x.float()
    # split-K shapes combine bf16-rounded partial sums: the absolute error scales with the partials (~ max|ref|), not
    # with each element's own magnitude -> 2^-8 of the largest entry on top of the single-rounding tolerance
    atol = 2e-2 * math.sqrt(M / 128) + 2 ** -6 * float(ref.abs().max())
    _close(dW, ref, atol=atol, what="wgrad (both MN-major)")
    ops.linear_wgrad(dy, x, dW, accumulate=True)  # gradient accumulation: dW += ...
    _close(dW, 2 * ref, rtol=2 ** -6, atol=2 * atol, what="wgrad accumulate")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(4, 1280, 1280), (4, 320, 1280), (4, 1280, 2816), (1, 1280, 320), (8, 640, 1280), (3, 72, 64)])
def test_few_row_linears(ops, M, N, K, monkeypatch):
    """smallm.cu: the per-sample-row linears (time_embedding / add_embedding / time_emb_proj; diffusers
    ResnetBlock2D.forward `self.time_emb_proj(self.nonlinearity(temb))`) forward, dgrad, wgrad on CUDA cores, every epilogue
    option of b2_gemm, against fp32 torch and against the tensor-core GEMM path for the same call."""
    x = _rand(M, K, seed=11)
    W = _rand(N, K, scale=1 / math.sqrt(K), seed=12)
    b = _rand(N, seed=13)
    r = _rand(M, N, seed=14)
    row = _rand(N, seed=15)
    dy = _rand(M, N, seed=16)
    ref = x.float() @ W.float().t()
    _close(ops.linear_fwd(x, W), ref, what="fwd plain")
    _close(ops.linear_fwd(x, W, bias=b, residual=r), ref + b.float() + r.float(), what="fwd bias+res")
    y = torch.empty(M, N, device="cuda", dtype=bf16)
    ops.gemm_raw(x, W, y, M, N, K, lda=K, ldb=K, ldd=N, bias=b, residual=row, ldr=0)  # broadcast row via ldr = 0
    _close(y, ref + b.float() + row.float(), what="fwd broadcast residual row")
    yf = torch.full((M, N), 0.5, device="cuda", dtype=torch.float32)
    ops.gemm_raw(x, W, yf, M, N, K, lda=K, ldb=K, ldd=N, alpha=0.25, accumulate=True, out_fp32=True)
    assert float((yf - (0.25 * ref + 0.5)).abs().max()) < 1e-3, "fwd alpha / accumulate / fp32 out"
    # input gradient, plain and accumulated into an existing bf16 buffer
    dref = dy.float() @ W.float()
    _close(ops.linear_dgrad(dy, W), dref, atol=2e-2, what="dgrad")
    dx = _rand(M, K, seed=17)
    dx0 = dx.float().clone()
    ops.linear_dgrad(dy, W, dx, accumulate=True)
    _close(dx, dx0 + dref, atol=3e-2, what="dgrad accumulate")
    # weight gradient
    wref = dy.float().t() @ x.float()
    dW = torch.zeros(N, K, device="cuda", dtype=bf16)
    ops.linear_wgrad(dy, x, dW, accumulate=False)
    _close(dW, wref, atol=2e-2, what="wgrad")
    ops.linear_wgrad(dy, x, dW, accumulate=True)
    _close(dW, 2 * wref, rtol=2 ** -6, atol=4e-2, what="wgrad accumulate")
    torch.cuda.synchronize()


@pytest.mark.parametrize("tile_n", [64, 128, 256])
def test_gemm_tile_variants(ops, tile_n):
    M, N, K = 512, 512, 448
    x = _rand(M, K, seed=8)
    W = _rand(N, K, scale=1 / math.sqrt(K), seed=9)
    out = torch.empty(M, N, device="cuda", dtype=bf16)
    ops.gemm_raw(x, W, out, M, N, K, lda=K, ldb=K, ldd=N, tile_n=tile_n)
    _close(out, x.float() @ W.float().t(), what=f"tile_n={tile_n}")
    # MN-major B with the wide tile (4 x 64-column blocks)
    Wt = W.t().contiguous()  # [K, N]: n contiguous
    ops.gemm_raw(x, Wt, out, M, N, K, b_mn=True, lda=K, ldb=N, ldd=N, tile_n=tile_n)
    _close(out, x.float() @ W.float().t(), what=f"tile_n={tile_n} b_mn")


def test_gemm_batched_attention_layout(ops):
    """QK^T and PV exactly as the attention path issues them: batch = (sample, head) with independent strides."""
    B, n, h, d = 2, 256, 5, 64
    Cc = h * d
    qkv = _rand(B * n, 3 * Cc, seed=10)
    S = torch.empty(B, h, n, n, device="cuda", dtype=torch.float32)
    ops.gemm_raw(qkv, qkv[:, Cc:], S, n, n, d, lda=3 * Cc, ldb=3 * Cc, ldd=n, nb_lo=h, nb_hi=B,
                 a_bs=(d, n * 3 * Cc), b_bs=(d, n * 3 * Cc), d_bs=(n * n, h * n * n), alpha=0.125, out_fp32=True)
    q = qkv[:, :Cc].float().view(B, n, h, d).transpose(1, 2)
    k = qkv[:, Cc:2 * Cc].float().view(B, n, h, d).transpose(1, 2)
    v = qkv[:, 2 * Cc:].float().view(B, n, h, d).transpose(1, 2)
    _close(S, 0.125 * q @ k.transpose(-1, -2), rtol=1e-4, atol=1e-3, what="QK^T fp32 out")
    P = torch.softmax(S, -1).to(bf16)
    O = torch.empty(B * n, Cc, device="cuda", dtype=bf16)
    ops.gemm_raw(P, qkv[:, 2 * Cc:], O, n, d, n, b_mn=True, lda=n, ldb=3 * Cc, ldd=Cc, nb_lo=h, nb_hi=B,
                 a_bs=(n * n, h * n * n), b_bs=(d, n * 3 * Cc), d_bs=(d, n * Cc))
    ref = (P.float() @ v).transpose(1, 2).reshape(B * n, Cc)
    _close(O, ref, what="PV")


def test_gemm_row_group_bias(ops):
    """conv1 epilogue: per-sample time-embedding row added to every pixel of that sample."""
    B, HW, Cc, K = 3, 96, 128, 192
    x = _rand(B * HW, K, seed=11)
    W = _rand(Cc, K, scale=1 / math.sqrt(K), seed=12)
    t = _rand(B, Cc, seed=13)
    out = ops.linear_fwd(x, W, bias=t, bias_rows_per_group=HW, bias_group_stride=Cc)
    ref = x.float() @ W.float().t() + t.float().repeat_interleave(HW, 0)
    _close(out, ref, what="row-group bias")


@pytest.mark.parametrize("stride,up", [(1, False), (2, False), (1, True)])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 32, 64), (1, 12, 20, 8, 16)])
def test_conv3x3_via_im2col(ops, stride, up, B, H, W, Cin, Cout):
    x = _rand(B * H * W, Cin, seed=14)
    w = _rand(Cout, Cin, 3, 3, scale=1 / math.sqrt(9 * Cin), seed=15)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()  # (kh,kw,cin) K order
    Ho, Wo = ops.conv_out_hw(H, W, stride, up)
    col = ops.im2col3x3(x, B, H, W, Cin, stride, up)
    y = ops.linear_fwd(col, wk)
    xn = x.float().view(B, H, W, Cin).permute(0, 3, 1, 2).requires_grad_(True)
    xin = F.interpolate(xn, scale_factor=2.0, mode="nearest") if up else xn
    ref = F.conv2d(xin, w.float(), stride=stride, padding=1)
    _close(y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2), ref, what="conv fwd")
    dy = _rand(B * Ho * Wo, Cout, seed=16)
    ref.backward(dy.float().view(B, Ho, Wo, Cout).permute(0, 3, 1, 2))
    dcol = ops.linear_dgrad(dy, wk)
    dx = torch.empty_like(x)
    ops.col2im3x3(dcol, dx, B, H, W, Cin, stride, up)
    _close(dx.view(B, H, W, Cin).permute(0, 3, 1, 2), xn.grad, atol=3e-2, what="conv dgrad")


SPLITK_SHAPES = [
    # M (rows of x / dy), N (out features), K (in features): wgrad reduces over M, dgrad over N
    (16384, 640, 640),     # wgrad: 3x3 tiles, 256 k-blocks -> split 8
    (4096, 1280, 1280),    # wgrad: 25 tiles -> split 2
    (4096, 10240, 1280),   # dgrad: N_gemm = 1280 -> 80 tiles on 74 clusters, K_gemm = 10240 -> split
    (65536, 320, 320),     # conv-like wgrad: 2x2 tiles, 1024 k-blocks
]


@pytest.mark.parametrize("M,N,K", SPLITK_SHAPES)
def test_gemm_split_k_gradients(ops, M, N, K):
    """Gradient GEMMs whose tile grid cannot fill 74 CTA pairs split K and combine bf16 partials with TMA reduce-adds.
    Checked against fp32 matmul, with and without accumulation, and against the unsplit kernel (allow_split_k=False)."""
    x = _rand(M, K, seed=40)
    W = _rand(N, K, scale=1 / math.sqrt(K), seed=41)
    dy = _rand(M, N, scale=0.25, seed=42)
    ref_dx = dy.float() @ W.float()
    dx = torch.full((M, K), 7.0, device="cuda", dtype=bf16)     # must be overwritten (zero-filled by the call)
    ops.linear_dgrad(dy, W, dx, accumulate=False)
    sdx = float(ref_dx.abs().max())
    _close(dx, ref_dx, atol=3e-2 + 2 ** -6 * sdx, what="split-K dgrad")
    assert float((dx.float() - ref_dx).norm() / ref_dx.norm()) <= 8e-3
    ops.linear_dgrad(dy, W, dx, accumulate=True)
    _close(dx, 2 * ref_dx, rtol=2 ** -6, atol=6e-2 + 2 ** -5 * sdx, what="split-K dgrad accumulate")
    ref_dw = dy.float().t() @ x.float()
    dW = torch.full((N, K), -3.0, device="cuda", dtype=bf16)
    ops.linear_wgrad(dy, x, dW, accumulate=False)
    scale = float(ref_dw.abs().max())
    # elementwise bound: each of up to 32 bf16 partials carries 2^-9 of ITS magnitude; aggregate bound: rel-L2
    _close(dW, ref_dw, rtol=2 ** -6, atol=2 ** -6 * scale, what="split-K wgrad")
    # s reduce-adds round the running bf16 total s times: rel-L2 ~ sqrt(s) * 2^-9 / sqrt(3), s <= 16
    assert float((dW.float() - ref_dw).norm() / ref_dw.norm()) <= 8e-3
    dW1 = torch.zeros_like(dW)
    ops.gemm_raw(dy, x, dW1, N, K, M, a_mn=True, b_mn=True, lda=N, ldb=K, ldd=K, allow_split_k=False)
    _close(dW, dW1, rtol=2 ** -6, atol=2 ** -6 * scale, what="split vs unsplit")
    assert float((dW.float() - dW1.float()).norm() / dW1.float().norm()) <= 8e-3
    ops.linear_wgrad(dy, x, dW, accumulate=True)
    _close(dW, 2 * ref_dw, rtol=2 ** -5, atol=2 ** -5 * scale, what="split-K wgrad accumulate")
    torch.cuda.synchronize()


IMPLICIT_CONV_SHAPES = [
    # B, H, W, Cin, Cout
    (2, 16, 16, 128, 64),     # W < 128: box {64, 16, 8}; dgrad with a single 128-wide N tile
    (1, 32, 32, 192, 320),    # Cout not a multiple of 128 (ragged N tile in fwd, M tail in wgrad)
    (2, 8, 16, 256, 128),     # H != W, one 128-pixel slab per image
    (1, 128, 128, 128, 64),   # W == 128: one image row per slab, half a row per wgrad k-block
    (1, 2, 256, 128, 64),     # W > 128: two slabs per image row
    (4, 32, 32, 320, 320),    # SDXL-like (320 = 5 x 64 channel blocks per tap)
]


@pytest.mark.parametrize("B,H,W,Cin,Cout", IMPLICIT_CONV_SHAPES)
def test_conv3x3_implicit(ops, B, H, W, Cin, Cout):
    """b2_conv3x3 (TMA-gathered implicit GEMM) fwd / dgrad / wgrad vs F.conv2d autograd on the same bf16 inputs."""
    assert ops.conv3x3_implicit_ok(B, H, W, Cin, Cout)
    M = B * H * W
    x = _rand(M, Cin, seed=30)
    w = _rand(Cout, Cin, 3, 3, scale=1 / math.sqrt(9 * Cin), seed=31)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    bias = _rand(Cout, seed=32)
    rowb = _rand(B, Cout, seed=33)
    res = _rand(M, Cout, seed=34)

    def nchw(t, Cc):
        return t.float().view(B, H, W, Cc).permute(0, 3, 1, 2)

    xn = nchw(x, Cin).clone().requires_grad_(True)
    wn = w.float().clone().requires_grad_(True)
    ref = F.conv2d(xn, wn, padding=1)
    y = ops.conv3x3_fwd(x, wk, B, H, W, Cin, Cout, bias=bias)
    _close(nchw(y, Cout), ref + bias.float().view(1, -1, 1, 1), what="implicit fwd + bias")
    y = ops.conv3x3_fwd(x, wk, B, H, W, Cin, Cout, bias=rowb, residual=res, bias_per_sample=True)
    _close(nchw(y, Cout), ref + rowb.float().view(B, Cout, 1, 1) + nchw(res, Cout), what="implicit fwd + rowbias + res")
    dy = _rand(M, Cout, seed=35)
    ref.backward(nchw(dy, Cout))
    dx = torch.empty_like(x)
    ops.conv3x3_dgrad(dy, wk, dx, B, H, W, Cin, Cout)
    sdx = 2 ** -6 * float(xn.grad.abs().max())   # split-K partial-sum rounding (see test_gemm_dgrad_wgrad)
    _close(nchw(dx, Cin), xn.grad, atol=3e-2 + sdx, what="implicit dgrad")
    ops.conv3x3_dgrad(dy, wk, dx, B, H, W, Cin, Cout, accumulate=True)
    _close(nchw(dx, Cin), 2 * xn.grad, rtol=2 ** -6, atol=6e-2 + 2 * sdx, what="implicit dgrad accumulate")
    dwk = torch.zeros_like(wk)
    ops.conv3x3_wgrad(dy, x, dwk, B, H, W, Cin, Cout, accumulate=False)
    refw = wn.grad.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)
    sdw = 2 ** -6 * float(refw.abs().max())
    _close(dwk, refw, atol=2e-2 * math.sqrt(M / 128) + sdw, what="implicit wgrad")
    ops.conv3x3_wgrad(dy, x, dwk, B, H, W, Cin, Cout, accumulate=True)
    _close(dwk, 2 * refw, rtol=2 ** -6, atol=4e-2 * math.sqrt(M / 128) + 2 * sdw, what="implicit wgrad accumulate")
    torch.cuda.synchronize()


def test_conv3x3_implicit_matches_im2col_path(ops):
    """Both conv paths share the GEMM kernel; on identical inputs they must agree to the last bit of the fp32 sum order
    (same k order (kh, kw, cin), same tiles) — checked as exact equality of the bf16 outputs."""
    B, H, W, Cin, Cout = 2, 32, 32, 128, 128
    x = _rand(B * H * W, Cin, seed=36)
    wk = _rand(Cout, 9 * Cin, scale=1 / math.sqrt(9 * Cin), seed=37)
    y1 = ops.conv3x3_fwd(x, wk, B, H, W, Cin, Cout)
    y2 = ops.linear_fwd(ops.im2col3x3(x, B, H, W, Cin), wk)
    assert torch.equal(y1, y2)
    assert not ops.conv3x3_implicit_ok(1, 12, 20, 128, 128)   # ragged geometry -> im2col path
    assert not ops.conv3x3_implicit_ok(1, 16, 16, 8, 64)      # conv_in


@pytest.mark.parametrize("B,HW,Cc,silu", [(2, 256, 320, True), (2, 100, 64, False), (1, 64, 2560, True), (3, 77, 960, True)])
def test_groupnorm(ops, B, HW, Cc, silu):
    G = 32
    x = _rand(B * HW, Cc, seed=17) * 2 + 0.5
    gamma = (1 + 0.1 * torch.randn(Cc, device="cuda")).to(bf16)
    beta = (0.1 * torch.randn(Cc, device="cuda")).to(bf16)
    mean, rstd = ops.gn_stats(x, B, HW, Cc, G, 1e-5)
    y = ops.gn_apply(x, mean, rstd, gamma, beta, B, HW, Cc, G, silu)
    xr = x.float().view(B, HW, Cc).transpose(1, 2).requires_grad_(True)
    gr = gamma.float().requires_grad_(True)
    br = beta.float().requires_grad_(True)
    ref = F.group_norm(xr, G, gr, br, 1e-5)
    if silu:
        ref = F.silu(ref)
    _close(y.view(B, HW, Cc).transpose(1, 2), ref, what="gn fwd")
    dy = _rand(B * HW, Cc, seed=18)
    ref.backward(dy.float().view(B, HW, Cc).transpose(1, 2))
    dgb = torch.zeros(2 * Cc, device="cuda")
    dx = ops.gn_bwd(x, dy, mean, rstd, gamma, beta, B, HW, Cc, G, silu, dgb)
    _close(dx.view(B, HW, Cc).transpose(1, 2), xr.grad, atol=2e-2, what="gn dx")
    _close(dgb[:Cc], gr.grad, rtol=1e-2, atol=5e-2, what="gn dgamma")
    _close(dgb[Cc:], br.grad, rtol=1e-2, atol=5e-2, what="gn dbeta")


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("M,Cc", [(512, 640), (300, 1280), (64, 128), (4096, 1280), (16384, 640), (1003, 320)])
def test_layernorm(ops, M, Cc, fused, monkeypatch):
    # fused: the opt-in one-pass backward (dx + dgamma / dbeta), selected per call through the environment
    if fused:
        monkeypatch.setenv("B2_LN_BWD_FUSED", "1")
    else:
        monkeypatch.delenv("B2_LN_BWD_FUSED", raising=False)
    x = _rand(M, Cc, seed=19) * 1.5 + 0.2
    gamma = (1 + 0.1 * torch.randn(Cc, device="cuda")).to(bf16)
    beta = (0.1 * torch.randn(Cc, device="cuda")).to(bf16)
    y, mean, rstd = ops.ln_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    gr = gamma.float().requires_grad_(True)
    br = beta.float().requires_grad_(True)
    ref = F.layer_norm(xr, (Cc,), gr, br, 1e-5)
    _close(y, ref, what="ln fwd")
    dy = _rand(M, Cc, seed=20)
    ref.backward(dy.float())
    dgb = torch.zeros(2 * Cc, device="cuda")
    dx = ops.ln_bwd(x, dy, gamma, mean, rstd, dgb)
    _close(dx, xr.grad, atol=2e-2, what="ln dx")
    _close(dgb[:Cc], gr.grad, rtol=1e-2, atol=5e-2, what="ln dgamma")
    _close(dgb[Cc:], br.grad, rtol=1e-2, atol=5e-2, what="ln dbeta")
    # accumulate: dx += and the fp32 staging of dgamma / dbeta += (what gradient accumulation and fan-out rely on)
    dx0 = _rand(M, Cc, seed=23)
    dx2 = dx0.clone()
    dgb2 = dgb.clone()
    ops.ln_bwd(x, dy, gamma, mean, rstd, dgb2, dx=dx2, accumulate=True)
    _close(dx2, dx0.float() + xr.grad, atol=3e-2, what="ln dx accumulate")
    _close(dgb2, 2 * dgb, rtol=1e-3, atol=1e-2 * max(1.0, float(dgb.abs().max())) * 1e-1, what="ln dgb accumulate")


@pytest.mark.parametrize("rows,n", [(64, 77), (40, 1024), (8, 4096)])
def test_softmax(ops, rows, n):
    ld = (n + 7) // 8 * 8
    S = torch.randn(rows, ld, device="cuda") * 3
    P = torch.empty(rows, ld, device="cuda", dtype=bf16)
    ops.softmax_fwd(S, P, rows, n)
    ref = torch.softmax(S[:, :n], -1)
    _close(P[:, :n], ref, atol=1e-3, what="softmax")
    assert float(P[:, n:].float().abs().sum()) == 0.0
    dP = torch.randn(rows, ld, device="cuda")
    dS = torch.empty_like(P)
    ops.softmax_bwd(P, dP, dS, rows, n, 0.125)
    Pf = P[:, :n].float()
    refd = Pf * (dP[:, :n] - (dP[:, :n] * Pf).sum(-1, keepdim=True)) * 0.125
    _close(dS[:, :n], refd, atol=1e-3, what="softmax bwd")


@pytest.mark.parametrize("M,F,Cc", [(4096, 5120, 1280), (16384, 2560, 640), (1000, 384, 128), (256, 256, 64)])
def test_dgrad_with_geglu_backward_epilogue_is_bit_identical(ops, M, F, Cc):
    """b2_linear_dgrad_geglu == b2_gemm(dgrad of the down-projection) + b2_geglu_bwd, bit for bit (dz is rounded to bf16 in the
    epilogue exactly where the un-fused path stores it); and against fp32 torch for the whole expression."""
    assert ops.linear_dgrad_geglu_ok(M, F, Cc)
    dy = _rand(M, Cc, seed=31)
    W2 = _rand(Cc, F, scale=1 / math.sqrt(F), seed=32)
    u = _rand(M, 2 * F, seed=33)
    dz = ops.linear_dgrad(dy, W2)                 # [M, F]
    ref = ops.geglu_bwd(u, dz, F)
    got = ops.linear_dgrad_geglu(dy, W2, u, F)
    torch.cuda.synchronize()
    assert torch.equal(got, ref), f"{int((got != ref).sum())} of {ref.numel()} elements differ"
    h, g = u[:, :F].float(), u[:, F:].float()
    dzf = (dy.float() @ W2.float())
    cdf = 0.5 * (1 + torch.erf(g / math.sqrt(2)))
    pdf = torch.exp(-0.5 * g * g) / math.sqrt(2 * math.pi)
    _close(got[:, :F], dzf * g * cdf, atol=3e-2, what="dh")
    _close(got[:, F:], dzf * h * (cdf + g * pdf), atol=3e-2, what="dg")


@pytest.mark.parametrize("M,F", [(4096, 5120), (300, 128), (1003, 2560), (64, 40)])
def test_geglu_bwd_with_bias_column_sums(ops, M, F):
    """b2_geglu_bwd_bias: du identical to b2_geglu_bwd, and db32 += column sums of the bf16 du (what b2_colsum_f32 over du gives)."""
    u = _rand(M, 2 * F, seed=21)
    dz = _rand(M, F, seed=22)
    du0 = ops.geglu_bwd(u, dz, F)
    db = torch.full((2 * F,), 0.25, device="cuda", dtype=torch.float32)
    du1 = ops.geglu_bwd(u, dz, F, dbias32=db)
    assert torch.equal(du0, du1)
    ref = 0.25 + du0.float().sum(0)
    assert float((db - ref).abs().max()) <= 1e-3 * max(1.0, float(ref.abs().max())), float((db - ref).abs().max())
    torch.cuda.synchronize()


def test_geglu_silu_add_colsum(ops):
    M, Fd = 96, 640
    u = _rand(M, 2 * Fd, seed=21)
    ur = u.float().requires_grad_(True)
    h, g = ur.chunk(2, -1)
    ref = h * F.gelu(g)
    _close(ops.geglu_fwd(u, Fd), ref, what="geglu")
    dz = _rand(M, Fd, seed=22)
    ref.backward(dz.float())
    _close(ops.geglu_bwd(u, dz, Fd), ur.grad, atol=2e-2, what="geglu bwd")
    x = _rand(1000, seed=23)
    _close(ops.silu_fwd(x), F.silu(x.float()), what="silu")
    xr = x.float().requires_grad_(True)
    dy = _rand(1000, seed=24)
    F.silu(xr).backward(dy.float())
    _close(ops.silu_bwd(x, dy), xr.grad, what="silu bwd")
    a, b = _rand(1003, seed=25), _rand(1003, seed=26)
    _close(ops.add(a, b), a.float() + b.float(), what="add")
    dyy = _rand(777, 320, seed=27)
    db = torch.zeros(320, device="cuda", dtype=bf16)
    ops.colsum(dyy, db, accumulate=False)
    _close(db, dyy.float().sum(0), rtol=1e-2, atol=5e-2, what="colsum")


def test_layout_timestep_noise_loss(ops):
    B, Cc, H, W = 2, 4, 8, 12
    x = torch.randn(B, Cc, H, W, device="cuda")
    y = ops.nchw_to_nhwc(x, 8)
    assert torch.equal(y.view(B, H * W, 8)[..., :4].transpose(1, 2).reshape(B, Cc, H, W), x.to(bf16))
    assert float(y.view(B, H * W, 8)[..., 4:].float().abs().sum()) == 0
    back = ops.nhwc_to_nchw(y, B, Cc, H, W, 8, dtype=torch.float32)
    assert torch.equal(back, x.to(bf16).float())
    t = torch.tensor([0.0, 17.0, 999.0, 0.4375], device="cuda")
    e = ops.timestep_embedding(t, 320)
    half = 160
    f = torch.exp(-math.log(10000.0) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    a = t[:, None] * f[None]
    _close(e, torch.cat([a.cos(), a.sin()], -1), rtol=0, atol=2 ** -8 + 2e-3, what="sinusoid")
    so = torch.tensor([1234, 0], device="cuda", dtype=torch.int64)
    z = ops.randn(1 << 20, so, 0, round_bf16=False)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1) < 5e-3
    assert abs(float((z ** 4).mean()) - 3) < 0.05
    z2 = ops.randn(1 << 20, so, 0, round_bf16=False)
    assert torch.equal(z, z2)
    ops.philox_advance(so, 1)
    z3 = ops.randn(1 << 20, so, 0, round_bf16=False)
    assert not torch.equal(z, z3) and abs(float((z * z3).mean())) < 5e-3
    # noising + loss
    HW = H * W
    xl = x.to(bf16).float()
    eps = ops.randn(B * Cc * HW, so, 1).view(B, Cc, HW)
    sig = torch.tensor([3.5, 20000.0], device="cuda")
    noisy, target = ops.make_noisy(xl, eps, sig, 0, True, True, B, Cc, HW, 8)
    xr = xl.view(B, Cc, HW)
    refn = torch.clamp(xr + sig.view(-1, 1, 1) * eps, -20000, 20000)
    assert torch.equal(noisy.view(B, HW, 8)[..., :4].transpose(1, 2), refn.to(bf16))
    _close(target, (eps - xr) / sig.view(-1, 1, 1), rtol=1e-6, atol=1e-9, what="velocity")
    tt = torch.tensor([0.25, 0.8125], device="cuda")
    noisy, target = ops.make_noisy(xl, eps, tt, 1, False, False, B, Cc, HW, 8)
    tb = tt.to(bf16).view(-1, 1, 1)
    ref_xt = (1 - tb) * eps.to(bf16) + tb * xr.to(bf16)
    assert torch.equal(noisy.view(B, HW, 8)[..., :4].transpose(1, 2), ref_xt)
    assert torch.equal(target, (xr.to(bf16) - eps.to(bf16)).float())
    pred = _rand(B * HW, 8, seed=28)
    ls = torch.zeros(1, device="cuda", dtype=torch.float64)
    dpred = torch.empty_like(pred)
    cnt = B * Cc * HW
    ops.mse_loss(pred, target, None, ls, dpred, 1.0 / cnt, B, Cc, HW, 8)
    loss = torch.empty(1, device="cuda")
    ok = torch.empty(1, device="cuda", dtype=torch.int32)
    ops.finalize_loss(ls, cnt, 1.0, loss, ok, dpred)
    pr = pred.float().view(B, HW, 8)[..., :4].transpose(1, 2).requires_grad_(True)
    refl = F.mse_loss(pr, target)
    refl.backward()
    assert abs(float(loss) - float(refl)) < 1e-5 * max(1, float(refl)) and int(ok) == 1
    _close(dpred.view(B, HW, 8)[..., :4].transpose(1, 2), pr.grad, atol=1e-4, what="dpred")


def test_adamw_and_sumsq(ops):
    n = 100003
    p32 = torch.randn(n, device="cuda")
    g = _rand(n, seed=29)
    p = p32.to(bf16)
    master = p.float()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    ref_p = master.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    gs = torch.zeros(1, device="cuda", dtype=torch.float64)
    for step in range(1, 4):
        gs.zero_()
        ops.sumsq(g, gs)
        assert abs(float(gs) - float((g.float() ** 2).sum())) < 1e-3 * float(gs)
        ops.adamw(p, master, g, m, v, lr=1e-3, step=step, gnorm_sq=gs, max_norm=1.0)
        ref_p.grad = g.float().clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
    _close(master, ref_p.detach(), rtol=1e-5, atol=1e-6, what="adamw master")
    assert torch.equal(p, master.to(bf16))


def test_small_grad_staging_and_group_colsum(ops):
    """fp32 staging path of the bias / norm-affine gradients and the per-sample (grouped) column sum."""
    dy = _rand(3 * 200, 320, seed=50)
    out = torch.zeros(3, 320, device="cuda", dtype=bf16)
    ops.colsum_groups(dy, out, 3, 200, accumulate=False)
    ref = dy.float().view(3, 200, 320).sum(1)
    _close(out, ref, rtol=1e-2, atol=5e-2, what="colsum_groups")
    ops.colsum_groups(dy, out, 3, 200, accumulate=True)
    _close(out, 2 * ref, rtol=2e-2, atol=1e-1, what="colsum_groups accumulate")
    st = torch.zeros(1000, device="cuda", dtype=torch.float32)
    ops.colsum_f32(dy, st[40:360])
    ops.colsum_f32(dy, st[40:360])
    _close(st[40:360], 2 * dy.float().sum(0), rtol=1e-5, atol=1e-3, what="colsum_f32 accumulates")
    assert float(st[:40].abs().sum()) == 0 and float(st[360:].abs().sum()) == 0
    grad = _rand(5000, seed=51)
    g0 = grad.clone()
    segs = torch.tensor([40, 1000, 320, 600, 16, 8], device="cuda", dtype=torch.int64)  # (src, dst, n) triples
    st[600:608] = 3.0
    ops.flush_small_grads(st, grad, segs, 2)
    exp = g0.float().clone()
    exp[1000:1320] += 2 * dy.float().sum(0)
    exp[16:24] += 3.0
    _close(grad, exp, rtol=2 ** -7, atol=2e-2, what="flush_small_grads")
    assert float(st.abs().sum()) == 0.0, "staging must be cleared by the flush"


@pytest.mark.parametrize("M,Fd,K", [(4096, 5120, 1280), (1024, 2560, 640), (300, 512, 128), (128, 128, 64)])
def test_linear_geglu_fused_epilogue_is_bit_identical_to_gemm_plus_geglu(ops, M, Fd, K):
    """b2_linear_geglu (gate in the GEMM epilogue, GEGLU mode of gemm2_kernel) vs the two-kernel path b2_gemm + b2_geglu_fwd:
    u and z BIT-identical (same accumulation order per output element, same bf16 rounding points), and vs torch:
    diffusers GEGLU.forward = proj -> chunk(2) -> hidden * F.gelu(gate) (exact erf)."""
    assert ops.linear_geglu_ok(M, Fd, K)
    x = _rand(M, K, seed=31)
    W1 = _rand(2 * Fd, K, scale=K ** -0.5, seed=32)
    b1 = _rand(2 * Fd, scale=0.1, seed=33)
    u, z = ops.linear_geglu_fwd(x, W1, b1, Fd)
    u2 = ops.linear_fwd(x, W1, bias=b1)
    z2 = ops.geglu_fwd(u2, Fd)
    if M >= 256:  # below that b2_gemm takes its generic (different accumulation order) kernel: compare with tolerance only
        assert torch.equal(u, u2), f"u differs in {int((u != u2).sum())} elements"
        assert torch.equal(z, z2), f"z differs in {int((z != z2).sum())} elements"
    ur = x.float() @ W1.float().t() + b1.float()
    _close(u, ur, what="geglu-fused u")
    h, g = u.float().chunk(2, -1)  # the gate is formed from the bf16 values u holds (one bf16 rounding of the product)
    _close(z, h * F.gelu(g), what="geglu-fused z")
