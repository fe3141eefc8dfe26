"""Thin Python wrappers: torch tensors (device memory + stream only) -> C-ABI calls of libsdxl_b200.so.

Every function enqueues hand-written sm_100a kernels on torch's current CUDA stream; nothing here computes
with torch ops.  Shapes follow the token-major convention [B*H*W, C] bf16.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

bf16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sdxl_b200 ops need CUDA tensors (no CPU fallback)")


# --------------------------------------------------------------------------------------- GEMM
def gemm_raw(A, B, D, M, N, K, *, a_mn=False, b_mn=False, lda, ldb, ldd, bias=None, residual=None, ldr=0,
             nb_lo=1, nb_hi=1, a_bs=(0, 0), b_bs=(0, 0), d_bs=(0, 0), r_bs=(0, 0), alpha=1.0, accumulate=False,
             out_fp32=False, bias_rows_per_group=0, bias_group_stride=0, tile_n=0, allow_split_k=False):
    _need_cuda(A, B, D, bias, residual)
    lib = _lib.load()
    g = _lib.GemmArgs()
    g.A, g.B, g.D, g.bias, g.residual = _p(A), _p(B), _p(D), _p(bias), _p(residual)
    g.M, g.N, g.K = int(M), int(N), int(K)
    g.nb_lo, g.nb_hi = int(nb_lo), int(nb_hi)
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    g.lda, g.ldb, g.ldd, g.ldr = int(lda), int(ldb), int(ldd), int(ldr)
    g.a_bs_lo, g.a_bs_hi = int(a_bs[0]), int(a_bs[1])
    g.b_bs_lo, g.b_bs_hi = int(b_bs[0]), int(b_bs[1])
    g.d_bs_lo, g.d_bs_hi = int(d_bs[0]), int(d_bs[1])
    g.r_bs_lo, g.r_bs_hi = int(r_bs[0]), int(r_bs[1])
    g.alpha = float(alpha)
    g.accumulate, g.out_fp32 = int(accumulate), int(out_fp32)
    g.bias_rows_per_group, g.bias_group_stride = int(bias_rows_per_group), int(bias_group_stride)
    g.tile_n = int(tile_n)
    g.allow_split_k = int(allow_split_k)
    _lib.check(lib.b2_gemm(C.byref(g), _stream()), "b2_gemm")
    return D


def linear_fwd(x, W, bias=None, residual=None, out=None, *, bias_rows_per_group=0, bias_group_stride=0, ldw=None):
    """out[M,N] = x[M,K] @ W[N,K]^T (+bias) (+residual).  x, W row-major (W may have row stride ldw)."""
    M, K = x.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=bf16)
    return gemm_raw(x, W, out, M, N, K, lda=x.stride(0), ldb=(ldw or W.stride(0)), ldd=out.stride(0), bias=bias,
                    residual=residual, ldr=(residual.stride(0) if residual is not None else 0),
                    bias_rows_per_group=bias_rows_per_group, bias_group_stride=bias_group_stride)


def linear_dgrad(dy, W, dx=None, accumulate=False, *, K=None, ldw=None):
    """dx[M,K] (+)= dy[M,N] @ W[N,K]."""
    M, N = dy.shape
    K = K or W.shape[1]
    if dx is None:
        dx = torch.empty((M, K), device=dy.device, dtype=bf16)
        accumulate = False
    return gemm_raw(dy, W, dx, M, K, N, a_mn=False, b_mn=True, lda=dy.stride(0), ldb=(ldw or W.stride(0)),
                    ldd=dx.stride(0), accumulate=accumulate, allow_split_k=True)


def linear_wgrad(dy, x, dW, accumulate=True, *, ldw=None, K=None):
    """dW[N,K] (+)= dy[M,N]^T @ x[M,K]."""
    M, N = dy.shape
    K = K or x.shape[1]
    return gemm_raw(dy, x, dW, N, K, M, a_mn=True, b_mn=True, lda=dy.stride(0), ldb=x.stride(0),
                    ldd=(ldw or dW.stride(0)), accumulate=accumulate, allow_split_k=True)


# --------------------------------------------------------------------------------------- conv staging
def conv_out_hw(H, W, stride=1, upsample=False):
    Hin, Win = (2 * H, 2 * W) if upsample else (H, W)
    return (Hin + 2 - 3) // stride + 1, (Win + 2 - 3) // stride + 1


def im2col3x3(x, B, H, W, Cc, stride=1, upsample=False, out=None):
    Ho, Wo = conv_out_hw(H, W, stride, upsample)
    ldc = 9 * Cc
    if out is None:
        out = torch.empty((B * Ho * Wo, ldc), device=x.device, dtype=bf16)
    _lib.check(_lib.load().b2_im2col3x3(_p(x), _p(out), B, H, W, Cc, stride, int(upsample), ldc, _stream()), "im2col")
    return out


def col2im3x3(dcol, dx, B, H, W, Cc, stride=1, upsample=False, accumulate=False):
    _lib.check(_lib.load().b2_col2im3x3(_p(dcol), _p(dx), B, H, W, Cc, stride, int(upsample), dcol.stride(0),
                                       int(accumulate), _stream()), "col2im")
    return dx


def conv3x3_implicit_ok(B, H, W, Cin, Cout) -> bool:
    return bool(_lib.load().b2_conv3x3_implicit_ok(int(B), int(H), int(W), int(Cin), int(Cout)))


def _conv_args(mode, x, w, y, B, H, W, Cin, Cout, bias=None, residual=None, bias_per_sample=False, accumulate=False):
    _need_cuda(x, w, y, bias, residual)
    a = _lib.ConvArgs()
    a.x, a.w, a.y, a.bias, a.residual = _p(x), _p(w), _p(y), _p(bias), _p(residual)
    a.mode, a.B, a.H, a.W, a.Cin, a.Cout = mode, int(B), int(H), int(W), int(Cin), int(Cout)
    a.ldx, a.ldy = x.stride(0), y.stride(0)
    a.ldr = residual.stride(0) if residual is not None else 0
    a.bias_per_sample, a.accumulate = int(bias_per_sample), int(accumulate)
    return a


def conv3x3_fwd(x, Wk, B, H, W, Cin, Cout, bias=None, residual=None, bias_per_sample=False, out=None):
    """y[B*H*W, Cout] = conv3x3(x[B*H*W, Cin] NHWC, Wk[Cout, 9*Cin]) + bias (+residual); implicit GEMM, no im2col."""
    if out is None:
        out = torch.empty((B * H * W, Cout), device=x.device, dtype=bf16)
    a = _conv_args(0, x, Wk, out, B, H, W, Cin, Cout, bias, residual, bias_per_sample)
    _lib.check(_lib.load().b2_conv3x3(C.byref(a), _stream()), "b2_conv3x3 fwd")
    return out


def conv3x3_dgrad(dy, Wk, dx, B, H, W, Cin, Cout, accumulate=False):
    a = _conv_args(1, dx, Wk, dy, B, H, W, Cin, Cout, accumulate=accumulate)
    _lib.check(_lib.load().b2_conv3x3(C.byref(a), _stream()), "b2_conv3x3 dgrad")
    return dx


def conv3x3_wgrad(dy, x, dWk, B, H, W, Cin, Cout, accumulate=True):
    a = _conv_args(2, x, dWk, dy, B, H, W, Cin, Cout, accumulate=accumulate)
    _lib.check(_lib.load().b2_conv3x3(C.byref(a), _stream()), "b2_conv3x3 wgrad")
    return dWk


# --------------------------------------------------------------------------------------- norms
def gn_stats(x, B, HW, Cc, G, eps):
    ws = torch.empty(B * G * 2, device=x.device, dtype=torch.float64)
    mean = torch.empty(B * G, device=x.device, dtype=torch.float32)
    rstd = torch.empty(B * G, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_gn_stats(_p(x), B, HW, Cc, G, eps, _p(ws), _p(mean), _p(rstd), _stream()), "gn_stats")
    return mean, rstd


def gn_apply(x, mean, rstd, gamma, beta, B, HW, Cc, G, silu, out=None):
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().b2_gn_apply(_p(x), _p(out), _p(mean), _p(rstd), _p(gamma), _p(beta), B, HW, Cc, G,
                                      int(silu), _stream()), "gn_apply")
    return out


def gn_bwd(x, dy, mean, rstd, gamma, beta, B, HW, Cc, G, silu, dgb, dx=None, accumulate=False):
    if dx is None:
        dx = torch.empty_like(x)
        accumulate = False
    ws = torch.empty(B * G * 2, device=x.device, dtype=torch.float64)
    _lib.check(_lib.load().b2_gn_bwd(_p(x), _p(dy), _p(dx), _p(mean), _p(rstd), _p(gamma), _p(beta), B, HW, Cc, G,
                                    int(silu), _p(ws), _p(dgb), int(accumulate), _stream()), "gn_bwd")
    return dx


def ln_fwd(x, gamma, beta, eps=1e-5):
    M, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_ln_fwd(_p(x), _p(y), _p(gamma), _p(beta), _p(mean), _p(rstd), M, Cc, eps, _stream()),
               "ln_fwd")
    return y, mean, rstd


def ln_bwd(x, dy, gamma, mean, rstd, dgb, dx=None, accumulate=False, parts=3):
    """parts: 1 = dx only, 2 = dgamma / dbeta accumulation into `dgb` only, 3 = both (two kernels either way)."""
    M, Cc = x.shape
    if dx is None and parts & 1:
        dx = torch.empty_like(x)
        accumulate = False
    _lib.check(_lib.load().b2_ln_bwd_parts(_p(x), _p(dy), _p(dx), _p(gamma), _p(mean), _p(rstd), _p(dgb), M, Cc,
                                          int(accumulate), int(parts), _stream()), "ln_bwd")
    return dx


# --------------------------------------------------------------------------------------- attention pieces
def softmax_fwd(S, P, rows, n_valid):
    _lib.check(_lib.load().b2_softmax_fwd(_p(S), _p(P), rows, n_valid, S.stride(-2), P.stride(-2), _stream()), "softmax")
    return P


def softmax_bwd(P, dP, dS, rows, n_valid, scale):
    _lib.check(_lib.load().b2_softmax_bwd(_p(P), _p(dP), _p(dS), rows, n_valid, dP.stride(-2), P.stride(-2),
                                         float(scale), _stream()), "softmax_bwd")
    return dS


def _attn_args(q, k, v, o, lse, B, H, n_q, n_k, scale, flags=0):
    """q/k/v/o: 2-D bf16 views [B*n, ld] whose first H*64 columns are this tensor's heads (row stride = .stride(0))."""
    a = _lib.AttnArgs()
    a.Q, a.K, a.V, a.O, a.LSE = _p(q), _p(k), _p(v), _p(o), _p(lse)
    a.B, a.H, a.n_q, a.n_k = int(B), int(H), int(n_q), int(n_k)
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    a.q_bs, a.k_bs, a.v_bs, a.o_bs = n_q * q.stride(0), n_k * k.stride(0), n_k * v.stride(0), n_q * o.stride(0)
    a.scale, a.flags = float(scale), int(flags)
    return a


def attn_lse_rows(n_q: int) -> int:
    return (n_q + 127) // 128 * 128


def attn_fwd(q, k, v, B, H, n_q, n_k, scale, out=None):
    """O = softmax(Q K^T * scale) V per (sample, head); head dim 64.  Returns (O [B*n_q, H*64] bf16, LSE fp32)."""
    _need_cuda(q, k, v)
    if out is None:
        out = torch.empty((B * n_q, H * 64), device=q.device, dtype=bf16)
    lse = torch.empty((B, H, attn_lse_rows(n_q)), device=q.device, dtype=torch.float32)
    a = _attn_args(q, k, v, out, lse, B, H, n_q, n_k, scale)
    _lib.check(_lib.load().b2_attn_fwd(C.byref(a), _stream()), "b2_attn_fwd")
    return out, lse


def attn_bwd(q, k, v, o, lse, do, dq, dk, dv, B, H, n_q, n_k, scale, flags=0):
    """dQ, dK, dV (written, not accumulated) of attn_fwd; dq/dk/dv are 2-D bf16 views like q/k/v."""
    _need_cuda(q, k, v, o, do, dq, dk, dv)
    a = _attn_args(q, k, v, o, lse, B, H, n_q, n_k, scale, flags)
    d = torch.empty_like(lse)
    a.dO, a.D, a.dQ, a.dK, a.dV = _p(do), _p(d), _p(dq), _p(dk), _p(dv)
    a.lddo, a.lddq, a.lddk, a.lddv = do.stride(0), dq.stride(0), dk.stride(0), dv.stride(0)
    a.do_bs, a.dq_bs, a.dk_bs, a.dv_bs = n_q * do.stride(0), n_q * dq.stride(0), n_k * dk.stride(0), n_k * dv.stride(0)
    _lib.check(_lib.load().b2_attn_bwd(C.byref(a), _stream()), "b2_attn_bwd")
    return dq, dk, dv


# --------------------------------------------------------------------------------------- elementwise
def geglu_fwd(u, F):
    M = u.shape[0]
    z = torch.empty((M, F), device=u.device, dtype=bf16)
    _lib.check(_lib.load().b2_geglu_fwd(_p(u), _p(z), M, F, _stream()), "geglu_fwd")
    return z


def xattn_q_core_ok(B, n_q, n_k, C) -> bool:
    return bool(_lib.load().b2_xattn_q_core_ok(int(B), int(n_q), int(n_k), int(C)))


def xattn_q_core(xn, Wq, k, v, B, n_q, n_k, scale):
    """(Q, O, LSE) of a cross-attention: Q = xn @ Wq^T and O = softmax(Q K^T * scale) V per 64-channel head in ONE launch.
    k, v: [B * n_k, C] column views of the grouped K / V projection (row stride = k.stride(0))."""
    _need_cuda(xn, Wq, k, v)
    M, Cc = xn.shape
    H = Cc // 64
    q = torch.empty((M, Cc), device=xn.device, dtype=bf16)
    o = torch.empty((M, Cc), device=xn.device, dtype=bf16)
    n_pad = int(_lib.load().b2_attn_lse_rows(int(n_q)))
    lse = torch.empty((B, H, n_pad), device=xn.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_xattn_q_core(_p(xn), _p(Wq), _p(k), _p(v), _p(q), _p(o), _p(lse), int(B), int(n_q), int(n_k),
                                           int(Cc), int(xn.stride(0)), int(Wq.stride(0)), Cc, Cc, int(k.stride(0)),
                                           int(v.stride(0)), int(n_k * k.stride(0)), int(n_k * v.stride(0)), float(scale),
                                           _stream()), "xattn_q_core")
    return q, o, lse


def linear_dgrad_geglu_ok(M, F, C) -> bool:
    return bool(_lib.load().b2_linear_dgrad_geglu_ok(int(M), int(F), int(C)))


def linear_dgrad_geglu(dy, W2, u, F):
    """du[M, 2F]: the down-projection's input gradient dz = dy @ W2 pushed through the GEGLU backward in the GEMM epilogue
    (dz itself is never stored).  dy [M, C], W2 [C, F] (Linear weight, row-major), u [M, 2F] = [h | g]."""
    _need_cuda(dy, W2, u)
    M, Cc = dy.shape
    du = torch.empty_like(u)
    _lib.check(_lib.load().b2_linear_dgrad_geglu(_p(dy), _p(W2), _p(u), _p(du), int(M), int(F), int(Cc), int(dy.stride(0)),
                                                 int(W2.stride(0)), int(u.stride(0)), int(du.stride(0)), _stream()),
               "linear_dgrad_geglu")
    return du


def linear_geglu_ok(M, F, K) -> bool:
    return bool(_lib.load().b2_linear_geglu_ok(int(M), int(F), int(K)))


def linear_geglu_fwd(x, W1, b1, F):
    """(u, z): u[M, 2F] = x @ W1^T + b1, z[M, F] = u[:, :F] * gelu(u[:, F:]) in ONE GEMM launch (gate in the epilogue)."""
    _need_cuda(x, W1, b1)
    M, K = x.shape
    u = torch.empty((M, 2 * F), device=x.device, dtype=bf16)
    z = torch.empty((M, F), device=x.device, dtype=bf16)
    _lib.check(_lib.load().b2_linear_geglu(_p(x), _p(W1), _p(b1), _p(u), _p(z), int(M), int(F), int(K), int(x.stride(0)),
                                           int(W1.stride(0)), 2 * F, F, _stream()), "linear_geglu")
    return u, z


def geglu_bwd(u, dz, F, dbias32=None):
    """du[M, 2F] from u = [h | g] and dz; with `dbias32` (fp32 [2F] staging slice) the column sums of du — the bias gradient
    of the up-projection — are accumulated into it in the same pass."""
    M = u.shape[0]
    du = torch.empty_like(u)
    if dbias32 is not None:
        _lib.check(_lib.load().b2_geglu_bwd_bias(_p(u), _p(dz), _p(du), _p(dbias32), M, F, _stream()), "geglu_bwd_bias")
    else:
        _lib.check(_lib.load().b2_geglu_bwd(_p(u), _p(dz), _p(du), M, F, _stream()), "geglu_bwd")
    return du


def silu_fwd(x):
    y = torch.empty_like(x)
    _lib.check(_lib.load().b2_silu_fwd(_p(x), _p(y), x.numel(), _stream()), "silu_fwd")
    return y


def silu_bwd(x, dy, dx=None, accumulate=False):
    if dx is None:
        dx = torch.empty_like(x)
        accumulate = False
    _lib.check(_lib.load().b2_silu_bwd(_p(x), _p(dy), _p(dx), x.numel(), int(accumulate), _stream()), "silu_bwd")
    return dx


def add(a, b, out=None):
    if out is None:
        out = torch.empty_like(a)
    _lib.check(_lib.load().b2_add(_p(a), _p(b), _p(out), a.numel(), _stream()), "add")
    return out


def copy2d(src, dst, rows, cols, lds, ldd, accumulate=False):
    _lib.check(_lib.load().b2_copy2d(_p(src), _p(dst), rows, cols, lds, ldd, int(accumulate), _stream()), "copy2d")
    return dst


def copy2d_any(src, dst, rows, cols, lds, ldd, accumulate=False):
    _lib.check(_lib.load().b2_copy2d_any(_p(src), _p(dst), rows, cols, lds, ldd, int(accumulate), _stream()), "copy2d_any")
    return dst


def colsum(dy, db, accumulate=True):
    M, N = dy.shape
    ws = torch.empty(N, device=dy.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_colsum(_p(dy), _p(db), M, N, dy.stride(0), int(accumulate), _p(ws), _stream()), "colsum")
    return db


def colsum_groups(dy, out, groups, rows_per_group, accumulate=False):
    """out[g, :] (+)= column sums of rows [g*rows_per_group, (g+1)*rows_per_group) of dy; out bf16 [groups, N]."""
    N = dy.shape[1]
    ws = torch.empty(groups * N, device=dy.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_colsum_groups(_p(dy), _p(out), int(groups), int(rows_per_group), N, dy.stride(0),
                                           int(accumulate), _p(ws), _stream()), "colsum_groups")
    return out


def colsum_f32(dy, out32):
    """out32[n] += sum_m dy[m, n]  (fp32 staging slice; see ParamStore.small32)."""
    M, N = dy.shape
    _lib.check(_lib.load().b2_colsum_f32(_p(dy), _p(out32), M, N, dy.stride(0), _stream()), "colsum_f32")


def flush_small_grads(staging, grad, segments, nseg):
    _lib.check(_lib.load().b2_flush_small_grads(_p(staging), _p(grad), _p(segments), int(nseg), _stream()), "flush_small_grads")


def accum_f32_to_bf16(src, dst, accumulate=True):
    _lib.check(_lib.load().b2_accum_f32_to_bf16(_p(src), _p(dst), src.numel(), int(accumulate), _stream()), "accum")
    return dst


def nchw_to_nhwc(x, Cpad):
    B, Cc, H, W = x.shape
    x = x.contiguous()
    if x.dtype not in (torch.float32, bf16):
        x = x.float()
    y = torch.empty((B * H * W, Cpad), device=x.device, dtype=bf16)
    _lib.check(_lib.load().b2_nchw_to_nhwc(_p(x), int(x.dtype == torch.float32), _p(y), B, Cc, H * W, Cpad, _stream()),
               "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x, B, Cc, H, W, Cpad, dtype=bf16):
    y = torch.empty((B, Cc, H, W), device=x.device, dtype=dtype)
    _lib.check(_lib.load().b2_nhwc_to_nchw(_p(x), _p(y), int(dtype == torch.float32), B, Cc, H * W, Cpad, _stream()),
               "nhwc_to_nchw")
    return y


def timestep_embedding(t_f32, dim, out=None, ldo=None):
    n = t_f32.numel()
    if out is None:
        out = torch.empty((n, dim), device=t_f32.device, dtype=bf16)
        ldo = dim
    _lib.check(_lib.load().b2_timestep_embedding(_p(t_f32), _p(out), n, dim, ldo, _stream()), "timestep_embedding")
    return out


# --------------------------------------------------------------------------------------- loss side
def randn(n, seed_offset, stream_id, round_bf16=True, out=None):
    if out is None:
        out = torch.empty(n, device=seed_offset.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_randn(_p(out), n, _p(seed_offset), int(stream_id), int(round_bf16), _stream()), "randn")
    return out


def philox_advance(seed_offset, inc=1):
    _lib.check(_lib.load().b2_philox_advance(_p(seed_offset), int(inc), _stream()), "philox_advance")


def make_noisy(x, eps, sigma_or_t, mode, v_prediction, clamp_ztsnr, B, Cc, HW, Cpad):
    noisy = torch.empty((B * HW, Cpad), device=x.device, dtype=bf16)
    target = torch.empty((B, Cc, HW), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().b2_make_noisy(_p(x), _p(eps), _p(sigma_or_t), mode, int(v_prediction), int(clamp_ztsnr),
                                        _p(noisy), _p(target), B, Cc, HW, Cpad, _stream()), "make_noisy")
    return noisy, target


def mse_loss(pred, target, weight, loss_sum, dpred, gscale, B, Cc, HW, Cpad):
    _lib.check(_lib.load().b2_mse_loss(_p(pred), _p(target), _p(weight), _p(loss_sum), _p(dpred), float(gscale), B, Cc,
                                      HW, Cpad, _stream()), "mse_loss")


def finalize_loss(loss_sum, count, scale, loss_out, ok, dpred=None):
    _lib.check(_lib.load().b2_finalize_loss(_p(loss_sum), float(count), float(scale), _p(loss_out), _p(ok), _p(dpred),
                                           0 if dpred is None else dpred.numel(), _stream()), "finalize_loss")


def abs_sq_sums(x, out, period=0, valid=0):
    _lib.check(_lib.load().b2_abs_sq_sums(_p(x), int(x.dtype == torch.float32), x.numel(), period, valid, _p(out),
                                         _stream()), "abs_sq_sums")


def scale_bf16(x, dev_scale=None, host_scale=1.0):
    _lib.check(_lib.load().b2_scale_bf16(_p(x), x.numel(), _p(dev_scale), float(host_scale), _stream()), "scale_bf16")


def sumsq(g, out):
    _lib.check(_lib.load().b2_sumsq(_p(g), g.numel(), _p(out), _stream()), "sumsq")


def adamw(p, master, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, step=1, gnorm_sq=None,
          max_norm=0.0, grad_scale=1.0, dev_step=None):
    """step >= 1: host step; step = 0: read the step from the device counter `dev_step` (uint64, CUDA-graph safe)."""
    _lib.check(_lib.load().b2_adamw(_p(p), _p(master), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps,
                                   weight_decay, int(step), _p(gnorm_sq), float(max_norm), float(grad_scale),
                                   _p(dev_step), _stream()), "adamw")


def adamw_bf16(p, g, m, v, shift, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=1, gnorm_sq=None, max_norm=0.0,
               grad_scale=1.0, seed_offset=None, as_written=True, rng_mode=0, test_rand16=None, zero_grad=False):
    """Fused AdamWBF16 step over flat bf16 buffers (reference: adamw_bfloat16/__init__.py:150-197); `zero_grad` clears
    `g` in the same pass (the `optimizer.zero_grad()` that follows every step)."""
    _need_cuda(p, g, m, v, shift)
    _lib.check(_lib.load().b2_adamw_bf16(_p(p), _p(g), _p(m), _p(v), _p(shift), p.numel(), lr, beta1, beta2, eps,
                                        int(step), _p(gnorm_sq), float(max_norm), float(grad_scale), _p(seed_offset),
                                        int(as_written), int(rng_mode), _p(test_rand16), int(bool(zero_grad)), _stream()),
               "adamw_bf16")


def axpy_bf16(y, x, alpha):
    _lib.check(_lib.load().b2_axpy_bf16(_p(y), _p(x), y.numel(), float(alpha), _stream()), "axpy_bf16")
