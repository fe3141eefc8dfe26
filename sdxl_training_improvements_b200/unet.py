"""B200-native SDXL UNet: forward + hand-written backward over the C-ABI kernels.

Stands in for `self.model.unet` of the reference (a diffusers `UNet2DConditionModel`; call sites
src/training/trainers/methods/ddpm_trainer.py:320-325, flow_matching_trainer.py:400-405,
src/training/trainers/sdxl_trainer.py:65-70).  The module graph is SURVEY.md Appendix A; nothing here calls a
torch compute op on the hot path — torch supplies device memory, the stream and (for the drop-in surface) one
autograd.Function around the whole network.

Execution model: activations are token-major [B*H*W, C] bf16.  `forward()` runs the kernels and records one
backward closure per operator on a tape; `backward()` replays the tape in reverse.  Gradients w.r.t. activations
are accumulated in place by the kernels' `accumulate` epilogues (no separate add passes for fan-out), parameter
gradients accumulate (`+=`) into the flat bf16 gradient buffer of `ParamStore`, which is what gradient
accumulation and the single NCCL all-reduce operate on.
"""
from __future__ import annotations

import math
import os
from contextlib import contextmanager
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional

import torch

from . import ops
from .params import SDXL_BASE, ParamStore

bf16 = torch.bfloat16
HEAD_DIM = 64
IN_PAD = 8     # latent channels (4) padded to 8 at conv_in: TMA's 16-byte rule
PRED_PAD = 64  # conv_out's 4 output channels padded to 64: N = 64 puts its forward and dgrad GEMMs on the CTA-pair kernel
               # (at N = 8 they ran on the first-generation kernel: 426 + 397 us per step for 3 GFLOP)


class Act:
    """An activation and its (lazily allocated) gradient."""
    __slots__ = ("d", "g", "bias_grad_done")

    def __init__(self, d: torch.Tensor):
        self.d = d
        self.g: Optional[torch.Tensor] = None
        self.bias_grad_done = False  # set by a backward closure that already produced the producing Linear's bias gradient


def _gslot(t: Act):
    """(gradient buffer, accumulate?) for writing a contribution into t.g."""
    if t.g is None:
        t.g = torch.empty_like(t.d)
        return t.g, False
    return t.g, True


def _ceil8(n: int) -> int:
    return (n + 7) // 8 * 8


class UNetEngine:
    """Kernel-level forward/backward of the UNet for one (B, H, W) problem."""

    def __init__(self, store: ParamStore):
        self.store = store
        self.cfg = store.cfg
        self.tape: List[Callable[[], None]] = []
        self._ws: Dict[str, torch.Tensor] = {}
        self._ws_retired: List[torch.Tensor] = []
        # Backward pass: the weight-gradient GEMM (+ bias column sum) of a layer and its input-gradient GEMM are independent.
        # They are issued on two streams (fork after dy, join right after the pair), so that inside the captured CUDA graph
        # they are parallel branches: the second persistent kernel's CTAs take over SMs as the first one's retire — no launch
        # gap between the two, the tail of one filled by the head of the other.  B2_BWD_OVERLAP=0 issues them back to back.
        self._overlap = os.environ.get("B2_BWD_OVERLAP", "1") != "0" and store.flat.is_cuda
        self._side: Optional[torch.cuda.Stream] = None
        st = store
        for pfx in self._attn_prefixes():
            assert st.adjacent(f"{pfx}.attn1.to_q.weight", f"{pfx}.attn1.to_k.weight", f"{pfx}.attn1.to_v.weight")
            assert st.adjacent(f"{pfx}.attn2.to_k.weight", f"{pfx}.attn2.to_v.weight")

    def _attn_prefixes(self):
        return sorted({n.rsplit(".attn1.", 1)[0] for n, _ in self.store.specs if ".attn1.to_q." in n})

    @contextmanager
    def _side_branch(self):
        """`with self._side_branch():` — the enclosed launches run on the side stream, ordered after everything issued so far
        on the current stream; `self._join()` makes the current stream wait for them."""
        if not self._overlap:
            yield
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.store.flat.device)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            yield

    def _join(self):
        if self._overlap and self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    # ------------------------------------------------------------------ workspaces
    def ws(self, key: str, numel: int, dtype) -> torch.Tensor:
        """Grow-only scratch buffers reused by every layer (stream-ordered reuse is safe on one stream)."""
        cur = self._ws.get(key)
        if cur is None or cur.numel() < numel or cur.dtype != dtype:
            dev = self.store.flat.device
            if cur is not None:
                # a CUDA graph captured for a smaller shape keeps the OLD buffer's address: it must stay allocated
                self._ws_retired.append(cur)
            cur = torch.empty(numel, device=dev, dtype=dtype)
            self._ws[key] = cur
        return cur[:numel]

    def reserve_workspaces(self, B, H, W):
        """Pre-size the scratch buffers for a problem (required before CUDA-graph capture)."""
        boc = self.cfg["block_out_channels"]
        heads = self.cfg["num_heads"]
        depth = self.cfg["transformer_layers_per_block"]
        M0 = B * H * W
        maxcol = 0
        maxS = 0
        for lvl, c in enumerate(boc):
            m = M0 >> (2 * lvl)
            cin_max = c + (boc[min(lvl + 1, len(boc) - 1)] if lvl < len(boc) - 1 else c)
            cin_max = max(cin_max, 2 * c, c + (boc[lvl - 1] if lvl > 0 else c))
            maxcol = max(maxcol, m * 9 * cin_max)
            if lvl > 0:
                maxcol = max(maxcol, m * 9 * boc[lvl])  # upsample conv output grid
            if depth[lvl] > 0:
                n = (H >> lvl) * (W >> lvl)
                maxS = max(maxS, B * heads[lvl] * n * _ceil8(n))
        maxcol = max(maxcol, (M0 // 4) * 9 * boc[1] * 4)  # Upsample2D(640) writes a 128^2 grid of 9*640
        self.ws("col", maxcol, bf16)
        self.ws("dcol", maxcol, bf16)

    # ------------------------------------------------------------------ primitive layers
    def _norm_dgb(self, wname: str, bname: str, Cc: int) -> torch.Tensor:
        """fp32 [2C] staging slice the norm backward kernels accumulate (dgamma | dbeta) into."""
        st = self.store
        assert st.small_off[bname] == st.small_off[wname] + Cc, "norm weight / bias must be adjacent in the staging buffer"
        return st.gs(wname, 2 * Cc)

    def linear(self, x: Act, wname: str, N: int, K: int, bname: Optional[str] = None, residual: Optional[Act] = None,
               need_dx: bool = True, res_ld0: Optional[torch.Tensor] = None) -> Act:
        st = self.store
        Wt = st.w(wname, N, K)
        bias = st.v(bname) if bname else None
        if res_ld0 is not None:  # broadcast row added through the residual port (ldr = 0)
            y = torch.empty((x.d.shape[0], N), device=x.d.device, dtype=bf16)
            ops.gemm_raw(x.d, Wt, y, x.d.shape[0], N, K, lda=x.d.stride(0), ldb=K, ldd=N, bias=bias,
                         residual=res_ld0, ldr=0)
        else:
            y = ops.linear_fwd(x.d, Wt, bias=bias, residual=residual.d if residual is not None else None)
        out = Act(y)

        def bwd():
            dy = out.g
            gW = st.g(wname, N, K)
            gb = st.gs(bname) if bname else None
            with self._side_branch():  # weight + bias gradients beside the input gradient
                ops.linear_wgrad(dy, x.d, gW, accumulate=True)
                if bname and not out.bias_grad_done:
                    ops.colsum_f32(dy, gb)
            if need_dx:
                buf, acc = _gslot(x)
                ops.linear_dgrad(dy, Wt, buf, acc)
            self._join()
            if residual is not None:
                self._add_grad(residual, dy)

        self.tape.append(bwd)
        return out

    def _linear_geglu(self, x: Act, wname: str, bname: str, F: int, K: int):
        """GEGLU up-projection with the gate fused into the GEMM epilogue: returns (u, z) Acts; the backward closure is the
        plain Linear's (weight / bias / input gradients from u.g, which the caller's geglu backward fills)."""
        st = self.store
        Wt = st.w(wname, 2 * F, K)
        ud, zd = ops.linear_geglu_fwd(x.d, Wt, st.v(bname), F)
        u, z = Act(ud), Act(zd)

        def bwd():
            dy = u.g
            gW, gb = st.g(wname, 2 * F, K), st.gs(bname)
            with self._side_branch():
                ops.linear_wgrad(dy, x.d, gW, accumulate=True)
                if not u.bias_grad_done:
                    ops.colsum_f32(dy, gb)
            buf, acc = _gslot(x)
            ops.linear_dgrad(dy, Wt, buf, acc)
            self._join()

        self.tape.append(bwd)
        return u, z

    def _add_grad(self, t: Act, dy: torch.Tensor):
        if t.g is None:
            t.g = dy  # alias: dy's readers are all enqueued before any later in-place accumulation
        else:
            ops.add(t.g, dy, t.g)

    def groupnorm(self, x: Act, B, HW, Cc, wname, bname, eps, silu) -> Act:
        st = self.store
        G = self.cfg["norm_num_groups"]
        gamma, beta = st.v(wname), st.v(bname)
        mean, rstd = ops.gn_stats(x.d, B, HW, Cc, G, eps)
        out = Act(ops.gn_apply(x.d, mean, rstd, gamma, beta, B, HW, Cc, G, silu))

        def bwd():
            buf, acc = _gslot(x)
            ops.gn_bwd(x.d, out.g, mean, rstd, gamma, beta, B, HW, Cc, G, silu, self._norm_dgb(wname, bname, Cc), buf, acc)

        self.tape.append(bwd)
        return out

    def layernorm(self, x: Act, Cc, wname, bname) -> Act:
        st = self.store
        gamma, beta = st.v(wname), st.v(bname)
        y, mean, rstd = ops.ln_fwd(x.d, gamma, beta, 1e-5)
        out = Act(y)

        def bwd():
            buf, acc = _gslot(x)
            dgb = self._norm_dgb(wname, bname, Cc)
            with self._side_branch():  # dgamma / dbeta beside dx: two kernels over the same x / dy
                ops.ln_bwd(x.d, out.g, gamma, mean, rstd, dgb, None, False, parts=2)
            ops.ln_bwd(x.d, out.g, gamma, mean, rstd, dgb, buf, acc, parts=1)
            self._join()

        self.tape.append(bwd)
        return out

    def conv3x3(self, x: Act, B, H, W, Cin, Cout, wname, bname, stride=1, up=False, rowbias: Optional[Act] = None,
                residual: Optional[Act] = None, need_dx=True, Wk=None, gWk=None, bias_t=None, Npad=None) -> Act:
        """3x3 conv as im2col (TMA-friendly K-major operand) + tcgen05 GEMM.  `rowbias` [B, Cout] already contains the
        conv bias (time-embedding projection path of ResnetBlock2D.conv1)."""
        st = self.store
        Ho, Wo = ops.conv_out_hw(H, W, stride, up)
        M = B * Ho * Wo
        K = 9 * Cin
        N = Npad or Cout
        implicit = (stride == 1 and not up and Wk is None and Npad is None
                    and ops.conv3x3_implicit_ok(B, H, W, Cin, Cout))
        Wk = st.w(wname, Cout, K) if Wk is None else Wk
        gWk_given = gWk
        res = residual.d if residual is not None else None
        if implicit:
            # implicit GEMM: TMA gathers the shifted pixel blocks straight from the NHWC activation (no col buffer)
            if rowbias is not None:
                y = ops.conv3x3_fwd(x.d, Wk, B, H, W, Cin, Cout, bias=rowbias.d, residual=res, bias_per_sample=True)
            else:
                y = ops.conv3x3_fwd(x.d, Wk, B, H, W, Cin, Cout, bias=st.v(bname), residual=res)
        else:
            col = ops.im2col3x3(x.d, B, H, W, Cin, stride, up, out=self.ws("col", M * K, bf16).view(M, K))
            y = torch.empty((M, N), device=x.d.device, dtype=bf16)
            if rowbias is not None:
                ops.gemm_raw(col, Wk, y, M, N, K, lda=K, ldb=K, ldd=N, bias=rowbias.d, residual=res, ldr=N,
                             bias_rows_per_group=Ho * Wo, bias_group_stride=N)
            else:
                bias = bias_t if bias_t is not None else st.v(bname)
                ops.gemm_raw(col, Wk, y, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, residual=res, ldr=N)
        out = Act(y)

        def bwd():
            dy = out.g
            # the gradient view is taken HERE, not in forward: ParamStore logs when each gradient is written (dp.py)
            gWk = st.g(wname, Cout, K) if gWk_given is None else gWk_given
            if implicit:
                with self._side_branch():  # joined at the end of this closure
                    ops.conv3x3_wgrad(dy, x.d, gWk, B, H, W, Cin, Cout, accumulate=True)
            else:
                colb = ops.im2col3x3(x.d, B, H, W, Cin, stride, up, out=self.ws("col", M * K, bf16).view(M, K))
                ops.gemm_raw(dy, colb, gWk, N, K, M, a_mn=True, b_mn=True, lda=N, ldb=K, ldd=K, accumulate=True,
                             allow_split_k=True)
            if rowbias is not None:
                if rowbias.g is None:
                    rowbias.g = torch.empty_like(rowbias.d)
                    acc = False
                else:
                    acc = True
                ops.colsum_groups(dy, rowbias.g, B, Ho * Wo, accumulate=acc)
            elif bias_t is None:
                ops.colsum_f32(dy, st.gs(bname))
            if need_dx:
                buf, acc = _gslot(x)
                if implicit:
                    ops.conv3x3_dgrad(dy, Wk, buf, B, H, W, Cin, Cout, accumulate=acc)
                else:
                    dcol = self.ws("dcol", M * K, bf16).view(M, K)
                    ops.gemm_raw(dy, Wk, dcol, M, K, N, b_mn=True, lda=N, ldb=K, ldd=K)
                    ops.col2im3x3(dcol, buf, B, H, W, Cin, stride, up, accumulate=acc)
            if implicit:
                self._join()
            if residual is not None:
                self._add_grad(residual, dy)

        self.tape.append(bwd)
        return out

    def concat(self, a: Act, b: Act, M, Ca, Cb) -> Act:
        y = torch.empty((M, Ca + Cb), device=a.d.device, dtype=bf16)
        ops.copy2d(a.d, y, M, Ca, Ca, Ca + Cb)
        ops.copy2d(b.d, y[:, Ca:], M, Cb, Cb, Ca + Cb)
        out = Act(y)

        def bwd():
            ga, acca = _gslot(a)
            ops.copy2d(out.g, ga, M, Ca, Ca + Cb, Ca, accumulate=acca)
            gb, accb = _gslot(b)
            ops.copy2d(out.g[:, Ca:], gb, M, Cb, Ca + Cb, Cb, accumulate=accb)

        self.tape.append(bwd)
        return out

    def silu(self, x: Act) -> Act:
        out = Act(ops.silu_fwd(x.d))

        def bwd():
            buf, acc = _gslot(x)
            ops.silu_bwd(x.d, out.g, buf, acc)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ composite blocks
    def resnet(self, x: Act, emb_silu: Act, B, H, W, Cin, Cout, pfx) -> Act:
        st = self.store
        temb = self.cfg["block_out_channels"][0] * 4
        eps = self.cfg["norm_eps"]
        a1 = self.groupnorm(x, B, H * W, Cin, f"{pfx}.norm1.weight", f"{pfx}.norm1.bias", eps, True)
        # time_emb_proj(SiLU(emb)) + conv1.bias  ->  one [B, Cout] row per sample, added in conv1's epilogue
        tproj = self.linear(emb_silu, f"{pfx}.time_emb_proj.weight", Cout, temb, f"{pfx}.time_emb_proj.bias",
                            res_ld0=st.v(f"{pfx}.conv1.bias"))
        h1 = self.conv3x3(a1, B, H, W, Cin, Cout, f"{pfx}.conv1.weight", f"{pfx}.conv1.bias", rowbias=tproj)

        def conv1_bias_grad():  # d conv1.bias = sum over samples of d tproj
            ops.colsum_f32(tproj.g, st.gs(f"{pfx}.conv1.bias"))

        # runs after conv1's backward (which fills tproj.g) and before the time_emb_proj linear's backward
        self.tape.insert(len(self.tape) - 1, conv1_bias_grad)
        a2 = self.groupnorm(h1, B, H * W, Cout, f"{pfx}.norm2.weight", f"{pfx}.norm2.bias", eps, True)
        if Cin != Cout:
            sc = self.linear(x, f"{pfx}.conv_shortcut.weight", Cout, Cin, f"{pfx}.conv_shortcut.bias")
        else:
            sc = x
        return self.conv3x3(a2, B, H, W, Cout, Cout, f"{pfx}.conv2.weight", f"{pfx}.conv2.bias", residual=sc)

    def attention(self, xn: Act, res: Act, B, n, Cc, pfx, ctx: Optional[torch.Tensor], n_ctx: int) -> Act:
        """x + to_out(softmax(Q K^T / 8) V): fused QKV (self) / Q + KV (cross) projections, the flash-style tcgen05
        attention core (no n x n tensor in HBM), then to_out + bias + residual in one GEMM epilogue."""
        st = self.store
        heads = Cc // HEAD_DIM
        scale = 1.0 / math.sqrt(HEAD_DIM)
        is_self = ctx is None
        if is_self:
            nk = n
            Wqkv = st.w(f"{pfx}.to_q.weight", 3 * Cc, Cc)
            qkv = ops.linear_fwd(xn.d, Wqkv)  # [M, 3C]
            q_t, k_t, v_t = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
        else:
            nk = n_ctx
            kv_bwd = self._kv_wgrad.pop(Cc, None)
            if kv_bwd is not None:
                self.tape.append(kv_bwd)
            Wq = st.w(f"{pfx}.to_q.weight", Cc, Cc)
            kv, dkv = self._kv_slices[pfx]  # [B*77, 2C] views of the grouped projection (computed once per forward)
            k_t, v_t = kv[:, :Cc], kv[:, Cc:]
            if ops.xattn_q_core_ok(B, n, nk, Cc):
                # ONE launch: to_q GEMM with the 77-key attention core in its epilogue (csrc/xattn.cu); Q is written for the
                # backward pass but never read back
                q, O, lse = ops.xattn_q_core(xn.d, Wq, k_t, v_t, B, n, nk, scale)
            else:
                q = ops.linear_fwd(xn.d, Wq)
            q_t = q
        if is_self or not ops.xattn_q_core_ok(B, n, nk, Cc):
            O, lse = ops.attn_fwd(q_t, k_t, v_t, B, heads, n, nk, scale)
        Oa = Act(O)
        out = self.linear(Oa, f"{pfx}.to_out.0.weight", Cc, Cc, f"{pfx}.to_out.0.bias", residual=res)

        def bwd():
            dO = Oa.g
            if is_self:
                dqkv = torch.empty_like(qkv)
                dq_t, dk_t, dv_t = dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:]
            else:
                dq = torch.empty_like(q)
                dq_t, dk_t, dv_t = dq, dkv[:, :Cc], dkv[:, Cc:]  # slice of the grouped dKV buffer (written, not accumulated)
            ops.attn_bwd(q_t, k_t, v_t, O, lse, dO, dq_t, dk_t, dv_t, B, heads, n, nk, scale)
            buf, acc = _gslot(xn)
            if is_self:
                gW = st.g(f"{pfx}.to_q.weight", 3 * Cc, Cc)
                with self._side_branch():
                    ops.linear_wgrad(dqkv, xn.d, gW, accumulate=True)
                ops.linear_dgrad(dqkv, Wqkv, buf, acc)
            else:
                gW = st.g(f"{pfx}.to_q.weight", Cc, Cc)
                with self._side_branch():
                    ops.linear_wgrad(dq, xn.d, gW, accumulate=True)
                ops.linear_dgrad(dq, Wq, buf, acc)
            self._join()
                # the K/V projection weight gradients of all blocks run as ONE GEMM per width (see _project_context)

        # attention-core backward must run after to_out's backward (already on the tape) -> append
        self.tape.append(bwd)
        # ...but the tape is replayed in reverse, so the closure appended last runs first: swap the two
        self.tape[-1], self.tape[-2] = self.tape[-2], self.tape[-1]
        return out

    def transformer_block(self, h: Act, B, n, Cc, pfx, ctx, n_ctx) -> Act:
        n1 = self.layernorm(h, Cc, f"{pfx}.norm1.weight", f"{pfx}.norm1.bias")
        h1 = self.attention(n1, h, B, n, Cc, f"{pfx}.attn1", None, 0)
        n2 = self.layernorm(h1, Cc, f"{pfx}.norm2.weight", f"{pfx}.norm2.bias")
        h2 = self.attention(n2, h1, B, n, Cc, f"{pfx}.attn2", ctx, n_ctx)
        n3 = self.layernorm(h2, Cc, f"{pfx}.norm3.weight", f"{pfx}.norm3.bias")
        if ops.linear_geglu_ok(n3.d.shape[0], 4 * Cc, Cc):
            u, z = self._linear_geglu(n3, f"{pfx}.ff.net.0.proj.weight", f"{pfx}.ff.net.0.proj.bias", 4 * Cc, Cc)
        else:
            u = self.linear(n3, f"{pfx}.ff.net.0.proj.weight", 8 * Cc, Cc, f"{pfx}.ff.net.0.proj.bias")
            z = Act(ops.geglu_fwd(u.d, 4 * Cc))

        def geglu_bwd():
            if u.g is not None:  # the down-projection's dgrad GEMM already pushed dz through the gate (see _linear_ff2)
                return
            # the up-projection's bias gradient (column sums of du) falls out of the same pass; its Linear closure skips it
            u.g = ops.geglu_bwd(u.d, z.g, 4 * Cc, dbias32=self.store.gs(f"{pfx}.ff.net.0.proj.bias"))
            u.bias_grad_done = True

        self.tape.append(geglu_bwd)
        if ops.linear_dgrad_geglu_ok(z.d.shape[0], 4 * Cc, Cc):
            return self._linear_ff2(z, u, f"{pfx}.ff.net.2.weight", f"{pfx}.ff.net.2.bias", Cc, 4 * Cc, h2)
        return self.linear(z, f"{pfx}.ff.net.2.weight", Cc, 4 * Cc, f"{pfx}.ff.net.2.bias", residual=h2)

    def _linear_ff2(self, z: Act, u: Act, wname: str, bname: str, N: int, F: int, residual: Act) -> Act:
        """FeedForward down-projection whose backward fuses the GEGLU backward into the input-gradient GEMM: du = f(dy W2, u)
        comes out of the dgrad epilogue (csrc/gemm2.cu, GEGLU-backward mode); dz is never written and the gate's own
        backward closure finds u.g already filled."""
        st = self.store
        Wt = st.w(wname, N, F)
        y = ops.linear_fwd(z.d, Wt, bias=st.v(bname), residual=residual.d)
        out = Act(y)

        def bwd():
            dy = out.g
            gW, gb = st.g(wname, N, F), st.gs(bname)
            with self._side_branch():  # weight + bias gradients beside the input gradient
                ops.linear_wgrad(dy, z.d, gW, accumulate=True)
                ops.colsum_f32(dy, gb)
            u.g = ops.linear_dgrad_geglu(dy, Wt, u.d, F)
            self._join()
            self._add_grad(residual, dy)

        self.tape.append(bwd)
        return out

    def transformer(self, x: Act, B, H, W, Cc, depth, pfx, ctx, n_ctx) -> Act:
        a = self.groupnorm(x, B, H * W, Cc, f"{pfx}.norm.weight", f"{pfx}.norm.bias", 1e-6, False)
        h = self.linear(a, f"{pfx}.proj_in.weight", Cc, Cc, f"{pfx}.proj_in.bias")
        for k in range(depth):
            h = self.transformer_block(h, B, H * W, Cc, f"{pfx}.transformer_blocks.{k}", ctx, n_ctx)
        return self.linear(h, f"{pfx}.proj_out.weight", Cc, Cc, f"{pfx}.proj_out.bias", residual=x)

    def _project_context(self, ctx: torch.Tensor):
        """Cross-attention K / V of EVERY transformer block in one GEMM per width: the text embeddings do not change
        through the network, and ParamStore stacks the [to_k; to_v] weights of all blocks of a width into one
        [L * 2C, 2048] matrix.  140 GEMMs with M = B*77 = 308 rows (and their 70 weight-gradient GEMMs with K = 308)
        become 2 + 2 launches.  Backward: each block's attention kernel writes its dK / dV slice of `dkv_all`; the
        closure built here accumulates all weight gradients of a width at once."""
        st = self.store
        self._kv_slices = {}
        self._kv_wgrad = {}
        for Cc, pfxs in st.kv_groups.items():
            W = st.kv_group_view(Cc)
            kv_all = ops.linear_fwd(ctx, W)                      # [B*77, L*2C]
            dkv_all = torch.empty_like(kv_all)
            for l, pfx in enumerate(pfxs):
                self._kv_slices[pfx] = (kv_all[:, l * 2 * Cc:(l + 1) * 2 * Cc], dkv_all[:, l * 2 * Cc:(l + 1) * 2 * Cc])

            def bwd(Cc=Cc, dkv_all=dkv_all):
                ops.linear_wgrad(dkv_all, ctx, st.kv_group_view(Cc, grad=True), accumulate=True)

            # recorded on the tape by the FIRST cross-attention of this width in the forward pass (see attention()), so
            # that in the reverse replay it runs right after the last block that writes into dkv_all — not at the very end
            # of the backward pass: the stacked K/V gradients (0.73 GB) then join an early data-parallel exchange chunk
            self._kv_wgrad[Cc] = bwd

    def embeddings(self, t_f32: torch.Tensor, pooled: torch.Tensor, time_ids_f32: torch.Tensor, B) -> Act:
        """time_embedding(sinus(t)) + add_embedding([pooled, sinus(time_ids)]) -> SiLU (shared by all resnets)."""
        cfg = self.cfg
        boc0 = cfg["block_out_channels"][0]
        temb = boc0 * 4
        adim = cfg["addition_time_embed_dim"]
        pin = cfg["projection_class_embeddings_input_dim"]
        pooled_dim = pin - 6 * adim
        dev = t_f32.device
        ts = Act(ops.timestep_embedding(t_f32, boc0))
        e1 = self.linear(ts, "time_embedding.linear_1.weight", temb, boc0, "time_embedding.linear_1.bias", need_dx=False)
        e1s = self.silu(e1)
        et = self.linear(e1s, "time_embedding.linear_2.weight", temb, temb, "time_embedding.linear_2.bias")
        tid = ops.timestep_embedding(time_ids_f32.reshape(-1), adim)  # [B*6, adim] == [B, 6*adim]
        add_in = torch.empty((B, pin), device=dev, dtype=bf16)
        ops.copy2d(pooled, add_in, B, pooled_dim, pooled_dim, pin)
        ops.copy2d(tid, add_in[:, pooled_dim:], B, 6 * adim, 6 * adim, pin)
        a1 = self.linear(Act(add_in), "add_embedding.linear_1.weight", temb, pin, "add_embedding.linear_1.bias",
                         need_dx=False)
        a1s = self.silu(a1)
        emb = self.linear(a1s, "add_embedding.linear_2.weight", temb, temb, "add_embedding.linear_2.bias", residual=et)
        return self.silu(emb)

    # ------------------------------------------------------------------ whole network
    def forward(self, x_nhwc8: torch.Tensor, t_f32: torch.Tensor, ctx: torch.Tensor, pooled: torch.Tensor,
                time_ids_f32: torch.Tensor, B: int, H: int, W: int) -> Act:
        """x_nhwc8: [B*H*W, IN_PAD] bf16 (4 latent channels + zero pad).  Returns pred as Act([B*H*W, PRED_PAD])."""
        cfg = self.cfg
        st = self.store
        self.tape = []
        boc = cfg["block_out_channels"]
        depth = cfg["transformer_layers_per_block"]
        L = cfg["layers_per_block"]
        n_ctx = ctx.shape[0] // B
        cin = cfg["in_channels"]
        assert cin <= 8 and cfg["out_channels"] <= 8
        dev = x_nhwc8.device

        self._project_context(ctx)  # first on the tape -> its weight-gradient GEMMs run last in the backward pass
        emb_silu = self.embeddings(t_f32, pooled, time_ids_f32, B)

        # conv_in: weight repacked to Cin padded to 8 (K = 72)
        w_in = st.v("conv_in.weight").view(boc[0] * 9, cin)
        wk_in = torch.zeros((boc[0], 72), device=dev, dtype=bf16)
        ops.copy2d_any(w_in, wk_in.view(boc[0] * 9, 8), boc[0] * 9, cin, cin, 8)
        gwk_in = torch.zeros((boc[0], 72), device=dev, dtype=bf16)
        h = self.conv3x3(Act(x_nhwc8), B, H, W, 8, boc[0], "conv_in.weight", "conv_in.bias", need_dx=False,
                         Wk=wk_in, gWk=gwk_in)

        def conv_in_wgrad():
            ops.copy2d_any(gwk_in.view(boc[0] * 9, 8), st.gv("conv_in.weight").view(boc[0] * 9, cin), boc[0] * 9, cin, 8,
                           cin, accumulate=True)

        self.tape.insert(len(self.tape) - 1, conv_in_wgrad)  # runs after conv_in's backward

        skips = [(h, boc[0])]
        ch, cH, cW = boc[0], H, W
        for i, cout in enumerate(boc):
            last = i == len(boc) - 1
            for j in range(L):
                h = self.resnet(h, emb_silu, B, cH, cW, ch, cout, f"down_blocks.{i}.resnets.{j}")
                ch = cout
                if depth[i] > 0:
                    h = self.transformer(h, B, cH, cW, ch, depth[i], f"down_blocks.{i}.attentions.{j}", ctx, n_ctx)
                skips.append((h, ch))
            if not last:
                h = self.conv3x3(h, B, cH, cW, ch, ch, f"down_blocks.{i}.downsamplers.0.conv.weight",
                                 f"down_blocks.{i}.downsamplers.0.conv.bias", stride=2)
                cH, cW = ops.conv_out_hw(cH, cW, 2, False)
                skips.append((h, ch))

        h = self.resnet(h, emb_silu, B, cH, cW, ch, ch, "mid_block.resnets.0")
        h = self.transformer(h, B, cH, cW, ch, depth[-1], "mid_block.attentions.0", ctx, n_ctx)
        h = self.resnet(h, emb_silu, B, cH, cW, ch, ch, "mid_block.resnets.1")

        rev, rdepth = list(reversed(boc)), list(reversed(depth))
        for i, cout in enumerate(rev):
            last = i == len(rev) - 1
            for j in range(L + 1):
                s, sc = skips.pop()
                h = self.concat(h, s, B * cH * cW, ch, sc)
                h = self.resnet(h, emb_silu, B, cH, cW, ch + sc, cout, f"up_blocks.{i}.resnets.{j}")
                ch = cout
                if rdepth[i] > 0:
                    h = self.transformer(h, B, cH, cW, ch, rdepth[i], f"up_blocks.{i}.attentions.{j}", ctx, n_ctx)
            if not last:
                h = self.conv3x3(h, B, cH, cW, ch, ch, f"up_blocks.{i}.upsamplers.0.conv.weight",
                                 f"up_blocks.{i}.upsamplers.0.conv.bias", up=True)
                cH, cW = cH * 2, cW * 2

        a = self.groupnorm(h, B, cH * cW, ch, "conv_norm_out.weight", "conv_norm_out.bias", cfg["norm_eps"], True)
        # conv_out: Cout padded to PRED_PAD rows (zero rows / zero bias): the GEMM's N and the dgrad's K then satisfy TMA and
        # the CTA-pair kernel's minimum tile
        co = cfg["out_channels"]
        K = 9 * ch
        NP = PRED_PAD
        wk_out = torch.zeros((NP, K), device=dev, dtype=bf16)
        ops.copy2d(st.w("conv_out.weight", co, K), wk_out, co, K, K, K)
        gwk_out = torch.zeros((NP, K), device=dev, dtype=bf16)
        b_out = torch.zeros(NP, device=dev, dtype=bf16)
        ops.copy2d_any(st.v("conv_out.bias").view(1, co), b_out.view(1, NP), 1, co, co, NP)
        pred = self.conv3x3(a, B, cH, cW, ch, co, "conv_out.weight", "conv_out.bias", Wk=wk_out, gWk=gwk_out,
                            bias_t=b_out, Npad=NP)

        def conv_out_pgrad():
            ops.copy2d(gwk_out, st.g("conv_out.weight", co, K), co, K, K, K, accumulate=True)
            gb = torch.zeros(NP, device=dev, dtype=bf16)
            ops.colsum(pred.g, gb, accumulate=False)
            ops.copy2d_any(gb.view(1, NP), st.gv("conv_out.bias").view(1, co), 1, co, NP, co, accumulate=True)

        self.tape.insert(len(self.tape) - 1, conv_out_pgrad)
        self._out = pred
        return pred

    def detach_tape(self):
        """Hand the recorded tape to the caller (one tape per forward call)."""
        t = (self.tape, self._out)
        self.tape, self._out = [], None
        return t

    def backward(self, dpred_nhwc8: torch.Tensor, saved=None, cuts=None, on_cut=None):
        """Replay a tape: fills ParamStore.grad (+=).  dpred: [B*H*W, PRED_PAD] bf16 (pad channels must be zero).
        `cuts` (ascending replay positions, the last one = end of tape) + `on_cut(k)`: data-parallel hook — called right
        after position cuts[k] has run, when chunk k of the gradient buffer is final (dp.plan_chunks); the hook flushes
        that chunk's small-parameter gradients and starts its exchange."""
        order = self.backward_begin(dpred_nhwc8, saved)
        if cuts is None:
            self.backward_span(order, 0, len(order))
            self.store.flush_small_grads()  # biases / norm affine: fp32 staging -> flat bf16 gradient buffer (+=)
        else:
            lo = 0
            for k, c in enumerate(cuts):
                self.backward_span(order, lo, c + 1)
                lo = c + 1
                on_cut(k)
            assert lo == len(order), "the last cut must be the end of the tape"
        order.clear()

    def backward_begin(self, dpred_nhwc8: torch.Tensor, saved=None) -> List[Callable[[], None]]:
        """Closures of one forward pass in replay (reverse) order; run them with backward_span()."""
        tape, out = saved if saved is not None else self.detach_tape()
        out.g = dpred_nhwc8
        order = list(reversed(tape))
        tape.clear()
        return order

    def backward_span(self, order, lo: int, hi: int):
        st = self.store
        for i in range(lo, hi):
            st._touch_pos = i
            order[i]()
            order[i] = None  # release the closure's activations as soon as it has run


class _UNetFunction(torch.autograd.Function):
    """Bridges the hand-written backward into torch autograd so `loss.backward()` of the reference loops works
    unchanged (ddpm_trainer.py:268-271, flow_matching_trainer.py:249-252).  The flat-parameter tensor is an input so
    that autograd considers the output differentiable; parameter gradients are written by the kernels directly
    into `p.grad` (accumulating), the returned gradient for it is None."""

    @staticmethod
    def forward(ctx, unet, sample, t_f32, ehs, pooled, time_ids, anchor):
        B, Cc, H, W = sample.shape
        eng = unet.engine
        x8 = ops.nchw_to_nhwc(sample, 8)
        pred = eng.forward(x8, t_f32, ehs, pooled, time_ids, B, H, W)
        ctx.saved_tape = eng.detach_tape()
        ctx.unet = unet
        ctx.shape = (B, Cc, H, W)
        ctx.out_dtype = sample.dtype if sample.dtype in (bf16, torch.float32) else bf16
        return ops.nhwc_to_nchw(pred.d, B, Cc, H, W, PRED_PAD, dtype=ctx.out_dtype)

    @staticmethod
    def backward(ctx, grad_out):
        B, Cc, H, W = ctx.shape
        g8 = ops.nchw_to_nhwc(grad_out.contiguous(), PRED_PAD)
        ctx.unet.engine.backward(g8, ctx.saved_tape)
        ctx.saved_tape = None
        return (None,) * 7


class B200UNet:
    """Duck-typed replacement for `StableDiffusionXL.unet` (src/models/sdxl.py:38-66).

    Surface honoured: `unet(sample, timestep, encoder_hidden_states, added_cond_kwargs=...)
    -> object with .sample`, `parameters()`, `named_parameters()`, `state_dict()/load_state_dict()`, `to()`,
    `train()/eval()`, `zero_grad()`, `save_pretrained()`, no-op `enable_gradient_checkpointing()` and
    `enable_xformers_memory_efficient_attention()` (flow_matching_trainer.py:60-72).
    """

    def __init__(self, cfg: Optional[dict] = None, device="cuda"):
        self.config = dict(SDXL_BASE if cfg is None else cfg)
        self.store = ParamStore(self.config, device=device)
        self.engine = UNetEngine(self.store)
        self.training = True
        self.dtype = bf16
        self._disk_config: Optional[dict] = None
        self._anchor = torch.zeros(1, device=device, requires_grad=True)

    # --- nn.Module-like surface ---
    @property
    def device(self):
        return self.store.flat.device

    def parameters(self):
        return iter(self.store.params.values())

    def named_parameters(self):
        return iter(self.store.params.items())

    def state_dict(self):
        return self.store.state_dict()

    def load_state_dict(self, sd, strict=True):
        return self.store.load_state_dict(sd, strict)

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)) and torch.device(a).type == "cuda":
                if self.store.flat.device != torch.device(a):
                    self.store.to(a)
                    self.engine = UNetEngine(self.store)
                    self._anchor = torch.zeros(1, device=a, requires_grad=True)
            elif isinstance(a, torch.dtype) and a not in (bf16,):
                raise RuntimeError("B200UNet computes in bf16 only (the reference casts the UNet to bf16, "
                                   "sdxl_trainer.py:52-55)")
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def requires_grad_(self, flag=True):
        for p in self.parameters():
            p.requires_grad_(flag)
        return self

    def zero_grad(self, set_to_none: bool = False):
        self.store.grad.zero_()

    def enable_gradient_checkpointing(self):  # 180 GB HBM: activations are kept
        return None

    def enable_xformers_memory_efficient_attention(self, *a, **k):  # attention is our own kernel
        return None

    def diffusers_config(self) -> dict:
        """The `config.json` diffusers' `UNet2DConditionModel.from_pretrained` needs to rebuild this network (the
        reference reloads checkpoints through `StableDiffusionXLPipeline.from_pretrained`, src/models/sdxl.py:25-40).
        A config read by `from_pretrained` is written back unchanged; otherwise the SDXL-architecture fields are emitted
        in diffusers' own vocabulary: `attention_head_dim` is the per-level head COUNT, every level lists a
        `transformer_layers_per_block` entry (levels without attention are `DownBlock2D` / `UpBlock2D` and ignore it)."""
        if getattr(self, "_disk_config", None) is not None:
            return dict(self._disk_config)
        c = self.config
        depth = list(c["transformer_layers_per_block"])
        down = ["CrossAttnDownBlock2D" if d > 0 else "DownBlock2D" for d in depth]
        up = ["CrossAttnUpBlock2D" if d > 0 else "UpBlock2D" for d in reversed(depth)]
        return {
            "_class_name": "UNet2DConditionModel",
            "_diffusers_version": "0.21.0",
            "act_fn": "silu",
            "addition_embed_type": "text_time",
            "addition_embed_type_num_heads": 64,
            "addition_time_embed_dim": c["addition_time_embed_dim"],
            "attention_head_dim": list(c["num_heads"]),
            "block_out_channels": list(c["block_out_channels"]),
            "center_input_sample": False,
            "class_embed_type": None,
            "class_embeddings_concat": False,
            "conv_in_kernel": 3,
            "conv_out_kernel": 3,
            "cross_attention_dim": c["cross_attention_dim"],
            "cross_attention_norm": None,
            "down_block_types": down,
            "downsample_padding": 1,
            "dual_cross_attention": False,
            "encoder_hid_dim": None,
            "encoder_hid_dim_type": None,
            "flip_sin_to_cos": True,
            "freq_shift": 0,
            "in_channels": c["in_channels"],
            "layers_per_block": c["layers_per_block"],
            "mid_block_only_cross_attention": None,
            "mid_block_scale_factor": 1,
            "mid_block_type": "UNetMidBlock2DCrossAttn",
            "norm_eps": c["norm_eps"],
            "norm_num_groups": c["norm_num_groups"],
            "num_attention_heads": None,
            "num_class_embeds": None,
            "only_cross_attention": False,
            "out_channels": c["out_channels"],
            "projection_class_embeddings_input_dim": c["projection_class_embeddings_input_dim"],
            "resnet_out_scale_factor": 1.0,
            "resnet_skip_time_act": False,
            "resnet_time_scale_shift": "default",
            "sample_size": 128,
            "time_cond_proj_dim": None,
            "time_embedding_act_fn": None,
            "time_embedding_dim": None,
            "time_embedding_type": "positional",
            "timestep_post_act": None,
            "transformer_layers_per_block": [max(1, d) for d in depth],
            "up_block_types": up,
            "upcast_attention": None,
            "use_linear_projection": True,
        }

    def save_pretrained(self, path, safe_serialization=True, **kw):
        import json
        import os
        os.makedirs(path, exist_ok=True)
        sd = {k: v.contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file({k: v.cpu() for k, v in sd.items()}, os.path.join(path, "diffusion_pytorch_model.safetensors"),
                      metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(path, "diffusion_pytorch_model.bin"))
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.diffusers_config(), f, indent=2, sort_keys=True)

    @classmethod
    def from_pretrained(cls, path, device="cuda", subfolder: Optional[str] = None, **kw):
        """Load a diffusers-format UNet directory (what `StableDiffusionXL.save_pretrained` writes per component,
        src/models/sdxl.py:246-288, and what `StableDiffusionXLPipeline.from_pretrained(...).unet` reads,
        src/models/sdxl.py:25-40): `config.json` + `diffusion_pytorch_model.safetensors` (single file or sharded with
        `diffusion_pytorch_model.safetensors.index.json`) or `.bin`.  Weights are diffusers' logical layout (conv OIHW,
        Linear [out, in]); ParamStore re-lays them out (channels-last conv taps, adjacent q/k/v) on the way in."""
        import json
        import os
        root = os.path.join(path, subfolder) if subfolder else path
        cfg = dict(SDXL_BASE)
        cj = os.path.join(root, "config.json")
        disk_config = None
        if os.path.exists(cj):
            with open(cj) as f:
                disk = json.load(f)
            disk_config = dict(disk)
            unsupported = {k: disk[k] for k, want in (("use_linear_projection", True), ("addition_embed_type", "text_time"),
                                                      ("act_fn", "silu"), ("resnet_time_scale_shift", "default"),
                                                      ("flip_sin_to_cos", True), ("freq_shift", 0))
                           if k in disk and disk[k] != want}
            if unsupported:
                raise ValueError(f"B200UNet implements the SDXL UNet architecture only; unsupported config: {unsupported}")
            if "num_heads" not in disk and "attention_head_dim" in disk:  # diffusers names the head COUNT this way
                disk["num_heads"] = disk["attention_head_dim"]
            if "transformer_layers_per_block" in disk and "down_block_types" in disk:
                tl = disk["transformer_layers_per_block"]
                tl = [tl] * len(disk["down_block_types"]) if isinstance(tl, int) else list(tl)
                disk["transformer_layers_per_block"] = [t if "CrossAttn" in b else 0
                                                        for t, b in zip(tl, disk["down_block_types"])]
            for k in cfg:
                if k in disk:
                    cfg[k] = tuple(disk[k]) if isinstance(disk[k], list) else disk[k]
        net = cls(cfg, device=device)
        net._disk_config = disk_config  # written back unchanged by save_pretrained
        st_single = os.path.join(root, "diffusion_pytorch_model.safetensors")
        st_index = st_single + ".index.json"
        sd = {}
        if os.path.exists(st_single):
            from safetensors.torch import load_file
            sd = load_file(st_single)
        elif os.path.exists(st_index):
            from safetensors.torch import load_file
            with open(st_index) as f:
                shards = sorted(set(json.load(f)["weight_map"].values()))
            for sh in shards:
                sd.update(load_file(os.path.join(root, sh)))
        elif os.path.exists(os.path.join(root, "diffusion_pytorch_model.bin")):
            sd = torch.load(os.path.join(root, "diffusion_pytorch_model.bin"), map_location="cpu")
        else:
            raise FileNotFoundError(f"no diffusion_pytorch_model.(safetensors|bin) under {root}")
        net.load_state_dict(sd, strict=True)
        return net

    # --- forward ---
    def __call__(self, sample, timestep, encoder_hidden_states=None, added_cond_kwargs=None, **kw):
        if added_cond_kwargs is None:
            raise ValueError("SDXL UNet needs added_cond_kwargs={'text_embeds','time_ids'}")
        dev = self.device
        B = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=dev)
        t = timestep.to(dev).float().reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        t = t.contiguous()
        ehs = encoder_hidden_states.to(dev, bf16).reshape(-1, encoder_hidden_states.shape[-1]).contiguous()
        pooled = added_cond_kwargs["text_embeds"].to(dev, bf16).reshape(B, -1).contiguous()
        tid = added_cond_kwargs["time_ids"].to(dev).float().reshape(B, -1).contiguous()
        sample = sample.to(dev)
        out = _UNetFunction.apply(self, sample, t, ehs, pooled, tid, self._anchor)
        return SimpleNamespace(sample=out)

    forward = __call__
