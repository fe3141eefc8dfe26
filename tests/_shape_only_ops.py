"""Shape-only stand-ins for sdxl_training_improvements_b200.ops (CPU tests of the tape / plan / step protocol): outputs have
the right shape on the inputs' device, nothing is computed.  Test infrastructure only."""
import torch

from sdxl_training_improvements_b200 import ops as real_ops

bf16 = torch.bfloat16


class ShapeOnlyOps:
    """Stand-ins for sdxl_training_improvements_b200.ops: outputs have the right shape on the inputs' device, nothing runs."""

    conv_out_hw = staticmethod(real_ops.conv_out_hw)

    @staticmethod
    def conv3x3_implicit_ok(B, H, W, Cin, Cout):
        return Cin % 64 == 0 and W % 8 == 0  # which path is taken does not change the tape

    def __getattr__(self, name):  # everything that only writes into buffers it was given
        return lambda *a, **k: None

    @staticmethod
    def _e(ref, *shape, dtype=bf16):
        return torch.empty(shape, device=ref.device, dtype=dtype)

    def linear_fwd(self, x, W, bias=None, residual=None, out=None, **k):
        return out if out is not None else self._e(x, x.shape[0], W.shape[0])

    def gemm_raw(self, A, B_, D, *a, **k):
        return D

    def conv3x3_fwd(self, x, Wk, B, H, W, Cin, Cout, **k):
        return self._e(x, B * H * W, Cout)

    def im2col3x3(self, x, B, H, W, Cc, stride=1, upsample=False, out=None):
        return out

    def gn_stats(self, x, B, HW, Cc, G, eps):
        return self._e(x, B * G, dtype=torch.float32), self._e(x, B * G, dtype=torch.float32)

    def gn_apply(self, x, *a, **k):
        return torch.empty_like(x)

    def ln_fwd(self, x, gamma, beta, eps=1e-5):
        return torch.empty_like(x), self._e(x, x.shape[0], dtype=torch.float32), self._e(x, x.shape[0], dtype=torch.float32)

    @staticmethod
    def xattn_q_core_ok(B, n_q, n_k, C):
        return False  # the fused and the two-launch cross-attention record the same tape

    @staticmethod
    def linear_geglu_ok(M, F, K):
        return F % 128 == 0  # which path is taken does not change the tape length

    def linear_geglu_fwd(self, x, W1, b1, F):
        return self._e(x, x.shape[0], 2 * F), self._e(x, x.shape[0], F)

    def geglu_fwd(self, u, F):
        return self._e(u, u.shape[0], F)

    def geglu_bwd(self, u, dz, F, dbias32=None):
        return torch.empty_like(u)

    @staticmethod
    def linear_dgrad_geglu_ok(M, F, C):
        return M >= 256 and F % 128 == 0  # as the library decides it: same tape either way

    def linear_dgrad_geglu(self, dy, W2, u, F):
        return torch.empty_like(u)

    def silu_fwd(self, x):
        return torch.empty_like(x)

    def timestep_embedding(self, t, dim, **k):
        return self._e(t, t.numel(), dim)

    def attn_fwd(self, q, k, v, B, H, n_q, n_k, scale, out=None):
        return self._e(q, B * n_q, H * 64), self._e(q, B, H, (n_q + 127) // 128 * 128, dtype=torch.float32)
