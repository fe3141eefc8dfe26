"""A few launches of one b2_gemm shape for `ncu --set full` (python tools/prof_gemm.py M N K [a_mn b_mn tile])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdxl_training_improvements_b200 import ops

M, N, K = (int(x) for x in sys.argv[1:4])
a_mn, b_mn, tile = (int(x) for x in sys.argv[4:7]) if len(sys.argv) > 6 else (0, 0, 0)
bf16 = torch.bfloat16
A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(bf16)
B = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.03).to(bf16)
D = torch.empty(M, N, device="cuda", dtype=bf16)
for _ in range(4):
    ops.gemm_raw(A, B, D, M, N, K, a_mn=a_mn, b_mn=b_mn, lda=A.stride(0), ldb=B.stride(0), ldd=N, tile_n=tile)
torch.cuda.synchronize()
print("done")
