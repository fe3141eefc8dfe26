"""A few launches of the fused attention kernels for `ncu --set full` (python tools/prof_attn.py B H n_q n_k)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdxl_training_improvements_b200 import ops

B, H, n_q, n_k = (int(x) for x in sys.argv[1:5])
bf16 = torch.bfloat16
Cc = H * 64
q = torch.randn(B * n_q, Cc, device="cuda").to(bf16)
kv = torch.randn(B * n_k, 2 * Cc, device="cuda").to(bf16)
k, v = kv[:, :Cc], kv[:, Cc:]
do = torch.randn(B * n_q, Cc, device="cuda").to(bf16)
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
for _ in range(3):
    o, lse = ops.attn_fwd(q, k, v, B, H, n_q, n_k, 0.125)
    ops.attn_bwd(q, k, v, o, lse, do, dq, dkv[:, :Cc], dkv[:, Cc:], B, H, n_q, n_k, 0.125)
torch.cuda.synchronize()
print("done")
