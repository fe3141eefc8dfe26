"""Event trace of one CTA of attn_bwd_dkv3_kernel (clock64 stamps behind b2_attn_set_debug): who waits for whom."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdxl_training_improvements_b200 import _lib, ops

bf16 = torch.bfloat16
for (B, H, n) in ((4, 20, 1024), (4, 10, 4096)):
    Cc = H * 64
    qkv = torch.randn(B * n, 3 * Cc, device="cuda").to(bf16)
    q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
    do = torch.randn(B * n, Cc, device="cuda").to(bf16)
    dqkv = torch.empty_like(qkv)
    o, lse = ops.attn_fwd(q, k, v, B, H, n, n, 0.125)
    for _ in range(2):
        ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:], B, H, n, n, 0.125)
    torch.cuda.synchronize()
    ctr = torch.zeros(512, device="cuda", dtype=torch.int64)
    _lib.load().b2_attn_set_debug(ctr.data_ptr())
    ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:], B, H, n, n, 0.125)
    torch.cuda.synchronize()
    _lib.load().b2_attn_set_debug(None)
    c = ctr.tolist()
    nb = min(64, n // 64)
    t0 = min(x for x in c[64:64 + 2 * nb] + c[192:192 + 2 * nb] if x > 0)
    print(f"--- dkv3 n={n}: block | S seen by group | P arrived | P seen by MMA | block issued   (cycles since first event)")
    for i in range(min(nb, 24)):
        print(f"{i:3d} g{i & 1} | {c[192 + 2 * i] - t0:7d} | {c[193 + 2 * i] - t0:7d} | {c[64 + 2 * i] - t0:7d} | {c[65 + 2 * i] - t0:7d}"
              f"   softmax {c[193 + 2 * i] - c[192 + 2 * i]:5d}  handoff P->MMA {c[64 + 2 * i] - c[193 + 2 * i]:5d}  issue {c[65 + 2 * i] - c[64 + 2 * i]:4d}")
