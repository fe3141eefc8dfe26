"""Per-phase cycle breakdown of attn_fwd3_kernel (clock64 counters behind b2_attn_set_debug)."""
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
    ops.attn_fwd(q, k, v, B, H, n, n, 0.125)
    torch.cuda.synchronize()
    ctr = torch.zeros(512, device="cuda", dtype=torch.int64)
    _lib.load().b2_attn_set_debug(ctr.data_ptr())
    ops.attn_fwd(q, k, v, B, H, n, n, 0.125)
    torch.cuda.synchronize()
    _lib.load().b2_attn_set_debug(None)
    c = ctr.tolist()
    # backward dK/dV kernel phases
    do = torch.randn(B * n, Cc, device="cuda").to(bf16)
    dqkv = torch.empty_like(qkv)
    o, lse = ops.attn_fwd(q, k, v, B, H, n, n, 0.125)
    ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:], B, H, n, n, 0.125)
    torch.cuda.synchronize()
    ctr2 = torch.zeros(512, device="cuda", dtype=torch.int64)
    _lib.load().b2_attn_set_debug(ctr2.data_ptr())
    ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:], B, H, n, n, 0.125)
    torch.cuda.synchronize()
    _lib.load().b2_attn_set_debug(None)
    c2 = ctr2.tolist()
    nqb = n // 64
    for g in (0, 1):
        d = c2[24 + g * 4: 28 + g * 4]
        cnt = max(1, c2[23]); blocks = (nqb - g + 1) // 2
        print(f"n={n} dkv3 group {g}: per own 64-query block cycles: wait_S {d[0] / cnt / blocks:.0f}, tmem_ld {d[1] / cnt / blocks:.0f}, "
              f"compute {d[2] / cnt / blocks:.0f}, st+arrive {d[3] / cnt / blocks:.0f}")
    nkb = n // 128
    for g in (0, 1):
        d = c[g * 8:(g + 1) * 8]
        cnt = max(1, d[7])
        blocks = (nkb - g + 1) // 2
        names = ["wait_S", "tmem_ld", "compute", "st_wait", "arrive"]
        per = [x / cnt / blocks for x in d[:5]]
        print(f"n={n} group {g}: per own block cycles: " + ", ".join(f"{a} {b:.0f}" for a, b in zip(names, per)) +
              f" | loop total/CTA {d[5] / cnt:.0f}, tail {d[6] / cnt:.0f}, blocks {blocks}")
    print(f"n={n} MMA thread per block: wait_P {c[16] / max(1, c[18]) / nkb:.0f}, wait_KV {c[17] / max(1, c[18]) / nkb:.0f}, issue_PV {c[19] / max(1, c[18]) / nkb:.0f}, issue_S {c[20] / max(1, c[18]) / nkb:.0f}, commits {c[21] / max(1, c[18]) / nkb:.0f}")
