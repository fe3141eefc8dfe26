"""clock64 event trace of the first tile of CTA 0 of xattn_q_core_kernel: where the attention phase spends its time."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdxl_training_improvements_b200 import _lib, ops

bf16 = torch.bfloat16
B, n, Cc, nk = 4, 1024, 1280, 77
xn = torch.randn(B * n, Cc, device="cuda").to(bf16)
Wq = (torch.randn(Cc, Cc, device="cuda") * 0.02).to(bf16)
kv = torch.randn(B * nk, 2 * Cc, device="cuda").to(bf16)
k, v = kv[:, :Cc], kv[:, Cc:]
for _ in range(3):
    ops.xattn_q_core(xn, Wq, k, v, B, n, nk, 0.125)
torch.cuda.synchronize()
buf = torch.zeros(64, device="cuda", dtype=torch.int64)
_lib.load().b2_xattn_set_debug(buf.data_ptr())
ops.xattn_q_core(xn, Wq, k, v, B, n, nk, 0.125)
torch.cuda.synchronize()
_lib.load().b2_xattn_set_debug(None)
c = buf.tolist()
t0 = c[0]
print(f"acc ready at {c[1] - t0} (cycles after the epilogue warps started waiting = mainloop incl. fill)")
print(f"B1 (5 x Q_h -> smem + store) {c[2] - c[1]} cycles; attention phase B2 {c[3] - c[2]}; tile total {c[3] - t0}")
for h in range(5):
    s0, s1, o0, o1 = (c[8 + 4 * h + i] - c[2] for i in range(4))
    print(f"head {h}: wait S from {s0} to {s1} ({s1 - s0}); softmax+P until O-wait starts {o0} ; O ready {o1} (waited {o1 - o0})")
