"""A few launches of the implicit-GEMM conv (fwd, dgrad, wgrad) for `ncu --set full` (python tools/prof_conv.py B H W Cin Cout)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdxl_training_improvements_b200 import ops

B, H, W, Cin, Cout = (int(x) for x in sys.argv[1:6])
bf16 = torch.bfloat16
M = B * H * W
x = torch.randn(M, Cin, device="cuda").to(bf16)
wk = (torch.randn(Cout, 9 * Cin, device="cuda") / math.sqrt(9 * Cin)).to(bf16)
dy = torch.randn(M, Cout, device="cuda").to(bf16)
dx = torch.empty_like(x)
dw = torch.zeros_like(wk)
bias = torch.zeros(Cout, device="cuda", dtype=bf16)
for _ in range(3):
    ops.conv3x3_fwd(x, wk, B, H, W, Cin, Cout, bias=bias)
    ops.conv3x3_dgrad(dy, wk, dx, B, H, W, Cin, Cout)
    ops.conv3x3_wgrad(dy, x, dw, B, H, W, Cin, Cout)
torch.cuda.synchronize()
print("done")
