"""ff1 (GEGLU up-projection) at M=4096, F=5120, K=1280: gate fused into the GEMM epilogue vs GEMM + gate kernel (graph replay)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sdxl_training_improvements_b200 import ops

bf = torch.bfloat16
for (M, F, K) in ((4096, 5120, 1280), (16384, 2560, 640)):
    x = torch.randn(M, K, device="cuda").to(bf)
    W1 = (torch.randn(2 * F, K, device="cuda") * 0.03).to(bf)
    b1 = torch.zeros(2 * F, device="cuda", dtype=bf)
    u = torch.empty(M, 2 * F, device="cuda", dtype=bf)
    tf = bench._graph_time_us(lambda: ops.linear_geglu_fwd(x, W1, b1, F)) if ops.linear_geglu_ok(M, F, K) else float("nan")
    tg = bench._graph_time_us(lambda: ops.linear_fwd(x, W1, bias=b1, out=u))
    tk = bench._graph_time_us(lambda: ops.geglu_fwd(u, F))
    fl = 2.0 * M * 2 * F * K
    print(f"M={M} F={F} K={K}: fused {tf:.1f} us ({fl / tf / 1e6:.0f} TFLOP/s) | gemm {tg:.1f} us ({fl / tg / 1e6:.0f}) + gate kernel {tk:.1f} us = {tg + tk:.1f} us")

# down-projection dgrad with the GEGLU backward in its epilogue vs dgrad GEMM + GEGLU-backward kernel
for (M, F, Cc) in ((4096, 5120, 1280), (16384, 2560, 640)):
    dy = torch.randn(M, Cc, device="cuda").to(bf)
    W2 = (torch.randn(Cc, F, device="cuda") * 0.02).to(bf)
    u = torch.randn(M, 2 * F, device="cuda").to(bf)
    dz = torch.empty(M, F, device="cuda", dtype=bf)
    db = torch.zeros(2 * F, device="cuda")
    tf = bench._graph_time_us(lambda: ops.linear_dgrad_geglu(dy, W2, u, F)) if ops.linear_dgrad_geglu_ok(M, F, Cc) else float("nan")
    tg = bench._graph_time_us(lambda: ops.linear_dgrad(dy, W2, dz))
    tk = bench._graph_time_us(lambda: ops.geglu_bwd(u, dz, F, dbias32=db))
    tc = bench._graph_time_us(lambda: ops.colsum_f32(u, db))
    print(f"M={M} F={F} C={Cc}: fused dgrad+geglu_bwd {tf:.1f} us (+ colsum {tc:.1f} on the side stream) | dgrad {tg:.1f} us + geglu_bwd_bias kernel {tk:.1f} us = {tg + tk:.1f} us")
