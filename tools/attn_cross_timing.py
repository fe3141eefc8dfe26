"""Cross-attention (77 keys) kernel timings in CUDA-graph replay: fwd, bwd (prep + dQ + dK/dV)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdxl_training_improvements_b200 import ops

bf16 = torch.bfloat16


def graph_time(fn, iters=20):
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for (B, H, n) in ((4, 20, 1024), (4, 10, 4096)):
    Cc = H * 64
    q = torch.randn(B * n, Cc, device="cuda").to(bf16)
    kv = torch.randn(B * 77, 2 * Cc, device="cuda").to(bf16)
    k, v = kv[:, :Cc], kv[:, Cc:]
    do = torch.randn(B * n, Cc, device="cuda").to(bf16)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    o, lse = ops.attn_fwd(q, k, v, B, H, n, 77, 0.125)
    tf = graph_time(lambda: ops.attn_fwd(q, k, v, B, H, n, 77, 0.125, out=o))
    tb = graph_time(lambda: ops.attn_bwd(q, k, v, o, lse, do, dq, dkv[:, :Cc], dkv[:, Cc:], B, H, n, 77, 0.125))
    byts = 2 * q.numel() * 2 + 2 * kv.numel() * 2
    print(f"cross-attn B={B} H={H} n_q={n}: fwd {tf:.1f} us ({byts / tf / 1e3:.0f} GB/s of compulsory traffic), bwd {tb:.1f} us")
