"""GPU diagnostics: (1) GPU-only time of small GEMMs (CUDA-graph replay removes host launch cost) vs the eager loop,
(2) implicit conv vs im2col+GEMM on the SDXL conv shapes.  Prints one line per case."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdxl_training_improvements_b200 import ops

bf16 = torch.bfloat16


def timeit(fn, iters=50, graph=False):
    fn(); fn()
    torch.cuda.synchronize()
    if graph:
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
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def gemm_case(name, M, N, K, mode):
    x = torch.randn(M, K, device="cuda").to(bf16)
    W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf16)
    if mode == "fwd":
        out = torch.empty(M, N, device="cuda", dtype=bf16)
        fn = lambda: ops.linear_fwd(x, W, out=out)
        fl = 2.0 * M * N * K
    elif mode == "dgrad":
        dy = torch.randn(M, N, device="cuda").to(bf16)
        dx = torch.empty(M, K, device="cuda", dtype=bf16)
        fn = lambda: ops.linear_dgrad(dy, W, dx)
        fl = 2.0 * M * N * K
    else:
        dy = torch.randn(M, N, device="cuda").to(bf16)
        dW = torch.zeros(N, K, device="cuda", dtype=bf16)
        fn = lambda: ops.linear_wgrad(dy, x, dW, accumulate=True)
        fl = 2.0 * M * N * K
    te = timeit(fn)
    tg = timeit(fn, graph=True)
    print(f"GEMM {name:14s} {mode:5s} M={M:6d} N={N:6d} K={K:6d}: eager {te:7.1f} us  graph {tg:7.1f} us "
          f"({fl / tg / 1e6:6.0f} TF/s)", flush=True)


def conv_case(B, H, W, Cin, Cout):
    M = B * H * W
    x = torch.randn(M, Cin, device="cuda").to(bf16)
    wk = (torch.randn(Cout, 9 * Cin, device="cuda") / math.sqrt(9 * Cin)).to(bf16)
    dy = torch.randn(M, Cout, device="cuda").to(bf16)
    y = torch.empty(M, Cout, device="cuda", dtype=bf16)
    dx = torch.empty(M, Cin, device="cuda", dtype=bf16)
    dw = torch.zeros_like(wk)
    col = torch.empty(M, 9 * Cin, device="cuda", dtype=bf16)
    fl = 2.0 * M * Cout * 9 * Cin
    bias = torch.zeros(Cout, device="cuda", dtype=bf16)
    r = {}
    r["fwd_impl"] = timeit(lambda: ops.conv3x3_fwd(x, wk, B, H, W, Cin, Cout, bias=bias, out=y), 20, True)
    r["fwd_im2col"] = timeit(lambda: (ops.im2col3x3(x, B, H, W, Cin, out=col), ops.linear_fwd(col, wk, bias=bias, out=y)), 20, True)
    r["dgrad_impl"] = timeit(lambda: ops.conv3x3_dgrad(dy, wk, dx, B, H, W, Cin, Cout), 20, True)
    r["dgrad_im2col"] = timeit(lambda: (ops.linear_dgrad(dy, wk, col), ops.col2im3x3(col, dx, B, H, W, Cin)), 20, True)
    r["wgrad_impl"] = timeit(lambda: ops.conv3x3_wgrad(dy, x, dw, B, H, W, Cin, Cout), 20, True)
    r["wgrad_im2col"] = timeit(lambda: (ops.im2col3x3(x, B, H, W, Cin, out=col),
                                        ops.gemm_raw(dy, col, dw, Cout, 9 * Cin, M, a_mn=True, b_mn=True, lda=Cout,
                                                     ldb=9 * Cin, ldd=9 * Cin, accumulate=True)), 20, True)
    print(f"CONV B={B} {H}x{W} {Cin}->{Cout} ({fl / 1e9:.1f} GF): " +
          "  ".join(f"{k} {v:7.1f}us ({fl / v / 1e6:5.0f}TF/s)" for k, v in r.items()), flush=True)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    for name, M, N, K in [("attn proj", 4096, 1280, 1280), ("qkv", 4096, 3840, 1280), ("ff1", 4096, 10240, 1280),
                          ("ff2", 4096, 1280, 5120), ("proj@640", 16384, 640, 640), ("ff1@640", 16384, 5120, 640)]:
        for mode in ("fwd", "dgrad", "wgrad"):
            gemm_case(name, M, N, K, mode)
    for c in [(4, 128, 128, 320, 320), (4, 64, 64, 640, 640), (4, 32, 32, 1280, 1280), (4, 32, 32, 2560, 1280),
              (4, 128, 128, 960, 320), (4, 64, 64, 1920, 640)]:
        conv_case(*c)
