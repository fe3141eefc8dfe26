"""clock64 trace of the GEGLU-mode GEMM (cluster 0): epilogue thread 0 and the MMA thread, first 8 tiles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdxl_training_improvements_b200 import _lib, ops

bf = torch.bfloat16
M, F, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 5120, 1280)
x = torch.randn(M, K, device="cuda").to(bf)
W1 = (torch.randn(2 * F, K, device="cuda") * 0.03).to(bf)
b1 = torch.zeros(2 * F, device="cuda", dtype=bf)
for _ in range(3):
    ops.linear_geglu_fwd(x, W1, b1, F)
torch.cuda.synchronize()
buf = torch.zeros(256, device="cuda", dtype=torch.int64)
_lib.load().b2_gemm2_set_debug(buf.data_ptr())
ops.linear_geglu_fwd(x, W1, b1, F)
torch.cuda.synchronize()
_lib.load().b2_gemm2_set_debug(None)
t = buf.tolist()
t0 = min(v for v in t if v > 0)
print(f"M={M} F={F} K={K}  (cycles relative to the first stamp; mainloop of one 256x256 tile = {K // 64 * 512} tensor cycles)")
print("epilogue thread 0: tile | wait acc start, acc ready | c0: smem free + tmem ld, math done, stores issued | c1: ... ")
for i in range(8):
    r = t[i * 8:i * 8 + 8]
    if r[0] == 0:
        continue
    print(f"  tile {i}: " + " ".join(f"{v - t0:7d}" for v in r) + f"   | acc wait {r[1] - r[0]:6d}  c0: wait {r[2] - r[1]:5d} math {r[3] - r[2]:5d} store {r[4] - r[3]:5d}"
          f"  c1: wait {r[5] - r[4]:5d} math {r[6] - r[5]:5d} store {r[7] - r[6]:5d}  total {r[7] - r[1]:6d}")
print("MMA thread: tile | before acc_empty wait, after, last commit issued")
for i in range(8):
    r = t[128 + i * 4:128 + i * 4 + 3]
    if r[0] == 0:
        continue
    print(f"  tile {i}: " + " ".join(f"{v - t0:7d}" for v in r) + f"   | wait {r[1] - r[0]:6d} issue {r[2] - r[1]:6d}")

# plain mode (no gate): is the epilogue (time between "acc ready" of tile i and "wait acc start" of tile i+1) shorter than the mainloop?
import bench
for (M2, N2, K2, res) in ((4096, 10240, 1280, False), (4096, 3840, 1280, False), (16384, 5120, 640, False), (16384, 1920, 640, False), (16384, 640, 640, True), (16384, 640, 2560, True)):
    x2 = torch.randn(M2, K2, device="cuda").to(bf)
    W2 = (torch.randn(N2, K2, device="cuda") * 0.03).to(bf)
    bb = torch.zeros(N2, device="cuda", dtype=bf)
    r2 = torch.randn(M2, N2, device="cuda").to(bf) if res else None
    o2 = torch.empty(M2, N2, device="cuda", dtype=bf)
    for _ in range(3):
        ops.linear_fwd(x2, W2, bias=bb, residual=r2, out=o2)
    torch.cuda.synchronize()
    buf.zero_()
    _lib.load().b2_gemm2_set_debug(buf.data_ptr())
    ops.linear_fwd(x2, W2, bias=bb, residual=r2, out=o2)
    torch.cuda.synchronize()
    _lib.load().b2_gemm2_set_debug(None)
    t = buf.tolist()
    us = bench._graph_time_us(lambda: ops.linear_fwd(x2, W2, bias=bb, residual=r2, out=o2))
    print(f"plain M={M2} N={N2} K={K2} residual={res}: {us:.1f} us = {2.0 * M2 * N2 * K2 / us / 1e6:.0f} TFLOP/s; mainloop {K2 // 64 * 512} tensor cycles per 256x256 tile")
    for i in range(2, 5):
        e, e1, m = t[i * 8:i * 8 + 2], t[(i + 1) * 8:(i + 1) * 8 + 2], t[128 + i * 4:128 + i * 4 + 3]
        if e[0] == 0 or e1[0] == 0:
            continue
        r = t[i * 8:i * 8 + 8]
        print(f"  tile {i}: chunk 0: sync+ld {r[2] - r[1]:5d} math {r[3] - r[2]:5d} fence+bar+store {r[4] - r[3]:5d} | chunk 1: sync+ld {r[5] - r[4]:5d} math {r[6] - r[5]:5d} fence+bar+store {r[7] - r[6]:5d}")
        print(f"  tile {i}: epilogue busy {e1[0] - e[1]:6d} (then waits {e1[1] - e1[0]:6d} for the next accumulator) | MMA thread: waits {m[1] - m[0]:6d} for a free accumulator, issues for {m[2] - m[1]:6d}")


# single-round GEMMs (one tile per CTA pair): where does the launch's time go?  Graph-replay time per launch next to the
# cycle stamps of cluster 0: entry -> set-up done -> first MMA issue possible -> last MMA issued -> accumulator ready ->
# last store issued -> stores complete.
for (M2, N2, K2, res, bmn) in ((4096, 1280, 1280, True, False), (4096, 1280, 1280, False, True), (4096, 1280, 5120, True, False)):
    x2 = torch.randn(M2, K2, device="cuda").to(bf)
    W2 = (torch.randn(N2, K2, device="cuda") * 0.03).to(bf)
    bb = torch.zeros(N2, device="cuda", dtype=bf)
    r2 = torch.randn(M2, N2, device="cuda").to(bf) if res else None
    o2 = torch.empty(M2, N2, device="cuda", dtype=bf)
    dy = torch.randn(M2, N2, device="cuda").to(bf)
    dx = torch.empty(M2, K2, device="cuda", dtype=bf)
    fn = (lambda: ops.linear_dgrad(dy, W2, dx)) if bmn else (lambda: ops.linear_fwd(x2, W2, bias=bb, residual=r2, out=o2))
    us = bench._graph_time_us(fn)
    buf.zero_()
    _lib.load().b2_gemm2_set_debug(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    _lib.load().b2_gemm2_set_debug(None)
    t = buf.tolist()
    e0 = t[187]
    m = t[128:131]
    ep = t[0:8]
    print(f"{'dgrad' if bmn else 'fwd'} M={M2} N={N2} K={K2} residual={res}: {us:.1f} us per launch in graph replay ({us * 1.85e3:.0f} cycles at 1.85 GHz); "
          f"mainloop {K2 // 64 * (640 if N2 % 320 == 0 and not bmn else 512)} tensor cycles")
    print(f"   cycles from kernel entry (cluster 0): set-up done {t[191] - e0}, dependency wait done {t[190] - e0}, MMA thread has accumulator {m[1] - e0}, "
          f"last MMA issued {m[2] - e0}, epilogue sees accumulator {ep[1] - e0}, last store issued {t[189] - e0}, stores complete {t[188] - e0}")


# GEGLU-backward mode (dgrad of the down-projection + gate backward in the epilogue): group 0's two chunks of each tile
M3, F3, C3 = 4096, 5120, 1280
dy3 = torch.randn(M3, C3, device="cuda").to(bf)
W3 = (torch.randn(C3, F3, device="cuda") * 0.02).to(bf)
u3 = torch.randn(M3, 2 * F3, device="cuda").to(bf)
for _ in range(3):
    ops.linear_dgrad_geglu(dy3, W3, u3, F3)
torch.cuda.synchronize()
buf.zero_()
_lib.load().b2_gemm2_set_debug(buf.data_ptr())
ops.linear_dgrad_geglu(dy3, W3, u3, F3)
torch.cuda.synchronize()
_lib.load().b2_gemm2_set_debug(None)
t = buf.tolist()
print("dgrad + GEGLU backward epilogue, group 0 (chunks 0 and 2 of each 256-column tile):")
for i in range(1, 5):
    r, r1, m = t[i * 8:i * 8 + 8], t[(i + 1) * 8:(i + 1) * 8 + 2], t[128 + i * 4:128 + i * 4 + 3]
    print(f"  tile {i}: acc wait {r[1] - r[0]:6d} | chunk 0: drain + issue loads {r[2] - r[1]:5d}, u tiles arrive {r[3] - r[2]:5d}, math {r[4] - r[3]:5d} | "
          f"chunk 2: fence+bar+store+drain+issue {r[5] - r[4]:5d}, arrive {r[6] - r[5]:5d}, math {r[7] - r[6]:5d} | to next tile {r1[0] - r[7]:5d} | MMA waits {m[1] - m[0]:6d} issues {m[2] - m[1]:6d}")
