"""Times b2_gemm on the SDXL step's dominant shapes (CUDA events, operands rotated through > L2 worth of buffers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdxl_training_improvements_b200 import ops

bf16 = torch.bfloat16
SHAPES = [
    # name, M, N, K, a_mn, b_mn
    ("ff1 fwd      ", 4096, 10240, 1280, 0, 0),
    ("ff2 fwd      ", 4096, 1280, 5120, 0, 0),
    ("attn proj fwd", 4096, 1280, 1280, 0, 0),
    ("qkv fwd      ", 4096, 3840, 1280, 0, 0),
    ("ff1 dgrad    ", 4096, 1280, 10240, 0, 1),
    ("ff1 wgrad    ", 10240, 1280, 4096, 1, 1),
    ("proj wgrad   ", 1280, 1280, 4096, 1, 1),
    ("ff1@640 fwd  ", 16384, 5120, 640, 0, 0),
    ("proj@640 fwd ", 16384, 640, 640, 0, 0),
    ("conv1280 fwd ", 4096, 1280, 11520, 0, 0),
    ("conv640 fwd  ", 16384, 640, 5760, 0, 0),
    ("conv320 fwd  ", 65536, 320, 2880, 0, 0),
    ("conv320 wgrad", 320, 2880, 65536, 1, 1),
    ("conv1280 wgr ", 1280, 11520, 4096, 1, 1),
]
# configs: "legacy" (single-CTA 128x128 kernel), "auto" (CTA-pair kernel, host-picked BN), "bnNNN" (forced BN)
tiles = sys.argv[1].split(",") if len(sys.argv) > 1 else ["auto"]


def run(M, N, K, a_mn, b_mn, cfg):
    os.environ.pop("B2_GEMM_LEGACY", None)
    os.environ.pop("B2_GEMM_BN", None)
    if cfg == "legacy":
        os.environ["B2_GEMM_LEGACY"] = "1"
    elif cfg.startswith("bn"):
        os.environ["B2_GEMM_BN"] = cfg[2:]
    tile = 0
    nset = max(2, int(300e6 / (2 * (M * K + N * K + M * N))) + 1)
    nset = min(nset, 8)
    As = [torch.randn((K, M) if a_mn else (M, K), device="cuda").to(bf16) for _ in range(nset)]
    Bs = [(torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.03).to(bf16) for _ in range(nset)]
    Ds = [torch.empty(M, N, device="cuda", dtype=bf16) for _ in range(nset)]

    def go(i):
        ops.gemm_raw(As[i], Bs[i], Ds[i], M, N, K, a_mn=a_mn, b_mn=b_mn, lda=As[i].stride(0), ldb=Bs[i].stride(0), ldd=N,
                     tile_n=tile)

    for i in range(nset):
        go(i)
    torch.cuda.synchronize()
    it = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(it):
        go(i % nset)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    # correctness spot check against torch
    A = As[0].float().t() if a_mn else As[0].float()
    Bm = Bs[0].float() if b_mn else Bs[0].float().t()
    go(0)
    ref = A[:256] @ Bm
    err = float((Ds[0][:256].float() - ref).abs().max() / (ref.abs().max() + 1e-9))
    return ms, err


for name, M, N, K, a_mn, b_mn in SHAPES:
    line = f"{name} M={M:6d} N={N:6d} K={K:6d}:"
    for t in tiles:
        try:
            ms, err = run(M, N, K, a_mn, b_mn, t)
            line += f"  {t}: {ms * 1e3:7.1f} us {2.0 * M * N * K / ms / 1e9:6.0f} TF/s (err {err:.1e})"
        except Exception as ex:  # noqa: BLE001
            line += f"  {t}: FAILED {str(ex)[:60]}"
    print(line, flush=True)
