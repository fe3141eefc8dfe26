"""Raw copy-engine bandwidth between IPC-mapped peer buffers, PULL (remote -> local) vs PUSH (local -> remote), through
b2_dpx_memcpy_async — the primitive behind the finding that decided the exchange design (csrc/dpx.cu): at N = 4 the pull-based
reduce-scatter reached 370 GB/s of NVLink ingress per GPU, the push-based one 698 GB/s (profiles/r1_dpx_transports_n4.txt).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/dp_copy_probe.py
Every rank copies `MB` megabytes from / to each peer at the same time (all-to-all pattern, one stream per peer), CUDA events,
max over ranks.  Prints GB/s per GPU and direction."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sdxl_training_improvements_b200 import _lib
    lib = _lib.load()
    mb = int(os.environ.get("MB", "512"))
    n = mb * 1024 * 1024
    src = torch.empty(n, device="cuda", dtype=torch.uint8)
    dst = torch.empty((world - 1) * n, device="cuda", dtype=torch.uint8)

    def export(t):
        h, off = (C.c_ubyte * 64)(), C.c_int64()
        _lib.check(lib.b2_dpx_ipc_export(C.c_void_p(t.data_ptr()), h, C.byref(off)), "ipc_export")
        return bytes(h), int(off.value)

    everyone = [None] * world
    dist.all_gather_object(everyone, (export(src), export(dst)))

    def imp(hb, off):
        p = C.c_void_p()
        _lib.check(lib.b2_dpx_ipc_import((C.c_ubyte * 64).from_buffer_copy(hb), off, C.byref(p)), "ipc_import")
        return p.value

    peers = [(rank + 1 + j) % world for j in range(world - 1)]
    psrc = {p: imp(*everyone[p][0]) for p in peers}
    pdst = {p: imp(*everyone[p][1]) for p in peers}
    streams = [torch.cuda.Stream() for _ in peers]

    def run(kind, iters=5):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for _ in range(iters):
            for j, p in enumerate(peers):
                if kind == "pull":   # read peer p's src into my slot j
                    a, b = dst.data_ptr() + j * n, psrc[p]
                else:                # write my src into peer p's slot (my index among ITS peers)
                    a, b = pdst[p] + ((rank - p - 1) % world) * n, src.data_ptr()
                _lib.check(lib.b2_dpx_memcpy_async(C.c_void_p(a), C.c_void_p(b), n, C.c_void_p(streams[j].cuda_stream)),
                           "memcpy")
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for kind in ("pull", "push", "pull", "push"):
        ms = run(kind)
        if rank == 0:
            print(f"{kind}: {world - 1} x {mb} MB per GPU in {ms:.2f} ms = {(world - 1) * n / ms / 1e6:.0f} GB/s per GPU and direction "
                  f"(N={world})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
