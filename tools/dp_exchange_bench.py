"""Times the peer-memory gradient exchange (csrc/dpx.cu) against one NCCL all-reduce on the full-size flat bf16 gradient
buffer (2.57 G elements = 5.13 GB), stand-alone (nothing else on the GPUs).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_exchange_bench.py
Prints per-step ms (max over ranks), algorithm bandwidth (buffer bytes / time) and per-GPU NVLink ingress bandwidth
(2 (N-1)/N buffer bytes / time) for: exchange_all (one chunk), the same buffer as 10 chunks, NCCL all_reduce."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sdxl_training_improvements_b200 import dp as D
    total = int(os.environ.get("DPX_ELEMS", 2_567_464_000)) // 8 * 8
    grad = torch.zeros(total, device="cuda", dtype=torch.bfloat16)
    x = D.PeerGradExchange(grad, self_test=True, mode="auto")
    nchunk = 10
    step = total // nchunk // 8 * 8
    base = 0
    for k in range(nchunk):
        base = x._set_chunk(k, [(k * step, step if k < nchunk - 1 else total - k * step)], base)

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def whole():
        x.exchange_all()
        x.finish()

    def chunks():
        for k in range(nchunk):
            x.exchange_chunk(k)
        x.finish()

    def nccl():
        dist.all_reduce(grad)

    res = {}
    chosen = x.mode
    for m in x.handles:
        x.mode = m
        res[f"{m}/whole"] = timed(whole)
        res[f"{m}/chunks10"] = timed(chunks)
    x.mode = chosen
    res["nccl"] = timed(nccl)
    if rank == 0:
        gb = total * 2 / 1e9
        for k, ms in res.items():
            print(f"{k:17s} {ms:8.2f} ms  alg {gb / ms * 1e3:7.1f} GB/s  ingress/GPU {2 * (world - 1) / world * gb / ms * 1e3:7.1f} GB/s"
                  f"  (N={world}, {gb:.2f} GB)", flush=True)
        print("autotune picked", chosen, x.timings, flush=True)
    x.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
