"""AdamWBF16 kernel alone at the SDXL UNet's parameter count: ms per launch and GB/s against the box's own copy bandwidth.
18 B of algorithmic HBM traffic per parameter (read p, g, m, v, shift; write p, m, v, shift)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdxl_training_improvements_b200 import ops

bf16 = torch.bfloat16
n = int(os.environ.get("N", 2_567_463_680))
g = torch.Generator(device="cuda").manual_seed(1)
bufs = []
for scale in (0.05, 1e-3, 1e-3, 1e-6, 1e-5):
    t = torch.empty(n, device="cuda", dtype=bf16)
    t.normal_(0, scale, generator=g)
    bufs.append(t)
p, gr, m, v, sh = bufs
v.abs_()
so = torch.tensor([7, 1], device="cuda", dtype=torch.int64)
gn = torch.tensor([4.0], device="cuda", dtype=torch.float64)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = [e[i].elapsed_time(e[i + 1]) for i in range(reps)]
    return min(ts), sum(ts) / reps


for label, kw in (("no clip", dict()), ("clip active", dict(gnorm_sq=gn, max_norm=1.0)),
                  ("no clip, zero_grad folded in: 20 B/param", dict(zero_grad=True))):
    best, avg = timed(lambda: ops.adamw_bf16(p, gr, m, v, sh, lr=1e-5, step=3, seed_offset=so, **kw))
    print(f"adamw_bf16 n={n} ({label}): best {best:.3f} ms avg {avg:.3f} ms = {18 * n / best / 1e6:.0f} GB/s")
a = torch.empty(n, device="cuda", dtype=bf16)
best, avg = timed(lambda: gr.zero_())
print(f"torch zero_ of the gradient vector {2 * n / 1e9:.2f} GB: best {best:.3f} ms = {2 * n / best / 1e6:.0f} GB/s (write only)")
best, avg = timed(lambda: a.copy_(p))
print(f"torch copy {2 * n / 1e9:.2f} GB: best {best:.3f} ms = {4 * n / best / 1e6:.0f} GB/s (read + write)")
