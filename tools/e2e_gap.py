"""Where the end-to-end step (plugin API, host buffers) loses time against the device-resident loop: GPU idle time between the
optimizer graph of step i and the micro-step graph of step i+1, and the host time of each phase of _execute_training_step."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sdxl_training_improvements_b200 import trainer as T
from sdxl_training_improvements_b200.unet import B200UNet

unet = B200UNet(device="cuda:0")
bench._init_weights_(unet, 1234)
opt = T.B200AdamWBF16(unet, lr=4e-7, weight_decay=1e-2)
tr = T.create_trainer(bench._config_ns("ddpm"), unet, opt, device="cuda:0", seed=1, cuda_graph=True)
batch = bench._synthetic_batch(4, 128, 128, 77, pin=True)
for _ in range(4):
    tr._execute_training_step(batch)
torch.cuda.synchronize()
gm = next(iter(tr._micro_graphs.values()))
og = tr._opt_graph
ev = {"m0": [], "m1": [], "o0": [], "o1": []}
host = {"micro_launch": [], "opt_launch": [], "sync": []}
g_replay, o_replay = gm.replay, og.replay


def mrep(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t = time.perf_counter(); r = g_replay(*a, **k); host["micro_launch"].append(time.perf_counter() - t); e1.record()
    ev["m0"].append(e0); ev["m1"].append(e1)
    return r


def orep(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t = time.perf_counter(); r = o_replay(*a, **k); host["opt_launch"].append(time.perf_counter() - t); e1.record()
    ev["o0"].append(e0); ev["o1"].append(e1)
    return r


gm.replay, og.replay = mrep, orep
hs = tr._host_event.synchronize


def hsync():
    t = time.perf_counter(); hs(); host["sync"].append(time.perf_counter() - t)


tr._host_event.synchronize = hsync
K = 10
t_all = []
for i in range(K):
    t = time.perf_counter()
    tr._execute_training_step(batch)
    t_all.append(time.perf_counter() - t)
torch.cuda.synchronize()
ms = lambda a, b: a.elapsed_time(b)
micro = [ms(ev["m0"][i], ev["m1"][i]) for i in range(K)]
optm = [ms(ev["o0"][i], ev["o1"][i]) for i in range(K)]
between = [ms(ev["m1"][i], ev["o0"][i]) for i in range(K)]
gap = [ms(ev["o1"][i], ev["m0"][i + 1]) for i in range(K - 1)]
step = [ms(ev["m0"][i], ev["m0"][i + 1]) for i in range(K - 1)]
avg = lambda v: sum(v) / len(v)
print(f"GPU timeline per step (ms): micro-step graph {avg(micro):.2f}, micro->optimizer {avg(between):.3f}, optimizer graph {avg(optm):.2f}, "
      f"optimizer end -> next micro-step start {avg(gap):.3f}; step period {avg(step):.2f}")
print(f"host per step (ms): whole call {1e3 * avg(t_all):.2f}, micro graph launch {1e3 * avg(host['micro_launch']):.3f}, "
      f"optimizer graph launch {1e3 * avg(host['opt_launch']):.3f}, blocked in the loss read {1e3 * avg(host['sync']):.2f}")
