"""One profiled training step for ncu (`ncu --profile-from-start off ... python tools/profile_step.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from sdxl_training_improvements_b200.trainer import B200AdamW, _prep_batch, create_trainer
from sdxl_training_improvements_b200.unet import B200UNet

method = sys.argv[1] if len(sys.argv) > 1 else "ddpm"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
unet = B200UNet(device="cuda:0")
bench._init_weights_(unet, 1234)
from sdxl_training_improvements_b200.trainer import B200AdamWBF16
opt = B200AdamWBF16(unet, lr=4e-7, weight_decay=1e-2)
tr = create_trainer(bench._config_ns(method), unet, opt, device="cuda:0", seed=1)
batch = bench._synthetic_batch(4, 128, 128, 77)
for _ in range(warm):
    tr._execute_training_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr._execute_training_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
