"""Full-size (SDXL-base) parity probe: kernel UNet vs the oracle in CUDA eager, same weights/inputs.

Prints per-stage statistics so that a divergence can be localised; not a pytest (needs ~30 GB HBM and ~1 min).
  python tools/fullsize_check.py [B] [H] [W]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from oracle import schedule as S
from oracle.unet_sdxl import OracleUNet
from sdxl_training_improvements_b200.unet import B200UNet

bf16 = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W = int(sys.argv[3]) if len(sys.argv) > 3 else 64


def rel(a, b):
    a = a.float().flatten(); b = b.float().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


net = B200UNet(device="cuda:0")
bench._init_weights_(net, 1234)
with torch.device("meta"):
    ref = OracleUNet()
ref = ref.to_empty(device="cuda").to(bf16)
ref.load_state_dict({k: v.to(bf16) for k, v in net.state_dict().items()})
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn(B, 4, H, W, device="cuda", generator=g).to(bf16)
ctx = torch.randn(B, 77, 2048, device="cuda", generator=g).to(bf16)
pooled = torch.randn(B, 1280, device="cuda", generator=g).to(bf16)
tid = torch.tensor([[8. * W, 8. * H, 0., 0., 8. * W, 8. * H]], device="cuda").repeat(B, 1)[:, None]
sig = S.schedule_sigmas()
for tval in (900, 500, 100, 5):
    t = torch.full((B,), tval, device="cuda", dtype=torch.long)
    noise = torch.randn(B, 4, H, W, device="cuda", generator=g).to(bf16)
    noisy = S.add_noise(x.float(), noise.float(), sig.cuda()[t]).to(bf16)
    with torch.no_grad():
        ro = ref(noisy, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
        ko = net(noisy, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
    tgt = S.get_velocity(x.float(), noise.float(), sig.cuda()[t])
    print(f"t={tval} sigma={float(sig[tval]):.4g} |noisy|max={float(noisy.float().abs().max()):.4g} "
          f"oracle: absmax {float(ro.float().abs().max()):.4g} finite {bool(torch.isfinite(ro.float()).all())} "
          f"loss {float(((ro.float() - tgt) ** 2).mean()):.5g} | kernel: absmax {float(ko.float().abs().max()):.4g} "
          f"finite {bool(torch.isfinite(ko.float()).all())} loss {float(((ko.float() - tgt) ** 2).mean()):.5g} "
          f"| rel-L2 {rel(ko, ro):.4g}", flush=True)

# ---- backward at full size: parameter gradients of the kernel path vs the oracle (both bf16, same weights / inputs)
t = torch.full((B,), 300, device="cuda", dtype=torch.long)
noise = torch.randn(B, 4, H, W, device="cuda", generator=g).to(bf16)
noisy = S.add_noise(x.float(), noise.float(), sig.cuda()[t]).to(bf16)
wgt = torch.randn(B, 4, H, W, device="cuda", generator=g)
net.zero_grad()
ko = net(noisy, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
(ko.float() * wgt).sum().backward()
ro = ref(noisy, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
(ro.float() * wgt).sum().backward()
rp = dict(ref.named_parameters())
names = ["conv_in.weight", "down_blocks.0.resnets.0.conv1.weight", "down_blocks.0.resnets.1.conv2.bias",
         "down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.weight",
         "down_blocks.1.attentions.1.transformer_blocks.1.ff.net.0.proj.weight",
         "down_blocks.2.attentions.0.transformer_blocks.3.attn2.to_k.weight",
         "down_blocks.2.attentions.1.transformer_blocks.9.norm2.weight",
         "mid_block.attentions.0.transformer_blocks.5.attn2.to_out.0.weight", "mid_block.resnets.1.conv2.weight",
         "up_blocks.0.resnets.0.conv1.weight", "up_blocks.0.attentions.2.transformer_blocks.9.ff.net.2.weight",
         "up_blocks.1.resnets.2.conv_shortcut.weight", "up_blocks.1.upsamplers.0.conv.weight",
         "up_blocks.2.resnets.2.conv2.weight", "up_blocks.2.resnets.0.norm1.weight", "conv_norm_out.bias", "conv_out.weight",
         "time_embedding.linear_2.weight", "add_embedding.linear_1.weight"]
num = den = 0.0
worst = ("", 0.0)
for k, p in net.named_parameters():
    gk, go = p.grad.float(), rp[k].grad.float()
    num += float((gk - go).norm() ** 2)
    den += float(go.norm() ** 2)
    r = rel(gk, go)
    if r > worst[1] and float(go.norm()) > 1e-6 * (den ** 0.5 + 1e-30):
        worst = (k, r)
    if k in names:
        print(f"grad {k:72s} rel-L2 {r:.4g}  |g| {float(go.norm()):.4g}", flush=True)
print(f"backward: aggregate rel-L2 over all {sum(1 for _ in net.parameters())} tensors {((num / den) ** 0.5):.4g}; worst {worst}")
print(f"forward rel-L2 {rel(ko, ro):.4g}")
