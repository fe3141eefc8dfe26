"""Algorithmic FLOPs (2*MAC) of the SDXL UNet forward, derived from the module graph.
Training step = 3x forward (fwd + dgrad + wgrad, no recompute).  Used by bench.py for the roofline figures."""
from __future__ import annotations

from typing import Dict

from .params import SDXL_BASE


def unet_forward_flops(cfg: dict = None, H: int = 128, W: int = 128, n_ctx: int = 77) -> Dict[str, float]:
    cfg = dict(SDXL_BASE if cfg is None else cfg)
    boc = cfg["block_out_channels"]
    depth = cfg["transformer_layers_per_block"]
    L = cfg["layers_per_block"]
    ctx = cfg["cross_attention_dim"]
    temb = boc[0] * 4
    f = dict(conv3x3=0.0, conv1x1=0.0, ffn=0.0, self_qkvo=0.0, self_core=0.0, cross_qo=0.0, cross_kv=0.0,
             cross_core=0.0, proj_inout=0.0, emb=0.0)

    def conv(cin, cout, hw):
        f["conv3x3"] += 2.0 * hw * cout * 9 * cin

    def resnet(cin, cout, hw):
        conv(cin, cout, hw)
        conv(cout, cout, hw)
        f["emb"] += 2.0 * temb * cout
        if cin != cout:
            f["conv1x1"] += 2.0 * hw * cin * cout

    def transformer(c, d, n):
        f["proj_inout"] += 2 * 2.0 * n * c * c
        for _ in range(d):
            f["self_qkvo"] += 4 * 2.0 * n * c * c
            f["self_core"] += 2 * 2.0 * n * n * c
            f["cross_qo"] += 2 * 2.0 * n * c * c
            f["cross_kv"] += 2 * 2.0 * n_ctx * ctx * c
            f["cross_core"] += 2 * 2.0 * n * n_ctx * c
            f["ffn"] += 2.0 * n * c * 8 * c + 2.0 * n * 4 * c * c

    f["emb"] += 2.0 * (boc[0] * temb + temb * temb + cfg["projection_class_embeddings_input_dim"] * temb + temb * temb)
    h, w = H, W
    conv(cfg["in_channels"], boc[0], h * w)
    skip = [boc[0]]
    cin = boc[0]
    for i, cout in enumerate(boc):
        last = i == len(boc) - 1
        for j in range(L):
            resnet(cin if j == 0 else cout, cout, h * w)
            if depth[i]:
                transformer(cout, depth[i], h * w)
            skip.append(cout)
        if not last:
            h, w = (h + 1) // 2, (w + 1) // 2
            conv(cout, cout, h * w)
            skip.append(cout)
        cin = cout
    resnet(boc[-1], boc[-1], h * w)
    transformer(boc[-1], depth[-1], h * w)
    resnet(boc[-1], boc[-1], h * w)
    rev, rdepth = list(reversed(boc)), list(reversed(depth))
    prev = boc[-1]
    for i, cout in enumerate(rev):
        last = i == len(rev) - 1
        for j in range(L + 1):
            resnet((prev if j == 0 else cout) + skip.pop(), cout, h * w)
            if rdepth[i]:
                transformer(cout, rdepth[i], h * w)
        if not last:
            h, w = h * 2, w * 2
            conv(cout, cout, h * w)
        prev = cout
    conv(boc[0], cfg["out_channels"], h * w)
    f["total"] = sum(f.values())
    return f


def train_step_flops(cfg: dict = None, H: int = 128, W: int = 128) -> float:
    return 3.0 * unet_forward_flops(cfg, H, W)["total"]
