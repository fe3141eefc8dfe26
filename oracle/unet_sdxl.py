"""ORACLE (test infrastructure, NOT product code): pure-PyTorch restatement of the SDXL-base UNet.

The reference calls ``self.model.unet(...)`` (src/training/trainers/methods/ddpm_trainer.py:320-325,
flow_matching_trainer.py:400-405, sdxl_trainer.py:65-70) where ``unet`` is diffusers'
``UNet2DConditionModel`` loaded from ``stabilityai/stable-diffusion-xl-base-1.0`` (src/models/sdxl.py:25-40).
diffusers is an un-vendored, un-pinned dependency (``diffusers>=0.21.0``, requirements.txt:2) and is not
installed here, so this file restates the published module graph with the diffusers
state-dict key names.  PARITY UNPINNED by reference tests (the reference has none); structural pins used
instead: 2,567,463,684 parameters and 1,680 state-dict tensors (tests/test_oracle_unet.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


SDXL_BASE = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280),
    layers_per_block=2,
    transformer_layers_per_block=(0, 2, 10),  # down block 0 has no attention
    num_heads=(5, 10, 20),  # diffusers "attention_head_dim" = number of heads; head dim 64
    cross_attention_dim=2048,
    addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=2816,
    norm_num_groups=32,
    norm_eps=1e-5,
)


def tiny_config(**over) -> dict:
    """A reduced-width config with the same topology (for CPU-sized parity tests)."""
    cfg = dict(SDXL_BASE)
    cfg.update(
        block_out_channels=(64, 128, 256),
        transformer_layers_per_block=(0, 1, 2),
        num_heads=(1, 2, 4),
        cross_attention_dim=128,
        addition_time_embed_dim=32,
        projection_class_embeddings_input_dim=6 * 32 + 96,  # pooled dim 96
    )
    cfg.update(over)
    return cfg


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)`` (Appendix A.1)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    freqs = torch.exp(exponent)
    args = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_ch: int, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int, kv_dim: Optional[int] = None):
        super().__init__()
        kv_dim = kv_dim or dim
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv_dim, dim, bias=False)
        self.to_v = nn.Linear(kv_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        B, n, C = x.shape
        h = self.heads
        q = self.to_q(x).view(B, n, h, C // h).transpose(1, 2)
        k = self.to_k(ctx).view(B, ctx.shape[1], h, C // h).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], h, C // h).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)  # scale = 1/sqrt(head_dim), no mask
        o = o.transpose(1, 2).reshape(B, n, C)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.proj = nn.Linear(cin, cout * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)  # exact (erf) GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, depth: int, ctx_dim: int, groups: int = 32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = self.proj_out(h)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return h + res


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, n_layers, depth, heads, ctx_dim, add_down, groups):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(n_layers)]
        )
        if depth > 0:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(cout, heads, depth, ctx_dim, groups) for _ in range(n_layers)]
            )
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, h, emb, ctx, skips):
        for i, res in enumerate(self.resnets):
            h = res(h, emb)
            if self.attentions is not None:
                h = self.attentions[i](h, ctx)
            skips.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            skips.append(h)
        return h


class MidBlock(nn.Module):
    def __init__(self, ch, temb, depth, heads, ctx_dim, groups):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, depth, ctx_dim, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups), ResnetBlock2D(ch, ch, temb, groups)])

    def forward(self, h, emb, ctx):
        h = self.resnets[0](h, emb)
        h = self.attentions[0](h, ctx)
        return self.resnets[1](h, emb)


class UpBlock(nn.Module):
    def __init__(self, res_in: Sequence[int], cout, temb, depth, heads, ctx_dim, add_up, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, cout, temb, groups) for c in res_in])
        if depth > 0:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(cout, heads, depth, ctx_dim, groups) for _ in res_in]
            )
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h, emb, ctx, skips):
        for i, res in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)
            h = res(h, emb)
            if self.attentions is not None:
                h = self.attentions[i](h, ctx)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class OracleUNet(nn.Module):
    """SDXL-base ``UNet2DConditionModel`` restated; state-dict keys per Appendix A.4."""

    def __init__(self, cfg: Optional[dict] = None):
        super().__init__()
        cfg = dict(SDXL_BASE if cfg is None else cfg)
        self.cfg = cfg
        boc = cfg["block_out_channels"]
        temb = boc[0] * 4
        G = cfg["norm_num_groups"]
        ctx = cfg["cross_attention_dim"]
        depth = cfg["transformer_layers_per_block"]
        heads = cfg["num_heads"]
        L = cfg["layers_per_block"]

        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.add_embedding = TimestepEmbedding(cfg["projection_class_embeddings_input_dim"], temb)

        self.down_blocks = nn.ModuleList()
        skip_ch = [boc[0]]
        cin = boc[0]
        for i, cout in enumerate(boc):
            last = i == len(boc) - 1
            self.down_blocks.append(DownBlock(cin, cout, temb, L, depth[i], heads[i], ctx, not last, G))
            skip_ch += [cout] * L + ([] if last else [cout])
            cin = cout

        self.mid_block = MidBlock(boc[-1], temb, depth[-1], heads[-1], ctx, G)

        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        rdepth = list(reversed(depth))
        rheads = list(reversed(heads))
        prev = boc[-1]
        for i, cout in enumerate(rev):
            res_in = []
            for j in range(L + 1):
                res_in.append((prev if j == 0 else cout) + skip_ch.pop())
            last = i == len(rev) - 1
            self.up_blocks.append(UpBlock(res_in, cout, temb, rdepth[i], rheads[i], ctx, not last, G))
            prev = cout

        self.conv_norm_out = nn.GroupNorm(G, boc[0], eps=cfg["norm_eps"])
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)

    def embed(self, sample_dtype, timestep, text_embeds, time_ids, B):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=text_embeds.device)
        if timestep.dim() == 0:
            timestep = timestep[None]
        timestep = timestep.expand(B)
        boc0 = self.cfg["block_out_channels"][0]
        t_emb = timestep_embedding(timestep, boc0).to(sample_dtype)
        emb = self.time_embedding(t_emb)
        tid = timestep_embedding(time_ids.flatten(), self.cfg["addition_time_embed_dim"]).reshape(B, -1)
        add = torch.cat([text_embeds.reshape(B, -1), tid.to(text_embeds.dtype)], dim=-1).to(emb.dtype)
        return emb + self.add_embedding(add)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs: Dict[str, torch.Tensor]):
        B = sample.shape[0]
        emb = self.embed(sample.dtype, timestep, added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"], B)
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            h = blk(h, emb, encoder_hidden_states, skips)
        h = self.mid_block(h, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            h = blk(h, emb, encoder_hidden_states, skips)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h)


def seeded_init_(model: nn.Module, seed: int = 0, out_gain: float = 1.0) -> nn.Module:
    """Deterministic synthetic weights (no SDXL checkpoint is available offline).

    Default torch init re-drawn from a seeded generator so that the oracle and the CUDA path can be given
    bit-identical weights through ``state_dict()``.
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.dim() >= 2:
                fan_in = p[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            elif name.endswith("weight"):  # norm scales
                p.copy_(1.0 + 0.1 * (torch.rand(p.shape, generator=g) * 2 - 1))
            else:
                p.copy_(0.05 * (torch.rand(p.shape, generator=g) * 2 - 1))
    return model
