"""Parameter inventory of the SDXL-base UNet in diffusers' state-dict naming, and the
flat bf16 parameter / gradient buffers the kernels and the single NCCL all-reduce operate on.

The replacement UNet must round-trip `unet.state_dict()` / `save_pretrained` (src/models/sdxl.py:86-94,246-288) and
expose `parameters()` whose `.grad` the reference optimizers read (src/training/optimizers/adamw_bfloat16/__init__.py:92-119).
Layout choices made here, all invisible through the state-dict view:
  * every parameter is a view into ONE flat bf16 buffer, in state-dict order, each start 16-byte aligned (TMA);
  * 3x3 conv weights keep their logical OIHW shape but are stored channels-last (O,kh,kw,I contiguous), which is
    exactly the K-major [Cout, 9*Cin] operand of the implicit-GEMM conv;
  * to_q/to_k/to_v (self) and to_k/to_v (cross) are adjacent, so the fused QKV / KV projection weight is a view.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

SDXL_BASE = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280),
    layers_per_block=2,
    transformer_layers_per_block=(0, 2, 10),  # 0 => block without attention (DownBlock2D / UpBlock2D)
    num_heads=(5, 10, 20),
    cross_attention_dim=2048,
    addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=2816,
    norm_num_groups=32,
    norm_eps=1e-5,
)


def _resnet(specs, pfx, cin, cout, temb):
    specs += [(f"{pfx}.norm1.weight", (cin,)), (f"{pfx}.norm1.bias", (cin,)),
              (f"{pfx}.conv1.weight", (cout, cin, 3, 3)), (f"{pfx}.conv1.bias", (cout,)),
              (f"{pfx}.time_emb_proj.weight", (cout, temb)), (f"{pfx}.time_emb_proj.bias", (cout,)),
              (f"{pfx}.norm2.weight", (cout,)), (f"{pfx}.norm2.bias", (cout,)),
              (f"{pfx}.conv2.weight", (cout, cout, 3, 3)), (f"{pfx}.conv2.bias", (cout,))]
    if cin != cout:
        specs += [(f"{pfx}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{pfx}.conv_shortcut.bias", (cout,))]


def _transformer(specs, pfx, dim, depth, ctx):
    specs += [(f"{pfx}.norm.weight", (dim,)), (f"{pfx}.norm.bias", (dim,)),
              (f"{pfx}.proj_in.weight", (dim, dim)), (f"{pfx}.proj_in.bias", (dim,))]
    for k in range(depth):
        b = f"{pfx}.transformer_blocks.{k}"
        specs += [(f"{b}.norm1.weight", (dim,)), (f"{b}.norm1.bias", (dim,)),
                  (f"{b}.attn1.to_q.weight", (dim, dim)), (f"{b}.attn1.to_k.weight", (dim, dim)),
                  (f"{b}.attn1.to_v.weight", (dim, dim)),
                  (f"{b}.attn1.to_out.0.weight", (dim, dim)), (f"{b}.attn1.to_out.0.bias", (dim,)),
                  (f"{b}.norm2.weight", (dim,)), (f"{b}.norm2.bias", (dim,)),
                  (f"{b}.attn2.to_q.weight", (dim, dim)), (f"{b}.attn2.to_k.weight", (dim, ctx)),
                  (f"{b}.attn2.to_v.weight", (dim, ctx)),
                  (f"{b}.attn2.to_out.0.weight", (dim, dim)), (f"{b}.attn2.to_out.0.bias", (dim,)),
                  (f"{b}.norm3.weight", (dim,)), (f"{b}.norm3.bias", (dim,)),
                  (f"{b}.ff.net.0.proj.weight", (dim * 8, dim)), (f"{b}.ff.net.0.proj.bias", (dim * 8,)),
                  (f"{b}.ff.net.2.weight", (dim, dim * 4)), (f"{b}.ff.net.2.bias", (dim,))]
    specs += [(f"{pfx}.proj_out.weight", (dim, dim)), (f"{pfx}.proj_out.bias", (dim,))]


def unet_param_specs(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (name, logical shape) list; order == diffusers / torch module registration order."""
    boc = cfg["block_out_channels"]
    temb = boc[0] * 4
    ctx = cfg["cross_attention_dim"]
    depth = cfg["transformer_layers_per_block"]
    L = cfg["layers_per_block"]
    s: List[Tuple[str, Tuple[int, ...]]] = []
    s += [("conv_in.weight", (boc[0], cfg["in_channels"], 3, 3)), ("conv_in.bias", (boc[0],))]
    s += [("time_embedding.linear_1.weight", (temb, boc[0])), ("time_embedding.linear_1.bias", (temb,)),
          ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]
    pin = cfg["projection_class_embeddings_input_dim"]
    s += [("add_embedding.linear_1.weight", (temb, pin)), ("add_embedding.linear_1.bias", (temb,)),
          ("add_embedding.linear_2.weight", (temb, temb)), ("add_embedding.linear_2.bias", (temb,))]
    skip = [boc[0]]
    cin = boc[0]
    for i, cout in enumerate(boc):
        last = i == len(boc) - 1
        # torch registers `resnets` before `attentions` in our oracle; diffusers registers attentions first for
        # cross-attn blocks.  Order inside the flat buffer is irrelevant to state-dict compatibility (a dict).
        for j in range(L):
            _resnet(s, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb)
        if depth[i] > 0:
            for j in range(L):
                _transformer(s, f"down_blocks.{i}.attentions.{j}", cout, depth[i], ctx)
        if not last:
            s += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
        skip += [cout] * L + ([] if last else [cout])
        cin = cout
    _transformer(s, "mid_block.attentions.0", boc[-1], depth[-1], ctx)
    _resnet(s, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    _resnet(s, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    rev, rdepth = list(reversed(boc)), list(reversed(depth))
    prev = boc[-1]
    for i, cout in enumerate(rev):
        last = i == len(rev) - 1
        res_in = [(prev if j == 0 else cout) + skip.pop() for j in range(L + 1)]
        for j, c in enumerate(res_in):
            _resnet(s, f"up_blocks.{i}.resnets.{j}", c, cout, temb)
        if rdepth[i] > 0:
            for j in range(L + 1):
                _transformer(s, f"up_blocks.{i}.attentions.{j}", cout, rdepth[i], ctx)
        if not last:
            s += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
        prev = cout
    s += [("conv_norm_out.weight", (boc[0],)), ("conv_norm_out.bias", (boc[0],)),
          ("conv_out.weight", (cfg["out_channels"], boc[0], 3, 3)), ("conv_out.bias", (cfg["out_channels"],))]
    return s


class ParamStore:
    """One flat bf16 parameter buffer + one flat bf16 gradient buffer with named views."""

    ALIGN = 8  # elements (16 bytes)

    def __init__(self, cfg: dict, device="cpu"):
        self.cfg = cfg
        self.specs = unet_param_specs(cfg)
        self.offsets: Dict[str, int] = {}
        off = 0
        # Placement order != state-dict order for one family: the cross-attention K / V projection weights of every block
        # with the same width are laid out back to back ([to_k; to_v] per block), so that ALL of them form one
        # [L * 2C, cross_dim] matrix — the text embeddings are step-invariant, so the 140 projections run as one GEMM per
        # width at the start of the forward pass and their weight gradients as one GEMM at the end of the backward pass.
        kv = [(n_, s_) for n_, s_ in self.specs if n_.endswith(".attn2.to_k.weight") or n_.endswith(".attn2.to_v.weight")]
        rest = [(n_, s_) for n_, s_ in self.specs if not (n_.endswith(".attn2.to_k.weight") or n_.endswith(".attn2.to_v.weight"))]
        self.kv_groups: Dict[int, List[str]] = {}     # width C -> block prefixes ("....attn2") in placement order
        for n_, s_ in kv:
            if n_.endswith(".to_k.weight"):
                self.kv_groups.setdefault(int(s_[0]), []).append(n_[:-len(".to_k.weight")])
        placed = list(rest)
        for C_ in sorted(self.kv_groups):
            for pfx in self.kv_groups[C_]:
                shp = dict(kv)
                placed += [(pfx + ".to_k.weight", shp[pfx + ".to_k.weight"]), (pfx + ".to_v.weight", shp[pfx + ".to_v.weight"])]
        for name, shape in placed:
            n = 1
            for d in shape:
                n *= d
            self.offsets[name] = off
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.numel = sum(int(torch.Size(s).numel()) for _, s in self.specs)
        self.flat = torch.zeros(self.total, dtype=torch.bfloat16, device=device)
        self.grad = torch.zeros(self.total, dtype=torch.bfloat16, device=device)
        # fp32 staging for the gradients of the small (1-D) parameters — biases and norm scale / shift: kernels accumulate
        # into it with atomics, `flush_small_grads()` folds it into `grad` once per backward pass.  Same order as the
        # specs, so a norm's (weight, bias) pair is one contiguous [2C] slice.
        self.small_off: Dict[str, int] = {}
        so = 0
        segs = []
        for name, shape in self.specs:
            if len(shape) == 1:
                self.small_off[name] = so
                segs += [so, self.offsets[name], int(shape[0])]
                so += (int(shape[0]) + 3) // 4 * 4
        self.small_total = so
        self.small32 = torch.zeros(max(so, 4), dtype=torch.float32, device=device)
        self._small_segs_host = torch.tensor(segs, dtype=torch.int64)
        self.small_segs = self._small_segs_host.to(device)
        self.n_small = len(segs) // 3
        self.params: "OrderedDict[str, torch.nn.Parameter]" = OrderedDict()
        # (tape position, "flat" | "small", offset, length) of every gradient view handed to a backward kernel while
        # logging is on — dp.plan_chunks() derives from it when each piece of the gradient buffer is final
        self.touch_log = None
        self._touch_pos = 0
        self._build_views()

    @staticmethod
    def _view(buf, off, shape):
        n = int(torch.Size(shape).numel())
        seg = buf[off:off + n]
        if len(shape) == 4 and shape[2] == 3:  # OIHW logical, O(kh)(kw)I physical == channels_last
            O, I, kh, kw = shape
            return seg.view(O, kh, kw, I).permute(0, 3, 1, 2)
        return seg.view(shape)

    def _build_views(self):
        self.params.clear()
        for name, shape in self.specs:
            p = torch.nn.Parameter(self._view(self.flat, self.offsets[name], shape), requires_grad=True)
            p.grad = self._view(self.grad, self.offsets[name], shape)
            self.params[name] = p

    def to(self, device):
        self.flat = self.flat.to(device)
        self.grad = self.grad.to(device)
        self.small32 = self.small32.to(device)
        self.small_segs = self._small_segs_host.to(device)
        self._build_views()
        return self

    # raw kernel-side views ------------------------------------------------------------------
    def w(self, name: str, rows: int, cols: int) -> torch.Tensor:
        """Row-major [rows, cols] view of the physical storage starting at `name` (may span adjacent params)."""
        off = self.offsets[name]
        return self.flat[off:off + rows * cols].view(rows, cols)

    def _touch(self, kind: str, off: int, n: int):
        if self.touch_log is not None:
            self.touch_log.append((self._touch_pos, kind, off, n))

    def g(self, name: str, rows: int, cols: int) -> torch.Tensor:
        off = self.offsets[name]
        self._touch("flat", off, rows * cols)
        return self.grad[off:off + rows * cols].view(rows, cols)

    def v(self, name: str) -> torch.Tensor:
        off = self.offsets[name]
        n = int(torch.Size(dict(self.specs)[name]).numel()) if False else self._numel[name]
        return self.flat[off:off + n]

    def gs(self, name: str, n: int = 0) -> torch.Tensor:
        """fp32 staging slice of a small parameter's gradient (n > 0: span `n` floats, e.g. a norm's weight+bias pair)."""
        off = self.small_off[name]
        self._touch("small", off, n or self._numel[name])
        return self.small32[off:off + (n or self._numel[name])]

    def flush_small_grads(self):
        from . import ops
        ops.flush_small_grads(self.small32, self.grad, self.small_segs, self.n_small)

    def gv(self, name: str) -> torch.Tensor:
        off = self.offsets[name]
        self._touch("flat", off, self._numel[name])
        return self.grad[off:off + self._numel[name]]

    @property
    def _numel(self):
        if not hasattr(self, "_numel_cache"):
            self._numel_cache = {n: int(torch.Size(s).numel()) for n, s in self.specs}
        return self._numel_cache

    def kv_group_view(self, C: int, grad: bool = False) -> torch.Tensor:
        """[L * 2C, cross_dim] view over the stacked cross-attention K/V projection weights (or their gradients)."""
        pfxs = self.kv_groups[C]
        cdim = self.cfg["cross_attention_dim"]
        first = pfxs[0] + ".to_k.weight"
        for a, b in zip(pfxs, pfxs[1:]):
            assert self.offsets[b + ".to_k.weight"] == self.offsets[a + ".to_k.weight"] + 2 * C * cdim
        off = self.offsets[first]
        if grad:
            self._touch("flat", off, len(pfxs) * 2 * C * cdim)
        buf = self.grad if grad else self.flat
        return buf[off:off + len(pfxs) * 2 * C * cdim].view(len(pfxs) * 2 * C, cdim)

    def adjacent(self, *names) -> bool:
        """True if the given params are stored back to back (fused-weight views are valid)."""
        for a, b in zip(names, names[1:]):
            if self.offsets[a] + self._numel[a] != self.offsets[b]:
                return False
        return True

    # state dict ------------------------------------------------------------------------------
    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, p.detach()) for k, p in self.params.items())

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.params if k not in sd]
        unexpected = [k for k in sd if k not in self.params]
        if strict and (missing or unexpected):
            raise KeyError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        with torch.no_grad():
            for k, p in self.params.items():
                if k in sd:
                    p.copy_(sd[k].to(device=p.device, dtype=p.dtype))
        return missing, unexpected
