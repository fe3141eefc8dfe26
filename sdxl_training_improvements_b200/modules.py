"""Single-module entry points over the UNet engine: one ResnetBlock2D or one BasicTransformerBlock, forward + backward,
on the same kernels and the same tape machinery the whole UNet uses.

Why this exists: BASELINE.json `configs[0]` is a module-level case ("single ResnetBlock2D (320ch, 64x64 latent) fwd/bwd on
CPU eager vs new kernel — loss match"), and SURVEY.md §4.1 T3 asks for module tests (ResnetBlock2D, BasicTransformerBlock)
against the oracle's modules (diffusers `resnet.py` / `attention.py` restated in oracle/unet_sdxl.py).  The runner builds the
smallest parameter store that contains the module — a one-level UNet config of the requested width — and drives
`UNetEngine.resnet` / `UNetEngine.transformer_block` directly, so there is no second implementation to keep in sync.

Layouts at this boundary are the reference's (NCHW activations for the resnet, [B, n, C] tokens for the transformer block);
the token-major / NHWC layout of the kernels stays internal.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops
from .params import SDXL_BASE
from .unet import Act, B200UNet, bf16


def one_level_config(width: int, depth: int = 0, cross_attention_dim: int = 2048) -> dict:
    """A UNet config with a single resolution level of `width` channels (`depth` transformer blocks per attention)."""
    cfg = dict(SDXL_BASE)
    cfg.update(block_out_channels=(width,), transformer_layers_per_block=(depth,), num_heads=(width // 64,),
               cross_attention_dim=cross_attention_dim)
    return cfg


class ModuleRunner:
    """Holds a one-level parameter store and runs single modules of it.

    Module prefixes available (diffusers names): `down_blocks.0.resnets.{0,1}` (width -> width),
    `up_blocks.0.resnets.{0,1,2}` (2*width -> width, with the 1x1 `conv_shortcut`), and with depth > 0
    `down_blocks.0.attentions.{0,1}.transformer_blocks.{k}` / `mid_block.attentions.0.transformer_blocks.{k}`.
    """

    def __init__(self, width: int, depth: int = 0, cross_attention_dim: int = 2048, device="cuda"):
        self.cfg = one_level_config(width, depth, cross_attention_dim)
        self.net = B200UNet(self.cfg, device=device)
        self.eng = self.net.engine
        self.store = self.net.store
        self.width = width
        self.temb_ch = width * 4
        self._tape = None

    # ---- parameters --------------------------------------------------------------------------------------------
    def load_module_state(self, prefix: str, sd: Dict[str, torch.Tensor]):
        """Copy a module's state dict (keys relative to `prefix`, diffusers layout: conv OIHW, Linear [out, in])."""
        params = dict(self.net.named_parameters())
        with torch.no_grad():
            for k, v in sd.items():
                p = params[f"{prefix}.{k}"]
                p.copy_(v.to(device=p.device, dtype=p.dtype).reshape(p.shape))

    def module_grads(self, prefix: str) -> Dict[str, torch.Tensor]:
        return {k[len(prefix) + 1:]: p.grad for k, p in self.net.named_parameters() if k.startswith(prefix + ".")}

    # ---- layout helpers ----------------------------------------------------------------------------------------
    @staticmethod
    def _to_tokens(x_nchw: torch.Tensor) -> torch.Tensor:
        B, Cc, H, W = x_nchw.shape
        return x_nchw.permute(0, 2, 3, 1).reshape(B * H * W, Cc).to(bf16).contiguous()

    @staticmethod
    def _to_nchw(t: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        return t.view(B, H, W, -1).permute(0, 3, 1, 2)

    def _run_tape_backward(self):
        tape = self.eng.tape
        self.eng.tape = []
        for i, f in enumerate(reversed(tape)):
            self.store._touch_pos = i
            f()
        self.store.flush_small_grads()

    # ---- ResnetBlock2D -----------------------------------------------------------------------------------------
    def resnet_forward(self, prefix: str, x_nchw: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
        """diffusers `ResnetBlock2D.forward(x, emb)`: x [B, Cin, H, W], emb [B, temb_ch] (SiLU applied inside)."""
        B, Cin, H, W = x_nchw.shape
        Cout = self.store._numel[f"{prefix}.conv1.bias"]
        self.eng.tape = []
        self._x = Act(self._to_tokens(x_nchw))
        self._emb = Act(emb.to(bf16).contiguous())
        emb_silu = self.eng.silu(self._emb)
        self._out = self.eng.resnet(self._x, emb_silu, B, H, W, Cin, Cout, prefix)
        self._shape = (B, H, W)
        return self._to_nchw(self._out.d, B, H, W)

    def resnet_backward(self, dout_nchw: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Accumulates parameter gradients into the store; returns (dx NCHW, demb)."""
        B, H, W = self._shape
        self._out.g = self._to_tokens(dout_nchw)
        self._run_tape_backward()
        return self._to_nchw(self._x.g, B, H, W), self._emb.g

    # ---- BasicTransformerBlock ---------------------------------------------------------------------------------
    def transformer_block_forward(self, prefix: str, x: torch.Tensor, ctx: torch.Tensor) -> torch.Tensor:
        """diffusers `BasicTransformerBlock.forward(x, encoder_hidden_states=ctx)`: x [B, n, C], ctx [B, n_ctx, cross_dim]."""
        B, n, Cc = x.shape
        n_ctx = ctx.shape[1]
        ctx2 = ctx.to(bf16).reshape(B * n_ctx, -1).contiguous()
        self.eng.tape = []
        self.eng._project_context(ctx2)   # grouped K / V projection of every block of this width (as in the full forward)
        self._x = Act(x.to(bf16).reshape(B * n, Cc).contiguous())
        self._out = self.eng.transformer_block(self._x, B, n, Cc, prefix, ctx2, n_ctx)
        self._shape = (B, n, Cc)
        return self._out.d.view(B, n, Cc)

    def transformer_block_backward(self, dout: torch.Tensor) -> torch.Tensor:
        B, n, Cc = self._shape
        self._out.g = dout.to(bf16).reshape(B * n, Cc).contiguous()
        # K / V weight gradients of blocks that did not run are zero: their dkv slices must not hold garbage
        for _, dkv in self.eng._kv_slices.values():
            dkv.zero_()
        pending = list(self.eng._kv_wgrad.values())   # widths whose stacked-wgrad closure was not claimed by a block
        self._run_tape_backward()
        for f in pending:
            f()
        return self._x.g.view(B, n, Cc)

    def zero_grad(self):
        self.net.zero_grad()
        self.store.small32.zero_()


def resnet_block_flops(B: int, H: int, W: int, Cin: int, Cout: int, temb_ch: int = 1280) -> float:
    """Forward 2*MAC of one ResnetBlock2D (SURVEY.md §8d "Conv roofline figure": the two 3x3 convs dominate)."""
    px = B * H * W
    f = 2.0 * px * Cout * 9 * Cin + 2.0 * px * Cout * 9 * Cout + 2.0 * B * Cout * temb_ch
    if Cin != Cout:
        f += 2.0 * px * Cout * Cin
    return f
