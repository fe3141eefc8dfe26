"""Drop-in training-step plugins for the reference's `src/training` surface, running on the B200 kernels.

Mirrors (same names, argument meaning and error behaviour):
  * `DDPMTrainer.training_step(batch)`            src/training/trainers/methods/ddpm_trainer.py:280-405
  * `FlowMatchingTrainer.compute_loss(model, batch, generator=None)`  .../flow_matching_trainer.py:261-356
  * `NoiseScheduler` (Karras / ZTSNR table, sample_timesteps)         src/training/schedulers/novelai_v3.py:101-184
  * accumulate / clip / step protocol             .../example_method.py:124-148,191-206 (defect B4 decision)
  * DDP contract: all-reduce of the UNet gradients only (src/core/distributed.py:142-163, B9)

Host-side logic (which timesteps, which sigma) is the reference's own tiny torch-CPU arithmetic restated; all tensor
work on latents/activations/gradients runs in libsdxl_b200 kernels: Philox noise, noising, target, UNet fwd/bwd,
MSE (+min-SNR / tag weights), clamp/fallback, grad-norm, AdamW.
"""
from __future__ import annotations

import math
import time
from types import SimpleNamespace
from typing import Any, Dict, Optional

import torch

from . import ops
from .unet import IN_PAD, PRED_PAD, B200UNet, bf16

LATENT_PAD = IN_PAD  # channel padding of the noisy latents handed to conv_in; the prediction comes back PRED_PAD wide


class FatalStepError(RuntimeError):
    """A training step failed after part of the data-parallel gradient exchange had been issued: peers already hold
    (and have flagged) some of this step's chunks, so the step cannot be skipped or retried — the job must stop."""


# ----------------------------------------------------------------------------------------------------------------
# schedule (host side; reference: novelai_v3.py)
# ----------------------------------------------------------------------------------------------------------------
def get_karras_sigmas(n_sigmas: int, sigma_min: float, sigma_max: float, rho: float = 7.0, device=None) -> torch.Tensor:
    """novelai_v3.py:160-184 — descending fp32 table (index 0 = sigma_max)."""
    ramp = torch.linspace(0, 1, n_sigmas, device=device)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho


class NoiseScheduler:
    """The subset of the reference's NoiseScheduler the training step uses (novelai_v3.py:101-151)."""

    def __init__(self, config=None, device="cpu"):
        m = getattr(config, "model", None)
        self.num_timesteps = int(getattr(m, "num_timesteps", 1000))
        self.sigma_min = float(getattr(m, "sigma_min", 0.002))
        self.sigma_max = float(getattr(m, "sigma_max", 20000.0))
        self.use_ztsnr = bool(getattr(m, "use_ztsnr", True))
        self.rho = float(getattr(m, "rho", 7.0))  # B2: ModelConfig has no rho -> function default
        self.sigma_data = 1.0
        self.device = device
        # the reference rebuilds this table on every call (novelai_v3.py:134-137); it is step-invariant
        self.sigmas = get_karras_sigmas(self.num_timesteps, self.sigma_min,
                                        20000.0 if self.use_ztsnr else self.sigma_max, self.rho)

    def get_sigmas(self, n: int) -> torch.Tensor:
        return get_karras_sigmas(n, self.sigma_min, 20000.0 if self.use_ztsnr else self.sigma_max, self.rho)

    def timestep_to_sigma(self, timesteps: torch.Tensor) -> torch.Tensor:
        return self.sigmas[timesteps.cpu()]

    def get_snr(self, timesteps: torch.Tensor) -> torch.Tensor:
        return (self.sigma_data / self.timestep_to_sigma(timesteps)) ** 2

    def sample_timesteps(self, batch_size: int, device=None, generator=None) -> torch.Tensor:
        """novelai_v3.py:139-151 (B1: `device` accepted and ignored; B13: uniform)."""
        if self.use_ztsnr:
            u = torch.rand(batch_size, generator=generator)
            return (u * self.num_timesteps).long()
        return torch.randint(0, self.num_timesteps, (batch_size,), generator=generator)


def sample_logit_normal(shape, dtype=bf16, mean=0.0, std=1.0, generator=None) -> torch.Tensor:
    """flow_matching_trainer.py:373-385, drawn in the model dtype (B20)."""
    normal = torch.randn(shape, dtype=dtype, generator=generator)
    return torch.sigmoid(mean + std * normal)


# ----------------------------------------------------------------------------------------------------------------
# fused loss: noise -> noising -> UNet -> MSE, as ONE autograd node
# ----------------------------------------------------------------------------------------------------------------
class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, core, anchor, args):
        loss = core._forward(**args)
        ctx.core = core
        ctx.saved = core._saved
        core._saved = None
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        tape, dpred = ctx.saved
        ctx.saved = None
        ops.scale_bf16(dpred, grad_out.reshape(1).float().contiguous(), 1.0)
        ctx.core._backward(dpred, tape)
        return None, None, None


class _PrecomputedBackward(torch.autograd.Function):
    """Loss of a graph-replayed micro-step: gradients are already in p.grad, so backward is a no-op."""

    @staticmethod
    def forward(ctx, loss, anchor):
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


class FusedLossCore:
    """Device-side part of one micro-step.  `method`: "ddpm" | "flow_matching"."""

    def __init__(self, unet: B200UNet, method: str, prediction_type: str = "v_prediction", use_ztsnr: bool = True,
                 seed: int = 0):
        self.unet = unet
        self.method = method
        self.prediction_type = prediction_type
        self.use_ztsnr = use_ztsnr
        dev = unet.device
        self.seed_offset = torch.tensor([seed, 0], device=dev, dtype=torch.int64)
        self.loss_sum = torch.zeros(1, device=dev, dtype=torch.float64)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float32)
        self.ok = torch.zeros(1, device=dev, dtype=torch.int32)
        self.stats = torch.zeros(6, device=dev, dtype=torch.float64)  # |eps|,eps^2,|pred|,pred^2,|x|,x^2
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._saved = None
        self.last: Dict[str, torch.Tensor] = {}
        # data parallel (dp.PeerGradExchange or None): when `dp_last` is set — the trainer announces the last accumulation
        # micro-step of an optimizer step — the backward pass hands each finished chunk of the gradient buffer to the
        # exchange while the rest of it is still running
        self.dp = None
        self.dp_last = False

    def _backward(self, dpred, tape):
        x = self.dp
        eng, st = self.unet.engine, self.unet.store
        if x is None:
            eng.backward(dpred, tape)
            return
        n_tape = len(tape[0])
        if x.plan is None or x.plan.n_tape != n_tape:
            # first pass: log which tape position last writes each gradient, derive the chunk plan (a pure function of
            # the log, hence identical on every rank), exchange the whole buffer in one piece this time
            from .dp import plan_chunks
            st.touch_log = []
            eng.backward(dpred, tape)
            log, st.touch_log = st.touch_log, None
            x.set_plan(plan_chunks(st, log, n_tape))
            if self.dp_last:
                x.exchange_all()
            return
        if not self.dp_last:
            eng.backward(dpred, tape)
            return

        def on_cut(k):
            x.flush_chunk(st, k)
            x.exchange_chunk(k)

        eng.backward(dpred, tape, cuts=x.plan.cuts, on_cut=on_cut)

    def _forward(self, latents, ctx, pooled, time_ids, t_embed, sig_or_t, weight, loss_scale, noise=None):
        """latents fp32 NCHW holding bf16-representable values; returns 0-d fp32 loss (device)."""
        B, Cc, H, W = latents.shape
        HW = H * W
        n = B * Cc * HW
        if noise is None:
            noise = ops.randn(n, self.seed_offset, 1, round_bf16=True)
            ops.philox_advance(self.seed_offset, 1)
        mode = 0 if self.method == "ddpm" else 1
        noisy, target = ops.make_noisy(latents, noise, sig_or_t, mode, self.prediction_type == "v_prediction",
                                       self.use_ztsnr, B, Cc, HW, LATENT_PAD)
        pred = self.unet.engine.forward(noisy, t_embed, ctx, pooled, time_ids, B, H, W)
        tape = self.unet.engine.detach_tape()
        dpred = torch.empty_like(pred.d)
        self.loss_sum.zero_()
        self.stats.zero_()
        ops.mse_loss(pred.d, target, weight, self.loss_sum, dpred, loss_scale / n, B, Cc, HW, PRED_PAD)
        ops.finalize_loss(self.loss_sum, n, loss_scale, self.loss, self.ok, dpred)
        ops.abs_sq_sums(noise, self.stats[0:2])
        ops.abs_sq_sums(pred.d, self.stats[2:4], PRED_PAD, Cc)
        ops.abs_sq_sums(latents, self.stats[4:6])
        self._saved = ((tape[0], tape[1]), dpred)
        self.last = {"pred": pred.d, "target": target, "noisy": noisy, "noise": noise, "numel": n}
        return self.loss.clone().reshape(())

    def loss_fn(self, **args) -> torch.Tensor:
        return _FusedLoss.apply(self, self._anchor, args)

    def step_no_autograd(self, grad_scale: float = 1.0, grad_scale_dev: Optional[torch.Tensor] = None, **args) -> torch.Tensor:
        """forward + backward without torch autograd (used by the graph-captured bench/step path).  `grad_scale_dev`
        (1-element fp32 device tensor) carries autograd's upstream scalar (1 / accumulation steps) in graph replays."""
        loss = self._forward(**args)
        tape, dpred = self._saved
        self._saved = None
        if grad_scale_dev is not None:
            ops.scale_bf16(dpred, grad_scale_dev, 1.0)
        elif grad_scale != 1.0:
            ops.scale_bf16(dpred, None, grad_scale)
        self._backward(dpred, tape)
        return loss


class GraphedMicroStep:
    """One micro-step (Philox noise -> noising -> UNet forward -> loss -> UNet backward into the flat gradient buffer)
    captured as a CUDA graph for a fixed (B, H, W): ~4,700 kernel launches become one graph launch, and the kernels run
    back to back without the stream-launch gaps.  Inputs live in static device buffers (`load()` copies into them);
    everything that varies per step (noise counter, sigma / t, grad scale) is read from device memory."""

    def __init__(self, core: "FusedLossCore", B: int, H: int, W: int, n_ctx: int = 77):
        self.core = core
        cfg = core.unet.config
        dev = core.unet.device
        self.shape = (B, H, W)
        pooled_dim = cfg["projection_class_embeddings_input_dim"] - 6 * cfg["addition_time_embed_dim"]
        self.latents = torch.zeros(B, cfg["in_channels"], H, W, device=dev, dtype=torch.float32)
        self.ctx = torch.zeros(B * n_ctx, cfg["cross_attention_dim"], device=dev, dtype=bf16)
        self.pooled = torch.zeros(B, pooled_dim, device=dev, dtype=bf16)
        self.time_ids = torch.zeros(B, 6, device=dev, dtype=torch.float32)
        self.t_embed = torch.zeros(B, device=dev, dtype=torch.float32)
        self.sig_or_t = torch.ones(B, device=dev, dtype=torch.float32)
        self.weight = torch.ones(B, device=dev, dtype=torch.float32)
        self.grad_scale = torch.ones(1, device=dev, dtype=torch.float32)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graphs: Optional[list] = None  # data parallel: one graph per exchange chunk
        self.loss: Optional[torch.Tensor] = None
        self.static_last: Dict[str, torch.Tensor] = {}
        self.launches_per_replay = 0
        self._pin: Optional[torch.Tensor] = None
        self._pin_i = 0

    def _run(self):
        return self.core.step_no_autograd(grad_scale_dev=self.grad_scale, latents=self.latents, ctx=self.ctx,
                                          pooled=self.pooled, time_ids=self.time_ids, t_embed=self.t_embed,
                                          sig_or_t=self.sig_or_t, weight=self.weight, loss_scale=1.0)

    def capture(self, pool=None):
        """Call at an optimizer-step boundary: the warm-up passes accumulate into the gradient buffer, which is zeroed
        again afterwards.  Data parallel (core.dp set): the micro-step is captured as one graph PER CHUNK of the exchange
        plan — segment k ends when chunk k of the gradient buffer is final — sharing one memory pool and replayed in
        order; `replay(last=True)` starts the exchange of chunk k between segment k and k+1."""
        from . import _lib
        core = self.core
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        was_last, core.dp_last = core.dp_last, False
        with torch.cuda.stream(side):
            for _ in range(2):  # lazy one-time setup (function attributes, grow-only workspaces, the dp chunk plan)
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the warm-up passes' activations go back to the driver before the pool grows
        n0 = _lib.launch_count()
        x = core.dp
        if x is None:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, pool=pool):
                self.loss = self._run()
        else:
            eng, st = core.unet.engine, core.unet.store
            cuts = x.plan.cuts
            self.graphs = []
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self.loss = core._forward(latents=self.latents, ctx=self.ctx, pooled=self.pooled, time_ids=self.time_ids,
                                          t_embed=self.t_embed, sig_or_t=self.sig_or_t, weight=self.weight, loss_scale=1.0)
                tape, dpred = core._saved
                core._saved = None
                ops.scale_bf16(dpred, self.grad_scale, 1.0)
                assert len(tape[0]) == x.plan.n_tape
                order = eng.backward_begin(dpred, tape)
                eng.backward_span(order, 0, cuts[0] + 1)
                x.flush_chunk(st, 0)
            self.graphs.append(g)
            lo = cuts[0] + 1
            for k in range(1, len(cuts)):
                gk = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gk, pool=g.pool()):
                    eng.backward_span(order, lo, cuts[k] + 1)
                    x.flush_chunk(st, k)
                lo = cuts[k] + 1
                self.graphs.append(gk)
            assert lo == len(order)
            order.clear()
        core.dp_last = was_last
        # static tensors of the captured micro-step (noise drawn in-graph, noisy latents, target, prediction): a replay
        # refreshes them in place; parity tests read the noise from here to feed the oracle the same draw
        self.static_last = dict(core.last)
        self.launches_per_replay = _lib.launch_count() - n0
        self.core.unet.store.grad.zero_()
        torch.cuda.synchronize()
        return self

    def _h2d(self, dst: torch.Tensor, src: torch.Tensor):
        """dst <- src without stalling the host: an async copy from PAGEABLE host memory makes the driver synchronise the
        stream first (the host then sits behind the 11 ms optimizer graph and the GPU idles while the next micro-step is
        being enqueued: measured ~2 ms per step end to end).  Small host tensors (timesteps, sigmas, weights) therefore go
        through a pinned staging row; two rows alternate so that a row is never rewritten while its copy may be in flight."""
        if src.is_cuda or src.is_pinned():
            dst.copy_(src, non_blocking=True)
            return
        if self._pin is None or self._pin.shape[1] < dst.numel():
            self._pin = torch.zeros(8, max(16, dst.numel()), dtype=torch.float32).pin_memory()
        row = self._pin[self._pin_i % 8, :dst.numel()].view(dst.shape)
        self._pin_i += 1
        row.copy_(src)
        dst.copy_(row, non_blocking=True)

    def load(self, latents, ctx, pooled, time_ids, t_embed, sig_or_t, weight=None, grad_scale: float = 1.0):
        self.latents.copy_(latents, non_blocking=True)
        self.ctx.copy_(ctx, non_blocking=True)
        self.pooled.copy_(pooled, non_blocking=True)
        self.time_ids.copy_(time_ids, non_blocking=True)
        self._h2d(self.t_embed, t_embed)
        self._h2d(self.sig_or_t, sig_or_t)
        if weight is None:
            self.weight.fill_(1.0)
        else:
            self._h2d(self.weight, weight)
        self.grad_scale.fill_(float(grad_scale))

    def replay(self, last: bool = False) -> torch.Tensor:
        """`last`: this is the last accumulation micro-step of an optimizer step -> exchange each chunk as it completes."""
        if self.graphs is None:
            self.graph.replay()
            return self.loss
        x = self.core.dp
        for k, g in enumerate(self.graphs):
            g.replay()
            if last:
                x.exchange_chunk(k)
        return self.loss


class GraphedOptimizerStep:
    """grad-norm + clip + fused optimizer update + gradient zeroing as one CUDA graph (the step counter and the clip
    coefficient are device-side, so replays stay correct).  Host-side bookkeeping that cannot be captured (the
    reference's per-tensor deferred weight decay) runs after the replay via `optimizer.after_graph_step()`."""

    def __init__(self, optimizer, max_norm: float, grad_scale: float, gnorm_ready: bool = False):
        self.optimizer = optimizer
        self.max_norm, self.gscale = max_norm, grad_scale
        self.gnorm_ready = gnorm_ready  # data parallel: the exchange delivers sum(grad^2); no norm pass in this graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_replay = 0
        self._hyper = None

    def _hyper_now(self):
        g = self.optimizer.param_groups[0]
        return (float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"]), float(g.get("weight_decay", 0.0)),
                float(self.max_norm), float(self.gscale))

    def capture(self):
        """Captures WITHOUT executing: optimizer state is not advanced by the capture itself.  lr / betas / eps / weight
        decay / clip norm are kernel ARGUMENTS, i.e. baked into the graph: `replay()` re-captures when any of them has
        changed (an lr schedule that changes every step should drive the un-captured `fused_step` instead)."""
        from . import _lib
        opt = self.optimizer
        self._hyper = self._hyper_now()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if isinstance(opt, B200AdamWBF16):  # zero_grad() folded into the optimizer kernel (no separate 5 GB fill)
                opt.fused_step(max_norm=self.max_norm, grad_scale=self.gscale, _in_graph=True, zero_grad=True,
                               gnorm_ready=self.gnorm_ready)
            else:
                opt.fused_step(max_norm=self.max_norm, grad_scale=self.gscale, _in_graph=True, gnorm_ready=self.gnorm_ready)
                opt.zero_grad()
        self.launches_per_replay = _lib.launch_count() - n0
        torch.cuda.synchronize()
        return self

    def replay(self):
        if self._hyper_now() != self._hyper:
            self.capture()
        self.graph.replay()
        self.optimizer.after_graph_step()


def _prep_batch(batch: Dict[str, Any], device) -> Dict[str, torch.Tensor]:
    """Batch-dict contract of the reference (dataset.py:209-229; ddpm_trainer.py:284-296).  Tensors may arrive on CPU
    or GPU in any float dtype; cast to the UNet dtype like the flow path does (flow_matching_trainer.py:288-291, B21)."""
    required = {"vae_latents", "prompt_embeds", "pooled_prompt_embeds", "time_ids", "metadata"}
    if not all(k in batch for k in required):
        raise ValueError(f"Batch missing required keys: {required - set(batch.keys())}")
    lat = batch["vae_latents"].to(device, non_blocking=True)
    B = lat.shape[0]
    return {
        "latents": lat.to(bf16).float().contiguous(),
        "ctx": batch["prompt_embeds"].to(device, non_blocking=True).to(bf16).reshape(-1, batch["prompt_embeds"].shape[-1]).contiguous(),
        "pooled": batch["pooled_prompt_embeds"].to(device, non_blocking=True).to(bf16).reshape(B, -1).contiguous(),
        "time_ids": batch["time_ids"].to(device, non_blocking=True).float().reshape(B, -1).contiguous(),
    }


def _tag_weight_mean(batch) -> Optional[float]:
    """ddpm_trainer.py:348-368: mean over samples of the mean tag weight, if every sample has weights.  The reference
    builds the per-sample weights as a tensor in the model dtype (bf16, :365) and takes `.mean()` of that."""
    md = batch.get("metadata")
    if isinstance(md, (list, tuple)) and md and all(isinstance(m, dict) and isinstance(m.get("tag_info"), dict)
                                                     and "tags" in m["tag_info"] for m in md):
        ws = []
        for m in md:
            iw = [td["weight"] for tags in m["tag_info"]["tags"].values() for td in tags]
            if not iw:
                return None
            ws.append(sum(iw) / len(iw))
        return float(torch.tensor(ws, dtype=bf16).mean())
    return None


class _StepBase:
    """Shared constructor signature of the reference trainers (base_router.py:15-31; sdxl_trainer.py:18-36)."""

    def __init__(self, model, optimizer, train_dataloader=None, device=None, wandb_logger=None, config=None, **kwargs):
        self.model = model
        self.unet: B200UNet = getattr(model, "unet", model)
        if not isinstance(self.unet, B200UNet):
            raise TypeError("model.unet must be a B200UNet (there is no CPU / diffusers fallback on this path)")
        self.optimizer = optimizer
        self.train_dataloader = train_dataloader
        self.device = torch.device(device) if device is not None else self.unet.device
        self.wandb_logger = wandb_logger
        self.config = config
        tr = getattr(config, "training", None)
        self.gradient_accumulation_steps = int(getattr(tr, "gradient_accumulation_steps", 1))
        self.clip_grad_norm = float(getattr(tr, "clip_grad_norm", 1.0))
        self.prediction_type = getattr(tr, "prediction_type", "v_prediction")
        self.world_size = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        # cuda_graph=True: the micro-step (forward + backward) and the optimizer step replay captured CUDA graphs, one
        # per latent shape.  In that mode the backward pass runs inside compute_loss() (gradients accumulate into
        # p.grad with the 1/accumulation scale the step protocol announces), and `loss.backward()` is a no-op.
        self.cuda_graph = bool(kwargs.get("cuda_graph", False))
        self._micro_graphs: Dict[Any, GraphedMicroStep] = {}
        self._graph_pool = None                      # shared by all captured shapes
        self.max_graphed_shapes = int(kwargs.get("max_graphed_shapes", 8))
        self.uncaptured_micro_steps = 0              # micro-steps cuda_graph mode ran un-captured (new shape mid-window / cap)
        # graph mode driven by _execute_training_step: the optimizer graph is enqueued right behind the last micro-step's
        # graph, BEFORE the host reads the step's loss / metrics (which travel through a pinned buffer + an event) — the GPU
        # does not idle while Python builds the metrics dict and walks the (no-op) autograd backward
        self._opt_after_micro = False
        self._opt_done_early = False
        self._host_vals: Optional[torch.Tensor] = None   # pinned [7] float64: loss, stats[0:6]
        self._host_event: Optional[torch.cuda.Event] = None
        self._opt_graph: Optional[GraphedOptimizerStep] = None
        self._pending_grad_scale = 1.0

    def _lr(self):
        try:
            return float(self.optimizer.param_groups[0]["lr"])
        except Exception:
            return 0.0

    # --- accumulate / clip / step protocol (example_method.py:124-148, 191-206; flow_matching_trainer.py:172-189) ---
    def _execute_training_step(self, batch, accumulate: bool = False, is_last_accumulation_step: bool = True):
        self._pending_grad_scale = 1.0 / self.gradient_accumulation_steps if accumulate else 1.0
        # the backward pass of the last micro-step hands finished gradient chunks to the data-parallel exchange
        self.core.dp_last = (not accumulate) or is_last_accumulation_step
        self._opt_after_micro = (self.cuda_graph and hasattr(self.optimizer, "fused_step")
                                 and ((not accumulate) or is_last_accumulation_step))
        self._opt_done_early = False
        try:
            out = self.compute_loss(batch) if not isinstance(self, B200FlowMatchingTrainer) else self.compute_loss(self.model, batch)
            loss = out["loss"]
            if accumulate:
                loss = loss / self.gradient_accumulation_steps
            loss.backward()
        except Exception as e:
            x = self.core.dp
            if x is not None and x.issued:
                # chunks of this step are already with the peers under the current sequence number: a "skip and continue"
                # (ddpm_trainer.py:202-204) would re-issue them unsynchronised and silently corrupt the gradients
                raise FatalStepError(f"step failed after the gradient exchange had started: {type(e).__name__}: {e}") from e
            raise
        finally:
            self.core.dp_last = False
            self._opt_after_micro = False
        if (not accumulate or is_last_accumulation_step) and not self._opt_done_early:
            self.optimizer_step()
        self._opt_done_early = False
        return out["loss"].detach(), out["metrics"]

    def _graphed_loss(self, t: Dict[str, torch.Tensor], t_embed, sig_or_t, weight) -> torch.Tensor:
        """Replay (capturing on first use) the micro-step graph for this latent shape; returns a loss tensor whose
        `.backward()` does nothing because the backward pass already ran inside the graph."""
        B, _, H, W = t["latents"].shape
        key = (B, H, W, t["ctx"].shape[0] // B)
        gm = self._micro_graphs.get(key)
        if gm is None:
            # A capture runs warm-up passes that accumulate into the gradient buffer, so it can only happen at an optimizer-
            # step boundary (gradients all zero).  A shape first seen in the middle of an accumulation window (aspect buckets
            # + gradient_accumulation_steps > 1) therefore runs this micro-step through the SAME kernels un-captured and is
            # captured the next time it shows up at a boundary; shapes beyond `max_graphed_shapes` always run un-captured.
            at_boundary = False
            if len(self._micro_graphs) < self.max_graphed_shapes:
                chk = torch.zeros(1, device=self.unet.device, dtype=torch.float64)
                ops.sumsq(self.unet.store.grad, chk)
                at_boundary = float(chk) == 0.0
            if not at_boundary:
                self.uncaptured_micro_steps += 1
                loss = self.core.step_no_autograd(grad_scale=self._pending_grad_scale, latents=t["latents"], ctx=t["ctx"],
                                                  pooled=t["pooled"], time_ids=t["time_ids"],
                                                  t_embed=t_embed.to(self.unet.device), sig_or_t=sig_or_t.to(self.unet.device),
                                                  weight=None if weight is None else weight.to(self.unet.device),
                                                  loss_scale=1.0)
                return _PrecomputedBackward.apply(loss.reshape(()), self.core._anchor)
            torch.cuda.empty_cache()
            if self._graph_pool is None:
                # ONE memory pool for every captured shape: a micro-step graph is self-contained (activations die inside
                # it) and graphs replay one at a time on one stream, so the pool's size is the LARGEST shape's activation
                # set (~40 GB at 1024^2, B=4), not the sum over buckets
                self._graph_pool = torch.cuda.graph_pool_handle()
            gm = GraphedMicroStep(self.core, B, H, W, key[3])
            gm.load(t["latents"], t["ctx"], t["pooled"], t["time_ids"], t_embed, sig_or_t, weight, 1.0)
            gm.capture(pool=self._graph_pool)
            self._micro_graphs[key] = gm
        gm.load(t["latents"], t["ctx"], t["pooled"], t["time_ids"], t_embed, sig_or_t, weight, self._pending_grad_scale)
        loss = gm.replay(last=self.core.dp_last)
        self.core.last = {"numel": t["latents"].numel()}
        out = _PrecomputedBackward.apply(loss.reshape(()), self.core._anchor)
        if self._opt_after_micro:
            # loss + metric sums -> pinned host memory behind the micro-step, then the optimizer graph right away
            if self._host_vals is None:
                self._host_vals = torch.zeros(7, dtype=torch.float64).pin_memory()
                self._host_event = torch.cuda.Event()
                self._dev_vals = torch.zeros(7, device=self.unet.device, dtype=torch.float64)
            self._dev_vals[0:1].copy_(out.detach().reshape(1))
            self._dev_vals[1:7].copy_(self.core.stats)
            self._host_vals.copy_(self._dev_vals, non_blocking=True)
            self._host_event.record()
            self.optimizer_step()
            self._opt_done_early = True
        return out

    def _read_loss_and_stats(self, loss: torch.Tensor):
        """(python float loss, list of the six metric sums): ONE host read per micro-step.  After _graphed_loss() enqueued the
        optimizer graph early the values come from the pinned buffer (waiting only for the micro-step, not the update)."""
        if self._opt_done_early and self._host_event is not None:
            self._host_event.synchronize()
            v = self._host_vals.tolist()
            return v[0], v[1:]
        st = self.core.stats.tolist()
        return float(loss.detach()), st

    def _init_dp(self):
        """Data parallel: gradients are exchanged over NVSwitch peer memory beside the backward pass (dp.py); if that
        cannot be set up on every rank, by ONE NCCL all-reduce of the flat buffer before the optimizer step."""
        if self.world_size > 1:
            from .dp import try_create_exchange
            self.core.dp = try_create_exchange(self.unet.store.grad)

    def optimizer_step(self):
        norm_ready = False
        if self.world_size > 1:
            x = self.core.dp
            if x is not None:
                if not x.issued:  # nobody announced the last micro-step: exchange the whole buffer now, no overlap
                    x.exchange_all()
                # the exchange's reduce kernels have already summed the squares of the reduced gradients: with a fused
                # optimizer and clipping on, finish() delivers the global norm and the optimizer skips its own 5 GB pass
                norm_ready = bool(getattr(x, "provides_norm", False) and hasattr(self.optimizer, "fused_step")
                                  and self.clip_grad_norm and self.clip_grad_norm > 0)
                if norm_ready:
                    x.finish(gnorm_sq_out=self.optimizer.gnorm_sq)
                else:
                    x.finish()
            else:
                allreduce_gradients(self.unet)
        if self.cuda_graph and hasattr(self.optimizer, "fused_step"):
            if self._opt_graph is None or self._opt_graph.gnorm_ready != norm_ready:
                self._opt_graph = GraphedOptimizerStep(self.optimizer, self.clip_grad_norm, 1.0 / self.world_size,
                                                       gnorm_ready=norm_ready).capture()
            self._opt_graph.replay()
            return
        if isinstance(self.optimizer, B200AdamWBF16):
            self.optimizer.fused_step(max_norm=self.clip_grad_norm, grad_scale=1.0 / self.world_size, zero_grad=True,
                                      gnorm_ready=norm_ready)
            return
        if hasattr(self.optimizer, "fused_step"):
            self.optimizer.fused_step(max_norm=self.clip_grad_norm, grad_scale=1.0 / self.world_size, gnorm_ready=norm_ready)
        else:  # a reference optimizer reading p.grad (adamw_bfloat16/__init__.py:92-119)
            if self.world_size > 1:
                self.unet.store.grad.mul_(1.0 / self.world_size)
            if self.clip_grad_norm and self.clip_grad_norm > 0:
                torch.nn.utils.clip_grad_norm_(list(self.unet.parameters()), self.clip_grad_norm)
            self.optimizer.step()
        self.optimizer.zero_grad()

    def train(self, num_epochs: int):
        """Minimal loop with the reference's protocol; logging/checkpoint plumbing is out of scope."""
        N = self.gradient_accumulation_steps
        history = []
        self.optimizer.zero_grad()
        for epoch in range(num_epochs):
            for step, batch in enumerate(self.train_dataloader):
                t0 = time.time()
                try:
                    loss, metrics = self._execute_training_step(batch, accumulate=N > 1,
                                                                is_last_accumulation_step=(step + 1) % N == 0)
                except FatalStepError:
                    raise
                except Exception:
                    if isinstance(self, B200FlowMatchingTrainer):
                        raise  # flow loop re-raises (flow_matching_trainer.py:226-228)
                    continue   # ddpm loop logs and continues (ddpm_trainer.py:202-204)
                metrics["step_time"] = time.time() - t0
                history.append(metrics)
                if self.wandb_logger is not None and hasattr(self.wandb_logger, "log_metrics"):
                    self.wandb_logger.log_metrics(metrics)
        return history


class B200DDPMTrainer(_StepBase):
    """`training.method: ddpm` — replaces DDPMTrainer (ddpm_trainer.py:26)."""
    name = "ddpm"

    def __init__(self, model, optimizer, train_dataloader=None, device=None, wandb_logger=None, config=None, **kwargs):
        super().__init__(model, optimizer, train_dataloader, device, wandb_logger, config, **kwargs)
        self.noise_scheduler = NoiseScheduler(config, "cpu")
        m = getattr(config, "model", None)
        self.min_snr_gamma = getattr(m, "min_snr_gamma", None)
        self.core = FusedLossCore(self.unet, "ddpm", self.prediction_type, self.noise_scheduler.use_ztsnr,
                                  seed=int(kwargs.get("seed", 0)))
        self._init_dp()

    def training_step(self, batch: Dict[str, Any], noise: Optional[torch.Tensor] = None,
                      timesteps: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        dev = self.device
        t = _prep_batch(batch, dev)
        B = t["latents"].shape[0]
        if timesteps is None:
            timesteps = self.noise_scheduler.sample_timesteps(B, device=dev)
        timesteps = timesteps.cpu().long()
        sig = self.noise_scheduler.timestep_to_sigma(timesteps).float()
        weight = None
        if self.min_snr_gamma is not None:  # B3: intended per-sample broadcast
            snr = (self.noise_scheduler.sigma_data / sig) ** 2
            weight = torch.minimum(snr, torch.ones_like(snr) * float(self.min_snr_gamma)).float()  # host; staged below
        tw = _tag_weight_mean(batch)
        if self.cuda_graph and noise is None and tw is None:
            loss = self._graphed_loss(t, timesteps.float(), sig, weight)
        else:
            loss = self.core.loss_fn(latents=t["latents"], ctx=t["ctx"], pooled=t["pooled"], time_ids=t["time_ids"],
                                     t_embed=timesteps.float().to(dev), sig_or_t=sig.to(dev),
                                     weight=None if weight is None else weight.to(dev),
                                     loss_scale=1.0 if tw is None else tw,
                                     noise=None if noise is None else noise.to(dev).to(bf16).float().reshape(-1).contiguous())
        loss_f, st = self._read_loss_and_stats(loss)  # one D2H for all metrics (the reference does 5 .item() syncs)
        n = self.core.last["numel"]
        metrics = {
            "loss": loss_f,
            "lr": self._lr(),
            "timestep_mean": float(timesteps.float().mean()),
            "noise_scale": st[0] / n,
            "pred_scale": st[2] / n,
            "batch_size": B,
        }
        if B > 1:
            metrics["timestep_std"] = float(timesteps.float().std())
        return {"loss": loss, "metrics": metrics}

    compute_loss = training_step


class B200FlowMatchingTrainer(_StepBase):
    """`training.method: flow_matching` — replaces FlowMatchingTrainer (flow_matching_trainer.py:23)."""
    name = "flow_matching"

    def __init__(self, model, optimizer, train_dataloader=None, device=None, wandb_logger=None, config=None, **kwargs):
        super().__init__(model, optimizer, train_dataloader, device, wandb_logger, config, **kwargs)
        self.core = FusedLossCore(self.unet, "flow_matching", seed=int(kwargs.get("seed", 0)))
        self._init_dp()

    def compute_loss(self, model, batch: Dict[str, Any], generator: Optional[torch.Generator] = None,
                     x0: Optional[torch.Tensor] = None, t: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        dev = self.device
        tb = _prep_batch(batch, dev)
        B = tb["latents"].shape[0]
        if t is None:
            t = sample_logit_normal((B,), bf16, generator=generator)  # bf16 draw (B20)
        t = t.to(bf16).cpu()
        weight = None
        loss_scale = 1.0
        if "tag_weights" in batch:  # flow_matching_trainer.py:326-328
            loss_scale = float(batch["tag_weights"].to(bf16).float().mean())
        if self.cuda_graph and x0 is None and loss_scale == 1.0:
            th = t.float()  # host tensor: the graph path stages it through pinned memory (no stream-synchronising copy)
            loss = self._graphed_loss(tb, th, th, weight)
        else:
            tf = t.float().to(dev)
            loss = self.core.loss_fn(latents=tb["latents"], ctx=tb["ctx"], pooled=tb["pooled"], time_ids=tb["time_ids"],
                                     t_embed=tf, sig_or_t=tf, weight=weight, loss_scale=loss_scale,
                                     noise=None if x0 is None else x0.to(dev).to(bf16).float().reshape(-1).contiguous())
        loss_f, st = self._read_loss_and_stats(loss)
        metrics = {
            "loss": loss_f,
            "x0_norm": math.sqrt(st[1]),
            "x1_norm": math.sqrt(st[5]),
            "time_mean": float(t.float().mean()),
            "time_std": float(t.float().std()) if B > 1 else 0.0,
            "velocity_norm": math.sqrt(st[3]),  # B7: from the single forward's prediction
            "batch_size": B,
            "lr": self._lr(),
        }
        return {"loss": loss, "metrics": metrics}


TRAINER_MAP = {"ddpm": B200DDPMTrainer, "flow_matching": B200FlowMatchingTrainer}


def create_trainer(config, model, optimizer, train_dataloader=None, device=None, wandb_logger=None, **kw):
    """Dispatch on `config.training.method` like SDXLTrainer.__init__ (sdxl_trainer.py:128-152)."""
    method = str(getattr(getattr(config, "training", None), "method", "ddpm")).lower()
    if method not in TRAINER_MAP:
        raise ValueError(f"Unsupported training method: {method}")
    return TRAINER_MAP[method](model, optimizer, train_dataloader, device, wandb_logger, config, **kw)


# ----------------------------------------------------------------------------------------------------------------
# optimizer + data parallel
# ----------------------------------------------------------------------------------------------------------------
class B200AdamW:
    """Fused AdamW over the flat buffers (one launch for 2.57 B parameters), optional fp32 master weights.
    Exposes the torch-optimizer surface the loops touch: step / zero_grad / param_groups / state_dict."""

    def __init__(self, unet: B200UNet, lr=4e-7, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, master_weights=True):
        self.unet = unet
        st = unet.store
        dev = st.flat.device
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, params=list(unet.parameters()))]
        self.m = torch.zeros(st.total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(st.total, device=dev, dtype=torch.float32)
        self.master = st.flat.float() if master_weights else None
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.steps = 0
        self.step_ctr = torch.zeros(2, device=dev, dtype=torch.int64)  # [1] = step, advanced on the device (graph-safe)

    def sync_master(self):
        if self.master is not None:
            self.master.copy_(self.unet.store.flat.float())

    def fused_step(self, max_norm: float = 0.0, grad_scale: float = 1.0, _in_graph: bool = False, gnorm_ready: bool = False):
        st = self.unet.store
        g = self.param_groups[0]
        gn = None
        if max_norm and max_norm > 0:
            if not gnorm_ready:  # else: the data-parallel exchange already wrote sum(grad^2) into gnorm_sq
                self.gnorm_sq.zero_()
                ops.sumsq(st.grad, self.gnorm_sq)
            gn = self.gnorm_sq
        ops.philox_advance(self.step_ctr, 1)
        ops.adamw(st.flat, self.master, st.grad, self.m, self.v, lr=g["lr"], beta1=g["betas"][0], beta2=g["betas"][1],
                  eps=g["eps"], weight_decay=g["weight_decay"], step=0, gnorm_sq=gn, max_norm=max_norm or 0.0,
                  grad_scale=grad_scale, dev_step=self.step_ctr[1:])
        if not _in_graph:
            self.after_graph_step()

    def after_graph_step(self):
        self.steps += 1

    def step(self):
        self.fused_step()

    def zero_grad(self, set_to_none: bool = False):
        self.unet.store.grad.zero_()

    def state_dict(self):
        return {"m": self.m, "v": self.v, "master": self.master, "steps": self.steps,
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}


class B200AdamWBF16:
    """`optimizer_type: adamw_bf16` — the reference's default optimizer (src/config.yaml;
    src/training/optimizers/adamw_bfloat16/__init__.py:26-148) on the flat buffers: bf16 exp_avg / exp_avg_sq / shift,
    stochastic rounding, deferred thresholded weight decay.  One fused launch per step (b2_adamw_bf16, 18 B/parameter)
    instead of ~20 eager kernels for each of the 1,680 tensors; the per-tensor `accumulated_decay` bookkeeping
    (:118-128) stays on the host exactly as written and fires a per-tensor b2_axpy_bf16 when its threshold trips.
    Same surface as the reference class: step(zero_grad=False) / zero_grad / state_dict / param_groups."""
    decay_threshold = 5e-3  # adamw_bfloat16/__init__.py:27

    def __init__(self, unet: B200UNet, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, seed: int = 0,
                 as_written: bool = True):
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        self.unet = unet
        st = unet.store
        dev = st.flat.device
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, params=list(unet.parameters()))]
        self.exp_avg = torch.zeros(st.total, device=dev, dtype=bf16)
        self.exp_avg_sq = torch.zeros(st.total, device=dev, dtype=bf16)
        self.shift = torch.zeros(st.total, device=dev, dtype=bf16)
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.seed_offset = torch.tensor([seed, 0], device=dev, dtype=torch.int64)
        self.as_written = as_written
        self.steps = 0
        g = torch.Generator().manual_seed(seed)
        # each tensor starts its decay accumulator at a random phase (:110-116)
        self.accumulated_decay = {name: float(torch.rand([], generator=g)) * self.decay_threshold for name, _ in st.specs}
        self._min_headroom = 0.0  # steps can skip the per-tensor scan while no accumulator can reach the threshold

    def fused_step(self, max_norm: float = 0.0, grad_scale: float = 1.0, rng_mode: int = 0, test_rand16=None,
                   _in_graph: bool = False, zero_grad: bool = False, gnorm_ready: bool = False):
        st = self.unet.store
        g = self.param_groups[0]
        gn = None
        if max_norm and max_norm > 0:
            if not gnorm_ready:  # else: the data-parallel exchange already wrote sum(grad^2) into gnorm_sq
                self.gnorm_sq.zero_()
                ops.sumsq(st.grad, self.gnorm_sq)
            gn = self.gnorm_sq
        ops.philox_advance(self.seed_offset, 1)  # seed_offset[1] = optimizer step, kept on the device (graph-safe)
        ops.adamw_bf16(st.flat, st.grad, self.exp_avg, self.exp_avg_sq, self.shift, lr=g["lr"], beta1=g["betas"][0],
                       beta2=g["betas"][1], eps=g["eps"], step=0, gnorm_sq=gn, max_norm=max_norm or 0.0,
                       grad_scale=grad_scale, seed_offset=self.seed_offset, as_written=self.as_written,
                       rng_mode=rng_mode, test_rand16=test_rand16, zero_grad=zero_grad)
        if not _in_graph:
            self.after_graph_step()

    def after_graph_step(self):
        """Host-side part of a step: the per-tensor deferred weight decay (adamw_bfloat16/__init__.py:118-128, 191-192)."""
        st = self.unet.store
        g = self.param_groups[0]
        self.steps += 1
        inc = g["weight_decay"] * g["lr"]
        if inc > 0:
            self._min_headroom -= inc
            if self._min_headroom > 0:  # nothing can trip yet: defer the per-tensor adds (same trip step up to host float rounding)
                self._lazy_inc = getattr(self, "_lazy_inc", 0.0) + inc
                return
            inc += getattr(self, "_lazy_inc", 0.0)
            self._lazy_inc = 0.0
            for name, _ in st.specs:
                acc = self.accumulated_decay[name] + inc
                if acc > self.decay_threshold:
                    off, numel = st.offsets[name], st._numel[name]
                    # torch's bf16 add_ rounds alpha to the tensor dtype on CPU (oracle/adamw_bf16.py:apply_decay)
                    alpha = float(torch.tensor(-acc, dtype=torch.float32).to(bf16))
                    ops.axpy_bf16(self.shift[off:off + numel], st.flat[off:off + numel], alpha)
                    acc = 0.0
                self.accumulated_decay[name] = acc
            self._min_headroom = self.decay_threshold - max(self.accumulated_decay.values())

    def step(self, zero_grad: bool = False):
        self.fused_step(zero_grad=zero_grad)  # the kernel clears the gradient vector behind its own read

    def zero_grad(self, set_to_none: bool = False):
        self.unet.store.grad.zero_()

    def state_dict(self):
        st = self.unet.store
        state = {}
        for name, _ in st.specs:
            off, numel = st.offsets[name], st._numel[name]
            state[name] = {"step": float(self.steps), "exp_avg": self.exp_avg[off:off + numel],
                           "exp_avg_sq": self.exp_avg_sq[off:off + numel], "shift": self.shift[off:off + numel],
                           "accumulated_decay": self.accumulated_decay[name] + getattr(self, "_lazy_inc", 0.0)}
        return {"state": state, "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}


def allreduce_gradients(unet: B200UNet):
    """ONE collective per optimizer step: sum of the flat bf16 gradient buffer over NVLink."""
    if torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(unet.store.grad, op=torch.distributed.ReduceOp.SUM)
