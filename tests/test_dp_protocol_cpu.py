"""Data-parallel step protocol of the trainers on CPU: WHEN the gradient exchange is driven.

The trainer, the engine's tape and the chunk plan are the real code; the kernels are replaced by shape-only stand-ins and
the transport by a recorder, so this runs without a GPU.  Pins:
  * first optimizer step: the backward pass is logged, the plan is built from it, the whole buffer is exchanged once;
  * later steps: chunk k is flushed and exchanged right after tape position plan.cuts[k], in order, and only in the LAST
    accumulation micro-step; `finish()` precedes the optimizer update exactly once per optimizer step;
  * a caller that never announces the last micro-step (compute_loss + backward + optimizer_step by hand) still gets one
    whole-buffer exchange before the update;
  * the 1/world factor goes to the optimizer's grad scale, not into the exchanged gradients.
The transport itself: tests/test_gpu_dp_exchange.py (needs >= 2 GPUs).  Reference behaviour replaced: DDP's reducer firing
during loss.backward() (src/core/distributed.py:142-163; ddpm_trainer.py:256-278 for the accumulate / step protocol).
"""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.unet_sdxl import tiny_config  # noqa: E402
from sdxl_training_improvements_b200 import trainer as trainer_mod, unet as unet_mod  # noqa: E402
from _shape_only_ops import ShapeOnlyOps  # noqa: E402


class TrainerOps(ShapeOnlyOps):
    """+ the loss-side entry points the trainers call."""

    def randn(self, n, seed_offset, stream_id, round_bf16=True, out=None):
        return torch.zeros(n, dtype=torch.float32)

    def make_noisy(self, x, eps, sig, mode, vpred, clamp, B, Cc, HW, Cpad):
        return torch.zeros(B * HW, Cpad, dtype=torch.bfloat16), torch.zeros(B, Cc, HW)

    def nchw_to_nhwc(self, x, Cpad):
        B, Cc, H, W = x.shape
        return torch.zeros(B * H * W, Cpad, dtype=torch.bfloat16)


class RecordingExchange:
    """Stands in for dp.PeerGradExchange: same surface, records the calls."""

    def __init__(self):
        self.plan = None
        self.issued = False
        self.calls = []

    def set_plan(self, plan):
        self.plan = plan
        self.calls.append(("plan", plan.n_chunks))

    def flush_chunk(self, store, k):
        self.calls.append(("flush", k))

    def exchange_chunk(self, k):
        self.calls.append(("chunk", k))
        self.issued = True

    def exchange_all(self):
        self.calls.append(("all",))
        self.issued = True

    provides_norm = False  # set True by the test of the norm hand-off

    def finish(self, gnorm_sq_out=None):
        self.calls.append(("finish",) if gnorm_sq_out is None else ("finish+norm", gnorm_sq_out))
        self.issued = False


@pytest.fixture()
def rig(monkeypatch):
    fake = TrainerOps()
    monkeypatch.setattr(unet_mod, "ops", fake)
    monkeypatch.setattr(trainer_mod, "ops", fake)
    cfg = tiny_config()
    net = unet_mod.B200UNet(cfg, device="cpu")
    net.store.flush_small_grads = lambda: None
    opt = trainer_mod.B200AdamW(net, lr=1e-3)
    scales = []
    real_fused = opt.fused_step
    opt.fused_step = lambda **kw: (scales.append(kw.get("grad_scale")),
                                   x.calls.append(("update",) if not kw.get("gnorm_ready") else ("update", "norm from exchange")))[0]
    conf = SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0, use_ztsnr=True,
                                                 min_snr_gamma=None),
                           training=SimpleNamespace(method="ddpm", prediction_type="v_prediction",
                                                    gradient_accumulation_steps=2, clip_grad_norm=1.0))
    tr = trainer_mod.B200DDPMTrainer(net, opt, None, "cpu", config=conf)
    x = RecordingExchange()
    tr.world_size = 2
    tr.core.dp = x
    spans = []
    eng = net.engine
    real_span = eng.backward_span
    eng.backward_span = lambda order, lo, hi: (spans.append((lo, hi)), real_span(order, lo, hi))[1]
    B, H, W = 2, 16, 16
    batch = {"vae_latents": torch.randn(B, 4, H, W), "prompt_embeds": torch.randn(B, 77, cfg["cross_attention_dim"]),
             "pooled_prompt_embeds": torch.randn(B, 96),
             "time_ids": torch.tensor([[128., 128., 0., 0., 128., 128.]]).repeat(B, 1)[:, None],
             "metadata": [{} for _ in range(B)]}
    del real_fused
    return SimpleNamespace(tr=tr, x=x, spans=spans, scales=scales, batch=batch)


def test_exchange_is_driven_by_the_backward_tape(rig):
    tr, x, batch = rig.tr, rig.x, rig.batch
    # ---- optimizer step 1: no plan yet -> logged pass, whole-buffer exchange, then the update
    tr._execute_training_step(batch)
    assert x.calls[0][0] == "plan" and x.calls[0][1] >= 2
    assert x.calls[1:] == [("all",), ("finish",), ("update",)]
    plan = x.plan
    K = plan.n_chunks
    # ---- optimizer step 2, two accumulation micro-steps: only the last one exchanges, chunk by chunk at the cuts
    x.calls.clear()
    rig.spans.clear()
    tr._execute_training_step(batch, accumulate=True, is_last_accumulation_step=False)
    assert x.calls == [], "a non-final micro-step must not touch the exchange"
    assert rig.spans == [(0, plan.n_tape)]
    rig.spans.clear()
    tr._execute_training_step(batch, accumulate=True, is_last_accumulation_step=True)
    want = [c for k in range(K) for c in (("flush", k), ("chunk", k))] + [("finish",), ("update",)]
    assert x.calls == want
    cuts = plan.cuts
    assert rig.spans == [(0 if k == 0 else cuts[k - 1] + 1, cuts[k] + 1) for k in range(K)]
    assert rig.spans[-1][1] == plan.n_tape
    assert tr.core.dp_last is False, "the announcement must not leak into the next micro-step"
    # the mean over ranks is the optimizer's business: grad_scale = 1 / world, the exchanged buffer holds the SUM
    assert rig.scales and all(s == 0.5 for s in rig.scales)


def test_unannounced_step_still_exchanges_once(rig):
    tr, x, batch = rig.tr, rig.x, rig.batch
    tr._execute_training_step(batch)  # builds the plan
    x.calls.clear()
    out = tr.compute_loss(batch)      # a loop that drives backward / step itself (README plugin contract)
    out["loss"].backward()
    assert x.calls == [], "without the announcement nothing may be exchanged during backward"
    tr.optimizer_step()
    assert x.calls == [("all",), ("finish",), ("update",)]


def test_exchange_hands_the_gradient_norm_to_the_fused_optimizer(rig):
    """Copy-engine transports sum (reduced gradient)^2 inside their reduce kernels: finish() must then be given the optimizer's
    gnorm_sq buffer and the fused step must be told to skip its own pass — and neither when clipping is off or the transport
    cannot provide it (clip_grad_norm_, flow_matching_trainer.py:181-186, needs the norm of the REDUCED gradients)."""
    tr, x, batch = rig.tr, rig.x, rig.batch
    x.provides_norm = True
    tr._execute_training_step(batch)
    assert x.calls[-2][0] == "finish+norm" and x.calls[-2][1] is tr.optimizer.gnorm_sq
    assert x.calls[-1] == ("update", "norm from exchange")
    x.calls.clear()
    tr.clip_grad_norm = 0.0           # no clipping: nobody needs a norm
    tr._execute_training_step(batch)
    assert x.calls[-2:] == [("finish",), ("update",)]
    x.calls.clear()
    tr.clip_grad_norm = 1.0
    x.provides_norm = False           # e.g. the "sm" transport
    tr._execute_training_step(batch)
    assert x.calls[-2:] == [("finish",), ("update",)]
