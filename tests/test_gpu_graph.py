"""CUDA-graph replay of the micro-step and of the optimizer step must reproduce the eager (kernel-by-kernel) path:
same Philox counters, same kernels, same order.  Split-K reduce-adds make the last bf16 bit of some gradients
order-dependent, so gradients / parameters are compared with a 1e-2 relative-L2 bound and the loss with 1e-3."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _conf(method, accum=1):
    return SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0, use_ztsnr=True,
                                                 min_snr_gamma=None),
                           training=SimpleNamespace(method=method, prediction_type="v_prediction",
                                                    gradient_accumulation_steps=accum, clip_grad_norm=1.0))


def _batch(cfg, B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return {"vae_latents": torch.randn(B, 4, H, W, generator=g),
            "prompt_embeds": torch.randn(B, 77, cfg["cross_attention_dim"], generator=g),
            "pooled_prompt_embeds": torch.randn(B, 96, generator=g),
            "time_ids": torch.tensor([[8. * W, 8. * H, 0., 0., 8. * W, 8. * H]]).repeat(B, 1)[:, None],
            "metadata": [{} for _ in range(B)]}


def _rel(a, b):
    a = a.float().flatten(); b = b.float().flatten()
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.mark.parametrize("method,optname", [("ddpm", "fp32"), ("flow_matching", "bf16")])
def test_graph_replay_matches_eager(method, optname):
    from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
    from sdxl_training_improvements_b200.trainer import B200AdamW, B200AdamWBF16, create_trainer
    from sdxl_training_improvements_b200.unet import B200UNet
    cfg = tiny_config()
    sd = seeded_init_(OracleUNet(cfg), 3).state_dict()
    nets, opts, trs = [], [], []
    for graph in (True, False):
        net = B200UNet(cfg, device="cuda")
        net.load_state_dict(sd)
        opt = B200AdamW(net, lr=1e-3) if optname == "fp32" else B200AdamWBF16(net, lr=1e-3, seed=9)
        nets.append(net); opts.append(opt)
        trs.append(create_trainer(_conf(method, accum=2), net, opt, device="cuda", seed=5, cuda_graph=graph))
    tg, te = trs
    B, H, W = 2, 16, 16
    # first optimizer step on the graph trainer only: captures both graphs (2 warm-up passes advance the noise counter)
    for k in range(2):
        tg._execute_training_step(_batch(cfg, B, H, W, 100 + k), accumulate=True, is_last_accumulation_step=k == 1)
    assert tg._micro_graphs and tg._opt_graph is not None
    # bring the eager twin to the same state
    nets[1].store.flat.copy_(nets[0].store.flat)
    te.core.seed_offset.copy_(tg.core.seed_offset)
    if optname == "fp32":
        for n in ("m", "v", "master", "step_ctr"):
            getattr(opts[1], n).copy_(getattr(opts[0], n))
    else:
        for n in ("exp_avg", "exp_avg_sq", "shift", "seed_offset"):
            getattr(opts[1], n).copy_(getattr(opts[0], n))
    opts[1].steps = opts[0].steps
    # one more optimizer step (2 accumulation micro-steps) on both, same data and timesteps
    before = opts[0].master.clone() if optname == "fp32" else nets[0].store.flat.float().clone()
    gen = torch.Generator().manual_seed(1)
    tg._pending_grad_scale = 0.5  # what _execute_training_step(accumulate=True) announces before compute_loss
    for k in range(2):
        b = _batch(cfg, B, H, W, 200 + k)
        if method == "ddpm":
            ts = torch.randint(0, 1000, (B,), generator=gen)
            og = tg.training_step(b, timesteps=ts); oe = te.training_step(b, timesteps=ts)
        else:
            t = torch.sigmoid(torch.randn(B, generator=gen)).to(bf16)
            og = tg.compute_loss(None, b, t=t); oe = te.compute_loss(None, b, t=t)
        lg, le = float(og["loss"]), float(oe["loss"])
        assert abs(lg - le) <= 1e-3 * max(1.0, abs(le)), (k, lg, le)
        assert set(og["metrics"]) == set(oe["metrics"])
        (oe["loss"] / 2).backward()   # eager: autograd runs the hand-written backward
        (og["loss"] / 2).backward()   # graph: no-op, the replay already accumulated 0.5 * dL/dw
    gg, ge = nets[0].store.grad, nets[1].store.grad
    assert float(ge.float().norm()) > 0
    assert _rel(gg, ge) <= 1e-2, _rel(gg, ge)
    tg.optimizer_step(); te.optimizer_step()
    torch.cuda.synchronize()
    assert float(nets[0].store.grad.float().abs().max()) == 0.0 and float(nets[1].store.grad.float().abs().max()) == 0.0
    assert _rel(nets[0].store.flat, nets[1].store.flat) <= 1e-3
    if optname == "fp32":
        assert _rel(opts[0].master - before, opts[1].master - before) <= 0.2
    assert opts[0].steps == opts[1].steps == 2
