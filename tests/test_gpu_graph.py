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


def test_new_shape_mid_accumulation_runs_uncaptured_then_captures_at_the_next_boundary():
    """ADVICE r1 (trainer.py:436): aspect buckets + gradient_accumulation_steps > 1 — a latent shape first seen on the second
    micro-step of an accumulation window cannot be captured (gradients are not zero).  It must run through the same
    kernels un-captured (no exception, nothing dropped, accumulation phase intact) and be captured the next time it shows
    up at a boundary; all captured shapes share ONE graph memory pool; an lr change re-captures the optimizer graph."""
    from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
    from sdxl_training_improvements_b200.trainer import B200AdamW, create_trainer
    from sdxl_training_improvements_b200.unet import B200UNet
    cfg = tiny_config()
    sd = seeded_init_(OracleUNet(cfg), 3).state_dict()
    nets, trs = [], []
    for graph in (True, False):
        net = B200UNet(cfg, device="cuda")
        net.load_state_dict(sd)
        nets.append(net)
        trs.append(create_trainer(_conf("flow_matching", accum=2), net, B200AdamW(net, lr=1e-3), device="cuda", seed=5,
                                  cuda_graph=graph))
    tg, te = trs
    gen = torch.Generator().manual_seed(4)
    shapes = [(16, 16), (16, 24), (16, 24), (16, 16), (24, 16), (16, 24)]   # window 1: A then NEW B; window 2: B, A; ...
    for k, (H, W) in enumerate(shapes):
        b = _batch(cfg, 2, H, W, 300 + k)
        t = torch.sigmoid(torch.randn(2, generator=gen)).to(bf16)
        last = k % 2 == 1
        tg._pending_grad_scale = 0.5   # what _execute_training_step(accumulate=True) announces before compute_loss
        og = tg.compute_loss(None, b, t=t)
        # the eager twin must see the graph trainer's in-graph / un-captured noise draw: same Philox counter
        te.core.seed_offset.copy_(tg.core.seed_offset - torch.tensor([0, 1], device="cuda"))
        oe = te.compute_loss(None, b, t=t)
        (og["loss"] / 2).backward(); (oe["loss"] / 2).backward()
        assert abs(float(og["loss"]) - float(oe["loss"])) <= 1e-3 * max(1.0, abs(float(oe["loss"]))), k
        if last:
            assert _rel(nets[0].store.grad, nets[1].store.grad) <= 1e-2, k
            if k == 3:
                tg.optimizer.param_groups[0]["lr"] = 5e-4   # a scheduler changes lr: the optimizer graph must follow
                te.optimizer.param_groups[0]["lr"] = 5e-4
            tg.optimizer_step(); te.optimizer_step()
            assert _rel(nets[0].store.flat, nets[1].store.flat) <= 1e-3, k
    # (16,24) first arrived mid-window -> un-captured once, captured at k=2 (a boundary); (24,16) arrived at a boundary
    assert tg.uncaptured_micro_steps == 1
    assert set(tg._micro_graphs) == {(2, 16, 16, 77), (2, 16, 24, 77), (2, 24, 16, 77)}
    pools = {tuple(g.graph.pool()) for g in tg._micro_graphs.values()}
    assert len(pools) == 1, "captured shapes must share one graph memory pool"
    tg.max_graphed_shapes = 3
    b = _batch(cfg, 2, 24, 24, 999)
    tg._pending_grad_scale = 1.0
    tg.compute_loss(None, b, t=torch.tensor([0.3, 0.6]).to(bf16))["loss"].backward()
    assert (2, 24, 24, 77) not in tg._micro_graphs and tg.uncaptured_micro_steps == 2   # cap reached: runs un-captured
