"""T5: 100 optimizer steps on identical seeds / weights / noise / timesteps — loss curve of the kernel
path (bf16 activations, fp32 accumulation, fused loss + clip + AdamW with fp32 master weights) vs the oracle
(fp32 PyTorch on CUDA, torch.optim.AdamW, clip_grad_norm_ 1.0).

Three arms, same seeds / initial weights / noise / timesteps / optimizer hyper-parameters:
  fp32   : oracle in fp32 (the "exact" curve),
  bf16   : oracle in CUDA-eager bf16 with fp32 master weights — the numerics of the reference path itself, which casts the
           whole UNet to bf16 (sdxl_trainer.py:52-55),
  kernel : this repo (bf16 activations, fp32 accumulation in TMEM, fp32 master weights).
Tolerance (stated): the north star asks for 1e-3 on the loss curve; SURVEY.md B23 notes that this is below the bf16
path's own resolution, so the bar here is   max_t |kernel - fp32| <= max(1e-3, 1.5 * max_t |bf16 - fp32|)   i.e. the
kernel path must track the exact curve at least as well as the reference's own precision does, and the first step
(pure forward parity, no optimizer drift) must be within 1e-3 * max(1, loss).
Reduced-width UNet (same topology) so that the oracles run in seconds; flow matching and ddpm/v-prediction.
"""
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
STEPS = 100


def _setup(seed=0):
    from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
    from sdxl_training_improvements_b200.unet import B200UNet
    cfg = tiny_config()
    ref = seeded_init_(OracleUNet(cfg), seed).cuda()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(p.to(bf16).float())
    net = B200UNet(cfg, device="cuda")
    net.load_state_dict(ref.state_dict())
    return cfg, ref, net


def _batches(cfg, B, H, W, n, seed):
    g = torch.Generator().manual_seed(seed)
    pooled_dim = cfg["projection_class_embeddings_input_dim"] - 6 * cfg["addition_time_embed_dim"]
    out = []
    for _ in range(n):
        out.append({
            "vae_latents": torch.randn(B, 4, H, W, generator=g).to(bf16).float(),
            "prompt_embeds": torch.randn(B, 77, cfg["cross_attention_dim"], generator=g).to(bf16).float(),
            "pooled_prompt_embeds": torch.randn(B, pooled_dim, generator=g).to(bf16).float(),
            "time_ids": torch.tensor([[8. * W, 8. * H, 0., 0., 8. * W, 8. * H]]).repeat(B, 1)[:, None],
            "metadata": [{} for _ in range(B)],
            "noise": torch.randn(B, 4, H, W, generator=g).to(bf16).float(),
            "t_ddpm": torch.randint(300, 800, (B,), generator=g),          # sigma in [0.6, 1800]: loss not clamped
            "t_flow": torch.sigmoid(torch.randn(B, generator=g)).to(bf16),
        })
    return out


def _conf(method):
    return SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0, use_ztsnr=True,
                                                 min_snr_gamma=None),
                           training=SimpleNamespace(method=method, prediction_type="v_prediction",
                                                    gradient_accumulation_steps=1, clip_grad_norm=1.0))


@pytest.mark.parametrize("method", ["flow_matching", "ddpm"])
def test_loss_curve_100_steps(method):
    from oracle import schedule as S
    from sdxl_training_improvements_b200.trainer import B200AdamW, create_trainer
    cfg, ref, net = _setup(seed=3)
    lr = 2e-4
    opt_k = B200AdamW(net, lr=lr, weight_decay=1e-2, master_weights=True)
    opt_o = torch.optim.AdamW(ref.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    tr = create_trainer(_conf(method), net, opt_k, device="cuda")
    import copy
    ref16 = copy.deepcopy(ref).to(bf16)                       # bf16 eager arm, fp32 master = `m16`
    m16 = [p.detach().float().clone().requires_grad_(True) for p in ref16.parameters()]
    opt_16 = torch.optim.AdamW(m16, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    l16 = []
    sig = S.schedule_sigmas().cuda()
    B, H, W = 2, 16, 16
    lk, lo = [], []
    for b in _batches(cfg, B, H, W, STEPS, seed=11):
        dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
        if method == "ddpm":
            out = tr.training_step(b, noise=b["noise"], timesteps=b["t_ddpm"])
            o = S.ddpm_loss(ref, dev["vae_latents"], dev["noise"], dev["t_ddpm"], dev["prompt_embeds"],
                            dev["pooled_prompt_embeds"], dev["time_ids"], sigmas=sig)
        else:
            out = tr.compute_loss(net, b, x0=b["noise"], t=b["t_flow"])
            o = S.flow_loss(ref, dev["vae_latents"], dev["noise"], dev["t_flow"].float(), dev["prompt_embeds"],
                            dev["pooled_prompt_embeds"], dev["time_ids"])
        out["loss"].backward()
        tr.optimizer_step()
        opt_o.zero_grad(set_to_none=True)
        o["loss"].backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt_o.step()
        c16 = {k: (v.to(bf16) if torch.is_tensor(v) and v.is_floating_point() and k != "time_ids" else v)
               for k, v in dev.items()}
        if method == "ddpm":
            o16 = S.ddpm_loss(ref16, c16["vae_latents"], c16["noise"], dev["t_ddpm"], c16["prompt_embeds"],
                              c16["pooled_prompt_embeds"], dev["time_ids"], sigmas=sig)
        else:
            o16 = S.flow_loss(ref16, c16["vae_latents"], c16["noise"], c16["t_flow"], c16["prompt_embeds"],
                              c16["pooled_prompt_embeds"], dev["time_ids"])
        for p in ref16.parameters():
            p.grad = None
        o16["loss"].backward()
        for mp_, p in zip(m16, ref16.parameters()):
            mp_.grad = p.grad.float()
        torch.nn.utils.clip_grad_norm_(m16, 1.0)
        opt_16.step()
        with torch.no_grad():
            for mp_, p in zip(m16, ref16.parameters()):
                p.copy_(mp_.to(bf16))
        l16.append(float(o16["loss"].detach()))
        lk.append(float(out["loss"].detach()))
        lo.append(float(o["loss"].detach()))
    d = [abs(a - b) for a, b in zip(lk, lo)]
    worst = max(range(STEPS), key=lambda i: d[i])
    print(f"\n{method}: loss[0] kernel {lk[0]:.6f} oracle {lo[0]:.6f}; loss[99] kernel {lk[-1]:.6f} oracle {lo[-1]:.6f}; "
          f"max |d| {d[worst]:.3e} at step {worst} (loss {lo[worst]:.4f}); mean |d| {sum(d) / STEPS:.3e}")
    d16 = [abs(a - b) for a, b in zip(l16, lo)]
    print(f"{method}: bf16-eager oracle vs fp32 oracle: max |d| {max(d16):.3e}, mean |d| {sum(d16) / STEPS:.3e}")
    assert lo[-1] < lo[0], "the oracle's loss did not go down: the trajectory test is not exercising learning"
    assert d[0] <= 1e-3 * max(1.0, lo[0]), f"first-step loss differs: {d[0]:.3e}"
    bar = max(1e-3, 1.5 * max(d16))
    assert max(d) <= bar, f"loss curves diverge: max |d| {max(d):.3e} > {bar:.3e}"


# ---------------------------------------------------------------------------------------------------------------------
# Mid-size trajectory: SDXL's real widths 320 / 640 / 1280 (heads 5 / 10 / 20, cross dim 2048) with
# transformer depth 0 / 1 / 2, latent 32x32, B=2 — ~0.8 B parameters, every production kernel path at its real channel
# counts (implicit-GEMM convs, CTA-pair GEMMs with split-K, wide tiles, attention at n = 256 / 64, grouped K/V).
# Both optimizers: fp32-master AdamW and the reference's default AdamWBF16 (the one bench.py times).
# Numbers are written to gpurun_out/r2_trajectory_<method>_<optimizer>.json (copied to profiles/r2_trajectory.json).
# ---------------------------------------------------------------------------------------------------------------------
def _mid_config():
    from oracle.unet_sdxl import tiny_config
    return tiny_config(block_out_channels=(320, 640, 1280), transformer_layers_per_block=(0, 1, 2), num_heads=(5, 10, 20),
                       cross_attention_dim=2048, addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816)


def _oracle_adamw_bf16_step(params, state, step, lr, wd_state, gen):
    """The reference's AdamWBF16 (oracle/adamw_bf16.py, pinned bit-exact to adamw_bfloat16/__init__.py) on bf16 params."""
    from oracle import adamw_bf16 as O
    for i, p in enumerate(params):
        g = p.grad.float()
        st = state.setdefault(i, {"m": torch.zeros_like(g), "v": torch.zeros_like(g), "s": torch.zeros_like(g)})
        r16 = torch.randint(0, 65536, (4,) + tuple(g.shape), device=g.device, generator=gen)
        p1, st["m"], st["v"], st["s"] = O.make_step(p.detach().float(), g, st["m"], st["v"], st["s"], beta1=0.9, beta2=0.999,
                                                      step=step, lr=lr, eps=1e-8, rand16=r16)
        with torch.no_grad():
            p.copy_(p1.to(bf16))


@pytest.mark.parametrize("optname", ["adamw_fp32_master", "adamw_bf16"])
@pytest.mark.parametrize("method", ["ddpm", "flow_matching"])
def test_mid_size_loss_curve_100_steps(method, optname):
    import copy
    import json
    from oracle import schedule as S
    from oracle.unet_sdxl import OracleUNet, seeded_init_
    from sdxl_training_improvements_b200.trainer import B200AdamW, B200AdamWBF16, create_trainer
    from sdxl_training_improvements_b200.unet import B200UNet
    cfg = _mid_config()
    ref = seeded_init_(OracleUNet(cfg), 3).cuda()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(p.to(bf16).float())
    host = B200UNet(cfg, device="cpu")
    host.load_state_dict({k: v.cpu() for k, v in ref.state_dict().items()})
    net = B200UNet(cfg, device="cuda")
    net.store.flat.copy_(host.store.flat)
    del host
    # lr: the reference ships 4e-7 (src/config.yaml:17).  At 1e-4 a random-init 0.85 B-parameter UNet on 10 batches is in a
    # chaotic regime where ANY two precisions diverge by 0.1-0.7 in loss within 100 steps (measured: bf16 eager vs fp32
    # 0.15 / 0.23 / 0.66, profiles/r2_trajectory_lr1e-4_chaotic.json) — the yardstick bar still holds there but the north
    # star's 1e-3 is meaningless.  5e-6 (12x the shipped value) keeps the curves smooth while the loss still goes down.
    lr = float(os.environ.get("B2_TRAJ_LR", "5e-6"))
    if optname == "adamw_fp32_master":
        opt_k = B200AdamW(net, lr=lr, weight_decay=1e-2, master_weights=True)
    else:
        opt_k = B200AdamWBF16(net, lr=lr, weight_decay=0.0, seed=21)
    tr = create_trainer(_conf(method), net, opt_k, device="cuda")
    opt_o = torch.optim.AdamW(ref.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8,
                              weight_decay=1e-2 if optname == "adamw_fp32_master" else 0.0)
    # yardstick arm: the reference's own numerics — the whole UNet cast to bf16 (sdxl_trainer.py:52-55), eager
    ref16 = copy.deepcopy(ref).to(bf16)
    p16 = list(ref16.parameters())
    m16 = [p.detach().float().clone().requires_grad_(True) for p in p16] if optname == "adamw_fp32_master" else None
    opt_16 = torch.optim.AdamW(m16, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2) if m16 else None
    st16, gen16 = {}, torch.Generator(device="cuda").manual_seed(5)
    sig = S.schedule_sigmas().cuda()
    B, H, W = 2, 32, 32
    batches = _batches(cfg, B, H, W, 10, seed=11)   # 10 distinct batches cycled 10 times: the curve must go DOWN
    lk, lo, l16 = [], [], []
    for step in range(STEPS):
        b = batches[step % len(batches)]
        dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
        c16 = {k: (v.to(bf16) if torch.is_tensor(v) and v.is_floating_point() and k != "time_ids" else v)
               for k, v in dev.items()}
        if method == "ddpm":
            out = tr.training_step(b, noise=b["noise"], timesteps=b["t_ddpm"])
            o = S.ddpm_loss(ref, dev["vae_latents"], dev["noise"], dev["t_ddpm"], dev["prompt_embeds"],
                            dev["pooled_prompt_embeds"], dev["time_ids"], sigmas=sig)
            o16 = S.ddpm_loss(ref16, c16["vae_latents"], c16["noise"], dev["t_ddpm"], c16["prompt_embeds"],
                              c16["pooled_prompt_embeds"], dev["time_ids"], sigmas=sig)
        else:
            out = tr.compute_loss(net, b, x0=b["noise"], t=b["t_flow"])
            o = S.flow_loss(ref, dev["vae_latents"], dev["noise"], dev["t_flow"].float(), dev["prompt_embeds"],
                            dev["pooled_prompt_embeds"], dev["time_ids"])
            o16 = S.flow_loss(ref16, c16["vae_latents"], c16["noise"], c16["t_flow"], c16["prompt_embeds"],
                              c16["pooled_prompt_embeds"], dev["time_ids"])
        out["loss"].backward()
        tr.optimizer_step()
        opt_o.zero_grad(set_to_none=True)
        o["loss"].backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt_o.step()
        for p in p16:
            p.grad = None
        o16["loss"].backward()
        if m16 is not None:
            for mp_, p in zip(m16, p16):
                mp_.grad = p.grad.float()
            torch.nn.utils.clip_grad_norm_(m16, 1.0)
            opt_16.step()
            with torch.no_grad():
                for mp_, p in zip(m16, p16):
                    p.copy_(mp_.to(bf16))
        else:
            torch.nn.utils.clip_grad_norm_(p16, 1.0)
            _oracle_adamw_bf16_step(p16, st16, step + 1, lr, None, gen16)
        lk.append(float(out["loss"].detach())); lo.append(float(o["loss"].detach())); l16.append(float(o16["loss"].detach()))
    d = [abs(a - b_) for a, b_ in zip(lk, lo)]
    d16 = [abs(a - b_) for a, b_ in zip(l16, lo)]
    # THE north-star comparison: the reference casts the whole UNet to bf16 (sdxl_trainer.py:52-55), so "the reference's loss
    # curve" is the bf16-eager arm; the fp32 arm shows how far BOTH bf16 paths sit from exact arithmetic (the weights they
    # train are rounded to bf16 after every update, the fp32 arm's are not — that, not the kernels, is the 1e-2 gap)
    dkr = [abs(a - b_) for a, b_ in zip(lk, l16)]
    rel_kr = max(x / max(1.0, abs(r)) for x, r in zip(dkr, l16))
    rec = {"method": method, "optimizer": optname, "steps": STEPS, "lr": lr,
           "config": "widths 320/640/1280, heads 5/10/20, depth 0/1/2, cross dim 2048, latent 32x32, B=2, 10 batches cycled",
           "params": int(net.store.numel),
           "loss_first": {"kernel": lk[0], "oracle_fp32": lo[0], "oracle_bf16_eager": l16[0]},
           "loss_last": {"kernel": lk[-1], "oracle_fp32": lo[-1], "oracle_bf16_eager": l16[-1]},
           "kernel_vs_reference_bf16_eager": {"max_abs": max(dkr), "mean_abs": sum(dkr) / STEPS,
                                              "argmax_step": dkr.index(max(dkr)), "max_rel_to_max1loss": rel_kr},
           "kernel_vs_fp32": {"max_abs": max(d), "mean_abs": sum(d) / STEPS, "argmax_step": d.index(max(d))},
           "bf16_eager_vs_fp32_yardstick": {"max_abs": max(d16), "mean_abs": sum(d16) / STEPS},
           "north_star_1e-3_vs_reference_met": rel_kr <= 1e-3,
           "curves": {"kernel": lk, "oracle_fp32": lo, "oracle_bf16_eager": l16}}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"r2_trajectory_{method}_{optname}_lr{lr:g}.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(f"\n{method}/{optname} lr {lr:g}: loss {lo[0]:.4f} -> {lo[-1]:.4f} (fp32 oracle); kernel vs REFERENCE (bf16 eager) max |d| "
          f"{max(dkr):.3e} (rel {rel_kr:.2e}) mean {sum(dkr) / STEPS:.3e}; kernel vs fp32 max |d| {max(d):.3e}; "
          f"bf16-eager vs fp32 max |d| {max(d16):.3e}")
    assert sum(lo[-10:]) < sum(lo[:10]), "the oracle's loss did not go down: the trajectory is not exercising learning"
    assert d[0] <= max(1e-3 * max(1.0, lo[0]), 1.5 * d16[0]), f"first-step loss differs: {d[0]:.3e} (bf16 eager: {d16[0]:.3e})"
    # (1) north star: within 1e-3 (relative to max(1, loss)) of the reference's own numerics — or, where stochastic rounding /
    #     a larger lr make two bf16 runs drift, at most a quarter of the distance between the reference and exact arithmetic
    #     (AdamWBF16: the kernel's Philox stream and the oracle's torch.randint stream round every update differently, so the
    #     two bf16 runs are two independent samples of the same stochastic process: they may sit as far from each other as
    #     each sits from the exact curve — factor 1.5 instead of 0.25)
    fac = 0.25 if optname == "adamw_fp32_master" else 1.5
    assert rel_kr <= 1e-3 or max(dkr) <= fac * max(d16), \
        f"kernel vs reference (bf16 eager): max |d| {max(dkr):.3e}, rel {rel_kr:.2e}; reference vs fp32 {max(d16):.3e}"
    # (2) and it tracks the exact curve as well as the reference's precision does
    bar = max(1e-3, 1.5 * max(d16))
    assert max(d) <= bar, f"loss curves diverge: max |d| {max(d):.3e} > {bar:.3e}"
