"""Every branch of the reference's loss plugins on the GPU, against the oracle on identical inputs.

Reference branches covered (src/training/trainers/methods/):
  * min-SNR weighting, the shipped default `min_snr_gamma: 5.0` (ddpm_trainer.py:334-343; src/config.yaml:13)
  * tag-weight scaling from `metadata[i]["tag_info"]` (ddpm_trainer.py:348-368) and `batch["tag_weights"]`
    (flow_matching_trainer.py:326-328)
  * epsilon prediction (ddpm_trainer.py:328-333)
  * clamp(max=1000): timesteps 998 / 999 -> sigma ~ 0.002 -> v-target ~ 500 (ddpm_trainer.py:384)
  * non-finite loss -> 1000.0 with no gradient (ddpm_trainer.py:378-382; flow_matching_trainer.py:331-335; decision B22)
each in eager mode (autograd bridge, `loss.backward()`) and — where the trainer takes the graph path — in CUDA-graph
mode (`cuda_graph=True`: the same kernels replayed; noise drawn in-graph is read back and handed to the oracle).

Tolerances (stated): loss |d| <= 2e-2 * max(1, |loss|) (bf16 activations vs the fp32 oracle on a random-init tiny UNet,
the bar of tests/test_gpu_unet.py); probe-gradient rel-L2 <= 5e-2; the clamp / fallback values are exact (1000.0, all
gradients exactly zero).  b2_make_noisy's ddpm arithmetic is also checked bit-for-bit against the golden vectors that
the reference's own add_noise / get_velocity produced (tests/golden/schedule_golden.json).
"""
import json
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
PROBE = "mid_block.resnets.0.conv1.weight"


def _conf(method="ddpm", pred="v_prediction", gamma=None):
    return SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0, use_ztsnr=True,
                                                 min_snr_gamma=gamma),
                           training=SimpleNamespace(method=method, prediction_type=pred, gradient_accumulation_steps=1,
                                                    clip_grad_norm=1.0))


def _setup(seed=5):
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


def _batch(cfg, B=2, H=16, W=16, seed=1, metadata=None):
    g = torch.Generator().manual_seed(seed)
    pooled_dim = cfg["projection_class_embeddings_input_dim"] - 6 * cfg["addition_time_embed_dim"]
    return {"vae_latents": torch.randn(B, 4, H, W, generator=g).to(bf16).float(),
            "prompt_embeds": torch.randn(B, 77, cfg["cross_attention_dim"], generator=g).to(bf16).float(),
            "pooled_prompt_embeds": torch.randn(B, pooled_dim, generator=g).to(bf16).float(),
            "time_ids": torch.tensor([[8. * W, 8. * H, 0., 0., 8. * W, 8. * H]]).repeat(B, 1)[:, None],
            "metadata": metadata if metadata is not None else [{} for _ in range(B)],
            "noise": torch.randn(B, 4, H, W, generator=g).to(bf16).float()}


TAGS = [{"tag_info": {"tags": {"subject": [{"tag": "cat", "weight": 1.5}, {"tag": "dog", "weight": 0.5}],
                               "style": [{"tag": "oil", "weight": 2.0}]}}},
        {"tag_info": {"tags": {"subject": [{"tag": "tree", "weight": 0.75}]}}}]
TAG_MEAN = float(torch.tensor([(1.5 + 0.5 + 2.0) / 3, 0.75], dtype=bf16).mean())  # model dtype, ddpm_trainer.py:365


def _grads_all_zero(net):
    return float(net.store.grad.float().abs().max()) == 0.0


def _run_ddpm(graph, *, gamma=None, pred="v_prediction", t=(700, 880), metadata=None, poison=False):
    """One ddpm micro-step through the plugin surface + the oracle on the same inputs; returns everything a test checks."""
    from oracle import schedule as S
    from sdxl_training_improvements_b200.trainer import B200AdamW, B200DDPMTrainer
    cfg, ref, net = _setup()
    if poison:  # a non-finite prediction: NaN in conv_out.bias reaches `pred` and nothing else
        sd = ref.state_dict()
        sd["conv_out.bias"] = sd["conv_out.bias"].clone()
        sd["conv_out.bias"][0] = float("nan")
        ref.load_state_dict(sd)
        net.load_state_dict(sd)
    tr = B200DDPMTrainer(net, B200AdamW(net, lr=1e-4), None, "cuda", config=_conf("ddpm", pred, gamma), cuda_graph=graph)
    b = _batch(cfg, metadata=metadata)
    ts = torch.tensor(t)
    net.zero_grad()
    if graph:
        out = tr.training_step(b, timesteps=ts)           # the graph path draws its own (Philox) noise in-graph
        assert tr._micro_graphs, "cuda_graph=True did not take the graph path"
        gm = next(iter(tr._micro_graphs.values()))
        noise = gm.static_last["noise"].view(b["noise"].shape).clone()
    else:
        out = tr.training_step(b, noise=b["noise"], timesteps=ts)
        noise = b["noise"].cuda()
    out["loss"].backward()
    torch.cuda.synchronize()
    tw = None
    if metadata is not None:
        tw = TAG_MEAN
    o = S.ddpm_loss(ref, b["vae_latents"].cuda(), noise, ts.cuda(), b["prompt_embeds"].cuda(),
                    b["pooled_prompt_embeds"].cuda(), b["time_ids"].cuda(), sigmas=S.schedule_sigmas(),
                    prediction_type=pred, min_snr_gamma=gamma, tag_weight_mean=tw)
    if o["loss"].requires_grad:
        o["loss"].backward()
    return SimpleNamespace(out=out, oracle=o, net=net, ref=ref, tr=tr)


def _check_close(r):
    lk, lo = float(r.out["loss"]), float(r.oracle["loss"])
    assert abs(lk - lo) <= 2e-2 * max(1.0, abs(lo)), (lk, lo)
    assert abs(r.out["metrics"]["loss"] - lk) < 1e-6
    gk = dict(r.net.named_parameters())[PROBE].grad.float()
    go = dict(r.ref.named_parameters())[PROBE].grad
    rel = float((gk - go).norm() / go.norm())
    assert rel <= 5e-2, rel
    return lk, lo, rel


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_ddpm_min_snr_gamma_default(graph):
    """src/config.yaml:13 ships min_snr_gamma = 5.0.  The table is descending (index 0 = sigma 20000): t = 700 -> sigma 18.9,
    snr 0.0028 < 5 (weight = snr); t = 880 -> sigma 0.33, snr 9.3 > 5 (weight = gamma)."""
    from oracle import schedule as S
    sig = S.schedule_sigmas()
    snr = S.get_snr(sig[torch.tensor([700, 880])])
    assert float(snr[0]) < 5.0 < float(snr[1]), "the two timesteps must sit on both sides of gamma"
    r = _run_ddpm(graph, gamma=5.0)
    lk, lo, rel = _check_close(r)
    # the weighting must actually matter for this draw: the unweighted loss is far from the weighted one
    r0 = _run_ddpm(False, gamma=None)
    assert abs(float(r0.oracle["loss"]) - lo) > 0.1 * abs(lo)
    print(f"\nmin-SNR {'graph' if graph else 'eager'}: kernel {lk:.6f} oracle {lo:.6f} grad rel-L2 {rel:.2e}")


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_ddpm_epsilon_prediction(graph):
    r = _run_ddpm(graph, pred="epsilon", gamma=5.0)
    _check_close(r)
    # epsilon target = the noise itself (ddpm_trainer.py:328-329): the kernel's fp32 target buffer is bit-equal to it
    core = r.tr.core
    last = next(iter(r.tr._micro_graphs.values())).static_last if graph else core.last
    assert torch.equal(last["target"].flatten(), last["noise"].flatten())


def test_ddpm_tag_weights_scale_the_loss():
    """metadata[i]["tag_info"] present on every sample -> loss * mean(per-sample mean tag weight); gradients scale too."""
    r = _run_ddpm(False, gamma=5.0, metadata=TAGS)
    lk, lo, _ = _check_close(r)
    r1 = _run_ddpm(False, gamma=5.0)
    assert abs(lk / float(r1.out["loss"]) - TAG_MEAN) < 1e-3  # same inputs, same noise: the ratio is the weight
    g_w = dict(r.net.named_parameters())[PROBE].grad.float()
    g_1 = dict(r1.net.named_parameters())[PROBE].grad.float()
    assert float((g_w - TAG_MEAN * g_1).norm() / g_w.norm()) < 4e-2  # two bf16 backward passes with differently scaled dL/dpred


def test_ddpm_tag_weights_under_cuda_graph_mode_use_the_eager_kernels():
    """cuda_graph=True: a tag-weighted batch (host-side scalar scale) runs the same kernels un-captured."""
    from sdxl_training_improvements_b200.trainer import B200AdamW, B200DDPMTrainer
    cfg, ref, net = _setup()
    tr = B200DDPMTrainer(net, B200AdamW(net, lr=1e-4), None, "cuda", config=_conf("ddpm", gamma=5.0), cuda_graph=True)
    b = _batch(cfg, metadata=TAGS)
    out = tr.training_step(b, timesteps=torch.tensor([700, 880]))
    out["loss"].backward()
    assert not tr._micro_graphs and float(net.store.grad.float().abs().max()) > 0.0
    r = _run_ddpm(False, gamma=5.0, metadata=TAGS)  # same weights / batch; different noise draw -> same magnitude only
    assert 0.2 < float(out["loss"]) / float(r.out["loss"]) < 5.0


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_ddpm_clamp_at_1000_has_zero_gradient(graph):
    """t = 998, 999: sigma ~ 0.002, v-target ~ (eps - x) / 0.002 -> mean squared error ~ 5e5 -> clamp(max=1000); the
    clamp passes no gradient (torch.clamp's backward), so every parameter gradient is exactly zero."""
    r = _run_ddpm(graph, t=(998, 999))
    assert float(r.out["loss"]) == 1000.0 and float(r.oracle["loss"]) == 1000.0
    assert _grads_all_zero(r.net)
    assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in r.ref.parameters())
    r.tr.optimizer_step()  # a zero-gradient step must not produce non-finite parameters
    assert bool(torch.isfinite(r.net.store.flat.float()).all())


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_ddpm_nonfinite_prediction_falls_back_to_1000(graph):
    """ddpm_trainer.py:380-382: a non-finite loss is replaced by 1000.0 — not an error, and (B22) no gradient."""
    r = _run_ddpm(graph, gamma=5.0, poison=True)
    assert not bool(torch.isfinite(r.oracle["pred"]).all())
    assert float(r.out["loss"]) == 1000.0 and float(r.oracle["loss"]) == 1000.0
    assert _grads_all_zero(r.net)
    assert r.out["metrics"]["loss"] == 1000.0


@pytest.mark.parametrize("case", ["tag_weights", "nonfinite", "plain_graph"])
def test_flow_branches(case):
    from oracle import schedule as S
    from sdxl_training_improvements_b200.trainer import B200AdamW, B200FlowMatchingTrainer
    cfg, ref, net = _setup(seed=6)
    if case == "nonfinite":
        sd = ref.state_dict()
        sd["conv_out.bias"] = sd["conv_out.bias"].clone()
        sd["conv_out.bias"][1] = float("inf")
        ref.load_state_dict(sd)
        net.load_state_dict(sd)
    graph = case == "plain_graph"
    tr = B200FlowMatchingTrainer(net, B200AdamW(net, lr=1e-4), None, "cuda", config=_conf("flow_matching"), cuda_graph=graph)
    b = _batch(cfg, seed=2)
    t = torch.sigmoid(torch.randn(2, generator=torch.Generator().manual_seed(9))).to(bf16)
    tw = None
    if case == "tag_weights":
        b["tag_weights"] = torch.tensor([1.25, 0.5])
        tw = float(b["tag_weights"].to(bf16).float().mean())
    net.zero_grad()
    if graph:
        out = tr.compute_loss(net, b, t=t)
        x0 = next(iter(tr._micro_graphs.values())).static_last["noise"].view(b["noise"].shape).clone()
    else:
        out = tr.compute_loss(net, b, x0=b["noise"], t=t)
        x0 = b["noise"].cuda()
    out["loss"].backward()
    # the reference casts everything to the model dtype (flow_matching_trainer.py:288-306): bf16 op-by-op path
    x1b, x0b = b["vae_latents"].cuda().to(bf16), x0.to(bf16)
    ref16_in = S.optimal_transport_path(x0b, x1b, t.cuda()).float()
    o = S.flow_loss(ref, b["vae_latents"].cuda(), x0, t.float().cuda(), b["prompt_embeds"].cuda(),
                    b["pooled_prompt_embeds"].cuda(), b["time_ids"].cuda(), tag_weight_mean=tw)
    assert float((ref16_in - o["noisy"]).abs().max()) <= 4e-2  # bf16 rounding of x_t only
    lk, lo = float(out["loss"]), float(o["loss"])
    if case == "nonfinite":
        assert lk == 1000.0 and lo == 1000.0 and _grads_all_zero(net)
        return
    o["loss"].backward()
    assert abs(lk - lo) <= 2e-2 * max(1.0, abs(lo)), (lk, lo)
    gk = dict(net.named_parameters())[PROBE].grad.float()
    go = dict(ref.named_parameters())[PROBE].grad
    assert float((gk - go).norm() / go.norm()) <= 5e-2
    m = out["metrics"]
    for k in ("loss", "x0_norm", "x1_norm", "time_mean", "time_std", "velocity_norm", "batch_size", "lr"):
        assert k in m, k  # flow_matching_trainer.py:338-347
    assert abs(m["x1_norm"] - float(b["vae_latents"].norm())) <= 1e-2 * m["x1_norm"]


def test_make_noisy_matches_reference_golden():
    """b2_make_noisy (ddpm mode) against add_noise / get_velocity values produced by the reference's own NoiseScheduler
    (novelai_v3.py:111-127) — fp32 targets bit-exact, noisy latents equal to the golden value rounded to bf16."""
    from oracle import schedule as S
    from sdxl_training_improvements_b200 import ops
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "schedule_golden.json")))
    x = torch.tensor(G["x"]).view(2, 4, 2, 2).cuda()
    n = torch.tensor(G["n"]).view(2, 4, 2, 2).cuda()
    sig = S.schedule_sigmas()[torch.tensor([10, 900])].float().cuda()
    noisy, target = ops.make_noisy(x.contiguous(), n.flatten().contiguous(), sig, 0, True, True, 2, 4, 4, 8)
    want_noisy = torch.tensor(G["add_noise"]).view(2, 4, 4)               # [B, C, HW]
    got_noisy = noisy.view(2, 4, 8)[:, :, :4].permute(0, 2, 1).float().cpu()  # [B, HW, Cpad] -> [B, C, HW]
    assert torch.equal(got_noisy, want_noisy.to(bf16).float())
    assert float(noisy.view(2, 4, 8)[:, :, 4:].float().abs().max()) == 0.0  # pad channels are zero
    assert target.flatten().cpu().tolist() == G["velocity"]
