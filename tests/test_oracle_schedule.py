"""The oracle's schedule / loss arithmetic against values produced by the reference's own functions
(tests/golden/schedule_golden.json <- tests/golden/make_schedule_golden.py)."""
import json
import os
import types

import pytest
import torch

from oracle import schedule as S

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "schedule_golden.json")))


def test_karras_table_matches_reference():
    sig = S.karras_sigmas(1000, 0.002, 20000.0, 7.0)
    got = [float(sig[i]) for i in G["karras_ztsnr_idx"]]
    assert got == G["karras_ztsnr"]  # bit-exact: same torch ops in the same order
    assert float(sig.double().sum()) == G["karras_ztsnr_sum"]
    sig80 = S.schedule_sigmas(1000, 0.002, 80.0, use_ztsnr=False)
    assert [float(sig80[i]) for i in G["karras_ztsnr_idx"]] == G["karras_80"]
    # ZTSNR forces sigma_max = 20000 regardless of config (novelai_v3.py:106)
    assert torch.equal(S.schedule_sigmas(1000, 0.002, 80.0, use_ztsnr=True), sig)
    assert sig[0] > sig[1] > sig[-1]  # descending: index 0 is the noisiest (B11)


def test_add_noise_velocity_snr_match_reference():
    sig = S.schedule_sigmas()
    t = torch.tensor([10, 900])
    x = torch.tensor(G["x"]).view(2, 4, 2, 2)
    n = torch.tensor(G["n"]).view(2, 4, 2, 2)
    assert S.add_noise(x, n, sig[t], True).flatten().tolist() == G["add_noise"]
    sig80 = S.schedule_sigmas(1000, 0.002, 80.0, use_ztsnr=False)
    assert S.add_noise(x, n, sig80[t], False).flatten().tolist() == G["add_noise_noztsnr"]
    assert S.get_velocity(x, n, sig[t]).flatten().tolist() == G["velocity"]
    assert [float(v) for v in S.get_snr(sig[t])] == G["snr_t10_t900"]


def test_samplers_match_reference():
    torch.manual_seed(0)
    assert S.sample_timesteps(4, use_ztsnr=True).tolist() == G["sample_timesteps_seed0"]
    torch.manual_seed(0)
    assert S.sample_timesteps(4, use_ztsnr=False).tolist() == G["sample_timesteps_seed0_noztsnr"]
    g = torch.Generator().manual_seed(0)
    t = S.sample_logit_normal((4,), "cpu", torch.float32, generator=g)
    assert t.tolist() == G["logit_normal_seed0"]
    x0 = torch.zeros(4, 1, 1, 1)
    x1 = torch.ones(4, 1, 1, 1)
    assert S.optimal_transport_path(x0, x1, t).flatten().tolist() == G["ot_path_0_1"]


def test_flow_loss_matches_reference():
    class Lin:
        def __call__(self, xt, t, encoder_hidden_states=None, added_cond_kwargs=None):
            return types.SimpleNamespace(sample=0.5 * xt + t.view(-1, 1, 1, 1))
    t = torch.tensor(G["logit_normal_seed0"])
    x0 = torch.tensor(G["flow_x0"]).view(4, 4, 2, 2)
    x1 = torch.tensor(G["flow_x1"]).view(4, 4, 2, 2)
    out = S.flow_loss(Lin(), x1, x0, t, None, None, None)
    per = torch.tensor(G["flow_loss_per_sample"])
    assert torch.allclose(out["loss"], per.mean(), rtol=1e-6, atol=0)


def test_ddpm_loss_semantics():
    class Zero:
        def __call__(self, x, t, e, added_cond_kwargs=None):
            return types.SimpleNamespace(sample=torch.zeros_like(x))
    sig = S.schedule_sigmas()
    torch.manual_seed(3)
    x = torch.randn(2, 4, 4, 4); n = torch.randn(2, 4, 4, 4)
    t = torch.tensor([300, 990])
    o = S.ddpm_loss(Zero(), x, n, t, None, None, None, sigmas=sig, prediction_type="epsilon")
    assert torch.allclose(o["loss"], (n ** 2).mean())
    o = S.ddpm_loss(Zero(), x, n, t, None, None, None, sigmas=sig, prediction_type="v_prediction")
    v = (n - x) / sig[t].view(-1, 1, 1, 1)
    assert torch.allclose(o["loss"], torch.clamp((v ** 2).mean(), max=1000.0))
    # clamp at 1000 (ddpm_trainer.py:384): t=998,999 -> sigma ~ 0.002 -> v^2 ~ 1e5
    o = S.ddpm_loss(Zero(), x, n, torch.tensor([998, 999]), None, None, None, sigmas=sig)
    assert float(o["loss"]) == 1000.0
    # min-SNR with the intended per-sample broadcast (B3)
    o = S.ddpm_loss(Zero(), x, n, t, None, None, None, sigmas=sig, prediction_type="epsilon", min_snr_gamma=5.0)
    w = torch.minimum(1.0 / sig[t] ** 2, torch.tensor(5.0)).view(-1, 1, 1, 1)
    assert torch.allclose(o["loss"], ((n ** 2) * w).mean())
    # tag weights multiply the mean loss (ddpm_trainer.py:366-368)
    o2 = S.ddpm_loss(Zero(), x, n, t, None, None, None, sigmas=sig, prediction_type="epsilon", tag_weight_mean=0.5)
    assert torch.allclose(o2["loss"], 0.5 * (n ** 2).mean())
