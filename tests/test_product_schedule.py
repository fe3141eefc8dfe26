"""The PRODUCT's host-side schedule functions (`trainer.NoiseScheduler`, `trainer.get_karras_sigmas`,
`trainer.sample_logit_normal`) against values produced by the reference's own functions
(tests/golden/schedule_golden.json <- tests/golden/make_schedule_golden.py, which imports
/root/reference/src/training/schedulers/novelai_v3.py and .../flow_matching_trainer.py).

tests/test_oracle_schedule.py pins the oracle's copy; this file pins the functions the trainers actually call, so that
the two cannot drift apart unnoticed.  CPU only: no kernel is launched.
"""
import json
import os
from types import SimpleNamespace

import torch

from sdxl_training_improvements_b200 import trainer as T

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "schedule_golden.json")))


def _sched(use_ztsnr=True, sigma_max=20000.0):
    return T.NoiseScheduler(SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=sigma_max,
                                                                  use_ztsnr=use_ztsnr)))


def test_karras_table_bit_exact():
    sig = T.get_karras_sigmas(1000, 0.002, 20000.0, 7.0)
    assert [float(sig[i]) for i in G["karras_ztsnr_idx"]] == G["karras_ztsnr"]
    assert float(sig.double().sum()) == G["karras_ztsnr_sum"]
    s = _sched(True, 80.0)  # ZTSNR forces sigma_max = 20000 regardless of config (novelai_v3.py:106)
    assert torch.equal(s.sigmas, sig) and torch.equal(s.get_sigmas(1000), sig)
    s80 = _sched(False, 80.0)
    assert [float(s80.sigmas[i]) for i in G["karras_ztsnr_idx"]] == G["karras_80"]
    assert s.rho == 7.0  # B2: ModelConfig has no rho -> the function default


def test_timestep_to_sigma_and_snr_bit_exact():
    s = _sched()
    t = torch.tensor([10, 900])
    assert [float(v) for v in s.get_snr(t)] == G["snr_t10_t900"]
    assert torch.equal(s.timestep_to_sigma(t), s.sigmas[t])
    # the device-side noising / velocity arithmetic (b2_make_noisy) is checked on the GPU against the same golden
    # vectors: tests/test_gpu_loss_branches.py::test_make_noisy_matches_reference_golden


def test_sample_timesteps_match_reference_stream():
    torch.manual_seed(0)
    assert _sched(True).sample_timesteps(4).tolist() == G["sample_timesteps_seed0"]
    torch.manual_seed(0)
    assert _sched(False, 80.0).sample_timesteps(4).tolist() == G["sample_timesteps_seed0_noztsnr"]
    # B1: a `device` argument is accepted
    torch.manual_seed(0)
    assert _sched(True).sample_timesteps(4, device="cpu").tolist() == G["sample_timesteps_seed0"]


def test_logit_normal_matches_reference_stream():
    g = torch.Generator().manual_seed(0)
    t = T.sample_logit_normal((4,), torch.float32, generator=g)
    assert t.tolist() == G["logit_normal_seed0"]
    g = torch.Generator().manual_seed(0)
    tb = T.sample_logit_normal((4,), generator=g)  # the trainers draw in the model dtype (B20)
    assert tb.dtype == torch.bfloat16 and bool(((tb > 0) & (tb < 1)).all())


def test_tag_weight_mean_follows_reference():
    """ddpm_trainer.py:348-368: per sample the mean of its tags' weights, then the mean over samples; any sample without
    tag_info (or without tags) disables the scaling."""
    md = [{"tag_info": {"tags": {"subject": [{"tag": "cat", "weight": 1.5}, {"tag": "dog", "weight": 0.5}],
                                 "style": [{"tag": "oil", "weight": 2.0}]}}},
          {"tag_info": {"tags": {"subject": [{"tag": "tree", "weight": 0.8}]}}}]
    want = float(torch.tensor([(1.5 + 0.5 + 2.0) / 3, 0.8], dtype=torch.bfloat16).mean())  # model dtype, :365
    assert T._tag_weight_mean({"metadata": md}) == want and abs(want - 1.0667) < 5e-3
    assert T._tag_weight_mean({"metadata": [md[0], {}]}) is None
    assert T._tag_weight_mean({"metadata": [{"tag_info": {"tags": {}}}]}) is None
    assert T._tag_weight_mean({"metadata": {}}) is None
