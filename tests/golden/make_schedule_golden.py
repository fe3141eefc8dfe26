"""Generate tests/golden/schedule_golden.json by calling the REFERENCE's own functions.

Run in the build container only (needs /root/reference):  python tests/golden/make_schedule_golden.py
The reference imports colorama / diffusers / spacy, which are not installed; three stub modules are
registered first.  Nothing here is imported by the product or at GPU-test time: the
JSON it writes is the committed fixture.
"""
import json
import os
import sys
import tempfile
import types

import torch

REF = "/root/reference"


def _stub_modules():
    col = types.ModuleType("colorama")
    class _Any:
        def __getattr__(self, k):
            return ""
    col.Fore = _Any(); col.Style = _Any(); col.Back = _Any(); col.init = lambda *a, **k: None
    sys.modules["colorama"] = col
    dif = types.ModuleType("diffusers")
    for n in ("DDPMScheduler", "StableDiffusionXLPipeline", "AutoencoderKL", "UNet2DConditionModel",
              "EulerDiscreteScheduler"):
        setattr(dif, n, type(n, (), {"__init__": lambda self, *a, **k: None}))
    sys.modules["diffusers"] = dif
    sp = types.ModuleType("spacy")
    sp.load = lambda *a, **k: None
    sys.modules["spacy"] = sp


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # importing the reference creates ./outputs/logs
    from src.training.schedulers.novelai_v3 import NoiseScheduler, get_karras_sigmas
    from src.training.trainers.methods.flow_matching_trainer import FlowMatchingTrainer
    os.chdir(cwd)

    out = {"reference_commit": "6083befd", "torch": torch.__version__}
    sig = get_karras_sigmas(1000, 0.002, 20000.0, rho=7.0)
    out["karras_ztsnr_idx"] = [0, 1, 10, 250, 500, 750, 900, 998, 999]
    out["karras_ztsnr"] = [float(sig[i]) for i in out["karras_ztsnr_idx"]]
    out["karras_ztsnr_sum"] = float(sig.double().sum())
    sig80 = get_karras_sigmas(1000, 0.002, 80.0, rho=7.0)
    out["karras_80"] = [float(sig80[i]) for i in out["karras_ztsnr_idx"]]

    cfg = types.SimpleNamespace(model=types.SimpleNamespace(
        sigma_min=0.002, sigma_max=80.0, use_ztsnr=True, rho=7.0, num_timesteps=1000))
    ns = object.__new__(NoiseScheduler)
    ns.config = cfg; ns.device = "cpu"; ns.sigma_data = 1.0
    t = torch.tensor([10, 900])
    out["snr_t10_t900"] = [float(v) for v in ns.get_snr(t)]
    torch.manual_seed(0)
    x = torch.randn(2, 4, 2, 2); n = torch.randn(2, 4, 2, 2)
    out["add_noise"] = ns.add_noise(x, n, t).flatten().tolist()
    out["velocity"] = ns.get_velocity(x, n, t).flatten().tolist()
    out["x"] = x.flatten().tolist(); out["n"] = n.flatten().tolist()
    torch.manual_seed(0)
    out["sample_timesteps_seed0"] = ns.sample_timesteps(4).tolist()
    cfg.model.use_ztsnr = False
    torch.manual_seed(0)
    out["sample_timesteps_seed0_noztsnr"] = ns.sample_timesteps(4).tolist()
    out["add_noise_noztsnr"] = ns.add_noise(x, n, t).flatten().tolist()

    g = torch.Generator().manual_seed(0)
    tl = FlowMatchingTrainer.sample_logit_normal(None, (4,), "cpu", torch.float32, generator=g)
    out["logit_normal_seed0"] = tl.tolist()
    x0 = torch.zeros(4, 1, 1, 1); x1 = torch.ones(4, 1, 1, 1)
    out["ot_path_0_1"] = FlowMatchingTrainer.optimal_transport_path(None, x0, x1, tl).flatten().tolist()

    # flow loss through the reference's own function with a linear stand-in model
    class Lin:
        def __call__(self, xt, t, encoder_hidden_states=None, added_cond_kwargs=None):
            return types.SimpleNamespace(sample=0.5 * xt + t.view(-1, 1, 1, 1))
    fm = types.SimpleNamespace(optimal_transport_path=lambda a, b, c: FlowMatchingTrainer.optimal_transport_path(None, a, b, c),
                               compute_velocity=lambda m, xt, tt, ce: FlowMatchingTrainer.compute_velocity(None, m, xt, tt, ce))
    torch.manual_seed(1)
    fx0 = torch.randn(4, 4, 2, 2); fx1 = torch.randn(4, 4, 2, 2)
    per = FlowMatchingTrainer.compute_flow_matching_loss(
        fm, Lin(), fx0, fx1, tl, {"prompt_embeds": None, "added_cond_kwargs": None})
    out["flow_loss_per_sample"] = per.tolist()
    out["flow_x0"] = fx0.flatten().tolist(); out["flow_x1"] = fx1.flatten().tolist()

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "schedule_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
