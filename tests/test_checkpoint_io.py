"""diffusers-format checkpoint round trip of the replacement UNet (CPU; no kernels involved).
`save_pretrained` must write what the reference's `StableDiffusionXL.save_pretrained` / `from_pretrained` exchange
(src/models/sdxl.py:25-40, 246-288): diffusers key names, conv weights in logical OIHW, Linear [out, in]."""
import json
import os

import torch

from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
from sdxl_training_improvements_b200.unet import B200UNet

bf16 = torch.bfloat16


def test_save_load_round_trip_and_oracle_compat(tmp_path):
    cfg = tiny_config()
    ref = seeded_init_(OracleUNet(cfg), 7)
    net = B200UNet(cfg, device="cpu")
    net.load_state_dict(ref.state_dict())
    d = str(tmp_path / "unet")
    net.save_pretrained(d, safe_serialization=True)
    assert sorted(os.listdir(d)) == ["config.json", "diffusion_pytorch_model.safetensors"]
    from safetensors.torch import load_file
    disk = load_file(os.path.join(d, "diffusion_pytorch_model.safetensors"))
    # key names, logical shapes and values are diffusers': the oracle (diffusers module tree) loads the file as is
    assert set(disk) == set(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert tuple(disk[k].shape) == tuple(v.shape), k
        assert disk[k].is_contiguous()
        assert torch.equal(disk[k].float(), v.to(bf16).float()), k
    ref2 = OracleUNet(cfg)
    ref2.load_state_dict({k: v.float() for k, v in disk.items()})
    # and back into the flat / channels-last layout
    net2 = B200UNet.from_pretrained(d, device="cpu")
    assert json.load(open(os.path.join(d, "config.json")))["_class_name"] == "UNet2DConditionModel"
    assert net2.config["block_out_channels"] == tuple(cfg["block_out_channels"])
    assert torch.equal(net2.store.flat, net.store.flat)          # physical layout identical
    k = "down_blocks.0.resnets.0.conv1.weight"
    p = dict(net2.named_parameters())[k]
    assert p.shape == ref.state_dict()[k].shape and not p.is_contiguous()  # OIHW view over (O, kh, kw, I) storage


def test_from_pretrained_bin_and_missing(tmp_path):
    cfg = tiny_config()
    net = B200UNet(cfg, device="cpu")
    with torch.no_grad():
        net.store.flat.copy_(torch.randn(net.store.total).to(bf16))
    d = str(tmp_path / "unet_bin")
    net.save_pretrained(d, safe_serialization=False)
    net2 = B200UNet.from_pretrained(d, device="cpu")
    for (k, a), (_, b) in zip(net.named_parameters(), net2.named_parameters()):
        assert torch.equal(a, b), k
    import pytest
    with pytest.raises(FileNotFoundError):
        B200UNet.from_pretrained(str(tmp_path / "nothing_here"), device="cpu")
