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


def test_config_json_is_a_complete_diffusers_unet_config(tmp_path):
    """diffusers' UNet2DConditionModel.from_pretrained must be able to rebuild the network from config.json
    (the reference reloads through StableDiffusionXLPipeline.from_pretrained, src/models/sdxl.py:25-40): block types,
    attention_head_dim (= head count per level), text_time addition embedding, linear projections."""
    net = B200UNet(device="cpu") if False else None  # full size is not needed: the config is a pure function of cfg
    from sdxl_training_improvements_b200.params import SDXL_BASE
    shell = B200UNet.__new__(B200UNet)
    shell.config, shell._disk_config = dict(SDXL_BASE), None
    c = shell.diffusers_config()
    assert c["_class_name"] == "UNet2DConditionModel"
    assert c["down_block_types"] == ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"]
    assert c["up_block_types"] == ["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"]
    assert c["mid_block_type"] == "UNetMidBlock2DCrossAttn"
    assert c["attention_head_dim"] == [5, 10, 20] and c["transformer_layers_per_block"] == [1, 2, 10]
    assert c["addition_embed_type"] == "text_time" and c["use_linear_projection"] is True
    assert c["block_out_channels"] == [320, 640, 1280] and c["cross_attention_dim"] == 2048
    assert c["projection_class_embeddings_input_dim"] == 2816 and c["addition_time_embed_dim"] == 256
    assert "num_heads" not in c  # no internal keys leak into the file
    # tiny config: write -> read -> the internal config is recovered; a config read from disk is written back unchanged
    cfg = tiny_config()
    net = B200UNet(cfg, device="cpu")
    d = str(tmp_path / "u")
    net.save_pretrained(d)
    disk = json.load(open(os.path.join(d, "config.json")))
    disk["_diffusers_version"] = "0.32.1"
    disk["some_future_key"] = 7
    json.dump(disk, open(os.path.join(d, "config.json"), "w"))
    net2 = B200UNet.from_pretrained(d, device="cpu")
    for k in ("block_out_channels", "transformer_layers_per_block", "num_heads", "cross_attention_dim"):
        assert tuple(net2.config[k]) == tuple(cfg[k]) if isinstance(cfg[k], tuple) else net2.config[k] == cfg[k], k
    d2 = str(tmp_path / "u2")
    net2.save_pretrained(d2)
    assert json.load(open(os.path.join(d2, "config.json"))) == disk
    # an architecture this UNet does not implement is refused rather than mis-built
    disk["use_linear_projection"] = False
    json.dump(disk, open(os.path.join(d, "config.json"), "w"))
    import pytest
    with pytest.raises(ValueError):
        B200UNet.from_pretrained(d, device="cpu")
