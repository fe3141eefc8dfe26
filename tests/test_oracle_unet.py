"""Structural pins of the oracle UNet: the reference has no tests for this path."""
import torch

from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config


def test_sdxl_base_param_count_and_keys():
    with torch.device("meta"):
        m = OracleUNet()
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 2_567_463_684
    assert len(sd) == 1680
    for k in ["conv_in.weight", "time_embedding.linear_1.weight", "add_embedding.linear_2.bias",
              "down_blocks.0.resnets.1.conv2.weight", "down_blocks.0.downsamplers.0.conv.bias",
              "down_blocks.1.attentions.1.transformer_blocks.1.attn2.to_k.weight",
              "down_blocks.2.attentions.0.transformer_blocks.9.ff.net.0.proj.weight",
              "mid_block.attentions.0.proj_out.weight", "mid_block.resnets.1.time_emb_proj.weight",
              "up_blocks.0.resnets.2.conv_shortcut.weight", "up_blocks.1.upsamplers.0.conv.weight",
              "up_blocks.2.resnets.0.norm1.weight", "conv_norm_out.bias", "conv_out.weight"]:
        assert k in sd, k
    assert "down_blocks.0.attentions.0.norm.weight" not in sd
    assert "up_blocks.2.attentions.0.norm.weight" not in sd
    assert tuple(sd["up_blocks.0.resnets.0.conv1.weight"].shape) == (1280, 2560, 3, 3)
    assert tuple(sd["up_blocks.0.resnets.2.conv1.weight"].shape) == (1280, 1920, 3, 3)
    assert tuple(sd["up_blocks.2.resnets.0.conv1.weight"].shape) == (320, 960, 3, 3)
    assert tuple(sd["down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight"].shape) == (640, 2048)
    assert tuple(sd["mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"].shape) == (10240, 1280)
    assert tuple(sd["add_embedding.linear_1.weight"].shape) == (1280, 2816)


def test_tiny_forward_backward_shapes():
    cfg = tiny_config()
    m = seeded_init_(OracleUNet(cfg), 0)
    B = 2
    for (H, W) in [(16, 16), (12, 20)]:
        x = torch.randn(B, 4, H, W)
        ctx = torch.randn(B, 77, cfg["cross_attention_dim"])
        pooled = torch.randn(B, 96)
        tid = torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).repeat(B, 1)[:, None]
        out = m(x, torch.tensor([10, 500]), ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
        assert out.shape == x.shape
        out.square().mean().backward()
    assert all(p.grad is not None for p in m.parameters())


def test_seeded_init_is_deterministic():
    a = seeded_init_(OracleUNet(tiny_config()), 7).state_dict()
    b = seeded_init_(OracleUNet(tiny_config()), 7).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
