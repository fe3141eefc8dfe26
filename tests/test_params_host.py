"""Host-side logic that needs no GPU: parameter inventory == oracle state dict, flat-buffer views, C-ABI exports."""
import ctypes
import os
import re

import torch

from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
from sdxl_training_improvements_b200 import _lib
from sdxl_training_improvements_b200.params import SDXL_BASE, ParamStore, unet_param_specs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_specs_match_oracle_state_dict_full():
    with torch.device("meta"):
        sd = OracleUNet().state_dict()
    specs = dict(unet_param_specs(SDXL_BASE))
    assert set(specs) == set(sd)
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in specs.items())
    assert len(specs) == 1680
    assert sum(int(torch.Size(v).numel()) for v in specs.values()) == 2_567_463_684


def test_param_store_views_and_roundtrip():
    cfg = tiny_config()
    ref = seeded_init_(OracleUNet(cfg), 3)
    st = ParamStore(cfg, device="cpu")
    st.load_state_dict(ref.state_dict())
    sd = st.state_dict()
    for k, v in ref.state_dict().items():
        assert torch.equal(sd[k], v.to(torch.bfloat16)), k
    # 3x3 conv weights: logical OIHW, physical O(kh)(kw)I == the K-major [Cout, 9*Cin] GEMM operand
    w = sd["down_blocks.1.resnets.0.conv1.weight"]
    O, I = w.shape[:2]
    assert w.is_contiguous(memory_format=torch.channels_last)
    wk = st.w("down_blocks.1.resnets.0.conv1.weight", O, 9 * I)
    assert torch.equal(wk, w.permute(0, 2, 3, 1).reshape(O, 9 * I))
    # fused QKV / KV views
    p = "down_blocks.1.attentions.0.transformer_blocks.0"
    Cc = sd[f"{p}.attn1.to_q.weight"].shape[0]
    assert st.adjacent(f"{p}.attn1.to_q.weight", f"{p}.attn1.to_k.weight", f"{p}.attn1.to_v.weight")
    qkv = st.w(f"{p}.attn1.to_q.weight", 3 * Cc, Cc)
    assert torch.equal(qkv, torch.cat([sd[f"{p}.attn1.to_q.weight"], sd[f"{p}.attn1.to_k.weight"],
                                       sd[f"{p}.attn1.to_v.weight"]]))
    kv = st.w(f"{p}.attn2.to_k.weight", 2 * Cc, cfg["cross_attention_dim"])
    assert torch.equal(kv, torch.cat([sd[f"{p}.attn2.to_k.weight"], sd[f"{p}.attn2.to_v.weight"]]))
    # grads are views of the flat gradient buffer, params of the flat parameter buffer
    for name, prm in st.params.items():
        assert prm.grad is not None and prm.grad.shape == prm.shape
        assert prm.grad.stride() == prm.stride()
    st.grad.fill_(1.0)
    assert float(st.params["conv_in.bias"].grad.sum()) == st.params["conv_in.bias"].numel()
    assert all(off % 8 == 0 for off in st.offsets.values())  # 16-byte alignment for TMA


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sdxl_b200.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sdxl_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().b2_version() >= 100
