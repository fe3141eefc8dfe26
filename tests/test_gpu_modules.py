"""T3 + BASELINE.json configs[0]: single modules on the kernels vs the oracle's modules on the CPU in fp32.

  * ResnetBlock2D(320 -> 320) on a 64x64 latent, B=4 — BASELINE config 1 ("fwd/bwd on CPU eager vs new kernel — loss match")
    and the 2C -> C variant with the 1x1 conv_shortcut (up-block resnets);
  * BasicTransformerBlock(1280, 20 heads, n=1024 tokens, 77 x 2048 context) — self-attention, cross-attention, GEGLU FFN.

Oracle: oracle/unet_sdxl.py ResnetBlock2D / BasicTransformerBlock (diffusers `resnet.py` / `attention.py` restated), run in
PyTorch eager fp32 on the HOST CPU with the same bf16-rounded weights and inputs.  Tolerances (stated):
forward rel-L2 <= 1e-2, loss |d| <= 1e-3 (loss = mean squared output, O(1)), every gradient rel-L2 <= 1e-2 ... except that
per-tensor gradients that pass through bf16 split-K partial sums get 2e-2 (tests/test_gpu_kernels.py states the same bar).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rel(a, b):
    a = a.detach().float().flatten().cpu(); b = b.detach().float().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _round_(m):
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.to(bf16).float())
    return m


@pytest.mark.parametrize("prefix,cin,cout", [("down_blocks.0.resnets.0", 320, 320), ("up_blocks.0.resnets.1", 640, 320)])
def test_resnet_block_320_64x64_vs_cpu_oracle(prefix, cin, cout):
    from oracle.unet_sdxl import ResnetBlock2D
    from sdxl_training_improvements_b200.modules import ModuleRunner
    torch.manual_seed(0)
    B, H, W, temb = 4, 64, 64, 1280
    ref = _round_(ResnetBlock2D(cin, cout, temb))          # torch default init, CPU fp32
    run = ModuleRunner(320)
    run.load_module_state(prefix, ref.state_dict())
    x = torch.randn(B, cin, H, W).to(bf16).float()
    emb = torch.randn(B, temb).to(bf16).float()
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    out_ref = ref(xr, er)
    loss_ref = out_ref.square().mean()
    loss_ref.backward()

    run.zero_grad()
    out = run.resnet_forward(prefix, x.cuda(), emb.cuda())
    loss = out.float().square().mean()
    dout = (2.0 / out.numel()) * out.float()               # d(mean square)/d out, from the kernel's own output
    dx, demb = run.resnet_backward(dout)
    torch.cuda.synchronize()

    assert _rel(out, out_ref) <= 1e-2, _rel(out, out_ref)
    assert abs(float(loss) - float(loss_ref)) <= 1e-3, (float(loss), float(loss_ref))
    assert _rel(dx, xr.grad) <= 1e-2, _rel(dx, xr.grad)
    assert _rel(demb, er.grad) <= 2e-2, _rel(demb, er.grad)
    got = run.module_grads(prefix)
    worst = ("", 0.0)
    for k, p in ref.named_parameters():
        r = _rel(got[k], p.grad)
        worst = max(worst, (k, r), key=lambda t: t[1])
        assert r <= 2e-2, (k, r)
    print(f"\nResnetBlock2D {cin}->{cout} @ {H}x{W}, B={B}: loss kernel {float(loss):.6f} cpu-oracle {float(loss_ref):.6f}; "
          f"fwd rel-L2 {_rel(out, out_ref):.2e}; dx {_rel(dx, xr.grad):.2e}; worst param grad {worst[0]} {worst[1]:.2e}")


def test_basic_transformer_block_1280_n1024_vs_cpu_oracle():
    from oracle.unet_sdxl import BasicTransformerBlock
    from sdxl_training_improvements_b200.modules import ModuleRunner
    torch.manual_seed(1)
    B, n, Cc, heads, n_ctx, cdim = 2, 1024, 1280, 20, 77, 2048
    ref = _round_(BasicTransformerBlock(Cc, heads, cdim))
    run = ModuleRunner(Cc, depth=1, cross_attention_dim=cdim)
    prefix = "mid_block.attentions.0.transformer_blocks.0"
    run.load_module_state(prefix, ref.state_dict())
    x = torch.randn(B, n, Cc).to(bf16).float()
    ctx = torch.randn(B, n_ctx, cdim).to(bf16).float()
    xr = x.clone().requires_grad_(True)
    out_ref = ref(xr, ctx)
    loss_ref = out_ref.square().mean()
    loss_ref.backward()

    run.zero_grad()
    out = run.transformer_block_forward(prefix, x.cuda(), ctx.cuda())
    loss = out.float().square().mean()
    dx = run.transformer_block_backward((2.0 / out.numel()) * out.float())
    torch.cuda.synchronize()

    assert _rel(out, out_ref) <= 1e-2, _rel(out, out_ref)
    assert abs(float(loss) - float(loss_ref)) <= 1e-3 * max(1.0, float(loss_ref)), (float(loss), float(loss_ref))
    assert _rel(dx, xr.grad) <= 1e-2, _rel(dx, xr.grad)
    got = run.module_grads(prefix)
    worst = ("", 0.0)
    for k, p in ref.named_parameters():
        r = _rel(got[k], p.grad)
        worst = max(worst, (k, r), key=lambda t: t[1])
        assert r <= 2e-2, (k, r)
    print(f"\nBasicTransformerBlock C={Cc} n={n} B={B}: loss kernel {float(loss):.6f} cpu-oracle {float(loss_ref):.6f}; "
          f"fwd rel-L2 {_rel(out, out_ref):.2e}; dx {_rel(dx, xr.grad):.2e}; worst param grad {worst[0]} {worst[1]:.2e}")
