"""T3/T4: the kernel UNet vs the oracle on identical weights/inputs — forward, input of the loss,
and every parameter gradient.  Tolerances (stated): the kernel path keeps activations in bf16 (like the reference's
whole-model bf16 cast) with fp32 accumulation, the oracle here runs fp32 on the same bf16-rounded weights, so
   forward  rel-L2 <= 2e-2,   parameter-gradient rel-L2 <= 5e-2 per tensor (<= 3e-2 in aggregate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rel(a, b):
    a = a.float().flatten(); b = b.float().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _make(cfg, seed=0):
    from oracle.unet_sdxl import OracleUNet, seeded_init_
    from sdxl_training_improvements_b200.unet import B200UNet
    ref = seeded_init_(OracleUNet(cfg), seed).cuda()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(p.to(bf16).float())
    net = B200UNet(cfg, device="cuda")
    net.load_state_dict(ref.state_dict())
    return ref, net


def _inputs(cfg, B, H, W, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    pooled_dim = cfg["projection_class_embeddings_input_dim"] - 6 * cfg["addition_time_embed_dim"]
    x = torch.randn(B, 4, H, W, device="cuda", generator=g).to(bf16)
    ctx = torch.randn(B, 77, cfg["cross_attention_dim"], device="cuda", generator=g).to(bf16)
    pooled = torch.randn(B, pooled_dim, device="cuda", generator=g).to(bf16)
    tid = torch.tensor([[8. * W, 8. * H, 0., 0., 8. * W, 8. * H]], device="cuda").repeat(B, 1)[:, None]
    t = torch.randint(0, 1000, (B,), device="cuda", generator=g)
    return x, t, ctx, pooled, tid


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 12, 20)])
def test_tiny_unet_forward_backward_vs_oracle(B, H, W):
    from oracle.unet_sdxl import tiny_config
    cfg = tiny_config()
    ref, net = _make(cfg)
    x, t, ctx, pooled, tid = _inputs(cfg, B, H, W)
    out = net(x, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
    ro = ref(x.float(), t, ctx.float(), added_cond_kwargs={"text_embeds": pooled.float(), "time_ids": tid}).sample
    assert out.shape == ro.shape
    e = _rel(out, ro)
    assert e <= 2e-2, f"forward rel-L2 {e}"
    w = torch.randn_like(ro)
    net.zero_grad()
    (out.float() * w).sum().backward()
    (ro * w).sum().backward()
    worst, num, den = ("", 0.0), 0.0, 0.0
    rp = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        g, rg = p.grad.float(), rp[k].grad
        r = _rel(g, rg)
        num += float((g - rg).norm() ** 2); den += float(rg.norm() ** 2)
        if r > worst[1]:
            worst = (k, r)
    agg = (num / den) ** 0.5
    assert worst[1] <= 5e-2, f"worst param grad {worst}"
    assert agg <= 3e-2, f"aggregate grad rel-L2 {agg}"
    # gradient accumulation semantics: a second backward doubles p.grad
    out2 = net(x, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
    (out2.float() * w).sum().backward()
    k = "mid_block.resnets.0.conv1.weight"
    assert _rel(dict(net.named_parameters())[k].grad, 2 * rp[k].grad) <= 5e-2


def test_flow_timestep_and_state_dict_surface():
    from oracle.unet_sdxl import tiny_config
    cfg = tiny_config()
    ref, net = _make(cfg, seed=5)
    x, _, ctx, pooled, tid = _inputs(cfg, 2, 8, 8, seed=6)
    t = torch.tensor([0.3125, 0.84375], device="cuda", dtype=bf16)  # flow matching passes raw bf16 t in (0,1)
    out = net(x, t, encoder_hidden_states=ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
    ro = ref(x.float(), t, ctx.float(), added_cond_kwargs={"text_embeds": pooled.float(), "time_ids": tid}).sample
    assert _rel(out, ro) <= 2e-2
    sd = net.state_dict()
    assert set(sd) == set(ref.state_dict())
    assert all(torch.equal(sd[k].float(), v) for k, v in ref.state_dict().items())
    assert sum(p.numel() for p in net.parameters()) == sum(p.numel() for p in ref.parameters())


def test_fullsize_sdxl_unet_forward_backward_vs_oracle():
    """T4 at the REAL size (2.57 B parameters, SDXL-base config, latent 64x64, B=1), identical weights / inputs:
      * truth      = the oracle module tree in fp32 (CUDA),
      * yardstick  = the same oracle in PyTorch eager bf16 (what the reference's diffusers path computes on a GPU),
      * candidate  = the kernel UNet (bf16 activations, fp32 accumulation).
    Bars: forward rel-L2 vs fp32 <= 2e-2; aggregate parameter-gradient rel-L2 vs fp32 <= max(1.5e-2, 2 x yardstick's);
    every tensor with a non-negligible gradient <= max(6e-2, 2.5 x the yardstick's error on that tensor) — random-init
    attention is peaky in the deep blocks and bf16 noise on dS is amplified there for ANY bf16 implementation.
    Exercises what the tiny config cannot: split-K plans, wide tiles, 1280-channel implicit convs, 10 / 20-head attention,
    77-key cross-attention kernels, the 10-deep transformer stacks."""
    import bench
    from oracle.unet_sdxl import OracleUNet
    from sdxl_training_improvements_b200.unet import B200UNet
    net = B200UNet(device="cuda")
    bench._init_weights_(net, 1234)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    B, H, W = 1, 64, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(B, 4, H, W, device="cuda", generator=g) * 3).to(bf16)
    ctx = torch.randn(B, 77, 2048, device="cuda", generator=g).to(bf16)
    pooled = torch.randn(B, 1280, device="cuda", generator=g).to(bf16)
    tid = torch.tensor([[8. * W, 8. * H, 0., 0., 8. * W, 8. * H]], device="cuda").repeat(B, 1)[:, None]
    t = torch.full((B,), 300, device="cuda", dtype=torch.long)
    wgt = torch.randn(B, 4, H, W, device="cuda", generator=g)
    ko = net(x, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
    (ko.float() * wgt).sum().backward()
    gk = {k: p.grad.float().clone() for k, p in net.named_parameters()}
    ko = ko.detach().float()
    del net
    torch.cuda.empty_cache()

    def run_oracle(dtype):
        with torch.device("meta"):
            ref = OracleUNet()
        ref = ref.to_empty(device="cuda").to(dtype)
        ref.load_state_dict({k: v.to(dtype) for k, v in sd.items()})
        out = ref(x.to(dtype), t, ctx.to(dtype), added_cond_kwargs={"text_embeds": pooled.to(dtype), "time_ids": tid}).sample
        (out.float() * wgt).sum().backward()
        grads = {k: p.grad.float().clone() for k, p in ref.named_parameters()}
        out = out.detach().float()
        del ref
        torch.cuda.empty_cache()
        return out, grads

    o32, g32 = run_oracle(torch.float32)
    o16, g16 = run_oracle(bf16)
    assert torch.isfinite(ko).all()
    ef, eo = _rel(ko, o32), _rel(o16, o32)
    assert ef <= 2e-2, f"full-size forward rel-L2 {ef} (eager bf16: {eo})"
    nk = no = den = 0.0
    bad = []
    for k in gk:
        d = float(g32[k].norm() ** 2)
        den += d
        nk += float((gk[k] - g32[k]).norm() ** 2)
        no += float((g16[k] - g32[k]).norm() ** 2)
    for k in gk:
        if float(g32[k].norm()) <= 1e-7 * den ** 0.5:
            continue
        rk, ro_ = _rel(gk[k], g32[k]), _rel(g16[k], g32[k])
        if rk > max(6e-2, 2.5 * ro_):
            bad.append((k, rk, ro_))
    agg_k, agg_o = (nk / den) ** 0.5, (no / den) ** 0.5
    assert agg_k <= max(1.5e-2, 2 * agg_o), f"aggregate gradient rel-L2 {agg_k} (eager bf16: {agg_o})"
    assert not bad, f"{len(bad)} tensors worse than the bf16 yardstick allows: {sorted(bad, key=lambda r: -r[1])[:5]}"
