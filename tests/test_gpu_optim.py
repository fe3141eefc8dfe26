"""AdamWBF16 kernel (b2_adamw_bf16) vs the reference's own `_make_step` (golden vectors generated from
/root/reference by tests/golden/make_adamw_bf16_golden.py) and vs the pinned oracle restatement.

Bit-exact: with the 16 random bits of every stochastic rounding injected (all-zero, all-one and a stored pseudo-random
pattern) the kernel must reproduce the reference's bf16 state after every step exactly.  With its own Philox bits the
kernel is checked statistically: every result is one of the two bf16 neighbours of the exact fp32 value and the
rounding is unbiased."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adamw_bf16_golden.pt")


@pytest.fixture(scope="module")
def ops():
    from sdxl_training_improvements_b200 import ops as o
    return o


@pytest.mark.parametrize("mode,rng_mode", [("zero", 1), ("ffff", 2), ("random", 3), ("zero", 3)])
def test_adamw_bf16_bit_exact_vs_reference_golden(ops, mode, rng_mode):
    gold = torch.load(GOLD, weights_only=False)
    h = gold["hyper"]
    case = gold["cases"][mode]
    p = case["p0"].cuda().clone()
    n = p.numel()
    m = torch.zeros(n, device="cuda", dtype=bf16)
    v = torch.zeros_like(m)
    sh = torch.zeros_like(m)
    for k, ref in enumerate(case["states"]):
        r16 = case["rand16"][k].cuda().contiguous()
        ops.adamw_bf16(p, case["grads"][k].cuda(), m, v, sh, lr=h["lr"], beta1=h["beta1"], beta2=h["beta2"], eps=h["eps"],
                       step=k + 1, rng_mode=rng_mode, test_rand16=r16)
        torch.cuda.synchronize()
        for name, got in (("p", p), ("m", m), ("v", v), ("shift", sh)):
            bad = int((got.cpu().view(torch.int16) != ref[name].view(torch.int16)).sum())
            assert bad == 0, f"{mode} step {k + 1} {name}: {bad}/{n} elements differ from the reference's _make_step"
    # deferred weight decay on top (adamw_bfloat16/__init__.py:191-192)
    k = len(case["states"])
    ops.adamw_bf16(p, case["grads"][0].cuda(), m, v, sh, lr=h["lr"], beta1=h["beta1"], beta2=h["beta2"], eps=h["eps"],
                   step=k + 1, rng_mode=rng_mode, test_rand16=case["rand16"][k].cuda().contiguous())
    alpha = float(torch.tensor(-case["decay_state"]["decay"]).to(bf16))
    ops.axpy_bf16(sh, p, alpha)
    assert torch.equal(sh.cpu(), case["decay_state"]["shift"])
    assert torch.equal(p.cpu(), case["decay_state"]["p"])


def test_adamw_bf16_hash_rounding_is_neighbouring_and_unbiased(ops):
    from oracle import adamw_bf16 as O
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(3)
    p0 = (torch.randn(n, device="cuda", generator=g) * 0.05).to(bf16)
    gr = (torch.randn(n, device="cuda", generator=g) * 1e-2).to(bf16)
    m0 = (torch.randn(n, device="cuda", generator=g) * 1e-2).to(bf16)
    v0 = (torch.rand(n, device="cuda", generator=g) * 1e-4).to(bf16)
    s0 = (torch.randn(n, device="cuda", generator=g) * 1e-5).to(bf16)
    p, m, v, sh = p0.clone(), m0.clone(), v0.clone(), s0.clone()
    so = torch.tensor([77, 0], device="cuda", dtype=torch.int64)
    ops.adamw_bf16(p, gr, m, v, sh, lr=1e-3, step=5, seed_offset=so)
    # exact fp32 pre-rounding value of exp_avg (as written: g + (1-b1) * bf16(b1*m)) and its two bf16 neighbours
    lo = O.make_step(p0.float().cpu(), gr.float().cpu(), m0.float().cpu(), v0.float().cpu(), s0.float().cpu(), beta1=0.9,
                     beta2=0.999, step=5.0, lr=1e-3, eps=1e-8, rand16=torch.zeros(4, n, dtype=torch.int32))
    hi = O.make_step(p0.float().cpu(), gr.float().cpu(), m0.float().cpu(), v0.float().cpu(), s0.float().cpu(), beta1=0.9,
                     beta2=0.999, step=5.0, lr=1e-3, eps=1e-8, rand16=torch.full((4, n), 65535, dtype=torch.int32))
    mk = m.float().cpu()
    assert bool(((mk == lo[1]) | (mk == hi[1])).all()), "exp_avg is not one of the two bf16 neighbours"
    assert torch.equal(v.float().cpu(), lo[2]), "exp_avg_sq is deterministic (round-to-nearest bf16 ops)"
    m1 = (m0.float() * 0.9).to(bf16).float()
    exact = torch.addcmul(gr.float(), torch.tensor(0.1, device="cuda"), m1)
    bias = float((m.float() - exact).double().mean())
    spread = float((hi[1] - lo[1]).abs().double().mean())
    assert abs(bias) < 0.02 * spread, f"stochastic rounding looks biased: mean err {bias:.3e} vs ulp {spread:.3e}"
    # different step -> different bits; same step -> identical (reproducible)
    p2, m2, v2, s2 = p0.clone(), m0.clone(), v0.clone(), s0.clone()
    ops.adamw_bf16(p2, gr, m2, v2, s2, lr=1e-3, step=5, seed_offset=so)
    assert torch.equal(m2, m) and torch.equal(p2, p) and torch.equal(s2, sh)
    p3, m3, v3, s3 = p0.clone(), m0.clone(), v0.clone(), s0.clone()
    ops.adamw_bf16(p3, gr, m3, v3, s3, lr=1e-3, step=6, seed_offset=so)
    assert not torch.equal(m3, m)
    # zero_grad folded into the kernel: same update, gradient vector cleared (odd length: the scalar tail too)
    for cnt in (n, n - 3):
        p5, m5, v5, s5, g5 = p0[:cnt].clone(), m0[:cnt].clone(), v0[:cnt].clone(), s0[:cnt].clone(), gr[:cnt].clone()
        p6, m6, v6, s6 = p0[:cnt].clone(), m0[:cnt].clone(), v0[:cnt].clone(), s0[:cnt].clone()
        ops.adamw_bf16(p5, g5, m5, v5, s5, lr=1e-3, step=5, seed_offset=so, zero_grad=True)
        ops.adamw_bf16(p6, gr[:cnt].clone(), m6, v6, s6, lr=1e-3, step=5, seed_offset=so)
        assert torch.equal(p5, p6) and torch.equal(m5, m6) and torch.equal(v5, v6) and torch.equal(s5, s6)
        assert int(g5.view(torch.int16).count_nonzero()) == 0
    # the parameter rounding draws from the SECOND random word: with g = 0 and exp_avg = 0 the update is p <- SR(p + shift)
    # exactly (shift passes through its own rounding unchanged), so its bias is measurable from the outputs alone
    z = torch.zeros(n, device="cuda", dtype=bf16)
    s_big = (torch.randn(n, device="cuda", generator=g) * 3e-4).to(bf16)
    p4, m4, v4, s4 = p0.clone(), z.clone(), v0.clone(), s_big.clone()
    ops.adamw_bf16(p4, z, m4, v4, s4, lr=1e-3, step=5, seed_offset=so)
    exact_p = p0.float() + s_big.float()
    down = (exact_p.view(torch.int32) & -65536).view(torch.float32)
    ulp = (((exact_p.view(torch.int32) & -65536) + 65536).view(torch.float32) - down).abs()
    err = (p4.float() - exact_p).double()
    assert bool(((p4.float() == down) | ((p4.float() - down).abs() == ulp)).all()), "p is not a bf16 neighbour of p + shift"
    assert abs(float(err.mean())) < 0.02 * float(ulp.double().mean()), f"p rounding biased: {float(err.mean()):.3e}"
    # and the two words are not the same draw: rounding direction of exp_avg and of p must be (nearly) uncorrelated
    up_m = (mk != lo[1]).double()
    up_p = (p4.float().cpu() != down.cpu()).double()
    corr = float(((up_m - up_m.mean()) * (up_p - up_p.mean())).mean() / (up_m.std() * up_p.std() + 1e-12))
    assert abs(corr) < 0.01, f"rounding directions of exp_avg and p correlate: {corr:.4f}"


def test_adamw_denominator_all_bf16_patterns(ops):
    """optim.cu computes bf16(bf16(sqrt v) + eps) with `sqrt.approx.f32`; the result must equal the IEEE statement for
    EVERY bf16 input (adamw_bfloat16/__init__.py:176 `exp_avg_sq.sqrt().add_(eps)` on bf16 tensors)."""
    from sdxl_training_improvements_b200 import _lib
    allv = torch.arange(65536, dtype=torch.int32).to(torch.int16).cuda().view(bf16)
    for eps in (1e-8, 1e-6, 1e-3):
        fast, ieee = torch.empty_like(allv), torch.empty_like(allv)
        _lib.check(_lib.load().b2_adamw_denom_test(allv.data_ptr(), fast.data_ptr(), ieee.data_ptr(), allv.numel(), eps, None),
                   "denom_test")
        torch.cuda.synchronize()
        a, b = fast.view(torch.int16), ieee.view(torch.int16)
        nan = torch.isnan(fast) & torch.isnan(ieee)
        assert bool(((a == b) | nan).all()), f"eps={eps}: {(~((a == b) | nan)).sum().item()} bf16 patterns differ"
        ref = (allv.float().sqrt().to(bf16).float() + eps).to(bf16)
        ok = (ref.view(torch.int16) == b) | (torch.isnan(ref) & torch.isnan(ieee))
        assert bool(ok.all()), "IEEE leg disagrees with torch"


def test_adamw_bf16_optimizer_class_on_tiny_unet(ops):
    """B200AdamWBF16 over the flat buffers == the oracle applied to the same flat buffers (truncating rounding), incl.
    the global-norm clip and the per-tensor deferred decay bookkeeping."""
    from oracle import adamw_bf16 as O
    from oracle.unet_sdxl import tiny_config
    from sdxl_training_improvements_b200.trainer import B200AdamWBF16
    from sdxl_training_improvements_b200.unet import B200UNet
    net = B200UNet(tiny_config(), device="cuda")
    st = net.store
    gen = torch.Generator(device="cuda").manual_seed(11)
    st.flat.copy_((torch.randn(st.total, device="cuda", generator=gen) * 0.05).to(bf16))
    opt = B200AdamWBF16(net, lr=1e-3, weight_decay=4.0, seed=5)   # wd*lr = 4e-3: the threshold trips on step 1 or 2
    p = st.flat.float().cpu()
    m = torch.zeros_like(p); v = torch.zeros_like(p); sh = torch.zeros_like(p)
    acc = dict(opt.accumulated_decay)
    zeros = torch.zeros(4, st.total, dtype=torch.int32)
    for step in (1, 2):
        st.grad.copy_((torch.randn(st.total, device="cuda", generator=gen) * 0.5).to(bf16))
        g = st.grad.float().cpu()
        norm = float(g.double().pow(2).sum().sqrt())
        clip = min(1.0, 1.0 / (norm + 1e-6))
        opt.fused_step(max_norm=1.0, rng_mode=1)
        p, m, v, sh = O.make_step(p, g, m, v, sh, beta1=0.9, beta2=0.999, step=float(step), lr=1e-3, eps=1e-8,
                                  rand16=zeros, clip=clip)
        for name, _ in st.specs:
            a = acc[name] + 4.0 * 1e-3
            if a > 5e-3:
                off, nn = st.offsets[name], st._numel[name]
                sh[off:off + nn] = O.apply_decay(sh[off:off + nn], p[off:off + nn], a)
                a = 0.0
            acc[name] = a
        torch.cuda.synchronize()
        for name, got, want in (("p", st.flat, p), ("m", opt.exp_avg, m), ("v", opt.exp_avg_sq, v), ("shift", opt.shift, sh)):
            gotf = got.float().cpu()
            bad = int((gotf != want).sum())
            # the clip coefficient is computed from an fp64 device sum vs this host sum: allow a few 1-ulp flips
            assert bad <= max(8, got.numel() // 2000), f"step {step} {name}: {bad} elements differ"
            assert float((gotf - want).abs().max()) <= 2 ** -6 * float(want.abs().max())
    assert opt.state_dict()["state"]["conv_in.weight"]["exp_avg"].numel() == 64 * 4 * 9


def test_adamw_random_bits_match_the_numpy_restatement(ops):
    """Pins the kernel's counter hash to tests/test_adamw_hash_cpu.py bit for bit: with g = 0 and exp_avg = 0 the parameter and
    the shift are rounded with the two halves of the SECOND random word, and with a plain gradient exp_avg is rounded with the
    low half of the FIRST — all three predicted exactly from the numpy hash (stochastic/__init__.py:46-71: add 16 random bits
    below the bf16 mantissa, truncate)."""
    import numpy as np
    from test_adamw_hash_cpu import keys, rand64
    n = 1 << 16
    seed, step = 1234567891234, 9
    g = torch.Generator(device="cuda").manual_seed(11)
    p0 = (torch.randn(n, device="cuda", generator=g) * 0.05).to(bf16)
    s0 = (torch.randn(n, device="cuda", generator=g) * 3e-4).to(bf16)
    v0 = (torch.rand(n, device="cuda", generator=g) * 1e-4).to(bf16)
    z = torch.zeros(n, device="cuda", dtype=bf16)
    so = torch.tensor([seed, 0], device="cuda", dtype=torch.int64)
    ka, kb = keys(seed, step)
    r01, r23 = rand64(np.arange(n, dtype=np.uint64), ka, kb)
    r01, r23 = r01.astype(np.uint32), r23.astype(np.uint32)

    def bits(t):
        return t.float().cpu().numpy().view(np.uint32)

    def sr(x32, r16):   # x32: float32 array
        return ((x32.view(np.uint32).astype(np.uint64) + r16.astype(np.uint64)) & np.uint64(0xFFFF0000)).astype(np.uint32)

    def rn_bf16(x32):
        b = x32.view(np.uint32).astype(np.uint64)
        return (((b + np.uint64(0x7FFF) + ((b >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)).astype(np.uint32)

    # ---- second word: p <- SR(p + shift, r23 low), shift <- SR(bf16(p_old - p_new) + shift, r23 high)
    p, m, v, sh = p0.clone(), z.clone(), v0.clone(), s0.clone()
    ops.adamw_bf16(p, z, m, v, sh, lr=1e-3, step=step, seed_offset=so)
    pf, sf = p0.float().cpu().numpy(), s0.float().cpu().numpy()
    p_pred = sr((sf + pf).astype(np.float32), r23 & np.uint32(0xFFFF))
    assert np.array_equal(bits(p), p_pred), f"{int((bits(p) != p_pred).sum())} parameters differ from the numpy prediction"
    d = rn_bf16((pf - p_pred.view(np.float32)).astype(np.float32)).view(np.float32)
    s_pred = sr((d + sf).astype(np.float32), r23 >> np.uint32(16))
    assert np.array_equal(bits(sh), s_pred), f"{int((bits(sh) != s_pred).sum())} shift values differ from the numpy prediction"
    # ---- first word, low half: exp_avg <- SR(g + (1 - b1) * bf16(b1 * exp_avg), r01 low)   (operand order as written)
    gr = (torch.randn(n, device="cuda", generator=g) * 1e-2).to(bf16)
    m0 = (torch.randn(n, device="cuda", generator=g) * 1e-2).to(bf16)
    p, m, v, sh = p0.clone(), m0.clone(), v0.clone(), s0.clone()
    ops.adamw_bf16(p, gr.clone(), m, v, sh, lr=1e-3, step=step, seed_offset=so)
    b1, omb1 = np.float32(0.9), np.float32(1.0 - 0.9)
    m1 = rn_bf16((m0.float().cpu().numpy().astype(np.float64) * np.float64(b1)).astype(np.float32)).view(np.float32)
    # fma = one rounding of the exact a*b + c: the 32-bit product plus an 8-bit addend is exact in 64-bit extended precision
    mr = (np.longdouble(omb1) * m1.astype(np.longdouble) + gr.float().cpu().numpy().astype(np.longdouble)).astype(np.float32)
    m_pred = sr(mr, r01 & np.uint32(0xFFFF))
    assert np.array_equal(bits(m), m_pred), f"{int((bits(m) != m_pred).sum())} exp_avg values differ from the numpy prediction"
