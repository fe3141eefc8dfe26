"""poly_exp2 (csrc/tc.cuh): 2^x on the FMA / integer pipes, used (opt-in) for a quarter of the softmax exponentials of the
attention forward kernel.  The device function is restated here in numpy float32 — same constants, same operation order,
same integer exponent arithmetic — and checked against 2^x in float64: the polynomial + exponent splice must stay far
below the bf16 rounding (2^-9) of the probabilities it produces, including negative integer parts, exact integers, the
clamp at -126 and the masked-logit case (-inf)."""
import numpy as np

C3, C2, C1, C0 = np.float32(0.05517132), np.float32(0.24261054), np.float32(0.69326097), np.float32(0.99992812)
MAGIC = np.float32(12582912.0)  # 1.5 * 2^23


def poly_exp2(x):
    x = np.maximum(np.asarray(x, np.float32), np.float32(-126.0))
    xr = (x + MAGIC).astype(np.float32)
    f = (x - (xr - MAGIC).astype(np.float32)).astype(np.float32)
    # fmaf: one rounding per step (float64 product + sum, rounded to float32)
    p = (np.float64(C3) * f + np.float64(C2)).astype(np.float32)
    p = (np.float64(p) * f + np.float64(C1)).astype(np.float32)
    p = (np.float64(p) * f + np.float64(C0)).astype(np.float32)
    bits = p.view(np.int32) + (xr.view(np.int32) << np.int32(23))
    return bits.astype(np.int32).view(np.float32)


def test_relative_error_far_below_bf16_rounding():
    x = np.concatenate([np.linspace(-100, 9, 200001), np.arange(-120, 9, 1.0), np.arange(-120, 9, 1.0) + 0.5,
                        np.arange(-120, 9, 1.0) - 0.49999]).astype(np.float32)
    got = poly_exp2(x).astype(np.float64)
    want = np.exp2(x.astype(np.float64))
    rel = np.abs(got / want - 1)
    assert rel.max() < 1e-4 < 2.0 ** -9 / 10, rel.max()
    assert (got > 0).all()


def test_split_is_round_to_nearest_and_exponent_splice_is_exact_on_integers():
    n = np.arange(-125, 10).astype(np.float32)  # at -126 the result is a denormal (the splice drops the implicit one)
    got = poly_exp2(n).astype(np.float64)
    assert np.allclose(got / np.exp2(n.astype(np.float64)), float(C0), rtol=1e-7)  # f = 0 -> p = c0 exactly


def test_clamp_and_masked_logits():
    got = poly_exp2(np.array([-np.inf, -1e30, -500.0, -126.0], np.float32))
    assert (got >= 0).all() and (got < 2e-38).all()  # masked keys contribute nothing measurable to l or to P V
