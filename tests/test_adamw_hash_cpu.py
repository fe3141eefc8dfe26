"""Random bits of the AdamWBF16 kernel (csrc/optim.cu: lowbias32 / rand64), restated in numpy and checked statistically on CPU.

The reference draws its stochastic-rounding bits with `torch.randint_like(..., 0, 1 << 16)` per rounding
(src/training/optimizers/adamw_bfloat16/stochastic/__init__.py:46-71): four independent 16-bit words per element and step.
The kernel replaces the generator by a counter hash — one full integer mixer of (element, keys) for the first 32-bit word and one
multiply / xor-shift round of that word under the second key for the other — because the kernel is HBM-bound only if the bits
cost ~20 integer instructions per element.  What stochastic rounding needs from them: each 16-bit field uniform, the four
fields of one element not correlated with each other, neighbouring elements and different steps not correlated.  The GPU test
(tests/test_gpu_optim.py) measures rounding bias and cross-field correlation through the kernel itself; this file pins the
construction (constants included) without a GPU.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def lowbias32(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(17)
    x = (x * np.uint64(0xED5AD4BB)) & M32
    x ^= x >> np.uint64(11)
    x = (x * np.uint64(0xAC4C1B51)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x31848BAB)) & M32
    x ^= x >> np.uint64(14)
    return x


def keys(seed, step):
    s = np.uint64(seed)
    a = lowbias32(np.array([(int(s) >> 32) + 0x9E3779B9 & 0xFFFFFFFF], dtype=np.uint64))[0]
    b = lowbias32(np.array([(step * 0x85EBCA6B + (step >> 32)) & 0xFFFFFFFF], dtype=np.uint64))[0]
    kmix = lowbias32(np.array([(int(s) & 0xFFFFFFFF) ^ int(a) ^ int(b)], dtype=np.uint64))[0]
    kb = int(lowbias32(np.array([int(kmix) ^ 0x68E31DA4], dtype=np.uint64))[0]) | 1
    return int(kmix), kb


def rand64(e, ka, kb):
    lo = e & M32
    hi = ((e >> np.uint64(32)) * np.uint64(0xC2B2AE35)) & M32
    r01 = lowbias32(lo ^ np.uint64(ka) ^ hi)
    t = ((r01 ^ np.uint64(kb)) * np.uint64(0x9E3779B1)) & M32
    r23 = t ^ (t >> np.uint64(15))
    return r01, r23


def fields(e, ka, kb):
    r01, r23 = rand64(e, ka, kb)
    return [(r01 & np.uint64(0xFFFF)).astype(np.float64), (r01 >> np.uint64(16)).astype(np.float64),
            (r23 & np.uint64(0xFFFF)).astype(np.float64), (r23 >> np.uint64(16)).astype(np.float64)]


def corr(a, b):
    a, b = a - a.mean(), b - b.mean()
    return float((a * b).mean() / (a.std() * b.std()))


N = 1 << 20


def test_each_16_bit_field_is_uniform():
    ka, kb = keys(77, 5)
    f = fields(np.arange(N, dtype=np.uint64), ka, kb)
    for k, x in enumerate(f):
        # mean of U{0..65535} = 32767.5, sigma of the mean over N draws = 18918 / sqrt(N) ~ 18.5
        assert abs(x.mean() - 32767.5) < 5 * 18918 / np.sqrt(N), (k, x.mean())
        h = np.bincount((x.astype(np.int64) >> 8), minlength=256)      # 256 buckets, expected N / 256 each
        chi2 = float(((h - N / 256) ** 2 / (N / 256)).sum())
        assert chi2 < 255 + 6 * np.sqrt(2 * 255), (k, chi2)            # chi-square with 255 dof: mean 255, sigma 22.6
        for bit in range(16):                                           # every bit balanced
            p = float(((x.astype(np.int64) >> bit) & 1).mean())
            assert abs(p - 0.5) < 5 * 0.5 / np.sqrt(N), (k, bit, p)


def test_fields_of_one_element_neighbouring_elements_and_steps_are_uncorrelated():
    ka, kb = keys(77, 5)
    e = np.arange(N, dtype=np.uint64)
    f = fields(e, ka, kb)
    lim = 5 / np.sqrt(N)
    for i in range(4):
        for j in range(i + 1, 4):
            assert abs(corr(f[i], f[j])) < lim, (i, j, corr(f[i], f[j]))
            # the ROUNDING DECISION is a threshold on the field: correlate the events "field above a threshold" too
            for thr in (8192, 32768, 57344):
                assert abs(corr((f[i] > thr).astype(np.float64), (f[j] > thr).astype(np.float64))) < lim, (i, j, thr)
    g = fields(e + np.uint64(1), ka, kb)                                # next element
    for i in range(4):
        for j in range(4):
            assert abs(corr(f[i], g[j])) < lim, ("neighbour", i, j)
    ka2, kb2 = keys(77, 6)                                              # next optimizer step
    h = fields(e, ka2, kb2)
    for i in range(4):
        for j in range(4):
            assert abs(corr(f[i], h[j])) < lim, ("step", i, j)
    ka3, kb3 = keys(78, 5)                                              # another seed
    s = fields(e, ka3, kb3)
    for i in range(4):
        assert abs(corr(f[i], s[i])) < lim, ("seed", i)


def test_mixer_avalanche_and_high_index_word():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 1 << 32, size=1 << 16, dtype=np.uint64)
    y = lowbias32(x)
    worst = 0.0
    for bit in range(32):
        d = y ^ lowbias32(x ^ np.uint64(1 << bit))
        for ob in range(32):
            p = float(((d >> np.uint64(ob)) & np.uint64(1)).mean())
            worst = max(worst, abs(p - 0.5))
    assert worst < 0.02, worst                                          # every output bit flips with p = 0.5 +- 0.02
    # elements beyond 2^32 (the index's high word is folded in through an odd multiplier): not a replay of the low range
    ka, kb = keys(3, 1)
    lo = fields(np.arange(N, dtype=np.uint64), ka, kb)
    hi = fields(np.arange(N, dtype=np.uint64) + (np.uint64(1) << np.uint64(32)), ka, kb)
    for i in range(4):
        assert abs(corr(lo[i], hi[i])) < 5 / np.sqrt(N)
        assert float((lo[i] == hi[i]).mean()) < 1e-3
