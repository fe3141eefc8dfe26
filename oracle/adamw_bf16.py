"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's AdamWBF16 update, op by op:
  src/training/optimizers/adamw_bfloat16/__init__.py:150-197 (`_make_step`), :118-128 (deferred weight decay)
  src/training/optimizers/adamw_bfloat16/stochastic/__init__.py:46-124 (copy_/add_/addcdiv_stochastic_)
The 16 random low bits of every stochastic rounding are injectable (`rand16`), so the restatement can be pinned
bit-for-bit against the reference's own functions run with a patched `torch.randint_like`
(tests/golden/make_adamw_bf16_golden.py -> tests/golden/adamw_bf16_golden.pt; checked by tests/test_oracle_adamw_bf16.py).

`as_written=True` keeps the operand order of `add_stochastic_` as the reference wrote it
(result = other + alpha * input), i.e. exp_avg <- SR(grad + (1-beta1) * beta1 * exp_avg).
"""
from __future__ import annotations

import torch

bf16 = torch.bfloat16


def _rn(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest-even bf16 -> fp32 (what every torch bf16 op does to its fp32 intermediate)."""
    return x.to(bf16).float()


def _sr(x: torch.Tensor, rand16: torch.Tensor) -> torch.Tensor:
    """stochastic/__init__.py:46-71."""
    bits = x.contiguous().view(torch.int32) + rand16.to(torch.int32)
    return (bits & -65536).view(torch.float32)


def make_step(p, g, m, v, shift, *, beta1, beta2, step, lr, eps, rand16, as_written=True, clip=1.0):
    """All state tensors fp32 holding bf16-representable values; returns new (p, m, v, shift).
    rand16: int tensor [4, n] of values in [0, 65535] for the four stochastic roundings (m, shift, p, shift)."""
    f32 = torch.float32
    g = g if clip == 1.0 else _rn(g * torch.tensor(clip, dtype=f32))
    m1 = _rn(m * torch.tensor(beta1, dtype=f32))
    # result.add_(_input, alpha=alpha): torch's add-with-alpha is a fused multiply-add
    mr = torch.add(g, m1, alpha=1.0 - beta1) if as_written else torch.add(m1, g, alpha=1.0 - beta1)
    m2 = _sr(mr, rand16[0])
    v1 = _rn(v * torch.tensor(beta2, dtype=f32))
    v2 = _rn(v1 + torch.tensor(1.0 - beta2, dtype=f32) * g * g)
    den = _rn(_rn(v2.sqrt()) + torch.tensor(eps, dtype=f32))
    value = torch.tensor(-lr * (1 - beta2 ** step) ** 0.5, dtype=f32)
    s1 = _sr(shift + value * m2 / den, rand16[1])
    p1 = _sr(s1 + p, rand16[2])
    d = _rn(p - p1)
    s2 = _sr(d + s1, rand16[3])
    return p1, m2, v2, s2


def apply_decay(shift, p, decay):
    """shift.add_(p, alpha=-decay) on bf16 tensors (adamw_bfloat16/__init__.py:191-192).  torch's CPU kernel casts
    `alpha` to the tensor dtype first (bf16) — that is what the golden vectors capture; torch's CUDA kernel keeps it in
    fp32.  The difference is one bf16 ulp of a term that fires once per ~1e6 steps at the default lr."""
    alpha = float(torch.tensor(-decay, dtype=torch.float32).to(bf16))
    return _rn(torch.add(shift, p, alpha=alpha))
