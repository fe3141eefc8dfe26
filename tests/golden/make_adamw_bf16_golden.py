"""Generate tests/golden/adamw_bf16_golden.pt by running the REFERENCE's own `_make_step`
(src/training/optimizers/adamw_bfloat16/__init__.py:150-197) on CPU, with `torch.randint_like` patched to return a
fixed 16-bit pattern so that the stochastic roundings are reproducible (0 = truncate, 65535 = always round away, and a
seeded pseudo-random pattern that is stored in the fixture).  Build container only (needs /root/reference):
    python tests/golden/make_adamw_bf16_golden.py
"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_schedule_golden import REF, _stub_modules  # noqa: E402


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    import src.training.optimizers.adamw_bfloat16 as ab
    import src.training.optimizers.adamw_bfloat16.stochastic as st
    os.chdir(cwd)

    n, steps = 1024, 3
    bf16 = torch.bfloat16
    hyper = dict(beta1=0.9, beta2=0.999, lr=1e-3, eps=1e-8)
    out = {"reference_commit": "6083befd", "torch": torch.__version__, "hyper": hyper, "n": n, "steps": steps, "cases": {}}
    real_randint_like = torch.randint_like
    for mode in ("zero", "ffff", "random"):
        gen = torch.Generator().manual_seed(1234)
        p = (torch.randn(n, generator=gen) * 0.05).to(bf16)
        grads = [(torch.randn(n, generator=gen) * 10 ** torch.empty(n).uniform_(-4, 0, generator=gen)).to(bf16)
                 for _ in range(steps)]
        m = torch.zeros(n, dtype=bf16); v = torch.zeros(n, dtype=bf16); sh = torch.zeros(n, dtype=bf16)
        rgen = torch.Generator().manual_seed(99)
        used = []

        def fake_randint_like(source, dtype=None, low=0, high=None, **kw):
            if mode == "zero":
                r = torch.zeros(source.shape, dtype=torch.int32)
            elif mode == "ffff":
                r = torch.full(source.shape, 65535, dtype=torch.int32)
            else:
                r = torch.randint(0, 1 << 16, source.shape, generator=rgen, dtype=torch.int32)
            used.append(r.clone())
            return r

        case = {"p0": p.clone(), "grads": [g.clone() for g in grads], "states": []}
        torch.randint_like = fake_randint_like
        try:
            for k in range(steps):
                ab._make_step(grads[k].clone(), p, sh, m, v, beta1=hyper["beta1"], beta2=hyper["beta2"], step=float(k + 1),
                              lr=hyper["lr"], eps=hyper["eps"], decay_this_iteration=0.0, zero_grad=False)
                case["states"].append({"p": p.clone(), "m": m.clone(), "v": v.clone(), "shift": sh.clone()})
            # one deferred-decay application on top of the last state
            ab._make_step(grads[0].clone(), p, sh, m, v, beta1=hyper["beta1"], beta2=hyper["beta2"], step=float(steps + 1),
                          lr=hyper["lr"], eps=hyper["eps"], decay_this_iteration=0.0075, zero_grad=False)
            case["decay_state"] = {"p": p.clone(), "m": m.clone(), "v": v.clone(), "shift": sh.clone(), "decay": 0.0075}
        finally:
            torch.randint_like = real_randint_like
        assert len(used) == 4 * (steps + 1)
        case["rand16"] = torch.stack(used).view(steps + 1, 4, n).to(torch.int32)
        out["cases"][mode] = case
    # copy_stochastic_ known answers (stochastic/__init__.py:46-71)
    src = torch.tensor([1.0, 1.00390625, -1.00390625, 3.14159265, 1e-30, 65504.0], dtype=torch.float32)
    tgt = torch.empty_like(src, dtype=bf16)
    ks = {}
    for name, val in (("zero", 0), ("ffff", 65535), ("8000", 0x8000)):
        torch.randint_like = lambda source, dtype=None, low=0, high=None, _v=val, **kw: torch.full(source.shape, _v, dtype=torch.int32)
        try:
            st.copy_stochastic_(tgt, src)
        finally:
            torch.randint_like = real_randint_like
        ks[name] = tgt.clone()
    out["copy_stochastic"] = {"src": src, "out": ks}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adamw_bf16_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
