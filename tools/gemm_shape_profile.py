"""Join the B2_GEMM_LOG shape log with an ncu launch list: time per GEMM / conv shape class.
usage: python tools/gemm_shape_profile.py launches.csv shapes.log [aggregate.json]
(the optional third argument receives {launches, ms, tflops} over ALL gemm2 launches of the profiled step — bench.py
reports it next to the best-shape roofline fraction)"""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = [r for r in csv.DictReader(lines[i:]) if "gemm2_kernel" in r["Kernel Name"]]
shapes = [l.strip() for l in open(sys.argv[2]) if l.startswith("B2GEMM")]
# the log covers warm-up steps too: the profiled step is the LAST len(rows) entries
shapes = shapes[-len(rows):]
assert len(shapes) == len(rows), (len(shapes), len(rows))
tot, cnt, fl = collections.Counter(), collections.Counter(), collections.Counter()
for r, s in zip(rows, shapes):
    t = float(r["Metric Value"].replace(",", ""))
    m = re.search(r"M=(\d+) N=(\d+) K=(\d+) a_mn=(\d) b_mn=(\d) conv=(\d) BN=(\d+) splits=(\d+)", s)
    M, N, K = int(m.group(1)), int(m.group(2)), int(m.group(3))
    key = f"M={M:6d} N={N:6d} K={K:6d} a{m.group(4)}b{m.group(5)} conv{m.group(6)} BN={m.group(7):>3s} s={m.group(8):>2s}"
    tot[key] += t
    cnt[key] += 1
    fl[key] += 2.0 * M * N * K
T = sum(tot.values())
print(f"gemm2 total {T / 1e6:.2f} ms over {len(rows)} launches, {sum(fl.values()) / T / 1e3:.0f} TF/s average")
if len(sys.argv) > 3:
    import json
    json.dump({"launches": len(rows), "ms": round(T / 1e6, 3), "tflops": round(sum(fl.values()) / T / 1e3, 1),
               "source": "ncu gpu__time_duration per launch (serialised, one training step bs=4 1024^2) joined with B2_GEMM_LOG"},
              open(sys.argv[3], "w"))
for k, t in tot.most_common(45):
    print(f"{t / 1e6:7.2f} ms {100 * t / T:5.1f}% {cnt[k]:4d} x {t / cnt[k] / 1e3:7.1f} us {fl[k] / t / 1e3:6.0f} TF/s  {k}")
