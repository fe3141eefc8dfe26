"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel."""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i:]))
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"])[:64]
    t = float(r["Metric Value"].replace(",", ""))
    tot[n] += t
    cnt[n] += 1
T = sum(tot.values())
print(f"total {T / 1e6:.2f} ms over {len(rows)} launches")
for n, t in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{t / 1e6:9.2f} ms {100 * t / T:5.1f}% {cnt[n]:5d}  {n}")
