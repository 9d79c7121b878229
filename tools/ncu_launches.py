"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ms and share."""
import csv, sys
from collections import defaultdict
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:48]
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"]) / 1e6
ours = {k: v for k, v in agg.items() if "at::" not in k}
tot = sum(v[1] for v in ours.values())
print(f"{'kernel':48s} {'launches':>8s} {'total_ms':>12s} {'share':>7s}")
for k, v in sorted(ours.items(), key=lambda x: -x[1][1]):
    print(f"{k:48s} {v[0]:8d} {v[1]:12.3f} {100 * v[1] / tot:6.1f}%")
print(f"(torch helper kernels excluded: {sum(v[0] for k, v in agg.items() if 'at::' in k)} launches)")
