"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel launch count, total ms, share of the listed time and (when captured) DRAM bytes."""
import csv, sys
from collections import defaultdict
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
agg = defaultdict(lambda: {"ids": set(), "ms": 0.0, "rd": 0.0, "wr": 0.0})
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:48]
    a = agg[k]
    a["ids"].add(r["ID"])
    val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    name = r["Metric Name"]
    if name.startswith("gpu__time_duration"):
        a["ms"] += val
    elif name.startswith("dram__bytes_read"):
        a["rd"] += val
    elif name.startswith("dram__bytes_write"):
        a["wr"] += val
ours = {k: v for k, v in agg.items() if "at::" not in k}
tot = sum(v["ms"] for v in ours.values())
have_dram = any(v["rd"] or v["wr"] for v in ours.values())
print(f"{'kernel':48s} {'launches':>8s} {'total_ms':>12s} {'share':>7s}" + (f" {'dram_rd_MB':>11s} {'dram_wr_MB':>11s}" if have_dram else ""))
for k, v in sorted(ours.items(), key=lambda x: -x[1]["ms"]):
    print(f"{k:48s} {len(v['ids']):8d} {v['ms']:12.3f} {100 * v['ms'] / tot:6.1f}%" + (f" {v['rd'] / 1e6:11.1f} {v['wr'] / 1e6:11.1f}" if have_dram else ""))
if have_dram:
    print(f"{'(all listed kernels)':48s} {'':8s} {tot:12.3f} {'':7s} {sum(v['rd'] for v in ours.values()) / 1e6:11.1f} {sum(v['wr'] for v in ours.values()) / 1e6:11.1f}")
print(f"(torch helper kernels excluded: {sum(len(v['ids']) for k, v in agg.items() if 'at::' in k)} launches)")
