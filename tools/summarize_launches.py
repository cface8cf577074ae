"""Summarise an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]) per kernel.
usage: summarize_launches.py raw.csv "<header comment>" > summary.csv"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
t, rd, wr, n = defaultdict(float), defaultdict(float), defaultdict(float), defaultdict(set)
scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}
bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
for r in rows:
    name = r[4].split("(")[0].replace("cc::", "").strip()
    metric, unit, val = r[12], r[13], float(r[14].replace(",", ""))
    n[name].add(r[0])
    if metric == "gpu__time_duration.sum":
        t[name] += val * scale.get(unit, 1.0)
    elif metric == "dram__bytes_read.sum":
        rd[name] += val * bscale.get(unit, 1e-6)
    elif metric == "dram__bytes_write.sum":
        wr[name] += val * bscale.get(unit, 1e-6)
tot = sum(t.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# gpu__time_duration.sum (+ dram bytes) per launch, --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)")
print(f"# total {tot / 1e3:.3f} ms over {sum(len(v) for v in n.values())} launches")
print("kernel,launches,total_us,share,dram_read_MB,dram_write_MB,dram_MB_per_launch")
for k in sorted(t, key=lambda k: -t[k]):
    L = len(n[k])
    print(f"{k},{L},{t[k]:.1f},{t[k] / tot:.4f},{rd[k]:.1f},{wr[k]:.1f},{(rd[k] + wr[k]) / L:.2f}")
