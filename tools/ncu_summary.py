"""Compact per-launch summary of an `ncu --set full` report: the counters the roofline discussion uses.
usage: ncu_summary.py report.ncu-rep "<header comment>" > profiles/<name>.csv      (runs `ncu -i ... --page raw --csv` here, no GPU)"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_of_active"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__cycles_active.avg", "smsp_cycles_active"), ("sm__cycles_elapsed.max", "cycles_elapsed")]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# ncu --set full --clock-control none (every launch replayed ~40x from a cold cache: times are NOT bench values)")
print("kernel," + ",".join(f"{short}[{units[idx[full]]}]" if full in idx and units[idx[full]] else short for full, short in COLS))
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0].replace("cc::", "").replace("void ", "").strip()
    print(name + "," + ",".join(r[idx[full]].replace(",", "") if full in idx else "" for full, _ in COLS))
