"""Summarise ncu artefacts into small text/JSON files for profiles/ (run in the build container, no GPU)."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m in METRICS:
            if m in hdr:
                d[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        res.append(d)
    return res


def launches(csv_path):
    rows = [r for r in csv.reader(open(csv_path)) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(csv_path)) if r and r[0] == "ID")
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)   # -> us
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v[1] for v in agg.values())
    return [{"kernel": k[:90], "launches": v[0], "total_us": round(v[1], 2), "avg_us": round(v[1] / v[0], 2),
             "share": round(v[1] / tot, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    print(json.dumps(raw_page(path) if kind == "raw" else launches(path), indent=1))
