#!/usr/bin/env python
"""Turn the raw ncu exports in gpurun_out/ (profiles/run_ncu_full.sh, profiles/run_launches.sh) into the tracked
summaries profiles/<tag>_ncu_summary.{json,md}.   python profiles/summarize_ncu.py r01"""
import csv
import json
import sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
KERNELS = ["k_pair_tiles", "k_exch_recur_cluster", "k_exch_coeff_tiles", "k_exch_forces", "k_integrate", "k_assemble"]   # (k_assemble: round 1 only)
METRICS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fp64.sum": "fp64_warp_insts",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "cycles",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
}
UNIT = {"us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3,
        "Mbyte": 1e6, "Gbyte": 1e9}


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(h):
            continue
        d = {"kernel": r[h.index("Kernel Name")].split("(")[0]}
        for m, name in METRICS.items():
            if m in h:
                i = h.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[name] = v * UNIT.get(units[i], 1.0)
        out.append(d)
    return out


summary = {"tag": tag, "note": "ncu --set full --clock-control none, launches 7-8 of `bench.py --steps 12 --warmup 5` "
           "(workload C3); per launch; caches flushed and kernels serialised by ncu, so small kernels read slower "
           "than in the running step", "kernels": {}}
for k in KERNELS:
    p = OUT / f"prof_{tag}_{k}_raw.csv"
    if not p.exists():
        continue
    launches = parse(p)
    if launches:
        d = launches[-1]
        d["dram_traffic_bytes"] = d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0)
        summary["kernels"][k] = d
ls = OUT / f"launches_{tag}_summary.txt"
if ls.exists():
    summary["launch_list"] = ls.read_text().strip().splitlines()
(ROOT / "profiles" / f"{tag}_ncu_summary.json").write_text(json.dumps(summary, indent=1))
lines = [f"# ncu summary {tag} (workload C3: He-4 Aziz, N=512, P=64, one B200)", "", summary["note"], "",
         "| kernel | duration us | grid x block | regs | FP64 pipe % | warps active % | issue active % | DRAM read | DRAM write |",
         "|---|---|---|---|---|---|---|---|---|"]
for k, d in summary["kernels"].items():
    lines.append(f"| {d['kernel']} | {d.get('duration_us', 0):.1f} | {int(d.get('grid', 0))} x {int(d.get('block', 0))} | "
                 f"{int(d.get('registers', 0))} | {d.get('fp64_pipe_pct', 0):.1f} | {d.get('warps_active_pct', 0):.1f} | "
                 f"{d.get('issue_active_pct', 0):.1f} | {d.get('dram_read_bytes', 0) / 1e6:.2f} MB | {d.get('dram_write_bytes', 0) / 1e6:.2f} MB |")
if "launch_list" in summary:
    lines += ["", "## launch list (gpu__time_duration.sum, 120 launches of the timed region; shares, not absolutes)", "", "```"]
    lines += summary["launch_list"] + ["```"]
c5 = OUT / f"prof_{tag}_c5_exchange_raw.csv"
if c5.exists():
    rows = parse(c5)
    summary["c5_exchange_chain"] = rows
    lines += ["", "## C5 (N = 8192, P = 256): the kernels of one exchange chain, ncu --set full", "",
              "| kernel | duration us | grid x block | regs | DRAM read | DRAM write | FP64 pipe % |", "|---|---|---|---|---|---|---|"]
    tot = 0.0
    for d in rows:
        tot += d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)
        lines.append(f"| {d['kernel']} | {d.get('duration_us', 0):.1f} | {int(d.get('grid', 0))} x {int(d.get('block', 0))} | {int(d.get('registers', 0))} | "
                     f"{d.get('dram_read_bytes', 0) / 1e6:.1f} MB | {d.get('dram_write_bytes', 0) / 1e6:.1f} MB | {d.get('fp64_pipe_pct', 0):.1f} |")
    lines.append(f"")
    lines.append(f"DRAM traffic of the captured kernels: {tot / 1e6:.0f} MB")
    (ROOT / "profiles" / f"{tag}_ncu_summary.json").write_text(json.dumps(summary, indent=1))
(ROOT / "profiles" / f"{tag}_ncu_summary.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
