#!/bin/bash
# ncu launch list of the bench command (cold-cache, serialised: compare SHARES, not absolutes)
# usage: profiles/run_launches.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr+1:]:
    name = r[ki].split("(")[0]; v = float(r[vi].replace(",", "")); u = r[ui]
    if u in ("ns", "nsecond"): v /= 1e3
    elif u in ("ms", "msecond"): v *= 1e3
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
with open("gpurun_out/launches_${TAG}_summary.txt", "w") as f:
    f.write("kernel, launches, avg_us, total_us, share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k}, {len(v)}, {sum(v)/len(v):.2f}, {sum(v):.1f}, {sum(v)/tot:.3f}\n")
print(open("gpurun_out/launches_${TAG}_summary.txt").read())
PY
