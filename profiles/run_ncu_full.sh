#!/bin/bash
# One `ncu --set full` capture per hot kernel (1 GPU; numbers printed by a run under ncu are never bench values).
# usage: profiles/run_ncu_full.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
for K in k_pair_tiles k_exch_recur k_integrate k_exch_forces k_assemble; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_$K \
      python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_${TAG}_$K.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${K}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | tail -20
