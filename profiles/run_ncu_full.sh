#!/bin/bash
# One `ncu --set full` capture per hot kernel (1 GPU; numbers printed by a run under ncu are never bench values).
# ncu serialises kernels, so the library takes its plain launch order (it sees the profiler's injection library in the
# environment; PIMDB_EXCH_SERIAL=1 forces it): factor tiles, then recurrences, then exterior forces.
# usage: profiles/run_ncu_full.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
for K in k_pair_tiles k_exch_recur_cluster k_exch_coeff_tiles k_integrate k_exch_forces k_assemble; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_$K \
      env PIMDB_EXCH_SERIAL=1 python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_${TAG}_$K.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${K}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | tail -20
