#!/bin/bash
# One `ncu --set full` capture per hot kernel (1 GPU; numbers printed by a run under ncu are never bench values).
# ncu serialises kernels; every dependency of the step is a real one (stream order / programmatic dependent launch), so the
# library runs unchanged under the profiler -- only the overlap between the exchange chain and the pair tiles is gone.
# usage: profiles/run_ncu_full.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
for K in k_pair_tiles k_exch_recur_cluster k_exch_coeff_tiles k_integrate k_exch_forces; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_$K \
      python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/ncu_${TAG}_$K.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${K}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$K.ncu-rep      # (gpurun brings back at most 64 MiB: the CSV export is what is kept)
done
# C5 (N = 8192, P = 256): the whole exchange chain once, for its DRAM traffic
ncu --set full --clock-control none -k regex:k_exch -s 12 -c 6 -f -o gpurun_out/prof_${TAG}_c5_exchange \
    python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/ncu_${TAG}_c5.log 2>&1
ncu -i gpurun_out/prof_${TAG}_c5_exchange.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_c5_exchange_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${TAG}_c5_exchange.ncu-rep
ls -la gpurun_out/ | tail -20
