#!/bin/bash
# ncu --set full of the pair-tile kernel alone (profiles/pair_only.py), main library and optional variants
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for v in main "$@"; do
  if [ $v = main ]; then unset PIMDB200_LIB; else export PIMDB200_LIB=$PWD/pimd_b_b200/_variants/lib_$v.so; fi
  ncu --set full --clock-control none --import-source on -k regex:k_pair_tiles -s 4 -c 1 -f -o gpurun_out/prof_pair_$v python profiles/pair_only.py c3 8 > gpurun_out/ncu_pair_$v.log 2>&1
  ncu -i gpurun_out/prof_pair_$v.ncu-rep --page raw --csv > gpurun_out/prof_pair_${v}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_pair_$v.ncu-rep --page source --csv > gpurun_out/prof_pair_${v}_source.csv 2>/dev/null
  rm -f gpurun_out/prof_pair_$v.ncu-rep
done
