#!/bin/bash
# A/B of the pair-tile kernel variants under pimd_b_b200/_variants/ (+ env switches) on one B200
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/pair_ab.txt
: > $out
run() { # name, env...
  local n=$1; shift
  if [ "$n" = main ]; then unset PIMDB200_LIB; else export PIMDB200_LIB=$PWD/pimd_b_b200/_variants/lib_$n.so; fi
  for w in c3 c4; do env "$@" timeout 200 python profiles/pair_probe.py $w 2>&1 | tail -1 >> $out; done
}
run main A=1
run main PIMDB_PAIR_SPLIT=2

for v in $(ls pimd_b_b200/_variants/ | sed 's/^lib_//; s/\.so$//'); do run $v A=1; done
unset PIMDB200_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_more.py -q -x -m gpu 2>&1 | tail -2 >> $out
cat $out
