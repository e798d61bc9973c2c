#!/bin/bash
# Kernel-variant shoot-out on one B200: every library under pimd_b_b200/_variants/ (same ABI, differently compiled
# pair_forces.cu) runs the force-parity tests and the C3 / C4 timing probe.  usage: profiles/variants.sh [names...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/variants.txt
: > $out
names="$@"
[ -z "$names" ] && names=$(ls pimd_b_b200/_variants/ | sed 's/^lib_//; s/\.so$//')
for n in main $names; do
  if [ "$n" = main ]; then unset PIMDB200_LIB; else export PIMDB200_LIB=$PWD/pimd_b_b200/_variants/lib_$n.so; fi
  echo "== $n" >> $out
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "forces" 2>&1 | tail -1 >> $out
  timeout 120 python profiles/quick_pair.py c3 1000 2>&1 | tail -1 >> $out
  timeout 120 python profiles/quick_pair.py c4 100 2>&1 | tail -1 >> $out
done
cat $out
