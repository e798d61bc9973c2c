#!/bin/bash
# experiment: warps per block of the exterior-force kernel (variants fw1 / fw2) x shared-memory padding of the recurrence blocks
cd "$(dirname "$0")/.."
for v in main fw2 fw1; do
  if [ $v = main ]; then unset PIMDB200_LIB; else export PIMDB200_LIB=$PWD/pimd_b_b200/_variants/lib_$v.so; fi
  for kb in 0 218; do echo "== $v PIMDB_RECUR_SMEM_KB=$kb"; PIMDB_RECUR_SMEM_KB=$kb python profiles/step_timeline.py c3 2>&1 | tail -11 | head -8 | tail -5; PIMDB_RECUR_SMEM_KB=$kb python profiles/pair_probe.py c3 | cut -c90-250; done
done
