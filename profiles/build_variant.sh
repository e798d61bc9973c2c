#!/bin/bash
# build a variant of the library with extra -D flags for one source file (default pair_forces.cu):  profiles/build_variant.sh name -DPIMDB_PAIR_WARPS=8 ...
# -> pimd_b_b200/_variants/lib_<name>.so (select with PIMDB200_LIB, see profiles/variants.sh)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
src=${SRC:-pair_forces}      # which source file gets the flags (SRC=exchange profiles/build_variant.sh ...)
mkdir -p pimd_b_b200/_variants /tmp/pv_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c pimd_b_b200/csrc/$src.cu -o /tmp/pv_$name/$src.o
objs=$(ls pimd_b_b200/_build/*.o | grep -v "/$src.o")
nvcc -shared -o pimd_b_b200/_variants/lib_$name.so /tmp/pv_$name/$src.o $objs -gencode arch=compute_100a,code=sm_100a -lcudart
echo built pimd_b_b200/_variants/lib_$name.so
