#!/bin/bash
# Build tuning variants of libptb200.so into build/variants/<name>.so:  tools/build_variants.sh name1="-DFOO=1 -DBAR=2" name2="..."
set -e
cd "$(dirname "$0")/../path_trace_golang_b200/csrc"
mkdir -p ../../build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC"
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  tmp=$(mktemp -d)
  nvcc $FLAGS $defs -Xptxas -v -c -o $tmp/integrator.o integrator.cu 2>&1 | grep -A2 "integrate_wf_kernelILb0" | grep -E "Used|spill" | tr '\n' ' '
  nvcc $FLAGS -fmad=false -c -o $tmp/primary_fp64.o primary_fp64.cu
  nvcc $FLAGS $defs -c -o $tmp/api.o api.cu
  nvcc $FLAGS -x cu -c -o $tmp/bvh.o bvh.cpp
  nvcc $FLAGS -x cu -c -o $tmp/host_scene.o host/scene.cpp
  nvcc $FLAGS -x cu -c -o $tmp/host_engine.o host/engine.cpp
  nvcc $ARCH -shared -o ../../build/variants/$name.so $tmp/*.o
  rm -rf $tmp
  echo; echo "built $name ($defs)"
done
