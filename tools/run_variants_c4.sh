#!/bin/bash
# On the GPU box: time every build/variants/*.so on the mesh workload.
for so in build/variants/*.so; do
  echo "== $(basename $so .so)"
  PTB200_LIB=$PWD/$so timeout 120 python tools/profile_kernel.py ${1:-C4_1M} 16 3 | tail -1
done
