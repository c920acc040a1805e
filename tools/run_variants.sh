#!/bin/bash
# On the GPU box: time every build/variants/*.so on a few workloads.
for so in build/variants/*.so; do
  echo "== $(basename $so .so)"
  for wl in ${@:-C3 C2}; do
    PTB200_LIB=$PWD/$so timeout 60 python tools/profile_kernel.py $wl 64 3 | tail -1
  done
done
