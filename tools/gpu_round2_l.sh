#!/bin/bash
# Round-2 GPU session L: 5 resident CTAs per SM (48 registers) against 4 (56 registers).
set -u
mkdir -p gpurun_out
{
for so in default b5; do echo "== $so"; for wl in C3 C2 C5 C1; do spp=64; [ $wl = C1 ] && spp=16; PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $wl $spp 3 | tail -1; done; done
} > gpurun_out/r02l_blocks.log 2>&1; cat gpurun_out/r02l_blocks.log
