#!/bin/bash
# Round-2 GPU session K: CTA shape for the mesh instantiation (128 threads x 4 slots: more rays per traversing lane).
set -u
mkdir -p gpurun_out
{
for so in default t128s4 t128s4b6; do echo "== $so"; for wl in C4_1M C4_10M C3; do spp=16; [ $wl = C3 ] && spp=64; PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $wl $spp 3 | tail -1; done; done
} > gpurun_out/r02k_shape.log 2>&1; cat gpurun_out/r02k_shape.log
