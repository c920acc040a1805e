#!/bin/bash
# Round-2 GPU session T: packed two-ray scan (FFMA2) against the scalar scan, then the GPU suite on the default (packed) build.
set -u
mkdir -p gpurun_out
{
for so in scalar packed packed_bg2 packed_sg4 packed_sg1; do echo "== $so"; for wl in C3 C2 C5 C1; do spp=64; [ $wl = C1 ] && spp=16; PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $wl $spp 3 | tail -1; done; done
echo "== packed C3 full"; PTB200_LIB=$PWD/build/variants/packed.so timeout 300 python tools/profile_kernel.py C3 256 3 | tail -1
echo "== scalar C3 full"; PTB200_LIB=$PWD/build/variants/scalar.so timeout 300 python tools/profile_kernel.py C3 256 3 | tail -1
echo "== packed C4"; PTB200_LIB=$PWD/build/variants/packed.so timeout 300 python tools/profile_kernel.py C4_1M 16 3 | tail -1
echo "== scalar C4"; PTB200_LIB=$PWD/build/variants/scalar.so timeout 300 python tools/profile_kernel.py C4_1M 16 3 | tail -1
} > gpurun_out/r02t_packed.log 2>&1; cat gpurun_out/r02t_packed.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02t_pytest.log
