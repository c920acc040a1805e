#!/bin/bash
# Round-2 GPU session U: packed sphere test only (boxes scalar) against the scalar scan, then the GPU suite on the default build.
set -u
mkdir -p gpurun_out
{
for so in scalar hybrid hybrid_u1; do echo "== $so"; for wl in C3 C2 C5 C1 C4_1M; do spp=64; [ $wl = C1 ] && spp=16; [ $wl = C4_1M ] && spp=16; PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $wl $spp 3 | tail -1; done; done
for so in scalar hybrid; do echo "== $so C3 full"; PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py C3 256 3 | tail -1; done
} > gpurun_out/r02u_packed.log 2>&1; cat gpurun_out/r02u_packed.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02u_pytest.log
