#!/bin/bash
# Round-2 GPU session Y2: item-switching slots sorted into a class of their own (CL_SWITCH) against switching inside the terminate chunks.
set -u
mkdir -p gpurun_out
{
for so in sw0 sw1; do echo "== $so"; for cfg in "C3 256" "C3 32" "C2 64" "C2 8" "C5 64" "C1 16" "C4_1M 16"; do set -- $cfg; PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $1 $2 3 | tail -1; done; done
} > gpurun_out/r02y2_switch.log 2>&1; cat gpurun_out/r02y2_switch.log
PTB200_LIB=$PWD/build/variants/sw1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_isolation.py tests/test_mesh.py -q -m gpu -x 2>&1 | tail -3
