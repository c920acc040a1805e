#!/bin/bash
# Round-2 GPU session D: traversal-list ordering + budget sweep on C4, mesh / parity tests, C3 check of the inline-PTX dispenser.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_mesh.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r02d_pytest.log; cat gpurun_out/r02d_pytest.log
for so in lpt_b16r4 tail8 tail4 b8r4 b24r4 b16r2 b32r8tail8; do echo "== $so"; PTB200_LIB=$PWD/build/variants/$so.so timeout 200 python tools/profile_kernel.py C4_1M 16 3 | tail -1; done > gpurun_out/r02d_c4.log 2>&1
for so in lpt_b16r4 tail8 b24r4; do echo "== $so"; PTB200_LIB=$PWD/build/variants/$so.so timeout 200 python tools/profile_kernel.py C4_10M 16 3 | tail -1; done >> gpurun_out/r02d_c4.log 2>&1; cat gpurun_out/r02d_c4.log
for wl in C3 C2 C5; do python tools/profile_kernel.py $wl 64 3 | tail -1; done > gpurun_out/r02d_c3.log 2>&1; cat gpurun_out/r02d_c3.log
