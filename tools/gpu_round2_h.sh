#!/bin/bash
# Round-2 GPU session H: postponed-leaf traversal — mesh parity tests, budget / round sweep, pipeline variant, ncu capture.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mesh.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r02h_pytest.log; cat gpurun_out/r02h_pytest.log
{
for so in post_b16r4 post_b24r6 post_b24r8 post_b16r2 post_b12r3; do echo "== $so"; for wl in C4_1M; do PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $wl 16 3 | tail -1; done; done
echo "== post_b16r4 10M"; timeout 300 python tools/profile_kernel.py C4_10M 16 3 | tail -1
echo "== pipeline (16 CTAs/SM of slots)"; for wl in C4_1M C4_10M; do PTB_MESH_PIPELINE=1 PTB_MP_CTAS_PER_SM=16 timeout 300 python tools/profile_kernel.py $wl 16 3 | tail -1; done
} > gpurun_out/r02h_c4.log 2>&1; cat gpurun_out/r02h_c4.log
NCU_SKIP=2 tools/gpu_profile_light.sh r02h C4_1M 4
