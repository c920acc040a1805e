#!/bin/bash
# Round-2 GPU session E: the mesh pipeline (wavefront across kernels) — parity tests, A/B against the in-kernel traversal, sweeps.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mesh.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r02e_pytest.log; cat gpurun_out/r02e_pytest.log
{
echo "== persistent (in-kernel traversal)"; for wl in C4_1M C4_10M; do PTB_MESH_PERSISTENT=1 timeout 200 python tools/profile_kernel.py $wl 16 3 | tail -1; done
for k in 4 8 16; do echo "== pipeline, $k CTAs/SM of slots"; for wl in C4_1M C4_10M; do PTB_MP_CTAS_PER_SM=$k timeout 200 python tools/profile_kernel.py $wl 16 3 | tail -1; done; done
for so in pipe_t5 pipe_t6 pipe_r2 pipe_r8; do echo "== $so"; for wl in C4_1M; do PTB200_LIB=$PWD/build/variants/$so.so timeout 200 python tools/profile_kernel.py $wl 16 3 | tail -1; done; done
} > gpurun_out/r02e_c4.log 2>&1; cat gpurun_out/r02e_c4.log
