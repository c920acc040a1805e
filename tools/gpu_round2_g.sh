#!/bin/bash
# Round-2 GPU session G: 256-bit node / triangle loads — mesh parity tests, C4 timings (in-kernel traversal and the pipeline), ncu capture.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mesh.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r02g_pytest.log; cat gpurun_out/r02g_pytest.log
{
echo "== in-kernel traversal, 256-bit loads"; for wl in C4_1M C4_4M C4_10M; do timeout 300 python tools/profile_kernel.py $wl 16 3 | tail -1; done
echo "== pipeline (16 CTAs/SM of slots), 256-bit loads"; for wl in C4_1M C4_10M; do PTB_MESH_PIPELINE=1 PTB_MP_CTAS_PER_SM=16 timeout 300 python tools/profile_kernel.py $wl 16 3 | tail -1; done
} > gpurun_out/r02g_c4.log 2>&1; cat gpurun_out/r02g_c4.log
NCU_SKIP=2 tools/gpu_profile_light.sh r02g C4_1M 4
