#!/bin/bash
# Round-2 GPU session X: software prefetch in the BVH traversal (pushed children, postponed leaf) — none / L2 / L1 / L2 all children.
set -u
mkdir -p gpurun_out
{
for so in pf0 pf1 pf2 pf3; do echo "== $so"; for wl in C4_1M C4_10M; do PTB200_LIB=$PWD/build/variants/$so.so timeout 300 python tools/profile_kernel.py $wl 16 3 | tail -1; done; done
} > gpurun_out/r02x_prefetch.log 2>&1; cat gpurun_out/r02x_prefetch.log
timeout 300 python -m pytest tests/test_gpu_isolation.py tests/test_mesh.py -q -m gpu -s 2>&1 | grep -E "bit-identical|passed|failed" > gpurun_out/r02x_tests.log; cat gpurun_out/r02x_tests.log
