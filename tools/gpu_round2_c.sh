#!/bin/bash
# Round-2 GPU session C: parity suite on the packed-state kernel + 4-wide BVH, A/B timings, ncu captures (C3 and C4 with 10 M triangles), bench line.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rs 2>&1 | tail -60 > gpurun_out/r02c_pytest.log; tail -6 gpurun_out/r02c_pytest.log
for so in base opt1 packed; do echo "== $so"; for wl in C3 C2 C5; do PTB200_LIB=$PWD/build/variants/$so.so timeout 60 python tools/profile_kernel.py $wl 64 3 | tail -1; done; done > gpurun_out/r02c_variants.log 2>&1; cat gpurun_out/r02c_variants.log
for so in packed c4b16r4 c4b12r6; do echo "== $so"; for wl in C4_1M C4_10M; do PTB200_LIB=$PWD/build/variants/$so.so timeout 200 python tools/profile_kernel.py $wl 16 3 | tail -1; done; done > gpurun_out/r02c_c4.log 2>&1; cat gpurun_out/r02c_c4.log
tools/gpu_profile_light.sh r02c C3 16
NCU_SKIP=2 tools/gpu_profile_light.sh r02d C4_10M 4
python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02c_bench.err
