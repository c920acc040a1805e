#!/bin/bash
# Round-2 GPU session A: parity suite, bench line, A/B of the TERM+REGEN class merge, phase timing, ncu capture.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -40 > gpurun_out/r02a_pytest.log; tail -5 gpurun_out/r02a_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02a_bench.err
tools/run_variants.sh C3 C2 C5 > gpurun_out/r02a_variants.log 2>&1; cat gpurun_out/r02a_variants.log
PTB200_LIB=$PWD/build/variants/timing.so PTB_DEBUG_TIMING=1 python tools/profile_kernel.py C3 64 2 > gpurun_out/r02a_timing.log 2>&1; tail -12 gpurun_out/r02a_timing.log
tools/gpu_profile_light.sh r02a C3 16
