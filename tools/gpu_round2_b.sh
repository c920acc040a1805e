#!/bin/bash
# Round-2 GPU session B: whole parity suite, A/B of dynamic SHADE scheduling and the zero-direction check, phase timing.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rs 2>&1 | tail -60 > gpurun_out/r02b_pytest.log; tail -8 gpurun_out/r02b_pytest.log
tools/run_variants.sh C3 C2 C5 > gpurun_out/r02b_variants.log 2>&1; cat gpurun_out/r02b_variants.log
for v in timing dyn1timing; do
  echo "== $v"; PTB200_LIB=$PWD/build/variants/$v.so PTB_DEBUG_TIMING=1 PROBE_SPP=32 python tools/timing_probe.py C3 2>&1 | tail -12
done > gpurun_out/r02b_timing.log 2>&1; cat gpurun_out/r02b_timing.log
