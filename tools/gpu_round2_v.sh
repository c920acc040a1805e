#!/bin/bash
# Round-2 closing GPU session (after the work-item policy and the PK instantiation): the whole parity suite, smoke(), the bench
# line + its ncu launch list, --set full captures of the three shipped instantiations (C3 scalar, C5 packed spheres, C4 10 M mesh),
# the C4 and reference arms.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rs 2>&1 | tail -14 > gpurun_out/r02v_pytest.log; cat gpurun_out/r02v_pytest.log
python -m pytest tests/test_gpu_isolation.py -q -m gpu -s -k packed 2>&1 | grep -E "bit-identical|passed|failed" > gpurun_out/r02v_packed_equal.log; cat gpurun_out/r02v_packed_equal.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; tail -2 gpurun_out/r02v_smoke.log
python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02v_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02v_ncu_list.log 2>&1; echo "launch list rc=$?"
tools/gpu_profile_light.sh r02v_c5 C5 4
NCU_SKIP=2 tools/gpu_profile_light.sh r02v_c4 C4_10M 4
python bench.py --workload C4_10M --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r02v_bench_C4_10M.json 2> gpurun_out/r02v_bench_C4_10M.err; echo "C4_10M rc=$?"
