#!/bin/bash
# Round-2 final GPU session: the whole parity suite, smoke(), the bench line + its ncu launch list, the final kernel capture, C4 and reference arms.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rs 2>&1 | tail -12 > gpurun_out/r02z_pytest.log; cat gpurun_out/r02z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02z_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02z_ncu_list.log 2>&1; echo "launch list rc=$?"
tools/gpu_profile_light.sh r02z C3 16
python bench.py --workload C4_10M --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r02z_bench_C4_10M.json 2> gpurun_out/r02z_bench_C4_10M.err; echo "C4_10M rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_ref.json 2> gpurun_out/r02z_ref.err; echo "ref rc=$?"
