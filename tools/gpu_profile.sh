#!/bin/bash
# Run on the GPU box (via gpurun): bench, then the ncu launch list and one full capture of the integrator.
# Usage: tools/gpu_profile.sh <tag>      -> gpurun_out/<tag>_*.{json,csv,ncu-rep}
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.json 2> gpurun_out/${TAG}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:integrate_ -s 1 -c 1 -o gpurun_out/${TAG}_integrate -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
