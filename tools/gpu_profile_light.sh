#!/bin/bash
# ncu full capture of the integrator on a reduced-spp workload. Usage: tools/gpu_profile_light.sh <tag> [workload] [spp]
set -u
TAG=${1:-p}; WL=${2:-C3}; SPP=${3:-32}
mkdir -p gpurun_out
CMD="timeout 120 python tools/profile_kernel.py $WL $SPP 3"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:integrate_ -s ${NCU_SKIP:-2} -c 1 -o gpurun_out/${TAG}_integrate -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/${TAG}_plain.log; tail -3 gpurun_out/${TAG}_ncu.log
