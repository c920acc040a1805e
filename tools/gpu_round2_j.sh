#!/bin/bash
set -u
mkdir -p gpurun_out
PTB_MESH_PIPELINE=1 PTB_MP_CTAS_PER_SM=16 ncu --set full --clock-control none --import-source on -k regex:mp_traverse -s 30 -c 1 -o gpurun_out/r02j_traverse -f python tools/profile_kernel.py C4_1M 16 1 > gpurun_out/r02j_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02j_ncu.log
