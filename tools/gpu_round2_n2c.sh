#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 tools/peer_check.py > gpurun_out/r02n2c_peer_check.json 2> gpurun_out/r02n2c_peer_check.err; echo "peer_check rc=$?"; grep "^{" gpurun_out/r02n2c_peer_check.json | cut -c1-1500; tail -3 gpurun_out/r02n2c_peer_check.err
