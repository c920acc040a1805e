#!/bin/bash
# Round-2 session on 8 GPUs: scaling lines of every BASELINE config with the fused peer exchange (and the NCCL baselines next to it),
# C5 by image tiles, one process driving all devices (ptb_multi_*), the multi-process exchange check.
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29800
run() { # name, extra args...
  name=$1; shift; P=$((P+1))
  timeout 400 $TR --master-port $P bench.py --gpus $N --no-cpu --no-extras "$@" > gpurun_out/r02bn${N}_$name.json 2> gpurun_out/r02bn${N}_$name.err; echo "$name rc=$?"
}
timeout 300 $TR --master-port 29799 tools/peer_check.py > gpurun_out/r02bn${N}_peer_check.json 2> gpurun_out/r02bn${N}_peer_check.err; echo "peer_check rc=$?"; tail -c 600 gpurun_out/r02bn${N}_peer_check.json
run C3_peer --steps 10 --warmup 3 --exchange peer
run C3_reduce --steps 10 --warmup 3 --exchange reduce
run C2_peer --steps 20 --warmup 5 --exchange peer --workload C2
run C2_reduce --steps 20 --warmup 5 --exchange reduce --workload C2
run C5_peer --steps 3 --warmup 3 --exchange peer --workload C5
run C4_10M_peer --steps 5 --warmup 3 --exchange peer --workload C4_10M
for wl in C3; do timeout 300 python tools/multi_bench.py --workload $wl --steps 3 > gpurun_out/r02bn${N}_multi_$wl.json 2> gpurun_out/r02bn${N}_multi_$wl.err; echo "multi $wl rc=$?"; cat gpurun_out/r02bn${N}_multi_$wl.json; done
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/r02bn${N}_C*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), 'Msamples/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']), d['clocks']['reasons'])
    except Exception as e: print(f, 'ERR', e)
P
