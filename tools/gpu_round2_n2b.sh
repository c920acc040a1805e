#!/bin/bash
# 2 GPUs: the flag-synchronised peer exchange — correctness against the other exchanges and short bench lines.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 tools/peer_check.py > gpurun_out/r02n2d_peer_check.json 2> gpurun_out/r02n2d_peer_check.err; echo "peer_check rc=$?"; cat gpurun_out/r02n2d_peer_check.json | cut -c1-900; tail -3 gpurun_out/r02n2d_peer_check.err
P=29710
for ex in peer reduce; do
  P=$((P+1)); timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-extras --exchange $ex > gpurun_out/r02n2d_C3_$ex.json 2> gpurun_out/r02n2d_C3_$ex.err; echo "C3 $ex rc=$?"
  P=$((P+1)); timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-extras --exchange $ex --workload C2 > gpurun_out/r02n2d_C2_$ex.json 2> gpurun_out/r02n2d_C2_$ex.err; echo "C2 $ex rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n2d_C*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02n2d_C3_peer.err
