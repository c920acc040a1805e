#!/bin/bash
# Round-2 closing N = 8 session (short: the budget allows ~2 minutes of an 8-GPU box): C2 and C3 through torchrun on the final build.
set -u
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 45 $TR --master-port 29802 bench.py --gpus $N --no-cpu --no-extras --steps 20 --warmup 5 --workload C2 > gpurun_out/r02n8d_C2_peer.json 2> gpurun_out/r02n8d_C2_peer.err; echo "C2 rc=$?"
timeout 50 $TR --master-port 29801 bench.py --gpus $N --no-cpu --no-extras --steps 8 --warmup 3 > gpurun_out/r02n8d_C3_peer.json 2> gpurun_out/r02n8d_C3_peer.err; echo "C3 rc=$?"
python - <<'P'
import json
for f in ['C2_peer','C3_peer']:
    try:
        x=json.loads([l for l in open(f'gpurun_out/r02n8d_{f}.json') if l.startswith('{')][-1]); print(f, round(x['value']), x['ms_per_step'], x['roofline']['kernel_ms'], 'e2e', round(x['e2e']['value']), x['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
