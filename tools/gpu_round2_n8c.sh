#!/bin/bash
# 8 GPUs: the slice kernel with plain unrolled peer loads — exchange-only timings and the C3 / C2 lines.
set -u
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29799 tools/peer_check.py > gpurun_out/r02cn8_peer_check.json 2> gpurun_out/r02cn8_peer_check.err; echo "peer_check rc=$?"
timeout 400 $TR --master-port 29801 bench.py --gpus $N --no-cpu --no-extras --steps 10 --warmup 3 --exchange peer > gpurun_out/r02cn8_C3_peer.json 2> gpurun_out/r02cn8_C3_peer.err; echo "C3 rc=$?"
timeout 400 $TR --master-port 29802 bench.py --gpus $N --no-cpu --no-extras --steps 20 --warmup 5 --exchange peer --workload C2 > gpurun_out/r02cn8_C2_peer.json 2> gpurun_out/r02cn8_C2_peer.err; echo "C2 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02cn8_peer_check.json') if l.startswith('{')][-1])
for c in d['cases']: print(c['scene'], c['size'], 'peer', round(c['exchange_only_peer_ms'],3), 'nccl reduce+finalize', round(c['exchange_only_nccl_reduce_ms'],3), 'ok', d['ok'])
for f in ['C3_peer','C2_peer']:
    x=json.loads(open(f'gpurun_out/r02cn8_{f}.json').read().strip().splitlines()[-1]); print(f, round(x['value']), x['ms_per_step'], x['roofline']['kernel_ms'])
P
