#!/bin/bash
# Round-2 closing N = 2 session: the two-GPU tests, the peer exchange check and the bench line through torchrun on the final build.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r02n2d_pytest.log; cat gpurun_out/r02n2d_pytest.log
timeout 300 $TR --master-port 29711 tools/peer_check.py > gpurun_out/r02n2d_peer_check.json 2> gpurun_out/r02n2d_peer_check.err; echo "peer_check rc=$?"; grep "^{" gpurun_out/r02n2d_peer_check.json | cut -c1-700
for wl in C3 C2; do
  timeout 300 $TR --master-port 29712 bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r02n2d_${wl}_peer.json 2> gpurun_out/r02n2d_${wl}_peer.err; echo "$wl peer rc=$?"
done
timeout 300 $TR --master-port 29713 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02n2d_default.json 2> gpurun_out/r02n2d_default.err; echo "default rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n2d_*.json')):
    if 'peer_check' in f: continue
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['gpu_launches_note'][:160])
    except Exception as e: print(f, 'ERR', e)
P
