#!/bin/bash
# Round-2 session on 2 GPUs: the multi-process exchanges (peer / scatter / reduce / rows) and ptb_multi_* against one GPU, short bench lines.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu -rs 2>&1 | tail -15 > gpurun_out/r02n2_pytest.log; cat gpurun_out/r02n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29701 tools/peer_check.py > gpurun_out/r02n2_peer_check.json 2> gpurun_out/r02n2_peer_check.err; echo "peer_check rc=$?"; cat gpurun_out/r02n2_peer_check.json; tail -5 gpurun_out/r02n2_peer_check.err
for ex in peer scatter reduce; do
  $TR --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-extras --exchange $ex > gpurun_out/r02n2_C3_$ex.json 2> gpurun_out/r02n2_C3_$ex.err; echo "C3 $ex rc=$?"
  $TR --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-extras --exchange $ex --workload C2 > gpurun_out/r02n2_C2_$ex.json 2> gpurun_out/r02n2_C2_$ex.err; echo "C2 $ex rc=$?"
done
$TR --master-port 29733 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-extras --partition rows > gpurun_out/r02n2_C3_rows.json 2> gpurun_out/r02n2_C3_rows.err; echo "rows rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n2_C*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
P
