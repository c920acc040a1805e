#!/bin/bash
# Last GPU session of round 2: what the driver runs at round end, on the final tree — GPU suite, smoke(), the default bench line.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02final_pytest.log; cat gpurun_out/r02final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1; tail -1 gpurun_out/r02final_smoke.log
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02final_bench.json') if l.startswith('{')][-1])
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d['roofline']['kernel'], 'parity', d['parity']['pass'], d['parity']['primary_hit_mismatches'], {k: round(v['msamples_s']) for k,v in d['configs'].items()})
P
