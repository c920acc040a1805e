#!/bin/bash
# Round-2 GPU session S: (1) per-scene converged-parity values (pytest -s), (2) work-item size sweep (PTB_SPLIT = sample planes per pixel).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_converged.py -m gpu -q -s > gpurun_out/r02s_parity.log 2>&1; echo "parity rc=$?"
grep -E "rel-RMSE|C2 full|C5 frame|passed|failed" gpurun_out/r02s_parity.log
{
for cfg in "C3 256" "C3 32" "C2 64" "C2 8" "C5 64" "C1 16"; do set -- $cfg
  for k in 1 2 4 8 16 32; do [ $k -gt $2 ] && continue; echo "== $1 spp=$2 split=$k"; PTB_SPLIT=$k timeout 300 python tools/profile_kernel.py $1 $2 3 | tail -1; done
done
} > gpurun_out/r02s_split.log 2>&1; cat gpurun_out/r02s_split.log
