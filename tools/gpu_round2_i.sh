#!/bin/bash
# Round-2 GPU session I: mesh pipeline replayed as a CUDA graph — equality with the in-kernel traversal, timings by pool size.
set -u
mkdir -p gpurun_out
python - <<'P' > gpurun_out/r02i_equal.log 2>&1
import os, sys, numpy as np
sys.path.insert(0, '.')
from bench import load_scene
from path_trace_golang_b200 import engine
ctx = engine.Context(0); ctx.upload(load_scene("C4_1M"))
cfg = ctx.cfg(640, 360, 4, 10, seed=3)
a = ctx.render_accum(cfg)
os.environ["PTB_MESH_PIPELINE"] = "1"
b = ctx.render_accum(cfg); k = ctx.last_kernel()
print("pipeline == in-kernel traversal:", bool((a == b).all()), float(np.abs(a - b).max()), k)
P
cat gpurun_out/r02i_equal.log
{
for k in 8 16 32; do echo "== pipeline graph, $k CTAs/SM of slots"; for wl in C4_1M C4_10M; do PTB_MESH_PIPELINE=1 PTB_MP_CTAS_PER_SM=$k timeout 300 python tools/profile_kernel.py $wl 16 3 | tail -1; done; done
} > gpurun_out/r02i_c4.log 2>&1; cat gpurun_out/r02i_c4.log
