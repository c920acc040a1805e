#!/usr/bin/env python3
"""Small driver for ncu: the integrator on a BASELINE workload at reduced spp (same instruction mix, short kernel).

    python tools/profile_kernel.py [workload=C3] [spp=32] [reps=3]
"""
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS, load_scene  # noqa: E402
from path_trace_golang_b200 import engine, scene  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
name, W, H, _, depth = WORKLOADS[wl]
ctx = engine.Context(0)
ctx.upload(load_scene(wl))
if ctx.bvh_info()["n_triangles"]:
    print("bvh:", ctx.bvh_info())
import torch  # noqa: E402
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
for i in range(reps):
    t0 = time.perf_counter()
    ctx.render_device(ctx.cfg(W, H, spp, depth, seed=1), out.data_ptr(), 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{wl} {W}x{H} spp={spp}: {dt * 1e3:.1f} ms, {W * H * spp / dt / 1e6:.0f} Msamples/s")
