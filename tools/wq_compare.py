import sys, pathlib, time
sys.path.insert(0, "/root/repo")
import numpy as np
from path_trace_golang_b200 import engine, scene
from bench import WORKLOADS, load_scene
ctx = engine.Context(0)
for wl, spp in [("C1", 16), ("C3", 16), ("C2", 16)]:
    name, W, H, _, depth = WORKLOADS[wl]
    ctx.upload(load_scene(wl))
    a = ctx.render_accum(ctx.cfg(W, H, spp, depth, seed=1, wavequeue=False))
    t_a = ctx.stats()["last_render_ms"]
    b = ctx.render_accum(ctx.cfg(W, H, spp, depth, seed=1, wavequeue=True))
    t_b = ctx.stats()["last_render_ms"]
    ok = (np.abs(a - b) <= 1e-4 * np.maximum(1.0, np.abs(a))).all(axis=2).mean()
    print(f"{wl}: barrier {t_a:.2f} ms  queue {t_b:.2f} ms  agree {ok:.5f}", flush=True)
