#!/usr/bin/env python
"""Run-to-run determinism of the mesh path (C4_1M at 4 spp): plain vs plain, stats vs stats, plain vs stats."""
import json, pathlib, sys
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import bench
from path_trace_golang_b200 import engine
ctx = engine.Context(0)
ctx.upload(bench.load_scene(sys.argv[1] if len(sys.argv) > 1 else "C4_1M"))
W, H = 1920, 1080
for spp, depth in ((1, 1), (1, 2), (1, 3), (4, 10)):
  def run(stats): return ctx.render_accum(ctx.cfg(W, H, spp, depth, seed=1, stats=stats))
  a, b, c, d = run(False), run(False), run(True), run(True)
  print("spp", spp, "depth", depth)
  for n, x, y in (("plain/plain", a, b), ("stats/stats", c, d), ("plain/stats", a, c)):
    diff = np.any(x != y, axis=2)
    print("  ", n, "differing pixels:", int(diff.sum()), "max abs", float(np.abs(x - y).max()))
    if diff.any():
        ys, xs = np.nonzero(diff)
        print("      first:", [(int(yy), int(xx), x[yy, xx].tolist(), y[yy, xx].tolist()) for yy, xx in zip(ys[:3], xs[:3])])
