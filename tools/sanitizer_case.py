#!/usr/bin/env python3
"""Smallest end-to-end case (both integrators, the fp64 parity kernel, the progressive epilogue, a mesh scene), meant for
`compute-sanitizer --tool memcheck|racecheck python tools/sanitizer_case.py`.  On this GPU pool compute-sanitizer is closed by the
operators (runs under it left GPUs needing a reset), so round 1 has no sanitizer log; DESIGN.md §6 argues race freedom by construction."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from path_trace_golang_b200 import engine, scene  # noqa: E402

ctx = engine.Context(0)
sc = scene.Load(ROOT / "scenes" / "metal_glass_room.json")
ctx.upload(sc)
W, H = 96, 54
a = ctx.render_accum(ctx.cfg(W, H, 3, 16, seed=1))
b = ctx.render_accum(ctx.cfg(W, H, 3, 16, seed=1, megakernel=True))
img = ctx.render(ctx.cfg(W, H, 3, 16, seed=1), progress=lambda: None)
ids, t = ctx.primary_hits(W, H)
doc = json.loads((ROOT / "scenes" / "example_simple.json").read_text())
doc["objects"].append({"type": "mesh", "position": {"x": 0, "y": 1, "z": 1}, "size": {"x": 6, "y": 1, "z": 6}, "material_id": "lambert-red",
                       "mesh": {"heightfield": {"nx": 24, "nz": 16, "seed": 3, "amplitude": 0.4, "frequency": 3, "octaves": 2}}})
ctx.upload(scene.Parse(json.dumps(doc)))
c = ctx.render_accum(ctx.cfg(W, H, 2, 8, seed=1, stats=True))
ids2, _ = ctx.primary_hits(W, H)
print("ok", float(a.sum()), float(b.sum()), int(img.sum()), int((ids >= 0).sum()), float(c.sum()), int((ids2 >= 0).sum()))
