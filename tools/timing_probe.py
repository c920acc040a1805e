import pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS
from path_trace_golang_b200 import engine, scene
import os
SPP = int(os.environ.get("PROBE_SPP", "16"))
for wl in sys.argv[1:] or ["C3", "C2"]:
    name, W, H, _, depth = WORKLOADS[wl]
    ctx = engine.Context(0)
    ctx.upload(scene.Load(ROOT / "scenes" / f"{name}.json"))
    for _ in range(2):
        ctx.render_accum(ctx.cfg(W, H, SPP, depth, seed=1))
    print(wl, "ms", ctx.stats()["last_render_ms"])
    ctx.close()
