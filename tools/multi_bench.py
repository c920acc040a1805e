#!/usr/bin/env python
"""Single-process multi-GPU render (ptb_multi_*): time C3 through the host-buffer call on every device of the box.
Usage: python tools/multi_bench.py [--workload C3] [--steps 3]"""
import argparse
import json
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))

import torch  # noqa: E402  (device count only)

import bench  # noqa: E402
from path_trace_golang_b200 import engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--devices", type=int, default=0)
    a = ap.parse_args()
    n = a.devices or torch.cuda.device_count()
    name, W, H, spp, depth = bench.WORKLOADS[a.workload][:5]
    sc = bench.load_scene(a.workload)
    m = engine.MultiContext(n)
    m.upload(sc)
    cfg = engine.Context.cfg(W, H, spp, depth, seed=1)
    m.render(engine.Context.cfg(W, H, max(n, 4), depth, seed=1))           # warm-up
    best = None
    for _ in range(a.steps):
        t0 = time.perf_counter()
        m.render(cfg)
        wall = (time.perf_counter() - t0) * 1e3
        t = m.last_timing()
        if best is None or wall < best["wall_ms"]:
            best = dict(wall_ms=wall, **t)
    print(json.dumps(dict(workload=a.workload, n_gpus=n, msamples_per_s=W * H * spp / best["wall_ms"] / 1e3, **best)))


if __name__ == "__main__":
    main()
