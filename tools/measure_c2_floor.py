#!/usr/bin/env python3
"""C2 (test_scene 1920x1080, 64 spp, depth 10 — the BASELINE config that names "RMSE vs Go render") at full size:
oracle-vs-oracle noise floor of the 8-bit image comparison the GPU test makes (renderer.go:189-221).

    python tools/measure_c2_floor.py [--threads 8]

Renders with the fp64 oracle: R = 256 spp (4x the config, seed 4004) and A = 64 spp (seed 3003), both through the
reference's epilogue (mean, sqrt, *255.999, truncate).  floor = RMSE over all R,G,B bytes of A vs R: what an exact
implementation with an independent RNG stream scores against R.  Writes tests/golden/c2_fullsize_oracle256.npz (R as
uint8, zlib) + the floor in its meta; tests/test_gpu_converged.py compares the device's 64-spp image with R.
"""
import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=8)
    args = ap.parse_args()
    from oracle import pyoracle
    W, H, spp, depth = 1920, 1080, 64, 10
    ora = pyoracle.OracleScene.load(ROOT / "scenes" / "test_scene.json")
    t0 = time.time()
    ref, _ = ora.render_rgba(W, H, 4 * spp, depth, seed=4004, threads=args.threads)
    a, _ = ora.render_rgba(W, H, spp, depth, seed=3003, threads=args.threads)
    d = a[..., :3].astype(np.float64) - ref[..., :3].astype(np.float64)
    meta = {"scene": "test_scene", "width": W, "height": H, "spp_config": spp, "spp_reference": 4 * spp, "max_depth": depth,
            "seed_reference": 4004, "seed_floor": 3003, "floor_rmse_8bit": float(np.sqrt((d ** 2).mean())),
            "floor_mean_abs_8bit": float(np.abs(d).mean()), "floor_bias_8bit": float(d.mean()), "seconds": time.time() - t0}
    np.savez_compressed(ROOT / "tests" / "golden" / "c2_fullsize_oracle256.npz", rgb=ref[..., :3].copy(), meta=json.dumps(meta))
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
