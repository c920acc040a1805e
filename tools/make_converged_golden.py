#!/usr/bin/env python3
"""Generates tests/golden/converged_<scene>.npz — converged block-mean radiance of the fp64 oracle (the restatement of
renderIntoCPU, renderer.go:171-221 + 286-404) for the five BASELINE scenes, and the oracle-vs-oracle noise floor the
GPU parity tolerance is derived from (VERDICT r01 "next round" #1).

    python tools/make_converged_golden.py [--spp 4096] [--threads 6] [scene ...]

Per scene: two INDEPENDENT estimates (RNG seeds 1001 and 2002) of the 480x270 frame at `spp` samples per pixel each,
linear radiance (sum / spp), binary64, reduced to 4x4-block means over rows 0..267 (270 is not a multiple of 4)
-> (67, 120, 3) per estimate, stored as float32.
    floor_rel_rmse = sqrt(mean((A - B)^2)) / mean((A + B) / 2)      two spp-sample estimates against each other
    floor_lum      = lum(A) / lum(B)                                  Rec.709 luminance of the whole cropped frame
The golden image is (A + B) / 2 (2 x spp samples).  A device render of n samples compared with it has the expected
rel-RMSE  floor * sqrt((1/n + 1/(2 spp)) / (2/spp)); tests/test_gpu_converged.py states the tolerance it derives.

This runs in the build container (CPU only, ~0.5 h on 8 cores for all five scenes); the .npz files are committed.
"""
from __future__ import annotations

import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

W, H, BLOCK = 480, 270, 4
SEEDS = (1001, 2002)
DEPTH = {"example_simple": 8, "test_scene": 10, "metal_glass_room": 16, "test_comprehensive": 10, "gpu_showcase": 12}


def block_means(img: np.ndarray) -> np.ndarray:
    """(H, W, 3) linear radiance -> 4x4-block means over the rows that tile ((H // 4) * 4 = 268 of 270)."""
    h = img.shape[0] // BLOCK * BLOCK
    w = img.shape[1] // BLOCK * BLOCK
    return img[:h, :w].reshape(h // BLOCK, BLOCK, w // BLOCK, BLOCK, 3).mean(axis=(1, 3))


def luminance(a: np.ndarray) -> float:
    return float((0.2126 * a[..., 0] + 0.7152 * a[..., 1] + 0.0722 * a[..., 2]).mean())


def rel_rmse(x: np.ndarray, ref: np.ndarray) -> float:
    return float(np.sqrt(((x - ref) ** 2).mean()) / ref.mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="*", default=list(DEPTH))
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    args = ap.parse_args()
    from oracle import pyoracle
    pyoracle.build()
    for name in args.scenes:
        ora = pyoracle.OracleScene.load(ROOT / "scenes" / f"{name}.json")
        est, stats, secs = [], [], []
        for seed in SEEDS:
            t0 = time.time()
            s, st = ora.render_sum(W, H, args.spp, DEPTH[name], seed=seed, precision=64, threads=args.threads)
            secs.append(time.time() - t0)
            est.append(block_means(s / args.spp))
            stats.append(st)
        a, b = est
        gold = 0.5 * (a + b)
        floor = float(np.sqrt(((a - b) ** 2).mean()) / gold.mean())
        floor_lum = luminance(a) / luminance(b)
        meta = {"scene": name, "width": W, "height": H, "block": BLOCK, "rows_used": H // BLOCK * BLOCK, "spp_each": args.spp,
                "seeds": list(SEEDS), "max_depth": DEPTH[name], "precision": "binary64 oracle (oracle/oracle.cpp)",
                "floor_rel_rmse": floor, "floor_lum_ratio": floor_lum, "mean_radiance": float(gold.mean()),
                "segments_per_sample": stats[0]["segments"] / stats[0]["samples"], "seconds": secs}
        out = pathlib.Path(args.out) / f"converged_{name}.npz"
        np.savez_compressed(out, a=a.astype(np.float32), b=b.astype(np.float32), meta=json.dumps(meta))
        print(json.dumps(meta), flush=True)


if __name__ == "__main__":
    main()
