"""Loader of REFERENCE-PRODUCED golden vectors (tests/golden/ref/, written by go/cmd/gengolden on a box with a Go toolchain).

Layout (all little-endian; <scene> in the five shipped scene names):
    <scene>.primary.bin    int32[H*W] world index of the primary hit (pixel centre, lens off; -1 = miss), then float64[H*W] t
    <scene>.camera.bin     float64[22]: origin, lowerLeftCorner, horizontal, vertical, u, v, w, lensRadius (newCamera at WxH)
    <scene>.converged.bin  float64[67*120*3]: 4x4-block means (rows 0..267) of the 480x270 mean linear radiance
    manifest.json          {"go", "goarch", "spp", "seed", "primary": {"width", "height", "xi_u", "xi_v"},
                            "converged": {"width", "height", "block", "rows_used"}, "world_size": {...}, "max_depth": {...}}
The directory does not exist in this repository yet: the build image has no Go toolchain, so the oracle is PARITY-UNPINNED and
tests/test_reference_golden.py says so in its skip message.  The loader itself is tested against a synthetic directory written
in the same layout (from the oracle's outputs)."""
import json
import pathlib

import numpy as np

REF_DIR = pathlib.Path(__file__).resolve().parent / "golden" / "ref"


def available(ref_dir=REF_DIR) -> bool:
    return (pathlib.Path(ref_dir) / "manifest.json").exists()


def load(name, ref_dir=REF_DIR):
    ref_dir = pathlib.Path(ref_dir)
    man = json.loads((ref_dir / "manifest.json").read_text())
    W, H = man["primary"]["width"], man["primary"]["height"]
    raw = (ref_dir / f"{name}.primary.bin").read_bytes()
    if len(raw) != W * H * 12:
        raise ValueError(f"{name}.primary.bin: {len(raw)} bytes, expected {W * H * 12}")
    ids = np.frombuffer(raw, dtype="<i4", count=W * H).reshape(H, W)
    t = np.frombuffer(raw, dtype="<f8", offset=W * H * 4, count=W * H).reshape(H, W)
    cam = np.frombuffer((ref_dir / f"{name}.camera.bin").read_bytes(), dtype="<f8")
    if cam.shape != (22,):
        raise ValueError(f"{name}.camera.bin: {cam.size} values, expected 22")
    c = man["converged"]
    shape = (c["rows_used"] // c["block"], c["width"] // c["block"], 3)
    conv = np.frombuffer((ref_dir / f"{name}.converged.bin").read_bytes(), dtype="<f8")
    if conv.size != shape[0] * shape[1] * 3:
        raise ValueError(f"{name}.converged.bin: {conv.size} values, expected {shape}")
    return {"ids": ids, "t": t, "camera": cam, "converged": conv.reshape(shape), "manifest": man,
            "world_size": man["world_size"][name], "max_depth": man["max_depth"][name]}


def write(ref_dir, name, ids, t, cam, conv, manifest):
    """The writer twin of go/cmd/gengolden (used by the loader's self-test and by anyone porting the generator)."""
    ref_dir = pathlib.Path(ref_dir)
    ref_dir.mkdir(parents=True, exist_ok=True)
    with open(ref_dir / f"{name}.primary.bin", "wb") as f:
        f.write(np.ascontiguousarray(ids, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(t, dtype="<f8").tobytes())
    (ref_dir / f"{name}.camera.bin").write_bytes(np.ascontiguousarray(cam, dtype="<f8").tobytes())
    (ref_dir / f"{name}.converged.bin").write_bytes(np.ascontiguousarray(conv, dtype="<f8").tobytes())
    (ref_dir / "manifest.json").write_text(json.dumps(manifest, indent=2))
