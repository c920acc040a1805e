"""Two independent restatements of the reference's CPU path must agree bit for bit.

oracle/oracle.cpp (C++, the checker of the CUDA path) and oracle/goref.py (plain Python, written from the Go sources alone)
both restate internal/engine/{math,objects,materials,camera,renderer}.go in binary64 without fused multiply-add, in the
expression order of the Go code, on the shared counter RNG.  The reference cannot be executed here (no Go toolchain) and ships
no golden vectors; agreement of two independent readings to the last bit is the strongest pin available without it, and any
difference names a statement to re-read against the Go source.  (CPU only; pure-Python loops, so the frames are small.)
"""
import json
import math
import struct

import numpy as np
import pytest

from conftest import SCENE_DEPTH, SCENES, scene_path


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


@pytest.fixture(scope="module")
def pairs(oracle_mod):
    from oracle import goref
    out = {}
    for name in SCENES:
        doc = json.loads(open(scene_path(name)).read())
        out[name] = (oracle_mod.OracleScene(doc), goref.Scene(doc), doc)
    return out


def test_rng_spec_is_the_same(oracle_mod):
    from oracle import goref
    for seed, pixel, sample in [(1, 0, 0), (77, 12345, 3), (0xFFFFFFFF, 8294399, 255), (2026, 1, 4095)]:
        r = goref.Rng(seed, pixel, sample)
        for i in range(6):
            assert r.float64() == oracle_mod.rng_uniform(seed, pixel, sample, i)


@pytest.mark.parametrize("name", SCENES)
def test_world_and_camera_bit_for_bit(name, pairs):
    """sceneToWorld + convertMaterial (objects.go:225-269, materials.go:28-55) and newCamera (camera.go:19-58)."""
    from oracle import goref
    ora, ref, doc = pairs[name]
    w = ora.world()
    assert len(w) == len(ref.world) > 0
    for e, (kind, a, b, mat) in zip(w, ref.world):
        assert e["type"] == kind and e["mat_type"] == mat.typ
        bb = (b, 0.0, 0.0) if kind == goref.SPHERE else b
        assert [bits(v) for v in e["a"]] == [bits(v) for v in a]
        assert [bits(v) for v in e["b"]] == [bits(v) for v in bb]
        for key, val in (("albedo", mat.albedo), ("emit", mat.emit), ("absorption", mat.absorption)):
            assert [bits(v) for v in e[key]] == [bits(v) for v in val], key
        assert bits(e["rough"]) == bits(mat.rough) and bits(e["ior"]) == bits(mat.ior)
    for (W, H) in [(1920, 1080), (333, 211), (64, 64)]:
        assert [bits(v) for v in ora.camera(W, H)] == [bits(v) for v in ref.camera22(W, H)]


@pytest.mark.parametrize("name", SCENES)
def test_primary_hits_bit_for_bit(name, pairs):
    """Closest-hit scan of the lens-free camera ray (camera.go:70-73, renderer.go:292-302): ids and t, two sub-pixel offsets."""
    ora, ref, _ = pairs[name]
    W, H = 96, 54
    for xi in [(0.5, 0.5), (0.0, 0.999)]:
        ids, t = ora.primary_hits(W, H, *xi)
        rids, rt = ref.primary_hits(W, H, *xi)
        assert ids.ravel().tolist() == rids
        assert [bits(v) for v in t.ravel()] == [bits(v) for v in rt]
        assert (ids >= 0).mean() > 0.3


@pytest.mark.parametrize("name", SCENES)
def test_radiance_sums_and_counters_bit_for_bit(name, pairs):
    """The whole pixel loop (renderer.go:171-187 + rayColorOpt :286-404 + camera.getRay with the lens branch + all five
    materials + the dielectric exit search + Russian roulette) at the scene's BASELINE depth: per-pixel sums of 4 samples, every
    event counter."""
    ora, ref, _ = pairs[name]
    W, H, spp, depth = 48, 27, 4, SCENE_DEPTH[name]
    got, st = ora.render_sum(W, H, spp, depth, seed=11, precision=64, threads=2)
    want, rst = ref.render_sum(W, H, spp, depth, seed=11)
    for k in ["samples", "segments", "exit_scans", "prim_tests", "scatters", "end_sky", "end_emissive", "end_rr", "end_depth", "end_noscatter"]:
        assert st[k] == rst[k], (k, st[k], rst[k])
    assert st["accepts"] == rst["accepts"]
    g = got.reshape(-1, 3)
    bad = [(i, g[i].tolist(), want[i]) for i in range(W * H) if [bits(v) for v in g[i]] != [bits(v) for v in want[i]]]
    assert not bad, (len(bad), bad[:3])
    assert st["scatters"] > W * H and rst["segments"] > 2 * W * H          # the frame really exercises the path ...
    assert rst["end_rr"] > 0 and rst["end_sky"] + rst["end_emissive"] > 0  # ... Russian roulette and both kinds of path end included
    if name != "example_simple" or rst["exit_scans"]:
        assert rst["exit_scans"] > 0                                       # (every shipped scene has glass in view)
    # a second sample range (the partition the multi-GPU path uses) and another seed, fewer pixels
    got2, _ = ora.render_sum(12, 7, 2, depth, seed=5, precision=64, threads=1, s_begin=3)
    want2, _ = ref.render_sum(12, 7, 2, depth, seed=5, s_begin=3)
    assert [[bits(v) for v in px] for px in got2.reshape(-1, 3)] == [[bits(v) for v in px] for px in want2]


def test_epilogue_bit_for_bit(oracle_mod):
    """renderer.go:189-221: mean, sqrt, x 255.999, clamp, truncation to uint8; A = 255."""
    from oracle import goref
    rng = np.random.default_rng(3)
    sums = np.concatenate([rng.random((40, 3)) * 64, rng.random((20, 3)) * 400, np.zeros((2, 3)), np.full((2, 3), 1e9)]).reshape(8, 8, 3)
    for spp in (1, 7, 64):
        img = oracle_mod.finalize(sums, spp)
        want = [goref.finalize_pixel(px, spp) for px in sums.reshape(-1, 3)]
        assert img.reshape(-1, 4).tolist() == want


def test_go_min_max_special_cases():
    """math.Min / math.Max as Go specifies them (used at math.go:49, materials.go:185, renderer.go:378,384)."""
    from oracle import goref
    assert goref.go_min(1.0, 2.0) == 1.0 and goref.go_max(1.0, 2.0) == 2.0
    assert math.isnan(goref.go_min(1.0, math.nan)) and math.isnan(goref.go_max(math.nan, 1.0))
    assert goref.go_min(math.nan, -math.inf) == -math.inf and goref.go_max(math.nan, math.inf) == math.inf
    assert math.copysign(1.0, goref.go_min(0.0, -0.0)) == -1.0 and math.copysign(1.0, goref.go_max(-0.0, 0.0)) == 1.0


def random_scene(seed):
    """A synthetic scene that reaches the branches the shipped scenes leave cold: rough lambert (materials.go:83-90), smoothness
    overriding rough (:36-39), ior 0 -> 1.5 (:42-45), absorbing glass boxes and spheres, duplicate / unknown material ids
    (objects.go:226-233), unknown object types (dropped, :237-266), a lens (camera.go:61-68), the three background kinds."""
    r = np.random.default_rng(seed)
    def col(lo=0.1, hi=0.95): return {"r": float(r.uniform(lo, hi)), "g": float(r.uniform(lo, hi)), "b": float(r.uniform(lo, hi))}
    mats = [
        {"id": "floor", "type": "lambert", "albedo": col(), "rough": float(r.choice([0.0, 0.4]))},
        {"id": "rough-lambert", "type": "lambert", "albedo": col(), "rough": 0.8},
        {"id": "metal-rough", "type": "metal", "albedo": col(), "rough": float(r.uniform(0.05, 0.9))},
        {"id": "metal-smooth", "type": "metal", "albedo": col(), "rough": 0.7, "smoothness": float(r.choice([1.0, 0.6, 1.7]))},
        {"id": "glass", "type": "dielectric", "albedo": col(), "ior": float(r.choice([0.0, 1.33, 1.5, 2.4])), "absorption": {"r": 0.0, "g": 0.0, "b": 0.0}},
        {"id": "tinted", "type": "dielectric", "ior": 1.5, "absorption": {"r": float(r.uniform(0, 2)), "g": 0.1, "b": float(r.uniform(0, 2))}},
        {"id": "lamp", "type": "emissive", "emit": col(0.5, 1.0), "power": float(r.uniform(2, 12))},
        {"id": "mirror", "type": "mirror", "albedo": col(0.7, 1.0)},
        {"id": "odd", "type": "plastic", "albedo": col(), "rough": 1.5},          # unknown type: lambert, rough clamped
        {"id": "glass", "type": "dielectric", "ior": 1.7, "absorption": {"r": 0.3, "g": 0.0, "b": 0.0}},   # duplicate id: the later one wins
    ]
    ids = ["floor", "rough-lambert", "metal-rough", "metal-smooth", "glass", "tinted", "lamp", "mirror", "odd", "nowhere"]
    objs = [{"type": "plane", "position": {"x": 0, "y": 0, "z": 0}, "size": {"x": 1, "y": 1, "z": 1}, "material_id": "floor"}]
    for k in range(14):
        t = ["sphere", "box", "sphere_light", "box", "sphere", "torus"][k % 6]
        s = float(r.uniform(0.3, 1.1))
        objs.append({"type": t, "position": {"x": float(r.uniform(-4, 4)), "y": float(r.uniform(0.3, 2.5)), "z": float(r.uniform(-4, 4))},
                     "size": {"x": s, "y": float(r.uniform(0.3, 2.0)), "z": float(r.uniform(0.3, 2.0))},
                     "material_id": "lamp" if t == "sphere_light" else ids[int(r.integers(0, len(ids)))]})
    doc = {"name": f"random-{seed}",
           "camera": {"position": {"x": float(r.uniform(-1, 1)), "y": float(r.uniform(1.5, 4)), "z": float(r.uniform(7, 10))},
                      "target": {"x": 0, "y": 1, "z": 0}, "up": {"x": 0, "y": 1, "z": 0}, "fov": float(r.uniform(30, 70)),
                      "aperture": float(r.choice([0.0, 0.15])), "focus_dist": float(r.choice([0.0, 8.0])), "aspect_ratio": float(r.choice([0.0, 1.5]))},
           "objects": objs, "materials": mats, "background": col(0.0, 0.4)}
    kind = seed % 3
    if kind == 0:
        doc["sky"] = {"type": "gradient", "horizon": col(), "zenith": col()}
    elif kind == 1:
        doc["sky"] = {"type": "solid", "color": col()}
    return doc


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_scenes_bit_for_bit(seed, oracle_mod):
    from oracle import goref
    doc = json.loads(json.dumps(random_scene(seed)))
    ora, ref = oracle_mod.OracleScene(doc), goref.Scene(doc)
    assert len(ora.world()) == len(ref.world) == 13            # the plane + 14 objects - the 2 of the unknown type
    W, H = 40, 24
    ids, t = ora.primary_hits(W, H)
    rids, rt = ref.primary_hits(W, H)
    assert ids.ravel().tolist() == rids and [bits(v) for v in t.ravel()] == [bits(v) for v in rt]
    for depth in (12, 3):                                       # (depth 3: Russian roulette from the first bounce on)
        got, st = ora.render_sum(W, H, 3, depth, seed=seed, precision=64, threads=2)
        want, rst = ref.render_sum(W, H, 3, depth, seed=seed)
        for k in rst:
            assert st[k] == rst[k], (k, st[k], rst[k])
        assert [[bits(v) for v in px] for px in got.reshape(-1, 3)] == [[bits(v) for v in px] for px in want]
