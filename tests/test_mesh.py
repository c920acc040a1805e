"""EXTENSION (BASELINE config C4): triangle meshes + BVH.  The reference has neither, so the
semantics are this repository's (DESIGN.md "Meshes"); the oracle implements them independently (brute force and its
own median-split BVH), the product with a binned-SAH BVH traversed on the GPU."""
import json

import numpy as np
import pytest

from conftest import SCENE_DEPTH, scene_json

CAM = {"position": {"x": 0, "y": 0, "z": 5}, "target": {"x": 0, "y": 0, "z": 0}, "up": {"x": 0, "y": 1, "z": 0}, "fov": 60}
SKY = {"type": "solid", "color": {"r": 1, "g": 1, "b": 1}}


def mesh_obj(vertices, triangles, pos=(0, 0, 0), size=(0, 0, 0), mat="m"):
    return {"type": "mesh", "position": dict(zip("xyz", pos)), "size": dict(zip("xyz", size)), "material_id": mat,
            "mesh": {"vertices": [float(x) for v in vertices for x in v], "triangles": [int(i) for t in triangles for i in t]}}


def with_heightfield(name, nx, nz, pos=(0, 0.8, 1), size=(6, 1, 6), mat=None, seed=7):
    sc = scene_json(name)
    mat = mat or sc["materials"][0]["id"]
    sc["objects"].append({"id": "terrain", "type": "mesh", "position": dict(zip("xyz", pos)), "size": dict(zip("xyz", size)),
                          "material_id": mat,
                          "mesh": {"heightfield": {"nx": nx, "nz": nz, "seed": seed, "amplitude": 0.4, "frequency": 3, "octaves": 3}}})
    return sc


LAMB = {"id": "m", "type": "lambert", "albedo": {"r": 0.5, "g": 0.5, "b": 0.5}}
QUAD_V = [(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)]
QUAD_T = [(0, 1, 2), (0, 2, 3)]


def test_triangle_hit_kats(oracle_mod):
    """Moeller-Trumbore, two-sided, geometric normal flipped against the ray."""
    sc = {"camera": CAM, "sky": SKY, "materials": [LAMB], "objects": [mesh_obj(QUAD_V, QUAD_T)]}
    o = oracle_mod.OracleScene(sc)
    ids, t, ff, _ = o.trace_path((0.25, 0.5, 5), (0, 0, -2), 1)
    assert list(ids) == [0] and t[0] == 2.5 and ff[0] == 1            # CCW seen from +z: front face
    ids, t, ff, _ = o.trace_path((0.25, 0.5, -4), (0, 0, 1), 1)
    assert list(ids) == [0] and t[0] == 4.0 and ff[0] == 0            # from behind: back face
    assert list(o.trace_path((1.5, 0, 5), (0, 0, -1), 1)[0]) == [-1]  # beside the quad
    assert list(o.trace_path((0, 0, 5), (1, 0, 0), 1)[0]) == [-1]     # parallel: det == 0
    assert list(o.trace_path((0, 0, 0.0005), (0, 0, -1), 1)[0]) == [-1]   # t < tMin = 0.001
    w = o.world()
    assert len(w) == 1 and w[0]["type"] == 3 and w[0]["a"] == [-1, -1, 0] and w[0]["b"] == [1, 1, 0]


def test_mesh_tie_rules(oracle_mod):
    """Meshes are scanned after the analytic objects and must beat them strictly; among triangles of equal t the
    lowest triangle id wins (here: the diagonal shared by both triangles, and two coincident quads)."""
    plane = {"type": "plane", "position": {"x": 0, "y": 0, "z": 0}, "size": {"x": 0, "y": 0, "z": 0}, "material_id": "m"}
    floor = mesh_obj([(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)], [(0, 2, 1), (0, 3, 2)])
    for objs, want in (([floor, plane], 1), ([plane, floor], 0)):          # coplanar with the plane: the plane wins either way
        sc = {"camera": CAM, "sky": SKY, "materials": [LAMB], "objects": objs}
        assert list(oracle_mod.OracleScene(sc).trace_path((0.3, 2, 0.1), (0, -1, 0), 1)[0]) == [want]
    two = {"camera": CAM, "sky": SKY, "materials": [LAMB], "objects": [mesh_obj(QUAD_V, QUAD_T), mesh_obj(QUAD_V, QUAD_T)]}
    assert list(oracle_mod.OracleScene(two).trace_path((0.3, 0.1, 5), (0, 0, -1), 1)[0]) == [0]   # first mesh = lower triangle ids


def test_host_mesh_flatten_generator_and_save(host_scenes):
    from path_trace_golang_b200 import scene
    doc = {"camera": CAM, "sky": SKY, "materials": [LAMB],
           "objects": [mesh_obj(QUAD_V, QUAD_T, pos=(1, 2, 3), size=(2, 0, 0.5)),
                       {"type": "mesh", "material_id": "m"},                                   # no triangles: dropped
                       {"type": "mesh", "position": {"x": 0, "y": 0, "z": 0}, "size": {"x": 4, "y": 1, "z": 2}, "material_id": "m",
                        "mesh": {"heightfield": {"nx": 5, "nz": 3, "seed": 11, "amplitude": 0.25, "frequency": 2, "octaves": 2}}}]}
    sc = scene.Parse(json.dumps(doc))
    flat = sc.flat()
    assert [flat.obj_type[i] for i in range(3)] == [3, -1, 3] and [flat.obj_mesh[i] for i in range(3)] == [0, -1, 1]
    tris = sc.mesh_triangles()
    assert tris[0].shape == (2, 9) and tris[2].shape == (5 * 3 * 2, 9)
    # world vertex = position + size * local (size 0 -> 1)
    np.testing.assert_array_equal(tris[0][0], [1 - 2, 2 - 1, 3, 1 + 2, 2 - 1, 3, 1 + 2, 2 + 1, 3])
    hf = tris[2].reshape(-1, 3)
    assert hf[:, 0].min() == -2 and hf[:, 0].max() == 2 and hf[:, 2].min() == -1 and hf[:, 2].max() == 1
    assert np.abs(hf[:, 1]).max() <= 0.25 * 1.5 + 1e-6 and hf[:, 1].std() > 0.01
    # deterministic, and Save keeps the extension (generator parameters, not 30 triangles)
    again = scene.Parse(sc.marshal())
    np.testing.assert_array_equal(again.mesh_triangles()[2], tris[2])
    np.testing.assert_array_equal(again.mesh_triangles()[0], tris[0])
    assert '"heightfield"' in sc.marshal() and '"vertices"' in sc.marshal()
    # shipped scenes are untouched by the extension
    assert host_scenes["example_simple"].flat().n_mesh == 0
    from path_trace_golang_b200 import PtbError
    with pytest.raises(PtbError, match="decode scene"):
        scene.Parse(json.dumps({"objects": [mesh_obj(QUAD_V, [(0, 1, 9)])]}))                  # index out of range


def test_oracle_bvh_equals_bruteforce(oracle_mod):
    from path_trace_golang_b200 import scene
    doc = with_heightfield("example_simple", 40, 30)
    tris = scene.Parse(json.dumps(doc)).mesh_triangles()
    ora = oracle_mod.OracleScene(doc, mesh_triangles=tris)
    a = ora.primary_hits(200, 112)
    sa, _ = ora.render_sum(64, 36, 2, 8, seed=3, precision=64)
    ora.set_mesh_accel(False)
    b = ora.primary_hits(200, 112)
    sb, _ = ora.render_sum(64, 36, 2, 8, seed=3, precision=64)
    assert (a[0] == b[0]).all() and (a[1].view(np.uint64) == b[1].view(np.uint64)).all() and (sa == sb).all()
    assert (a[0] == len(doc["objects"]) - 1).mean() > 0.05


def test_bvh_build_service_emits_aligned_wide_nodes():
    """ptb_bvh_build (host only): the flattened arrays the device traverses — 128-byte aligned 4-wide nodes whose links cover every
    triangle exactly once in leaves of at most 4, triangle records carrying the original index, depth about half a binary tree's."""
    import ctypes as C
    from path_trace_golang_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    n = 5000
    base = rng.random((n, 1, 3)).astype(np.float32) * 10
    tris = (base + rng.random((n, 3, 3)).astype(np.float32) * 0.3).reshape(n, 9)
    tris = np.ascontiguousarray(tris)
    b = _lib.PtbBvh()
    assert L.ptb_bvh_build(tris.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(b)) == 0
    try:
        assert b.info.n_triangles == n and b.info.node_bytes == 128 and b.info.triangle_bytes == 64
        assert C.addressof(b.nodes.contents) % 64 == 0
        nodes = np.ctypeslib.as_array(b.nodes, shape=(b.info.n_nodes, 32)).copy()
        tr = np.ctypeslib.as_array(b.triangles, shape=(n, 16)).copy()
        links = nodes[:, 24:28].view(np.int32)
        half = nodes[:, [3, 9, 15, 21]]
        used = half >= 0
        assert (links[~used] == 0x7fffffff).all() and (nodes[:, 28:] == 0).all()
        inner = links[used & (links >= 0)]
        assert sorted(inner.tolist()) == list(range(1, b.info.n_nodes))          # every node but the root has exactly one parent
        leaf = ~links[used & (links < 0)]
        first, cnt = leaf >> 2, (leaf & 3) + 1
        covered = np.zeros(n, dtype=np.int32)
        for f, c in zip(first, cnt):
            covered[f:f + c] += 1
        assert (covered == 1).all()
        ids = tr[:, 3].view(np.int32)
        assert sorted(ids.tolist()) == list(range(n))
        np.testing.assert_array_equal(tr[:, 0:3], tris[ids, 0:3])
        assert 4 <= b.info.max_depth <= 14 and used.sum(axis=1).mean() > 2.5       # log4(5000 / 4) ~ 5; inner levels are full, nodes over two leaves stay 2-wide
    finally:
        L.ptb_bvh_free(C.byref(b))
    assert L.ptb_bvh_selfcheck(tris.ctypes.data_as(C.POINTER(C.c_float)), n, None, None) == 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,nx,nz,res", [("example_simple", 40, 30, (1280, 720)), ("test_comprehensive", 300, 200, (1920, 1080))])
def test_gpu_primary_hits_with_mesh_bit_exact(name, nx, nz, res, ctx, oracle_mod):
    """ids and t of the binary64 kernel (BVH traversal + Moeller-Trumbore) == the oracle's, bit for bit."""
    from path_trace_golang_b200 import scene
    doc = with_heightfield(name, nx, nz, pos=(0, 1.2, 2), size=(14, 2.5, 10))
    sc = scene.Parse(json.dumps(doc))
    ctx.upload(sc)
    info = ctx.bvh_info()
    assert info["n_triangles"] == nx * nz * 2 and info["node_bytes"] == 128 and info["triangle_bytes"] == 64 and info["max_depth"] < 38
    ids, t = ctx.primary_hits(*res)
    oids, ot = oracle_mod.OracleScene(doc, mesh_triangles=sc.mesh_triangles()).primary_hits(*res)
    assert (ids == oids).all(), f"{(ids != oids).sum()} id mismatches"
    assert (t.view(np.uint64) == ot.view(np.uint64)).all()
    mesh_idx = len(ctx.world()) - 1
    assert ctx.world()[mesh_idx]["type"] == 3 and (ids == mesh_idx).mean() > 0.05


@pytest.mark.gpu
def test_gpu_mesh_render_vs_oracle(ctx, oracle_mod):
    """fp32 integrator with a mesh in the scene: path-for-path vs the binary32 oracle (>= 99 % of pixels within 1e-3),
    counters within 1 %, and the BVH counters are alive."""
    from path_trace_golang_b200 import scene, PtbError
    doc = with_heightfield("example_simple", 60, 40, pos=(0, 1.0, 1), size=(6, 1, 6), mat="metal-rough")
    sc = scene.Parse(json.dumps(doc))
    ctx.upload(sc)
    ora = oracle_mod.OracleScene(doc, mesh_triangles=sc.mesh_triangles())
    W, H, depth = 320, 180, 8
    dev = ctx.render_accum(ctx.cfg(W, H, 1, depth, seed=7)).astype(np.float64)
    ref, _ = ora.render_sum(W, H, 1, depth, seed=7, precision=32)
    ok = (np.abs(dev - ref) <= 1e-3 * np.maximum(1.0, np.abs(ref))).all(axis=2).mean()
    print(f"mesh scene path-for-path match {ok:.5f}")
    assert ok >= 0.99
    ctx.render_accum(ctx.cfg(W, H, 4, depth, seed=3, stats=True))
    d = ctx.stats()
    _, o = ora.render_sum(W, H, 4, depth, seed=3, precision=32)
    for k in ["segments", "scatters", "end_sky", "end_emissive"]:
        assert abs(d[k] - o[k]) <= 0.01 * max(o[k], 1000), (k, d[k], o[k])
    assert d["accepts_mesh"] > 0.02 * d["segments"] and d["bvh_nodes_visited"] > d["segments"] and d["bvh_tris_tested"] > 0
    assert sum(d["accepts"]) + d["accepts_mesh"] == d["segments"] - d["end_sky"]
    with pytest.raises(PtbError, match="megakernel"):
        ctx.render_accum(ctx.cfg(W, H, 1, depth, megakernel=True))


@pytest.mark.gpu
def test_gpu_c4_million_triangles(ctx, oracle_mod):
    """C4: test_comprehensive + a 1M-triangle heightfield at 1920x1080.  Full-size properties: BVH stats, determinism,
    primary ids vs the oracle (its own BVH) on a 480x270 sub-sampled grid of the same camera, bounded traversal work."""
    from path_trace_golang_b200 import scene
    doc = with_heightfield("test_comprehensive", 1000, 500, pos=(0, 1.2, 2), size=(14, 2.5, 10), mat="lambert-green")
    sc = scene.Parse(json.dumps(doc))
    ctx.upload(sc)
    info = ctx.bvh_info()
    print("C4 BVH:", info)
    assert info["n_triangles"] == 1_000_000 and info["n_nodes"] < 600_000
    ids, t = ctx.primary_hits(480, 270)
    oids, ot = oracle_mod.OracleScene(doc, mesh_triangles=sc.mesh_triangles()).primary_hits(480, 270)
    assert (ids == oids).all() and (t.view(np.uint64) == ot.view(np.uint64)).all()
    cfg = ctx.cfg(1920, 1080, 4, 10, seed=1, stats=True)
    a = ctx.render_accum(cfg)
    st = ctx.stats()
    b = ctx.render_accum(ctx.cfg(1920, 1080, 4, 10, seed=1))
    b2 = ctx.render_accum(ctx.cfg(1920, 1080, 4, 10, seed=1))
    assert np.array_equal(b, b2)                                                     # deterministic run to run (the ray list order is not)
    # the counting build is a separate instantiation: ptxas may contract a*b+c differently, which moves a hit by an ulp and,
    # after ten bounces, a fraction of a percent of the pixels onto another path
    close = np.all(np.abs(a - b) <= 1e-5 * np.maximum(1.0, np.abs(b)), axis=2)
    print(f"C4 stats vs plain build: {np.all(a == b, axis=2).mean():.4f} of the pixels bit-identical, {close.mean():.4f} within 1e-5")
    assert close.mean() > 0.99 and abs(a.mean() - b.mean()) < 2e-3 * b.mean()
    nodes_per_ray = st["bvh_nodes_visited"] / st["segments"]
    tris_per_ray = st["bvh_tris_tested"] / st["segments"]
    print(f"C4: {st['last_render_ms']:.1f} ms (stats variant), {nodes_per_ray:.1f} nodes/ray, {tris_per_ray:.1f} tris/ray, mesh hits {st['accepts_mesh'] / st['segments']:.3f}")
    assert 1 < nodes_per_ray < 200 and tris_per_ray < 50


@pytest.mark.gpu
def test_gpu_bvh_is_reused_for_identical_meshes(ctx):
    """RenderInto hands the scene over on every call (renderer.go:34): uploading the same triangles again must not
    rebuild the BVH (same build_ms = the kept build), a different mesh must, and results are unchanged."""
    from path_trace_golang_b200 import scene
    doc = with_heightfield("example_simple", 64, 48)
    sc = scene.Parse(json.dumps(doc))
    ctx.upload(sc)
    first = ctx.bvh_info()
    a = ctx.render_accum(ctx.cfg(160, 90, 2, 8, seed=1))
    ctx.upload(scene.Parse(json.dumps(doc)))                       # a fresh Scene object with identical content
    second = ctx.bvh_info()
    assert second == first                                          # build_ms identical: nothing was rebuilt
    assert np.array_equal(a, ctx.render_accum(ctx.cfg(160, 90, 2, 8, seed=1)))
    ctx.upload(scene.Parse(json.dumps(with_heightfield("example_simple", 64, 48, seed=8))))
    assert ctx.bvh_info()["build_ms"] != first["build_ms"] and ctx.bvh_info()["n_triangles"] == first["n_triangles"]
    ctx.upload(scene.Parse(json.dumps(scene_json("example_simple"))))
    assert ctx.bvh_info()["n_triangles"] == 0


@pytest.mark.gpu
def test_gpu_mesh_generation_skips_the_triangles_and_pipeline_matches(ctx, monkeypatch):
    """(1) ptb_scene.mesh_generation: the host mirror stamps every flattened scene with the identity of its mesh data, so the
    second upload of an unchanged Scene reuses the BVH without hashing the triangles, and a scene whose triangles differ gets
    another generation and another BVH.  (2) The mesh pipeline (wavefront across kernels, PTB_MESH_PIPELINE=1 — an alternative
    to the in-kernel traversal kept for A/B) renders the same per-pixel sums bit for bit: same helpers, same
    per-slot sample order."""
    from path_trace_golang_b200 import scene
    sc = scene.Parse(json.dumps(with_heightfield("test_comprehensive", 200, 150, pos=(0, 1.2, 2), size=(14, 2.5, 10))))
    assert sc.flat().mesh_generation != 0 and sc.flat().mesh_generation == sc.flat().mesh_generation
    other = scene.Parse(json.dumps(with_heightfield("test_comprehensive", 200, 150, pos=(0, 1.2, 2), size=(14, 2.5, 10), seed=5)))
    assert other.flat().mesh_generation != sc.flat().mesh_generation
    ctx.upload(sc)
    first = ctx.bvh_info()
    ctx.upload(sc)
    assert ctx.bvh_info() == first
    cfg = ctx.cfg(480, 270, 3, 10, seed=2)
    a = ctx.render_accum(cfg)
    assert ctx.last_kernel().startswith("integrate_wf_kernel<0, 1")
    monkeypatch.setenv("PTB_MESH_PIPELINE", "1")
    b = ctx.render_accum(cfg)
    assert ctx.last_kernel().startswith("mp_shade_scan_kernel")
    monkeypatch.delenv("PTB_MESH_PIPELINE")
    assert (a == b).all()
    ctx.render_accum(ctx.cfg(480, 270, 3, 10, seed=2, stats=True))
    assert ctx.stats()["bvh_stack_overflows"] == 0 and ctx.stats()["bvh_nodes_visited"] > 0


def test_bvh_builder_selfcheck_cpu():
    """Host-only: the emitted node array is sound — every child box (centre / half extent, binary32) contains all the
    triangles below it, every triangle is in exactly one leaf (ptb_bvh_selfcheck, no CUDA involved)."""
    import ctypes as C
    from path_trace_golang_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(7)
    cases = []
    cases.append(rng.uniform(-5, 5, size=(20000, 9)).astype(np.float32))                      # soup
    g = np.stack(np.meshgrid(np.arange(120), np.arange(80), indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    h = lambda p: np.sin(p[:, 0] * 0.3) * np.cos(p[:, 1] * 0.2)
    def vert(p): return np.stack([p[:, 0] * 0.1 - 6, h(p), p[:, 1] * 0.1 - 4], -1)
    a, b, c, d = vert(g), vert(g + [1, 0]), vert(g + [0, 1]), vert(g + [1, 1])
    cases.append(np.concatenate([np.concatenate([a, b, c], 1), np.concatenate([b, d, c], 1)]).astype(np.float32))   # heightfield
    flat = rng.uniform(-1, 1, size=(3000, 9)).astype(np.float32); flat[:, 1::3] = 0.25            # axis-aligned, zero thickness
    cases.append(flat)
    cases.append((rng.uniform(-1, 1, size=(5000, 9)) * 1e4 + 3e5).astype(np.float32))            # large coordinates
    cases.append(rng.uniform(-1, 1, size=(3, 9)).astype(np.float32))                             # fewer triangles than a leaf holds
    for tri in cases:
        tri = np.ascontiguousarray(tri)
        n_nodes, depth = C.c_int64(), C.c_int32()
        bad = L.ptb_bvh_selfcheck(tri.ctypes.data_as(C.POINTER(C.c_float)), len(tri), C.byref(n_nodes), C.byref(depth))
        assert bad == 0, (len(tri), bad)
        assert 1 <= n_nodes.value <= max(1, len(tri)) and 1 <= depth.value <= 38
