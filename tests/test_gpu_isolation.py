"""Isolation, limits and persistence of the CUDA backend (VERDICT r01 "weak" #9):
 * a launch carries its scene BY VALUE (kernel parameters): two contexts on one device render concurrently, frames queued on
   a stream survive a later ptb_scene_upload / a frame of another size;
 * no compiled-in limit on the number of analytic objects (sceneToWorld has none, objects.go:225-269): 2 000 spheres;
 * the shared-memory opt-in / occupancy are per device and per scene size: a small scene first, a large one later;
 * a progressive render and a render resumed from stored sums give the bytes of one call."""
import json
import threading

import numpy as np
import pytest

from conftest import SCENE_DEPTH, scene_json

pytestmark = pytest.mark.gpu


def sphere_field(n, seed=5, box_every=0):
    """A synthetic world of n small spheres (every box_every-th object a box) over a plane, in the reference's JSON schema."""
    rng = np.random.default_rng(seed)
    mats = [{"id": "floor", "type": "lambert", "albedo": {"r": 0.6, "g": 0.6, "b": 0.6}},
            {"id": "lamp", "type": "emissive", "emit": {"r": 1, "g": 0.9, "b": 0.8}, "power": 6},
            {"id": "glass", "type": "dielectric", "ior": 1.5, "absorption": {"r": 0.1, "g": 0.02, "b": 0.02}},
            {"id": "steel", "type": "metal", "albedo": {"r": 0.8, "g": 0.8, "b": 0.9}, "rough": 0.2}]
    for k in range(24):
        mats.append({"id": f"m{k}", "type": "lambert", "albedo": {"r": float(rng.uniform(0.1, 0.9)), "g": float(rng.uniform(0.1, 0.9)), "b": float(rng.uniform(0.1, 0.9))}})
    objs = [{"id": "ground", "type": "plane", "position": {"x": 0, "y": 0, "z": 0}, "size": {"x": 0, "y": 0, "z": 0}, "material_id": "floor"}]
    side = int(np.ceil(np.sqrt(n)))
    for i in range(n):
        gx, gz = i % side, i // side
        pos = {"x": float((gx - side / 2) * 0.5 + rng.uniform(-0.1, 0.1)), "y": float(rng.uniform(0.12, 0.6)), "z": float(-gz * 0.5 + rng.uniform(-0.1, 0.1))}
        r = float(rng.uniform(0.08, 0.2))
        mat = ["lamp", "glass", "steel"][i % 3] if i % 11 == 0 else f"m{i % 24}"
        if box_every and i % box_every == 0:
            objs.append({"id": f"b{i}", "type": "box", "position": pos, "size": {"x": 2 * r, "y": 2 * r, "z": 2 * r}, "material_id": mat})
        else:
            objs.append({"id": f"s{i}", "type": "sphere", "position": pos, "size": {"x": r, "y": 0, "z": 0}, "material_id": mat})
    return {"name": f"field{n}", "camera": {"position": {"x": 0, "y": 6, "z": 9}, "target": {"x": 0, "y": 0, "z": -side * 0.2}, "up": {"x": 0, "y": 1, "z": 0},
                                          "fov": 50, "aperture": 0, "focus_dist": 0, "aspect_ratio": 0},
            "objects": objs, "materials": mats, "settings": {"width": 320, "height": 180, "samples_per_px": 4, "max_depth": 6},
            "background": {"r": 0.2, "g": 0.3, "b": 0.5}, "sky": {"type": "gradient", "color": {"r": 0, "g": 0, "b": 0},
                                                                "horizon": {"r": 0.8, "g": 0.8, "b": 0.9}, "zenith": {"r": 0.2, "g": 0.4, "b": 0.9}}}


@pytest.mark.parametrize("n,box_every", [(2000, 0), (700, 3)])
def test_large_world_matches_oracle(n, box_every, ctx, oracle_mod):
    """More objects than any constant bank held (the r01 limit was 512): the BIG instantiation scans global-memory tables.
    Primary-hit ids and t bit-exact against the oracle at 640x360, 1-spp image path-for-path against the binary32 oracle,
    counters within 0.5 %."""
    from path_trace_golang_b200 import scene
    doc = sphere_field(n, box_every=box_every)
    sc = scene.Parse(json.dumps(doc))
    ora = oracle_mod.OracleScene(doc)
    ctx.upload(sc)
    assert len(ctx.world()) == n + 1
    ids, t = ctx.primary_hits(640, 360)
    oids, ot = ora.primary_hits(640, 360)
    assert (ids == oids).all() and (t.view(np.uint64) == ot.view(np.uint64)).all()
    assert len(np.unique(ids)) > n // 3                                   # the field really is in view
    W, H, depth = 256, 144, 6
    dev = ctx.render_accum(ctx.cfg(W, H, 1, depth, seed=3, stats=True)).astype(np.float64)
    st = ctx.stats()
    ref, ost = ora.render_sum(W, H, 1, depth, seed=3, precision=32)
    ok = (np.abs(dev - ref) <= 1e-3 * np.maximum(1.0, np.abs(ref))).all(axis=2).mean()
    print(f"{n} objects: path-for-path {ok:.4f}")
    # thousands of pixel-sized spheres: far more silhouette pixels than in the shipped scenes, where binary32 with approximate
    # reciprocals and the binary32 oracle may decide hit / miss differently (measured 0.990 with 2 000 spheres; shipped scenes 0.998+)
    assert ok >= 0.98
    # counters: 1 % of the oracle's, plus an allowance for the paths that diverged (each changes a counter by a few units; rare
    # events such as exit scans — ~450 in this frame — would otherwise be held to +-10 while ~370 paths differ)
    n_div = (1.0 - ok) * W * H
    for k in ["segments", "exit_scans", "scatters", "end_sky", "end_emissive"]:
        assert abs(st[k] - ost[k]) <= 0.01 * ost[k] + 0.05 * n_div + 5, (k, st[k], ost[k])


def test_parameter_table_boundary_matches_global_tables(ctx, oracle_mod, monkeypatch):
    """The same mid-size world through both table paths (kernel parameters vs global memory, forced with PTB_FORCE_BIG):
    identical sums, bit for bit — the two instantiations are the same arithmetic."""
    from path_trace_golang_b200 import scene
    doc = sphere_field(150, box_every=2)
    sc = scene.Parse(json.dumps(doc))
    cfg = ctx.cfg(320, 180, 4, 6, seed=8)
    ctx.upload(sc)
    small = ctx.render_accum(cfg)
    monkeypatch.setenv("PTB_FORCE_BIG", "1")
    ctx.upload(sc)
    big = ctx.render_accum(cfg)
    monkeypatch.delenv("PTB_FORCE_BIG")
    ctx.upload(sc)
    assert (small == big).all()
    ora = oracle_mod.OracleScene(doc)
    ref, _ = ora.render_sum(320, 180, 4, 6, seed=8, precision=32)
    assert (np.abs(small - ref) <= 4e-3 * np.maximum(1.0, np.abs(ref))).all(axis=2).mean() >= 0.99


def test_packed_sphere_instantiation_matches_scalar(ctx, host_scenes, monkeypatch):
    """Sphere-rich scenes run the instantiation that tests a thread's two rays against a sphere in packed binary32 arithmetic
    (integrator.cu "two rays per instruction"; chosen by kPackedSphereMin, forced here with PTB_PACKED).  Each packed half is the
    same round-to-nearest operation as the scalar form, so the sums must agree; the selection itself is checked by kernel name."""
    cfg = ctx.cfg(480, 270, 8, 10, seed=21)                        # (a small frame: rendered as sample sub-range work items)
    ctx.upload(host_scenes["metal_glass_room"])                    # 2 spheres: scalar
    ctx.render_accum(cfg)
    assert ctx.last_kernel() == "integrate_wf_kernel<0, 0, 0, 0> + finalize_planes_kernel"
    ctx.upload(host_scenes["test_scene"])                          # 11 spheres: packed
    auto = ctx.render_accum(cfg)
    assert ctx.last_kernel() == "integrate_wf_kernel<0, 0, 0, 1> + finalize_planes_kernel"
    monkeypatch.setenv("PTB_PACKED", "0")
    scalar = ctx.render_accum(cfg)
    assert ctx.last_kernel().startswith("integrate_wf_kernel<0, 0, 0, 0>")
    monkeypatch.setenv("PTB_PACKED", "1")
    packed = ctx.render_accum(cfg)
    monkeypatch.delenv("PTB_PACKED")
    assert (packed == auto).all()
    same = (packed == scalar).all(axis=2).mean()
    print(f"packed vs scalar sphere test: {same:.6f} of the pixels bit-identical")
    assert same == 1.0


def test_small_scene_then_large_scene_same_process(host_scenes):
    """the >48 KB shared-memory opt-in used to be taken once per process on the first scene's size.  A fresh
    context renders a 10-object scene, then one whose records need the opt-in (320 objects, 28 materials: the largest world that
    still takes the kernel-parameter path)."""
    from path_trace_golang_b200 import engine, scene
    c = engine.Context(0)
    try:
        c.upload(host_scenes["example_simple"])
        a = c.render(c.cfg(160, 90, 2, 8, seed=1))
        c.upload(scene.Parse(json.dumps(sphere_field(320, box_every=4))))      # 38.7 KB of path state + 11.6 KB of records > 48 KB
        b = c.render(c.cfg(160, 90, 2, 6, seed=1))
        c.upload(host_scenes["example_simple"])
        a2 = c.render(c.cfg(160, 90, 2, 8, seed=1))
        assert b[..., :3].any() and (a == a2).all()
    finally:
        c.close()


def test_rejected_upload_keeps_the_previous_scene(ctx, host_scenes):
    """a failed ptb_scene_upload used to leave the context half-updated.  The upload is staged now: a scene that is
    rejected (material index out of range) changes nothing, and the next render equals the one before."""
    import ctypes as C
    from path_trace_golang_b200 import _lib, engine
    sc = host_scenes["test_scene"]
    ctx.upload(sc)
    cfg = ctx.cfg(200, 120, 3, 10, seed=4)
    before = ctx.render(cfg)
    flat = sc.flat()
    bad = _lib.PtbScene()
    C.memmove(C.byref(bad), C.byref(flat), C.sizeof(flat))
    mats = (C.c_int32 * flat.n_obj)(*[flat.obj_mat[i] for i in range(flat.n_obj)])
    mats[flat.n_obj - 1] = flat.n_mat + 5                       # out of range: build_world rejects the scene
    bad.obj_mat = mats
    rc = _lib.lib().ptb_scene_upload(ctx._h, C.byref(bad))
    assert rc == _lib.PTB_ERR_INVALID and b"out of range" in _lib.lib().ptb_last_error(ctx._h)
    assert len(ctx.world()) > 0 and (ctx.render(cfg) == before).all()


def test_two_contexts_render_concurrently_on_one_device(host_scenes):
    """Two contexts, two host threads, one device, different scenes and frame sizes, many frames each: every frame equals the
    frame the same context renders alone.  (r01 kept the scene in one __constant__ symbol per device: this corrupted frames.)"""
    from path_trace_golang_b200 import engine
    jobs = [("metal_glass_room", 480, 270, 8), ("test_scene", 400, 300, 6)]
    ctxs, want = [], []
    for name, W, H, spp in jobs:
        c = engine.Context(0)
        c.upload(host_scenes[name])
        ctxs.append(c)
        want.append(c.render(c.cfg(W, H, spp, SCENE_DEPTH[name], seed=4)))
    bad = [0, 0]

    def work(k):
        name, W, H, spp = jobs[k]
        for _ in range(12):
            img = ctxs[k].render(ctxs[k].cfg(W, H, spp, SCENE_DEPTH[name], seed=4))
            bad[k] += int((img != want[k]).any())

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    [c.close() for c in ctxs]
    assert bad == [0, 0]


def test_queued_frames_survive_a_later_upload_and_other_sizes(ctx, host_scenes):
    """frames queued asynchronously (ptb_render_accum_device on a caller stream) used to read the pinned scene
    header at execution time.  Queue a 4K-ish frame, then a small frame of another size, then upload ANOTHER scene: the
    results equal the frames rendered one at a time."""
    import torch
    dev = torch.device("cuda", 0)
    name = "metal_glass_room"
    ctx.upload(host_scenes[name])
    big_cfg = ctx.cfg(1920, 1080, 16, 16, seed=2)
    small_cfg = ctx.cfg(333, 211, 4, 16, seed=2)
    want_big = torch.from_numpy(ctx.render_accum(big_cfg))
    want_small = torch.from_numpy(ctx.render_accum(small_cfg))
    a = torch.empty((1080, 1920, 3), dtype=torch.float32, device=dev)
    b = torch.empty((211, 333, 3), dtype=torch.float32, device=dev)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        ctx.render_accum_device(big_cfg, a.data_ptr(), s.cuda_stream)
        ctx.render_accum_device(small_cfg, b.data_ptr(), s.cuda_stream)
    ctx.upload(host_scenes["test_scene"])                  # (waits for the device, then replaces every table)
    s.synchronize()
    assert torch.equal(a.cpu(), want_big) and torch.equal(b.cpu(), want_small)


@pytest.mark.parametrize("W,H,spp", [(400, 225, 20), (320, 180, 64), (1600, 900, 12)])
def test_progressive_render_equals_one_launch(W, H, spp, ctx, host_scenes):
    """ptb_render with a progress callback (sample batches, asynchronous double-buffered read-back) ends with exactly the
    bytes of the single fused launch, also for SPLIT small frames (the batches used to regroup the partial sums);
    progress() fires after every refresh and once more at the end (renderer.go:226-245), partial images are non-decreasing
    in coverage."""
    name = "example_simple"
    ctx.upload(host_scenes[name])
    cfg = ctx.cfg(W, H, spp, 8, seed=6)
    one = ctx.render(cfg)
    calls, partial = [], []
    out = np.zeros((H, W, 4), dtype=np.uint8)

    def progress():
        calls.append(1)
        partial.append(int(out[..., :3].astype(np.int64).sum()))

    ctx.render(cfg, out=out, progress=progress)
    assert (out == one).all()
    assert 3 <= len(calls) <= 12 and partial[0] > 0


def test_resume_from_stored_sums_is_bit_exact(ctx, host_scenes, tmp_path):
    """ptb_render_resume: [0, 7) + [7, 19) + [19, 32) with the sums round-tripped through a file == one call over [0, 32),
    sums and image, bit for bit; then the same through the checkpoint file of the host mirror, interrupted twice."""
    from path_trace_golang_b200 import engine
    name, W, H, spp, depth = "metal_glass_room", 640, 360, 32, 16
    sc = host_scenes[name]
    ctx.upload(sc)
    whole = np.zeros((H, W, 3), dtype=np.float32)
    img_whole = np.zeros((H, W, 4), dtype=np.uint8)
    ctx.render_resume(ctx.cfg(W, H, spp, depth, seed=3), whole, img_whole)
    sums = np.zeros((H, W, 3), dtype=np.float32)
    img = np.zeros((H, W, 4), dtype=np.uint8)
    for b, e in [(0, 7), (7, 19), (19, 32)]:
        ctx.render_resume(ctx.cfg(W, H, spp, depth, seed=3, sample_begin=b, sample_count=e - b), sums, img)
        np.save(tmp_path / "sums.npy", sums)
        sums = np.load(tmp_path / "sums.npy")
    assert (sums == whole).all() and (img == img_whole).all()
    assert (ctx.finalize_host(whole, spp) == img_whole).all()

    ck = tmp_path / "render.ptbacc"
    out = np.zeros((H, W, 4), dtype=np.uint8)
    cfg = engine.RenderConfig(W, H, spp, depth)
    assert engine.RenderCheckpointed(sc, cfg, out, ck, samples_per_call=5, max_calls=2, ctx=ctx, seed=3) == 10
    assert ck.exists() and (out != img_whole).any()
    assert engine.RenderCheckpointed(sc, cfg, out, ck, samples_per_call=9, max_calls=1, ctx=ctx, seed=3) == 19
    assert engine.RenderCheckpointed(sc, cfg, out, ck, samples_per_call=9, ctx=ctx, seed=3) == 32
    assert (out == img_whole).all()
    # a finished checkpoint only re-runs the epilogue; a checkpoint of another render (other seed) is ignored
    out2 = np.zeros_like(out)
    assert engine.RenderCheckpointed(sc, cfg, out2, ck, samples_per_call=9, ctx=ctx, seed=3) == 32 and (out2 == img_whole).all()
    assert engine.RenderCheckpointed(sc, cfg, out2, ck, samples_per_call=32, ctx=ctx, seed=4) == 32 and (out2 != img_whole).any()


def test_pinned_caller_image_direct_copy(ctx, host_scenes):
    """ptb_host_buffer_pin: the image lands in the caller's page-locked buffer without the staging copy — same bytes,
    also with a row stride."""
    name = "example_simple"
    ctx.upload(host_scenes[name])
    cfg = ctx.cfg(320, 180, 4, 8, seed=2)
    want = ctx.render(cfg)
    buf = np.zeros((180, 320 + 16, 4), dtype=np.uint8)
    ctx.pin(buf)
    try:
        ctx.render(cfg, out=buf[:, :320])
        assert (buf[:, :320] == want).all() and (buf[:, 320:] == 0).all()
        tight = np.zeros((180, 320, 4), dtype=np.uint8)
        ctx.pin(tight)
        try:
            ctx.render(cfg, out=tight)
            assert (tight == want).all()
        finally:
            ctx.unpin(tight)
    finally:
        ctx.unpin(buf)


def test_exact_zero_direction_component_misses_box_outside_slab(ctx, oracle_mod):
    """A direction component of exactly 0: the reference computes tNear = tFar = +-inf on that axis and misses a box
    whose slab excludes the origin (objects.go:149-165).  Camera looking exactly along -z, pixel-centre column with d.x == 0,
    box off to the side: the binary64 parity kernel must miss it like the oracle.  (The binary32 integrator keeps 1/d finite
    for such components — make_ray, integrator.cu — but jittered samples cannot produce an exact 0 on demand; its handling
    is covered statistically by the path-for-path tests.)"""
    from path_trace_golang_b200 import scene
    doc = {"name": "zero", "camera": {"position": {"x": 0, "y": 0, "z": 5}, "target": {"x": 0, "y": 0, "z": 0}, "up": {"x": 0, "y": 1, "z": 0},
                                      "fov": 40, "aperture": 0, "focus_dist": 0, "aspect_ratio": 0},
           "objects": [{"id": "b", "type": "box", "position": {"x": 1.0, "y": 0, "z": 0}, "size": {"x": 1, "y": 50, "z": 1}, "material_id": "e"}],
           "materials": [{"id": "e", "type": "emissive", "emit": {"r": 1, "g": 1, "b": 1}, "power": 1}],
           "settings": {"width": 65, "height": 65, "samples_per_px": 1, "max_depth": 4}, "background": {"r": 0, "g": 0, "b": 0}, "sky": None}
    sc = scene.Parse(json.dumps(doc))
    ctx.upload(sc)
    W = H = 65
    # column 32 of a 65-wide frame with xi_u = 0 is u = 0.5 exactly: horizontal * 0.5 cancels llc.x exactly -> d.x == 0
    ids, _ = ctx.primary_hits(W, H, 0.0, 0.0)
    oids, _ = oracle_mod.OracleScene(doc).primary_hits(W, H, 0.0, 0.0)
    assert (ids == oids).all() and (ids[:, 32] == -1).all()
