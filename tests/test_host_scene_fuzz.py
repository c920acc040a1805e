"""Property test of the C++ scene decoder/flattener (csrc/host/scene.cpp) against Python's json + the flattening rules of
objects.go:225-269 / materials.go:28-55 written out here: random scenes (unknown keys, missing fields, duplicate
material ids, unknown types, odd numbers) must flatten to exactly the same SoA arrays, and Save -> Load must round-trip."""
import json

import numpy as np
from hypothesis import given, settings, strategies as st

MAT_CODE = {"metal": 1, "dielectric": 2, "emissive": 3, "mirror": 4}
OBJ_CODE = {"sphere": 0, "sphere_light": 0, "plane": 1, "box": 2}

num = st.one_of(st.integers(-50, 50), st.floats(-1e3, 1e3, allow_nan=False, allow_infinity=False, width=64),
                st.sampled_from([0, 0.0, 1e-9, 1.5, 1e21, -2.5e-7, 123456789.125]))
vec = st.fixed_dictionaries({}, optional={"x": num, "y": num, "z": num, "w": num})
col = st.fixed_dictionaries({}, optional={"r": num, "g": num, "b": num})
ident = st.sampled_from(["a", "b", "c", "glass", "m-1", "", "ünï", 'q"uote'])
material = st.fixed_dictionaries({}, optional={
    "id": ident, "type": st.sampled_from(["lambert", "metal", "dielectric", "emissive", "mirror", "velvet", ""]),
    "albedo": col, "rough": num, "ior": num, "emit": col, "power": num, "absorption": col, "smoothness": num,
    "reflectivity": num, "tint": col, "absorption_scale": num, "comment": st.text(max_size=8)})
obj = st.fixed_dictionaries({}, optional={
    "id": ident, "type": st.sampled_from(["sphere", "plane", "box", "sphere_light", "torus", ""]),
    "position": vec, "size": vec, "material_id": ident, "extra": st.lists(st.integers(), max_size=3)})
sky = st.one_of(st.none(), st.fixed_dictionaries({}, optional={"type": st.sampled_from(["solid", "gradient", "other"]),
                                                               "color": col, "horizon": col, "zenith": col}))
scene_doc = st.fixed_dictionaries({}, optional={
    "name": st.text(max_size=12), "camera": st.fixed_dictionaries({}, optional={
        "position": vec, "target": vec, "up": vec, "fov": num, "aperture": num, "focus_dist": num, "aspect_ratio": num}),
    "objects": st.lists(obj, max_size=6), "materials": st.lists(material, max_size=5), "background": col, "sky": sky,
    "settings": st.fixed_dictionaries({}, optional={"width": st.integers(0, 4000), "height": st.integers(0, 4000),
                                                    "samples_per_px": st.integers(0, 2000), "max_depth": st.integers(0, 100)}),
    "unknown_top_level": st.dictionaries(st.text(max_size=4), st.integers(), max_size=2)})


def f(d, k):
    return float((d or {}).get(k, 0))


def arr(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype).copy() if n else np.zeros(0, dtype)


@settings(max_examples=150, deadline=None)
@given(scene_doc)
def test_decoder_and_flattener_match_the_rules(doc):
    from path_trace_golang_b200 import scene
    sc = scene.Parse(json.dumps(doc))
    flat = sc.flat()
    mats, objs = doc.get("materials") or [], doc.get("objects") or []
    assert flat.n_mat == len(mats) and flat.n_obj == len(objs)
    assert list(arr(flat.mat_type, flat.n_mat, np.int32)) == [MAT_CODE.get(m.get("type", ""), 0) for m in mats]
    assert list(arr(flat.obj_type, flat.n_obj, np.int32)) == [OBJ_CODE.get(o.get("type", ""), -1) for o in objs]
    last = {}
    for i, m in enumerate(mats):
        last[m.get("id", "")] = i                                    # later duplicate id wins (objects.go:226-229)
    assert list(arr(flat.obj_mat, flat.n_obj, np.int32)) == [last.get(o.get("material_id", ""), -1) for o in objs]
    pos, size = arr(flat.obj_pos, 3 * flat.n_obj, np.float64), arr(flat.obj_size, 3 * flat.n_obj, np.float64)
    for i, o in enumerate(objs):
        assert list(pos[3 * i:3 * i + 3]) == [f(o.get("position"), k) for k in "xyz"]
        assert list(size[3 * i:3 * i + 3]) == [f(o.get("size"), k) for k in "xyz"]
    alb, emit = arr(flat.mat_albedo, 3 * flat.n_mat, np.float64), arr(flat.mat_emit, 3 * flat.n_mat, np.float64)
    for i, m in enumerate(mats):
        assert list(alb[3 * i:3 * i + 3]) == [f(m.get("albedo"), k) for k in "rgb"]
        assert list(emit[3 * i:3 * i + 3]) == [f(m.get("emit"), k) for k in "rgb"]
        assert flat.mat_rough[i] == float(m.get("rough", 0)) and flat.mat_ior[i] == float(m.get("ior", 0))
        assert flat.mat_power[i] == float(m.get("power", 0)) and flat.mat_smoothness[i] == float(m.get("smoothness", 0))
    cam = doc.get("camera") or {}
    assert list(flat.camera.position) == [f(cam.get("position"), k) for k in "xyz"]
    assert (flat.camera.fov, flat.camera.aperture, flat.camera.focus_dist, flat.camera.aspect_ratio) == tuple(
        float(cam.get(k, 0)) for k in ("fov", "aperture", "focus_dist", "aspect_ratio"))
    s = doc.get("sky")
    if s and s.get("type") == "gradient":                            # renderer.go:56-79
        assert flat.sky.kind == 1 and list(flat.sky.horizon) == [f(s.get("horizon"), k) for k in "rgb"]
        assert list(flat.sky.zenith) == [f(s.get("zenith"), k) for k in "rgb"]
    else:                                                            # renderer.go:80-92
        want = s.get("color") if (s and s.get("type") == "solid") else doc.get("background")
        assert flat.sky.kind == 0 and list(flat.sky.color) == [f(want, k) for k in "rgb"]
    st_ = doc.get("settings") or {}
    got = sc.Settings
    assert (got.Width, got.Height, got.SamplesPerPx, got.MaxDepth) == tuple(st_.get(k, 0) for k in ("width", "height", "samples_per_px", "max_depth"))
    # Save -> Load round trip: identical flattening, and the marshalled text is a fixed point
    again = scene.Parse(sc.marshal())
    assert again.marshal() == sc.marshal()
    f2 = again.flat()
    assert list(arr(f2.obj_pos, 3 * f2.n_obj, np.float64)) == list(pos) and list(arr(f2.mat_albedo, 3 * f2.n_mat, np.float64)) == list(alb)
    assert json.loads(sc.marshal())["name"] == doc.get("name", "")
