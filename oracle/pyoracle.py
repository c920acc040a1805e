"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/oracle.h.  Import from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs,
never from the product package.  PARITY UNPINNED (no reference golden vectors); cross-checked bit for bit
against the independent Python restatement oracle/goref.py (tests/test_oracle_crosscheck.py).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


class RawMaterial(C.Structure):
    _fields_ = [("id", C.c_char_p), ("type", C.c_char_p), ("albedo", C.c_double * 3), ("rough", C.c_double),
                ("ior", C.c_double), ("emit", C.c_double * 3), ("power", C.c_double),
                ("absorption", C.c_double * 3), ("smoothness", C.c_double)]


class RawObject(C.Structure):
    _fields_ = [("type", C.c_char_p), ("position", C.c_double * 3), ("size", C.c_double * 3),
                ("material_id", C.c_char_p), ("tri_vertices", C.POINTER(C.c_float)), ("n_tri", C.c_int64)]


class RawCamera(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("target", C.c_double * 3), ("up", C.c_double * 3),
                ("fov", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
                ("aspect_ratio", C.c_double)]


class RawSky(C.Structure):
    _fields_ = [("has_sky", C.c_int), ("sky_type", C.c_char_p), ("color", C.c_double * 3),
                ("horizon", C.c_double * 3), ("zenith", C.c_double * 3), ("background", C.c_double * 3)]


class WorldEntry(C.Structure):
    _fields_ = [("type", C.c_int32), ("mat_type", C.c_int32), ("a", C.c_double * 3), ("b", C.c_double * 3),
                ("albedo", C.c_double * 3), ("rough", C.c_double), ("ior", C.c_double),
                ("emit", C.c_double * 3), ("absorption", C.c_double * 3)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("segments", C.c_uint64), ("exit_scans", C.c_uint64),
                ("prim_tests", C.c_uint64), ("accepts", C.c_uint64 * 3), ("scatters", C.c_uint64),
                ("end_sky", C.c_uint64), ("end_emissive", C.c_uint64), ("end_rr", C.c_uint64),
                ("end_depth", C.c_uint64), ("end_noscatter", C.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "accepts"}
        d["accepts"] = list(self.accepts)
        return d


def build(force: bool = False) -> pathlib.Path:
    so = _HERE / "liboracle.so"
    src = [_HERE / "oracle.cpp", _HERE / "oracle.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in src if s.exists()):
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(RawObject), C.c_int, C.POINTER(RawMaterial), C.c_int,
                                       C.POINTER(RawCamera), C.POINTER(RawSky)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_world_size.argtypes = [C.c_void_p]
        L.orc_world_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(WorldEntry)]
        L.orc_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_primary_hits.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_render_sum.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                     C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.orc_finalize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_render_rgba.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                      C.c_void_p, C.POINTER(Stats)]
        L.orc_trace_path.restype = C.c_int
        L.orc_trace_path.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                     C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_double)]
        L.orc_set_mesh_accel.argtypes = [C.c_void_p, C.c_int]
        L.orc_rng_uniform.restype = C.c_double
        L.orc_rng_uniform.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _LIB = L
    return _LIB


def _v3(d, keys):
    d = d or {}
    return (C.c_double * 3)(*(float(d.get(k, 0) or 0) for k in keys))


def _b(s):
    return (s or "").encode("utf-8")


class OracleScene:
    """Scene handle built from a decoded scene JSON dict (schema: internal/scene/scene.go)."""

    def __init__(self, sc: dict, mesh_triangles: dict | None = None):
        """mesh_triangles: {object index: float32 array (n_tri, 9) of WORLD-space triangles} for objects of type "mesh"
        (EXTENSION).  Tests take them from the product's flattener so that the mesh generator is not part of the
        parity surface; inline "mesh": {"vertices", "triangles"} objects are expanded here."""
        L = lib()
        mats = sc.get("materials") or []
        objs = sc.get("objects") or []
        self._keep = []
        RM = (RawMaterial * max(1, len(mats)))()
        for i, m in enumerate(mats):
            RM[i] = RawMaterial(_b(m.get("id")), _b(m.get("type")), _v3(m.get("albedo"), "rgb"),
                                float(m.get("rough", 0) or 0), float(m.get("ior", 0) or 0),
                                _v3(m.get("emit"), "rgb"), float(m.get("power", 0) or 0),
                                _v3(m.get("absorption"), "rgb"), float(m.get("smoothness", 0) or 0))
        RO = (RawObject * max(1, len(objs)))()
        for i, o in enumerate(objs):
            RO[i] = RawObject(_b(o.get("type")), _v3(o.get("position"), "xyz"), _v3(o.get("size"), "xyz"),
                              _b(o.get("material_id")), None, 0)
            if o.get("type") == "mesh":
                tri = (mesh_triangles or {}).get(i)
                if tri is None and isinstance(o.get("mesh"), dict) and "vertices" in o["mesh"]:
                    v = np.asarray(o["mesh"]["vertices"], dtype=np.float32).reshape(-1, 3).astype(np.float64)
                    idx = np.asarray(o["mesh"]["triangles"], dtype=np.int64).reshape(-1, 3)
                    pos = np.array([float((o.get("position") or {}).get(k, 0)) for k in "xyz"])
                    size = np.array([float((o.get("size") or {}).get(k, 0)) or 1.0 for k in "xyz"])
                    tri = (pos + size * v)[idx].reshape(-1, 9).astype(np.float32)
                if tri is not None and len(tri):
                    tri = np.ascontiguousarray(tri, dtype=np.float32).reshape(-1, 9)
                    self._keep.append(tri)
                    RO[i].tri_vertices = tri.ctypes.data_as(C.POINTER(C.c_float))
                    RO[i].n_tri = len(tri)
        cam = sc.get("camera") or {}
        rc = RawCamera(_v3(cam.get("position"), "xyz"), _v3(cam.get("target"), "xyz"), _v3(cam.get("up"), "xyz"),
                       float(cam.get("fov", 0) or 0), float(cam.get("aperture", 0) or 0),
                       float(cam.get("focus_dist", 0) or 0), float(cam.get("aspect_ratio", 0) or 0))
        sky = sc.get("sky")
        rs = RawSky(1 if sky is not None else 0, _b((sky or {}).get("type")), _v3((sky or {}).get("color"), "rgb"),
                    _v3((sky or {}).get("horizon"), "rgb"), _v3((sky or {}).get("zenith"), "rgb"),
                    _v3(sc.get("background"), "rgb"))
        self._h = L.orc_scene_create(RO, len(objs), RM, len(mats), C.byref(rc), C.byref(rs))

    @classmethod
    def load(cls, path):
        with open(path) as f:
            return cls(json.load(f))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_mesh_accel(self, enabled: bool):
        lib().orc_set_mesh_accel(self._h, 1 if enabled else 0)

    def world(self):
        L = lib()
        out = []
        for i in range(L.orc_world_size(self._h)):
            e = WorldEntry()
            L.orc_world_get(self._h, i, C.byref(e))
            out.append(dict(type=e.type, mat_type=e.mat_type, a=list(e.a), b=list(e.b), albedo=list(e.albedo),
                            rough=e.rough, ior=e.ior, emit=list(e.emit), absorption=list(e.absorption)))
        return out

    def camera(self, w, h):
        out = (C.c_double * 22)()
        lib().orc_camera(self._h, w, h, out)
        return np.array(out, dtype=np.float64)

    def primary_hits(self, w, h, xi_u=0.5, xi_v=0.5):
        ids = np.empty((h, w), dtype=np.int32)
        t = np.empty((h, w), dtype=np.float64)
        lib().orc_primary_hits(self._h, w, h, xi_u, xi_v, ids.ctypes.data, t.ctypes.data)
        return ids, t

    def render_sum(self, w, h, spp, max_depth, seed=1, precision=64, threads=None, s_begin=0):
        out = np.empty((h, w, 3), dtype=np.float64)
        st = Stats()
        threads = threads or (os.cpu_count() or 1)
        lib().orc_render_sum(self._h, w, h, s_begin, s_begin + spp, max_depth, seed, precision, threads,
                             out.ctypes.data, C.byref(st))
        return out, st.as_dict()

    def render_rgba(self, w, h, spp, max_depth, seed=1, threads=None):
        out = np.empty((h, w, 4), dtype=np.uint8)
        st = Stats()
        threads = threads or (os.cpu_count() or 1)
        lib().orc_render_rgba(self._h, w, h, spp, max_depth, seed, threads, out.ctypes.data, C.byref(st))
        return out, st.as_dict()

    def trace_path(self, orig, direction, max_depth, seed=1, cap=256):
        ids = np.full(cap, -2, dtype=np.int32)
        t = np.zeros(cap, dtype=np.float64)
        ff = np.zeros(cap, dtype=np.int32)
        rgb = (C.c_double * 3)()
        n = lib().orc_trace_path(self._h, (C.c_double * 3)(*orig), (C.c_double * 3)(*direction), max_depth, seed,
                                 cap, ids.ctypes.data, t.ctypes.data, ff.ctypes.data, rgb)
        return ids[:n], t[:n], ff[:n], np.array(rgb)


def finalize(rgb_sum: np.ndarray, spp: int) -> np.ndarray:
    h, w, _ = rgb_sum.shape
    rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float64)
    out = np.empty((h, w, 4), dtype=np.uint8)
    lib().orc_finalize(rgb_sum.ctypes.data, w, h, spp, out.ctypes.data)
    return out


def rng_uniform(seed, pixel, sample, i):
    return lib().orc_rng_uniform(seed, pixel, sample, i)
