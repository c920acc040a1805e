"""Python veneer over the C++ mirror of internal/scene (scene.go, io.go): Load / Save / settings.

All parsing, flattening and marshalling happens in C++ (csrc/host/scene.cpp); this module only holds the
handle and exposes the reference's names.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import PtbError, PtbScene


class RenderSettings:
    """scene.RenderSettings (scene.go:92-97)."""

    def __init__(self, Width=0, Height=0, SamplesPerPx=0, MaxDepth=0):
        self.Width, self.Height, self.SamplesPerPx, self.MaxDepth = Width, Height, SamplesPerPx, MaxDepth

    def __repr__(self):
        return f"RenderSettings({self.Width}x{self.Height}, spp={self.SamplesPerPx}, depth={self.MaxDepth})"


class Scene:
    """Handle to a decoded scene (scene.Scene, scene.go:146-158) living in the C++ host layer."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().ptb_host_scene_free(h)
            except Exception:
                pass

    @property
    def Settings(self) -> RenderSettings:
        out = (C.c_int32 * 4)()
        _lib.lib().ptb_host_scene_settings(self._h, out)
        return RenderSettings(*out)

    def counts(self):
        a, b = C.c_int32(), C.c_int32()
        _lib.lib().ptb_host_scene_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def flat(self) -> PtbScene:
        """The SoA view passed to ptb_scene_upload (valid while this Scene is alive)."""
        s = PtbScene()
        rc = _lib.lib().ptb_host_scene_flat(self._h, C.byref(s))
        if rc:
            raise PtbError(rc, _lib.lib().ptb_host_last_error().decode())
        return s

    def mesh_triangles(self) -> dict:
        """EXTENSION: {object index: float32 (n_tri, 9) world-space triangles} of the "mesh" objects, as flattened."""
        import numpy as np
        f = self.flat()
        out = {}
        for i in range(f.n_obj):
            m = f.obj_mesh[i] if f.n_mesh > 0 else -1
            if m < 0:
                continue
            t0, t1 = f.mesh_tri_begin[m], f.mesh_tri_begin[m + 1]
            if t1 > t0:
                out[i] = np.ctypeslib.as_array(f.tri_vertices, shape=(f.mesh_tri_begin[f.n_mesh] * 9,))[t0 * 9:t1 * 9].reshape(-1, 9).copy()
        return out

    def marshal(self) -> str:
        L = _lib.lib()
        n = L.ptb_host_scene_marshal(self._h, None, 0)
        buf = C.create_string_buffer(n)
        L.ptb_host_scene_marshal(self._h, buf, n)
        return buf.raw.decode("utf-8")


def Load(path) -> Scene:
    """scene.Load (io.go:10-22)."""
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.ptb_host_scene_load(str(path).encode(), C.byref(h))
    if rc:
        raise PtbError(rc, L.ptb_host_last_error().decode())
    return Scene(h)


def Parse(text: str) -> Scene:
    L = _lib.lib()
    h = C.c_void_p()
    b = text.encode("utf-8")
    rc = L.ptb_host_scene_parse(b, len(b), C.byref(h))
    if rc:
        raise PtbError(rc, L.ptb_host_last_error().decode())
    return Scene(h)


def Save(path, sc: Scene) -> None:
    """scene.Save (io.go:25-38)."""
    L = _lib.lib()
    rc = L.ptb_host_scene_save(sc._h, str(path).encode())
    if rc:
        raise PtbError(rc, L.ptb_host_last_error().decode())
