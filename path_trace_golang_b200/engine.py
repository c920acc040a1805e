"""Python veneer over libptb200.so with the reference's engine API names (internal/engine):
RenderConfig, Render, RenderInto, RenderScene, RenderSettingsForMode, SavePNG, SetBackend/GetBackend.

Images are numpy uint8 arrays of shape (H, W, 4) — the Pix/Stride layout of image.RGBA.  All compute is
in the CUDA library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PROGRESS_FN, PtbBvhInfo, PtbCfg, PtbDeviceInfo, PtbError, PtbStats, PTB_FLAG_MEGAKERNEL, PTB_FLAG_STATS
from .scene import RenderSettings, Scene

BackendCPU, BackendGPU, BackendCUDA = 0, 1, 2      # backend.go:7-10 + the CUDA backend
_backend = BackendCUDA
_default_ctx = None


class RenderConfig:
    """engine.RenderConfig (renderer.go:17-22)."""

    def __init__(self, Width, Height, SamplesPerPx, MaxDepth):
        self.Width, self.Height, self.SamplesPerPx, self.MaxDepth = int(Width), int(Height), int(SamplesPerPx), int(MaxDepth)


def SetBackend(b):
    """backend.go:16-23. Unknown values select the CPU backend — which this build does not have."""
    global _backend
    _backend = b if b in (BackendCPU, BackendGPU, BackendCUDA) else BackendCPU


def GetBackend():
    return _backend


class Context:
    """One ptb_ctx = one CUDA device."""

    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.ptb_create(int(device), C.byref(h))
        if rc:
            raise PtbError(rc, self._L.ptb_last_error(None).decode())
        self._h = h
        self.device = device
        self._scene = None

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.ptb_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise PtbError(rc, self._L.ptb_last_error(self._h).decode())

    def device_info(self) -> dict:
        d = PtbDeviceInfo()
        self._check(self._L.ptb_get_device_info(self._h, C.byref(d)))
        return dict(name=d.name.decode(), sm_count=d.sm_count, cc=(d.cc_major, d.cc_minor), clock_khz=d.clock_khz,
                    global_mem_bytes=d.global_mem_bytes)

    def upload(self, sc: Scene):
        flat = sc.flat()
        self._check(self._L.ptb_scene_upload(self._h, C.byref(flat)))
        self._scene = sc

    def world(self):
        n = self._L.ptb_world_size(self._h)
        if n < 0:
            self._check(n)
        out = []
        buf = (C.c_double * 19)()
        for i in range(n):
            self._check(self._L.ptb_world_get(self._h, i, buf))
            v = list(buf)
            out.append(dict(type=int(v[0]), mat_type=int(v[1]), a=v[2:5], b=v[5:8], albedo=v[8:11], rough=v[11],
                            ior=v[12], emit=v[13:16], absorption=v[16:19]))
        return out

    @staticmethod
    def cfg(width, height, spp, max_depth, seed=1, sample_begin=0, sample_count=0, stats=False, megakernel=False,
            row_offset=0, row_step=0) -> PtbCfg:
        flags = (PTB_FLAG_STATS if stats else 0) | (PTB_FLAG_MEGAKERNEL if megakernel else 0)
        return PtbCfg(int(width), int(height), int(spp), int(max_depth), int(seed) & 0xFFFFFFFF, int(sample_begin),
                      int(sample_count), flags, int(row_offset), int(row_step))

    @staticmethod
    def rows_of(cfg: PtbCfg) -> int:
        """Rows a call with this cfg outputs (row partition: a compact subset of the frame's rows)."""
        return cfg.height if cfg.row_step <= 1 else (cfg.height - cfg.row_offset + cfg.row_step - 1) // cfg.row_step

    def render(self, cfg: PtbCfg, out: np.ndarray | None = None, progress=None) -> np.ndarray:
        """ptb_render: host RGBA8 image (H, W, 4)."""
        if out is None:
            out = np.zeros((self.rows_of(cfg), cfg.width, 4), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.ndim == 3 and out.shape[2] == 4 and out.strides[1] == 4 and out.strides[2] == 1
        cb = PROGRESS_FN(lambda _u: progress()) if progress is not None else None
        self._check(self._L.ptb_render(self._h, C.byref(cfg), out.ctypes.data, out.strides[0],
                                       C.cast(cb, C.c_void_p) if cb else None, None))
        return out

    def render_accum(self, cfg: PtbCfg) -> np.ndarray:
        """ptb_render_accum: host fp32 sums (H, W, 3) of the cfg's sample range."""
        out = np.empty((self.rows_of(cfg), cfg.width, 3), dtype=np.float32)
        self._check(self._L.ptb_render_accum(self._h, C.byref(cfg), out.ctypes.data))
        return out

    def render_resume(self, cfg: PtbCfg, sums: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """ptb_render_resume: continue the fp32 sums (H, W, 3) over cfg's sample range, in place; returns them."""
        assert sums.dtype == np.float32 and sums.flags.c_contiguous and sums.shape == (cfg.height, cfg.width, 3)
        if out is not None:
            assert out.dtype == np.uint8 and out.shape == (cfg.height, cfg.width, 4) and out.strides[1] == 4
        self._check(self._L.ptb_render_resume(self._h, C.byref(cfg), sums.ctypes.data, out.ctypes.data if out is not None else None,
                                              out.strides[0] if out is not None else 0))
        return sums

    def finalize_host(self, sums: np.ndarray, spp_total: int) -> np.ndarray:
        """ptb_finalize_host: pixel epilogue (on the device) of host fp32 sums (H, W, 3) -> RGBA8 (H, W, 4)."""
        sums = np.ascontiguousarray(sums, dtype=np.float32)
        out = np.empty(sums.shape[:2] + (4,), dtype=np.uint8)
        self._check(self._L.ptb_finalize_host(self._h, sums.ctypes.data, sums.shape[1], sums.shape[0], int(spp_total), out.ctypes.data, out.strides[0]))
        return out

    def pin(self, arr: np.ndarray):
        """ptb_host_buffer_pin: page-lock a caller-owned image so that renders copy straight into it."""
        self._check(self._L.ptb_host_buffer_pin(self._h, arr.ctypes.data, arr.nbytes))

    def unpin(self, arr: np.ndarray):
        self._check(self._L.ptb_host_buffer_unpin(self._h, arr.ctypes.data))

    def render_accum_device(self, cfg: PtbCfg, d_ptr: int, stream: int = 0):
        self._check(self._L.ptb_render_accum_device(self._h, C.byref(cfg), C.c_void_p(d_ptr), C.c_void_p(stream)))

    def finalize_device(self, d_accum: int, width, height, spp_total, d_rgba: int, stream: int = 0):
        self._check(self._L.ptb_finalize_device(self._h, C.c_void_p(d_accum), width, height, spp_total,
                                                C.c_void_p(d_rgba), C.c_void_p(stream)))

    def render_device(self, cfg: PtbCfg, d_rgba: int, stream: int = 0):
        self._check(self._L.ptb_render_device(self._h, C.byref(cfg), C.c_void_p(d_rgba), C.c_void_p(stream)))

    def primary_hits(self, width, height, xi_u=0.5, xi_v=0.5):
        cfg = self.cfg(width, height, 1, 1)
        ids = np.empty((height, width), dtype=np.int32)
        t = np.empty((height, width), dtype=np.float64)
        self._check(self._L.ptb_primary_hits(self._h, C.byref(cfg), xi_u, xi_v, ids.ctypes.data, t.ctypes.data))
        return ids, t

    def stats(self) -> dict:
        s = PtbStats()
        self._check(self._L.ptb_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def last_kernel(self) -> str:
        """Symbol of the integrator instantiation the last render launched (as ncu prints it)."""
        return self._L.ptb_last_kernel(self._h).decode()

    def bvh_info(self) -> dict:
        """EXTENSION: the BVH ptb_scene_upload built over the mesh triangles (zeros without meshes)."""
        b = PtbBvhInfo()
        self._check(self._L.ptb_get_bvh_info(self._h, C.byref(b)))
        return b.as_dict()

    def fp32_peak_tflops(self) -> float:
        v = C.c_double()
        self._check(self._L.ptb_measure_fp32_peak(self._h, C.byref(v)))
        return v.value


class MultiContext:
    """ptb_multi: every listed CUDA device driven from ONE process (what a Go host owning the whole box would do).
    Device k traces its slice of the samples; device 0 sums all slices over NVLink peer loads inside the epilogue kernel."""

    def __init__(self, devices):
        self._L = _lib.lib()
        devs = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        rc = self._L.ptb_multi_create(arr, len(devs), C.byref(h))
        if rc:
            raise PtbError(rc, self._L.ptb_multi_last_error(None).decode())
        self._h = h
        self.devices = devs

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.ptb_multi_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise PtbError(rc, self._L.ptb_multi_last_error(self._h).decode())

    def upload(self, sc: Scene):
        flat = sc.flat()
        self._check(self._L.ptb_multi_scene_upload(self._h, C.byref(flat)))
        self._scene = sc

    def render(self, cfg: PtbCfg, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.zeros((cfg.height, cfg.width, 4), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.ndim == 3 and out.shape[2] == 4 and out.strides[1] == 4 and out.strides[2] == 1
        self._check(self._L.ptb_multi_render(self._h, C.byref(cfg), out.ctypes.data, out.strides[0]))
        return out

    def last_timing(self) -> dict:
        a, b = C.c_double(), C.c_double()
        self._check(self._L.ptb_multi_last_timing(self._h, C.byref(a), C.byref(b)))
        return dict(render_ms=a.value, reduce_ms=b.value)


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def RenderInto(sc: Scene, cfg: RenderConfig, img: np.ndarray, progress=None, *, ctx: Context | None = None, seed: int = 1):
    """engine.RenderInto (renderer.go:34-41) on the CUDA backend, via the C++ host mirror.

    img: uint8 (H, W, 4).  A size mismatch returns silently (renderer.go:46-49).  Errors raise PtbError
    (the Go shim returns `error`); nothing falls back to a CPU renderer.
    """
    if _backend != BackendCUDA:
        raise PtbError(_lib.PTB_ERR_INVALID, "only BackendCUDA exists in this build (no CPU fallback)")
    ctx = ctx or default_context()
    L = _lib.lib()
    cb = PROGRESS_FN(lambda _u: progress()) if progress is not None else None
    rc = L.ptb_engine_render_into(ctx._h, sc._h, cfg.Width, cfg.Height, cfg.SamplesPerPx, cfg.MaxDepth,
                                  int(seed) & 0xFFFFFFFF, img.ctypes.data, img.strides[0], img.shape[1], img.shape[0],
                                  C.cast(cb, C.c_void_p) if cb else None, None)
    if rc:
        raise PtbError(rc, L.ptb_host_last_error().decode())


def RenderCheckpointed(sc: Scene, cfg: RenderConfig, img: np.ndarray, path, samples_per_call: int, max_calls: int = 0, progress=None, *,
                       ctx: Context | None = None, seed: int = 1) -> int:
    """RenderInto with an accumulation checkpoint at `path` (ptb_engine_render_checkpointed): continues a matching checkpoint,
    rewrites it after every call of samples_per_call samples; max_calls > 0 stops early.  Returns the samples per pixel done."""
    ctx = ctx or default_context()
    L = _lib.lib()
    cb = PROGRESS_FN(lambda _u: progress()) if progress is not None else None
    done = C.c_int32()
    rc = L.ptb_engine_render_checkpointed(ctx._h, sc._h, cfg.Width, cfg.Height, cfg.SamplesPerPx, cfg.MaxDepth, int(seed) & 0xFFFFFFFF,
                                          img.ctypes.data, img.strides[0], img.shape[1], img.shape[0], str(path).encode(),
                                          int(samples_per_call), int(max_calls), C.byref(done), C.cast(cb, C.c_void_p) if cb else None, None)
    if rc:
        raise PtbError(rc, L.ptb_host_last_error().decode())
    return done.value


def Render(sc: Scene, cfg: RenderConfig, **kw) -> np.ndarray:
    """engine.Render (renderer.go:25-29)."""
    img = np.zeros((cfg.Height, cfg.Width, 4), dtype=np.uint8)
    RenderInto(sc, cfg, img, None, **kw)
    return img


def RenderScene(sc: Scene, settings: RenderSettings, **kw) -> np.ndarray:
    """engine.RenderScene (util.go:13-22)."""
    return Render(sc, RenderConfig(settings.Width, settings.Height, settings.SamplesPerPx, settings.MaxDepth), **kw)


def RenderSettingsForMode(mode: str) -> RenderSettings:
    """engine.RenderSettingsForMode (util.go:25-42)."""
    out = (C.c_int32 * 4)()
    _lib.lib().ptb_engine_settings_for_mode((mode or "").encode(), out)
    return RenderSettings(*out)


def SavePNG(path, img: np.ndarray) -> None:
    """engine.SavePNG (util.go:45-55)."""
    L = _lib.lib()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    rc = L.ptb_engine_save_png(str(path).encode(), img.ctypes.data, img.strides[0], img.shape[1], img.shape[0])
    if rc:
        raise PtbError(rc, L.ptb_host_last_error().decode())
