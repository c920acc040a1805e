"""ctypes binding of libptb200.so (include/ptb200.h + include/ptb200_host.h).

The product path is the CUDA library and nothing else: if the shared object is missing or does not
load, importing this module raises — there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import pathlib

import os

_PKG = pathlib.Path(__file__).resolve().parent
# PTB200_LIB: load another build of the SAME library (tuning variants made by tools/build_variants.sh)
SO_PATH = pathlib.Path(os.environ["PTB200_LIB"]) if os.environ.get("PTB200_LIB") else _PKG / "libptb200.so"

PTB_OK, PTB_ERR_INVALID, PTB_ERR_CUDA, PTB_ERR_NO_SCENE, PTB_ERR_LIMIT = 0, -1, -2, -3, -4
PTB_OBJ_SPHERE, PTB_OBJ_PLANE, PTB_OBJ_BOX, PTB_OBJ_MESH = 0, 1, 2, 3
PTB_MAT_LAMBERT, PTB_MAT_METAL, PTB_MAT_DIELECTRIC, PTB_MAT_EMISSIVE, PTB_MAT_MIRROR = 0, 1, 2, 3, 4
PTB_SKY_CONST, PTB_SKY_GRADIENT = 0, 1
PTB_FLAG_STATS = 1
PTB_FLAG_MEGAKERNEL = 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class PtbCamera(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("target", C.c_double * 3), ("up", C.c_double * 3),
                ("fov", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
                ("aspect_ratio", C.c_double)]


class PtbSky(C.Structure):
    _fields_ = [("kind", C.c_int32), ("color", C.c_double * 3), ("horizon", C.c_double * 3),
                ("zenith", C.c_double * 3)]


class PtbScene(C.Structure):
    _fields_ = [("n_obj", C.c_int32), ("obj_type", _ip), ("obj_mat", _ip), ("obj_pos", _dp), ("obj_size", _dp),
                ("n_mat", C.c_int32), ("mat_type", _ip), ("mat_albedo", _dp), ("mat_rough", _dp), ("mat_ior", _dp),
                ("mat_emit", _dp), ("mat_power", _dp), ("mat_absorption", _dp), ("mat_smoothness", _dp),
                ("camera", PtbCamera), ("sky", PtbSky),
                ("n_mesh", C.c_int32), ("obj_mesh", _ip), ("mesh_tri_begin", C.POINTER(C.c_int64)),
                ("tri_vertices", C.POINTER(C.c_float)), ("mesh_generation", C.c_uint64)]


class PtbCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples_per_px", C.c_int32),
                ("max_depth", C.c_int32), ("seed", C.c_uint32), ("sample_begin", C.c_int32),
                ("sample_count", C.c_int32), ("flags", C.c_uint32), ("row_offset", C.c_int32), ("row_step", C.c_int32)]


class PtbStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("segments", C.c_uint64), ("exit_scans", C.c_uint64),
                ("accepts", C.c_uint64 * 3), ("scatters", C.c_uint64), ("end_sky", C.c_uint64),
                ("end_emissive", C.c_uint64), ("end_rr", C.c_uint64), ("end_depth", C.c_uint64),
                ("end_noscatter", C.c_uint64), ("lane_iters_active", C.c_uint64),
                ("lane_iters_total", C.c_uint64), ("last_render_ms", C.c_double),
                ("accepts_mesh", C.c_uint64), ("bvh_nodes_visited", C.c_uint64), ("bvh_tris_tested", C.c_uint64),
                ("bvh_stack_overflows", C.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "accepts"}
        d["accepts"] = list(self.accepts)
        return d


class PtbBvhInfo(C.Structure):
    _fields_ = [("n_triangles", C.c_int64), ("n_nodes", C.c_int64), ("max_depth", C.c_int32), ("node_bytes", C.c_int32),
                ("triangle_bytes", C.c_int32), ("reserved", C.c_int32), ("sah_cost", C.c_double), ("build_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class PtbDeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("sm_count", C.c_int32), ("cc_major", C.c_int32),
                ("cc_minor", C.c_int32), ("clock_khz", C.c_int32), ("global_mem_bytes", C.c_uint64)]


class PtbBvh(C.Structure):
    _fields_ = [("nodes", C.POINTER(C.c_float)), ("triangles", C.POINTER(C.c_float)), ("info", PtbBvhInfo)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p)

# name -> (restype, argtypes): every symbol include/ptb200.h and include/ptb200_host.h declare
SYMBOLS = {
    "ptb_abi_version": (C.c_int, []),
    "ptb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ptb_destroy": (None, [C.c_void_p]),
    "ptb_last_error": (C.c_char_p, [C.c_void_p]),
    "ptb_get_device_info": (C.c_int, [C.c_void_p, C.POINTER(PtbDeviceInfo)]),
    "ptb_scene_upload": (C.c_int, [C.c_void_p, C.POINTER(PtbScene)]),
    "ptb_world_size": (C.c_int, [C.c_void_p]),
    "ptb_world_get": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "ptb_rows_of": (C.c_int, [C.POINTER(PtbCfg)]),
    "ptb_render": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ptb_render_accum": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_void_p]),
    "ptb_render_accum_device": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_void_p, C.c_void_p]),
    "ptb_finalize_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ptb_render_device": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_void_p, C.c_void_p]),
    "ptb_render_resume": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_void_p, C.c_void_p, C.c_size_t]),
    "ptb_finalize_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]),
    "ptb_host_buffer_pin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ptb_host_buffer_unpin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ptb_primary_hits": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "ptb_get_stats": (C.c_int, [C.c_void_p, C.POINTER(PtbStats)]),
    "ptb_last_kernel": (C.c_char_p, [C.c_void_p]),
    "ptb_get_bvh_info": (C.c_int, [C.c_void_p, C.POINTER(PtbBvhInfo)]),
    "ptb_measure_fp32_peak": (C.c_int, [C.c_void_p, _dp]),
    "ptb_scene_device_order": (C.c_int, [C.POINTER(PtbScene), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "ptb_launch_plan": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "ptb_bvh_selfcheck": (C.c_int64, [C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "ptb_bvh_build": (C.c_int, [C.POINTER(C.c_float), C.c_int64, C.POINTER(PtbBvh)]),
    "ptb_bvh_free": (None, [C.POINTER(PtbBvh)]),
    "ptb_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "ptb_multi_destroy": (None, [C.c_void_p]),
    "ptb_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "ptb_multi_scene_upload": (C.c_int, [C.c_void_p, C.POINTER(PtbScene)]),
    "ptb_multi_render": (C.c_int, [C.c_void_p, C.POINTER(PtbCfg), C.c_void_p, C.c_size_t]),
    "ptb_multi_last_timing": (C.c_int, [C.c_void_p, _dp, _dp]),
    "ptb_peer_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ptb_peer_destroy": (None, [C.c_void_p]),
    "ptb_peer_handles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptb_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptb_peer_status": (C.c_int, [C.c_void_p]),
    "ptb_peer_accum": (C.c_void_p, [C.c_void_p]),
    "ptb_peer_image": (C.c_void_p, [C.c_void_p]),
    "ptb_peer_slice": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ptb_peer_reduce_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    # host mirror
    "ptb_host_last_error": (C.c_char_p, []),
    "ptb_host_scene_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ptb_host_scene_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ptb_host_scene_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "ptb_host_scene_marshal": (C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "ptb_host_scene_free": (None, [C.c_void_p]),
    "ptb_host_scene_flat": (C.c_int, [C.c_void_p, C.POINTER(PtbScene)]),
    "ptb_host_scene_settings": (C.c_int, [C.c_void_p, _ip]),
    "ptb_host_scene_counts": (C.c_int, [C.c_void_p, _ip, _ip]),
    "ptb_engine_settings_for_mode": (None, [C.c_char_p, _ip]),
    "ptb_engine_render_into": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_uint32, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p]),
    "ptb_engine_render_checkpointed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32,
                                                 C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_int32,
                                                 C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    "ptb_engine_save_png": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32]),
}

_LIB = None


def lib():
    """Load libptb200.so (once). Raises if it is missing: the CUDA library IS the product."""
    global _LIB
    if _LIB is None:
        if not SO_PATH.exists():
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C path_trace_golang_b200/csrc`. There is no CPU fallback.")
        L = C.CDLL(str(SO_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class PtbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ptb200 error {code}: {message}")
        self.code = code
        self.message = message
