"""ctypes binding of libtsplat.so (include/tsplat.h).  No fallback: if the CUDA library is missing or fails to load,
importing a renderer raises -- the product path never routes around the hand-written kernels."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libtsplat.so"

MODE_DENSITY, MODE_WEIGHTED, MODE_RGB, MODE_DEPTH, MODE_SURFACE = 0, 1, 2, 3, 4
MODE_CHANNELS = {MODE_DENSITY: 1, MODE_WEIGHTED: 2, MODE_RGB: 4, MODE_DEPTH: 2, MODE_SURFACE: 2}
FMT_RGBA8, FMT_RGBA16F, FMT_RGBA32F = 0, 1, 2
CMAP_DENSITY, CMAP_WEIGHTED, CMAP_BIVARIATE, CMAP_BIVARIATE_WEIGHTED, CMAP_RGB = 0, 1, 2, 3, 4
LUT_TOTAL = 5440

ERR_INVALID, ERR_STATE, ERR_CUDA, ERR_NOMEM = -1, -2, -3, -4


class ColormapParams(ctypes.Structure):
    _fields_ = [("vmin", ctypes.c_float), ("vmax", ctypes.c_float),
                ("density_vmin", ctypes.c_float), ("density_vmax", ctypes.c_float),
                ("window_aspect_ratio", ctypes.c_float), ("gamma", ctypes.c_float),
                ("kind", ctypes.c_int32), ("log_scale", ctypes.c_int32)]


class SurfaceParams(ctypes.Structure):
    _fields_ = [("depth_scale", ctypes.c_float), ("light_direction", ctypes.c_float * 3), ("light_color", ctypes.c_float * 3),
                ("ambient_color", ctypes.c_float * 3), ("vmin", ctypes.c_float), ("vmax", ctypes.c_float),
                ("window_aspect_ratio", ctypes.c_float), ("material_colormap", ctypes.c_int32), ("log_scale", ctypes.c_int32)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("particles_submitted", "particles_culled", "particles_direct",
                                              "particles_tiled", "particles_huge", "tile_pairs", "kernel_launches",
                                              "direct_vector_reds")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ContentStats(ctypes.Structure):
    _fields_ = [("lin_min", ctypes.c_float), ("lin_max", ctypes.c_float), ("log_min", ctypes.c_float),
                ("log_max", ctypes.c_float), ("n_finite_lin", ctypes.c_int64), ("n_finite_log", ctypes.c_int64),
                ("any_negative", ctypes.c_int32)]


CONTENT_CH0, CONTENT_RATIO, CONTENT_ALL, CONTENT_CH0_WHERE_CH1 = 0, 1, 2, 3


class TsplatError(RuntimeError):
    pass


_lib = None

_vp, _i64, _i32, _f = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float

REDUCE_SUM, REDUCE_ZMAX = 0, 1

_SIGNATURES = {
    "tsplat_last_error": (ctypes.c_char_p, []),
    "tsplat_abi_version": (_i32, []),
    "tsplat_mode_channels": (_i32, [_i32]),
    "tsplat_create": (_i32, [_i32, _i32, ctypes.POINTER(_vp)]),
    "tsplat_destroy": (_i32, [_vp]),
    "tsplat_set_kernel_lut": (_i32, [_vp, _vp, _i32]),
    "tsplat_set_camera": (_i32, [_vp, _vp, _f]),
    "tsplat_set_particles": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64]),
    "tsplat_set_weights": (_i32, [_vp, _vp, _vp, _vp]),
    "tsplat_set_image": (_i32, [_vp, _vp, _i32]),
    "tsplat_scratch_bytes": (_i64, [_i32, _i64]),
    "tsplat_set_scratch": (_i32, [_vp, _vp, _i64]),
    "tsplat_render": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "tsplat_colormap": (_i32, [_vp, _vp, _i32, _i32, ctypes.POINTER(ColormapParams), _vp, _i32, _i32, _vp, _i32, _i32,
                               _i32, _vp]),
    "tsplat_reduce_colormap": (_i32, [_vp, ctypes.POINTER(_vp), _i32, _i32, _i32, _i32, ctypes.POINTER(ColormapParams), _vp,
                                      _i32, _i32, _vp, _i32, _vp, _vp]),
    "tsplat_allreduce_image": (_i32, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i32, _i32, _i32, _i32,
                                      _i32, _vp]),
    "tsplat_enable_peer_access": (_i32, [_i32, _i32]),
    "tsplat_enable_kernel_timing": (_i32, [_vp, _i32]),
    "tsplat_kernel_timing": (_i32, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(ctypes.c_double)]),
    "tsplat_set_surface": (_i32, [_vp, _vp, _i32, _f]),
    "tsplat_bilateral_filter": (_i32, [_vp, _vp, _vp, _i32, _i32, _f, _f, _i32, _vp]),
    "tsplat_surface_shade": (_i32, [_vp, _vp, _i32, ctypes.POINTER(SurfaceParams), _vp, _i32, _vp, _i32, _i32, _i32, _vp]),
    "tsplat_periodic_accumulate": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp]),
    "tsplat_content_stats": (_i32, [_vp, _vp, _i32, _i32, _i32, _f, ctypes.POINTER(ContentStats), _vp]),
    "tsplat_content_select": (_i32, [_vp, _vp, _i32, _i32, _i32, _f, _i32, _vp, _i32, _vp, _vp]),
    "tsplat_image_axpy": (_i32, [_vp, _vp, _vp, _f, _i64, _vp]),
    "tsplat_cell_layout_work_bytes": (_i64, [_i64, _i32]),
    "tsplat_cell_layout": (_i32, [_i32, _vp, _i64, _i32, ctypes.c_double, ctypes.c_double, _i32, _vp, _vp, _vp, _vp,
                                  _i64, _vp]),
    "tsplat_cell_layout_shuffled": (_i32, [_i32, _vp, _i64, _i32, ctypes.c_double, ctypes.c_double, _i32, ctypes.c_uint32, _vp, _vp,
                                           _vp, _vp, _i64, _vp]),
    "tsplat_gather_f32": (_i32, [_i32, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp]),
    "tsplat_memcpy_h2d": (_i32, [_vp, _vp, _i64, _vp]),
    "tsplat_memcpy_d2h": (_i32, [_vp, _vp, _i64, _vp]),
    "tsplat_stream_sync": (_i32, [_vp]),
    "tsplat_get_stats": (_i32, [_vp, ctypes.POINTER(Stats)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load libtsplat.so (built in-tree by ``__graft_entry__.build()`` / ``make -C topsy_b200/csrc``)."""
    global _lib
    if _lib is None:
        path = os.environ.get("TSPLAT_LIBRARY", str(LIB_PATH))
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                              f"g.build()' or make -C topsy_b200/csrc). There is no CPU fallback.")
        L = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.tsplat_abi_version() != 1:
            raise ImportError("libtsplat.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int):
    """Map tsplat_status to the exception types the reference raises at the same places."""
    if rc == 0:
        return
    msg = lib().tsplat_last_error().decode(errors="replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ERR_STATE:
        raise RuntimeError(msg)
    raise TsplatError(f"tsplat error {rc}: {msg}")
