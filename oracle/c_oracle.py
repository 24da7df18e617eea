"""ctypes front-end of oracle/splat_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).parent
_SO = _HERE / "_build" / "libsplat_oracle.so"
_lib = None

MODE_DENSITY, MODE_WEIGHTED, MODE_RGB, MODE_DEPTH = 0, 1, 2, 3
MODE_CHANNELS = {0: 1, 1: 2, 2: 4, 3: 2}


def build(force: bool = False) -> Path:
    src = _HERE / "splat_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "_build/libsplat_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        fp = ctypes.POINTER(ctypes.c_float); ip = ctypes.POINTER(ctypes.c_int64)
        for name, acc in (("oracle_splat_f64", ctypes.c_double), ("oracle_splat_f32", ctypes.c_float)):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [fp, fp, fp, fp, fp, fp, fp, ctypes.c_int64, ip, ip, ctypes.c_int, fp, ctypes.c_float, fp,
                           ctypes.c_int, ctypes.c_int, ctypes.POINTER(acc), ctypes.c_int, ctypes.c_int]
        _lib.oracle_count_updates.restype = ctypes.c_int
        _lib.oracle_count_updates.argtypes = [fp, fp, fp, fp, ctypes.c_int64, fp, ctypes.c_float, ctypes.c_int, ip, ip]
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_splat_surface.restype = ctypes.c_int
        _lib.oracle_splat_surface.argtypes = [fp, fp, fp, fp, fp, fp, ctypes.c_int64, ip, ip, ctypes.c_int, fp, ctypes.c_float,
                                              fp, ctypes.c_int, ctypes.c_float, fp, fp, ctypes.c_int, ctypes.c_int]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ty=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def splat(x, y, z, h, weights, M, sf, R, mode, lut, ranges=None, out=None, clear=True, accum=np.float64,
          nthreads=0):
    """Same contract as topsy_oracle.splat; returns (R,R,C) image of dtype ``accum``."""
    L = lib()
    x, y, z, h = _f32(x), _f32(y), _f32(z), _f32(h)
    w = [_f32(a) for a in weights] + [None, None, None]
    C = MODE_CHANNELS[mode]
    img = np.zeros((R, R, C), accum) if out is None else out
    assert img.dtype == accum and img.flags.c_contiguous
    M = _f32(np.asarray(M).reshape(16)); lut = _f32(lut)
    if ranges is None:
        st = ln = None; nr = 0
    else:
        st = np.ascontiguousarray(ranges[0], np.int64); ln = np.ascontiguousarray(ranges[1], np.int64); nr = len(st)
    fn = L.oracle_splat_f64 if accum == np.float64 else L.oracle_splat_f32
    acc_t = ctypes.c_double if accum == np.float64 else ctypes.c_float
    rc = fn(_ptr(x), _ptr(y), _ptr(z), _ptr(h), _ptr(w[0]), _ptr(w[1]), _ptr(w[2]), len(x),
            _ptr(st, ctypes.c_int64), _ptr(ln, ctypes.c_int64), nr, _ptr(M), ctypes.c_float(float(sf)), _ptr(lut),
            int(R), int(mode), _ptr(img, acc_t), int(bool(clear)), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle splat failed rc={rc}")
    return img


def splat_surface(x, y, z, h, m, q, M, sf, R, lut, density_cut, ranges=None, out=None, zbuf=None, clear=True,
                  clamp_depth=False):
    """Z-buffered surface splat (see oracle_splat_surface): (R, R, 2) float32 = (quantity, depth)."""
    L = lib()
    x, y, z, h, m, q = (_f32(a) for a in (x, y, z, h, m, q))
    img = np.zeros((R, R, 2), np.float32) if out is None else out
    if clamp_depth and zbuf is None:
        zbuf = np.zeros((R, R), np.float32)
    M = _f32(np.asarray(M).reshape(16)); lut = _f32(lut)
    if ranges is None:
        st = ln = None; nr = 0
    else:
        st = np.ascontiguousarray(ranges[0], np.int64); ln = np.ascontiguousarray(ranges[1], np.int64); nr = len(st)
    rc = L.oracle_splat_surface(_ptr(x), _ptr(y), _ptr(z), _ptr(h), _ptr(m), _ptr(q), len(x), _ptr(st, ctypes.c_int64),
                                _ptr(ln, ctypes.c_int64), nr, _ptr(M), ctypes.c_float(float(sf)), _ptr(lut), int(R),
                                ctypes.c_float(float(density_cut)), _ptr(img), _ptr(zbuf), int(bool(clear)),
                                int(bool(clamp_depth)))
    if rc != 0:
        raise RuntimeError(f"oracle surface splat failed rc={rc}")
    return img


def count_updates(x, y, z, h, M, sf, R):
    L = lib()
    x, y, z, h = _f32(x), _f32(y), _f32(z), _f32(h)
    M = _f32(np.asarray(M).reshape(16))
    upd = ctypes.c_int64(0); cul = ctypes.c_int64(0)
    L.oracle_count_updates(_ptr(x), _ptr(y), _ptr(z), _ptr(h), len(x), _ptr(M), ctypes.c_float(float(sf)), int(R),
                           ctypes.byref(upd), ctypes.byref(cul))
    return upd.value, cul.value


def num_threads():
    return lib().oracle_num_threads()
