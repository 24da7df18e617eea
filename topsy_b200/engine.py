"""Thin host object around one tsplat context: PyTorch owns every device buffer, the C ABI borrows raw pointers.

This is the layer the drop-in classes (``sph.SPH``, ``particle_buffers.ParticleBuffers``, ``colormap.*``) talk to; it has
no reference counterpart because the reference talks to wgpu directly.  It is deliberately small: allocate, hand
pointers over, launch on the current CUDA stream.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native as N
from .kernel_lut import kernel_lut


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class SplatEngine:
    """One per (GPU, render resolution)."""

    def __init__(self, resolution: int, device=None, max_particles_per_call: int = 2 ** 25):
        if not torch.cuda.is_available():
            raise RuntimeError("topsy_b200 needs a CUDA device: there is no CPU fallback for the SPH projection path")
        self.lib = N.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        self.resolution = int(resolution)
        handle = ctypes.c_void_p()
        N.check(self.lib.tsplat_create(self.device.index, self.resolution, ctypes.byref(handle)))
        self._ctx = handle
        self._images = {}
        self._particles = None
        self._weights = None
        self._scratch = None
        self._scratch_particles = 0
        self._max_particles_per_call = int(max_particles_per_call)
        self.set_kernel_lut(kernel_lut())
        self._bound_image = None

    # -- lifetime -------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.tsplat_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- constants ------------------------------------------------------------------------------------------
    def set_kernel_lut(self, lut: np.ndarray):
        lut = np.ascontiguousarray(lut, dtype=np.float32)
        N.check(self.lib.tsplat_set_kernel_lut(self._ctx, lut.ctypes.data_as(ctypes.c_void_p), lut.size))
        self.kernel_lut = lut

    def set_camera(self, matrix_row_major: np.ndarray, scale_factor: float):
        m = np.ascontiguousarray(matrix_row_major, dtype=np.float32).reshape(16)
        N.check(self.lib.tsplat_set_camera(self._ctx, m.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(scale_factor)))

    # -- particle data (borrowed) ---------------------------------------------------------------------------
    def set_particles(self, x, y, z, h):
        for t in (x, y, z, h):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.device != self.device:
                raise ValueError("particle arrays must be contiguous float32 CUDA tensors on the engine's device")
        n = x.numel()
        if not (y.numel() == z.numel() == h.numel() == n):
            raise ValueError("particle arrays differ in length")
        N.check(self.lib.tsplat_set_particles(self._ctx, _ptr(x), _ptr(y), _ptr(z), _ptr(h), n))
        self._particles = (x, y, z, h)      # keep alive
        self._weights = None
        self._ensure_scratch(min(n, self._max_particles_per_call))

    def set_weights(self, w0, w1=None, w2=None):
        n = self._particles[0].numel() if self._particles else 0
        for t in (w0, w1, w2):
            if t is None:
                continue
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.numel() != n:
                raise ValueError("weight arrays must be contiguous float32 CUDA tensors of the particle count")
        N.check(self.lib.tsplat_set_weights(self._ctx, _ptr(w0), _ptr(w1), _ptr(w2)))
        self._weights = (w0, w1, w2)

    def _ensure_scratch(self, n_particles):
        n_particles = max(int(n_particles), 1 << 20)
        if self._scratch is None or n_particles > self._scratch_particles:
            nbytes = self.lib.tsplat_scratch_bytes(self.resolution, n_particles)
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._scratch_particles = n_particles
            N.check(self.lib.tsplat_set_scratch(self._ctx, _ptr(self._scratch), nbytes))

    # -- render target --------------------------------------------------------------------------------------
    def image(self, channels: int) -> torch.Tensor:
        """(R, R, C) float32 accumulation image, row 0 = top; allocated on first use and reused."""
        if channels not in self._images:
            self._images[channels] = torch.zeros((self.resolution, self.resolution, channels), dtype=torch.float32,
                                                 device=self.device)
        return self._images[channels]

    def bind_image(self, image: torch.Tensor):
        if image.dtype != torch.float32 or not image.is_contiguous() or image.shape[:2] != (self.resolution,) * 2:
            raise ValueError("image must be a contiguous (R, R, C) float32 tensor")
        N.check(self.lib.tsplat_set_image(self._ctx, _ptr(image), image.shape[2]))
        self._bound_image = image

    # -- the hot path ---------------------------------------------------------------------------------------
    def render(self, mode: int, starts=None, lens=None, clear: bool = True, image: torch.Tensor | None = None):
        """Splat the particle ranges (None = all) into the mode's image on the current CUDA stream."""
        channels = N.MODE_CHANNELS[mode]
        img = self.image(channels) if image is None else image
        if self._bound_image is not img:
            self.bind_image(img)
        if starts is None:
            sp = lp = None
            n = 0
        else:
            s = np.ascontiguousarray(starts, dtype=np.int64)
            l = np.ascontiguousarray(lens, dtype=np.int64)
            if s.shape != l.shape or s.ndim != 1:
                raise ValueError("starts and lens must be 1-D arrays of equal length")
            n = len(s)
            if n == 0:
                if clear:
                    img.zero_()
                return img
            sp = s.ctypes.data_as(ctypes.c_void_p)
            lp = l.ctypes.data_as(ctypes.c_void_p)
        N.check(self.lib.tsplat_render(self._ctx, sp, lp, n, mode, int(bool(clear)), _stream(self.device)))
        return img

    def colormap(self, image: torch.Tensor, params: N.ColormapParams, lut: torch.Tensor | None, out: torch.Tensor,
                 out_fmt: int):
        """Fused normalise + log/linear + LUT pass.  ``out`` is (H, W, 4) uint8 / float16 / float32."""
        res, channels = image.shape[0], image.shape[2]
        if lut is None:
            lw = lh = 0
        elif lut.dim() == 2:
            lw, lh = lut.shape[0], 1
        else:
            lh, lw = lut.shape[0], lut.shape[1]
        N.check(self.lib.tsplat_colormap(self._ctx, _ptr(image), res, channels, ctypes.byref(params), _ptr(lut), lw, lh,
                                         _ptr(out), out.shape[1], out.shape[0], out_fmt, _stream(self.device)))
        return out

    # -- surface render mode --------------------------------------------------------------------------------------
    def set_surface(self, lut: np.ndarray | None, density_cut: float):
        """Local-sphere kernel LUT (None keeps the one already uploaded) and density cut of MODE_SURFACE renders."""
        if lut is None:
            N.check(self.lib.tsplat_set_surface(self._ctx, None, 0, ctypes.c_float(density_cut)))
        else:
            lut = np.ascontiguousarray(lut, dtype=np.float32)
            N.check(self.lib.tsplat_set_surface(self._ctx, lut.ctypes.data_as(ctypes.c_void_p), lut.size, ctypes.c_float(density_cut)))

    def bilateral_filter(self, image: torch.Tensor, out: torch.Tensor, spatial_sigma: float, range_sigma: float, kernel_size: int):
        """Bilateral filter of channel 1 of an (H, W, 2) float32 image into ``out`` (channel 0 is copied)."""
        if image.dim() != 3 or image.shape[2] != 2 or image.dtype != torch.float32 or not image.is_contiguous():
            raise ValueError("Input array must be 3D with shape (height, width, 2), contiguous float32")
        if out.shape != image.shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("output must match the input")
        N.check(self.lib.tsplat_bilateral_filter(self._ctx, _ptr(image), _ptr(out), image.shape[1], image.shape[0],
                                                 ctypes.c_float(spatial_sigma), ctypes.c_float(range_sigma), int(kernel_size),
                                                 _stream(self.device)))
        return out

    def surface_shade(self, smoothed: torch.Tensor, params: N.SurfaceParams, lut: torch.Tensor | None, out: torch.Tensor,
                      out_fmt: int):
        N.check(self.lib.tsplat_surface_shade(self._ctx, _ptr(smoothed), smoothed.shape[0], ctypes.byref(params), _ptr(lut),
                                              0 if lut is None else lut.shape[0], _ptr(out), out.shape[1], out.shape[0],
                                              out_fmt, _stream(self.device)))
        return out

    # -- device-side autorange ------------------------------------------------------------------------------------
    def content_stats(self, image: torch.Tensor, content: int, scale: float) -> N.ContentStats:
        st = N.ContentStats()
        N.check(self.lib.tsplat_content_stats(self._ctx, _ptr(image), image.shape[0], image.shape[2], content,
                                              ctypes.c_float(scale), ctypes.byref(st), _stream(self.device)))
        return st

    def content_percentiles(self, image: torch.Tensor, content: int, scale: float, use_log: bool, n_finite: int,
                            percentiles) -> list:
        """np.percentile(values[finite], percentiles) (linear interpolation) without leaving the device: exact order
        statistics from the radix-select kernel, interpolated here."""
        ranks, fracs = [], []
        for p in percentiles:
            pos = p / 100.0 * (n_finite - 1)
            lo = int(np.floor(pos))
            ranks += [lo, min(lo + 1, n_finite - 1)]
            fracs.append(pos - lo)
        out = []
        for k in range(0, len(ranks), 4):                       # the kernel resolves up to 4 ranks per sweep
            chunk = np.ascontiguousarray(ranks[k:k + 4], dtype=np.int64)
            vals = np.zeros(len(chunk), dtype=np.float32)
            N.check(self.lib.tsplat_content_select(self._ctx, _ptr(image), image.shape[0], image.shape[2], content,
                                                   ctypes.c_float(scale), int(bool(use_log)),
                                                   chunk.ctypes.data_as(ctypes.c_void_p), len(chunk),
                                                   vals.ctypes.data_as(ctypes.c_void_p), _stream(self.device)))
            out += list(vals)
        return [float(out[2 * i]) + (float(out[2 * i + 1]) - float(out[2 * i])) * fracs[i] for i in range(len(percentiles))]

    def axpy(self, dst: torch.Tensor, src: torch.Tensor, scale: float):
        N.check(self.lib.tsplat_image_axpy(self._ctx, _ptr(dst), _ptr(src), ctypes.c_float(scale), dst.numel(),
                                           _stream(self.device)))

    def enable_kernel_timing(self, enable: bool = True):
        """Bracket every K1 launch with CUDA events (for the live roofline of bench.py)."""
        N.check(self.lib.tsplat_enable_kernel_timing(self._ctx, int(bool(enable))))

    def kernel_timing(self):
        """(number of K1 launches, their summed duration in ms) since the last call; synchronises on those launches."""
        n, ms = ctypes.c_int64(0), ctypes.c_double(0.0)
        N.check(self.lib.tsplat_kernel_timing(self._ctx, ctypes.byref(n), ctypes.byref(ms)))
        return n.value, ms.value

    def stats(self) -> dict:
        st = N.Stats()
        N.check(self.lib.tsplat_get_stats(self._ctx, ctypes.byref(st)))
        return st.as_dict()

    # -- host staging (end-to-end path) ---------------------------------------------------------------------
    def upload(self, dst: torch.Tensor, src_host: np.ndarray | torch.Tensor):
        """cudaMemcpyAsync host -> device on the current stream (src should be pinned for true async)."""
        if isinstance(src_host, torch.Tensor):
            nbytes = src_host.numel() * src_host.element_size()
            sp = ctypes.c_void_p(src_host.data_ptr())
        else:
            nbytes = src_host.nbytes
            sp = src_host.ctypes.data_as(ctypes.c_void_p)
        if nbytes != dst.numel() * dst.element_size():
            raise ValueError("size mismatch in upload")
        N.check(self.lib.tsplat_memcpy_h2d(_ptr(dst), sp, nbytes, _stream(self.device)))

    def download(self, dst_host: torch.Tensor, src: torch.Tensor):
        nbytes = src.numel() * src.element_size()
        N.check(self.lib.tsplat_memcpy_d2h(ctypes.c_void_p(dst_host.data_ptr()), _ptr(src), nbytes, _stream(self.device)))

    def synchronize(self):
        N.check(self.lib.tsplat_stream_sync(_stream(self.device)))
