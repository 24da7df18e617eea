"""SPH kernel look-up table: the 64^2 / 32^2 / 16^2 / 8^2 mip chain topsy uploads as its kernel texture.

Host-side, start-up only (reference: SPH._get_kernel_at_resolution / _get_kernel_image_normalization /
_setup_kernel_texture, src/topsy/sph.py:372-426).  The reference evaluates ``pynbody.sph.kernels.Kernel2D.get_value``
5440 times; pynbody is not a dependency here, so the projected M4 cubic spline is integrated directly with
Gauss-Legendre quadrature, split at the spline's knot (r = 1) so every panel is smooth.
"""
from __future__ import annotations

import numpy as np

LEVEL_SIZES = (64, 32, 16, 8)
LEVEL_OFFSETS = (0, 4096, 5120, 5376)
LUT_TOTAL = 5440

_GL_X, _GL_W = np.polynomial.legendre.leggauss(48)


def cubic_spline_3d(r):
    """M4 spline with support 2 (h = 1), normalised in 3-D: W(r) = f(r)/pi."""
    r = np.asarray(r, dtype=np.float64)
    f = np.where(r < 1.0, 1.0 - 1.5 * r ** 2 + 0.75 * r ** 3, np.where(r < 2.0, 0.25 * (2.0 - r) ** 3, 0.0))
    return f / np.pi


def _panel(d, z0, z1):
    """int_{z0}^{z1} W(sqrt(z^2 + d^2)) dz for arrays d, z0, z1 (48-point Gauss-Legendre)."""
    half = 0.5 * (z1 - z0)
    mid = 0.5 * (z1 + z0)
    z = mid[..., None] + half[..., None] * _GL_X
    return half * np.sum(_GL_W * cubic_spline_3d(np.sqrt(z * z + d[..., None] ** 2)), axis=-1)


def projected_kernel(d):
    """Line-of-sight integral K2D(d) = 2 int_0^sqrt(4-d^2) W(sqrt(z^2+d^2)) dz; zero for d >= 2."""
    d = np.asarray(d, dtype=np.float64)
    zmax = np.sqrt(np.clip(4.0 - d * d, 0.0, None))
    zknot = np.sqrt(np.clip(1.0 - d * d, 0.0, None))          # 0 where d >= 1: first panel is empty
    return 2.0 * (_panel(d, np.zeros_like(d), zknot) + _panel(d, zknot, zmax))


def kernel_level(n_samples: int) -> np.ndarray:
    """Kernel sampled at the centres of an n x n grid spanning [-2, 2]^2, scaled so that the discrete integral
    sum(T) * (4/n)^2 is exactly 1 (sph.py:372-394)."""
    centres = np.linspace(-2 + 2.0 / n_samples, 2 - 2.0 / n_samples, n_samples)
    xx, yy = np.meshgrid(centres, centres)
    image = projected_kernel(np.sqrt(xx ** 2 + yy ** 2))
    image *= (n_samples / 4) ** 2 / image.sum()
    return image


def kernel_lut() -> np.ndarray:
    """All four levels, row-major, concatenated: float32[5440] in the layout tsplat_set_kernel_lut expects."""
    return np.concatenate([kernel_level(n).astype(np.float32).ravel() for n in LEVEL_SIZES])


def local_sphere_level(n_samples: int) -> np.ndarray:
    """Surface mode's kernel image (reference: LocalSphereKernel, sph.py:446-455, sampled by _get_kernel_at_resolution with
    normalisation 1.0, :497-501): depth of a sphere of radius 2h below its silhouette, -0.01 outside it (the fragment
    shader discards negative samples)."""
    centres = np.linspace(-2 + 2.0 / n_samples, 2 - 2.0 / n_samples, n_samples)
    xx, yy = np.meshgrid(centres, centres)
    d2 = xx ** 2 + yy ** 2
    return np.where(np.sqrt(d2) < 2.0, np.sqrt(np.clip(4.0 - d2, 0.0, None)), -0.01)


def local_sphere_lut() -> np.ndarray:
    """float32[5440] in the layout tsplat_set_surface expects."""
    return np.concatenate([local_sphere_level(n).astype(np.float32).ravel() for n in LEVEL_SIZES])
