"""Data sources (reference: src/topsy/loader.py).

``AbstractDataLoader`` is the contract the renderer relies on; ``TestDataLoader`` is the reference's seeded synthetic
snapshot (the fixture behind every golden vector); ``ArrayDataLoader`` wraps plain numpy arrays and applies the cell
layout + within-cell shuffle exactly like the reference's pynbody loaders; the pynbody-backed loaders themselves exist
only when pynbody is importable (it is an optional dependency here and absent from the build image)."""
from __future__ import annotations

import logging
from abc import ABC, abstractmethod

import numpy as np

from . import cell_layout, config

logger = logging.getLogger(__name__)


class AbstractDataLoader(ABC):
    def __init__(self, device):
        self._device = device

    @abstractmethod
    def __len__(self): ...

    @abstractmethod
    def get_positions(self): ...

    @abstractmethod
    def get_smooth(self): ...

    @abstractmethod
    def get_mass(self): ...

    @abstractmethod
    def get_named_quantity(self, name): ...

    @abstractmethod
    def get_quantity_label(self, quantity_name): ...

    @abstractmethod
    def get_rgb_masses(self): ...

    @abstractmethod
    def get_position_units(self) -> str: ...

    def get_pos_smooth(self):
        """(N,4) float32: x, y, z, smoothing length (loader.py:52-56)."""
        out = np.empty((len(self), 4), dtype=np.float32)
        out[:, :3] = self.get_positions()
        out[:, 3] = self.get_smooth()
        return out

    def get_periodicity_scale(self):
        return np.inf

    def get_render_progression(self):
        from . import progressive_render
        if hasattr(self, '_cell_layout'):
            return progressive_render.RenderProgressionWithCells(self._cell_layout, len(self))
        return progressive_render.RenderProgression(len(self))

    def get_initial_center(self):
        return np.zeros(3, dtype=np.float32)

    def get_initial_view_width(self):
        period = self.get_periodicity_scale()
        return period / 2 if period is not None else config.DEFAULT_SCALE

    def _keep_particles(self, indices):
        """Drop every particle not in ``indices`` (ascending positions in the loader's current order): how a rank of a
        multi-GPU run keeps its stripe (topsy_b200.distributed.shard_loader).  No reference counterpart."""
        raise TypeError(f"{type(self).__name__} does not support multi-GPU sharding")


class TestDataLoader(AbstractDataLoader):
    """Three-component Gaussian mixture with h = 2 / (number density)^(1/3)  (loader.py:241-332).

    The legacy global numpy stream is seeded and consumed in the same order as the reference, so a given (n, seed)
    yields bit-identical particles -- the reference's golden images depend on it."""
    __test__ = False     # not a pytest class

    def __init__(self, device, n_particles: int = config.TEST_DATA_NUM_PARTICLES_DEFAULT, n_cells=10, seed: int = 1337,
                 with_cells=False, periodic=False):
        self._n_particles = n_particles
        self._gmm_weights = [0.5, 0.4, 0.1]
        self._gmm_means = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [6.0, 10.0, 0.0]])
        self._gmm_std = np.array([[20.0, 20.0, 20.0], [4.0, 0.2, 4.0], [2.0, 2.0, 3.0]])
        self._gmm_pos = self._draw_positions(seed)
        self._gmm_den = self._number_density(self._gmm_pos)
        self._periodic = periodic
        if with_cells:
            self._cell_layout, order = cell_layout.CellLayout.from_positions(
                self._gmm_pos, self._gmm_pos.min() - 1e-3, self._gmm_pos.max() + 1, n_cells)
            self._gmm_pos = self._gmm_pos[order]
            self._gmm_den = self._gmm_den[order]
        super().__init__(device)

    def __len__(self):
        return self._n_particles

    def _draw_positions(self, seed):
        np.random.seed(seed)
        pos = np.empty((self._n_particles, 3), dtype=np.float32)
        if self._n_particles == 1:
            pos[0] = self._gmm_means[0]
            return np.random.permutation(pos)
        filled = 0
        for weight, mean, std in zip(self._gmm_weights, self._gmm_means, self._gmm_std):
            count = int(self._n_particles * weight)
            unit = np.random.normal(size=(count, 3), scale=1.0).astype(np.float32)
            pos[filled:filled + count] = unit * std[np.newaxis, :] + mean
            filled += count
        assert filled == self._n_particles
        return np.random.permutation(pos)

    def _number_density(self, pos):
        den = np.zeros(len(pos))
        for weight, mean, std in zip(self._gmm_weights, self._gmm_means, self._gmm_std):
            den += weight * np.exp(-np.sum((pos - mean) ** 2 / std ** 2, axis=1)) / ((2 * np.pi) ** 1.5 * np.prod(std))
        return den * self._n_particles

    def _keep_particles(self, indices):
        self._gmm_pos = self._gmm_pos[indices]
        self._gmm_den = self._gmm_den[indices]
        self._n_particles = len(indices)

    def get_positions(self):
        return self._gmm_pos

    def get_smooth(self):
        return 2.0 / self._gmm_den ** 0.333333

    def get_mass(self):
        return np.repeat(np.float32(1e-8), self._n_particles)

    def get_named_quantity(self, name):
        if name != "test-quantity":
            raise KeyError("Unknown quantity name")
        p = self._gmm_pos
        return np.sin(p[:, 0]) * np.cos(p[:, 1]) * np.cos(p[:, 2]) * 1e-4

    def get_quantity_names(self):
        return ["test-quantity"]

    def get_quantity_label(self, quantity_name):
        if quantity_name is None:
            return r"test density / $M_{\odot} / \mathrm{kpc}^2$"
        return "test quantity" if quantity_name == "test-quantity" else "unknown"

    def get_position_units(self):
        return "kpc"

    def get_filename(self):
        return "test data"

    def get_periodicity_scale(self):
        return 100.0 if self._periodic else None

    def get_rgb_masses(self):
        p = self._gmm_pos
        rgb = np.empty((len(p), 3), dtype=np.float32)
        rgb[:, 0] = abs(np.sin(p[:, 0] / 10.0))
        rgb[:, 1] = abs(np.cos(p[:, 1] / 10.0))
        rgb[:, 2] = abs(np.cos(p[:, 2] / 10.0))
        return rgb


class ArrayDataLoader(AbstractDataLoader):
    """Snapshot given as arrays.  Applies what the reference's PynbodyDataInMemory does at load (loader.py:84-98): pad the
    bounding cube by CELL_LAYOUT_FRACTIONAL_PADDING, bucket into DEFAULT_CELLS_NSIDE^3 cells, shuffle inside cells."""

    def __init__(self, device, pos, smooth, mass, quantities=None, rgb=None, position_units="kpc", boxsize=None,
                 use_cells=True, nside=None, layout_on_device=None):
        super().__init__(device)
        pos = np.asarray(pos)
        n = len(pos)
        self._quantities = dict(quantities or {})
        self._position_units = position_units
        self._boxsize = boxsize
        lo = pos.min(); hi = pos.max()
        extent = hi - lo
        self._initial_view_width = extent
        self._dev = None                  # device-resident columns (x, y, z, h, m [, r, g, b]) when the layout ran on the GPU
        self._order_dev = None
        if layout_on_device is None:      # default: on the GPU whenever the loader was given one
            layout_on_device = use_cells and hasattr(device, "torch_device") and n > 0
        if use_cells and layout_on_device:
            self._init_on_device(pos, smooth, mass, rgb, lo, hi, extent, nside or config.DEFAULT_CELLS_NSIDE)
            return
        if use_cells:
            lo = lo - config.CELL_LAYOUT_FRACTIONAL_PADDING * extent
            hi = hi + config.CELL_LAYOUT_FRACTIONAL_PADDING * extent
            self._cell_layout, order = cell_layout.CellLayout.from_positions(pos, lo, hi, nside or config.DEFAULT_CELLS_NSIDE)
            self._host_order = order[self._cell_layout.randomize_within_cells()]
        else:
            self._host_order = np.arange(n)
        self._pos = pos.astype(np.float32)[self._host_order]
        self._smooth = np.asarray(smooth).astype(np.float32)[self._host_order]
        self._mass = np.asarray(mass).astype(np.float32)[self._host_order]
        self._rgb = None if rgb is None else np.asarray(rgb, dtype=np.float32)[self._host_order]

    def _init_on_device(self, pos, smooth, mass, rgb, lo, hi, extent, nside):
        """The reference's load-time pipeline (loader.py:84-98: pad the bounding cube, bucket into cells, shuffle inside
        cells, reorder every array) with all O(N) work on the GPU: one upload of the raw arrays, the K4 counting sort with
        the within-cell shuffle fused in, and one gather per column.  The columns stay on the device for ParticleBuffers;
        the host getters download them on demand."""
        import torch
        dev = self._device.torch_device
        lo = lo - config.CELL_LAYOUT_FRACTIONAL_PADDING * extent
        hi = hi + config.CELL_LAYOUT_FRACTIONAL_PADDING * extent
        if pos.dtype not in (np.float32, np.float64):
            pos = pos.astype(np.float64)
        pos_d = torch.from_numpy(np.ascontiguousarray(pos)).to(dev)
        seed = int(np.random.randint(1, 2 ** 31 - 1))          # the reference shuffles with the unseeded global numpy stream
        self._cell_layout, order = cell_layout.CellLayout.from_positions(pos_d, lo, hi, nside, shuffle_seed=seed)
        self._order_dev = order
        self._dev = {}
        for k, name in enumerate("xyz"):
            self._dev[name] = self._gather(pos_d, stride=3, offset=k)
        del pos_d
        self._dev["h"] = self._gather(self._to_device(smooth))
        self._dev["m"] = self._gather(self._to_device(mass))
        if rgb is not None:
            rgb_d = self._to_device(np.asarray(rgb))
            for k, name in enumerate("rgb"):
                self._dev["rgb_" + name] = self._gather(rgb_d, stride=3, offset=k).nan_to_num_(0.0)      # loader.py:119
        self._pos = self._smooth = self._mass = self._rgb = None      # host copies are made on demand
        self._has_rgb = rgb is not None

    def _to_device(self, array):
        import torch
        array = np.asarray(array)
        if array.dtype not in (np.float32, np.float64):
            array = array.astype(np.float32)
        return torch.from_numpy(np.ascontiguousarray(array)).to(self._device.torch_device)

    def _gather(self, src, stride=1, offset=0):
        """float32 copy of src[order * stride + offset] on the device (tsplat_gather_f32)."""
        import ctypes

        import torch

        from . import _native as N
        n = self._order_dev.numel()
        out = torch.empty(n, dtype=torch.float32, device=src.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
        N.check(N.lib().tsplat_gather_f32(src.device.index, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                          src.element_size(), stride, offset, ctypes.c_void_p(self._order_dev.data_ptr()), n,
                                          stream))
        return out

    @property
    def _particle_order(self):
        if self._order_dev is not None:
            if getattr(self, "_host_order", None) is None:
                self._host_order = self._order_dev.cpu().numpy()
            return self._host_order
        return self._host_order

    def device_columns(self, names):
        """Device-resident float32 columns ('x','y','z','h','m','rgb_r','rgb_g','rgb_b') for ParticleBuffers, or None when
        the loader keeps its data on the host."""
        if self._dev is None or any(k not in self._dev for k in names):
            return None
        return [self._dev[k] for k in names]

    def device_quantity(self, name):
        """A named quantity reordered on the device (None on the host path)."""
        if self._dev is None:
            return None
        q = np.asarray(self._quantities[name])
        if q.ndim == 2:
            q = q[:, 0]
        return self._gather(self._to_device(q))

    def __len__(self):
        return len(self._pos) if self._dev is None else int(self._dev["x"].numel())

    def _keep_particles(self, indices):
        if self._dev is not None:
            import torch
            idx = torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int64)).to(self._order_dev.device)
            self._dev = {k: v[idx].contiguous() for k, v in self._dev.items()}
            self._order_dev = self._order_dev[idx].contiguous()
            self._host_order = None
            return
        self._pos, self._smooth, self._mass = self._pos[indices], self._smooth[indices], self._mass[indices]
        if self._rgb is not None:
            self._rgb = self._rgb[indices]
        self._host_order = self._host_order[indices]

    def _host(self, *names):
        cols = [self._dev[k].cpu().numpy() for k in names]
        return cols[0] if len(cols) == 1 else np.stack(cols, axis=1)

    def get_positions(self):
        return self._pos if self._dev is None else self._host("x", "y", "z")

    def get_smooth(self):
        return self._smooth if self._dev is None else self._host("h")

    def get_mass(self):
        return self._mass if self._dev is None else self._host("m")

    def get_named_quantity(self, name):
        q = np.asarray(self._quantities[name])
        if q.ndim == 2:
            q = q[:, 0]
        return q.astype(np.float32)[self._particle_order]

    def get_quantity_names(self):
        return list(self._quantities)

    def get_quantity_label(self, quantity_name):
        return r"density / $M_{\odot} / \mathrm{kpc}^2$" if quantity_name is None else str(quantity_name)

    def get_rgb_masses(self):
        if self._dev is not None and self._has_rgb:
            rgb = self._host("rgb_r", "rgb_g", "rgb_b")
        elif self._dev is None and self._rgb is not None:
            rgb = self._rgb.copy()
        else:
            raise ValueError("this snapshot has no rgb (band luminosity) arrays")
        rgb[np.isnan(rgb)] = 0.0
        return rgb

    def get_position_units(self):
        return self._position_units

    def get_periodicity_scale(self):
        return self._boxsize

    def get_initial_view_width(self):
        return self._initial_view_width

    def get_filename(self):
        return "in-memory arrays"


def _require_pynbody():
    try:
        import pynbody
        return pynbody
    except ImportError as e:
        raise ImportError("pynbody is not installed: use TestDataLoader / ArrayDataLoader, or install pynbody to load "
                          "simulation files") from e


class PynbodyDataInMemory(ArrayDataLoader):
    """A pynbody SimSnap already in memory (loader.py:79-155).  Needs pynbody."""
    _name_smooth_array = 'smooth'

    def __init__(self, device, snapshot):
        _require_pynbody()
        self.snapshot = snapshot
        boxsize = float(snapshot.properties['boxsize'].in_units("kpc")) if 'boxsize' in snapshot.properties else None
        super().__init__(device, np.asarray(snapshot['pos']), np.asarray(snapshot[self._name_smooth_array]),
                         np.asarray(snapshot['mass']), quantities=_LazySnapshotArrays(snapshot), rgb=None,
                         position_units=str(snapshot['pos'].units), boxsize=boxsize)

    def _effective_mass_for_band(self, band):
        return (10 ** (-0.4 * np.asarray(self.snapshot[band + "_mag"])))[self._particle_order]

    def get_rgb_masses(self):
        # derived on demand, only when the rgb render modes ask for it (reference loader.py:112-121); a snapshot without
        # the *_mag arrays raises the snapshot's own KeyError, as the reference does
        rgb = np.empty((len(self), 3), dtype=np.float32)
        rgb[:, 0] = self._effective_mass_for_band('I') * 0.5
        rgb[:, 1] = self._effective_mass_for_band('V')
        rgb[:, 2] = self._effective_mass_for_band('U')
        rgb[np.isnan(rgb)] = 0.0
        return rgb

    def get_quantity_label(self, quantity_name):
        if quantity_name is None:
            return r"density / $M_{\odot} / \mathrm{kpc}^2$"
        lunit = self.snapshot[quantity_name].units.latex()
        if lunit != "":
            lunit = "$/" + lunit + "$"
        return quantity_name + lunit

    def get_quantity_names(self):
        return self.snapshot.loadable_keys()

    def get_filename(self):
        return self.snapshot.filename


class _LazySnapshotArrays(dict):
    def __init__(self, snapshot):
        super().__init__()
        self._snapshot = snapshot

    def __getitem__(self, name):
        return np.asarray(self._snapshot[name])


class PynbodyDataLoader(PynbodyDataInMemory):
    """Load a simulation file with pynbody, centre it and (re)use cached smoothing lengths (loader.py:157-238)."""
    _name_smooth_array = 'topsy_smooth'

    def __init__(self, device, filename: str, center: str, particle: str, take_region=None):
        import pickle
        pynbody = _require_pynbody()
        snapshot = pynbody.load(filename) if take_region is None else pynbody.load(filename, take_region=take_region)
        snapshot.physical_units()
        self.filename = filename
        family = pynbody.family.get_family(particle)
        snapshot = snapshot[family]
        if np.ptp(snapshot['pos']) < 1.0:
            snapshot.physical_units('au')
        if center.startswith("halo-"):
            cen = pynbody.analysis.halo.center(snapshot.ancestor.halos()[int(center[5:])], return_cen=True)
        elif center == 'zoom':
            dm = snapshot.ancestor.dm
            cen = pynbody.analysis.halo.center(dm[dm['mass'] < 1.01 * dm['mass'].min()], return_cen=True)
        elif center == 'all':
            cen = pynbody.analysis.halo.center(snapshot, return_cen=True)
        elif center == 'none':
            cen = np.zeros(3)
        else:
            raise ValueError("Unknown centering type")
        self._initial_center = cen
        cache = f"{filename}-topsy-smooth-{family.name}.pkl"
        try:
            smooth = pickle.load(open(cache, 'rb'))
            if len(smooth) != len(snapshot):
                raise ValueError("stale smoothing cache")
            snapshot[self._name_smooth_array] = smooth
        except Exception:
            snapshot[self._name_smooth_array] = pynbody.sph.smooth(snapshot)
            try:
                pickle.dump(snapshot[self._name_smooth_array], open(cache, 'wb'))
            except IOError:
                logger.warning("Unable to save smoothing data to disk")
        super().__init__(device, snapshot)

    def get_initial_center(self):
        return self._initial_center
