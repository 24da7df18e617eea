"""Colormap holder (API of the reference's src/topsy/colormap/__init__.py:12-159).

The presentation stage can change kind while a visualizer lives (density -> weighted mean -> bivariate -> rgb ...).  The
holder owns the parameter dictionary contract and swaps the implementation object whenever the merged parameters are no
longer accepted by the current one; callers only ever talk to the holder.  Implementations register themselves by
subclassing ``ColormapBase``; the first registered class whose ``accepts_parameters`` says yes wins, searching
depth-first through the class tree exactly like the reference, so class precedence is identical.
"""
from __future__ import annotations

import numpy as np

from .. import config
from . import implementation, luts  # noqa: F401
from .implementation import (BivariateColormap, Colormap, ColormapBase, NoColormap, RGBColormap,  # noqa: F401
                             RGBHDRColormap)
from .surface import ColorAsSurfaceMap  # noqa: F401  (registers the 'surface' implementation)

_UNINITIALISED = {'colormap_name': config.DEFAULT_COLORMAP, 'vmin': None, 'vmax': None, 'log': False, 'type': 'none'}


def _walk(cls):
    """Depth-first pre-order over the subclasses of ``cls``."""
    for child in cls.__subclasses__():
        yield child
        yield from _walk(child)


class ColormapHolder:
    def __init__(self, device, input_texture, output_format):
        self._binding = (device, input_texture, output_format)
        self._device, self._input_texture, self._output_format = self._binding
        self._impl = self.instance_from_parameters(dict(_UNINITIALISED), *self._binding)

    # -- implementation lookup ----------------------------------------------------------------------------------
    @classmethod
    def _iter_classes(cls, base_class=ColormapBase):
        return _walk(base_class)

    @classmethod
    def _class_from_parameters(cls, parameters):
        return next((k for k in _walk(ColormapBase) if k.accepts_parameters(parameters)), None)

    @classmethod
    def instance_from_parameters(cls, parameters, device, input_texture, output_format):
        chosen = cls._class_from_parameters(parameters)
        if chosen is None:
            raise ValueError(f"No colormap class found for parameters: {parameters}")
        return chosen(device, input_texture, output_format, parameters)

    def _require_real_colormap(self):
        if self._impl is None or isinstance(self._impl, NoColormap):
            raise ValueError("ColormapHolder is not fully initialized")

    _check_valid = _require_real_colormap

    # -- parameters ---------------------------------------------------------------------------------------------
    def update_parameters(self, parameters: dict):
        """Merge ``parameters`` into the current set.  Returns True if that needed a different implementation object,
        False if the current one absorbed the change."""
        wanted = {**self.get_parameters(), **parameters}
        keep_current = self._impl is not None and self._impl.accepts_parameters(wanted)
        if keep_current:
            self._impl.update_parameters(parameters)
            return False
        if self._impl is None and self._class_from_parameters(wanted) is None:
            return None
        self._impl = self.instance_from_parameters(wanted, *self._binding)
        return True

    def get_parameters(self) -> dict:
        return self._impl.get_parameters()

    def get_parameter(self, name: str):
        return self._impl.get_parameter(name)

    def __getitem__(self, key: str):
        return self._impl.get_parameter(key)

    def __setitem__(self, key: str, value):
        self.update_parameters({key: value})

    # -- forwarding to the live implementation ----------------------------------------------------------------------
    def _forward(name):  # noqa: N805 -- tiny descriptor factory, evaluated at class creation
        def call(self, *args, **kwargs):
            self._require_real_colormap()
            return getattr(self._impl, name)(*args, **kwargs)
        call.__name__ = name
        return call

    encode_render_pass = _forward("encode_render_pass")
    set_scaling = _forward("set_scaling")
    sph_raw_output_to_image = _forward("sph_raw_output_to_image")
    sph_raw_output_to_content = _forward("sph_raw_output_to_content")
    del _forward

    def autorange_texture(self, mass_scale: float = 1.0):
        """Same decisions as ``autorange(image)`` taken on the device from the bound SPH texture (unscaled accumulators
        times ``mass_scale``), without reading the image back."""
        self._require_real_colormap()
        self._impl.autorange_device(self._input_texture.tensor, mass_scale)

    def autorange(self, sph_render_output: np.ndarray):
        """Re-derive vmin / vmax (and log vs linear) from an SPH image."""
        self._require_real_colormap()
        self._impl.autorange_vmin_vmax(sph_render_output)
