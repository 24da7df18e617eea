"""topsy_b200 -- B200-native SPH projection path with topsy's Python surface (Visualizer / SPH / colormap / ...)."""
from __future__ import annotations

__version__ = "0.1.0"

from . import config  # noqa: F401
