"""Profiling hooks.  The reference emits macOS signposts (src/topsy/performance.py:3-21); here the same call sites emit
NVTX ranges/marks (visible in Nsight Systems / ncu --nvtx) when CUDA is available and are no-ops otherwise."""
from __future__ import annotations

import contextlib


class _NvtxSignposter:
    def __init__(self):
        try:
            import torch
            self._nvtx = torch.cuda.nvtx if torch.cuda.is_available() else None
        except Exception:   # pragma: no cover
            self._nvtx = None

    def emit_event(self, name, *args, **kwargs):
        if self._nvtx is not None:
            self._nvtx.mark(str(name))

    def begin_interval(self, name, *args, **kwargs):
        if self._nvtx is not None:
            self._nvtx.range_push(str(name))
        return name

    def end_interval(self, *args, **kwargs):
        if self._nvtx is not None:
            self._nvtx.range_pop()

    @contextlib.contextmanager
    def use_interval(self, name, *args, **kwargs):
        self.begin_interval(name)
        try:
            yield
        finally:
            self.end_interval()


signposter = _NvtxSignposter()
