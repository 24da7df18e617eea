"""Headless canvas (reference: src/topsy/canvas/offscreen.py, built on rendercanvas' OffscreenRenderCanvas).

rendercanvas is not a dependency here: ``OffscreenCanvas`` provides the few things the visualizer uses -- an event
handler registry, ``request_draw`` coalescing, a context that hands out a target texture of the canvas format, and
``draw()`` which runs the pending draw function and returns the frame as a numpy array."""
from __future__ import annotations

from . import VisualizerCanvasBase

_PRESENT_FORMATS = {"rgba8unorm": "rgba-u8", "bgra8unorm": "bgra-u8", "rgba16float": "rgba-f16"}


class _Context:
    def __init__(self, canvas):
        self._canvas = canvas
        self._device = None
        self._format = None
        self._texture = None

    def get_preferred_format(self, adapter=None):
        return "rgba8unorm"

    def configure(self, device, format):
        allowed = self._canvas._rc_get_present_methods()["bitmap"]["formats"]
        if _PRESENT_FORMATS.get(format) not in allowed:
            raise ValueError(f"Canvas cannot present format {format} (supports {allowed})")
        self._device, self._format, self._texture = device, format, None

    def get_current_texture(self):
        w, h = self._canvas.width_physical, self._canvas.height_physical
        if self._texture is None or self._texture.size[:2] != (w, h):
            self._texture = self._device.create_texture((w, h, 1), self._format, label="canvas")
        return self._texture


class OffscreenCanvas:
    def __init__(self, *args, size=(640, 480), pixel_ratio=1, title="", **kwargs):
        self._handlers = []
        self._pending_draw = None
        self._later = []
        self._context = _Context(self)
        self._logical_size = size
        self._title = title

    def _rc_get_present_methods(self):
        return {"bitmap": {"formats": ["rgba-u8", "rgba-f16"]}}

    def add_event_handler(self, handler, *types):
        self._handlers.append((handler, types))

    def submit_event(self, event):
        for handler, types in self._handlers:
            if "*" in types or event.get('event_type') in types:
                handler(event)

    def get_context(self, kind="wgpu"):
        return self._context

    def request_draw(self, draw_function=None):
        if draw_function is not None:
            self._pending_draw = draw_function

    def draw(self):
        """Run the pending draw (if any) and return the presented frame as an (H, W, 4) array."""
        self._run_later()
        fn, self._pending_draw = self._pending_draw, None
        if fn is not None:
            fn()
        return self._context.get_current_texture().tensor.cpu().numpy()

    def _run_later(self):
        pending, self._later = self._later, []
        for fn, args in pending:
            fn(*args)

    def set_logical_size(self, width, height):
        self._logical_size = (width, height)
        self.submit_event({'event_type': 'resize', 'width': width, 'height': height, 'pixel_ratio': 1})

    def close(self):
        pass


class VisualizerCanvas(VisualizerCanvasBase, OffscreenCanvas):
    def call_later(self, delay, fn, *args):
        self._later.append((fn, args))
