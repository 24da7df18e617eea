"""Canvas layer (reference: src/topsy/canvas/__init__.py).  Only the toolkit-neutral base class and the offscreen canvas
are part of the B200 hot path; the Qt / Jupyter front-ends of the reference are windowing code and are not provided."""
from __future__ import annotations

import copy
import time

import numpy as np

from .. import config


class VisualizerCanvasBase:
    """Mouse / keyboard semantics shared by every canvas: drag rotates, shift-drag pans, wheel zooms, double click
    re-centres on the point under the cursor using the depth image (canvas/__init__.py:16-160)."""

    def __init__(self, *args, **kwargs):
        self._visualizer = kwargs.pop("visualizer")
        self._last_x = 0
        self._last_y = 0
        self.width_physical, self.height_physical = 640, 480      # until the first resize event
        self.pixel_ratio = 1
        super().__init__(*args, **kwargs)
        self.add_event_handler(self.event_handler, "*")

    def event_handler(self, event):
        kind = event['event_type']
        if kind == 'pointer_move':
            if len(event['buttons']) > 0:
                move = self.drag if len(event['modifiers']) == 0 else self.shift_drag
                move(event['x'] - self._last_x, event['y'] - self._last_y)
            self._last_x, self._last_y = event['x'], event['y']
        elif kind == 'wheel':
            self.mouse_wheel(event['dx'], event['dy'])
        elif kind == 'key_up':
            self.key_up(event['key'])
        elif kind == 'resize':
            self.resize_complete(event['width'], event['height'], event['pixel_ratio'])
        elif kind == 'double_click':
            self.double_click(event['x'], event['y'])
        elif kind == 'pointer_up':
            self.release_drag()

    def drag(self, dx, dy):
        self._visualizer.rotate(dx * 0.01, dy * 0.01)

    def _screen_to_world(self, dx, dy):
        span = max(self.width_physical, self.height_physical)
        shift = 2.0 * self.pixel_ratio * np.array([dx, dy, 0], dtype=np.float32) / span * self._visualizer.scale
        return self._visualizer.rotation_matrix.T @ shift

    def shift_drag(self, dx, dy):
        self._visualizer.position_offset += self._screen_to_world(dx, -dy)
        self._visualizer.display_status("centre = [{:.2f}, {:.2f}, {:.2f}]".format(*self._visualizer._sph.position_offset))
        self._visualizer.crosshairs_visible = True

    def key_up(self, key):
        if key == 's':
            self._visualizer.save()
        elif key == 'r':
            self._visualizer.colormap_autorange()
        elif key == 'h':
            self._visualizer.reset_view()
        elif key == 'w':
            offset = np.array2string(self._visualizer.position_offset, separator=",")
            rot = np.array2string(self._visualizer.rotation_matrix, separator=",")
            print(f".translate({offset}).transform(np.array({rot}))")

    def mouse_wheel(self, delta_x, delta_y):
        self._visualizer.scale *= np.exp(delta_y / 1000)

    def release_drag(self):
        if self._visualizer.crosshairs_visible:
            self._visualizer.crosshairs_visible = False
            self._visualizer.invalidate()

    def resize_complete(self, width, height, pixel_ratio=1):
        self.width_physical = int(width * pixel_ratio)
        self.height_physical = int(height * pixel_ratio)
        self.pixel_ratio = pixel_ratio

    def double_click(self, x, y):
        vis = self._visualizer
        start = copy.copy(vis.position_offset)
        cx = self.width_physical / (2 * self.pixel_ratio)
        cy = self.height_physical / (2 * self.pixel_ratio)
        vis.position_offset += self._screen_to_world(cx - x, y - cy)
        depth = vis.get_depth_image()
        central = depth[depth.shape[0] // 2, depth.shape[1] // 2]
        if not np.isnan(central):
            vis.position_offset += vis.rotation_matrix.T @ np.array([0, 0, -central], dtype=np.float32)
        target = vis.position_offset
        vis.position_offset = start          # the work is done; now glide there so the motion is readable
        t0 = time.time()

        def ease(t):
            w = np.arctan(5 * (t * 2 - 1)) / np.pi + 0.5
            return (1 - w) * start + w * target

        def glide():
            t = (time.time() - t0) / config.GLIDE_TIME
            if t > 1:
                vis.position_offset = target
            else:
                self.call_later(0.0, glide)
                vis.position_offset = ease(t)

        self.call_later(1.0 / config.TARGET_FPS, glide)

    @classmethod
    def call_later(cls, delay, fn, *args):
        raise NotImplementedError()


def __getattr__(name):
    # the reference picks a Qt or Jupyter canvas here; the B200 build is headless, so the default is offscreen
    if name == "VisualizerCanvas":
        from .offscreen import VisualizerCanvas
        return VisualizerCanvas
    raise AttributeError(name)
