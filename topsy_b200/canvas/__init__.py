"""Canvas layer (API of the reference's src/topsy/canvas/__init__.py).  Only the toolkit-neutral interaction logic and
the offscreen canvas belong to the B200 hot path; the Qt / Jupyter front-ends of the reference are windowing code and
are not provided."""
from __future__ import annotations

import time

import numpy as np

from .. import config

ROTATE_RADIANS_PER_PIXEL = 0.01
WHEEL_ZOOM_DIVISOR = 1000.0


class VisualizerCanvasBase:
    """Interaction semantics shared by every canvas (canvas/__init__.py:16-160 of the reference):
    drag = rotate, modifier + drag = pan, wheel = zoom, double click = re-centre on the matter under the cursor
    (found through the depth image), keys s / r / h / w = save, autorange, home view, print the camera."""

    def __init__(self, *args, **kwargs):
        self._visualizer = kwargs.pop("visualizer")
        self._pointer = (0, 0)
        self.width_physical, self.height_physical = 640, 480      # replaced by the first resize event
        self.pixel_ratio = 1
        super().__init__(*args, **kwargs)
        self._dispatch = {
            'pointer_move': self._on_pointer_move,
            'wheel': lambda ev: self.mouse_wheel(ev['dx'], ev['dy']),
            'key_up': lambda ev: self.key_up(ev['key']),
            'resize': lambda ev: self.resize_complete(ev['width'], ev['height'], ev['pixel_ratio']),
            'double_click': lambda ev: self.double_click(ev['x'], ev['y']),
            'pointer_up': lambda ev: self.release_drag(),
        }
        self.add_event_handler(self.event_handler, "*")

    # -- event plumbing -----------------------------------------------------------------------------------------
    def event_handler(self, event):
        handler = self._dispatch.get(event['event_type'])
        if handler is not None:
            handler(event)

    def _on_pointer_move(self, event):
        dx, dy = event['x'] - self._pointer[0], event['y'] - self._pointer[1]
        if event['buttons']:
            (self.shift_drag if event['modifiers'] else self.drag)(dx, dy)
        self._pointer = (event['x'], event['y'])

    # kept for code that pokes the reference's attribute names
    @property
    def _last_x(self):
        return self._pointer[0]

    @property
    def _last_y(self):
        return self._pointer[1]

    # -- gestures -----------------------------------------------------------------------------------------------
    def drag(self, dx, dy):
        self._visualizer.rotate(dx * ROTATE_RADIANS_PER_PIXEL, dy * ROTATE_RADIANS_PER_PIXEL)

    def _pixels_to_world(self, right, up):
        """Displacement in simulation coordinates of a screen-space move of (right, up) logical pixels."""
        vis = self._visualizer
        extent = max(self.width_physical, self.height_physical)
        in_view = np.array([right, up, 0], dtype=np.float32) * (2.0 * self.pixel_ratio * vis.scale / extent)
        return vis.rotation_matrix.T @ in_view

    def shift_drag(self, dx, dy):
        vis = self._visualizer
        vis.position_offset += self._pixels_to_world(dx, -dy)
        vis.display_status("centre = [{:.2f}, {:.2f}, {:.2f}]".format(*vis._sph.position_offset))
        vis.crosshairs_visible = True

    def release_drag(self):
        vis = self._visualizer
        if vis.crosshairs_visible:
            vis.crosshairs_visible = False
            vis.invalidate()

    def mouse_wheel(self, delta_x, delta_y):
        self._visualizer.scale *= np.exp(delta_y / WHEEL_ZOOM_DIVISOR)

    def key_up(self, key):
        vis = self._visualizer
        actions = {'s': vis.save, 'r': vis.colormap_autorange, 'h': vis.reset_view}
        if key in actions:
            actions[key]()
        elif key == 'w':
            shift = np.array2string(vis.position_offset, separator=",")
            turn = np.array2string(vis.rotation_matrix, separator=",")
            print(f".translate({shift}).transform(np.array({turn}))")

    def resize(self, *args):
        """Window-system resize hook of the reference (canvas/__init__.py:96-98); there is no window here."""

    def resize_complete(self, width, height, pixel_ratio=1):
        self.pixel_ratio = pixel_ratio
        self.width_physical, self.height_physical = int(width * pixel_ratio), int(height * pixel_ratio)

    def double_click(self, x, y):
        """Move the clicked point to the view centre, in depth too, then animate the move (GLIDE_TIME)."""
        vis = self._visualizer
        origin = np.array(vis.position_offset, copy=True)
        half_w = self.width_physical / (2 * self.pixel_ratio)
        half_h = self.height_physical / (2 * self.pixel_ratio)
        vis.position_offset += self._pixels_to_world(half_w - x, y - half_h)
        depth_map = vis.get_depth_image()
        depth_here = depth_map[depth_map.shape[0] // 2, depth_map.shape[1] // 2]
        if not np.isnan(depth_here):
            vis.position_offset += vis.rotation_matrix.T @ np.array([0, 0, -depth_here], dtype=np.float32)
        destination = vis.position_offset
        vis.position_offset = origin
        began = time.time()

        def blend(t):                      # smooth-step made of an arctangent, as in the reference
            w = np.arctan(5 * (2 * t - 1)) / np.pi + 0.5
            return origin + w * (destination - origin)

        def step():
            t = (time.time() - began) / config.GLIDE_TIME
            if t <= 1:
                self.call_later(0.0, step)
                vis.position_offset = blend(t)
            else:
                vis.position_offset = destination

        self.call_later(1.0 / config.TARGET_FPS, step)

    @classmethod
    def call_later(cls, delay, fn, *args):
        raise NotImplementedError()


def __getattr__(name):
    # the reference picks a Qt or Jupyter canvas here; the B200 build is headless, so the default is offscreen
    if name == "VisualizerCanvas":
        from .offscreen import VisualizerCanvas
        return VisualizerCanvas
    raise AttributeError(name)
