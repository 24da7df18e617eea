"""Progressive rendering: which particles to splat in the next render call (reference: src/topsy/progressive_render.py).

A frame is a sequence of *blocks*; a block is a logical index range ``[start, start + length)`` into the particle
ordering.  Interactive frames render as many particles as the measured splat rate allows in 1/TARGET_FPS and leave the
rest to REFINE frames that keep accumulating; EXPORT frames walk the whole snapshot in chunks.  The image is always
rescaled by N / particles-rendered (``end_frame_get_scalefactor``).  With a cell layout, a logical range is mapped to the
same fraction of every selected cell, so any block is a spatially fair subsample.
"""
from __future__ import annotations

import math

import numpy as np

from . import config
from .cell_layout import CellLayout
from .drawreason import DrawReason

_KEEP_IMAGE = (DrawReason.PRESENTATION_CHANGE, DrawReason.REFINE)


class RenderProgression:
    """Block recommender driven by the time the previous blocks took (progressive_render.py:8-137)."""

    def __init__(self, total_particles, initial_particles=None):
        if initial_particles is None:
            initial_particles = int(config.INITIAL_PARTICLES_TO_RENDER)
        self._max_num_particles = total_particles
        self._recommended_num_particles_to_render = min(initial_particles, total_particles)
        self._start_index = 0
        self._last_num_to_render = 1
        self._current_draw_reason = None

    def get_max_particle_regions_per_block(self):
        return 1

    # -- frame protocol -------------------------------------------------------------------------------------
    def start_frame(self, draw_reason: DrawReason) -> bool:
        """Begin a frame; True means the accumulation image must be cleared (the frame restarts at particle 0)."""
        self._current_draw_reason = draw_reason
        self._first_block_in_frame = True
        self._total_num_rendered_in_frame = 0
        if draw_reason in _KEEP_IMAGE:
            return False
        self._start_index = 0
        return True

    def get_block(self, time_elapsed_in_frame: float):
        """``([start], [length])`` of the next block, or None when the frame is over."""
        reason = self._current_draw_reason
        if reason is None:
            raise RuntimeError("get_block called without a current frame")
        if reason == DrawReason.PRESENTATION_CHANGE:
            return None
        remaining = self._max_num_particles - self._start_index
        if remaining <= 0:
            return None

        if reason == DrawReason.EXPORT:
            # chunked so that one call never carries more than the export budget *inside the selected volume*
            chunk = int(config.MAX_PARTICLES_PER_EXPORT_RENDERCALL / self.get_fraction_volume_selected())
            count = min(remaining, chunk)
        else:
            frame_budget = 1.0 / config.TARGET_FPS
            if self._first_block_in_frame:
                self._first_block_in_frame = False
                time_available = frame_budget
            else:
                time_available = frame_budget - time_elapsed_in_frame
            if time_available <= 0.4 * frame_budget:
                return None         # one block per interactive frame; leftovers go to the next (REFINE) frame
            count = min(remaining, int(self._recommended_num_particles_to_render * time_available * config.TARGET_FPS))
        self._last_num_to_render = count
        return ([self._start_index], [count])

    def end_block(self, time_elapsed_in_frame: float):
        self._start_index += self._last_num_to_render
        self._total_num_rendered_in_frame += self._last_num_to_render
        self._time_in_frame = time_elapsed_in_frame

    def set_time_in_frame(self, time_elapsed_in_frame: float):
        """The frame's device time, when it was not known yet at the last ``end_block`` (EXPORT frames are not
        synchronised block by block); feeds the particle-budget update of ``end_frame_get_scalefactor``."""
        self._time_in_frame = time_elapsed_in_frame

    def end_frame_get_scalefactor(self):
        """Close the frame, adapt the particle budget, return N / particles accumulated so far."""
        self._perform_particle_number_update()
        self._current_draw_reason = None
        return self._max_num_particles / self._start_index

    def _perform_particle_number_update(self):
        achievable = int(self._total_num_rendered_in_frame / (self._time_in_frame * config.TARGET_FPS))
        achievable = max(1, min(achievable, self._max_num_particles))
        if self._current_draw_reason == DrawReason.REFINE:
            return
        current = self._recommended_num_particles_to_render
        mismatch = abs(math.log2(achievable) - math.log2(current))
        if mismatch > 1.5:
            self._recommended_num_particles_to_render = achievable                          # way off: jump
        elif mismatch > 0.3:
            self._recommended_num_particles_to_render = int(achievable ** 0.3 * current ** 0.7)   # drift: damped step

    def needs_refine(self):
        return self._start_index < self._max_num_particles

    # -- spatial selection (no-ops without cells) -------------------------------------------------------------
    def select_sphere(self, cen, radius):
        pass

    def select_all(self):
        pass

    def get_fraction_volume_selected(self):
        return 1.0


class RenderProgressionWithCells(RenderProgression):
    """Maps logical blocks onto per-cell ranges and restricts them to the selected cells (:139-215)."""

    def __init__(self, cell_layout: CellLayout, total_particles: int, initial_particles=None):
        super().__init__(total_particles, initial_particles)
        self._cell_layout = cell_layout
        # a fixed pseudo-random phase per cell decorrelates the integer truncation between cells, so even a block
        # with < 1 particle per cell on average picks *some* particles, evenly over space
        self._cell_phase_shifts = np.random.RandomState(1337).permutation(cell_layout.get_num_cells())
        self._selected_cells_hash = 0
        self.select_all()

    def get_max_particle_regions_per_block(self):
        return self._cell_layout.get_num_cells()

    def _map_logical_range_to_actual_ranges(self, start, length):
        layout = self._cell_layout
        n_total = layout.get_num_particles()
        n_cells = layout.get_num_cells()
        frac_lo = start / n_total
        frac_hi = (start + length) / n_total
        phase = self._cell_phase_shifts / n_cells
        per_cell = layout._lengths.astype(np.float64)
        first = (frac_lo * per_cell + phase).astype(np.intp)
        last = (frac_hi * per_cell + phase).astype(np.intp)
        starts = (first + layout._offsets)[self._selected_cells]
        counts = (last - first)[self._selected_cells]
        keep = counts > 0
        return starts[keep], counts[keep]

    def get_block(self, time_elapsed_in_frame: float):
        block = super().get_block(time_elapsed_in_frame)
        if block is None:
            return None
        (start,), (count,) = block
        if count == self._max_num_particles:
            return block                     # everything in one go: no per-cell mapping (and no spatial selection)
        return self._map_logical_range_to_actual_ranges(start, count)

    def select_all(self):
        self._selected_cells = np.arange(self._cell_layout.get_num_cells())
        self._note_selection()

    def select_sphere(self, cen, r):
        self._selected_cells = self._cell_layout.cells_in_sphere(cen, r)
        self._note_selection()

    def _note_selection(self):
        digest = hash(self._selected_cells.tobytes())
        if digest != self._selected_cells_hash:
            self._selected_cells_hash = digest
            self._update_particle_ranges = True

    def get_fraction_volume_selected(self):
        return max(1, len(self._selected_cells)) / self._cell_layout.get_num_cells()
