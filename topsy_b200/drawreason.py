"""Why a frame is being drawn (same members and values as the reference's src/topsy/drawreason.py:3-9)."""
import enum


class DrawReason(enum.Enum):
    INITIAL_UPDATE = 1        # first render of a freshly created pipeline
    CHANGE = 2                # camera / data changed: restart the progressive render from particle 0
    REFINE = 3                # keep the accumulated image and add the next block of particles
    PRESENTATION_CHANGE = 4   # only the colormap changed: do not touch the SPH image
    EXPORT = 5                # full-quality render of every particle, regardless of time budget
