"""Camera -> clip-space transform of the SPH pass (reference: SPH._get_transform_params, src/topsy/sph.py:268-299)."""
from __future__ import annotations

import numpy as np


def transform_matrix(rotation_matrix, position_offset, scale) -> np.ndarray:
    """Row-major float32 4x4 M such that clip = M @ (x, y, z, 1).

    Built in float64 as  clip_z_squash @ (rotation/scale (+) 1) @ translate(position_offset)  and rounded once to
    float32, exactly like the reference (which then uploads the transpose because WGSL matrices are column-major).
    x, y land in [-1, 1] across the view; z is squashed so that |z_rotated| <= scale maps onto [0, 1]."""
    translate = np.eye(4)
    translate[:3, 3] = np.asarray(position_offset, dtype=np.float64)
    squash_z = np.eye(4)
    squash_z[2, 2] = 0.5
    squash_z[2, 3] = 0.5
    rot = np.zeros((4, 4))
    rot[:3, :3] = np.asarray(rotation_matrix, dtype=np.float64)
    rot = rot / scale
    rot[3, 3] = 1.0
    return (squash_z @ rot @ translate).astype(np.float32)


def rotation_about_y(angle):
    """What the reference calls ``_x_rotation_matrix`` (visualizer.py:353-357): horizontal mouse drag."""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rotation_about_x(angle):
    """What the reference calls ``_y_rotation_matrix`` (visualizer.py:347-351): vertical mouse drag."""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def rotate(rotation_matrix, x_angle, y_angle):
    """Visualizer.rotate (visualizer.py:194-197)."""
    return rotation_about_y(x_angle) @ rotation_about_x(y_angle) @ np.asarray(rotation_matrix)
