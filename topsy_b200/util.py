"""Small helpers (reference: src/topsy/util.py).  The WGSL loader / preprocessor of the reference has no counterpart
here -- the shader variants are template parameters / runtime switches of the CUDA kernels."""
from __future__ import annotations

import os

import numpy as np
import torch


class TimeGpuOperation:
    """Times what is enqueued inside ``with timer:`` on the GPU and keeps a running mean over the last frames.

    The reference brackets each submit with two blocking ``on_submitted_work_done_sync`` calls and host clocks
    (util.py:76-115); here a pair of CUDA events on the current stream measures device time, and only the closing event
    is waited on (the progression needs the duration before it sizes the next block)."""

    def __init__(self, device, n_frames_smooth: int = 10):
        self.device = device
        self.n_frames_smooth = n_frames_smooth
        self._recent_times = []
        self._current_frame_duration = 0.0
        self._start = torch.cuda.Event(enable_timing=True)
        self._stop = torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self._start.record(torch.cuda.current_stream(self.device.torch_device))
        return self

    def __exit__(self, *exc):
        self._stop.record(torch.cuda.current_stream(self.device.torch_device))
        self._stop.synchronize()
        self._current_frame_duration += self._start.elapsed_time(self._stop) * 1e-3

    def end_frame(self):
        self.last_duration = self._current_frame_duration
        self._current_frame_duration = 0.0
        self._recent_times.append(self.last_duration)
        if len(self._recent_times) > self.n_frames_smooth:
            self._recent_times.pop(0)

    def total_time_in_frame(self):
        return self._current_frame_duration

    @property
    def running_mean_duration(self):
        return np.mean(self._recent_times)


def is_inside_ipython():
    try:
        __IPYTHON__  # noqa: F821
        return True
    except NameError:
        return False


def is_inside_jupyter_notebook():
    return "JPY_SESSION_NAME" in os.environ
