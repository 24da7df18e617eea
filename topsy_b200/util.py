"""Small helpers (reference: src/topsy/util.py).  The WGSL loader / preprocessor of the reference has no counterpart
here -- the shader variants are template parameters / runtime switches of the CUDA kernels."""
from __future__ import annotations

import os

import numpy as np
import torch


class TimeGpuOperation:
    """Times what is enqueued inside ``with timer:`` on the GPU and keeps a running mean over the last frames.

    The reference brackets each submit with two blocking ``on_submitted_work_done_sync`` calls and host clocks
    (util.py:76-115); here a pair of CUDA events on the current stream measures device time.  The host waits for a
    block's closing event only when somebody asks for the time: an interactive frame does after every block (the
    progression sizes the next block from it), an EXPORT frame -- whose block sizes do not depend on the time -- asks
    once at the end of the frame, so its blocks are enqueued back to back (``total_time_in_frame(wait=False)``)."""

    def __init__(self, device, n_frames_smooth: int = 10):
        self.device = device
        self.n_frames_smooth = n_frames_smooth
        self._recent_times = []
        self._current_frame_duration = 0.0
        self._pending = []            # (start, stop) event pairs whose duration has not been read yet
        self._spare = []              # event pairs to reuse
        self._open = None

    def _stream(self):
        return torch.cuda.current_stream(self.device.torch_device)

    def __enter__(self):
        pair = self._spare.pop() if self._spare else (torch.cuda.Event(enable_timing=True),
                                                      torch.cuda.Event(enable_timing=True))
        pair[0].record(self._stream())
        self._open = pair
        return self

    def __exit__(self, *exc):
        self._open[1].record(self._stream())
        self._pending.append(self._open)
        self._open = None

    def _resolve(self):
        for start, stop in self._pending:
            stop.synchronize()
            self._current_frame_duration += start.elapsed_time(stop) * 1e-3
        self._spare.extend(self._pending)
        self._pending.clear()

    def end_frame(self):
        self._resolve()
        self.last_duration = self._current_frame_duration
        self._current_frame_duration = 0.0
        self._recent_times.append(self.last_duration)
        if len(self._recent_times) > self.n_frames_smooth:
            self._recent_times.pop(0)

    def total_time_in_frame(self, wait: bool = True):
        """Device time of the frame's blocks so far; ``wait=False`` counts only the blocks already read back."""
        if wait:
            self._resolve()
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
