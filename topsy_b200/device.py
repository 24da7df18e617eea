"""Stand-ins for the two wgpu objects the hot path touches: the device and its textures.

The reference passes a ``wgpu.GPUDevice`` around and allocates ``GPUBuffer`` / ``GPUTexture`` objects on it
(visualizer.py:156-168, sph.py:56-63).  Here a ``Device`` names one CUDA GPU; buffers are PyTorch tensors (PyTorch is
only the allocator) and a ``Texture`` is a tensor plus the wgpu-style format string the rest of topsy's code inspects.
"""
from __future__ import annotations

import numpy as np
import torch

_FORMATS = {            # format -> (channels, torch dtype)
    "r32float": (1, torch.float32), "rg32float": (2, torch.float32), "rgba32float": (4, torch.float32),
    "rgba8unorm": (4, torch.uint8), "bgra8unorm": (4, torch.uint8), "rgba16float": (4, torch.float16),
}


class Texture:
    """2-D image on the device.  ``tensor`` is (height, width, channels); row 0 is the top of the picture."""

    def __init__(self, tensor_or_getter, format: str, label: str = ""):
        self._source = tensor_or_getter
        self.format = format
        self.label = label

    @property
    def tensor(self) -> torch.Tensor:
        return self._source() if callable(self._source) else self._source

    @property
    def size(self):
        t = self.tensor
        return (t.shape[1], t.shape[0], 1)

    @property
    def width(self):
        return self.tensor.shape[1]

    @property
    def height(self):
        return self.tensor.shape[0]

    def create_view(self):
        return self


class Device:
    """One CUDA GPU.  Shared by every Visualizer in the process, like the reference's class-level wgpu device."""

    def __init__(self, index: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("topsy_b200 needs a CUDA device: the SPH projection path has no CPU fallback")
        self.index = torch.cuda.current_device() if index is None else int(index)
        self.torch_device = torch.device("cuda", self.index)
        self._engines = {}
        self.queue = self           # ``device.queue.write_buffer`` spelling of the reference keeps working

    # -- engines (one tsplat context per render resolution) ----------------------------------------------------
    def engine(self, resolution: int):
        from .engine import SplatEngine
        eng = self._engines.get(resolution)
        if eng is None:
            eng = self._engines[resolution] = SplatEngine(resolution, device=self.index)
        return eng

    # -- buffers / textures -----------------------------------------------------------------------------------
    def create_buffer(self, size: int, usage=None, dtype=torch.uint8) -> torch.Tensor:
        return torch.empty(int(size), dtype=dtype, device=self.torch_device)

    def write_buffer(self, buffer: torch.Tensor, data, offset: int = 0) -> None:
        src = torch.from_numpy(np.ascontiguousarray(data)).view(torch.uint8).reshape(-1)
        buffer.view(torch.uint8).reshape(-1)[offset:offset + src.numel()].copy_(src, non_blocking=False)

    def upload(self, array, dtype=np.float32) -> torch.Tensor:
        """Contiguous host array -> new device tensor."""
        return torch.from_numpy(np.ascontiguousarray(array, dtype=dtype)).to(self.torch_device)

    def create_texture(self, size, format: str, usage=None, label: str = "", **_ignored) -> Texture:
        if format not in _FORMATS:
            raise ValueError(f"Unsupported texture format {format}")
        channels, dtype = _FORMATS[format]
        width, height = int(size[0]), int(size[1])
        return Texture(torch.zeros((height, width, channels), dtype=dtype, device=self.torch_device), format, label)

    def read_texture(self, texture: Texture) -> np.ndarray:
        return texture.tensor.cpu().numpy()

    def synchronize(self):
        torch.cuda.synchronize(self.torch_device)
