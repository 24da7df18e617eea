"""K6 (k_reduce_colormap) and K6b (k_allreduce_image) over NVLink from ONE process that owns two GPUs, so that the kernels
can be profiled with ncu (a multi-rank job cannot be wrapped in ncu).  GPU 0 reduces the lower half of the rows of a
4096^2 density image from its own and GPU 1's partial image and colormaps it; timing with CUDA events, NVLink bytes = the
peer's slab.
  ncu --set full -k regex:'k_reduce_colormap|k_allreduce_image' -o gpurun_out/prof_k6 python profiles/k6_peer_ncu.py
"""
import ctypes
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from topsy_b200 import _native as N            # noqa: E402
from topsy_b200.colormap import luts           # noqa: E402
from topsy_b200.engine import SplatEngine      # noqa: E402

R, C = 4096, 1
assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
img0 = torch.rand((R, R, C), device="cuda:0") + 0.1
img1 = torch.rand((R, R, C), device="cuda:1") + 0.1
tmp = torch.empty_like(img0); tmp.copy_(img1)
assert torch.cuda.can_device_access_peer(0, 1)
eng = SplatEngine(R, device=0)
N.check(eng.lib.tsplat_enable_peer_access(0, 1))        # kernels on GPU 0 may load / store GPU 1's memory
out = torch.empty((R, R, 4), dtype=torch.uint8, device="cuda:0")
red0, red1 = torch.empty_like(img0), torch.empty_like(img1)
scale0, scale1 = torch.ones(4, device="cuda:0"), torch.ones(4, device="cuda:1")
lut = torch.from_numpy(luts.colormap_table_1d("viridis", 1000)).to("cuda:0")
params = N.ColormapParams(vmin=-1.0, vmax=0.5, density_vmin=0, density_vmax=1, window_aspect_ratio=1.0, gamma=1.0,
                          kind=N.CMAP_DENSITY, log_scale=1)
arr = ctypes.c_void_p * 2
peers = arr(img0.data_ptr(), img1.data_ptr())
outs = arr(red0.data_ptr(), red1.data_ptr())
scales = arr(scale0.data_ptr(), scale1.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream(0).cuda_stream)
rows = R // 2


def k6():
    N.check(eng.lib.tsplat_reduce_colormap(eng._ctx, peers, 2, C, 0, rows, ctypes.byref(params), ctypes.c_void_p(lut.data_ptr()),
                                           1000, 1, ctypes.c_void_p(out.data_ptr()), N.FMT_RGBA8, None, stream))


def k6b():
    N.check(eng.lib.tsplat_allreduce_image(eng._ctx, peers, outs, scales, 2, C, 0, rows, N.REDUCE_SUM, stream))


res = {"resolution": R, "channels": C, "rows_reduced": rows, "peer_bytes_read": rows * R * C * 4}
for name, fn, link_bytes in (("k_reduce_colormap", k6, rows * R * C * 4), ("k_allreduce_image", k6b, 2 * rows * R * C * 4)):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 20
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = {"ms": ms, "nvlink_bytes": link_bytes, "nvlink_GBs": link_bytes / ms / 1e6,
                 "frac_of_measured_peer_copy_770GBs": link_bytes / ms / 1e6 / 770.0}
# correctness of the peer path
want = (img0[:rows] + tmp[:rows])
assert torch.allclose(red0[:rows], want) and torch.allclose(red1[:rows].to("cuda:0"), want)
print(json.dumps(res))
