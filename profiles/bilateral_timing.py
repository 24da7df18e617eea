"""K10 timing: bilateral filter of a (R, R, 2) image at the reference's largest window (kernel_size 101), new tiled
kernel vs the direct evaluation it replaced (forced through kernel sizes whose tile does not fit).
  python profiles/bilateral_timing.py [R]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from topsy_b200.engine import SplatEngine   # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = SplatEngine(R)
g = torch.Generator(device="cuda"); g.manual_seed(1)
img = torch.rand((R, R, 2), generator=g, device="cuda")
out = torch.empty_like(img)
res = {"resolution": R}
for ks in (11, 41, 101):
    for _ in range(2):
        eng.bilateral_filter(img, out, 0.02 * R, 0.04, ks)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 3
    for _ in range(n):
        eng.bilateral_filter(img, out, 0.02 * R, 0.04, ks)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[f"kernel_size_{ks}_ms"] = ms
    res[f"kernel_size_{ks}_Gtaps_per_s"] = R * R * (2 * (ks // 2) + 1) ** 2 / ms / 1e6
print(json.dumps(res))
