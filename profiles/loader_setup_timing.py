"""Set-up time of a snapshot through the drop-in loader: ArrayDataLoader with the cell layout, the
within-cell shuffle and the reordering on the GPU, against the host (numpy) path.
  python profiles/loader_setup_timing.py [n_particles] [--host]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from topsy_b200 import loader          # noqa: E402
from topsy_b200.device import Device   # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 50_000_000
rs = np.random.default_rng(5)
pos = rs.uniform(-50, 50, (n, 3)).astype(np.float32)
smooth = rs.uniform(0.01, 0.1, n).astype(np.float32)
mass = np.full(n, 1.0 / n, np.float32)
q = rs.normal(size=n).astype(np.float32)
dev = Device()
out = {"n_particles": n}
loader.ArrayDataLoader(dev, pos[:100000], smooth[:100000], mass[:100000])      # warm-up (context, kernels)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    ld = loader.ArrayDataLoader(dev, pos, smooth, mass, quantities={"q": q})
    cols = ld.device_columns(["x", "y", "z", "h", "m"])
    qd = ld.device_quantity("q")
    torch.cuda.synchronize()
    out["device_path_s"] = time.perf_counter() - t0
    del ld, cols, qd
if "--host" in sys.argv:
    t0 = time.perf_counter()
    ld = loader.ArrayDataLoader(dev, pos, smooth, mass, quantities={"q": q}, layout_on_device=False)
    cols = [dev.upload(c) for c in (ld.get_positions()[:, 0], ld.get_positions()[:, 1], ld.get_positions()[:, 2], ld.get_smooth(), ld.get_mass(), ld.get_named_quantity("q"))]
    torch.cuda.synchronize()
    out["host_path_s"] = time.perf_counter() - t0
print(json.dumps(out))
