"""One EXPORT frame of the reference's synthetic GMM snapshot through the drop-in API (for ncu launch lists)."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import topsy_b200 as topsy
from topsy_b200.canvas import offscreen
from topsy_b200.drawreason import DrawReason
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 512
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 40.0
vis = topsy.test(n, render_resolution=R, canvas_class=offscreen.VisualizerCanvas)
vis.scale = scale
for i in range(3):
    vis.rotate(0.01, 0.0)
    vis.render_sph(DrawReason.EXPORT)
torch.cuda.synchronize()
eng = vis._sph._engine
print(eng.stats())
h = vis.data_loader.get_pos_smooth()[:, 3]
w = 2 * h * R / scale
print("wpx quantiles 10/50/90/99:", np.quantile(w, [0.1, 0.5, 0.9, 0.99]), "sum w^2 = %.3e" % (w.astype(np.float64) ** 2).sum())
