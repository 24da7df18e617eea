"""Host-side cost of an interactive frame through the drop-in API when the GPU work is small (cells on, 4096 ranges)."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import time, cProfile, pstats, numpy as np, torch
import topsy_b200 as topsy
from topsy_b200.canvas import offscreen
from topsy_b200.drawreason import DrawReason
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
vis = topsy.test(n, render_resolution=512, canvas_class=offscreen.VisualizerCanvas, with_cells=True)
vis.scale = 300.0
def frame(reason):
    torch.cuda.synchronize(); a = time.perf_counter()
    vis.render_sph(reason); torch.cuda.synchronize()
    return (time.perf_counter() - a) * 1e3
for reason in (DrawReason.EXPORT, DrawReason.CHANGE):
    ts = []
    for i in range(12):
        vis.rotate(0.01, 0.0)
        ts.append(frame(reason))
    print(f"{reason.name}: median {np.median(ts):.2f} ms  min {min(ts):.2f}  scale {vis._sph.last_render_mass_scale:.2f}  recommended {vis._sph._render_progression._recommended_num_particles_to_render}")
vis._sph._render_progression._recommended_num_particles_to_render = 20000
pr = cProfile.Profile(); pr.enable()
vis.rotate(0.01, 0.0); t = frame(DrawReason.CHANGE)
pr.disable()
print("cold-budget CHANGE frame %.2f ms, fraction %.3f" % (t, 1 / vis._sph.last_render_mass_scale))
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
