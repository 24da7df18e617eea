import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import time, numpy as np, torch
import topsy_b200 as topsy
from topsy_b200.canvas import offscreen
from topsy_b200.drawreason import DrawReason
for n, R in [(1_000_000, 512), (10_000_000, 1024)]:
    t0 = time.perf_counter()
    vis = topsy.test(n, render_resolution=R, canvas_class=offscreen.VisualizerCanvas)
    vis.scale = 40.0
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"n={n} R={R}: construct {t1-t0:.2f}s")
    for reason in (DrawReason.EXPORT, DrawReason.CHANGE, DrawReason.REFINE):
        ts = []
        for i in range(6):
            vis.rotate(0.01, 0.0) if reason != DrawReason.REFINE else None
            torch.cuda.synchronize(); a = time.perf_counter()
            vis.render_sph(reason)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - a)
        print(f"   render_sph({reason.name}): {np.median(ts)*1e3:.2f} ms  (min {min(ts)*1e3:.2f})")
    ts = []
    for i in range(5):
        vis.rotate(0.01, 0.0)
        torch.cuda.synchronize(); a = time.perf_counter()
        img = vis.get_sph_presentation_image()
        ts.append(time.perf_counter() - a)
    print(f"   get_sph_presentation_image: {np.median(ts)*1e3:.2f} ms")
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for i in range(5):
        vis.rotate(0.01, 0.0); vis.render_sph(DrawReason.EXPORT)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
