"""Run under torchrun on >= 2 GPUs: sharded splat + fused P2P reduce/colormap against the single-GPU result and the
CPU oracle.  Invoked by tests/test_distributed_gpu.py (pytest -m gpu) and usable by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from oracle import c_oracle as co                     # noqa: E402
from oracle import topsy_oracle as o                  # noqa: E402
from topsy_b200 import _native as N                   # noqa: E402
from topsy_b200 import distributed as D               # noqa: E402
from topsy_b200.cell_layout import CellLayout         # noqa: E402
from topsy_b200.colormap import luts                  # noqa: E402


def visualizer_check(rank, world):
    """The drop-in classes under torchrun: every rank's Visualizer loads the same snapshot, keeps its stripe
    (distributed.shard_loader), splats it and all-reduces the image (distributed.ImageExchange / K6b); images, autorange and
    presentation must equal those of an unsharded Visualizer of the same snapshot on every rank."""
    from topsy_b200 import loader
    from topsy_b200.canvas import offscreen
    from topsy_b200.drawreason import DrawReason
    from topsy_b200.visualizer import Visualizer

    def make(sharded, **kw):
        previous = D.set_sharding(sharded)
        try:
            return Visualizer(canvas_class=offscreen.VisualizerCanvas, render_resolution=200, **kw)
        finally:
            D.set_sharding(previous)

    cases = [
        ("density, the reference's 1000-particle fixture", dict(data_loader_args=(1000,)), None, 200.0, "get_sph_image"),
        ("weighted quantity, fixture with cells", dict(data_loader_args=(20000,), data_loader_kwargs=dict(with_cells=True)),
         "test-quantity", 20.0, "get_sph_image"),
        ("rgb", dict(data_loader_args=(5000,), render_mode="rgb"), None, 20.0, "get_sph_image"),
        ("depth image", dict(data_loader_args=(5000,)), None, 20.0, "get_depth_image"),
    ]
    for what, kw, quantity, scale, getter in cases:
        ref_vis, vis = make(False, **kw), make(True, **kw)
        assert len(vis.data_loader) < len(ref_vis.data_loader) == vis.data_loader.global_num_particles, what
        for v in (ref_vis, vis):
            if quantity is not None:
                v.quantity_name = quantity
            v.scale = scale
            v.rotate(0.2, 0.4)
        if getter == "get_depth_image":
            ref_vis._sph.render(DrawReason.EXPORT); vis._sph.render(DrawReason.EXPORT)
            want, got = ref_vis._sph.get_depth_image(DrawReason.EXPORT), vis._sph.get_depth_image(DrawReason.EXPORT)
            ok = np.isfinite(want)
            assert np.array_equal(ok, np.isfinite(got)) and np.abs(got[ok] - want[ok]).max() <= 1e-3 * scale, what
            continue
        ref_vis.render_sph(DrawReason.EXPORT); vis.render_sph(DrawReason.EXPORT)
        want, got = ref_vis._sph.get_image().astype(np.float64), vis._sph.get_image().astype(np.float64)
        for c in range(want.shape[2]):
            mag = np.abs(want[..., c])
            big = mag > 1e-6 * mag.max()
            if quantity is not None and c == 1:      # signed channel: relative to the accumulated magnitude (see test_gpu_parity)
                big &= np.abs(want[..., 1]) > 1e-3 * np.abs(want[..., 1]).max()
            if not big.any():                        # e.g. the all-zero quantity channel of a plain density render
                assert not got[..., c].any(), (what, c)
                continue
            rel = np.abs(got[..., c][big] - want[..., c][big]) / mag[big]
            assert rel.max() <= 1e-4, (what, c, rel.max())
        # device autorange and presentation run on the all-reduced image on every rank
        ref_vis.colormap_autorange(); vis.colormap_autorange()
        for k in ("vmin", "vmax"):
            a, b = ref_vis.colormap.get_parameter(k), vis.colormap.get_parameter(k)
            assert abs(a - b) <= 1e-3 * max(1.0, abs(a)), (what, k, a, b)
        pa, pb = ref_vis.get_sph_presentation_image(), vis.get_sph_presentation_image()
        d = np.abs(pa.astype(np.float64) - pb.astype(np.float64))
        assert d.max() <= (2 if pa.dtype == np.uint8 else 2e-2) and (d > 0).mean() < 2e-2, (what, d.max(), (d > 0).mean())
    # interactive frames: ranks may cover different fractions of their stripes; the reduce weights them by their mass scale
    ref_vis, vis = make(False, data_loader_args=(200000,), data_loader_kwargs=dict(with_cells=True)), \
        make(True, data_loader_args=(200000,), data_loader_kwargs=dict(with_cells=True))
    for v in (ref_vis, vis):
        v.scale = 30.0
    # force one block per frame and unequal fractions per rank (a B200 would otherwise finish the stripe inside one frame)
    def interactive_sequence(v, per_frame):
        rp = v._sph._render_progression
        rp._recommended_num_particles_to_render = per_frame
        rp._perform_particle_number_update = lambda: None
        v._sph._render_timer.total_time_in_frame = lambda wait=True: 1.0
        v.render_sph(DrawReason.CHANGE)
        frames = 1
        while v._sph.needs_refine() and frames < 100:     # collective when sharded: true until the slowest rank is done
            v.render_sph(DrawReason.REFINE); frames += 1
        return frames

    # the unsharded reference runs the same kind of sequence: interactive frames draw only the cells selected by
    # select_sphere(-offset, 1.2 scale) (sph.py:313), an EXPORT frame draws everything (reference quirk)
    assert interactive_sequence(ref_vis, 47000) > 1
    # (never the whole stripe in one block: a block that is the whole set skips the cell selection, progressive_render.py:197-198)
    assert interactive_sequence(vis, len(vis.data_loader) // 4 + 300 * rank) > 1
    want, got = ref_vis._sph.get_image()[..., 0].astype(np.float64), vis._sph.get_image()[..., 0].astype(np.float64)
    big = want > 1e-6 * want.max()
    rel = np.abs(got[big] - want[big]) / want[big]
    assert rel.max() <= 1e-4, ("progressive", rel.max())
    # and a partially refined state: after ONE frame the ranks have drawn different fractions (25 % + 300 rank) of their
    # stripes; the mass-scale weighted reduce must still estimate the full image (statistically: mean ratio ~ 1)
    vis.invalidate(DrawReason.CHANGE)
    vis.render_sph(DrawReason.CHANGE)
    part = vis._sph.get_image()[..., 0].astype(np.float64)
    bright = want > 0.05 * want.max()
    ratio = part[bright].sum() / want[bright].sum()
    assert abs(ratio - 1.0) < 0.03, ("partial frame estimate", ratio)
    if rank == 0:
        print("VISUALIZER_SHARDING_OK: density / weighted / rgb / depth / progressive equal the unsharded Visualizer")


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R, n = 256, 300_000
    rs = np.random.RandomState(7)                      # identical snapshot on every rank; each keeps its stripe
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (0.01 * np.exp(rs.normal(size=n) * 0.7)).astype(np.float32)
    m = rs.uniform(0.5, 1.5, n).astype(np.float32); q = np.exp(rs.normal(size=n)).astype(np.float32)
    layout, order = CellLayout.from_positions(pos, np.float32(-1.001), np.float32(1.001), 8)
    pos, h, m, q = pos[order], h[order], m[order], q[order]
    mine = D.shard_indices(layout._offsets, layout._lengths, rank, world)
    M = o.transform_matrix(o.rotate(np.eye(3), 0.3, 0.4), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    lut = o.kernel_lut()
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], h, (m, q), M, sf, R, o.MODE_WEIGHTED, lut)

    cm_lut = torch.from_numpy(luts.colormap_table_1d("viridis", 1000)).cuda()
    params = N.ColormapParams(vmin=-1.0, vmax=1.0, density_vmin=0, density_vmax=1, window_aspect_ratio=1.0, gamma=1.0,
                              kind=N.CMAP_WEIGHTED, log_scale=1)
    want_rgba = o.to_unorm8(o.colormap_scalar(ref, {"vmin": np.float32(-1), "vmax": np.float32(1)},
                                              luts.colormap_table_1d("viridis", 1000), True, True))
    results = {}
    for method in ("p2p", "nccl"):
        sh = D.ShardedSplat(R, 2, reduce=method)
        assert sh.method == method, (sh.method, getattr(sh, "_fallback_reason", ""))
        sh.engine.set_kernel_lut(lut)
        sh.engine.set_camera(M, sf)
        dev = [torch.from_numpy(np.ascontiguousarray(a[mine])).cuda() for a in (pos[:, 0], pos[:, 1], pos[:, 2], h, m, q)]
        sh.engine.set_particles(*dev[:4]); sh.engine.set_weights(dev[4], dev[5])
        for frame in range(3):                          # repeated frames exercise the barriers / image reuse
            sh.splat(N.MODE_WEIGHTED, [(0, len(mine) // 2), (len(mine) // 2, len(mine) - len(mine) // 2)])
            out = sh.present(params, cm_lut)
        torch.cuda.synchronize()
        total = sh.reduced_image().cpu().numpy().astype(np.float64)
        big = ref[..., 0] > 1e-6 * ref[..., 0].max()
        for c in range(2):
            rel = np.abs(total[..., c][big] - ref[..., c][big]) / np.abs(ref[..., c][big])
            assert rel.max() <= 1e-4, (method, c, rel.max())
        if rank == 0:
            got = out.cpu().numpy()
            d = np.abs(got.astype(int) - want_rgba.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 5e-3, (method, d.max(), (d > 0).mean())
            results[method] = got
        sh.close()
    if rank == 0:
        assert np.abs(results["p2p"].astype(int) - results["nccl"].astype(int)).max() <= 1
    visualizer_check(rank, world)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
