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
        print("MULTI_GPU_CHECK_OK world", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
