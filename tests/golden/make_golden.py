"""Generate tests/golden/*.npz from the read-only reference checkout.

Run in the build container only (``python tests/golden/make_golden.py``); /root/reference does not exist on the
GPU box, which is why the vectors are committed.  Two kinds of fixtures:

1. ``reference_goldens.npz`` -- the known-answer arrays the reference's own test-suite pins for the SPH projection
   path (tests/test_render_output.py of the reference).  They are extracted from the test source with ``ast`` (the
   tests themselves cannot run here: wgpu / pynbody / matplotlib are absent).
2. ``scheduling_goldens.npz`` -- outputs of the reference's own pure-Python modules (config, cell_layout,
   progressive_render, split-buffer address maths, TestDataLoader) executed here with ``wgpu``/``pynbody`` stubbed
   in ``sys.modules``.  Nothing is copied from the reference: it is imported and run.
"""
from __future__ import annotations

import ast
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).parent


def _literal_arrays(func_node):
    """Map of assigned name -> numpy array for list / np.array(list) literals assigned inside a test function."""
    found = {}
    for node in ast.walk(func_node):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            val = node.value
            if isinstance(val, ast.Call) and val.args and isinstance(val.args[0], ast.List):
                val = val.args[0]
            if isinstance(val, ast.List):
                try:
                    arr = np.array(ast.literal_eval(val), dtype=np.float64)
                except Exception:
                    continue
                if arr.size >= 40:
                    found[node.targets[0].id] = arr
        # known answer passed inline to assert_allclose (test_particle_pos_smooth)
        if isinstance(node, ast.Call) and getattr(node.func, 'attr', '') == 'assert_allclose' and len(node.args) >= 2:
            if isinstance(node.args[1], ast.List):
                try:
                    arr = np.array(ast.literal_eval(node.args[1]), dtype=np.float64)
                    found['_inline'] = arr
                except Exception:
                    pass
    return found


def reference_goldens():
    src = (REF / "tests" / "test_render_output.py").read_text()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name.startswith("test_"):
            for name, arr in _literal_arrays(node).items():
                out[f"{node.name}__{name}"] = arr
    wanted = ["test_render__reference_result", "test_hdr_rgb_render__result_ref", "test_particle_pos_smooth___inline",
              "test_sph_weighted_output__expect", "test_sph_output__expect", "test_periodic_sph_output__expect",
              "test_depth_output__expect", "test_bivariate_render__expect_den", "test_bivariate_render__expect_qty",
              "test_bivariate_render__expect_rgba"]
    missing = [w for w in wanted if w not in out]
    assert not missing, missing
    np.savez_compressed(OUT / "reference_goldens.npz", **{k: out[k] for k in wanted})
    print("reference_goldens.npz:", {k: out[k].shape for k in wanted})


def surface_goldens():
    """Known answers of the surface render mode: tests/test_render_output.py::test_surface_render (:448-556) and
    tests/test_smooth.py::test_smoothing_operation (bilateral filter, atol 1e-6) -> surface_goldens.npz"""
    out = {}
    for rel, fn, names in (("test_render_output.py", "test_surface_render",
                            ["quantity_expectation", "depth_expectation", "presentation_expectation"]),
                           ("test_smooth.py", "test_smoothing_operation", ["expected_global_samples", "expected_edge_check"])):
        tree = ast.parse((REF / "tests" / rel).read_text())
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fn)
        found = _literal_arrays(node)
        for name in names:
            out[f"{fn}__{name}"] = found[name]
    np.savez_compressed(OUT / "surface_goldens.npz", **out)
    print("surface_goldens.npz:", {k: v.shape for k, v in out.items()})


def load_reference_modules():
    wgpu = types.ModuleType("wgpu"); wgpu.GPUDevice = object
    pynbody = types.ModuleType("pynbody"); pynbody.snapshot = types.ModuleType("pynbody.snapshot")
    pynbody.snapshot.SimSnap = object
    filt = types.ModuleType("pynbody.filt"); filt.Filter = object; filt.geometry_selection = None
    pynbody.filt = filt
    sys.modules.update({"wgpu": wgpu, "pynbody": pynbody, "pynbody.snapshot": pynbody.snapshot, "pynbody.filt": filt})
    pkg = types.ModuleType("topsy"); pkg.__path__ = [str(REF / "src" / "topsy")]
    sys.modules["topsy"] = pkg
    mods = {}
    for name in ["config", "drawreason", "performance", "cell_layout", "progressive_render", "loader"]:
        spec = importlib.util.spec_from_file_location(f"topsy.{name}", REF / "src" / "topsy" / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"topsy.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods[name] = mod
    return mods


def scheduling_goldens():
    m = load_reference_modules()
    out = {}
    # --- fixture data (loader.py:241-332) -------------------------------------------------------------------
    for n in (1, 1000, 4000):
        ld = m["loader"].TestDataLoader(None, n)
        out[f"gmm{n}_pos_smooth"] = ld.get_pos_smooth()
        out[f"gmm{n}_mass"] = ld.get_mass()
        out[f"gmm{n}_qty"] = ld.get_named_quantity("test-quantity")
        out[f"gmm{n}_rgb"] = ld.get_rgb_masses()
    ldc = m["loader"].TestDataLoader(None, 1000, with_cells=True)
    out["gmm1000_cells_pos_smooth"] = ldc.get_pos_smooth()
    out["gmm1000_cells_lengths"] = ldc._cell_layout._lengths
    out["gmm1000_cells_offsets"] = ldc._cell_layout._offsets
    out["gmm1000_cells_centres"] = ldc._cell_layout._centres

    # --- cell layout (cell_layout.py:63-113): float64 uniform, float32 gaussian -------------------------------
    rs = np.random.RandomState(42)
    pos64 = rs.uniform(-1.0, 1.0, (20000, 3))
    cl, order = m["cell_layout"].CellLayout.from_positions(pos64, -1.0, 1.0, 10)
    out["cl64_pos"] = pos64; out["cl64_order"] = order; out["cl64_lengths"] = cl._lengths
    out["cl64_offsets"] = cl._offsets; out["cl64_centres"] = cl._centres
    out["cl64_sphere"] = cl.cells_in_sphere((0.1, -0.2, 0.3), 0.35)
    pos32 = (rs.normal(size=(30000, 3)) * [20, 5, 1]).astype(np.float32)
    bmin = pos32.min(); bmax = pos32.max(); rng = bmax - bmin
    bmin -= m["config"].CELL_LAYOUT_FRACTIONAL_PADDING * rng; bmax += m["config"].CELL_LAYOUT_FRACTIONAL_PADDING * rng
    cl32, order32 = m["cell_layout"].CellLayout.from_positions(pos32, bmin, bmax, m["config"].DEFAULT_CELLS_NSIDE)
    out["cl32_pos"] = pos32; out["cl32_box"] = np.array([bmin, bmax]); out["cl32_order"] = order32
    out["cl32_lengths"] = cl32._lengths; out["cl32_offsets"] = cl32._offsets; out["cl32_centres"] = cl32._centres
    out["cl32_sphere"] = cl32.cells_in_sphere((1.0, 2.0, 0.5), 12.0)

    # --- progression with cells (progressive_render.py:139-215) ---------------------------------------------
    DR = m["drawreason"].DrawReason
    rp = m["progressive_render"].RenderProgressionWithCells(cl, len(pos64), 100)
    out["rp_phase"] = rp._cell_phase_shifts
    blocks = []
    rp.start_frame(DR.CHANGE)
    t = 0.0
    frames = 0
    while True:
        blk = rp.get_block(0.0)
        blocks.append(np.stack([np.asarray(blk[0]), np.asarray(blk[1])]))
        rp.end_block(0.0001)
        rp.end_frame_get_scalefactor()
        frames += 1
        if rp.needs_refine() and frames < 6:
            rp.start_frame(DR.REFINE)
        else:
            break
    for i, b in enumerate(blocks):
        out[f"rp_block{i}"] = b
    out["rp_nblocks"] = np.array(len(blocks))
    rp2 = m["progressive_render"].RenderProgressionWithCells(cl, len(pos64), 100)
    rp2.select_sphere((0.1, -0.2, 0.3), 0.35)
    out["rp_sphere_fraction"] = np.array(rp2.get_fraction_volume_selected())
    rp2.start_frame(DR.EXPORT)
    blk = rp2.get_block(0.0)
    out["rp_sphere_export_block"] = np.stack([np.asarray(blk[0]), np.asarray(blk[1])])
    s, l = rp2._map_logical_range_to_actual_ranges(1234, 4321)
    out["rp_sphere_map_1234_4321"] = np.stack([s, l])

    # --- plain progression: adaptive particle-number update (progressive_render.py:88-110) -----------------
    rpp = m["progressive_render"].RenderProgression(10 ** 7)
    trace = []
    for frame_time in [0.01, 0.2, 0.05, 0.033, 0.001, 0.5]:
        rpp.start_frame(DR.CHANGE)
        blk = rpp.get_block(0.0)
        rpp.end_block(frame_time)
        sf = rpp.end_frame_get_scalefactor()
        trace.append([blk[0][0], blk[1][0], sf, rpp._recommended_num_particles_to_render])
    out["rpp_trace"] = np.array(trace, dtype=np.float64)

    np.savez_compressed(OUT / "scheduling_goldens.npz", **out)
    print("scheduling_goldens.npz:", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "surface":      # added later; leaves the older fixture files untouched
        surface_goldens()
    else:
        reference_goldens()
        scheduling_goldens()
        surface_goldens()
