"""The C-ABI library loads on a CPU-only box and exports every symbol include/tsplat.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "tsplat.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsplat_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("tsplat_create", "tsplat_render", "tsplat_colormap", "tsplat_cell_layout", "tsplat_set_particles"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from topsy_b200 import _native
    if not _native.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(str(_native.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/tsplat.h but not exported"
    assert set(_native.EXPORTED_SYMBOLS) == set(declared_symbols())


def test_pure_host_entry_points():
    from topsy_b200 import _native as N
    lib = N.lib()
    assert lib.tsplat_abi_version() == 1
    assert [lib.tsplat_mode_channels(m) for m in range(5)] == [1, 2, 4, 2, 2]
    assert lib.tsplat_mode_channels(9) == -1
    assert lib.tsplat_scratch_bytes(2048, 1 << 20) > (1 << 20) * 32
    assert lib.tsplat_cell_layout_work_bytes(1000, 16) > 0
    # argument validation happens before any CUDA call
    assert lib.tsplat_set_kernel_lut(None, None, 0) == N.ERR_INVALID
    with pytest.raises(ValueError):
        N.check(lib.tsplat_render(None, None, None, 0, 0, 1, None))
