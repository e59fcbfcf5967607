"""CPU checks of the drop-in boundary: the C-ABI shared library loads, exports every symbol
include/values_b200.h declares (and nothing is declared twice), and the compute entry points
fail loudly -- never silently fall back -- when no CUDA device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "values_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)           # drop comments
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)            # drop preprocessor lines
    names = re.findall(r"\b(values_[a-z0-9_]+)\s*\(", text)
    return names


def test_header_declares_each_function_once():
    names = declared_functions()
    assert len(names) >= 12
    assert len(names) == len(set(names)), sorted(n for n in names if names.count(n) > 1)


def test_library_exports_every_declared_symbol():
    import values_b200  # noqa: F401  (raises ValuesExtensionMissing if the .so is not built)
    from values_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/values_b200.h but not exported: {missing}"
    assert set(_lib.EXPORTED) == set(declared_functions())   # the ctypes table binds all of them
    assert lib.values_abi_version() == _lib.ABI_VERSION
    ver = int(re.search(r"#define\s+VALUES_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert ver == _lib.ABI_VERSION


def test_workspace_queries_are_pure_host_functions():
    from values_b200 import _lib

    assert _lib.lib.values_uncertainty_workspace_bytes(2, 4096, _lib.F32) > 0
    assert _lib.lib.values_uncertainty_workspace_bytes(0, 4096, _lib.F32) == 0
    sh, pa = _lib.i64x3([64, 64, 64]), _lib.i64x3([10, 10, 10])
    assert _lib.lib.values_patch_max_workspace_bytes(3, sh, pa, 0) > 0
    assert _lib.lib.values_patch_max_workspace_bytes(3, _lib.i64x3([8, 8, 8]), pa, 0) == 0  # patch > image
    assert _lib.lib.values_patch_max_workspace_bytes(3, sh, pa, 1) == 0  # unknown implementation
    assert _lib.lib.values_map_reduce_workspace_bytes(3, 1000) > 0
    # maps whose innermost extent is not a multiple of 4 get room for the pitched scratch copy the strip
    # filter runs on (M x D0 x D1 x round_up(D2, 4) floats); aligned shapes and the exact path do not
    aligned = _lib.lib.values_patch_max_workspace_bytes(3, _lib.i64x3([40, 40, 64]), pa, 0)
    odd = _lib.lib.values_patch_max_workspace_bytes(3, _lib.i64x3([40, 40, 63]), pa, 0)
    assert odd >= aligned - 4096 + 3 * 40 * 40 * 64 * 4 and aligned < 3 * 40 * 40 * 64 * 4
    assert _lib.lib.values_patch_max_workspace_bytes(3, _lib.i64x3([40, 40, 63]), pa, 5) < 3 * 40 * 40 * 64 * 4


def test_map_set_tables():
    """The *_set entry points take HOST arrays of device pointers and counts (include/values_b200.h)."""
    from values_b200 import _lib

    maps = [torch.zeros(5), torch.zeros(0), torch.zeros(7, 3)]      # CPU tensors: only the table is built here
    ptrs, counts, n = _lib.map_set(maps)
    assert n == 3 and list(counts) == [5, 0, 21]
    assert ptrs[0] == maps[0].data_ptr() and ptrs[2] == maps[2].data_ptr()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    """Without a device the reference-facing entry points raise; nothing computes on the host."""
    import numpy as np

    import values_b200 as vb

    x = torch.softmax(torch.randn(3, 2, 4, 4, 4), dim=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vb.calculate_uncertainty(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vb.calculate_one_minus_msr(x[0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vb.image_level_aggregation(np.ones((4, 4)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vb.patch_level_aggregation(np.ones((12, 12)), 10)
    with pytest.raises(RuntimeError):
        vb.uncertainty_fused(x.unsqueeze(0))


def test_invalid_arguments_are_rejected_before_any_launch():
    from values_b200 import _lib

    rc = _lib.lib.values_uncertainty_fused(None, _lib.F32, 1, 0, 2, 16, 0, 0, 0, None, None, None, 0,
                                           None, None, None, 3, None, None, 0, 0, 0, None)
    assert rc == _lib.ERR_INVALID_ARG and b"bad sizes" in _lib.lib.values_last_error()
    rc = _lib.lib.values_uncertainty_fused(None, 7, 1, 2, 2, 16, 32, 16, 16, None, None, None, 0,
                                           None, None, None, 3, None, None, 0, 0, 0, None)
    assert rc == _lib.ERR_INVALID_ARG   # NULL stack / unknown dtype
    sh, pa = _lib.i64x3([8, 8, 8]), _lib.i64x3([10, 10, 10])
    rc = _lib.lib.values_patch_max(None, _lib.F32, 1, 512, sh, pa, 0, 1e-5, 1e-8, None, 1, None, _lib.I64, 3, None, 0, 0, None)
    assert rc == _lib.ERR_INVALID_ARG and b"valid" in _lib.lib.values_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
