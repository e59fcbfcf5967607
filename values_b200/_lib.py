"""ctypes binding of libvalues_b200.so (the C-ABI declared in include/values_b200.h).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is present
when a compute entry point is called, this module raises -- loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# VALUES_B200_LIB: an alternative build of the same library (the compute-sanitizer build, values_b200/build.py)
LIB_PATH = os.environ.get("VALUES_B200_LIB") or os.path.join(_PKG, "lib", "libvalues_b200.so")

F32, F64, BF16, U8, I32, I64 = 0, 1, 2, 3, 4, 5
ABI_VERSION = 8  # include/values_b200.h VALUES_ABI_VERSION
_DTYPES = {torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16}
_LABEL_DTYPES = {torch.uint8: U8, torch.int32: I32, torch.int64: I64}

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4


class ValuesExtensionMissing(ImportError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ValuesExtensionMissing(
            f"{LIB_PATH} not found. The values_b200 CUDA extension is required (there is no CPU "
            "fallback); build it with `python -m values_b200.build` or `__graft_entry__.build()`."
        )
    lib = C.CDLL(LIB_PATH)
    i64, vp, sz, dbl = C.c_int64, C.c_void_p, C.c_size_t, C.c_double
    pi64, pdbl = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    sig = {
        "values_abi_version": (C.c_int, []),
        "values_last_error": (C.c_char_p, []),
        "values_launch_count": (i64, []),
        "values_uncertainty_workspace_bytes": (sz, [i64, i64, C.c_int]),
        "values_uncertainty_fused": (C.c_int, [vp, C.c_int, i64, i64, i64, i64, i64, i64, i64,
                                               vp, vp, vp, i64, vp, vp, vp, i64, pdbl, vp, sz, C.c_int, C.c_int, vp]),
        "values_one_minus_msr": (C.c_int, [vp, C.c_int, i64, i64, i64, i64, i64, vp, vp]),
        "values_map_reduce_workspace_bytes": (sz, [i64, i64]),
        "values_map_reduce": (C.c_int, [vp, C.c_int, i64, i64, i64, pdbl, C.c_int, vp, vp, sz, vp]),
        "values_patch_max_workspace_bytes": (sz, [i64, pi64, pi64, C.c_int]),
        "values_patch_max": (C.c_int, [vp, C.c_int, i64, i64, pi64, pi64, C.c_int, dbl, dbl,
                                       vp, i64, vp, C.c_int, i64, vp, sz, C.c_int, vp]),
        "values_stitch_accumulate": (C.c_int, [vp, C.c_int, i64, i64, vp, vp, i64, i64, i64,
                                               pi64, pi64, vp, C.c_int, vp, C.c_int, C.c_int, vp]),
        "values_stitch_accumulate_weighted": (C.c_int, [vp, C.c_int, i64, i64, vp, vp, vp, i64, i64, i64,
                                                        pi64, pi64, vp, C.c_int, vp, C.c_int, C.c_int, vp]),
        "values_stitch_accumulate_separable": (C.c_int, [vp, C.c_int, i64, i64, vp, vp, vp, vp, vp, i64, i64, i64,
                                                         pi64, pi64, vp, C.c_int, vp, C.c_int, C.c_int, vp]),
        "values_normalize_maps": (C.c_int, [vp, C.c_int, i64, i64, i64, vp, dbl, vp, vp]),
        "values_seg_loss_workspace_bytes": (sz, [i64, C.c_int, i64]),
        "values_seg_loss_terms": (C.c_int, [vp, C.c_int, i64, vp, C.c_int, i64, i64, C.c_int, i64, vp, vp, sz, vp]),
        "values_count_nonzero": (C.c_int, [vp, C.c_int, i64, vp, vp]),
        "values_radix_histogram": (C.c_int, [vp, C.c_int, i64, C.c_uint64, C.c_int, C.c_int, vp, vp]),
        "values_min_key_above": (C.c_int, [vp, C.c_int, i64, C.c_uint64, vp, vp]),
        "values_radix_histogram_dev": (C.c_int, [vp, C.c_int, i64, vp, C.c_int, vp, vp]),
        "values_radix_select": (C.c_int, [vp, C.c_int, vp, vp]),
        "values_radix_histogram_set": (C.c_int, [C.POINTER(vp), pi64, i64, C.c_int, C.c_uint64, C.c_int, vp,
                                                 C.c_int, vp, vp]),
        "values_min_key_above_set": (C.c_int, [C.POINTER(vp), pi64, i64, C.c_int, C.c_uint64, vp, vp]),
        "values_pair_moments_workspace_bytes": (sz, [i64, i64]),
        "values_pair_moments": (C.c_int, [vp, C.c_int, i64, vp, C.c_int, i64, i64, i64, vp, vp, vp, sz, vp]),
        "values_calib_bins_workspace_bytes": (sz, [i64]),
        "values_calib_bins": (C.c_int, [vp, C.c_int, vp, C.c_int, i64, pdbl, C.c_int, vp, vp, sz, vp]),
        "values_calib_bins_fused": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, i64, i64, dbl, dbl, C.c_int,
                                              i64, pdbl, C.c_int, vp, vp, sz, vp]),
        "values_confusion_counts": (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, i64, C.c_int, vp, vp]),
        "values_reverse_axes": (C.c_int, [vp, vp, C.c_int, i64, i64, i64, vp]),
        "values_patch_filter_err_coef": (dbl, [C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
        fn.restype, fn.argtypes = res, args
    if lib.values_abi_version() != ABI_VERSION:
        raise ValuesExtensionMissing("libvalues_b200.so ABI version mismatch; rebuild it")
    return lib


lib = _load()
EXPORTED = [
    "values_abi_version", "values_last_error", "values_launch_count",
    "values_uncertainty_workspace_bytes", "values_uncertainty_fused", "values_one_minus_msr",
    "values_map_reduce_workspace_bytes", "values_map_reduce",
    "values_patch_max_workspace_bytes", "values_patch_max", "values_stitch_accumulate",
    "values_stitch_accumulate_weighted", "values_stitch_accumulate_separable",
    "values_normalize_maps", "values_count_nonzero", "values_radix_histogram",
    "values_min_key_above", "values_radix_histogram_dev", "values_radix_select",
    "values_radix_histogram_set", "values_min_key_above_set",
    "values_pair_moments_workspace_bytes", "values_pair_moments",
    "values_calib_bins_workspace_bytes", "values_calib_bins", "values_calib_bins_fused",
    "values_confusion_counts", "values_reverse_axes", "values_patch_filter_err_coef",
    "values_seg_loss_workspace_bytes", "values_seg_loss_terms",
]


def map_set(maps):
    """(pointer array, count array, n) of a sequence of CUDA tensors, for the *_set entry points."""
    n = len(maps)
    ptrs = (C.c_void_p * n)(*[m.data_ptr() for m in maps])
    counts = (C.c_int64 * n)(*[m.numel() for m in maps])
    return ptrs, counts, n


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = (lib.values_last_error() or b"").decode()
    if rc == ERR_INVALID_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"values_b200 error {rc}: {msg}")


def launch_count() -> int:
    return int(lib.values_launch_count())


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"values_b200: unsupported dtype {dt} (float32, float64, bfloat16)") from None


def label_dtype_code(dt: torch.dtype) -> int:
    try:
        return _LABEL_DTYPES[dt]
    except KeyError:
        raise TypeError(f"values_b200: unsupported label dtype {dt} (uint8, int32, int64)") from None


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "values_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback. "
            "For a CPU oracle see oracle/values_oracle.py (test infrastructure only)."
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def i64x3(vals):
    return (C.c_int64 * 3)(*[int(v) for v in vals])


def dbl_array(vals):
    vals = [float(v) for v in vals]
    return (C.c_double * len(vals))(*vals)
