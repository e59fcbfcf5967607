"""Threshold finding on B200: drop-in mirrors of
evaluation/uncertainty_aggregation/find_threshold.py over the K4 statistics kernels.

    calculate_foreground_quantile_image(image)                       :11-13
    get_foreground_quantile(exp_dataloader)                          :16-28
    save_foreground_quantiles(results_dict, save_path)               :31-40
    calculate_threshold_image(quantile_path, image, method)          :63-68
    find_threshold(results_dict, quantile_path, save_path)           :71-117

The threshold is `np.quantile` (linear interpolation) over EVERY voxel of every validation
map of a prediction model -- 10^7..10^9 values.  The reference stacks them into one host
array and lets numpy partition it; here each map stays where it is (a list of CUDA tensors,
sharded over ranks if desired) and the two order statistics the interpolation needs are found
by a most-significant-digit radix select: per digit one histogram sweep over the maps
(values_radix_histogram accumulates), one all-reduce of 2048 counters when distributed, one
2048-entry read-back.  The result is exact (bit-identical to numpy on the same values).
"""
from __future__ import annotations

import json
import os
from itertools import chain
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

_DIGITS = {32: (11, 11, 10), 64: (11, 11, 11, 11, 11, 9)}
_NP_DTYPE = {torch.float32: np.float32, torch.float64: np.float64}


# ------------------------------------------------------------------ device-side primitives
def _as_cuda(x, device: torch.device, float_only: bool = False) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        if x.dtype == np.bool_:
            x = x.astype(np.uint8)
        elif float_only and x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        elif x.dtype in (np.int8, np.int16, np.uint16):
            x = x.astype(np.int32)
        elif x.dtype in (np.uint32, np.uint64):
            x = x.astype(np.int64)
        elif x.dtype == np.float16:
            x = x.astype(np.float32)
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected a numpy array or torch tensor, got {type(x)}")
    if x.dtype == torch.bool:
        x = x.to(torch.uint8)
    if float_only and x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float64)
    if x.device.type != "cuda":
        x = x.to(device, non_blocking=True)
    return x.contiguous()


def _any_dtype_code(dt: torch.dtype) -> int:
    if dt in (torch.float32, torch.float64):
        return _lib.dtype_code(dt)
    return _lib.label_dtype_code(dt)


def count_nonzero(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (CUDA; uint8 / int32 / int64 / float32 / float64) -> int64 device scalar; with `out`
    the count is ADDED to out[0] (no sync)."""
    if x.device.type != "cuda":
        raise RuntimeError("count_nonzero expects a CUDA tensor (no CPU fallback)")
    x = x.contiguous()
    if out is None:
        out = torch.zeros(1, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib.values_count_nonzero(x.data_ptr(), _any_dtype_code(x.dtype), x.numel(),
                                           out.data_ptr(), _lib.stream_ptr(x.device))
    _lib.check(rc)
    return out


def _key_to_value(key: int, bits: int):
    top = 1 << (bits - 1)
    raw = (key & (top - 1)) if key & top else (~key) & ((1 << bits) - 1)
    if bits == 32:
        return np.array([raw], dtype=np.uint32).view(np.float32)[0]
    return np.array([raw], dtype=np.uint64).view(np.float64)[0]


def _all_reduce(t: torch.Tensor, op, group) -> None:
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=op, group=group)


def order_statistic(maps: Sequence[torch.Tensor], k: int, distributed: bool = False, group=None
                    ) -> Tuple[int, int, int, int]:
    """Key of the k-th smallest value (0-based, NaN last) over all elements of `maps` (CUDA
    tensors of one float dtype; with distributed=True over the maps of every rank).
    Returns (key, count_lt, count_eq, bits)."""
    dt = maps[0].dtype
    bits = 32 if dt == torch.float32 else 64
    dev = maps[0].device
    code = _lib.dtype_code(dt)
    hist = torch.zeros(1 << 11, dtype=torch.int64, device=dev)
    # the digit walk stays on the device: {prefix, prefix_bits, rank, count_eq, count of the top bucket of the
    # first digit}; ONE readback at the end (with distributed=True the histogram all-reduce is on-stream too)
    state = torch.tensor([0, 0, int(k), 0, 0], dtype=torch.int64, device=dev)
    ptrs, counts, n_maps = _lib.map_set(maps)
    with torch.cuda.device(dev):
        for db in _DIGITS[bits]:
            # one launch per digit and 96 maps (a launch per map made the walk launch-bound)
            _lib.check(_lib.lib.values_radix_histogram_set(ptrs, counts, n_maps, code, 0, 0, state.data_ptr(), db,
                                                           hist.data_ptr(), _lib.stream_ptr(dev)))
            if distributed:
                _all_reduce(hist, dist.ReduceOp.SUM, group)
            _lib.check(_lib.lib.values_radix_select(hist.data_ptr(), db, state.data_ptr(), _lib.stream_ptr(dev)))
    prefix, prefix_bits, rank_left, count_eq, top0 = (int(v) for v in state.cpu().tolist())
    if prefix_bits != bits:
        raise IndexError(f"order_statistic: rank {k} out of range")
    order_statistic.last_top_bucket = top0      # fp32: number of NaNs (their key is the only one in that bucket)
    return prefix & ((1 << 64) - 1), int(k) - rank_left, count_eq, bits


def _min_key_above(maps: Sequence[torch.Tensor], key: int, bits: int, distributed: bool, group) -> int:
    dev = maps[0].device
    out = torch.full((1,), -1, dtype=torch.int64, device=dev)   # 0xffff... as unsigned
    code = _lib.dtype_code(maps[0].dtype)
    ptrs, counts, n_maps = _lib.map_set(maps)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.values_min_key_above_set(ptrs, counts, n_maps, code, key, out.data_ptr(),
                                                     _lib.stream_ptr(dev)))
    if distributed:   # unsigned min through a signed all-reduce: flip the top bit
        out ^= torch.iinfo(torch.int64).min
        _all_reduce(out, dist.ReduceOp.MIN, group)
        out ^= torch.iinfo(torch.int64).min
    return int(out.item()) & ((1 << 64) - 1)


def _virtual_index(n: int, q, np_dtype):
    """numpy's linear-method virtual index, in numpy's own arithmetic: since numpy 2.0 a Python
    float q takes the array's dtype (so fp32 maps index in fp32), before that it stays fp64."""
    if isinstance(q, (int, float)) and np.lib.NumpyVersion(np.__version__) >= "2.0.0":
        qa = np.asanyarray(q, dtype=np_dtype)
    else:
        qa = np.asanyarray(q)
    if qa.ndim != 0:
        raise ValueError("quantile: q must be a scalar")
    if not (0.0 <= float(qa) <= 1.0):
        raise ValueError("Quantiles must be in the range [0, 1]")
    return np.asanyarray((n - 1) * qa)   # numpy's 'linear' get_virtual_index


def _lerp(a, b, t):
    """numpy.lib._function_base_impl._lerp for 0-d operands."""
    a, b = np.asanyarray(a), np.asanyarray(b)
    diff = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff * t))
    if t >= 0.5:
        out = np.asanyarray(np.subtract(b, diff * (1 - t))).astype(out.dtype)
    return out[()]


def quantile(maps: Union[torch.Tensor, np.ndarray, Sequence], q: float, distributed: bool = False,
             group=None):
    """Exact `np.quantile(np.array(maps), q)` (method="linear") with the maps resident on the GPU.

    maps: one array / tensor or a sequence of them (float32 or float64, one dtype); numpy inputs
    are uploaded.  distributed=True: the quantile over the maps of EVERY rank of `group`
    (histograms all-reduced over NCCL / gloo); every rank returns the same value.
    Returns a numpy scalar of the maps' dtype, NaN if any value is NaN (as numpy)."""
    dev = _lib.require_cuda()
    if isinstance(maps, (torch.Tensor, np.ndarray)):
        maps = [maps]
    maps = [_as_cuda(m, dev, float_only=True) for m in maps]
    if not maps:
        raise ValueError("quantile: no maps")
    dt = maps[0].dtype
    if any(m.dtype != dt for m in maps):
        maps = [m.to(torch.float64) for m in maps]
        dt = torch.float64
    np_dtype = _NP_DTYPE[dt]
    n = sum(m.numel() for m in maps)
    if distributed:
        n_t = torch.tensor([n], dtype=torch.int64, device=dev)
        _all_reduce(n_t, dist.ReduceOp.SUM, group)
        n = int(n_t.item())
    if n == 0:
        raise IndexError("quantile of an empty set of maps")
    virtual = _virtual_index(n, q, np_dtype)
    prev = np.floor(virtual)
    nxt = prev + 1
    if virtual >= n - 1:
        prev = nxt = n - 1
    if virtual < 0:
        prev = nxt = 0
    k_lo, k_hi = int(prev), int(nxt)
    key_lo, below, eq, bits = order_statistic(maps, k_lo, distributed, group)
    # NaN sorts last: any NaN <=> the largest key is all ones
    nan_key = (1 << bits) - 1
    if key_lo == nan_key:
        return np_dtype(np.nan)
    if bits == 32:
        # fp32: the top bucket of the first digit (sign flipped, exponent 255, mantissa 11...) holds NaNs only
        if order_statistic.last_top_bucket > 0:
            return np_dtype(np.nan)
    else:   # fp64: that bucket also holds finite values >= 2^1023; count the all-ones key itself
        last_db = _DIGITS[bits][-1]
        hist = torch.zeros(1 << last_db, dtype=torch.int64, device=dev)
        ptrs, counts, n_maps = _lib.map_set(maps)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.values_radix_histogram_set(
                ptrs, counts, n_maps, _lib.dtype_code(dt), nan_key >> last_db, bits - last_db, None,
                last_db, hist.data_ptr(), _lib.stream_ptr(dev)))
        if distributed:
            _all_reduce(hist, dist.ReduceOp.SUM, group)
        if int(hist[-1].item()) > 0:
            return np_dtype(np.nan)
    a_lo = _key_to_value(key_lo, bits)
    if k_hi < below + eq:
        a_hi = a_lo
    else:
        a_hi = _key_to_value(_min_key_above(maps, key_lo, bits, distributed, group), bits)
    gamma = np.asanyarray(virtual - np.asanyarray(prev), dtype=virtual.dtype)
    return _lerp(a_lo, a_hi, gamma)


# ------------------------------------------------------------------ reference entry points
def calculate_foreground_quantile_image(image) -> float:
    """Drop-in for find_threshold.py:11-13: 1 - count_nonzero(image) / image.size."""
    dev = _lib.require_cuda()
    img = _as_cuda(image, dev)
    foreground = int(count_nonzero(img).item())
    return 1 - (foreground / img.numel())


def get_foreground_quantile(exp_dataloader) -> Dict:
    """Drop-in for find_threshold.py:16-28 (the data loader's own getters do the file I/O)."""
    dev = _lib.require_cuda()
    quantile_dict = {exp_dataloader.exp_version.pred_model: {}}
    all_quantiles: List[float] = []
    for image_id in exp_dataloader.image_ids:
        for pred_seg in exp_dataloader.get_pred_segs(image_id):
            all_quantiles.append(calculate_foreground_quantile_image(pred_seg))
    quantile_dict[exp_dataloader.exp_version.pred_model][
        exp_dataloader.exp_version.version_name] = all_quantiles
    return quantile_dict


def save_foreground_quantiles(results_dict: Dict, save_path) -> None:
    """Drop-in for find_threshold.py:31-40 (host-side json)."""
    methods_results_dict = {}
    for method, versions in results_dict.items():
        methods_results_dict[method] = float(np.mean(list(chain.from_iterable(versions.values()))))
    if not os.path.isfile(save_path):
        save_path = Path(save_path) / "quantile_analysis.json"
    with open(save_path, "w") as f:
        json.dump(methods_results_dict, f, indent=2)


def calculate_threshold_image(quantile_path, image, method: str, distributed: bool = False,
                              group=None):
    """Drop-in for find_threshold.py:63-68: np.quantile(image, quantiles[method]).  `image` may be
    one array or a sequence of maps (the reference passes np.array(list of maps))."""
    with open(quantile_path) as f:
        all_quantiles = json.load(f)
    return quantile(image, all_quantiles[method], distributed=distributed, group=group)


def find_threshold(results_dict: Dict, quantile_path, save_path, load_fn=None) -> Dict:
    """Drop-in for find_threshold.py:71-117, to its evident intent: the reference calls
    calculate_threshold_image with two arguments (:94) although it takes three; the quantile file
    is the missing first argument.  `load_fn(path) -> ndarray or CUDA tensor` defaults to
    values_b200.formats.load_to_device (medpy.io.load's result, on the device).
    Returns the dict it writes to threshold_analysis.json."""
    if load_fn is None:
        from .formats import load_to_device

        def load_fn(path):
            return load_to_device(path)[0]

    if not os.path.isfile(quantile_path):
        quantile_path = Path(quantile_path) / "quantile_analysis.json"
    if not os.path.isfile(save_path):
        save_path = Path(save_path) / "threshold_analysis.json"
    dev = _lib.require_cuda()
    pred_model_path_dict: Dict[str, Dict[str, list]] = {}
    for pred_model, versions in results_dict.items():
        pred_model_path_dict[pred_model] = {}
        for version, uncs in versions.items():
            for unc, paths in uncs.items():
                pred_model_path_dict[pred_model].setdefault(unc, []).extend(paths)
    threshold_dict: Dict[str, Dict[str, float]] = {}
    for pred_model, uncs in pred_model_path_dict.items():
        threshold_dict[pred_model] = {}
        for unc, paths in uncs.items():
            unc_images = [_as_cuda(load_fn(path), dev, float_only=True) for path in paths]
            threshold = calculate_threshold_image(quantile_path, unc_images, pred_model)
            threshold_dict[pred_model][f"Mean {unc.split('_')[0]} threshold"] = float(threshold)
    all_aleatoric, all_epistemic, all_predictive = [], [], []
    for key, value in threshold_dict.items():
        if key != "Softmax":
            all_aleatoric.append(value["Mean aleatoric threshold"])
            all_epistemic.append(value["Mean epistemic threshold"])
        all_predictive.append(value["Mean predictive threshold"])
    threshold_dict["Mean"] = {
        "Mean aleatoric threshold": float(np.mean(all_aleatoric)),
        "Mean epistemic threshold": float(np.mean(all_epistemic)),
        "Mean predictive threshold": float(np.mean(all_predictive)),
    }
    with open(save_path, "w") as f:
        json.dump(threshold_dict, f, indent=2)
    return threshold_dict
