"""C3 aggregation strategies on B200: drop-in mirrors of
evaluation/uncertainty_aggregation/aggregate_uncertainties.py over kernels K2a / K2b.

    patch_level_aggregation(image, patch_size, mean=False, **kwargs)     :13-31
    image_level_aggregation(image, mean=False, **kwargs)                 :34-37
    threshold_aggregation(image, threshold=None, threshold_path=None,
                          pred_model=None, unc_type=None, mean=True)     :40-67
    aggregate_uncertainties(exp_dataloader, aggregations)                :70-96

`image` may be a numpy array (as medpy hands it to the reference) or a torch tensor; CUDA
tensors are used in place.  Batched device-side forms (`patch_max`, `map_reduce`) return
tensors and never synchronise.
"""
from __future__ import annotations

import importlib
import json
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

ISCLOSE_RTOL, ISCLOSE_ATOL = 1e-5, 1e-8  # np.isclose defaults (aggregate_uncertainties.py:20)


def _as_map(image, device: torch.device) -> torch.Tensor:
    """numpy / torch -> contiguous CUDA fp32 or fp64 tensor (other dtypes widen to fp64,
    as numpy/scipy would when combining with their fp64 ones-kernel)."""
    if isinstance(image, np.ndarray):
        if image.dtype not in (np.float32, np.float64):
            image = image.astype(np.float64)
        if image.ndim in (2, 3) and image.flags.f_contiguous and not image.flags.c_contiguous:
            # what medpy.io.load returns (a transposed view of the x-fastest payload): upload it as it
            # lies in memory and reverse the axes on the GPU instead of a strided host copy
            from .formats import reverse_axes

            return reverse_axes(torch.from_numpy(image.T).to(device, non_blocking=True))
        image = torch.from_numpy(np.ascontiguousarray(image))
    if not isinstance(image, torch.Tensor):
        raise TypeError(f"image must be a numpy array or torch tensor, got {type(image)}")
    if image.dtype not in (torch.float32, torch.float64):
        image = image.to(torch.float64)
    if image.device.type != "cuda":
        image = image.to(device, non_blocking=True)
    return image.contiguous()


# ------------------------------------------------------------------ device-side batched ops
def patch_max_workspace_bytes(M: int, spatial: Sequence[int], patch_size, path: int = 0) -> int:
    nd = len(spatial)
    if isinstance(patch_size, (int, np.integer)):
        patch_size = nd * [int(patch_size)]
    return int(_lib.lib.values_patch_max_workspace_bytes(
        M, _lib.i64x3([1] * (3 - nd) + list(spatial)), _lib.i64x3([1] * (3 - nd) + list(patch_size)), int(path)))


def patch_max(maps: torch.Tensor, patch_size, mean: bool = False,
              rtol: float = ISCLOSE_RTOL, atol: float = ISCLOSE_ATOL,
              out_score: Optional[torch.Tensor] = None, out_bbox: Optional[torch.Tensor] = None,
              workspace: Optional[torch.Tensor] = None, path: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """maps [M, *S] (CUDA fp32/fp64, 1 <= len(S) <= 3) -> (max_score fp64 [M], bbox_lo [M, len(S)]).
    out_score fp64 [M] (any stride) / out_bbox [M, 3] int64 or fp64 (rows evenly strided, e.g. columns
    of an fp64 score table) / workspace (uint8) may be preallocated and are written in place.
    path: implementation for this call (include/values_b200.h: 0 automatic, 5 exact march without the
    fp32 filter, 4 fused tile kernel, 2 generic tiled path); all give identical results."""
    if maps.device.type != "cuda":
        raise RuntimeError("patch_max expects a CUDA tensor (no CPU fallback)")
    nd = maps.dim() - 1
    if not 1 <= nd <= 3:
        raise NotImplementedError("patch_level_aggregation supports 1-, 2- and 3-D images")
    if isinstance(patch_size, (int, np.integer)):
        patch_size = nd * [int(patch_size)]
    if len(patch_size) != nd:
        raise ValueError("patch_size must have one entry per image dimension")
    maps = maps.contiguous()
    M = maps.shape[0]
    shape3 = [1] * (3 - nd) + list(maps.shape[1:])
    patch3 = [1] * (3 - nd) + [int(p) for p in patch_size]
    dev = maps.device
    score = out_score if out_score is not None else torch.empty(M, dtype=torch.float64, device=dev)
    bbox = out_bbox if out_bbox is not None else torch.empty((M, 3), dtype=torch.int64, device=dev)
    if (tuple(score.shape) != (M,) or score.dtype != torch.float64 or (M > 1 and score.stride(0) < 1)
            or tuple(bbox.shape) != (M, 3) or bbox.dtype not in (torch.int64, torch.float64)
            or bbox.stride(1) != 1 or (M > 1 and bbox.stride(0) < 3)):
        raise ValueError("out_score must be fp64 [M] and out_bbox int64 / fp64 [M, 3] with contiguous rows")
    sh, pa = _lib.i64x3(shape3), _lib.i64x3(patch3)
    ws_bytes = _lib.lib.values_patch_max_workspace_bytes(M, sh, pa, int(path))
    if workspace is not None and workspace.numel() * workspace.element_size() >= ws_bytes:
        ws = workspace
    else:
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    V = int(np.prod(shape3))
    with torch.cuda.device(dev):
        rc = _lib.lib.values_patch_max(maps.data_ptr(), _lib.dtype_code(maps.dtype), M, V, sh, pa,
                                       int(bool(mean)), float(rtol), float(atol), score.data_ptr(),
                                       score.stride(0) if M > 1 else 1, bbox.data_ptr(),
                                       _lib.I64 if bbox.dtype == torch.int64 else _lib.F64,
                                       bbox.stride(0) if M > 1 else 3, ws.data_ptr(), ws_bytes, int(path),
                                       _lib.stream_ptr(dev))
    _lib.check(rc)
    return score, bbox[:, 3 - nd:]


def map_reduce(maps: torch.Tensor, thresholds: Optional[Sequence[float]] = None) -> torch.Tensor:
    """maps [M, *S] (CUDA fp32/fp64) -> fp64 [M, 3] = {sum, sum over v>=thr, count v>=thr};
    map m uses thresholds[m % len(thresholds)]."""
    if maps.device.type != "cuda":
        raise RuntimeError("map_reduce expects a CUDA tensor (no CPU fallback)")
    maps = maps.contiguous()
    M = maps.shape[0]
    V = int(np.prod(maps.shape[1:])) if maps.dim() > 1 else 1
    dev = maps.device
    out = torch.empty((M, 3), dtype=torch.float64, device=dev)
    ws_bytes = _lib.lib.values_map_reduce_workspace_bytes(M, V)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    thr, n_thr = None, 0
    if thresholds is not None:
        n_thr = len(thresholds)
        thr = _lib.dbl_array(thresholds)
    with torch.cuda.device(dev):
        rc = _lib.lib.values_map_reduce(maps.data_ptr(), _lib.dtype_code(maps.dtype), M, V, V, thr,
                                        n_thr, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        _lib.stream_ptr(dev))
    _lib.check(rc)
    return out


def normalize_maps(maps: torch.Tensor, count: torch.Tensor, clip_min: float = 1.0) -> torch.Tensor:
    """maps [M, *S] / clip(count [*S], 1) -> fp64 [M, *S] (data_carrier_3D.py:215-217, 326-329).
    clip_min=0: divide by a weighted stitch's weight sum, uncovered voxels unscaled."""
    maps = maps.contiguous()
    count = count.to(torch.float64).contiguous()
    M = maps.shape[0]
    V = count.numel()
    if maps[0].numel() != V:
        raise ValueError("count must match the spatial shape of the maps")
    out = torch.empty(maps.shape, dtype=torch.float64, device=maps.device)
    with torch.cuda.device(maps.device):
        rc = _lib.lib.values_normalize_maps(maps.data_ptr(), _lib.dtype_code(maps.dtype), M, V, V,
                                            count.data_ptr(), float(clip_min), out.data_ptr(),
                                            _lib.stream_ptr(maps.device))
    _lib.check(rc)
    return out


# ------------------------------------------------------------------ reference entry points
def patch_level_aggregation(image, patch_size, mean: bool = False, **kwargs) -> Dict:
    """Drop-in for aggregate_uncertainties.py:13-31."""
    dev = _lib.require_cuda()
    img = _as_map(image, dev)
    try:
        # `_k2b_path` (tests only): the K2b implementation for this call, see patch_max
        score, bbox = patch_max(img.unsqueeze(0), patch_size, mean=mean, path=int(kwargs.get("_k2b_path", 0)))
    except ValueError as e:  # image smaller than the patch: scipy's message (probe)
        raise ValueError(str(e)) from None
    lo = bbox[0].tolist()
    if lo and lo[0] < 0:  # NaN map: the reference's np.where(...)[0] raises IndexError
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    if isinstance(patch_size, (int, np.integer)):
        patch_size = len(lo) * [int(patch_size)]
    return {
        "max_score": float(score[0].item()),
        "bounding_box": [(int(i), int(i + patch_size[d])) for d, i in enumerate(lo)],
    }


def image_level_aggregation(image, mean: bool = False, **kwargs):
    """Drop-in for aggregate_uncertainties.py:34-37 (bare float when mean=True)."""
    dev = _lib.require_cuda()
    img = _as_map(image, dev)
    total = float(map_reduce(img.unsqueeze(0))[0, 0].item())
    if mean:
        return float(total / img.numel())
    return {"max_score": total}


def _resolve_threshold(threshold, threshold_path, pred_model, unc_type):
    if threshold is None:  # aggregate_uncertainties.py:48-60
        if threshold_path is None:
            raise Exception("A threshold needs to be provided for threshold aggregation!")
        with open(threshold_path) as f:
            threshold_json = json.load(f)
        if pred_model is None or unc_type is None:
            raise Exception(
                "If you want to load the threshold from a json file, you have to provide the "
                "prediction model and the uncertainty type"
            )
        unc_type_split = unc_type.split("_")[0]
        threshold = threshold_json[pred_model][f"Mean {unc_type_split} threshold"]
    return threshold


def threshold_aggregation(image, threshold=None, threshold_path=None, pred_model=None,
                          unc_type=None, mean: bool = True) -> Dict:
    """Drop-in for aggregate_uncertainties.py:40-67.  `max_score` is a Python float (the
    reference returns numpy scalars, np.float32(0.0) on an empty fp32 selection, which
    json.dumps cannot serialise -- SURVEY.md section 8a9)."""
    threshold = _resolve_threshold(threshold, threshold_path, pred_model, unc_type)
    dev = _lib.require_cuda()
    img = _as_map(image, dev)
    _, s, n = map_reduce(img.unsqueeze(0), [threshold])[0].tolist()
    if mean and n > 0:
        return {"max_score": s / n, "threshold": threshold}
    return {"max_score": s, "threshold": threshold}


# dotted paths the reference's hydra configs use (evaluation/configs/tasks/*.yaml) -> ours
_REF_MODULE = "evaluation.uncertainty_aggregation.aggregate_uncertainties"
TARGETS = {
    f"{_REF_MODULE}.patch_level_aggregation": patch_level_aggregation,
    f"{_REF_MODULE}.image_level_aggregation": image_level_aggregation,
    f"{_REF_MODULE}.threshold_aggregation": threshold_aggregation,
}


def _instantiate(cfg, **kwargs):
    """Minimal stand-in for hydra.utils.instantiate(cfg, **kwargs) for `_target_` callables
    (aggregate_uncertainties.py:81-86).  Reference dotted paths resolve to this module."""
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    fn = TARGETS.get(target)
    if fn is None:
        mod, _, name = target.rpartition(".")
        fn = getattr(importlib.import_module(mod), name)
    cfg.update(kwargs)
    return fn(**cfg)


def aggregate_uncertainties(exp_dataloader, aggregations, load_fn=None, save: bool = True) -> Dict:
    """Drop-in for aggregate_uncertainties.py:70-96: for every uncertainty type and image, run
    every configured aggregation and write `aggregated_<unc>.json`.

    `load_fn(path) -> ndarray or CUDA tensor` defaults to values_b200.formats.load_to_device (the
    file's payload is uploaded as it lies on disk and brought into medpy's [x, y, z] index order on
    the GPU).  Unlike the reference the image is loaded and uploaded ONCE per image, not once per
    aggregation (:77-79).  Returns {unc: {image_key: {aggregation: result}}}.
    """
    if load_fn is None:
        from .formats import load_to_device

        def load_fn(path):
            return load_to_device(path)[0]

    dev = _lib.require_cuda()
    results = {}
    for unc, unc_path in exp_dataloader.unc_path_dict.items():
        all_uncs = {}
        for image_id in exp_dataloader.image_ids:
            key = f"{image_id}{exp_dataloader.exp_version.unc_ending}"
            all_uncs[key] = {}
            unc_image = _as_map(load_fn(unc_path / key), dev)
            for aggregation in aggregations:
                all_uncs[key][aggregation] = _instantiate(
                    aggregations[aggregation], image=unc_image,
                    pred_model=exp_dataloader.exp_version.pred_model, unc_type=unc)
        results[unc] = all_uncs
        if save:
            with open(exp_dataloader.dataset_path / f"aggregated_{unc}.json", "w") as f:
                json.dump(all_uncs, f, indent=4)
    return results
