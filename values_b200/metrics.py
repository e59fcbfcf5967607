"""Per-voxel downstream metrics on B200: drop-in mirrors of evaluation/metrics/ncc.py and the
binning half of evaluation/metrics/ace.py over the K4 statistics kernels.

    compute_ncc(gt_unc_map, pred_unc_map)                    ncc.py:9-25
    ncc_main(exp_dataloader)                                 ncc.py:28-47   (`main` there)
    platt_scale_confid(uncalib_confid, file, uncertainty)    ace.py:42-46
    calib_stats(correct, calib_confids)                      ace.py:49-81
    calc_ace(correct, calib_confids)                         ace.py:84-86
    calibration_error(exp_dataloader, ignore_value=None)     ace.py:89-133

Both are single sweeps over an uncertainty map next to a label / reference map -- the only other
voxel-bandwidth consumers of the C2 maps.  The Platt fit itself
(`platt_scale_params`, sklearn's `_sigmoid_calibration`) is an iterative host optimiser and stays
with sklearn.
"""
from __future__ import annotations

import json
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .threshold import _as_cuda

N_BINS = 20  # ace.py:51


def _edges() -> np.ndarray:
    return np.linspace(0.0, 1.0 + 1e-8, N_BINS + 1)  # ace.py:68


# ------------------------------------------------------------------ NCC
def pair_moments(a: torch.Tensor, b: torch.Tensor, shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a, b [M, *S] (CUDA fp32/fp64) -> fp64 [M, 5] = {sum(a-sa), sum(b-sb), sum (a-sa)^2,
    sum (b-sb)^2, sum (a-sa)(b-sb)}; shift fp64 [M, 2] on the device or None.  No sync."""
    if a.device.type != "cuda" or b.device.type != "cuda":
        raise RuntimeError("pair_moments expects CUDA tensors (no CPU fallback)")
    if a.shape != b.shape:
        raise ValueError("pair_moments: maps must have the same shape")
    a, b = a.contiguous(), b.contiguous()
    M = a.shape[0]
    V = a[0].numel() if M else 0
    dev = a.device
    out = torch.empty((M, 5), dtype=torch.float64, device=dev)
    ws_bytes = _lib.lib.values_pair_moments_workspace_bytes(M, V)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    if shift is not None:
        shift = shift.to(torch.float64).contiguous()
        if tuple(shift.shape) != (M, 2):
            raise ValueError("pair_moments: shift must be [M, 2]")
    with torch.cuda.device(dev):
        rc = _lib.lib.values_pair_moments(a.data_ptr(), _lib.dtype_code(a.dtype), V, b.data_ptr(),
                                          _lib.dtype_code(b.dtype), V, M, V, _lib.ptr(shift),
                                          out.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(dev))
    _lib.check(rc)
    return out


def ncc_batched(gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """gt, pred [M, *S] -> fp64 [M] normalised cross correlation per pair (two sweeps: means, then
    centred moments, as the reference's two-pass formula).  No sync."""
    V = gt[0].numel()
    sums = pair_moments(gt, pred)
    means = sums[:, :2] / V
    c = pair_moments(gt, pred, shift=means)
    sigma_gt = torch.sqrt(c[:, 2] / (V - 1))      # np.std(ddof=1)
    sigma_pred = torch.sqrt(c[:, 3] / (V - 1))
    return (1.0 / (V * sigma_gt * sigma_pred)) * c[:, 4]


def compute_ncc(gt_unc_map, pred_unc_map) -> float:
    """Drop-in for ncc.py:9-25."""
    dev = _lib.require_cuda()
    g = _as_cuda(gt_unc_map, dev, float_only=True)
    p = _as_cuda(pred_unc_map, dev, float_only=True)
    return float(ncc_batched(g.unsqueeze(0), p.unsqueeze(0))[0].item())


def ncc_main(exp_dataloader, save: bool = True) -> Dict:
    """Drop-in for ncc.py:28-47 (`main`): per image and uncertainty type the NCC between the
    rater-variability map and the predicted map, plus the mean; written to ambiguity_modeling.json."""
    ncc_dict: Dict = {"mean": {}}
    for unc_type in exp_dataloader.exp_version.unc_types:
        nccs_unc = []
        for image_id in exp_dataloader.image_ids:
            ncc_dict.setdefault(image_id, {})
            ncc = compute_ncc(exp_dataloader.get_gt_unc_map(image_id),
                              exp_dataloader.get_unc_map(image_id, unc_type))
            ncc_dict[image_id][unc_type] = {"metrics": {"ncc": ncc}}
            nccs_unc.append(ncc)
        ncc_dict["mean"][unc_type] = {"metrics": {"ncc": float(np.mean(np.array(nccs_unc)))}}
    if save:
        with open(exp_dataloader.dataset_path / "ambiguity_modeling.json", "w") as f:
            json.dump(ncc_dict, f, indent=2)
    return ncc_dict


# ------------------------------------------------------------------ ACE
def calib_bins(prob: torch.Tensor, correct: torch.Tensor) -> torch.Tensor:
    """prob (CUDA fp32/fp64) and correct (uint8/int32/int64, non-zero = true), same numel ->
    fp64 [3, 21] = {bin_total, bin_sums, bin_true} (ace.py:72-74).  No sync."""
    if prob.device.type != "cuda" or correct.device.type != "cuda":
        raise RuntimeError("calib_bins expects CUDA tensors (no CPU fallback)")
    prob, correct = prob.contiguous(), correct.contiguous()
    if prob.numel() != correct.numel():
        raise ValueError("calib_bins: prob and correct must have the same number of elements")
    n = prob.numel()
    dev = prob.device
    out = torch.empty((3, N_BINS + 1), dtype=torch.float64, device=dev)
    ws_bytes = _lib.lib.values_calib_bins_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib.values_calib_bins(prob.data_ptr(), _lib.dtype_code(prob.dtype), correct.data_ptr(),
                                        _lib.label_dtype_code(correct.dtype), n, _lib.dbl_array(_edges()),
                                        N_BINS, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        _lib.stream_ptr(dev))
    _lib.check(rc)
    return out


def calib_bins_fused(unc_map: torch.Tensor, pred_seg: torch.Tensor, reference_segs: torch.Tensor,
                     a: float, b: float, ignore_value: Optional[int] = None) -> torch.Tensor:
    """One sweep of the per-image body of calibration_error (ace.py:96-127): unc_map [*S] (CUDA
    fp32/fp64), pred_seg [*S], reference_segs [R, *S] (one integer dtype) -> fp64 [3, 21]."""
    if unc_map.device.type != "cuda":
        raise RuntimeError("calib_bins_fused expects CUDA tensors (no CPU fallback)")
    unc_map = unc_map.contiguous()
    V = unc_map.numel()
    if reference_segs.dim() == unc_map.dim():
        reference_segs = reference_segs.unsqueeze(0)
    R = reference_segs.shape[0]
    if pred_seg.numel() != V or reference_segs[0].numel() != V:
        raise ValueError("calib_bins_fused: pred_seg / reference_segs do not match the map")
    ldt = pred_seg.dtype if pred_seg.dtype == reference_segs.dtype else torch.int64
    if ldt not in (torch.uint8, torch.int32, torch.int64):
        ldt = torch.int64
    # gt_seg files are fp64 (label sum / count, data_carrier_3D.py:232-241) and the reference compares
    # them with pred_seg and ignore_value in floating point (ace.py:108-117): a non-integral label
    # (overlapping patches whose raters disagree) equals no class.  Such voxels get a label no
    # prediction can take instead of being truncated to a class.
    if reference_segs.is_floating_point():
        frac = reference_segs != reference_segs.round()
        reference_segs = reference_segs.round().to(torch.int64)
        if bool(frac.any()):
            ldt = torch.int64
            reference_segs = torch.where(frac, torch.full_like(reference_segs, torch.iinfo(torch.int32).max),
                                         reference_segs)
    if pred_seg.is_floating_point():
        pred_seg = pred_seg.round()
    pred_seg = pred_seg.to(ldt).contiguous()
    reference_segs = reference_segs.to(ldt).contiguous()
    dev = unc_map.device
    out = torch.empty((3, N_BINS + 1), dtype=torch.float64, device=dev)
    ws_bytes = _lib.lib.values_calib_bins_workspace_bytes(V)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib.values_calib_bins_fused(
            unc_map.data_ptr(), _lib.dtype_code(unc_map.dtype), pred_seg.data_ptr(),
            reference_segs.data_ptr(), _lib.label_dtype_code(ldt), V, R, float(a), float(b),
            int(ignore_value is not None), int(ignore_value or 0), _lib.dbl_array(_edges()), N_BINS,
            out.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(dev))
    _lib.check(rc)
    return out


def _stats_from_bins(bins: np.ndarray) -> Tuple[np.ndarray, np.ndarray, int]:
    """ace.py:64-81 from the three bincounts.  label_binarize quirk kept: with a single label
    present (all correct or all wrong) sklearn's label_binarize yields all zeros (ace.py:60-66)."""
    bin_total, bin_sums, bin_true = bins[0], bins[1], bins[2].copy()
    n, n_true = bin_total.sum(), bin_true.sum()
    if n_true == 0 or n_true == n:
        bin_true[:] = 0.0
    nonzero = bin_total != 0
    num_nonzero = int(nonzero.sum())
    prob_true = bin_true[nonzero] / bin_total[nonzero]
    prob_pred = bin_sums[nonzero] / bin_total[nonzero]
    prob_total = bin_total[nonzero] / bin_total.sum()
    return np.abs(prob_true - prob_pred), prob_total, num_nonzero


def calib_stats(correct, calib_confids):
    """Drop-in for ace.py:49-81 -> (bin_discrepancies, prob_total, num_nonzero)."""
    dev = _lib.require_cuda()
    y_prob = _as_cuda(calib_confids, dev, float_only=True).reshape(-1)
    y_true = _as_cuda(correct, dev).reshape(-1)
    if y_true.dtype in (torch.float32, torch.float64):
        y_true = y_true.to(torch.int64)
    if y_prob.numel():
        lo, hi = torch.aminmax(y_prob)
        if float(lo) < 0 or float(hi) > 1:
            raise ValueError("y_prob has values outside [0, 1] and normalize is set to False.")
    labels = torch.unique(y_true)
    if labels.numel() > 2:
        raise ValueError(f"Only binary classification is supported. Provided labels {labels.cpu().numpy()}.")
    if labels.numel() == 2:   # label_binarize: the larger label is the positive class
        y_true = (y_true == labels[1]).to(torch.uint8)
    else:
        y_true = torch.zeros_like(y_true, dtype=torch.uint8)
    return _stats_from_bins(calib_bins(y_prob, y_true).cpu().numpy())


def calc_ace(correct, calib_confids) -> float:
    """Drop-in for ace.py:84-86."""
    bin_discrepancies, _, num_nonzero = calib_stats(correct, calib_confids)
    return float((1 / num_nonzero) * np.sum(bin_discrepancies))


def platt_scale_confid(uncalib_confid, platt_scale_file, uncertainty):
    """Drop-in for ace.py:42-46 on the device (returns a CUDA tensor)."""
    with open(platt_scale_file) as f:
        params = json.load(f)[uncertainty]
    dev = _lib.require_cuda()
    x = _as_cuda(uncalib_confid, dev, float_only=True)
    return 1 / (1 + torch.exp(x * params["a"] + params["b"]))


def calibration_error_image(unc_map, pred_seg, reference_segs, a: float, b: float,
                            ignore_value: Optional[int] = None) -> float:
    """ACE of one image: the loop body of calibration_error (ace.py:96-127) in one fused sweep."""
    dev = _lib.require_cuda()
    unc = _as_cuda(unc_map, dev, float_only=True)
    pred = _as_cuda(pred_seg, dev)
    refs = _as_cuda(reference_segs, dev)
    if pred.shape != unc.shape:   # 2d unc map is loaded in shape (W, H)  (ace.py:101-102)
        unc = unc.swapaxes(0, 1).contiguous()
    bins = calib_bins_fused(unc, pred, refs, a, b, ignore_value).cpu().numpy()
    disc, _, num_nonzero = _stats_from_bins(bins)
    return float((1 / num_nonzero) * np.sum(disc))


def calibration_error(exp_dataloader, ignore_value=None, save: bool = True) -> Dict:
    """Drop-in for ace.py:89-133: per image and uncertainty type the ACE of the Platt-scaled
    confidence against rater agreement; written to calibration.json."""
    calib_dict: Dict = {"mean": {}}
    platt_scale_file = exp_dataloader.exp_version.exp_path / "platt_scale_params.json"
    with open(platt_scale_file) as f:
        params_dict = json.load(f)
    for unc_type in exp_dataloader.exp_version.unc_types:
        aces_unc = []
        params = params_dict[unc_type]
        for image_id in exp_dataloader.image_ids:
            calib_dict.setdefault(image_id, {})
            ace = calibration_error_image(
                exp_dataloader.get_unc_map(image_id, unc_type),
                exp_dataloader.get_mean_pred_seg(image_id),
                exp_dataloader.get_reference_segs(image_id), params["a"], params["b"], ignore_value)
            calib_dict[image_id][unc_type] = {"metrics": {"ace": ace}}
            aces_unc.append(ace)
        calib_dict["mean"][unc_type] = {"metrics": {"ace": float(np.mean(np.array(aces_unc)))}}
    if save:
        with open(exp_dataloader.dataset_path / "calibration.json", "w") as f:
            json.dump(calib_dict, f, indent=2)
    return calib_dict
