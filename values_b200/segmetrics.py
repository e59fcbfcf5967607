"""Segmentation-quality metrics on the stack K1 already streams: drop-in
mirror of calculate_ged (uncertainty_modeling/test_3D.py:284-358) and of the torchmetrics `dice`
calls it is built from, over kernel K4 `values_confusion_counts`.

Every Dice / GED term is a function of pairwise CONFUSION MATRICES between label maps -- the
per-sample arg-max maps (K1's `sample_argmax` output, one HBM sweep of the stack) and the rater
segmentations.  The kernel counts them exactly; the few scalars after that are host arithmetic.

PARITY UNPINNED: the arithmetic lives in torchmetrics 0.11.4 (requirements.txt:103), which is not
vendored under the reference and not installed in the build image, and the reference has no test
for it.  `dice_from_confusion` restates torchmetrics' documented micro-average Dice
(2 tp / (2 tp + fp + fn) over the classes left after dropping `ignore_index`, zero_division = 0);
the counts are exact integers, the ratios are evaluated in fp64 (torchmetrics divides in fp32).
The SoftDice + NLL loss of calculate_test_metrics (test_3D.py:250-281) is `calculate_test_metrics` below
(kernel `values_seg_loss_terms`); that part IS pinned: against the reference's own SoftDiceLoss
(uncertainty_modeling/loss_modules.py:7-90) and torch.nn.NLLLoss, tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .uncertainty import uncertainty_fused


def _labels(x: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.device != dev:
        x = x.to(dev)
    if x.dtype not in (torch.uint8, torch.int32, torch.int64):
        x = x.to(torch.int64)
    return x.contiguous()


def confusion_counts(a: torch.Tensor, b: torch.Tensor, n_classes: int) -> torch.Tensor:
    """a [Na, *S], b [Nb, *S] integer label maps (CUDA, one dtype) -> int64 [Na, Nb, C, C] with
    out[i, j, p, q] = #{voxels : a[i] == p and b[j] == q}.  Exact; no sync."""
    if a.device.type != "cuda" or b.device.type != "cuda":
        raise RuntimeError("confusion_counts expects CUDA tensors (no CPU fallback)")
    if a.dtype != b.dtype:
        a, b = a.to(torch.int64), b.to(torch.int64)
    a, b = a.contiguous(), b.contiguous()
    Na, Nb = a.shape[0], b.shape[0]
    V = a[0].numel() if Na else 0
    if Nb and b[0].numel() != V:
        raise ValueError("confusion_counts: label maps must have the same number of voxels")
    out = torch.zeros((Na, Nb, n_classes, n_classes), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib.values_confusion_counts(a.data_ptr(), Na, V, b.data_ptr(), Nb, V,
                                              _lib.label_dtype_code(a.dtype), V, int(n_classes),
                                              out.data_ptr(), _lib.stream_ptr(a.device))
    _lib.check(rc)
    return out


def dice_from_confusion(conf: np.ndarray, ignore_index: Optional[int] = None) -> float:
    """torchmetrics 0.11.4 `dice(preds, target, ignore_index=...)` (micro average, zero_division=0)
    from a confusion matrix conf[pred, target] (any leading axes are pooled, as the reference pools
    repeated pairs into one call, test_3D.py:290-317)."""
    conf = np.asarray(conf, dtype=np.float64)
    conf = conf.reshape((-1,) + conf.shape[-2:]).sum(axis=0)
    keep = [c for c in range(conf.shape[0]) if c != ignore_index]
    tp = sum(conf[c, c] for c in keep)
    fp = sum(conf[c, :].sum() - conf[c, c] for c in keep)
    fn = sum(conf[:, c].sum() - conf[c, c] for c in keep)
    den = 2 * tp + fp + fn
    return float(2 * tp / den) if den > 0 else 0.0


def calculate_ged(output_softmax: torch.Tensor, ground_truth: torch.Tensor, ignore_index: int = 0,
                  ged_only: bool = False) -> Dict[str, float]:
    """Drop-in for uncertainty_modeling/test_3D.py:284-358.
    output_softmax [N, C, *S] (float), ground_truth [R, *S] (integer labels < C)."""
    dev = _lib.require_cuda()
    sm = output_softmax if isinstance(output_softmax, torch.Tensor) else torch.from_numpy(np.asarray(output_softmax))
    if sm.device != dev:
        sm = sm.to(dev)
    if sm.dtype not in (torch.float32, torch.float64, torch.bfloat16):
        sm = sm.float()
    n_pred, n_cls = sm.shape[:2]
    gt = _labels(ground_truth, dev)
    n_rater = gt.shape[0]
    # per-sample arg-max: one sweep of the stack (first maximum wins, as torch.argmax)
    pred = uncertainty_fused(sm.unsqueeze(0), maps=False, sample_argmax=True).sample_argmax[0]
    return ged_from_labels(pred.reshape(n_pred, -1), gt.reshape(n_rater, -1), n_cls, ignore_index, ged_only)


def ged_from_labels(pred: torch.Tensor, gt: torch.Tensor, n_cls: int, ignore_index: int = 0,
                    ged_only: bool = False) -> Dict[str, float]:
    """The arithmetic of calculate_ged (test_3D.py:290-358) on label maps: pred [N, V] (the per-sample arg-max),
    gt [R, V] (CUDA integer labels < n_cls).  Three launches of the confusion-count kernel, the rest on the host."""
    n_pred, n_rater = pred.shape[0], gt.shape[0]
    gt_flat = gt.to(torch.uint8) if n_cls <= 255 and gt.dtype != torch.uint8 and int(gt.max()) < 256 and int(gt.min()) >= 0 else gt
    if gt_flat.dtype != pred.dtype:
        pred = pred.to(gt_flat.dtype)
    conf_pg = confusion_counts(pred, gt_flat, n_cls).cpu().numpy()      # [N, R, C, C]
    conf_pp = confusion_counts(pred, pred, n_cls).cpu().numpy()         # [N, N, C, C]
    conf_gg = confusion_counts(gt_flat, gt_flat, n_cls).cpu().numpy()   # [R, R, C, C]
    dist_gt_pred_2 = 1 - dice_from_confusion(conf_pg, ignore_index)
    dist_pred_pred_2 = 1 - dice_from_confusion(conf_pp, ignore_index if ignore_index == 0 else None)
    gt_has_ignore = bool(conf_gg[:, :, ignore_index, :].sum() > 0) if 0 <= ignore_index < n_cls else False
    dist_gt_gt_2 = 1 - dice_from_confusion(conf_gg, ignore_index if gt_has_ignore else None)
    ged = 2 * dist_gt_pred_2 - dist_pred_pred_2 - dist_gt_gt_2
    ged_dict = {"ged": float(ged)}
    if n_rater > 1 and not ged_only:
        pair = np.array([[dice_from_confusion(conf_pg[n, r], ignore_index) for r in range(n_rater)]
                         for n in range(n_pred)])
        for idx in range(n_rater):          # max over predictions, floor 0 (test_3D.py:322-333)
            ged_dict["max dice rater {}".format(idx)] = float(max(0.0, pair[:, idx].max()))
        ged_dict["max dice pred"] = float(np.mean(np.maximum(0.0, pair.max(axis=1))))
    return ged_dict


def mean_prediction_dice(mean_argmax: torch.Tensor, ground_truth: torch.Tensor, n_classes: int,
                         ignore_index: int = 0) -> float:
    """The `dice` entry of calculate_test_metrics (test_3D.py:271-279): Dice of the arg-max of the
    mean softmax (K1's `mean_argmax`) against every rater, averaged over raters."""
    dev = mean_argmax.device
    gt = _labels(ground_truth, dev)
    gt = gt.reshape(gt.shape[0], -1)
    pred = mean_argmax.reshape(1, -1).to(gt.dtype)
    conf = confusion_counts(pred, gt, n_classes).cpu().numpy()
    return float(np.mean([dice_from_confusion(conf[0, r], ignore_index) for r in range(gt.shape[0])]))


def seg_loss_terms(probs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """probs [C, V] (CUDA fp32 / fp64), labels [R, V] (CUDA u8 / i32 / i64) -> fp64 [R, 3 CT + 1] on the
    device, CT = C rounded up to 2, 4 or 8: per rater {intersect[c], count[c], sum_x[c]} and
    sum_v log probs[label_v][v] (include/values_b200.h, values_seg_loss_terms).  One sweep, no sync."""
    if probs.device.type != "cuda" or labels.device.type != "cuda":
        raise RuntimeError("seg_loss_terms expects CUDA tensors (no CPU fallback)")
    C, V = probs.shape
    R = labels.shape[0]
    if labels.shape[1] != V:
        raise ValueError("seg_loss_terms: probabilities and labels must cover the same voxels")
    probs, labels = probs.contiguous(), labels.contiguous()
    ct = 2 if C <= 2 else (4 if C <= 4 else 8)
    out = torch.empty((R, 3 * ct + 1), dtype=torch.float64, device=probs.device)
    ws_bytes = _lib.lib.values_seg_loss_workspace_bytes(R, C, V)
    ws = torch.empty((max(1, ws_bytes),), dtype=torch.uint8, device=probs.device)
    with torch.cuda.device(probs.device):
        rc = _lib.lib.values_seg_loss_terms(probs.data_ptr(), _lib.dtype_code(probs.dtype), V, labels.data_ptr(),
                                            _lib.label_dtype_code(labels.dtype), V, R, int(C), V, out.data_ptr(),
                                            ws.data_ptr(), ws_bytes, _lib.stream_ptr(probs.device))
    _lib.check(rc)
    return out


def calculate_test_metrics(output_softmax: torch.Tensor, ground_truth: torch.Tensor, ignore_index: int = 0,
                           smooth: float = 1e-5) -> Dict[str, float]:
    """Drop-in for calculate_test_metrics (uncertainty_modeling/test_3D.py:250-281): per rater
    SoftDiceLoss()(x, gt) + NLLLoss()(log x, gt) and torchmetrics `dice(x, gt, ignore_index=0)`, averaged over
    the raters.  output_softmax [1, C, *S] (or [C, *S]), ground_truth [R, *S] integer labels < C, C <= 8.
    The sums come from one sweep of `values_seg_loss_terms` (fp64; the reference sums in the input dtype),
    the Dice from the confusion counts of the arg-max (first maximum wins, as torch.argmax); the few
    scalars after that are host arithmetic."""
    dev = _lib.require_cuda()
    x = output_softmax if isinstance(output_softmax, torch.Tensor) else torch.from_numpy(np.asarray(output_softmax))
    if x.device != dev:
        x = x.to(dev)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    gt = _labels(ground_truth, dev)
    if x.dim() == gt.dim() + 1:          # [1, C, *S] against [R, *S]
        if x.shape[0] != 1:
            raise ValueError("calculate_test_metrics: one prediction (batch 1) against R raters")
        x = x[0]
    n_cls, n_rater = x.shape[0], gt.shape[0]
    V = x[0].numel()
    ct = 2 if n_cls <= 2 else (4 if n_cls <= 4 else 8)
    terms = seg_loss_terms(x.reshape(n_cls, V), gt.reshape(n_rater, V)).cpu().numpy()
    inter, count, sumx, logp = terms[:, :n_cls], terms[:, ct:ct + n_cls], terms[:, 2 * ct:2 * ct + n_cls], terms[:, 3 * ct]
    soft_dice = np.mean(-((2.0 * inter + smooth) / (sumx + count + smooth)), axis=1)      # loss_modules.py:86-90
    nll = -logp / V                                                                       # NLLLoss, mean reduction
    pred = uncertainty_fused(x.reshape((1, 1, n_cls) + tuple(x.shape[1:])), maps=False, mean_argmax=True).mean_argmax[0]
    return {"loss": float(np.mean(soft_dice + nll)),
            "dice": mean_prediction_dice(pred, gt, n_cls, ignore_index)}
