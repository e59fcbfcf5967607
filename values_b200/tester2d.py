"""The 2-D caller of the hot path: a mirror of `Tester.process_output` / `calculate_test_metrics`
(uncertainty_modeling/test_2D.py:161-173, 204-254) for a whole batch on the device.

The reference takes the stacked predictions [N, B, C, H, W] of one test batch, appends an all-zero channel
(a full copy of the stack, :206-218) so that torchmetrics can address the ignore label as class C, relabels
`gt == ignore_index` to C, and then loops over the images: mean over the samples, Dice of its arg-max against
every rater, GED over the per-sample arg-max maps, the uncertainty maps, two file writes.

Here the batch goes through ONE K1 launch on the stack as it lies in memory (the [N, B, ...] -> [B, N, ...]
permutation is a strided view, include/values_b200.h): maps, the arg-max of the mean and the per-sample
arg-max of every image.  The zero channel is never materialised: it contributes nothing to the entropies
(0 log 0 is skipped, test_3D.py:493-494) and cannot win an arg-max against a softmax row, so class C only
exists in the relabelled ground truth and in the size of the confusion matrices.  Metrics are host arithmetic
on the confusion counts (values_b200.segmetrics; the Dice itself is torchmetrics' and parity-unpinned, see
there).  File writes stay with the caller (values_b200.formats)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .segmetrics import _labels, confusion_counts, dice_from_confusion, ged_from_labels
from .uncertainty import calculate_one_minus_msr, uncertainty_fused

MAP_KEYS = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")


def process_output(all_preds: Dict, is_ssn: bool = False, ignore_index: int = 255) -> Dict[str, Dict]:
    """all_preds: {"softmax_pred": [N, B, C, H, W] CUDA float stack (torch.stack of the N predictions, :317),
    "gt": [B, R, H, W] integer labels (ignore_index marks unlabeled pixels), "image_id": B ids,
    "dataset": B names}.  Returns {image_id: {"dataset", "metrics": {"dice", "ged"}, "uncertainty": the
    reference's dict of fp32 [H, W] maps, "mean_argmax": u8 [H, W], "sample_argmax": u8 [N, H, W],
    "ignore_index_map": bool [H, W]}} -- what process_output leaves in results_dict plus the tensors it hands to
    save_prediction / save_uncertainty.  `all_preds` is not modified (the reference relabels gt in place)."""
    dev = _lib.require_cuda()
    sm = all_preds["softmax_pred"]
    if sm.device != dev:
        sm = sm.to(dev)
    if sm.dim() != 5:
        raise ValueError("softmax_pred must be [N, B, C, H, W]")
    n_pred, n_img, n_cls = sm.shape[:3]
    spatial = tuple(sm.shape[3:])
    gt = _labels(all_preds["gt"], dev)
    if gt.dim() == 3:
        gt = gt.unsqueeze(1)
    ignore_map = gt == ignore_index
    gt = torch.where(ignore_map, torch.full_like(gt, n_cls), gt)            # :219-222, without touching the input
    stack = sm.permute(1, 0, 2, 3, 4)                                        # [B, N, C, H, W], a view
    if n_pred > 1:
        res = uncertainty_fused(stack, mean_argmax=True)
        sample_argmax = uncertainty_fused(stack, maps=False, sample_argmax=True).sample_argmax
    else:
        res = None
        sample_argmax = uncertainty_fused(stack, maps=False, sample_argmax=True).sample_argmax
    out: Dict[str, Dict] = {}
    for b in range(n_img):
        image_id = all_preds["image_id"][b]
        mean_argmax = res.mean_argmax[b] if res is not None else sample_argmax[b, 0]
        gt_b = gt[b].reshape(gt.shape[1], -1)
        metrics = {"dice": _mean_dice(mean_argmax, gt_b, n_cls + 1, n_cls)}                      # :161-173
        metrics.update(ged_from_labels(sample_argmax[b].reshape(n_pred, -1), gt_b, n_cls + 1,    # :234-241
                                       ignore_index=n_cls, ged_only=True))
        if res is not None:
            unc = res.as_dict(b, ssn=is_ssn)                                                    # :242-243
        else:
            unc = calculate_one_minus_msr(stack[b, 0])                                          # :244-245
        out[image_id] = {"dataset": all_preds["dataset"][b] if "dataset" in all_preds else None,
                         "metrics": metrics, "uncertainty": unc, "mean_argmax": mean_argmax,
                         "sample_argmax": sample_argmax[b], "ignore_index_map": ignore_map[b, 0]}
    return out


def _mean_dice(mean_argmax: torch.Tensor, gt: torch.Tensor, n_classes: int, ignore_index: int) -> float:
    pred = mean_argmax.reshape(1, -1)
    if gt.dtype != pred.dtype:
        pred = pred.to(gt.dtype)
    conf = confusion_counts(pred, gt, n_classes).cpu().numpy()
    return float(np.mean([dice_from_confusion(conf[0, r], ignore_index) for r in range(gt.shape[0])]))


def results_dict_with_mean(results: Dict[str, Dict]) -> Dict[str, Dict]:
    """save_results_dict (test_2D.py:256-270) without the file: {"<id>": {"dataset", "metrics"}, ...,
    "mean": {"metrics": mean of every metric over the images}} ready for json.dump."""
    out = {k: {"dataset": v["dataset"], "metrics": dict(v["metrics"])} for k, v in results.items()}
    names = []
    for v in results.values():
        names += [m for m in v["metrics"] if m not in names]
    out["mean"] = {"metrics": {m: float(np.asarray([v["metrics"][m] for v in results.values() if m in v["metrics"]]).mean())
                               for m in names}}
    return out
