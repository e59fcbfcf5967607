"""DataCarrier3D on B200: the stitch accumulator + result container of
uncertainty_modeling/data_carrier_3D.py, with the accumulators resident in HBM.

Same entry point and `.data` layout as the reference (data_carrier_3D.py:99-135):

    carrier.concat_data(batch, softmax_pred, n_pred=1, pred_idx=0, sigma=None)
    carrier.data[image_path] = {label_paths, softmax_pred [n_pred, C, X,Y,Z] fp64 raw sums,
                                num_predictions [C, X,Y,Z] fp64, data [X,Y,Z] fp64,
                                seg [R, X,Y,Z] int32, (sigma), + the three uncertainty maps}

The arrays are CUDA tensors instead of numpy arrays (no device->host copy per patch, which
is what the reference does at :161); `numpy_data()` materialises the reference's numpy
layout.  The reference hardcodes C=2 (:120); here C is taken from the softmax batch.
`normalized()` is the arithmetic of the save path (:208-217, 253-259, 281-285, 323-363) on the
device; `save_data()` writes the reference's directory layout of `.nii.gz` files (:181-371)
through values_b200.formats (axis reversal on the GPU, one D2H copy per file, gzip on the host).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .aggregation import normalize_maps
from .stitching import stitch_accumulate
from .uncertainty import MAP_KEYS, uncertainty_fused


class DataCarrier3D:
    def __init__(self, device: Optional[torch.device] = None, accum_dtype: torch.dtype = torch.float64,
                 patch_weight: Optional[torch.Tensor] = None, stitch_path: int = 0):
        self.data: Dict[str, Dict] = {}
        self.save_dir = None
        self.device = device
        self.accum_dtype = accum_dtype  # fp64 = reference parity (np.zeros default)
        # opt-in importance map [p, p, p] (e.g. stitching.gaussian_importance_map): softmax sums
        # and `num_predictions` become weighted sums.  None = uniform, the reference's behaviour.
        self.patch_weight = patch_weight
        self.stitch_path = stitch_path   # K3 implementation for this carrier's calls (tests; 0 = automatic)

    @staticmethod
    def load_image(sample: Dict) -> Dict:
        """Drop-in for data_carrier_3D.py:59-97 (called by predict_cases, test_3D.py:369, 418): the
        input dict of ONE crop of a 3-D `.npy` volume -- host-side file slicing, nothing for the GPU.
        Keys: image_paths, label_paths, crop_idx, org_image_size, data [1, p,p,p], seg [R, 1, p,p,p] (int32)."""
        (x0, x1), (y0, y1), (z0, z1) = sample["crop_idx"]
        volume = np.load(sample["image_path"], mmap_mode="r")
        out = {
            "image_paths": [sample["image_path"]],
            "label_paths": [sample["label_paths"]],
            "crop_idx": [sample["crop_idx"]],
            "org_image_size": [volume.shape],
            "data": np.expand_dims(volume[x0:x1, y0:y1, z0:z1], 0),
        }
        if sample["label_paths"] is not None:
            crops = [np.load(path, mmap_mode="r")[x0:x1, y0:y1, z0:z1] for path in sample["label_paths"]]
            out["seg"] = np.expand_dims(np.array(crops, dtype=np.intc), 1)
        return out

    def _dev(self) -> torch.device:
        if self.device is None:
            self.device = _lib.require_cuda()
        return self.device

    def concat_data(self, batch: Dict, softmax_pred: torch.Tensor, n_pred: int = 1,
                    pred_idx: int = 0, sigma: torch.Tensor = None) -> None:
        """Drop-in for data_carrier_3D.py:99-179.  `softmax_pred` [B, C, p,p,p] (any device);
        batch["crop_idx"][i] = ((x0,x1),(y0,y1),(z0,z1)).  Patches of one call that belong to
        the same image may overlap: the kernel sums them output-stationary in list order."""
        dev = self._dev()
        sp = softmax_pred.detach()
        if sp.device != dev:
            sp = sp.to(dev, non_blocking=True)
        if sp.dtype not in (torch.float32, torch.float64, torch.bfloat16):
            sp = sp.float()
        sg = None
        if sigma is not None:
            sg = sigma.detach().to(dev)
            if sg.dtype not in (torch.float32, torch.float64, torch.bfloat16):
                sg = sg.float()
        n_cls = sp.shape[1]
        groups: Dict[str, list] = {}
        for index, image_path in enumerate(batch["image_paths"]):
            if image_path not in self.data:
                size = tuple(int(s) for s in batch["org_image_size"][index])
                entry = {
                    "label_paths": batch["label_paths"][index],
                    "softmax_pred": torch.zeros((n_pred, n_cls) + size, dtype=self.accum_dtype, device=dev),
                    "_count": torch.zeros(size, dtype=torch.float64, device=dev),
                    "data": torch.zeros(size, dtype=torch.float64, device=dev),
                }
                if sg is not None:
                    entry["sigma"] = torch.zeros((n_pred, n_cls) + size, dtype=self.accum_dtype, device=dev)
                n_raters = len(batch["label_paths"][index]) if batch["label_paths"][index] is not None else 0
                entry["seg"] = torch.zeros((n_raters,) + size, dtype=torch.int32, device=dev)
                # reference layout: [C, X, Y, Z] count (all class planes identical, :127-129)
                entry["num_predictions"] = entry["_count"].unsqueeze(0).expand((n_cls,) + size)
                self.data[image_path] = entry
            groups.setdefault(image_path, []).append(index)
        for image_path, idxs in groups.items():
            entry = self.data[image_path]
            # crop origins and patch indices of the call in ONE host-to-device copy: [3 n] origins, then [n] indices
            n_sel = len(idxs)
            meta = np.empty(4 * n_sel, dtype=np.int32)
            meta[:3 * n_sel] = [batch["crop_idx"][i][d][0] for i in idxs for d in range(3)]
            meta[3 * n_sel:] = idxs
            meta_dev = torch.from_numpy(meta).to(dev)
            crop_lo = meta_dev[:3 * n_sel].view(n_sel, 3)
            pidx = meta_dev[3 * n_sel:]
            stitch_accumulate(sp.unsqueeze(0), crop_lo, entry["softmax_pred"][pred_idx:pred_idx + 1],
                              entry["_count"] if pred_idx == 0 else None, patch_index=pidx, accumulate=True,
                              weight=self.patch_weight, path=self.stitch_path)
            if sg is not None:
                stitch_accumulate(sg.unsqueeze(0), crop_lo, entry["sigma"][pred_idx:pred_idx + 1],
                                  None, patch_index=pidx, accumulate=True, weight=self.patch_weight,
                                  path=self.stitch_path)
            if pred_idx == 0:  # image / label slabs (:138-153): bookkeeping, plain torch slicing
                for i in idxs:
                    (x0, x1), (y0, y1), (z0, z1) = batch["crop_idx"][i]
                    if "data" in batch and batch["data"] is not None:
                        img = batch["data"][i].detach().to(dev).squeeze()
                        entry["data"][x0:x1, y0:y1, z0:z1] += img.to(torch.float64)
                    if "seg" in batch and batch["seg"] is not None and entry["seg"].shape[0] > 0:
                        seg = batch["seg"][:, i].detach().to(dev).to(torch.int32)
                        entry["seg"][:, x0:x1, y0:y1, z0:z1] += seg.reshape(
                            (entry["seg"].shape[0], x1 - x0, y1 - y0, z1 - z0))

    # ------------------------------------------------------------------ save-path arithmetic
    def normalized(self, image_path: str) -> Dict[str, torch.Tensor]:
        """What save_data writes (data_carrier_3D.py:208-217, 253-259, 281-285, 323-363), on device:
        softmax / clip(count,1), its mean over samples, mean_seg / per-sample arg-max (uint8)
        and every uncertainty map / clip(count,1) as fp64."""
        v = self.data[image_path]
        cnt = v["_count"]
        n_pred, n_cls = v["softmax_pred"].shape[:2]
        size = tuple(cnt.shape)
        clip_min = 1.0 if self.patch_weight is None else 0.0
        sm = normalize_maps(v["softmax_pred"].reshape((n_pred * n_cls,) + size), cnt, clip_min)
        sm = sm.reshape((n_pred, n_cls) + size)
        res = uncertainty_fused(sm.unsqueeze(0), maps=False, mean_argmax=True, sample_argmax=True)
        mean_sm = None
        if n_pred > 1:   # np.mean(axis=0) (:254): samples added in index order, one IEEE division
            mean_sm = sm[0].clone()
            for i in range(1, n_pred):
                mean_sm += sm[i]
            # (a CUDA divisor: torch turns division by a host scalar into a multiplication by 1/n)
            mean_sm /= torch.tensor(float(n_pred), dtype=mean_sm.dtype, device=mean_sm.device)
        out = {
            "softmax_pred": sm,
            "mean_softmax_pred": mean_sm,
            "mean_seg": res.mean_argmax[0],
            "pred_seg": res.sample_argmax[0],
        }
        present = [k for k in MAP_KEYS if k in v]
        if present:
            maps = torch.stack([v[k].to(cnt.device) for k in present])
            norm = normalize_maps(maps, cnt, clip_min)
            for i, k in enumerate(present):
                out[k] = norm[i]
        return out

    # ------------------------------------------------------------------ file output
    def _create_save_dirs(self, root_dir: str, exp_name: str, version: int, sigma_save_dir: bool,
                          test_split: str = "id") -> None:
        """data_carrier_3D.py:19-57 (its `if id is None` branch is dead: `id` is the builtin)."""
        import os

        self.save_dir = os.path.join(root_dir, exp_name, "test_results", str(version), test_split)
        self.save_input_dir = os.path.join(self.save_dir, "input")
        self.save_gt_dir = os.path.join(self.save_dir, "gt_seg")
        self.save_pred_dir = os.path.join(self.save_dir, "pred_seg")
        self.save_pred_prob_dir = os.path.join(self.save_dir, "pred_prob")
        dirs = [self.save_dir, self.save_input_dir, self.save_gt_dir, self.save_pred_dir, self.save_pred_prob_dir]
        if sigma_save_dir:
            self.save_pred_sigma_dir = os.path.join(self.save_dir, "sigma")
            dirs.append(self.save_pred_sigma_dir)
        for d in dirs:
            os.makedirs(d, exist_ok=True)

    def save_data(self, root_dir: str, exp_name: str, version: int, org_data_path: str = None,
                  test_split: str = "id") -> None:
        """Drop-in for data_carrier_3D.py:181-371: same directories, file names, dtypes (fp64 maps,
        probabilities, image and labels; uint8 segmentations) and arithmetic; files are written by
        values_b200.formats.save_from_device instead of medpy.io.save."""
        import os

        from . import formats

        sigma_save_dir = "sigma" in list(self.data.values())[0]
        self._create_save_dirs(root_dir, exp_name, version, sigma_save_dir, test_split)
        clip_min = 1.0 if self.patch_weight is None else 0.0
        for key, value in self.data.items():
            name = key.split("/")[-1].split(".")[0]
            cnt = value["_count"]
            norm = self.normalized(key)
            header = False
            if org_data_path:
                _, header = formats.load(os.path.join(org_data_path, name + ".nii.gz"))
            data = normalize_maps(value["data"].unsqueeze(0), cnt, clip_min)[0]
            formats.save_from_device(data, os.path.join(self.save_input_dir, name + ".nii.gz"), header)
            if value["seg"].shape[0] > 0:
                gt_seg = normalize_maps(value["seg"].to(torch.float64), cnt, clip_min)   # int32 / fp64 -> fp64
                for seg_idx in range(gt_seg.shape[0]):
                    formats.save_from_device(gt_seg[seg_idx], os.path.join(
                        self.save_gt_dir, "{}_{}.nii.gz".format(name, str(seg_idx).zfill(2))), header)
            softmax_pred = norm["softmax_pred"]
            if softmax_pred.shape[0] > 1:
                formats.save_from_device(norm["mean_seg"], os.path.join(
                    self.save_pred_dir, "{}_{}.nii.gz".format(name, "mean")), header)
                mean_softmax_pred = norm["mean_softmax_pred"]
                for class_idx in range(mean_softmax_pred.shape[0]):
                    formats.save_from_device(mean_softmax_pred[class_idx], os.path.join(
                        self.save_pred_prob_dir,
                        "{}_{}_{}.nii.gz".format(name, "mean", str(class_idx + 1).zfill(2))), header)
            if sigma_save_dir and "sigma" in value:
                n_pred, n_cls = value["sigma"].shape[:2]
                sigma = normalize_maps(value["sigma"].reshape((n_pred * n_cls,) + tuple(cnt.shape)), cnt,
                                       clip_min).reshape(value["sigma"].shape)
            for pred_idx in range(softmax_pred.shape[0]):
                formats.save_from_device(norm["pred_seg"][pred_idx], os.path.join(
                    self.save_pred_dir, "{}_{}.nii.gz".format(name, str(pred_idx + 1).zfill(2))), header)
                for class_idx in range(softmax_pred.shape[1]):
                    formats.save_from_device(softmax_pred[pred_idx, class_idx], os.path.join(
                        self.save_pred_prob_dir, "{}_{}_{}.nii.gz".format(
                            name, str(pred_idx + 1).zfill(2), str(class_idx + 1).zfill(2))), header)
                    if "sigma" in value and pred_idx == 0:
                        formats.save_from_device(sigma[pred_idx, class_idx], os.path.join(
                            self.save_pred_sigma_dir, "{}_{}.nii.gz".format(name, str(class_idx + 1).zfill(2))),
                            header)
            for unc in MAP_KEYS:
                if unc in value:
                    unc_dir = os.path.join(self.save_dir, unc)
                    os.makedirs(unc_dir, exist_ok=True)
                    formats.save_from_device(norm[unc], os.path.join(unc_dir, name + ".nii.gz"), header)

    def log_metrics(self) -> None:
        """Drop-in for data_carrier_3D.py:373-391: `metrics.json` in the results directory, one entry
        per image plus the mean of every metric over the images."""
        import json
        import os

        per_image = {path: dict(value["metrics"]) for path, value in self.data.items()}
        names = []
        for scores in per_image.values():
            names += [m for m in scores if m not in names]
        per_image["mean"] = {m: np.asarray([s[m] for p, s in per_image.items() if p != "mean" and m in s]).mean()
                             for m in names}
        with open(os.path.join(self.save_dir, "metrics.json"), "w") as f:
            json.dump(per_image, f, indent=2)

    def numpy_data(self) -> Dict[str, Dict]:
        """The reference's `.data` layout with numpy arrays (torch fp32 maps stay torch CPU
        tensors, as test_3D.py:533 stores them)."""
        out = {}
        for key, v in self.data.items():
            e = {}
            for name, val in v.items():
                if name == "_count":
                    continue
                if isinstance(val, torch.Tensor):
                    e[name] = val.cpu() if name in MAP_KEYS else val.cpu().numpy()
                else:
                    e[name] = val
            out[key] = e
        return out


def calculate_metrics(test_datacarrier) -> None:
    """Drop-in for calculate_metrics (uncertainty_modeling/test_3D.py:537-575) on a device-resident
    carrier: the mean softmax prediction and the rater labels are normalised on the GPU
    (`normalize_maps`, the `/ clip(count, 1)` of :545-547, 553-566); GED / max-Dice come from
    values_b200.segmetrics.calculate_ged, the SoftDice + NLL loss and the Dice of the mean prediction
    (calculate_test_metrics, test_3D.py:250-281) from values_b200.segmetrics.calculate_test_metrics --
    all on the device.  A reference (numpy) DataCarrier3D is passed through to the reference's own
    function."""
    import sys

    from .segmetrics import calculate_ged, calculate_test_metrics

    ref = sys.modules.get("uncertainty_modeling.test_3D") or sys.modules.get("test_3D")
    if not isinstance(test_datacarrier, DataCarrier3D):
        original = getattr(ref, "_values_b200_original_calculate_metrics", None) if ref else None
        if original is None:
            raise TypeError("calculate_metrics expects a values_b200.DataCarrier3D")
        return original(test_datacarrier)
    for key, value in test_datacarrier.data.items():
        cnt = value["_count"]
        size = tuple(cnt.shape)
        n_pred, n_cls = value["softmax_pred"].shape[:2]
        clip_min = 1.0 if test_datacarrier.patch_weight is None else 0.0
        sm = normalize_maps(value["softmax_pred"].reshape((n_pred * n_cls,) + size), cnt, clip_min)
        sm = sm.reshape((n_pred, n_cls) + size)
        metrics_dict = {}
        if value["seg"].shape[0] > 0 and n_cls <= 8:
            # :545-552 -- the reference averages RAW sums / count over samples, gt = raw label sums
            metrics_dict.update(calculate_test_metrics(torch.mean(sm, dim=0), value["seg"]))
        if value["seg"].shape[0] > 1 or n_pred > 1:
            gt = normalize_maps(value["seg"].to(torch.float64), cnt, clip_min).to(torch.int32)   # np.asarray(.., intc)
            metrics_dict.update(calculate_ged(sm, gt))
        value["metrics"] = metrics_dict
