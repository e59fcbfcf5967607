"""ExperimentDataloader: the file-level side of the evaluation loop
(evaluation/experiment_dataloader.py:10-161), on values_b200.formats instead of medpy.

Same constructor, attributes and methods as the reference class, so `aggregate_uncertainties`,
`find_threshold`, `calibration_error` and `ncc.main` run on a test-results directory written by
`DataCarrier3D.save_data` (or by the reference itself).  Differences, all deliberate:
  * `medpy.io.load / save` -> `values_b200.formats.load / save`;
  * `device=True` makes the `get_*` methods return CUDA tensors (payload uploaded in file order,
    axes reversed on the GPU) instead of numpy arrays -- the aggregation entry points take both;
  * a hydra `datamodule_config` is not instantiated here (hydra is absent): pass `dataloader=`
    with an object exposing `.dataset.image_ids` and `.dataset.__getitem__` as the reference's
    test dataloader does, or leave both None to read `gt_seg/` files;
  * `gt_unc_map_loading` / `pred_seg_loading` may be plain callables (the reference instantiates
    hydra `_target_` configs for them, :133-137, :150-154).
"""
from __future__ import annotations

import os
import random
from pathlib import Path

import numpy as np
import torch

from . import _lib, formats


def set_seed(seed: int) -> None:
    """evaluation/utils/set_seed.py:9-19 without pytorch_lightning (pl.seed_everything seeds the
    same three generators)."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    np.random.seed(seed)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)


class ExperimentDataloader:
    def __init__(self, exp_version, dataset_split, dataloader=None, device: bool = False):
        self.exp_version = exp_version
        if "seed" in getattr(exp_version, "version_params", {}):
            set_seed(int(exp_version.version_params["seed"]))
        self.dataset_split = dataset_split
        exp_path = Path(exp_version.exp_path)
        self.dataset_path = exp_path / dataset_split if dataset_split else exp_path
        self.pred_seg_dir = self.dataset_path / "pred_seg"
        self.pred_prob_dir = (self.dataset_path / "pred_prob"
                              if os.path.exists(self.dataset_path / "pred_prob") else None)
        self.device = device
        self.image_ids = sorted(self._get_image_ids())
        if self.exp_version.pred_model == "Softmax":
            self._setup_pred_entropy_softmax()
        self.unc_path_dict = self._setup_unc_path_dict()
        if dataloader is not None:
            self.dataloader = dataloader
            self.ref_seg_dir = None
        elif getattr(exp_version, "datamodule_config", None) is not None:
            raise NotImplementedError(
                "hydra datamodule configs are not instantiated by values_b200; pass dataloader=")
        else:
            self.dataloader = None
            self.ref_seg_dir = self.dataset_path / "gt_seg"

    # ------------------------------------------------------------------ loading
    def _load(self, path):
        if self.device:
            return formats.load_to_device(path)[0]
        return formats.load(path)[0]

    def get_max_softmax_pred(self, image_id: str):
        """1 - max_c p_c from the pred_prob/<id>_01_<cc> files (:38-49)."""
        probs = []
        for class_prob in range(self.exp_version.n_classes):
            prob_file = os.path.join(
                self.pred_prob_dir,
                f"{image_id}_01_{str(class_prob + 1).zfill(2)}{self.exp_version.unc_ending}")
            probs.append(self._load(prob_file))
        if self.device:
            from .uncertainty import calculate_one_minus_msr

            return calculate_one_minus_msr(torch.stack(probs))["pred_entropy"]
        probs = np.array(probs)
        return 1 - np.max(probs, axis=0)

    def _setup_pred_entropy_softmax(self):
        target = self.dataset_path / "pred_entropy"
        if not os.path.exists(target):
            os.makedirs(target)
            for image_id in self.image_ids:
                max_softmax = self.get_max_softmax_pred(image_id)
                path = target / f"{image_id}{self.exp_version.unc_ending}"
                if isinstance(max_softmax, torch.Tensor):
                    formats.save_from_device(max_softmax, path)
                else:
                    formats.save(max_softmax, path)

    def _setup_unc_path_dict(self):
        unc_path_dict = {}
        for unc_type in self.exp_version.unc_types:
            if unc_type == "predictive_uncertainty":
                unc_path_dict[unc_type] = self.dataset_path / "pred_entropy"
            else:
                unc_path_dict[unc_type] = self.dataset_path / unc_type
        return unc_path_dict

    def _get_image_ids(self):
        return set("_".join(image_name.split("_")[:-1])
                   for image_name in os.listdir(self.pred_seg_dir)
                   if image_name.endswith(self.exp_version.image_ending))

    def get_pred_seg_paths(self, image_id):
        return [self.pred_seg_dir / image_path
                for image_path in os.listdir(self.pred_seg_dir)
                if image_path.startswith(image_id) and image_path.endswith(self.exp_version.image_ending)]

    def get_pred_segs(self, image_id):
        return [self._load(p) for p in self.get_pred_seg_paths(image_id)]

    def get_aggregated_unc_files_dict(self):
        out = {}
        for unc in self.unc_path_dict.keys():
            if os.path.isfile(self.dataset_path / f"aggregated_{unc}.json"):
                out[unc] = self.dataset_path / f"aggregated_{unc}.json"
        return out

    def _reference_seg_stack(self, image_id):
        paths = [self.ref_seg_dir / f"{image_id}_{i:02d}{self.exp_version.image_ending}"
                 for i in range(self.exp_version.n_reference_segs)]
        segs = [self._load(p) for p in paths]
        return torch.stack(segs) if self.device else np.array(segs)

    def get_reference_segs(self, image_id):
        if self.dataloader is not None:
            idx = self.dataloader.dataset.image_ids.index(image_id)
            data = self.dataloader.dataset.__getitem__(idx)
            return data["seg"].squeeze().numpy()
        return self._reference_seg_stack(image_id)

    def get_gt_unc_map(self, image_id):
        """Per-pixel variance of the reference segmentations (:117-138)."""
        loading = getattr(self.exp_version, "gt_unc_map_loading", None)
        if loading is None:
            segs = self._reference_seg_stack(image_id)
            if self.device:
                return torch.var(segs.to(torch.float64), dim=0, unbiased=False)
            return np.var(segs, axis=0)
        if callable(loading):
            return loading(image_id=image_id, dataloader=self.dataloader)
        raise NotImplementedError("hydra gt_unc_map_loading configs need hydra; pass a callable")

    def get_mean_pred_seg(self, image_id):
        name = "mean" if self.exp_version.pred_model != "Softmax" else "01"
        pred_seg_path = self.pred_seg_dir / f"{image_id}_{name}{self.exp_version.image_ending}"
        loading = getattr(self.exp_version, "pred_seg_loading", None)
        if loading is None:
            return self._load(pred_seg_path)
        if callable(loading):
            return loading(pred_seg_path=pred_seg_path)
        raise NotImplementedError("hydra pred_seg_loading configs need hydra; pass a callable")

    def get_unc_map(self, image_id, unc_type):
        return self._load(self.unc_path_dict[unc_type] / f"{image_id}{self.exp_version.unc_ending}")
