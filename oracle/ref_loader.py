"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (values_b200/).

Imports the UNMODIFIED reference (IML-DKFZ/values) in place from /root/reference so
that (a) the oracle restatement in oracle/values_oracle.py can be pinned against it and
(b) golden vectors can be generated (tests/golden/make_golden.py).

The reference imports several third-party packages at module scope that are not
installed in this image (hydra, medpy, batchgenerators, torchmetrics,
pytorch_lightning, albumentations, jsbeautifier, ...).  None of the hot-path function
bodies touch them, so they are replaced by MagicMock modules.

/root/reference does not exist on the GPU box: callers must check `available()`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """VALUES_REFERENCE_ROOT, else the read-only mount of the build container, else the offline
    pip install of the UNMODIFIED reference under baseline/_ref (git-ignored; it travels to the
    GPU box with the snapshot, which is how `bench.py --impl reference` runs the real thing there):
        python -m pip install --no-index --no-build-isolation --no-deps --ignore-requires-python \
            --target baseline/_ref <copy of /root/reference>
    (--ignore-requires-python because setup.py pins python ==3.10; --no-deps because its pinned
    requirements are not in the wheelhouse -- the hot-path bodies need only torch/numpy/scipy.)"""
    for cand in (os.environ.get("VALUES_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "uncertainty_modeling")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_STUBBED = [
    "hydra", "hydra.utils", "omegaconf", "medpy", "medpy.io", "batchgenerators",
    "batchgenerators.dataloading", "batchgenerators.dataloading.data_loader",
    "batchgenerators.dataloading.multi_threaded_augmenter",
    "batchgenerators.dataloading.single_threaded_augmenter",
    "batchgenerators.transforms", "batchgenerators.transforms.abstract_transforms",
    "batchgenerators.transforms.noise_transforms",
    "batchgenerators.transforms.spatial_transforms",
    "batchgenerators.transforms.color_transforms",
    "batchgenerators.transforms.crop_and_pad_transforms",
    "batchgenerators.transforms.utility_transforms",
    "batchgenerators.transforms.sample_normalization_transforms",
    "batchgenerators.augmentations", "batchgenerators.augmentations.utils",
    "batchgenerators.augmentations.crop_and_pad_augmentations",
    "torchmetrics", "torchmetrics.functional", "torchmetrics.functional.classification",
    "pytorch_lightning", "pytorch_lightning.loggers", "pytorch_lightning.callbacks",
    "albumentations", "albumentations.pytorch", "jsbeautifier", "SimpleITK", "nibabel",
    "matplotlib", "matplotlib.pyplot", "seaborn", "tifffile", "skimage", "tqdm",
    "pydantic",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "uncertainty_modeling"))


class _Anything(types.ModuleType):
    """A module whose every attribute is a MagicMock (so `from x import y` works)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


_loaded = {}


def load():
    """Return a namespace with the reference's hot-path callables (unmodified)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)

    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for name in _STUBBED:
        try:
            importlib.import_module(name)
        except Exception:
            mod = _Anything(name)
            mod.__path__ = []  # behave like a package
            sys.modules[name] = mod
    # tqdm is used as `for x in tqdm(iterable)`: make it the identity.
    if isinstance(sys.modules.get("tqdm"), _Anything):
        sys.modules["tqdm"].tqdm = lambda it, *a, **k: it
    for p in (
        REFERENCE_ROOT,
        os.path.join(REFERENCE_ROOT, "uncertainty_modeling"),
        os.path.join(REFERENCE_ROOT, "evaluation"),
    ):
        if p not in sys.path:
            sys.path.insert(0, p)
    t3d = importlib.import_module("uncertainty_modeling.test_3D")
    dc = importlib.import_module("uncertainty_modeling.data_carrier_3D")
    agg = importlib.import_module(
        "evaluation.uncertainty_aggregation.aggregate_uncertainties"
    )
    thr = importlib.import_module("evaluation.uncertainty_aggregation.find_threshold")
    ncc = importlib.import_module("evaluation.metrics.ncc")
    ace = importlib.import_module("evaluation.metrics.ace")
    _loaded.update(
        calculate_uncertainty=t3d.calculate_uncertainty,
        calculate_one_minus_msr=t3d.calculate_one_minus_msr,
        caculcate_uncertainty_multiple_pred=t3d.caculcate_uncertainty_multiple_pred,
        DataCarrier3D=dc.DataCarrier3D,
        patch_level_aggregation=agg.patch_level_aggregation,
        image_level_aggregation=agg.image_level_aggregation,
        threshold_aggregation=agg.threshold_aggregation,
        calculate_foreground_quantile_image=thr.calculate_foreground_quantile_image,
        compute_ncc=ncc.compute_ncc, calib_stats=ace.calib_stats, calc_ace=ace.calc_ace,
        platt_scale_confid=ace.platt_scale_confid, calibration_error=ace.calibration_error,
        ncc_main=ncc.main, calculate_threshold_image=thr.calculate_threshold_image,
        modules=dict(test_3D=t3d, data_carrier_3D=dc, aggregate_uncertainties=agg,
                     find_threshold=thr, ncc=ncc, ace=ace),
    )
    return types.SimpleNamespace(**_loaded)


def load_datamodules():
    """The reference's two 3-D datamodules (unmodified), whose `get_val_test_data_samples` holds the
    crop-index loop of the sliding-window path (lidc_idri_datamodule_3D.py:719-736,
    toy_datamodule_3D.py:637-654).  Their batchgenerators / lightning imports are stubbed like the rest."""
    load()
    out = []
    for name in ("uncertainty_modeling.lidc_idri_datamodule_3D", "uncertainty_modeling.toy_datamodule_3D"):
        for _ in range(8):   # stub whatever further third-party sub-module the import asks for
            try:
                out.append(importlib.import_module(name))
                break
            except ModuleNotFoundError as e:
                if not e.name or e.name.split(".")[0] not in {n.split(".")[0] for n in _STUBBED}:
                    raise
                mod = _Anything(e.name)
                mod.__path__ = []
                sys.modules[e.name] = mod
        else:
            raise RuntimeError(f"could not import {name}")
    return out
