"""GPU tests of the file-level hand-off: the axis-reversal kernel, device
load / save, DataCarrier3D.save_data against the arrays the reference's own save_data handed to
medpy.io.save, and ExperimentDataloader + aggregate_uncertainties against the aggregated_<unc>.json
the reference's loop wrote (tests/golden/save_data_3d.{npz,json}, made by tests/golden/make_golden.py)."""
import json
import os
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def vb():
    import values_b200

    return values_b200


@pytest.mark.parametrize("shape", [(5, 4, 3), (33, 1, 70), (64, 65, 31), (128, 128, 128), (37, 53), (1024, 478), (9,)])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int16, torch.float32, torch.float64])
def test_reverse_axes_matches_permute(vb, shape, dtype):
    g = torch.Generator().manual_seed(len(shape))
    x = (torch.rand(shape, generator=g) * 200).to(dtype).cuda()
    y = vb.reverse_axes(x)
    want = x.permute(*reversed(range(x.dim()))).contiguous()
    assert y.is_contiguous() and y.shape == want.shape and torch.equal(y, want)
    assert torch.equal(vb.reverse_axes(y), x)                # an involution
    with pytest.raises(RuntimeError):
        vb.reverse_axes(x.cpu())


@pytest.mark.parametrize("dtype", [np.uint8, np.int32, np.float32, np.float64])
def test_device_load_save_match_host(vb, tmp_path, dtype):
    rng = np.random.default_rng(1)
    a = (rng.random((40, 33, 27)) * 100).astype(dtype)
    vb.formats.save(a, tmp_path / "h.nii.gz")
    t, hdr = vb.load_to_device(tmp_path / "h.nii.gz")
    assert t.is_cuda and t.is_contiguous() and tuple(t.shape) == a.shape
    np.testing.assert_array_equal(t.cpu().numpy(), a)
    vb.save_from_device(t, tmp_path / "d.nii.gz")
    assert (tmp_path / "d.nii.gz").read_bytes() == (tmp_path / "h.nii.gz").read_bytes()   # same file, byte for byte
    img = rng.random((37, 53)).astype(np.float32)
    vb.formats.save(img.T, tmp_path / "u.tif")
    t2, _ = vb.load_to_device(tmp_path / "u.tif")
    np.testing.assert_array_equal(t2.cpu().numpy(), img.T)


def test_aggregation_takes_medpy_style_views(vb, tmp_path):
    """A Fortran-ordered view (what medpy.io.load returns) is uploaded as it lies in memory and
    re-ordered on the GPU; results equal those of the C-contiguous copy."""
    rng = np.random.default_rng(2)
    a = rng.random((31, 40, 45))
    vb.formats.save(a, tmp_path / "m.nii.gz")
    view, _ = vb.formats.load(tmp_path / "m.nii.gz")
    assert view.flags.f_contiguous and not view.flags.c_contiguous
    assert vb.patch_level_aggregation(view, 10) == vb.patch_level_aggregation(np.ascontiguousarray(a), 10)
    assert vb.image_level_aggregation(view) == vb.image_level_aggregation(a)
    assert vb.threshold_aggregation(view, threshold=0.5) == vb.threshold_aggregation(a, threshold=0.5)


def _golden_carrier(vb):
    g = np.load(os.path.join(GOLDEN, "stitch_3d.npz"))
    shape, p = tuple(g["shape"].tolist()), int(g["patch"])
    crops = vb.patch_grid(shape, p, float(g["overlap"]))
    patches = torch.from_numpy(g["patches"])
    n_pred = patches.shape[0]
    carrier = vb.DataCarrier3D()
    for pred_idx in range(n_pred):
        for s in range(0, len(crops), 5):
            idx = list(range(s, min(s + 5, len(crops))))
            batch = {"image_paths": ["vol_a.npy"] * len(idx), "label_paths": [["lab_a.npy"]] * len(idx),
                     "org_image_size": [shape] * len(idx), "crop_idx": [crops[i] for i in idx],
                     "data": torch.zeros(len(idx), 1, p, p, p),
                     "seg": torch.zeros(1, len(idx), p, p, p, dtype=torch.int32)}
            carrier.concat_data(batch, patches[pred_idx, idx], n_pred=n_pred, pred_idx=pred_idx)
    vb.caculcate_uncertainty_multiple_pred(carrier)
    return carrier


def test_save_data_matches_reference_files_and_aggregation(vb, tmp_path):
    gold = np.load(os.path.join(GOLDEN, "save_data_3d.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "save_data_3d.json")))
    carrier = _golden_carrier(vb)
    carrier.save_data(root_dir=str(tmp_path), exp_name="Dropout", version=0, org_data_path=None, test_split="id")
    written = sorted(os.path.relpath(os.path.join(d, f), tmp_path) for d, _, fs in os.walk(tmp_path) for f in fs)
    assert written == sorted(k.replace("|", os.sep) for k in gold.files)          # same tree, same names
    unc_dirs = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")
    for key in gold.files:
        rel = key.replace("|", os.sep)
        arr, _ = vb.formats.load(tmp_path / rel)
        want = gold[key]
        assert arr.dtype == want.dtype and arr.shape == want.shape, rel
        if rel.split(os.sep)[-2] in unc_dirs:   # fp32 maps (K1) / count, stored as fp64 like the reference
            np.testing.assert_allclose(arr, want, rtol=1e-5, atol=1e-6, err_msg=rel)
        elif "pred_seg" in rel:                # arg-max: exact away from 2-ulp ties of the class means
            assert (arr != want).mean() < 1e-3, rel
        else:
            np.testing.assert_array_equal(arr, want, err_msg=rel)                # fp64 sums / counts: bit-exact
    # the evaluation side on the directory just written, maps loaded straight to the device
    version = SimpleNamespace(
        exp_path=Path(tmp_path) / "Dropout" / "test_results" / "0", version_params={"seed": 123},
        pred_model="Dropout", n_classes=2, image_ending=".nii.gz", unc_ending=".nii.gz",
        unc_types=list(meta["unc_dirs"]), n_reference_segs=1, datamodule_config=None,
        gt_unc_map_loading=None, pred_seg_loading=None)
    loader = vb.ExperimentDataloader(version, "id", device=True)
    assert loader.image_ids == meta["image_ids"]
    assert {k: os.path.relpath(v, tmp_path) for k, v in loader.unc_path_dict.items()} == meta["unc_dirs"]
    assert sorted(os.path.relpath(p, tmp_path) for p in loader.get_pred_seg_paths("vol_a")) == meta["pred_seg_files"]
    assert float(loader.get_gt_unc_map("vol_a").sum()) == meta["gt_unc_map_sum"]
    got = vb.aggregate_uncertainties(loader, meta["aggregations"])
    for unc, per_image in meta["aggregated"].items():
        on_disk = json.load(open(loader.dataset_path / f"aggregated_{unc}.json"))
        assert on_disk == json.loads(json.dumps(got[unc]))
        for image_key, aggs in per_image.items():
            mine = got[unc][image_key]
            assert [list(b) for b in mine["patch_level"]["bounding_box"]] == aggs["patch_level"]["bounding_box"]
            for name in aggs:
                np.testing.assert_allclose(mine[name]["max_score"], aggs[name]["max_score"], rtol=1e-5, atol=1e-6)
            assert mine["threshold"]["threshold"] == aggs["threshold"]["threshold"]
