"""Build container only: the oracle against the UNMODIFIED reference imported in place
(oracle/ref_loader.py).  Skipped where /root/reference is absent (the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import values_oracle as vo

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("n,c,spatial,dtype", [
    (5, 2, (12, 12, 12), torch.float64),
    (16, 4, (10, 9, 8), torch.float32),
    (10, 20, (24, 30), torch.float32),
    (3, 7, (17,), torch.float32),
])
def test_c2_bit_exact(ref, n, c, spatial, dtype):
    g = torch.Generator().manual_seed(n * 100 + c)
    x = torch.softmax(3 * torch.randn(n, c, *spatial, generator=g, dtype=torch.float64), 1).to(dtype)
    x[0, 0].view(-1)[:3] = 0.0
    a, b = ref.calculate_uncertainty(x), vo.calculate_uncertainty(x)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    a, b = ref.calculate_uncertainty(x, ssn=True), vo.calculate_uncertainty(x, ssn=True)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_c3_random(ref):
    rng = np.random.default_rng(3)
    for shape, p in [((20, 21, 22), 10), ((40, 33), 10), ((11, 12, 13), [2, 3, 4])]:
        m = rng.random(shape)
        a = ref.patch_level_aggregation(m, p)
        b = vo.patch_level_aggregation(m, p)
        assert a["bounding_box"] == b["bounding_box"]
        np.testing.assert_allclose(a["max_score"], b["max_score"], rtol=1e-12)
        assert ref.image_level_aggregation(m) == vo.image_level_aggregation(m)
        ra, rb = ref.threshold_aggregation(m, threshold=0.7), vo.threshold_aggregation(m, threshold=0.7)
        assert float(ra["max_score"]) == float(rb["max_score"])


@pytest.mark.parametrize("c,r,shape,dtype", [(2, 4, (12, 11, 10), torch.float64), (4, 2, (20, 18), torch.float64),
                                             (2, 1, (9, 9, 9), torch.float32)])
def test_f2_test_metrics_loss_is_the_reference_loss(ref, c, r, shape, dtype):
    """calculate_test_metrics (test_3D.py:250-281) run UNMODIFIED, with only the absent torchmetrics `dice`
    bound to the oracle's dice_micro: the SoftDiceLoss + NLLLoss half of the oracle is pinned bit for bit."""
    t3d = ref.modules["test_3D"]
    g = torch.Generator().manual_seed(c * 10 + r)
    x = torch.softmax(2 * torch.randn(1, c, *shape, generator=g, dtype=torch.float64), dim=1).to(dtype)
    gt = torch.randint(0, c, (r,) + shape, generator=g)
    saved = t3d.dice
    t3d.dice = lambda p, t, ignore_index=None: torch.tensor(
        vo.dice_micro(torch.argmax(p, dim=1).numpy(), t.numpy(), p.shape[1], ignore_index), dtype=torch.float64)
    try:
        want = t3d.calculate_test_metrics(x, gt)
    finally:
        t3d.dice = saved
    got = vo.calculate_test_metrics(x, gt)
    assert got["loss"] == want["loss"] and got["dice"] == want["dice"]
    assert float(t3d.SoftDiceLoss()(x, gt[:1])) == float(vo.soft_dice_loss(x, gt[:1]))


def test_k4_stats_random(ref):
    rng = np.random.default_rng(11)
    for dt in (np.float64, np.float32):
        a, b = rng.random((9, 8, 7)).astype(dt), rng.random((9, 8, 7)).astype(dt)
        assert ref.compute_ncc(a, b) == vo.compute_ncc(a, b)
        conf = rng.random(4000).astype(dt)
        for corr in ((rng.random(4000) < conf).astype(int), np.ones(4000, dtype=int), np.zeros(4000, dtype=int)):
            ra, rb = ref.calib_stats(corr, conf), vo.calib_stats(corr, conf)
            np.testing.assert_array_equal(ra[0], rb[0])
            np.testing.assert_array_equal(ra[1], rb[1])
            assert ra[2] == rb[2] and ref.calc_ace(corr, conf) == vo.calc_ace(corr, conf)
    seg = (rng.random((6, 7, 8)) < 0.2).astype(np.uint8)
    assert ref.calculate_foreground_quantile_image(seg) == vo.calculate_foreground_quantile_image(seg)


def test_experiment_dataloader_mirror_matches_reference_class(ref, tmp_path):
    """the reference's ExperimentDataloader (medpy.io replaced by values_b200.formats,
    medpy being absent) and our mirror must agree on ids, paths and every array they hand out."""
    import importlib
    from pathlib import Path

    from test_formats import exp_version, make_results_dir   # tests/ is on sys.path (rootdir conftest)
    from values_b200 import formats
    from values_b200.experiment_dataloader import ExperimentDataloader

    edl = importlib.import_module("evaluation.experiment_dataloader")
    edl.load, edl.save = formats.load, formats.save
    for pred_model in ("Dropout", "Softmax"):
        root = Path(tmp_path) / pred_model
        d, _ = make_results_dir(root, pred_model=pred_model)
        if pred_model == "Softmax":
            import shutil

            shutil.rmtree(d / "pred_entropy")
        a = edl.ExperimentDataloader(exp_version(root, pred_model=pred_model), "id")
        b = ExperimentDataloader(exp_version(root, pred_model=pred_model), "id")
        assert a.image_ids == b.image_ids and a.unc_path_dict == b.unc_path_dict
        assert a.dataset_path == b.dataset_path and a.ref_seg_dir == b.ref_seg_dir
        for image_id in a.image_ids:
            assert sorted(a.get_pred_seg_paths(image_id)) == sorted(b.get_pred_seg_paths(image_id))
            np.testing.assert_array_equal(a.get_mean_pred_seg(image_id), b.get_mean_pred_seg(image_id))
            np.testing.assert_array_equal(a.get_reference_segs(image_id), b.get_reference_segs(image_id))
            np.testing.assert_array_equal(a.get_gt_unc_map(image_id), b.get_gt_unc_map(image_id))
            np.testing.assert_array_equal(a.get_max_softmax_pred(image_id), b.get_max_softmax_pred(image_id))
            for unc in a.unc_path_dict:
                np.testing.assert_array_equal(a.get_unc_map(image_id, unc), b.get_unc_map(image_id, unc))


def test_patch_install_rebinds_and_restores_reference_names(ref):
    """values_b200.patch.install() puts the B200 path behind the
    reference's own module attributes (which is also how hydra resolves its `_target_` strings),
    uninstall() restores them.  Rebinding only -- nothing is computed here."""
    import importlib

    import values_b200 as vb
    import values_b200.patch as patch

    t3d, dc, agg = ref.modules["test_3D"], ref.modules["data_carrier_3D"], ref.modules["aggregate_uncertainties"]
    edl = importlib.import_module("evaluation.experiment_dataloader")
    before = (t3d.calculate_uncertainty, dc.DataCarrier3D, agg.patch_level_aggregation, dc.save, edl.ExperimentDataloader)
    saved = patch.install(file_io=True)
    try:
        assert t3d.calculate_uncertainty is vb.calculate_uncertainty
        assert t3d.caculcate_uncertainty_multiple_pred is vb.caculcate_uncertainty_multiple_pred
        assert dc.DataCarrier3D is vb.DataCarrier3D
        target = "evaluation.uncertainty_aggregation.aggregate_uncertainties.patch_level_aggregation"
        mod, _, name = target.rpartition(".")
        assert getattr(importlib.import_module(mod), name) is vb.patch_level_aggregation      # hydra's lookup
        assert agg.aggregate_uncertainties is vb.aggregate_uncertainties
        assert ref.modules["find_threshold"].find_threshold is vb.find_threshold
        assert ref.modules["ncc"].compute_ncc is vb.compute_ncc and ref.modules["ace"].calc_ace is vb.calc_ace
        assert dc.save is vb.formats.save and edl.load is vb.formats.load
        assert edl.ExperimentDataloader is vb.ExperimentDataloader
    finally:
        patch.uninstall(saved)
    after = (t3d.calculate_uncertainty, dc.DataCarrier3D, agg.patch_level_aggregation, dc.save, edl.ExperimentDataloader)
    assert all(a is b for a, b in zip(before, after))


A5_CASES = [((20, 17, 33), 8, 0.5), ((64, 64, 64), 16, 1), ((70, 45, 33), 16, 0.75), ((40, 40, 40), 10, 0.25),
            ((33, 20, 17), 8, 0.3), ((16, 16, 16), 16, 1), ((15, 40, 40), 16, 0.5), ((128, 128, 128), 64, 0.5)]


def reference_crop_loop(module, shape, patch_size, overlap, tmp_path, toy):
    """Run the reference's own get_val_test_data_samples (the crop-index loop of the sliding-window
    path) on a temporary .npy volume and return its crop tuples."""
    import os

    sub = "Tr" if toy else ""
    os.makedirs(tmp_path / f"images{sub}", exist_ok=True)
    os.makedirs(tmp_path / f"labels{sub}", exist_ok=True)
    np.save(tmp_path / f"images{sub}" / "vol.npy", np.zeros(shape, np.uint8))
    np.save(tmp_path / f"labels{sub}" / ("vol_00.npy" if toy else "vol_00_mask.npy"), np.zeros(shape, np.uint8))
    samples = module.get_val_test_data_samples(str(tmp_path), subject_ids=["vol.npy"], num_raters=1,
                                               patch_size=patch_size, patch_overlap=overlap)
    assert all(s["image_path"].endswith("vol.npy") and s["label_paths"] is not None for s in samples)
    return [s["crop_idx"] for s in samples]


@pytest.mark.parametrize("shape,patch_size,overlap", A5_CASES)
def test_a5_patch_grid_is_the_reference_crop_loop(shape, patch_size, overlap, tmp_path):
    """lidc_idri_datamodule_3D.py:719-736 and toy_datamodule_3D.py:637-654, run
    unmodified, against the oracle's patch_grid (the product's patch_grid is compared with the
    committed fixture tests/golden/patch_grid.json, which make_golden.py writes from this loop)."""
    lidc, toy = ref_loader.load_datamodules()
    want = reference_crop_loop(lidc, shape, patch_size, overlap, tmp_path / "lidc", toy=False)
    assert want == reference_crop_loop(toy, shape, patch_size, overlap, tmp_path / "toy", toy=True)
    assert want == vo.patch_grid(shape, patch_size, overlap)
    if min(shape) < patch_size:
        assert want == []
