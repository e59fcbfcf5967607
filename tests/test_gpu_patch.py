"""values_b200.patch.install() end to end: the reference's own
load_image -> concat_data -> caculcate_uncertainty_multiple_pred -> calculate_metrics -> save_data ->
log_metrics sequence (what run_test drives, test_3D.py:605-700, and what LightningExperiment's
test_step / on_test_end drive, lightning_experiment.py:399-410) is run twice through the names bound
in the reference's modules -- once unpatched on the host (the unmodified reference, from
/root/reference or the offline install under baseline/_ref that travels to the GPU box), once after
install() -- and the two result trees are compared file by file."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference not present")]


def micro_dice(preds, target, ignore_index=None, **_):
    """Stand-in for torchmetrics.functional.dice (absent from this image; SURVEY 8 f2 is parity-unpinned):
    micro-average Dice over the classes left after dropping ignore_index, float predictions arg-maxed."""
    if preds.is_floating_point():
        n_cls = preds.shape[1]
        preds = preds.argmax(dim=1)
    else:
        n_cls = int(max(preds.max(), target.max())) + 1
    p, t = preds.reshape(-1).cpu(), target.reshape(-1).cpu()
    tp = fp = fn = 0
    for c in range(n_cls):
        if c == ignore_index:
            continue
        tp += int(((p == c) & (t == c)).sum()); fp += int(((p == c) & (t != c)).sum()); fn += int(((p != c) & (t == c)).sum())
    den = 2 * tp + fp + fn
    return torch.tensor(2 * tp / den if den else 0.0, dtype=torch.float32)


def run_flow(t3d, samples, probs, n_pred, root):
    """The reference's sequence, written against the names bound in its module `t3d`."""
    carrier = t3d.DataCarrier3D()
    for i, sample in enumerate(samples):
        inp = carrier.load_image(sample)
        inp["data"] = np.expand_dims(inp["data"], axis=0)                      # test_3D.py:419
        batch = {k: torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
                 for k, v in inp.items()}                                        # NumpyToTensor
        for pred_idx in range(n_pred):
            carrier.concat_data(batch=batch, softmax_pred=probs[pred_idx, i:i + 1], n_pred=n_pred, pred_idx=pred_idx)
    t3d.caculcate_uncertainty_multiple_pred(carrier)
    t3d.calculate_metrics(carrier)
    carrier.save_data(root_dir=str(root), exp_name="Ensemble", version=0, org_data_path=None, test_split="id")
    carrier.log_metrics()
    return carrier


def test_install_runs_the_reference_sequence(tmp_path):
    import values_b200 as vb
    import values_b200.patch

    ref = ref_loader.load()
    lidc, _ = ref_loader.load_datamodules()
    t3d, dc = ref.modules["test_3D"], ref.modules["data_carrier_3D"]
    # the reference imports some of its files under bare names too (lightning_experiment.py:23)
    sys.modules.setdefault("data_carrier_3D", dc)
    rng = np.random.default_rng(5)
    # patch_overlap = 1, the shipped setting: with overlapping patches the reference's `seg` holds raw label
    # SUMS and its own calculate_test_metrics fails (one-hot scatter of labels >= C, loss_modules.py:57)
    shape, p, overlap, n_pred = (24, 20, 16), 8, 1, 3
    os.makedirs(tmp_path / "in" / "images"); os.makedirs(tmp_path / "in" / "labels")
    np.save(tmp_path / "in" / "images" / "case.npy", rng.random(shape).astype(np.float32))
    for r in range(2):
        np.save(tmp_path / "in" / "labels" / f"case_0{r}_mask.npy", (rng.random(shape) < 0.3 + 0.1 * r).astype(np.uint8))
    samples = lidc.get_val_test_data_samples(str(tmp_path / "in"), subject_ids=["case.npy"], num_raters=2,
                                             patch_size=p, patch_overlap=overlap)
    assert [s["crop_idx"] for s in samples] == vb.patch_grid(shape, p, overlap)
    g = torch.Generator().manual_seed(3)
    probs = torch.softmax(2 * torch.randn(n_pred, len(samples), 2, p, p, p, generator=g, dtype=torch.float64), dim=2)

    saved_io = {m: (m.save if hasattr(m, "save") else None) for m in (dc,)}
    dc.save = vb.formats.save                      # medpy.io.save is absent here; same writer for both runs
    t3d.dice = micro_dice
    try:
        a = run_flow(t3d, samples, probs, n_pred, tmp_path / "ref")
        assert type(a).__module__.endswith("data_carrier_3D") and isinstance(next(iter(a.data.values()))["softmax_pred"], np.ndarray)
        saved = vb.patch.install()
        try:
            assert t3d.DataCarrier3D is vb.DataCarrier3D and dc.DataCarrier3D is vb.DataCarrier3D
            assert sys.modules["data_carrier_3D"].DataCarrier3D is vb.DataCarrier3D
            b = run_flow(t3d, samples, probs, n_pred, tmp_path / "b200")
            assert isinstance(b, vb.DataCarrier3D) and next(iter(b.data.values()))["softmax_pred"].is_cuda
            # the reference's numpy carrier still works through the patched calculate_metrics
            t3d.calculate_metrics(a)
        finally:
            vb.patch.uninstall(saved)
        assert t3d.DataCarrier3D is not vb.DataCarrier3D
    finally:
        dc.save = saved_io[dc]

    def tree(root):
        return sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs in os.walk(root) for f in fs)

    assert tree(tmp_path / "ref") == tree(tmp_path / "b200")
    assert any(f.endswith("metrics.json") for f in tree(tmp_path / "ref"))
    for rel in tree(tmp_path / "ref"):
        if rel.endswith("metrics.json"):
            ma, mb = json.load(open(tmp_path / "ref" / rel)), json.load(open(tmp_path / "b200" / rel))
            assert ma.keys() == mb.keys()
            for k in ma:
                assert ma[k].keys() == mb[k].keys(), k
                for name in ma[k]:
                    np.testing.assert_allclose(mb[k][name], ma[k][name], rtol=1e-5, atol=1e-6, err_msg=f"{k}/{name}")
            continue
        x, _ = vb.formats.load(tmp_path / "ref" / rel)
        y, _ = vb.formats.load(tmp_path / "b200" / rel)
        assert x.dtype == y.dtype and x.shape == y.shape, rel
        part = rel.split(os.sep)[-2]
        if part in ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty"):
            np.testing.assert_allclose(y, x, rtol=1e-5, atol=1e-6, err_msg=rel)
        elif part == "pred_seg":
            assert (x != y).mean() < 1e-3, rel
        else:
            np.testing.assert_array_equal(y, x, err_msg=rel)
