"""GPU parity tests of the K4 statistics kernels (threshold finding,
NCC, ACE binning) through the reference-facing Python mirrors / the C-ABI, against the golden
vectors the reference itself produced (tests/golden/k4_stats.npz) and against the CPU oracle
on larger seeded inputs.

Tolerances: counts, order statistics and quantiles bit-exact; NCC / ACE 1e-12 relative for
fp64 maps (the reference sums sequentially / pairwise in fp64, the kernels block-then-grid),
1e-5 for fp32 maps (the reference's own fp32 pairwise sums and fp32 exp differ from a
correctly rounded result by that much; bin counts may move by single voxels there).
"""
import json
import os
import types
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def vb():
    import values_b200

    return values_b200


@pytest.fixture(scope="module")
def vo():
    from oracle import values_oracle

    return values_oracle


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "k4_stats.npz"))


class FakeExpDataloader:
    """Duck-typed ExperimentDataloader (same as tests/golden/make_golden.py)."""

    def __init__(self, root, unc_maps, pred_segs, ref_segs, gt_unc):
        self.dataset_path = Path(root)
        self.exp_version = types.SimpleNamespace(exp_path=Path(root), unc_types=sorted(unc_maps),
                                                 pred_model="Dropout", version_name="v0")
        self.image_ids = sorted(pred_segs)
        self._unc, self._pred, self._ref, self._gt = unc_maps, pred_segs, ref_segs, gt_unc

    def get_unc_map(self, image_id, unc_type):
        return self._unc[unc_type][image_id]

    def get_mean_pred_seg(self, image_id):
        return self._pred[image_id]

    def get_pred_segs(self, image_id):
        return [self._pred[image_id]]

    def get_reference_segs(self, image_id):
        return self._ref[image_id]

    def get_gt_unc_map(self, image_id):
        return self._gt[image_id]


# ------------------------------------------------------------------ f1: threshold finding
def test_foreground_quantile_golden(vb, g):
    assert vb.calculate_foreground_quantile_image(g["fg_seg"]) == float(g["fg_quantile"])


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.int64, torch.float32, torch.float64])
@pytest.mark.parametrize("n,offset", [(0, 0), (1, 0), (37, 3), (4099, 1), (1 << 20, 0), ((1 << 20) + 13, 5)])
def test_count_nonzero_exact(vb, dtype, n, offset):
    gen = torch.Generator().manual_seed(n + offset)
    base = (torch.rand(n + offset, generator=gen) < 0.3).to(dtype)
    if dtype.is_floating_point and n > 2:
        base[offset] = float("nan")   # NaN is non-zero
        base[offset + 1] = -0.0       # -0.0 is zero
    x = base.cuda()[offset:]          # unaligned start exercises the scalar head
    got = int(vb.count_nonzero(x).item())
    assert got == int(np.count_nonzero(base[offset:].numpy()))


def test_quantile_golden(vb, g):
    for j, q in enumerate(g["thr_q"].tolist()):
        r64 = vb.quantile(g["thr_maps64"], q)
        assert r64.dtype == np.float64 and r64 == g[f"thr64_{j}"], (j, q)
        r32 = vb.quantile(list(g["thr_maps32"]), q)     # a list of maps == the stacked array
        assert r32.dtype == np.float32 and r32 == g[f"thr32_{j}"], (j, q)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_quantile_matches_numpy_large(vb, dtype):
    rng = np.random.default_rng(5)
    maps = [(rng.random((64, 64, 64)) ** 3).astype(dtype) for _ in range(3)]
    maps[1][:8] = 0.0                       # a large block of duplicates
    maps[2][0, 0, :4] = [-1.5, -0.0, np.inf, 1e-30]
    stacked = np.array(maps)
    for q in (0.0, 1e-7, 0.01, 0.5, 0.98, 0.999999, 1.0):
        got = vb.quantile([torch.from_numpy(m).cuda() for m in maps], q)
        want = np.quantile(stacked, q)
        assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), (q, got, want)  # q=1: inf-inf=NaN in numpy too


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_quantile_over_many_ragged_maps(vb, dtype):
    """A validation set of many separate images (find_threshold.py:90-96): more maps than one launch's table
    holds (96), different sizes, empty maps, views that start off a 16-byte boundary -- one launch per digit
    and table, same exact result as np.quantile of the concatenation."""
    rng = np.random.default_rng(11)
    sizes = [int(s) for s in rng.integers(0, 5000, size=203)]
    sizes[5] = 0
    sizes[96] = 0
    host = [(rng.random(s + 3) ** 2).astype(dtype) for s in sizes]
    dev = [torch.from_numpy(h).cuda()[(i % 3):(i % 3) + sizes[i]] for i, h in enumerate(host)]
    flat = np.concatenate([h[(i % 3):(i % 3) + sizes[i]] for i, h in enumerate(host)])
    for q in (0.0, 0.25, 0.98, 1.0):
        got = vb.quantile(dev, q)
        want = np.quantile(flat, q)
        assert got.dtype == want.dtype and got == want, (q, got, want)


def test_quantile_nan_and_errors(vb):
    x = torch.rand(1000, dtype=torch.float64)
    x[17] = float("nan")
    assert np.isnan(vb.quantile(x.cuda(), 0.3))
    assert np.isnan(np.quantile(x.numpy(), 0.3))
    with pytest.raises(ValueError):
        vb.quantile(torch.rand(10).cuda(), 1.5)


def test_calculate_threshold_image_and_find_threshold(vb, g, tmp_path):
    qp = tmp_path / "quantile_analysis.json"
    qp.write_text(json.dumps({"Dropout": 0.98, "Softmax": 0.9312345}))
    assert vb.calculate_threshold_image(qp, g["thr_maps64"], "Dropout") == g["thr64_3"]
    # find_threshold over "files": the loader is injected (medpy is host I/O, out of scope)
    store = {f"{u}_{i}": g["thr_maps64"][i] * s for u, s in (("aleatoric", 1.0), ("epistemic", 0.5), ("predictive", 2.0))
             for i in range(3)}
    results = {m: {"v0": {f"{u}_uncertainty": [f"{u}_{i}" for i in range(3)]
                          for u in (("predictive",) if m == "Softmax" else ("aleatoric", "epistemic", "predictive"))}}
               for m in ("Dropout", "Softmax")}
    out = vb.find_threshold(results, tmp_path, tmp_path, load_fn=lambda p: store[p])
    saved = json.loads((tmp_path / "threshold_analysis.json").read_text())
    assert saved == out
    for m, q in (("Dropout", 0.98), ("Softmax", 0.9312345)):
        for u, s in (("aleatoric", 1.0), ("epistemic", 0.5), ("predictive", 2.0)):
            if m == "Softmax" and u != "predictive":
                continue
            assert out[m][f"Mean {u} threshold"] == float(np.quantile(g["thr_maps64"] * s, q))
    assert out["Mean"]["Mean predictive threshold"] == np.mean(
        [out["Dropout"]["Mean predictive threshold"], out["Softmax"]["Mean predictive threshold"]])
    # the threshold feeds threshold_aggregation through the json file, as in the reference
    r = vb.threshold_aggregation(g["thr_maps64"][0], threshold_path=tmp_path / "threshold_analysis.json",
                                 pred_model="Dropout", unc_type="aleatoric_uncertainty")
    assert r["threshold"] == out["Dropout"]["Mean aleatoric threshold"]


# ------------------------------------------------------------------ f3: NCC
def test_ncc_golden(vb, g):
    for k in range(3):
        for i in range(3):
            got = vb.compute_ncc(g[f"3d_gt_{i}"], g[f"3d_unc_{k}_{i}"])
            np.testing.assert_allclose(got, float(g[f"3d_ncc_{k}_{i}"]), rtol=1e-12)
    np.testing.assert_allclose(vb.compute_ncc(g["ncc32_a"], g["ncc32_b"]), float(g["ncc32"]), rtol=1e-5)


@pytest.mark.parametrize("shape", [(128, 128, 128), (256, 478), (5,)])
def test_ncc_vs_oracle(vb, vo, shape):
    rng = np.random.default_rng(9)
    a = rng.random(shape)
    b = 0.3 * a + rng.random(shape) + 5.0          # correlated, large mean (cancellation)
    np.testing.assert_allclose(vb.compute_ncc(a, b), vo.compute_ncc(a, b), rtol=1e-11)
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    np.testing.assert_allclose(vb.compute_ncc(a32, b32), vo.compute_ncc(a32.astype(np.float64), b32.astype(np.float64)),
                               rtol=1e-11)       # fp32 inputs, fp64 arithmetic on both sides
    np.testing.assert_allclose(vb.compute_ncc(a32, b), vo.compute_ncc(a32.astype(np.float64), b), rtol=1e-11)


def test_ncc_batched_and_loop(vb, g, tmp_path):
    gt = torch.from_numpy(np.stack([g[f"3d_gt_{i}"] for i in range(3)])).cuda()
    pr = torch.from_numpy(np.stack([g[f"3d_unc_0_{i}"] for i in range(3)])).cuda()
    got = vb.ncc_batched(gt, pr).cpu().numpy()
    np.testing.assert_allclose(got, [float(g[f"3d_ncc_0_{i}"]) for i in range(3)], rtol=1e-12)
    names = [str(u) for u in g["unc_names"]]
    ids = ["img_a", "img_b", "img_c"]
    dl = FakeExpDataloader(tmp_path, {u: {ids[i]: g[f"3d_unc_{k}_{i}"] for i in range(3)} for k, u in enumerate(names)},
                           {ids[i]: g[f"3d_pred_{i}"] for i in range(3)}, {ids[i]: g[f"3d_refs_{i}"] for i in range(3)},
                           {ids[i]: g[f"3d_gt_{i}"] for i in range(3)})
    out = vb.ncc_main(dl)
    assert json.loads((tmp_path / "ambiguity_modeling.json").read_text()) == out
    for k, u in enumerate(names):
        np.testing.assert_allclose(out["mean"][u]["metrics"]["ncc"], float(g[f"3d_ncc_mean_{k}"]), rtol=1e-12)


# ------------------------------------------------------------------ f3: ACE
def test_calib_stats_golden(vb, g):
    disc, tot, nz = vb.calib_stats(g["cs_correct"], g["cs_conf"])
    assert nz == int(g["cs_nonzero"])
    np.testing.assert_array_equal(tot, g["cs_total"])          # counts: exact
    np.testing.assert_allclose(disc, g["cs_disc"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(vb.calc_ace(g["cs_correct"], g["cs_conf"]), float(g["cs_ace"]), rtol=1e-12)


def test_calib_stats_edges_and_errors(vb, vo):
    edges = np.linspace(0.0, 1.0 + 1e-8, 21)
    conf = np.concatenate([edges[:-1], np.nextafter(edges[1:-1], 0), np.nextafter(edges[1:-1], 2), [1.0, 0.0]])
    corr = (np.arange(conf.size) % 2).astype(np.int64)
    bins = vb.metrics.calib_bins(torch.from_numpy(conf).cuda(), torch.from_numpy(corr).cuda()).cpu().numpy()
    want = vo.calib_bincounts(corr, conf)
    np.testing.assert_array_equal(bins[0], want[0])
    np.testing.assert_array_equal(bins[2], want[2])
    np.testing.assert_allclose(bins[1], want[1], rtol=1e-14)
    for c in (np.ones(50, dtype=int), np.zeros(50, dtype=int)):      # single label: sklearn quirk
        a, b = vb.calib_stats(c, conf[:50]), vo.calib_stats(c, conf[:50])
        np.testing.assert_allclose(a[0], b[0], rtol=1e-13)
        assert a[2] == b[2]
    with pytest.raises(ValueError):
        vb.calib_stats(corr, conf + 0.5)
    with pytest.raises(ValueError):
        vb.calib_stats(np.arange(conf.size) % 3, conf)


@pytest.mark.parametrize("tag", ["3d", "2d"])
def test_calibration_error_golden(vb, g, tag, tmp_path):
    ignore = int(g[f"{tag}_ignore"])
    ignore = None if ignore < 0 else ignore
    rtol = 1e-12 if tag == "3d" else 1e-5
    names = [str(u) for u in g["unc_names"]]
    ids = ["img_a", "img_b", "img_c"]
    for k in range(3):
        a, b = g[f"{tag}_platt_{k}"].tolist()
        for i in range(3):
            got = vb.calibration_error_image(g[f"{tag}_unc_{k}_{i}"], g[f"{tag}_pred_{i}"], g[f"{tag}_refs_{i}"],
                                             a, b, ignore)
            np.testing.assert_allclose(got, float(g[f"{tag}_ace_{k}_{i}"]), rtol=rtol, err_msg=f"{tag} {k} {i}")
    dl = FakeExpDataloader(tmp_path, {u: {ids[i]: g[f"{tag}_unc_{k}_{i}"] for i in range(3)} for k, u in enumerate(names)},
                           {ids[i]: g[f"{tag}_pred_{i}"] for i in range(3)}, {ids[i]: g[f"{tag}_refs_{i}"] for i in range(3)},
                           {ids[i]: g[f"{tag}_gt_{i}"] for i in range(3)})
    (tmp_path / "platt_scale_params.json").write_text(json.dumps(
        {u: {"a": float(g[f"{tag}_platt_{k}"][0]), "b": float(g[f"{tag}_platt_{k}"][1])} for k, u in enumerate(names)}))
    out = vb.calibration_error(dl, ignore_value=ignore)
    assert json.loads((tmp_path / "calibration.json").read_text()) == out
    for k, u in enumerate(names):
        np.testing.assert_allclose(out["mean"][u]["metrics"]["ace"], float(g[f"{tag}_ace_mean_{k}"]), rtol=rtol)


@pytest.mark.parametrize("dtype,ldt", [(np.float64, np.uint8), (np.float64, np.int64), (np.float32, np.int32)])
def test_calib_fused_vs_oracle_large(vb, vo, dtype, ldt):
    rng = np.random.default_rng(21)
    shape, R = (96, 100, 101), 4
    unc = (rng.random(shape) ** 2 * 0.7).astype(dtype)
    pred = (rng.random(shape) < 0.4).astype(ldt)
    refs = np.stack([np.where(rng.random(shape) < 0.9, pred, 1 - pred) for _ in range(R)]).astype(ldt)
    refs[rng.random(refs.shape) < 0.05] = 7
    for ignore in (None, 7):
        bins = vb.metrics.calib_bins_fused(torch.from_numpy(unc).cuda(), torch.from_numpy(pred).cuda(),
                                           torch.from_numpy(refs).cuda(), 4.0, -1.0, ignore).cpu().numpy()
        want = vo.calibration_error_image(unc, pred, refs, 4.0, -1.0, ignore, return_bins=True)
        if dtype == np.float64:
            np.testing.assert_array_equal(bins[0], want[0])
            # bin_true before the label_binarize quirk == plain count of correct voxels per bin
            np.testing.assert_array_equal(bins[2], want[2])
            np.testing.assert_allclose(bins[1], want[1], rtol=1e-12)
        else:   # fp32 exp: numpy's and a correctly rounded one differ in the last bit near bin edges
            assert np.abs(bins[0] - want[0]).sum() <= 1e-5 * want[0].sum()
            np.testing.assert_allclose(bins[1], want[1], rtol=1e-5)
        np.testing.assert_allclose(vb.calibration_error_image(unc, pred, refs, 4.0, -1.0, ignore),
                                   vo.calibration_error_image(unc, pred, refs, 4.0, -1.0, ignore),
                                   rtol=1e-12 if dtype == np.float64 else 1e-5)


# ------------------------------------------------------------------ f2: confusion counts / Dice / GED
@pytest.mark.parametrize("ldt", [torch.uint8, torch.int32, torch.int64])
def test_confusion_counts_exact(vb, ldt):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 5, size=(3, 37, 41, 29))
    b = rng.integers(0, 6, size=(2, 37, 41, 29))      # label 5 >= C is dropped
    got = vb.confusion_counts(torch.from_numpy(a).to(ldt).cuda(), torch.from_numpy(b).to(ldt).cuda(), 5).cpu().numpy()
    for i in range(3):
        for j in range(2):
            want = np.zeros((5, 5), dtype=np.int64)
            ok = b[j] < 5
            np.add.at(want, (a[i][ok], b[j][ok]), 1)
            np.testing.assert_array_equal(got[i, j], want)


@pytest.mark.parametrize("c,r,shape,dtype,ldt", [(2, 4, (40, 36, 33), torch.float64, torch.int32),
                                                 (4, 3, (48, 65), torch.float64, torch.int64),
                                                 (2, 1, (31, 30, 29), torch.float32, torch.uint8),
                                                 (7, 2, (20, 21, 22), torch.float64, torch.int32)])
def test_test_metrics_vs_oracle(vb, vo, c, r, shape, dtype, ldt):
    """calculate_test_metrics (test_3D.py:250-281; SoftDiceLoss loss_modules.py:7-90 + NLLLoss + Dice) against
    the oracle, whose loss half is pinned to the reference (tests/test_oracle_vs_reference.py)."""
    g = torch.Generator().manual_seed(c * 7 + r)
    x = torch.softmax(2 * torch.randn(1, c, *shape, generator=g, dtype=torch.float64), dim=1).to(dtype)
    gt = torch.randint(0, c, (r,) + shape, generator=g)
    want = vo.calculate_test_metrics(x, gt)
    got = vb.calculate_test_metrics(x.cuda(), gt.to(ldt).cuda())
    tol = 1e-12 if dtype == torch.float64 else 2e-6      # the reference sums fp32 inputs in fp32, the kernel in fp64
    np.testing.assert_allclose(got["loss"], want["loss"], rtol=tol)
    np.testing.assert_allclose(got["dice"], want["dice"], rtol=1e-12)
    # the raw sums of the kernel against numpy
    terms = vb.seg_loss_terms(x[0].reshape(c, -1).cuda(), gt.reshape(r, -1).to(ldt).cuda()).cpu().numpy()
    ct = 2 if c <= 2 else (4 if c <= 4 else 8)
    xn, gn = x[0].reshape(c, -1).double().numpy(), gt.reshape(r, -1).numpy()
    for rr in range(r):
        for cc in range(c):
            np.testing.assert_allclose(terms[rr, cc], xn[cc][gn[rr] == cc].sum(), rtol=1e-12)
            assert terms[rr, ct + cc] == float((gn[rr] == cc).sum())
            np.testing.assert_allclose(terms[rr, 2 * ct + cc], xn[cc].sum(), rtol=1e-12)
        np.testing.assert_allclose(terms[rr, 3 * ct], np.log(np.take_along_axis(xn, gn[rr][None], 0)).sum(), rtol=1e-12)
    # a probability of exactly 0 under the label: -inf log, +inf loss, as torch
    x0 = x.clone()
    x0[0, int(gt[0].reshape(-1)[0])].view(-1)[0] = 0.0
    assert vb.calculate_test_metrics(x0.cuda(), gt[:1].to(ldt).cuda())["loss"] == float("inf")
    assert vo.calculate_test_metrics(x0, gt[:1])["loss"] == float("inf")


@pytest.mark.parametrize("n,c,r,shape,ignore", [(5, 2, 4, (40, 36, 32), 0), (10, 4, 3, (48, 64), 0),
                                                 (3, 3, 1, (20, 20, 20), 0), (4, 3, 2, (24, 24, 24), 1)])
def test_ged_vs_oracle(vb, vo, n, c, r, shape, ignore):
    """calculate_ged against the numpy restatement (parity unpinned upstream: torchmetrics absent)."""
    g = torch.Generator().manual_seed(n * 10 + c)
    base = 2.0 * torch.randn(1, c, *shape, generator=g)
    sm = torch.softmax(base + torch.randn(n, c, *shape, generator=g), dim=1)
    lab = torch.argmax(base[0], dim=0)
    gt = torch.stack([torch.where(torch.rand(shape, generator=g) < 0.9, lab, (lab + 1) % c) for _ in range(r)])
    if ignore == 1:
        gt[gt == 0] = 2          # no voxel carries label `ignore`... keep 1s: exercises both gt-gt branches
    got = vb.calculate_ged(sm.cuda(), gt.cuda(), ignore_index=ignore)
    want = vo.calculate_ged(sm.numpy(), gt.numpy(), ignore_index=ignore)
    assert set(got) == set(want)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-12, atol=1e-15, err_msg=k)
    assert set(vb.calculate_ged(sm.cuda(), gt.cuda(), ignore_index=ignore, ged_only=True)) == {"ged"}
    # identical predictions and raters: every distance is 0
    one = torch.nn.functional.one_hot(lab, c).movedim(-1, 0).float().unsqueeze(0).repeat(3, *([1] * (lab.dim() + 1)))
    z = vb.calculate_ged(one.cuda(), lab.unsqueeze(0).repeat(2, *([1] * lab.dim())).cuda())
    assert abs(z["ged"]) < 1e-15 and z["max dice pred"] == 1.0
