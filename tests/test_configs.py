"""One parity test per BASELINE.json config, through the reference-facing API, against the CPU
oracle on the same seeded inputs (sizes reduced where the oracle would take minutes; the
full-size cases are covered by size-independent properties here and in test_gpu_parity.py).

cfg1  toy 3D, MC-dropout N=5, C=2: PE/EE/MI + image-level aggregation (sum and mean)
cfg2  LIDC 64^3 patches, 5-member ensemble, C=2: patch-level + threshold aggregation, batched
cfg3  sliding-window stitching with TTA (N=8 and the reference's true N=16) + MI maps on raw sums
cfg4  GTA5/Cityscapes 2D, N=10, C=19 + the zero channel (and the reference-true C=24+1 at
      256x478), [N, B, C, H, W] strided layout, all C3 aggregations, batch sweep; fp32 and bf16
cfg5  volume-sharded sweep shape N=16, C=4: fused pipeline, score table layout, sharding ranges
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
MAPS = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")


@pytest.fixture(scope="module")
def vb():
    import values_b200

    return values_b200


@pytest.fixture(scope="module")
def vo():
    from oracle import values_oracle

    return values_oracle


def stack(seed, shape, class_dim, dtype=torch.float32, shared=False):
    g = torch.Generator().manual_seed(seed)
    logits = 3.0 * torch.randn(shape, generator=g, dtype=torch.float64)
    if shared:  # samples mostly agree: the small-MI regime of real ensembles
        base_shape = list(shape)
        base_shape[class_dim - 1] = 1
        logits = 3.0 * torch.randn(base_shape, generator=g, dtype=torch.float64) + 0.1 * logits
    return torch.softmax(logits, dim=class_dim).to(dtype)


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


# ----------------------------------------------------------------------------------------- cfg1
def test_cfg1_toy_mcdropout_image_level(vb, vo):
    B, N, C, S = 4, 5, 2, (64, 64, 64)          # Case_1 test set has 20 volumes; 4 keep the oracle quick
    x = stack(101, (B, N, C) + S, 2, torch.float64, shared=True)   # the 3D path feeds fp64 (test_3D.py:532)
    for b in range(B):
        got = vb.calculate_uncertainty(x[b])          # CPU tensor in -> CPU tensors out
        ref = vo.calculate_uncertainty(x[b])
        for k in MAPS:
            assert got[k].dtype == torch.float32 and got[k].device.type == "cpu"
            close(got[k], ref[k])
            m = ref[k].numpy()
            close(vb.image_level_aggregation(got[k].numpy())["max_score"], vo.image_level_aggregation(m)["max_score"], 1e-5, 1e-3)
            close(vb.image_level_aggregation(got[k].numpy(), mean=True), vo.image_level_aggregation(m, mean=True), 1e-5, 1e-8)
    # Softmax model (N == 1): 1 - MSR stored under pred_entropy (test_3D.py:521-525)
    one = vb.calculate_one_minus_msr(x[0, 0])
    close(one["pred_entropy"], 1 - x[0, 0].max(dim=0).values, 0, 0)


# ----------------------------------------------------------------------------------------- cfg2
def test_cfg2_lidc_ensemble_patch_and_threshold(vb, vo):
    B, N, C, S = 24, 5, 2, (64, 64, 64)
    x = stack(202, (B, N, C) + S, 2, torch.float32)
    thr = (0.45, 0.40, 0.03)
    cfg = vb.AggregationConfig(patch_size=10, thresholds=thr)
    res = vb.UncertaintyPipeline(cfg).run(x.cuda(), keep_maps=True, mean_argmax=True)
    ids = [f"LIDC-{b:04d}" for b in range(B)]
    d = res.to_dicts(ids)
    for b in (0, 7, B - 1):                       # oracle on a sample of the batch
        ref = vo.calculate_uncertainty(x[b])
        for k, key in enumerate(MAPS):
            m = res.maps[k, b].cpu().numpy()
            close(m, ref[key])
            pl = vo.patch_level_aggregation(m.astype(np.float64), 10)
            e = d[key][ids[b]]
            assert e["patch_level"]["bounding_box"] == pl["bounding_box"]
            close(e["patch_level"]["max_score"], pl["max_score"], 1e-12, 0)
            th = vo.threshold_aggregation(m, threshold=thr[k])
            close(e["threshold"]["max_score"], float(th["max_score"]), 1e-6, 0)
            assert e["threshold"]["threshold"] == thr[k]
    # the reference-facing single-image calls agree with the batched pipeline bit for bit
    m0 = res.maps[0, 3].cpu().numpy()
    assert vb.patch_level_aggregation(m0, 10) == d["pred_entropy"][ids[3]]["patch_level"]
    assert vb.threshold_aggregation(m0, threshold=thr[0])["max_score"] == d["pred_entropy"][ids[3]]["threshold"]["max_score"]


# ----------------------------------------------------------------------------------------- cfg3
@pytest.mark.parametrize("shape,overlap,n_pred", [
    ((128, 128, 128), 1.0, 8),      # shipped settings generalised: disjoint patches
    ((128, 128, 128), 0.5, 16),     # 3^3 = 27 overlapping patches, the reference's true TTA count
    ((100, 90, 70), 0.5, 8),        # non-multiple size: uncovered remainder stays zero
])
def test_cfg3_stitching_tta_then_mi(vb, vo, shape, overlap, n_pred):
    p, C = 32, 2
    crops = vb.patch_grid(shape, p, overlap)
    assert crops == vo.patch_grid(shape, p, overlap)
    g = torch.Generator().manual_seed(len(crops) + n_pred)
    patches = torch.softmax(3 * torch.randn(n_pred, len(crops), C, p, p, p, generator=g, dtype=torch.float64), dim=2)
    st = vo.StitchOracle(n_classes=C)
    carrier = vb.DataCarrier3D()
    bs = 6
    for pi in range(n_pred):
        for s in range(0, len(crops), bs):
            idx = list(range(s, min(s + bs, len(crops))))
            batch = {"image_paths": ["vol"] * len(idx), "label_paths": [None] * len(idx),
                     "org_image_size": [shape] * len(idx), "crop_idx": [crops[i] for i in idx]}
            st.concat_data(batch, patches[pi, idx], n_pred=n_pred, pred_idx=pi)
            carrier.concat_data(dict(batch, data=None, seg=None), patches[pi, idx], n_pred=n_pred, pred_idx=pi)
    v = carrier.data["vol"]
    ref_sum, ref_cnt = st.data["vol"]["softmax_pred"], st.data["vol"]["num_predictions"]
    np.testing.assert_array_equal(v["softmax_pred"].cpu().numpy(), ref_sum)          # fp64 bit-exact
    np.testing.assert_array_equal(v["num_predictions"].cpu().numpy(), ref_cnt)
    if overlap == 0.5 and shape == (100, 90, 70):
        assert (ref_cnt[0][96:, :, :] == 0).all() and v["softmax_pred"][:, :, 96:].abs().max().item() == 0
    # MI maps on the RAW accumulated sums (test_3D.py:532; reference quirk H5), then / clip(count, 1)
    vb.caculcate_uncertainty_multiple_pred(carrier)
    ref = vo.calculate_uncertainty(torch.from_numpy(ref_sum))
    scale = max(1.0, float(ref_cnt.max())) * n_pred   # overlap sums are up to 8x larger than probabilities
    for k in MAPS:
        close(v[k].cpu(), ref[k], RTOL, ATOL * scale)
    norm = carrier.normalized("vol")
    close(norm["epistemic_uncertainty"].cpu(), vo.normalize_map(ref["epistemic_uncertainty"], ref_cnt), RTOL, ATOL * scale)
    np.testing.assert_array_equal(norm["mean_seg"].cpu().numpy(), vo.stitched_argmax(ref_sum, ref_cnt))


# ----------------------------------------------------------------------------------------- cfg4
@pytest.mark.parametrize("B", [1, 2, 6])
def test_cfg4_2d_strided_batch_all_aggregations(vb, vo, B):
    N, C, H, W = 10, 19, 128, 256                # 1024x2048 at full size: see test_cfg4_full_size_properties
    full = stack(400 + B, (N, B, C, H, W), 2)
    full = torch.cat([full, torch.zeros(N, B, 1, H, W)], dim=2)   # zero channel, test_2D.py:208-218
    dev = full.cuda()
    thr = (0.8, 0.7, 0.05)
    res = vb.UncertaintyPipeline(vb.AggregationConfig(patch_size=10, thresholds=thr)).run(
        dev.permute(1, 0, 2, 3, 4), keep_maps=True, mean_argmax=True)      # [B, N, C, H, W] view, no copy
    d = res.to_dicts([f"img{i}" for i in range(B)])
    for i in range(B):
        ref = vo.calculate_uncertainty(full[:, i])
        single = vb.calculate_uncertainty(dev[:, i])                          # per-image call as test_2D.py:245
        for k, key in enumerate(MAPS):
            close(res.maps[k, i].cpu(), ref[key])
            close(single[key].cpu(), ref[key])
            m = res.maps[k, i].cpu().numpy()
            e = d[key][f"img{i}"]
            pl = vo.patch_level_aggregation(m.astype(np.float64), 10)
            assert e["patch_level"]["bounding_box"] == pl["bounding_box"] and len(pl["bounding_box"]) == 2
            close(e["patch_level"]["max_score"], pl["max_score"], 1e-12, 0)
            close(e["image_level"]["max_score"], m.astype(np.float64).sum(), 1e-12, 0)
            close(e["threshold"]["max_score"], float(vo.threshold_aggregation(m, threshold=thr[k])["max_score"]), 1e-6, 0)
        np.testing.assert_array_equal(res.mean_argmax[i].cpu().numpy(), vo.mean_argmax(full[:, i]).numpy())


def test_cfg4_reference_true_shape_and_bf16(vb, vo):
    N, C, H, W = 10, 24, 256, 478                # 24 classes + zero channel at 256x478
    x = torch.cat([stack(44, (N, C, H, W), 1), torch.zeros(N, 1, H, W)], dim=1)
    ref = vo.calculate_uncertainty(x)
    got = vb.calculate_uncertainty(x.cuda())
    for k in MAPS:
        close(got[k].cpu(), ref[k])
    # axis order of the 2D tif path is (W, H) (evaluation/metrics/ace.py:21-23): any 2D shape works
    m = got["pred_entropy"].cpu().numpy().T.copy()
    assert vb.patch_level_aggregation(m, 10) == vo.patch_level_aggregation(m.astype(np.float64), 10) or \
        vb.patch_level_aggregation(m, 10)["bounding_box"] == vo.patch_level_aggregation(m.astype(np.float64), 10)["bounding_box"]
    xb = x.to(torch.bfloat16)
    refb = vo.calculate_uncertainty(xb.float())
    gotb = vb.calculate_uncertainty(xb.cuda())
    for k in MAPS:
        close(gotb[k].cpu(), refb[k], 1e-3, 1e-5)


def test_cfg4_full_size_properties(vb):
    N, C, H, W = 10, 20, 1024, 2048
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.softmax(3 * torch.randn(N, C - 1, H, W, generator=g, device="cuda"), dim=1)
    x = torch.cat([x, torch.zeros(N, 1, H, W, device="cuda")], dim=1)
    r = vb.UncertaintyPipeline(vb.AggregationConfig(patch_size=10, thresholds=(0.0, 0.0, -1.0))).run(
        x.unsqueeze(0), keep_maps=True, mean_argmax=True)
    pe, ee, mi = r.maps[0, 0], r.maps[1, 0], r.maps[2, 0]
    assert torch.equal(mi, pe - ee) and pe.min().item() >= 0 and pe.max().item() <= np.log(19) * (1 + 1e-6)
    assert mi.min().item() >= -1e-6
    assert r.mean_argmax.max().item() <= 18               # the zero channel never wins
    sc = r.scores.cpu().numpy()[0]
    assert sc[0, 2] == H * W
    np.testing.assert_allclose(sc[:, 0], r.maps[:, 0].double().sum(dim=(1, 2)).cpu().numpy(), rtol=1e-9)
    # without the zero channel every map is bit-identical (NaN-skip rule)
    r2 = vb.uncertainty_fused(x[:, :-1].unsqueeze(0))
    assert torch.equal(r2.pred_entropy[0], pe) and torch.equal(r2.mutual_information[0], mi)
    # patch box lies inside the image and its score bounds the mean box
    box = r.scores[0, :, 4:].cpu().numpy()
    assert (box[:, 0] == 0).all() and (box[:, 1] <= H - 10).all() and (box[:, 2] <= W - 10).all()
    assert np.all(sc[:, 3] >= sc[:, 0] / (H * W) * 100 * (1 - 1e-9))


# ----------------------------------------------------------------------------------------- cfg5
def test_cfg5_sharded_sweep_shape(vb, vo):
    B, N, C, S = 6, 16, 4, (48, 48, 48)           # 128^3 at full size: test_gpu_parity.test_full_size_properties
    x = stack(505, (B, N, C) + S, 2)
    thr = (0.9, 0.5, 0.3)
    pipe = vb.UncertaintyPipeline(vb.AggregationConfig(patch_size=10, thresholds=thr, chunk_bytes=3 * 48 ** 3 * 4 * 4))
    whole = pipe.run(x.cuda())
    # volume sharding: every rank's slice gives exactly the rows of the global table
    from values_b200.sharding import shard_range

    for world in (2, 4):
        rows = []
        for rank in range(world):
            lo, hi = shard_range(B, rank, world)
            rows.append(pipe.run(x[lo:hi].cuda()).scores)
        assert torch.equal(torch.cat(rows), whole.scores)
    tab = whole.scores.cpu().numpy()
    for b in (0, B - 1):
        ref = vo.calculate_uncertainty(x[b])
        for k, key in enumerate(MAPS):
            m = ref[key].numpy().astype(np.float64)
            close(tab[b, k, 0], m.sum(), 1e-5, 1e-3)
            th = vo.threshold_aggregation(ref[key].numpy(), threshold=thr[k], mean=False)
            close(tab[b, k, 1], float(th["max_score"]), 2e-5, 1e-2)   # voxels within 1e-6 of the threshold may flip
            close(tab[b, k, 3], vo.patch_level_aggregation(m, 10)["max_score"], 1e-5, 1e-3)
