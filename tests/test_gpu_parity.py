"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same
seeded inputs and against the golden vectors produced by the reference itself.

Tolerances (BASELINE.json north_star): arg-max / threshold counts / patch indices bit-exact;
fp32 maps and scores within 1e-5 relative (+1e-6 absolute for the PE-EE cancellation, SURVEY
H1); bf16 inputs within 1e-3.
"""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
MAPS = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
C2_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "c2_*.npz")))


@pytest.fixture(scope="module")
def vb():
    import values_b200

    return values_b200


@pytest.fixture(scope="module")
def vo():
    from oracle import values_oracle

    return values_oracle


@pytest.fixture(params=[0, 5, 4, 2], ids=["k2b-auto", "k2b-march-nofilter", "k2b-fused", "k2b-tiled"])
def patch_path(request):
    """Every K2b implementation (exact march kernel behind the fp32 TMA strip filter / alone, fused
    tile kernel, generic tiled path) must give the same scores and the same bounding boxes.  The
    implementation is an argument of the call (values_patch_max `path`), not process state."""
    return request.param


def softmax_stack(seed, n, c, spatial, dtype=torch.float32, shared=False, sharp=3.0):
    g = torch.Generator().manual_seed(seed)
    if shared:
        logits = sharp * torch.randn(1, c, *spatial, generator=g, dtype=torch.float64) + \
            0.3 * torch.randn(n, c, *spatial, generator=g, dtype=torch.float64)
    else:
        logits = sharp * torch.randn(n, c, *spatial, generator=g, dtype=torch.float64)
    return torch.softmax(logits, dim=1).to(dtype)


def assert_maps_close(got, ref, rtol=RTOL, atol=ATOL):
    for k in MAPS:
        a, b = got[k].detach().cpu().numpy(), ref[k].detach().cpu().numpy() if hasattr(ref[k], "detach") else ref[k]
        assert a.dtype == np.float32 and a.shape == b.shape, k
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=k)


def assert_argmax(got_u8, stack, exact_ref):
    """bit-exact, except where the two largest class means are within 2 ulp (the reference's
    own mean rounds differently in torch's vector tail vs numpy)."""
    got = got_u8.cpu().numpy()
    ref = np.asarray(exact_ref)
    bad = got != ref
    if bad.any():
        m = np.sort(np.mean(stack.double().numpy(), axis=0), axis=0)
        gap = (m[-1] - m[-2])[bad]
        assert np.all(gap <= 4e-7 * np.abs(m[-1][bad])), f"{bad.sum()} arg-max mismatches beyond near-ties"


def test_native_library_loaded(vb):
    with open("/proc/self/maps") as f:
        assert "libvalues_b200.so" in f.read()
    assert vb._lib.lib.values_abi_version() == vb._lib.ABI_VERSION >= 3


@pytest.mark.parametrize("name", C2_CASES)
def test_c2_golden(vb, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    x = torch.from_numpy(g["softmax"])
    before = vb._lib.launch_count()
    d = vb.calculate_uncertainty(x, ssn=bool(g["ssn"]))
    assert vb._lib.launch_count() > before
    assert all(v.device.type == "cpu" for v in d.values())  # same device as the input
    assert_maps_close(d, {k: g[k] for k in MAPS})
    res = vb.uncertainty_fused(x.cuda().unsqueeze(0), mean_argmax=True, sample_argmax=True)
    assert_argmax(res.mean_argmax[0], x, g["mean_argmax"])
    np.testing.assert_array_equal(res.sample_argmax[0].cpu().numpy(), g["sample_argmax"])


@pytest.mark.parametrize("n,c,spatial,dtype", [
    (5, 2, (64, 64, 64), torch.float32),      # cfg1/2 shape
    (5, 2, (33, 31, 29), torch.float64),      # 3D path dtype, odd sizes -> scalar path
    (16, 4, (40, 40, 40), torch.float32),     # cfg5 class/sample counts
    (10, 20, (96, 128), torch.float32),       # cfg4 class count -> shared-memory class sums
    (10, 25, (61, 77), torch.float32),        # reference-true GTA class count, odd width
    (8, 2, (48, 48, 48), torch.float64),      # cfg3 stitched stack dtype
    (3, 7, (1000,), torch.float32),
    (17, 3, (24, 24, 24), torch.float32),     # N > 16
    (1, 4, (32, 32), torch.float32),
    (6, 8, (20, 20, 20), torch.float64),
    (4, 11, (30, 30), torch.float64),
])
def test_c2_vs_oracle(vb, vo, n, c, spatial, dtype):
    x = softmax_stack(1000 + n * 31 + c, n, c, spatial, dtype)
    x.view(n, c, -1)[0, 0, :5] = 0.0
    ref = vo.calculate_uncertainty(x)
    res = vb.uncertainty_fused(x.cuda().unsqueeze(0), mean_argmax=True, sample_argmax=True)
    assert_maps_close(res.as_dict(0), ref)
    assert_argmax(res.mean_argmax[0], x, np.argmax(np.mean(x.numpy(), axis=0), axis=0))
    np.testing.assert_array_equal(res.sample_argmax[0].cpu().numpy(), vo.sample_argmax(x).numpy())


@pytest.mark.parametrize("n,c,spatial,dtype", [
    (16, 4, (64, 64, 40), torch.float32),   # RS = 8 (two sub-batches of 4); variants 2 / 3: RS = 4 / 8 x 3 stages
    (12, 3, (40, 64, 24), torch.float32),   # RS = 4
    (10, 20, (96, 132), torch.float32),     # RS = 5
    (6, 3, (50, 64), torch.float32),        # RS = 3, ragged last tile
    (3, 4, (24, 40, 20), torch.float32),    # RS = 3, one stage per class
    (14, 3, (40, 52), torch.float32),       # RS = 2
    (7, 5, (40, 52), torch.float32),        # 4-row stages, ragged last stage (4 + 3)
    (17, 4, (24, 40, 20), torch.float32),   # ... 4 x 4 + 1
    (11, 3, (33, 64), torch.bfloat16),      # ... 4 + 4 + 3, bf16
    (1, 6, (40, 52), torch.float32),        # RS = 1
    (8, 4, (33, 64), torch.bfloat16),
    (16, 4, (24, 40, 26), torch.float64),   # fp64: class-outer ring kernel (per-sample accumulators) vs
    (8, 2, (30, 50), torch.float64),        # the sample-outer kernel (variant 4)
    (5, 2, (32, 32, 32), torch.float64),    # the reference's own 3-D case: 5-member ensemble, fp64
    (10, 3, (40, 36), torch.float64),
    (2, 2, (24, 24, 20), torch.float64),    # the other fp64 sample counts with a ring instantiation
    (3, 2, (30, 50), torch.float64),
    (4, 3, (20, 36), torch.float64),
    (6, 2, (24, 24, 24), torch.float64),
    (12, 2, (40, 36), torch.float64),
    (20, 2, (24, 40), torch.float64),
])
def test_k1_kernels_agree(vb, n, c, spatial, dtype):
    """The bulk-copy (TMA ring) kernel and the register-stream kernel share their arithmetic and
    its order: every output must be bit-identical, including the fused fp64 scores, for any
    tiles-per-CTA setting."""
    x = softmax_stack(n * 7 + c, 3 * n, c, spatial).reshape(3, n, c, *spatial).to(dtype).cuda()
    outs = []
    # (variant, tiles per CTA): automatic ring, ring with fixed tiles per CTA, register-stream kernel,
    # 4-row stages, 8-row stages x 3, sample-outer kernel (fp64) -- per-call arguments, no process state
    for variant, it in [(0, 0), (0, 1), (0, 3), (1, 0), (1, 2), (2, 0), (3, 0), (4, 0)]:
        r = vb.uncertainty_fused(x, mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05),
                                 variant=variant, tiles_per_cta=it)
        outs.append((r.pred_entropy, r.expected_entropy, r.mutual_information, r.mean_argmax, r.scores))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


@pytest.mark.parametrize("n,c,spatial,dtype", [
    (16, 4, (31, 33, 29), torch.float32),   # odd voxel count: rows at every 16-byte phase; RS = 4
    (5, 2, (63, 63, 63), torch.float32),    # RS = 5, many tiles, ragged end
    (10, 20, (61, 77), torch.float32),      # RS = 5, C = 20
    (6, 3, (45, 51), torch.float32),        # RS = 2
    (7, 5, (33, 39), torch.float32),        # ragged last stage in the element-strided mode
    (13, 2, (31, 45), torch.float32),
    (3, 2, (7, 5), torch.float32),          # less than one tile
    (8, 4, (33, 61), torch.bfloat16),       # rows 2 bytes off
    (16, 4, (21, 23, 25), torch.float64),   # fp64: rows 8 bytes off
    (8, 2, (31, 49), torch.float64),
    (5, 2, (33, 31, 29), torch.float64),
    (10, 3, (41, 37), torch.float64),
])
def test_k1_unaligned_stacks_take_the_ring(vb, vo, n, c, spatial, dtype):
    """Stacks whose rows are not 16-byte aligned (voxel count not a multiple of the vector) run through the
    bulk-copy ring in its element-strided mode instead of the scalar kernel: maps and arg-max bit-identical
    to the scalar kernel (variant 1) and the sample-outer kernel (variant 4); the fp64 scores group the
    voxels differently (1e-13); one volume against the oracle; stale bytes past a ragged end, negative and
    NaN inputs (the flagged cold paths) included."""
    x = softmax_stack(n * 5 + c, 3 * n, c, spatial).reshape(3, n, c, *spatial).to(dtype)
    flat = x.view(3, n, c, -1)
    if dtype != torch.bfloat16:                    # cold paths: negative, NaN, zero inputs (bf16: the exact
        flat[1, 0, 0, 3] = -0.25                   # recompute uses the polynomial log for every voxel of the
        flat[1, n - 1, c - 1, 17] = float("nan")   # flagged thread, the fast path MUFU.LG2: ownership shows)
    flat[2, 0, 0, -1] = 0.0
    xg = x.cuda()
    outs = []
    for variant, it in [(0, 0), (0, 1), (0, 2), (1, 0), (4, 0)]:
        r = vb.uncertainty_fused(xg, mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05),
                                 variant=variant, tiles_per_cta=it)
        outs.append((r.pred_entropy, r.expected_entropy, r.mutual_information, r.mean_argmax, r.scores))
    for o in outs[1:]:
        for a, b in zip(outs[0][:4], o[:4]):
            assert torch.equal(torch.nan_to_num(a.float(), nan=-7.0), torch.nan_to_num(b.float(), nan=-7.0))
        np.testing.assert_allclose(o[4][0].cpu().numpy(), outs[0][4][0].cpu().numpy(), rtol=1e-13, atol=0)
    ref = vo.calculate_uncertainty(x[0].float() if dtype == torch.bfloat16 else x[0])
    res = vb.uncertainty_fused(xg[:1])
    if dtype == torch.bfloat16:
        assert_maps_close(res.as_dict(0), ref, rtol=1e-3, atol=1e-5)
    else:
        assert_maps_close(res.as_dict(0), ref)
    # an offset base: the same stack as a view that starts one element into its storage
    store = torch.zeros(xg.numel() + 1, dtype=dtype, device="cuda")
    store[1:] = xg.reshape(-1)
    view = store[1:].view_as(xg)
    r2 = vb.uncertainty_fused(view, mean_argmax=True)
    for a, b in zip(outs[0][:4], (r2.pred_entropy, r2.expected_entropy, r2.mutual_information, r2.mean_argmax)):
        assert torch.equal(torch.nan_to_num(a.float(), nan=-7.0), torch.nan_to_num(b.float(), nan=-7.0))


@pytest.mark.parametrize("n,c,spatial,dtype", [
    (16, 4, (40, 40, 40), torch.float32),
    (10, 20, (96, 132), torch.float32),     # cfg4 class count
    (5, 2, (32, 32, 32), torch.float64),    # the 3-D save path: data_carrier_3D.py:253-285
    (8, 2, (24, 40, 26), torch.float64),
    (1, 7, (30, 52), torch.float32),        # one sample: arg-max of the sample == arg-max of the mean
    (3, 1, (16, 24), torch.float32),        # one class
    (7, 5, (33, 64), torch.bfloat16),
    (6, 3, (21, 19, 17), torch.float32),    # unaligned rows: the shared-memory kernel (both variants alike)
])
def test_argmax_only_sweeps(vb, vo, n, c, spatial, dtype):
    """maps=False calls (mean_seg / pred_seg at save time, the per-sample arg-max of test_2D.py:119-127, the
    Dice / GED inputs): the segmented sweeps (k1_argmax_kernel) against the sample-outer shared-memory kernel
    (variant 4) and the oracle -- first maximum wins, NaN counts as maximal, ties and zero channels included."""
    x = softmax_stack(n * 3 + c, 2 * n, c, spatial).reshape(2, n, c, *spatial).to(dtype)
    flat = x.view(2, n, c, -1)
    if c > 1:
        flat[0, 0, 1, 5] = flat[0, 0, 0, 5]                 # a tie: the first class wins
        flat[1, n - 1, c - 1, 9] = float("nan")             # NaN is maximal
        flat[1, :, :, 11] = 0.0                             # all-zero voxel -> class 0
    xg = x.cuda()
    for kw in (dict(mean_argmax=True, sample_argmax=True), dict(sample_argmax=True), dict(mean_argmax=True)):
        new = vb.uncertainty_fused(xg, maps=False, **kw)
        old = vb.uncertainty_fused(xg, maps=False, variant=4, **kw)
        if "sample_argmax" in kw:
            assert torch.equal(new.sample_argmax, old.sample_argmax)
            for b in range(2):
                np.testing.assert_array_equal(new.sample_argmax[b].cpu().numpy(), vo.sample_argmax(x[b].float() if dtype == torch.bfloat16 else x[b]).numpy())
        if "mean_argmax" in kw:
            assert torch.equal(new.mean_argmax, old.mean_argmax)
    withmaps = vb.uncertainty_fused(xg, mean_argmax=True)
    assert torch.equal(withmaps.mean_argmax, vb.uncertainty_fused(xg, maps=False, mean_argmax=True).mean_argmax)
    # a strided [N, B, C, H, W] stack (test_2D.py:317) sliced per image, no copy
    perm = xg.permute(1, 0, 2, *range(3, xg.dim())).contiguous().permute(1, 0, 2, *range(3, xg.dim()))
    r = vb.uncertainty_fused(perm, maps=False, mean_argmax=True, sample_argmax=True)
    ref = vb.uncertainty_fused(xg, maps=False, mean_argmax=True, sample_argmax=True, variant=4)
    assert torch.equal(r.sample_argmax, ref.sample_argmax) and torch.equal(r.mean_argmax, ref.mean_argmax)


def test_fp64_class_mean_is_true_division(vb):
    """The fp64 kernels form the class mean as S * RN(1/N) refined by two fmas instead of a division: the
    arg-max of the mean and PE must be those of true division for every magnitude (raw overlap sums can be
    large), including the exponent ranges that fall back to a real division."""
    g = torch.Generator().manual_seed(5)
    for n in (5, 8, 10, 16, 3, 7):
        for scale in (1.0, 1e-150, 1e150, 1e-290, 3e300 / n):
            x = torch.rand(n, 3, 2048, generator=g, dtype=torch.float64) * scale
            res = vb.uncertainty_fused(x.cuda().unsqueeze(0), mean_argmax=True)
            mean = x.sum(dim=0) / n if False else torch.stack([sum(x[i, c] for i in range(n)) for c in range(3)]) / n
            assert torch.equal(res.mean_argmax[0].cpu().to(torch.int64), mean.argmax(dim=0)), (n, scale)
            t = mean * torch.log(mean)
            pe = torch.zeros(2048, dtype=torch.float32)
            for c in range(3):                         # fp32 accumulator, fp64 add, NaN terms skipped
                ok = ~torch.isnan(t[c])
                pe = torch.where(ok, (pe.double() + t[c]).float(), pe)
            np.testing.assert_allclose(res.pred_entropy[0].cpu().numpy(), (-pe).numpy(), rtol=2e-7, atol=0)


def test_c2_low_mi_regime(vb, vo):
    x = softmax_stack(5, 5, 2, (48, 48, 48), torch.float32, shared=True)
    ref = vo.calculate_uncertainty(x)
    got = vb.calculate_uncertainty(x.cuda())
    assert got["pred_entropy"].device.type == "cuda"
    assert_maps_close(got, ref)
    # PE and EE themselves (no cancellation) hold 1e-5 relative with a tiny absolute floor
    for k in MAPS[:2]:
        np.testing.assert_allclose(got[k].cpu().numpy(), ref[k].numpy(), rtol=1e-5, atol=1e-9)


def test_c2_bf16(vb, vo):
    x = softmax_stack(9, 8, 4, (32, 32, 32), torch.float32).to(torch.bfloat16)
    ref = vo.calculate_uncertainty(x.float())  # bf16 is widened exactly, then the fp32 algorithm
    res = vb.uncertainty_fused(x.cuda().unsqueeze(0), mean_argmax=True)
    assert_maps_close(res.as_dict(0), ref, rtol=1e-3, atol=1e-5)
    assert_argmax(res.mean_argmax[0], x.float(), vo.mean_argmax(x.float()).numpy())
    x = softmax_stack(10, 6, 19, (40, 56), torch.float32).to(torch.bfloat16)  # smem path
    ref = vo.calculate_uncertainty(x.float())
    assert_maps_close(vb.calculate_uncertainty(x.cuda()), ref, rtol=1e-3, atol=1e-5)


def test_c2_strided_2d_layout_and_ssn(vb, vo):
    """[N, B, C, H, W] stack sliced per image as test_2D.py:223-227 does (strided in N)."""
    n, b, c, h, w = 6, 3, 5, 24, 40
    full = softmax_stack(77, n * b, c, (h, w)).reshape(n, b, c, h, w)
    full = torch.cat([full, torch.zeros(n, b, 1, h, w)], dim=2)  # zero channel, test_2D.py:208-218
    dev = full.cuda()
    res = vb.uncertainty_fused(dev.permute(1, 0, 2, 3, 4))  # batched, no copy
    for i in range(b):
        ref = vo.calculate_uncertainty(full[:, i], ssn=True)
        assert_maps_close(vb.calculate_uncertainty(dev[:, i], ssn=True), ref)
        assert_maps_close(res.as_dict(i, ssn=True), ref)


def test_c2_properties(vb):
    x = softmax_stack(3, 10, 4, (64, 64)).cuda()
    base = vb.calculate_uncertainty(x)
    # appended all-zero channel leaves every map bit-identical (NaN-skip rule)
    xz = torch.cat([x, torch.zeros(10, 1, 64, 64, device="cuda")], dim=1)
    withz = vb.calculate_uncertainty(xz)
    for k in MAPS:
        assert torch.equal(base[k], withz[k])
    # identical samples -> MI == 0 up to one rounding of PE
    same = x[:1].expand(10, -1, -1, -1).contiguous()
    d = vb.calculate_uncertainty(same)
    assert d["epistemic_uncertainty"].abs().max().item() <= 5e-7
    # uniform p -> PE = EE = log C
    u = torch.full((4, 5, 16, 16), 0.2, device="cuda")
    d = vb.calculate_uncertainty(u)
    np.testing.assert_allclose(d["pred_entropy"].cpu().numpy(), np.log(5.0), rtol=1e-6)
    # one-hot samples cycling over K classes -> PE = log K, EE = 0 exactly
    oh = torch.zeros(4, 4, 8, 8, device="cuda")
    for i in range(4):
        oh[i, i] = 1.0
    d = vb.calculate_uncertainty(oh)
    np.testing.assert_allclose(d["pred_entropy"].cpu().numpy(), np.log(4.0), rtol=1e-6)
    assert d["aleatoric_uncertainty"].abs().max().item() == 0.0
    # empty volume
    e = vb.uncertainty_fused(torch.zeros(1, 3, 2, 0, 4, device="cuda"))
    assert e.pred_entropy.shape == (1, 0, 4)


@pytest.mark.parametrize("name", ["msr_f32", "msr_f64"])
def test_msr_golden(vb, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = vb.calculate_one_minus_msr(torch.from_numpy(g["softmax"]))
    np.testing.assert_array_equal(d["pred_entropy"].numpy(), g["pred_entropy"])  # bit-exact


def test_fused_scores_vs_oracle(vb, vo):
    B, n, c, spatial = 3, 5, 3, (20, 24, 28)
    x = softmax_stack(21, B * n, c, spatial).reshape(B, n, c, *spatial)
    thr = (0.5, 0.4, 0.05)
    res = vb.uncertainty_fused(x.cuda(), scores=True, thresholds=thr)
    sc = res.scores.cpu().numpy()
    for b in range(B):
        ref = vo.calculate_uncertainty(x[b])
        for k, key in enumerate(MAPS):
            got_map = res.as_dict(b)[key].cpu().numpy().astype(np.float64)
            # counts are exact w.r.t. the map this kernel produced
            assert sc[b, k, 2] == float((got_map >= thr[k]).sum())
            np.testing.assert_allclose(sc[b, k, 0], vo.image_level_aggregation(ref[key].numpy())["max_score"], rtol=1e-5)
            np.testing.assert_allclose(sc[b, k, 0], got_map.sum(), rtol=1e-12)
            np.testing.assert_allclose(sc[b, k, 1], got_map[got_map >= thr[k]].sum(), rtol=1e-12)


# ------------------------------------------------------------------------------------ C3
def test_c3_golden(vb, patch_path):
    g = np.load(os.path.join(GOLDEN, "c3_aggregations.npz"))
    for i in range(int(g["n_patch_cases"])):
        for mean in (0, 1):
            key = f"patch_{i}_{mean}"
            m = g["map_" + str(g[key + "_map"])]
            p = g[key + "_patch"].tolist()
            p = p[0] if len(p) == 1 else p
            r = vb.patch_level_aggregation(m, p, mean=bool(mean), _k2b_path=patch_path)
            rtol = 1e-12 if m.dtype == np.float64 else 1e-6  # reference FFT runs fp32 images in complex64
            np.testing.assert_allclose(r["max_score"], float(g[key + "_score"]), rtol=rtol, atol=1e-300)
            assert [list(b) for b in r["bounding_box"]] == g[key + "_bbox"].tolist(), key
            assert all(isinstance(v, int) for b in r["bounding_box"] for v in b)
    for mname in ("m3d_f64", "m2d_f32", "m3d_zero", "m3d_planted"):
        m = g["map_" + mname]
        rt = 1e-12 if m.dtype == np.float64 else 1e-6
        np.testing.assert_allclose(vb.image_level_aggregation(m)["max_score"], float(g["image_sum_" + mname]), rtol=rt)
        np.testing.assert_allclose(vb.image_level_aggregation(m, mean=True), float(g["image_mean_" + mname]), rtol=rt)
        for j, thr in enumerate(g["thresholds"].tolist()):
            for mean in (1, 0):
                r = vb.threshold_aggregation(m, threshold=thr, mean=bool(mean))
                np.testing.assert_allclose(r["max_score"], float(g[f"thr_{mname}_{j}_{mean}"]), rtol=rt)
                assert r["threshold"] == thr and isinstance(r["max_score"], float)


@pytest.mark.parametrize("shape,patch,dtype", [
    ((64, 64, 64), 10, np.float64), ((70, 45, 33), 10, np.float32), ((128, 96), 10, np.float32),
    ((40, 50, 60), [3, 7, 12], np.float64), ((10, 10, 10), 10, np.float64), ((300,), 10, np.float32),
    ((35, 200, 37), [20, 1, 30], np.float32), ((256, 478), 10, np.float32),
    ((20, 30, 100), [3, 4, 40], np.float64),   # patch wider than a warp -> tiled shared-memory fallback
    ((9, 70, 33), [9, 33, 32], np.float32), ((1024, 2048), 10, np.float32),
    ((300, 500), 10, np.float32), ((75, 260), 10, np.float32),     # 2-D: the strip filter marches along y
    ((40, 45, 252), 10, np.float32),                               # three x tiles of the strip filter
])
def test_c3_patch_vs_oracle(vb, vo, shape, patch, dtype, patch_path):
    rng = np.random.default_rng(sum(shape) * 7 + len(shape))
    m = rng.random(shape).astype(dtype)
    for mean in (False, True):
        a = vb.patch_level_aggregation(m, patch, mean=mean, _k2b_path=patch_path)
        b = vo.patch_level_aggregation(m, patch, mean=mean)
        assert a["bounding_box"] == b["bounding_box"]
        np.testing.assert_allclose(a["max_score"], b["max_score"], rtol=1e-12)


def test_c3_batched_and_isclose_rule(vb, vo, patch_path):
    rng = np.random.default_rng(11)
    maps = rng.random((7, 30, 31, 32)).astype(np.float32)
    maps[3] = 0.0                                   # all-zero map -> score 0, box at origin
    maps[4, 5:15, 5:15, 5:15] += 1.0
    maps[5] = maps[4] * (1 - 4e-6)                  # everything scales: same bbox
    score, bbox = vb.patch_max(torch.from_numpy(maps).cuda(), 10, path=patch_path)
    for i in range(7):
        r = vo.patch_level_aggregation(maps[i], 10)
        np.testing.assert_allclose(score[i].item(), r["max_score"], rtol=1e-12)
        assert bbox[i].tolist() == [b[0] for b in r["bounding_box"]]
    assert bbox[3].tolist() == [0, 0, 0] and score[3].item() == 0.0


@pytest.mark.parametrize("M,shape", [(1, (200, 300)), (20, (200, 300)), (7, (530, 136)), (40, (64, 64))])
def test_c3_2d_batches_march_along_y(vb, vo, M, shape):
    """2-D images go through the y-march of the strip filter; how many CTA rows a CTA walks depends on the
    number of maps in the call.  Same results as the exact march alone and as the oracle, with values on both
    sides of 2.0 (the exponent fields of 1.x and 2.x once fooled the filter's bound on max |input|)."""
    rng = np.random.default_rng(M * 1000 + shape[0])
    maps = (rng.random((M,) + shape) * 3.0).astype(np.float32)
    maps[M // 2, shape[0] // 3:shape[0] // 3 + 10, 5:15] += 0.5
    t = torch.from_numpy(maps).cuda()
    s0, b0 = vb.patch_max(t, 10, path=0)
    s5, b5 = vb.patch_max(t, 10, path=5)
    np.testing.assert_allclose(s0.cpu().numpy(), s5.cpu().numpy(), rtol=1e-13, atol=0)
    assert torch.equal(b0, b5)
    for i in range(0, M, max(1, M // 4)):
        r = vo.patch_level_aggregation(maps[i], 10)
        np.testing.assert_allclose(s0[i].item(), r["max_score"], rtol=1e-12)
        assert b0[i].tolist() == [b[0] for b in r["bounding_box"]]


def _filter_cases():
    """fp32 maps at march-kernel sizes that stress the fp32 filter pass (it may only ever DROP
    sub-chunks that cannot hold a window np.isclose to the maximum)."""
    rng = np.random.default_rng(2024)
    shape = (60, 70, 140)
    cases = {}
    base = rng.random(shape).astype(np.float32)
    cases["random"] = base
    cases["constant"] = np.full(shape, 0.37, np.float32)          # every window ties: list overflows
    cases["zeros"] = np.zeros(shape, np.float32)
    m = np.zeros(shape, np.float32)
    m[3:13, 5:15, 100:110] = 0.5
    m[40:50, 50:60, 7:17] = 0.5 * (1 + 4e-6)                      # later in C order, 4e-6 larger: the
    cases["near_tie_close"] = m                                   # first block is still isclose -> wins
    m = m.copy()
    m[40:50, 50:60, 7:17] = 0.5 * (1 + 5e-5)                      # 5e-5 larger: not isclose -> second wins
    cases["near_tie_far"] = m
    cases["signed"] = (base - 0.5).astype(np.float32)             # negative values (MI can dip below 0)
    m = (base * 1e-3).astype(np.float32)
    m[20, 30, 40] = 1e4                                           # one huge voxel: large error bound
    cases["spike"] = m
    cases["huge"] = (base * 1e36).astype(np.float32)              # fp32 box sums overflow -> exact pass
    m = base.copy()
    m[30:50, 20:40, 60:90] = 0.0                                  # sparse blobs, as real maps
    m[m < 0.9] = 0.0
    cases["sparse"] = m
    cases["smooth"] = (np.add.outer(np.add.outer(np.hanning(60), np.hanning(70)), np.hanning(140)) / 3).astype(np.float32)
    return cases


@pytest.mark.parametrize("mean", [False, True])
def test_c3_filter_pass_is_exact(vb, vo, mean):
    """The fp32 filter + listed exact pass (path 0) must give the same boxes and scores as the exact
    march over every tile (path 5) -- scores to fp64 rounding, because a listed sub-chunk restarts
    its z-window sums -- and both must match the oracle."""
    cases = _filter_cases()
    names = list(cases)
    maps = torch.from_numpy(np.stack([cases[k] for k in names])).cuda()
    res = {}
    for path in (0, 5):
        s, b = vb.patch_max(maps, 10, mean=mean, path=path)
        res[path] = (s.cpu().numpy(), b.cpu().numpy())
    for path in (0,):
        np.testing.assert_allclose(res[path][0], res[5][0], rtol=1e-13, atol=0)
        assert np.array_equal(res[path][1], res[5][1])
    for i, k in enumerate(names):
        r = vo.patch_level_aggregation(cases[k], 10, mean=mean)
        np.testing.assert_allclose(res[0][0][i], r["max_score"], rtol=1e-12, err_msg=k)
        assert res[0][1][i].tolist() == [b[0] for b in r["bounding_box"]], k


@pytest.mark.parametrize("M,shape", [(5, (23, 50, 77)), (3, (41, 43, 127)), (4, (20, 30, 61)), (6, (130, 301)),
                                     (2, (256, 478)), (3, (33, 35, 254))])
def test_c3_unaligned_rows_take_the_pitched_copy(vb, vo, M, shape):
    """fp32 maps whose innermost extent is not a multiple of 4 (no tensor map can describe their rows) are
    copied into a scratch with pitched rows and go through the strip filter + listed passes there: same boxes
    and scores as the exact march on the maps themselves (path 5) and as the oracle; negative values, a
    planted near-tie in the last columns, a batch taken as a strided view."""
    rng = np.random.default_rng(sum(shape) + M)
    host = (rng.random((M + 1,) + shape) * 2.5 - 0.5).astype(np.float32)
    lo = tuple(s - 10 for s in shape)
    sl = tuple(slice(l, l + 10) for l in lo)
    host[1][sl] += 3.0                                   # the maximum sits against the far corner (last columns)
    host[2] = host[1] * (1 - 4e-6)
    store = torch.from_numpy(host).cuda()
    maps = store[:M] if M % 2 else store[1:M + 1]        # even M: a view that starts one map into the storage
    href = host[:M] if M % 2 else host[1:M + 1]
    s0, b0 = vb.patch_max(maps, 10, path=0)
    s5, b5 = vb.patch_max(maps, 10, path=5)
    np.testing.assert_allclose(s0.cpu().numpy(), s5.cpu().numpy(), rtol=1e-13, atol=0)
    assert torch.equal(b0, b5)
    for i in range(M):
        r = vo.patch_level_aggregation(href[i], 10)
        np.testing.assert_allclose(s0[i].item(), r["max_score"], rtol=1e-12)
        assert b0[i].tolist()[-len(shape):] == [b[0] for b in r["bounding_box"]]


@pytest.mark.parametrize("p0", [1, 4, 16, 33])
def test_c3_filter_pass_other_window_depths(vb, vo, p0):
    """The march kernel and its filter fix only the in-plane window (10 x 10); the window depth p0 is
    a run-time value (z-slide length, leaving-plane offset, error coefficient)."""
    rng = np.random.default_rng(40 + p0)
    maps = rng.random((3, 33, 52, 88)).astype(np.float32)
    maps[1, 10:20, 30:40, 60:70] += 0.25
    t = torch.from_numpy(maps).cuda()
    res = {}
    for path in (0, 5):
        s, b = vb.patch_max(t, [p0, 10, 10], path=path)
        res[path] = (s.cpu().numpy(), b.cpu().numpy())
    for path in (0,):
        np.testing.assert_allclose(res[path][0], res[5][0], rtol=1e-13, atol=0)
        assert np.array_equal(res[path][1], res[5][1])
    for i in range(3):
        r = vo.patch_level_aggregation(maps[i], [p0, 10, 10])
        np.testing.assert_allclose(res[0][0][i], r["max_score"], rtol=1e-12)
        assert res[0][1][i].tolist() == [b[0] for b in r["bounding_box"]]


def test_c3_filter_pass_nonfinite(vb):
    """NaN / inf voxels send the map through the exact pass: same result as without the filter."""
    rng = np.random.default_rng(7)
    maps = rng.random((4, 40, 64, 96)).astype(np.float32)
    maps[1, 20, 30, 50] = np.nan
    maps[2, 5, 6, 7] = np.inf
    maps[3, 39, 63, 95] = -np.inf
    t = torch.from_numpy(maps).cuda()
    res = {}
    for path in (0, 5):
        s, b = vb.patch_max(t, 10, path=path)
        res[path] = (s.cpu().numpy(), b.cpu().numpy())
    for path in (0,):
        np.testing.assert_allclose(res[path][0], res[5][0], rtol=1e-13, atol=0, equal_nan=True)
        assert np.array_equal(res[path][1], res[5][1])
    assert np.isnan(res[0][0][1]) and res[0][1][1].tolist() == [-1, -1, -1]


@pytest.mark.parametrize("shape", [(37, 50, 70), (128, 64, 96), (75, 128, 128)])
def test_c3_workspace_garbage_is_harmless(vb, vo, shape, patch_path):
    """Every workspace entry the finish pass reads must have been written by this call: a workspace
    full of +inf / NaN bit patterns (what a reused scratch buffer may hold) must not leak into the
    scores -- z-chunks whose last sub-chunk is short are the case that once did."""
    rng = np.random.default_rng(shape[0])
    maps = torch.from_numpy(rng.random((5,) + shape).astype(np.float32)).cuda()
    ws_bytes = vb.aggregation.patch_max_workspace_bytes(5, shape, 10, path=patch_path)
    for fill in (0x7f, 0xff):
        ws = torch.full((ws_bytes + 64,), fill, dtype=torch.uint8, device="cuda")
        score, bbox = vb.patch_max(maps, 10, workspace=ws, path=patch_path)
        for i in range(5):
            r = vo.patch_level_aggregation(maps[i].cpu().numpy(), 10)
            np.testing.assert_allclose(score[i].item(), r["max_score"], rtol=1e-12)
            assert bbox[i].tolist() == [b[0] for b in r["bounding_box"]]


def test_c3_errors(vb):
    img = np.ones((8, 8))
    with pytest.raises(ValueError):
        vb.patch_level_aggregation(img, 9)
    with pytest.raises(Exception, match="A threshold needs to be provided"):
        vb.threshold_aggregation(img)
    r = vb.threshold_aggregation(img.astype(np.float32), threshold=2.0)  # empty selection
    assert r == {"max_score": 0.0, "threshold": 2.0}
    assert vb.image_level_aggregation(np.ones((33, 35, 37)))["max_score"] == 33 * 35 * 37
    # extra kwargs the reference always passes (aggregate_uncertainties.py:81-86)
    vb.patch_level_aggregation(img, 4, pred_model="Softmax", unc_type="pred_entropy")
    vb.image_level_aggregation(img, pred_model="Softmax", unc_type="pred_entropy")


def test_c3_map_reduce_vs_oracle(vb, vo):
    rng = np.random.default_rng(5)
    for shape, dtype in [((64, 64, 64), np.float64), ((1024, 513), np.float32), ((7,), np.float32)]:
        m = rng.random(shape).astype(dtype)
        np.testing.assert_allclose(vb.image_level_aggregation(m)["max_score"],
                                   vo.image_level_aggregation(m)["max_score"], rtol=1e-6)
        for thr in (0.0, 0.3, 0.999, 1.5):
            a = vb.threshold_aggregation(m, threshold=thr)
            b = vo.threshold_aggregation(m, threshold=thr)
            np.testing.assert_allclose(a["max_score"], float(b["max_score"]), rtol=1e-6)
        out = vb.map_reduce(torch.from_numpy(m).cuda().unsqueeze(0), [0.3])[0].tolist()
        assert out[2] == float((m >= 0.3).sum())  # count bit-exact


# --------------------------------------------------------------------------------- stitch
@pytest.fixture(params=[0, 1, 2], ids=["k3-auto", "k3-scalar", "k3-vector"])
def stitch_path(request):
    """The box kernel (tensor-map copies through a shared-memory ring; the default when rows are 16-byte
    aligned), the register-staged vector kernel and the scalar kernel must all be bit-exact."""
    return request.param


def test_stitch_golden(vb, vo, stitch_path):
    g = np.load(os.path.join(GOLDEN, "stitch_3d.npz"))
    shape, p = tuple(g["shape"].tolist()), int(g["patch"])
    crops = vb.patch_grid(shape, p, float(g["overlap"]))
    assert np.array(crops).tolist() == g["crops"].tolist()
    patches = torch.from_numpy(g["patches"])
    n_pred = patches.shape[0]
    # (a) drop-in DataCarrier3D.concat_data, batched like the reference's loop
    carrier = vb.DataCarrier3D(stitch_path=stitch_path)
    for pred_idx in range(n_pred):
        for s in range(0, len(crops), 5):
            idx = list(range(s, min(s + 5, len(crops))))
            batch = {"image_paths": ["vol_a.npy"] * len(idx), "label_paths": [["lab"]] * len(idx),
                     "org_image_size": [shape] * len(idx), "crop_idx": [crops[i] for i in idx],
                     "data": torch.ones(len(idx), 1, p, p, p),
                     "seg": torch.ones(1, len(idx), p, p, p, dtype=torch.int32)}
            carrier.concat_data(batch, patches[pred_idx, idx], n_pred=n_pred, pred_idx=pred_idx)
    v = carrier.data["vol_a.npy"]
    np.testing.assert_array_equal(v["softmax_pred"].cpu().numpy(), g["softmax_sum"])  # bit-exact fp64
    np.testing.assert_array_equal(v["num_predictions"].cpu().numpy(), g["num_predictions"])
    np.testing.assert_array_equal(v["data"].cpu().numpy(), g["num_predictions"][0])
    np.testing.assert_array_equal(v["seg"][0].cpu().numpy(), g["num_predictions"][0].astype(np.int32))
    vb.caculcate_uncertainty_multiple_pred(carrier)
    assert_maps_close({k: v[k] for k in MAPS}, {k: g[k] for k in MAPS})
    norm = carrier.normalized("vol_a.npy")
    np.testing.assert_allclose(norm["pred_entropy"].cpu().numpy(), g["pred_entropy_saved"], rtol=RTOL, atol=ATOL)
    assert_argmax(norm["mean_seg"], torch.from_numpy(g["softmax_sum"] / np.clip(g["num_predictions"], 1, None)), g["mean_seg"])
    ref_layout = carrier.numpy_data()["vol_a.npy"]
    assert ref_layout["softmax_pred"].dtype == np.float64 and ref_layout["num_predictions"].shape == (2,) + shape
    # (b) all patches at once, written exactly once
    total, cnt = vb.stitch_volume(patches.cuda(), crops, shape, path=stitch_path)
    np.testing.assert_array_equal(total.cpu().numpy(), g["softmax_sum"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["num_predictions"][0])


@pytest.mark.parametrize("shape,p,overlap,dtype", [
    ((64, 64, 64), 64, 1.0, torch.float64),      # shipped settings: one patch per volume
    ((96, 80, 72), 32, 0.5, torch.float32),
    ((50, 50, 40), 16, 0.75, torch.float32),     # non-multiple size -> uncovered remainder
    ((40, 24, 200), 8, 1.0, torch.bfloat16),
    ((24, 40, 256), 16, 0.5, torch.float32),     # full-width vector tiles (64 threads x 4 voxels)
    ((20, 24, 136), 12, 0.5, torch.float64),     # stride 6: z origins not multiples of 4
    ((33, 31, 45), 16, 0.5, torch.float32),      # Z % 4 != 0: rows are not whole groups of four voxels
    ((24, 40, 70), 8, 1.0, torch.float32),       # ... Z % 4 == 2, fully covered along z up to 64
    ((17, 19, 23), 8, 0.5, torch.float64),
    ((20, 24, 27), 12, 0.5, torch.bfloat16),     # ... and z origins the copy engine cannot fetch
])
def test_stitch_vs_oracle(vb, vo, shape, p, overlap, dtype, stitch_path):
    crops = vo.patch_grid(shape, p, overlap)
    assert crops == vb.patch_grid(shape, p, overlap)
    n_pred, c = 2, 3
    g = torch.Generator().manual_seed(len(crops))
    patches = torch.rand(n_pred, len(crops), c, p, p, p, generator=g, dtype=torch.float64).to(dtype)
    st = vo.StitchOracle(n_classes=c)
    for pi in range(n_pred):
        st.concat_data({"image_paths": ["v"] * len(crops), "org_image_size": [shape] * len(crops),
                        "crop_idx": crops}, patches[pi].double(), n_pred=n_pred, pred_idx=pi)
    total, cnt = vb.stitch_volume(patches.cuda(), crops, shape, path=stitch_path)
    np.testing.assert_array_equal(total.cpu().numpy(), st.data["v"]["softmax_pred"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), st.data["v"]["num_predictions"][0])
    if dtype == torch.float32:
        t32, _ = vb.stitch_volume(patches.cuda(), crops, shape, out_dtype=torch.float32, path=stitch_path)
        np.testing.assert_allclose(t32.cpu().numpy(), st.data["v"]["softmax_pred"], rtol=1e-6)
    # a second pass on top of the sums (read-modify-write of every voxel): the oracle adds the patches again;
    # and the separable-weight form agrees with the weight map (Z % 4 != 0: the same element-wise epilogue)
    lo = vb.stitching.crops_to_lo(crops, "cuda")
    vb.stitch_accumulate(patches.cuda(), lo, total, cnt, accumulate=True, path=stitch_path)
    for pi in range(n_pred):
        st.concat_data({"image_paths": ["v"] * len(crops), "org_image_size": [shape] * len(crops),
                        "crop_idx": crops}, patches[pi].double(), n_pred=n_pred, pred_idx=pi)
    np.testing.assert_array_equal(total.cpu().numpy(), st.data["v"]["softmax_pred"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), st.data["v"]["num_predictions"][0])
    if shape[2] % 4 and p <= 128:
        fac = vb.gaussian_importance_factors((p, p, p), sigma_scale=0.25)
        wa = vb.stitch_volume(patches.cuda(), crops, shape, weight=tuple(f.cuda() for f in fac), path=stitch_path)
        wb = vb.stitch_volume(patches.cuda(), crops, shape, weight=vb.importance_map_from_factors(fac).cuda(), path=1)
        assert torch.equal(wa[0], wb[0]) and torch.equal(wa[1], wb[1])


def test_stitch_selected_patches_of_a_padded_stack(vb, vo, stitch_path):
    """A subset of the patches (patch_index), in a different order than they are stored, out of a stack whose
    sample stride is larger than patches x patch size (a view of a bigger tensor): the box kernel addresses
    (sample, patch) rows through one merged tensor-map dimension."""
    shape, p = (48, 32, 64), 16
    crops = vo.patch_grid(shape, p, 0.5)
    g = torch.Generator().manual_seed(8)
    store = torch.rand(3, len(crops) + 5, 2, p, p, p, generator=g, dtype=torch.float32).cuda()
    patches = store[:, 2:2 + len(crops)]                      # stride_n = (P + 5) * stride_p, offset base
    sel = [i for i in range(len(crops)) if i % 3 != 1][::-1]   # a subset, reversed: summation order = list order
    lo = vb.stitching.crops_to_lo([crops[i] for i in sel], "cuda")
    out = torch.empty((3, 2) + shape, dtype=torch.float64, device="cuda")
    cnt = torch.empty(shape, dtype=torch.float64, device="cuda")
    vb.stitch_accumulate(patches, lo, out, cnt, patch_index=torch.tensor(sel, dtype=torch.int32, device="cuda"),
                         accumulate=False, path=stitch_path)
    want = np.zeros((3, 2) + shape)
    want_cnt = np.zeros(shape)
    ph = patches.cpu().numpy().astype(np.float64)
    for i in sel:
        (x0, x1), (y0, y1), (z0, z1) = crops[i]
        want[:, :, x0:x1, y0:y1, z0:z1] += ph[:, i]
        want_cnt[x0:x1, y0:y1, z0:z1] += 1
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    np.testing.assert_array_equal(cnt.cpu().numpy(), want_cnt)
    # a second call accumulates on top (read-modify-write of every voxel)
    vb.stitch_accumulate(patches, lo, out, cnt, patch_index=torch.tensor(sel, dtype=torch.int32, device="cuda"),
                         accumulate=True, path=stitch_path)
    for i in sel:                                              # the same slabs added again, in list order
        (x0, x1), (y0, y1), (z0, z1) = crops[i]
        want[:, :, x0:x1, y0:y1, z0:z1] += ph[:, i]
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    np.testing.assert_array_equal(cnt.cpu().numpy(), 2 * want_cnt)


def test_stitch_one_sample_with_patch_index(vb, vo, stitch_path):
    """DataCarrier3D's form (data_carrier_3D.py:154-162 per image): ONE sample and a patch_index into a larger
    batch -- the row count behind the pointer is unknown to the library, so the box kernel's tensor map takes
    its row extent from the allocation (tma_host.cuh, bytes_to_allocation_end).  The last selected row is the
    last row of the tensor."""
    shape, p = (32, 32, 32), 16
    crops = vo.patch_grid(shape, p, 0.5)
    g = torch.Generator().manual_seed(21)
    batch = torch.rand(len(crops) + 3, 2, p, p, p, generator=g, dtype=torch.float32).cuda()
    sel = list(range(3, len(crops) + 3))[::-1]                 # ends at the last row of the batch
    lo = vb.stitching.crops_to_lo(crops[::-1], "cuda")
    out = torch.zeros((1, 2) + shape, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(shape, dtype=torch.float64, device="cuda")
    vb.stitch_accumulate(batch.unsqueeze(0), lo, out, cnt, patch_index=torch.tensor(sel, dtype=torch.int32, device="cuda"),
                         accumulate=True, path=stitch_path)
    want = np.zeros((2,) + shape)
    want_cnt = np.zeros(shape)
    bh = batch.cpu().numpy().astype(np.float64)
    for i, crop in zip(sel, crops[::-1]):
        (x0, x1), (y0, y1), (z0, z1) = crop
        want[:, x0:x1, y0:y1, z0:z1] += bh[i]
        want_cnt[x0:x1, y0:y1, z0:z1] += 1
    np.testing.assert_array_equal(out[0].cpu().numpy(), want)
    np.testing.assert_array_equal(cnt.cpu().numpy(), want_cnt)


def test_stitch_a_batch_at_a_time_equals_one_call(vb, vo, stitch_path):
    """DataCarrier3D.concat_data's pattern (data_carrier_3D.py:99-179): the patches of a volume arrive a few at
    a time and are accumulated on top of the sums; boxes a call does not reach must keep theirs (the box
    kernel skips them instead of reading and re-storing the whole volume)."""
    shape, p = (72, 56, 64), 16
    crops = vo.patch_grid(shape, p, 0.5)
    g = torch.Generator().manual_seed(31)
    patches = torch.rand(2, len(crops), 3, p, p, p, generator=g, dtype=torch.float32).cuda()
    lo = vb.stitching.crops_to_lo(crops, "cuda")
    want, want_cnt = vb.stitch_volume(patches, crops, shape, path=stitch_path)
    out = torch.zeros_like(want)
    cnt = torch.zeros_like(want_cnt)
    order = torch.randperm(len(crops), generator=g).tolist()
    order.sort()                                    # list order = summation order: keep it
    for i in range(0, len(order), 3):
        sel = order[i:i + 3]
        vb.stitch_accumulate(patches, lo[sel], out, cnt, patch_index=torch.tensor(sel, dtype=torch.int32, device="cuda"),
                             accumulate=True, path=stitch_path)
    assert torch.equal(out, want) and torch.equal(cnt, want_cnt)


def test_stitch_many_patches_chunked_list(vb, vo, stitch_path):
    shape, p = (44, 44, 44), 8
    crops = vo.patch_grid(shape, p, 0.25)  # stride 2 -> 19^3 = 6859 patches > list chunk
    g = torch.Generator().manual_seed(1)
    patches = torch.rand(1, len(crops), 1, p, p, p, generator=g, dtype=torch.float32)
    st = vo.StitchOracle(n_classes=1)
    st.concat_data({"image_paths": ["v"] * len(crops), "org_image_size": [shape] * len(crops),
                    "crop_idx": crops}, patches[0], pred_idx=0)
    total, cnt = vb.stitch_volume(patches.cuda(), crops, shape, path=stitch_path)
    np.testing.assert_allclose(total.cpu().numpy(), st.data["v"]["softmax_pred"], rtol=1e-14)
    np.testing.assert_array_equal(cnt.cpu().numpy(), st.data["v"]["num_predictions"][0])


def test_stitch_gaussian_weighted(vb, vo, stitch_path):
    """Opt-in Gaussian-weighted stitching (north_star; the reference is uniform): bit-exact against
    the numpy restatement `sum += w * patch`, `count += w`, and normalisation by the weight sum."""
    shape, p = (40, 36, 72), 16
    crops = vb.patch_grid(shape, p, 0.5)
    g = torch.Generator().manual_seed(4)
    patches = torch.rand(2, len(crops), 2, p, p, p, generator=g, dtype=torch.float32)
    w = vb.gaussian_importance_map((p, p, p))
    np.testing.assert_array_equal(w.numpy(), vo.gaussian_importance_map((p, p, p)))
    total, cnt = vb.stitch_volume(patches.cuda(), crops, shape, weight=w.cuda(), path=stitch_path)
    ref_total, ref_cnt = vo.stitch_weighted(patches.numpy(), crops, shape, w.numpy())
    np.testing.assert_array_equal(total.cpu().numpy(), ref_total)
    np.testing.assert_array_equal(cnt.cpu().numpy(), ref_cnt)
    # uniform weight of ones == the reference's unweighted accumulator, bit for bit
    t1, c1 = vb.stitch_volume(patches.cuda(), crops, shape, weight=torch.ones(p, p, p, dtype=torch.float64).cuda(), path=stitch_path)
    t0, c0 = vb.stitch_volume(patches.cuda(), crops, shape, path=stitch_path)
    assert torch.equal(t1, t0) and torch.equal(c1, c0)
    # carrier with a weight: normalised softmax = weighted mean, uncovered remainder stays 0
    carrier = vb.DataCarrier3D(patch_weight=w, stitch_path=stitch_path)
    for pred_idx in range(2):
        batch = {"image_paths": ["v"] * len(crops), "label_paths": [None] * len(crops),
                 "org_image_size": [shape] * len(crops), "crop_idx": crops, "data": None, "seg": None}
        carrier.concat_data(batch, patches[pred_idx], n_pred=2, pred_idx=pred_idx)
    norm = carrier.normalized("v")["softmax_pred"].cpu().numpy()
    want = ref_total / np.where(ref_cnt > 0, ref_cnt, 1.0)
    np.testing.assert_allclose(norm, want, rtol=1e-15)
    assert np.all(norm[:, :, :, 32:, :] == 0) and np.all(ref_cnt[:, 32:, :] == 0)
    # the same carrier fed with the separable factors of the map: bit-identical accumulators
    carrier_f = vb.DataCarrier3D(patch_weight=vb.gaussian_importance_factors((p, p, p)), stitch_path=stitch_path)
    for pred_idx in range(2):
        batch = {"image_paths": ["v"] * len(crops), "label_paths": [None] * len(crops),
                 "org_image_size": [shape] * len(crops), "crop_idx": crops, "data": None, "seg": None}
        carrier_f.concat_data(batch, patches[pred_idx], n_pred=2, pred_idx=pred_idx)
    assert torch.equal(carrier_f.data["v"]["softmax_pred"], carrier.data["v"]["softmax_pred"])
    assert torch.equal(carrier_f.data["v"]["_count"], carrier.data["v"]["_count"])


def test_stitch_separable_weights(vb, vo, stitch_path):
    """The Gaussian weight passed as its three 1-D factors (box kernel: factors in shared memory, no weight
    map traffic; other paths: the materialised map): bit-identical to the numpy statement on the map
    (wx * wy) * wz, on aligned grids, on z origins the copy engine cannot fetch, on a partially covered
    volume and accumulating on top of earlier sums."""
    fac = vb.gaussian_importance_factors((16, 16, 16))
    for a, b in zip(fac, vo.gaussian_importance_factors((16, 16, 16))):
        np.testing.assert_array_equal(a.numpy(), b)
    np.testing.assert_array_equal(vb.importance_map_from_factors(fac).numpy(), vo.gaussian_importance_map((16, 16, 16)))
    g = torch.Generator().manual_seed(8)
    for shape, p, crops in [
        ((40, 36, 72), (16, 16, 16), None),
        ((24, 40, 44), (8, 16, 12), [((0, 8), (0, 16), (0, 12)), ((3, 11), (5, 21), (6, 18)), ((16, 24), (24, 40), (32, 44)),
                                      ((2, 10), (1, 17), (3, 15)), ((8, 16), (8, 24), (20, 32))]),
    ]:
        if crops is None:
            crops = vb.patch_grid(shape, p[0], 0.5)
        patches = torch.rand(2, len(crops), 2, *p, generator=g, dtype=torch.float32)
        fac = vb.gaussian_importance_factors(p, sigma_scale=0.25)
        w = vo.gaussian_importance_map(p, sigma_scale=0.25)
        ref_total, ref_cnt = vo.stitch_weighted(patches.numpy(), crops, shape, w)
        total, cnt = vb.stitch_volume(patches.cuda(), crops, shape, weight=tuple(f.cuda() for f in fac), path=stitch_path)
        np.testing.assert_array_equal(total.cpu().numpy(), ref_total)
        np.testing.assert_array_equal(cnt.cpu().numpy(), ref_cnt)
        # the same through the map entry point, and a second accumulation on top
        t2, c2 = vb.stitch_volume(patches.cuda(), crops, shape, weight=torch.from_numpy(w).cuda(), path=stitch_path)
        assert torch.equal(t2, total) and torch.equal(c2, cnt)
        lo = vb.stitching.crops_to_lo(crops, "cuda")
        vb.stitch_accumulate(patches.cuda(), lo, total, cnt, accumulate=True, weight=fac, path=stitch_path)
        vb.stitch_accumulate(patches.cuda(), lo, t2, c2, accumulate=True, weight=torch.from_numpy(w), path=stitch_path)
        assert torch.equal(t2, total) and torch.equal(c2, cnt)
        r2, rc2 = ref_total.copy(), ref_cnt.copy()
        for i, ((x0, x1), (y0, y1), (z0, z1)) in enumerate(crops):
            for n in range(2):
                r2[n, :, x0:x1, y0:y1, z0:z1] += w * patches[n, i].numpy().astype(np.float64)
            rc2[x0:x1, y0:y1, z0:z1] += w
        np.testing.assert_array_equal(total.cpu().numpy(), r2)
        np.testing.assert_array_equal(cnt.cpu().numpy(), rc2)
    # patch edges beyond the separable path's table: the library refuses, the wrapper materialises the map
    big = (132, 8, 8)
    patches = torch.rand(1, 1, 1, *big, generator=g, dtype=torch.float32)
    fac = vb.gaussian_importance_factors(big)
    total, cnt = vb.stitch_volume(patches.cuda(), [((0, 132), (0, 8), (0, 8))], (132, 8, 8), weight=fac, path=stitch_path)
    w = vo.gaussian_importance_map(big)
    np.testing.assert_array_equal(total.cpu().numpy()[0, 0], w * patches[0, 0, 0].numpy().astype(np.float64))
    np.testing.assert_array_equal(cnt.cpu().numpy(), w)


# ------------------------------------------------------------------------------- pipeline
def test_pipeline_vs_oracle(vb, vo):
    B, n, c, spatial = 5, 5, 2, (32, 36, 40)
    x = softmax_stack(99, B * n, c, spatial).reshape(B, n, c, *spatial)
    thr = (0.45, 0.4, 0.03)
    cfg = vb.AggregationConfig(patch_size=10, thresholds=thr, chunk_bytes=3 * 32 * 36 * 40 * 4 * 2)  # two volumes per chunk
    res = vb.UncertaintyPipeline(cfg).run(x.cuda(), keep_maps=True, mean_argmax=True)
    ids = [f"img{b}" for b in range(B)]
    dicts = res.to_dicts(ids)
    for b in range(B):
        ref = vo.calculate_uncertainty(x[b])
        for k, key in enumerate(MAPS):
            got_map = res.maps[k, b].cpu().numpy()
            np.testing.assert_allclose(got_map, ref[key].numpy(), rtol=RTOL, atol=ATOL)
            e = dicts[key][ids[b]]
            # aggregations of OUR map against the oracle's aggregation of the same map: tight
            m64 = got_map.astype(np.float64)
            pl = vo.patch_level_aggregation(m64, 10)
            assert e["patch_level"]["bounding_box"] == pl["bounding_box"]
            np.testing.assert_allclose(e["patch_level"]["max_score"], pl["max_score"], rtol=1e-12)
            np.testing.assert_allclose(e["image_level"]["max_score"], m64.sum(), rtol=1e-12)
            th = vo.threshold_aggregation(m64, threshold=thr[k])
            np.testing.assert_allclose(e["threshold"]["max_score"], float(th["max_score"]), rtol=1e-12)
            # and against the oracle end to end (reference maps): within the float tolerance
            np.testing.assert_allclose(e["patch_level"]["max_score"],
                                       vo.patch_level_aggregation(ref[key].numpy().astype(np.float64), 10)["max_score"], rtol=1e-5)
        assert_argmax(res.mean_argmax[b], x[b], np.argmax(np.mean(x[b].numpy(), axis=0), axis=0))


def test_pipeline_chunking_and_overlap_do_not_change_results(vb):
    """One chunk, many chunks, and K2b on a second stream under the next chunk's K1 must give
    bit-identical score tables and maps (the two-stream schedule only reorders launches)."""
    B, n, c, spatial = 7, 4, 3, (40, 44, 64)
    x = softmax_stack(123, B * n, c, spatial).reshape(B, n, c, *spatial).cuda()
    thr = (0.5, 0.4, 0.03)
    per_vol = 3 * int(np.prod(spatial)) * 4
    outs = []
    for chunk, overlap in ((1 << 30, False), (2 * per_vol, False), (2 * per_vol, True), (per_vol, True)):
        cfg = vb.AggregationConfig(patch_size=10, thresholds=thr, chunk_bytes=chunk, overlap=overlap)
        pipe = vb.UncertaintyPipeline(cfg)
        for _ in range(2):   # second run reuses the scratch buffers
            r = pipe.run(x, keep_maps=False, mean_argmax=True)
        torch.cuda.synchronize()
        outs.append((r.scores.clone(), r.mean_argmax.clone()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])


def test_pipeline_overlap_across_runs(vb):
    """Overlap mode lets run() return with K2b and the score assembly still queued on the side stream:
    back-to-back runs on DIFFERENT inputs (scratch maps alternate, per-run score buffers) must each
    give the single-stream result, whenever the tables are read."""
    B, n, c, spatial = 3, 4, 3, (40, 44, 64)
    xs = [softmax_stack(200 + i, B * n, c, spatial).reshape(B, n, c, *spatial).cuda() for i in range(5)]
    thr = (0.5, 0.4, 0.03)
    plain = vb.UncertaintyPipeline(vb.AggregationConfig(patch_size=10, thresholds=thr))
    want = [plain.run(x, mean_argmax=True).scores.clone() for x in xs]
    torch.cuda.synchronize()
    pipe = vb.UncertaintyPipeline(vb.AggregationConfig(patch_size=10, thresholds=thr, overlap=True))
    for _ in range(3):
        res = [pipe.run(x, mean_argmax=True) for x in xs]          # nothing waits in between
        assert all(r.ready is not None for r in res)
        for r, w in zip(reversed(res), reversed(want)):             # read in any order
            assert torch.equal(r.scores, w)
    table, ready = res[0].table_async()
    assert ready is res[0].ready and table is res[0]._scores


def test_full_size_properties(vb):
    """BASELINE sizes, checked through size-independent properties (no oracle at this size)."""
    n, c, s = 16, 4, (128, 128, 128)
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.softmax(3 * torch.randn(1, n, c, *s, generator=g, device="cuda"), dim=2)
    pipe = vb.UncertaintyPipeline(vb.AggregationConfig(patch_size=10, thresholds=(0.0, 0.0, -1.0)))
    r = pipe.run(x, keep_maps=True)
    pe, ee, mi = r.maps[0, 0], r.maps[1, 0], r.maps[2, 0]
    assert torch.equal(mi, pe - ee)
    assert pe.min().item() >= 0 and pe.max().item() <= np.log(4) * (1 + 1e-6)
    assert mi.min().item() >= -1e-6  # Jensen: MI >= 0 up to rounding
    sc = r.scores.cpu().numpy()[0]
    V = float(np.prod(s))
    assert sc[0, 2] == V and sc[2, 2] == V            # threshold 0 / -1 keeps every voxel
    np.testing.assert_allclose(sc[:, 0], sc[:, 1], rtol=1e-15)  # so thr_sum == image sum
    np.testing.assert_allclose(sc[:, 0], r.maps[:, 0].double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=1e-9)
    # patch score bounded by 1000 * map max and >= mean box; sample permutation invariance
    assert np.all(sc[:, 3] <= 1000 * r.maps[:, 0].amax(dim=(1, 2, 3)).cpu().numpy() * (1 + 1e-12))
    assert np.all(sc[:, 3] >= sc[:, 0] / V * 1000 * (1 - 1e-9))
    r2 = pipe.run(x[:, torch.randperm(n, device="cuda")], keep_maps=True)
    np.testing.assert_allclose(r2.maps.cpu().numpy(), r.maps.cpu().numpy(), rtol=1e-5, atol=1e-6)
