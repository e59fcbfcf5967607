"""The 2-D caller (uncertainty_modeling/test_2D.py:204-254, Tester.process_output) on the device against the
golden fixture written by the reference's own method (tests/golden/make_golden.py::process_output_2d_case)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "process_output_2d.npz")


@pytest.fixture(scope="module")
def vb():
    import values_b200

    return values_b200


@pytest.mark.parametrize("name", ["mc", "ssn", "single"])
def test_process_output_matches_the_reference_method(vb, name):
    g = np.load(GOLDEN)
    sm, gt, ssn = torch.from_numpy(g[name + "/softmax_pred"]), torch.from_numpy(g[name + "/gt"]), bool(g[name + "/ssn"])
    n, b, c = sm.shape[:3]
    ids = [f"{name}_{i}" for i in range(b)]
    gt_before = gt.clone()
    res = vb.tester2d.process_output({"softmax_pred": sm.cuda(), "gt": gt.cuda(), "image_id": ids,
                                      "dataset": ["synthetic"] * b}, is_ssn=ssn, ignore_index=255)
    assert torch.equal(gt, gt_before) and list(res) == ids
    for i, key in enumerate(ids):
        r = res[key]
        want_metrics = {k.rsplit("/", 1)[1]: float(g[k]) for k in g.files if k.startswith(f"{name}/{i}/metrics/")}
        assert set(r["metrics"]) == set(want_metrics) == {"dice", "ged"}
        for m, v in want_metrics.items():
            np.testing.assert_allclose(r["metrics"][m], v, rtol=1e-12, atol=1e-15, err_msg=f"{key}/{m}")
        np.testing.assert_array_equal(r["sample_argmax"].cpu().numpy(), g[f"{name}/{i}/sample_argmax"])
        np.testing.assert_array_equal(r["mean_argmax"].cpu().numpy(), g[f"{name}/{i}/mean_argmax"])
        np.testing.assert_array_equal(r["ignore_index_map"].cpu().numpy(), g[f"{name}/{i}/ignore"])
        want_unc = {k.rsplit("/", 1)[1]: g[k] for k in g.files if k.startswith(f"{name}/{i}/unc/")}
        assert set(r["uncertainty"]) == set(want_unc)
        for k, v in want_unc.items():
            got = r["uncertainty"][k].cpu().numpy()
            assert got.dtype == v.dtype == np.float32 and got.shape == v.shape
            np.testing.assert_allclose(got, v, rtol=1e-5, atol=1e-6, err_msg=f"{key}/{k}")
    summary = vb.tester2d.results_dict_with_mean(res)
    assert set(summary) == set(ids) | {"mean"}
    np.testing.assert_allclose(summary["mean"]["metrics"]["ged"], np.mean([res[k]["metrics"]["ged"] for k in ids]))


def test_process_output_reads_the_stack_in_place(vb):
    """The [N, B, C, H, W] stack is consumed as a permuted view (no copy, no zero channel): a stack that is
    itself a strided slice of a larger buffer gives the same results as its contiguous copy."""
    g = torch.Generator().manual_seed(5)
    big = torch.softmax(torch.randn(6, 3, 7, 20, 32, generator=g), dim=2).cuda()
    view = big[1:5, :, :6]                      # N = 4 of 6 samples, 6 of 7 channels: strided in N and C
    gt = torch.randint(0, 6, (3, 1, 20, 32), generator=g).cuda()
    args = {"gt": gt, "image_id": ["a", "b", "c"], "dataset": ["s"] * 3}
    r1 = vb.tester2d.process_output(dict(args, softmax_pred=view))
    r2 = vb.tester2d.process_output(dict(args, softmax_pred=view.contiguous()))
    for k in r1:
        assert r1[k]["metrics"] == r2[k]["metrics"]
        for m in r1[k]["uncertainty"]:
            assert torch.equal(r1[k]["uncertainty"][m], r2[k]["uncertainty"][m])
        assert torch.equal(r1[k]["sample_argmax"], r2[k]["sample_argmax"])
