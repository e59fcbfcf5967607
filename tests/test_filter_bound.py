"""CPU check of the error bound behind K2b's fp32 filter pass (values_b200/csrc/aggregate.cu,
`filter_err_coef`): the filter may drop a sub-chunk only if its fp32 maximum is provably too small,
so |fp32 box sum - exact box sum| <= coef * max|input| must hold for the kernels' exact operation
order.  Here that order is emulated in numpy float32 (every add / subtract rounds to nearest, as
FADD / FADD2 do) and compared with fp64 box sums on random and adversarial maps; the coefficient
comes from the library itself (values_patch_filter_err_coef, a pure host function)."""
import numpy as np
import pytest

F = np.float32
P = 10


def tree10(v):
    """tree_sum_t<10>: ((v0+v1) + (v2+(v3+v4))) + ((v5+v6) + (v7+(v8+v9))), fp32 at every node."""
    def t5(a):
        return (a[0] + a[1]) + (a[2] + (a[3] + a[4]))
    return t5(v[:5]) + t5(v[5:])


def ytree_vec(c):
    """box_filter_kernel's y-stage start: q_j = c[2j] + c[2j+1]; ((q0+q1) + (q2+q3)) + q4."""
    q = [c[2 * j] + c[2 * j + 1] for j in range(5)]
    return ((q[0] + q[1]) + (q[2] + q[3])) + q[4]


def z_stage(m, p0, zc):
    D0 = m.shape[0]
    O0 = D0 - p0 + 1
    out = np.zeros((O0,) + m.shape[1:], F)
    for c0 in range(0, O0, zc):                       # the sliding sum restarts in every z-chunk
        zs = np.zeros(m.shape[1:], F)
        for zi in range(min(zc, O0 - c0) + p0 - 1):
            nw = m[c0 + zi]
            od = m[c0 + zi - p0] if zi >= p0 else np.zeros_like(nw)
            zs = zs + (nw - od)                       # fl(zs + fl(new - old))
            if zi >= p0 - 1:
                out[c0 + zi - (p0 - 1)] = zs
    return out


def slide_axis(a, axis, run, chains, start_fn):
    """Box sums of width P along `axis`: tasks of `run` outputs starting at multiples of `run`,
    each as `chains` independent chains (tree / pair start, then s = fl(s + fl(in - out)))."""
    a = np.moveaxis(a, axis, 0)
    n_out = a.shape[0] - P + 1
    out = np.zeros((n_out,) + a.shape[1:], F)
    step = run // chains
    for s0 in range(0, n_out, step):                  # every chain start is a fresh tree sum
        s = start_fn([a[s0 + k] for k in range(P)])
        out[s0] = s
        for i in range(1, min(step, n_out - s0)):
            s = s + (a[s0 + i + P - 1] - a[s0 + i - 1])
            out[s0 + i] = s
    return np.moveaxis(out, 0, axis)


def emulate(m, p0, zc, vector_kernel):
    z = z_stage(m.astype(F), p0, zc)
    if vector_kernel:   # x: 16 outputs, one chain; y: 8 outputs, one chain, pairwise start
        x = slide_axis(z, 2, 16, 1, tree10)
        return slide_axis(x, 1, 8, 1, ytree_vec)
    x = slide_axis(z, 2, 16, 2, tree10)               # march kernel: two chains of 8 / of 4, tree starts
    return slide_axis(x, 1, 8, 2, tree10)


def exact(m, p0):
    c = np.cumsum(np.cumsum(np.cumsum(np.pad(m.astype(np.longdouble), ((1, 0), (1, 0), (1, 0))), 0), 1), 2)
    a, b, d = p0, P, P
    return (c[a:, b:, d:] - c[:-a, b:, d:] - c[a:, :-b, d:] - c[a:, b:, :-d]
            + c[:-a, :-b, d:] + c[:-a, b:, :-d] + c[a:, :-b, :-d] - c[:-a, :-b, :-d])


def maps():
    rng = np.random.default_rng(99)
    shape = (40, 27, 45)
    out = {"uniform": rng.random(shape), "signed": rng.standard_normal(shape)}
    m = rng.random(shape) * 1e-3
    m[7, 9, 11] = 1e4
    m[25, 20, 30] = -3e3
    out["spikes"] = m                                             # big values leaving a window
    out["lognormal"] = np.exp(6 * rng.standard_normal(shape))     # 20 binades of dynamic range
    m = rng.random(shape)
    m[::2] *= 1e5                                                 # alternate huge / small planes: the
    out["alternating_planes"] = m                                 # sliding sum cancels at every step
    m = rng.random(shape)
    m[:, :, ::2] *= -1e4
    out["alternating_columns"] = m
    return {k: v.astype(F) for k, v in out.items()}


@pytest.mark.parametrize("vector_kernel", [1, 0], ids=["box_filter_kernel", "march_fp32"])
@pytest.mark.parametrize("p0,zc", [(10, 31), (10, 8), (3, 38), (1, 40)])
def test_fp32_box_sums_stay_within_the_filter_bound(vector_kernel, p0, zc):
    from values_b200 import _lib

    coef = _lib.lib.values_patch_filter_err_coef(zc, p0, vector_kernel)
    assert coef > 0
    for name, m in maps().items():
        got = emulate(m, p0, zc, vector_kernel).astype(np.longdouble)
        want = exact(m, p0)
        amax = float(np.abs(m).max())
        err = float(np.abs(got - want).max())
        assert err <= coef * amax, (name, err, coef * amax)
        assert err <= 0.25 * coef * amax, (name, err / (coef * amax))   # and with room to spare


def test_fp64_log_table_accuracy():
    """The table-driven fp64 log of K1's fp64 path (fast_log_f64, values_b200/csrc/uncertainty.cu):
    the algorithm restated in numpy (tools/check_log64.py) stays within 5e-14 of a 120-bit log, and the
    table in the kernel source is the one that script builds."""
    import os
    import re
    import sys

    pytest.importorskip("mpmath")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import check_log64

    check_log64.main()                                   # asserts the error bound
    inv, lnc = check_log64.build()
    src = open(os.path.join(root, "values_b200", "csrc", "uncertainty.cu")).read()
    body = src[src.index("kLog64Tab[129] = {"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[+-]?\d+", body)]
    assert len(vals) == 258
    np.testing.assert_array_equal(np.asarray(vals[0::2]), inv)
    np.testing.assert_array_equal(np.asarray(vals[1::2]), lnc)
