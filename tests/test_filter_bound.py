"""CPU check of the error bound behind K2b's fp32 filter pass (values_b200/csrc/aggregate.cu,
`filter_err_coef`): the filter may drop a sub-chunk only if its fp32 maximum is provably too small,
so |fp32 box sum - exact box sum| <= coef * max|input| must hold for the kernel's exact operation
order.  Here that order is emulated in numpy float32 (every add / subtract rounds to nearest, as
FADD / FADD2 do) and compared with fp64 box sums on random and adversarial maps; the coefficient
comes from the library itself (values_patch_filter_err_coef, a pure host function)."""
import numpy as np
import pytest

F = np.float32
P = 10


def tree10(v):
    """tree_sum4<10>: ((v0+v1) + (v2+(v3+v4))) + ((v5+v6) + (v7+(v8+v9))), fp32 at every node."""
    def t5(a):
        return (a[0] + a[1]) + (a[2] + (a[3] + a[4]))
    return t5(v[:5]) + t5(v[5:])


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


def y_stage(z):
    """box_strip_filter_kernel's y-stage: a strip of K = 8 output rows starts with a tree over its
    first 10 z-window sums, then 7 slides o = fl(o + fl(in - out)); strips start at multiples of 8."""
    K = 8
    n_out = z.shape[1] - P + 1
    out = np.zeros((z.shape[0], n_out, z.shape[2]), F)
    for r0 in range(0, n_out, K):
        o = tree10([z[:, r0 + k] for k in range(P)])
        out[:, r0] = o
        for k in range(1, min(K, n_out - r0)):
            o = o + (z[:, r0 + k + P - 1] - z[:, r0 + k - 1])
            out[:, r0 + k] = o
    return out


def x_stage(y, xs, later_tile_wins):
    """x-stage: x tiles of `xs` outputs (119 for 128-float staged rows, 55 for 64-float rows) at a pitch of
    xs & ~3 (TMA origins are 16-byte aligned, so neighbouring tiles both compute 3 outputs: either value
    may be the one that reaches the maximum); lane l owns outputs 4l .. 4l+3 of the tile:
    S0 = (Q_l + Q_{l+1}) + A_{l+2} (Q = (e0+e1) + (e2+e3), A = e0+e1), then three slides.  Columns past
    the map edge arrive as zeros (TMA fill) and only feed masked outputs."""
    D2 = y.shape[2]
    n_out = D2 - P + 1
    e = np.concatenate([y, np.zeros(y.shape[:2] + (16,), F)], axis=2)
    out = np.zeros(y.shape[:2] + (n_out,), F)
    origins = list(range(0, max(n_out - xs, 0) + (xs & ~3), xs & ~3)) if n_out > xs else [0]
    for x0 in (origins if later_tile_wins else origins[::-1]):
        for l4 in range(0, min(xs, n_out - x0), 4):
            b = x0 + l4
            q0 = (e[..., b] + e[..., b + 1]) + (e[..., b + 2] + e[..., b + 3])
            q1 = (e[..., b + 4] + e[..., b + 5]) + (e[..., b + 6] + e[..., b + 7])
            a2 = e[..., b + 8] + e[..., b + 9]
            s = [(q0 + q1) + a2]
            for j in range(1, 4):
                s.append(s[-1] + (e[..., b + 9 + j] - e[..., b + j - 1]))
            for j in range(4):
                if b + j < min(n_out, x0 + xs):
                    out[..., b + j] = s[j]
    return out


def emulate(m, p0, zc, later_tile_wins):
    xs = 55 if m.shape[2] <= 64 else 119              # StripNarrow / StripWide
    return x_stage(y_stage(z_stage(m.astype(F), p0, zc)), xs, later_tile_wins)


def exact(m, p0):
    c = np.cumsum(np.cumsum(np.cumsum(np.pad(m.astype(np.longdouble), ((1, 0), (1, 0), (1, 0))), 0), 1), 2)
    a, b, d = p0, P, P
    return (c[a:, b:, d:] - c[:-a, b:, d:] - c[a:, :-b, d:] - c[a:, b:, :-d]
            + c[:-a, :-b, d:] + c[:-a, b:, :-d] + c[a:, :-b, :-d] - c[:-a, :-b, :-d])


def maps(shape):
    rng = np.random.default_rng(99)
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


@pytest.mark.parametrize("shape", [(40, 27, 45), (30, 35, 150)], ids=["narrow", "wide-two-x-tiles"])
@pytest.mark.parametrize("p0,zc", [(10, 31), (10, 8), (3, 38), (1, 40)])
def test_fp32_box_sums_stay_within_the_filter_bound(shape, p0, zc):
    from values_b200 import _lib

    coef = _lib.lib.values_patch_filter_err_coef(zc, p0)
    assert coef > 0
    for name, m in maps(shape).items():
        want = exact(m, p0)
        amax = float(np.abs(m).max())
        for later in (False, True) if shape[2] > 64 + 9 else (True,):
            got = emulate(m, p0, zc, later).astype(np.longdouble)
            assert np.isfinite(got).all() and got.shape == want.shape
            err = float(np.abs(got - want).max())
            assert err <= coef * amax, (name, err, coef * amax)
            assert err <= 0.25 * coef * amax, (name, err / (coef * amax))   # and with room to spare


def test_fp64_log_table_accuracy():
    """The table-driven fp64 log of K1's fp64 path (fast_log_f64, values_b200/csrc/uncertainty.cu):
    the algorithm restated in numpy (tools/check_log64.py) stays within 3e-14 of a 120-bit log, and the
    table the kernel includes (csrc/log64_table.inc) is the one that script builds."""
    import os
    import re
    import sys

    pytest.importorskip("mpmath")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import check_log64

    check_log64.main()                                   # asserts the error bound
    inv, lnc = check_log64.build()
    body = open(os.path.join(root, "values_b200", "csrc", "log64_table.inc")).read()
    vals = [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[+-]?\d+", body)]
    assert len(vals) == 2 * check_log64.N_TAB == 1026
    np.testing.assert_array_equal(np.asarray(vals[0::2]), inv)
    np.testing.assert_array_equal(np.asarray(vals[1::2]), lnc)
    src = open(os.path.join(root, "values_b200", "csrc", "uncertainty.cu")).read()
    assert "kLog64Tab[513]" in src and hex(check_log64.BASE).lower() in src.lower()
