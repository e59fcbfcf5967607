#!/usr/bin/env python
"""Build and check the table behind fast_log_f64 (values_b200/csrc/uncertainty.cu): prints the
513-entry {inv, -log(inv)} table as the C initialiser used there and measures the algorithm's error
against a 120-bit reference (mpmath) over values from 1e-300 to 8, densely around 1.

    python tools/check_log64.py [--print-table]
"""
import math
import sys

import mpmath as mp
import numpy as np

LO = 0x3FE6A09E00000000          # the octave [LO, 2 LO) starts just below sqrt(2)/2: only the HIGH word of p is
                                 # needed to split exponent and mantissa (k, table index: 32-bit integer ops)
IDX_BITS = 9                     # top mantissa bits that select the interval
SHIFT = 52 - IDX_BITS            # 43
BASE = LO >> SHIFT               # 0x7FCD4
N_TAB = 513


def build():
    mp.mp.prec = 200
    inv, lnc = np.zeros(N_TAB), np.zeros(N_TAB)
    for i in range(N_TAB):
        hb = BASE + i            # sign / exponent / top 9 mantissa bits of the interval's left end
        left = np.array([hb << SHIFT], dtype=np.uint64).view(np.float64)[0]
        right = np.array([(hb + 1) << SHIFT], dtype=np.uint64).view(np.float64)[0]
        c = 1.0 if left <= 1.0 <= right else 0.5 * (left + right)
        iv = 1.0 / c
        e = math.floor(math.log2(iv))
        iv = round(iv / 2.0 ** (e - 11)) * 2.0 ** (e - 11)      # 12 significant bits
        inv[i] = iv
        lnc[i] = float(-mp.log(mp.mpf(iv))) if iv != 1.0 else 0.0
    return inv, lnc


def fast_log(p, inv, lnc):
    bits = p.view(np.uint64).astype(np.int64)
    k = (bits - LO) >> 52
    mb = bits - (k << 52)
    idx = (mb >> SHIFT) - BASE
    assert idx.min() >= 0 and idx.max() < N_TAB
    # the kernel uses one fma here: m * inv (53 + 12 bits) is held exactly by the x87 long double
    r = (mb.view(np.float64).astype(np.longdouble) * inv[idx].astype(np.longdouble) - 1.0).astype(np.float64)
    r2 = r * r
    q = r2 * (0.2 * r - 0.25) + (r / 3.0 - 0.5)       # Estrin, as the kernel (its fmas round once, these twice)
    return k * 0.6931471805599453 + (lnc[idx] + (r2 * q + r)), np.abs(r).max()


def main():
    inv, lnc = build()
    if "--print-table" in sys.argv:
        for i in range(0, N_TAB, 3):
            print("    " + " ".join("{%s, %s}," % (float.hex(a), float.hex(b)) for a, b in zip(inv[i:i + 3], lnc[i:i + 3])))
    rng = np.random.default_rng(0)
    p = np.concatenate([rng.random(200000), 1 - 10 ** rng.uniform(-15, -1, 100000), 1 + 10 ** rng.uniform(-15, -1, 50000),
                        10 ** rng.uniform(-300, 2, 100000), rng.random(50000) * 8])
    p = p[(p > 0) & (p != 1.0)]
    got, rmax = fast_log(p, inv, lnc)
    mp.mp.prec = 120
    sub = rng.choice(len(p), 20000, replace=False)
    err = max(abs((mp.mpf(float(got[i])) - mp.log(mp.mpf(float(p[i])))) / mp.log(mp.mpf(float(p[i])))) for i in sub)
    print(f"max |r| {rmax:.3e} (2^-9 = {2 ** -9:.3e}); max relative error vs 120-bit log: {float(err):.2e}")
    assert float(err) < 3e-14


if __name__ == "__main__":
    main()
