#!/usr/bin/env python
"""Build and check the table behind fast_log_f64 (values_b200/csrc/uncertainty.cu): prints the
129-entry {inv, -log(inv)} table as the C initialiser used there and measures the algorithm's error
against a 120-bit reference (mpmath) over values from 1e-300 to 8, densely around 1.

    python tools/check_log64.py [--print-table]
"""
import math
import sys

import mpmath as mp
import numpy as np

LO = 0x3FE6A09E667F3BCD          # bits of sqrt(2)/2
BASE = LO >> 45                  # 0x1FF35


def build():
    mp.mp.prec = 200
    inv, lnc = np.zeros(129), np.zeros(129)
    for i in range(129):
        hb = BASE + i            # sign / exponent / top 7 mantissa bits of the interval's left end
        left = np.array([hb << 45], dtype=np.uint64).view(np.float64)[0]
        right = np.array([(hb + 1) << 45], dtype=np.uint64).view(np.float64)[0]
        c = 1.0 if left <= 1.0 <= right else 0.5 * (left + right)
        iv = 1.0 / c
        e = math.floor(math.log2(iv))
        iv = round(iv / 2.0 ** (e - 9)) * 2.0 ** (e - 9)      # 10 significant bits
        inv[i] = iv
        lnc[i] = float(-mp.log(mp.mpf(iv))) if iv != 1.0 else 0.0
    return inv, lnc


def fast_log(p, inv, lnc):
    bits = p.view(np.uint64).astype(np.int64)
    k = (bits - LO) >> 52
    mb = bits - (k << 52)
    idx = (mb >> 45) - BASE
    r = mb.view(np.float64) * inv[idx] - 1.0          # the kernel uses one fma here
    q = 1.0 / 7.0
    for c in (-1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5):
        q = q * r + c
    return k * 0.6931471805599453 + (lnc[idx] + (r * r * q + r)), np.abs(r).max()


def main():
    inv, lnc = build()
    if "--print-table" in sys.argv:
        for i in range(0, 129, 2):
            print("    " + " ".join("{%s, %s}," % (float.hex(a), float.hex(b)) for a, b in zip(inv[i:i + 2], lnc[i:i + 2])))
    rng = np.random.default_rng(0)
    p = np.concatenate([rng.random(200000), 1 - 10 ** rng.uniform(-15, -1, 100000), 1 + 10 ** rng.uniform(-15, -1, 50000),
                        10 ** rng.uniform(-300, 2, 100000), rng.random(50000) * 8])
    p = p[(p > 0) & (p != 1.0)]
    got, rmax = fast_log(p, inv, lnc)
    mp.mp.prec = 120
    sub = rng.choice(len(p), 20000, replace=False)
    err = max(abs((mp.mpf(float(got[i])) - mp.log(mp.mpf(float(p[i])))) / mp.log(mp.mpf(float(p[i])))) for i in sub)
    print(f"max |r| {rmax:.3e} (2^-7 = {2 ** -7:.3e}); max relative error vs 120-bit log: {float(err):.2e}")
    assert float(err) < 5e-14


if __name__ == "__main__":
    main()
