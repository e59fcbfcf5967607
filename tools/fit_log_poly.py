"""Fit the polynomial used by fast_logf in values_b200/csrc/uncertainty.cu and report its
error over EVERY fp32 mantissa in the reduced range [2/3, 4/3).

    log(m) = f + f^2 * Q(f),  f = m - 1,  |f| <= 1/3,  Q of degree DEG (Horner, fp32 FFMA)
"""
import sys
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 7
a, b = -1.0 / 3.0, 1.0 / 3.0 + 1e-3


def q(f):
    f = np.asarray(f, dtype=np.float64)
    out = np.empty_like(f)
    small = np.abs(f) < 1e-4
    fs = f[small]
    out[small] = -0.5 + fs / 3 - fs**2 / 4 + fs**3 / 5
    fl = f[~small]
    out[~small] = (np.log1p(fl) - fl) / fl**2
    return out


# near-minimax: Chebyshev interpolation, then a few Remez-like reweighting rounds (Lawson)
xs = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * (b - a) / 2 + (a + b) / 2
w = np.ones_like(xs)
for it in range(60):
    # weight: relative error of log(m) = f + f^2 Q  ->  error*f^2/|log1p(f)|
    scale = xs**2 / np.maximum(np.abs(np.log1p(xs)), 1e-30)
    A = np.vander(xs, DEG + 1, increasing=True)
    W = np.sqrt(w) * scale
    coef, *_ = np.linalg.lstsq(A * W[:, None], q(xs) * W, rcond=None)
    err = np.abs((A @ coef - q(xs)) * scale)
    w = w * (err / err.max() + 1e-3)
    w /= w.sum()
coef32 = coef.astype(np.float32)
print("degree", DEG, "fit max rel err (fp64 eval)", err.max())
print("coefficients c0..c%d (Q(f) = sum c_k f^k):" % DEG)
for k, c in enumerate(coef32):
    print(f"  c{k} = {float(c)!r}f   // {c.view(np.uint32):#010x}")

# exhaustive fp32 check
lo = np.float32(2.0 / 3.0).view(np.uint32)
hi = np.float32(4.0 / 3.0).view(np.uint32)
bits = np.arange(lo, hi, dtype=np.uint32)
m = bits.view(np.float32)
f = (m - np.float32(1.0)).astype(np.float32)          # exact (Sterbenz)
acc = np.full_like(f, coef32[DEG])
for k in range(DEG - 1, -1, -1):                       # Horner with fused multiply-add
    acc = (acc.astype(np.float64) * f + np.float64(coef32[k])).astype(np.float32)
t = (acc.astype(np.float64) * f).astype(np.float32)    # Q*f
res = (t.astype(np.float64) * f + f).astype(np.float32)  # fma(Q*f, f, f)
ref = np.log(m.astype(np.float64))
ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
nz = ref != 0
e_ulp = np.abs(res[nz] - ref[nz]) / ulp[nz]
e_rel = np.abs(res[nz] - ref[nz]) / np.abs(ref[nz])
print("exhaustive over", len(m), "floats: max ulp err", e_ulp.max(), "max rel err", e_rel.max(),
      "mean ulp", e_ulp.mean())
