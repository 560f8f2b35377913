"""How many int8 slices would an error-free (Ozaki) split of the projection GEMM K2 need?  CPU study, exact integer
arithmetic: coef = proj_op . pp with both operands cut into signed 7-bit planes (first plane 6 bits + sign), products
of plane pairs accumulated exactly, pairs kept while i + j <= T.  Error measured against a long-double product, next to the
error of the plain float64 GEMM, on the pressure profiles of the golden parameter draws.
    python scripts/ozaki_slices_study.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from joxsz_b200.packer import PackedSetup
from helpers import oracle_setup_from_fit, orc

fit = bench.build_cluster("cl1226")
pk = PackedSetup(fit, max_walkers=64, device=0)
B = np.asarray(pk.proj_op, dtype=np.float64)                     # [484, 313]
g = np.load(os.path.join(ROOT, "tests", "golden", "cl1226_golden.npz"))
fin = np.isfinite(g["ll"])
A = np.asarray(g["pp"], dtype=np.float64)[fin]                   # [W, 313]
ref = (A.astype(np.longdouble) @ B.T.astype(np.longdouble))
f64 = A @ B.T
scale = np.abs(A) @ np.abs(B.T)                                   # sum |a||b|: the natural error scale of a dot product
print(f"K2: {A.shape[0]} walkers x {B.shape[0]} outputs, K = {A.shape[1]}")
print(f"float64 GEMM: max |err| / sum|a||b| = {np.max(np.abs((f64 - ref).astype(np.float64)) / scale):.2e}; "
      f"max |err| / |coef| = {np.max(np.abs((f64 - ref).astype(np.float64)) / np.maximum(np.abs(f64), 1e-300)):.2e}")
print(f"cancellation: median sum|a||b| / |coef| = {np.median(scale / np.maximum(np.abs(f64), 1e-300)):.1f}, max = {np.max(scale / np.maximum(np.abs(f64), 1e-300)):.2e}")


def planes(X, n):
    """X [rows, K] -> integer planes [n, rows, K] and the per-row exponent e: X = 2^e * sum_i P_i 2^(-6 - 7 i)."""
    e = np.ceil(np.log2(np.max(np.abs(X), axis=1, keepdims=True)))
    r = X * np.exp2(-e) * 64.0                                     # |r| <= 64
    out = []
    for _ in range(n):
        d = np.rint(r)
        out.append(d.astype(np.int64))
        r = (r - d) * 128.0
    return np.stack(out), e


for n in (4, 5, 6, 7, 8):
    PA, ea = planes(A, n)
    PB, eb = planes(B, n)
    assert np.abs(PA).max() <= 64 and np.abs(PB).max() <= 64
    for T in range(n - 1, 2 * n - 1):
        acc = np.zeros(ref.shape, dtype=np.longdouble)
        npairs = 0
        for i in range(n):
            for j in range(n):
                if i + j <= T:
                    prod = PA[i] @ PB[j].T                           # exact: |sum| <= 313 * 64 * 64 < 2^31
                    acc += prod.astype(np.longdouble) * np.longdouble(2.0) ** (-12 - 7 * (i + j))
                    npairs += 1
        val = (acc * np.exp2(ea.astype(np.longdouble)) * np.exp2(eb.T.astype(np.longdouble)))
        err = np.abs((val - ref).astype(np.float64))
        print(f"planes {n}, pairs with i+j <= {T} ({npairs:2d} products): max |err| / sum|a||b| = {np.max(err / scale):.2e}, "
              f"max |err| / |coef| = {np.max(err / np.maximum(np.abs(f64), 1e-300)):.2e}")
