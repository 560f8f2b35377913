"""Is a walker's log-likelihood bit-identical whatever batch it is evaluated in?  (It must be: the chain of an N-GPU
run is the 1-GPU chain only if it is.)  Evaluates the same 65,536 parameter vectors as one batch, as two halves and
in ragged pieces, and, where bits differ, walks the parity taps to find the stage.   python scripts/batch_invariance_check.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from joxsz_b200.batched import BatchedLikelihood

fit = bench.build_cluster("cl1226")
W = 65536
eng = BatchedLikelihood(fit, max_walkers=W, device=0)
theta = bench.ensemble(fit, W)
t = torch.from_numpy(theta).cuda()
full = eng.loglike_device(t).cpu().numpy()
again = eng.loglike_device(t).cpu().numpy()
print("same batch twice: differing", int((full.view(np.int64) != again.view(np.int64)).sum()))
for pieces in ([32768, 32768], [16384] * 4, [8192] * 8, [4096] * 16, [10000, 20000, 5536, 30000], [1] * 3 + [65533]):
    out, lo = [], 0
    for n in pieces:
        out.append(eng.loglike_device(t[lo:lo + n].contiguous()).cpu().numpy()); lo += n
    part = np.concatenate(out)
    bad = np.nonzero(full.view(np.int64) != part.view(np.int64))[0]
    print(f"pieces {pieces[:4]}{'...' if len(pieces) > 4 else ''}: differing walkers {bad.size}", bad[:8], (full[bad[:4]] - part[bad[:4]]) if bad.size else "")
    if bad.size and pieces[0] == 32768:
        w = int(bad[0]); lo = 0 if w < 32768 else 32768
        sel_full, sel_part = theta[:W], theta[lo:lo + 32768]
        for name, fn, keys in (("profiles", eng.profiles, ("pp", "tsz", "ne_ann", "tx_ann", "prior")), ("sz_project", eng.sz_project, ("coef",)),
                               ("sz_profile", eng.sz_profile, ("row", "bright", "model", "chisq", "cint")), ("xray", eng.xray, ("pred", "cash"))):
            a = fn(sel_full[: 32768] if lo == 0 else sel_full[32768:]) if False else fn(theta[max(0, w - 20000): w + 1][-20001:])   # context A: walker is the last row of a 20001-batch
            b = fn(theta[w: w + 4097])                                                                                                  # context B: first row of a 4097-batch
            for k in keys:
                xa, xb = np.asarray(a[k])[-1], np.asarray(b[k])[0]
                nd = int((np.atleast_1d(xa).view(np.int64) != np.atleast_1d(xb).view(np.int64)).sum())
                print(f"   tap {name}.{k}: differing elements {nd}")
