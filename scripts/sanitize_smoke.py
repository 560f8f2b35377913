"""Small exercise of every kernel (both map paths, taps, sampler) for compute-sanitizer runs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from joxsz_b200 import cluster
from joxsz_b200.batched import BatchedLikelihood
from joxsz_b200.mb import mb
from joxsz_b200.sampler import EnsembleSampler
from joxsz_b200.synthetic import draw_parameters

mb.fit.debugfit = False
base = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for name, inp in (("cl1226", base), ("synth255", cluster.synthetic_inputs(map_half=127, nr=512, base=base)),
                  ("synth511", cluster.synthetic_inputs(map_half=255, nr=1024, base=base))):
    if which not in ("all", name):
        continue
    fit, _ = cluster.build_fit(inp, savedir=None)
    eng = BatchedLikelihood(fit, max_walkers=512)
    th = draw_parameters(fit.thawed, n=37, seed=3, spread=0.03, frac_bad=0.2)
    ll = eng(th)
    eng.profiles(th[:5]); eng.sz_project(th[:3]); eng.sz_maps(th[:2]); eng.sz_profile(th[:5]); eng.xray(th[:5])
    if name == "cl1226":
        s = EnsembleSampler(64, eng.ndim, eng, seed=5)
        s.initialize(draw_parameters(fit.thawed, n=64, seed=8, spread=0.01))
        for _ in range(2):
            s.step()
        fit.press.press_fun(fit.pars, np.linspace(10.0, 2000.0, 33))
        fit.mass_cmpt.mass_fun(fit.pars, np.linspace(10.0, 2000.0, 33))
    print(name, "finite", int(np.isfinite(ll).sum()), "of", ll.size, flush=True)
    eng.close()
print("sanitize_smoke done")
