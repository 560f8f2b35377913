"""``mbproj2.Fit`` work-alike: parameter bookkeeping, X-ray profile prediction, simplex refit."""
import numpy as np

debugfit = True


class Fit:
    def __init__(self, pars, model, data):
        self.pars = pars
        self.model = model
        self.data = data
        self.refreshThawed()
        self.bestlike = -1e99

    # Methods bound on the INSTANCE (joxsz_b200.cluster.build_fit does that so that several fits can live in one
    # process; the reference binds on the class) refer back to the object: pickle would resolve them with getattr on
    # the half-built copy.  They travel as plain functions and are re-bound on the copy.
    def __getstate__(self):
        from types import MethodType
        d = dict(self.__dict__)
        bound = {k: v.__func__ for k, v in d.items() if isinstance(v, MethodType) and v.__self__ is self}
        for k in bound:
            del d[k]
        d["__bound__"] = bound
        return d

    def __setstate__(self, d):
        from types import MethodType
        d = dict(d)
        bound = d.pop("__bound__", {})
        self.__dict__.update(d)
        for k, f in bound.items():
            setattr(self, k, MethodType(f, self))

    def refreshThawed(self):
        self.thawed = [name for name, par in sorted(self.pars.items()) if not par.frozen]

    def thawedParVals(self):
        return [self.pars[name].val for name in self.thawed]

    def updateThawed(self, vals):
        for val, name in zip(vals, self.thawed):
            self.pars[name].val = val

    def calcProfiles(self):
        """Predicted X-ray profiles per band for the current parameters, computed on the GPU (K1 + K4)."""
        from ..funcs import calcProfiles
        return calcProfiles(self)

    def getLikelihood(self, vals=None):
        raise NotImplementedError("bind joxsz_funcs.getLikelihood (joxsz_main.py:187)")

    def doFitting(self, silent=False, maxiter=10):
        """Alternate Nelder-Mead / Powell on -getLikelihood until it improves by < 0.1."""
        from scipy import optimize

        def neg(vals):
            like = self.getLikelihood(vals)
            return -like if np.isfinite(like) else 1e99

        like = self.getLikelihood(self.thawedParVals())
        for _ in range(maxiter):
            for method in ("Nelder-Mead", "Powell"):
                res = optimize.minimize(neg, np.array(self.thawedParVals(), dtype=float), method=method)
                self.updateThawed(np.atleast_1d(res.x))
            newlike = self.getLikelihood(self.thawedParVals())
            if not silent:
                print("Fit: %g -> %g" % (like, newlike))
            done = abs(newlike - like) < 0.1
            like = newlike
            if done:
                break
        return like
