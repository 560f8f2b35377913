"""Helpers with the names and semantics of ``mbproj2.utils`` that JoXSZ touches."""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np


def uprint(*args, **kwargs):
    print(*args, **kwargs)


def projectionVolume(R1, R2, y1, y2):
    """Volume of the shell R1<r<R2 seen between projected radii y1<y<y2 (one hemisphere).

    ``2/3 pi [(p1^3 - p2^3) + (p4^3 - p3^3)]`` with ``p = sqrt(max(R^2 - y^2, 0))``
    (mbproj2 ``utils.projectionVolume``).
    """
    def rt(x):
        return np.sqrt(np.clip(x, 0.0, None))

    p1 = rt(R1 ** 2 - y2 ** 2)
    p2 = rt(R1 ** 2 - y1 ** 2)
    p3 = rt(R2 ** 2 - y2 ** 2)
    p4 = rt(R2 ** 2 - y1 ** 2)
    return (2.0 / 3.0) * np.pi * ((p1 ** 3 - p2 ** 3) + (p4 ** 3 - p3 ** 3))


def projectionVolumeMatrix(radii):
    """[shell, annulus] matrix of volumes (front + back) for shells/annuli with shared edges."""
    radii = np.asarray(radii, dtype=np.float64)
    i_s, i_a = np.indices((len(radii) - 1, len(radii) - 1))
    return 2.0 * projectionVolume(radii[i_s], radii[i_s + 1], radii[i_a], radii[i_a + 1])


def cashLogLikelihood(data, model):
    """mbproj2's Cash log-likelihood ``sum(data*log(model)) - sum(model)``.  Evaluated on the GPU in this
    package (``jx_cash_from_profiles`` behind ``fit.mylikeFromProfs``); no host implementation."""
    raise NotImplementedError("cashLogLikelihood has no host implementation here: use fit.mylikeFromProfs(profs)")


class AtomicWriteFile:
    """Write to a temporary file and rename over the target on clean exit."""

    def __init__(self, filename):
        self.filename = filename
        self._tmp = None
        self._fh = None

    def __enter__(self):
        d = os.path.dirname(os.path.abspath(self.filename))
        fd, self._tmp = tempfile.mkstemp(dir=d, prefix=".tmp_fit_")
        self._fh = os.fdopen(fd, "w")
        return self._fh

    def __exit__(self, exc_type, exc, tb):
        self._fh.close()
        if exc_type is None:
            os.replace(self._tmp, self.filename)
        else:
            os.unlink(self._tmp)
        return False


class WithLock:
    """Directory-based lock (mkdir is atomic)."""

    def __init__(self, dirname):
        self.dirname = dirname

    def __enter__(self):
        import time
        while True:
            try:
                os.mkdir(self.dirname)
                return self
            except FileExistsError:
                time.sleep(0.05)

    def __exit__(self, *a):
        os.rmdir(self.dirname)
        return False
