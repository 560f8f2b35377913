"""``mbproj2.Cosmology`` work-alike: Ned Wright's cosmology calculator (flat or curved LCDM).

Used at set-up only (``kpc_per_arcsec``, reference ``joxsz_main.py:28-31,96``).  Restated from
memory of the public calculator: 1000-point midpoint rule, radiation
density ``4.165e-5/h^2``.
"""
import math


class Cosmology:
    c_km_s = 299792.458

    def __init__(self, z, H0=70.0, WM=0.3, WV=0.7):
        self.z = z
        self.H0 = H0
        self.WM = WM
        self.WV = WV

    def _distances(self):
        h = self.H0 / 100.0
        WR = 4.165e-5 / (h * h)
        WK = 1.0 - self.WM - WR - self.WV
        az = 1.0 / (1.0 + self.z)
        n = 1000
        dcmr = 0.0
        for i in range(n):
            a = az + (1.0 - az) * (i + 0.5) / n
            adot = math.sqrt(WK + self.WM / a + WR / (a * a) + self.WV * a * a)
            dcmr += 1.0 / (a * adot)
        dcmr *= (1.0 - az) / n
        x = math.sqrt(abs(WK)) * dcmr
        if x > 0.1:
            ratio = (0.5 * (math.exp(x) - math.exp(-x)) / x) if WK > 0 else (math.sin(x) / x)
        else:
            y = x * x
            if WK < 0:
                y = -y
            ratio = 1.0 + y / 6.0 + y * y / 120.0
        dcmt = ratio * dcmr
        da = az * dcmt
        dl = da / (az * az)
        scale = self.c_km_s / self.H0
        return da * scale, dl * scale

    @property
    def D_A(self):
        return self._distances()[0]

    @property
    def D_L(self):
        return self._distances()[1]

    @property
    def kpc_per_arcsec(self):
        return self.D_A / 206.264806
