"""Background expansion rate: the part of ``victor.BackgroundCosmology`` the likelihood path uses.

The reference builds an ``astropy.cosmology.LambdaCDM(H0, Om0, Ode0)`` (victor/cosmology.py:26-45; no radiation,
since astropy's default ``Tcmb0`` is 0) and the model takes ``iaH = (1 + z) / (100 E(z))`` from it
(victor/ccf_model.py:44-45).  Here the same closed form, without astropy.  Distances, growth-rate fits and the
other conveniences of the reference class are outside the B200 path and are not provided.
"""
import numpy as np


class BackgroundCosmology:
    def __init__(self, cosmology=None):
        cosmology = cosmology or {}
        self.c = 299792.458                                    # km/s
        self.OmegaM = cosmology.get("Omega_m", 0.31)
        self.OmegaK = cosmology.get("Omega_K", 0)
        self.OmegaL = 1 - self.OmegaM - self.OmegaK
        self.H0 = cosmology.get("H0", 100 * cosmology.get("h", 0.675))
        self.rd = cosmology.get("sound_horizon", 148.1)
        self.sigma8 = cosmology.get("sigma8", 0.81)

    def Ez(self, z):
        """Normalised Hubble parameter H(z) / H0."""
        zp1 = 1.0 + np.asarray(z, dtype=np.float64)
        curv = 1.0 - self.OmegaM - self.OmegaL                 # astropy derives the curvature from Om0 + Ode0
        out = np.sqrt(self.OmegaM * zp1 ** 3 + curv * zp1 ** 2 + self.OmegaL)
        return float(out) if out.ndim == 0 else out

    def H(self, z):
        """Hubble parameter in km/s/Mpc."""
        return self.H0 * self.Ez(z)

    def Om(self, z):
        """Matter density parameter at redshift z."""
        zp1 = 1.0 + np.asarray(z, dtype=np.float64)
        out = self.OmegaM * zp1 ** 3 / np.asarray(self.Ez(z)) ** 2
        return float(out) if np.ndim(out) == 0 else out
