"""oracle/ecology.py -- NumPy restatement of the ecology SUB-DAILY path (test infrastructure only).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arms may import this module; the
product (``qingdai_b200``) never does.  Pinned against ``tests/golden/eco_golden.npz`` (recorded from the
reference by ``tests/golden/make_golden.py eco``) in ``tests/test_oracle_eco.py``.

Restated (reference file:line):
  * spectral bands / band weights / leaf template    pygcm/ecology/spectral.py:23-55,58-85,150-172
  * gene absorbance -> species reflectance           pygcm/ecology/genes.py:50-63,100-113, adapter.py:86-112
  * EcologyAdapter.step_subdaily                     pygcm/ecology/adapter.py:140-186
  * PopulationManager.step_subdaily + canopy cache   pygcm/ecology/population.py:252-286,831-841,895-915
  * effective_leaf_reflectance_bands / bands albedo  pygcm/ecology/population.py:855-892
Daily ecology (LAI growth, spread, seeds, individuals) is out of scope.
"""
from __future__ import annotations

import numpy as np


def make_bands(nbands=16, lam0=380.0, lam1=780.0):
    """spectral.py:23-55 -> (edges, centres, widths)."""
    edges = np.linspace(float(lam0), float(lam1), int(nbands) + 1)
    return edges, 0.5 * (edges[:-1] + edges[1:]), edges[1:] - edges[:-1]


def band_weights(centers, mode="simple", t0=0.9, lref=550.0, eta=4.0):
    """spectral.py:150-172 (rayleigh weight :58-69)."""
    if mode == "rayleigh":
        lam = np.maximum(1e-6, centers)
        w = np.clip(t0 * (lam / max(1e-6, lref)) ** float(eta), 0.0, None)
    else:
        w = np.ones_like(centers, dtype=float)
    return w / (float(np.sum(w)) + 1e-12)


def default_leaf_reflectance(centers):
    """spectral.py:72-85: 0.25 baseline + 0.15 Gaussian bump at 550 nm (sigma 60 nm)."""
    return np.clip(0.25 + 0.15 * np.exp(-((centers - 550.0) ** 2) / (2.0 * 60.0 ** 2)), 0.0, 1.0)


def gene_absorbance(centers, peaks):
    """genes.py:100-113: sum of Gaussian peaks (centre, sigma, height), clipped to [0,1]."""
    A = np.zeros_like(centers, dtype=float)
    for c, w, h in peaks:
        if w <= 0 or h <= 0:
            continue
        A += h * np.exp(-((centers - c) ** 2) / (2 * (w ** 2)))
    return np.clip(A, 0.0, 1.0)


DEFAULT_PEAKS = ((450.0, 40.0, 0.6), (680.0, 30.0, 0.8))            # genes.py:70


class EcoState:
    """Sub-daily state of one EcologyAdapter + PopulationManager pair."""

    def __init__(self, land_mask, lai_layers, alpha_leaf_scalar, k_canopy=0.5, update_every_hours=6.0,
                 lai_delta=0.05, soil_ref=0.20, substep_every=1):
        self.land = (np.asarray(land_mask) == 1)
        self.lai_layers = np.array(lai_layers, dtype=float)             # [S, K, lat, lon]
        self.alpha_leaf_scalar = float(alpha_leaf_scalar)
        self.k_canopy, self.every, self.lai_delta = float(k_canopy), float(update_every_hours), float(lai_delta)
        self.soil_ref, self.substep_every = float(soil_ref), max(1, int(substep_every))
        self.E_day = np.zeros(self.land.shape)
        self.hours = 0.0
        self.next_hours = self.every                                     # population.py:71
        self.f_cached = None
        self.snapshot = self.total_lai().copy()                          # population.py:70
        self.step_count = 0

    def total_lai(self):
        return np.sum(self.lai_layers, axis=(0, 1))                      # population.py:288-292

    def _should_recompute(self):
        """population.py:895-909."""
        if self.f_cached is None:
            return True
        if self.hours >= self.next_hours:
            return True
        now = self.total_lai()
        delta = np.nanmean(np.abs(now - self.snapshot))
        base = np.nanmean(np.maximum(self.snapshot, 1e-6))
        ratio = (delta / base) if base > 0 else delta
        return bool(ratio >= self.lai_delta)

    def _recompute(self):
        self.f_cached = 1.0 - np.exp(-self.k_canopy * np.maximum(self.total_lai(), 0.0))   # population.py:911-915

    def step_subdaily(self, isr, dt):
        """adapter.py:140-186 with the LAI manager present.  Returns the land alpha map (NaN on ocean) or
        None when this call is not on the QD_ECO_SUBSTEP_EVERY_NPHYS cadence."""
        self.step_count += 1
        self.E_day += np.nan_to_num(isr) * float(dt)                     # population.py:268-269
        self.hours += float(dt) / 3600.0
        if self._should_recompute():
            self._recompute()
            self.snapshot = self.total_lai().copy()
            self.next_hours = self.hours + self.every
        if self.step_count % self.substep_every != 0:
            return None
        f = np.where(self.land, self.f_cached, np.nan)                   # population.py:831-841
        alpha = np.full(self.land.shape, np.nan)
        alpha[self.land] = np.clip(self.alpha_leaf_scalar * f[self.land] + (1.0 - f[self.land]) * self.soil_ref, 0.0, 1.0)
        return alpha

    def surface_albedo_bands(self, R_species, species_weights):
        """population.py:855-892 through adapter.get_surface_albedo_bands: A[NB, lat, lon], NaN on ocean."""
        if self.f_cached is None:
            self._recompute()
        R_eff = np.clip(np.tensordot(species_weights, np.clip(R_species, 0.0, 1.0), axes=(0, 0)), 0.0, 1.0)
        f = np.where(self.land, self.f_cached, np.nan)
        A = np.full((R_eff.shape[0],) + self.land.shape, np.nan)
        for b in range(R_eff.shape[0]):
            Ab = R_eff[b] * f + (1.0 - f) * self.soil_ref
            A[b][self.land] = np.clip(Ab[self.land], 0.0, 1.0)
        return A
