"""Sub-daily face of the reference's individual pool (pygcm/ecology/individuals.py:23-191) on the device.

``IndividualPool(grid, land_mask, eco_adapter)`` samples land cells and draws individuals exactly like the reference
(same NumPy generator calls, seed 42), keeps their per-band weights in HBM and runs ``try_substep`` -- the NB-band split
of the dual-star insolation (spectral.py:388-426) and the per-individual energy / water-stress accumulation -- as one
kernel (csrc/qd_indiv.cuh).  The end-of-day aggregation (``step_daily``) is host ecology and out of scope; it can read
and reset ``indiv_E_day`` / ``indiv_water_stress_days`` through the properties below."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import constants as const
from .engine import _ptr

_T_SUN, _H, _C, _KB = 5778.0, 6.62607015e-34, 2.99792458e8, 1.380649e-23


def estimate_teff_from_LM(L_ratio, M_ratio, j=0.8):
    """spectral.py:236-246."""
    return float(_T_SUN * (float(max(L_ratio, 1e-12)) ** 0.25) * (float(max(M_ratio, 1e-12)) ** (-0.5 * j)))


def blackbody_band_weights(T_eff, bands):
    """spectral.py:249-285: Planck radiance at the band centres x band width, normalised to sum 1."""
    lam = np.maximum(np.asarray(bands.lambda_centers, dtype=float) * 1e-9, 1e-20)
    x = np.clip((_H * _C) / (lam * _KB * max(1e-12, float(T_eff))), 1e-8, 1e3)
    B = np.clip((1.0 / (lam ** 5)) * (1.0 / (np.expm1(x) + 1e-30)), 0.0, np.inf)
    w = B * np.asarray(bands.delta_lambda, dtype=float)
    return w / (float(np.sum(w)) + 1e-30)


def star_band_spectra(bands, env=None):
    """(specA, specB, T_ray) of dual_star_insolation_to_bands (spectral.py:330-386)."""
    env = os.environ if env is None else env
    T = []
    for star, L, M in (("A", const.L_A, const.M_A), ("B", const.L_B, const.M_B)):
        t = env.get(f"QD_STAR_{star}_TEFF_K")
        T.append(float(t) if t else estimate_teff_from_LM(L / const.L_SUN, M / const.M_SUN, j=float(env.get(f"QD_STAR_{star}_J", "0.8"))))
    lam = np.asarray(bands.lambda_centers, dtype=float)
    if env.get("QD_ECO_TOA_TO_SURF_MODE", "simple").strip().lower() == "rayleigh":
        t0, lref, eta = float(env.get("QD_ECO_RAYLEIGH_T0", "0.9")), float(env.get("QD_ECO_RAYLEIGH_LREF_NM", "550")), float(env.get("QD_ECO_RAYLEIGH_ETA", "4.0"))
        T_ray = np.clip(t0 * (np.maximum(1e-6, lam) / max(1e-6, lref)) ** float(eta), 0.0, None)
    else:
        T_ray = np.ones(lam.shape[0])
    return blackbody_band_weights(T[0], bands), blackbody_band_weights(T[1], bands), np.clip(T_ray, 0.0, np.inf)


class IndividualPool:
    def __init__(self, grid, land_mask, eco_adapter, *, sample_frac=0.02, per_cell=100, substeps_per_day=10, diag=True, env=None):
        env = os.environ if env is None else env
        self.grid = grid
        self.land_mask = (np.asarray(land_mask) == 1)
        self.h, self.w = self.land_mask.shape
        self.sample_frac = float(env.get("QD_ECO_INDIV_SAMPLE_FRAC", str(sample_frac)))
        self.per_cell = int(env.get("QD_ECO_INDIV_PER_CELL", str(per_cell)))
        self.substeps_per_day = max(1, int(env.get("QD_ECO_INDIV_SUBSTEPS_PER_DAY", str(substeps_per_day))))
        self.bands = eco_adapter.bands
        self.nb = int(self.bands.nbands)
        pop = getattr(eco_adapter, "pop", None)
        if pop is None:
            raise RuntimeError("IndividualPool requires EcologyAdapter.pop")
        spw = np.asarray(pop.species_weights, dtype=float)
        self.ns = int(spw.size)
        ssum = float(np.sum(spw))
        self.sp_weights = (spw / ssum) if ssum > 0 else np.full((self.ns,), 1.0 / float(self.ns))
        # sampling and draws in the reference's order (individuals.py:76-113)
        land_idx = np.flatnonzero(self.land_mask.ravel())
        n_land = int(land_idx.size)
        n_cells_sampled = max(1, int(self.sample_frac * n_land))
        rng = np.random.default_rng(seed=42)
        sampled = land_idx if n_cells_sampled >= n_land else rng.choice(land_idx, size=n_cells_sampled, replace=False)
        self.sample_j = np.asarray(sampled // self.w, dtype=np.int32)
        self.sample_i = np.asarray(sampled % self.w, dtype=np.int32)
        self.n_cells = int(self.sample_j.size)
        self.n_indiv = int(self.n_cells * self.per_cell)
        self.indiv_cell_index = np.repeat(np.arange(self.n_cells, dtype=np.int32), self.per_cell)
        self.indiv_species_id = rng.choice(np.arange(self.ns, dtype=np.int32), size=self.n_indiv, p=self.sp_weights)
        species_R = getattr(pop, "_species_R_leaf", None)
        if species_R is None or species_R.shape[0] != self.ns:
            species_R = np.full((self.ns, self.nb), 0.5)
        if species_R.shape[1] != self.nb:
            species_R = species_R[:, :self.nb] if species_R.shape[1] > self.nb else np.pad(species_R, ((0, 0), (0, self.nb - species_R.shape[1])), mode="edge")
        Ab = species_R[self.indiv_species_id, :] + rng.normal(0.0, 0.02, size=(self.n_indiv, self.nb))
        self.indiv_Ab = np.clip(Ab, 0.0, 1.0)
        tol = np.asarray(getattr(eco_adapter, "species_drought_tol", [0.5] * self.ns), dtype=float)
        self.species_drought_tol = np.clip(tol if tol.size == self.ns else np.full((self.ns,), 0.5), 0.0, 1.0)
        self.indiv_tol = self.species_drought_tol[self.indiv_species_id]
        self._substep_period = None
        self._substep_accum = 0.0
        # device tables
        self._engine = e = eco_adapter.engine
        specA, specB, T_ray = star_band_spectra(self.bands, env)
        cell = np.ascontiguousarray((self.sample_j.astype(np.int64) * self.w + self.sample_i)[self.indiv_cell_index].astype(np.int32))
        ab = np.ascontiguousarray(self.indiv_Ab, dtype=np.float64)
        tl = np.ascontiguousarray(self.indiv_tol, dtype=np.float64)
        e._chk(e.lib.qd_indiv_setup(e.ctx, self.n_indiv, self.nb, _ptr(cell), _ptr(ab), _ptr(tl),
                                    _ptr(np.ascontiguousarray(specA)), _ptr(np.ascontiguousarray(specB)), _ptr(np.ascontiguousarray(T_ray))), "qd_indiv_setup")
        self._soil_dev = torch.zeros((self.h, self.w), dtype=torch.float64, device=e.device)
        if diag and int(env.get("QD_ECO_DIAG", "1")) == 1:
            print(f"[EcoIndiv] initialized: cells={self.n_cells}, per_cell={self.per_cell}, N={self.n_indiv}, NB={self.nb}, K={self.substeps_per_day}")

    def _state(self):
        e = self._engine
        a, b = np.empty(self.n_indiv), np.empty(self.n_indiv)
        e._chk(e.lib.qd_indiv_state(e.ctx, _ptr(a), _ptr(b), 0), "qd_indiv_state")
        return a, b

    @property
    def indiv_E_day(self):
        return self._state()[0]

    @property
    def indiv_water_stress_days(self):
        return self._state()[1]

    def set_state(self, E_day, stress_days):
        e = self._engine
        a = np.ascontiguousarray(E_day, dtype=np.float64)
        b = np.ascontiguousarray(stress_days, dtype=np.float64)
        e._chk(e.lib.qd_indiv_state(e.ctx, _ptr(a), _ptr(b), 1), "qd_indiv_state")

    def try_substep(self, isr_A, isr_B, eco_adapter, soil_W_land, dt_seconds, day_length_seconds):
        """individuals.py:142-191.  ``isr_A`` / ``isr_B`` = None uses the engine's own per-star insolation fields (the
        fused loop leaves them in HBM); host arrays are uploaded like the reference passes them."""
        if self._substep_period is None:
            self._substep_period = float(day_length_seconds) / float(self.substeps_per_day)
            self._substep_accum = 0.0
        self._substep_accum += float(dt_seconds)
        if self._substep_accum < self._substep_period:
            return False
        self._substep_accum -= self._substep_period
        e = self._engine
        if isr_A is not None:
            e.set("isr_a", np.asarray(isr_A, dtype=np.float64))
            e.set("isr_b", np.asarray(isr_B, dtype=np.float64))
        soil_ptr, soil_scalar = None, 0.0
        if soil_W_land is None:
            pass
        elif np.isscalar(soil_W_land):
            soil_scalar = float(soil_W_land)
        else:
            soil = np.asarray(soil_W_land, dtype=np.float64)
            if soil.shape != (self.h, self.w):
                soil_scalar = float(np.nanmean(soil))
            else:
                self._soil_dev.copy_(torch.from_numpy(np.array(soil, order="C")))
                soil_ptr = _ptr(self._soil_dev)
        e._chk(e.lib.qd_indiv_substep(e.ctx, soil_ptr, soil_scalar, float(self._substep_period), float(day_length_seconds)), "qd_indiv_substep")
        return True
