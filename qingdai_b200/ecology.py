"""Ecology SUB-DAILY coupling behind the reference's interface (pygcm/ecology/adapter.py:23-186,
pygcm/ecology/population.py:36-140,252-292,831-915, pygcm/ecology/spectral.py:23-172, genes.py:44-113).

``EcologyAdapter(grid, land_mask)`` keeps the reference's constructor, ``step_subdaily(I_total, cloud_eff,
dt_seconds)``, ``get_surface_albedo_bands()`` and ``pop.*`` attributes, but the per-step work -- daily-energy
accumulation, the canopy-cache policy (two nanmean reductions over the summed LAI layers every step), the
canopy factor and the land alpha map / band albedo -- runs in the CUDA kernels of ``csrc/qd_eco.cuh``.  Inside
the fused loop (``Simulation(with_eco=True)``) nothing of it returns to the host.

Host-only pieces are the O(NB) spectral constants (band edges, band weights, leaf / gene reflectance) which
the reference also evaluates once in ``__init__``.  The DAILY ecology (LAI growth, spread, seeds,
individuals, gene export) is out of scope: a host owner of the LAI layers calls
``pop.set_lai_layers`` when they change.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .engine import Engine, F, S, _ptr


# ------------------------------------------------------------------------------------ spectral constants
@dataclass
class SpectralBands:                                   # spectral.py:8-20
    nbands: int
    lambda_edges: np.ndarray
    lambda_centers: np.ndarray
    delta_lambda: np.ndarray


def _envf(env, key, default):
    try:
        return float(env.get(key, str(default)))
    except (TypeError, ValueError):
        return float(default)


def make_bands(env=None) -> SpectralBands:
    """QD_ECO_SPECTRAL_BANDS equal bands over QD_ECO_SPECTRAL_RANGE_NM (spectral.py:23-55)."""
    env = os.environ if env is None else env
    try:
        nb = max(1, int(env.get("QD_ECO_SPECTRAL_BANDS", "16")))
    except ValueError:
        nb = 16
    try:
        lo, hi = [float(x.strip()) for x in env.get("QD_ECO_SPECTRAL_RANGE_NM", "380,780").split(",")]
    except ValueError:
        lo, hi = 380.0, 780.0
    if hi <= lo:
        lo, hi = 380.0, 780.0
    edges = np.linspace(lo, hi, nb + 1)
    return SpectralBands(nb, edges, 0.5 * (edges[:-1] + edges[1:]), edges[1:] - edges[:-1])


def band_weights_from_mode(bands: SpectralBands, env=None) -> np.ndarray:
    """Normalised band weights, 'simple' (flat) or 'rayleigh' (spectral.py:150-172, :58-69)."""
    env = os.environ if env is None else env
    mode = env.get("QD_ECO_TOA_TO_SURF_MODE", "simple").strip().lower()
    lam = bands.lambda_centers
    if mode == "rayleigh":
        t0, lref, eta = _envf(env, "QD_ECO_RAYLEIGH_T0", 0.9), _envf(env, "QD_ECO_RAYLEIGH_LREF_NM", 550.0), _envf(env, "QD_ECO_RAYLEIGH_ETA", 4.0)
        w = np.clip(t0 * (np.maximum(1e-6, lam) / max(1e-6, lref)) ** float(eta), 0.0, None)
    else:
        w = np.ones_like(lam, dtype=float)
    return w / (float(np.sum(w)) + 1e-12)


def default_leaf_reflectance(bands: SpectralBands) -> np.ndarray:
    """Green-ish leaf template (spectral.py:72-85)."""
    lam = bands.lambda_centers
    return np.clip(0.25 + 0.15 * np.exp(-((lam - 550.0) ** 2) / (2.0 * 60.0 ** 2)), 0.0, 1.0)


def gene_peaks_from_env(prefix, env=None):
    """``center:width:height`` triples of QD_ECO_*_PEAKS; the default two-band absorber (genes.py:56-70)."""
    env = os.environ if env is None else env
    peaks = []
    for part in env.get(prefix + "PEAKS", "").split(","):
        try:
            c, w, h = part.strip().split(":")
            peaks.append((float(c), float(w), float(h)))
        except ValueError:
            continue
    return peaks or [(450.0, 40.0, 0.6), (680.0, 30.0, 0.8)]


def reflectance_from_peaks(bands: SpectralBands, peaks) -> np.ndarray:
    """R = 1 - clip(sum of Gaussian absorption peaks) (genes.py:100-120)."""
    lam = bands.lambda_centers
    A = np.zeros_like(lam, dtype=float)
    for c, w, h in peaks:
        if w <= 0 or h <= 0:
            continue
        A += h * np.exp(-((lam - c) ** 2) / (2 * (w ** 2)))
    return np.clip(1.0 - np.clip(A, 0.0, 1.0), 0.0, 1.0)


# ------------------------------------------------------------------------------------ device-side population state
class PopulationState:
    """The sub-daily face of PopulationManager (population.py:36-140): LAI layers [S, K, lat, lon] in HBM,
    E_day, canopy cache and its clock on the device."""

    def __init__(self, engine: Engine, land_mask, env=None, member=0):
        env = os.environ if env is None else env
        self.engine, self.member = engine, member
        self.land = (np.asarray(land_mask) == 1)
        self.shape = self.land.shape
        self.k_canopy = _envf(env, "QD_ECO_LAI_K", 0.5)
        self.light_update_every_hours = _envf(env, "QD_ECO_LIGHT_UPDATE_EVERY_HOURS", 6.0)
        self.lai_recompute_delta = _envf(env, "QD_ECO_LIGHT_RECOMPUTE_LAI_DELTA", 0.05)
        self.K = max(1, int(_envf(env, "QD_ECO_COHORT_K", 1)))
        wenv = env.get("QD_ECO_SPECIES_WEIGHTS", "").strip()
        if wenv:
            w = [float(x) for x in wenv.split(",") if x.strip() != ""]
        else:
            ns = max(1, int(_envf(env, "QD_ECO_NS", 20)))
            w = [1.0 / float(ns)] * ns
        s = sum(w) if w else 1.0
        self.species_weights = np.asarray([max(0.0, x) for x in w], dtype=float)
        if s <= 0:
            ns = max(1, int(_envf(env, "QD_ECO_NS", 20)))
            self.species_weights = np.full((ns,), 1.0 / float(ns))
        else:
            self.species_weights /= s
        self.Ns = int(self.species_weights.shape[0])
        LAI = np.zeros(self.shape)
        LAI[self.land] = _envf(env, "QD_ECO_LAI_INIT", 0.2)
        layers = np.zeros((self.Ns, self.K) + self.shape)
        for si in range(self.Ns):                                         # population.py:117-122
            for k in range(self.K):
                layers[si, k] = float(self.species_weights[si]) * (LAI / float(self.K))
        self._species_R_leaf = None
        self._lai_dev = None
        self.set_lai_layers(layers, reset_clock=True)

    # -- LAI hand-over from the (host) daily ecology
    def set_lai_layers(self, layers, reset_clock=False):
        layers = np.ascontiguousarray(np.asarray(layers, dtype=np.float64))
        assert layers.shape[2:] == self.shape
        e = self.engine
        if e.batch != 1:
            raise ValueError("PopulationState drives one ensemble member per engine")
        self.Ns, self.K = layers.shape[0], layers.shape[1]
        flat = torch.from_numpy(layers.reshape(1, self.Ns * self.K, *self.shape))
        if self._lai_dev is None or tuple(self._lai_dev.shape) != tuple(flat.shape):
            self._lai_dev = flat.to(e.device).contiguous()
            e.bind_eco(self._lai_dev, self.Ns * self.K, self.k_canopy, self.light_update_every_hours,
                       self.lai_recompute_delta, getattr(self, "_substep_every", 1))
        else:
            self._lai_dev.copy_(flat)
        if reset_clock:
            e.eco_reset(0.0, self.light_update_every_hours, cached=False, step_count=0)   # population.py:57-71

    @property
    def LAI_layers_SK(self):
        return self._lai_dev.cpu().numpy().reshape(self.Ns, self.K, *self.shape)

    @LAI_layers_SK.setter
    def LAI_layers_SK(self, value):
        self.set_lai_layers(value)

    def total_LAI(self):
        return np.sum(self.LAI_layers_SK, axis=(0, 1))

    @property
    def E_day(self):
        return self.engine.get("eday", self.member)

    @E_day.setter
    def E_day(self, value):
        self.engine.set("eday", value, self.member)

    def clock(self):
        s = self.engine.scalars()[self.member]
        return float(s[S["eco_hours"]]), float(s[S["eco_next"]])

    def canopy_reflectance_factor(self):
        """f(LAI) on land, NaN on ocean (population.py:831-841); the cache is the device field."""
        f = self.engine.get("fcanopy", self.member)
        return np.where(self.land, f, np.nan)

    def lai_snapshot(self):
        return self.engine.get("lai_snap", self.member)

    def set_species_reflectance_bands(self, R):
        R = np.asarray(R, dtype=float)
        self._species_R_leaf = np.clip(R, 0.0, 1.0) if R.ndim == 2 else None

    def effective_leaf_reflectance_bands(self, nb):
        """population.py:855-873."""
        if self._species_R_leaf is None:
            return np.full((nb,), 0.5)
        Ns, NB = self._species_R_leaf.shape
        if NB != nb:
            return np.full((nb,), float(np.nanmean(self._species_R_leaf)))
        w = self.species_weights if self.species_weights.size == Ns else np.full((Ns,), 1.0 / max(1, Ns))
        return np.clip(np.tensordot(w, self._species_R_leaf, axes=(0, 0)), 0.0, 1.0)


# ------------------------------------------------------------------------------------ adapter
class EcologyAdapter:
    """Drop-in for pygcm.ecology.EcologyAdapter's sub-daily face (adapter.py:23-186, :276-300)."""

    def __init__(self, grid, land_mask, engine: Optional[Engine] = None, env=None, lib=None, device=None):
        env = dict(os.environ) if env is None else dict(env)
        self.grid = grid
        self.land_mask = (np.asarray(land_mask) == 1)
        nlat, nlon = self.land_mask.shape
        self.substep_every_nphys = max(1, int(_envf(env, "QD_ECO_SUBSTEP_EVERY_NPHYS", 1)))
        self.lai_albedo_weight = _envf(env, "QD_ECO_LAI_ALBEDO_WEIGHT", 1.0)
        self.soil_reflect = _envf(env, "QD_ECO_SOIL_REFLECT", 0.20)
        self.bands = make_bands(env)
        self.w_b = band_weights_from_mode(self.bands, env)
        self.R_leaf = default_leaf_reflectance(self.bands)
        self.alpha_leaf_scalar = float(np.sum(self.R_leaf * self.w_b))    # adapter.py:57-60
        self._own_engine = engine is None
        if engine is None:
            engine = Engine(nlat, nlon, batch=1, device=device, lib=lib)
            engine.set_mask("land", self.land_mask.astype(np.uint8))
        self.engine = engine
        for p in engine.params:
            p.eco_lai_albedo_weight, p.eco_soil_reflect = self.lai_albedo_weight, self.soil_reflect
        engine.set_eco(True, self.alpha_leaf_scalar)
        use_lai = int(_envf(env, "QD_ECO_USE_LAI", 1)) == 1
        self.pop = None
        if use_lai:
            self.pop = PopulationState.__new__(PopulationState)
            self.pop._substep_every = self.substep_every_nphys
            PopulationState.__init__(self.pop, engine, self.land_mask, env)
            # adapter.py:93-100: every species reads its own QD_ECO_SPECIES_{i}_* genes (defaults when unset)
            R = [reflectance_from_peaks(self.bands, gene_peaks_from_env(f"QD_ECO_SPECIES_{i}_", env)) for i in range(self.pop.Ns)]
            self.species_drought_tol = [_envf(env, f"QD_ECO_SPECIES_{i}_DROUGHT_TOL", 0.3) for i in range(self.pop.Ns)]   # genes.py:84
            self.pop.set_species_reflectance_bands(np.stack(R, axis=0))   # adapter.py:86-112
        else:                                                            # M1: alpha = clip(alpha_leaf_scalar) on land
            engine.bind_eco(None, 0, 0.5, 6.0, 0.05, self.substep_every_nphys)
            engine.set("fcanopy", np.ones(engine.shape))
            engine.eco_reset(0.0, 6.0, cached=True, step_count=0)
        self._alpha_dev = torch.empty((1, nlat, nlon), dtype=torch.float64, device=engine.device)
        self._isr_dev = torch.empty((1, nlat, nlon), dtype=torch.float64, device=engine.device)

    def step_subdaily(self, I_total, cloud_eff, dt_seconds):
        """Host-array form (adapter.py:140-186): upload isr, run the device path, download alpha (or None)."""
        e = self.engine
        isr = np.array(np.broadcast_to(np.asarray(I_total, dtype=np.float64), e.shape), dtype=np.float64, order="C")
        self._isr_dev.copy_(torch.from_numpy(isr).reshape(self._isr_dev.shape))
        produced = C.c_int(0)
        e._chk(e.lib.qd_eco_subdaily(e.ctx, _ptr(self._isr_dev), float(dt_seconds), _ptr(self._alpha_dev), C.byref(produced)), "qd_eco_subdaily")
        if not produced.value:
            return None
        e.sync()
        return self._alpha_dev[0].cpu().numpy()

    def get_surface_albedo_bands(self):
        """(A[NB, lat, lon], w_b) from the cached canopy factor (adapter.py:276-300, population.py:875-892)."""
        if self.pop is None:
            return None, None
        e = self.engine
        nb = int(self.bands.nbands)
        r_eff = np.ascontiguousarray(self.pop.effective_leaf_reflectance_bands(nb))
        out = torch.empty((1, nb) + tuple(e.shape), dtype=torch.float64, device=e.device)
        e._chk(e.lib.qd_eco_bands(e.ctx, nb, _ptr(r_eff), float(self.soil_reflect), _ptr(out)), "qd_eco_bands")
        e.sync()
        return out[0].cpu().numpy(), self.w_b.copy()
