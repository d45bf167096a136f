"""Seeded synthetic topography for benchmarks and smoke tests (there is no network for real data).

Not a port of pygcm/topography.py (init-time host code, out of scope): a band-limited random field
built from a few hundred Fourier modes, periodic in longitude, thresholded by an area-weighted
quantile to the requested land fraction; albedo/friction follow the reference's conventions
(ocean 0.08 / land 0.30 albedo, ocean 1e-6 / land 1e-5 s^-1 friction)."""
from __future__ import annotations

import numpy as np


def make_topography(nlat, nlon, seed=42, land_frac=0.29, scale_m=4500.0):
    rng = np.random.default_rng(seed)
    lat = np.deg2rad(np.linspace(-90, 90, nlat))[:, None]
    lon = np.deg2rad(np.linspace(0, 360, nlon))[None, :]
    elev = np.zeros((nlat, nlon))
    for _ in range(160):
        m = int(rng.integers(0, 9))
        n = int(rng.integers(1, 9))
        amp = rng.standard_normal() / (1.0 + m * m + n * n) ** 0.75
        ph1, ph2 = rng.uniform(0, 2 * np.pi, 2)
        elev += amp * np.cos(m * lon + ph1) * np.cos(n * lat + ph2) * np.cos(lat) ** (0.5 * (m > 0))
    elev = (elev - elev.mean()) / (elev.std() + 1e-12) * scale_m
    w = np.maximum(np.cos(lat), 0.0) * np.ones((nlat, nlon))
    order = np.argsort(elev, axis=None)
    cw = np.cumsum(w.ravel()[order]) / w.sum()
    sea = elev.ravel()[order][np.searchsorted(cw, 1.0 - land_frac)]
    land = (elev >= sea).astype(np.uint8)
    elevation = np.where(land == 1, elev - sea, 0.0)
    base_albedo = np.where(land == 1, 0.30, 0.08).astype(np.float64)
    friction = np.where(land == 1, 1.0e-5, 1.0e-6).astype(np.float64)
    return dict(land_mask=land, elevation=elevation, base_albedo=base_albedo, friction=friction)


def load_reference_topography(path, nlat, nlon, roll_columns=0):
    """QD_TOPO_NC path of the script (run_simulation.py:1198-1203): a topography file written by the reference's
    ``scripts/generate_topography.py`` read through ``load_topography_from_netcdf`` (bilinear / nearest regrid when the
    grid differs, as the reference's loader does for a coarser file).  ``roll_columns`` rotates the planet in longitude
    (the 0/360 seam column stays a duplicate): distinct but equally reference-made masks for ensemble members."""
    from .grid import SphericalGrid
    from .restart import load_topography_from_netcdf
    elev, mask, alb, fric = load_topography_from_netcdf(path, SphericalGrid(nlat, nlon))
    out = dict(land_mask=np.ascontiguousarray(mask, dtype=np.uint8), elevation=np.ascontiguousarray(elev, dtype=np.float64),
               base_albedo=np.ascontiguousarray(alb, dtype=np.float64), friction=np.ascontiguousarray(fric, dtype=np.float64))
    k = int(roll_columns) % max(nlon - 1, 1)
    if k:
        for name, a in out.items():
            r = np.roll(a[:, :-1], k, axis=1)
            out[name] = np.ascontiguousarray(np.concatenate([r, r[:, :1]], axis=1))
    return out
