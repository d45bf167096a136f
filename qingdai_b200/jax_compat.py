"""Drop-in for the reference's backend seam ``pygcm.jax_compat`` (jax_compat.py:66-216).

The reference imports ``is_enabled, to_numpy, laplacian_sphere, hyperdiffuse, advect_semilag`` from
that module (dynamics.py:15, ocean.py:24, ecology/phyto.py:6).  Pointing those imports here swaps the
three array kernels for the sm_100a ones (host NumPy arrays in, host NumPy arrays out, through the
``*_host`` C entry points); there is no JAX, XLA or NumPy path behind them.
"""
from __future__ import annotations

import numpy as np

from . import constants as const
from .engine import Engine

_engines = {}


def _engine(shape):
    eng = _engines.get(shape)
    if eng is None:
        eng = _engines[shape] = Engine(shape[0], shape[1])
    return eng


def is_enabled() -> bool:
    return True


def backend() -> str:
    return "b200"


def to_numpy(x):
    return x if isinstance(x, np.ndarray) else np.array(x, copy=True)


def _check_grid(eng, dlat, dlon, a):
    if not (np.isclose(dlat, eng.dlat, rtol=1e-12) and np.isclose(dlon, eng.dlon, rtol=1e-12) and a == const.PLANET_RADIUS):
        raise ValueError("qingdai_b200 kernels are built on the reference's regular grid (linspace(-90,90,n), linspace(0,360,n))")


def laplacian_sphere(F, dlat, dlon, coslat, a):
    F = np.asarray(F, dtype=np.float64)
    eng = _engine(F.shape)
    _check_grid(eng, dlat, dlon, a)
    return eng.op_laplacian(F, coslat)


def hyperdiffuse(F, k4, dt, n_substeps, dlat, dlon, coslat, a):
    F = np.asarray(F, dtype=np.float64)
    eng = _engine(F.shape)
    _check_grid(eng, dlat, dlon, a)
    return eng.op_hyperdiffuse(F, k4, dt, n_substeps, coslat)


def advect_semilag(field, u, v, dt, a, dlat, dlon, coslat):
    field = np.asarray(field, dtype=np.float64)
    eng = _engine(field.shape)
    _check_grid(eng, dlat, dlon, a)
    cos = np.asarray(coslat, dtype=np.float64)
    cos = cos[:, 0] if cos.ndim == 2 else cos
    return eng.op_advect(field, u, v, dt, np.maximum(1e-6, cos))        # jax_compat.py:197 floors again
