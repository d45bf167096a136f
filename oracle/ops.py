"""Numeric operators of the Qingdai step, restated in plain NumPy (oracle; test-only).

Every function cites the reference call site it follows (paths relative to the
reference checkout).  SciPy is *not* imported: the three scipy.ndimage routines the
reference uses are restated from their published algorithms (scipy 1.18.1,
``ndimage/src/ni_interpolation.c`` and ``ni_filters.c``) and are bit-identical to
SciPy on every case in ``tests/test_oracle_ops.py``.
"""
from __future__ import annotations

import numpy as np

DBL_MAX = np.finfo(np.float64).max


# --------------------------------------------------------------------------- hygiene
def nan_to_num(x):
    """np.nan_to_num semantics (NaN->0, +-inf->+-DBL_MAX); dynamics.py:661-667."""
    return np.nan_to_num(np.asarray(x, dtype=np.float64))


# --------------------------------------------------------------------------- gather
def _wrap_coord_legacy(x, n):
    """scipy ``mode='wrap'`` coordinate map (period n-1), ni_interpolation.c map_coordinate."""
    x = np.array(x, dtype=np.float64, copy=True)
    if n <= 1:
        return np.zeros_like(x)
    sz = float(n - 1)
    neg = x < 0
    x[neg] = x[neg] + sz * (np.trunc(-x[neg] / sz) + 1.0)
    big = x > (n - 1)
    x[big] = x[big] - sz * np.trunc(x[big] / sz)
    return x


def _wrap_index_legacy(k, n):
    k = np.array(k, dtype=np.int64, copy=True)
    if n <= 1:
        return np.zeros_like(k)
    sz = n - 1
    neg = k < 0
    k[neg] = k[neg] + sz * ((-k[neg]) // sz + 1)
    big = k > n - 1
    k[big] = k[big] - sz * (k[big] // sz)
    return k


def bilinear_wrap(F, J, I):
    """``map_coordinates(F,[J,I],order=1,mode='wrap',prefilter=False)``.

    SciPy's order-1 weights are w0 = 1 - t and w1 = 1 - w0 (NOT t: the last spline weight is
    formed as one minus the others), and the accumulation order is t = v00*wy0*wx0;
    t += v01*wy0*wx1; t += v10*wy1*wx0; t += v11*wy1*wx1
    (dynamics.py:117, ocean.py:193, run_simulation.py:1157).
    """
    F = np.asarray(F, dtype=np.float64)
    nj, ni = F.shape
    y = _wrap_coord_legacy(J, nj)
    x = _wrap_coord_legacy(I, ni)
    fy = np.floor(y)
    fx = np.floor(x)
    j0 = fy.astype(np.int64)
    i0 = fx.astype(np.int64)
    ty = y - fy
    tx = x - fx
    j1 = _wrap_index_legacy(j0 + 1, nj)
    i1 = _wrap_index_legacy(i0 + 1, ni)
    j0 = _wrap_index_legacy(j0, nj)
    i0 = _wrap_index_legacy(i0, ni)
    wy0 = 1.0 - ty
    wy1 = 1.0 - wy0
    wx0 = 1.0 - tx
    wx1 = 1.0 - wx0
    t = F[j0, i0] * wy0 * wx0
    t = t + F[j0, i1] * wy0 * wx1
    t = t + F[j1, i0] * wy1 * wx0
    t = t + F[j1, i1] * wy1 * wx1
    return t


def advect_semilag(F, u, v, dt, a, dlat, dlon, cos_rows):
    """Semi-Lagrangian bilinear gather (dynamics.py:90-118 == jax_compat.py:190-216).

    ``cos_rows`` is the per-row metric cosine *already floored* by the caller
    (1e-6 atmosphere, 0.5 ocean / script cloud tracer).
    """
    F = np.asarray(F, dtype=np.float64)
    nj, ni = F.shape
    c = np.asarray(cos_rows, dtype=np.float64).reshape(nj, 1)
    dx = (u * dt / (a * c)) / dlon
    dy = (v * dt / a) / dlat
    JJ, II = np.meshgrid(np.arange(nj), np.arange(ni), indexing="ij")
    return bilinear_wrap(F, JJ - dy, II - dx)


# --------------------------------------------------------------------------- stencils
def grad_rows(F, h):
    """np.gradient(F, h, axis=0), edge_order=1 (dynamics.py:167-168, :489)."""
    F = np.asarray(F, dtype=np.float64)
    out = np.empty_like(F)
    out[1:-1] = (F[2:] - F[:-2]) / (2.0 * h)
    out[0] = (F[1] - F[0]) / h
    out[-1] = (F[-1] - F[-2]) / h
    return out


def grad_cols(F, h):
    """np.gradient(F, h, axis=1): one-sided at columns 0 and n-1 (dynamics.py:488)."""
    return grad_rows(np.asarray(F).T, h).T


def laplacian(F, dlat, dlon, cos_rows, a):
    """Divergence-form spherical Laplacian (dynamics.py:144-173, ocean.py:100-117)."""
    F = nan_to_num(F)
    c = np.asarray(cos_rows, dtype=np.float64).reshape(-1, 1)
    G = grad_rows(F, dlat)
    term_phi = (1.0 / c) * grad_rows(c * G, dlat)
    d2 = (np.roll(F, -1, axis=1) - 2.0 * F + np.roll(F, 1, axis=1)) / (dlon ** 2)
    term_lam = d2 / (c ** 2)
    return (term_phi + term_lam) / (a ** 2)


def hyperdiffuse(F, k4, dt, nsub, dlat, dlon, cos_rows, a):
    """F <- F - k4 * lap(lap F) * dt/n, n times (dynamics.py:175-212, ocean.py:119-152).

    ``k4`` is a scalar or an array broadcastable to F.  Early-outs mirror the reference.
    """
    if dt <= 0.0:
        return F
    if np.isscalar(k4):
        k4 = float(k4)
        if k4 <= 0.0:
            return F
    else:
        k4 = nan_to_num(k4)
        if np.all(k4 <= 0.0):
            return F
    n = max(1, int(nsub))
    sub = dt / n
    out = nan_to_num(F)
    for _ in range(n):
        L2 = laplacian(laplacian(out, dlat, dlon, cos_rows, a), dlat, dlon, cos_rows, a)
        out = out - k4 * L2 * sub
    return nan_to_num(out)


def shapiro(F, n=2):
    """n x [1-2-1 along lon (periodic, period n_lon) then 1-2-1 along lat (edge replicate)].

    scipy.ndimage.convolve accumulates (x[k-1]*.25 + x[k]*.5) + x[k+1]*.25
    (dynamics.py:215-231, ocean.py:154-164).
    """
    out = nan_to_num(F)
    for _ in range(max(1, int(n))):
        out = (np.roll(out, 1, axis=1) * 0.25 + out * 0.5) + np.roll(out, -1, axis=1) * 0.25
        up = np.vstack([out[:1], out[:-1]])
        dn = np.vstack([out[1:], out[-1:]])
        out = (up * 0.25 + out * 0.5) + dn * 0.25
    return out


def gaussian_weights(sigma, truncate=4.0):
    """scipy ``_gaussian_kernel1d``: radius int(truncate*sigma+0.5), normalised weights."""
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return r, w / w.sum()


def _extend_index(n, r, mode):
    idx = np.arange(-r, n + r)
    if mode == "reflect":          # d c b a | a b c d | d c b a
        p = 2 * n
        m = np.mod(idx, p)
        return np.where(m >= n, p - 1 - m, m)
    if mode == "wrap":
        return np.mod(idx, n)
    if mode == "nearest":
        return np.clip(idx, 0, n - 1)
    raise ValueError(mode)


def _correlate1d_symmetric(F, w, r, axis, mode):
    """ni_filters.c NI_Correlate1D symmetric branch: centre first, then pairs outermost->innermost."""
    F = np.moveaxis(np.asarray(F, dtype=np.float64), axis, 0)
    n = F.shape[0]
    E = F[_extend_index(n, r, mode)]
    out = E[r:r + n] * w[r]
    for jj in range(-r, 0):
        out = out + (E[r + jj:r + jj + n] + E[r - jj:r - jj + n]) * w[r + jj]
    return np.moveaxis(out, 0, axis)


def gaussian(F, sigma, mode="reflect"):
    """scipy.ndimage.gaussian_filter (axis 0 then axis 1); physics.py:44,69,111,159,330
    use the default 'reflect' on both axes, run_simulation.py:1931 uses 'wrap'."""
    if not sigma or sigma <= 1e-15:
        return np.asarray(F, dtype=np.float64)
    r, w = gaussian_weights(sigma)
    o = _correlate1d_symmetric(F, w, r, 0, mode)
    return _correlate1d_symmetric(o, w, r, 1, mode)


def zonal_bandstop(F, cutoff=0.75, damp=0.5):
    """Per-row rfft, bins >= kcut scaled by (1-damp), irfft (dynamics.py:233-258)."""
    cutoff = float(cutoff)
    damp = float(damp)
    arr = nan_to_num(F)
    if damp <= 0.0 or cutoff <= 0.0:
        return arr
    n = arr.shape[1]
    spec = np.fft.rfft(arr, axis=1)
    bins = spec.shape[1]
    if bins <= 1:
        return arr
    kN = bins - 1
    kcut = int(max(1, min(kN, int(cutoff * kN))))
    fac = np.ones(bins)
    fac[kcut:] *= max(0.0, 1.0 - min(1.0, damp))
    spec = spec * fac[None, :]
    return nan_to_num(np.fft.irfft(spec, n=n, axis=1))


def divergence(u, v, lat_deg, dlat, dlon, a):
    """grid.py:41-68 (np.roll both axes; phi-term rows 0,n-1 zeroed; /(a*max(cos,1e-6)))."""
    cosl = np.cos(np.deg2rad(np.asarray(lat_deg, dtype=np.float64))).reshape(-1, 1)
    capped = np.maximum(cosl, 1e-6)
    du = (np.roll(u, -1, axis=1) - np.roll(u, 1, axis=1)) / (2 * dlon)
    vc = v * cosl
    dv = (np.roll(vc, -1, axis=0) - np.roll(vc, 1, axis=0)) / (2 * dlat)
    dv[0, :] = 0
    dv[-1, :] = 0
    return (1 / (a * capped)) * (du + dv)


def vorticity(u, v, lat_deg, dlat, dlon, a):
    """grid.py:70-88."""
    cosl = np.cos(np.deg2rad(np.asarray(lat_deg, dtype=np.float64))).reshape(-1, 1)
    capped = np.maximum(cosl, 1e-6)
    dv = (np.roll(v, -1, axis=1) - np.roll(v, 1, axis=1)) / (2 * dlon)
    uc = u * cosl
    du = (np.roll(uc, -1, axis=0) - np.roll(uc, 1, axis=0)) / (2 * dlat)
    du[0, :] = 0
    du[-1, :] = 0
    return (1 / (a * capped)) * (dv - du)


# --------------------------------------------------------------------------- reductions
def median_pos(x, empty=np.nan):
    """np.median(x[x>0]) with the caller's empty-set fallback (physics.py:298-304,
    run_simulation.py:1867-1875, dynamics.py:344-348)."""
    x = np.asarray(x, dtype=np.float64)
    p = x[x > 0]
    return float(np.median(p)) if p.size > 0 else float(empty)


def area_weights(lat_deg):
    """max(cos(lat),0) row weights (energy.py:520-521, hydrology.py:263-265)."""
    return np.maximum(np.cos(np.deg2rad(np.asarray(lat_deg, dtype=np.float64))), 0.0)


def wmean(x, w_rows):
    """sum(x*w)/(sum(w)+1e-15) with w broadcast over columns (energy.py:524-525)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.broadcast_to(np.asarray(w_rows, dtype=np.float64).reshape(-1, 1), x.shape)
    return float(np.sum(x * w) / (np.sum(w) + 1e-15))
