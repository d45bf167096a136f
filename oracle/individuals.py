"""oracle/individuals.py -- NumPy restatement of the sub-daily individual pool (test infrastructure only).
Reference: pygcm/ecology/individuals.py:142-191 (try_substep), pygcm/ecology/spectral.py:236-426 (effective
temperatures, Planck band weights, dual_star_insolation_to_bands).  Pinned against tests/golden/indiv_golden.npz."""
import numpy as np

_T_SUN, _h, _c, _kB = 5778.0, 6.62607015e-34, 2.99792458e8, 1.380649e-23
L_SUN, M_SUN = 3.828e26, 1.989e30


def teff(L_ratio, M_ratio, j=0.8):
    """spectral.py:236-246."""
    return float(_T_SUN * (max(L_ratio, 1e-12) ** 0.25) * (max(M_ratio, 1e-12) ** (-0.5 * j)))


def blackbody_band_weights(T, centers, widths):
    """spectral.py:249-285."""
    lam = np.maximum(np.asarray(centers, dtype=float) * 1e-9, 1e-20)
    x = np.clip((_h * _c) / (lam * _kB * max(1e-12, float(T))), 1e-8, 1e3)
    B = np.clip((1.0 / (lam ** 5)) * (1.0 / (np.expm1(x) + 1e-30)), 0.0, np.inf)
    w = B * np.asarray(widths, dtype=float)
    return w / (float(np.sum(w)) + 1e-30)


def insolation_to_bands(insA, insB, specA, specB, T_ray):
    """spectral.py:388-426: per-pixel band intensities [NB, lat, lon]."""
    NB = specA.shape[0]
    I_tot = insA + insB
    I_b = np.zeros((NB,) + insA.shape)
    for b in range(NB):
        I_b[b] = (specA[b] * insA + specB[b] * insB) * T_ray[b]
    S_sum = np.sum(I_b, axis=0)
    pos = (S_sum > 1e-12) & (I_tot > 1e-12)
    if np.any(pos):
        for b in range(NB):
            tmp = np.zeros_like(S_sum)
            tmp[pos] = (I_b[b][pos] / S_sum[pos]) * I_tot[pos]
            I_b[b] = tmp
    else:
        I_b[:] = 0.0
    return np.nan_to_num(I_b, nan=0.0, posinf=0.0, neginf=0.0)


class Pool:
    """State touched by try_substep (individuals.py:142-191)."""

    def __init__(self, sample_j, sample_i, cell_index, Ab, tol, substeps_per_day):
        self.sample_j, self.sample_i, self.cell_index = sample_j, sample_i, cell_index
        self.Ab, self.tol, self.K = Ab, tol, int(substeps_per_day)
        self.E_day = np.zeros(Ab.shape[0])
        self.stress = np.zeros(Ab.shape[0])
        self.period = None
        self.accum = 0.0

    def try_substep(self, insA, insB, spec, soil, dt, day_length):
        if self.period is None:
            self.period = float(day_length) / float(self.K)
            self.accum = 0.0
        self.accum += float(dt)
        if self.accum < self.period:
            return False
        self.accum -= self.period
        I_b = insolation_to_bands(insA, insB, *spec)
        I_cells = I_b[:, self.sample_j, self.sample_i].T
        dE = np.einsum("ij,ij->i", self.Ab, I_cells[self.cell_index, :]) * float(self.period)
        self.E_day += np.maximum(0.0, dE)
        soil_ind = np.asarray(soil, dtype=float)[self.sample_j, self.sample_i][self.cell_index]
        self.stress[soil_ind < self.tol] += float(self.period) / float(day_length)
        return True
