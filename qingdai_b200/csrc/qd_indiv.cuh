// qd_indiv.cuh -- sub-daily step of the individual pool (pygcm/ecology/individuals.py:142-191) with the
// NB-band split of the dual-star insolation (spectral.dual_star_insolation_to_bands, spectral.py:388-426).
// The reference builds I_b[NB, lat, lon] for the whole grid and then gathers the sampled cells; only the sampled cells
// are ever read, so one thread per individual evaluates the band split of ITS cell from the two insolation fields:
//   S_b = (specA_b insA + specB_b insB) T_ray_b;  I_b = S_b / sum(S) * (insA + insB) where both exceed 1e-12, else 0;
//   E_day += max(0, dot(Ab_i, I_b) * period);  stress_days += period / day where soil(cell) < tolerance_i.
#pragma once
#include "qd_ocean.cuh"

#define QD_INDIV_MAX_BANDS 32

struct QdIndivArgs {
  int n, nb;
  const int* cell;            // [n] flat cell index j * nlon + i of the individual's sampled cell
  const double* ab;           // [n][nb] per-band absorbance / reflectance weights
  const double* tol;          // [n]
  double *e_day, *stress;     // [n]
  const double *isr_a, *isr_b, *soil;   // fields [nlat][nlon]; soil may be null -> soil_scalar
  double soil_scalar, period, stress_inc;
  double spec_a[QD_INDIV_MAX_BANDS], spec_b[QD_INDIV_MAX_BANDS], t_ray[QD_INDIV_MAX_BANDS];
};

__global__ void __launch_bounds__(QD_THREADS) k_indiv_substep(QdIndivArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n) return;
  const int c = A.cell[t];
  const double ia = A.isr_a[c], ib = A.isr_b[c];
  const double itot = ia + ib;
  double S[QD_INDIV_MAX_BANDS];
  double ssum = 0.0;
  for (int k = 0; k < A.nb; ++k) { S[k] = (A.spec_a[k] * ia + A.spec_b[k] * ib) * A.t_ray[k]; ssum = (k == 0) ? S[0] : ssum + S[k]; }   // np.sum(axis=0): slice by slice
  const bool pos = (ssum > 1e-12) && (itot > 1e-12);
  const double* ab = A.ab + (size_t)t * A.nb;
  double dE = 0.0;
  for (int k = 0; k < A.nb; ++k) {
    double v = pos ? (S[k] / ssum) * itot : 0.0;
    if (!(fabs(v) <= DBL_MAX)) v = 0.0;                    // nan_to_num(nan=0, posinf=0, neginf=0)
    dE = (k == 0) ? ab[0] * v : dE + ab[k] * v;
  }
  dE = dE * A.period;
  A.e_day[t] = A.e_day[t] + qd_max(0.0, dE);
  const double soil = A.soil ? A.soil[c] : A.soil_scalar;
  if (soil < A.tol[t]) A.stress[t] = A.stress[t] + A.stress_inc;
}
