// qd_hyper4.cuh -- fused del^4 hyperdiffusion step: F' = nan_to_num(nan_to_num(F) - k4 * lap(lap(F)) * dt)
// (dynamics.py:175-212, ocean.py:119-152) as ONE shared-memory tiled kernel per field group.
//
// The two-kernel form (lap -> scratch, lap -> update) moves 5 x 8 B per cell and evaluates the Laplacian
// with ~10 global loads each; ncu (profiles/r01_ncu_full_ens64_v1.csv) showed it issue-bound at
// 12-16 % of HBM peak.  Here a block stages an (TJ+8) x (TI+4) halo tile of the cleaned field in shared
// memory, evaluates lap(F) on the (TJ+4) x (TI+2) ring it needs, then lap(lap F) and the update on its
// TJ x TI interior: HBM traffic is read F once + write F' once, every operand of the two stencils comes
// from shared memory.  Halo radius: 4 rows (np.gradient twice per Laplacian, one-sided at the poles:
// no wrap in latitude), 2 columns (periodic, period n_lon: np.roll).
// Out of place (a neighbouring block still needs the old halo): the orchestrator ping-pongs slots.
#pragma once
#include "qd_ocean.cuh"

#define QD_H4_TI 64

struct QdHyper4Args {
  int n;                                   // fields in this launch
  const double* src[QD_MAX_FIELDS];
  double* dst[QD_MAX_FIELDS];
  const double* k4rows[QD_MAX_FIELDS];     // per-row coefficient table
  double scale[QD_MAX_FIELDS];             // multiplier on the table (0.5 for eta)
  int raw_k4[QD_MAX_FIELDS];               // 1: table already holds k4 (atmosphere / overrides); 0: k4 = table / max(1e-12, sub_dt)
  const double* cosr;                      // cosine rows followed by 1/c and 1/c^2
  double dt;                               // step handed to the operator (atmosphere); ocean reads sub_dt from the scalar table
  int nsub;                                // inner sub-division (QD_K4_NSUB / QD_OCEAN_K4_NSUB): sub = dt / nsub
  int ocean;                               // 1: per-member sub_dt and early exit through the device sub-step counter
  QdSubCtl sc;
};

// Laplacian at (global row j, tile column c) from a tile accessor A(jglobal, tile_col).
template <class Acc>
QD_HD double qd_lap_rel(const Acc& A, int j, int c, const QdGeo& g, const double* cr) {
  const int nlat = g.nlat;
  const double* ic = cr + nlat;
  const double* ic2 = cr + 2 * nlat;
  auto G = [&](int jj) -> double {
    if (jj == 0) return (A(1, c) - A(0, c)) * g.inv_dlat;
    if (jj == nlat - 1) return (A(nlat - 1, c) - A(nlat - 2, c)) * g.inv_dlat;
    return (A(jj + 1, c) - A(jj - 1, c)) * g.inv_2dlat;
  };
  double gphi;
  if (j == 0) gphi = (cr[1] * G(1) - cr[0] * G(0)) * g.inv_dlat;
  else if (j == nlat - 1) gphi = (cr[j] * G(j) - cr[j - 1] * G(j - 1)) * g.inv_dlat;
  else gphi = (cr[j + 1] * G(j + 1) - cr[j - 1] * G(j - 1)) * g.inv_2dlat;
  const double term_phi = ic[j] * gphi;
  const double d2 = ((A(j, c + 1) - 2.0 * A(j, c)) + A(j, c - 1)) * g.inv_dlon_sq;
  return (term_phi + d2 * ic2[j]) * g.inv_a_sq;
}

template <int TJ>
__global__ void __launch_bounds__(QD_THREADS) k_hyper4_tile(QdGeo g, QdHyper4Args A) {
  constexpr int TI = QD_H4_TI;
  constexpr int RF = TJ + 8, CF = TI + 4, RL = TJ + 4, CL = TI + 2;
  __shared__ double Fs[RF * CF];
  __shared__ double Ls[RL * CL];
  const int b = blockIdx.y;
  if (A.ocean && qd_sub_done(g, b, A.sc)) return;
  const int tiles_i = (g.nlon + TI - 1) / TI;
  const int tj = blockIdx.x / tiles_i, ti = blockIdx.x - tj * tiles_i;
  const int j0 = tj * TJ, i0 = ti * TI;
  const int jF0 = j0 - 4, jL0 = j0 - 2, iF0 = i0 - 2;
  const int nlat = g.nlat, nlon = g.nlon;
  const size_t off = (size_t)b * g.ncell;
  double sub_dt = A.dt;
  if (A.ocean) sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const double inner = sub_dt / (double)(A.nsub > 1 ? A.nsub : 1);
  {
    const int k = blockIdx.z;                 // one field per block (gridDim.z = number of fields)
    const double* F = A.src[k] + off;
    QD_BLOCK_FIRST_FOR(e, RF * CF) {
      const int r = e / CF, cc = e - r * CF;
      const int gj = jF0 + r;
      if (gj >= 0 && gj < nlat) {
        int gi = (iF0 + cc) % nlon;
        if (gi < 0) gi += nlon;
        Fs[e] = qd_nan_to_num(F[(size_t)gj * nlon + gi]);
      }
    }
    __syncthreads();
    auto AF = [&](int jj, int c) -> double { return Fs[(jj - jF0) * CF + c]; };
    QD_BLOCK_FIRST_FOR(e, RL * CL) {
      const int r = e / CL, cc = e - r * CL;
      const int gj = jL0 + r;
      if (gj >= 0 && gj < nlat) Ls[e] = qd_nan_to_num(qd_lap_rel(AF, gj, cc + 1, g, A.cosr));
    }
    __syncthreads();
    auto AL = [&](int jj, int c) -> double { return Ls[(jj - jL0) * CL + c]; };
    const double* k4r = A.k4rows[k];
    for (int e = threadIdx.x; e < TJ * TI; e += blockDim.x) {
      const int r = e / TI, cc = e - r * TI;
      const int gj = j0 + r, gi = i0 + cc;
      if (gj < nlat && gi < nlon) {
        const double L2 = qd_lap_rel(AL, gj, cc + 1, g, A.cosr);
        double k4 = k4r[gj];
        if (!A.raw_k4[k]) k4 = k4 / fmax(1e-12, sub_dt);       // ocean.py:347
        k4 = A.scale[k] * k4;
        const double cur = Fs[(r + 4) * CF + cc + 2];
        A.dst[k][off + (size_t)gj * nlon + gi] = qd_nan_to_num(cur - k4 * L2 * inner);
      }
    }
  }
}
