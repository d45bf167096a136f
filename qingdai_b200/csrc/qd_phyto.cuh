// qd_phyto.cuh -- transport of the phytoplankton tracers by the ocean currents, once per physics step
// (PhytoManager.advect_diffuse, pygcm/ecology/phyto.py:496-547; called at scripts/run_simulation.py:2256-2258;
// SURVEY 8f row 1): per species semi-Lagrangian gather (phyto.py:470-493, cos floor 0.5) blended with weight
// QD_PHYTO_ADV_ALPHA, explicit lateral diffusion dt*K_h*lap(C) (phyto.py:453-468), clip to >= 0, zero on land, then
// the polar ring means.  Ten tracers by default: twice the tracer work of the atmosphere, batched here over
// blockIdx.y = species so that all of them move in three launches.
#pragma once
#include "qd_ocean.cuh"

struct QdPhytoArgs {
  const double *uo, *vo;     // [nlat][nlon] currents
  double* C;                 // [S][nlat][nlon], in place
  double* tmp;               // [S][nlat][nlon] scratch
  const uint8_t* land;
  double dt, alpha, dtkh;    // dtkh = float(dt) * K_h (0 when K_h <= 0)
};
// phyto.py:515-518: C_new = (1 - a) C + a * map_coordinates(C, departure)
__global__ void __launch_bounds__(QD_THREADS) k_phyto_advect(QdGeo g, QdPhytoArgs A) {
  QD_CELL_PROLOGUE(g)                      // b = species
  if (!active) return;
  double y, x;
  qd_departure(A.uo[idx], A.vo[idx], A.dt, g, qd_row(g, QD_R_COS_ADV_HALF)[j], qd_row(g, QD_R_INV_ACOS_HALF)[j], j, i, &y, &x);
  const double* Cs = A.C + off;
  const double adv = qd_bilinear_wrap(Cs, g.nlat, g.nlon, y, x);
  A.tmp[off + idx] = (1.0 - A.alpha) * Cs[idx] + A.alpha * adv;
}
// phyto.py:520-529
__global__ void __launch_bounds__(QD_THREADS) k_phyto_finish(QdGeo g, QdPhytoArgs A) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  double v = A.tmp[off + idx];
  if (A.dtkh > 0.0) {
    QdCleanLoad F{A.tmp + off, g.nlon};
    const double lap = qd_lap_cell(F, j, i, g, qd_row(g, QD_R_COS_ADV_HALF));
    v = qd_nan_to_num(v) + A.dtkh * lap;
  }
  v = qd_clip(v, 0.0, INFINITY);
  if (A.land[idx] != 0) v = 0.0;
  A.C[off + idx] = v;
}
// phyto.py:531-546: each pole row's ocean cells take the row's ocean mean.  grid (2 poles, S species).
__global__ void __launch_bounds__(QD_THREADS) k_phyto_polar(QdGeo g, QdPhytoArgs A) {
  const int north = blockIdx.x, s = blockIdx.y;
  const int j = north ? g.nlat - 1 : 0, n = g.nlon;
  double* row = A.C + (size_t)s * g.ncell + (size_t)j * n;
  const uint8_t* land = A.land + (size_t)j * n;
  const double cnt = qd_block_sum_n<0>(n, [&](int k) { return land[k] == 0 ? 1.0 : 0.0; });
  if (cnt > 0.0) {
    const double sum = qd_block_sum_n<1>(n, [&](int k) { return land[k] == 0 ? row[k] : 0.0; });
    const double m = sum / cnt;
    __syncthreads();
    QD_BLOCK_FIRST_FOR(k, n) { if (land[k] == 0) row[k] = m; }
  }
}
