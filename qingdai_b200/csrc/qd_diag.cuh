// qd_diag.cuh -- the periodic global diagnostics of the loop in ONE launch per call: area-weighted means
// (energy.compute_energy_diagnostics energy.py:494-538, hydrology.diagnose_water_closure hydrology.py:270-340,
// WindDrivenSlabOcean.diagnostics ocean.py:535-561, the script's periodic prints run_simulation.py:2273-2285,
// 2418-2424) and the extrema (eta min/max, max |U_ocean|, max |u|, T_s min/max).  Warp-shuffle block reductions,
// per-block partials combined in block order by the last block (deterministic).  The energy terms are re-evaluated
// from the state with the same device functions the step uses (qd_shortwave / qd_longwave / qd_sensible).
#pragma once
#include "qd_ocean.cuh"

// layout of the result, one row of QD_DIAG_COUNT doubles per member (sums are area weighted with max(cos lat, 0))
enum {
  QD_D_WSUM = 0,
  QD_D_TS, QD_D_H, QD_D_Q, QD_D_CLOUD, QD_D_HICE, QD_D_WLAND, QD_D_SSNOW, QD_D_EFLUX, QD_D_PRECIP, QD_D_RLAND, QD_D_ALBEDO, QD_D_SST,
  QD_D_I, QD_D_R, QD_D_OLR, QD_D_SW_SFC, QD_D_LW_SFC, QD_D_SH, QD_D_LH, QD_D_KE_OCEAN,
  QD_D_NSUM,
  QD_D_ETA_MIN = QD_D_NSUM, QD_D_ETA_MAX, QD_D_UOCEAN_MAX, QD_D_UABS_MAX, QD_D_TS_MIN, QD_D_TS_MAX,
  QD_DIAG_COUNT
};

struct QdDiagArgs {
  const double *ts, *h, *q, *cloud, *hice, *wland, *ssnow, *eflux, *precip, *rland, *albedo, *sst, *isr, *cloud_eff, *lh, *u, *v, *uo, *vo, *eta;
  const uint8_t* land;
  int has_cloud_eff;
  double* part;          // [B][QD_DIAG_COUNT][nvb]
  unsigned* ticket;
  double* out;           // [B][QD_DIAG_COUNT]
};

QD_D double qd_diag_combine(int q, double a, double b) {
  if (q < QD_D_NSUM) return a + b;
  if (q == QD_D_ETA_MIN || q == QD_D_TS_MIN) return b < a ? b : a;
  return b > a ? b : a;
}
QD_D double qd_diag_identity(int q) {
  if (q < QD_D_NSUM) return 0.0;
  if (q == QD_D_ETA_MIN || q == QD_D_TS_MIN) return DBL_MAX;
  return -DBL_MAX;
}

#ifndef QD_LB_DIAG
#define QD_LB_DIAG 1
#endif
__global__ void __launch_bounds__(QD_THREADS, QD_LB_DIAG) k_diag(QdGeo g, QdDiagArgs A) {
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  double* part = A.part + ((size_t)blockIdx.y * QD_DIAG_COUNT) * g.nvb;
  QD_VB_LOOP(g) {
  double acc[QD_DIAG_COUNT];
#pragma unroll
  for (int q = 0; q < QD_DIAG_COUNT; ++q) acc[q] = qd_diag_identity(q);
  QD_VB_CELLS(g, g.ncomp) {
    QD_CELL_JI(g)
    if (!qd_owned(g, j)) continue;
    const size_t c = off + idx;
    const double w = qd_row(g, QD_R_W)[j];
    const double ts = A.ts[c], h = A.h[c], hice = A.hice[c], u = A.u[c], v = A.v[c];
    const int land = A.land[c] == 1;
    const double ce = A.has_cloud_eff ? A.cloud_eff[c] : A.cloud[c];
    const QdSW sw = qd_shortwave(A.isr[c], A.albedo[c], ce, P[QD_P_SW_A0], P[QD_P_SW_KC]);
    const double Ta = 288.0 + (9.81 / 1004.0) * h;
    const QdLW lw = qd_longwave(ts, Ta, ce, land, qd_ice_frac(hice, P[QD_P_HICE_REF]), P);
    const double SH = qd_sensible(ts, Ta, u, v, P);
    const double uo = A.uo[c], vo = A.vo[c], eta = A.eta[c];
    acc[QD_D_WSUM] += w;
    acc[QD_D_TS] += ts * w; acc[QD_D_H] += h * w; acc[QD_D_Q] += A.q[c] * w; acc[QD_D_CLOUD] += A.cloud[c] * w;
    acc[QD_D_HICE] += hice * w; acc[QD_D_WLAND] += A.wland[c] * w; acc[QD_D_SSNOW] += A.ssnow[c] * w;
    acc[QD_D_EFLUX] += A.eflux[c] * w; acc[QD_D_PRECIP] += A.precip[c] * w; acc[QD_D_RLAND] += A.rland[c] * w;
    acc[QD_D_ALBEDO] += A.albedo[c] * w; acc[QD_D_SST] += A.sst[c] * w;
    acc[QD_D_I] += qd_max(0.0, A.isr[c]) * w; acc[QD_D_R] += sw.R * w; acc[QD_D_OLR] += lw.olr * w;
    acc[QD_D_SW_SFC] += sw.sfc * w; acc[QD_D_LW_SFC] += lw.sfc * w; acc[QD_D_SH] += SH * w; acc[QD_D_LH] += A.lh[c] * w;
    acc[QD_D_KE_OCEAN] += (0.5 * (uo * uo + vo * vo)) * w;
    const double so = sqrt(uo * uo + vo * vo), ua = fabs(u);
    if (eta < acc[QD_D_ETA_MIN]) acc[QD_D_ETA_MIN] = eta;
    if (eta > acc[QD_D_ETA_MAX]) acc[QD_D_ETA_MAX] = eta;
    if (so > acc[QD_D_UOCEAN_MAX]) acc[QD_D_UOCEAN_MAX] = so;
    if (ua > acc[QD_D_UABS_MAX]) acc[QD_D_UABS_MAX] = ua;
    if (ts < acc[QD_D_TS_MIN]) acc[QD_D_TS_MIN] = ts;
    if (ts > acc[QD_D_TS_MAX]) acc[QD_D_TS_MAX] = ts;
  }
  // block reduction of every quantity: warp shuffles, then one value per warp through shared memory
#if QD_EMU
  static thread_local double eacc[QD_DIAG_COUNT];
  static thread_local unsigned ecnt = 0;
  if (ecnt == 0) for (int q = 0; q < QD_DIAG_COUNT; ++q) eacc[q] = qd_diag_identity(q);
  for (int q = 0; q < QD_DIAG_COUNT; ++q) eacc[q] = qd_diag_combine(q, eacc[q], acc[q]);
  if (++ecnt == blockDim.x) { for (int q = 0; q < QD_DIAG_COUNT; ++q) part[(size_t)q * g.nvb + vb_] = eacc[q]; ecnt = 0; }
#else
  __shared__ double sm[QD_DIAG_COUNT][QD_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < QD_DIAG_COUNT; ++q) {
    double v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v = qd_diag_combine(q, v, __shfl_down_sync(0xffffffffu, v, o));
    if (lane == 0) sm[q][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < QD_DIAG_COUNT) {
    const int q = threadIdx.x;
    double v = sm[q][0];
    for (int k = 1; k < QD_THREADS / 32; ++k) v = qd_diag_combine(q, v, sm[q][k]);
    part[(size_t)q * g.nvb + vb_] = v;
  }
  __syncthreads();                         // sm[][] is reused by the next virtual block
#endif
  }
  if (qd_block_is_last(A.ticket + blockIdx.y, gridDim.x)) {
    QD_BLOCK_LAST_FOR(q, QD_DIAG_COUNT) {
      double v = qd_diag_identity(q);
      for (int k = 0; k < g.nvb; ++k) v = qd_diag_combine(q, v, QD_LDCG(part + (size_t)q * g.nvb + k));
      A.out[(size_t)blockIdx.y * QD_DIAG_COUNT + q] = v;
    }
  }
}
