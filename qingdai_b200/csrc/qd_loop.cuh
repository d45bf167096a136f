// qd_loop.cuh -- per-step physics of the script loop around the two cores
// (scripts/run_simulation.py:1766-1934): hybrid precipitation diagnosis (pygcm/physics.py:253-354),
// cloud from precipitation / cloud source / blend / tracer advection (physics.py:48-114,
// run_simulation.py:1866-1934).
#pragma once
#include "qd_ocean.cuh"

QD_HD double qd_scal(const QdGeo& g, int b, int id) { return g.scal[(size_t)b * QD_S_COUNT + id]; }

// ---- phase A: convergence field pos = max(0, -(div - D_crit)), sum(Pq*w), raw orographic factor
struct QdPrecipAArgs {
  const double *u, *v, *pcond, *nx, *ny;
  double *pos, *orog_raw, *part;
  unsigned* ticket;
  const qd_forcing_t* forcing; const int* step_idx; double* hcos;    // first kernel of the step: also fills cos(hour angle) per column
};
__global__ void __launch_bounds__(QD_THREADS, QD_LB_PRECIPA) k_precip_a(QdGeo g, QdPrecipAArgs A) {
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  if (A.hcos && blockIdx.y == 0) {           // forcing.py:118-131, consumed by k_column later in the step
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.nlon; i += gridDim.x * blockDim.x) qd_forcing_col(g, A.forcing, A.step_idx, A.hcos, i);
  }
  double t;
  double* part = A.part + (size_t)blockIdx.y * g.nvb;
  QD_VB_LOOP(g) {
    double contrib = 0.0;
    QD_VB_CELLS(g, g.ncomp) {
      QD_CELL_JI(g)
      const size_t c = off + idx;
      const double div = qd_div_cell(A.u + off, A.v + off, j, i, g);
      A.pos[c] = qd_max(0.0, -(div - P[QD_P_D_CRIT]));
      if (qd_owned(g, j)) contrib += qd_max(0.0, A.pcond[c]) * qd_row(g, QD_R_W)[j];
      if (P[QD_P_OROG] != 0.0 && P[QD_P_HAS_ELEVATION] != 0.0) {           // physics.py:154-156
        const double up = qd_max(0.0, A.u[c] * A.nx[c] + A.v[c] * A.ny[c]);
        A.orog_raw[c] = qd_clip(1.0 + P[QD_P_K_OROG] * up, 1.0, 2.0);
      }
    }
    if (qd_block_sum<0>(contrib, &t)) part[vb_] = t;
  }
  if (qd_block_is_last(A.ticket + blockIdx.y, gridDim.x)) {
    if (qd_final_sum<1>(part, g.nvb, &t)) g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_SUM_PQW] = t;
  }
}

// ---- phase B: P_raw = Pq * (1 + beta*clip(pos/scale,0,5)) * F_orog, sum(P_raw*w)   (physics.py:296-322)
struct QdPrecipBArgs {
  const double *pos, *pcond, *orog;
  double *praw, *part;
  unsigned* ticket;
};
#ifndef QD_LB_PRECIPB
#define QD_LB_PRECIPB 7
#endif
__global__ void __launch_bounds__(QD_THREADS, QD_LB_PRECIPB) k_precip_b(QdGeo g, QdPrecipBArgs A) {
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  // the median scale is the same for every cell of a member (one reciprocal per thread, amortised over its cells);
  // pos is zero wherever the flow diverges
  const QdRcp scale = qd_rcp(fmax(qd_scal(g, (int)blockIdx.y, QD_S_MED_POS), 1e-12));
  double t;
  double* part = A.part + (size_t)blockIdx.y * g.nvb;
  QD_VB_LOOP(g) {
    double contrib = 0.0;
    QD_VB_CELLS(g, g.ncomp) {
      QD_CELL_JI(g)
      const size_t c = off + idx;
      double F_div = 0.0;
      if (qd_scal(g, b, QD_S_CNT_POS) > 0.0) F_div = qd_clip(qd_div_u(A.pos[c], scale), 0.0, 5.0);
      double F_or = 1.0;
      if (P[QD_P_OROG] != 0.0 && P[QD_P_HAS_ELEVATION] != 0.0) F_or = qd_clip(A.orog[c], 1.0, 3.0);
      const double F = (1.0 + P[QD_P_BETA_DIV] * F_div) * F_or;
      const double praw = qd_max(0.0, A.pcond[c]) * F;
      A.praw[c] = praw;
      if (qd_owned(g, j)) contrib += praw * qd_row(g, QD_R_W)[j];
    }
    if (qd_block_sum<0>(contrib, &t)) part[vb_] = t;
  }
  if (qd_block_is_last(A.ticket + blockIdx.y, gridDim.x)) {
    if (qd_final_sum<1>(part, g.nvb, &t)) g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_SUM_PRAWW] = t;
  }
}

QD_HD double qd_precip_renorm(const QdGeo& g, int b) {        // physics.py:321-323
  const double num = qd_scal(g, b, QD_S_SUM_PQW);
  const double den = qd_scal(g, b, QD_S_SUM_PRAWW) + 1e-20;
  return den > 0 ? num / den : 1.0;
}
QD_HD bool qd_precip_fallback(const QdGeo& g, int b) {        // physics.py:344-348
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  if (P[QD_P_P_FALLBACK] == 0.0) return false;
  return (qd_scal(g, b, QD_S_SUM_PQW) / P[QD_P_WSUM_ALL]) < P[QD_P_PQ_MIN];
}

// ---- phase C: Gaussian (sigma=1, reflect) along latitude of P_raw*s and, when the weak-humidity
//      fallback is active, of the legacy k_precip*pos field (physics.py:12-46 with cloud gating off)
struct QdPrecipCArgs { const double *praw, *pos; double *g0, *g1; };
__global__ void __launch_bounds__(QD_THREADS) k_precip_c(QdGeo g, QdPrecipCArgs A, QdGaussW w) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double s = qd_precip_renorm(g, b);
  const int nlon = g.nlon;
  const double* p0 = A.praw + off + i;
  auto E0 = [&](int jj) -> double { return p0[(size_t)jj * nlon] * s; };
  A.g0[off + idx] = qd_gauss_tap(E0, j, g.nlat, w);
  if (qd_precip_fallback(g, b)) {
    const double* p1 = A.pos + off + i;
    const double kp = P[QD_P_K_PRECIP];
    auto E1 = [&](int jj) -> double { return kp * p1[(size_t)jj * nlon]; };
    A.g1[off + idx] = qd_gauss_tap(E1, j, g.nlat, w);
  }
}
// ---- phase D: Gaussian along longitude, blend, clip -> precip
struct QdPrecipDArgs { const double *g0, *g1; double* precip; };
__global__ void __launch_bounds__(QD_THREADS) k_precip_d(QdGeo g, QdPrecipDArgs A, QdGaussW w) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double* r0 = A.g0 + off + (size_t)j * g.nlon;
  auto E0 = [&](int ii) -> double { return r0[ii]; };
  double Pv = qd_gauss_tap(E0, i, g.nlon, w);
  if (qd_precip_fallback(g, b)) {
    const double* r1 = A.g1 + off + (size_t)j * g.nlon;
    auto E1 = [&](int ii) -> double { return r1[ii]; };
    const double Pdyn = qd_gauss_tap(E1, i, g.nlon, w);
    Pv = (1.0 - P[QD_P_P_BLEND]) * Pv + P[QD_P_P_BLEND] * Pdyn;
  }
  A.precip[off + idx] = (Pv != Pv) ? Pv : (Pv < 0.0 ? 0.0 : Pv);     // np.clip(P, 0, None)
}

// ---- cloud phase A: C_raw = C_max*tanh(precip/(P_ref+1e-12)) and the raw cloud source
//      (physics.py:48-70, :72-108; P_ref = median(precip>0) run_simulation.py:1867-1875)
struct QdCloudAArgs { const double *precip, *ts, *u, *v; double *craw, *sraw; };
__global__ void __launch_bounds__(QD_THREADS, QD_LB_CLOUDA) k_cloud_a(QdGeo g, QdCloudAArgs A) {
  QD_CELL_PROLOGUE(g)
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const QdRcp* D = g.udiv + (size_t)b * QD_U_COUNT;
  if (!active) return;
  const size_t c = off + idx;
  const int nlon = g.nlon, nlat = g.nlat;
  double P_ref = 1e-6;
  if (qd_scal(g, b, QD_S_CNT_PRECIP) > 0.0) {
    const double ov = P[QD_P_PREF];
    P_ref = (ov == ov) ? ov : qd_scal(g, b, QD_S_PREF);
  }
  A.craw[c] = P[QD_P_CMAX] * QD_TANH(qd_div_z(A.precip[c], P_ref + 1e-12));      // dry cells have precip == 0
  // source
  const double* T = A.ts + off;
  const double Ts = T[idx], u = A.u[c], v = A.v[c];
  double src = 0.5 * qd_clip(QD_TANH(qd_div_u(Ts - 285.0, D[QD_U_12K])), 0.0, 1.0);
  const double vort = qd_vort_cell(A.u + off, A.v + off, j, i, g);
  const double rel = qd_div_z(vort, qd_row(g, QD_R_FCOR)[j] + 1e-12);
  src = src + 0.4 * qd_clip(QD_TANH((rel - 0.5) / 2.0), 0.0, 1.0);
  const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
  const int jp = j + 1 < nlat ? j + 1 : 0, jm = j > 0 ? j - 1 : nlat - 1;
  const double dx = g.dlon * g.a * qd_row(g, QD_R_COS_ADV_ATM)[j];
  const double gx = qd_div_z(T[(size_t)j * nlon + ip] - T[(size_t)j * nlon + im], 2 * dx);
  const double gy = qd_div_u(T[(size_t)jp * nlon + i] - T[(size_t)jm * nlon + i], D[QD_U_2DY]);
  const double adv = -(u * gx + v * gy);
  src = src + 0.3 * qd_clip(QD_TANH(qd_div_u(fabs(adv), D[QD_U_ADV_REF])), 0.0, 1.0);
  A.sraw[c] = src;
}
// ---- cloud phase B: Gaussian along longitude of both fields, then the blend
//      (run_simulation.py:1890-1913)
struct QdCloudBArgs { const double *g0, *g1; double* cloud; double dt; /* dt / (6 * 3600), formed on the host */ };
__global__ void __launch_bounds__(QD_THREADS) k_cloud_b(QdGeo g, QdCloudBArgs A, QdGaussW w) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double* r0 = A.g0 + off + (size_t)j * g.nlon;
  const double* r1 = A.g1 + off + (size_t)j * g.nlon;
  auto E0 = [&](int ii) -> double { return r0[ii]; };
  auto E1 = [&](int ii) -> double { return r1[ii]; };
  const double C_P = qd_clip(qd_gauss_tap(E0, i, g.nlon, w), 0.0, 1.0);
  const double src = qd_clip(qd_gauss_tap(E1, i, g.nlon, w), 0.0, 1.0);
  const double tend = src * A.dt;
  double cl = A.cloud[c];
  cl = P[QD_P_W_MEM] * cl + P[QD_P_W_P] * C_P + P[QD_P_W_SRC] * qd_clip(cl + tend, 0.0, 1.0);
  if (P[QD_P_CLOUD_FLOOR] > 0.0) cl = qd_max(cl, qd_clip(P[QD_P_CLOUD_FLOOR] * C_P, 0.0, 1.0));
  A.cloud[c] = qd_clip(cl, 0.0, 1.0);
}
// ---- cloud phase C: Gaussian(sigma=0.2, wrap) along longitude of the advected tracer + blend
//      (run_simulation.py:1916-1934); with sigma<=0 the lat pass is skipped and w.r = 0
struct QdCloudCArgs { const double* g0; double* cloud; };
__global__ void __launch_bounds__(QD_THREADS) k_cloud_c(QdGeo g, QdCloudCArgs A, QdGaussW w) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double* r0 = A.g0 + off + (size_t)j * g.nlon;
  auto E0 = [&](int ii) -> double { return r0[ii]; };
  const double adv = (w.r > 0) ? qd_gauss_tap(E0, i, g.nlon, w) : r0[i];
  const double al = P[QD_P_CLOUD_ADV_ALPHA];
  A.cloud[c] = qd_clip((1.0 - al) * A.cloud[c] + al * adv, 0.0, 1.0);
}

// ---- routing accumulate (routing.py:232-236): buffer_kg += where(land, R*area*dt, 0)
__global__ void __launch_bounds__(QD_THREADS) k_route_accumulate(QdGeo g, const double* rland, const uint8_t* land,
                                                                double* buffer, double dt) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const size_t c = off + idx;
  const double incr = (land[idx] == 1) ? rland[c] * qd_row(g, QD_R_AREA)[j] * dt : 0.0;   // network mask, shared by members
  buffer[c] = buffer[c] + incr;
}

__global__ void k_set_forcing(qd_forcing_t* dst, qd_forcing_t f, int* step_idx) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { dst[0] = f; *step_idx = 0; }
}
__global__ void k_step_advance(int* step_idx) { if (threadIdx.x == 0 && blockIdx.x == 0) *step_idx = *step_idx + 1; }
