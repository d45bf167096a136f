// qd_eco.cuh -- ecology SUB-DAILY path on the device (pygcm/ecology/adapter.py:140-186,
// pygcm/ecology/population.py:252-286 step_subdaily, :288-292 total_LAI, :831-841 canopy factor,
// :895-915 recompute policy and cache, :855-892 band albedo).  The daily ecology (LAI growth, spread,
// seeds, individuals) stays host Python and only hands the LAI layers over.
//
// Per physics step the reference evaluates two np.nanmean reductions over the summed LAI_layers_SK
// [S, K, lat, lon] to decide whether the canopy cache f = 1 - exp(-k * LAI_tot) must be rebuilt (time
// based every QD_ECO_LIGHT_UPDATE_EVERY_HOURS, or mean|dLAI| / mean(LAI_snapshot) >= 0.05).  Here the
// clock, the decision and the cache live on the device: k_eco_stats reduces and decides (last block),
// k_eco_canopy rebuilds where the member's flag is set; the per-cell albedo and E_day accumulation are
// fused into k_column (qd_phys.cuh).  Nothing returns to the host.
#pragma once
#include "qd_loop.cuh"

struct QdEcoArgs {
  const double* lai;        // [B][nl][ncell], caller-owned (LAI_layers_SK flattened over species x layers)
  int nl;
  double* snap;             // QD_F_LAI_SNAP
  double* fcanopy;          // QD_F_FCANOPY
  double* part[4];
  unsigned* ticket;
  double dt_hours, every_hours, delta_thr, k_canopy;
};

// np.sum(LAI_layers_SK, axis=(0, 1)): slices are accumulated in storage order
QD_D double qd_eco_total(const QdEcoArgs& A, const QdGeo& g, int b, int idx) {
  const double* p = A.lai + ((size_t)b * A.nl) * g.ncell + idx;
  double t = p[0];
  for (int l = 1; l < A.nl; ++l) t = t + p[(size_t)l * g.ncell];
  return t;
}

__global__ void __launch_bounds__(QD_THREADS) k_eco_stats(QdGeo g, QdEcoArgs A) {
  double t;
  const size_t pb = (size_t)blockIdx.y * g.nvb;
  QD_VB_LOOP(g) {
    double sd = 0.0, nd = 0.0, sb = 0.0, nb = 0.0;
    QD_VB_CELLS(g, g.ncomp) {
      QD_CELL_JI(g)
      if (!qd_owned(g, j)) continue;
      const double now = qd_eco_total(A, g, b, idx);
      const double s = A.snap[off + idx];
      const double d = fabs(now - s);
      if (d == d) { sd += d; nd += 1.0; }                     // nanmean skips NaN
      const double m = qd_max(s, 1e-6);
      if (m == m) { sb += m; nb += 1.0; }
    }
    if (qd_block_sum<0>(sd, &t)) A.part[0][pb + vb_] = t;
    if (qd_block_sum<1>(nd, &t)) A.part[1][pb + vb_] = t;
    if (qd_block_sum<2>(sb, &t)) A.part[2][pb + vb_] = t;
    if (qd_block_sum<3>(nb, &t)) A.part[3][pb + vb_] = t;
  }
  if (qd_block_is_last(A.ticket + b, gridDim.x)) {
    double Sd = 0.0, Nd = 0.0, Sb = 0.0, Nb = 0.0;
    const bool o0 = qd_final_sum<4>(A.part[0] + pb, g.nvb, &Sd);
#if !QD_EMU
    __shared__ double keep[4];
    if (o0) keep[0] = Sd;
    if (qd_final_sum<5>(A.part[1] + pb, g.nvb, &Nd)) keep[1] = Nd;
    if (qd_final_sum<6>(A.part[2] + pb, g.nvb, &Sb)) keep[2] = Sb;
    if (qd_final_sum<7>(A.part[3] + pb, g.nvb, &Nb)) keep[3] = Nb;
    __syncthreads();
    Sd = keep[0]; Nd = keep[1]; Sb = keep[2]; Nb = keep[3];
    const bool one = threadIdx.x == 0;
#else
    qd_final_sum<5>(A.part[1] + pb, g.nvb, &Nd);
    qd_final_sum<6>(A.part[2] + pb, g.nvb, &Sb);
    qd_final_sum<7>(A.part[3] + pb, g.nvb, &Nb);
    const bool one = o0;
#endif
    if (one) {
      double* S = g.scal + (size_t)b * QD_S_COUNT;
      const double hours = S[QD_S_ECO_HOURS] + A.dt_hours;                 // population.py:272
      S[QD_S_ECO_HOURS] = hours;
      bool rec = (S[QD_S_ECO_CACHED] == 0.0) || (hours >= S[QD_S_ECO_NEXT]);
      if (!rec) {                                                          // population.py:902-907
        const double delta = Sd / Nd, base = Sb / Nb;                      // 0/0 -> NaN like nanmean of all-NaN
        const double ratio = (base > 0.0) ? delta / base : delta;
        rec = ratio >= A.delta_thr;
      }
      S[QD_S_ECO_FLAG] = rec ? 1.0 : 0.0;
      if (rec) { S[QD_S_ECO_NEXT] = hours + A.every_hours; S[QD_S_ECO_CACHED] = 1.0; }
    }
  }
}

// f = 1 - exp(-k * max(LAI_tot, 0)); snapshot = LAI_tot            (population.py:911-915, :274-276)
__global__ void __launch_bounds__(QD_THREADS) k_eco_canopy(QdGeo g, QdEcoArgs A) {
  QD_CELL_PROLOGUE(g)
  if (!active || g.scal[(size_t)b * QD_S_COUNT + QD_S_ECO_FLAG] == 0.0) return;
  const double tot = qd_eco_total(A, g, b, idx);
  A.fcanopy[off + idx] = 1.0 - QD_EXP(-A.k_canopy * qd_max(tot, 0.0));
  A.snap[off + idx] = tot;
}

__global__ void __launch_bounds__(QD_THREADS) k_eco_snapshot(QdGeo g, QdEcoArgs A) {
  QD_CELL_PROLOGUE(g)
  if (active) A.snap[off + idx] = qd_eco_total(A, g, b, idx);
}

// Stand-alone form of the per-cell part (the drop-in EcologyAdapter.step_subdaily): E_day += nan_to_num(isr)*dt,
// alpha = clip(leaf*f + (1-f)*soil, 0, 1) on land, NaN on ocean                         (adapter.py:159-176)
struct QdEcoCellArgs { const double *isr, *fcanopy; double *eday, *alpha; const uint8_t* land; double dt; int want_alpha; };
__global__ void __launch_bounds__(QD_THREADS) k_eco_cell(QdGeo g, QdEcoCellArgs A) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  A.eday[c] = A.eday[c] + qd_nan_to_num(A.isr[c]) * A.dt;
  if (A.want_alpha) {
    double a = NAN;
    if (A.land[c] == 1) {
      const double fc = A.fcanopy[c];
      a = qd_clip(P[QD_P_ECO_ALPHA_LEAF] * fc + (1.0 - fc) * P[QD_P_ECO_SOIL_REFLECT], 0.0, 1.0);
    }
    A.alpha[c] = a;
  }
}

// A_b = clip(R_eff[b] * f + (1 - f) * soil, 0, 1) on land, NaN on ocean; out [B][nb][ncell]   (population.py:875-892)
struct QdEcoBandArgs { const double* fcanopy; const uint8_t* land; double* out; int nb; double soil; double r_eff[QD_ECO_MAX_BANDS]; };
__global__ void __launch_bounds__(QD_THREADS) k_eco_bands(QdGeo g, QdEcoBandArgs A) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const size_t c = off + idx;
  const bool land = A.land[c] == 1;
  const double f = A.fcanopy[c];
  for (int k = 0; k < A.nb; ++k) {
    double v = NAN;
    if (land) v = qd_clip(A.r_eff[k] * f + (1.0 - f) * A.soil, 0.0, 1.0);
    A.out[((size_t)b * A.nb + k) * g.ncell + idx] = v;
  }
}
