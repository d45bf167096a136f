// qd_ocean.cuh -- WindDrivenSlabOcean.step (pygcm/ocean.py:265-533) as phase kernels.
//
// Slots: the persistent state lives in (UO, VO, SST) = "A"; each CFL sub-step writes momentum to
// "B" scratch (UB, VB), runs del^4 in place on B, and the closing kernel of the sub-step reads B
// with neighbours (mean4 outlier fix) and writes A again -- the loop closes for any sub-step
// count without pointer swaps.  The sub-step count is data dependent (ocean.py:293-303): it is
// computed on the device per ensemble member and every sub-step kernel exits early for members
// that are already done.
#pragma once
#include "qd_phys.cuh"

// The current sub-step index lives in device memory (reset by k_ocean_nsub, advanced by
// k_ocean_sub_advance) so that the sub-step body can be the body of a CUDA-graph WHILE node.
struct QdSubCtl { const int* ctr; };

QD_HD bool qd_sub_done(const QdGeo& g, int b, const QdSubCtl& sc) {
  return *sc.ctr >= (int)g.scal[(size_t)b * QD_S_COUNT + QD_S_NSUB];
}

// Wind stress (ocean.py:285-290) + the two maxima behind n_sub (ocean.py:298-299).
QD_D void qd_ocean_nsub_member(const QdGeo& g, int b, double dt);
struct QdOcPrepArgs {
  const double *u, *v, *uo, *vo;
  double *taux, *tauy, *part_u, *part_va;
  unsigned* ticket;
  double dt; int* sub_ctr;   // sub_ctr != nullptr: the last block also derives n_sub (no cross-rank maxima pending)
};
__global__ void __launch_bounds__(QD_THREADS) k_ocean_prep(QdGeo g, QdOcPrepArgs A) {
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  double mu = 0.0, mva = 0.0;
  QD_CELL_LOOP(g) {
    QD_CELL_JI(g)
    const size_t c = off + idx;
    const double uo = A.uo[c], vo = A.vo[c];
    const double ur = A.u[c] - uo, vr = A.v[c] - vo;
    const double Va = sqrt(ur * ur + vr * vr);
    const double Ve = qd_min(Va, P[QD_P_OC_VCAP]);
    A.taux[c] = P[QD_P_OC_TAU_SCALE] * (P[QD_P_OC_RHO_A] * P[QD_P_OC_CD] * Ve * ur);
    A.tauy[c] = P[QD_P_OC_TAU_SCALE] * (P[QD_P_OC_RHO_A] * P[QD_P_OC_CD] * Ve * vr);
    const double sp = sqrt(uo * uo + vo * vo);
    if (sp > mu) mu = sp;                       // NaN-ignoring running maxima, like qd_block_max
    if (Va > mva) mva = Va;
  }
  double t;
  double* pu = A.part_u + (size_t)b * gridDim.x;
  double* pv = A.part_va + (size_t)b * gridDim.x;
  if (qd_block_max<0>(mu, &t)) pu[blockIdx.x] = t;
  if (qd_block_max<1>(mva, &t)) pv[blockIdx.x] = t;
  if (qd_block_is_last(A.ticket + b, gridDim.x)) {
    double m1 = 0.0, m2 = 0.0;
    const bool o1 = qd_final_max<2>(pu, gridDim.x, &m1);
    const bool o2 = qd_final_max<3>(pv, gridDim.x, &m2);
    // the owner thread of both reductions is the same thread (thread 0 / the serial host thread)
    if (o1 && o2) {
      double* S = g.scal + (size_t)b * QD_S_COUNT;
      S[QD_S_MAX_UOCEAN] = m1;
      S[QD_S_MAX_VA] = m2;
      if (A.sub_ctr) { if (b == 0) *A.sub_ctr = 0; qd_ocean_nsub_member(g, b, A.dt); }
    }
  }
}

// n_sub = clip(ceil(max(c, uadv) * (dt / max(1e-12, dx_min)) / max(1e-3, cfl)), 1, 500)  (ocean.py:297-303)
QD_D void qd_ocean_nsub_member(const QdGeo& g, int b, double dt) {
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  double* S = g.scal + (size_t)b * QD_S_COUNT;
  const double c = sqrt(P[QD_P_OC_G] * P[QD_P_OC_H]);
  const double uadv = fmax(S[QD_S_MAX_UOCEAN], S[QD_S_MAX_VA]);
  const double target = fmax(1e-3, P[QD_P_OC_CFL]);
  double n = ceil(fmax(c, uadv) * (dt / fmax(1e-12, P[QD_P_OC_DX_MIN])) / target);
  if (!(n >= 1.0)) n = 1.0;
  if (n > 500.0) n = 500.0;
  S[QD_S_NSUB] = n;
  S[QD_S_SUB_DT] = dt / n;
}
__global__ void k_ocean_nsub(QdGeo g, double dt, int* sub_ctr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *sub_ctr = 0;
  if (b >= g.batch) return;
  qd_ocean_nsub_member(g, b, dt);
}

// ice mask from the ice thickness (run_simulation.py:2201) on the rows of the launch geometry: latitude bands rebuild it
// on their halo rows after exchanging h_ice (masks themselves are never exchanged)
__global__ void __launch_bounds__(QD_THREADS) k_ice_mask(QdGeo g, const double* hice, uint8_t* ice) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  ice[off + idx] = (uint8_t)(hice[off + idx] > 0.0);
}

// Momentum + land zeroing + polar sponge (ocean.py:306-336): A -> B.
struct QdOcMomArgs {
  const double *eta, *uo, *vo, *taux, *tauy;
  double *ub, *vb;
  const uint8_t* land;
  const double *uo_alt, *vo_alt;      // fused sub-step path (qd_ocean_fused.cuh): the currents ping-pong, parity decided on the device
};
QD_D int qd_oc_src(const QdGeo& g, int b, const QdSubCtl& sc) {      // 0: this sub-step reads the home currents, 1: the alternate pair
  const int n = (int)g.scal[(size_t)b * QD_S_COUNT + QD_S_NSUB], s = *sc.ctr;
  if ((n & 1) && s == 0) return 0;
  return ((n - s) & 1) ? 1 : 0;
}
__global__ void __launch_bounds__(QD_THREADS) k_ocean_momentum(QdGeo g, QdOcMomArgs A, QdSubCtl sc) {
  QD_CELL_PROLOGUE(g)
  if (!active || qd_sub_done(g, b, sc)) return;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const int nlon = g.nlon, nlat = g.nlat;
  const double* eta = A.eta + off;
  const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
  const int jp = j + 1 < nlat ? j + 1 : 0, jm = j > 0 ? j - 1 : nlat - 1;     // np.roll wraps pole to pole
  // metric divisions as reciprocal multiplies (<= 1 ulp each): /(2 dlon), /(2 dphi), /(a cos), /a, /(rho_w H)
  const double de_dl = (eta[(size_t)j * nlon + ip] - eta[(size_t)j * nlon + im]) * g.inv_2dlon;
  const double de_dp = (eta[(size_t)jp * nlon + i] - eta[(size_t)jm * nlon + i]) * g.inv_2dlat;
  const double gx = de_dl * qd_row(g, QD_R_INV_ACOS_HALF)[j];
  const double gy = de_dp * g.inv_a;
  const double f = qd_row(g, QD_R_FCOR)[j];
  const bool alt = A.uo_alt && qd_oc_src(g, b, sc) == 1;
  double uo = alt ? A.uo_alt[c] : A.uo[c], vo = alt ? A.vo_alt[c] : A.vo[c];
  const double irH = P[QD_P_OC_INV_RHO_H];
  const double du = (f * vo - P[QD_P_OC_G] * gx + A.taux[c] * irH - P[QD_P_OC_R_BOT] * uo);
  const double dv = (-f * uo - P[QD_P_OC_G] * gy + A.tauy[c] * irH - P[QD_P_OC_R_BOT] * vo);
  uo = uo + sub_dt * du;
  vo = vo + sub_dt * dv;
  if (A.land[c] == 1) { uo = 0.0; vo = 0.0; }
  const double rex = qd_mrow(g, QD_R_OC_SPONGE, b)[j];
  uo = uo - sub_dt * rex * uo;
  vo = vo - sub_dt * rex * vo;
  A.ub[c] = uo;
  A.vb[c] = vo;
}

// del^4 on (UB, VB, ETA) with k4 = sigma4*dx^4/max(1e-12, sub_dt) (ocean.py:341-356): two passes.
__global__ void __launch_bounds__(QD_THREADS) k_ocean_lap(QdGeo g, QdFields f, QdSubCtl sc) {
  QD_CELL_PROLOGUE(g)
  if (!active || qd_sub_done(g, b, sc)) return;
  const double* cosr = qd_row(g, QD_R_COS_ADV_HALF);
  for (int k = 0; k < f.n; ++k) {
    QdCleanLoad F{f.src[k] + off, g.nlon};
    f.dst[k][off + idx] = qd_lap_cell(F, j, i, g, cosr);
  }
}
// aux[k] = row table (sigma4*dx^4, or a constant row for QD_OCEAN_K4_* overrides with over[k]=1)
__global__ void __launch_bounds__(QD_THREADS) k_ocean_hyper(QdGeo g, QdFields f, QdSubCtl sc, int k4_nsub, int over0, int over1, int over2) {
  QD_CELL_PROLOGUE(g)
  if (!active || qd_sub_done(g, b, sc)) return;
  const double* cosr = qd_row(g, QD_R_COS_ADV_HALF);
  const double sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const double inner = sub_dt / (double)(k4_nsub > 1 ? k4_nsub : 1);
  const int over[3] = {over0, over1, over2};
  for (int k = 0; k < f.n; ++k) {
    QdCleanLoad L{f.src[k] + off, g.nlon};
    const double L2 = qd_lap_cell(L, j, i, g, cosr);
    double k4 = f.aux[k][j];
    if (!over[k]) { k4 = k4 / fmax(1e-12, sub_dt); k4 = f.scale[k] * k4; }
    const double cur = qd_nan_to_num(f.dst[k][off + idx]);
    f.dst[k][off + idx] = qd_nan_to_num(cur - k4 * L2 * inner);
  }
}

// Continuity (ocean.py:364-367) + the area-weighted ocean sum of eta for the mean removal (:369-375), fused with the
// SST semi-Lagrangian blend (ocean.py:380-382): both read the post-del^4 currents, so one pass loads them once.
// The mean removal and hygiene of eta (ocean.py:375,436-443) need the global sum and are finished by
// k_ocean_sst_finish.  Grid-stride, one resident wave (ends in a grid-wide reduction).
struct QdOcContArgs {
  const double *ub, *vb, *eta_in;      // eta_in: where del^4 left eta (may be a scratch slot)
  const double* sst;
  double *eta, *tb, *part;
  const uint8_t* land;
  unsigned* ticket;
  QdBandCtl band;                      // world > 1: the last block all-reduces the eta sum with the other ranks (publish + pull)
};
__global__ void __launch_bounds__(QD_THREADS) k_ocean_continuity(QdGeo g, QdOcContArgs A, QdSubCtl sc) {
  const bool done = qd_sub_done(g, blockIdx.y, sc);
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  const double sub_dt = g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_SUB_DT];
  const double al = P[QD_P_OC_ADV_ALPHA];
  double t;
  double* part = A.part + (size_t)blockIdx.y * g.nvb;
  QD_VB_LOOP(g) {
    double contrib = 0.0;
    QD_VB_CELLS(g, done ? 0 : g.ncomp) {
      QD_CELL_JI(g)
      const size_t c = off + idx;
      const double div = qd_div_cell(A.ub + off, A.vb + off, j, i, g);
      double e = A.eta_in[c] + (-sub_dt * P[QD_P_OC_H] * div);
      const bool land = A.land[c] == 1;
      if (land) e = 0.0;
      A.eta[c] = e;
      if (qd_owned(g, j)) contrib += e * (qd_row(g, QD_R_W)[j] * (land ? 0.0 : 1.0));
      double y, x;
      qd_departure(A.ub[c], A.vb[c], sub_dt, g, qd_row(g, QD_R_COS_ADV_HALF)[j], qd_row(g, QD_R_INV_ACOS_HALF)[j], j, i, &y, &x);
      const double adv = qd_bilinear_wrap(A.sst + off, g.nlat, g.nlon, y, x);
      A.tb[c] = (1.0 - al) * A.sst[c] + al * adv;
    }
    if (qd_block_sum<0>(contrib, &t)) part[vb_] = t;
  }
  if (qd_block_is_last(A.ticket + blockIdx.y, gridDim.x)) {
    if (qd_final_sum<1>(part, g.nvb, &t)) {
      if (A.band.world > 1) {
        // latitude bands: the all-reduce of the eta sum rides in this kernel's tail -- publish the rank's partial, pull the
        // world's partials (rank order: identical bits on every rank); no separate all-reduce launch per CFL sub-step.
        // Measured alternative (every block of the consumer pulling): +32 us per sub-step, dropped.
        qd_band_publish(A.band, done ? 0.0 : t);
        t = qd_band_pull_sum(A.band);
      }
      if (!done) g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_ETA_NUM] = t;
    }
  }
}

#if !QD_EMU
// The same pass with TWO adjacent cells per thread (n_lon even): one row / column index, 16-byte loads of the pair's own
// values and of its north / south neighbours, shared row constants (the one-cell kernel spends 44 % of its issue slots on
// index arithmetic: profiles/README.md).  Virtual block vb owns PAIRS vb*256 + t + k*nvb*256; a thread adds its cells in
// that order, so the eta sum is still a function of the grid and the device only (not of the batch size).
__global__ void __launch_bounds__(QD_THREADS, QD_LB_CONT2) k_ocean_continuity2(QdGeo g, QdOcContArgs A, QdSubCtl sc) {
  const bool done = qd_sub_done(g, blockIdx.y, sc);
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  const double sub_dt = g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_SUB_DT];
  const double al = P[QD_P_OC_ADV_ALPHA], mH = -sub_dt * P[QD_P_OC_H];
  const int nlon = g.nlon, nlat = g.nlat;
  const double* cosr = qd_row(g, QD_R_COS);
  const double* iacc = qd_row(g, QD_R_INV_ACOS_CAP);
  const double* wrow = qd_row(g, QD_R_W);
  const double* cosh = qd_row(g, QD_R_COS_ADV_HALF);
  const double* iach = qd_row(g, QD_R_INV_ACOS_HALF);
  double t;
  double* part = A.part + (size_t)blockIdx.y * g.nvb;
  QD_VB_LOOP(g) {
    double contrib = 0.0;
    QD_VB_CELLS(g, done ? 0 : g.ncomp / 2) {
      const int t2 = 2 * t_;
      const int r_ = qd_div_nlon(g, t2);
      const int i = t2 - r_ * nlon;
      const int j = qd_seg_row(g, r_);
      const int idx = j * nlon + i;
      const double* ub = A.ub + off;
      const double* vb = A.vb + off;
      const double2 u2 = *reinterpret_cast<const double2*>(ub + idx);
      const double2 v2 = *reinterpret_cast<const double2*>(vb + idx);
      const double uw = ub[i > 0 ? idx - 1 : idx + nlon - 1];
      const double ue = ub[i + 2 < nlon ? idx + 2 : idx + 2 - nlon];
      const double du0 = (u2.y - uw) * g.inv_2dlon, du1 = (ue - u2.x) * g.inv_2dlon;
      double dv0 = 0.0, dv1 = 0.0;
      if (j > 0 && j < nlat - 1) {
        const double2 vn = *reinterpret_cast<const double2*>(vb + idx + nlon);
        const double2 vs = *reinterpret_cast<const double2*>(vb + idx - nlon);
        const double cp = cosr[j + 1], cm = cosr[j - 1];
        dv0 = (vn.x * cp - vs.x * cm) * g.inv_2dlat;
        dv1 = (vn.y * cp - vs.y * cm) * g.inv_2dlat;
      }
      const double ia = iacc[j];
      const double div0 = ia * (du0 + dv0), div1 = ia * (du1 + dv1);
      const double2 ein = *reinterpret_cast<const double2*>(A.eta_in + off + idx);
      const uchar2 l2 = *reinterpret_cast<const uchar2*>(A.land + off + idx);
      double e0 = ein.x + (mH * div0), e1 = ein.y + (mH * div1);
      const bool ld0 = l2.x == 1, ld1 = l2.y == 1;
      if (ld0) e0 = 0.0;
      if (ld1) e1 = 0.0;
      *reinterpret_cast<double2*>(A.eta + off + idx) = make_double2(e0, e1);
      if (qd_owned(g, j)) {
        const double w = wrow[j];
        contrib += e0 * (w * (ld0 ? 0.0 : 1.0));
        contrib += e1 * (w * (ld1 ? 0.0 : 1.0));
      }
      const double cj = cosh[j], ij = iach[j];
      double y0, x0, y1, x1;
      qd_departure(u2.x, v2.x, sub_dt, g, cj, ij, j, i, &y0, &x0);
      qd_departure(u2.y, v2.y, sub_dt, g, cj, ij, j, i + 1, &y1, &x1);
      const double adv0 = qd_bilinear_wrap(A.sst + off, nlat, nlon, y0, x0);
      const double adv1 = qd_bilinear_wrap(A.sst + off, nlat, nlon, y1, x1);
      const double2 s2 = *reinterpret_cast<const double2*>(A.sst + off + idx);
      *reinterpret_cast<double2*>(A.tb + off + idx) = make_double2((1.0 - al) * s2.x + al * adv0, (1.0 - al) * s2.y + al * adv1);
    }
    if (qd_block_sum<0>(contrib, &t)) part[vb_] = t;
  }
  if (qd_block_is_last(A.ticket + blockIdx.y, gridDim.x)) {
    if (qd_final_sum<1>(part, g.nvb, &t)) {
      if (A.band.world > 1) {
        qd_band_publish(A.band, done ? 0.0 : t);
        t = qd_band_pull_sum(A.band);
      }
      if (!done) g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_ETA_NUM] = t;
    }
  }
}
#endif

// SST diffusion + Q_net heating (ocean.py:384-406), outlier handling of currents (ocean.py:408-434): B -> A.
// On the member's last sub-step the non-polar rows also get the final Ts clip (ocean.py:531-533) and,
// in loop mode, the SST injection into the atmosphere's T_s (run_simulation.py:2252-2253); the two
// polar rows are finished by k_ocean_polar.
struct QdOcSstBArgs {
  const double *tb, *ub, *vb, *qnet;
  double *sst, *uo, *vo, *ts_atm, *eta;
  const uint8_t *land, *ice;
  int has_q, has_ice, inject;
  // fuse_mom: the closing pass of sub-step s also does the momentum step of sub-step s+1 (ocean.py:306-336) for every
  // member that has one: eta comes from eta_mid (continuity's output; the finished eta goes to `eta`, a different slot,
  // because the momentum stencil reads the neighbours' eta_mid), the post-momentum currents go to ub_next / vb_next (the
  // del^4 input slots, dead by now); the home currents uo / vo are only written on a member's last sub-step.
  int fuse_mom;
  const double *eta_mid, *taux, *tauy;
  double *ub_next, *vb_next;
};
// momentum step of ONE cell from its finished currents and the finished eta of its four neighbours (k_ocean_momentum,
// operand for operand)
QD_HD void qd_oc_momentum_cell(const QdGeo& g, const double* P, double sub_dt, double f, double iach, double rex, bool land,
                               double e_e, double e_w, double e_n, double e_s, double taux, double tauy, double* uo_io, double* vo_io) {
  double uo = *uo_io, vo = *vo_io;
  const double de_dl = (e_e - e_w) * g.inv_2dlon;
  const double de_dp = (e_n - e_s) * g.inv_2dlat;
  const double gx = de_dl * iach;
  const double gy = de_dp * g.inv_a;
  const double irH = P[QD_P_OC_INV_RHO_H];
  const double du = (f * vo - P[QD_P_OC_G] * gx + taux * irH - P[QD_P_OC_R_BOT] * uo);
  const double dv = (-f * uo - P[QD_P_OC_G] * gy + tauy * irH - P[QD_P_OC_R_BOT] * vo);
  uo = uo + sub_dt * du;
  vo = vo + sub_dt * dv;
  if (land) { uo = 0.0; vo = 0.0; }
  uo = uo - sub_dt * rex * uo;
  vo = vo - sub_dt * rex * vo;
  *uo_io = uo; *vo_io = vo;
}
// finished eta of a cell from continuity's output (ocean.py:375,436-443): what the closing pass stores for that cell
QD_HD double qd_oc_eta_done(double e_mid, bool any_ocean, double eta_mean, double eta_cap) {
  double e = e_mid;
  if (any_ocean) e = e - eta_mean;
  return qd_clip(qd_nan_to_num(e), -eta_cap, eta_cap);
}
// One cell of the closing pass, general form (any row: one-sided Laplacian next to the poles, every value cleaned).
QD_D void qd_sstf_cell_general(const QdGeo& g, const QdOcSstBArgs& A, const QdSubCtl& sc, int b, int j, int i, double eta_mean) {
  const size_t off = (size_t)b * g.ncell;
  const int idx = j * g.nlon + i;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double* S = g.scal + (size_t)b * QD_S_COUNT;
  const double sub_dt = S[QD_S_SUB_DT];
  const int nlon = g.nlon, nlat = g.nlat;
  {   // eta: mean removal over the ocean + hygiene (ocean.py:375,436-443); the sum comes from k_ocean_continuity
    double e = A.fuse_mom ? A.eta_mid[c] : A.eta[c];
    if (P[QD_P_OC_ANY_OCEAN] != 0.0) e = e - eta_mean;
    A.eta[c] = qd_clip(qd_nan_to_num(e), -P[QD_P_OC_ETA_CAP], P[QD_P_OC_ETA_CAP]);
  }
  double T = A.tb[c];
  if (P[QD_P_OC_K_H] > 0.0) {
    QdCleanLoad F{A.tb + off, nlon};
    const double lap = qd_lap_cell(F, j, i, g, qd_row(g, QD_R_COS_ADV_HALF));
    T = qd_nan_to_num(T) + sub_dt * P[QD_P_OC_K_H] * lap;
  }
  const bool ocean = A.land[c] != 1;
  const bool ice = A.has_ice ? (A.ice[c] != 0) : false;
  if (P[QD_P_OC_USE_QNET] != 0.0 && A.has_q) {
    const double tend = A.qnet[c] * P[QD_P_OC_INV_RHO_CP_H];      // Q / (rho_w cp_w H), reciprocal from the host
    if (ocean && !ice) T = T + sub_dt * tend;
    else if (ocean && ice && A.has_ice && P[QD_P_OC_ICE_QFAC] > 0.0) T = T + sub_dt * P[QD_P_OC_ICE_QFAC] * tend;
  }
  T = qd_nan_to_num(T);
  // currents
  const double* ub = A.ub + off;
  const double* vb = A.vb + off;
  double uo = qd_nan_to_num(ub[idx]), vo = qd_nan_to_num(vb[idx]);
  const double cap = P[QD_P_OC_MAX_U];
  const double s2 = uo * uo + vo * vo;
  if (s2 >= P[QD_P_OC_SPEED2_CAP]) {           // == sqrt(s2) > cap (host-computed threshold, engine.py:speed2_threshold): no sqrt on the common path
    const double speed = sqrt(s2);
    // ocean.py:408-434.  Cells at or below the cap (and NaN speeds) fall through unchanged: there the reference
    // multiplies by exactly 1.0, so skipping the second sqrt and the scale is bit-identical.
    if (P[QD_P_OC_MEAN4] != 0.0) {
      const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
      const int jp = j + 1 < nlat ? j + 1 : 0, jm = j > 0 ? j - 1 : nlat - 1;
      const size_t n_ = (size_t)jp * nlon + i, s_ = (size_t)jm * nlon + i, e_ = (size_t)j * nlon + ip, w_ = (size_t)j * nlon + im;
      uo = 0.25 * (qd_nan_to_num(ub[n_]) + qd_nan_to_num(ub[s_]) + qd_nan_to_num(ub[e_]) + qd_nan_to_num(ub[w_]));
      vo = 0.25 * (qd_nan_to_num(vb[n_]) + qd_nan_to_num(vb[s_]) + qd_nan_to_num(vb[e_]) + qd_nan_to_num(vb[w_]));
      const double sp2 = sqrt(uo * uo + vo * vo);
      const double sc2 = (sp2 > cap) ? cap / (sp2 + 1e-12) : 1.0;
      uo = uo * sc2;
      vo = vo * sc2;
    } else {
      const double sc1 = cap / (speed + 1e-12);
      uo = uo * sc1;
      vo = vo * sc1;
    }
  }
  const bool last = (*sc.ctr == (int)S[QD_S_NSUB] - 1);
  if (!A.fuse_mom || last) { A.uo[c] = uo; A.vo[c] = vo; }
  if (A.fuse_mom && !last) {          // momentum step of the next sub-step (np.roll: pole-to-pole wrap in latitude)
    const bool any = P[QD_P_OC_ANY_OCEAN] != 0.0;
    const double cap_e = P[QD_P_OC_ETA_CAP];
    const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
    const int jp = j + 1 < nlat ? j + 1 : 0, jm = j > 0 ? j - 1 : nlat - 1;
    const double* em = A.eta_mid + off;
    qd_oc_momentum_cell(g, P, sub_dt, qd_row(g, QD_R_FCOR)[j], qd_row(g, QD_R_INV_ACOS_HALF)[j], qd_mrow(g, QD_R_OC_SPONGE, b)[j], !ocean,
                        qd_oc_eta_done(em[(size_t)j * nlon + ip], any, eta_mean, cap_e), qd_oc_eta_done(em[(size_t)j * nlon + im], any, eta_mean, cap_e),
                        qd_oc_eta_done(em[(size_t)jp * nlon + i], any, eta_mean, cap_e), qd_oc_eta_done(em[(size_t)jm * nlon + i], any, eta_mean, cap_e),
                        A.taux[c], A.tauy[c], &uo, &vo);
    A.ub_next[c] = uo; A.vb_next[c] = vo;
  }
  if (last && j > 0 && j < nlat - 1) {
    T = qd_clip(T, P[QD_P_OC_TS_MIN], P[QD_P_OC_TS_MAX]);
    if (A.inject && ocean && !ice) A.ts_atm[c] = T;
  }
  A.sst[c] = T;
}
__global__ void __launch_bounds__(QD_THREADS) k_ocean_sst_finish(QdGeo g, QdOcSstBArgs A, QdSubCtl sc) {
  QD_CELL_PROLOGUE(g)
  // the ocean-mean of eta (one true division) once per block instead of once per cell
  __shared__ double s_eta_mean;
  if (threadIdx.x == 0) {
    const double* Pm = g.prm + (size_t)b * QD_P_COUNT;
    s_eta_mean = g.scal[(size_t)b * QD_S_COUNT + QD_S_ETA_NUM] / (Pm[QD_P_OC_WSUM_OCEAN] + 1e-15);
  }
#if !QD_EMU
  __syncthreads();
#endif
  if (!active || qd_sub_done(g, b, sc)) return;
  qd_sstf_cell_general(g, A, sc, b, j, i, s_eta_mean);
}

#if !QD_EMU
// ---- the same pass with TWO adjacent cells per thread (n_lon even) and np.nan_to_num applied lazily.
// The one-cell kernel spent 38 % of its issue slots on index arithmetic, 25 % on the select chains of nan_to_num
// and 8 % on parameter loads against 8 % of fp64 arithmetic (profiles/README.md, SASS mix of the r02 capture): here
// a thread forms one row / column index for two cells, loads them as 16-byte vectors, shares the row constants and
// parameters, and only tracks whether a value the reference would have cleaned was non-finite (one DSETP per
// value); a pair that saw one recomputes with the cleaning applied, from the same registers -- identical bits.
// Rows whose Laplacian is one-sided (j < 2, j > n_lat-3) and cells above the speed cap take the general cell code.
template <bool CLEAN>
__device__ __forceinline__ double qd_lz(double x, bool& bad) {
  if (CLEAN) return qd_nan_to_num(x);
  bad = bad || !(fabs(x) <= DBL_MAX);
  return x;
}
struct QdSstfK {                       // block-uniform constants of the closing pass
  double eta_mean, eta_cap, kh, sub_dt, irch, qfac, ts_min, ts_max, speed2_cap;
  bool any_ocean, use_q, has_ice, ice_q, last, inject;
};
struct QdSstfRow { double cp, cm, ic, ic2; };      // cos(j+1), cos(j-1), 1/cos(j), 1/cos(j)^2 of the half-level table
template <bool CLEAN>
__device__ __forceinline__ void qd_sstf_fast(const QdGeo& g, const QdSstfK& K, const QdSstfRow& R, double e_in, double t0, double tn, double ts,
                                             double te, double tw, double qn, double u, double v, bool ocean, bool ice,
                                             double* eta_o, double* T_o, double* uo_o, double* vo_o, bool* over, bool& bad) {
  double e = e_in;
  if (K.any_ocean) e = e - K.eta_mean;
  *eta_o = qd_clip(qd_lz<CLEAN>(e, bad), -K.eta_cap, K.eta_cap);
  double T = t0;
  if (K.kh > 0.0) {
    // qd_lap_cell's centred form, operand for operand
    const double f0 = qd_lz<CLEAN>(t0, bad);
    const double gp = (qd_lz<CLEAN>(tn, bad) - f0) * g.inv_2dlat, gm = (f0 - qd_lz<CLEAN>(ts, bad)) * g.inv_2dlat;
    const double gphi = (R.cp * gp - R.cm * gm) * g.inv_2dlat;
    const double term_phi = R.ic * gphi;
    const double d2 = ((qd_lz<CLEAN>(te, bad) - 2.0 * f0) + qd_lz<CLEAN>(tw, bad)) * g.inv_dlon_sq;
    const double lap = (term_phi + d2 * R.ic2) * g.inv_a_sq;
    T = f0 + K.sub_dt * K.kh * lap;
  }
  if (K.use_q) {
    const double tend = qn * K.irch;
    if (ocean && !ice) T = T + K.sub_dt * tend;
    else if (ocean && ice && K.ice_q) T = T + K.sub_dt * K.qfac * tend;
  }
  T = qd_lz<CLEAN>(T, bad);
  const double uo = qd_lz<CLEAN>(u, bad), vo = qd_lz<CLEAN>(v, bad);
  const double s2 = uo * uo + vo * vo;
  *over = s2 >= K.speed2_cap;
  if (K.last) T = qd_clip(T, K.ts_min, K.ts_max);
  *T_o = T; *uo_o = uo; *vo_o = vo;
}
// a current above the speed cap (ocean.py:408-434): the general code on that cell's neighbours
__device__ __forceinline__ void qd_sstf_over(const QdGeo& g, const QdOcSstBArgs& A, int b, int j, int ii, double* uo_io, double* vo_io) {
  const size_t off = (size_t)b * g.ncell;
  const int nlon = g.nlon;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double* ub = A.ub + off;
  const double* vb = A.vb + off;
  const double cap = P[QD_P_OC_MAX_U];
  double uo = *uo_io, vo = *vo_io;
  if (P[QD_P_OC_MEAN4] != 0.0) {
    const int ip = ii + 1 < nlon ? ii + 1 : 0, im = ii > 0 ? ii - 1 : nlon - 1;
    const size_t n_ = (size_t)(j + 1) * nlon + ii, s_ = (size_t)(j - 1) * nlon + ii, e_ = (size_t)j * nlon + ip, w_ = (size_t)j * nlon + im;
    uo = 0.25 * (qd_nan_to_num(ub[n_]) + qd_nan_to_num(ub[s_]) + qd_nan_to_num(ub[e_]) + qd_nan_to_num(ub[w_]));
    vo = 0.25 * (qd_nan_to_num(vb[n_]) + qd_nan_to_num(vb[s_]) + qd_nan_to_num(vb[e_]) + qd_nan_to_num(vb[w_]));
    const double sp2 = sqrt(uo * uo + vo * vo);
    const double sc2 = (sp2 > cap) ? cap / (sp2 + 1e-12) : 1.0;
    uo = uo * sc2;
    vo = vo * sc2;
  } else {
    const double sc1 = cap / (sqrt(uo * uo + vo * vo) + 1e-12);
    uo = uo * sc1;
    vo = vo * sc1;
  }
  *uo_io = uo; *vo_io = vo;
}
__global__ void __launch_bounds__(QD_THREADS, QD_LB_FIN2) k_ocean_sst_finish2(QdGeo g, QdOcSstBArgs A, QdSubCtl sc) {
  const int b = blockIdx.y;
  __shared__ QdSstfK sK;
  if (threadIdx.x == 0) {
    const double* P = g.prm + (size_t)b * QD_P_COUNT;
    const double* S = g.scal + (size_t)b * QD_S_COUNT;
    QdSstfK K;
    K.eta_mean = S[QD_S_ETA_NUM] / (P[QD_P_OC_WSUM_OCEAN] + 1e-15);
    K.eta_cap = P[QD_P_OC_ETA_CAP]; K.kh = P[QD_P_OC_K_H]; K.sub_dt = S[QD_S_SUB_DT]; K.irch = P[QD_P_OC_INV_RHO_CP_H];
    K.qfac = P[QD_P_OC_ICE_QFAC]; K.ts_min = P[QD_P_OC_TS_MIN]; K.ts_max = P[QD_P_OC_TS_MAX]; K.speed2_cap = P[QD_P_OC_SPEED2_CAP];
    K.any_ocean = P[QD_P_OC_ANY_OCEAN] != 0.0; K.use_q = (P[QD_P_OC_USE_QNET] != 0.0) && A.has_q; K.has_ice = A.has_ice != 0;
    K.ice_q = A.has_ice && P[QD_P_OC_ICE_QFAC] > 0.0; K.last = (*sc.ctr == (int)S[QD_S_NSUB] - 1); K.inject = A.inject != 0;
    sK = K;
  }
  __syncthreads();
  const int t2 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (t2 >= g.ncomp || qd_sub_done(g, b, sc)) return;
  const int nlon = g.nlon, nlat = g.nlat;
  const int r_ = qd_div_nlon(g, t2);
  const int i = t2 - r_ * nlon;                      // even: n_lon is even
  const int j = qd_seg_row(g, r_);
  if (j < 2 || j > nlat - 3) {
    for (int h = 0; h < 2; ++h) qd_sstf_cell_general(g, A, sc, b, j, i + h, sK.eta_mean);
    return;
  }
  const QdSstfK& K = sK;
  const size_t off = (size_t)b * g.ncell;
  const int idx = j * nlon + i;
  const double* __restrict__ tb = A.tb + off;
  const double* cr = qd_row(g, QD_R_COS_ADV_HALF);
  const QdSstfRow R{cr[j + 1], cr[j - 1], cr[nlat + j], cr[2 * nlat + j]};
  const double* __restrict__ esrc = (A.fuse_mom ? A.eta_mid : A.eta) + off;
  const double2 e2 = *reinterpret_cast<const double2*>(esrc + idx);
  const double2 t2c = *reinterpret_cast<const double2*>(tb + idx);
  double2 tn2 = t2c, ts2 = t2c;
  double tw = 0.0, te = 0.0;
  if (K.kh > 0.0) {
    tn2 = *reinterpret_cast<const double2*>(tb + idx + 2 * nlon);
    ts2 = *reinterpret_cast<const double2*>(tb + idx - 2 * nlon);
    tw = tb[i > 0 ? idx - 1 : idx + nlon - 1];
    te = tb[i + 2 < nlon ? idx + 2 : idx + 2 - nlon];
  }
  double2 q2 = make_double2(0.0, 0.0);
  if (K.use_q) q2 = *reinterpret_cast<const double2*>(A.qnet + off + idx);
  const double2 u2 = *reinterpret_cast<const double2*>(A.ub + off + idx);
  const double2 v2 = *reinterpret_cast<const double2*>(A.vb + off + idx);
  const uchar2 l2 = *reinterpret_cast<const uchar2*>(A.land + off + idx);
  uchar2 c2 = make_uchar2(0, 0);
  if (K.has_ice) c2 = *reinterpret_cast<const uchar2*>(A.ice + off + idx);
  const bool oc0 = l2.x != 1, oc1 = l2.y != 1, ic0 = c2.x != 0, ic1 = c2.y != 0;
  double eo0, eo1, T0, T1, uo0, uo1, vo0, vo1;
  bool ov0, ov1, bad = false;
  qd_sstf_fast<false>(g, K, R, e2.x, t2c.x, tn2.x, ts2.x, t2c.y, tw, q2.x, u2.x, v2.x, oc0, ic0, &eo0, &T0, &uo0, &vo0, &ov0, bad);
  qd_sstf_fast<false>(g, K, R, e2.y, t2c.y, tn2.y, ts2.y, te, t2c.x, q2.y, u2.y, v2.y, oc1, ic1, &eo1, &T1, &uo1, &vo1, &ov1, bad);
  if (bad) {
    qd_sstf_fast<true>(g, K, R, e2.x, t2c.x, tn2.x, ts2.x, t2c.y, tw, q2.x, u2.x, v2.x, oc0, ic0, &eo0, &T0, &uo0, &vo0, &ov0, bad);
    qd_sstf_fast<true>(g, K, R, e2.y, t2c.y, tn2.y, ts2.y, te, t2c.x, q2.y, u2.y, v2.y, oc1, ic1, &eo1, &T1, &uo1, &vo1, &ov1, bad);
  }
  if (ov0) qd_sstf_over(g, A, b, j, i, &uo0, &vo0);
  if (ov1) qd_sstf_over(g, A, b, j, i + 1, &uo1, &vo1);
  *reinterpret_cast<double2*>(A.eta + off + idx) = make_double2(eo0, eo1);
  if (!A.fuse_mom || K.last) {
    *reinterpret_cast<double2*>(A.uo + off + idx) = make_double2(uo0, uo1);
    *reinterpret_cast<double2*>(A.vo + off + idx) = make_double2(vo0, vo1);
  }
  if (A.fuse_mom && !K.last) {
    // momentum step of sub-step s+1 on the pair (rows 2 .. n_lat-3: no wrap in latitude): the finished eta of the six
    // neighbours is formed from continuity's output exactly as their own threads form it
    const double* P = g.prm + (size_t)b * QD_P_COUNT;
    const double2 en2 = *reinterpret_cast<const double2*>(esrc + idx + nlon), es2 = *reinterpret_cast<const double2*>(esrc + idx - nlon);
    const double ew = esrc[i > 0 ? idx - 1 : idx + nlon - 1], ee = esrc[i + 2 < nlon ? idx + 2 : idx + 2 - nlon];
    const double2 tx2 = *reinterpret_cast<const double2*>(A.taux + off + idx), ty2 = *reinterpret_cast<const double2*>(A.tauy + off + idx);
    const double f = qd_row(g, QD_R_FCOR)[j], iach = qd_row(g, QD_R_INV_ACOS_HALF)[j], rex = qd_mrow(g, QD_R_OC_SPONGE, b)[j];
    double mu0 = uo0, mv0 = vo0, mu1 = uo1, mv1 = vo1;
    qd_oc_momentum_cell(g, P, K.sub_dt, f, iach, rex, !oc0, eo1, qd_oc_eta_done(ew, K.any_ocean, K.eta_mean, K.eta_cap),
                        qd_oc_eta_done(en2.x, K.any_ocean, K.eta_mean, K.eta_cap), qd_oc_eta_done(es2.x, K.any_ocean, K.eta_mean, K.eta_cap),
                        tx2.x, ty2.x, &mu0, &mv0);
    qd_oc_momentum_cell(g, P, K.sub_dt, f, iach, rex, !oc1, qd_oc_eta_done(ee, K.any_ocean, K.eta_mean, K.eta_cap), eo0,
                        qd_oc_eta_done(en2.y, K.any_ocean, K.eta_mean, K.eta_cap), qd_oc_eta_done(es2.y, K.any_ocean, K.eta_mean, K.eta_cap),
                        tx2.y, ty2.y, &mu1, &mv1);
    *reinterpret_cast<double2*>(A.ub_next + off + idx) = make_double2(mu0, mu1);
    *reinterpret_cast<double2*>(A.vb_next + off + idx) = make_double2(mv0, mv1);
  }
  *reinterpret_cast<double2*>(A.sst + off + idx) = make_double2(T0, T1);
  if (K.last && K.inject) {
    if (oc0 && !ic0) A.ts_atm[off + idx] = T0;
    if (oc1 && !ic1) A.ts_atm[off + idx + 1] = T1;
  }
}
#endif

// Polar rows: ring means (ocean.py:197-262), final clip, SST injection.  grid = (2 poles, B).
template <int SLOT, class Fn>
QD_D double qd_block_sum_n(int n, Fn fn) {
  __shared__ double res;
#if QD_EMU
  if (threadIdx.x == 0) { double s = 0.0; for (int k = 0; k < n; ++k) s += fn(k); res = s; }
  return res;
#else
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += fn(k);
  double tot;
  if (qd_block_sum<SLOT>(s, &tot)) res = tot;
  __syncthreads();
  return res;
#endif
}
struct QdOcPolarArgs {
  double *sst, *uo, *vo, *ts_atm;
  const uint8_t *land, *ice;
  int has_ice, inject;
  int* step_idx;           // loop mode: the step counter of the forcing table advances here (last kernel of the step)
};
// Two blocks only (one per pole): 1024 threads each, so a ring of 2880 cells takes 3 trips per sweep instead of 12
// (25 -> ~12 us at 1441x2880; the kernel is a chain of three dependent sweeps).
#define QD_POLAR_THREADS 1024
__global__ void __launch_bounds__(QD_POLAR_THREADS) k_ocean_polar(QdGeo g, QdOcPolarArgs A) {
  const int b = blockIdx.y, north = blockIdx.x;
  if (A.step_idx && b == 0 && north == 0 && threadIdx.x == 0) *A.step_idx = *A.step_idx + 1;
  const int j = north ? g.nlat - 1 : 0, n = g.nlon;
  if (!qd_owned(g, j)) return;               // latitude bands: the rank that owns the pole row (full longitude circle)
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const size_t base = (size_t)b * g.ncell + (size_t)j * n;
  double* T = A.sst + base;
  double* U = A.uo + base;
  double* V = A.vo + base;
  const uint8_t* land = A.land + base;
  const double* sl = g.cols + (size_t)QD_C_SIN_LON * n;
  const double* cl = g.cols + (size_t)QD_C_COS_LON * n;
  const double sgn = north ? -1.0 : 1.0;
  if (P[QD_P_OC_POLAR_FIX] != 0.0) {
    // one sweep over the ring for the four sums (count, T, tangent-plane x / y)
    __shared__ double tot[4];
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    QD_BLOCK_FIRST_FOR(k, n) {
      if (land[k] != 1) {
        const double u = U[k], v = V[k];
        c0 += 1.0; c1 += T[k];
        c2 += (-sl[k] * u + sgn * cl[k] * v);
        c3 += (cl[k] * u + sgn * sl[k] * v);
      }
    }
    double t;
#if QD_EMU
    if (threadIdx.x == 0) { tot[0] = c0; tot[1] = c1; tot[2] = c2; tot[3] = c3; }
    (void)t;
#else
    if (qd_block_sum<0>(c0, &t)) tot[0] = t;
    if (qd_block_sum<1>(c1, &t)) tot[1] = t;
    if (qd_block_sum<2>(c2, &t)) tot[2] = t;
    if (qd_block_sum<3>(c3, &t)) tot[3] = t;
    __syncthreads();
#endif
    const double cnt = tot[0];
    if (cnt > 0.0) {
      const double mt = tot[1] / cnt, mx = tot[2] / cnt, my = tot[3] / cnt;
      QD_BLOCK_FIRST_FOR(k, n) {
        if (land[k] != 1) {
          T[k] = mt;
          U[k] = -sl[k] * mx + cl[k] * my;
          V[k] = sgn * cl[k] * mx + sgn * sl[k] * my;
        }
      }
    }
  }
  __syncthreads();
  QD_BLOCK_FIRST_FOR(k, n) {
    const double t = qd_clip(T[k], P[QD_P_OC_TS_MIN], P[QD_P_OC_TS_MAX]);
    T[k] = t;
    const bool ice = A.has_ice ? (A.ice[base + k] != 0) : false;
    if (A.inject && land[k] != 1 && !ice) A.ts_atm[base + k] = t;
  }
}

// End of one sub-step: advance the device-side counter and, inside a CUDA graph, tell the WHILE node
// whether any ensemble member still has sub-steps to do.
#if !QD_EMU
__global__ void k_ocean_sub_advance(QdGeo g, int* sub_ctr, cudaGraphConditionalHandle handle, int use_handle) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int nmax = 1;
  for (int b = 0; b < g.batch; ++b) { const int n = (int)g.scal[(size_t)b * QD_S_COUNT + QD_S_NSUB]; if (n > nmax) nmax = n; }
  const int s = min(*sub_ctr + 1, nmax);                      // saturates: a group of sub-steps may run past the last one (no-ops)
  *sub_ctr = s;
  if (use_handle) cudaGraphSetConditional(handle, s < nmax ? 1u : 0u);
}
#else
__global__ void k_ocean_sub_advance(QdGeo g, int* sub_ctr, unsigned long long handle, int use_handle) {
  (void)handle; (void)use_handle;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int nmax = 1;
  for (int b = 0; b < g.batch; ++b) { const int n = (int)g.scal[(size_t)b * QD_S_COUNT + QD_S_NSUB]; if (n > nmax) nmax = n; }
  *sub_ctr = *sub_ctr + 1 < nmax ? *sub_ctr + 1 : nmax;
}
#endif
