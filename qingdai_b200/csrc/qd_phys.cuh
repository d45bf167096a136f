// qd_phys.cuh -- per-cell physics of the Qingdai step and the fused phase kernels built on it.
//
// Each QD_HD function restates one reference routine for ONE cell with the reference's operand
// order (file:line cited); the __global__ kernels below fuse the pointwise chains of the step
// so that a prognostic field is read and written as few times as the global dependencies
// (medians, area sums, stencils of freshly written fields) allow.
#pragma once
#include "qd_ops.cuh"
#include "qd_select.cuh"
// Minimum resident blocks per SM (__launch_bounds__) of the large cell kernels.  They are latency-bound (30-58 % of DRAM
// peak AND 35-70 % issue utilisation in the r02b ncu capture), so more resident warps pay even at the price of a few
// spilled registers: measured per kernel at 1441x2880 on one B200 (profiles/README.md), e.g. k_energy 193 -> 141 us at 7
// blocks (36 registers instead of 54), k_column 235 -> 212 us at 5; the whole step 2.13 -> 1.99 ms.  -D overrides for A/B.
#ifndef QD_LB_COL
#define QD_LB_COL 5
#endif
#ifndef QD_LB_EN
#define QD_LB_EN 7
#endif
#ifndef QD_LB_TAIL
#define QD_LB_TAIL 5
#endif
#ifndef QD_LB_ADVMOM
#define QD_LB_ADVMOM 7
#endif
#ifndef QD_LB_CONT2
#define QD_LB_CONT2 6
#endif
#ifndef QD_LB_FIN2
#define QD_LB_FIN2 3
#endif
#ifndef QD_LB_CLOUDA
#define QD_LB_CLOUDA 7
#endif
#ifndef QD_LB_PRECIPA
#define QD_LB_PRECIPA 7
#endif

// ------------------------------------------------------------------------------ humidity (humidity.py)
QD_HD double qd_qsat(double T, double p0) {                       // humidity.py:85-101
  const double Tc = qd_clip(T - 273.15, -80.0, 60.0);
  const double es = 610.94 * QD_EXP(17.625 * Tc / (Tc + 243.04));
  const double den = qd_max(p0 - (1.0 - 0.622) * es, 1.0);
  return qd_clip(0.622 * es / den, 0.0, 0.5);
}
QD_HD double qd_evap_factor(int land, double h_ice, double s_ocean, double s_land, double s_ice) {  // humidity.py:116-142
  if (land) return s_land;
  return (h_ice > 1e-6) ? s_ice : s_ocean;
}

// ------------------------------------------------------------------------------ energy (energy.py)
struct QdSW { double atm, sfc, R; };
QD_HD QdSW qd_shortwave(double I, double albedo, double cloud, double a0, double kc) {   // energy.py:77-98
  const double alpha = qd_clip(albedo, 0.0, 1.0);
  const double Ic = qd_max(0.0, I);
  QdSW o;
  o.R = Ic * alpha;
  const double A = qd_clip(a0 + kc * qd_clip(cloud, 0.0, 1.0), 0.0, 0.95);
  o.atm = Ic * A;
  o.sfc = qd_max(0.0, Ic - o.R - o.atm);
  return o;
}
QD_HD double qd_pow4(double x) { const double x2 = x * x; return x2 * x2; }

struct QdLW { double atm, sfc, olr; };
// energy.py:161-234 (v2: cloud optical depth + surface emissivity map energy.py:141-158) and
// energy.py:101-137 (v1); greenhouse lock overrides OLR/DLR/LW_sfc in both.
QD_HD QdLW qd_longwave(double Ts, double Ta, double cloud, int land, double ice_frac, const double* P) {
  const double sig = QD_SIGMA_SB;
  QdLW o;
  if (P[QD_P_LW_V2] != 0.0) {
    const double Tsc = qd_max(0.0, Ts), Tac = qd_max(0.0, Ta);
    const double Ts4 = qd_pow4(Tsc), Ta4 = qd_pow4(Tac);
    const double eps_clear = qd_clip(P[QD_P_LW_EPS0], 0.0, 1.0);
    const double ce = qd_clip(cloud, 0.0, 1.0);
    const double eps_cloud = qd_clip(1.0 - QD_EXP(-P[QD_P_LW_KTAU] * (P[QD_P_LW_TAU0] * ce)), 0.0, 1.0);
    const double eps_eff = 1.0 - (1.0 - eps_clear) * (1.0 - eps_cloud);
    double es;
    if (land) es = P[QD_P_EPS_LAND];
    else { const double fi = qd_clip(ice_frac, 0.0, 1.0); es = (1.0 - fi) * P[QD_P_EPS_OCEAN] + fi * P[QD_P_EPS_ICE]; }
    es = qd_clip(qd_nan_to_num(es), 0.0, 1.0);
    o.olr = eps_eff * sig * Ta4 + (1.0 - eps_eff) * sig * es * Ts4;
    const double dlr = eps_eff * sig * Ta4;
    o.sfc = dlr - sig * es * Ts4;
    o.atm = eps_eff * (sig * es * Ts4 - 2.0 * sig * Ta4);
    if (P[QD_P_GH_LOCK] != 0.0) {
      const double gt = P[QD_P_GH_FACTOR_LW];
      o.olr = (1.0 - gt) * sig * Ts4;
      o.sfc = gt * sig * Ts4 - sig * es * Ts4;
    }
  } else {
    const double Ts4 = qd_pow4(qd_max(0.0, Ts)), Ta4 = qd_pow4(qd_max(0.0, Ta));
    const double eps = qd_clip(P[QD_P_LW_EPS0] + P[QD_P_LW_KC] * qd_clip(cloud, 0.0, 1.0), 0.0, 1.0);
    o.olr = eps * sig * Ta4 + (1.0 - eps) * sig * Ts4;
    o.sfc = eps * sig * Ta4 - sig * Ts4;
    o.atm = eps * (sig * Ts4 - 2.0 * sig * Ta4);
    if (P[QD_P_GH_LOCK] != 0.0) {
      const double gt = P[QD_P_GH_FACTOR_LW];
      o.olr = (1.0 - gt) * sig * Ts4;
      o.sfc = gt * sig * Ts4 - sig * Ts4;
    }
  }
  return o;
}
QD_HD double qd_ice_frac(double h_ice, double href) {              // dynamics.py:362, run_simulation.py:2065
  return 1.0 - QD_EXP(-qd_max(h_ice, 0.0) / qd_max(1e-6, href));
}
QD_HD double qd_ice_frac_u(double h_ice, const QdRcp& href) {      // same, href = RN-reciprocal pair of max(1e-6, H_ice_ref)
  return 1.0 - QD_EXP(qd_div_u(-qd_max(h_ice, 0.0), href));
}
QD_HD double qd_sensible(double Ts, double Ta, double u, double v, const double* P) {   // energy.py:439-442
  const double V = sqrt(u * u + v * v);
  return P[QD_P_RHO_A] * P[QD_P_CP_AIR] * P[QD_P_C_H] * V * (Ts - Ta);
}
QD_HD double qd_seaice_cs(double Cs) { return (isfinite(Cs) && (Cs > 1e3)) ? Cs : 1e3; }    // energy.py:395-399
// energy.py:291-420 for one cell.  pole_row: 1 = south row, 2 = north row, 0 otherwise.  D: the member's divisor table
// (rho_i L_f, dt and the three sanitised heat capacities are the same for every cell).
QD_HD void qd_seaice_cell_u(double Ts, double Q, double dt, int land, double h_ice, int pole_row,
                            const double* P, const QdRcp* D, double* Ts_out, double* hice_out) {
  const bool ocean = !land;
  double Tn = Ts, hi = h_ice;
  const double rho_i = P[QD_P_RHO_I], L_f = P[QD_P_L_F], t_frz = P[QD_P_T_FREEZE];
  if ((hi > 0.0) && ocean && (Q > 0.0)) {
    const double dh = qd_min(qd_div_u(Q * dt, D[QD_U_RHO_I_LF]), hi);
    hi = hi - dh;
    Q = Q - qd_div_u(dh * rho_i * L_f, D[QD_U_DT]);
  }
  if (ocean && (Q < 0.0) && (Tn <= (t_frz + 0.5))) {
    hi = hi + qd_div_u(-Q * dt, D[QD_U_RHO_I_LF]);
    Q = 0.0;
    Tn = qd_min(Tn, t_frz);
  }
  const QdRcp& Cs = land ? D[QD_U_CS_LAND] : ((hi > 0.0) ? D[QD_U_CS_ICE] : D[QD_U_CS_OCEAN]);
  Tn = Tn + qd_div_u(Q, Cs) * dt;
  if ((pole_row == 1 && P[QD_P_POLAR_FIX_S] != 0.0) || (pole_row == 2 && P[QD_P_POLAR_FIX_N] != 0.0)) {
    if (ocean && (Q < 0.0) && (Tn > t_frz)) Tn = t_frz;
  }
  if ((hi > 0.0) && ocean) Tn = qd_min(Tn, t_frz);
  Tn = qd_max(P[QD_P_T_FLOOR], Tn);
  *Ts_out = qd_nan_to_num(Tn);
  *hice_out = qd_nan_to_num(hi);
}

// ------------------------------------------------------------------------------ forcing columns
// cos(theta + lon - alpha) per column and star (forcing.py:126-131); one thread per column.
QD_D void qd_forcing_col(const QdGeo& g, const qd_forcing_t* forcing, const int* step_idx, double* hcos /*[2][nlon]*/, int i) {
  const qd_forcing_t F = forcing[*step_idx];
  const double lon = g.cols[(size_t)QD_C_LON_RAD * g.nlon + i];
  hcos[i] = cos(F.theta + lon - F.alpha_a);
  hcos[g.nlon + i] = cos(F.theta + lon - F.alpha_b);
}
__global__ void k_forcing_cols(QdGeo g, const qd_forcing_t* forcing, const int* step_idx, double* hcos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.nlon) qd_forcing_col(g, forcing, step_idx, hcos, i);
}

// ------------------------------------------------------------------------------ column physics
// One pointwise pass that fuses, for the script loop (mode_loop=1): dual-star insolation
// (forcing.py:105-136), P019 lapse/phase/snowpack/glacier (run_simulation.py:1948-2008,
// hydrology.py:100-177), ecology sub-daily albedo (adapter.py:140-186), albedo synthesis
// (run_simulation.py:2064-2146, physics.py:164-250), Teq (forcing.py:138-165), and then the
// pointwise head of SpectralModel.time_step: humidity (dynamics.py:274-297), Newtonian Ts
// (dynamics.py:304-322), radiative h relaxation (dynamics.py:464-467), plus the land bucket
// (hydrology.py:219-260, run_simulation.py:2304-2339) which only needs this step's E.
// For the stand-alone time_step (mode_loop=0) Teq / albedo / isr are read from their fields.
// With has_albedo the energy branch needs median(P_cond>0) first, so it runs in k_energy below.
struct QdColArgs {
  // prognostics (in place unless noted)
  double *u, *v, *h, *ts, *q, *cloud, *hice, *wland, *ssnow, *eday;
  // outputs
  double *ts_pre, *q_pre;              // Ts / q before the semi-Lagrangian blend (scratch)
  double *isr, *isr_a, *isr_b, *olr, *eflux, *pcond, *lh, *lhrel, *albedo, *teq, *csnow, *rland, *alpha_eco;
  // inputs
  const double *precip, *cloud_eff, *base_albedo, *elevation, *fcanopy, *hcos;
  const uint8_t* land; uint8_t* glacier;
  const qd_forcing_t* forcing; const int* step_idx;
  double dt;
  int mode_loop, has_albedo, has_cloud_eff, with_hydrology, with_eco, store_isr_ab;
};

__global__ void __launch_bounds__(QD_THREADS, QD_LB_COL) k_column(QdGeo g, QdColArgs A) {
  QD_CELL_PROLOGUE(g)
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const QdRcp* D = g.udiv + (size_t)b * QD_U_COUNT;
  const double dt = A.dt;
  if (!active) return;
  const size_t c = off + idx;
  const int land = A.land[c] == 1;
  const double h = A.h[c], Ts = A.ts[c], q0 = A.q[c], u = A.u[c], v = A.v[c], hice = A.hice[c];
  // ---- humidity (dynamics.py:274-297, humidity.py:145-183): start-of-step Ts, q, u, v, h
  const double Ta = 288.0 + D[QD_Q_G_CP].r * h;
  const double fac = qd_evap_factor(land, hice, P[QD_P_EVAP_OCEAN], P[QD_P_EVAP_LAND], P[QD_P_EVAP_ICE]);
  const double V = sqrt(u * u + v * v);
  const double deficit = qd_max(0.0, qd_qsat(Ts, P[QD_P_P0]) - q0);
  const double E = qd_nan_to_num(P[QD_P_RHO_A] * P[QD_P_C_E] * V * deficit * fac);
  const double LH = P[QD_P_L_V] * E;
  const double M_col = D[QD_U_MCOL].b;
  const double q_evap = q0 + qd_div_u(E, D[QD_U_MCOL]) * dt;
  const double excess = qd_max(0.0, q_evap - qd_qsat(Ta, P[QD_P_P0]));
  double P_cond = qd_div_u(excess, D[QD_U_TAU_COND]) * M_col;
  double q_next = q_evap - qd_div_u(P_cond, D[QD_U_MCOL]) * dt;
  q_next = qd_clip(qd_nan_to_num(q_next), 0.0, 0.5);
  P_cond = qd_nan_to_num(P_cond);
  A.eflux[c] = E; A.pcond[c] = P_cond; A.lh[c] = LH; A.lhrel[c] = P[QD_P_L_V] * P_cond;
  A.q_pre[c] = q_next;

  double Teq;
  if (A.mode_loop) {
    // ---- insolation
    const qd_forcing_t F = A.forcing[*A.step_idx];
    const double sl = qd_row(g, QD_R_SIN)[j], cl = qd_row(g, QD_R_COS)[j];
    const double cza = sl * F.sin_delta_a + cl * F.cos_delta_a * A.hcos[i];
    const double czb = sl * F.sin_delta_b + cl * F.cos_delta_b * A.hcos[g.nlon + i];
    const double ia = F.flux_a * qd_max(0.0, cza), ib = F.flux_b * qd_max(0.0, czb);
    const double isr = ia + ib;
    A.isr[c] = isr;
    if (A.store_isr_ab) { A.isr_a[c] = ia; A.isr_b[c] = ib; }

    // ---- P019: lapse-adjusted air temperature, smooth phase split, snowpack, glacier mask
    const double precip = A.precip[c];
    const double Ta_proxy = 288.0 + (9.81 / 1004.0) * h;
    const double Hb = (P[QD_P_HAS_ELEVATION] != 0.0) ? A.elevation[c] : 0.0;
    const double S0 = A.ssnow[c];
    const double hs_geom = land ? qd_div_u(qd_max(S0, 0.0), D[QD_U_RHO_SNOW]) : 0.0;
    const double hs_eff = (qd_mrow(g, QD_R_POLAR, b)[j] != 0.0) ? qd_min(hs_geom, P[QD_P_POLAR_ICE_THICK_MAX]) : hs_geom;
    const double H_eff = qd_min(Hb + hs_eff, P[QD_P_LAND_ELEV_MAX]);
    const double T_hat = (P[QD_P_LAPSE_ENABLE] != 0.0) ? Ta_proxy - P[QD_P_LAPSE_KPM] * qd_div_u(H_eff, D[QD_U_1000]) : Ta_proxy;
    const double f_snow = qd_clip(1.0 / (1.0 + QD_EXP(qd_div_u(T_hat - P[QD_P_SNOW_THRESH], D[QD_U_SNOW_BAND]))), 0.0, 1.0);
    const double P_snow = qd_nan_to_num(f_snow * precip);
    const double P_rain = qd_nan_to_num((1.0 - f_snow) * precip);
    double C_snow = 0.0, S_next = S0, melt = 0.0;
    int glacier;
    if (P[QD_P_SWE_ENABLE] != 0.0) {
      const double Ps_land = P_snow * (land ? 1.0 : 0.0);
      double mflux;
      if (P[QD_P_SNOW_DEGREE_DAY] != 0.0) mflux = D[QD_Q_DDF].r * qd_max(T_hat - P[QD_P_SNOW_MELT_TREF], 0.0);
      else mflux = (T_hat >= P[QD_P_SNOW_THRESH]) ? D[QD_Q_MELT].r : 0.0;
      const double amt = qd_min(qd_max(S0, 0.0), mflux * dt);
      S_next = S0 + Ps_land * dt - amt;
      if (P[QD_P_SWE_MAX] > 0.0) S_next = qd_min(S_next, P[QD_P_SWE_MAX]);
      S_next = qd_max(0.0, S_next);
      melt = qd_nan_to_num(qd_div_u(amt, D[QD_U_DT]));
      C_snow = qd_clip(1.0 - QD_EXP(qd_div_u(-qd_max(S_next, 0.0), D[QD_U_SWE_REF])), 0.0, 1.0);
      S_next = qd_nan_to_num(S_next);
      glacier = land && ((C_snow >= P[QD_P_GLACIER_FRAC]) || (S_next >= P[QD_P_GLACIER_SWE]));
      const double rain_gl = (P_rain * (land ? 1.0 : 0.0)) * (glacier ? 1.0 : 0.0);
      if (rain_gl != 0.0) S_next = S_next + rain_gl * dt;   // np.any() guard: adding 0*dt elsewhere is a no-op
    } else {
      glacier = land && (C_snow >= P[QD_P_GLACIER_FRAC]);
    }
    A.csnow[c] = C_snow;
    A.glacier[c] = (uint8_t)glacier;

    // ---- albedo synthesis
    const double ice_frac = qd_ice_frac_u(hice, D[QD_U_HICE_REF]);
    const double cloud_rad = A.has_cloud_eff ? A.cloud_eff[c] : A.cloud[c];
    double base = (P[QD_P_USE_TOPO_ALBEDO] != 0.0) ? A.base_albedo[c] : P[QD_P_ALPHA_WATER];
    if (A.with_eco) {
      // adapter.py:140-186: E_day += nan_to_num(isr)*dt; alpha = leaf*f + (1-f)*soil on land, NaN on ocean
      A.eday[c] = A.eday[c] + qd_nan_to_num(isr) * dt;
      // with_eco: 1 = the adapter produces a new alpha map this step, 2 = the script re-uses the last map
      // (run_simulation.py:2083-2086), 3 = no map yet
      double alpha_e = NAN;
      if (A.with_eco == 1) {
        if (land) {
          const double fc = A.fcanopy[c];
          alpha_e = qd_clip(P[QD_P_ECO_ALPHA_LEAF] * fc + (1.0 - fc) * P[QD_P_ECO_SOIL_REFLECT], 0.0, 1.0);
        }
        A.alpha_eco[c] = alpha_e;
      } else if (A.with_eco == 2) {
        alpha_e = A.alpha_eco[c];
      }
      if (land && !glacier && isfinite(alpha_e))
        base = (1.0 - P[QD_P_ECO_W_LAI]) * base + P[QD_P_ECO_W_LAI] * alpha_e;
    }
    if (P[QD_P_SWE_ENABLE] != 0.0 && land)
      base = qd_clip((1.0 - C_snow) * base + C_snow * P[QD_P_SNOW_ALBEDO_FRESH], 0.0, 1.0);
    const double Cc = qd_clip(cloud_rad, 0.0, 1.0);
    const double fi = qd_clip(ice_frac, 0.0, 1.0) * (land ? 0.0 : 1.0);
    const double surf = base * (1.0 - fi) + P[QD_P_ALPHA_ICE] * fi;
    const double albedo = qd_clip(surf * (1.0 - Cc) + P[QD_P_ALPHA_CLOUD] * Cc, 0.0, 1.0);
    A.albedo[c] = albedo;
    double num = isr * (1 - albedo);
    if (num < 0) num = 0.0;
    Teq = sqrt(sqrt(qd_div_u(num, D[QD_U_SIGMA])));
    A.teq[c] = Teq;

    // ---- hydrology commit (run_simulation.py:2304-2339, hydrology.py:219-260) with this step's E
    if (A.with_hydrology) {
      A.ssnow[c] = S_next;
      const double non_gl = (land && !glacier) ? 1.0 : 0.0;
      const double P_in = (P_rain * (land ? 1.0 : 0.0) + melt) * non_gl;
      const double W = A.wland[c];
      const double R_base = qd_div_u(W, D[QD_U_RUNOFF_TAU]);
      const double E_land = (E * (land ? 1.0 : 0.0)) * non_gl;
      double W_next = qd_max(0.0, W + (P_in - E_land - R_base) * dt);
      double R_fast = 0.0;
      if (P[QD_P_WLAND_CAP] > 0.0) {
        const double over = qd_max(0.0, W_next - P[QD_P_WLAND_CAP]);
        W_next = W_next - over;
        R_fast = qd_div_u(over, D[QD_U_DT]);
      }
      A.wland[c] = qd_nan_to_num(W_next);
      A.rland[c] = qd_nan_to_num(R_base + R_fast) + melt * (glacier ? 1.0 : 0.0);
    }
  } else {
    Teq = A.teq[c];
  }

  // ---- Newtonian surface update (dynamics.py:304-322)
  const double olr_old = QD_SIGMA_SB * qd_pow4(Ts);
  const double net_old = QD_SIGMA_SB * qd_pow4(Teq) + P[QD_P_GH_NEWTON] * QD_SIGMA_SB * qd_pow4(Ta) - olr_old;
  const double Ts_newton = Ts + qd_div_u(net_old, D[QD_U_C_SFC]) * dt;
  A.ts_pre[c] = Ts_newton;
  if (!A.has_albedo) {
    A.olr[c] = olr_old;
    // radiative relaxation of h (dynamics.py:464-467); with albedo it is applied in k_energy
    const double h_eq = D[QD_Q_R_G].r * Teq;
    A.h[c] = h + qd_div_u(h_eq - h, D[QD_U_TAU_RAD]) * dt;
  }
}

// Energy branch of time_step (dynamics.py:326-449,470-478) for one cell; runs after k_column and the
// median of the fresh P_cond.  Reads the OLD Ts and h (k_column left them untouched when has_albedo).
struct QdEnergyArgs {
  double *h, *hice, *ts_pre, *olr, *cloud_eff;
  const double *ts, *q_pre, *cloud, *pcond, *isr, *albedo, *teq, *u, *v, *lh, *lhrel, *cs_map;
  const uint8_t* land;
  double dt;
};
__global__ void __launch_bounds__(QD_THREADS, QD_LB_EN) k_energy(QdGeo g, QdEnergyArgs A) {
  QD_CELL_PROLOGUE(g)
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const QdRcp* D = g.udiv + (size_t)b * QD_U_COUNT;
  const double dt = A.dt;
  if (!active) return;
  const size_t c = off + idx;
  const int land = A.land[c] == 1;
  const double h = A.h[c], Ts = A.ts[c], hice = A.hice[c], u = A.u[c], v = A.v[c];
  const double Ta = 288.0 + D[QD_Q_G_CP].r * h;
  double cloud_eff = A.cloud[c];
  if (P[QD_P_CLOUD_COUPLE] != 0.0) {
    const double RH = qd_clip(qd_div_z(A.q_pre[c], qd_max(1e-12, qd_qsat(Ta, P[QD_P_P0]))), 0.0, 1.5);
    const double rh_ex = qd_max(0.0, RH - P[QD_P_RH0]);
    const double Pc = A.pcond[c];
    const double pref_over = P[QD_P_PCOND_REF];
    const double P_ref = (pref_over == pref_over) ? pref_over : g.scal[(size_t)b * QD_S_COUNT + QD_S_PREF_ATM];
    const double p_term = QD_TANH((P_ref > 0) ? qd_div_z(Pc, P_ref) : 0.0);          // Pc is zero wherever nothing condenses
    cloud_eff = qd_clip(cloud_eff + P[QD_P_K_Q] * rh_ex + P[QD_P_K_P] * p_term, 0.0, 1.0);
  }
  A.cloud_eff[c] = cloud_eff;
  const QdSW sw = qd_shortwave(A.isr[c], A.albedo[c], cloud_eff, P[QD_P_SW_A0], P[QD_P_SW_KC]);
  const QdLW lw = qd_longwave(Ts, Ta, cloud_eff, land, qd_ice_frac_u(hice, D[QD_U_HICE_REF]), P);
  const double SH = qd_sensible(Ts, Ta, u, v, P);
  const double LH = A.lh[c];
  double Ts_e, hi_next = hice;
  if (P[QD_P_SEAICE] != 0.0) {
    const double Q = sw.sfc - lw.sfc - SH - LH;
    const int pole = (j == 0) ? 1 : ((j == g.nlat - 1) ? 2 : 0);
    qd_seaice_cell_u(Ts, Q, dt, land, hice, pole, P, D, &Ts_e, &hi_next);
  } else {
    const double net = sw.sfc - lw.sfc - SH - LH;
    double Cs = A.cs_map ? A.cs_map[c] : P[QD_P_C_SFC];
    if (A.cs_map) { if (!(isfinite(Cs) && (Cs > 1e3))) Cs = 1e3; }
    const double dT = A.cs_map ? qd_div_z(net, Cs) : qd_div_u(net, D[QD_U_C_SFC]);
    Ts_e = qd_nan_to_num(qd_max(P[QD_P_T_FLOOR], Ts + dT * dt));
  }
  A.olr[c] = lw.olr;
  const double w = fmin(1.0, fmax(0.0, P[QD_P_ENERGY_W]));
  A.ts_pre[c] = (1.0 - w) * A.ts_pre[c] + w * Ts_e;
  if (P[QD_P_SEAICE] != 0.0) A.hice[c] = hi_next;
  // h: radiative relaxation then M3 energy coupling (energy.py:452-491)
  const double h_eq = D[QD_Q_R_G].r * A.teq[c];
  double hn = h + qd_div_u(h_eq - h, D[QD_U_TAU_RAD]) * dt;
  if (P[QD_P_ENERGY_W] > 0.0) {
    const double F_atm = sw.atm + lw.atm + SH + A.lhrel[c];
    hn = qd_nan_to_num(hn + P[QD_P_ENERGY_W] * qd_div_u(F_atm, D[QD_U_ATM]) * dt);
  }
  A.h[c] = hn;
}

// ------------------------------------------------------------------------------ advect Ts,q + momentum
// dynamics.py:454-461 (gentle semi-Lagrangian blend, alpha=0.2) and :484-530 (geostrophic
// relaxation or primitive momentum with np.gradient one-sided at BOTH lon edges).  u,v are only
// read at the own cell, so the momentum update can be done in place in the same pass.
struct QdAdvMomArgs {
  const double *ts_pre, *q_pre, *h, *friction;
  double *ts, *q, *u, *v;
  double dt;
};
__global__ void __launch_bounds__(QD_THREADS, QD_LB_ADVMOM) k_advect_momentum(QdGeo g, QdAdvMomArgs A) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double dt = A.dt;
  const double u0 = A.u[c], v0 = A.v[c];
  double y, x;
  qd_departure(u0, v0, dt, g, qd_row(g, QD_R_COS_ADV_ATM)[j], qd_row(g, QD_R_IAC_ADV_ATM)[j], j, i, &y, &x);
  const double adv_t = qd_bilinear_wrap(A.ts_pre + off, g.nlat, g.nlon, y, x);
  const double adv_q = qd_bilinear_wrap(A.q_pre + off, g.nlat, g.nlon, y, x);
  A.ts[c] = (1.0 - 0.2) * A.ts_pre[c] + 0.2 * adv_t;
  A.q[c] = qd_clip(qd_nan_to_num((1.0 - 0.2) * A.q_pre[c] + 0.2 * adv_q), 0.0, 0.5);
  // gradients of the UPDATED h
  const double* hh = A.h + off;
  const int nlon = g.nlon, nlat = g.nlat;
  // np.gradient: one-sided at both lon edges and at the poles; the four divisors are grid constants whose
  // correctly rounded reciprocals come with the geometry
  const QdRcp d_lon1{g.dlon, g.inv_dlon}, d_lon2{2.0 * g.dlon, g.inv_2dlon};
  const QdRcp d_lat1{g.dlat, g.inv_dlat}, d_lat2{2.0 * g.dlat, g.inv_2dlat};
  double dh_dlon, dh_dlat;
  if (i == 0) dh_dlon = qd_div_u(hh[(size_t)j * nlon + 1] - hh[(size_t)j * nlon], d_lon1);
  else if (i == nlon - 1) dh_dlon = qd_div_u(hh[(size_t)j * nlon + i] - hh[(size_t)j * nlon + i - 1], d_lon1);
  else dh_dlon = qd_div_u(hh[(size_t)j * nlon + i + 1] - hh[(size_t)j * nlon + i - 1], d_lon2);
  if (j == 0) dh_dlat = qd_div_u(hh[(size_t)nlon + i] - hh[i], d_lat1);
  else if (j == nlat - 1) dh_dlat = qd_div_u(hh[(size_t)j * nlon + i] - hh[(size_t)(j - 1) * nlon + i], d_lat1);
  else dh_dlat = qd_div_u(hh[(size_t)(j + 1) * nlon + i] - hh[(size_t)(j - 1) * nlon + i], d_lat2);
  const double cosc = qd_row(g, QD_R_COS_CAP)[j];
  const double fr = A.friction[c];
  double un, vn;
  if (P[QD_P_MOM_PRIMITIVE] != 0.0) {
    const double f = qd_row(g, QD_R_FCOR)[j];
    const double PGx = -(P[QD_P_G] / (g.a * cosc)) * dh_dlon;
    const double PGy = -(P[QD_P_G] / g.a) * dh_dlat;
    un = qd_clip(u0 + (PGx + f * v0 - fr * u0) * dt, -200.0, 200.0);
    vn = qd_clip(v0 + (PGy - f * u0 - fr * v0) * dt, -200.0, 200.0);
  } else {
    const double fs = qd_row(g, QD_R_FSAFE)[j];
    const double ug = qd_clip(-(P[QD_P_G] / (fs * g.a * cosc)) * dh_dlat, -200.0, 200.0);
    const double vg = qd_clip((P[QD_P_G] / (fs * g.a)) * dh_dlon, -200.0, 200.0);
    un = u0 * 0.8 + ug * 0.2;
    vn = v0 * 0.8 + vg * 0.2;
    un = un + (-fr * un) * dt;
    vn = vn + (-fr * vn) * dt;
  }
  A.u[c] = un;
  A.v[c] = vn;
}

// ------------------------------------------------------------------------------ atmosphere tail (+ Q_net)
// dynamics.py:644-667: cloud <- advected cloud * (1 - dt/2d), x diff_factor on u,v,h,cloud,q,
// nan_to_num on all six; fused with the loop's surface heat flux (run_simulation.py:2201-2239)
// and the block maxima that decide the ocean sub-step count (ocean.py:285-303).
struct QdTailArgs {
  const double *u_in, *v_in, *h_in, *q_in;      // where del^4 / filters left the fields
  double *u, *v, *h, *ts, *q, *cloud;
  const double *cloud_adv, *hice, *isr, *albedo, *cloud_eff, *lh, *uo, *vo;
  double *qnet; uint8_t* ice; const uint8_t* land;
  double *part_max_u, *part_max_va; unsigned* ticket;
  double dt; int with_qnet, has_cloud_eff, with_max;
  // with_max (loop mode with the dynamic ocean, one GPU): this pass also does the ocean step's preparation
  // (k_ocean_prep: wind stress ocean.py:285-290, the two maxima and n_sub ocean.py:293-303) -- it has the final winds and
  // the currents in registers already, so the ocean step starts without re-reading four fields
  double *taux, *tauy; int* sub_ctr;
};
QD_D void qd_ocean_nsub_member(const QdGeo& g, int b, double dt);      // qd_ocean.cuh
__global__ void __launch_bounds__(QD_THREADS, QD_LB_TAIL) k_tail(QdGeo g, QdTailArgs A) {
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  double mu = 0.0, mva = 0.0;
  const QdRcp* D = g.udiv + (size_t)blockIdx.y * QD_U_COUNT;
  QD_CELL_LOOP(g) {
    QD_CELL_JI(g)
    const size_t c = off + idx;
    const double df = P[QD_P_DIFF_FACTOR];
    const double rate = A.dt / (2.0 * 24 * 3600);
    const double u = qd_nan_to_num(A.u_in[c] * df), v = qd_nan_to_num(A.v_in[c] * df), h = qd_nan_to_num(A.h_in[c] * df);
    const double q = qd_nan_to_num(A.q_in[c] * df), ts = qd_nan_to_num(A.ts[c]);
    const double cl = qd_nan_to_num((A.cloud_adv[c] * (1 - rate)) * df);
    A.u[c] = u; A.v[c] = v; A.h[c] = h; A.q[c] = q; A.ts[c] = ts; A.cloud[c] = cl;
    if (A.with_qnet) {
      const int land = A.land[c] == 1;
      const double hice = A.hice[c];
      A.ice[c] = (uint8_t)(hice > 0.0);
      const double ce = A.has_cloud_eff ? A.cloud_eff[c] : cl;
      const QdSW sw = qd_shortwave(A.isr[c], A.albedo[c], ce, P[QD_P_SW_A0], P[QD_P_SW_KC]);
      const double Ta = 288.0 + (9.81 / 1004.0) * h;
      const QdLW lw = qd_longwave(ts, Ta, ce, land, qd_ice_frac_u(hice, D[QD_U_HICE_REF]), P);
      const double SH = qd_sensible(ts, Ta, u, v, P);
      A.qnet[c] = sw.sfc - lw.sfc - SH - A.lh[c];
    }
    if (A.with_max) {                           // k_ocean_prep's cell code, operand for operand
      const double uo = A.uo[c], vo = A.vo[c];
      const double ur = u - uo, vr = v - vo;
      const double sp = sqrt(uo * uo + vo * vo), va = sqrt(ur * ur + vr * vr);
      const double Ve = qd_min(va, P[QD_P_OC_VCAP]);
      A.taux[c] = P[QD_P_OC_TAU_SCALE] * (P[QD_P_OC_RHO_A] * P[QD_P_OC_CD] * Ve * ur);
      A.tauy[c] = P[QD_P_OC_TAU_SCALE] * (P[QD_P_OC_RHO_A] * P[QD_P_OC_CD] * Ve * vr);
      if (sp > mu) mu = sp;                     // NaN-ignoring running maxima, like qd_block_max
      if (va > mva) mva = va;
    }
  }
  if (A.with_max) {
    double t;
    if (qd_block_max<0>(mu, &t)) A.part_max_u[(size_t)b * gridDim.x + blockIdx.x] = t;
    if (qd_block_max<1>(mva, &t)) A.part_max_va[(size_t)b * gridDim.x + blockIdx.x] = t;
    if (qd_block_is_last(A.ticket + b, gridDim.x)) {
      double* S = g.scal + (size_t)b * QD_S_COUNT;
      double m1 = 0.0, m2 = 0.0;
      const bool o1 = qd_final_max<2>(A.part_max_u + (size_t)b * gridDim.x, gridDim.x, &m1);
      const bool o2 = qd_final_max<3>(A.part_max_va + (size_t)b * gridDim.x, gridDim.x, &m2);
      if (o1 && o2) {                           // the owner thread of both reductions is the same thread
        S[QD_S_MAX_UOCEAN] = m1;
        S[QD_S_MAX_VA] = m2;
        if (A.sub_ctr) { if (b == 0) *A.sub_ctr = 0; qd_ocean_nsub_member(g, b, A.dt); }
      }
    }
  }
}
