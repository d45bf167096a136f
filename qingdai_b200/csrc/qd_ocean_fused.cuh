// qd_ocean_fused.cuh -- one CFL sub-step of WindDrivenSlabOcean.step (pygcm/ocean.py:305-444) in TWO kernels
// instead of four (+ the del^4 pole tiles):
//
//   k_ocean_fused   momentum (:306-336) -> del^4 of (uo, vo, eta) (:341-356) -> continuity + eta sum (:364-375) ->
//                   SST semi-Lagrangian blend (:380-382) -> outlier handling of the currents (:408-434)
//   k_ocean_close   eta mean removal + hygiene (:375,436-443), SST diffusion + Q_net heating (:384-406), final clip and
//                   SST injection on the member's last sub-step
//
// The four-kernel form re-reads (uo, vo, eta) between momentum, del^4, continuity and the closing kernel: 830 MB of
// DRAM traffic per sub-step at 1441x2880 against ~0.5 GB here (profiles/README.md).  k_ocean_fused is the warp-streaming
// shape of k_hyper4_stream extended on both ends: ONE WARP marches down a strip of 24 longitudes (lanes 4..27; lanes
// 0..3 and 28..31 carry the +-4 column halo that momentum (1) + del^4 (2) + divergence (1) consume), keeping sliding
// latitude windows of the post-momentum currents, their Laplacians and the post-del^4 fields in registers; east / west
// neighbours come from warp shuffles.  No shared memory, no block barriers.  Every expression is the one of the
// unfused kernels (k_ocean_momentum, k_hyper4_stream, k_ocean_continuity, k_ocean_sst_finish), operand for operand, so
// both paths produce identical bits (tests/qdcheck.py:check_ocean_fused_matches_unfused).
//
// Rows: the streaming kernel produces rows [ja, jb) whose whole dependency cone uses centred differences and no
// pole-to-pole wrap; the few rows next to the poles are done by the cell kernels restricted to those rows
// (k_ocean_momentum -> k_hyper4_tile -> k_ocean_cont_pole), launched first so that the streaming kernel's last block
// can add their partial eta sums to its own.
//
// Currents ping-pong between the home slots (QD_F_UO/VO) and one alternate pair; which one a sub-step reads is
// decided ON THE DEVICE from (n_sub, sub-step index) so that the LAST sub-step always writes the home slots and the
// captured WHILE-node body stays valid for any n_sub: with r = n_sub - s sub-steps left, r even reads home / writes
// alternate, r odd reads alternate / writes home; for odd n_sub the first sub-step reads home, writes alternate, and
// k_ocean_close copies alternate -> home.
#pragma once
#include "qd_hyper4.cuh"

#if !QD_EMU
#define QD_OF_COLS 24
#define QD_OF_HALO 4
#define QD_OF_WARPS 4

// which current buffers sub-step s of member b reads (src) / writes (1 - src); copy_back: k_ocean_close copies dst -> home
__device__ __forceinline__ void qd_oc_parity(const QdGeo& g, int b, const QdSubCtl& sc, int* src, int* copy_back) {
  const int n = (int)g.scal[(size_t)b * QD_S_COUNT + QD_S_NSUB], s = *sc.ctr;
  if ((n & 1) && s == 0) { *src = 0; *copy_back = 1; }
  else { *src = ((n - s) & 1) ? 1 : 0; *copy_back = 0; }
}

struct QdOcFusedArgs {
  double *uo[2], *vo[2];                   // [0] home, [1] alternate
  const double *eta, *taux, *tauy, *sst;
  double *eta_out, *tb;                    // eta after continuity (mean not yet removed), SST after the advective blend
  const uint8_t* land;
  const double* k4rows[3]; long long k4_bstride[3]; double scale[3]; int raw_k4[3];
  double* part;                            // [B][npart]: slots [0, part_off) belong to the pole pass, the rest to the warps of this kernel
  int part_off, npart;
  unsigned* ticket;
  int ja, jb;
};

template <int R>
__global__ void __launch_bounds__(32 * QD_OF_WARPS) k_ocean_fused(QdGeo g, QdOcFusedArgs A, QdSubCtl sc) {
  static_assert(R == 32 || R == 64 || R == 96, "k4 rows are staged in up to three registers per lane");
  const int b = blockIdx.y;
  const bool done = qd_sub_done(g, b, sc);
  const int lane = threadIdx.x & 31;
  const int nlat = g.nlat, nlon = g.nlon;
  const int nstrips = (nlon + QD_OF_COLS - 1) / QD_OF_COLS;
  const int w = blockIdx.x * QD_OF_WARPS + (threadIdx.x >> 5);
  const int chunk = w / nstrips, strip = w - chunk * nstrips;
  const int j0 = A.ja + chunk * R;
  double contrib = 0.0;
  if (!done && j0 < A.jb) {
    const int j1 = min(j0 + R, A.jb);
    const size_t off = (size_t)b * g.ncell;
    const double* __restrict__ P = g.prm + (size_t)b * QD_P_COUNT;
    const double sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
    int src, copy_back;
    qd_oc_parity(g, b, sc, &src, &copy_back);
    const double* __restrict__ uo_in = A.uo[src] + off;
    const double* __restrict__ vo_in = A.vo[src] + off;
    double* __restrict__ uo_out = A.uo[1 - src] + off;
    double* __restrict__ vo_out = A.vo[1 - src] + off;
    const double* __restrict__ eta_in = A.eta + off;
    const double* __restrict__ taux = A.taux + off;
    const double* __restrict__ tauy = A.tauy + off;
    const double* __restrict__ sst = A.sst + off;
    const uint8_t* __restrict__ land = A.land + off;
    double* __restrict__ eta_out = A.eta_out + off;
    double* __restrict__ tb = A.tb + off;
    // row tables
    const double* __restrict__ cosh_r = qd_row(g, QD_R_COS_ADV_HALF);
    const double* __restrict__ cap = cosh_r + 3 * (size_t)nlat;        // centred-stencil coefficients of del^4 (qd_cos_companions)
    const double* __restrict__ cam = cosh_r + 4 * (size_t)nlat;
    const double* __restrict__ cbl = cosh_r + 5 * (size_t)nlat;
    const double* __restrict__ iach = qd_row(g, QD_R_INV_ACOS_HALF);
    const double* __restrict__ fcor = qd_row(g, QD_R_FCOR);
    const double* __restrict__ spng = qd_mrow(g, QD_R_OC_SPONGE, b);
    const double* __restrict__ cosr = qd_row(g, QD_R_COS);
    const double* __restrict__ iacc = qd_row(g, QD_R_INV_ACOS_CAP);
    const double* __restrict__ wrow = qd_row(g, QD_R_W);
    // parameters
    const double pG = P[QD_P_OC_G], irH = P[QD_P_OC_INV_RHO_H], rbot = P[QD_P_OC_R_BOT], pH = P[QD_P_OC_H];
    const double al = P[QD_P_OC_ADV_ALPHA], ucap = P[QD_P_OC_MAX_U];
    const bool mean4 = P[QD_P_OC_MEAN4] != 0.0;
    const double inner = sub_dt / 1.0;
    // k4 of this chunk's rows, per field: lane l holds rows j0-1+l (+32, +64) -- del^4 is produced for rows j0-1 .. j1
    double kq[3][R / 32 + 1];
    {
      const double k4div = fmax(1e-12, sub_dt);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* __restrict__ k4r = A.k4rows[k] + (size_t)b * A.k4_bstride[k];
        const bool raw = A.raw_k4[k] != 0;
#pragma unroll
        for (int q = 0; q < R / 32 + 1; ++q) {
          const int j = j0 - 1 + 32 * q + lane;
          double v = (j < nlat) ? k4r[j] : 0.0;
          if (!raw) v = v / k4div;                                   // ocean.py:347
          kq[k][q] = A.scale[k] * v;
        }
      }
    }
    int gi = strip * QD_OF_COLS - QD_OF_HALO + lane;
    if (gi < 0) gi += nlon;
    if (gi >= nlon) gi -= nlon;
    if (gi >= nlon) gi -= nlon;
    const int icol = strip * QD_OF_COLS + lane - QD_OF_HALO;             // unwrapped column of this lane
    const bool writer = lane >= QD_OF_HALO && lane < QD_OF_HALO + QD_OF_COLS && icol < nlon;
    // ---- sliding windows (row m = the momentum row of the current step; continuity row j = m - 5)
    //   raw eta rows m-1, m, m+1 (momentum);  cleaned eta rows m-4 .. m-2 (+ the raw ones, cleaned at use)
    //   cleaned post-momentum currents ub, vb rows m-4 .. m;  Laplacians rows m-6 .. m-2;  post-del^4 rows m-6 .. m-4
    double er_m1, er_0, er_p1;
    double ec4 = 0.0, ec3 = 0.0, ec2 = 0.0;
    double u4 = 0.0, u3 = 0.0, u2 = 0.0, u1 = 0.0;
    double v4 = 0.0, v3 = 0.0, v2 = 0.0, v1 = 0.0;
    double lu6 = 0.0, lu5 = 0.0, lu4 = 0.0, lu3 = 0.0;
    double lv6 = 0.0, lv5 = 0.0, lv4 = 0.0, lv3 = 0.0;
    double le6 = 0.0, le5 = 0.0, le4 = 0.0, le3 = 0.0;
    double U6 = 0.0, U5 = 0.0, V6 = 0.0, V5 = 0.0, E5 = 0.0;
    const int m_first = j0 - 5;                                         // first momentum row: U'(j0-1) <- lap(j0-3) <- ub(j0-5); >= 3 since ja >= 8
    unsigned landbits = 0u;                                             // land flags of rows m, m-1, ... (bit k = row m-k)
    {
      const size_t r0 = (size_t)(m_first - 1) * nlon + gi;
      er_m1 = eta_in[r0]; er_0 = eta_in[r0 + nlon];
    }
    for (int m = m_first; m < j1 + 5; ++m) {
      const size_t rc = (size_t)m * nlon + gi;
      // ---- loads of this step (issued together)
      er_p1 = eta_in[rc + nlon];
      double uo = uo_in[rc], vo = vo_in[rc];
      const double tx = taux[rc], ty = tauy[rc];
      const bool ld = land[rc] == 1;
      landbits = (landbits << 1) | (ld ? 1u : 0u);
      // ---- momentum at row m (ocean.py:306-336), raw eta
      {
        const double ee = __shfl_down_sync(0xffffffffu, er_0, 1), ew = __shfl_up_sync(0xffffffffu, er_0, 1);
        const double de_dl = (ee - ew) * g.inv_2dlon;
        const double de_dp = (er_p1 - er_m1) * g.inv_2dlat;
        const double gx = de_dl * iach[m];
        const double gy = de_dp * g.inv_a;
        const double f = fcor[m];
        const double du = (f * vo - pG * gx + tx * irH - rbot * uo);
        const double dv = (-f * uo - pG * gy + ty * irH - rbot * vo);
        uo = uo + sub_dt * du;
        vo = vo + sub_dt * dv;
        if (ld) { uo = 0.0; vo = 0.0; }
        const double rex = spng[m];
        uo = uo - sub_dt * rex * uo;
        vo = vo - sub_dt * rex * vo;
      }
      const double u0 = qd_clean_sel(uo), v0 = qd_clean_sel(vo), e0 = qd_clean_sel(er_0);
      // ---- Laplacians at row m-2 from rows m-4, m-2, m
      double lu2, lv2, le2;
      {
        const int q = m - 2;                                            // >= 1; warm-up rows only feed values that are never used
        const double ap = cap[q], am = cam[q], bl = cbl[q];
        const double ue = __shfl_down_sync(0xffffffffu, u2, 1), uw = __shfl_up_sync(0xffffffffu, u2, 1);
        lu2 = qd_clean_sel((ap * (u0 - u2) - am * (u2 - u4)) + bl * ((ue - 2.0 * u2) + uw));
        const double ve = __shfl_down_sync(0xffffffffu, v2, 1), vw = __shfl_up_sync(0xffffffffu, v2, 1);
        lv2 = qd_clean_sel((ap * (v0 - v2) - am * (v2 - v4)) + bl * ((ve - 2.0 * v2) + vw));
        const double ee = __shfl_down_sync(0xffffffffu, ec2, 1), ew = __shfl_up_sync(0xffffffffu, ec2, 1);
        le2 = qd_clean_sel((ap * (e0 - ec2) - am * (ec2 - ec4)) + bl * ((ee - 2.0 * ec2) + ew));
      }
      // ---- del^4 update at row m-4 from Laplacian rows m-6, m-4, m-2
      double U4, V4, E4;
      {
        const int q = m - 4 < 0 ? 0 : m - 4;
        const double ap = cap[q], am = cam[q], bl = cbl[q];
        const int r = q - (j0 - 1);                                     // index into the per-lane k4 registers (valid when 0 <= r)
        const int rs = r < 0 ? 0 : r;
        double k4u = kq[0][0], k4v = kq[1][0], k4e = kq[2][0];
#pragma unroll
        for (int qq = 1; qq < R / 32 + 1; ++qq) if (rs >= 32 * qq) { k4u = kq[0][qq]; k4v = kq[1][qq]; k4e = kq[2][qq]; }
        k4u = __shfl_sync(0xffffffffu, k4u, rs & 31); k4v = __shfl_sync(0xffffffffu, k4v, rs & 31); k4e = __shfl_sync(0xffffffffu, k4e, rs & 31);
        const double lue = __shfl_down_sync(0xffffffffu, lu4, 1), luw = __shfl_up_sync(0xffffffffu, lu4, 1);
        const double L2u = (ap * (lu2 - lu4) - am * (lu4 - lu6)) + bl * ((lue - 2.0 * lu4) + luw);
        U4 = qd_clean_sel(u4 - k4u * L2u * inner);
        const double lve = __shfl_down_sync(0xffffffffu, lv4, 1), lvw = __shfl_up_sync(0xffffffffu, lv4, 1);
        const double L2v = (ap * (lv2 - lv4) - am * (lv4 - lv6)) + bl * ((lve - 2.0 * lv4) + lvw);
        V4 = qd_clean_sel(v4 - k4v * L2v * inner);
        const double lee = __shfl_down_sync(0xffffffffu, le4, 1), lew = __shfl_up_sync(0xffffffffu, le4, 1);
        const double L2e = (ap * (le2 - le4) - am * (le4 - le6)) + bl * ((lee - 2.0 * le4) + lew);
        E4 = qd_clean_sel(ec4 - k4e * L2e * inner);
      }
      // ---- continuity, SST blend and current hygiene at row j = m-5 (post-del^4 rows j-1 = m-6, j = m-5, j+1 = m-4)
      const int j = m - 5;
      if (j >= j0) {                                                    // warp-uniform
        const size_t c = (size_t)j * nlon + gi;
        const double Ue = __shfl_down_sync(0xffffffffu, U5, 1), Uw = __shfl_up_sync(0xffffffffu, U5, 1);
        const double du = (Ue - Uw) * g.inv_2dlon;
        const double dv = (V4 * cosr[j + 1] - V6 * cosr[j - 1]) * g.inv_2dlat;
        const double div = iacc[j] * (du + dv);
        const bool ldj = ((landbits >> 5) & 1u) != 0u;
        double e = E5 + (-sub_dt * pH * div);
        if (ldj) e = 0.0;
        double uo2 = U5, vo2 = V5;                                      // nan_to_num of a cleaned value is the value
        const double speed = sqrt(uo2 * uo2 + vo2 * vo2);
        const bool over = speed > ucap;
        if (__any_sync(0xffffffffu, over)) {                            // ocean.py:408-434, rare
          const double Ve = __shfl_down_sync(0xffffffffu, V5, 1), Vw = __shfl_up_sync(0xffffffffu, V5, 1);
          if (over) {
            if (mean4) {
              uo2 = 0.25 * (U4 + U6 + Ue + Uw);
              vo2 = 0.25 * (V4 + V6 + Ve + Vw);
              const double sp2 = sqrt(uo2 * uo2 + vo2 * vo2);
              const double sc2 = (sp2 > ucap) ? ucap / (sp2 + 1e-12) : 1.0;
              uo2 = uo2 * sc2;
              vo2 = vo2 * sc2;
            } else {
              const double sc1 = ucap / (speed + 1e-12);
              uo2 = uo2 * sc1;
              vo2 = vo2 * sc1;
            }
          }
        }
        if (writer) {
          eta_out[c] = e;
          if (qd_owned(g, j)) contrib += e * (wrow[j] * (ldj ? 0.0 : 1.0));
          double y, x;
          qd_departure(U5, V5, sub_dt, g, cosh_r[j], iach[j], j, icol, &y, &x);
          const double adv = qd_bilinear_wrap(sst, nlat, nlon, y, x);
          tb[c] = (1.0 - al) * sst[c] + al * adv;
          uo_out[c] = uo2;
          vo_out[c] = vo2;
        }
      }
      // ---- shift the windows by one row
      ec4 = ec3; ec3 = ec2; ec2 = qd_clean_sel(er_m1);                 // cleaned eta of row m-1 enters the del^4 window
      er_m1 = er_0; er_0 = er_p1;
      u4 = u3; u3 = u2; u2 = u1; u1 = u0;
      v4 = v3; v3 = v2; v2 = v1; v1 = v0;
      lu6 = lu5; lu5 = lu4; lu4 = lu3; lu3 = lu2;
      lv6 = lv5; lv5 = lv4; lv4 = lv3; lv3 = lv2;
      le6 = le5; le5 = le4; le4 = le3; le3 = le2;
      U6 = U5; U5 = U4; V6 = V5; V5 = V4; E5 = E4;
    }
  }
  // ---- eta sum: one partial per warp, combined in warp order by the last block (together with the pole pass's partials)
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
  double* part = A.part + (size_t)b * A.npart;
  if (lane == 0 && A.part_off + w < A.npart) part[A.part_off + w] = done ? 0.0 : contrib;
  if (qd_block_is_last(A.ticket + b, gridDim.x)) {
    double t;
    if (qd_final_sum<1>(part, (unsigned)A.npart, &t)) { if (!done) g.scal[(size_t)b * QD_S_COUNT + QD_S_ETA_NUM] = t; }
  }
}

// ---- rows next to the poles: continuity + eta partial sums + SST blend + current hygiene as a cell kernel (general
// one-sided / wrapping stencils), on the rows of the launch geometry.  ub / vb / eta_in are the post-del^4 fields of the
// pole pass.  Partials go to slots [0, nvb) of the shared partial array; the streaming kernel (launched after this one)
// forms the total.
struct QdOcContPoleArgs {
  const double *ub, *vb, *eta_in, *sst;
  double *uo[2], *vo[2];
  double *eta_out, *tb, *part;
  int npart;
  const uint8_t* land;
};
__global__ void __launch_bounds__(QD_THREADS) k_ocean_cont_pole(QdGeo g, QdOcContPoleArgs A, QdSubCtl sc) {
  const bool done = qd_sub_done(g, blockIdx.y, sc);
  const double* P = g.prm + (size_t)blockIdx.y * QD_P_COUNT;
  const double sub_dt = g.scal[(size_t)blockIdx.y * QD_S_COUNT + QD_S_SUB_DT];
  const double al = P[QD_P_OC_ADV_ALPHA], cap = P[QD_P_OC_MAX_U];
  int src, copy_back;
  qd_oc_parity(g, blockIdx.y, sc, &src, &copy_back);
  double t;
  double* part = A.part + (size_t)blockIdx.y * A.npart;
  const int nlon = g.nlon, nlat = g.nlat;
  QD_VB_LOOP(g) {
    double contrib = 0.0;
    QD_VB_CELLS(g, done ? 0 : g.ncomp) {
      QD_CELL_JI(g)
      const size_t c = off + idx;
      const double* ub = A.ub + off;
      const double* vb = A.vb + off;
      const double div = qd_div_cell(ub, vb, j, i, g);
      double e = A.eta_in[c] + (-sub_dt * P[QD_P_OC_H] * div);
      const bool land = A.land[c] == 1;
      if (land) e = 0.0;
      A.eta_out[c] = e;
      if (qd_owned(g, j)) contrib += e * (qd_row(g, QD_R_W)[j] * (land ? 0.0 : 1.0));
      double y, x;
      qd_departure(ub[idx], vb[idx], sub_dt, g, qd_row(g, QD_R_COS_ADV_HALF)[j], qd_row(g, QD_R_INV_ACOS_HALF)[j], j, i, &y, &x);
      const double adv = qd_bilinear_wrap(A.sst + off, nlat, nlon, y, x);
      A.tb[c] = (1.0 - al) * A.sst[c] + al * adv;
      // currents (ocean.py:408-434)
      double uo = qd_nan_to_num(ub[idx]), vo = qd_nan_to_num(vb[idx]);
      const double speed = sqrt(uo * uo + vo * vo);
      if (speed > cap) {
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
      A.uo[1 - src][c] = uo;
      A.vo[1 - src][c] = vo;
    }
    if (qd_block_sum<0>(contrib, &t)) part[vb_] = t;
  }
}

// ---- closing kernel of a sub-step: eta mean removal + hygiene, SST diffusion + heating, final clip / injection on the
// member's last sub-step, and (first sub-step of an odd n_sub) the copy of the currents back to their home slots.
struct QdOcCloseArgs {
  const double *tb, *eta_mid, *qnet;
  double *uo[2], *vo[2];
  double *sst, *ts_atm, *eta;
  const uint8_t *land, *ice;
  int has_q, has_ice, inject;
};
__global__ void __launch_bounds__(QD_THREADS) k_ocean_close(QdGeo g, QdOcCloseArgs A, QdSubCtl sc) {
  QD_CELL_PROLOGUE(g)
  __shared__ double s_eta_mean;
  if (threadIdx.x == 0) {
    const double* Pm = g.prm + (size_t)b * QD_P_COUNT;
    s_eta_mean = g.scal[(size_t)b * QD_S_COUNT + QD_S_ETA_NUM] / (Pm[QD_P_OC_WSUM_OCEAN] + 1e-15);
  }
  __syncthreads();
  if (!active || qd_sub_done(g, b, sc)) return;
  const size_t c = off + idx;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  const double* S = g.scal + (size_t)b * QD_S_COUNT;
  const double sub_dt = S[QD_S_SUB_DT];
  const int nlat = g.nlat;
  {
    double e = A.eta_mid[c];
    if (P[QD_P_OC_ANY_OCEAN] != 0.0) e = e - s_eta_mean;
    A.eta[c] = qd_clip(qd_nan_to_num(e), -P[QD_P_OC_ETA_CAP], P[QD_P_OC_ETA_CAP]);
  }
  double T = A.tb[c];
  if (P[QD_P_OC_K_H] > 0.0) {
    QdCleanLoad F{A.tb + off, g.nlon};
    const double lap = qd_lap_cell(F, j, i, g, qd_row(g, QD_R_COS_ADV_HALF));
    T = qd_nan_to_num(T) + sub_dt * P[QD_P_OC_K_H] * lap;
  }
  const bool ocean = A.land[c] != 1;
  const bool ice = A.has_ice ? (A.ice[c] != 0) : false;
  if (P[QD_P_OC_USE_QNET] != 0.0 && A.has_q) {
    const double tend = A.qnet[c] * P[QD_P_OC_INV_RHO_CP_H];
    if (ocean && !ice) T = T + sub_dt * tend;
    else if (ocean && ice && A.has_ice && P[QD_P_OC_ICE_QFAC] > 0.0) T = T + sub_dt * P[QD_P_OC_ICE_QFAC] * tend;
  }
  T = qd_nan_to_num(T);
  int src, copy_back;
  qd_oc_parity(g, b, sc, &src, &copy_back);
  if (copy_back) { A.uo[0][c] = A.uo[1][c]; A.vo[0][c] = A.vo[1][c]; }
  const bool last = (*sc.ctr == (int)S[QD_S_NSUB] - 1);
  if (last && j > 0 && j < nlat - 1) {
    T = qd_clip(T, P[QD_P_OC_TS_MIN], P[QD_P_OC_TS_MAX]);
    if (A.inject && ocean && !ice) A.ts_atm[c] = T;
  }
  A.sst[c] = T;
}
#endif
