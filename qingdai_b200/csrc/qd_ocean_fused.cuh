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
  *src = qd_oc_src(g, b, sc);
  *copy_back = ((n & 1) && s == 0) ? 1 : 0;
}

struct QdOcFusedArgs {
  double *uo[2], *vo[2];                   // [0] home, [1] alternate
  const double *eta, *taux, *tauy, *sst;
  double *eta_out, *tb;                    // eta after continuity (mean not yet removed), SST after the advective blend
  const uint8_t* land;
  const double* k4tab;                     // [B][3][nlat]: k4 of (uo, vo, eta) per row for this step's sub_dt (k_ocean_k4tab)
  double* part;                            // [B][npart]: slots [0, part_off) belong to the pole pass, the rest to the warps of this kernel
  int part_off, npart;
  unsigned* ticket;
  int ja, jb, R;                           // rows [ja, jb) in chunks of R rows per warp
};

// Once per ocean step, after n_sub is known: k4 = scale * (raw ? row : row / max(1e-12, sub_dt)) per row and field
// (ocean.py:343-352) -- one thread per (row, field, member).
struct QdOcK4Args { const double* k4rows[3]; long long k4_bstride[3]; double scale[3]; int raw_k4[3]; double* tab; };
__global__ void k_ocean_k4tab(QdGeo g, QdOcK4Args A) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y, b = blockIdx.z;
  if (j >= g.nlat) return;
  const double sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  double v = A.k4rows[k][(size_t)b * A.k4_bstride[k] + j];
  if (!A.raw_k4[k]) v = v / fmax(1e-12, sub_dt);
  A.tab[((size_t)b * 3 + k) * g.nlat + j] = A.scale[k] * v;
}

__device__ __forceinline__ void qd_cp_async8(void* smem, const void* gptr) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gptr) : "memory");
}
__device__ __forceinline__ void qd_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void qd_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Input staging: every lane copies its own column of the five momentum inputs (uo, vo, taux, tauy of row r and eta of row
// r+1) and of the SST row r, and lanes 0..12 one per-row constant each (metric / Coriolis / sponge / del^4 coefficient /
// k4 rows), into per-warp shared-memory rings with cp.async (LDGSTS), QD_OF_D rows ahead of their use: the loads of
// QD_OF_D rows per warp are in flight without holding registers, and the ~17 row-constant reads of a step are
// shared-memory broadcasts instead of global loads that the streaming traffic keeps evicting from L1.  The SST ring
// keeps rows j-1 .. j+1 of the continuity row j = m-5 alive so that the semi-Lagrangian gather (CFL < 1: the departure
// point lies within one cell) reads its four taps from shared memory instead of global memory.
#define QD_OF_D 3                                   // rows in flight
#define QD_OF_NM 4                                  // slots of the momentum-input ring (rows m .. m+D)
#define QD_OF_NS 16                                 // slots of the SST / row-constant rings (rows m-6 .. m+D)
enum { QD_RC_IACH = 0, QD_RC_FCOR, QD_RC_SPNG, QD_RC_CAP, QD_RC_CAM, QD_RC_CBL, QD_RC_K4U, QD_RC_K4V, QD_RC_K4E, QD_RC_COS, QD_RC_IACC,
       QD_RC_W, QD_RC_COSH, QD_RC_COUNT };

// End of a warp's work: publish its partial eta sum; the LAST WARP of the member's grid to arrive (one atomic ticket per
// warp, no block barrier -- idle warps leave early and the working warps keep their shuffles at the top level of the
// kernel) adds the partials in slot order with a fixed shuffle tree: the total depends on the geometry only.
__device__ __forceinline__ void qd_of_finish(const QdGeo& g, const QdOcFusedArgs& A, int b, int w, int lane, double contrib, bool done) {
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
  double* part = A.part + (size_t)b * A.npart;
  int last = 0;
  if (lane == 0) {
    if (A.part_off + w < A.npart) part[A.part_off + w] = done ? 0.0 : contrib;
    __threadfence();
    const unsigned total = gridDim.x * QD_OF_WARPS;
    const unsigned t = atomicAdd(A.ticket + b, 1u);
    last = (t == total - 1);
    if (last) { A.ticket[b] = 0u; __threadfence(); }
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  double s = 0.0;
  for (int k = lane; k < A.npart; k += 32) s += __ldcg(part + k);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0 && !done) g.scal[(size_t)b * QD_S_COUNT + QD_S_ETA_NUM] = s;
}

// np.nan_to_num bookkeeping of the streaming kernel.  The reference cleans every intermediate field; on finite data that
// is the identity.  The FAST instantiation (CLEAN = false) therefore computes without the select chains and only tracks
// whether any value the reference would have cleaned was non-finite (one DSETP per value); if that ever happens in a
// warp, the warp re-runs its whole chunk with CLEAN = true (the outputs are rewritten; inputs are never modified by this
// kernel).  Identical bits either way.
template <bool CLEAN>
__device__ __forceinline__ double qd_of_cl(double x, bool& bad) {
  if (CLEAN) return qd_clean_sel(x);
  bad = bad || !(fabs(x) <= DBL_MAX);
  return x;
}

// One chunk of one warp: rows [j0, j1) of strip `strip`.  Returns the warp's partial eta sum in *contrib_out and whether a
// non-finite intermediate was seen (CLEAN = false only).
//
// Register windows are RINGS indexed by the row number modulo the ring length with compile-time indices: the loop is
// unrolled by 4 (= the longest ring) and every unrolled copy names its ring slots by literal index, so no value is ever
// moved between registers when the window advances.  With k = m - m_first and S = k & 3:
//   UB, VB [4]   cleaned post-momentum currents of rows m-4 .. m-1    (row r in slot (r - m_first) & 3)
//   EC     [4]   cleaned eta of rows m-4 .. m-1
//   ER     [4]   raw eta of rows m-2 .. m+1 (momentum uses rows m-1, m, m+1)
//   LU, LV, LE [4]  Laplacians of rows m-6 .. m-3
//   UP, VP [2]   post-del^4 currents of rows m-6, m-5;  E5 = post-del^4 eta of row m-5
template <bool CLEAN>
__device__ __forceinline__ bool qd_of_chunk(const QdGeo& g, const QdOcFusedArgs& A, const int b, const int lane, const int strip,
                                            const int j0, const int j1, const int src, double* s_in_w, double* s_sst_w, double* s_rc_w,
                                            double* contrib_out) {
  const int nlat = g.nlat, nlon = g.nlon;
  const size_t off = (size_t)b * g.ncell;
  const double* __restrict__ P = g.prm + (size_t)b * QD_P_COUNT;
  const double sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const double speed2_cap = P[QD_P_OC_SPEED2_CAP];
  const double* __restrict__ uo_in = src ? A.uo[1] : A.uo[0];
  const double* __restrict__ vo_in = src ? A.vo[1] : A.vo[0];
  double* __restrict__ uo_out = src ? A.uo[0] : A.uo[1];
  double* __restrict__ vo_out = src ? A.vo[0] : A.vo[1];
  // the per-row constant this lane stages (lanes 0 .. QD_RC_COUNT-1)
  const double* rc_src = nullptr;
  {
    const double* rows = g.rows;
    const double* cosh_r = rows + (size_t)QD_R_COS_ADV_HALF * nlat;
    const double* k4t = A.k4tab + (size_t)b * 3 * nlat;
    switch (lane) {
      case QD_RC_IACH: rc_src = rows + (size_t)QD_R_INV_ACOS_HALF * nlat; break;
      case QD_RC_FCOR: rc_src = rows + (size_t)QD_R_FCOR * nlat; break;
      case QD_RC_SPNG: rc_src = rows + (size_t)b * g.row_bstride + (size_t)QD_R_OC_SPONGE * nlat; break;
      case QD_RC_CAP: rc_src = cosh_r + 3 * (size_t)nlat; break;
      case QD_RC_CAM: rc_src = cosh_r + 4 * (size_t)nlat; break;
      case QD_RC_CBL: rc_src = cosh_r + 5 * (size_t)nlat; break;
      case QD_RC_K4U: rc_src = k4t; break;
      case QD_RC_K4V: rc_src = k4t + nlat; break;
      case QD_RC_K4E: rc_src = k4t + 2 * (size_t)nlat; break;
      case QD_RC_COS: rc_src = rows + (size_t)QD_R_COS * nlat; break;
      case QD_RC_IACC: rc_src = rows + (size_t)QD_R_INV_ACOS_CAP * nlat; break;
      case QD_RC_W: rc_src = rows + (size_t)QD_R_W * nlat; break;
      case QD_RC_COSH: rc_src = cosh_r; break;
      default: break;
    }
  }
  const double pG = P[QD_P_OC_G], irH = P[QD_P_OC_INV_RHO_H], rbot = P[QD_P_OC_R_BOT], pH = P[QD_P_OC_H];
  const double al = P[QD_P_OC_ADV_ALPHA], ucap = P[QD_P_OC_MAX_U];
  const bool mean4 = P[QD_P_OC_MEAN4] != 0.0;
  const double inner = sub_dt / 1.0;
  int gi = strip * QD_OF_COLS - QD_OF_HALO + lane;
  if (gi < 0) gi += nlon;
  if (gi >= nlon) gi -= nlon;
  if (gi >= nlon) gi -= nlon;
  const int icol = strip * QD_OF_COLS - QD_OF_HALO + lane;              // unwrapped column of this lane
  const double icd = (double)icol;
  const bool writer = lane >= QD_OF_HALO && lane < QD_OF_HALO + QD_OF_COLS && icol < nlon;
  double UB[4] = {0.0, 0.0, 0.0, 0.0}, VB[4] = {0.0, 0.0, 0.0, 0.0}, EC[4] = {0.0, 0.0, 0.0, 0.0}, ER[4] = {0.0, 0.0, 0.0, 0.0};
  double LU[4] = {0.0, 0.0, 0.0, 0.0}, LV[4] = {0.0, 0.0, 0.0, 0.0}, LE[4] = {0.0, 0.0, 0.0, 0.0};
  double UP[2] = {0.0, 0.0}, VP[2] = {0.0, 0.0}, E5 = 0.0;
  double contrib = 0.0;
  bool bad = false;
  const int m_first = j0 - 5;                                           // first momentum row: U'(j0-1) <- lap(j0-3) <- ub(j0-5); >= 3 since ja >= 8
  const int m_end = j1 + 5;
  unsigned landbits = 0u;                                               // land flags of rows m, m-1, ... (bit q = row m-q)
  size_t e = off + (size_t)m_first * nlon + gi;                         // element index of (row m, this lane's column)
  ER[3] = A.eta[e - nlon]; ER[0] = A.eta[e];                            // raw eta of rows m_first-1 (slot (k-1)&3), m_first (slot k&3)
  double jd = (double)(j0 - 10);                                        // (double)(m - 5), advanced by 1.0 per row (exact)
  // shared-memory rings of this warp as 32-bit shared addresses (computed once: no generic -> shared conversion per copy)
  const unsigned a_in = (unsigned)__cvta_generic_to_shared(s_in_w + lane);
  const unsigned a_sst = (unsigned)__cvta_generic_to_shared(s_sst_w + lane);
  const unsigned a_rc = (unsigned)__cvta_generic_to_shared(s_rc_w + lane);
  auto cp8 = [](unsigned saddr, const double* gptr) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gptr) : "memory"); };
  // stage row r = m_first + kk (element index er of this lane's column in that row) into its ring slots
  auto stage = [&](int r, int slot_m, int slot_s, size_t er) {
    const unsigned d = a_in + (unsigned)slot_m * (5 * 32 * 8);
    cp8(d, uo_in + er);
    cp8(d + 256, vo_in + er);
    cp8(d + 512, A.taux + er);
    cp8(d + 768, A.tauy + er);
    cp8(d + 1024, A.eta + er + nlon);
    cp8(a_sst + (unsigned)slot_s * (32 * 8), A.sst + er);
    if (lane < QD_RC_COUNT) cp8(a_rc + (unsigned)slot_s * (16 * 8), rc_src + r);
  };
  // row constants of the 6 rows below m_first (del^4 coefficients of rows m-2, m-4; cos of rows j+-1 of the first steps)
  if (lane < QD_RC_COUNT) {
#pragma unroll
    for (int q = 1; q <= 6; ++q) s_rc_w[((QD_OF_NS - q) & (QD_OF_NS - 1)) * 16 + lane] = rc_src[m_first - q >= 0 ? m_first - q : 0];
  }
  size_t es = e;                                                        // element index of the next row to stage (row m + D)
  unsigned char LDR[4] = {0, 0, 0, 0};                                  // land flags of rows m .. m+2 (ring, loaded D rows ahead like the staged fields)
#pragma unroll
  for (int d = 0; d < QD_OF_D; ++d) {
    if (m_first + d < m_end) { stage(m_first + d, d, d, es); LDR[d] = A.land[es]; }
    qd_cp_async_commit();
    es += nlon;
  }
  int m = m_first, k = 0;                                               // k = m - m_first
  // Slot of row k + S + d in the 16-slot rings, S and d literals: the group bases gm2, gm1, g0, gp1 (slots of rows k-8,
  // k-4, k, k+4; k is a multiple of 4 at the top of a group) are formed once per group, the rest is a compile-time offset.
#define QD_OF_SLOT(S, d) (((S) + (d)) >= 4 ? gp1 + ((S) + (d) - 4) : ((S) + (d)) >= 0 ? g0 + ((S) + (d)) : ((S) + (d)) >= -4 ? gm1 + ((S) + (d) + 4) : gm2 + ((S) + (d) + 8))
#define QD_OF_GROUP(GUARD)                                                                                        \
  {                                                                                                               \
    const int g0 = k & 12, gm1 = (k - 4) & 12, gm2 = (k - 8) & 12, gp1 = (k + 4) & 12;                            \
    QD_OF_STEP(0, GUARD) QD_OF_STEP(1, GUARD) QD_OF_STEP(2, GUARD) QD_OF_STEP(3, GUARD)                           \
  }
#define QD_OF_STEP(S, GUARD)                                                                                      \
  {                                                                                                               \
    const bool ld = LDR[(S) & 3] == 1;                                                                            \
    if (!(GUARD) || m + QD_OF_D < m_end) { stage(m + QD_OF_D, ((S) + QD_OF_D) & 3, QD_OF_SLOT(S, QD_OF_D), es); LDR[((S) + QD_OF_D) & 3] = A.land[es]; } \
    qd_cp_async_commit();                                                                                         \
    landbits = (landbits << 1) | (ld ? 1u : 0u);                                                                  \
    qd_cp_async_wait<QD_OF_D>();                         /* this lane's copies of row m have landed */            \
    __syncwarp();                                        /* ... and every other lane's (SST taps, row constants) */ \
    const double* in = s_in_w + ((S) & (QD_OF_NM - 1)) * (5 * 32) + lane;   /* k & 3 == S */                      \
    const double* rc0 = s_rc_w + QD_OF_SLOT(S, 0) * 16;                                                           \
    const double* rc2 = s_rc_w + QD_OF_SLOT(S, -2) * 16;                                                          \
    const double* rc4 = s_rc_w + QD_OF_SLOT(S, -4) * 16;                                                          \
    double uo = in[0], vo = in[32];                                                                               \
    const double tx = in[64], ty = in[96];                                                                        \
    const double er_p1 = in[128], er_0 = ER[(S) & 3], er_m1 = ER[((S) + 3) & 3];                                  \
    ER[((S) + 1) & 3] = er_p1;                                                                                    \
    /* ---- momentum at row m (ocean.py:306-336), raw eta */                                                      \
    {                                                                                                             \
      const double ee = __shfl_down_sync(0xffffffffu, er_0, 1), ew = __shfl_up_sync(0xffffffffu, er_0, 1);        \
      const double de_dl = (ee - ew) * g.inv_2dlon;                                                               \
      const double de_dp = (er_p1 - er_m1) * g.inv_2dlat;                                                         \
      const double gx = de_dl * rc0[QD_RC_IACH];                                                                  \
      const double gy = de_dp * g.inv_a;                                                                          \
      const double f = rc0[QD_RC_FCOR];                                                                           \
      const double du = (f * vo - pG * gx + tx * irH - rbot * uo);                                                \
      const double dv = (-f * uo - pG * gy + ty * irH - rbot * vo);                                               \
      uo = uo + sub_dt * du;                                                                                      \
      vo = vo + sub_dt * dv;                                                                                      \
      if (ld) { uo = 0.0; vo = 0.0; }                                                                             \
      const double rex = rc0[QD_RC_SPNG];                                                                         \
      uo = uo - sub_dt * rex * uo;                                                                                \
      vo = vo - sub_dt * rex * vo;                                                                                \
    }                                                                                                             \
    const double u0 = qd_of_cl<CLEAN>(uo, bad), v0 = qd_of_cl<CLEAN>(vo, bad), e0 = qd_of_cl<CLEAN>(er_0, bad);   \
    /* ---- Laplacians at row m-2 from rows m-4, m-2, m */                                                        \
    const double u4 = UB[(S) & 3], u2 = UB[((S) + 2) & 3], v4 = VB[(S) & 3], v2 = VB[((S) + 2) & 3];              \
    const double ec4 = EC[(S) & 3], ec2 = EC[((S) + 2) & 3];                                                      \
    double lu2, lv2, le2;                                                                                         \
    {                                                                                                             \
      const double ap = rc2[QD_RC_CAP], am = rc2[QD_RC_CAM], bl = rc2[QD_RC_CBL];                                 \
      const double ue = __shfl_down_sync(0xffffffffu, u2, 1), uw = __shfl_up_sync(0xffffffffu, u2, 1);            \
      lu2 = qd_of_cl<CLEAN>((ap * (u0 - u2) - am * (u2 - u4)) + bl * ((ue - 2.0 * u2) + uw), bad);                \
      const double ve = __shfl_down_sync(0xffffffffu, v2, 1), vw = __shfl_up_sync(0xffffffffu, v2, 1);            \
      lv2 = qd_of_cl<CLEAN>((ap * (v0 - v2) - am * (v2 - v4)) + bl * ((ve - 2.0 * v2) + vw), bad);                \
      const double ee = __shfl_down_sync(0xffffffffu, ec2, 1), ew = __shfl_up_sync(0xffffffffu, ec2, 1);          \
      le2 = qd_of_cl<CLEAN>((ap * (e0 - ec2) - am * (ec2 - ec4)) + bl * ((ee - 2.0 * ec2) + ew), bad);            \
    }                                                                                                             \
    /* ---- del^4 update at row m-4 from Laplacian rows m-6, m-4, m-2 */                                          \
    double U4, V4, E4;                                                                                            \
    {                                                                                                             \
      const double ap = rc4[QD_RC_CAP], am = rc4[QD_RC_CAM], bl = rc4[QD_RC_CBL];                                 \
      const double k4u = rc4[QD_RC_K4U], k4v = rc4[QD_RC_K4V], k4e = rc4[QD_RC_K4E];                              \
      const double lu4 = LU[(S) & 3], lu6 = LU[((S) + 2) & 3], lv4 = LV[(S) & 3], lv6 = LV[((S) + 2) & 3];        \
      const double le4 = LE[(S) & 3], le6 = LE[((S) + 2) & 3];                                                    \
      const double lue = __shfl_down_sync(0xffffffffu, lu4, 1), luw = __shfl_up_sync(0xffffffffu, lu4, 1);        \
      const double L2u = (ap * (lu2 - lu4) - am * (lu4 - lu6)) + bl * ((lue - 2.0 * lu4) + luw);                  \
      U4 = qd_of_cl<CLEAN>(u4 - k4u * L2u * inner, bad);                                                          \
      const double lve = __shfl_down_sync(0xffffffffu, lv4, 1), lvw = __shfl_up_sync(0xffffffffu, lv4, 1);        \
      const double L2v = (ap * (lv2 - lv4) - am * (lv4 - lv6)) + bl * ((lve - 2.0 * lv4) + lvw);                  \
      V4 = qd_of_cl<CLEAN>(v4 - k4v * L2v * inner, bad);                                                          \
      const double lee = __shfl_down_sync(0xffffffffu, le4, 1), lew = __shfl_up_sync(0xffffffffu, le4, 1);        \
      const double L2e = (ap * (le2 - le4) - am * (le4 - le6)) + bl * ((lee - 2.0 * le4) + lew);                  \
      E4 = qd_of_cl<CLEAN>(ec4 - k4e * L2e * inner, bad);                                                         \
    }                                                                                                             \
    /* ---- continuity, SST blend and current hygiene at row j = m-5 (post-del^4 rows j-1 = m-6, j, j+1 = m-4).     \
       Straight-line for every lane (the halo lanes compute harmless garbage, stores are predicated); the two rare   \
       cases -- a current above the speed cap, a departure point outside the fast gather's reach -- are repaired     \
       afterwards behind ONE warp-uniform branch, so that the four unrolled rows form long basic blocks. */         \
    const double U6 = UP[(S) & 1], U5 = UP[((S) + 1) & 1], V6 = VP[(S) & 1], V5 = VP[((S) + 1) & 1];              \
    const double Ue = __shfl_down_sync(0xffffffffu, U5, 1), Uw = __shfl_up_sync(0xffffffffu, U5, 1);              \
    {                                                                                                             \
      const int j = m - 5;                                                                                        \
      const bool emit = writer && (!(GUARD) || (j >= j0 && m < m_end));                                           \
      const double* rc5 = s_rc_w + QD_OF_SLOT(S, -5) * 16;                                                        \
      const double* rc6 = s_rc_w + QD_OF_SLOT(S, -6) * 16;                                                        \
      const size_t c = e - 5 * (size_t)nlon;                                                                      \
      const double du = (Ue - Uw) * g.inv_2dlon;                                                                  \
      const double dv = (V4 * rc4[QD_RC_COS] - V6 * rc6[QD_RC_COS]) * g.inv_2dlat;                                \
      const double div = rc5[QD_RC_IACC] * (du + dv);                                                             \
      const bool ldj = ((landbits >> 5) & 1u) != 0u;                                                              \
      double ev = E5 + (-sub_dt * pH * div);                                                                      \
      if (ldj) ev = 0.0;                                                                                          \
      const double s2 = U5 * U5 + V5 * V5;                                                                        \
      const bool over = emit && (s2 >= speed2_cap);      /* == sqrt(s2) > cap, engine.py:speed2_threshold */      \
      /* departure point (ocean.py:380, dynamics.py:104-115 form): exact quotients by reciprocal + residual FMAs */ \
      const double cosj = rc5[QD_RC_COSH];                                                                        \
      const double ddx = qd_div_exact(qd_div_exact(U5 * sub_dt, g.a * cosj, rc5[QD_RC_IACH]), g.dlon, g.inv_dlon); \
      const double ddy = qd_div_exact(qd_div_exact(V5 * sub_dt, g.a, g.inv_a), g.dlat, g.inv_dlat);               \
      const double y = jd - ddy, x = icd - ddx;                                                                   \
      const double own = s_sst_w[QD_OF_SLOT(S, -5) * 32 + lane];                                                  \
      /* fast gather: |displacement| < 1 cell, no wrap (scipy's 'wrap' has period n-1, unlike the stencils): the   \
         four taps sit in SST ring rows j-1 .. j+1 and in this warp's columns.  floor(y) = j-1 or j. */           \
      const bool yin = (y >= jd - 1.0) && (y < jd + 1.0), xin = (x >= icd - 1.0) && (x < icd + 1.0);              \
      const bool ylo = y < jd, xlo = x < icd;                                                                     \
      const bool fast = yin && xin && x >= 0.0 && icol + (xlo ? -1 : 0) <= nlon - 2;                              \
      const double fy = ylo ? jd - 1.0 : jd, fx = xlo ? icd - 1.0 : icd;                                          \
      const double wy0 = 1.0 - (y - fy), wx0 = 1.0 - (x - fx);                                                    \
      const double wy1 = 1.0 - wy0, wx1 = 1.0 - wx0;                                                              \
      const int l0 = (xlo && lane > 0) ? lane - 1 : lane;        /* clamped for the halo lanes (never stored) */  \
      const double* r0 = s_sst_w + (ylo ? QD_OF_SLOT(S, -6) : QD_OF_SLOT(S, -5)) * 32 + l0;                       \
      const double* r1 = s_sst_w + (ylo ? QD_OF_SLOT(S, -5) : QD_OF_SLOT(S, -4)) * 32 + l0;                       \
      double adv = r0[0] * wy0 * wx0;                                                                             \
      adv = adv + r0[1] * wy0 * wx1;                                                                              \
      adv = adv + r1[0] * wy1 * wx0;                                                                              \
      adv = adv + r1[1] * wy1 * wx1;                                                                              \
      const bool slow = emit && !fast;                                                                            \
      if (emit) {                                                                                                 \
        A.eta_out[c] = ev;                                                                                        \
        if (qd_owned(g, j)) contrib += ev * (rc5[QD_RC_W] * (ldj ? 0.0 : 1.0));                                   \
        A.tb[c] = (1.0 - al) * own + al * adv;                                                                    \
        uo_out[c] = U5;                                  /* nan_to_num of a cleaned value is the value */         \
        vo_out[c] = V5;                                                                                           \
      }                                                                                                           \
      if (__any_sync(0xffffffffu, over || slow)) {       /* rare repairs */                                       \
        const double Ve = __shfl_down_sync(0xffffffffu, V5, 1), Vw = __shfl_up_sync(0xffffffffu, V5, 1);          \
        if (over) {                                      /* ocean.py:408-434 */                                   \
          double uo2 = U5, vo2 = V5;                                                                              \
          if (mean4) {                                                                                            \
            uo2 = 0.25 * (U4 + U6 + Ue + Uw);                                                                     \
            vo2 = 0.25 * (V4 + V6 + Ve + Vw);                                                                     \
            const double sp2 = sqrt(uo2 * uo2 + vo2 * vo2);                                                       \
            const double sc2 = (sp2 > ucap) ? ucap / (sp2 + 1e-12) : 1.0;                                         \
            uo2 = uo2 * sc2;                                                                                      \
            vo2 = vo2 * sc2;                                                                                      \
          } else {                                                                                                \
            const double sc1 = ucap / (sqrt(s2) + 1e-12);                                                         \
            uo2 = uo2 * sc1;                                                                                      \
            vo2 = vo2 * sc1;                                                                                      \
          }                                                                                                       \
          uo_out[c] = uo2;                                                                                        \
          vo_out[c] = vo2;                                                                                        \
        }                                                                                                         \
        if (slow) A.tb[c] = (1.0 - al) * own + al * qd_bilinear_wrap(A.sst + off, nlat, nlon, y, x);              \
      }                                                                                                           \
    }                                                                                                             \
    /* ---- advance the rings: each new value replaces the oldest row of its ring */                              \
    UB[(S) & 3] = u0; VB[(S) & 3] = v0; EC[(S) & 3] = e0;                                                         \
    LU[((S) + 2) & 3] = lu2; LV[((S) + 2) & 3] = lv2; LE[((S) + 2) & 3] = le2;                                    \
    UP[(S) & 1] = U4; VP[(S) & 1] = V4; E5 = E4;                                                                  \
    e += nlon; es += nlon; jd += 1.0; ++m;                                                                        \
    __syncwarp();                                        /* ring slots staged next step were read by neighbour lanes in this one */ \
  }
  // warm-up groups (rows before the first continuity row j0), steady state without any guard, guarded tail; the last
  // group may run up to 3 rows past m_end: nothing is staged, loaded from global or stored for them
  while (m < j0 + 5) { QD_OF_GROUP(true) k += 4; }
  while (m + 3 + QD_OF_D < m_end) { QD_OF_GROUP(false) k += 4; }
  while (m < m_end) { QD_OF_GROUP(true) k += 4; }
#undef QD_OF_STEP
#undef QD_OF_GROUP
#undef QD_OF_SLOT
  qd_cp_async_wait<0>();
  *contrib_out = contrib;
  return bad;
}

__global__ void __launch_bounds__(32 * QD_OF_WARPS, 3) k_ocean_fused(QdGeo g, QdOcFusedArgs A, QdSubCtl sc) {
  __shared__ double s_in[QD_OF_WARPS][QD_OF_NM * 5 * 32];
  __shared__ double s_sst[QD_OF_WARPS][QD_OF_NS * 32];
  __shared__ double s_rc[QD_OF_WARPS][QD_OF_NS * 16];
  const int b = blockIdx.y;
  const bool done = qd_sub_done(g, b, sc);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nstrips = (g.nlon + QD_OF_COLS - 1) / QD_OF_COLS;
  const int w = blockIdx.x * QD_OF_WARPS + wib;
  const int chunk = w / nstrips, strip = w - chunk * nstrips;
  const int j0 = A.ja + chunk * A.R;
  if (done || j0 >= A.jb) { qd_of_finish(g, A, b, w, lane, 0.0, done); return; }
  const int j1 = min(j0 + A.R, A.jb);
  int src, copy_back;
  qd_oc_parity(g, b, sc, &src, &copy_back);
  double contrib = 0.0;
  const bool bad = qd_of_chunk<false>(g, A, b, lane, strip, j0, j1, src, s_in[wib], s_sst[wib], s_rc[wib], &contrib);
  if (__any_sync(0xffffffffu, bad)) {                                   // a non-finite intermediate: redo the chunk with np.nan_to_num applied
    __syncwarp();
    qd_of_chunk<true>(g, A, b, lane, strip, j0, j1, src, s_in[wib], s_sst[wib], s_rc[wib], &contrib);
  }
  qd_of_finish(g, A, b, w, lane, contrib, false);
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
      const double s2 = uo * uo + vo * vo;
      if (s2 >= P[QD_P_OC_SPEED2_CAP]) {
        const double speed = sqrt(s2);
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
      (src ? A.uo[0] : A.uo[1])[c] = uo;
      (src ? A.vo[0] : A.vo[1])[c] = vo;
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
