// qd_ops.cuh -- numeric operators of the Qingdai step as sm_100a kernels.
//
// Layout: every field is float64 [B][nlat][nlon]; a thread owns one cell, threads run along
// longitude (linear cell index) so that every warp touches consecutive addresses whatever
// n_lon is; blockIdx.y is the ensemble member.  These kernels are HBM/L2-bandwidth work:
// no tensor cores, no reshaping into GEMMs.  Operand order follows the reference's NumPy
// expressions exactly (compile with -fmad=false) so fields agree to the last few ulp.
#pragma once
#include "qd_rt.h"
#include "qd_math.cuh"
#include "../../include/qd_b200.h"

#define QD_THREADS 256
#define QD_MAX_FIELDS 6
#define QD_GAUSS_MAXR 8
#define QD_SIGMA_SB 5.670374e-8   /* constants.py:10 */

// A divisor with its correctly rounded reciprocal (qd_div_u below), or a ready quotient in .r
struct alignas(16) QdRcp { double b, r; };
// Per-member table of the parameter-only divisors and quotients of the physics kernels, formed on the host whenever
// the parameters or dt change (qd_derive in qd_api.cu).  The expressions are the reference's, operand for operand.
enum QdUdivId {
  QD_U_MCOL = 0,      // max(1e-6, rho_a h_mbl)                humidity.py:160-176
  QD_U_TAU_COND,      // max(1e-6, tau_cond)
  QD_U_RHO_SNOW,      // max(rho_snow, 1e-6)                   run_simulation.py:1962
  QD_U_1000,          // 1000 (m per km of the lapse rate)
  QD_U_SNOW_BAND,     // max(1e-6, snow_t_band)                hydrology.py:121
  QD_U_DT,            // dt
  QD_U_SWE_REF,       // max(1e-6, swe_ref)
  QD_U_SIGMA,         // Stefan-Boltzmann constant
  QD_U_RUNOFF_TAU,    // max(1, runoff_tau_days 86400)         hydrology.py:244
  QD_U_C_SFC,         // max(1e-12, c_sfc)                     dynamics.py:316
  QD_U_TAU_RAD,       // tau_rad
  QD_U_HICE_REF,      // max(1e-6, H_ice_ref)
  QD_U_RHO_I_LF,      // rho_i L_f                             energy.py:340-372
  QD_U_CS_LAND, QD_U_CS_ICE, QD_U_CS_OCEAN,   // heat capacities after the 1e3 sanitiser, energy.py:395-399
  QD_U_ATM,           // max(1e-6, rho_a) max(1, H_atm) g      energy.py:476-480
  QD_U_12K,           // 12 K                                  physics.py:94
  QD_U_ADV_REF,       // 2e-5 K/s                              physics.py:107
  QD_U_2DY,           // 2 (dlat a)                            physics.py:103
  QD_U_NDIV,
  QD_Q_G_CP = QD_U_NDIV,   // .r = g / 1004
  QD_Q_R_G,                // .r = 287 / g
  QD_Q_DDF,                // .r = snow_ddf / 86400
  QD_Q_MELT,               // .r = snow_melt_rate / 86400
  QD_U_COUNT
};

struct QdGeo {
  int nlat, nlon, ncell, batch;
  double a, dlat, dlon, a_sq, dlon_sq;
  double inv_dlat, inv_2dlat, inv_dlon_sq, inv_a_sq;      // reciprocals used by the stencil kernels
  double inv_2dlon, inv_a, inv_dlon;
  const double* rows;    // [B][QD_R_COUNT + 3*QD_NUSER_ROWS][nlat]; member 0's copy serves the geometry rows
  long long row_bstride; // doubles between two members' tables
  const double* cols;    // [QD_C_COUNT][nlon]
  const double* prm;     // [B][QD_P_COUNT]
  double* scal;          // [B][QD_S_COUNT]
  const QdRcp* udiv;     // [B][QD_U_COUNT]
  // Latitude-band decomposition (qd_band.cuh).  Every rank stores full-size fields but computes only the
  // rows of up to two segments [sa0,sa1) u [sb0,sb1) (its own rows [own0,own1) widened by the halo the
  // inputs allow; the second segment is the part that wraps over a pole).  One rank: sa = own = [0, nlat).
  int own0, own1, sa0, sa1, sb0, sb1, ncomp;
  int nvb;                       // virtual blocks per member of the sum reductions (QD_VB_LOOP): a function of the grid size only
  unsigned long long div_nlon;   // floor(2^40 / nlon) + 1: t / nlon == (t * div_nlon) >> 40 for t * nlon < 2^40
};
QD_HD int qd_div_nlon(const QdGeo& g, int t) { return (int)(((unsigned long long)(unsigned)t * g.div_nlon) >> 40); }

struct QdFields {          // up to QD_MAX_FIELDS field pointers passed by value
  int n;
  const double* src[QD_MAX_FIELDS];
  double* dst[QD_MAX_FIELDS];
  const double* aux[QD_MAX_FIELDS];   // per-field row table (k4 rows) or second input
  double scale[QD_MAX_FIELDS];
};

struct QdGaussW { int r; int wrap; double w[2 * QD_GAUSS_MAXR + 1]; };

// row r of the compute region -> global row
QD_HD int qd_seg_row(const QdGeo& g, int r) { const int na = g.sa1 - g.sa0; return r < na ? g.sa0 + r : g.sb0 + (r - na); }
QD_HD bool qd_owned(const QdGeo& g, int j) { return j >= g.own0 && j < g.own1; }

#define QD_CELL_PROLOGUE(geo)                                         \
  const int b = blockIdx.y;                                           \
  const int t_ = blockIdx.x * blockDim.x + threadIdx.x;               \
  const bool active = t_ < (geo).ncomp;                               \
  const int r_ = active ? qd_div_nlon((geo), t_) : 0;                 \
  const int i = active ? t_ - r_ * (geo).nlon : 0;                    \
  const int j = active ? qd_seg_row((geo), r_) : 0;                   \
  const int idx = j * (geo).nlon + i;                                 \
  const size_t off = (size_t)b * (geo).ncell;                         \
  (void)i; (void)j; (void)off; (void)idx;

// Grid-stride form for kernels that end in a grid-wide reduction: the grid is capped at a few blocks per SM
// (QD_KR in qd_api.cu), so the "last block" ticket -- one atomic round trip on every block's critical path,
// ~40 us per launch at 1441x2880 with one block per 256 cells (profiles/r01_ncu_hires_step.md) -- is paid
// by ~1e3 blocks instead of ~1.6e4.  Partials stay per block and are combined in block order: deterministic.
#define QD_CELL_LOOP(geo)                                             \
  const int b = blockIdx.y;                                           \
  const size_t off = (size_t)b * (geo).ncell;                         \
  (void)off;                                                          \
  _Pragma("unroll 2")                                                 \
  for (int t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < (geo).ncomp; t_ += gridDim.x * blockDim.x)
// same loop over the first `limit` cells of the compute region, unrolled twice so that the loads of two cells overlap
#define QD_CELL_LOOP_N(geo, limit)                                    \
  const int b = blockIdx.y;                                           \
  const size_t off = (size_t)b * (geo).ncell;                         \
  (void)off;                                                          \
  _Pragma("unroll 2")                                                 \
  for (int t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < (limit); t_ += gridDim.x * blockDim.x)
// Sums whose value feeds back into the fields (precipitation renormalisation, the ocean's eta mean, diagnostics) are
// formed per VIRTUAL block: geo.nvb of them per member, a function of the grid size only.  A physical block takes
// virtual blocks vb = blockIdx.x, blockIdx.x + gridDim.x, ...; virtual block vb owns cells vb*256 + t + k*nvb*256.
// Partials are stored per virtual block and combined in that order, so the result does not depend on how many
// physical blocks the launch has -- i.e. not on how many ensemble members share the GPU (B = 8 gives the bits of B = 1).
#define QD_VB_LOOP(geo)                                               \
  const int b = blockIdx.y;                                           \
  const size_t off = (size_t)b * (geo).ncell;                         \
  (void)off;                                                          \
  for (int vb_ = blockIdx.x; vb_ < (geo).nvb; vb_ += gridDim.x)
#define QD_VB_CELLS(geo, limit)                                       \
  _Pragma("unroll 2")                                                 \
  for (int t_ = vb_ * blockDim.x + threadIdx.x; t_ < (limit); t_ += (geo).nvb * blockDim.x)
#define QD_CELL_JI(geo)                                               \
  const int r_ = qd_div_nlon((geo), t_); const int i = t_ - r_ * (geo).nlon; \
  const int j = qd_seg_row((geo), r_); const int idx = j * (geo).nlon + i; (void)i; (void)j; (void)idx;

QD_HD const double* qd_row(const QdGeo& g, int id) { return g.rows + (size_t)id * g.nlat; }
// rows that follow a member's parameters (K4, ocean sponge, polar flag): member b's copy of the table
QD_HD const double* qd_mrow(const QdGeo& g, int id, int b) { return g.rows + (size_t)b * g.row_bstride + (size_t)id * g.nlat; }
QD_HD double qd_prm(const QdGeo& g, int b, int id) { return g.prm[(size_t)b * QD_P_COUNT + id]; }

// ------------------------------------------------------------------------------ Laplacian
// dynamics.py:144-173 / ocean.py:100-117: (1/a^2)[(1/c) d_phi(c d_phi F) + d2_lambda F / c^2],
// np.gradient in phi (one-sided at rows 0, n-1), periodic-n 3-point in lambda, nan_to_num(F) first.
struct QdCleanLoad {
  const double* p; int nlon;
  QD_HD double operator()(int jj, int ii) const { return qd_nan_to_num(p[(size_t)jj * nlon + ii]); }
};

// `c` points at a cosine row table that is followed by its 1/c and 1/c^2 rows (QD_R_*COS* layout).
template <class Acc>
QD_HD double qd_lap_cell(const Acc& F, int j, int i, const QdGeo& g, const double* c) {
  const int nlat = g.nlat, nlon = g.nlon;
  const double* ic = c + nlat;
  const double* ic2 = c + 2 * nlat;
  if (j >= 2 && j <= nlat - 3) {
    // every np.gradient involved is centred: the same operations as the general form below, without its
    // per-row case analysis (bit-identical)
    const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
    const double f0 = F(j, i);
    const double gp = (F(j + 2, i) - f0) * g.inv_2dlat, gm = (f0 - F(j - 2, i)) * g.inv_2dlat;
    const double gphi = (c[j + 1] * gp - c[j - 1] * gm) * g.inv_2dlat;
    const double term_phi = ic[j] * gphi;
    const double d2 = ((F(j, ip) - 2.0 * f0) + F(j, im)) * g.inv_dlon_sq;
    return (term_phi + d2 * ic2[j]) * g.inv_a_sq;
  }
  const int jm = j > 0 ? j - 1 : 0, jp = j < nlat - 1 ? j + 1 : nlat - 1;
  // G(jj): np.gradient(F, dphi, axis=0) at row jj (edge_order=1)
  auto G = [&](int jj) -> double {
    if (jj == 0) return (F(1, i) - F(0, i)) * g.inv_dlat;
    if (jj == nlat - 1) return (F(nlat - 1, i) - F(nlat - 2, i)) * g.inv_dlat;
    return (F(jj + 1, i) - F(jj - 1, i)) * g.inv_2dlat;
  };
  double gphi;
  if (j == 0) gphi = (c[1] * G(1) - c[0] * G(0)) * g.inv_dlat;
  else if (j == nlat - 1) gphi = (c[j] * G(j) - c[j - 1] * G(j - 1)) * g.inv_dlat;
  else gphi = (c[jp] * G(jp) - c[jm] * G(jm)) * g.inv_2dlat;
  const double term_phi = ic[j] * gphi;
  const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
  const double d2 = ((F(j, ip) - 2.0 * F(j, i)) + F(j, im)) * g.inv_dlon_sq;
  const double term_lam = d2 * ic2[j];
  return (term_phi + term_lam) * g.inv_a_sq;
}

// out[k] = lap(src[k]) for k < n fields; aux[0] = cosine row table (floored by the caller).
__global__ void __launch_bounds__(QD_THREADS) k_laplacian(QdGeo g, QdFields f, const double* cosr) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  for (int k = 0; k < f.n; ++k) {
    QdCleanLoad F{f.src[k] + off, g.nlon};
    f.dst[k][off + idx] = qd_lap_cell(F, j, i, g, cosr);
  }
}

// F <- nan_to_num(nan_to_num(F) - k4*lap(L)*sub_dt)   (dynamics.py:205-212).  src = L, dst = F (in
// place), aux = k4 row table, scale = multiplier applied to the table (1, 0.5, 0.25 are exact).
// If k4_over_dt >= 0 the coefficient is (aux*scale)/k4_over_dt... see ocean: k4 = s4dx4 / sub_dt.
__global__ void __launch_bounds__(QD_THREADS) k_hyper_update(QdGeo g, QdFields f, const double* cosr,
                                                            double sub_dt, double k4_div) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  for (int k = 0; k < f.n; ++k) {
    QdCleanLoad L{f.src[k] + off, g.nlon};
    const double L2 = qd_lap_cell(L, j, i, g, cosr);
    double k4 = f.aux[k][j];
    if (k4_div > 0.0) k4 = k4 / k4_div;        // ocean.py:347  sigma4*dx^4 / max(1e-12, sub_dt)
    k4 = f.scale[k] * k4;                      // 0.5*k4_map etc. (dynamics.py:568-570, ocean.py:352)
    const double cur = qd_nan_to_num(f.dst[k][off + idx]);
    f.dst[k][off + idx] = qd_nan_to_num(cur - k4 * L2 * sub_dt);
  }
}

// ------------------------------------------------------------------------------ semi-Lagrangian gather
// scipy map_coordinates(order=1, mode='wrap') semantics (see oracle/ops.py:bilinear_wrap):
// legacy wrap with period n-1 on BOTH axes, weights w0 = 1-t, w1 = 1-w0, fixed accumulation order.
QD_HD double qd_wrap_coord(double x, int n) {
  if (n <= 1) return 0.0;
  const double sz = (double)(n - 1);
  if (!(fabs(x) < 1.0e15)) return 0.0;          // NaN / absurd departure points: reference is UB here
  if (x < 0.0) x += sz * ((double)(long long)(-x / sz) + 1.0);
  else if (x > sz) x -= sz * (double)(long long)(x / sz);
  return x;
}
QD_HD int qd_wrap_next(int k, int n) {           // index k+1 reduced like scipy does for the 2nd tap
  const int k1 = k + 1;
  if (n <= 1) return 0;
  if (k1 > n - 1) return k1 - (n - 1) * (k1 / (n - 1));
  return k1;
}
QD_HD double qd_bilinear_wrap(const double* F, int nlat, int nlon, double y, double x) {
  y = qd_wrap_coord(y, nlat);
  x = qd_wrap_coord(x, nlon);
  const double fy = floor(y), fx = floor(x);
  int j0 = (int)fy, i0 = (int)fx;
  if (j0 > nlat - 1) j0 = nlat - 1;
  if (i0 > nlon - 1) i0 = nlon - 1;
  const int j1 = qd_wrap_next(j0, nlat), i1 = qd_wrap_next(i0, nlon);
  const double wy0 = 1.0 - (y - fy), wx0 = 1.0 - (x - fx);
  const double wy1 = 1.0 - wy0, wx1 = 1.0 - wx0;
  const double* r0 = F + (size_t)j0 * nlon;
  const double* r1 = F + (size_t)j1 * nlon;
  double t = r0[i0] * wy0 * wx0;
  t = t + r0[i1] * wy0 * wx1;
  t = t + r1[i0] * wy1 * wx0;
  t = t + r1[i1] * wy1 * wx1;
  return t;
}
// x / b with y = RN(1 / b) from the host: reciprocal multiply + two exact-residual corrections.  By Markstein's
// theorem the result is the correctly rounded quotient, i.e. the same bits as the IEEE division the reference
// performs, in 5 instructions instead of ~25 (the four divisions of a departure point were 45 % of the gather
// kernels' instructions).  Zero / NaN / inf quotients only occur for winds that the wrap below maps to 0 anyway.
#if !QD_EMU
__device__ __forceinline__ double qd_div_exact(double x, double b, double y) {
  double q = x * y;
  double r = __fma_rn(-b, q, x);
  q = __fma_rn(r, y, q);
  r = __fma_rn(-b, q, x);
  return __fma_rn(r, y, q);
}
#endif
// ---- divisions by block-uniform divisors
// nvcc's inline fp64 division is ~25 instructions and sends the whole warp through a 64-instruction out-of-line slow
// path whenever a numerator is zero -- dry cells, open ocean, the night side: 7.5 slow-path calls per warp in k_column,
// a quarter of its instructions (profiles/README.md).  For divisors that are the same for every cell of a block
// (parameters, dt) the host forms RN(1/b) once (QdGeo::udiv); the cell code then gets the correctly rounded quotient
// from qd_div_exact.  Outside the range where its residuals are exact it falls back to
// an exact shortcut (zero numerator) or to the plain division, so every path returns the bits of x / b.
QD_HD QdRcp qd_rcp(double b) {
  QdRcp d;
  d.b = b;
  d.r = (fabs(b) > 1e-100 && fabs(b) < 1e100) ? 1.0 / b : NAN;      // NaN: always take the plain division
  return d;
}
QD_HD double qd_div_u(double x, const QdRcp& d) {
#if !QD_EMU && defined(__CUDA_ARCH__)
  const double q = x * d.r;
  const unsigned e = ((unsigned)__double2hiint(q) >> 20) & 0x7ffu;  // biased exponent of the first estimate
  if (e - 523u < 1000u) {                                           // 2^-500 <= |q| < 2^500: residuals are exact
    double r = __fma_rn(-d.b, q, x);
    const double q1 = __fma_rn(r, d.r, q);
    r = __fma_rn(-d.b, q1, x);
    return __fma_rn(r, d.r, q1);
  }
  if (x == 0.0 && d.r == d.r) return q;                             // +-0 with the sign of x / b
  return x / d.b;
#else
  return x / d.b;
#endif
}
// x / b for a per-cell divisor: only skips the slow path of a zero numerator (0 / b = +-0 for any non-zero, non-NaN b)
QD_HD double qd_div_z(double x, double b) {
#if !QD_EMU && defined(__CUDA_ARCH__)
  if (x == 0.0 && b == b && b != 0.0)
    return __hiloint2double((__double2hiint(x) ^ __double2hiint(b)) & 0x80000000, 0);
#endif
  return x / b;
}
// departure point of cell (j,i): dynamics.py:104-115.  iac = RN(1 / (a cosj)) from the row tables, or 0 when the
// caller has no such table (then the quotients are plain divisions).
QD_HD void qd_departure(double u, double v, double dt, const QdGeo& g, double cosj, double iac,
                        int j, int i, double* y, double* x) {
#if !QD_EMU && defined(__CUDA_ARCH__)
  if (iac != 0.0) {
    const double dx = qd_div_exact(qd_div_exact(u * dt, g.a * cosj, iac), g.dlon, g.inv_dlon);
    const double dy = qd_div_exact(qd_div_exact(v * dt, g.a, g.inv_a), g.dlat, g.inv_dlat);
    *y = (double)j - dy;
    *x = (double)i - dx;
    return;
  }
#endif
  (void)iac;
  const double dx = (u * dt / (g.a * cosj)) / g.dlon;
  const double dy = (v * dt / g.a) / g.dlat;
  *y = (double)j - dy;
  *x = (double)i - dx;
}

#ifndef QD_LB_ADVECT
#define QD_LB_ADVECT 7
#endif
__global__ void __launch_bounds__(QD_THREADS, QD_LB_ADVECT) k_advect(QdGeo g, QdFields f, const double* u, const double* v,
                                                      double dt, const double* cosr, const double* iacr) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  double y, x;
  qd_departure(u[off + idx], v[off + idx], dt, g, cosr[j], iacr ? iacr[j] : 0.0, j, i, &y, &x);
  for (int k = 0; k < f.n; ++k)
    f.dst[k][off + idx] = qd_bilinear_wrap(f.src[k] + off, g.nlat, g.nlon, y, x);
}

// ------------------------------------------------------------------------------ Shapiro 1-2-1
// dynamics.py:215-231: lon pass periodic (period n_lon), lat pass edge-replicate; scipy's
// accumulation order (x[k-1]*.25 + x[k]*.5) + x[k+1]*.25.  clean=1 applies nan_to_num on load.
__global__ void __launch_bounds__(QD_THREADS) k_shapiro_lon(QdGeo g, QdFields f, int clean) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const int ip = i + 1 < g.nlon ? i + 1 : 0, im = i > 0 ? i - 1 : g.nlon - 1;
  for (int k = 0; k < f.n; ++k) {
    const double* r = f.src[k] + off + (size_t)j * g.nlon;
    double xm = r[im], x0 = r[i], xp = r[ip];
    if (clean) { xm = qd_nan_to_num(xm); x0 = qd_nan_to_num(x0); xp = qd_nan_to_num(xp); }
    f.dst[k][off + idx] = (xm * 0.25 + x0 * 0.5) + xp * 0.25;
  }
}
__global__ void __launch_bounds__(QD_THREADS) k_shapiro_lat(QdGeo g, QdFields f) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  const int jm = j > 0 ? j - 1 : 0, jp = j < g.nlat - 1 ? j + 1 : g.nlat - 1;
  for (int k = 0; k < f.n; ++k) {
    const double* p = f.src[k] + off;
    f.dst[k][off + idx] = (p[(size_t)jm * g.nlon + i] * 0.25 + p[idx] * 0.5) + p[(size_t)jp * g.nlon + i] * 0.25;
  }
}

// ------------------------------------------------------------------------------ Gaussian (separable)
// scipy gaussian_filter1d -> NI_Correlate1D symmetric branch: centre*w0, then pairs from the outermost
// tap inwards.  reflect = (d c b a | a b c d | d c b a); wrap = period n.
QD_HD int qd_extend(int k, int n, int wrap) {
  if (wrap) { k %= n; return k < 0 ? k + n : k; }
  const int p = 2 * n;
  k %= p; if (k < 0) k += p;
  return k >= n ? p - 1 - k : k;
}
template <class Load>
QD_HD double qd_gauss_tap(const Load& E, int c, int n, const QdGaussW& w) {
  double out = E(c) * w.w[w.r];
  if (c - w.r >= 0 && c + w.r < n) {           // no boundary extension: skip the modulo arithmetic (same sums, same order)
    for (int jj = -w.r; jj < 0; ++jj) out = out + (E(c + jj) + E(c - jj)) * w.w[w.r + jj];
    return out;
  }
  for (int jj = -w.r; jj < 0; ++jj)
    out = out + (E(qd_extend(c + jj, n, w.wrap)) + E(qd_extend(c - jj, n, w.wrap))) * w.w[w.r + jj];
  return out;
}
__global__ void __launch_bounds__(QD_THREADS) k_gauss_lat(QdGeo g, QdFields f, QdGaussW w) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  for (int k = 0; k < f.n; ++k) {
    const double* p = f.src[k] + off + i;
    const int nlon = g.nlon;
    auto E = [&](int jj) -> double { return p[(size_t)jj * nlon]; };
    f.dst[k][off + idx] = qd_gauss_tap(E, j, g.nlat, w);
  }
}
__global__ void __launch_bounds__(QD_THREADS) k_gauss_lon(QdGeo g, QdFields f, QdGaussW w) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  for (int k = 0; k < f.n; ++k) {
    const double* p = f.src[k] + off + (size_t)j * g.nlon;
    auto E = [&](int ii) -> double { return p[ii]; };
    f.dst[k][off + idx] = qd_gauss_tap(E, i, g.nlon, w);
  }
}

// ------------------------------------------------------------------------------ divergence / vorticity
// grid.py:41-88: np.roll centred differences; the phi-term is zeroed on rows 0 and n-1.  The three
// divisions (/(2 dlon), /(2 dphi), the 1/(a cos) factor) are reciprocal multiplies (<= 1 ulp each).
QD_HD double qd_div_cell(const double* u, const double* v, int j, int i, const QdGeo& g) {
  const int nlon = g.nlon, nlat = g.nlat;
  const double* cosr = qd_row(g, QD_R_COS);
  const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
  const double du = (u[(size_t)j * nlon + ip] - u[(size_t)j * nlon + im]) * g.inv_2dlon;
  double dv = 0.0;
  if (j > 0 && j < nlat - 1)
    dv = (v[(size_t)(j + 1) * nlon + i] * cosr[j + 1] - v[(size_t)(j - 1) * nlon + i] * cosr[j - 1]) * g.inv_2dlat;
  return qd_row(g, QD_R_INV_ACOS_CAP)[j] * (du + dv);
}
QD_HD double qd_vort_cell(const double* u, const double* v, int j, int i, const QdGeo& g) {
  const int nlon = g.nlon, nlat = g.nlat;
  const double* cosr = qd_row(g, QD_R_COS);
  const int ip = i + 1 < nlon ? i + 1 : 0, im = i > 0 ? i - 1 : nlon - 1;
  const double dv = (v[(size_t)j * nlon + ip] - v[(size_t)j * nlon + im]) * g.inv_2dlon;
  double du = 0.0;
  if (j > 0 && j < nlat - 1)
    du = (u[(size_t)(j + 1) * nlon + i] * cosr[j + 1] - u[(size_t)(j - 1) * nlon + i] * cosr[j - 1]) * g.inv_2dlat;
  return qd_row(g, QD_R_INV_ACOS_CAP)[j] * (dv - du);
}
__global__ void __launch_bounds__(QD_THREADS) k_divvort(QdGeo g, const double* u, const double* v, double* out, int vort) {
  QD_CELL_PROLOGUE(g)
  if (!active) return;
  out[off + idx] = vort ? qd_vort_cell(u + off, v + off, j, i, g) : qd_div_cell(u + off, v + off, j, i, g);
}

// ------------------------------------------------------------------------------ zonal band-stop
// dynamics.py:233-258: rfft -> bins >= kcut scaled by (1-damp) -> irfft.  Row-local equivalent used
// here: y = x - d * HP(x), HP = projection on the stop band kcut..kN, evaluated as a real DFT
// restricted to that band (n_lon*n_stop MACs forward and back per row); d = 1 - max(0, 1-min(1,damp)).
// One block per (row, member).  twid = cos|sin(2 pi m / n) [2][nlon] from the host.  re/im (global scratch) are only
// read back by the block that wrote them.
__global__ void __launch_bounds__(QD_THREADS) k_zonal_bandstop(QdGeo g, double* f, const double* twid,
                                                              int kcut, double d, double* coef, double* outbuf) {
  const int b = blockIdx.y, j = qd_seg_row(g, blockIdx.x), n = g.nlon;      // one block per row of the compute region
  double* row = f + (size_t)b * g.ncell + (size_t)j * n;
  const int kN = n / 2;                                   // rfft bins = n/2 + 1
  const int nstop = kN - kcut + 1;
  double* re = coef + ((size_t)b * g.nlat + j) * 2 * (size_t)(kN + 1);
  double* im = re + (kN + 1);
#if !QD_EMU
  // Shared-memory real DFT restricted to the stop band: the cleaned row and the cos / sin tables are staged once
  // (3 n doubles, 69 KB at n = 2880); the phase index (k m) mod n advances by addition instead of a 64-bit modulo.
  extern __shared__ double smem[];
  double* xs = smem;                 // [n] nan_to_num(row)
  double* ct = smem + n;             // [n]
  double* st = smem + 2 * (size_t)n; // [n]
  for (int m = threadIdx.x; m < n; m += blockDim.x) { xs[m] = qd_nan_to_num(row[m]); ct[m] = twid[m]; st[m] = twid[n + m]; }
  __syncthreads();
  for (int k = threadIdx.x; k < nstop; k += blockDim.x) {
    const int kk = kcut + k;
    double sr = 0.0, si = 0.0;
    int ph = 0;
    for (int m = 0; m < n; ++m) {
      const double x = xs[m];
      sr = sr + x * ct[ph];
      si = si - x * st[ph];
      ph += kk; if (ph >= n) ph -= n;
    }
    re[k] = sr; im[k] = si;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    double acc = 0.0;
    int ph = (int)(((long long)kcut * m) % n);
    for (int k = 0; k < nstop; ++k) {
      const int kk = kcut + k;
      const bool nyq = (2 * kk == n);
      const double wgt = (kk == 0 || nyq) ? 1.0 : 2.0;
      const double term = nyq ? re[k] * ct[ph] : (re[k] * ct[ph] - im[k] * st[ph]);
      acc = acc + wgt * term;
      ph += m; if (ph >= n) ph -= n;
    }
    row[m] = qd_nan_to_num(xs[m] - d * (acc / (double)n));
  }
  (void)outbuf;
#else
  double* orow = outbuf + (size_t)b * g.ncell + (size_t)j * n;
  const double* ct = twid;
  const double* st = twid + n;
  QD_BLOCK_FIRST_FOR(k, nstop) {
    const int kk = kcut + k;
    double sr = 0.0, si = 0.0;
    for (int m = 0; m < n; ++m) {
      const int ph = (int)(((long long)kk * m) % n);
      const double x = qd_nan_to_num(row[m]);
      sr = sr + x * ct[ph];
      si = si - x * st[ph];
    }
    re[k] = sr; im[k] = si;
  }
  __syncthreads();
  QD_BLOCK_FIRST_FOR(m, n) {
    double acc = 0.0;
    for (int k = 0; k < nstop; ++k) {
      const int kk = kcut + k;
      const int ph = (int)(((long long)kk * m) % n);
      const bool nyq = (2 * kk == n);
      const double wgt = (kk == 0 || nyq) ? 1.0 : 2.0;
      // Re(X_k e^{+i 2 pi k m / n}); irfft drops the imaginary part of the Nyquist bin
      const double term = nyq ? re[k] * ct[ph] : (re[k] * ct[ph] - im[k] * st[ph]);
      acc = acc + wgt * term;
    }
    orow[m] = qd_nan_to_num(qd_nan_to_num(row[m]) - d * (acc / (double)n));
  }
  __syncthreads();
  QD_BLOCK_FIRST_FOR(m, n) { row[m] = orow[m]; }
#endif
}

// ------------------------------------------------------------------------------ reductions
// sum over cells of x * w[j] (energy.py:524, hydrology.py:267) -> out[b*stride]; deterministic: block
// partials are combined in a fixed order by the last block to finish.
__global__ void __launch_bounds__(QD_THREADS) k_wsum(QdGeo g, const double* x, const double* wrow,
                                                    double* partial, unsigned* ticket, double* out, int out_stride) {
  double tot;
  double* part = partial + (size_t)blockIdx.y * g.nvb;
  QD_VB_LOOP(g) {
    double v = 0.0;
    QD_VB_CELLS(g, g.ncomp) { QD_CELL_JI(g) if (qd_owned(g, j)) v += x[off + idx] * (wrow ? wrow[j] : 1.0); }
    if (qd_block_sum<0>(v, &tot)) part[vb_] = tot;
  }
  if (qd_block_is_last(ticket + blockIdx.y, gridDim.x)) {
    if (qd_final_sum<1>(part, g.nvb, &tot)) out[(size_t)blockIdx.y * out_stride] = tot;
  }
}
