// qd_gauss2d.cuh -- separable Gaussian (scipy gaussian_filter: axis 0 then axis 1, NI_Correlate1D symmetric summation
// order, 'reflect' or 'wrap' extension) as ONE shared-memory tile kernel per call site instead of a
// latitude-pass kernel, an intermediate field in HBM and a longitude-pass kernel.  A block stages the
// (TJ + 2r) x (TI + 2r) halo tile of each input (boundary extension resolved once per staged element instead of once
// per tap), filters along latitude into a TJ x (TI + 2r) tile, then along longitude, and applies the call site's
// point-wise epilogue (precipitation blend, cloud blend, tracer blend).  Sums and their order are those of the
// two-pass kernels (qd_gauss_tap), so results are bit-identical.  GPU only: the sequential host check build keeps
// the two-pass kernels.
#pragma once
#include "qd_loop.cuh"

#if !QD_EMU
#define QD_G2_TJ 16
#define QD_G2_TI 64
#define QD_G2_RMAX 4
#define QD_G2_NX 64
#define QD_G2_NY 4

enum { QD_G2_PLAIN = 0, QD_G2_PRECIP = 1, QD_G2_CLOUD_B = 2, QD_G2_CLOUD_C = 3 };
struct QdG2Args {
  int n;                               // inputs
  const double* src[2];
  double* dst[2];                      // PLAIN: one output per input; otherwise dst[0] is the call site's output field
  double dt;                           // CLOUD_B: dt / (6 * 3600), the quotient formed on the host
  int row0, row1;                      // output rows of this launch (latitude bands: one launch per segment)
};

// R = compile-time radius (4: sigma = 1; 1: sigma = 0.2; 0: generic radius <= QD_G2_RMAX taken from w.r).
// TJ x TI = output tile of a block of 256 threads: 16 x 64 on large grids, 8 x 32 where that would leave most SMs
// without a tile (181x360: 276 tiles instead of 72).
template <int MODE, int R, int TJ = QD_G2_TJ, int TI = QD_G2_TI>
__global__ void __launch_bounds__(QD_G2_NX * QD_G2_NY) k_gauss2d_tile(QdGeo g, QdG2Args A, QdGaussW w) {
  constexpr int RM = R > 0 ? R : QD_G2_RMAX, NX = TI, NY = (QD_G2_NX * QD_G2_NY) / TI;
  constexpr int PER = TJ / NY;                                  // outputs per thread
  __shared__ double in[(TJ + 2 * RM) * (TI + 2 * RM)];
  __shared__ double mid[TJ * (TI + 2 * RM)];
  const int b = blockIdx.y, r = R > 0 ? R : w.r;
  const int nlat = g.nlat, nlon = g.nlon;
  const int tiles_i = (nlon + TI - 1) / TI;
  const int tj = blockIdx.x / tiles_i, ti = blockIdx.x - tj * tiles_i;
  const int j0 = A.row0 + tj * TJ, i0 = ti * TI;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * NX + tx;
  const int CW = TI + 2 * r, RH = TJ + 2 * r;                    // staged tile extent for this radius
  const size_t off = (size_t)b * g.ncell;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  int nf = A.n;
  double scale[2] = {1.0, 1.0};
  if (MODE == QD_G2_PRECIP) {                                   // physics.py:321-352: P_raw * s and, in the weak-humidity fallback, k_precip * pos
    scale[0] = qd_precip_renorm(g, b);
    scale[1] = P[QD_P_K_PRECIP];
    nf = qd_precip_fallback(g, b) ? 2 : 1;
  }
  double res[2][PER];
  for (int f = 0; f < nf; ++f) {
    const double* __restrict__ S = A.src[f] + off;
    const double sc = scale[f];
    const bool inside = (j0 - r >= 0) && (j0 + TJ + r <= nlat) && (i0 - r >= 0) && (i0 + TI + r <= nlon);   // block-uniform
    if (inside) {                                                // no boundary extension anywhere in this tile
      const double* __restrict__ S0 = S + (size_t)(j0 - r) * nlon + (i0 - r);
      for (int e = tid; e < RH * CW; e += NX * NY) {
        const int rr = e / CW, cc = e - rr * CW;
        const double v = S0[(size_t)rr * nlon + cc];
        in[rr * (TI + 2 * RM) + cc] = (MODE == QD_G2_PRECIP) ? (f == 0 ? v * sc : sc * v) : v;
      }
    } else {
      for (int e = tid; e < RH * CW; e += NX * NY) {
        const int rr = e / CW, cc = e - rr * CW;
        const int gj = qd_extend(j0 - r + rr, nlat, w.wrap), gi = qd_extend(i0 - r + cc, nlon, w.wrap);
        const double v = S[(size_t)gj * nlon + gi];
        in[rr * (TI + 2 * RM) + cc] = (MODE == QD_G2_PRECIP) ? (f == 0 ? v * sc : sc * v) : v;
      }
    }
    __syncthreads();
    for (int e = tid; e < TJ * CW; e += NX * NY) {                // latitude pass (axis 0)
      const int rr = e / CW, cc = e - rr * CW;
      const double* col = in + (rr + r) * (TI + 2 * RM) + cc;
      double o = col[0] * w.w[r];
#pragma unroll
      for (int jj = -RM; jj < 0; ++jj) if (jj >= -r) o = o + (col[jj * (TI + 2 * RM)] + col[-jj * (TI + 2 * RM)]) * w.w[r + jj];
      mid[rr * (TI + 2 * RM) + cc] = o;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < PER; ++m) {                                // longitude pass (axis 1)
      const int rr = ty + m * NY;
      const double* row = mid + rr * (TI + 2 * RM) + tx + r;
      double o = row[0] * w.w[r];
#pragma unroll
      for (int jj = -RM; jj < 0; ++jj) if (jj >= -r) o = o + (row[jj] + row[-jj]) * w.w[r + jj];
      res[f][m] = o;
    }
    __syncthreads();
  }
  const int gi = i0 + tx;
  if (gi >= nlon) return;
#pragma unroll
  for (int m = 0; m < PER; ++m) {
    const int gj = j0 + ty + m * NY;
    if (gj >= A.row1) continue;
    const size_t c = off + (size_t)gj * nlon + gi;
    if (MODE == QD_G2_PLAIN) {
      for (int f = 0; f < nf; ++f) A.dst[f][c] = res[f][m];
    } else if (MODE == QD_G2_PRECIP) {                            // k_precip_d
      double Pv = res[0][m];
      if (nf == 2) Pv = (1.0 - P[QD_P_P_BLEND]) * Pv + P[QD_P_P_BLEND] * res[1][m];
      A.dst[0][c] = (Pv != Pv) ? Pv : (Pv < 0.0 ? 0.0 : Pv);
    } else if (MODE == QD_G2_CLOUD_B) {                           // k_cloud_b
      const double C_P = qd_clip(res[0][m], 0.0, 1.0);
      const double src = qd_clip(res[1][m], 0.0, 1.0);
      const double tend = src * A.dt;
      double cl = A.dst[0][c];
      cl = P[QD_P_W_MEM] * cl + P[QD_P_W_P] * C_P + P[QD_P_W_SRC] * qd_clip(cl + tend, 0.0, 1.0);
      if (P[QD_P_CLOUD_FLOOR] > 0.0) cl = qd_max(cl, qd_clip(P[QD_P_CLOUD_FLOOR] * C_P, 0.0, 1.0));
      A.dst[0][c] = qd_clip(cl, 0.0, 1.0);
    } else {                                                      // k_cloud_c
      const double al = P[QD_P_CLOUD_ADV_ALPHA];
      A.dst[0][c] = qd_clip((1.0 - al) * A.dst[0][c] + al * res[0][m], 0.0, 1.0);
    }
  }
}

// ---- sigma = 1 (radius 4) on large grids: register sliding windows + TMA box loads ---------------------------------------
// The generic tile kernel above reads 9 shared-memory taps per output in both passes and spends ~160 thread instructions
// per cell and field (profiles/r02_ncu_full_hires_step.csv: 13-22 % of DRAM peak, issue-bound).  Here
//   * the (16+8) x (56+8) halo tile of a block-uniform INSIDE tile (no boundary extension anywhere in it) is fetched by ONE
//     thread with a TMA box load (cp.async.bulk.tensor.2d -> shared memory, completion on an mbarrier; SASS: UTMALDG) from
//     a tensor map of the field block [B * n_lat][n_lon] built at qd_bind; tiles that touch a boundary keep the per-element
//     staging with the extension resolved per element (TMA can only zero-fill, scipy reflects / wraps);
//   * latitude pass: a thread owns ONE column and FOUR consecutive output rows, i.e. a sliding window of 12 values:
//     3 shared loads per output instead of 9;
//   * longitude pass: a thread owns FOUR consecutive output columns of one row: 12 values as six 16-byte loads.
// Sums and their order are those of qd_gauss_tap, so results stay bit-identical to the two-pass kernels.
#include <cuda.h>
#ifndef QD_G3_TJ
#define QD_G3_TJ 16                     // measured: 16-row tiles 57.8 / 52.2 / 31.3 us (cloud_b / precip / plain at 1441x2880), 32-row tiles 67.9 / 56.1 / 33.1
#endif
#define QD_G3_TI 56
#define QD_G3_CW 64                     // TI + 2 * 4
#define QD_G3_RH (QD_G3_TJ + 8)         // TJ + 2 * 4

__device__ __forceinline__ void qd_mbar_init(unsigned long long* mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void qd_mbar_expect_tx(unsigned long long* mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qd_mbar_wait(unsigned long long* mbar, unsigned phase) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(phase) : "memory");
  }
}
__device__ __forceinline__ void qd_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, unsigned long long* mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(mbar)) : "memory");
}

template <int MODE, bool TMA>
#ifndef QD_LB_G2R4
#define QD_LB_G2R4 1
#endif
__global__ void __launch_bounds__(256, QD_LB_G2R4) k_gauss2d_r4(QdGeo g, QdG2Args A, QdGaussW w, const __grid_constant__ CUtensorMap tm0,
                                                   const __grid_constant__ CUtensorMap tm1) {
  constexpr int TJ = QD_G3_TJ, TI = QD_G3_TI, CW = QD_G3_CW, RH = QD_G3_RH;
  __shared__ __align__(128) double in2[2][RH * CW];          // one buffer per input field: both TMA loads are in flight from the start
  __shared__ __align__(16) double mid[TJ * CW];
  __shared__ __align__(8) unsigned long long mbar2[2];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int nlat = g.nlat, nlon = g.nlon;
  const int tiles_i = (nlon + TI - 1) / TI;
  const int tj = blockIdx.x / tiles_i, ti = blockIdx.x - tj * tiles_i;
  const int j0 = A.row0 + tj * TJ, i0 = ti * TI;
  const size_t off = (size_t)b * g.ncell;
  const double* P = g.prm + (size_t)b * QD_P_COUNT;
  int nf = A.n;
  double scale[2] = {1.0, 1.0};
  if (MODE == QD_G2_PRECIP) {                                   // physics.py:321-352: P_raw * s and, in the weak-humidity fallback, k_precip * pos
    scale[0] = qd_precip_renorm(g, b);
    scale[1] = P[QD_P_K_PRECIP];
    nf = qd_precip_fallback(g, b) ? 2 : 1;
  }
  const bool inside = (j0 - 4 >= 0) && (j0 + TJ + 4 <= nlat) && (i0 - 4 >= 0) && (i0 + TI + 4 <= nlon);   // block-uniform
  if (TMA) {
    if (tid == 0) { qd_mbar_init(&mbar2[0], 1); qd_mbar_init(&mbar2[1], 1); }
    __syncthreads();
    if (inside && tid == 0) {
      for (int f = 0; f < nf; ++f) {
        qd_mbar_expect_tx(&mbar2[f], RH * CW * 8);
        qd_tma_load_2d(in2[f], f == 0 ? &tm0 : &tm1, i0 - 4, b * nlat + j0 - 4, &mbar2[f]);
      }
    }
  }
  const double w0 = w.w[0], w1 = w.w[1], w2 = w.w[2], w3 = w.w[3], w4 = w.w[4];
  // longitude-pass items: (row lr, output columns 4 * lg .. 4 * lg + 3 of the tile), TJ * 14 of them over 256 threads
  constexpr int NIT = (TJ * 14 + 255) / 256, RPT = TJ / 4;     // items per thread; latitude-pass rows per thread
  double res[2][NIT][4];
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    if (f >= nf) break;
    const double* __restrict__ S = A.src[f] + off;
    const double sc = scale[f];
    double* in = in2[f];
    if (TMA && inside) {
      qd_mbar_wait(&mbar2[f], 0);
    } else {
      if (inside) {
        const double* __restrict__ S0 = S + (size_t)(j0 - 4) * nlon + (i0 - 4);
        for (int e = tid; e < RH * CW; e += 256) in[e] = S0[(size_t)(e >> 6) * nlon + (e & 63)];
      } else {
        for (int e = tid; e < RH * CW; e += 256) {
          const int gj = qd_extend(j0 - 4 + (e >> 6), nlat, w.wrap), gi = qd_extend(i0 - 4 + (e & 63), nlon, w.wrap);
          in[e] = S[(size_t)gj * nlon + gi];
        }
      }
      __syncthreads();
    }
    {   // latitude pass (axis 0): column cc, output rows RPT * q .. RPT * q + RPT - 1 (sliding window of RPT + 8 values)
      const int cc = tid & 63, q = tid >> 6;
      double v[RPT + 8];
#pragma unroll
      for (int k = 0; k < RPT + 8; ++k) {
        const double x = in[(RPT * q + k) * CW + cc];
        v[k] = (MODE == QD_G2_PRECIP) ? (f == 0 ? x * sc : sc * x) : x;
      }
#pragma unroll
      for (int m = 0; m < RPT; ++m) {
        double o = v[m + 4] * w4;
        o = o + (v[m] + v[m + 8]) * w0;
        o = o + (v[m + 1] + v[m + 7]) * w1;
        o = o + (v[m + 2] + v[m + 6]) * w2;
        o = o + (v[m + 3] + v[m + 5]) * w3;
        mid[(RPT * q + m) * CW + cc] = o;
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < NIT; ++it) {   // longitude pass (axis 1)
      const int item = tid + 256 * it;
      if (item < TJ * 14) {
        const int lr = item / 14, lg = item - lr * 14;
        const double2* row = reinterpret_cast<const double2*>(mid + lr * CW + 4 * lg);
        double v[12];
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double2 t = row[k]; v[2 * k] = t.x; v[2 * k + 1] = t.y; }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          double o = v[m + 4] * w4;
          o = o + (v[m] + v[m + 8]) * w0;
          o = o + (v[m + 1] + v[m + 7]) * w1;
          o = o + (v[m + 2] + v[m + 6]) * w2;
          o = o + (v[m + 3] + v[m + 5]) * w3;
          res[f][it][m] = o;
        }
      }
    }
    __syncthreads();                                            // `mid` is reused by the next field
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
  const int item = tid + 256 * it;
  if (item >= TJ * 14) continue;
  const int lr = item / 14, lg = item - lr * 14;
  const int gj = j0 + lr;
  if (gj >= A.row1) continue;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int gi = i0 + 4 * lg + m;
    if (gi >= nlon) continue;
    const size_t c = off + (size_t)gj * nlon + gi;
    if (MODE == QD_G2_PLAIN) {
      A.dst[0][c] = res[0][it][m];
      if (nf > 1) A.dst[1][c] = res[1][it][m];
    } else if (MODE == QD_G2_PRECIP) {                            // k_precip_d
      double Pv = res[0][it][m];
      if (nf == 2) Pv = (1.0 - P[QD_P_P_BLEND]) * Pv + P[QD_P_P_BLEND] * res[1][it][m];
      A.dst[0][c] = (Pv != Pv) ? Pv : (Pv < 0.0 ? 0.0 : Pv);
    } else if (MODE == QD_G2_CLOUD_B) {                           // k_cloud_b
      const double C_P = qd_clip(res[0][it][m], 0.0, 1.0);
      const double src = qd_clip(res[1][it][m], 0.0, 1.0);
      const double tend = src * A.dt;
      double cl = A.dst[0][c];
      cl = P[QD_P_W_MEM] * cl + P[QD_P_W_P] * C_P + P[QD_P_W_SRC] * qd_clip(cl + tend, 0.0, 1.0);
      if (P[QD_P_CLOUD_FLOOR] > 0.0) cl = qd_max(cl, qd_clip(P[QD_P_CLOUD_FLOOR] * C_P, 0.0, 1.0));
      A.dst[0][c] = qd_clip(cl, 0.0, 1.0);
    } else {                                                      // k_cloud_c
      const double al = P[QD_P_CLOUD_ADV_ALPHA];
      A.dst[0][c] = qd_clip((1.0 - al) * A.dst[0][c] + al * res[0][it][m], 0.0, 1.0);
    }
  }
  }
}
#endif
