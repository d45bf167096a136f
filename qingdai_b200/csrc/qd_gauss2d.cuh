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
#endif
