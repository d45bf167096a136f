// qd_select.cuh -- exact median of the positive entries of a field, one cooperative kernel.
//
// np.median(x[x>0]) (physics.py:298-301, run_simulation.py:1872-1873, dynamics.py:344-348) is an
// exact order statistic; the loop needs three of them per step.  Positive IEEE doubles order like
// their 63-bit patterns, so the lower-middle element is found by an MSD radix select with digits
// of 11 (exponent) + 4 x 13 bits.  Every pass builds a shared-memory histogram per block, merges it
// into a per-(pass, member) global histogram, and after one grid-wide sync every block locates the
// digit redundantly (no second sync, no host round trip).  A closing pass finds the smallest
// element above the lower median for even counts (np.median = mean of the two middle values).
// The whole selection is ONE persistent cooperative launch: 6 grid syncs instead of ~15 launches.
#pragma once
#include "qd_ops.cuh"

#define QD_SEL_PASSES 5
#define QD_SEL_MAXBINS 8192
#define QD_SEL_THREADS 512

struct QdSelOut { double* value; double* count; int stride; double empty_value; };

#if !QD_EMU
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// Finds, for one member, the bin holding 0-based rank `rank` in hist[0..nbins), returns the bin and
// sets *below (elements in lower bins), *inbin (elements in that bin), *total.  Block-cooperative.
__device__ __forceinline__ int qd_sel_locate(const unsigned* __restrict__ hist, int nbins, unsigned long long rank,
                                              unsigned* sh, unsigned long long* part,
                                              unsigned long long* below, unsigned long long* inbin,
                                              unsigned long long* total, int* next_nonempty) {
  __shared__ int s_bin, s_next;
  __shared__ unsigned long long s_below, s_inbin, s_total;
  const int t = threadIdx.x;
  for (int k = t; k < nbins; k += QD_SEL_THREADS) sh[k] = __ldcg(hist + k);
  __syncthreads();
  const int per = QD_SEL_MAXBINS / QD_SEL_THREADS;        // 16 bins per thread
  unsigned long long s = 0;
  for (int k = 0; k < per; ++k) { const int idx = t * per + k; if (idx < nbins) s += sh[idx]; }
  part[t] = s;
  __syncthreads();
  if (t < 32) {
    const int pl = QD_SEL_THREADS / 32;                   // 16 partials per lane
    unsigned long long ls = 0;
    for (int k = 0; k < pl; ++k) ls += part[t * pl + k];
    unsigned long long inc = ls;
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o); if (t >= o) inc += y; }
    const unsigned long long tot = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned long long exc = inc - ls;
    const unsigned ball = __ballot_sync(0xffffffffu, inc > rank);
    const int lane = ball ? (__ffs(ball) - 1) : 31;
    if (t == lane) {
      unsigned long long cum = exc;
      int c = t * pl;
      for (; c < t * pl + pl - 1; ++c) { if (cum + part[c] > rank) break; cum += part[c]; }
      int k = c * per;
      const int kend = min(k + per, nbins) - 1;
      for (; k < kend; ++k) { if (cum + sh[k] > rank) break; cum += sh[k]; }
      int nx = -1;
      for (int q = k + 1; q < nbins; ++q) if (sh[q]) { nx = q; break; }
      s_bin = k; s_below = cum; s_inbin = sh[k]; s_total = tot; s_next = nx;
    }
  }
  __syncthreads();
  *below = s_below; *inbin = s_inbin; *total = s_total; *next_nonempty = s_next;
  const int r = s_bin;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(QD_SEL_THREADS) k_select_coop(QdGeo g, const double* __restrict__ x, unsigned* hist,
                                                               unsigned long long* mingt, QdSelOut out) {
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned sh[QD_SEL_MAXBINS];
  __shared__ unsigned long long part[QD_SEL_THREADS];
  const int b = blockIdx.y;
  const size_t off = (size_t)b * g.ncell;
  const int stride = gridDim.x * QD_SEL_THREADS;
  const int t0 = blockIdx.x * QD_SEL_THREADS + threadIdx.x;
  const int shifts[QD_SEL_PASSES] = {52, 39, 26, 13, 0};
  const int nbits[QD_SEL_PASSES] = {11, 13, 13, 13, 13};
  unsigned long long prefix = 0, rank = 0, count = 0, le = 0;
  for (int pass = 0; pass < QD_SEL_PASSES; ++pass) {
    const int shift = shifts[pass], nb = 1 << nbits[pass];
    unsigned* gh = hist + ((size_t)pass * g.batch + b) * QD_SEL_MAXBINS;
    for (int k = threadIdx.x; k < nb; k += QD_SEL_THREADS) sh[k] = 0;
    __syncthreads();
    const int hi = shift + nbits[pass];
    for (int idx = t0; idx < g.ncell; idx += stride) {
      const double v = x[off + idx];
      if (v > 0.0) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(v);
        if (pass == 0 || (key >> hi) == (prefix >> hi)) atomicAdd(&sh[(key >> shift) & (unsigned long long)(nb - 1)], 1u);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += QD_SEL_THREADS) { const unsigned c = sh[k]; if (c) atomicAdd(gh + k, c); }
    __threadfence();
    grid.sync();
    unsigned long long below, inbin, total; int nx;
    const int bin = qd_sel_locate(gh, nb, rank, sh, part, &below, &inbin, &total, &nx);
    if (pass == 0) {
      count = total;
      rank = count ? (count - 1) / 2 : 0;
      // rank changed from 0 -> relocate with the real rank (members with no positives keep going:
      // every block of the grid must reach every grid.sync)
      const int bin2 = qd_sel_locate(gh, nb, rank, sh, part, &below, &inbin, &total, &nx);
      prefix |= ((unsigned long long)bin2) << shift;
    } else {
      prefix |= ((unsigned long long)bin) << shift;
    }
    le += below;
    rank -= below;
    if (pass == QD_SEL_PASSES - 1) le += inbin;            // all elements of the last bucket equal L
  }
  // closing pass: smallest element above L when the upper middle element is not L itself
  const double L = __longlong_as_double((long long)prefix);
  const bool need_upper = count > 0 && !(count & 1ull) && !(le > count / 2);
  if (need_upper) {
    unsigned long long m = ~0ull;
    for (int idx = t0; idx < g.ncell; idx += stride) {
      const double v = x[off + idx];
      if (v > L) { const unsigned long long key = (unsigned long long)__double_as_longlong(v); if (key < m) m = key; }
    }
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_down_sync(0xffffffffu, m, o); if (y < m) m = y; }
    if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(mingt + b, m);
    __threadfence();
  }
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double r = out.empty_value;
    if (count > 0) {
      if (count & 1ull) r = L;
      else {
        const double U = need_upper ? __longlong_as_double((long long)__ldcg(mingt + b)) : L;
        r = (L + U) / 2.0;                                  // np.mean of the two middle values
      }
    }
    out.value[(size_t)b * out.stride] = r;
    if (out.count) out.count[(size_t)b * out.stride] = (double)count;
  }
}
#else
// Host check build: same contract, evaluated serially (scaffolding only; the CUDA kernel above is
// what the GPU tests exercise, including tests/qdcheck.py:check_median_edge_cases).
#include <algorithm>
#include <vector>
static void qd_select_host(const QdGeo& g, const double* x, QdSelOut out) {
  for (int b = 0; b < g.batch; ++b) {
    std::vector<double> p;
    for (int k = 0; k < g.ncell; ++k) { const double v = x[(size_t)b * g.ncell + k]; if (v > 0.0) p.push_back(v); }
    double r = out.empty_value;
    if (!p.empty()) {
      std::sort(p.begin(), p.end());
      const size_t n = p.size();
      r = (n & 1) ? p[n / 2] : (p[n / 2 - 1] + p[n / 2]) / 2.0;
    }
    out.value[(size_t)b * out.stride] = r;
    if (out.count) out.count[(size_t)b * out.stride] = (double)p.size();
  }
}
#endif
