// qd_select.cuh -- exact median of the positive entries of a field, one cooperative kernel.
//
// np.median(x[x>0]) (physics.py:298-301, run_simulation.py:1872-1873, dynamics.py:344-348) is an
// exact order statistic; the loop needs three of them per step.  Positive IEEE doubles order like
// their 63-bit patterns, so the lower-middle element is found by an MSD radix select with 11-bit digits.
// Every pass builds a shared-memory histogram per block, merges it into a per-(pass, member) global
// histogram, and after one grid-wide sync every block locates the digit redundantly (no second sync, no
// host round trip).  As soon as <= QD_SEL_CAP candidates remain they are gathered and sorted by one block,
// which also yields the upper middle element for even counts (np.median = mean of the two middle values).
// The whole selection is ONE persistent cooperative launch.
//
// Speculation on the first digit: the median of a field moves slowly from step to step, and its first digit (sign +
// exponent of the double) practically never changes.  Every call site remembers the first-digit bucket of its last
// call (`spec`, device memory, per member); the FIRST sweep then also builds the second-digit histogram restricted to
// that bucket.  If the located first digit equals the remembered one (checked exactly, every time), the second sweep
// -- and, with latitude bands, its cross-rank round -- is skipped: 2 sweeps + 2 grid syncs (+ 2 rounds) instead of
// 3 + 3 (+ 3).  A miss costs nothing but the wasted shared-memory atomics: the normal second pass runs.
#pragma once
#include "qd_band.cuh"           // QD_SEL_* sizes, QdBandCtl and the cross-rank pieces used when the field is split over ranks

#define QD_SEL_UNROLL 4
struct QdSelOut { double* value; double* count; int stride; double empty_value; };

#if !QD_EMU
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// Finds, for one member, the bin holding 0-based rank `*rank` in hist[0..nbins) and sets *below (elements
// in lower bins), *inbin (elements in that bin), *total.  With first != 0 the rank is the lower-median
// index (total - 1) / 2 of the whole histogram and is returned through *rank.  Block-cooperative.
__device__ __forceinline__ int qd_sel_locate(const unsigned* __restrict__ hist, int nbins, unsigned long long* rank, int first,
                                              unsigned* sh, unsigned long long* part,
                                              unsigned long long* below, unsigned long long* inbin,
                                              unsigned long long* total) {
  __shared__ int s_bin;
  __shared__ unsigned long long s_below, s_inbin, s_total, s_rank;
  const int t = threadIdx.x;
  for (int k = t; k < nbins; k += QD_SEL_THREADS) sh[k] = __ldcg(hist + k);
  __syncthreads();
  const int per = QD_SEL_MAXBINS / QD_SEL_THREADS;        // 4 bins per thread
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
    const unsigned long long rk = first ? (tot ? (tot - 1) / 2 : 0) : *rank;
    const unsigned long long exc = inc - ls;
    const unsigned ball = __ballot_sync(0xffffffffu, inc > rk);
    const int lane = ball ? (__ffs(ball) - 1) : 31;
    if (t == lane) {
      unsigned long long cum = exc;
      int c = t * pl;
      for (; c < t * pl + pl - 1; ++c) { if (cum + part[c] > rk) break; cum += part[c]; }
      int k = c * per;
      const int kend = min(k + per, nbins) - 1;
      for (; k < kend; ++k) { if (cum + sh[k] > rk) break; cum += sh[k]; }
      s_bin = k; s_below = cum; s_inbin = sh[k]; s_total = tot; s_rank = rk;
    }
  }
  __syncthreads();
  *below = s_below; *inbin = s_inbin; *total = s_total; *rank = s_rank;
  const int r = s_bin;
  __syncthreads();
  return r;
}

// Radix passes of 11 bits (the first is the exponent) narrow the candidates
// until at most QD_SEL_CAP share the prefix -- two passes for continuous data of any size here -- then ONE
// gather pass appends them to a list (and records the smallest key above the prefix bucket); block 0 sorts
// the list in shared memory and reads both middle elements.  Heavily duplicated data simply keeps taking
// radix passes down to bit 0.  3 sweeps + 3 grid syncs in the common case instead of 6 + 6.
// The histograms are left zeroed for the next launch and the list counter / mingt are reset at the start: no
// memset nodes in the step graph.
// hist slots: [0] first digit, [1] speculative second digit, [p + 1] digit p >= 1 -- slots 0 and 1 are adjacent for one
// member, so latitude bands all-reduce both in one round
#define QD_SEL_SLOTS (QD_SEL_PASSES + 1)
#define QD_SEL_NOSPEC 0xffffffffu
__global__ void __launch_bounds__(QD_SEL_THREADS, 2) k_select_coop(QdGeo g, const double* __restrict__ x, unsigned* hist,
                                                               unsigned long long* list, unsigned* lcount,
                                                               unsigned long long* mingt, int* more_flag, QdSelOut out, QdBandCtl B,
                                                               unsigned* spec /* [B] remembered first digit, or null */,
                                                               unsigned* spec_stat /* [2] calls, speculation hits of this site */) {
  cg::grid_group grid = cg::this_grid();
  __shared__ __align__(16) unsigned sh[QD_SEL_MAXBINS > 2 * QD_SEL_CAP ? QD_SEL_MAXBINS : 2 * QD_SEL_CAP];   // histogram, later the candidate keys
  __shared__ unsigned long long part[QD_SEL_THREADS];
  const int b = blockIdx.y;
  const size_t off = (size_t)b * g.ncell;
  const int stride = gridDim.x * QD_SEL_THREADS;
  const int t0 = blockIdx.x * QD_SEL_THREADS + threadIdx.x;
  const int c0 = g.own0 * g.nlon, c1 = g.own1 * g.nlon;        // this rank's own rows (all rows without latitude bands)
  // 11-bit digits: the first is exactly the exponent (positive doubles), the second the top of the mantissa; a
  // 2048-bin histogram keeps the per-pass merge + locate short (every block scans it redundantly)
  const int shifts[QD_SEL_PASSES] = {52, 41, 30, 19, 8, 0};
  const int nbits[QD_SEL_PASSES] = {11, 11, 11, 11, 11, 8};
  unsigned long long prefix = 0, rank = 0, count = 0, inbin = ~0ull;
  int npass = 0, lo_shift = 63;
  // latitude bands: epoch of the k-th cross-rank collective of this launch = word at launch + k (same in every block)
  unsigned long long sel_epoch = B.world > 1 ? qd_bflags(B, B.rank)[QD_BF_EPOCH_SEL] : 0ull;
  const unsigned spec1 = spec ? __ldcg(spec + b) : QD_SEL_NOSPEC;          // first digit of this call site's last median (block-uniform)
  const bool spec_on = spec1 != QD_SEL_NOSPEC;
  unsigned* ghs = hist + ((size_t)1 * g.batch + b) * QD_SEL_MAXBINS;
  // One radix pass.  Every block of the GRID takes part in the sync; members whose candidates already fit
  // the list (work == false) skip the sweep.
  auto radix_pass = [&](int pass, bool work) {
    const int shift = shifts[pass], nb = 1 << nbits[pass];
    unsigned* gh = hist + ((size_t)(pass ? pass + 1 : 0) * g.batch + b) * QD_SEL_MAXBINS;
    const bool spec_pass = pass == 0 && spec_on;
    if (work) {
      for (int k = threadIdx.x; k < (spec_pass ? 2 * nb : nb); k += QD_SEL_THREADS) sh[k] = 0;
      __syncthreads();
      const int hi = shift + nbits[pass];
      // four independent loads per trip: with one load in flight per thread the sweep is bound by HBM latency
      // (27 dependent trips of ~1 us at 1441x2880), not by bandwidth
      // First digit: nearly every key of a field shares ONE exponent, so the plain histogram serialises a warp's 32
      // atomics on one shared-memory word -- that contention, not the loads, is the cost of this sweep.  Keys with the
      // remembered digit are therefore counted in a register (one atomic per warp at the end); their second digit
      // spreads over 2048 bins.
      unsigned nspec = 0u;
      for (int idx = c0 + t0; idx < c1; idx += QD_SEL_UNROLL * stride) {
        double v[QD_SEL_UNROLL];
#pragma unroll
        for (int u = 0; u < QD_SEL_UNROLL; ++u) { const int i2 = idx + u * stride; v[u] = (i2 < c1) ? x[off + i2] : 0.0; }
#pragma unroll
        for (int u = 0; u < QD_SEL_UNROLL; ++u)
          if (v[u] > 0.0) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[u]);
            if (pass == 0) {
              const unsigned d1 = (unsigned)(key >> 52);
              if (spec_pass && d1 == spec1) { ++nspec; atomicAdd(&sh[QD_SEL_MAXBINS + (unsigned)((key >> 41) & 2047ull)], 1u); }
              else atomicAdd(&sh[d1], 1u);
            } else if ((key >> hi) == (prefix >> hi)) atomicAdd(&sh[(key >> shift) & (unsigned long long)(nb - 1)], 1u);
          }
      }
      if (spec_pass) {
        for (int o = 16; o > 0; o >>= 1) nspec += __shfl_down_sync(0xffffffffu, nspec, o);
        if ((threadIdx.x & 31) == 0 && nspec) atomicAdd(&sh[spec1 & (QD_SEL_MAXBINS - 1)], nspec);
      }
      __syncthreads();
      for (int k = threadIdx.x; k < nb; k += QD_SEL_THREADS) { const unsigned c = sh[k]; if (c) atomicAdd(gh + k, c); }
      if (spec_pass) for (int k = threadIdx.x; k < nb; k += QD_SEL_THREADS) { const unsigned c = sh[QD_SEL_MAXBINS + k]; if (c) atomicAdd(ghs + k, c); }
      __threadfence();
    }
    grid.sync();
    if (B.world > 1 && work) {                             // latitude bands (one member): histogram of the whole domain
      qd_band_hist_allreduce(B, gh, spec_pass ? 2 * nb : nb, ++sel_epoch);      // slots 0 and 1 are adjacent (batch = 1)
      grid.sync();
    }
    if (work) {
      unsigned long long below, total;
      const int bin = qd_sel_locate(gh, nb, &rank, pass == 0, sh, part, &below, &inbin, &total);
      if (pass == 0) count = total;
      prefix |= ((unsigned long long)bin) << shift;
      rank -= below;
      npass = pass + 1; lo_shift = shift;
    }
  };
  // Gather pass: candidates sharing the prefix -> list (when they fit); smallest key above the prefix
  // bucket -> mingt (needed when the upper middle element of an even count lies outside the bucket).
  auto gather_pass = [&](bool work) {
    const bool fits = inbin <= QD_SEL_CAP;
    const bool need_above = count > 0 && !(count & 1ull) && (rank + 1 >= inbin);
    if (work && ((fits && inbin > 0) || need_above)) {
      unsigned long long m = ~0ull;
      unsigned long long* lst = list + (size_t)b * QD_SEL_CAP;
      const unsigned long long pk = prefix >> lo_shift;
      for (int idx = c0 + t0; idx < c1; idx += QD_SEL_UNROLL * stride) {
        double v[QD_SEL_UNROLL];
#pragma unroll
        for (int u = 0; u < QD_SEL_UNROLL; ++u) { const int i2 = idx + u * stride; v[u] = (i2 < c1) ? x[off + i2] : 0.0; }
#pragma unroll
        for (int u = 0; u < QD_SEL_UNROLL; ++u)
          if (v[u] > 0.0) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[u]);
            const unsigned long long kk = key >> lo_shift;
            if (kk == pk) { if (fits) lst[atomicAdd(lcount + b, 1u)] = key; }
            else if (kk > pk && key < m) m = key;
          }
      }
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_down_sync(0xffffffffu, m, o); if (y < m) m = y; }
      if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(mingt + b, m);
      __threadfence();
    }
    grid.sync();
  };
  // list counter / "next above" key: reset here (they are first touched in the gather pass, two grid syncs later)
  if (blockIdx.x == 0 && threadIdx.x == 0) { lcount[b] = 0u; mingt[b] = ~0ull; }
  radix_pass(0, true);
  // speculation check (exact): the located first digit is the remembered one -> its second-digit histogram is already
  // complete in slot 1; locate there instead of sweeping again
  const bool spec_hit = spec_on && (unsigned)(prefix >> 52) == spec1 && count > 0;
  if (spec_hit && inbin > QD_SEL_CAP) {
    unsigned long long below, total;
    const int bin = qd_sel_locate(ghs, 1 << nbits[1], &rank, 0, sh, part, &below, &inbin, &total);
    prefix |= ((unsigned long long)bin) << shifts[1];
    rank -= below;
    lo_shift = shifts[1];
  }
  {
    // every block of the grid passes through radix_pass(1)'s sync; members that hit (or already fit) do no work in it
    const bool work1 = !spec_hit && inbin > QD_SEL_CAP;
    radix_pass(1, work1);
  }
  const bool more = inbin > QD_SEL_CAP;                    // heavily duplicated data: keep narrowing
  if (more && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(more_flag, 1);
  gather_pass(!more);
  if (__ldcg(more_flag)) {                                 // grid-uniform (set before the sync above)
    radix_pass(2, more);
    radix_pass(3, more && inbin > QD_SEL_CAP);
    radix_pass(4, more && inbin > QD_SEL_CAP);
    radix_pass(5, more && inbin > QD_SEL_CAP);
    gather_pass(more);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *more_flag = 0;
  }
  // leave the histograms of the passes that ran zeroed for the next launch (every block is past its reads)
  for (int p = 0; p < npass; ++p) {
    unsigned* gh = hist + ((size_t)(p ? p + 1 : 0) * g.batch + b) * QD_SEL_MAXBINS;
    for (int k = t0; k < (1 << nbits[p]); k += stride) gh[k] = 0u;
  }
  if (spec_on) for (int k = t0; k < QD_SEL_MAXBINS; k += stride) ghs[k] = 0u;
  if (spec && blockIdx.x == 0 && threadIdx.x == 0) {
    spec[b] = count > 0 ? (unsigned)(prefix >> 52) : QD_SEL_NOSPEC;
    if (spec_stat) { atomicAdd(spec_stat, 1u); if (spec_hit) atomicAdd(spec_stat + 1, 1u); }       // calls, hits (qd_median_stats)
  }
  const bool fits = inbin <= QD_SEL_CAP;                   // false only when one VALUE fills the last bucket
  unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sh);      // QD_SEL_CAP x u64
  __shared__ unsigned long long s_mingt;
  int mband = 0;
  if (B.world > 1 && count > 0) {                          // merge every rank's candidates and "next above" key
    unsigned long long mg = ~0ull;
    mband = qd_band_list_allgather(B, list + (size_t)b * QD_SEL_CAP, fits ? __ldcg(lcount + b) : 0u, __ldcg(mingt + b), skeys, &mg, ++sel_epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) s_mingt = mg;
  } else if (threadIdx.x == 0) s_mingt = __ldcg(mingt + b);
  if (blockIdx.x != 0) return;
  __syncthreads();
  const bool even = count > 0 && !(count & 1ull);
  const bool need_above = even && (rank + 1 >= inbin);
  unsigned long long Lk = prefix, Uk = prefix;
  if (count > 0 && fits) {
    const int m = (int)inbin;
    int n2 = 1; while (n2 < m) n2 <<= 1;
    const unsigned long long* lst = list + (size_t)b * QD_SEL_CAP;
    if (B.world > 1) { for (int k = mband + threadIdx.x; k < n2; k += QD_SEL_THREADS) skeys[k] = ~0ull; }
    else for (int k = threadIdx.x; k < n2; k += QD_SEL_THREADS) skeys[k] = k < m ? __ldcg(lst + k) : ~0ull;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
      for (int st = size >> 1; st > 0; st >>= 1) {
        for (int k = threadIdx.x; k < n2; k += QD_SEL_THREADS) {
          const int q = k ^ st;
          if (q > k) {
            const unsigned long long a = skeys[k], c2 = skeys[q];
            const bool up = (k & size) == 0;
            if ((a > c2) == up) { skeys[k] = c2; skeys[q] = a; }
          }
        }
        __syncthreads();
      }
    Lk = skeys[rank];
    Uk = (rank + 1 < (unsigned long long)m) ? skeys[rank + 1] : Lk;
  }
  if (threadIdx.x == 0) {
    double r = out.empty_value;
    if (count > 0) {
      const double L = __longlong_as_double((long long)Lk);
      if (!even) r = L;
      else {
        const double U = need_above ? __longlong_as_double((long long)s_mingt) : __longlong_as_double((long long)Uk);
        r = (L + U) / 2.0;                                  // np.mean of the two middle values
      }
    }
    out.value[(size_t)b * out.stride] = r;
    if (out.count) out.count[(size_t)b * out.stride] = (double)count;
    if (B.world > 1) qd_bflags(B, B.rank)[QD_BF_EPOCH_SEL] = sel_epoch;
  }
}
#else
// Host check build: same contract, evaluated serially (scaffolding only; the CUDA kernel above is
// what the GPU tests exercise, including tests/qdcheck.py:check_median_edge_cases).
#include <algorithm>
#include <vector>
static void qd_select_host(const QdGeo& g, const double* x, QdSelOut out, const QdBandCtl& B) {
  for (int b = 0; b < g.batch; ++b) {
    std::vector<double> p;
    for (int k = g.own0 * g.nlon; k < g.own1 * g.nlon; ++k) { const double v = x[(size_t)b * g.ncell + k]; if (v > 0.0) p.push_back(v); }
    if (B.world > 1) {             // every rank's positives -> every rank (host check build only: unbounded lists)
      unsigned long long* mine = qd_bflags(B, B.rank);
      const unsigned long long epoch = mine[QD_BF_EPOCH_SEL] + 1ull;
      const size_t cap = (size_t)g.ncell + 1;
      for (int r = 0; r < B.world; ++r) {
        double* dst = (double*)(B.peer[r] + B.off_emu) + (size_t)B.rank * cap;
        dst[0] = (double)p.size();
        for (size_t k = 0; k < p.size(); ++k) dst[1 + k] = p[k];
      }
      qd_fence_sys();
      mine[QD_BF_EPOCH_SEL] = epoch;
      for (int r = 0; r < B.world; ++r) qd_st_sys(qd_bflags(B, r) + QD_BF_SEL + B.rank, epoch);
      for (int r = 0; r < B.world; ++r) qd_band_wait(B, mine + QD_BF_SEL + r, epoch);
      p.clear();
      for (int r = 0; r < B.world; ++r) {
        const double* src = (const double*)(B.peer[B.rank] + B.off_emu) + (size_t)r * cap;
        const size_t n = (size_t)src[0];
        p.insert(p.end(), src + 1, src + 1 + n);
      }
      // a rank may only reuse the boxes after every rank has read them: second round trip
      const unsigned long long e2 = epoch + 1ull;
      qd_fence_sys();
      mine[QD_BF_EPOCH_SEL] = e2;
      for (int r = 0; r < B.world; ++r) qd_st_sys(qd_bflags(B, r) + QD_BF_SEL + B.rank, e2);
      for (int r = 0; r < B.world; ++r) qd_band_wait(B, mine + QD_BF_SEL + r, e2);
    }
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
