// qd_hyper4.cuh -- fused del^4 hyperdiffusion step: F' = nan_to_num(nan_to_num(F) - k4 * lap(lap(F)) * dt)
// (dynamics.py:175-212, ocean.py:119-152) as ONE shared-memory tiled kernel per field group.
//
// The two-kernel form (lap -> scratch, lap -> update) moves 5 x 8 B per cell and evaluates the Laplacian
// with ~10 global loads each; ncu (profiles/r01_ncu_full_ens64_v1.csv) showed it issue-bound at
// 12-16 % of HBM peak.  Here a block stages an (TJ+8) x (TI+4) halo tile of the cleaned field in shared
// memory, evaluates lap(F) on the (TJ+4) x (TI+2) ring it needs, then lap(lap F) and the update on its
// TJ x TI interior: HBM traffic is read F once + write F' once, every operand of the two stencils comes
// from shared memory.  Halo radius: 4 rows (np.gradient twice per Laplacian, one-sided at the poles:
// no wrap in latitude), 2 columns (periodic, period n_lon: np.roll).
// Out of place (a neighbouring block still needs the old halo): the orchestrator ping-pongs slots.
#pragma once
#include "qd_ocean.cuh"

#define QD_H4_TI 64

struct QdHyper4Args {
  int n;                                   // fields in this launch
  const double* src[QD_MAX_FIELDS];
  double* dst[QD_MAX_FIELDS];
  const double* k4rows[QD_MAX_FIELDS];     // per-row coefficient table
  long long k4_bstride[QD_MAX_FIELDS];     // member stride of that table (0: one table shared by every member)
  double scale[QD_MAX_FIELDS];             // multiplier on the table (0.5 for eta)
  int raw_k4[QD_MAX_FIELDS];               // 1: table already holds k4 (atmosphere / overrides); 0: k4 = table / max(1e-12, sub_dt)
  const double* cosr;                      // cosine rows followed by 1/c and 1/c^2
  double dt;                               // step handed to the operator (atmosphere); ocean reads sub_dt from the scalar table
  int nsub;                                // inner sub-division (QD_K4_NSUB / QD_OCEAN_K4_NSUB): sub = dt / nsub
  int ocean;                               // 1: per-member sub_dt and early exit through the device sub-step counter
  QdSubCtl sc;
  int tj_lo, tj_skip;                      // tile kernel: tile rows >= tj_lo are shifted by tj_skip (pole tiles only, see launch_hyper4)
  int ja, jb;                              // streaming kernel: output rows [ja, jb), all with a centred dependency cone
  int row0, row1;                          // tile kernel: output rows [row0, row1) (tile rows count from row0)
};

// Laplacian at (global row j, tile column c) from a tile accessor A(jglobal, tile_col).
template <class Acc>
QD_HD double qd_lap_rel(const Acc& A, int j, int c, const QdGeo& g, const double* cr) {
  const int nlat = g.nlat;
  const double* ic = cr + nlat;
  const double* ic2 = cr + 2 * nlat;
  auto G = [&](int jj) -> double {
    if (jj == 0) return (A(1, c) - A(0, c)) * g.inv_dlat;
    if (jj == nlat - 1) return (A(nlat - 1, c) - A(nlat - 2, c)) * g.inv_dlat;
    return (A(jj + 1, c) - A(jj - 1, c)) * g.inv_2dlat;
  };
  double gphi;
  if (j == 0) gphi = (cr[1] * G(1) - cr[0] * G(0)) * g.inv_dlat;
  else if (j == nlat - 1) gphi = (cr[j] * G(j) - cr[j - 1] * G(j - 1)) * g.inv_dlat;
  else gphi = (cr[j + 1] * G(j + 1) - cr[j - 1] * G(j - 1)) * g.inv_2dlat;
  const double term_phi = ic[j] * gphi;
  const double d2 = ((A(j, c + 1) - 2.0 * A(j, c)) + A(j, c - 1)) * g.inv_dlon_sq;
  return (term_phi + d2 * ic2[j]) * g.inv_a_sq;
}

// 2-D cooperative tile loops: strided over the (x, y) threads of the block on the GPU, run completely
// by the first thread in the sequential host check build.
#if QD_EMU
#define QD_TILE_FIRST_LOOP(r, R, cc, Cn) \
  if (threadIdx.x == 0 && threadIdx.y == 0) for (int r = 0; r < (R); ++r) for (int cc = 0; cc < (Cn); ++cc)
#define QD_TILE_FIRST_ROWS(r, R) if (threadIdx.x == 0 && threadIdx.y == 0) for (int r = 0; r < (R); ++r)
#else
#define QD_TILE_FIRST_LOOP(r, R, cc, Cn) \
  for (int r = threadIdx.y; r < (R); r += blockDim.y) for (int cc = threadIdx.x; cc < (Cn); cc += blockDim.x)
#define QD_TILE_FIRST_ROWS(r, R) for (int r = threadIdx.y * blockDim.x + threadIdx.x; r < (R); r += blockDim.x * blockDim.y)
#endif
#define QD_H4_NX 64
#define QD_H4_NY 4

QD_HD double qd_clean_fast(double x) { return (fabs(x) <= DBL_MAX) ? x : qd_nan_to_num(x); }
#if !QD_EMU
// np.nan_to_num without branches (selects only): keeps the streaming kernel's loads and arithmetic of
// neighbouring rows free to overlap; identical results to qd_nan_to_num.
__device__ __forceinline__ double qd_clean_sel(double x) {
  const int hi = __double2hiint(x);
  const double big = __hiloint2double((hi & 0x80000000) | 0x7fefffff, 0xffffffff);     // copysign(DBL_MAX, x)
  const double alt = (x != x) ? 0.0 : big;
  return (fabs(x) <= DBL_MAX) ? x : alt;
}
#endif

// Interior rows (2 <= j <= n_lat-3, every np.gradient centred) use three per-row coefficients staged in
// shared memory:  lap F = ap*(F[j+2]-F[j]) - am*(F[j]-F[j-2]) + bl*((F[i+1]-2F)+F[i-1])  with
//   ap = (1/c_j)(1/2dphi)(c_{j+1}/2dphi)/a^2,  am = (1/c_j)(1/2dphi)(c_{j-1}/2dphi)/a^2,  bl = (1/c_j^2)(1/dlam^2)/a^2.
// The four rows next to the poles take the general one-sided form (qd_lap_rel).
#if !QD_EMU
// GPU kernel.  64 x 4 threads; a warp is 32 consecutive longitudes of ONE tile row, so every row
// predicate (inside the domain? interior?) is warp-uniform.  All loops have compile-time trip counts
// and shared-memory offsets are immediates (ncu showed the first, generic version spending >70 % of
// its issue slots on index arithmetic and branch bookkeeping: profiles/r01_ncu_hyper4.md).
template <int TJ>
__global__ void __launch_bounds__(QD_H4_NX * QD_H4_NY) k_hyper4_tile(QdGeo g, QdHyper4Args A) {
  constexpr int TI = QD_H4_TI, NX = QD_H4_NX, NY = QD_H4_NY;
  constexpr int RF = TJ + 8, CF = TI + 4, RL = TJ + 4, CL = TI + 2;
  static_assert(RF % NY == 0 && RL % NY == 0 && TJ % NY == 0 && TI == NX, "tile shape");
  __shared__ double Fs[RF * CF];
  __shared__ double Ls[RL * CL];
  __shared__ double cap[RF], cam[RF], cbl[RF], k4s[TJ];
  const int b = blockIdx.y;
  if (A.ocean && qd_sub_done(g, b, A.sc)) return;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * NX + tx;
  const int tiles_i = (g.nlon + TI - 1) / TI;
  const int tj0 = blockIdx.x / tiles_i, ti = blockIdx.x - tj0 * tiles_i;
  const int tj = tj0 >= A.tj_lo ? tj0 + A.tj_skip : tj0;
  const int j0 = A.row0 + tj * TJ, i0 = ti * TI;
  const int jF0 = j0 - 4, jL0 = j0 - 2, iF0 = i0 - 2;
  const int nlat = g.nlat, nlon = g.nlon;
  const size_t off = (size_t)b * g.ncell;
  double sub_dt = A.dt;
  if (A.ocean) sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const double inner = sub_dt / (double)(A.nsub > 1 ? A.nsub : 1);
  const int k = blockIdx.z;                   // one field per block (gridDim.z = number of fields)
  const double* __restrict__ F = A.src[k] + off;
  double* __restrict__ D = A.dst[k] + off;
  const double* __restrict__ cr = A.cosr;
  if (tid < RF) {
    const int gj = jF0 + tid;
    double ap = 0.0, am = 0.0, bl = 0.0;
    if (gj >= 2 && gj <= nlat - 3) {
      const double icj = cr[nlat + gj] * g.inv_2dlat;
      ap = (icj * (cr[gj + 1] * g.inv_2dlat)) * g.inv_a_sq;
      am = (icj * (cr[gj - 1] * g.inv_2dlat)) * g.inv_a_sq;
      bl = (cr[2 * nlat + gj] * g.inv_dlon_sq) * g.inv_a_sq;
    }
    cap[tid] = ap; cam[tid] = am; cbl[tid] = bl;
    if (tid < TJ) {
      const int oj = j0 + tid;
      double k4 = 0.0;
      if (oj < nlat) {
        k4 = A.k4rows[k][(size_t)b * A.k4_bstride[k] + oj];
        if (!A.raw_k4[k]) k4 = k4 / fmax(1e-12, sub_dt);       // ocean.py:347
        k4 = A.scale[k] * k4;
      }
      k4s[tid] = k4;
    }
  }
  // columns of this thread: main column and (first 4 threads of a row) one halo column; periodic wrap
  int gi0 = iF0 + tx;
  while (gi0 < 0) gi0 += nlon;
  while (gi0 >= nlon) gi0 -= nlon;
  int gi1 = iF0 + NX + (tx & 3);
  while (gi1 >= nlon) gi1 -= nlon;
  // ---- phase 1: halo tile of the field.  np.nan_to_num is applied lazily: values are staged raw, a
  // block-wide flag records whether anything non-finite was seen, and only then a cleaning pass runs
  // (identical results; the common all-finite case costs one compare per element instead of a branchy clean).
  const bool rows_ok = (jF0 >= 0) && (jF0 + RF <= nlat);          // block-uniform: no pole in this tile
  const size_t rstride = (size_t)NY * nlon;
  int bad = 0;
  {
    const double* p0 = F + (size_t)(rows_ok ? jF0 + ty : 0) * nlon + gi0;
    const double* p1 = F + (size_t)(rows_ok ? jF0 + ty : 0) * nlon + gi1;
#pragma unroll
    for (int m = 0; m < RF / NY; ++m) {
      const int r = ty + m * NY;
      double v0 = 0.0, v1 = 0.0;
      if (rows_ok) {
        v0 = p0[0];
        if (tx < 4) v1 = p1[0];
        p0 += rstride; p1 += rstride;
      } else {
        const int gj = jF0 + r;
        if (gj >= 0 && gj < nlat) {
          const double* row = F + (size_t)gj * nlon;
          v0 = row[gi0];
          if (tx < 4) v1 = row[gi1];
        }
      }
      bad |= !(fabs(v0) <= DBL_MAX) | !(fabs(v1) <= DBL_MAX);
      Fs[r * CF + tx] = v0;
      if (tx < 4) Fs[r * CF + NX + tx] = v1;
    }
  }
  if (__syncthreads_or(bad)) {
    for (int e = tid; e < RF * CF; e += NX * NY) Fs[e] = qd_nan_to_num(Fs[e]);
    __syncthreads();
  }
  auto AF = [&](int jj, int c) -> double { return Fs[(jj - jF0) * CF + c]; };
  // ---- phase 2: L = lap(F) on the ring needed by the second Laplacian (cleaned lazily like F)
  bad = 0;
#pragma unroll
  for (int m = 0; m < RL / NY; ++m) {
    const int r = ty + m * NY, gj = jL0 + r, fr = r + 2;
    const bool valid = gj >= 0 && gj < nlat, inter = gj >= 2 && gj <= nlat - 3;
    const double ap = cap[fr], am = cam[fr], bl = cbl[fr];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && tx >= 2) break;
      const int cc = h ? NX + tx : tx;
      const int e = fr * CF + cc + 1;
      double L = 0.0;
      if (inter) {
        const double f0 = Fs[e];
        L = (ap * (Fs[e + 2 * CF] - f0) - am * (f0 - Fs[e - 2 * CF])) + bl * ((Fs[e + 1] - 2.0 * f0) + Fs[e - 1]);
      } else if (valid) {
        L = qd_lap_rel(AF, gj, cc + 1, g, cr);
      }
      bad |= !(fabs(L) <= DBL_MAX);
      Ls[r * CL + cc] = L;
    }
  }
  if (__syncthreads_or(bad)) {
    for (int e = tid; e < RL * CL; e += NX * NY) Ls[e] = qd_nan_to_num(Ls[e]);
    __syncthreads();
  }
  auto AL = [&](int jj, int c) -> double { return Ls[(jj - jL0) * CL + c]; };
  // ---- phase 3: lap(L) and the update on the tile interior
  const int gi = i0 + tx;
  if (gi < nlon) {
#pragma unroll
    for (int m = 0; m < TJ / NY; ++m) {
      const int r = ty + m * NY, gj = j0 + r;
      if (gj < A.row1) {
        const int e = (r + 2) * CL + tx + 1, fr = r + 4;
        double L2;
        if (gj >= 2 && gj <= nlat - 3) {
          const double l0 = Ls[e];
          L2 = (cap[fr] * (Ls[e + 2 * CL] - l0) - cam[fr] * (l0 - Ls[e - 2 * CL])) + cbl[fr] * ((Ls[e + 1] - 2.0 * l0) + Ls[e - 1]);
        } else {
          L2 = qd_lap_rel(AL, gj, tx + 1, g, cr);
        }
        D[(size_t)gj * nlon + gi] = qd_clean_fast(Fs[fr * CF + tx + 2] - k4s[r] * L2 * inner);
      }
    }
  }
}

// ---- streaming form for large grids -------------------------------------------------------------------
// One WARP marches down a strip of 28 longitudes (lanes 2..29; lanes 0,1,30,31 carry the +-2 column halo),
// keeping F[j..j+4] and lap(F)[j-2..j+2] in registers; east / west neighbours come from warp shuffles.  No
// shared memory, no barriers, one coalesced load and one store per cell and row: ~50 thread instructions
// per cell instead of ~340 for the tile kernel (profiles/r01_ncu_hires_step.md), which was issue-bound at
// 15 % of HBM peak.  Only rows whose whole dependency cone uses centred differences (4 <= j <= n_lat-5) are
// produced here; the tile kernel does the few rows next to the poles.  Same operand order as the tile
// kernel, so both produce identical bits.
#define QD_H4S_COLS 28
#define QD_H4S_WARPS 4
// `cr` tables carry, behind the cosine row and its 1/c, 1/c^2 rows, the three centred-stencil coefficient
// rows ap, am, bl of the tile kernel (filled on the host by qd_cos_companions with the same expressions).
// np.nan_to_num is applied LAZILY (the pattern of qd_ocean_fused.cuh): the reference cleans the field, lap(F) and the
// result; on finite data each of those is the identity.  The FAST instantiation (CLEAN = false) only records whether
// any value the reference would have cleaned was non-finite (one DSETP per value instead of the ~8-instruction select
// chain, which was 30 % of this kernel's issue slots: profiles/README.md); a warp that saw one re-runs its chunk with
// CLEAN = true.  The kernel is out of place, so the re-run reads unmodified inputs and rewrites the outputs:
// identical bits either way.
template <bool CLEAN>
__device__ __forceinline__ double qd_h4s_cl(double x, bool& bad) {
  if (CLEAN) return qd_clean_sel(x);
  bad = bad || !(fabs(x) <= DBL_MAX);
  return x;
}
template <int R, bool CLEAN>
__device__ __forceinline__ bool qd_h4s_chunk(const QdGeo& g, const QdHyper4Args& A, const int b, const int k, const int lane,
                                             const int strip, const int j0, const int j1) {
  const int nlat = g.nlat, nlon = g.nlon;
  const size_t off = (size_t)b * g.ncell;
  const double* __restrict__ cap = A.cosr + 3 * (size_t)nlat;
  const double* __restrict__ cam = A.cosr + 4 * (size_t)nlat;
  const double* __restrict__ cbl = A.cosr + 5 * (size_t)nlat;
  double sub_dt = A.dt;
  if (A.ocean) sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const double inner = sub_dt / (double)(A.nsub > 1 ? A.nsub : 1);
  // k4 of this chunk's rows: lane l holds rows j0+l (and j0+32+l); one division per lane instead of one per row
  double kq0, kq1 = 0.0;
  {
    const double* __restrict__ k4r = A.k4rows[k] + (size_t)b * A.k4_bstride[k];
    const double k4div = fmax(1e-12, sub_dt), k4scale = A.scale[k];
    const bool raw = A.raw_k4[k] != 0;
    auto k4row = [&](int j) {
      double v = (j < nlat) ? k4r[j] : 0.0;
      if (!raw) v = v / k4div;                              // ocean.py:347
      return k4scale * v;
    };
    kq0 = k4row(j0 + lane);
    if (R > 32) kq1 = k4row(j0 + 32 + lane);
  }
  int gi = strip * QD_H4S_COLS - 2 + lane;
  if (gi < 0) gi += nlon;
  if (gi >= nlon) gi -= nlon;
  if (gi >= nlon) gi -= nlon;
  const bool writer = lane >= 2 && lane < 2 + QD_H4S_COLS && (strip * QD_H4S_COLS + lane - 2) < nlon;
  const double* __restrict__ p = A.src[k] + off + (size_t)(j0 - 4) * nlon + gi;
  double* __restrict__ d = A.dst[k] + off + (size_t)j0 * nlon + gi;
  bool bad = false;
  auto lap_row = [&](double fm2, double fc, double fp2, double ap, double am, double bl) {
    const double fe = __shfl_down_sync(0xffffffffu, fc, 1), fw = __shfl_up_sync(0xffffffffu, fc, 1);
    return qd_h4s_cl<CLEAN>((ap * (fp2 - fc) - am * (fc - fm2)) + bl * ((fe - 2.0 * fc) + fw), bad);
  };
  // F window f0..f3 = F[j..j+3]; lap window l0..l3 = lap(F)[j-2..j+1]; c?m / c?0 = coefficients of rows j, j+1
  double f0 = qd_h4s_cl<CLEAN>(p[0], bad), f1 = qd_h4s_cl<CLEAN>(p[nlon], bad), f2 = qd_h4s_cl<CLEAN>(p[2 * (size_t)nlon], bad),
         f3 = qd_h4s_cl<CLEAN>(p[3 * (size_t)nlon], bad);
  p += 4 * (size_t)nlon;
  double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
  double apm = 0.0, amm = 0.0, blm = 0.0, ap0 = 0.0, am0 = 0.0, bl0 = 0.0;
  // One group = 4 rows: the four loads of new F rows are issued together at the top of the group so that their
  // latency overlaps; the twelve coefficient loads of a group are warp-uniform L1 hits issued next to their use.
  // Unrolling by the window length turns the register shifts into renames.  jr = row of the first new Laplacian value
  // of the group; emit = false for the warm-up group.  Measured and dropped (profiles/README.md): loading group q+1
  // before computing group q (96 registers, 5 blocks per SM: slower), and staging rows 8 / 12 ahead in a per-warp
  // cp.async shared-memory ring (64 -> 74 us for the ocean's three fields).
#define QD_H4S_LOADS(more)                                                                           \
    const double n0 = p[0], n1 = p[nlon], n2 = p[2 * (size_t)nlon], n3 = p[3 * (size_t)nlon];        \
    p += 4 * (size_t)nlon;
#define QD_H4S_ROTATE
#define QD_H4S_GROUP(jr, emit, more)                                                                 \
  {                                                                                                  \
    QD_H4S_LOADS(more)                                                                               \
    QD_H4S_STEP(n0, cap[(jr)], cam[(jr)], cbl[(jr)], 0, emit)                                        \
    QD_H4S_STEP(n1, cap[(jr) + 1], cam[(jr) + 1], cbl[(jr) + 1], 1, emit)                            \
    QD_H4S_STEP(n2, cap[(jr) + 2], cam[(jr) + 2], cbl[(jr) + 2], 2, emit)                            \
    QD_H4S_STEP(n3, cap[(jr) + 3], cam[(jr) + 3], cbl[(jr) + 3], 3, emit)                            \
    QD_H4S_ROTATE                                                                                    \
  }
#define QD_H4S_STEP(nv, apx, amx, blx, s, emit)                                                      \
  {                                                                                                  \
    const double ap = (apx), am = (amx), bl = (blx);       /* warp-uniform L1 hits, loaded next to their use */ \
    const double f4 = qd_h4s_cl<CLEAN>(nv, bad);                                                     \
    const double l4 = lap_row(f0, f2, f4, ap, am, bl);                                               \
    if (emit) {                                                                                      \
      const int r = j - j0 + (s);                                                                    \
      const double le = __shfl_down_sync(0xffffffffu, l2, 1), lw = __shfl_up_sync(0xffffffffu, l2, 1); \
      const double L2 = (apm * (l4 - l2) - amm * (l2 - l0)) + blm * ((le - 2.0 * l2) + lw);          \
      const double k4 = __shfl_sync(0xffffffffu, (R > 32 && r >= 32) ? kq1 : kq0, r & 31);           \
      const double o = qd_h4s_cl<CLEAN>(f0 - k4 * L2 * inner, bad);                                  \
      if (writer && j + (s) < j1) d[(size_t)(s) * nlon] = o;                                         \
    }                                                                                                \
    l0 = l1; l1 = l2; l2 = l3; l3 = l4;                                                              \
    apm = ap0; amm = am0; blm = bl0; ap0 = ap; am0 = am; bl0 = bl;                                   \
    f0 = f1; f1 = f2; f2 = f3; f3 = f4;                                                              \
  }
  {
    const int j = j0;
    QD_H4S_GROUP(j0 - 2, false, true)                      // warm-up: lap(F) at rows j0-2 .. j0+1
  }
  for (int j = j0; j < j1; j += 4) {
    QD_H4S_GROUP(j + 2, true, j + 4 < j1)
    d += 4 * (size_t)nlon;
  }
#undef QD_H4S_GROUP
#undef QD_H4S_STEP
#undef QD_H4S_LOADS
#undef QD_H4S_ROTATE
  return bad;
}
#ifndef QD_H4S_MINB
#define QD_H4S_MINB 6
#endif
template <int R>
__global__ void __launch_bounds__(32 * QD_H4S_WARPS, QD_H4S_MINB) k_hyper4_stream(QdGeo g, QdHyper4Args A) {
  static_assert(R == 16 || R == 32 || R == 64, "k4 rows are staged in one or two registers per lane");
  const int b = blockIdx.y;
  if (A.ocean && qd_sub_done(g, b, A.sc)) return;
  const int lane = threadIdx.x & 31;
  const int nstrips = (g.nlon + QD_H4S_COLS - 1) / QD_H4S_COLS;
  const int w = blockIdx.x * QD_H4S_WARPS + (threadIdx.x >> 5);
  const int chunk = w / nstrips, strip = w - chunk * nstrips;
  const int j0 = A.ja + chunk * R;
  if (j0 >= A.jb) return;
  const int j1 = min(j0 + R, A.jb);
  const int k = blockIdx.z;
  const bool bad = qd_h4s_chunk<R, false>(g, A, b, k, lane, strip, j0, j1);
  if (__any_sync(0xffffffffu, bad)) {                      // a non-finite value: redo the chunk with np.nan_to_num applied
    __syncwarp();
    qd_h4s_chunk<R, true>(g, A, b, k, lane, strip, j0, j1);
  }
}
#else
// Host check build: the same tile algorithm with cooperative loops written so that one sequential
// "thread" can run a whole phase (see qd_rt.h); the GPU kernel above is exercised by tests/test_gpu.py.
template <int TJ>
__global__ void __launch_bounds__(QD_H4_NX * QD_H4_NY) k_hyper4_tile(QdGeo g, QdHyper4Args A) {
  constexpr int TI = QD_H4_TI;
  constexpr int RF = TJ + 8, CF = TI + 4, RL = TJ + 4, CL = TI + 2;
  __shared__ double Fs[RF * CF];
  __shared__ double Ls[RL * CL];
  __shared__ double cap[RF], cam[RF], cbl[RF], k4s[TJ];
  const int b = blockIdx.y;
  if (A.ocean && qd_sub_done(g, b, A.sc)) return;
  const int tiles_i = (g.nlon + TI - 1) / TI;
  const int tj0 = blockIdx.x / tiles_i, ti = blockIdx.x - tj0 * tiles_i;
  const int tj = tj0 >= A.tj_lo ? tj0 + A.tj_skip : tj0;
  const int j0 = A.row0 + tj * TJ, i0 = ti * TI;
  const int jF0 = j0 - 4, jL0 = j0 - 2, iF0 = i0 - 2;
  const int nlat = g.nlat, nlon = g.nlon;
  const size_t off = (size_t)b * g.ncell;
  double sub_dt = A.dt;
  if (A.ocean) sub_dt = g.scal[(size_t)b * QD_S_COUNT + QD_S_SUB_DT];
  const double inner = sub_dt / (double)(A.nsub > 1 ? A.nsub : 1);
  const int k = blockIdx.z;                   // one field per block (gridDim.z = number of fields)
  const double* F = A.src[k] + off;
  const double* cr = A.cosr;
  QD_TILE_FIRST_ROWS(r, RF) {
    const int gj = jF0 + r;
    double ap = 0.0, am = 0.0, bl = 0.0;
    if (gj >= 2 && gj <= nlat - 3) {
      const double icj = cr[nlat + gj] * g.inv_2dlat;
      ap = (icj * (cr[gj + 1] * g.inv_2dlat)) * g.inv_a_sq;
      am = (icj * (cr[gj - 1] * g.inv_2dlat)) * g.inv_a_sq;
      bl = (cr[2 * nlat + gj] * g.inv_dlon_sq) * g.inv_a_sq;
    }
    cap[r] = ap; cam[r] = am; cbl[r] = bl;
    if (r < TJ) {
      const int oj = j0 + r;
      double k4 = 0.0;
      if (oj < nlat) {
        k4 = A.k4rows[k][(size_t)b * A.k4_bstride[k] + oj];
        if (!A.raw_k4[k]) k4 = k4 / fmax(1e-12, sub_dt);       // ocean.py:347
        k4 = A.scale[k] * k4;
      }
      k4s[r] = k4;
    }
  }
  QD_TILE_FIRST_LOOP(r, RF, cc, CF) {
    const int gj = jF0 + r;
    double v = 0.0;
    if (gj >= 0 && gj < nlat) {
      int gi = iF0 + cc;
      while (gi < 0) gi += nlon;
      while (gi >= nlon) gi -= nlon;
      v = qd_clean_fast(F[(size_t)gj * nlon + gi]);
    }
    Fs[r * CF + cc] = v;
  }
  __syncthreads();
  auto AF = [&](int jj, int c) -> double { return Fs[(jj - jF0) * CF + c]; };
  QD_TILE_FIRST_LOOP(r, RL, cc, CL) {
    const int gj = jL0 + r;
    double L = 0.0;
    if (gj >= 0 && gj < nlat) {
      const int fr = r + 2, e = fr * CF + cc + 1;
      if (gj >= 2 && gj <= nlat - 3) {
        const double f0 = Fs[e];
        L = (cap[fr] * (Fs[e + 2 * CF] - f0) - cam[fr] * (f0 - Fs[e - 2 * CF])) + cbl[fr] * ((Fs[e + 1] - 2.0 * f0) + Fs[e - 1]);
      } else {
        L = qd_lap_rel(AF, gj, cc + 1, g, cr);
      }
      L = qd_clean_fast(L);
    }
    Ls[r * CL + cc] = L;
  }
  __syncthreads();
  auto AL = [&](int jj, int c) -> double { return Ls[(jj - jL0) * CL + c]; };
  const int cc = threadIdx.x;
  const int gi = i0 + cc;
  if (gi < nlon) {
#pragma unroll 4
    for (int r = threadIdx.y; r < TJ; r += QD_H4_NY) {
      const int gj = j0 + r;
      if (gj < A.row1) {
        const int lr = r + 2, e = lr * CL + cc + 1, fr = r + 4;
        double L2;
        if (gj >= 2 && gj <= nlat - 3) {
          const double l0 = Ls[e];
          L2 = (cap[fr] * (Ls[e + 2 * CL] - l0) - cam[fr] * (l0 - Ls[e - 2 * CL])) + cbl[fr] * ((Ls[e + 1] - 2.0 * l0) + Ls[e - 1]);
        } else {
          L2 = qd_lap_rel(AL, gj, cc + 1, g, cr);
        }
        const double cur = Fs[fr * CF + cc + 2];
        A.dst[k][off + (size_t)gj * nlon + gi] = qd_clean_fast(cur - k4s[r] * L2 * inner);
      }
    }
  }
}
#endif
