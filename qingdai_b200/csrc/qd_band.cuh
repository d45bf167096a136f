// qd_band.cuh -- latitude-band decomposition of ONE large domain over the GPUs of a node
// (BASELINE configs[4]): device side of the halo exchange and of the cross-rank reductions.
//
// Storage is replicated, compute is partitioned: every rank holds full-size fields (1.5 GB at 1441x2880,
// <1 % of a B200's HBM) but computes only its own rows [own0, own1) plus as many halo rows as its inputs are
// valid for (qd_api.cu: band_prep keeps a "valid halo width" per field; stencils and gathers consume it,
// an exchange refills it to H rows).  Full longitude circles stay on one rank, so the periodic-longitude
// stencils, the zonal band-stop DFT, the Shapiro/Gaussian longitude passes and the polar ring means are
// rank-local; the ring of ranks is closed (rank G-1 <-> rank 0) because the reference's map_coordinates
// (mode='wrap') and np.roll(axis=0) wrap over the poles.
//
// Transport: each rank owns one exchange buffer (cudaMalloc + cudaIpc handle; POSIX shared memory in the
// host check build) that its peers map.  A sender WRITES its boundary rows / partial sums / histograms
// straight into the receiver's inbox over NVLink and then raises an epoch flag there; the receiver spins
// on its own flag and unpacks.  Everything is ordinary kernels on the step's stream (graph-capturable, no
// host round trip, no NCCL on the data path).  Inboxes are double-buffered by epoch parity: a sender can
// be at most one collective ahead of a receiver.  Spins are bounded and raise an error word instead of
// hanging the GPU.
#pragma once
#include "qd_ops.cuh"
#if QD_EMU
#include <sched.h>
#endif

// sizes of the exact-median kernel (qd_select.cuh), shared with the mailboxes below
#define QD_SEL_PASSES 6
#define QD_SEL_MAXBINS 2048
#define QD_SEL_THREADS 512
#define QD_SEL_CAP 2048          // candidates finished by an in-block sort instead of further radix passes

#define QD_BAND_MAXW 8            // ranks (GPUs of one node)
#define QD_BAND_MAXX 10           // fields per halo exchange
#define QD_BAND_MAXR 8            // scalars per all-reduce
#define QD_BAND_SPIN (1u << 23)   // bounded spin (~ seconds) before the error word is raised

// flag words inside a rank's buffer (unsigned long long each)
enum { QD_BF_HALO_S = 0, QD_BF_HALO_N = 1, QD_BF_RED = 8, QD_BF_SEL = 16, QD_BF_ERR = 24, QD_BF_EPOCH_HALO = 32,
       QD_BF_EPOCH_RED = 33, QD_BF_EPOCH_SEL = 34, QD_BF_EPOCH_PUB = 35, QD_BF_PUB_FLAG = 36, QD_BF_TICKET = 40,
       QD_BF_PUB_VAL = 48 /* two doubles, by epoch parity */, QD_BF_WORDS = 64 };

struct QdBandCtl {
  int rank, world, H, nlon, nlat;
  char* peer[QD_BAND_MAXW];                 // base of every rank's buffer as mapped here (peer[rank] = own)
  unsigned long long off_inbox, off_red, off_hist, off_list, off_emu, off_sflag;   // byte offsets, identical on every rank
};
QD_HD unsigned long long* qd_bflags(const QdBandCtl& B, int r) { return (unsigned long long*)B.peer[r]; }
QD_HD double* qd_binbox(const QdBandCtl& B, int r, int parity, int dir, int slot) {
  return (double*)(B.peer[r] + B.off_inbox) + (((size_t)parity * 2 + dir) * QD_BAND_MAXX + slot) * (size_t)B.H * B.nlon;
}
QD_HD double* qd_bred(const QdBandCtl& B, int r, int parity, int src) {
  return (double*)(B.peer[r] + B.off_red) + ((size_t)parity * QD_BAND_MAXW + src) * QD_BAND_MAXR;
}
QD_HD unsigned* qd_bhist(const QdBandCtl& B, int r, int parity, int src) {
  return (unsigned*)(B.peer[r] + B.off_hist) + ((size_t)parity * QD_BAND_MAXW + src) * QD_SEL_MAXBINS;
}
QD_HD unsigned long long* qd_blist(const QdBandCtl& B, int r, int parity, int src) {   // [0] count, [1] mingt, [2..] keys
  return (unsigned long long*)(B.peer[r] + B.off_list) + ((size_t)parity * QD_BAND_MAXW + src) * (QD_SEL_CAP + 2);
}

#if QD_EMU
static inline void qd_fence_sys() { __sync_synchronize(); }
static inline unsigned long long qd_ld_sys(const unsigned long long* p) { __sync_synchronize(); return *(volatile const unsigned long long*)p; }
static inline void qd_st_sys(unsigned long long* p, unsigned long long v) { __sync_synchronize(); *(volatile unsigned long long*)p = v; __sync_synchronize(); }
#else
__device__ __forceinline__ void qd_fence_sys() { __threadfence_system(); }
__device__ __forceinline__ unsigned long long qd_ld_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void qd_st_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif
// spin until *flag >= epoch; returns false (and raises the error word) when the bound is hit
QD_D bool qd_band_wait(const QdBandCtl& B, const unsigned long long* flag, unsigned long long epoch) {
  if (qd_bflags(B, B.rank)[QD_BF_ERR] != 0ull) return false;             // already failed: do not stack timeouts
  for (unsigned n = 0; n < QD_BAND_SPIN; ++n) {
    if (qd_ld_sys(flag) >= epoch) return true;
#if QD_EMU
    if ((n & 1023u) == 1023u) sched_yield();
#endif
  }
  qd_bflags(B, B.rank)[QD_BF_ERR] = 1ull;
  return false;
}

// ---------------------------------------------------------------------------------------------- halo rows
struct QdBandList { int n; double* f[QD_BAND_MAXX]; };     // member-0 base pointers of the fields to exchange

// ONE kernel per exchange, grid (gx slices, n fields, 2 directions), every block resident (gx * n * 2 <= one wave).
// Block (x, k, dir) stores slice x of field k's boundary rows into the neighbour on side `dir` (dir 0: my lowest H
// rows to my SOUTH neighbour, dir 1: my top H rows to my NORTH neighbour), publishes the slice's own epoch flag
// there, then waits for the matching slice from the opposite neighbour and copies it next to my own rows (mod
// n_lat: the ring is closed over the poles).  No grid-wide dependency: a slice is unpacked as soon as it arrived.
#define QD_BAND_GX 24
QD_HD unsigned long long* qd_bslice_flag(const QdBandCtl& B, int r, int dir, int k, int x) {
  return (unsigned long long*)(B.peer[r] + B.off_sflag) + ((size_t)dir * QD_BAND_MAXX + k) * QD_BAND_GX + x;
}
// phase: 1 = push, 2 = receive, 3 = both (GPU; the sequential host check build launches 1 then 2: its blocks do not overlap)
__global__ void __launch_bounds__(QD_THREADS) k_band_exchange(QdBandCtl B, QdBandList L, int own0, int own1, int phase) {
  const int x = blockIdx.x, k = blockIdx.y, dir = blockIdx.z;
  unsigned long long* mine = qd_bflags(B, B.rank);
  const unsigned long long epoch = mine[QD_BF_EPOCH_HALO] + 1ull;      // bumped by the last block to FINISH: all are resident, all read it first
  const int parity = (int)(epoch & 1ull);
  const int south = (B.rank + B.world - 1) % B.world, north = (B.rank + 1) % B.world;
  const int n = B.H * B.nlon;
  const int per = ((n + gridDim.x - 1) / gridDim.x + 1) & ~1;           // slice length, even (16-byte stores)
  const int e0 = x * per, e1 = (e0 + per < n) ? e0 + per : n;
  // ---- push: (dir 0) rows [own0, own0+H) -> south neighbour's "from north" inbox; (dir 1) rows [own1-H, own1) -> north's "from south"
  if (phase & 1) {
    const int nbr = dir == 0 ? south : north, rdir = dir == 0 ? 1 : 0;
    const double* src = L.f[k] + (size_t)(dir == 0 ? own0 : own1 - B.H) * B.nlon;
    double* dst = qd_binbox(B, nbr, parity, rdir, k);
#if !QD_EMU
    if ((((size_t)src | (size_t)dst) & 15) == 0 && (e1 & 1) == 0) {
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2 = reinterpret_cast<double2*>(dst);
      for (int e = e0 / 2 + threadIdx.x; e < e1 / 2; e += blockDim.x) d2[e] = s2[e];
    } else
#endif
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) dst[e] = src[e];
#if !QD_EMU
    __syncthreads();
#endif
    QD_BLOCK_LAST_ONE { qd_fence_sys(); qd_st_sys(qd_bslice_flag(B, nbr, rdir, k, x), epoch); }
  }
  // ---- receive: slice x of field k arriving on side `dir` (dir 0: from the south neighbour -> rows below mine)
  if (phase & 2) {
#if QD_EMU
    if (threadIdx.x == 0) qd_band_wait(B, qd_bslice_flag(B, B.rank, dir, k, x), epoch);
#else
    if (threadIdx.x == 0) qd_band_wait(B, qd_bslice_flag(B, B.rank, dir, k, x), epoch);
    __syncthreads();
#endif
    int first = dir == 0 ? own0 - B.H : own1;               // H <= rows of any rank: the block never straddles the wrap
    if (first < 0) first += B.nlat;
    if (first >= B.nlat) first -= B.nlat;
    const double* src = qd_binbox(B, B.rank, parity, dir, k);
    double* dst = L.f[k] + (size_t)first * B.nlon;
#if !QD_EMU
    if ((((size_t)src | (size_t)dst) & 15) == 0 && (e1 & 1) == 0) {
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2 = reinterpret_cast<double2*>(dst);
      for (int e = e0 / 2 + threadIdx.x; e < e1 / 2; e += blockDim.x) d2[e] = __ldcg(s2 + e);
    } else
      for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) dst[e] = __ldcg(src + e);
#else
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) dst[e] = src[e];
#endif
  }
  const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
  if ((phase & 2) && qd_block_is_last((unsigned*)(mine + QD_BF_TICKET), nblocks)) { QD_BLOCK_LAST_ONE { mine[QD_BF_EPOCH_HALO] = epoch; } }
}

// ---------------------------------------------------------------------------------------------- scalars
// All-reduce (sum or max) of up to QD_BAND_MAXR entries of the scalar table: every rank writes its partial
// into every rank's mailbox, then combines the world's partials in rank order (identical bits everywhere).
struct QdBandRed { int n; int id[QD_BAND_MAXR]; int is_max[QD_BAND_MAXR]; };
__global__ void k_band_allreduce(QdBandCtl B, QdBandRed R, double* scal) {
  unsigned long long* mine = qd_bflags(B, B.rank);
  const unsigned long long epoch = mine[QD_BF_EPOCH_RED] + 1ull;
  const int parity = (int)(epoch & 1ull);
  const int t = threadIdx.x;
#if QD_EMU
  if (t != 0) return;
  for (int q = 0; q < R.n; ++q) for (int r = 0; r < B.world; ++r) qd_bred(B, r, parity, B.rank)[q] = scal[R.id[q]];
  qd_fence_sys();
  mine[QD_BF_EPOCH_RED] = epoch;
  for (int r = 0; r < B.world; ++r) qd_st_sys(qd_bflags(B, r) + QD_BF_RED + B.rank, epoch);
  for (int r = 0; r < B.world; ++r) qd_band_wait(B, mine + QD_BF_RED + r, epoch);
  for (int q = 0; q < R.n; ++q) {
    double acc = qd_bred(B, B.rank, parity, 0)[q];
    for (int r = 1; r < B.world; ++r) { const double v = qd_bred(B, B.rank, parity, r)[q]; acc = R.is_max[q] ? (v > acc ? v : acc) : acc + v; }
    scal[R.id[q]] = acc;
  }
#else
  if (t < R.n) { const double v = scal[R.id[t]]; for (int r = 0; r < B.world; ++r) qd_bred(B, r, parity, B.rank)[t] = v; }
  qd_fence_sys();
  __syncthreads();
  if (t == 0) mine[QD_BF_EPOCH_RED] = epoch;
  if (t < B.world) { qd_st_sys(qd_bflags(B, t) + QD_BF_RED + B.rank, epoch); qd_band_wait(B, mine + QD_BF_RED + t, epoch); }
  __syncthreads();
  if (t < R.n) {
    double acc = __ldcg(qd_bred(B, B.rank, parity, 0) + t);
    for (int r = 1; r < B.world; ++r) { const double v = __ldcg(qd_bred(B, B.rank, parity, r) + t); acc = R.is_max[t] ? (v > acc ? v : acc) : acc + v; }
    scal[R.id[t]] = acc;
  }
#endif
}

// ---- one scalar, published and pulled: the producer kernel's last block stores its partial into its OWN buffer, raises
// a flag, then reads the world's partials over NVLink (peers map the buffer) and adds them in rank order.  No separate
// all-reduce kernel between producer and consumer (the ocean's eta sum: once per CFL sub-step).  Values are double-buffered by epoch parity: a rank can publish epoch e+2 only after every peer has
// consumed epoch e (it had to read their e+1 first).
QD_D void qd_band_publish(const QdBandCtl& B, double v) {
  unsigned long long* mine = qd_bflags(B, B.rank);
  const unsigned long long epoch = mine[QD_BF_EPOCH_PUB] + 1ull;
  ((double*)(mine + QD_BF_PUB_VAL))[epoch & 1ull] = v;
  qd_fence_sys();
  mine[QD_BF_EPOCH_PUB] = epoch;
  qd_st_sys(mine + QD_BF_PUB_FLAG, epoch);
}
// called by ONE thread of a block; the rank's own epoch word already counts the publish that precedes this kernel
QD_D double qd_band_pull_sum(const QdBandCtl& B) {
#if QD_EMU
  const unsigned long long epoch = *(volatile unsigned long long*)(qd_bflags(B, B.rank) + QD_BF_EPOCH_PUB);
#else
  const unsigned long long epoch = __ldcg(qd_bflags(B, B.rank) + QD_BF_EPOCH_PUB);
#endif
  double acc = 0.0;
  for (int r = 0; r < B.world; ++r) {
    unsigned long long* fl = qd_bflags(B, r);
    qd_band_wait(B, fl + QD_BF_PUB_FLAG, epoch);
#if QD_EMU
    __sync_synchronize();
    const double v = ((volatile double*)(fl + QD_BF_PUB_VAL))[epoch & 1ull];
#else
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"((double*)(fl + QD_BF_PUB_VAL) + (epoch & 1ull)) : "memory");
#endif
    acc = r == 0 ? v : acc + v;
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------- selection
// Block-cooperative pieces used by block 0 of k_select_coop when the field is split over ranks.
#if !QD_EMU
// Collectives of the selection kernel.  `epoch` = (value of the rank's selection epoch word at kernel start) +
// (index of the collective inside this launch); every block computes it the same way, block 0 stores the last one
// back when the kernel ends.  Blocks 0..world-1 each serve one peer, block 0 combines.
//
// gh[0..nb) <- sum over ranks of their gh (element-wise, exact integers).  Call after a grid.sync (gh complete);
// the caller grid.syncs afterwards.
__device__ __forceinline__ void qd_band_hist_allreduce(const QdBandCtl& B, unsigned* gh, int nb, unsigned long long epoch) {
  unsigned long long* mine = qd_bflags(B, B.rank);
  const int parity = (int)(epoch & 1ull);
  for (int r = blockIdx.x; r < B.world; r += gridDim.x) {               // nb is a multiple of 4, boxes are 256-byte aligned
    uint4* dst = reinterpret_cast<uint4*>(qd_bhist(B, r, parity, B.rank));
    const uint4* src = reinterpret_cast<const uint4*>(gh);
    for (int k = threadIdx.x; k < nb / 4; k += blockDim.x) dst[k] = __ldcg(src + k);
    __syncthreads();
    if (threadIdx.x == 0) { qd_fence_sys(); qd_st_sys(qd_bflags(B, r) + QD_BF_SEL + B.rank, epoch); }
  }
  if (blockIdx.x != 0) return;
  if (threadIdx.x < B.world) qd_band_wait(B, mine + QD_BF_SEL + threadIdx.x, epoch);
  __syncthreads();
  for (int k = threadIdx.x; k < nb / 4; k += blockDim.x) {
    uint4 s = make_uint4(0u, 0u, 0u, 0u);
    for (int r = 0; r < B.world; ++r) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(qd_bhist(B, B.rank, parity, r)) + k);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<uint4*>(gh)[k] = s;
  }
  __threadfence();
}
// every rank's candidate list + "smallest key above the bucket" -> all ranks.  Blocks 0..world-1 push (call from
// every block after the grid.sync that completes the local list); block 0 returns the merged count with the keys in
// skeys[0..m) (shared memory, unsorted) and the global mingt; other blocks return -1.
__device__ __forceinline__ int qd_band_list_allgather(const QdBandCtl& B, const unsigned long long* lst, unsigned cnt, unsigned long long mingt_local,
                                                      unsigned long long* skeys, unsigned long long* mingt_out, unsigned long long epoch) {
  unsigned long long* mine = qd_bflags(B, B.rank);
  const int parity = (int)(epoch & 1ull);
  if (cnt > QD_SEL_CAP) cnt = QD_SEL_CAP;
  for (int r = blockIdx.x; r < B.world; r += gridDim.x) {
    unsigned long long* dst = qd_blist(B, r, parity, B.rank);
    if (threadIdx.x == 0) { dst[0] = cnt; dst[1] = mingt_local; }
    for (int k = threadIdx.x; k < (int)cnt; k += blockDim.x) dst[2 + k] = __ldcg(lst + k);
    __syncthreads();
    if (threadIdx.x == 0) { qd_fence_sys(); qd_st_sys(qd_bflags(B, r) + QD_BF_SEL + B.rank, epoch); }
  }
  if (blockIdx.x != 0) return -1;
  if (threadIdx.x < B.world) qd_band_wait(B, mine + QD_BF_SEL + threadIdx.x, epoch);
  __syncthreads();
  int m = 0;
  unsigned long long mg = ~0ull;
  for (int r = 0; r < B.world; ++r) {
    const unsigned long long* src = qd_blist(B, B.rank, parity, r);
    const int n = (int)__ldcg(src);
    const unsigned long long g2 = __ldcg(src + 1);
    if (g2 < mg) mg = g2;
    for (int k = threadIdx.x; k < n && m + k < QD_SEL_CAP; k += blockDim.x) skeys[m + k] = __ldcg(src + 2 + k);
    m += n;
  }
  __syncthreads();
  *mingt_out = mg;
  return m < QD_SEL_CAP ? m : QD_SEL_CAP;
}
#endif
