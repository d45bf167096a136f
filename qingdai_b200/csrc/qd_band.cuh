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
// straight into the receiver's inbox over NVLink as flagged 16-byte lines (value + arrival flag in one posted
// store, see qd_ll_store below); the receiver polls its own inbox and takes each value as it lands.
// Everything is ordinary kernels on the step's stream (graph-capturable, no host round trip, no NCCL on the
// data path).  Inboxes are double-buffered by epoch parity: a sender can be at most one collective ahead of
// a receiver.  Spins are bounded and raise an error word instead of hanging the GPU.
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
QD_HD double* qd_bred(const QdBandCtl& B, int r, int parity, int src) {
  return (double*)(B.peer[r] + B.off_red) + ((size_t)parity * QD_BAND_MAXW + src) * QD_BAND_MAXR;
}
struct QdLine;
QD_HD QdLine* qd_bhist(const QdBandCtl& B, int r, int parity, int src);      // QD_SEL_MAXBINS lines of two bins each (two histograms)
QD_HD QdLine* qd_blist(const QdBandCtl& B, int r, int parity, int src);      // [0] count, [1] mingt, [2..] keys: one line each

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

// ---- flagged lines ("LL" protocol, the low-latency form NCCL uses): one double travels as a 16-byte line
// {lo32, flag32, hi32, flag32}; each 8-byte half is written atomically, so a reader that sees the expected flag in
// BOTH halves has the whole value.  Data and "it has arrived" are ONE posted store: no fence, no separate flag
// store, no second NVLink round trip; the receiver polls its OWN memory.  Lines are double-buffered by epoch parity
// and the flag is the (never zero) low word of the epoch, so a stale line can never be mistaken for a fresh one.
struct QdLine { unsigned long long w[2]; };
#if QD_EMU
static inline void qd_ll_store(QdLine* l, double v, unsigned flag) {
  unsigned long long u; memcpy(&u, &v, 8);
  const unsigned long long a = (u & 0xffffffffull) | ((unsigned long long)flag << 32), b = (u >> 32) | ((unsigned long long)flag << 32);
  __sync_synchronize();
  *(volatile unsigned long long*)&l->w[0] = a; *(volatile unsigned long long*)&l->w[1] = b;
  __sync_synchronize();
}
static inline bool qd_ll_load(const QdLine* l, unsigned flag, double* v) {
  __sync_synchronize();
  const unsigned long long a = *(volatile const unsigned long long*)&l->w[0], b = *(volatile const unsigned long long*)&l->w[1];
  if ((unsigned)(a >> 32) != flag || (unsigned)(b >> 32) != flag) return false;
  const unsigned long long u = (a & 0xffffffffull) | (b << 32);
  memcpy(v, &u, 8);
  return true;
}
#else
__device__ __forceinline__ void qd_ll_store(QdLine* l, double v, unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(l), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ bool qd_ll_load(const QdLine* l, unsigned flag, double* v) {
  unsigned a, fa, b, fb;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(l) : "memory");
  if (fa != flag || fb != flag) return false;
  *v = __hiloint2double((int)b, (int)a);
  return true;
}
#endif
#if !QD_EMU
// the same line carrying two 32-bit words (histogram bins) or one 64-bit key
__device__ __forceinline__ void qd_ll_store2(QdLine* l, unsigned a, unsigned b, unsigned flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(l), "r"(a), "r"(flag), "r"(b), "r"(flag) : "memory");
}
__device__ __forceinline__ bool qd_ll_load2(const QdLine* l, unsigned flag, unsigned* a, unsigned* b) {
  unsigned fa, fb;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(*a), "=r"(fa), "=r"(*b), "=r"(fb) : "l"(l) : "memory");
  return fa == flag && fb == flag;
}
#endif
// spin until the line carries `flag`; false (and the error word) when the bound is hit
QD_D bool qd_ll_wait(const QdBandCtl& B, const QdLine* l, unsigned flag, double* v) {
  if (qd_bflags(B, B.rank)[QD_BF_ERR] != 0ull) { *v = 0.0; return false; }
  for (unsigned n = 0; n < QD_BAND_SPIN; ++n) {
    if (qd_ll_load(l, flag, v)) return true;
#if QD_EMU
    if ((n & 1023u) == 1023u) sched_yield();
#endif
  }
  qd_bflags(B, B.rank)[QD_BF_ERR] = 1ull;
  *v = 0.0;
  return false;
}
QD_HD unsigned qd_ll_flag(unsigned long long epoch) { const unsigned f = (unsigned)epoch; return f ? f : 0x80000000u; }
#if !QD_EMU
__device__ __forceinline__ bool qd_ll_wait2(const QdBandCtl& B, const QdLine* l, unsigned flag, unsigned* a, unsigned* b) {
  if (qd_bflags(B, B.rank)[QD_BF_ERR] != 0ull) { *a = *b = 0u; return false; }
  for (unsigned n = 0; n < QD_BAND_SPIN; ++n) if (qd_ll_load2(l, flag, a, b)) return true;
  qd_bflags(B, B.rank)[QD_BF_ERR] = 1ull;
  *a = *b = 0u;
  return false;
}
#endif
QD_HD QdLine* qd_bhist(const QdBandCtl& B, int r, int parity, int src) {
  return (QdLine*)(B.peer[r] + B.off_hist) + ((size_t)parity * QD_BAND_MAXW + src) * QD_SEL_MAXBINS;
}
QD_HD QdLine* qd_blist(const QdBandCtl& B, int r, int parity, int src) {
  return (QdLine*)(B.peer[r] + B.off_list) + ((size_t)parity * QD_BAND_MAXW + src) * (QD_SEL_CAP + 2);
}
// scalar mailboxes: line [set][parity][src][slot]; set 0 = k_band_allreduce, set 1 = publish / pull
QD_HD QdLine* qd_bline(const QdBandCtl& B, int r, int set, int parity, int src, int slot) {
  return (QdLine*)(B.peer[r] + B.off_red) + (((size_t)set * 2 + parity) * QD_BAND_MAXW + src) * QD_BAND_MAXR + slot;
}

// ---------------------------------------------------------------------------------------------- halo rows
struct QdBandList { int n; double* f[QD_BAND_MAXX]; };     // member-0 base pointers of the fields to exchange

// ONE kernel per exchange, grid (gx slices, n fields, 2 directions), every block resident (gx * n * 2 <= one wave).
// Block (x, k, dir) sends slice x of field k's boundary rows to the neighbour on side `dir` (dir 0: my lowest H rows to
// my SOUTH neighbour, dir 1: my top H rows to my NORTH neighbour) as flagged lines -- value and arrival flag in one
// 16-byte posted store per element, straight into the neighbour's inbox over NVLink: no block barrier, no fence, no
// separate flag store on the sender -- then collects the matching slice from the opposite neighbour out of its OWN
// inbox, element by element as the lines land, and writes it next to my own rows (mod n_lat: the ring is closed over
// the poles).  Measured on 2 B200s (tools/band_micro.py, profiles/README.md) against the earlier copy + fence + slice
// flag + unpack form.  No grid-wide dependency; inboxes are double-buffered by epoch parity.
#define QD_BAND_GX 48
QD_HD QdLine* qd_binbox(const QdBandCtl& B, int r, int parity, int dir, int slot) {
  return (QdLine*)(B.peer[r] + B.off_inbox) + (((size_t)parity * 2 + dir) * QD_BAND_MAXX + slot) * (size_t)B.H * B.nlon;
}
// phase: 1 = push, 2 = receive, 3 = both (GPU; the sequential host check build launches 1 then 2: its blocks do not overlap)
__global__ void __launch_bounds__(QD_THREADS) k_band_exchange(QdBandCtl B, QdBandList L, int own0, int own1, int phase) {
  const int x = blockIdx.x, k = blockIdx.y, dir = blockIdx.z;
  unsigned long long* mine = qd_bflags(B, B.rank);
  const unsigned long long epoch = mine[QD_BF_EPOCH_HALO] + 1ull;      // bumped by the last block to FINISH: all are resident, all read it first
  const int parity = (int)(epoch & 1ull);
  const unsigned flag = qd_ll_flag(epoch);
  const int south = (B.rank + B.world - 1) % B.world, north = (B.rank + 1) % B.world;
  const int n = B.H * B.nlon;
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int e0 = x * per, e1 = (e0 + per < n) ? e0 + per : n;
  // ---- push: (dir 0) rows [own0, own0+H) -> south neighbour's "from north" inbox; (dir 1) rows [own1-H, own1) -> north's "from south"
  if (phase & 1) {
    const int nbr = dir == 0 ? south : north, rdir = dir == 0 ? 1 : 0;
    const double* src = L.f[k] + (size_t)(dir == 0 ? own0 : own1 - B.H) * B.nlon;
    QdLine* dst = qd_binbox(B, nbr, parity, rdir, k);
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) qd_ll_store(dst + e, src[e], flag);
  }
  // ---- receive: slice x of field k arriving on side `dir` (dir 0: from the south neighbour -> rows below mine)
  if (phase & 2) {
    int first = dir == 0 ? own0 - B.H : own1;               // H <= rows of any rank: the block never straddles the wrap
    if (first < 0) first += B.nlat;
    if (first >= B.nlat) first -= B.nlat;
    const QdLine* src = qd_binbox(B, B.rank, parity, dir, k);
    double* dst = L.f[k] + (size_t)first * B.nlon;
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) { double v; qd_ll_wait(B, src + e, flag, &v); dst[e] = v; }
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
  const unsigned flag = qd_ll_flag(epoch);
  // thread (q, r): scalar q to / from rank r -- one flagged line each way, combined in rank order by thread (q, 0)
#if QD_EMU
  for (int q = 0; q < R.n; ++q) for (int r = 0; r < B.world; ++r) qd_ll_store(qd_bline(B, r, 0, parity, B.rank, q), scal[R.id[q]], flag);
  mine[QD_BF_EPOCH_RED] = epoch;
  for (int q = 0; q < R.n; ++q) {
    double acc = 0.0;
    for (int r = 0; r < B.world; ++r) {
      double v; qd_ll_wait(B, qd_bline(B, B.rank, 0, parity, r, q), flag, &v);
      acc = r == 0 ? v : (R.is_max[q] ? (v > acc ? v : acc) : acc + v);
    }
    scal[R.id[q]] = acc;
  }
#else
  __shared__ double got[QD_BAND_MAXR][QD_BAND_MAXW];
  const int q = threadIdx.x / QD_BAND_MAXW, r = threadIdx.x % QD_BAND_MAXW;
  __syncthreads();                                                      // every thread has read the epoch word
  if (q < R.n && r < B.world) qd_ll_store(qd_bline(B, r, 0, parity, B.rank, q), scal[R.id[q]], flag);
  if (threadIdx.x == 0) mine[QD_BF_EPOCH_RED] = epoch;
  if (q < R.n && r < B.world) { double v; qd_ll_wait(B, qd_bline(B, B.rank, 0, parity, r, q), flag, &v); got[q][r] = v; }
  __syncthreads();
  if (q < R.n && r == 0) {
    double acc = got[q][0];
    for (int k = 1; k < B.world; ++k) { const double v = got[q][k]; acc = R.is_max[q] ? (v > acc ? v : acc) : acc + v; }
    scal[R.id[q]] = acc;
  }
#endif
}

// ---- one scalar all-reduced in the tail of its producer kernel (the ocean's eta sum, once per CFL sub-step): the one
// thread that holds the rank's partial PUSHES it as a flagged line into every rank's mailbox (posted stores, all in
// flight together) and then collects the world's lines from its OWN mailbox, adding them in rank order (identical bits
// on every rank).  One NVLink one-way flight instead of the earlier publish-locally / poll-remotely scheme, whose
// remote polls were a round trip per rank.  Lines are double-buffered by epoch parity: a rank can publish epoch e+2
// only after it has collected every peer's e+1, which those peers sent after collecting its e.
QD_D void qd_band_publish(const QdBandCtl& B, double v) {
  unsigned long long* mine = qd_bflags(B, B.rank);
  const unsigned long long epoch = mine[QD_BF_EPOCH_PUB] + 1ull;
  const unsigned flag = qd_ll_flag(epoch);
  for (int r = 0; r < B.world; ++r) qd_ll_store(qd_bline(B, r, 1, (int)(epoch & 1ull), B.rank, 0), v, flag);
  mine[QD_BF_EPOCH_PUB] = epoch;
}
// called by the SAME thread right after qd_band_publish
QD_D double qd_band_pull_sum(const QdBandCtl& B) {
  const unsigned long long epoch = qd_bflags(B, B.rank)[QD_BF_EPOCH_PUB];
  const unsigned flag = qd_ll_flag(epoch);
  double acc = 0.0;
  for (int r = 0; r < B.world; ++r) {
    double v;
    qd_ll_wait(B, qd_bline(B, B.rank, 1, (int)(epoch & 1ull), r, 0), flag, &v);
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
// the caller grid.syncs afterwards.  Flagged lines of two bins each: block 0 posts this rank's histogram into every
// peer's mailbox, then adds the world's lines out of its own mailbox as they land.
__device__ __forceinline__ void qd_band_hist_allreduce(const QdBandCtl& B, unsigned* gh, int nb, unsigned long long epoch) {
  const int parity = (int)(epoch & 1ull);
  const unsigned flag = qd_ll_flag(epoch);
  if (blockIdx.x != 0) return;                                          // block 0 posts to every peer and combines: gh is only overwritten
  for (int r = 0; r < B.world; ++r) {                                   // after this block has read it for all of them (nb is even)
    QdLine* dst = qd_bhist(B, r, parity, B.rank);
    const uint2* src = reinterpret_cast<const uint2*>(gh);
    for (int k = threadIdx.x; k < nb / 2; k += blockDim.x) { const uint2 v = __ldcg(src + k); qd_ll_store2(dst + k, v.x, v.y, flag); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nb / 2; k += blockDim.x) {
    unsigned sx = 0u, sy = 0u;
    for (int r = 0; r < B.world; ++r) {
      unsigned a, c;
      qd_ll_wait2(B, qd_bhist(B, B.rank, parity, r) + k, flag, &a, &c);
      sx += a; sy += c;
    }
    reinterpret_cast<uint2*>(gh)[k] = make_uint2(sx, sy);
  }
  __threadfence();
}
// every rank's candidate list + "smallest key above the bucket" -> all ranks.  Blocks 0..world-1 push (call from
// every block after the grid.sync that completes the local list); block 0 returns the merged count with the keys in
// skeys[0..m) (shared memory, unsorted) and the global mingt; other blocks return -1.
__device__ __forceinline__ int qd_band_list_allgather(const QdBandCtl& B, const unsigned long long* lst, unsigned cnt, unsigned long long mingt_local,
                                                      unsigned long long* skeys, unsigned long long* mingt_out, unsigned long long epoch) {
  const int parity = (int)(epoch & 1ull);
  const unsigned flag = qd_ll_flag(epoch);
  if (cnt > QD_SEL_CAP) cnt = QD_SEL_CAP;
  for (int r = blockIdx.x; r < B.world; r += gridDim.x) {
    QdLine* dst = qd_blist(B, r, parity, B.rank);
    if (threadIdx.x == 0) { qd_ll_store2(dst, cnt, 0u, flag); qd_ll_store2(dst + 1, (unsigned)mingt_local, (unsigned)(mingt_local >> 32), flag); }
    for (int k = threadIdx.x; k < (int)cnt; k += blockDim.x) { const unsigned long long key = __ldcg(lst + k); qd_ll_store2(dst + 2 + k, (unsigned)key, (unsigned)(key >> 32), flag); }
  }
  if (blockIdx.x != 0) return -1;
  int m = 0;
  unsigned long long mg = ~0ull;
  for (int r = 0; r < B.world; ++r) {
    const QdLine* src = qd_blist(B, B.rank, parity, r);
    unsigned n, z, glo, ghi;
    qd_ll_wait2(B, src, flag, &n, &z);
    qd_ll_wait2(B, src + 1, flag, &glo, &ghi);
    const unsigned long long g2 = ((unsigned long long)ghi << 32) | glo;
    if (g2 < mg) mg = g2;
    for (int k = threadIdx.x; k < (int)n && m + k < QD_SEL_CAP; k += blockDim.x) {
      unsigned lo, hi;
      qd_ll_wait2(B, src + 2 + k, flag, &lo, &hi);
      skeys[m + k] = ((unsigned long long)hi << 32) | lo;
    }
    m += (int)n;
  }
  __syncthreads();
  *mingt_out = mg;
  return m < QD_SEL_CAP ? m : QD_SEL_CAP;
}
#endif
