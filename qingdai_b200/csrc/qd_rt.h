// qd_rt.h -- thin runtime layer under the kernels.
//
// Normal build (nvcc, sm_100a): plain CUDA.  Test build (-DQD_HOST_EMU, g++): the SAME kernel
// sources are compiled for the host and run one "thread" at a time so that indexing / operand
// order / boundary logic can be checked in CPU-only CI before GPU minutes are spent.  The host
// build is test scaffolding under tests/hostcheck; the product never loads it.
//
// Kernel authoring rules that make both builds behave identically:
//   * cooperative loops use QD_BLOCK_FIRST_FOR / QD_BLOCK_LAST_FOR (strided over the block on
//     the GPU; run by the first / last thread of the block in the sequential host build);
//   * block reductions go through qd_block_sum / qd_block_max, which hand the block total to
//     exactly one thread;
//   * a reduction result is consumed by a LATER kernel (or by the last block via qd_last_block).
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>
#include <float.h>

#ifdef QD_HOST_EMU
// ------------------------------------------------------------------------------ host emulation
#include <stdlib.h>
#include <functional>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
static inline void __syncthreads() {}
static inline void __threadfence() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
enum { cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return 0; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "host-emulation error"; }

void qd_emu_launch(dim3 grid, dim3 block, const std::function<void()>& body);
#define QD_LAUNCH(kern, grid, block, stream, ...) \
  qd_emu_launch((grid), (block), [&]() { kern(__VA_ARGS__); })
#define QD_EMU 1
#else
// ------------------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define QD_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define QD_EMU 0
#endif

#define QD_HD __host__ __device__ __forceinline__
#define QD_D __device__ __forceinline__

// Cooperative loops over a block (see header comment).
#if QD_EMU
#define QD_BLOCK_FIRST_FOR(k, n) if (threadIdx.x == 0) for (int k = 0; k < (int)(n); ++k)
#define QD_BLOCK_LAST_FOR(k, n) if (threadIdx.x == blockDim.x - 1) for (int k = 0; k < (int)(n); ++k)
#else
#define QD_BLOCK_FIRST_FOR(k, n) for (int k = threadIdx.x; k < (int)(n); k += blockDim.x)
#define QD_BLOCK_LAST_FOR(k, n) for (int k = threadIdx.x; k < (int)(n); k += blockDim.x)
#endif

// np.nan_to_num for one value (NaN -> 0, +-inf -> +-DBL_MAX).  Device form: selects only, no branches (the
// branchy form cost 6-10 issue slots and a reconvergence point per value in every stencil load).
QD_HD double qd_nan_to_num(double x) {
#if !QD_EMU && defined(__CUDA_ARCH__)
  const int hi = __double2hiint(x);
  const double big = __hiloint2double((hi & 0x80000000) | 0x7fefffff, 0xffffffff);     // copysign(DBL_MAX, x)
  const double alt = (x != x) ? 0.0 : big;
  return (fabs(x) <= DBL_MAX) ? x : alt;
#else
  if (x != x) return 0.0;
  if (x > DBL_MAX) return DBL_MAX;
  if (x < -DBL_MAX) return -DBL_MAX;
  return x;
#endif
}
// np.clip = minimum(maximum(x, lo), hi) for non-NaN bounds; a NaN x fails both comparisons and comes out unchanged, so
// no separate NaN test is needed (it was 7-14 % of the issue slots of the column / energy / tail kernels)
QD_HD double qd_clip(double x, double lo, double hi) {
  const double y = x < lo ? lo : x;
  return y > hi ? hi : y;
}
// np.maximum / np.minimum: NaN propagates from either argument.  (a > b || a != a) ? a : b -- a NaN b fails the
// comparison and is returned by the else branch.
QD_HD double qd_max(double a, double b) { return (a > b || a != a) ? a : b; }
QD_HD double qd_min(double a, double b) { return (a < b || a != a) ? a : b; }

// ------------------------------------------------------------------------------ block reductions
// Each returns true in exactly ONE thread of the block, with *total = reduction over the block.
// SLOT separates independent reductions inside one kernel.  Must be called by every thread.
#define QD_MAX_WARPS 32
template <int SLOT>
QD_D bool qd_block_sum(double v, double* total) {
#if QD_EMU
  static thread_local double acc = 0.0;
  static thread_local unsigned cnt = 0;
  acc += v;
  if (++cnt == blockDim.x * blockDim.y) { *total = acc; acc = 0.0; cnt = 0; return true; }
  return false;
#else
  __shared__ double sm[QD_MAX_WARPS];
  const unsigned tid = threadIdx.y * blockDim.x + threadIdx.x;
  const unsigned nthr = blockDim.x * blockDim.y;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) sm[tid >> 5] = v;
  __syncthreads();
  bool owner = false;
  if (tid < 32) {
    double x = (tid < (nthr + 31) / 32) ? sm[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (tid == 0) { *total = x; owner = true; }
  }
  __syncthreads();
  return owner;
#endif
}

template <int SLOT>
QD_D bool qd_block_max(double v, double* total) {   // NaN-ignoring max of non-negative values
#if QD_EMU
  static thread_local double acc = 0.0;
  static thread_local unsigned cnt = 0;
  if (v > acc) acc = v;
  if (++cnt == blockDim.x * blockDim.y) { *total = acc; acc = 0.0; cnt = 0; return true; }
  return false;
#else
  __shared__ double sm[QD_MAX_WARPS];
  const unsigned tid = threadIdx.y * blockDim.x + threadIdx.x;
  const unsigned nthr = blockDim.x * blockDim.y;
  for (int o = 16; o > 0; o >>= 1) { double w = __shfl_down_sync(0xffffffffu, v, o); if (w > v) v = w; }
  if ((tid & 31) == 0) sm[tid >> 5] = v;
  __syncthreads();
  bool owner = false;
  if (tid < 32) {
    double x = (tid < (nthr + 31) / 32) ? sm[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { double w = __shfl_down_sync(0xffffffffu, x, o); if (w > x) x = w; }
    if (tid == 0) { *total = x; owner = true; }
  }
  __syncthreads();
  return owner;
#endif
}

// "Last block done" election.  Called by EVERY thread of the block after the block's global
// writes / atomics.  GPU: returns true in all threads of the last of `nblocks` blocks to arrive
// (so they can run a cooperative epilogue); host build: true only in the last thread of that
// block, which then runs the QD_BLOCK_LAST_* loops serially.  Resets the ticket for the next use.
QD_D bool qd_block_is_last(unsigned* ticket, unsigned nblocks, bool system_scope = false) {
#if QD_EMU
  static thread_local unsigned cnt = 0;
  if (++cnt == blockDim.x * blockDim.y) {
    cnt = 0;
    unsigned t = (*ticket)++;
    if (t == nblocks - 1) { *ticket = 0; return true; }
  }
  return false;
#else
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    if (system_scope) __threadfence_system();     // the latitude-band kernels publish stores to peer GPUs this way
    else __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    last = (t == nblocks - 1);
    if (last) *ticket = 0;
    __threadfence();
  }
  __syncthreads();
  return last != 0;
#endif
}

// Loads that must observe other blocks' writes (bypass L1).
#if QD_EMU
#define QD_LDCG(p) (*(p))
#define QD_BLOCK_LAST_ONE if (threadIdx.x == blockDim.x - 1)
#else
#define QD_LDCG(p) __ldcg(p)
#define QD_BLOCK_LAST_ONE if (threadIdx.x == 0)
#endif

// Deterministic sum of `n` per-block partials inside the last block's epilogue; true in one thread.
template <int SLOT>
QD_D bool qd_final_sum(const double* part, unsigned n, double* total) {
#if QD_EMU
  double s = 0.0;
  for (unsigned k = 0; k < n; ++k) s += part[k];
  *total = s;
  return true;
#else
  double s = 0.0;
  for (unsigned k = threadIdx.x; k < n; k += blockDim.x) s += __ldcg(part + k);
  return qd_block_sum<SLOT>(s, total);
#endif
}
template <int SLOT>
QD_D bool qd_final_max(const double* part, unsigned n, double* total) {
#if QD_EMU
  double s = 0.0;
  for (unsigned k = 0; k < n; ++k) if (part[k] > s) s = part[k];
  *total = s;
  return true;
#else
  double s = 0.0;
  for (unsigned k = threadIdx.x; k < n; k += blockDim.x) { double v = __ldcg(part + k); if (v > s) s = v; }
  return qd_block_max<SLOT>(s, total);
#endif
}
