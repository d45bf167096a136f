// qd_api.cu -- context, C ABI and step orchestration of libqd_b200 (see include/qd_b200.h).
//
// The step is a fixed sequence of phase kernels on one CUDA stream; everything data dependent
// (medians, renormalisation sums, the weak-humidity fallback, the ocean's CFL sub-step count)
// stays on the device in the per-member scalar table and is consumed by later kernels.
#include "qd_loop.cuh"
#include "qd_hyper4.cuh"
#include "qd_ocean_fused.cuh"
#include "qd_eco.cuh"
#include "qd_gauss2d.cuh"
#include "qd_phyto.cuh"
#include "qd_indiv.cuh"
#include "qd_diag.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <map>
#include <string>

#ifdef QD_HOST_EMU
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
void qd_emu_launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  gridDim = grid; blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = dim3(bx, by, bz);
        for (unsigned ty = 0; ty < block.y; ++ty)
          for (unsigned tx = 0; tx < block.x; ++tx) { threadIdx = dim3(tx, ty, 0); body(); }
      }
}
#endif

#define QD_NUSER_ROWS 6
#define QD_NPART 4
#define QD_NVB_MAX 1184            // virtual blocks of the sum reductions (8 x 148: one resident wave of 256-thread blocks for one member)
#define QD_FORCING_CAP 64          // steps per forcing-table upload; longer qd_loop_step calls are chunked

struct qd_route {
  int ready = 0, n_levels = 0, n_lakes = 0;
  long long n_order = 0;
  std::vector<int> level_off;                       // host: [n_levels+1]
  int *d_level_cells = nullptr;                     // cells sorted by level
  int *d_don_off = nullptr, *d_don = nullptr;       // early donors CSR per cell (sorted by order position)
  int *d_late_off = nullptr, *d_late = nullptr;     // late donors CSR
  int *d_ocean_list = nullptr; long long n_ocean = 0;     // ocean-draining cells in flow_order order
  int *d_lake_list = nullptr, *d_lake_of = nullptr; long long n_lake_store = 0;
  unsigned char* d_in_order = nullptr;
  unsigned char* d_land = nullptr;                  // the NETWORK's land mask (routing.py:112,232)
  double *d_buffer = nullptr, *d_mass = nullptr, *d_after = nullptr, *d_out = nullptr;
};

#define QD_SEL_SITES 4          // call sites of the exact median that remember their last first digit (op_median)
struct qd_ctx {
  int nlat, nlon, ncell, batch, device, nblk, cur_nblk, red_blk, h4_stream, polar_advances_step, spec_attr_set, g2_fused;
  cudaStream_t stream;
  QdGeo geo;
  double *d_rows, *d_cols, *d_prm, *d_scal, *h_prm;
  QdRcp* d_udiv = nullptr; double udiv_dt = 0.0; int udiv_valid = 0;   // derived divisor table (qd_derive)
  double* fields; uint8_t* masks;
  double* d_part[QD_NPART]; unsigned* d_ticket;
  unsigned* d_hist; unsigned long long* d_mingt; int sel_gx;
  unsigned long long* d_sel_list; unsigned* d_sel_cnt; int* d_sel_more; unsigned* d_sel_spec = nullptr;
  qd_forcing_t* d_forcing; int forcing_cap; int* d_step_idx; double* d_hcos;
  double *d_twid, *d_spec_coef, *d_spec_out;
  double* d_stage[5];
  long long launches;
  int atm_counter, oc_counter, has_cloud_eff;
  int last_nsub_max;
  QdGaussW w_sigma1, w_cloud; int w_set;
  int* d_sub_ctr; int use_graphs;
  int tail_did_prep = 0;                            // set by atmos_core (loop mode): k_tail also did the ocean step's preparation
  int graph_failures = 0;                           // step / ocean graph captures that failed (the step then runs in stream mode)
#ifndef QD_HOST_EMU
  cudaStream_t cap_stream, cap_stream2;
  cudaGraph_t capture_graph;                        // non-null while a whole loop step is being captured

  std::map<const void*, int> red_cache;
  std::map<unsigned long long, cudaGraphExec_t> ocean_graphs;
  std::map<unsigned long long, std::pair<cudaGraphExec_t, long long>> step_graphs;   // variant -> (exec, launches per step)
  bool ocean_fm_on = false;                                // set by ocean_core: closing pass fused with the next sub-step's momentum
  bool ocean_fused = false;                                // set by ocean_substep_body: the last body took the fused two-kernel path
  int ocean_fused_enable = 0;                              // qd_set_ocean_fused: opt-in (measured slower than the four-kernel form on B200)
  long long ocean_body_launches(bool do_hyper, bool do_shap, const qd_step_cfg_t* cfg) const {
    if (ocean_fused) return 6 + 1;                           // pole pass (momentum, 2 x del^4 tiles, continuity) + k_ocean_fused + k_ocean_close + advance
    const long long tiles_i = (nlon + 63) / 64;                  // launch_hyper4: large grids take the stream + pole-tile pair
    const bool stream = h4_stream && nlat >= 96 && nlon >= 64 && tiles_i * ((nlat + 31) / 32) * batch * 3 >= 2 * 148;
    return (ocean_fm_on ? 2 : 3) + (do_hyper ? (stream ? 2 : 1) * std::max(1, cfg->oc_k4_nsub) : 0) + (do_shap ? 2 * std::max(1, cfg->oc_shapiro_n) : 0) + 1;
  }
#endif
  char err[512];
  qd_route route;
  void* prof;
  // latitude bands (qd_band.cuh): control block, exchange buffer, per-field valid halo width
  QdBandCtl band; int band_on; size_t band_bytes; char* band_base; void* band_peer_map[QD_BAND_MAXW];
  int band_valid[QD_F_COUNT + QD_M_COUNT]; char band_shm[64]; int band_maxext;
#ifndef QD_HOST_EMU
  CUtensorMap tmaps[QD_F_COUNT]; int tma_ok = 0;           // one 2-D tensor map per field slot ([B * n_lat][n_lon] doubles, box 64 x 24): k_gauss2d_r4
#endif
  double* d_oc_k4 = nullptr;                               // [B][3][nlat] k4 rows of the fused ocean sub-step for the current sub_dt
  double* d_oc_part = nullptr; int oc_npart = 0;           // eta partial sums of the fused ocean sub-step (pole pass + one per warp)
  QdIndivArgs indiv; int indiv_ready;                     // individual pool (qd_indiv.cuh); device arrays owned here
  double *d_diag_part, *d_diag_out;                       // qd_diag scratch
  double *h_diag = nullptr, *h_diag_dev = nullptr;         // pinned, device-mapped result block: k_diag writes straight to the host
  double* d_phyto_tmp; size_t phyto_cap;                  // scratch of qd_phyto_advect_diffuse
  // ecology sub-daily (qd_eco.cuh)
  const double* d_lai; int eco_nl, eco_every_nphys, eco_steps, eco_have_alpha;
  double eco_k, eco_every_hours, eco_delta;
};

#ifndef QD_HOST_EMU
#include <map>
#include <string>
struct qd_prof_rec { const char* name; cudaEvent_t a, b; };
struct qd_prof {
  int on = 0;
  std::vector<qd_prof_rec> pending;
  std::map<std::string, std::pair<long long, double>> acc;   // name -> (count, ms)
};
static qd_prof* qd_prof_of(qd_ctx* c);
static int qd_prof_begin(qd_ctx* c, const char* name) {
  qd_prof* p = qd_prof_of(c);
  if (!p || !p->on) return -1;
  qd_prof_rec r; r.name = name;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, c->stream);
  p->pending.push_back(r);
  return (int)p->pending.size() - 1;
}
static void qd_prof_end(qd_ctx* c, int idx) {
  if (idx < 0) return;
  qd_prof* p = qd_prof_of(c);
  cudaEventRecord(p->pending[idx].b, c->stream);
}
static void qd_prof_harvest(qd_ctx* c) {
  qd_prof* p = qd_prof_of(c);
  if (!p) return;
  cudaStreamSynchronize(c->stream);
  for (auto& r : p->pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    auto& e = p->acc[r.name];
    e.first += 1; e.second += ms;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  p->pending.clear();
}
#else
static inline int qd_prof_begin(qd_ctx*, const char*) { return -1; }
static inline void qd_prof_end(qd_ctx*, int) {}
#endif

static int qd_fail(qd_ctx* c, int code, const char* what, cudaError_t e) {
  if (c) snprintf(c->err, sizeof(c->err), "%s: %s", what, e != cudaSuccess ? cudaGetErrorString(e) : "invalid argument");
  return code;
}
#define QD_CUDA(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return qd_fail((c), QD_E_CUDA, #call, e_); } while (0)
#define QD_CHECK_LAUNCH(c) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return qd_fail((c), QD_E_CUDA, "kernel launch", e_); } while (0)
#define QD_REQUIRE(c, cond) do { if (!(cond)) return qd_fail((c), QD_E_INVALID, #cond, cudaSuccess); } while (0)
#define QD_BOUND(c) do { if (!(c)->fields || !(c)->masks) return qd_fail((c), QD_E_UNBOUND, "qd_bind was not called", cudaSuccess); } while (0)

// Blocks per member of a grid-stride reduction kernel: exactly ONE resident wave of that kernel (a partial second
// wave ran k_ocean_continuity at 55 % occupancy and doubled its time, profiles/r01_ncu_full_hires_step_v2.csv).
// one resident wave of `kern` (256-thread blocks) on this device
static int qd_wave_blocks(qd_ctx* c, const void* kern) {
#ifdef QD_HOST_EMU
  (void)kern;
  return c->red_blk * c->batch;
#else
  auto it = c->red_cache.find(kern);
  if (it != c->red_cache.end()) return it->second;
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QD_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  const int n = std::max(1, per_sm * std::max(1, sms));
  c->red_cache[kern] = n;
  return n;
#endif
}
static int qd_red_blocks(qd_ctx* c, const void* kern) { return std::max(1, qd_wave_blocks(c, kern) / c->batch); }

// Every launch goes through QD_KG so that launches are counted and, in profiling mode, bracketed by
// CUDA events on the launching stream (per-kernel device time for bench.py's roofline object).
#define QD_KGN(c, name, kern, grid, block, ...) do { \
    const int pi_ = qd_prof_begin((c), (name)); \
    QD_LAUNCH(kern, (grid), (block), (c)->stream, __VA_ARGS__); \
    qd_prof_end((c), pi_); (c)->launches++; } while (0)
#define QD_KG(c, kern, grid, block, ...) do { \
    const int pi_ = qd_prof_begin((c), #kern); \
    QD_LAUNCH(kern, (grid), (block), (c)->stream, __VA_ARGS__); \
    qd_prof_end((c), pi_); (c)->launches++; } while (0)
#define QD_K(c, kern, ...) QD_KG(c, kern, dim3((c)->cur_nblk, (c)->batch), dim3(QD_THREADS), __VA_ARGS__)
// grid-stride kernels that end in a grid-wide reduction (QD_CELL_LOOP): at most red_blk blocks per member
#ifdef QD_HOST_EMU
#define QD_KR(c, kern, ...) QD_KG(c, kern, dim3((c)->geo.nvb, (c)->batch), dim3(QD_THREADS), __VA_ARGS__)
#else
// Virtual blocks of a sum-reduction kernel: ONE resident wave of that kernel for a single member -- a property of the
// kernel and the device, not of how many members share the GPU (B = 8 gives the bits of B = 1).  With B members the
// physical grid per member is wave / B blocks and every block takes B virtual blocks (rounded so that the blocks differ by
// at most one virtual block).
static inline int qd_nvb_for(qd_ctx* c, const void* kern) { return std::max(1, std::min(c->nblk, qd_wave_blocks(c, kern))); }
static inline int qd_kr_grid(qd_ctx* c, const void* kern) {
  const int cap = std::max(1, std::min(qd_red_blocks(c, kern), c->cur_nblk)), nvb = c->geo.nvb;
  if (cap >= nvb) return nvb;
  const int k = (nvb + cap - 1) / cap;
  return (nvb + k - 1) / k;
}
#define QD_KR(c, kern, ...) do { (c)->geo.nvb = qd_nvb_for((c), (const void*)kern); \
    QD_KG(c, kern, dim3(qd_kr_grid((c), (const void*)kern), (c)->batch), dim3(QD_THREADS), __VA_ARGS__); } while (0)
#endif

struct BIn;
static void band_release(qd_ctx* c);
// Captured graphs bake in kernel arguments (pointers, Gaussian taps, switches read from h_prm on the host).  Every entry
// point that changes one of those drops the cached graphs; the next step captures again.
static void qd_drop_graphs(qd_ctx* c) {
#ifndef QD_HOST_EMU
  if (c->step_graphs.empty() && c->ocean_graphs.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto& kv : c->ocean_graphs) if (kv.second) cudaGraphExecDestroy(kv.second);
  for (auto& kv : c->step_graphs) if (kv.second.first) cudaGraphExecDestroy(kv.second.first);
  c->ocean_graphs.clear(); c->step_graphs.clear();
#else
  (void)c;
#endif
}
#ifndef QD_HOST_EMU
static qd_prof* qd_prof_of(qd_ctx* c) { return (qd_prof*)c->prof; }
static bool qd_prof_on(qd_ctx* c) { return c->prof && qd_prof_of(c)->on; }
#endif
extern "C" int qd_profile(qd_ctx* c, int enable) {
  if (!c) return QD_E_INVALID;
#ifndef QD_HOST_EMU
  if (!c->prof) c->prof = new qd_prof();
  qd_prof* p = qd_prof_of(c);
  qd_prof_harvest(c);
  if (enable && !p->on) p->acc.clear();
  p->on = enable ? 1 : 0;
#else
  (void)enable;
#endif
  return QD_OK;
}
// "name count total_ms\n" per kernel, largest total first; returns the number of bytes written
extern "C" int qd_profile_report(qd_ctx* c, char* buf, int buflen) {
  if (!c || !buf || buflen < 1) return QD_E_INVALID;
  buf[0] = 0;
#ifndef QD_HOST_EMU
  qd_prof* p = qd_prof_of(c);
  if (!p) return 0;
  qd_prof_harvest(c);
  std::vector<std::pair<double, std::string>> order;
  for (auto& kv : p->acc) order.push_back({-kv.second.second, kv.first});
  std::sort(order.begin(), order.end());
  int n = 0;
  for (auto& o : order) {
    auto& e = p->acc[o.second];
    int w = snprintf(buf + n, buflen - n, "%s %lld %.6f\n", o.second.c_str(), e.first, e.second);
    if (w < 0 || w >= buflen - n) break;
    n += w;
  }
  return n;
#else
  return 0;
#endif
}

static inline double* F(qd_ctx* c, int id) { return c->fields + (size_t)id * c->batch * c->ncell; }
static inline uint8_t* M(qd_ctx* c, int id) { return c->masks + (size_t)id * c->batch * c->ncell; }
static inline const double* ROW(qd_ctx* c, int id) { return c->d_rows + (size_t)id * c->nlat; }   // member 0's copy
static inline bool qd_in_row_table(qd_ctx* c, const double* p) {
  return p >= c->d_rows && p < c->d_rows + (size_t)c->batch * c->geo.row_bstride;
}

// Companion rows of a cosine table t[0..nlat): t[1] = 1/c, t[2] = 1/c^2 and the centred-stencil coefficients of
// the del^4 kernels t[3] = ap, t[4] = am, t[5] = bl (same expressions as k_hyper4_tile evaluates per block).
static void qd_cos_companions(const QdGeo& g, int nlat, double* t) {
  for (int j = 0; j < nlat; ++j) { const double x = t[j]; t[nlat + j] = 1.0 / x; t[2 * (size_t)nlat + j] = 1.0 / (x * x); }
  for (int j = 0; j < nlat; ++j) {
    double ap = 0.0, am = 0.0, bl = 0.0;
    if (j >= 2 && j <= nlat - 3) {
      const double icj = t[nlat + j] * g.inv_2dlat;
      ap = (icj * (t[j + 1] * g.inv_2dlat)) * g.inv_a_sq;
      am = (icj * (t[j - 1] * g.inv_2dlat)) * g.inv_a_sq;
      bl = (t[2 * (size_t)nlat + j] * g.inv_dlon_sq) * g.inv_a_sq;
    }
    t[3 * (size_t)nlat + j] = ap; t[4 * (size_t)nlat + j] = am; t[5 * (size_t)nlat + j] = bl;
  }
}
// one member's built-in row table: host rows -> companions of the two Laplacian cosine tables -> device
static int qd_upload_rows(qd_ctx* c, int member, const double* rows) {
  const int nlat = c->nlat;
  std::vector<double> h(rows, rows + (size_t)QD_R_COUNT * nlat);
  qd_cos_companions(c->geo, nlat, h.data() + (size_t)QD_R_COS_ADV_HALF * nlat);
  qd_cos_companions(c->geo, nlat, h.data() + (size_t)QD_R_COS_LAP_ATM * nlat);
  if (cudaMemcpy(c->d_rows + (size_t)member * c->geo.row_bstride, h.data(), h.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return QD_E_CUDA;
  return QD_OK;
}

// ------------------------------------------------------------------------------ lifecycle
extern "C" int qd_version(void) { return 100; }
extern "C" const char* qd_last_error(const qd_ctx* c) { return c ? c->err : "null context"; }

extern "C" int qd_destroy(qd_ctx* c);
extern "C" int qd_create(int nlat, int nlon, int batch, int device, double a, double dlat, double dlon,
                         double a_sq, double dlon_sq, const double* rows_host, const double* cols_host,
                         const double* params_host, qd_ctx** out) {
  if (!out || nlat < 3 || nlon < 3 || batch < 1 || !rows_host || !cols_host || !params_host) return QD_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return QD_E_NODEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return QD_E_NODEVICE;
  qd_ctx* c = new qd_ctx();
  memset(c->err, 0, sizeof(c->err));
  c->nlat = nlat; c->nlon = nlon; c->ncell = nlat * nlon; c->batch = batch; c->device = device;
  c->nblk = (c->ncell + QD_THREADS - 1) / QD_THREADS;
  c->cur_nblk = c->nblk;
  c->h4_stream = 1; c->polar_advances_step = 0; c->spec_attr_set = 0; c->g2_fused = 1;
  c->red_blk = std::max(1, c->nblk / 3);          // host check build: exercise the grid-stride loops
  c->stream = 0; c->fields = nullptr; c->masks = nullptr; c->launches = 0;
  c->atm_counter = 0; c->oc_counter = 0; c->has_cloud_eff = 0; c->last_nsub_max = 1;
  memset(&c->band, 0, sizeof(c->band)); c->band.world = 1; c->band_on = 0; c->band_bytes = 0; c->band_base = nullptr;
  memset(c->band_peer_map, 0, sizeof(c->band_peer_map)); memset(c->band_shm, 0, sizeof(c->band_shm)); c->band_maxext = 0;
  for (int k = 0; k < QD_F_COUNT + QD_M_COUNT; ++k) c->band_valid[k] = 1 << 28;
  c->d_phyto_tmp = nullptr; c->phyto_cap = 0;
  memset(&c->indiv, 0, sizeof(c->indiv)); c->indiv_ready = 0;
  c->d_lai = nullptr; c->eco_nl = 0; c->eco_every_nphys = 1; c->eco_steps = 0; c->eco_have_alpha = 0;
  c->eco_k = 0.5; c->eco_every_hours = 6.0; c->eco_delta = 0.05;
  c->forcing_cap = 0; c->d_forcing = nullptr; c->w_set = 0; c->prof = nullptr; c->use_graphs = 2;
#ifndef QD_HOST_EMU
  c->cap_stream = nullptr; c->cap_stream2 = nullptr; c->capture_graph = nullptr;
#endif
  const size_t nrows = (size_t)(QD_R_COUNT + 7 * QD_NUSER_ROWS) * nlat;   // one table PER MEMBER (K4, sponge, polar rows follow the member's parameters)
  // a failed allocation releases everything allocated so far (qd_destroy frees null pointers harmlessly)
#define QD_ALLOC(ptr, bytes) do { if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) { cudaGetLastError(); qd_destroy(c); return QD_E_CUDA; } cudaMemset((ptr), 0, (bytes)); } while (0)
  QD_ALLOC(c->d_rows, (size_t)batch * nrows * 8);
  QD_ALLOC(c->d_cols, (size_t)QD_C_COUNT * nlon * 8);
  QD_ALLOC(c->d_prm, (size_t)batch * QD_P_COUNT * 8);
  QD_ALLOC(c->d_scal, (size_t)batch * QD_S_COUNT * 8);
  QD_ALLOC(c->d_udiv, (size_t)batch * QD_U_COUNT * sizeof(QdRcp));
  for (int k = 0; k < QD_NPART; ++k) QD_ALLOC(c->d_part[k], (size_t)batch * c->nblk * 8);
  QD_ALLOC(c->d_ticket, (size_t)batch * 8 * sizeof(unsigned));
  QD_ALLOC(c->d_hist, (size_t)(QD_SEL_PASSES + 1) * batch * QD_SEL_MAXBINS * sizeof(unsigned));      // + the speculative second-digit slot (qd_select.cuh)
  QD_ALLOC(c->d_sel_spec, ((size_t)QD_SEL_SITES * batch + 2 * QD_SEL_SITES) * sizeof(unsigned));    // remembered digits, then (calls, hits) per site
  cudaMemset(c->d_sel_spec, 0xff, (size_t)QD_SEL_SITES * batch * sizeof(unsigned));                 // no remembered digit yet
  cudaMemset(c->d_sel_spec + (size_t)QD_SEL_SITES * batch, 0, 2 * QD_SEL_SITES * sizeof(unsigned));
  QD_ALLOC(c->d_mingt, (size_t)batch * sizeof(unsigned long long));
  cudaMemset(c->d_mingt, 0xff, (size_t)batch * sizeof(unsigned long long));
  QD_ALLOC(c->d_sel_list, (size_t)batch * QD_SEL_CAP * sizeof(unsigned long long));
  QD_ALLOC(c->d_sel_cnt, (size_t)batch * sizeof(unsigned));
  QD_ALLOC(c->d_sel_more, sizeof(int));
  c->sel_gx = 1;
#ifndef QD_HOST_EMU
  {
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_select_coop, QD_SEL_THREADS, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int resident = std::max(1, per_sm * sms);
    c->red_blk = std::max(1, std::min(c->nblk, (8 * sms + batch - 1) / batch));     // one resident wave of 256-thread blocks
    const int want = (c->ncell + QD_SEL_THREADS - 1) / QD_SEL_THREADS;
    c->sel_gx = std::max(1, std::min(want, resident / batch));
    if ((long long)c->sel_gx * batch > resident) { qd_destroy(c); return QD_E_INVALID; }   // ensemble too large for one cooperative grid
  }
#endif
  QD_ALLOC(c->d_diag_part, (size_t)batch * QD_DIAG_COUNT * c->nblk * 8);
  QD_ALLOC(c->d_diag_out, (size_t)batch * QD_DIAG_COUNT * 8);
  if (cudaHostAlloc((void**)&c->h_diag, (size_t)batch * QD_DIAG_COUNT * 8, cudaHostAllocMapped) == cudaSuccess &&
      cudaHostGetDevicePointer((void**)&c->h_diag_dev, c->h_diag, 0) != cudaSuccess) { cudaFreeHost(c->h_diag); c->h_diag = nullptr; }
  if (!c->h_diag) { c->h_diag_dev = nullptr; cudaGetLastError(); }        // no mapped memory: fall back to a device buffer + copy
  QD_ALLOC(c->d_step_idx, sizeof(int));
  QD_ALLOC(c->d_sub_ctr, sizeof(int));
#ifndef QD_HOST_EMU
  // eta partial sums of the fused ocean sub-step: nvb slots of the pole pass + one per strip warp (32-row chunks at most)
  c->oc_npart = std::max(1, std::min(c->nblk, QD_NVB_MAX)) + (((nlon + QD_OF_COLS - 1) / QD_OF_COLS) * ((nlat + 15) / 16) + 2 * QD_OF_WARPS);
  QD_ALLOC(c->d_oc_part, (size_t)batch * c->oc_npart * 8);
  QD_ALLOC(c->d_oc_k4, (size_t)batch * 3 * nlat * 8);
  cudaFuncSetAttribute(k_ocean_fused, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#endif
  c->forcing_cap = QD_FORCING_CAP;                  // fixed: captured step graphs hold this pointer
  QD_ALLOC(c->d_forcing, (size_t)QD_FORCING_CAP * sizeof(qd_forcing_t));
  QD_ALLOC(c->d_hcos, (size_t)2 * nlon * 8);
  QD_ALLOC(c->d_twid, (size_t)2 * nlon * 8);
  QD_ALLOC(c->d_spec_coef, (size_t)batch * nlat * 2 * (nlon / 2 + 1) * 8);
  QD_ALLOC(c->d_spec_out, (size_t)batch * c->ncell * 8);
  for (int k = 0; k < 5; ++k) QD_ALLOC(c->d_stage[k], (size_t)c->ncell * 8);
#undef QD_ALLOC
  c->h_prm = (double*)malloc((size_t)batch * QD_P_COUNT * 8);
  memcpy(c->h_prm, params_host, (size_t)batch * QD_P_COUNT * 8);
  cudaMemcpy(c->d_cols, cols_host, (size_t)QD_C_COUNT * nlon * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(c->d_prm, params_host, (size_t)batch * QD_P_COUNT * 8, cudaMemcpyHostToDevice);
  {
    std::vector<double> tw(2 * (size_t)nlon);
    for (int m = 0; m < nlon; ++m) {
      const double ang = 2.0 * 3.14159265358979323846 * (double)m / (double)nlon;
      tw[m] = cos(ang); tw[nlon + m] = sin(ang);
    }
    cudaMemcpy(c->d_twid, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
  }
  QdGeo& g = c->geo;
  g.nlat = nlat; g.nlon = nlon; g.ncell = c->ncell; g.batch = batch;
  g.a = a; g.dlat = dlat; g.dlon = dlon; g.a_sq = a_sq; g.dlon_sq = dlon_sq;
  g.inv_dlat = 1.0 / dlat; g.inv_2dlat = 1.0 / (2.0 * dlat); g.inv_dlon_sq = 1.0 / dlon_sq; g.inv_a_sq = 1.0 / a_sq;
  g.inv_2dlon = 1.0 / (2.0 * dlon); g.inv_a = 1.0 / a; g.inv_dlon = 1.0 / dlon;
  g.own0 = 0; g.own1 = nlat; g.sa0 = 0; g.sa1 = nlat; g.sb0 = 0; g.sb1 = 0; g.ncomp = c->ncell;
#ifdef QD_HOST_EMU
  g.nvb = c->red_blk;      // sequential host threads: one virtual block per physical block (QD_KR), still fewer blocks than cells / 256
#else
  g.nvb = std::max(1, std::min(c->nblk, QD_NVB_MAX));
#endif
  g.div_nlon = (1ull << 40) / (unsigned long long)nlon + 1ull;
  if ((unsigned long long)c->ncell * (unsigned long long)nlon >= (1ull << 40)) { qd_destroy(c); return QD_E_INVALID; }
  g.rows = c->d_rows; g.row_bstride = (long long)nrows; g.cols = c->d_cols; g.prm = c->d_prm; g.scal = c->d_scal; g.udiv = c->d_udiv;
  for (int b = 0; b < batch; ++b) if (qd_upload_rows(c, b, rows_host) != QD_OK) { qd_destroy(c); return QD_E_CUDA; }
  if (cudaGetLastError() != cudaSuccess) { qd_destroy(c); return QD_E_CUDA; }
  *out = c;
  return QD_OK;
}

static void qd_route_free(qd_route& r) {
  cudaFree(r.d_level_cells); cudaFree(r.d_don_off); cudaFree(r.d_don); cudaFree(r.d_late_off); cudaFree(r.d_late);
  cudaFree(r.d_ocean_list); cudaFree(r.d_lake_list); cudaFree(r.d_lake_of); cudaFree(r.d_in_order); cudaFree(r.d_land);
  cudaFree(r.d_buffer); cudaFree(r.d_mass); cudaFree(r.d_after); cudaFree(r.d_out);
  r = qd_route();
}

extern "C" int qd_destroy(qd_ctx* c) {
  if (!c) return QD_E_INVALID;
  cudaStreamSynchronize(c->stream);
  cudaFree(c->d_rows); cudaFree(c->d_cols); cudaFree(c->d_prm); cudaFree(c->d_scal); cudaFree(c->d_udiv);
  for (int k = 0; k < QD_NPART; ++k) cudaFree(c->d_part[k]);
  cudaFree(c->d_ticket); cudaFree(c->d_hist); cudaFree(c->d_mingt); cudaFree(c->d_sel_list); cudaFree(c->d_sel_cnt); cudaFree(c->d_sel_more); cudaFree(c->d_sel_spec); cudaFree(c->d_step_idx); cudaFree(c->d_sub_ctr); cudaFree(c->d_hcos);
  cudaFree(c->d_twid); cudaFree(c->d_spec_coef); cudaFree(c->d_spec_out); cudaFree(c->d_forcing);
  for (int k = 0; k < 5; ++k) cudaFree(c->d_stage[k]);
  qd_route_free(c->route);
  band_release(c);
  cudaFree(c->d_oc_part); cudaFree(c->d_oc_k4);
  cudaFree(c->d_phyto_tmp); cudaFree(c->d_diag_part); cudaFree(c->d_diag_out); if (c->h_diag) cudaFreeHost(c->h_diag);
  cudaFree((void*)c->indiv.cell); cudaFree((void*)c->indiv.ab); cudaFree((void*)c->indiv.tol); cudaFree(c->indiv.e_day); cudaFree(c->indiv.stress);
  free(c->h_prm);
#ifndef QD_HOST_EMU
  if (c->prof) { qd_prof_harvest(c); delete qd_prof_of(c); }
  qd_drop_graphs(c);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  if (c->cap_stream2) cudaStreamDestroy(c->cap_stream2);
#endif
  delete c;
  return QD_OK;
}

#ifndef QD_HOST_EMU
// Tensor maps for the TMA box loads of k_gauss2d_r4: field slot f as a 2-D tensor [B * n_lat][n_lon] of doubles, box
// QD_G3_CW x QD_G3_RH.  cuTensorMapEncodeTiled is fetched through the runtime (no link against libcuda); without it, or
// when a row is not a multiple of 16 bytes (odd n_lon), the kernels stage their tiles with ordinary loads.
static void qd_build_tmaps(qd_ctx* c) {
  c->tma_ok = 0;
  if ((c->nlon & 1) || c->nlon < QD_G3_CW || c->nlat < QD_G3_RH) return;
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return; }
  const cuuint64_t dims[2] = {(cuuint64_t)c->nlon, (cuuint64_t)c->nlat * (cuuint64_t)c->batch};
  const cuuint64_t strides[1] = {(cuuint64_t)c->nlon * 8};
  const cuuint32_t box[2] = {QD_G3_CW, QD_G3_RH}, estr[2] = {1, 1};
  for (int f = 0; f < QD_F_COUNT; ++f) {
    void* base = (void*)(c->fields + (size_t)f * c->batch * c->ncell);
    if (((size_t)base & 15) != 0) return;
    if (((encode_fn)fn)(&c->tmaps[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return;
  }
  c->tma_ok = 1;
}
#endif
extern "C" int qd_set_stream(qd_ctx* c, void* s) { if (!c) return QD_E_INVALID; c->stream = (cudaStream_t)s; return QD_OK; }
extern "C" int qd_synchronize(qd_ctx* c) { if (!c) return QD_E_INVALID; QD_CUDA(c, cudaStreamSynchronize(c->stream)); return QD_OK; }
extern "C" int qd_bind(qd_ctx* c, double* fields, uint8_t* masks) {
  if (!c || !fields || !masks) return QD_E_INVALID;
  qd_drop_graphs(c);
  c->fields = fields; c->masks = masks;
#ifndef QD_HOST_EMU
  qd_build_tmaps(c);
#endif
  return QD_OK;
}
extern "C" int qd_set_params(qd_ctx* c, const double* p) {
  if (!c || !p) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  qd_drop_graphs(c);                               // host-side branches on h_prm shape the captured launch sequence
  memcpy(c->h_prm, p, (size_t)c->batch * QD_P_COUNT * 8);
  QD_CUDA(c, cudaMemcpy(c->d_prm, p, (size_t)c->batch * QD_P_COUNT * 8, cudaMemcpyHostToDevice));
  c->udiv_valid = 0;
  return QD_OK;
}
extern "C" int qd_set_rows(qd_ctx* c, const double* rows) {
  if (!c || !rows) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int b = 0; b < c->batch; ++b) { int rc = qd_upload_rows(c, b, rows); if (rc) return qd_fail(c, rc, "qd_set_rows", cudaSuccess); }
  return QD_OK;
}
extern "C" int qd_set_rows_member(qd_ctx* c, int member, const double* rows) {
  if (!c || !rows || member < 0 || member >= c->batch) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  { int rc = qd_upload_rows(c, member, rows); if (rc) return qd_fail(c, rc, "qd_set_rows_member", cudaSuccess); }
  return QD_OK;
}
extern "C" int qd_get_scalars(qd_ctx* c, double* out) {
  if (!c || !out) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  QD_CUDA(c, cudaMemcpy(out, c->d_scal, (size_t)c->batch * QD_S_COUNT * 8, cudaMemcpyDeviceToHost));
  return QD_OK;
}
extern "C" const double* qd_row_dev(qd_ctx* c, int id) { return (c && id >= 0 && id < QD_R_COUNT + 7 * QD_NUSER_ROWS) ? ROW(c, id) : nullptr; }
#define QD_USER_ROW(c, slot) ((c)->d_rows + (size_t)(QD_R_COUNT + 7 * (slot)) * (c)->nlat)
extern "C" const double* qd_user_row(qd_ctx* c, int slot, const double* rows_host) {
  if (!c || slot < 0 || slot >= QD_NUSER_ROWS || !rows_host) return nullptr;
  cudaStreamSynchronize(c->stream);
  std::vector<double> tmp(7 * (size_t)c->nlat);
  for (int j = 0; j < c->nlat; ++j) tmp[j] = rows_host[j];
  qd_cos_companions(c->geo, c->nlat, tmp.data());
  for (int j = 0; j < c->nlat; ++j) tmp[6 * (size_t)c->nlat + j] = 1.0 / (c->geo.a * rows_host[j]);    // RN(1/(a c)): gather departure points
  double* dst = QD_USER_ROW(c, slot);
  for (int b = 0; b < c->batch; ++b)
    if (cudaMemcpy(dst + (size_t)b * c->geo.row_bstride, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return dst;
}
extern "C" int qd_user_row_member(qd_ctx* c, int slot, int member, const double* rows_host) {
  if (!c || slot < 0 || slot >= QD_NUSER_ROWS || !rows_host || member < 0 || member >= c->batch) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<double> tmp(7 * (size_t)c->nlat);
  for (int j = 0; j < c->nlat; ++j) tmp[j] = rows_host[j];
  qd_cos_companions(c->geo, c->nlat, tmp.data());
  for (int j = 0; j < c->nlat; ++j) tmp[6 * (size_t)c->nlat + j] = 1.0 / (c->geo.a * rows_host[j]);    // RN(1/(a c)): gather departure points
  QD_CUDA(c, cudaMemcpy(QD_USER_ROW(c, slot) + (size_t)member * c->geo.row_bstride, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice));
  return QD_OK;
}
// test / tuning switch: 0 forces the shared-memory tile kernel for del^4 at every size (default 1: large grids stream)
// test / tuning switch: 0 forces the two-pass Gaussian kernels at every size (default 1: large grids use the fused tile kernel)
// 0: two one-axis passes; 1 (default): fused tiles, radius 4 on large grids through k_gauss2d_r4 with TMA box loads; 2: the
// generic fused tile kernel everywhere; 3: k_gauss2d_r4 without TMA (per-element staging) -- 2 and 3 exist for A/B runs
extern "C" int qd_set_gauss2d(qd_ctx* c, int mode) { if (!c || mode < 0 || mode > 3) return QD_E_INVALID; qd_drop_graphs(c); c->g2_fused = mode; return QD_OK; }
// 1: CFL sub-steps of the ocean run as k_ocean_fused + k_ocean_close (qd_ocean_fused.cuh) where the grid allows it; default 0
extern "C" int qd_set_ocean_fused(qd_ctx* c, int enable) {
  if (!c) return QD_E_INVALID;
#ifndef QD_HOST_EMU
  qd_drop_graphs(c); c->ocean_fused_enable = enable ? 1 : 0;
#else
  (void)enable;                   // GPU-only kernels: the host check build keeps the four-kernel form
#endif
  return QD_OK;
}
extern "C" int qd_set_h4_stream(qd_ctx* c, int enable) { if (!c) return QD_E_INVALID; qd_drop_graphs(c); c->h4_stream = enable ? 1 : 0; return QD_OK; }
extern "C" int qd_launch_count(qd_ctx* c, long long* out) { if (!c || !out) return QD_E_INVALID; *out = c->launches; return QD_OK; }
extern "C" int qd_set_counters(qd_ctx* c, int a, int o, int ce) { if (!c) return QD_E_INVALID; c->atm_counter = a; c->oc_counter = o; c->has_cloud_eff = ce; return QD_OK; }
extern "C" int qd_get_counters(qd_ctx* c, int* a, int* o, int* ce) {
  if (!c) return QD_E_INVALID;
  if (a) *a = c->atm_counter; if (o) *o = c->oc_counter; if (ce) *ce = c->has_cloud_eff;
  return QD_OK;
}

static int qd_xfer(qd_ctx* c, void* dev_base, size_t elem, int member, void* host, bool up) {
  QD_BOUND(c);
  QD_REQUIRE(c, member >= -1 && member < c->batch && host);
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  const size_t one = (size_t)c->ncell * elem;
  char* d = (char*)dev_base + (member < 0 ? 0 : (size_t)member * one);
  const size_t n = member < 0 ? one * c->batch : one;
  if (up) QD_CUDA(c, cudaMemcpy(d, host, n, cudaMemcpyHostToDevice));
  else QD_CUDA(c, cudaMemcpy(host, d, n, cudaMemcpyDeviceToHost));
  return QD_OK;
}
extern "C" int qd_upload_field(qd_ctx* c, int f, int m, const double* h) { if (!c || f < 0 || f >= QD_F_COUNT) return QD_E_INVALID; return qd_xfer(c, F(c, f), 8, m, (void*)h, true); }
extern "C" int qd_download_field(qd_ctx* c, int f, int m, double* h) { if (!c || f < 0 || f >= QD_F_COUNT) return QD_E_INVALID; return qd_xfer(c, F(c, f), 8, m, h, false); }
extern "C" int qd_upload_mask(qd_ctx* c, int f, int m, const uint8_t* h) { if (!c || f < 0 || f >= QD_M_COUNT) return QD_E_INVALID; return qd_xfer(c, M(c, f), 1, m, (void*)h, true); }
extern "C" int qd_download_mask(qd_ctx* c, int f, int m, uint8_t* h) { if (!c || f < 0 || f >= QD_M_COUNT) return QD_E_INVALID; return qd_xfer(c, M(c, f), 1, m, h, false); }

// ------------------------------------------------------------------------------ latitude bands (host side)
#ifdef QD_HOST_EMU
#include <sys/mman.h>
#include <fcntl.h>
#include <unistd.h>
#endif
#define QD_VALID_ALL (1 << 28)
static const int QD_BAND_STATIC_F[] = {QD_F_FRICTION, QD_F_BASE_ALBEDO, QD_F_ELEVATION, QD_F_CS_MAP, QD_F_OROG_NX, QD_F_OROG_NY};
static inline bool band_is_static(int id) {
  if (id == QD_F_COUNT + QD_M_LAND) return true;
  for (int s : QD_BAND_STATIC_F) if (s == id) return true;
  return false;
}
// Rows of a rank.  The two ranks that hold a pole launch extra kernels per stencil phase (the del^4 tile rows next to the
// pole, the polar ring means) while every other rank waits for them at the next exchange; `trim` rows are therefore
// moved from each pole rank to the ranks in between (QD_BAND_POLE_TRIM overrides; 0 = equal shares).
static void band_rows_of(int nlat, int world, int rank, int* r0, int* r1) {
  // Measured at 1441x2880 on 8 B200s: 32 rows (of 180) -> 0.803 -> 0.776 ms/step.  The extra work of a pole rank does not
  // depend on its share, so the default is 32 rows whenever a share is at least 128 rows.
  int trim = (world > 2 && nlat / world >= 128) ? 32 : 0;
  if (world > 2) if (const char* ov = getenv("QD_BAND_POLE_TRIM")) trim = std::max(0, std::min(atoi(ov), nlat / world / 2));
  const int inner = nlat + 2 * trim;                         // equal shares of a grid padded by the trimmed rows ...
  const int base = inner / world, rem = inner % world;
  auto start = [&](int r) { return r * base + std::min(r, rem) - (r > 0 ? trim : 0); };      // ... shifted back by `trim` behind rank 0
  *r0 = rank == 0 ? 0 : start(rank);
  *r1 = rank == world - 1 ? nlat : start(rank + 1);
}
// compute region = own rows widened by `ext` rows on both sides (mod n_lat) -> c->geo segments, launch size
static void band_set_ext(qd_ctx* c, int ext) {
  QdGeo& g = c->geo;
  if (!c->band_on) { g.sa0 = 0; g.sa1 = c->nlat; g.sb0 = g.sb1 = 0; }
  else {
    const int lo = g.own0 - ext, hi = g.own1 + ext, n = c->nlat;
    if (hi - lo >= n) { g.sa0 = 0; g.sa1 = n; g.sb0 = g.sb1 = 0; }
    else if (lo < 0) { g.sa0 = 0; g.sa1 = hi; g.sb0 = n + lo; g.sb1 = n; }
    else if (hi > n) { g.sa0 = lo; g.sa1 = n; g.sb0 = 0; g.sb1 = hi - n; }
    else { g.sa0 = lo; g.sa1 = hi; g.sb0 = g.sb1 = 0; }
  }
  g.ncomp = ((g.sa1 - g.sa0) + (g.sb1 - g.sb0)) * c->nlon;
  c->cur_nblk = std::max(1, (g.ncomp + QD_THREADS - 1) / QD_THREADS);
}
static int band_fid(qd_ctx* c, const void* p) {             // field / mask slot of a pointer into the bound blocks, -1 otherwise
  const size_t fsz = (size_t)c->batch * c->ncell;
  const double* d = (const double*)p;
  if (c->fields && d >= c->fields && d < c->fields + (size_t)QD_F_COUNT * fsz) return (int)((d - c->fields) / fsz);
  const uint8_t* m = (const uint8_t*)p;
  if (c->masks && m >= c->masks && m < c->masks + (size_t)QD_M_COUNT * fsz) return QD_F_COUNT + (int)((m - c->masks) / fsz);
  return -1;
}
static int band_exchange(qd_ctx* c, const int* ids, int n) {
  if (getenv("QD_BAND_TRACE") && c->band.rank == 0) { fprintf(stderr, "[band] exchange"); for (int k = 0; k < n; ++k) fprintf(stderr, " %d", ids[k]); fprintf(stderr, "\n"); }
  for (int k0 = 0; k0 < n; k0 += QD_BAND_MAXX) {
    QdBandList L; memset(&L, 0, sizeof(L));
    L.n = std::min(QD_BAND_MAXX, n - k0);
    for (int k = 0; k < L.n; ++k) { L.f[k] = F(c, ids[k0 + k]); c->band_valid[ids[k0 + k]] = c->band.H; }
    // every block waits on a peer: the whole grid must be resident (gx * fields * 2 <= 960 blocks of 256 threads; 148 SMs hold 1184)
    const int gx = std::max(1, std::min(QD_BAND_GX, (c->band.H * c->nlon + 2 * QD_THREADS - 1) / (2 * QD_THREADS)));   // ~2 lines per thread
#ifdef QD_HOST_EMU
    QD_KG(c, k_band_exchange, dim3(gx, L.n, 2), dim3(QD_THREADS), c->band, L, c->geo.own0, c->geo.own1, 1);
    QD_KG(c, k_band_exchange, dim3(gx, L.n, 2), dim3(QD_THREADS), c->band, L, c->geo.own0, c->geo.own1, 2);
#else
    QD_KG(c, k_band_exchange, dim3(gx, L.n, 2), dim3(QD_THREADS), c->band, L, c->geo.own0, c->geo.own1, 3);
#endif
  }
  return QD_OK;
}
// rows a semi-Lagrangian gather may reach: winds are clipped to +-200 m/s (dynamics.py:522-527), the second
// bilinear tap adds one row, one more for safety
static int band_radv(qd_ctx* c, double dt) { return (int)floor(200.0 * fabs(dt) / (c->geo.a * c->geo.dlat)) + 2; }

struct BIn { const void* p; int r; };
#define BP(c, ins, outs) do { int rc_ = band_prep((c), ins, outs); if (rc_) return rc_; } while (0)
#define BL(...) {__VA_ARGS__}
#define BPV(c, ins, outs) do { int rc_ = band_prep_v((c), ins, outs); if (rc_) return rc_; } while (0)
// Called before every field kernel of the step.  ins: (pointer, stencil / gather radius in rows); outs: fields the
// kernel writes on its whole compute region.  Chooses the widest compute region the inputs allow, exchanging
// halos first when an input is not even valid for the kernel's own rows; records the outputs' valid width.
static int band_prep_v(qd_ctx* c, const std::vector<BIn>& ins, const std::vector<const void*>& outs);
static int band_prep(qd_ctx* c, std::initializer_list<BIn> ins, std::initializer_list<const void*> outs) {
  if (!c->band_on) return QD_OK;
  return band_prep_v(c, std::vector<BIn>(ins), std::vector<const void*>(outs));
}
static int band_prep_v(qd_ctx* c, const std::vector<BIn>& ins, const std::vector<const void*>& outs) {
  if (!c->band_on) return QD_OK;
  const int H = c->band.H, T = H / 2;
  bool need = false;
  for (const BIn& in : ins) {
    const int id = band_fid(c, in.p);
    if (id < 0) continue;
    if (c->band_valid[id] < in.r) {
      if (id >= QD_F_COUNT) return qd_fail(c, QD_E_STATE, "latitude bands: a mask would need a halo exchange", cudaSuccess);
      need = true;
    }
  }
  if (need) {
    if (getenv("QD_BAND_TRACE") && c->band.rank == 0) {
      fprintf(stderr, "[band] prep needs:");
      for (const BIn& in : ins) { const int id = band_fid(c, in.p); if (id >= 0) fprintf(stderr, " %d(v%d r%d)", id, c->band_valid[id] > 99 ? 99 : c->band_valid[id], in.r); }
      fprintf(stderr, "\n");
    }
    int ids[32]; int n = 0;
    for (const BIn& in : ins) {
      const int id = band_fid(c, in.p);
      if (id < 0 || id >= QD_F_COUNT || band_is_static(id)) continue;
      if (c->band_valid[id] - in.r >= T) continue;
      bool dup = false;
      for (int k = 0; k < n; ++k) dup = dup || ids[k] == id;
      if (!dup && n < 32) ids[n++] = id;
    }
    int rc = band_exchange(c, ids, n); if (rc) return rc;
  }
  int ext = std::min(H, c->band_maxext);
  for (const BIn& in : ins) {
    const int id = band_fid(c, in.p);
    if (id < 0) continue;
    if (in.r > H) return qd_fail(c, QD_E_STATE, "latitude bands: halo width smaller than a stencil radius", cudaSuccess);
    ext = std::min(ext, c->band_valid[id] - in.r);
  }
  if (c->band.rank == 0) if (const char* tr = getenv("QD_BAND_TRACE")) if (atoi(tr) >= 2) {
    fprintf(stderr, "[band] prep ext %d:", ext);
    for (const BIn& in : ins) { const int id = band_fid(c, in.p); if (id >= 0) fprintf(stderr, " %d(v%d r%d)", id, c->band_valid[id] > 99 ? 99 : c->band_valid[id], in.r); }
    fprintf(stderr, "\n");
  }
  if (ext < 0) return qd_fail(c, QD_E_STATE, "latitude bands: negative compute extent", cudaSuccess);
  band_set_ext(c, ext);
  for (const void* o : outs) { const int id = band_fid(c, o); if (id >= 0) c->band_valid[id] = ext; }
  return QD_OK;
}
// Refill the halos of the listed fields NOW (one exchange) when they are not fully valid.  band_prep exchanges lazily,
// right before the first kernel that needs a halo, and only what that kernel reads; fields that a later kernel of the
// same phase will need anyway ride along here, which halves the number of exchanges per step (each one is a
// neighbour-to-neighbour round trip that no amount of bandwidth shortens).
static int band_hint(qd_ctx* c, std::initializer_list<int> fids) {
  if (!c->band_on) return QD_OK;
  int ids[QD_BAND_MAXX]; int n = 0;
  for (int id : fids) if (c->band_valid[id] < c->band.H && n < QD_BAND_MAXX) ids[n++] = id;
  return n ? band_exchange(c, ids, n) : QD_OK;
}
static void band_invalidate_dynamic(qd_ctx* c) {            // start of a step / of a sub-step body: fixed point for graph replay
  if (!c->band_on) return;
  for (int id = 0; id < QD_F_COUNT + QD_M_COUNT; ++id) c->band_valid[id] = band_is_static(id) ? QD_VALID_ALL : 0;
}
static int band_allreduce(qd_ctx* c, std::initializer_list<int> ids, bool is_max) {
  if (!c->band_on) return QD_OK;
  QdBandRed R; memset(&R, 0, sizeof(R));
  for (int id : ids) { R.id[R.n] = id; R.is_max[R.n] = is_max ? 1 : 0; R.n++; }
  QD_KG(c, k_band_allreduce, dim3(1), dim3(QD_BAND_MAXR * QD_BAND_MAXW), c->band, R, c->d_scal);
  return QD_OK;
}

extern "C" int qd_band_init(qd_ctx* c, int rank, int world, int halo_rows) {
  if (!c || world < 1 || world > QD_BAND_MAXW || rank < 0 || rank >= world) return QD_E_INVALID;
  if (world == 1) { c->band_on = 0; return QD_OK; }
  if (c->batch != 1) return qd_fail(c, QD_E_INVALID, "latitude bands drive one member per context", cudaSuccess);
  const int min_rows = c->nlat / world;
  int H = std::min(halo_rows, min_rows / 2);
  if (H < 4) return qd_fail(c, QD_E_INVALID, "latitude bands need >= 8 rows per rank (del^4 halo of 4)", cudaSuccess);
  QdBandCtl& B = c->band;
  memset(&B, 0, sizeof(B));
  B.rank = rank; B.world = world; B.H = H; B.nlon = c->nlon; B.nlat = c->nlat;
  size_t off = QD_BF_WORDS * 8;
  auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  B.off_inbox = take((size_t)2 * 2 * QD_BAND_MAXX * H * c->nlon * sizeof(QdLine));      // flagged lines, 16 bytes per element
  B.off_sflag = take((size_t)2 * QD_BAND_MAXX * QD_BAND_GX * 8);
  B.off_red = take((size_t)2 * 2 * QD_BAND_MAXW * QD_BAND_MAXR * sizeof(QdLine));     // flagged lines: [set][parity][src][slot]
  B.off_hist = take((size_t)2 * QD_BAND_MAXW * QD_SEL_MAXBINS * sizeof(QdLine));       // two histograms (first digit + speculative second) of two bins per line
  B.off_list = take((size_t)2 * QD_BAND_MAXW * (QD_SEL_CAP + 2) * sizeof(QdLine));
#ifdef QD_HOST_EMU
  B.off_emu = take((size_t)QD_BAND_MAXW * ((size_t)c->ncell + 1) * 8);
#else
  B.off_emu = off;
#endif
  c->band_bytes = off;
#ifdef QD_HOST_EMU
  snprintf(c->band_shm, sizeof(c->band_shm), "/qd_band_%d_%d_%p", (int)getpid(), rank, (void*)c);
  int fd = shm_open(c->band_shm, O_CREAT | O_RDWR | O_TRUNC, 0600);
  if (fd < 0 || ftruncate(fd, (off_t)c->band_bytes) != 0) return qd_fail(c, QD_E_CUDA, "shm_open", cudaSuccess);
  c->band_base = (char*)mmap(nullptr, c->band_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (c->band_base == (char*)MAP_FAILED) return qd_fail(c, QD_E_CUDA, "mmap", cudaSuccess);
  memset(c->band_base, 0, c->band_bytes);
#else
  QD_CUDA(c, cudaMalloc((void**)&c->band_base, c->band_bytes));
  QD_CUDA(c, cudaMemset(c->band_base, 0, c->band_bytes));
  QD_CUDA(c, cudaDeviceSynchronize());
#endif
  B.peer[rank] = c->band_base;
  band_rows_of(c->nlat, world, rank, &c->geo.own0, &c->geo.own1);
  c->band_maxext = H;
  return QD_OK;
}
// 64 opaque bytes that let the other ranks map this rank's exchange buffer (cudaIpcMemHandle_t / shm name)
extern "C" int qd_band_export(qd_ctx* c, void* handle64) {
  if (!c || !handle64 || !c->band_base) return QD_E_INVALID;
  memset(handle64, 0, 64);
#ifdef QD_HOST_EMU
  memcpy(handle64, c->band_shm, std::min(sizeof(c->band_shm), (size_t)64));
#else
  static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  QD_CUDA(c, cudaIpcGetMemHandle(&h, c->band_base));
  memcpy(handle64, &h, sizeof(h));
#endif
  return QD_OK;
}
extern "C" int qd_band_connect(qd_ctx* c, const void* handles /* [world][64], rank order */) {
  if (!c || !handles || !c->band_base) return QD_E_INVALID;
  QdBandCtl& B = c->band;
  for (int r = 0; r < B.world; ++r) {
    if (r == B.rank) continue;
    const char* h = (const char*)handles + (size_t)r * 64;
#ifdef QD_HOST_EMU
    char name[65]; memcpy(name, h, 64); name[64] = 0;
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return qd_fail(c, QD_E_CUDA, "shm_open(peer)", cudaSuccess);
    void* m = mmap(nullptr, c->band_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return qd_fail(c, QD_E_CUDA, "mmap(peer)", cudaSuccess);
    c->band_peer_map[r] = m; B.peer[r] = (char*)m;
#else
    cudaIpcMemHandle_t ih; memcpy(&ih, h, sizeof(ih));
    void* m = nullptr;
    QD_CUDA(c, cudaIpcOpenMemHandle(&m, ih, cudaIpcMemLazyEnablePeerAccess));
    c->band_peer_map[r] = m; B.peer[r] = (char*)m;
#endif
  }
  qd_drop_graphs(c);
  c->band_on = 1;
  band_invalidate_dynamic(c);
  band_set_ext(c, 0);
  return QD_OK;
}
extern "C" int qd_band_info(qd_ctx* c, int* own0, int* own1, int* halo, int* error_word) {
  if (!c) return QD_E_INVALID;
  if (own0) *own0 = c->geo.own0;
  if (own1) *own1 = c->geo.own1;
  if (halo) *halo = c->band_on ? c->band.H : 0;
  if (error_word) {
    *error_word = 0;
    if (c->band_on) {
      unsigned long long e = 0;
#ifdef QD_HOST_EMU
      e = ((unsigned long long*)c->band_base)[QD_BF_ERR];
#else
      QD_CUDA(c, cudaStreamSynchronize(c->stream));
      QD_CUDA(c, cudaMemcpy(&e, c->band_base + QD_BF_ERR * 8, 8, cudaMemcpyDeviceToHost));
#endif
      *error_word = (int)e;
    }
  }
  return QD_OK;
}
extern "C" int qd_band_exchange_bench(qd_ctx* c, int nfields, int iters, float* ms_out) {
  if (!c || !ms_out || nfields < 1 || nfields > QD_BAND_MAXX || iters < 1) return QD_E_INVALID;
  if (!c->band_on) return qd_fail(c, QD_E_STATE, "qd_band_exchange_bench needs latitude bands", cudaSuccess);
  *ms_out = 0.f;
#ifndef QD_HOST_EMU
  int ids[QD_BAND_MAXX];
  for (int k = 0; k < nfields; ++k) ids[k] = k;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) { int rc = band_exchange(c, ids, nfields); if (rc) return rc; }
  cudaEventRecord(a, c->stream);
  for (int k = 0; k < iters; ++k) { int rc = band_exchange(c, ids, nfields); if (rc) return rc; }
  cudaEventRecord(b, c->stream);
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(ms_out, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  band_invalidate_dynamic(c);
#endif
  return QD_OK;
}
static void band_release(qd_ctx* c) {
  if (!c->band_base) return;
#ifdef QD_HOST_EMU
  for (int r = 0; r < QD_BAND_MAXW; ++r) if (c->band_peer_map[r]) munmap(c->band_peer_map[r], c->band_bytes);
  munmap(c->band_base, c->band_bytes);
  shm_unlink(c->band_shm);
#else
  for (int r = 0; r < QD_BAND_MAXW; ++r) if (c->band_peer_map[r]) cudaIpcCloseMemHandle(c->band_peer_map[r]);
  cudaFree(c->band_base);
#endif
  c->band_base = nullptr; c->band_on = 0;
}

// ------------------------------------------------------------------------------ operator building blocks
static QdFields mk_fields(int n) { QdFields f; memset(&f, 0, sizeof(f)); f.n = n; for (int k = 0; k < QD_MAX_FIELDS; ++k) f.scale[k] = 1.0; return f; }

static int op_laplacian(qd_ctx* c, int n, const double* const* src, double* const* dst, const double* cosr) {
  QdFields f = mk_fields(n);
  std::vector<BIn> bi; std::vector<const void*> bo;
  for (int k = 0; k < n; ++k) { f.src[k] = src[k]; f.dst[k] = dst[k]; bi.push_back({src[k], 2}); bo.push_back(dst[k]); }
  BPV(c, bi, bo);
  QD_K(c, k_laplacian, c->geo, f, cosr);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
// Fused del^4 (csrc/qd_hyper4.cuh): n fields, src[k] -> dst[k] (out of place), nsub sub-divisions
// ping-ponging between the two buffer sets; returns in *final_in_dst whether the result ended in dst.
#ifndef QD_HOST_EMU
// Rows per warp of the streaming del^4 kernel.  A warp marches down its chunk row by row (~0.3 us per row of dependent
// work): 64-row chunks keep the 8-row warm-up at 12 % but need enough strips x chunks to fill the machine (148 SMs x 24
// resident warps); smaller problems (latitude bands on 4+ GPUs, batches of 181x360 members) take 32- or 16-row chunks --
// more warps, shorter chains: the largest chunk that still fills the machine.  Identical bits (the chunk length only decides who computes a row).
static int h4s_rows(qd_ctx* c, int nstrips, int rows, int nfields) {
  if (const char* ov = getenv("QD_H4S_ROWS")) { const int r = atoi(ov); return r == 16 ? 16 : (r == 32 ? 32 : 64); }      // A/B override
  for (int R = 64; R > 16; R >>= 1)
    if ((long long)nstrips * ((rows + R - 1) / R) * nfields * c->batch >= 148LL * 24) return R;
  return 16;
}
#endif
static int launch_hyper4(qd_ctx* c, QdHyper4Args& H) {
  const int tiles_i = (c->nlon + QD_H4_TI - 1) / QD_H4_TI;
  const long long blocks32 = (long long)tiles_i * ((c->nlat + 31) / 32) * c->batch * H.n;
  // profile names carry the number of fields in the launch (bench.py: 16 B per cell and field)
  static const char* const nm32[] = {"", "k_hyper4_tile<32>[1]", "k_hyper4_tile<32>[2]", "k_hyper4_tile<32>[3]", "k_hyper4_tile<32>[4]", "k_hyper4_tile<32>[5]"};
  static const char* const nm8[] = {"", "k_hyper4_tile<8>[1]", "k_hyper4_tile<8>[2]", "k_hyper4_tile<8>[3]", "k_hyper4_tile<8>[4]", "k_hyper4_tile<8>[5]"};
  static const char* const nms[] = {"", "k_hyper4_stream[1]", "k_hyper4_stream[2]", "k_hyper4_stream[3]", "k_hyper4_stream[4]", "k_hyper4_stream[5]"};
  const int nn = H.n < 1 ? 1 : (H.n > 5 ? 5 : H.n);
  H.tj_lo = 1 << 30; H.tj_skip = 0; H.ja = 0; H.jb = 0; H.row0 = 0; H.row1 = c->nlat;
  {
    std::vector<BIn> bi; std::vector<const void*> bo;
    for (int k = 0; k < H.n; ++k) { bi.push_back({H.src[k], 4}); bo.push_back(H.dst[k]); }
    BPV(c, bi, bo);
  }
  if (c->band_on) {
    // latitude bands: the tile kernel over each segment of the compute region (large segments stream their
    // centred rows); pole handling inside the kernels works on global row indices
    const int seg[2][2] = {{c->geo.sa0, c->geo.sa1}, {c->geo.sb0, c->geo.sb1}};
    for (int q = 0; q < 2; ++q) {
      int s0 = seg[q][0], s1 = seg[q][1];
      if (s1 <= s0) continue;
#ifndef QD_HOST_EMU
      const int ja = std::max(s0, 8), jb = std::min(s1, ((c->nlat + 7) / 8 - 2) * 8);
      if (c->h4_stream && c->nlon >= 64 && jb - ja >= 32 && (long long)(jb - ja) * c->nlon * H.n >= (1 << 20)) {
        H.ja = ja; H.jb = jb;
        const int nstrips = (c->nlon + QD_H4S_COLS - 1) / QD_H4S_COLS;
        const int R = h4s_rows(c, nstrips, jb - ja, H.n);
        const int nwarps = nstrips * ((jb - ja + R - 1) / R);
        if (s0 < ja) { H.row0 = s0; H.row1 = ja; QD_KGN(c, nm8[nn], k_hyper4_tile<8>, dim3(tiles_i * ((ja - s0 + 7) / 8), c->batch, H.n), dim3(QD_H4_NX, QD_H4_NY), c->geo, H); }
        if (jb < s1) { H.row0 = jb; H.row1 = s1; QD_KGN(c, nm8[nn], k_hyper4_tile<8>, dim3(tiles_i * ((s1 - jb + 7) / 8), c->batch, H.n), dim3(QD_H4_NX, QD_H4_NY), c->geo, H); }
        if (R == 16) QD_KGN(c, nms[nn], k_hyper4_stream<16>, dim3((nwarps + QD_H4S_WARPS - 1) / QD_H4S_WARPS, c->batch, H.n), dim3(32 * QD_H4S_WARPS), c->geo, H);
        else if (R == 32) QD_KGN(c, nms[nn], k_hyper4_stream<32>, dim3((nwarps + QD_H4S_WARPS - 1) / QD_H4S_WARPS, c->batch, H.n), dim3(32 * QD_H4S_WARPS), c->geo, H);
        else QD_KGN(c, nms[nn], k_hyper4_stream<64>, dim3((nwarps + QD_H4S_WARPS - 1) / QD_H4S_WARPS, c->batch, H.n), dim3(32 * QD_H4S_WARPS), c->geo, H);
        continue;
      }
#endif
      H.row0 = s0; H.row1 = s1;
      QD_KGN(c, nm8[nn], k_hyper4_tile<8>, dim3(tiles_i * ((s1 - s0 + 7) / 8), c->batch, H.n), dim3(QD_H4_NX, QD_H4_NY), c->geo, H);
    }
    return QD_OK;
  }
#ifndef QD_HOST_EMU
  const int ntj8 = (c->nlat + 7) / 8;
  if (blocks32 >= 2 * 148 && c->nlat >= 96 && c->nlon >= 64 && c->h4_stream) {
    // large grid: warp-streaming kernel on the rows with a centred dependency cone, tile kernel on the
    // tile rows that touch a pole (tile row 0 and the last two)
    H.ja = 8; H.jb = (ntj8 - 2) * 8;
    const int nstrips = (c->nlon + QD_H4S_COLS - 1) / QD_H4S_COLS;
    const int R = h4s_rows(c, nstrips, H.jb - H.ja, H.n);
    const int nwarps = nstrips * ((H.jb - H.ja + R - 1) / R);
    H.tj_lo = 1; H.tj_skip = ntj8 - 3;
    QD_KGN(c, nm8[nn], k_hyper4_tile<8>, dim3(tiles_i * 3, c->batch, H.n), dim3(QD_H4_NX, QD_H4_NY), c->geo, H);
    if (R == 16) QD_KGN(c, nms[nn], k_hyper4_stream<16>, dim3((nwarps + QD_H4S_WARPS - 1) / QD_H4S_WARPS, c->batch, H.n), dim3(32 * QD_H4S_WARPS), c->geo, H);
    else if (R == 32) QD_KGN(c, nms[nn], k_hyper4_stream<32>, dim3((nwarps + QD_H4S_WARPS - 1) / QD_H4S_WARPS, c->batch, H.n), dim3(32 * QD_H4S_WARPS), c->geo, H);
    else QD_KGN(c, nms[nn], k_hyper4_stream<64>, dim3((nwarps + QD_H4S_WARPS - 1) / QD_H4S_WARPS, c->batch, H.n), dim3(32 * QD_H4S_WARPS), c->geo, H);
    return QD_OK;
  }
#endif
  if (blocks32 >= 2 * 148) {
    QD_KGN(c, nm32[nn], k_hyper4_tile<32>, dim3(tiles_i * ((c->nlat + 31) / 32), c->batch, H.n), dim3(QD_H4_NX, QD_H4_NY), c->geo, H);
  } else {
    QD_KGN(c, nm8[nn], k_hyper4_tile<8>, dim3(tiles_i * ((c->nlat + 7) / 8), c->batch, H.n), dim3(QD_H4_NX, QD_H4_NY), c->geo, H);
  }
  return QD_OK;
}
static int op_hyper(qd_ctx* c, int n, double* const* fld, double* const* scratch, const double* const* k4rows,
                    const double* scale, double dt, int nsub, const double* cosr, bool* final_in_scratch) {
  if (final_in_scratch) *final_in_scratch = false;
  if (dt <= 0.0 || n <= 0) return QD_OK;
  const int ns = nsub > 1 ? nsub : 1;
  bool in_scratch = false;
  for (int s = 0; s < ns; ++s) {
    QdHyper4Args H; memset(&H, 0, sizeof(H));
    H.n = n; H.cosr = cosr; H.dt = dt; H.nsub = ns; H.ocean = 0; H.sc.ctr = nullptr;
    for (int k = 0; k < n; ++k) {
      H.src[k] = in_scratch ? scratch[k] : fld[k];
      H.dst[k] = in_scratch ? fld[k] : scratch[k];
      H.k4rows[k] = k4rows[k]; H.scale[k] = scale ? scale[k] : 1.0; H.raw_k4[k] = 1;
      H.k4_bstride[k] = qd_in_row_table(c, k4rows[k]) ? c->geo.row_bstride : 0;
    }
    int rc = launch_hyper4(c, H); if (rc) return rc;
    in_scratch = !in_scratch;
  }
  QD_CHECK_LAUNCH(c);
  if (final_in_scratch) *final_in_scratch = in_scratch;
  return QD_OK;
}
static int op_shapiro(qd_ctx* c, int n, double* const* fld, double* const* scratch, int passes) {
  const int np = passes > 1 ? passes : 1;
  for (int p = 0; p < np; ++p) {
    QdFields a = mk_fields(n), b2 = mk_fields(n);
    std::vector<BIn> i1, i2; std::vector<const void*> o1, o2;
    for (int k = 0; k < n; ++k) {
      a.src[k] = fld[k]; a.dst[k] = scratch[k]; b2.src[k] = scratch[k]; b2.dst[k] = fld[k];
      i1.push_back({fld[k], 0}); o1.push_back(scratch[k]); i2.push_back({scratch[k], 1}); o2.push_back(fld[k]);
    }
    BPV(c, i1, o1);
    QD_K(c, k_shapiro_lon, c->geo, a, p == 0 ? 1 : 0);
    BPV(c, i2, o2);
    QD_K(c, k_shapiro_lat, c->geo, b2);
  }
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
// Gaussian weight tables supplied by the host (NumPy-evaluated, bit-identical to scipy's).
static QdGaussW mk_gauss_w(int radius, int wrap, const double* weights) {
  QdGaussW w; memset(&w, 0, sizeof(w));
  w.r = radius; w.wrap = wrap;
  for (int k = 0; k <= 2 * radius; ++k) w.w[k] = weights[k];
  return w;
}
extern "C" int qd_set_gauss(qd_ctx* c, int which, int radius, int wrap, const double* weights) {
  if (!c || radius < 0 || radius > QD_GAUSS_MAXR || !weights) return QD_E_INVALID;
  qd_drop_graphs(c);                               // the taps travel by value in captured kernel arguments
  if (which == 0) c->w_sigma1 = mk_gauss_w(radius, wrap, weights); else c->w_cloud = mk_gauss_w(radius, wrap, weights);
  c->w_set |= (1 << (which ? 1 : 0));
  return QD_OK;
}
#ifndef QD_HOST_EMU
// Fused two-axis Gaussian (qd_gauss2d.cuh) over every segment of the compute region.
#define QD_G2S_TJ 8
#define QD_G2S_TI 32
// large grids: 16 x 64 tiles; grids on which those would leave most SMs idle (181x360: 72 tiles): 8 x 32 tiles
static inline bool gauss2d_small(qd_ctx* c) {
  const long long tiles = (long long)((c->nlon + QD_G2_TI - 1) / QD_G2_TI) * ((c->nlat + QD_G2_TJ - 1) / QD_G2_TJ) * c->batch;
  return tiles < 2 * 148;
}
template <int MODE, int TJ, int TI>
static int launch_gauss2d_t(qd_ctx* c, QdG2Args A, const QdGaussW& w, const char* name) {
  const int tiles_i = (c->nlon + TI - 1) / TI;
  const int seg[2][2] = {{c->geo.sa0, c->geo.sa1}, {c->geo.sb0, c->geo.sb1}};
  for (int q = 0; q < 2; ++q) {
    if (seg[q][1] <= seg[q][0]) continue;
    A.row0 = seg[q][0]; A.row1 = seg[q][1];
    const int tiles_j = (A.row1 - A.row0 + TJ - 1) / TJ;
    const dim3 grid(tiles_i * tiles_j, c->batch), block(TI, (QD_G2_NX * QD_G2_NY) / TI);
    if (w.r == 4) QD_KGN(c, name, (k_gauss2d_tile<MODE, 4, TJ, TI>), grid, block, c->geo, A, w);          // sigma = 1 (physics.py:44,69,111,159,330)
    else if (w.r == 1) QD_KGN(c, name, (k_gauss2d_tile<MODE, 1, TJ, TI>), grid, block, c->geo, A, w);     // sigma = 0.2 (run_simulation.py:1931)
    else QD_KGN(c, name, (k_gauss2d_tile<MODE, 0, TJ, TI>), grid, block, c->geo, A, w);
  }
  return QD_OK;
}
// sigma = 1 on large grids: k_gauss2d_r4 (sliding windows; TMA box loads when the tensor maps exist)
template <int MODE>
static int launch_gauss2d_r4(qd_ctx* c, QdG2Args A, const QdGaussW& w, const char* name) {
  const int tiles_i = (c->nlon + QD_G3_TI - 1) / QD_G3_TI;
  const int f0 = band_fid(c, A.src[0]), f1 = A.n > 1 ? band_fid(c, A.src[1]) : f0;
  const bool tma = c->g2_fused == 1 && c->tma_ok && f0 >= 0 && f0 < QD_F_COUNT && f1 >= 0 && f1 < QD_F_COUNT;
  const CUtensorMap& t0 = c->tmaps[tma ? f0 : 0];
  const CUtensorMap& t1 = c->tmaps[tma ? f1 : 0];
  const int seg[2][2] = {{c->geo.sa0, c->geo.sa1}, {c->geo.sb0, c->geo.sb1}};
  for (int q = 0; q < 2; ++q) {
    if (seg[q][1] <= seg[q][0]) continue;
    A.row0 = seg[q][0]; A.row1 = seg[q][1];
    const dim3 grid(tiles_i * ((A.row1 - A.row0 + QD_G3_TJ - 1) / QD_G3_TJ), c->batch), block(256);
    if (tma) QD_KGN(c, name, (k_gauss2d_r4<MODE, true>), grid, block, c->geo, A, w, t0, t1);
    else QD_KGN(c, name, (k_gauss2d_r4<MODE, false>), grid, block, c->geo, A, w, t0, t1);
  }
  return QD_OK;
}
template <int MODE>
static int launch_gauss2d(qd_ctx* c, QdG2Args A, const QdGaussW& w, const std::vector<BIn>& ins, const std::vector<const void*>& outs, const char* name) {
  BPV(c, ins, outs);
  if (gauss2d_small(c)) return launch_gauss2d_t<MODE, QD_G2S_TJ, QD_G2S_TI>(c, A, w, name);
  if (w.r == 4 && c->g2_fused != 2 && c->nlon >= QD_G3_CW) return launch_gauss2d_r4<MODE>(c, A, w, name);
  return launch_gauss2d_t<MODE, QD_G2_TJ, QD_G2_TI>(c, A, w, name);
}
// Fused two-axis Gaussian or the two one-axis passes?  Default: fused wherever at least 148 tiles (of either size) exist.
static inline bool gauss2d_ok(qd_ctx* c, const QdGaussW& w) {
  const long long small_tiles = (long long)((c->nlon + QD_G2S_TI - 1) / QD_G2S_TI) * ((c->nlat + QD_G2S_TJ - 1) / QD_G2S_TJ) * c->batch;
#ifdef QD_G2_LARGE_ONLY
  if (gauss2d_small(c)) return false;
#endif
  return c->g2_fused && w.r >= 1 && w.r <= QD_G2_RMAX && small_tiles >= 148;
}
#endif
// in place on fld[k] (two-pass form); with `out` given and the fused kernel available the result is left in out[k]
// instead (fld untouched, no intermediate field) and *in_out is set
static int op_gauss(qd_ctx* c, int n, double* const* fld, double* const* scratch, const QdGaussW& w, bool* in_out = nullptr) {
  if (in_out) *in_out = false;
  if (w.r == 0) return QD_OK;
#ifndef QD_HOST_EMU
  if (in_out && gauss2d_ok(c, w)) {
    for (int k0 = 0; k0 < n; k0 += 2) {
      QdG2Args A; memset(&A, 0, sizeof(A));
      A.n = std::min(2, n - k0);
      std::vector<BIn> bi; std::vector<const void*> bo;
      for (int k = 0; k < A.n; ++k) { A.src[k] = fld[k0 + k]; A.dst[k] = scratch[k0 + k]; bi.push_back({fld[k0 + k], w.r}); bo.push_back(scratch[k0 + k]); }
      int rc = launch_gauss2d<QD_G2_PLAIN>(c, A, w, bi, bo, "k_gauss2d_tile<plain>"); if (rc) return rc;
    }
    QD_CHECK_LAUNCH(c);
    *in_out = true;
    return QD_OK;
  }
#endif
  QdFields a = mk_fields(n), b2 = mk_fields(n);
  std::vector<BIn> i1, i2; std::vector<const void*> o1, o2;
  for (int k = 0; k < n; ++k) {
    a.src[k] = fld[k]; a.dst[k] = scratch[k]; b2.src[k] = scratch[k]; b2.dst[k] = fld[k];
    i1.push_back({fld[k], w.r}); o1.push_back(scratch[k]); i2.push_back({scratch[k], 0}); o2.push_back(fld[k]);
  }
  BPV(c, i1, o1);
  QD_K(c, k_gauss_lat, c->geo, a, w);
  BPV(c, i2, o2);
  QD_K(c, k_gauss_lon, c->geo, b2, w);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
static int op_bandstop(qd_ctx* c, double* fld, double cutoff, double damp) {
  if (damp <= 0.0 || cutoff <= 0.0) return QD_OK;      // reference only cleans NaNs here; fields are finite
  const int bins = c->nlon / 2 + 1;
  if (bins <= 1) return QD_OK;
  const int kN = bins - 1;
  int kcut = (int)(cutoff * kN);
  kcut = std::max(1, std::min(kN, kcut));
  const double fac = std::max(0.0, 1.0 - std::min(1.0, damp));
  BP(c, BL({fld, 0}), BL(fld));
  {
    const dim3 grid((c->geo.sa1 - c->geo.sa0) + (c->geo.sb1 - c->geo.sb0), c->batch);
#ifdef QD_HOST_EMU
    QD_KG(c, k_zonal_bandstop, grid, dim3(QD_THREADS), c->geo, fld, c->d_twid, kcut, 1.0 - fac, c->d_spec_coef, c->d_spec_out);
#else
    const size_t shm = 3 * (size_t)c->nlon * sizeof(double);
    if (shm > 200 * 1024) return qd_fail(c, QD_E_INVALID, "zonal band-stop: n_lon too large for the shared-memory DFT", cudaSuccess);
    if (shm > 48 * 1024 && !c->spec_attr_set) {
      QD_CUDA(c, cudaFuncSetAttribute(k_zonal_bandstop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
      c->spec_attr_set = 1;
    }
    const int pi_ = qd_prof_begin(c, "k_zonal_bandstop");
    k_zonal_bandstop<<<grid, dim3(QD_THREADS), shm, c->stream>>>(c->geo, fld, c->d_twid, kcut, 1.0 - fac, c->d_spec_coef, c->d_spec_out);
    qd_prof_end(c, pi_); c->launches++;
#endif
  }
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
// exact median of positives of x -> dst[b*stride] (and count -> cnt[b*stride])
// site: which of the loop's medians this is (0 positive part of the diagnosed precipitation, 1 precipitation, 2 P_cond,
// 3 the operator API) -- every site remembers the first digit of its last result for the kernel's speculation
static int op_median(qd_ctx* c, const double* x, double empty_value, double* dst, double* cnt, int stride, int site) {
  QdSelOut out; out.value = dst; out.count = cnt; out.stride = stride; out.empty_value = empty_value;
#ifdef QD_HOST_EMU
  qd_select_host(c->geo, x, out, c->band);
  c->launches++;
#else
  // the kernel leaves hist / list counter / mingt reset for the next launch (no memset nodes in the step graph)
  QdGeo geo = c->geo;
  unsigned* spec = (c->d_sel_spec && !getenv("QD_NO_SELSPEC")) ? c->d_sel_spec + (size_t)site * c->batch : nullptr;
  unsigned* spec_stat = c->d_sel_spec ? c->d_sel_spec + (size_t)QD_SEL_SITES * c->batch + 2 * site : nullptr;
  void* args[] = {(void*)&geo, (void*)&x, (void*)&c->d_hist, (void*)&c->d_sel_list, (void*)&c->d_sel_cnt, (void*)&c->d_mingt, (void*)&c->d_sel_more, (void*)&out, (void*)&c->band, (void*)&spec, (void*)&spec_stat};
  const int pi = qd_prof_begin(c, "k_select_coop");
  QD_CUDA(c, cudaLaunchCooperativeKernel((const void*)k_select_coop, dim3(c->sel_gx, c->batch), dim3(QD_SEL_THREADS), args, 0, c->stream));
  qd_prof_end(c, pi);
  c->launches++;
#endif
  return QD_OK;
}

// RN(1/(a c)) row that belongs to a cosine row table handed to the operator API, or null (plain divisions)
static const double* qd_iac_for(qd_ctx* c, const double* cosr) {
  if (cosr == ROW(c, QD_R_COS_ADV_ATM)) return ROW(c, QD_R_IAC_ADV_ATM);
  if (cosr == ROW(c, QD_R_COS_ADV_HALF)) return ROW(c, QD_R_INV_ACOS_HALF);
  for (int slot = 0; slot < QD_NUSER_ROWS; ++slot) if (cosr == QD_USER_ROW(c, slot)) return cosr + 6 * (size_t)c->nlat;
  return nullptr;
}

// ------------------------------------------------------------------------------ operator C ABI
extern "C" int qd_laplacian(qd_ctx* c, const double* in, double* out, const double* cosr) {
  if (!c || !in || !out || !cosr) return QD_E_INVALID;
  const double* s[1] = {in}; double* d[1] = {out};
  return op_laplacian(c, 1, s, d, cosr);
}
extern "C" int qd_hyperdiffuse(qd_ctx* c, double* f, double* scratch, const double* k4rows, double k4_scale, double dt, int nsub, const double* cosr) {
  if (!c || !f || !scratch || !k4rows || !cosr) return QD_E_INVALID;
  double* fl[1] = {f}; double* sc[1] = {scratch}; const double* kr[1] = {k4rows}; double s[1] = {k4_scale};
  bool in_scratch = false;
  int rc = op_hyper(c, 1, fl, sc, kr, s, dt, nsub, cosr, &in_scratch);
  if (rc) return rc;
  if (in_scratch) QD_CUDA(c, cudaMemcpyAsync(f, scratch, (size_t)c->batch * c->ncell * 8, cudaMemcpyDeviceToDevice, c->stream));
  return QD_OK;
}
extern "C" int qd_advect(qd_ctx* c, const double* in, const double* u, const double* v, double* out, double dt, const double* cosr) {
  if (!c || !in || !u || !v || !out || !cosr) return QD_E_INVALID;
  QdFields f = mk_fields(1); f.src[0] = in; f.dst[0] = out;
  QD_K(c, k_advect, c->geo, f, u, v, dt, cosr, qd_iac_for(c, cosr));
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
extern "C" int qd_shapiro(qd_ctx* c, double* f, double* scratch, int n) {
  if (!c || !f || !scratch) return QD_E_INVALID;
  double* fl[1] = {f}; double* sc[1] = {scratch};
  return op_shapiro(c, 1, fl, sc, n);
}
extern "C" int qd_gaussian(qd_ctx* c, double* f, double* scratch, int radius, int wrap, const double* weights) {
  if (!c || !f || !scratch || !weights || radius < 0 || radius > QD_GAUSS_MAXR) return QD_E_INVALID;
  const QdGaussW w = mk_gauss_w(radius, wrap, weights);
  double* fl[1] = {f}; double* sc[1] = {scratch};
  bool moved = false;
  int rc = op_gauss(c, 1, fl, sc, w, &moved);
  if (rc) return rc;
  if (moved) QD_CUDA(c, cudaMemcpyAsync(f, scratch, (size_t)c->batch * c->ncell * 8, cudaMemcpyDeviceToDevice, c->stream));
  return QD_OK;
}
extern "C" int qd_zonal_bandstop(qd_ctx* c, double* f, double cutoff, double damp) {
  if (!c || !f) return QD_E_INVALID;
  return op_bandstop(c, f, cutoff, damp);
}
extern "C" int qd_divergence(qd_ctx* c, const double* u, const double* v, double* out) {
  if (!c || !u || !v || !out) return QD_E_INVALID;
  QD_K(c, k_divvort, c->geo, u, v, out, 0);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
extern "C" int qd_vorticity(qd_ctx* c, const double* u, const double* v, double* out) {
  if (!c || !u || !v || !out) return QD_E_INVALID;
  QD_K(c, k_divvort, c->geo, u, v, out, 1);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
extern "C" int qd_median_pos(qd_ctx* c, const double* in, double empty_value, double* out_host) {
  if (!c || !in || !out_host) return QD_E_INVALID;
  int rc = op_median(c, in, empty_value, c->d_scal + QD_S_TMP0, c->d_scal + QD_S_TMP1, QD_S_COUNT, 3);
  if (rc) return rc;
  std::vector<double> s((size_t)c->batch * QD_S_COUNT);
  rc = qd_get_scalars(c, s.data());
  if (rc) return rc;
  for (int b = 0; b < c->batch; ++b) out_host[b] = s[(size_t)b * QD_S_COUNT + QD_S_TMP0];
  return QD_OK;
}
extern "C" int qd_median_stats(qd_ctx* c, long long* out /* [4][2] */) {
  if (!c || !out) return QD_E_INVALID;
  for (int k = 0; k < 2 * QD_SEL_SITES; ++k) out[k] = 0;
#ifndef QD_HOST_EMU
  unsigned h[2 * QD_SEL_SITES];
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  QD_CUDA(c, cudaMemcpy(h, c->d_sel_spec + (size_t)QD_SEL_SITES * c->batch, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 2 * QD_SEL_SITES; ++k) out[k] = h[k];
#endif
  return QD_OK;
}
extern "C" int qd_wsum(qd_ctx* c, const double* in, double* out_host) {
  if (!c || !in || !out_host) return QD_E_INVALID;
  QD_KR(c, k_wsum, c->geo, in, ROW(c, QD_R_W), c->d_part[0], c->d_ticket + 2 * c->batch, c->d_scal + QD_S_TMP2, QD_S_COUNT);
  QD_CHECK_LAUNCH(c);
  std::vector<double> s((size_t)c->batch * QD_S_COUNT);
  int rc = qd_get_scalars(c, s.data());
  if (rc) return rc;
  for (int b = 0; b < c->batch; ++b) out_host[b] = s[(size_t)b * QD_S_COUNT + QD_S_TMP2];
  return QD_OK;
}
extern "C" int qd_minmax(qd_ctx* c, const double* in, double* out_host) {
  if (!c || !in || !out_host) return QD_E_INVALID;
  // diagnostics only (every ~100-200 steps in the reference): staged through the host
  std::vector<double> h((size_t)c->ncell);
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int b = 0; b < c->batch; ++b) {
    QD_CUDA(c, cudaMemcpy(h.data(), in + (size_t)b * c->ncell, (size_t)c->ncell * 8, cudaMemcpyDeviceToHost));
    double lo = h[0], hi = h[0];
    for (int k = 1; k < c->ncell; ++k) { if (h[k] < lo) lo = h[k]; if (h[k] > hi) hi = h[k]; }
    out_host[2 * b] = lo; out_host[2 * b + 1] = hi;
  }
  return QD_OK;
}

extern "C" int qd_math_check(qd_ctx* c, const double* x_host, long long n, double* out_host, int which) {
  if (!c || !x_host || !out_host || n <= 0) return QD_E_INVALID;
#ifdef QD_HOST_EMU
  for (long long i = 0; i < n; ++i) out_host[i] = out_host[n + i] = which ? tanh(x_host[i]) : exp(x_host[i]);
#else
  double *dx = nullptr, *dout = nullptr;
  QD_CUDA(c, cudaMalloc((void**)&dx, (size_t)n * 8));
  if (cudaMalloc((void**)&dout, (size_t)n * 16) != cudaSuccess) { cudaFree(dx); return qd_fail(c, QD_E_CUDA, "cudaMalloc", cudaGetLastError()); }
  cudaMemcpyAsync(dx, x_host, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
  k_math_check<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(dx, dout, n, which);
  cudaMemcpyAsync(out_host, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream);
  const cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dx); cudaFree(dout);
  if (e != cudaSuccess) return qd_fail(c, QD_E_CUDA, "qd_math_check", e);
#endif
  return QD_OK;
}

// host-buffer forms of the jax_compat seam (single member, staged through private device buffers)
static int stage_up(qd_ctx* c, int k, const double* h) { QD_CUDA(c, cudaMemcpyAsync(c->d_stage[k], h, (size_t)c->ncell * 8, cudaMemcpyHostToDevice, c->stream)); return QD_OK; }
static int stage_down(qd_ctx* c, int k, double* h) {
  QD_CUDA(c, cudaMemcpyAsync(h, c->d_stage[k], (size_t)c->ncell * 8, cudaMemcpyDeviceToHost, c->stream));
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  return QD_OK;
}
struct BatchOne { qd_ctx* c; int saved; BatchOne(qd_ctx* c_) : c(c_), saved(c_->batch) { c->batch = 1; c->geo.batch = 1; } ~BatchOne() { c->batch = saved; c->geo.batch = saved; } };

extern "C" int qd_laplacian_host(qd_ctx* c, const double* in, double* out, const double* cos_rows) {
  if (!c || !in || !out || !cos_rows) return QD_E_INVALID;
  const double* cr = qd_user_row(c, 0, cos_rows);
  if (!cr) return qd_fail(c, QD_E_CUDA, "qd_user_row", cudaSuccess);
  BatchOne one(c);
  int rc = stage_up(c, 0, in); if (rc) return rc;
  rc = qd_laplacian(c, c->d_stage[0], c->d_stage[1], cr); if (rc) return rc;
  return stage_down(c, 1, out);
}
extern "C" int qd_hyperdiffuse_host(qd_ctx* c, const double* in, double* out, const double* k4_map, double k4_scalar,
                                    double dt, int nsub, const double* cos_rows) {
  if (!c || !in || !out || !cos_rows) return QD_E_INVALID;
  // k4 maps in the reference are functions of latitude only (dynamics.py:557-563, ocean.py:343-347):
  // the first column is the row table.  A scalar k4 becomes a constant row.
  std::vector<double> k4r((size_t)c->nlat);
  bool allneg = true;
  for (int j = 0; j < c->nlat; ++j) {
    double v = k4_map ? k4_map[(size_t)j * c->nlon] : k4_scalar;
    if (v != v) v = 0.0;
    k4r[j] = v;
    if (k4_map) { for (int i = 0; i < c->nlon; ++i) if (k4_map[(size_t)j * c->nlon + i] != v) return qd_fail(c, QD_E_INVALID, "k4 map must be constant along longitude", cudaSuccess); }
    if (v > 0.0) allneg = false;
  }
  if (dt <= 0.0 || allneg) { if (out != in) memcpy(out, in, (size_t)c->ncell * 8); return QD_OK; }   // early-outs dynamics.py:190-204
  const double* cr = qd_user_row(c, 0, cos_rows);
  const double* kr = qd_user_row(c, 1, k4r.data());
  if (!cr || !kr) return qd_fail(c, QD_E_CUDA, "qd_user_row", cudaSuccess);
  BatchOne one(c);
  int rc = stage_up(c, 0, in); if (rc) return rc;
  rc = qd_hyperdiffuse(c, c->d_stage[0], c->d_stage[1], kr, 1.0, dt, nsub, cr); if (rc) return rc;
  return stage_down(c, 0, out);
}
extern "C" int qd_advect_host(qd_ctx* c, const double* in, const double* u, const double* v, double* out, double dt, const double* cos_rows) {
  if (!c || !in || !u || !v || !out || !cos_rows) return QD_E_INVALID;
  const double* cr = qd_user_row(c, 0, cos_rows);
  if (!cr) return qd_fail(c, QD_E_CUDA, "qd_user_row", cudaSuccess);
  BatchOne one(c);
  int rc = stage_up(c, 0, in); if (rc) return rc;
  rc = stage_up(c, 1, u); if (rc) return rc;
  rc = stage_up(c, 2, v); if (rc) return rc;
  rc = qd_advect(c, c->d_stage[0], c->d_stage[1], c->d_stage[2], c->d_stage[3], dt, cr); if (rc) return rc;
  return stage_down(c, 3, out);
}

// ------------------------------------------------------------------------------ atmosphere step

// ------------------------------------------------------------------------------ ecology sub-daily
static QdEcoArgs eco_args(qd_ctx* c, double dt) {
  QdEcoArgs E; memset(&E, 0, sizeof(E));
  E.lai = c->d_lai; E.nl = c->eco_nl; E.snap = F(c, QD_F_LAI_SNAP); E.fcanopy = F(c, QD_F_FCANOPY);
  for (int k = 0; k < 4; ++k) E.part[k] = c->d_part[k];
  E.ticket = c->d_ticket + 7 * c->batch;
  E.dt_hours = dt / 3600.0; E.every_hours = c->eco_every_hours; E.delta_thr = c->eco_delta; E.k_canopy = c->eco_k;
  return E;
}
// clock + recompute decision + conditional cache rebuild (no LAI manager bound: QD_ECO_USE_LAI=0, nothing to do)
static int eco_policy(qd_ctx* c, double dt) {
  if (!c->d_lai) return QD_OK;
  QdEcoArgs E = eco_args(c, dt);
  QD_KR(c, k_eco_stats, c->geo, E);
  QD_K(c, k_eco_canopy, c->geo, E);
  return QD_OK;
}
extern "C" int qd_eco_bind(qd_ctx* c, const double* lai, int nl, double k_canopy, double every_hours, double lai_delta, int every_nphys) {
  if (!c || (lai && nl < 1)) return QD_E_INVALID;
  qd_drop_graphs(c);
  c->d_lai = lai; c->eco_nl = lai ? nl : 0; c->eco_k = k_canopy; c->eco_every_hours = every_hours; c->eco_delta = lai_delta;
  c->eco_every_nphys = every_nphys > 1 ? every_nphys : 1;
  return QD_OK;
}
extern "C" int qd_eco_reset(qd_ctx* c, double hours, double next_hours, int cached, int step_count) {
  if (!c) return QD_E_INVALID;
  QD_BOUND(c);
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<double> s((size_t)c->batch * QD_S_COUNT);
  QD_CUDA(c, cudaMemcpy(s.data(), c->d_scal, s.size() * 8, cudaMemcpyDeviceToHost));
  for (int b = 0; b < c->batch; ++b) {
    double* S = s.data() + (size_t)b * QD_S_COUNT;
    S[QD_S_ECO_HOURS] = hours; S[QD_S_ECO_NEXT] = next_hours; S[QD_S_ECO_CACHED] = cached ? 1.0 : 0.0; S[QD_S_ECO_FLAG] = 0.0;
  }
  QD_CUDA(c, cudaMemcpy(c->d_scal, s.data(), s.size() * 8, cudaMemcpyHostToDevice));
  c->eco_steps = step_count; c->eco_have_alpha = 0;
  if (c->d_lai) {                                   // _lai_snapshot = total_LAI().copy()  (population.py:70)
    QdEcoArgs E = eco_args(c, 0.0);
    QD_K(c, k_eco_snapshot, c->geo, E);
    QD_CHECK_LAUNCH(c);
  }
  return QD_OK;
}
// Host-side counters of the ecology cadence (QD_ECO_SUBSTEP_EVERY_NPHYS): calls so far and whether an alpha map exists.
// Read (set = 0) or written (set = 1) by the checkpoint code.
extern "C" int qd_eco_state(qd_ctx* c, int* step_count, int* have_alpha, int set) {
  if (!c || !step_count || !have_alpha) return QD_E_INVALID;
  if (set) { c->eco_steps = *step_count; c->eco_have_alpha = *have_alpha ? 1 : 0; }
  else { *step_count = c->eco_steps; *have_alpha = c->eco_have_alpha; }
  return QD_OK;
}
extern "C" int qd_eco_subdaily(qd_ctx* c, const double* isr, double dt, double* alpha, int* produced) {
  if (!c || !isr) return QD_E_INVALID;
  QD_BOUND(c);
  int rc = eco_policy(c, dt); if (rc) return rc;
  c->eco_steps += 1;
  const int want = (c->eco_steps % std::max(1, c->eco_every_nphys) == 0) && alpha;
  QdEcoCellArgs A; A.isr = isr; A.fcanopy = F(c, QD_F_FCANOPY); A.eday = F(c, QD_F_EDAY); A.alpha = alpha;
  A.land = M(c, QD_M_LAND); A.dt = dt; A.want_alpha = want;
  QD_K(c, k_eco_cell, c->geo, A);
  QD_CHECK_LAUNCH(c);
  if (produced) *produced = want;
  return QD_OK;
}
// Individual pool: static per-individual tables (cell index, per-band weights, drought tolerance) and the per-star
// band spectra / Rayleigh factors (host NumPy, spectral.py:236-303); state (E_day, stress days) starts at zero.
extern "C" int qd_indiv_setup(qd_ctx* c, int n, int nb, const int* cell_host, const double* ab_host, const double* tol_host,
                              const double* spec_a, const double* spec_b, const double* t_ray) {
  if (!c || n < 1 || nb < 1 || nb > QD_INDIV_MAX_BANDS || !cell_host || !ab_host || !tol_host || !spec_a || !spec_b || !t_ray) return QD_E_INVALID;
  for (int k = 0; k < n; ++k) if (cell_host[k] < 0 || cell_host[k] >= c->ncell) return qd_fail(c, QD_E_INVALID, "qd_indiv_setup: cell index out of range", cudaSuccess);
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  QdIndivArgs& A = c->indiv;
  cudaFree((void*)A.cell); cudaFree((void*)A.ab); cudaFree((void*)A.tol); cudaFree(A.e_day); cudaFree(A.stress);
  memset(&A, 0, sizeof(A)); c->indiv_ready = 0;
  A.n = n; A.nb = nb;
  QD_CUDA(c, cudaMalloc((void**)&A.cell, (size_t)n * sizeof(int)));
  QD_CUDA(c, cudaMalloc((void**)&A.ab, (size_t)n * nb * 8));
  QD_CUDA(c, cudaMalloc((void**)&A.tol, (size_t)n * 8));
  QD_CUDA(c, cudaMalloc((void**)&A.e_day, (size_t)n * 8));
  QD_CUDA(c, cudaMalloc((void**)&A.stress, (size_t)n * 8));
  QD_CUDA(c, cudaMemcpy((void*)A.cell, cell_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  QD_CUDA(c, cudaMemcpy((void*)A.ab, ab_host, (size_t)n * nb * 8, cudaMemcpyHostToDevice));
  QD_CUDA(c, cudaMemcpy((void*)A.tol, tol_host, (size_t)n * 8, cudaMemcpyHostToDevice));
  QD_CUDA(c, cudaMemset(A.e_day, 0, (size_t)n * 8));
  QD_CUDA(c, cudaMemset(A.stress, 0, (size_t)n * 8));
  for (int k = 0; k < nb; ++k) { A.spec_a[k] = spec_a[k]; A.spec_b[k] = spec_b[k]; A.t_ray[k] = t_ray[k]; }
  c->indiv_ready = 1;
  return QD_OK;
}
// One fired sub-step (individuals.py:159-191) from the context's per-star insolation fields (member 0).
// soil_dev: [nlat][nlon] soil index or NULL for the scalar.
extern "C" int qd_indiv_substep(qd_ctx* c, const double* soil_dev, double soil_scalar, double period, double day_length) {
  if (!c || !c->indiv_ready) return QD_E_STATE;
  QD_BOUND(c);
  QdIndivArgs A = c->indiv;
  A.isr_a = F(c, QD_F_ISR_A); A.isr_b = F(c, QD_F_ISR_B); A.soil = soil_dev; A.soil_scalar = soil_scalar;
  A.period = period; A.stress_inc = period / day_length;
  QD_KG(c, k_indiv_substep, dim3((A.n + QD_THREADS - 1) / QD_THREADS), dim3(QD_THREADS), A);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
// upload = 0: device -> host, 1: host -> device (the daily ecology on the host resets / edits the buffers)
extern "C" int qd_indiv_state(qd_ctx* c, double* e_day_host, double* stress_host, int upload) {
  if (!c || !c->indiv_ready || !e_day_host || !stress_host) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  const size_t nb = (size_t)c->indiv.n * 8;
  if (upload) {
    QD_CUDA(c, cudaMemcpy(c->indiv.e_day, e_day_host, nb, cudaMemcpyHostToDevice));
    QD_CUDA(c, cudaMemcpy(c->indiv.stress, stress_host, nb, cudaMemcpyHostToDevice));
  } else {
    QD_CUDA(c, cudaMemcpy(e_day_host, c->indiv.e_day, nb, cudaMemcpyDeviceToHost));
    QD_CUDA(c, cudaMemcpy(stress_host, c->indiv.stress, nb, cudaMemcpyDeviceToHost));
  }
  return QD_OK;
}
// Global diagnostics of every member in one launch (qd_diag.cuh); out_host [B][QD_DIAG_COUNT].  Sync.
extern "C" int qd_diag_count(void) { return QD_DIAG_COUNT; }
extern "C" int qd_diag(qd_ctx* c, double* out_host) {
  if (!c || !out_host) return QD_E_INVALID;
  QD_BOUND(c);
  if (c->band_on) return qd_fail(c, QD_E_STATE, "latitude bands: use Engine.gather_rows for diagnostics", cudaSuccess);
  QdDiagArgs A; memset(&A, 0, sizeof(A));
  A.ts = F(c, QD_F_TS); A.h = F(c, QD_F_H); A.q = F(c, QD_F_Q); A.cloud = F(c, QD_F_CLOUD); A.hice = F(c, QD_F_HICE);
  A.wland = F(c, QD_F_WLAND); A.ssnow = F(c, QD_F_SSNOW); A.eflux = F(c, QD_F_EFLUX); A.precip = F(c, QD_F_PRECIP);
  A.rland = F(c, QD_F_RLAND); A.albedo = F(c, QD_F_ALBEDO); A.sst = F(c, QD_F_SST); A.isr = F(c, QD_F_ISR);
  A.cloud_eff = F(c, QD_F_CLOUD_EFF); A.lh = F(c, QD_F_LH); A.u = F(c, QD_F_U); A.v = F(c, QD_F_V);
  A.uo = F(c, QD_F_UO); A.vo = F(c, QD_F_VO); A.eta = F(c, QD_F_ETA); A.land = M(c, QD_M_LAND);
  A.has_cloud_eff = c->has_cloud_eff; A.part = c->d_diag_part; A.ticket = c->d_ticket + 2 * c->batch; A.out = c->h_diag_dev ? c->h_diag_dev : c->d_diag_out;
  QD_KR(c, k_diag, c->geo, A);
  QD_CHECK_LAUNCH(c);
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  // the last block wrote the 27 numbers per member into mapped pinned memory: one synchronisation, no copy call
  if (c->h_diag_dev) memcpy(out_host, c->h_diag, (size_t)c->batch * QD_DIAG_COUNT * 8);
  else QD_CUDA(c, cudaMemcpy(out_host, c->d_diag_out, (size_t)c->batch * QD_DIAG_COUNT * 8, cudaMemcpyDeviceToHost));
  return QD_OK;
}
// PhytoManager.advect_diffuse on S tracers [S][nlat][nlon] (in place).  uo / vo: device currents [nlat][nlon], or
// NULL for the context's own ocean state (member 0).
extern "C" int qd_phyto_advect_diffuse(qd_ctx* c, double* conc, int S, const double* uo, const double* vo, double dt,
                                       double adv_alpha, double k_h) {
  if (!c || !conc || S < 1) return QD_E_INVALID;
  QD_BOUND(c);
  if (c->band_on) return qd_fail(c, QD_E_STATE, "latitude bands: tracer transport is not partitioned", cudaSuccess);
  if (dt <= 0.0) return QD_OK;                                         // phyto.py:507-508
  const size_t need = (size_t)S * c->ncell;
  if (need > c->phyto_cap) {
    QD_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_phyto_tmp); c->d_phyto_tmp = nullptr; c->phyto_cap = 0;
    QD_CUDA(c, cudaMalloc((void**)&c->d_phyto_tmp, need * 8));
    c->phyto_cap = need;
  }
  QdPhytoArgs A; memset(&A, 0, sizeof(A));
  A.uo = uo ? uo : F(c, QD_F_UO); A.vo = vo ? vo : F(c, QD_F_VO); A.C = conc; A.tmp = c->d_phyto_tmp; A.land = M(c, QD_M_LAND);
  A.dt = dt; A.alpha = adv_alpha; A.dtkh = k_h > 0.0 ? dt * k_h : 0.0;
  QD_KG(c, k_phyto_advect, dim3(c->nblk, S), dim3(QD_THREADS), c->geo, A);
  QD_KG(c, k_phyto_finish, dim3(c->nblk, S), dim3(QD_THREADS), c->geo, A);
  QD_KG(c, k_phyto_polar, dim3(2, S), dim3(QD_THREADS), c->geo, A);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
extern "C" int qd_eco_bands(qd_ctx* c, int nb, const double* r_eff, double soil, double* out) {
  if (!c || nb < 1 || nb > QD_ECO_MAX_BANDS || !r_eff || !out) return QD_E_INVALID;
  QD_BOUND(c);
  QdEcoBandArgs A; memset(&A, 0, sizeof(A));
  A.fcanopy = F(c, QD_F_FCANOPY); A.land = M(c, QD_M_LAND); A.out = out; A.nb = nb; A.soil = soil;
  for (int k = 0; k < nb; ++k) A.r_eff[k] = r_eff[k];
  QD_K(c, k_eco_bands, c->geo, A);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}

// The per-member table of parameter-only divisors (QdUdivId in qd_ops.cuh): every entry is the reference's divisor
// expression evaluated in fp64 on the host plus its correctly rounded reciprocal, rebuilt only when the parameters or
// dt change.  Called by the step entry points before anything is enqueued or captured.
static int qd_derive(qd_ctx* c, double dt) {
  if (c->udiv_valid && c->udiv_dt == dt) return QD_OK;
  std::vector<QdRcp> T((size_t)c->batch * QD_U_COUNT);
  for (int b = 0; b < c->batch; ++b) {
    const double* P = c->h_prm + (size_t)b * QD_P_COUNT;
    QdRcp* D = T.data() + (size_t)b * QD_U_COUNT;
    D[QD_U_MCOL] = qd_rcp(fmax(1e-6, P[QD_P_RHO_A] * P[QD_P_H_MBL]));
    D[QD_U_TAU_COND] = qd_rcp(fmax(1e-6, P[QD_P_TAU_COND]));
    D[QD_U_RHO_SNOW] = qd_rcp(qd_max(P[QD_P_RHO_SNOW], 1e-6));
    D[QD_U_1000] = qd_rcp(1000.0);
    D[QD_U_SNOW_BAND] = qd_rcp(qd_max(1e-6, P[QD_P_SNOW_T_BAND]));
    D[QD_U_DT] = qd_rcp(dt);
    D[QD_U_SWE_REF] = qd_rcp(qd_max(1e-6, P[QD_P_SWE_REF]));
    D[QD_U_SIGMA] = qd_rcp(QD_SIGMA_SB);
    D[QD_U_RUNOFF_TAU] = qd_rcp(fmax(1.0, P[QD_P_RUNOFF_TAU_DAYS] * 86400.0));
    D[QD_U_C_SFC] = qd_rcp(fmax(1e-12, P[QD_P_C_SFC]));
    D[QD_U_TAU_RAD] = qd_rcp(P[QD_P_TAU_RAD]);
    D[QD_U_HICE_REF] = qd_rcp(qd_max(1e-6, P[QD_P_HICE_REF]));
    D[QD_U_RHO_I_LF] = qd_rcp(P[QD_P_RHO_I] * P[QD_P_L_F]);
    D[QD_U_CS_LAND] = qd_rcp(qd_seaice_cs(P[QD_P_CS_LAND]));
    D[QD_U_CS_ICE] = qd_rcp(qd_seaice_cs(P[QD_P_CS_ICE]));
    D[QD_U_CS_OCEAN] = qd_rcp(qd_seaice_cs(P[QD_P_CS_OCEAN]));
    D[QD_U_ATM] = qd_rcp(fmax(1e-6, P[QD_P_RHO_A]) * fmax(1.0, P[QD_P_ATM_H]) * P[QD_P_G]);
    D[QD_U_12K] = qd_rcp(12.0);
    D[QD_U_ADV_REF] = qd_rcp(2e-5);
    D[QD_U_2DY] = qd_rcp(2 * (c->geo.dlat * c->geo.a));
    D[QD_Q_G_CP].b = D[QD_Q_G_CP].r = P[QD_P_G] / 1004.0;
    D[QD_Q_R_G].b = D[QD_Q_R_G].r = 287 / P[QD_P_G];
    D[QD_Q_DDF].b = D[QD_Q_DDF].r = P[QD_P_SNOW_DDF] / 86400.0;
    D[QD_Q_MELT].b = D[QD_Q_MELT].r = P[QD_P_SNOW_MELT_RATE] / 86400.0;
  }
  QD_CUDA(c, cudaStreamSynchronize(c->stream));          // a step in flight may still be reading the old table
  QD_CUDA(c, cudaMemcpy(c->d_udiv, T.data(), T.size() * sizeof(QdRcp), cudaMemcpyHostToDevice));
  c->udiv_dt = dt; c->udiv_valid = 1;
  return QD_OK;
}

static int atmos_core(qd_ctx* c, const qd_step_cfg_t* cfg, int mode_loop) {
  const double dt = cfg->dt;
  const int has_alb = mode_loop ? cfg->loop_with_albedo : cfg->has_albedo;
  QdColArgs A; memset(&A, 0, sizeof(A));
  A.u = F(c, QD_F_U); A.v = F(c, QD_F_V); A.h = F(c, QD_F_H); A.ts = F(c, QD_F_TS); A.q = F(c, QD_F_Q);
  A.cloud = F(c, QD_F_CLOUD); A.hice = F(c, QD_F_HICE); A.wland = F(c, QD_F_WLAND); A.ssnow = F(c, QD_F_SSNOW);
  A.eday = F(c, QD_F_EDAY);
  A.ts_pre = F(c, QD_F_X2); A.q_pre = F(c, QD_F_X3);
  A.isr = F(c, QD_F_ISR); A.isr_a = F(c, QD_F_ISR_A); A.isr_b = F(c, QD_F_ISR_B); A.olr = F(c, QD_F_OLR);
  A.eflux = F(c, QD_F_EFLUX); A.pcond = F(c, QD_F_PCOND); A.lh = F(c, QD_F_LH); A.lhrel = F(c, QD_F_LHREL);
  A.albedo = F(c, QD_F_ALBEDO); A.teq = F(c, QD_F_TEQ); A.csnow = F(c, QD_F_CSNOW); A.rland = F(c, QD_F_RLAND);
  A.alpha_eco = F(c, QD_F_ALPHA_ECO);
  A.precip = F(c, QD_F_PRECIP); A.cloud_eff = F(c, QD_F_CLOUD_EFF); A.base_albedo = F(c, QD_F_BASE_ALBEDO);
  A.elevation = F(c, QD_F_ELEVATION); A.fcanopy = F(c, QD_F_FCANOPY); A.hcos = c->d_hcos;
  A.land = M(c, QD_M_LAND); A.glacier = M(c, QD_M_GLACIER);
  A.forcing = c->d_forcing; A.step_idx = c->d_step_idx;
  A.dt = dt; A.mode_loop = mode_loop; A.has_albedo = has_alb; A.has_cloud_eff = c->has_cloud_eff;
  A.with_hydrology = cfg->with_hydrology; A.with_eco = 0; A.store_isr_ab = cfg->store_isr_ab;
  if (cfg->with_eco && mode_loop) {
    // adapter.py:148-157: the adapter's call counter decides whether a new alpha map is produced this step;
    // otherwise the script keeps using the last one (run_simulation.py:2083-2086)
    int rc = eco_policy(c, dt); if (rc) return rc;
    c->eco_steps += 1;
    if (c->eco_steps % std::max(1, c->eco_every_nphys) == 0) { A.with_eco = 1; c->eco_have_alpha = 1; }
    else A.with_eco = c->eco_have_alpha ? 2 : 3;
  }
  if (cfg->with_eco && c->band_on) return qd_fail(c, QD_E_STATE, "latitude bands: the ecology coupling is not partitioned yet", cudaSuccess);
  BP(c, BL({A.u, 0}, {A.v, 0}, {A.h, 0}, {A.ts, 0}, {A.q, 0}, {A.cloud, 0}, {A.hice, 0}, {A.wland, 0}, {A.ssnow, 0}, {A.eday, 0},
           {A.precip, 0}, {A.cloud_eff, 0}, {A.base_albedo, 0}, {A.elevation, 0}, {A.fcanopy, 0}, {A.land, 0}, {A.pcond, 0},
           {A.isr_a, 0}, {A.isr_b, 0}, {A.alpha_eco, 0}, {A.eflux, 0}, {A.lh, 0}, {A.lhrel, 0}),
     BL(A.h, A.ts, A.q, A.hice, A.wland, A.ssnow, A.eday, A.ts_pre, A.q_pre, A.isr, A.isr_a, A.isr_b, A.olr, A.eflux, A.pcond, A.lh,
        A.lhrel, A.albedo, A.teq, A.csnow, A.rland, A.alpha_eco, A.glacier));
  QD_K(c, k_column, c->geo, A);

  if (has_alb) {
    bool need_median = false;
    for (int b = 0; b < c->batch; ++b) {
      const double* P = c->h_prm + (size_t)b * QD_P_COUNT;
      if (P[QD_P_CLOUD_COUPLE] != 0.0 && P[QD_P_PCOND_REF] != P[QD_P_PCOND_REF]) need_median = true;
    }
    if (need_median) { int rc = op_median(c, F(c, QD_F_PCOND), 1e-6, c->d_scal + QD_S_PREF_ATM, c->d_scal + QD_S_CNT_PCOND, QD_S_COUNT, 2); if (rc) return rc; }
    QdEnergyArgs E; memset(&E, 0, sizeof(E));
    E.h = F(c, QD_F_H); E.hice = F(c, QD_F_HICE); E.ts_pre = F(c, QD_F_X2); E.olr = F(c, QD_F_OLR); E.cloud_eff = F(c, QD_F_CLOUD_EFF);
    E.ts = F(c, QD_F_TS); E.q_pre = F(c, QD_F_X3); E.cloud = F(c, QD_F_CLOUD); E.pcond = F(c, QD_F_PCOND); E.isr = F(c, QD_F_ISR);
    E.albedo = F(c, QD_F_ALBEDO); E.teq = F(c, QD_F_TEQ); E.u = F(c, QD_F_U); E.v = F(c, QD_F_V); E.lh = F(c, QD_F_LH); E.lhrel = F(c, QD_F_LHREL);
    E.cs_map = F(c, QD_F_CS_MAP); E.land = M(c, QD_M_LAND); E.dt = dt;
    BP(c, BL({E.h, 0}, {E.hice, 0}, {E.ts_pre, 0}, {E.olr, 0}, {E.cloud_eff, 0}, {E.ts, 0}, {E.q_pre, 0}, {E.cloud, 0}, {E.pcond, 0},
             {E.isr, 0}, {E.albedo, 0}, {E.teq, 0}, {E.u, 0}, {E.v, 0}, {E.lh, 0}, {E.lhrel, 0}, {E.cs_map, 0}, {E.land, 0}),
       BL(E.h, E.hice, E.ts_pre, E.olr, E.cloud_eff));
    QD_K(c, k_energy, c->geo, E);
    c->has_cloud_eff = 1;
  }
  c->atm_counter += 1;                       // dynamics.py:451, before the cadence tests
  const int sc = c->atm_counter;

  QdAdvMomArgs AM; memset(&AM, 0, sizeof(AM));
  AM.ts_pre = F(c, QD_F_X2); AM.q_pre = F(c, QD_F_X3); AM.h = F(c, QD_F_H); AM.friction = F(c, QD_F_FRICTION);
  AM.ts = F(c, QD_F_TS); AM.q = F(c, QD_F_Q); AM.u = F(c, QD_F_U); AM.v = F(c, QD_F_V); AM.dt = dt;
  BP(c, BL({AM.ts_pre, band_radv(c, dt)}, {AM.q_pre, band_radv(c, dt)}, {AM.h, 1}, {AM.friction, 0}, {AM.u, 0}, {AM.v, 0}, {AM.ts, 0}, {AM.q, 0}),
     BL(AM.ts, AM.q, AM.u, AM.v));
  QD_K(c, k_advect_momentum, c->geo, AM);

  // From here on u, v, h, q, cloud may live in scratch slots (the fused del^4 kernel is out of place);
  // cur[] tracks where each field currently is, k_tail writes everything back to its home slot.
  const double* cos_lap = ROW(c, QD_R_COS_LAP_ATM);
  enum { iU, iV, iH, iQ, iC };
  double* home[5] = {F(c, QD_F_U), F(c, QD_F_V), F(c, QD_F_H), F(c, QD_F_Q), F(c, QD_F_CLOUD)};
  double* alt[5] = {F(c, QD_F_X4), F(c, QD_F_X5), F(c, QD_F_X6), F(c, QD_F_X7), F(c, QD_F_X8)};
  double* cur[5] = {home[0], home[1], home[2], home[3], home[4]};
  auto other = [&](int k) { return cur[k] == home[k] ? alt[k] : home[k]; };
  if (cfg->diff_enable && (sc % std::max(1, cfg->diff_every) == 0)) {
    const double* kr[5] = {ROW(c, QD_R_K4_U), ROW(c, QD_R_K4_V), ROW(c, QD_R_K4_H), ROW(c, QD_R_K4_Q), ROW(c, QD_R_K4_C)};
    auto run = [&](const int* ids, int n, int nsub) -> int {
      double* fl[5]; double* ot[5]; const double* kk[5];
      for (int k = 0; k < n; ++k) { fl[k] = cur[ids[k]]; ot[k] = other(ids[k]); kk[k] = kr[ids[k]]; }
      bool moved = false;
      int rc = op_hyper(c, n, fl, ot, kk, nullptr, dt, nsub, cos_lap, &moved);
      if (rc) return rc;
      if (moved) for (int k = 0; k < n; ++k) cur[ids[k]] = ot[k];
      return QD_OK;
    };
    int ids[5] = {iU, iV, iH, 0, 0}; int n = 3, rc;
    if (cfg->k4_nsub <= 1) {
      if (cfg->apply_q) ids[n++] = iQ;
      if (cfg->apply_cloud) ids[n++] = iC;
      if ((rc = run(ids, n, 1))) return rc;
    } else {
      if ((rc = run(ids, 3, cfg->k4_nsub))) return rc;
      int id2[2]; int m = 0;
      if (cfg->apply_q) id2[m++] = iQ;
      if (cfg->apply_cloud) id2[m++] = iC;
      if (m && (rc = run(id2, m, 1))) return rc;
    }
  }
  if (cfg->shapiro_every > 0 && (sc % cfg->shapiro_every == 0)) {
    double* fl[3] = {cur[iU], cur[iV], cur[iH]};
    double* sx[3] = {F(c, QD_F_X0), F(c, QD_F_X1), F(c, QD_F_X2)};
    int rc = op_shapiro(c, 3, fl, sx, cfg->shapiro_n); if (rc) return rc;
    double* f2[2]; int n = 0;
    if (cfg->shapiro_q) f2[n++] = cur[iQ];
    if (cfg->shapiro_cloud) f2[n++] = cur[iC];
    if (n) { rc = op_shapiro(c, n, f2, sx, std::max(1, cfg->shapiro_n - 1)); if (rc) return rc; }
  }
  if (cfg->spec_every > 0 && (sc % cfg->spec_every == 0)) {
    int rc;
    if ((rc = op_bandstop(c, cur[iU], cfg->spec_cutoff, cfg->spec_damp))) return rc;
    if ((rc = op_bandstop(c, cur[iV], cfg->spec_cutoff, cfg->spec_damp))) return rc;
    if ((rc = op_bandstop(c, cur[iH], cfg->spec_cutoff, cfg->spec_damp))) return rc;
  }
  // cloud tail: advect with the NEW winds, then dissipation / damping / hygiene (+ Q_net in loop mode)
  {
    QdFields f = mk_fields(1); f.src[0] = cur[iC]; f.dst[0] = F(c, QD_F_X9);
    BP(c, BL({f.src[0], band_radv(c, dt)}, {cur[iU], 0}, {cur[iV], 0}), BL(f.dst[0]));
    QD_K(c, k_advect, c->geo, f, cur[iU], cur[iV], dt, ROW(c, QD_R_COS_ADV_ATM), ROW(c, QD_R_IAC_ADV_ATM));
    QdTailArgs T; memset(&T, 0, sizeof(T));
    T.u_in = cur[iU]; T.v_in = cur[iV]; T.h_in = cur[iH]; T.q_in = cur[iQ];
    T.u = F(c, QD_F_U); T.v = F(c, QD_F_V); T.h = F(c, QD_F_H); T.ts = F(c, QD_F_TS); T.q = F(c, QD_F_Q); T.cloud = F(c, QD_F_CLOUD);
    T.cloud_adv = F(c, QD_F_X9); T.hice = F(c, QD_F_HICE); T.isr = F(c, QD_F_ISR); T.albedo = F(c, QD_F_ALBEDO);
    T.cloud_eff = F(c, QD_F_CLOUD_EFF); T.lh = F(c, QD_F_LH); T.uo = F(c, QD_F_UO); T.vo = F(c, QD_F_VO);
    T.qnet = F(c, QD_F_QNET); T.ice = M(c, QD_M_ICE); T.land = M(c, QD_M_LAND);
    T.part_max_u = c->d_part[0]; T.part_max_va = c->d_part[1]; T.ticket = c->d_ticket + 3 * c->batch;
    T.dt = dt; T.with_qnet = (mode_loop && cfg->with_ocean) ? 1 : 0; T.has_cloud_eff = c->has_cloud_eff;
    // loop mode on one GPU: the ocean step's preparation rides along (ocean_core then skips k_ocean_prep); latitude bands
    // keep the separate kernel (n_sub needs the all-reduced maxima)
    T.with_max = (mode_loop && cfg->with_ocean && !c->band_on && !getenv("QD_NO_TAILPREP")) ? 1 : 0;
    T.taux = F(c, QD_F_X0); T.tauy = F(c, QD_F_X1); T.sub_ctr = c->d_sub_ctr;
    c->tail_did_prep = T.with_max;
    BP(c, BL({T.u_in, 0}, {T.v_in, 0}, {T.h_in, 0}, {T.q_in, 0}, {T.ts, 0}, {T.cloud_adv, 0}, {T.hice, 0}, {T.isr, 0}, {T.albedo, 0},
             {T.cloud_eff, 0}, {T.lh, 0}, {T.uo, 0}, {T.vo, 0}, {T.land, 0}, {T.qnet, 0}, {T.ice, 0}),
       BL(T.u, T.v, T.h, T.ts, T.q, T.cloud, T.qnet, T.ice));
    QD_KR(c, k_tail, c->geo, T);
  }
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}

extern "C" int qd_atmos_step(qd_ctx* c, const qd_step_cfg_t* cfg) {
  if (!c || !cfg) return QD_E_INVALID;
  QD_BOUND(c);
  { const int rc = qd_derive(c, cfg->dt); if (rc) return rc; }
  return atmos_core(c, cfg, 0);
}

#ifndef QD_HOST_EMU
// Fused form of one CFL sub-step (qd_ocean_fused.cuh): available on one rank, with the default del^4 cadence (one
// application per sub-step, no ocean Shapiro), on grids large enough to fill the machine with strip warps.
// Rows per warp chunk of the fused sub-step, 0 = use the four-kernel form.  A warp streams R output rows after a
// 10-row warm-up; blocks of 4 warps run 3 per SM.  R is chosen to minimise (waves of blocks) x (R + 10): one full wave
// where the grid allows it (1441x2880: 14 chunks of 102 rows = 420 blocks on 444 slots).
static int ocean_fused_rows(qd_ctx* c, const qd_step_cfg_t* cfg, bool do_hyper, bool do_shap) {
  if (!c->ocean_fused_enable || c->band_on || !c->h4_stream || !do_hyper || do_shap || cfg->oc_k4_nsub > 1) return 0;
  if (c->nlat < 96 || c->nlon < 64) return 0;
  const int ntj8 = (c->nlat + 7) / 8, ja = 8, jb = (ntj8 - 2) * 8, rows = jb - ja;
  const long long strips = (long long)((c->nlon + QD_OF_COLS - 1) / QD_OF_COLS) * c->batch;
  if (strips * ((rows + 31) / 32) < 2 * 148) return 0;                  // too few strip warps even with 32-row chunks
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  const long long slots = 3LL * sms;
  if (const char* ov = getenv("QD_OCEAN_FUSED_ROWS")) { const int r = atoi(ov); if (r >= 16) return std::min(r, rows); }   // tuning override
  int best = 0; long long best_cost = 1LL << 60;
  for (int nch = 1; nch <= rows / 16; ++nch) {
    const int R = (rows + nch - 1) / nch;
    const long long blocks = (((long long)((c->nlon + QD_OF_COLS - 1) / QD_OF_COLS) * ((rows + R - 1) / R) + QD_OF_WARPS - 1) / QD_OF_WARPS) * c->batch;
    const long long cost = ((blocks + slots - 1) / slots) * (R + 10);
    if (cost < best_cost) { best_cost = cost; best = R; }
  }
  return best;
}
// c->geo restricted to rows [a0, a1) u [b0, b1)
static QdGeo geo_rows(qd_ctx* c, int a0, int a1, int b0, int b1, int* nblk) {
  QdGeo g = c->geo;
  g.sa0 = a0; g.sa1 = a1; g.sb0 = b0; g.sb1 = b1;
  g.ncomp = ((a1 - a0) + (b1 - b0)) * c->nlon;
  *nblk = std::max(1, (g.ncomp + QD_THREADS - 1) / QD_THREADS);
  return g;
}
static int ocean_substep_fused(qd_ctx* c, const qd_step_cfg_t* cfg, int inject, int R) {
  const double* P = c->h_prm;
  const int ovu = P[QD_P_OC_K4_U] == P[QD_P_OC_K4_U], ovv = P[QD_P_OC_K4_V] == P[QD_P_OC_K4_V], ove = P[QD_P_OC_K4_ETA] == P[QD_P_OC_K4_ETA];
  QdSubCtl sc{c->d_sub_ctr};
  const int nlat = c->nlat, ntj8 = (nlat + 7) / 8, ja = 8, jb = (ntj8 - 2) * 8;
  const int nstrips = (c->nlon + QD_OF_COLS - 1) / QD_OF_COLS;
  const int nwarps = nstrips * ((jb - ja + R - 1) / R);
  const int nblocks = (nwarps + QD_OF_WARPS - 1) / QD_OF_WARPS;
  const int pole_nvb = std::max(1, std::min(c->nblk, QD_NVB_MAX));
  const int npart = pole_nvb + nblocks * QD_OF_WARPS;
  if (npart > c->oc_npart) return qd_fail(c, QD_E_STATE, "fused ocean sub-step: partial-sum table too small", cudaSuccess);   // sized in qd_create (no allocation inside a captured step)
  double* uo2[2] = {F(c, QD_F_UO), F(c, QD_F_X9)};
  double* vo2[2] = {F(c, QD_F_VO), F(c, QD_F_X10)};
  int nb;
  // ---- pole pass: momentum rows [0, 13) u [jb-5, nlat) -> del^4 rows [0, 9) u [jb-1, nlat) -> continuity rows [0, 8) u [jb, nlat)
  {
    QdOcMomArgs Mo; memset(&Mo, 0, sizeof(Mo));
    Mo.eta = F(c, QD_F_ETA); Mo.uo = uo2[0]; Mo.vo = vo2[0]; Mo.uo_alt = uo2[1]; Mo.vo_alt = vo2[1];
    Mo.taux = F(c, QD_F_X0); Mo.tauy = F(c, QD_F_X1); Mo.ub = F(c, QD_F_X2); Mo.vb = F(c, QD_F_X3); Mo.land = M(c, QD_M_LAND);
    const QdGeo gm = geo_rows(c, 0, 13, jb - 5, nlat, &nb);
    QD_KG(c, k_ocean_momentum, dim3(nb, c->batch), dim3(QD_THREADS), gm, Mo, sc);
  }
  QdHyper4Args H; memset(&H, 0, sizeof(H));
  H.n = 3; H.cosr = ROW(c, QD_R_COS_ADV_HALF); H.dt = 0.0; H.nsub = 1; H.ocean = 1; H.sc = sc;
  {
    const double* src3[3] = {F(c, QD_F_X2), F(c, QD_F_X3), F(c, QD_F_ETA)};
    double* dst3[3] = {F(c, QD_F_X4), F(c, QD_F_X5), F(c, QD_F_X6)};
    for (int k = 0; k < 3; ++k) { H.src[k] = src3[k]; H.dst[k] = dst3[k]; H.scale[k] = 1.0; H.k4_bstride[k] = c->geo.row_bstride; }
    H.k4rows[0] = ovu ? QD_USER_ROW(c, 2) : ROW(c, QD_R_OC_S4DX4); H.raw_k4[0] = ovu;
    H.k4rows[1] = ovv ? QD_USER_ROW(c, 3) : ROW(c, QD_R_OC_S4DX4); H.raw_k4[1] = ovv;
    H.k4rows[2] = ove ? QD_USER_ROW(c, 4) : ROW(c, QD_R_OC_S4DX4); H.raw_k4[2] = ove;
    if (!ove) H.scale[2] = 0.5;                                       // ocean.py:352
    H.tj_lo = 1 << 30; H.tj_skip = 0; H.ja = 0; H.jb = 0;
    const int tiles_i = (c->nlon + QD_H4_TI - 1) / QD_H4_TI;
    H.row0 = 0; H.row1 = 9;
    QD_KGN(c, "k_hyper4_tile<8>[3]", k_hyper4_tile<8>, dim3(tiles_i * 2, c->batch, 3), dim3(QD_H4_NX, QD_H4_NY), c->geo, H);
    H.row0 = jb - 1; H.row1 = nlat;
    QD_KGN(c, "k_hyper4_tile<8>[3]", k_hyper4_tile<8>, dim3(tiles_i * ((nlat - (jb - 1) + 7) / 8), c->batch, 3), dim3(QD_H4_NX, QD_H4_NY), c->geo, H);
  }
  {
    QdOcContPoleArgs Cp; memset(&Cp, 0, sizeof(Cp));
    Cp.ub = F(c, QD_F_X4); Cp.vb = F(c, QD_F_X5); Cp.eta_in = F(c, QD_F_X6); Cp.sst = F(c, QD_F_SST);
    Cp.uo[0] = uo2[0]; Cp.uo[1] = uo2[1]; Cp.vo[0] = vo2[0]; Cp.vo[1] = vo2[1];
    Cp.eta_out = F(c, QD_F_X7); Cp.tb = F(c, QD_F_X8); Cp.part = c->d_oc_part; Cp.npart = npart; Cp.land = M(c, QD_M_LAND);
    QdGeo gp = geo_rows(c, 0, 8, jb, nlat, &nb);
    gp.nvb = pole_nvb;
    QD_KG(c, k_ocean_cont_pole, dim3(std::min(nb, gp.nvb), c->batch), dim3(QD_THREADS), gp, Cp, sc);
  }
  // ---- rows [8, jb): the streaming kernel; its last block totals the eta sum (pole partials included)
  {
    QdOcFusedArgs A; memset(&A, 0, sizeof(A));
    A.uo[0] = uo2[0]; A.uo[1] = uo2[1]; A.vo[0] = vo2[0]; A.vo[1] = vo2[1];
    A.eta = F(c, QD_F_ETA); A.taux = F(c, QD_F_X0); A.tauy = F(c, QD_F_X1); A.sst = F(c, QD_F_SST);
    A.eta_out = F(c, QD_F_X7); A.tb = F(c, QD_F_X8); A.land = M(c, QD_M_LAND);
    A.k4tab = c->d_oc_k4;
    A.part = c->d_oc_part; A.part_off = pole_nvb; A.npart = npart; A.ticket = c->d_ticket + 5 * c->batch;
    A.ja = ja; A.jb = jb;
    A.R = R;
    QD_KG(c, k_ocean_fused, dim3(nblocks, c->batch), dim3(32 * QD_OF_WARPS), c->geo, A, sc);
  }
  {
    QdOcCloseArgs Cl; memset(&Cl, 0, sizeof(Cl));
    Cl.tb = F(c, QD_F_X8); Cl.eta_mid = F(c, QD_F_X7); Cl.qnet = F(c, QD_F_QNET);
    Cl.uo[0] = uo2[0]; Cl.uo[1] = uo2[1]; Cl.vo[0] = vo2[0]; Cl.vo[1] = vo2[1];
    Cl.sst = F(c, QD_F_SST); Cl.ts_atm = F(c, QD_F_TS); Cl.eta = F(c, QD_F_ETA);
    Cl.land = M(c, QD_M_LAND); Cl.ice = M(c, QD_M_ICE);
    Cl.has_q = cfg->oc_has_q; Cl.has_ice = cfg->oc_has_ice; Cl.inject = inject;
    QD_K(c, k_ocean_close, c->geo, Cl, sc);
  }
  return QD_OK;
}
#endif

// ------------------------------------------------------------------------------ ocean step
// One CFL sub-step body (ocean.py:305-444).  Launched either from a host loop (stream mode) or captured
// once as the body of a CUDA-graph WHILE node (graph mode); the sub-step index is read from device memory.
// Closing pass of sub-step s fused with the momentum step of sub-step s+1 (qd_ocean.cuh: QdOcSstBArgs::fuse_mom): the
// loop body is del^4 -> continuity -> closing+momentum, one momentum launch in front of the loop does sub-step 0.  Saves
// one kernel and ~40 B per cell and sub-step (eta, uo, vo are not re-read, the home currents are only stored at the end).
// Needs the two-cells-per-thread kernels, one del^4 pass per sub-step (its outputs must not alias the momentum outputs),
// no latitude bands (their halo exchange sits between the closing pass and the momentum step).
static bool ocean_fm(qd_ctx* c, const qd_step_cfg_t* cfg, bool do_hyper, bool do_shap) {
#ifdef QD_HOST_EMU
  (void)c; (void)cfg; (void)do_hyper; (void)do_shap;
  return false;
#else
  if (c->band_on || !do_hyper || do_shap || cfg->oc_k4_nsub > 1 || (c->nlon & 1) || getenv("QD_NO_PAIRS") || getenv("QD_NO_FM")) return false;
  return ocean_fused_rows(c, cfg, do_hyper, do_shap) == 0;
#endif
}
static int ocean_momentum_launch(qd_ctx* c) {
  QdSubCtl sc{c->d_sub_ctr};
  QdOcMomArgs Mo; memset(&Mo, 0, sizeof(Mo));
  Mo.eta = F(c, QD_F_ETA); Mo.uo = F(c, QD_F_UO); Mo.vo = F(c, QD_F_VO); Mo.taux = F(c, QD_F_X0); Mo.tauy = F(c, QD_F_X1);
  Mo.ub = F(c, QD_F_X2); Mo.vb = F(c, QD_F_X3); Mo.land = M(c, QD_M_LAND);
  BP(c, BL({Mo.eta, 1}, {Mo.uo, 0}, {Mo.vo, 0}, {Mo.taux, 0}, {Mo.tauy, 0}, {Mo.land, 0}), BL(Mo.ub, Mo.vb));
  QD_K(c, k_ocean_momentum, c->geo, Mo, sc);
  return QD_OK;
}
static int ocean_substep_body(qd_ctx* c, const qd_step_cfg_t* cfg, int inject, bool do_hyper, bool do_shap, bool first_in_group = true) {
#ifndef QD_HOST_EMU
  const int fused_rows = ocean_fused_rows(c, cfg, do_hyper, do_shap);
  c->ocean_fused = fused_rows > 0;
  if (c->ocean_fused) return ocean_substep_fused(c, cfg, inject, fused_rows);
#endif
  const double* P = c->h_prm;   // K4 overrides are shared by all ensemble members (switch-like)
  const int ovu = P[QD_P_OC_K4_U] == P[QD_P_OC_K4_U], ovv = P[QD_P_OC_K4_V] == P[QD_P_OC_K4_V], ove = P[QD_P_OC_K4_ETA] == P[QD_P_OC_K4_ETA];
  QdSubCtl sc{c->d_sub_ctr};
  QdOcMomArgs Mo; memset(&Mo, 0, sizeof(Mo));
  Mo.eta = F(c, QD_F_ETA); Mo.uo = F(c, QD_F_UO); Mo.vo = F(c, QD_F_VO); Mo.taux = F(c, QD_F_X0); Mo.tauy = F(c, QD_F_X1);
  Mo.ub = F(c, QD_F_X2); Mo.vb = F(c, QD_F_X3); Mo.land = M(c, QD_M_LAND);
  if (c->band_on && first_in_group) {   // loop-carried fields start every GROUP of sub-steps from "own rows only": the body replays unchanged (WHILE node)
    for (int id : {(int)QD_F_UO, (int)QD_F_VO, (int)QD_F_ETA, (int)QD_F_SST, (int)QD_F_TS, (int)QD_F_X2, (int)QD_F_X3, (int)QD_F_X4,
                   (int)QD_F_X5, (int)QD_F_X6, (int)QD_F_X7, (int)QD_F_X8, (int)QD_F_X9}) c->band_valid[id] = 0;
    // ONE exchange per sub-step: everything the body's stencils read with a halo (eta: momentum and del^4; currents: del^4,
    // divergence; SST: gather and diffusion).  The wind stress was exchanged once in ocean_core.
    { int rcb = band_hint(c, {QD_F_ETA, QD_F_UO, QD_F_VO, QD_F_SST}); if (rcb) return rcb; }
  }
  const bool fm = ocean_fm(c, cfg, do_hyper, do_shap);
  if (!fm) {
    BP(c, BL({Mo.eta, 1}, {Mo.uo, 0}, {Mo.vo, 0}, {Mo.taux, 0}, {Mo.tauy, 0}, {Mo.land, 0}), BL(Mo.ub, Mo.vb));
    QD_K(c, k_ocean_momentum, c->geo, Mo, sc);
  }
  double* ub = F(c, QD_F_X2); double* vb = F(c, QD_F_X3); double* eta_cur = F(c, QD_F_ETA);
  if (do_hyper) {
    const int ns = std::max(1, cfg->oc_k4_nsub);
    double* A3[3] = {F(c, QD_F_X2), F(c, QD_F_X3), F(c, QD_F_ETA)};
    double* B3[3] = {F(c, QD_F_X4), F(c, QD_F_X5), F(c, QD_F_X6)};
    bool inB = false;
    for (int q = 0; q < ns; ++q) {
      QdHyper4Args H; memset(&H, 0, sizeof(H));
      H.n = 3; H.cosr = ROW(c, QD_R_COS_ADV_HALF); H.dt = 0.0; H.nsub = ns; H.ocean = 1; H.sc = sc;
      for (int k = 0; k < 3; ++k) { H.src[k] = inB ? B3[k] : A3[k]; H.dst[k] = inB ? A3[k] : B3[k]; H.scale[k] = 1.0; }
      // rows: sigma4*dx^4 (divided by sub_dt on device) or user rows 2..4 holding the QD_OCEAN_K4_* overrides
      H.k4rows[0] = ovu ? QD_USER_ROW(c, 2) : ROW(c, QD_R_OC_S4DX4); H.raw_k4[0] = ovu;
      H.k4rows[1] = ovv ? QD_USER_ROW(c, 3) : ROW(c, QD_R_OC_S4DX4); H.raw_k4[1] = ovv;
      H.k4rows[2] = ove ? QD_USER_ROW(c, 4) : ROW(c, QD_R_OC_S4DX4); H.raw_k4[2] = ove;
      H.k4_bstride[0] = H.k4_bstride[1] = H.k4_bstride[2] = c->geo.row_bstride;
      if (!ove) H.scale[2] = 0.5;                                       // ocean.py:352
      int rc = launch_hyper4(c, H); if (rc) return rc;
      inB = !inB;
    }
    if (inB) { ub = B3[0]; vb = B3[1]; eta_cur = B3[2]; }
  }
  if (do_shap) {
    // optional (QD_OCEAN_SHAPIRO_N, off by default); valid when every member shares n_sub
    double* fl[3] = {ub, vb, eta_cur};
    double* sx[3] = {F(c, QD_F_X7), F(c, QD_F_X8), F(c, QD_F_X9)};
    int rc = op_shapiro(c, 3, fl, sx, cfg->oc_shapiro_n); if (rc) return rc;
  }
  QdOcContArgs Co; memset(&Co, 0, sizeof(Co));
  Co.ub = ub; Co.vb = vb; Co.eta_in = eta_cur; Co.eta = fm ? F(c, QD_F_X8) : F(c, QD_F_ETA); Co.part = c->d_part[2]; Co.land = M(c, QD_M_LAND);
  Co.sst = F(c, QD_F_SST); Co.tb = F(c, QD_F_X7);
  Co.ticket = c->d_ticket + 5 * c->batch;
  Co.band = c->band; if (!c->band_on) Co.band.world = 1;
  BP(c, BL({Co.ub, 1}, {Co.vb, 1}, {Co.eta_in, 0}, {Co.land, 0}, {Co.sst, 2}), BL(Co.eta, Co.tb));
#ifndef QD_HOST_EMU
  const bool pairs = (c->nlon & 1) == 0 && !getenv("QD_NO_PAIRS");       // two cells per thread (qd_ocean.cuh)
  if (pairs) QD_KR(c, k_ocean_continuity2, c->geo, Co, sc);
  else
#endif
  QD_KR(c, k_ocean_continuity, c->geo, Co, sc);
  // latitude bands: no all-reduce kernel here -- the last block of k_ocean_continuity publishes the rank's partial and pulls the world's
  QdOcSstBArgs Sb; memset(&Sb, 0, sizeof(Sb));
  Sb.tb = F(c, QD_F_X7); Sb.ub = ub; Sb.vb = vb; Sb.qnet = F(c, QD_F_QNET);
  Sb.sst = F(c, QD_F_SST); Sb.uo = F(c, QD_F_UO); Sb.vo = F(c, QD_F_VO); Sb.ts_atm = F(c, QD_F_TS); Sb.eta = F(c, QD_F_ETA);
  Sb.land = M(c, QD_M_LAND); Sb.ice = M(c, QD_M_ICE);
  Sb.has_q = cfg->oc_has_q; Sb.has_ice = cfg->oc_has_ice; Sb.inject = inject;
  if (fm) { Sb.fuse_mom = 1; Sb.eta_mid = F(c, QD_F_X8); Sb.taux = F(c, QD_F_X0); Sb.tauy = F(c, QD_F_X1); Sb.ub_next = F(c, QD_F_X2); Sb.vb_next = F(c, QD_F_X3); }
  // T_s is only WRITTEN here (SST injection on a member's last sub-step, also on whatever halo rows the launch covers --
  // the neighbour computes the same values for them); nothing inside the loop reads it, and ocean_core resets its valid
  // width after the loop, so it constrains neither the compute region nor is it marked valid
  BP(c, BL({Sb.tb, 2}, {Sb.ub, 1}, {Sb.vb, 1}, {Sb.qnet, 0}, {Sb.land, 0}, {Sb.ice, 0}, {Sb.sst, 0}, {Sb.eta, 0}),
     BL(Sb.sst, Sb.uo, Sb.vo, Sb.eta));
#ifndef QD_HOST_EMU
  if (pairs) {                                                 // two cells per thread, lazy nan_to_num (qd_ocean.cuh)
    QD_KG(c, k_ocean_sst_finish2, dim3((c->geo.ncomp / 2 + QD_THREADS - 1) / QD_THREADS, c->batch), dim3(QD_THREADS), c->geo, Sb, sc);
    return QD_OK;
  }
#endif
  QD_K(c, k_ocean_sst_finish, c->geo, Sb, sc);
  return QD_OK;
}

// Latitude bands: a sub-step consumes 8 halo rows (eta gradient 1 + del^4 4 + divergence 1 + SST Laplacian 2), so ONE
// exchange of H rows serves H / 8 consecutive sub-steps (the first one computes on its own rows widened by H - 8 rows,
// redundantly on both sides of a cut): half as many neighbour round trips per ocean step at H = 16.  The body of the
// WHILE node is then a GROUP of sub-steps; members whose n_sub is not a multiple of the group size find qd_sub_done()
// true in the trailing sub-steps of their last group (the kernels exit at once).  The valid-halo tracker (band_prep)
// keeps this correct whatever the group size: a kernel whose inputs are no longer valid triggers an exchange itself.
static int ocean_group_size(qd_ctx* c, bool do_shap) {
  if (!c->band_on || do_shap) return 1;
  // Measured on 2 B200s (profiles/README.md): 6.0 instead of 7.85 exchanges per step, but the step is 1 % SLOWER (inside
  // the step graph an exchange costs less than the extra halo rows, the ice-mask kernel and the no-op tail of a group),
  // so the default group is ONE sub-step; QD_OCEAN_GROUP=2..4 opts in.
  if (const char* ov = getenv("QD_OCEAN_GROUP")) return std::max(1, std::min(std::min(4, c->band.H / 8), atoi(ov)));
  return 1;
}
#ifdef QD_HOST_EMU
typedef unsigned long long qd_cond_handle_t;
#else
typedef cudaGraphConditionalHandle qd_cond_handle_t;
#endif
// one group: K x (body, advance).  raw: inside a capture of the WHILE body (launch not counted / profiled, handle live)
static int ocean_substep_group(qd_ctx* c, const qd_step_cfg_t* cfg, int inject, bool do_hyper, bool do_shap, bool raw, qd_cond_handle_t handle) {
  const int K = ocean_group_size(c, do_shap);
  for (int q = 0; q < K; ++q) {
    int rc = ocean_substep_body(c, cfg, inject, do_hyper, do_shap, q == 0);
    if (rc) return rc;
    if (raw) QD_LAUNCH(k_ocean_sub_advance, dim3(1), dim3(32), c->stream, c->geo, c->d_sub_ctr, handle, 1);
    else QD_KG(c, k_ocean_sub_advance, dim3(1), dim3(32), c->geo, c->d_sub_ctr, handle, 0);
  }
  return QD_OK;
}

#ifndef QD_HOST_EMU
// WHILE-node graph around the sub-step body, keyed by the switches baked into the captured launches.
static cudaGraphExec_t ocean_while_graph(qd_ctx* c, const qd_step_cfg_t* cfg, int inject, bool do_hyper, bool do_shap) {
  const unsigned long long key = ((unsigned long long)inject) | ((unsigned long long)do_hyper << 1) | ((unsigned long long)do_shap << 2) |
                                 ((unsigned long long)(cfg->oc_has_q & 1) << 3) | ((unsigned long long)(cfg->oc_has_ice & 1) << 4) |
                                 ((unsigned long long)(cfg->oc_k4_nsub & 0xff) << 8) | ((unsigned long long)(cfg->oc_shapiro_n & 0xff) << 16);
  auto it = c->ocean_graphs.find(key);
  if (it != c->ocean_graphs.end()) return it->second;
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  cudaStream_t saved = c->stream;
  const long long saved_launches = c->launches;
  bool ok = false;
  do {
    if (!c->cap_stream && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) break;
    if (cudaGraphCreate(&graph, 0) != cudaSuccess) break;
    cudaGraphConditionalHandle handle;
    if (cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) != cudaSuccess) break;
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    if (cudaGraphAddNode(&node, graph, nullptr, 0, &np) != cudaSuccess) break;
    cudaGraph_t body = np.conditional.phGraph_out[0];
    if (cudaStreamBeginCaptureToGraph(c->cap_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) break;
    c->stream = c->cap_stream;
    int rc = ocean_substep_group(c, cfg, inject, do_hyper, do_shap, true, handle);
    c->stream = saved;
    cudaGraph_t dummy = nullptr;
    if (cudaStreamEndCapture(c->cap_stream, &dummy) != cudaSuccess || rc != QD_OK) break;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { exec = nullptr; break; }
    ok = true;
  } while (0);
  c->stream = saved;
  c->launches = saved_launches;
  cudaGetLastError();
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    exec = nullptr; c->graph_failures++;
    snprintf(c->err, sizeof(c->err), "CUDA-graph capture of the ocean sub-step loop failed: running it as a host loop (stream mode)");
  }
  c->ocean_graphs[key] = exec;       // nullptr = graphs unavailable -> host loop with one read-back (reported by qd_graph_status)
  return exec;
}
#endif

static int ocean_core(qd_ctx* c, const qd_step_cfg_t* cfg, int inject, const double* wind_u = nullptr, const double* wind_v = nullptr) {
  const double dt = cfg->dt;
  c->oc_counter += 1;                        // ocean.py:281
  QdOcPrepArgs P0; memset(&P0, 0, sizeof(P0));
  // winds: the atmosphere's own fields, or the arrays the caller of the stand-alone step passed (ocean.step(dt, u_atm, v_atm))
  P0.u = wind_u ? wind_u : F(c, QD_F_U); P0.v = wind_v ? wind_v : F(c, QD_F_V); P0.uo = F(c, QD_F_UO); P0.vo = F(c, QD_F_VO);
  P0.taux = F(c, QD_F_X0); P0.tauy = F(c, QD_F_X1); P0.part_u = c->d_part[0]; P0.part_va = c->d_part[1];
  P0.ticket = c->d_ticket + 4 * c->batch;
  P0.dt = dt; P0.sub_ctr = c->band_on ? nullptr : c->d_sub_ctr;      // bands: n_sub needs the all-reduced maxima first
  const bool prep_done = c->tail_did_prep && !wind_u && !wind_v && !c->band_on;      // k_tail of this step already did it
  c->tail_did_prep = 0;
  if (!prep_done) {
    BP(c, BL({P0.u, 0}, {P0.v, 0}, {P0.uo, 0}, {P0.vo, 0}), BL(P0.taux, P0.tauy));
    QD_KR(c, k_ocean_prep, c->geo, P0);
  }
  { int rcb = band_allreduce(c, {QD_S_MAX_UOCEAN, QD_S_MAX_VA}, true); if (rcb) return rcb; }
  if (c->band_on) QD_KG(c, k_ocean_nsub, dim3((c->batch + 63) / 64), dim3(64), c->geo, dt, c->d_sub_ctr);
  const bool do_hyper = (cfg->oc_diff_every > 0) && (c->oc_counter % cfg->oc_diff_every == 0);
  const bool do_shap = (cfg->oc_shapiro_n > 0) && (cfg->oc_shapiro_every > 0) && (c->oc_counter % cfg->oc_shapiro_every == 0);
  if (c->band_on && inject && ocean_group_size(c, do_shap) > 1) {
    // groups of sub-steps per exchange: the closing kernel of a group's first sub-steps also runs on halo rows, so
    // everything it reads there must be valid -- the wind stress, Q_net (constant over the sub-steps: ONE
    // exchange per step) and the ice mask, rebuilt on the halo rows from the exchanged ice thickness exactly as
    // k_tail built it on the rank's own rows (run_simulation.py:2201: ice = h_ice > 0)
    { int rcb = band_hint(c, {QD_F_X0, QD_F_X1, QD_F_QNET, QD_F_HICE}); if (rcb) return rcb; }
    BP(c, BL({F(c, QD_F_HICE), 0}), BL(M(c, QD_M_ICE)));
    QD_K(c, k_ice_mask, c->geo, F(c, QD_F_HICE), M(c, QD_M_ICE));
  } else {
    int rcb = band_hint(c, {QD_F_X0, QD_F_X1}); if (rcb) return rcb;      // wind stress: constant over the sub-steps, exchanged once
  }
  QD_CHECK_LAUNCH(c);
  int rc;
  bool launched = false;
#ifndef QD_HOST_EMU
  if (ocean_fused_rows(c, cfg, do_hyper, do_shap) > 0) {      // k4 rows of this step's sub_dt for the fused sub-steps
    const double* Pp = c->h_prm;
    const int ovu = Pp[QD_P_OC_K4_U] == Pp[QD_P_OC_K4_U], ovv = Pp[QD_P_OC_K4_V] == Pp[QD_P_OC_K4_V], ove = Pp[QD_P_OC_K4_ETA] == Pp[QD_P_OC_K4_ETA];
    QdOcK4Args K; memset(&K, 0, sizeof(K));
    K.k4rows[0] = ovu ? QD_USER_ROW(c, 2) : ROW(c, QD_R_OC_S4DX4); K.raw_k4[0] = ovu;
    K.k4rows[1] = ovv ? QD_USER_ROW(c, 3) : ROW(c, QD_R_OC_S4DX4); K.raw_k4[1] = ovv;
    K.k4rows[2] = ove ? QD_USER_ROW(c, 4) : ROW(c, QD_R_OC_S4DX4); K.raw_k4[2] = ove;
    for (int k = 0; k < 3; ++k) { K.k4_bstride[k] = c->geo.row_bstride; K.scale[k] = 1.0; }
    if (!ove) K.scale[2] = 0.5;                                 // ocean.py:352
    K.tab = c->d_oc_k4;
    QD_KG(c, k_ocean_k4tab, dim3((c->nlat + 127) / 128, 3, c->batch), dim3(128), c->geo, K);
  }
#endif
#ifndef QD_HOST_EMU
  c->ocean_fm_on = ocean_fm(c, cfg, do_hyper, do_shap);
#endif
  if (ocean_fm(c, cfg, do_hyper, do_shap)) { int rcm = ocean_momentum_launch(c); if (rcm) return rcm; }   // momentum of sub-step 0; the body's closing pass does the others
#ifndef QD_HOST_EMU
  if (c->capture_graph) {
    // whole-step capture: splice a WHILE node into the graph being captured (CUDA programming guide,
    // "conditional nodes with stream capture"): current capture dependencies -> node -> rest of the step
    cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t gg; const cudaGraphNode_t* deps = nullptr; size_t ndeps = 0;
    QD_CUDA(c, cudaStreamGetCaptureInfo_v2(c->stream, &st, &id, &gg, &deps, &ndeps));
    cudaGraphConditionalHandle handle;
    QD_CUDA(c, cudaGraphConditionalHandleCreate(&handle, c->capture_graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.conditional.handle = handle; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
    cudaGraphNode_t node;
    QD_CUDA(c, cudaGraphAddNode(&node, c->capture_graph, deps, ndeps, &np));
    QD_CUDA(c, cudaStreamUpdateCaptureDependencies(c->stream, &node, 1, cudaStreamSetCaptureDependencies));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    if (!c->cap_stream2) QD_CUDA(c, cudaStreamCreateWithFlags(&c->cap_stream2, cudaStreamNonBlocking));
    QD_CUDA(c, cudaStreamBeginCaptureToGraph(c->cap_stream2, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    cudaStream_t outer = c->stream;
    c->stream = c->cap_stream2;
    rc = ocean_substep_group(c, cfg, inject, do_hyper, do_shap, true, handle);
    c->stream = outer;
    cudaGraph_t dummy = nullptr;
    cudaError_t ee = cudaStreamEndCapture(c->cap_stream2, &dummy);
    if (rc) return rc;
    if (ee != cudaSuccess) return qd_fail(c, QD_E_CUDA, "capture of the ocean sub-step body", ee);
    c->last_nsub_max = -1;
    launched = true;
  } else if (c->use_graphs && !qd_prof_on(c)) {
    cudaGraphExec_t exec = ocean_while_graph(c, cfg, inject, do_hyper, do_shap);
    if (exec) {
      QD_CUDA(c, cudaGraphLaunch(exec, c->stream));
      c->launches += c->ocean_body_launches(do_hyper, do_shap, cfg);      // per executed sub-step; n_sub is known only on the device
      c->last_nsub_max = -1;
      launched = true;
    }
  }
#endif
  if (!launched) {
    // stream mode: the data-dependent sub-step count costs one small read-back per step
    std::vector<double> s((size_t)c->batch * QD_S_COUNT);
    rc = qd_get_scalars(c, s.data()); if (rc) return rc;
    int nmax = 1;
    for (int b = 0; b < c->batch; ++b) nmax = std::max(nmax, (int)s[(size_t)b * QD_S_COUNT + QD_S_NSUB]);
    c->last_nsub_max = nmax;
    for (int sub = 0; sub < nmax; sub += ocean_group_size(c, do_shap)) {
      rc = ocean_substep_group(c, cfg, inject, do_hyper, do_shap, false, 0); if (rc) return rc;
    }
  }
  QdOcPolarArgs Po; memset(&Po, 0, sizeof(Po));
  Po.sst = F(c, QD_F_SST); Po.uo = F(c, QD_F_UO); Po.vo = F(c, QD_F_VO); Po.ts_atm = F(c, QD_F_TS);
  Po.land = M(c, QD_M_LAND); Po.ice = M(c, QD_M_ICE); Po.has_ice = cfg->oc_has_ice; Po.inject = inject;
  Po.step_idx = c->polar_advances_step ? c->d_step_idx : nullptr;
  QD_KG(c, k_ocean_polar, dim3(2, c->batch), dim3(QD_POLAR_THREADS), c->geo, Po);
  if (c->band_on) for (int id : {(int)QD_F_SST, (int)QD_F_UO, (int)QD_F_VO, (int)QD_F_TS}) c->band_valid[id] = 0;   // pole rows changed on their owners only
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}

extern "C" int qd_ocean_step(qd_ctx* c, const qd_step_cfg_t* cfg) {
  if (!c || !cfg) return QD_E_INVALID;
  QD_BOUND(c);
  return ocean_core(c, cfg, 0);
}
// Same step driven by caller-owned wind arrays [B][nlat][nlon] (device) instead of QD_F_U / QD_F_V: the reference's
// ocean.step(dt, u_atm, v_atm, ...) takes the winds as arguments, and a caller may pass something else than the
// atmosphere's current winds without disturbing them.
extern "C" int qd_ocean_step_winds(qd_ctx* c, const qd_step_cfg_t* cfg, const double* u_dev, const double* v_dev) {
  if (!c || !cfg || !u_dev || !v_dev) return QD_E_INVALID;
  QD_BOUND(c);
  return ocean_core(c, cfg, 0, u_dev, v_dev);
}
// live = captured step / ocean-loop graphs in the cache; failed = captures that fell back to stream mode since qd_create
extern "C" int qd_graph_status(qd_ctx* c, int* live, int* failed) {
  if (!c) return QD_E_INVALID;
  int n = 0;
#ifndef QD_HOST_EMU
  for (auto& kv : c->step_graphs) if (kv.second.first) ++n;
  for (auto& kv : c->ocean_graphs) if (kv.second) ++n;
#endif
  if (live) *live = n;
  if (failed) *failed = c->graph_failures;
  return QD_OK;
}
extern "C" int qd_use_graphs(qd_ctx* c, int level) { if (!c) return QD_E_INVALID; c->use_graphs = level < 0 ? 0 : (level > 2 ? 2 : level); return QD_OK; }
extern "C" int qd_last_nsub(qd_ctx* c, int* out) {
  if (!c || !out) return QD_E_INVALID;
  std::vector<double> s((size_t)c->batch * QD_S_COUNT);
  int rc = qd_get_scalars(c, s.data()); if (rc) return rc;
  for (int b = 0; b < c->batch; ++b) out[b] = (int)s[(size_t)b * QD_S_COUNT + QD_S_NSUB];
  return QD_OK;
}

// ------------------------------------------------------------------------------ script loop step
static int loop_physics(qd_ctx* c, const qd_step_cfg_t* cfg) {
  const double dt = cfg->dt;
  const double* P = c->h_prm;
  const bool orog = P[QD_P_OROG] != 0.0 && P[QD_P_HAS_ELEVATION] != 0.0;
  if ((c->w_set & 3) != 3) return qd_fail(c, QD_E_STATE, "qd_set_gauss(0|1) must be called before qd_loop_step", cudaSuccess);
  const QdGaussW w1 = c->w_sigma1;
  int rc;
  // latitude bands: the halos the whole diagnosis phase reads (winds: divergence, vorticity, advection; T_s: cloud source
  // gradient; cloud: blend + tracer advection) in one exchange instead of three
  if ((rc = band_hint(c, {QD_F_U, QD_F_V, QD_F_PCOND, QD_F_TS, QD_F_CLOUD}))) return rc;
  // precipitation (physics.py:253-354)
  QdPrecipAArgs Pa; memset(&Pa, 0, sizeof(Pa));
  Pa.u = F(c, QD_F_U); Pa.v = F(c, QD_F_V); Pa.pcond = F(c, QD_F_PCOND); Pa.nx = F(c, QD_F_OROG_NX); Pa.ny = F(c, QD_F_OROG_NY);
  Pa.pos = F(c, QD_F_X0); Pa.orog_raw = F(c, QD_F_X1); Pa.part = c->d_part[0]; Pa.ticket = c->d_ticket + 6 * c->batch;
  Pa.forcing = c->d_forcing; Pa.step_idx = c->d_step_idx; Pa.hcos = c->d_hcos;      // this step's hour-angle cosines ride along
  BP(c, BL({Pa.u, 1}, {Pa.v, 1}, {Pa.pcond, 0}, {Pa.nx, 0}, {Pa.ny, 0}), BL(Pa.pos, Pa.orog_raw));
  QD_KR(c, k_precip_a, c->geo, Pa);          // latitude bands: sum(Pq w) is all-reduced together with sum(P_raw w) below (both are first read by the precipitation Gaussian)
  double* orog_f = F(c, QD_F_X1);
  if (orog) {
    double* fl[1] = {F(c, QD_F_X1)}; double* sx[1] = {F(c, QD_F_X5)}; bool moved = false;
    if ((rc = op_gauss(c, 1, fl, sx, w1, &moved))) return rc;
    if (moved) orog_f = sx[0];
  }
  if ((rc = op_median(c, F(c, QD_F_X0), 0.0, c->d_scal + QD_S_MED_POS, c->d_scal + QD_S_CNT_POS, QD_S_COUNT, 0))) return rc;
  QdPrecipBArgs Pb; memset(&Pb, 0, sizeof(Pb));
  Pb.pos = F(c, QD_F_X0); Pb.pcond = F(c, QD_F_PCOND); Pb.orog = orog_f; Pb.praw = F(c, QD_F_X2);
  Pb.part = c->d_part[0]; Pb.ticket = c->d_ticket + 6 * c->batch;
  BP(c, BL({Pb.pos, 0}, {Pb.pcond, 0}, {Pb.orog, 0}), BL(Pb.praw));
  QD_KR(c, k_precip_b, c->geo, Pb);
  if ((rc = band_allreduce(c, {QD_S_SUM_PQW, QD_S_SUM_PRAWW}, false))) return rc;
  bool fused = false;
#ifndef QD_HOST_EMU
  fused = gauss2d_ok(c, w1);
  if (fused) {       // Gaussian of P_raw * s (and of k_precip * pos in the fallback) + blend + clip in one tile kernel
    QdG2Args A; memset(&A, 0, sizeof(A));
    A.n = 2; A.src[0] = F(c, QD_F_X2); A.src[1] = F(c, QD_F_X0); A.dst[0] = F(c, QD_F_PRECIP);
    if ((rc = launch_gauss2d<QD_G2_PRECIP>(c, A, w1, {{A.src[0], w1.r}, {A.src[1], w1.r}}, {A.dst[0]}, "k_gauss2d_tile<precip>"))) return rc;
  }
#endif
  if (!fused) {
    QdPrecipCArgs Pc; Pc.praw = F(c, QD_F_X2); Pc.pos = F(c, QD_F_X0); Pc.g0 = F(c, QD_F_X3); Pc.g1 = F(c, QD_F_X4);
    BP(c, BL({Pc.praw, w1.r}, {Pc.pos, w1.r}, {Pc.g1, 0}), BL(Pc.g0, Pc.g1));
    QD_K(c, k_precip_c, c->geo, Pc, w1);
    QdPrecipDArgs Pd; Pd.g0 = F(c, QD_F_X3); Pd.g1 = F(c, QD_F_X4); Pd.precip = F(c, QD_F_PRECIP);
    BP(c, BL({Pd.g0, 0}, {Pd.g1, 0}), BL(Pd.precip));
    QD_K(c, k_precip_d, c->geo, Pd, w1);
  }
  // clouds (run_simulation.py:1866-1934)
  if ((rc = op_median(c, F(c, QD_F_PRECIP), 1e-6, c->d_scal + QD_S_PREF, c->d_scal + QD_S_CNT_PRECIP, QD_S_COUNT, 1))) return rc;
  QdCloudAArgs Ca; Ca.precip = F(c, QD_F_PRECIP); Ca.ts = F(c, QD_F_TS); Ca.u = F(c, QD_F_U); Ca.v = F(c, QD_F_V);
  Ca.craw = F(c, QD_F_X0); Ca.sraw = F(c, QD_F_X1);
  BP(c, BL({Ca.precip, 0}, {Ca.ts, 1}, {Ca.u, 1}, {Ca.v, 1}), BL(Ca.craw, Ca.sraw));
  QD_K(c, k_cloud_a, c->geo, Ca);
#ifndef QD_HOST_EMU
  if (fused) {
    QdG2Args A; memset(&A, 0, sizeof(A));
    A.n = 2; A.src[0] = F(c, QD_F_X0); A.src[1] = F(c, QD_F_X1); A.dst[0] = F(c, QD_F_CLOUD); A.dt = dt / (6 * 3600);
    if ((rc = launch_gauss2d<QD_G2_CLOUD_B>(c, A, w1, {{A.src[0], w1.r}, {A.src[1], w1.r}, {A.dst[0], 0}}, {A.dst[0]}, "k_gauss2d_tile<cloud_b>"))) return rc;
  }
#endif
  if (!fused) {
    QdFields f = mk_fields(2); f.src[0] = F(c, QD_F_X0); f.src[1] = F(c, QD_F_X1); f.dst[0] = F(c, QD_F_X2); f.dst[1] = F(c, QD_F_X3);
    BP(c, BL({f.src[0], w1.r}, {f.src[1], w1.r}), BL(f.dst[0], f.dst[1]));
    QD_K(c, k_gauss_lat, c->geo, f, w1);
    QdCloudBArgs Cb; Cb.g0 = F(c, QD_F_X2); Cb.g1 = F(c, QD_F_X3); Cb.cloud = F(c, QD_F_CLOUD); Cb.dt = dt / (6 * 3600);
    BP(c, BL({Cb.g0, 0}, {Cb.g1, 0}, {Cb.cloud, 0}), BL(Cb.cloud));
    QD_K(c, k_cloud_b, c->geo, Cb, w1);
  }
  if (P[QD_P_CLOUD_ADVECT] != 0.0) {
    QdFields f = mk_fields(1); f.src[0] = F(c, QD_F_CLOUD); f.dst[0] = F(c, QD_F_X0);
    BP(c, BL({f.src[0], band_radv(c, dt)}, {F(c, QD_F_U), 0}, {F(c, QD_F_V), 0}), BL(f.dst[0]));
    QD_K(c, k_advect, c->geo, f, F(c, QD_F_U), F(c, QD_F_V), dt, ROW(c, QD_R_COS_ADV_HALF), ROW(c, QD_R_INV_ACOS_HALF));
    const double sig = P[QD_P_CLOUD_SMOOTH_SIGMA];
    QdGaussW wc = c->w_cloud;
    if (!(sig > 0.0)) wc.r = 0;
    const double* src = F(c, QD_F_X0);
    bool fused_c = false;
#ifndef QD_HOST_EMU
    fused_c = gauss2d_ok(c, wc);
    if (fused_c) {
      QdG2Args A; memset(&A, 0, sizeof(A));
      A.n = 1; A.src[0] = F(c, QD_F_X0); A.dst[0] = F(c, QD_F_CLOUD);
      if ((rc = launch_gauss2d<QD_G2_CLOUD_C>(c, A, wc, {{A.src[0], wc.r}, {A.dst[0], 0}}, {A.dst[0]}, "k_gauss2d_tile<cloud_c>"))) return rc;
    }
#endif
    if (!fused_c) {
      if (wc.r > 0) {
        QdFields g1 = mk_fields(1); g1.src[0] = F(c, QD_F_X0); g1.dst[0] = F(c, QD_F_X1);
        BP(c, BL({g1.src[0], wc.r}), BL(g1.dst[0]));
        QD_K(c, k_gauss_lat, c->geo, g1, wc);
        src = F(c, QD_F_X1);
      }
      QdCloudCArgs Cc; Cc.g0 = src; Cc.cloud = F(c, QD_F_CLOUD);
      BP(c, BL({Cc.g0, 0}, {Cc.cloud, 0}), BL(Cc.cloud));
      QD_K(c, k_cloud_c, c->geo, Cc, wc);
    }
  }
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}

static int loop_step_enqueue(qd_ctx* c, const qd_step_cfg_t* cfg) {
  int rc;
  band_invalidate_dynamic(c);               // every step starts from "own rows only": one graph serves every step
  if (c->band_on && cfg->with_routing && c->route.ready)
    return qd_fail(c, QD_E_STATE, "latitude bands: river routing is a global DAG and is not partitioned (SURVEY 8e)", cudaSuccess);
  if ((rc = loop_physics(c, cfg))) return rc;
  if ((rc = atmos_core(c, cfg, 1))) return rc;
  // the ocean's closing kernel advances the forcing-table index when it is the last field kernel of the step
  c->polar_advances_step = (cfg->with_ocean && !(cfg->with_routing && c->route.ready)) ? 1 : 0;
  if (cfg->with_ocean) { rc = ocean_core(c, cfg, 1); c->polar_advances_step = 0; if (rc) return rc; }
  if (cfg->with_routing && c->route.ready) {
    QD_K(c, k_route_accumulate, c->geo, F(c, QD_F_RLAND), c->route.d_land, c->route.d_buffer, cfg->dt);
  }
  if (!(cfg->with_ocean && !(cfg->with_routing && c->route.ready))) QD_KG(c, k_step_advance, dim3(1), dim3(32), c->d_step_idx);
  if (c->band_on) band_set_ext(c, 0);
  return QD_OK;
}

#ifndef QD_HOST_EMU
// One whole loop step as ONE CUDA graph (kernels, memsets, the cooperative selects and the ocean WHILE
// node).  Variants differ only in host-known cadences and switches, which form the cache key.
static int loop_step_graph(qd_ctx* c, const qd_step_cfg_t* cfg) {
  const int an = c->atm_counter + 1, on = c->oc_counter + 1;
  const bool a_hyp = cfg->diff_enable && (an % std::max(1, cfg->diff_every) == 0);
  const bool a_shp = cfg->shapiro_every > 0 && (an % cfg->shapiro_every == 0);
  const bool a_spc = cfg->spec_every > 0 && (an % cfg->spec_every == 0);
  const bool o_hyp = (cfg->oc_diff_every > 0) && (on % cfg->oc_diff_every == 0);
  const bool o_shp = (cfg->oc_shapiro_n > 0) && (cfg->oc_shapiro_every > 0) && (on % cfg->oc_shapiro_every == 0);
  unsigned long long key = 1469598103934665603ull;
  auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
  const int en = c->eco_steps + 1;
  const bool eco_new = cfg->with_eco && (en % std::max(1, c->eco_every_nphys) == 0);
  mix(a_hyp); mix(a_shp); mix(a_spc); mix(o_hyp); mix(o_shp); mix(c->has_cloud_eff); mix(c->route.ready);
  mix(cfg->with_eco ? (eco_new ? 1 : (c->eco_have_alpha ? 2 : 3)) : 0); mix(c->d_lai != nullptr); mix((unsigned long long)c->eco_nl);
  { const unsigned char* p = (const unsigned char*)cfg; for (size_t k = 0; k < sizeof(*cfg); ++k) mix(p[k]); }
  auto it = c->step_graphs.find(key);
  if (it != c->step_graphs.end()) {
    if (!it->second.first) return QD_E_STATE;
    QD_CUDA(c, cudaGraphLaunch(it->second.first, c->stream));
    c->atm_counter = an;
    if (cfg->with_ocean) c->oc_counter = on;
    if (cfg->loop_with_albedo) c->has_cloud_eff = 1;
    if (cfg->with_eco) { c->eco_steps = en; if (eco_new) c->eco_have_alpha = 1; }
    c->launches += it->second.second;
    return QD_OK;
  }
  // build: capture this step once, then launch it
  const int sa = c->atm_counter, so = c->oc_counter, sce = c->has_cloud_eff, ses = c->eco_steps, sea = c->eco_have_alpha;
  const long long sl = c->launches;
  cudaStream_t saved = c->stream;
  cudaGraph_t G = nullptr; cudaGraphExec_t exec = nullptr;
  bool ok = false;
  int rc = QD_OK;
  do {
    if (!c->cap_stream && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) break;
    if (cudaGraphCreate(&G, 0) != cudaSuccess) break;
    if (cudaStreamBeginCaptureToGraph(c->cap_stream, G, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) break;
    c->stream = c->cap_stream; c->capture_graph = G;
    rc = loop_step_enqueue(c, cfg);
    c->stream = saved; c->capture_graph = nullptr;
    cudaGraph_t out = nullptr;
    if (cudaStreamEndCapture(c->cap_stream, &out) != cudaSuccess || rc != QD_OK) break;
    if (cudaGraphInstantiate(&exec, G, 0) != cudaSuccess) { exec = nullptr; break; }
    ok = true;
  } while (0);
  c->stream = saved; c->capture_graph = nullptr;
  cudaGetLastError();
  if (G) cudaGraphDestroy(G);
  const long long per_step = c->launches - sl;
  if (!ok) {
    c->atm_counter = sa; c->oc_counter = so; c->has_cloud_eff = sce; c->launches = sl; c->eco_steps = ses; c->eco_have_alpha = sea;
    c->step_graphs[key] = std::make_pair((cudaGraphExec_t) nullptr, 0ll);
    c->graph_failures++;
    if (rc == QD_OK) snprintf(c->err, sizeof(c->err), "CUDA-graph capture of the loop step failed: this step variant runs in stream mode");
    return rc != QD_OK ? rc : QD_E_STATE;
  }
  c->step_graphs[key] = std::make_pair(exec, per_step);
  QD_CUDA(c, cudaGraphLaunch(exec, c->stream));     // counters were advanced by the capture pass
  return QD_OK;
}
#endif

extern "C" int qd_loop_step(qd_ctx* c, const qd_step_cfg_t* cfg, const qd_forcing_t* forcing, int nsteps) {
  if (!c || !cfg || !forcing || nsteps < 1) return QD_E_INVALID;
  QD_BOUND(c);
  { const int rc = qd_derive(c, cfg->dt); if (rc) return rc; }
  if (nsteps > c->forcing_cap) {
    // the forcing table has a fixed capacity (captured graphs hold its address): longer calls run as chunks; the
    // upload of chunk k+1 is ordered behind the kernels of chunk k on the same stream
    for (int s0 = 0; s0 < nsteps; s0 += c->forcing_cap) {
      const int rc = qd_loop_step(c, cfg, forcing + s0, std::min(c->forcing_cap, nsteps - s0));
      if (rc) return rc;
    }
    return QD_OK;
  }
  if (nsteps == 1) {
    // one step per call (the interactive / end-to-end pattern): the 80 bytes of orbital scalars travel as a kernel
    // argument -- one launch instead of a pageable-memory copy plus a memset on the critical path of every step
    QD_LAUNCH(k_set_forcing, dim3(1), dim3(1), c->stream, c->d_forcing, forcing[0], c->d_step_idx);
    c->launches++;
  } else {
    QD_CUDA(c, cudaMemcpyAsync(c->d_forcing, forcing, (size_t)nsteps * sizeof(qd_forcing_t), cudaMemcpyHostToDevice, c->stream));
    QD_CUDA(c, cudaMemsetAsync(c->d_step_idx, 0, sizeof(int), c->stream));
  }
  for (int s = 0; s < nsteps; ++s) {
    int rc;
#ifndef QD_HOST_EMU
    if (c->use_graphs >= 2 && !qd_prof_on(c)) {
      rc = loop_step_graph(c, cfg);
      if (rc == QD_OK) continue;
      if (rc != QD_E_STATE) return rc;        // QD_E_STATE: graphs unavailable for this variant -> stream mode
    }
#endif
    if ((rc = loop_step_enqueue(c, cfg))) return rc;
  }
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}

#include "qd_route.cuh"
#include "qd_netbuild.h"

// Routing-network builder (host; no context, no device): see qd_netbuild.h.  elev is filled in place.
extern "C" int qd_net_build(int nlat, int nlon, double* elev /* in: elevation, out: pit-filled */, const uint8_t* land_mask,
                            const double* dist /* [3][nlat][3][3] */, int pit_iters, double pit_eps,
                            int64_t* flow_to /* [nlat*nlon] */, int64_t* flow_order /* [nlat*nlon] */, int64_t* n_order,
                            uint8_t* lake_mask, int32_t* lake_id, int32_t* lake_outlet /* [nlat*nlon] capacity */, int* n_lakes, int* sweeps) {
  if (nlat < 2 || nlon < 2 || !elev || !land_mask || !dist || !flow_to || !flow_order || !n_order || !lake_mask || !lake_id || !lake_outlet || !n_lakes)
    return QD_E_INVALID;
  const int it = qd_net_pit_fill(nlat, nlon, elev, land_mask, pit_iters, pit_eps);
  if (sweeps) *sweeps = it;
  qd_net_flow_to(nlat, nlon, elev, land_mask, dist, flow_to);
  *n_lakes = qd_net_lakes(nlat, nlon, flow_to, land_mask, lake_mask, lake_id);
  if (*n_lakes > 0) qd_net_outlets(nlat, nlon, elev, lake_mask, lake_id, land_mask, *n_lakes, lake_outlet);
  *n_order = qd_net_topo_order(nlat, nlon, flow_to, land_mask, flow_order);
  return QD_OK;
}
