// qd_route.cuh -- RiverRouting (pygcm/routing.py:211-335) as an integer-indexed gather.
//
// The reference pushes mass sequentially along flow_order (routing.py:261-298).  The same
// floating-point result is obtained in parallel by (i) giving every cell the list of its EARLY
// donors -- cells that drain into it and come earlier in flow_order -- sorted by their position
// in flow_order, (ii) processing cells level by level (level = 1 + max level of early donors) and
// (iii) letting each receiver add its donors' masses in that order, starting from its own
// buffered mass: exactly the additions of the serial loop, in the same order, so flow
// accumulation is bit-exact.  The ocean inflow is an ordered serial sum over the ocean-draining
// cells in flow_order order (one thread; a few thousand adds every 6 model hours).  Donors that
// come LATER in flow_order leave their mass in the receiver as residual (routing.py:300-301).
// Included at the end of qd_api.cu (single member: routing is replicated).
#pragma once

__global__ void k_route_level(const int* cells, int n, const int* don_off, const int* don,
                              const double* buffer, double* mass) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int r = cells[t];
  double m = buffer[r];
  for (int k = don_off[r]; k < don_off[r + 1]; ++k) {
    const double md = mass[don[k]];
    if (md > 0.0) m = m + md;                       // donors with m <= 0 are skipped (routing.py:263)
  }
  mass[r] = m;
}
// residual left in every cell after the event: own mass if it was not pushed (m <= 0 or the cell is
// not in flow_order) plus the late donors' masses, in flow_order order
__global__ void k_route_after(int ncell, const unsigned char* in_order, const int* late_off, const int* late,
                              const double* buffer, const double* mass, double* after, double* flow_acc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ncell) return;
  const double m = in_order[r] ? mass[r] : buffer[r];
  const bool pushed = in_order[r] && (m > 0.0);
  double a = pushed ? 0.0 : m;
  for (int k = late_off[r]; k < late_off[r + 1]; ++k) {
    const double md = mass[late[k]];
    if (md > 0.0) a = a + md;
  }
  after[r] = a;
  flow_acc[r] = pushed ? m : 0.0;
}
// ordered serial sums: out[0] = ocean inflow, out[1 + lid] = lake stores without outlet
__global__ void k_route_sums(const int* ocean_list, long long n_ocean, const int* lake_list, const int* lake_of,
                             long long n_lake, const double* mass, double* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double s = 0.0;
  for (long long k = 0; k < n_ocean; ++k) { const double m = mass[ocean_list[k]]; if (m > 0.0) s += m; }
  out[0] = s;
  for (long long k = 0; k < n_lake; ++k) { const double m = mass[lake_list[k]]; if (m > 0.0) out[1 + lake_of[k]] += m; }
}

template <class T>
static int up_vec(qd_ctx* c, T** dptr, const std::vector<T>& v) {
  const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  QD_CUDA(c, cudaMalloc((void**)dptr, bytes));
  if (!v.empty()) QD_CUDA(c, cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return QD_OK;
}

extern "C" int qd_route_setup(qd_ctx* c, int n_order, const int64_t* flow_order, const int64_t* flow_to,
                              const uint8_t* land, const uint8_t* lake, const int32_t* lake_id,
                              int n_lakes, const int64_t* lake_outlet) {
  if (!c || n_order < 0 || !flow_order || !flow_to || !land) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  qd_drop_graphs(c);                               // captured steps hold the old network's buffer pointers
  qd_route_free(c->route);
  qd_route& R = c->route;
  const int n = c->ncell;
  const bool has_lakes = lake && lake_id && lake_outlet && n_lakes > 0;
  // static target of every ordered cell: >=0 cell, -1 ocean, -2-lid lake store, -1000000 none
  std::vector<int> pos(n, -1), target(n, -1000000);
  for (int k = 0; k < n_order; ++k) {
    const int64_t idx = flow_order[k];
    if (idx < 0 || idx >= n) return qd_fail(c, QD_E_INVALID, "flow_order index out of range", cudaSuccess);
    pos[idx] = k;
  }
  std::vector<int> ocean_list, lake_list, lake_of;
  for (int k = 0; k < n_order; ++k) {
    const int idx = (int)flow_order[k];
    if (pos[idx] != k) continue;                    // duplicated entries: only the last position counts below
    int tg;
    if (has_lakes && lake[idx] > 0) {
      const int lid = lake_id[idx];
      if (lid > 0 && lid <= n_lakes) {
        const int64_t o = lake_outlet[lid - 1];
        if (o < 0) tg = -1; else if (o < n && land[o] == 1) tg = (int)o; else tg = -1;
      } else if (lid > 0) tg = -2 - (lid - 1);
      else tg = -1000000;
    } else {
      const int64_t dn = flow_to[idx];
      tg = (dn < 0 || dn >= n || land[dn] != 1) ? -1 : (int)dn;
    }
    target[idx] = tg;
    if (tg == -1) ocean_list.push_back(idx);
    else if (tg <= -2 && tg > -1000000) { lake_list.push_back(idx); lake_of.push_back(-2 - tg); }
  }
  // donors (early / late relative to the receiver's own position), each list sorted by position
  std::vector<std::vector<int>> early(n), late(n);
  for (int k = 0; k < n_order; ++k) {
    const int d = (int)flow_order[k];
    if (pos[d] != k) continue;
    const int r = target[d];
    if (r < 0) continue;
    if (pos[r] >= 0 && pos[d] < pos[r]) early[r].push_back(d); else late[r].push_back(d);
  }
  // levels over the early-donor DAG, in flow_order order (donors always precede receivers)
  std::vector<int> level(n, 0);
  int n_levels = 0;
  for (int k = 0; k < n_order; ++k) {
    const int r = (int)flow_order[k];
    if (pos[r] != k) continue;
    int lv = 0;
    for (int d : early[r]) lv = std::max(lv, level[d] + 1);
    level[r] = lv;
    n_levels = std::max(n_levels, lv + 1);
  }
  std::vector<int> level_off(n_levels + 1, 0), level_cells;
  for (int k = 0; k < n_order; ++k) { const int r = (int)flow_order[k]; if (pos[r] == k) level_off[level[r] + 1]++; }
  for (int l = 0; l < n_levels; ++l) level_off[l + 1] += level_off[l];
  level_cells.resize(level_off[n_levels]);
  { std::vector<int> cur(level_off.begin(), level_off.end() - 1);
    for (int k = 0; k < n_order; ++k) { const int r = (int)flow_order[k]; if (pos[r] == k) level_cells[cur[level[r]]++] = r; } }
  std::vector<int> don_off(n + 1, 0), don, late_off(n + 1, 0), latev;
  for (int r = 0; r < n; ++r) { don_off[r + 1] = don_off[r] + (int)early[r].size(); late_off[r + 1] = late_off[r] + (int)late[r].size(); }
  don.reserve(don_off[n]); latev.reserve(late_off[n]);
  for (int r = 0; r < n; ++r) { for (int d : early[r]) don.push_back(d); for (int d : late[r]) latev.push_back(d); }
  std::vector<unsigned char> in_order(n, 0);
  for (int r = 0; r < n; ++r) in_order[r] = pos[r] >= 0;
  int rc;
  if ((rc = up_vec(c, &R.d_level_cells, level_cells))) return rc;
  if ((rc = up_vec(c, &R.d_don_off, don_off))) return rc;
  if ((rc = up_vec(c, &R.d_don, don))) return rc;
  if ((rc = up_vec(c, &R.d_late_off, late_off))) return rc;
  if ((rc = up_vec(c, &R.d_late, latev))) return rc;
  if ((rc = up_vec(c, &R.d_ocean_list, ocean_list))) return rc;
  if ((rc = up_vec(c, &R.d_lake_list, lake_list))) return rc;
  if ((rc = up_vec(c, &R.d_lake_of, lake_of))) return rc;
  if ((rc = up_vec(c, &R.d_in_order, in_order))) return rc;
  { std::vector<unsigned char> lv(land, land + n); if ((rc = up_vec(c, &R.d_land, lv))) return rc; }
  const size_t fb = (size_t)c->batch * n * 8;
  QD_CUDA(c, cudaMalloc((void**)&R.d_buffer, fb)); QD_CUDA(c, cudaMemset(R.d_buffer, 0, fb));
  QD_CUDA(c, cudaMalloc((void**)&R.d_mass, (size_t)n * 8));
  QD_CUDA(c, cudaMalloc((void**)&R.d_after, (size_t)n * 8));
  QD_CUDA(c, cudaMalloc((void**)&R.d_out, (size_t)(n + 1 + std::max(n_lakes, 1)) * 8));
  R.level_off = level_off; R.n_levels = n_levels; R.n_lakes = n_lakes; R.n_order = n_order;
  R.n_ocean = (long long)ocean_list.size(); R.n_lake_store = (long long)lake_list.size();
  R.ready = 1;
  return QD_OK;
}

extern "C" int qd_route_accumulate(qd_ctx* c, double dt) {
  if (!c) return QD_E_INVALID;
  QD_BOUND(c);
  if (!c->route.ready) return qd_fail(c, QD_E_STATE, "qd_route_setup was not called", cudaSuccess);
  QD_K(c, k_route_accumulate, c->geo, F(c, QD_F_RLAND), c->route.d_land, c->route.d_buffer, dt);
  QD_CHECK_LAUNCH(c);
  return QD_OK;
}
extern "C" int qd_route_levels(qd_ctx* c) { return (c && c->route.ready) ? c->route.n_levels : -1; }

// One routing event for ensemble member `member`: routes the accumulated buffer, clears it, and
// returns flow accumulation (kg per cell), ocean inflow, the per-cell residual and the input buffer
// (the host forms the NumPy sums of routing.py:252,301 from these so they are bit-identical).
extern "C" int qd_route_event(qd_ctx* c, int member, double* flow_accum_host, double* ocean_inflow_kg,
                              double* after_host, double* input_host, double* lake_store_host) {
  if (!c || member < 0 || member >= c->batch) return QD_E_INVALID;
  qd_route& R = c->route;
  if (!R.ready) return qd_fail(c, QD_E_STATE, "qd_route_setup was not called", cudaSuccess);
  const int n = c->ncell;
  double* buf = R.d_buffer + (size_t)member * n;
  double* flow = R.d_out + 1 + std::max(R.n_lakes, 1);
  QD_CUDA(c, cudaMemsetAsync(R.d_out, 0, (size_t)(1 + std::max(R.n_lakes, 1)) * 8, c->stream));
  QD_CUDA(c, cudaMemsetAsync(R.d_mass, 0, (size_t)n * 8, c->stream));
  for (int l = 0; l < R.n_levels; ++l) {
    const int cnt = R.level_off[l + 1] - R.level_off[l];
    if (cnt <= 0) continue;
    QD_LAUNCH(k_route_level, dim3((cnt + 127) / 128), dim3(128), c->stream, R.d_level_cells + R.level_off[l], cnt,
              R.d_don_off, R.d_don, buf, R.d_mass);
    c->launches++;
  }
  QD_LAUNCH(k_route_after, dim3((n + 127) / 128), dim3(128), c->stream, n, R.d_in_order, R.d_late_off, R.d_late, buf, R.d_mass, R.d_after, flow);
  QD_LAUNCH(k_route_sums, dim3(1), dim3(32), c->stream, R.d_ocean_list, R.n_ocean, R.d_lake_list, R.d_lake_of, R.n_lake_store, R.d_mass, R.d_out);
  c->launches += 2;
  QD_CHECK_LAUNCH(c);
  if (input_host) QD_CUDA(c, cudaMemcpyAsync(input_host, buf, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (flow_accum_host) QD_CUDA(c, cudaMemcpyAsync(flow_accum_host, flow, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (after_host) QD_CUDA(c, cudaMemcpyAsync(after_host, R.d_after, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (ocean_inflow_kg) QD_CUDA(c, cudaMemcpyAsync(ocean_inflow_kg, R.d_out, 8, cudaMemcpyDeviceToHost, c->stream));
  if (lake_store_host && R.n_lakes > 0) QD_CUDA(c, cudaMemcpyAsync(lake_store_host, R.d_out + 1, (size_t)R.n_lakes * 8, cudaMemcpyDeviceToHost, c->stream));
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  QD_CUDA(c, cudaMemsetAsync(buf, 0, (size_t)n * 8, c->stream));
  return QD_OK;
}
extern "C" int qd_route_buffer(qd_ctx* c, int member, double* host, int upload) {
  if (!c || !host || member < 0 || member >= c->batch || !c->route.ready) return QD_E_INVALID;
  QD_CUDA(c, cudaStreamSynchronize(c->stream));
  double* buf = c->route.d_buffer + (size_t)member * c->ncell;
  if (upload) QD_CUDA(c, cudaMemcpy(buf, host, (size_t)c->ncell * 8, cudaMemcpyHostToDevice));
  else QD_CUDA(c, cudaMemcpy(host, buf, (size_t)c->ncell * 8, cudaMemcpyDeviceToHost));
  return QD_OK;
}
