// qd_netbuild.h -- host C++ port of the D8 routing-network builder of the reference
// (scripts/generate_hydrology_maps.py:48-273).  Order-dependent sequential algorithms (Gauss-Seidel
// pit filling, Kahn topological sort with a FIFO seeded in ascending index order, scan-order lake labelling) are
// kept sequential so that every output is bit-identical; the Python loops they replace take seconds at 61x120 and
// minutes to hours at the benchmark grids.  Distances between neighbouring cell centres depend only on the row and
// the offset, so the caller passes them as a table dist[3][nlat][3][3] (first / interior / last column) evaluated with the reference's own NumPy
// expression (spherical_distance :65-82) -- no libm difference can leak into the slope comparisons.
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include <deque>
#include <limits>

// D8 neighbours in the reference's order: dj in (-1,0,1), di in (-1,0,1), skipping (0,0); longitude periodic,
// latitude clamped (:48-62)
template <class Fn>
static inline void qd_net_neighbors(int i, int j, int nlon, int nlat, Fn fn) {
  for (int dj = -1; dj <= 1; ++dj)
    for (int di = -1; di <= 1; ++di) {
      if (di == 0 && dj == 0) continue;
      const int jj = j + dj;
      if (jj < 0 || jj >= nlat) continue;
      int ii = (i + di) % nlon;
      if (ii < 0) ii += nlon;
      if (!fn(ii, jj, di, dj)) return;
    }
}

// pit_fill (:85-111): returns the number of sweeps
static int qd_net_pit_fill(int nlat, int nlon, double* e, const uint8_t* land, int max_iters, double eps) {
  bool changed = true;
  int it = 0;
  while (changed && it < max_iters) {
    changed = false;
    ++it;
    for (int j = 0; j < nlat; ++j)
      for (int i = 0; i < nlon; ++i) {
        if (land[(size_t)j * nlon + i] != 1) continue;
        double mn = std::numeric_limits<double>::infinity();
        bool any = false;
        qd_net_neighbors(i, j, nlon, nlat, [&](int ii, int jj, int, int) {
          const double v = e[(size_t)jj * nlon + ii];
          if (!any || v < mn) mn = v;          // Python's min(): first minimum, NaN-free inputs
          any = true;
          return true;
        });
        if (!any) continue;
        double& c = e[(size_t)j * nlon + i];
        if (c <= mn) {
          const double nv = mn + eps;
          if (nv > c) { c = nv; changed = true; }
        }
      }
  }
  return it;
}

// compute_flow_to_index (:114-152); dist[cls][j][dj+1][di+1], cls = 0: i = 0, 1: interior, 2: i = nlon-1
static void qd_net_flow_to(int nlat, int nlon, const double* elev, const uint8_t* land, const double* dist, int64_t* flow_to) {
  for (int j = 0; j < nlat; ++j)
    for (int i = 0; i < nlon; ++i) {
      const size_t c = (size_t)j * nlon + i;
      flow_to[c] = -1;
      if (land[c] != 1) continue;
      const double z0 = elev[c];
      double best = -std::numeric_limits<double>::infinity();
      int bi = -1, bj = -1;
      qd_net_neighbors(i, j, nlon, nlat, [&](int ii, int jj, int di, int dj) {
        const int cls = i == 0 ? 0 : (i == nlon - 1 ? 2 : 1);
        const double d = dist[(((size_t)cls * nlat + j) * 3 + (dj + 1)) * 3 + (di + 1)];
        if (d <= 0) return true;
        const double slope = (z0 - elev[(size_t)jj * nlon + ii]) / d;
        if (slope > best) { best = slope; bi = ii; bj = jj; }
        return true;
      });
      if (best > 0 && bi >= 0 && land[(size_t)bj * nlon + bi] == 1) flow_to[c] = (int64_t)bj * nlon + bi;
    }
}

// topo_sort_flow_order (:155-191): returns the number of entries written to order[] (= number of land cells)
static int64_t qd_net_topo_order(int nlat, int nlon, const int64_t* flow_to, const uint8_t* land, int64_t* order) {
  const int64_t n = (int64_t)nlat * nlon;
  std::vector<int64_t> indeg((size_t)n, 0);
  for (int64_t k = 0; k < n; ++k) {
    if (land[k] != 1) continue;
    const int64_t dn = flow_to[k];
    if (dn >= 0 && land[dn] == 1) indeg[(size_t)dn] += 1;
  }
  std::deque<int64_t> q;
  for (int64_t k = 0; k < n; ++k) if (land[k] == 1 && indeg[(size_t)k] == 0) q.push_back(k);
  std::vector<uint8_t> seen((size_t)n, 0);
  int64_t m = 0;
  while (!q.empty()) {
    const int64_t u = q.front(); q.pop_front();
    order[m++] = u; seen[(size_t)u] = 1;
    const int64_t dn = flow_to[u];
    if (dn >= 0 && land[dn] == 1) { if (--indeg[(size_t)dn] == 0) q.push_back(dn); }
  }
  for (int64_t k = 0; k < n; ++k) if (land[k] == 1 && !seen[(size_t)k]) order[m++] = k;   // cycles (should not happen): appended in index order
  return m;
}

// identify_lakes (:194-227): returns n_lakes
static int qd_net_lakes(int nlat, int nlon, const int64_t* flow_to, const uint8_t* land, uint8_t* lake_mask, int32_t* lake_id) {
  const size_t n = (size_t)nlat * nlon;
  memset(lake_mask, 0, n); memset(lake_id, 0, n * sizeof(int32_t));
  std::vector<uint8_t> term(n), visited(n, 0);
  bool any = false;
  for (size_t k = 0; k < n; ++k) { term[k] = (land[k] == 1 && flow_to[k] == -1); any = any || term[k]; }
  if (!any) return 0;
  int count = 0;
  std::vector<std::pair<int, int>> stack;
  for (int j = 0; j < nlat; ++j)
    for (int i = 0; i < nlon; ++i) {
      const size_t c = (size_t)j * nlon + i;
      if (!term[c] || visited[c]) continue;
      ++count;
      stack.clear(); stack.push_back({i, j}); visited[c] = 1;
      while (!stack.empty()) {
        const auto cur = stack.back(); stack.pop_back();
        const size_t cc = (size_t)cur.second * nlon + cur.first;
        lake_mask[cc] = 1; lake_id[cc] = count;
        qd_net_neighbors(cur.first, cur.second, nlon, nlat, [&](int ni, int nj, int, int) {
          const size_t nc = (size_t)nj * nlon + ni;
          if (term[nc] && !visited[nc]) { visited[nc] = 1; stack.push_back({ni, nj}); }
          return true;
        });
      }
    }
  return count;
}

// compute_lake_outlets (:230-273)
static void qd_net_outlets(int nlat, int nlon, const double* elev_filled, const uint8_t* lake_mask, const int32_t* lake_id,
                           const uint8_t* land, int n_lakes, int32_t* out) {
  std::vector<std::vector<int64_t>> cells((size_t)n_lakes + 1);
  for (int64_t k = 0; k < (int64_t)nlat * nlon; ++k) { const int id = lake_id[k]; if (id >= 1 && id <= n_lakes) cells[(size_t)id].push_back(k); }   // np.where order
  for (int k = 1; k <= n_lakes; ++k) {
    int64_t best_idx = -1;
    double best_z = std::numeric_limits<double>::infinity();
    bool ocean = false;
    for (int64_t c : cells[(size_t)k]) {
      const int j = (int)(c / nlon), i = (int)(c - (int64_t)j * nlon);
      qd_net_neighbors(i, j, nlon, nlat, [&](int ii, int jj, int, int) {
        const size_t nc = (size_t)jj * nlon + ii;
        if (lake_mask[nc] == 1) return true;
        if (land[nc] == 0) { ocean = true; return false; }
        const double z = elev_filled[nc];
        if (z < best_z) { best_z = z; best_idx = (int64_t)nc; }
        return true;
      });
      if (ocean) break;
    }
    out[k - 1] = ocean ? -1 : (best_idx >= 0 ? (int32_t)best_idx : -1);
  }
}
