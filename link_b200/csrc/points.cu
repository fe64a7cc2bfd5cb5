// points -> voxels on the device (reference: det3d/ops/point_cloud/point_cloud_ops.py:8-183,
// `points_to_voxel`: a single-threaded numba loop in the data loader).
//
// Reference semantics, reproduced bit for bit: a point is kept iff floor((p - range_min) / voxel_size)
// lies inside the grid (float32 arithmetic, like the numpy arrays of the reference); voxels are
// numbered in order of FIRST APPEARANCE in the point sequence, only the first `max_voxels` of them
// exist; a voxel keeps its first `max_points` points in point order.
//
// Parallel formulation (two radix sorts of the existing sort/unique primitive, no atomics, so the
// result is deterministic):
//   keys kernel  : key = linear cell index (z,y,x) per point, `cells` (one past the last) if rejected
//   sort/unique  : stable, so inside a cell the point indices ascend -> first[u] = order[seg[u]]
//   sort first[] : rank of a cell among the first appearances = its voxel id
//   fill kernel  : one warp per cell copies up to max_points rows, writes coordinate and count
#include "common.cuh"


struct PvGrid {
  float lo[3], vs[3];
  int gs[3];
};

__global__ void __launch_bounds__(256) pv_keys_kernel(const float* __restrict__ pts, int64_t n, int ndim,
                                                      PvGrid g, unsigned long long invalid,
                                                      unsigned long long* __restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    bool ok = true;
    int c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = floorf(__fdiv_rn(__fsub_rn(pts[i * ndim + j], g.lo[j]), g.vs[j]));
      ok = ok && v >= 0.f && v < (float)g.gs[j];
      c[j] = ok ? (int)v : 0;
    }
    // (z, y, x) linear index, z most significant: ascending key == ascending (z, y, x)
    keys[i] = ok ? ((unsigned long long)c[2] * g.gs[1] + c[1]) * g.gs[0] + c[0] : invalid;
  }
}

// first appearance of each cell as a sort key (the invalid cell, if present, sorts last and is dropped)
__global__ void __launch_bounds__(256) pv_first_kernel(const unsigned long long* __restrict__ uniq,
                                                       const int* __restrict__ order,
                                                       const int* __restrict__ seg,
                                                       const int* __restrict__ d_num, int64_t cap,
                                                       unsigned long long invalid_cell,
                                                       unsigned long long* __restrict__ first_keys) {
  int64_t m = *d_num;
  if (m > cap) m = cap;
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < cap;
       u += (int64_t)gridDim.x * blockDim.x)
    first_keys[u] = (u < m && uniq[u] != invalid_cell) ? (unsigned long long)order[seg[u]] : (unsigned long long)cap;
}

// one warp per voxel id v (rank of the cell's first appearance): cell u = cell_of[v]
__global__ void __launch_bounds__(256) pv_fill_kernel(
    const float* __restrict__ pts, int ndim, const unsigned long long* __restrict__ uniq,
    const int* __restrict__ order, const int* __restrict__ seg, const int* __restrict__ cell_of,
    const unsigned long long* __restrict__ first_sorted, const int* __restrict__ d_cells, int64_t cap,
    PvGrid g, int max_points, int max_voxels, float* __restrict__ voxels, int* __restrict__ coors,
    int* __restrict__ num_points, int* __restrict__ d_voxel_num) {
  int64_t m = *d_cells;
  if (m > cap) m = cap;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < m && v < max_voxels;
       v += warps_total) {
    if (first_sorted[v] == (unsigned long long)cap) continue;   // the rejected-points cell (sorted last)
    const int u = cell_of[v];
    const int s0 = seg[u], cnt = seg[u + 1] - s0;
    const int keep = cnt < max_points ? cnt : max_points;
    const int row = max_points * ndim;
    for (int t = lane; t < keep * ndim; t += 32) {
      const int j = t / ndim, d = t - j * ndim;
      voxels[v * (int64_t)row + t] = pts[(int64_t)order[s0 + j] * ndim + d];
    }
    for (int t = keep * ndim + lane; t < row; t += 32) voxels[v * (int64_t)row + t] = 0.f;
    if (lane == 0) {
      unsigned long long k = uniq[u];
      const int x = (int)(k % g.gs[0]); k /= g.gs[0];
      const int y = (int)(k % g.gs[1]); k /= g.gs[1];
      coors[v * 3 + 0] = (int)k; coors[v * 3 + 1] = y; coors[v * 3 + 2] = x;      // (z, y, x)
      num_points[v] = keep;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // number of voxels = valid cells, capped
    int64_t valid = m;
    if (m > 0 && first_sorted[m - 1] == (unsigned long long)cap) valid = m - 1;
    *d_voxel_num = (int)(valid < max_voxels ? valid : max_voxels);
  }
}

static inline int64_t pv_al(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t lk_points_to_voxel_ws_bytes(int64_t n) {
  // keys, uniq, first keys, first sorted (8 B each) + order, seg, cell_of (4 B each) + num + 2 sort workspaces
  return 4 * pv_al(n * 8) + 2 * pv_al(n * 4) + pv_al((n + 1) * 4) + 512 + 2 * pv_al(lk_sort_unique_ws_bytes(n));
}

extern "C" int lk_points_to_voxel(const float* d_points, int64_t n, int ndim, const float* voxel_size,
                                  const float* coors_range, int max_points, int max_voxels,
                                  float* d_voxels, int32_t* d_coors, int32_t* d_num_points,
                                  int32_t* d_voxel_num, void* d_ws, int64_t ws_bytes, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && ndim >= 3 && voxel_size && coors_range && max_points > 0 && max_voxels > 0 &&
                 d_voxel_num, "lk_points_to_voxel: bad arguments");
  cudaStream_t st = (cudaStream_t)s;
  if (n == 0) {
    LK_CUDA(cudaMemsetAsync(d_voxel_num, 0, 4, st));
    lk_count_launch();
    return LK_OK;
  }
  LK_REQUIRE(d_points && d_voxels && d_coors && d_num_points && d_ws, "lk_points_to_voxel: null pointer");
  if (ws_bytes < lk_points_to_voxel_ws_bytes(n)) {
    lk_set_error("lk_points_to_voxel: workspace %lld < %lld bytes", (long long)ws_bytes,
                 (long long)lk_points_to_voxel_ws_bytes(n));
    return LK_ENOSPC;
  }
  PvGrid g;
  int64_t cells = 1;
  for (int j = 0; j < 3; ++j) {
    g.lo[j] = coors_range[j];
    g.vs[j] = voxel_size[j];
    // grid_size = round((max - min) / voxel_size) in float32, as numpy does for float32 inputs
    g.gs[j] = (int)nearbyintf((coors_range[3 + j] - coors_range[j]) / voxel_size[j]);
    LK_REQUIRE(g.gs[j] > 0, "lk_points_to_voxel: empty grid");
    cells *= g.gs[j];
  }
  int cell_bits = 1;                     // keys are in [0, cells]  (cells = rejected)
  while (cell_bits < 63 && (1LL << cell_bits) <= cells) ++cell_bits;
  int n_bits = 1;                        // first-appearance keys are in [0, n]  (n = no valid cell)
  while ((1LL << n_bits) <= n) ++n_bits;
  char* p = (char*)d_ws;
  unsigned long long* keys = (unsigned long long*)p; p += pv_al(n * 8);
  unsigned long long* uniq = (unsigned long long*)p; p += pv_al(n * 8);
  unsigned long long* first_keys = (unsigned long long*)p; p += pv_al(n * 8);
  unsigned long long* first_sorted = (unsigned long long*)p; p += pv_al(n * 8);
  int* order = (int*)p; p += pv_al(n * 4);
  int* cell_of = (int*)p; p += pv_al(n * 4);
  int* seg = (int*)p; p += pv_al((n + 1) * 4);
  int* num_cells = (int*)p; p += 256;
  int* num2 = (int*)p; p += 256;
  void* sws1 = p; p += pv_al(lk_sort_unique_ws_bytes(n));
  void* sws2 = p;
  pv_keys_kernel<<<lk_grid(n, 256, 8), 256, 0, st>>>(d_points, n, ndim, g, (unsigned long long)cells, keys);
  LK_LAUNCHED();
  int rc = lk_sort_unique((const uint64_t*)keys, n, cell_bits, (uint64_t*)uniq, nullptr, order, seg, nullptr,
                          num_cells, sws1, lk_sort_unique_ws_bytes(n), s);
  if (rc) return rc;
  pv_first_kernel<<<lk_grid(n, 256, 8), 256, 0, st>>>(uniq, order, seg, num_cells, n, (unsigned long long)cells, first_keys);
  LK_LAUNCHED();
  // rank the cells by first appearance (keys < n, or all ones): order2[v] = cell with the v-th first point
  rc = lk_sort_unique((const uint64_t*)first_keys, n, n_bits, (uint64_t*)first_sorted, nullptr, cell_of, nullptr,
                      nullptr, num2, sws2, lk_sort_unique_ws_bytes(n), s);
  if (rc) return rc;
  const int64_t warps = n < max_voxels ? n : max_voxels;
  pv_fill_kernel<<<lk_grid(warps * 32, 256, 8), 256, 0, st>>>(
      d_points, ndim, uniq, order, seg, cell_of, first_sorted, num_cells, n, g, max_points, max_voxels,
      d_voxels, d_coors, d_num_points, d_voxel_num);
  LK_LAUNCHED();
  return LK_OK;
}
