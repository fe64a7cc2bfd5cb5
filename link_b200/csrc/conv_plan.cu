// Tile-skipping plan for the tensor-core sparse convolution (conv_tc.cu).
//
// The tcgen05 kernel works on 128-row output tiles and one kernel offset at a time; a (tile, offset)
// step whose 128 rows have NO neighbour at that offset is pure zero work.  On a LiDAR scan only
// ~5.5 of the 27 offsets of a 3^3 submanifold kernel are present per voxel, and in storage order
// half of all (tile, offset) steps are empty.  Grouping output rows by the SIGNS of the offsets
// they use (is there any neighbour with dx < 0, dx > 0, dy < 0, ... : a 6-bit class; ground voxels
// have no dz != 0 neighbours, walls no dx or dy ones, ...) makes tiles homogeneous: 27 % of the
// dense steps remain, and the rows of an active step are 75 % populated (measured on the bench
// scan).  The plan is a counting sort of the output rows by that class:
//
//   class kernel   : per row, the K-bit mask of present offsets and its class; class histogram
//   scatter kernel : perm[position] = row      (class-major order, order inside a class arbitrary)
//                    and tile_mask[position / 128] |= mask of the row
//
// Output values do not depend on the plan: every output row is still the same sum over k in the
// same order, only the tile it is computed in changes.  The reference has no counterpart (its
// gather -> GEMM -> scatter loop touches only existing pairs but pays 3 launches per offset and a
// scatter with atomics, convolution_cuda.cu:101-164).
#include "common.cuh"

#define CP_TILE 128

static inline int64_t cp_al(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t lk_conv_plan_ws_bytes(int64_t n_out) {
  return cp_al(n_out * 4) + cp_al(n_out) + 2 * cp_al(256 * 4);
}

__global__ void __launch_bounds__(256, 4) plan_class_kernel(const int* __restrict__ nbr, int64_t n_out,
                                                         int K, const int* __restrict__ offsets,
                                                         unsigned* __restrict__ rowmask,
                                                         unsigned char* __restrict__ cls,
                                                         unsigned* __restrict__ hist,
                                                         unsigned* __restrict__ tile_mask, int tiles) {
  lk_pdl_enter();
  __shared__ unsigned h[256];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tiles;
       t += (int64_t)gridDim.x * blockDim.x)
    tile_mask[t] = 0u;                             // OR-accumulated by the scatter kernel
  __shared__ unsigned code[32];
  h[threadIdx.x] = 0;
  if (threadIdx.x < 32) {
    unsigned c = 0;
    if ((int)threadIdx.x < K && offsets) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        int d = offsets[3 * threadIdx.x + a];
        if (d < 0) c |= 1u << (2 * a);
        if (d > 0) c |= 1u << (2 * a + 1);
      }
    }
    code[threadIdx.x] = c;
  }
  __syncthreads();
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_out;
       o += (int64_t)gridDim.x * blockDim.x) {
    // all K column loads of the row in flight before the first one is tested (a loop with a branch per
    // load kept one L2 round trip in flight per thread: 12 us for a 13 MB streaming read)
    int v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < K ? __ldg(nbr + (int64_t)k * n_out + o) : -1;
    unsigned m = 0, c = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (v[k] >= 0) { m |= 1u << k; c |= code[k]; }
    if (K <= 8 || !offsets) c = m & 255u;       // small kernels: the mask itself is the class
    rowmask[o] = m;
    cls[o] = (unsigned char)c;
    atomicAdd(&h[c], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) plan_scatter_kernel(const unsigned char* __restrict__ cls,
                                                           int64_t n_out,
                                                           const unsigned* __restrict__ hist,
                                                           unsigned* __restrict__ cursor,
                                                           const unsigned* __restrict__ rowmask,
                                                           int* __restrict__ perm,
                                                           unsigned* __restrict__ tile_mask) {
  lk_pdl_enter();
  __shared__ unsigned base[256];
  __shared__ unsigned wsum[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // exclusive scan of the 256 class counts
  unsigned v = hist[tid], incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  unsigned wb = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) wb += (w < warp) ? wsum[w] : 0u;
  base[tid] = wb + incl - v;
  __syncthreads();
  // Positions are reserved per (CTA, class): ranks inside the CTA come from shared-memory atomics
  // and ONE global atomic per class present in the CTA claims the range, so the 64 hot class
  // cursors see ~10 atomics per CTA instead of one per warp and class.
  __shared__ unsigned cnt_s[256];
  __shared__ unsigned cta_base[256];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (n_out + stride - 1) / stride;
  for (int64_t it = 0; it < rounds; ++it) {       // uniform trip count (block-wide barriers inside)
    cnt_s[tid] = 0;
    __syncthreads();
    const int64_t o = it * stride + blockIdx.x * (int64_t)blockDim.x + tid;
    const bool ok = o < n_out;
    const unsigned c = ok ? cls[o] : 0u;
    const unsigned m = ok ? rowmask[o] : 0u;
    unsigned r = 0;
    if (ok) r = atomicAdd(&cnt_s[c], 1u);
    __syncthreads();
    if (cnt_s[tid]) cta_base[tid] = base[tid] + atomicAdd(&cursor[tid], cnt_s[tid]);
    __syncthreads();
    const unsigned pos = ok ? cta_base[c] + r : 0xFFFFFFFFu;
    if (ok) perm[pos] = (int)o;
    // tile masks: one atomicOr per (warp, tile) -- lanes of a warp mostly land in 1-3 tiles
    const unsigned tile = ok ? pos / CP_TILE : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xffffffffu, tile);
    const unsigned tm = __reduce_or_sync(peers, m);
    if (ok && lane == __ffs(peers) - 1) atomicOr(&tile_mask[tile], tm);
  }
}

extern "C" int lk_conv_plan(const int32_t* d_nbr, int64_t n_out, int k, const int32_t* d_offsets,
                            int32_t* d_perm, uint32_t* d_tile_mask, void* d_ws, int64_t ws_bytes,
                            lk_stream_t s) {
  LK_REQUIRE(n_out >= 0 && k > 0 && k <= 32, "lk_conv_plan: needs 1 <= K <= 32");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_nbr && d_perm && d_tile_mask && d_ws, "lk_conv_plan: null pointer");
  if (ws_bytes < lk_conv_plan_ws_bytes(n_out)) {
    lk_set_error("lk_conv_plan: workspace %lld < %lld bytes", (long long)ws_bytes,
                 (long long)lk_conv_plan_ws_bytes(n_out));
    return LK_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)s;
  char* p = (char*)d_ws;
  unsigned* rowmask = (unsigned*)p; p += cp_al(n_out * 4);
  unsigned char* cls = (unsigned char*)p; p += cp_al(n_out);
  unsigned* hist = (unsigned*)p; p += cp_al(256 * 4);
  unsigned* cursor = (unsigned*)p;
  const int tiles = (int)((n_out + CP_TILE - 1) / CP_TILE);
  LK_CUDA(cudaMemsetAsync(hist, 0, 2 * cp_al(256 * 4), st));
  lk_count_launch();
  const int grid = lk_grid(n_out, 256, 4);
  LK_PDL_LAUNCH(plan_class_kernel, grid, 256, 0, st, d_nbr, n_out, k, d_offsets, rowmask, cls, hist, d_tile_mask, tiles);
  LK_LAUNCHED();
  LK_PDL_LAUNCH(plan_scatter_kernel, grid, 256, 0, st, cls, n_out, hist, cursor, rowmask, d_perm, d_tile_mask);
  LK_LAUNCHED();
  return LK_OK;
}
