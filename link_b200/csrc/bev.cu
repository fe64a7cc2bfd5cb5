// Sparse -> dense bird's-eye-view scatter of the detection backbone's last level and its transpose.
//
// Reference: `ret = self.extra_conv(x).dense(); ret.view(N, C * D, H, W)` (detection/det3d/models/
// backbones/scn.py:612-617): spconv's dense() allocates a zero [B, D, H, W, C] tensor, scatters the rows
// and permutes it to channels-first -- three passes over the 66 MB dense tensor.  Here the zero fill
// is one memset and the rows go straight to their channels-first position: a CTA stages 32 sparse rows
// (coalesced 128-bit loads) in shared memory and writes them channel by channel with lane = row, so
// rows that are neighbours along x (the usual order of the sparse tensor) land in one 128-byte line.
// The backward pass is the same walk with the copies reversed (a pure gather, no atomics).
#include "common.cuh"

#define BEV_ROWS 32
#define BEV_THREADS 256

template <bool SCATTER>
__global__ void __launch_bounds__(BEV_THREADS) bev_kernel(float* __restrict__ sparse /*[n, c]*/,
                                                          const int4* __restrict__ indices /*(b, z, y, x)*/,
                                                          int64_t n, int c, int D, int H, int W,
                                                          float* __restrict__ dense /*[B, c, D, H, W]*/) {
  extern __shared__ float tile[];                 // [BEV_ROWS][c + 1]
  __shared__ int64_t base_s[BEV_ROWS];            // offset of (b, 0, z, y, x), or -1
  const int pitch = c + 1;
  const int64_t plane = (int64_t)D * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ntiles = (n + BEV_ROWS - 1) / BEV_ROWS;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t row0 = t * BEV_ROWS;
    if (threadIdx.x < BEV_ROWS) {
      const int64_t r = row0 + threadIdx.x;
      int64_t b = -1;
      if (r < n) {
        const int4 i = __ldg(indices + r);
        b = (int64_t)i.x * c * plane + ((int64_t)i.y * H + i.z) * W + i.w;
      }
      base_s[threadIdx.x] = b;
    }
    if (SCATTER) {
      for (int e = threadIdx.x; e < BEV_ROWS * c; e += BEV_THREADS) {
        const int r = e / c, ch = e - r * c;
        tile[r * pitch + ch] = row0 + r < n ? __ldg(sparse + (row0 + r) * c + ch) : 0.f;
      }
    }
    __syncthreads();
    const int64_t b = base_s[lane];
    for (int ch = warp; ch < c; ch += BEV_THREADS / 32) {       // lane = row: neighbours along x coalesce
      if (b >= 0) {
        if (SCATTER) dense[b + ch * plane] = tile[lane * pitch + ch];
        else tile[lane * pitch + ch] = __ldg(dense + b + ch * plane);
      }
    }
    if (!SCATTER) {
      __syncthreads();
      for (int e = threadIdx.x; e < BEV_ROWS * c; e += BEV_THREADS) {
        const int r = e / c, ch = e - r * c;
        if (row0 + r < n) sparse[(row0 + r) * c + ch] = tile[r * pitch + ch];
      }
    }
    __syncthreads();
  }
}

static int bev_launch(bool scatter, float* sparse, const int32_t* indices, int64_t n, int c, int B, int D,
                      int H, int W, float* dense, cudaStream_t st) {
  LK_REQUIRE(n >= 0 && c > 0 && c <= 1024 && B > 0 && D > 0 && H > 0 && W > 0, "lk_bev: bad sizes");
  if (scatter) {
    LK_REQUIRE(dense, "lk_bev_scatter: null output");
    LK_CUDA(cudaMemsetAsync(dense, 0, (size_t)B * c * D * H * W * sizeof(float), st));
    lk_count_launch();
  }
  if (n == 0) return LK_OK;
  LK_REQUIRE(sparse && indices && dense && (uintptr_t)indices % 16 == 0, "lk_bev: null or misaligned pointer");
  const size_t smem = (size_t)BEV_ROWS * (c + 1) * sizeof(float);
  const int grid = lk_grid((n + BEV_ROWS - 1) / BEV_ROWS, 1, 8);
  if (scatter)
    bev_kernel<true><<<grid, BEV_THREADS, smem, st>>>(sparse, (const int4*)indices, n, c, D, H, W, dense);
  else
    bev_kernel<false><<<grid, BEV_THREADS, smem, st>>>(sparse, (const int4*)indices, n, c, D, H, W, dense);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_bev_scatter(const float* d_feats, const int32_t* d_indices, int64_t n, int c, int batch,
                              int depth, int height, int width, float* d_dense, lk_stream_t s) {
  return bev_launch(true, (float*)d_feats, d_indices, n, c, batch, depth, height, width, d_dense, (cudaStream_t)s);
}

extern "C" int lk_bev_gather(const float* d_dense, const int32_t* d_indices, int64_t n, int c, int batch,
                             int depth, int height, int width, float* d_feats, lk_stream_t s) {
  return bev_launch(false, d_feats, d_indices, n, c, batch, depth, height, width, (float*)d_dense, (cudaStream_t)s);
}
