// Generic scatter-mean / weighted-gather ops with the reference's semantics.
// Replaces backend/voxelize/voxelize_cuda.cu and backend/devoxelize/devoxelize_cuda.cu.
// The reference launches <<<N, c>>> (one CTA per row, c threads); here a thread owns a
// 4-channel vector of one row, grids are grid-stride over rows*vectors, feature rows move
// as 128-bit transactions and the scatter uses one vector red.global.add per 4 floats.
#include "common.cuh"

template <int VEC>
__global__ void __launch_bounds__(256) voxelize_fwd_kernel(const float* __restrict__ feats,
                                                           const int* __restrict__ idx,
                                                           const int* __restrict__ counts,
                                                           int64_t n, int64_t m, int c,
                                                           float* out) {
  int vpr = c / VEC;
  int64_t total = n * vpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / vpr;
    int j = (int)(t - i * vpr) * VEC;
    int pos = idx[i];
    if (pos < 0 || pos >= m) continue;
    int cnt = counts[pos];
    if (cnt == 0) continue;
    float fc = (float)cnt;
    if (VEC == 4) {
      float4 v = *(const float4*)(feats + i * c + j);
      v.x /= fc; v.y /= fc; v.z /= fc; v.w /= fc;
      lk_red_add_v4(out + (int64_t)pos * c + j, v);
    } else {
      atomicAdd(out + (int64_t)pos * c + j, feats[i * c + j] / fc);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) voxelize_bwd_kernel(const float* __restrict__ top,
                                                           const int* __restrict__ idx,
                                                           const int* __restrict__ counts,
                                                           int64_t n, int64_t m, int c,
                                                           float* __restrict__ bottom) {
  int vpr = c / VEC;
  int64_t total = n * vpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / vpr;
    int j = (int)(t - i * vpr) * VEC;
    int pos = idx[i];
    bool ok = pos >= 0 && pos < m;
    int cnt = ok ? counts[pos] : 0;
    float fc = (float)cnt;
    if (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cnt > 0) {
        v = *(const float4*)(top + (int64_t)pos * c + j);
        v.x /= fc; v.y /= fc; v.z /= fc; v.w /= fc;
      }
      *(float4*)(bottom + i * c + j) = v;
    } else {
      bottom[i * c + j] = cnt > 0 ? top[(int64_t)pos * c + j] / fc : 0.f;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) devoxelize_fwd_kernel(const float* __restrict__ feat,
                                                             const int* __restrict__ idx,
                                                             const float* __restrict__ w,
                                                             int64_t N, int R, int c,
                                                             float* __restrict__ out) {
  int vpr = c / VEC;
  int64_t total = N * vpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / vpr;
    int j = (int)(t - i * vpr) * VEC;
    const int* ii = idx + i * R;
    const float* ww = w + i * R;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < R; ++k) {
      int src = ii[k];
      if (src < 0) continue;
      float wk = ww[k];
      if (VEC == 4) {
        float4 v = __ldg((const float4*)(feat + (int64_t)src * c + j));
        acc.x += wk * v.x; acc.y += wk * v.y; acc.z += wk * v.z; acc.w += wk * v.w;
      } else {
        acc.x += wk * __ldg(feat + (int64_t)src * c + j);
      }
    }
    if (VEC == 4) *(float4*)(out + i * c + j) = acc;
    else out[i * c + j] = acc.x;
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) devoxelize_bwd_kernel(const float* __restrict__ top,
                                                             const int* __restrict__ idx,
                                                             const float* __restrict__ w,
                                                             int64_t N, int R, int c, int64_t n,
                                                             float* bottom) {
  int vpr = c / VEC;
  int64_t total = N * vpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / vpr;
    int j = (int)(t - i * vpr) * VEC;
    const int* ii = idx + i * R;
    const float* ww = w + i * R;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (VEC == 4) g = *(const float4*)(top + i * c + j);
    else g.x = top[i * c + j];
    for (int k = 0; k < R; ++k) {
      int dst = ii[k];
      if (dst < 0 || dst >= n) continue;
      float wk = ww[k];
      if (VEC == 4)
        lk_red_add_v4(bottom + (int64_t)dst * c + j,
                      make_float4(wk * g.x, wk * g.y, wk * g.z, wk * g.w));
      else
        atomicAdd(bottom + (int64_t)dst * c + j, wk * g.x);
    }
  }
}

static inline bool vec_ok(const void* a, const void* b, int c) {
  return c % 4 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}

extern "C" int lk_voxelize_fwd(const float* d_feats, const int32_t* d_idx, const int32_t* d_counts,
                               int64_t n, int64_t m, int c, float* d_out, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && m >= 0 && c > 0, "lk_voxelize_fwd: bad sizes");
  cudaStream_t st = (cudaStream_t)s;
  if (m > 0) {
    LK_REQUIRE(d_out, "lk_voxelize_fwd: null output");
    LK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)m * c * sizeof(float), st));
    lk_count_launch();
  }
  if (n == 0 || m == 0) return LK_OK;
  LK_REQUIRE(d_feats && d_idx && d_counts, "lk_voxelize_fwd: null input");
  if (vec_ok(d_feats, d_out, c))
    voxelize_fwd_kernel<4><<<lk_grid(n * (c / 4), 256, 8), 256, 0, st>>>(d_feats, d_idx, d_counts, n, m, c, d_out);
  else
    voxelize_fwd_kernel<1><<<lk_grid(n * c, 256, 8), 256, 0, st>>>(d_feats, d_idx, d_counts, n, m, c, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_voxelize_bwd(const float* d_top, const int32_t* d_idx, const int32_t* d_counts,
                               int64_t n, int64_t m, int c, float* d_bottom, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && m >= 0 && c > 0, "lk_voxelize_bwd: bad sizes");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_idx && d_counts && d_bottom && (m == 0 || d_top), "lk_voxelize_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)s;
  if (vec_ok(d_top, d_bottom, c))
    voxelize_bwd_kernel<4><<<lk_grid(n * (c / 4), 256, 8), 256, 0, st>>>(d_top, d_idx, d_counts, n, m, c, d_bottom);
  else
    voxelize_bwd_kernel<1><<<lk_grid(n * c, 256, 8), 256, 0, st>>>(d_top, d_idx, d_counts, n, m, c, d_bottom);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_devoxelize_fwd(const float* d_feat, const int32_t* d_idx, const float* d_w,
                                 int64_t N, int r3, int c, float* d_out, lk_stream_t s) {
  LK_REQUIRE(N >= 0 && r3 > 0 && c > 0, "lk_devoxelize_fwd: bad sizes");
  if (N == 0) return LK_OK;
  LK_REQUIRE(d_idx && d_w && d_out, "lk_devoxelize_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)s;
  if (vec_ok(d_feat, d_out, c))
    devoxelize_fwd_kernel<4><<<lk_grid(N * (c / 4), 256, 8), 256, 0, st>>>(d_feat, d_idx, d_w, N, r3, c, d_out);
  else
    devoxelize_fwd_kernel<1><<<lk_grid(N * c, 256, 8), 256, 0, st>>>(d_feat, d_idx, d_w, N, r3, c, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_devoxelize_bwd(const float* d_top, const int32_t* d_idx, const float* d_w,
                                 int64_t N, int r3, int c, int64_t n, float* d_bottom,
                                 lk_stream_t s) {
  LK_REQUIRE(N >= 0 && n >= 0 && r3 > 0 && c > 0, "lk_devoxelize_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)s;
  if (n > 0) {
    LK_REQUIRE(d_bottom, "lk_devoxelize_bwd: null output");
    LK_CUDA(cudaMemsetAsync(d_bottom, 0, (size_t)n * c * sizeof(float), st));
    lk_count_launch();
  }
  if (N == 0 || n == 0) return LK_OK;
  LK_REQUIRE(d_top && d_idx && d_w, "lk_devoxelize_bwd: null input");
  if (vec_ok(d_top, d_bottom, c))
    devoxelize_bwd_kernel<4><<<lk_grid(N * (c / 4), 256, 8), 256, 0, st>>>(d_top, d_idx, d_w, N, r3, c, n, d_bottom);
  else
    devoxelize_bwd_kernel<1><<<lk_grid(N * c, 256, 8), 256, 0, st>>>(d_top, d_idx, d_w, N, r3, c, n, d_bottom);
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- classifier head gather
// out[i, g*c : (g+1)*c] = relu( src_g[idx_g[i], :] + bias[g*c : (g+1)*c] ) for g < G; idx_g == NULL
// means the identity map.  Used by the ELKEncoder head (linkencoder.py:371-379): the grouped 1x1
// conv is applied at each level's own (coarse) resolution and only its c-channel result is
// upsampled, instead of upsampling 64-channel features and concatenating them.
struct GatherArgs {
  const float* src[8];
  const int64_t* idx[8];
};

__global__ void __launch_bounds__(256) gather_concat_bias_relu_kernel(GatherArgs a, int G, int c,
                                                                      int64_t n,
                                                                      const float* __restrict__ bias,
                                                                      int relu, float* __restrict__ out) {
  const int vpr = c >> 2;                 // float4 per source row
  const int64_t total = n * G * vpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / (G * vpr);
    int rem = (int)(t - i * (G * vpr));
    int g = rem / vpr, j = (rem - g * vpr) * 4;
    int64_t row = a.idx[g] ? __ldg(a.idx[g] + i) : i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0) v = __ldg((const float4*)(a.src[g] + row * c + j));
    if (bias) {
      float4 b = __ldg((const float4*)(bias + g * c + j));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *(float4*)(out + i * (int64_t)(G * c) + g * c + j) = v;
  }
}

extern "C" int lk_gather_concat(const float* const* d_src, const int64_t* const* d_idx, int groups,
                                int c, int64_t n, const float* d_bias, int relu, float* d_out,
                                lk_stream_t s) {
  LK_REQUIRE(groups >= 1 && groups <= 8 && c > 0 && c % 4 == 0 && n >= 0, "lk_gather_concat: bad sizes");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_src && d_idx && d_out, "lk_gather_concat: null pointer");
  GatherArgs a;
  for (int g = 0; g < 8; ++g) {
    a.src[g] = g < groups ? d_src[g] : nullptr;
    a.idx[g] = g < groups ? d_idx[g] : nullptr;
    LK_REQUIRE(g >= groups || a.src[g], "lk_gather_concat: null source");
  }
  gather_concat_bias_relu_kernel<<<lk_grid(n * groups * (c / 4), 256, 8), 256, 0, (cudaStream_t)s>>>(
      a, groups, c, n, d_bias, relu, d_out);
  LK_LAUNCHED();
  return LK_OK;
}
