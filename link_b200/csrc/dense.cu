// Fused Linear (no bias) + LayerNorm over feature rows: out = LN(x @ W^T; gamma, beta, eps).
// This is ELKBlock.pre_mix (linkencoder.py:112-115, 132).  The reference runs a cuBLAS SGEMM and
// a separate LayerNorm kernel (which alone costs 182 us at N=120k, C=64 on B200: one row per
// warp with C=64 leaves the machine mostly idle); fusing keeps the [N,C] GEMM result in registers
// through the normalisation, so x is read once and out is written once.
//
// Tile: 64 rows x C outputs per CTA (C = 16*CPT), 16x16 thread grid, 4 rows x CPT columns per
// thread; the 16 threads that share a row are a half-warp, so LayerNorm statistics are four
// xor-shuffles.  fp32 FFMA (the reference's accumulate type).
#include "common.cuh"

#define LL_ROWS 64
#define LL_THREADS 256

template <int CPT>
__global__ void __launch_bounds__(LL_THREADS) linear_ln_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int64_t n, float* __restrict__ out) {
  constexpr int C = 16 * CPT;
  extern __shared__ __align__(16) float smem[];
  float (*ws)[C] = (float (*)[C])smem;                    // ws[k][o] = W[o][k]
  float (*xs)[LL_ROWS + 4] = (float (*)[LL_ROWS + 4])(smem + C * C);   // xs[k][r]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  for (int i = tid; i < C * C; i += LL_THREADS) {
    int o = i / C, k = i % C;                             // coalesced read of W[o][k]
    ws[k][o] = __ldg(w + i);
  }
  for (int64_t row0 = (int64_t)blockIdx.x * LL_ROWS; row0 < n; row0 += (int64_t)gridDim.x * LL_ROWS) {
    __syncthreads();
    for (int i = tid; i < LL_ROWS * (C / 4); i += LL_THREADS) {
      int r = i / (C / 4), k4 = (i % (C / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < n) v = lk_ldg_stream((const float4*)(x + (row0 + r) * C + k4));
      xs[k4 + 0][r] = v.x; xs[k4 + 1][r] = v.y; xs[k4 + 2][r] = v.z; xs[k4 + 3][r] = v.w;
    }
    __syncthreads();
    float acc[4][CPT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < C; ++k) {
      float4 a = *(const float4*)&xs[k][ty * 4];
      float b[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) b[j] = ws[k][tx * CPT + j];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        acc[0][j] += a.x * b[j]; acc[1][j] += a.y * b[j];
        acc[2][j] += a.z * b[j]; acc[3][j] += a.w * b[j];
      }
    }
    float g[CPT], bt[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { g[j] = __ldg(gamma + tx * CPT + j); bt[j] = __ldg(beta + tx * CPT + j); }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) s += acc[i][j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      float mean = s / (float)C, q = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) { acc[i][j] -= mean; q += acc[i][j] * acc[i][j]; }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      float rstd = 1.0f / sqrtf(q / (float)C + eps);
      int64_t r = row0 + ty * 4 + i;
      if (r < n) {
        float* dst = out + r * C + tx * CPT;
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = acc[i][j] * rstd * g[j] + bt[j];
        if (CPT % 4 == 0) {
#pragma unroll
          for (int j = 0; j < CPT; j += 4)
            lk_stg_stream((float4*)(dst + j), make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < CPT; ++j) dst[j] = acc[i][j];
        }
      }
    }
  }
}

template <int CPT>
static int launch_linear_ln(const float* x, const float* w, const float* g, const float* b, float eps,
                            int64_t n, float* out, cudaStream_t st) {
  constexpr int C = 16 * CPT;
  size_t smem = (size_t)(C * C + C * (LL_ROWS + 4)) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    LK_CUDA(cudaFuncSetAttribute(linear_ln_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int grid = lk_grid(n, LL_ROWS, 6);
  linear_ln_kernel<CPT><<<grid, LL_THREADS, smem, st>>>(x, w, g, b, eps, n, out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_linear_ln_fwd(const float* d_x, const float* d_w, const float* d_gamma,
                                const float* d_beta, float eps, int64_t n, int c, float* d_out,
                                lk_stream_t s) {
  LK_REQUIRE(n >= 0 && (c == 16 || c == 32 || c == 64 || c == 128),
             "lk_linear_ln_fwd: C must be 16, 32, 64 or 128");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_x && d_w && d_gamma && d_beta && d_out, "lk_linear_ln_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_x % 16 == 0 && (uintptr_t)d_out % 16 == 0, "lk_linear_ln_fwd: alignment");
  cudaStream_t st = (cudaStream_t)s;
  switch (c) {
    case 16: return launch_linear_ln<1>(d_x, d_w, d_gamma, d_beta, eps, n, d_out, st);
    case 32: return launch_linear_ln<2>(d_x, d_w, d_gamma, d_beta, eps, n, d_out, st);
    case 64: return launch_linear_ln<4>(d_x, d_w, d_gamma, d_beta, eps, n, d_out, st);
    default: return launch_linear_ln<8>(d_x, d_w, d_gamma, d_beta, eps, n, d_out, st);
  }
}
