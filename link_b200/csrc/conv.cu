// Sparse 3D convolution on an output-stationary kernel map.
// Replaces convolution_{forward,backward}_cuda (backend/convolution/convolution_cuda.cu:53-278):
// the reference loops over the K offsets on the HOST and launches gather -> cuBLAS mm -> scatter
// per offset (3 launches + 2 full-size temporaries each).  Here one kernel owns a tile of BM
// output rows, walks the K offsets itself and, per offset, compacts the rows that actually have a
// neighbour (nu/K ~ 30 % for LiDAR) so no FLOP is spent on missing neighbours; partial products
// are accumulated in a shared-memory output tile, so there are no atomics, no temporaries in HBM
// and the output is written exactly once.
//
// v1 arithmetic: fp32 FFMA register tiles (bit-for-bit the reference's fp32 accumulate type).
#include "common.cuh"

#define CV_BM 256      // output rows per CTA (~ nu/K * 256 = 64-96 compacted rows per offset)
#define CV_SR 64       // compacted rows per GEMM sub-tile
#define CV_BN 64       // output channels per CTA (grid.y tiles wider layers)
#define CV_BK 32       // input-channel chunk
#define CV_THREADS 256
#define CV_SMEM_BYTES                                                                     \
  ((size_t)CV_BM * (CV_BN + 4) * 4 + (size_t)CV_BK * (CV_SR + 4) * 4 + (size_t)CV_BK * CV_BN * 4 + \
   (size_t)2 * CV_BM * 4)

__global__ void __launch_bounds__(CV_THREADS) conv_fwd_kernel(
    const float* __restrict__ in, const float* __restrict__ w, const int* __restrict__ nbr,
    int64_t n_out, int K, int c_in, int c_out, lk_conv_epilogue_t ep,
    float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*s_acc)[CV_BN + 4] = (float (*)[CV_BN + 4])smem_raw;                 // output tile
  float (*s_a)[CV_SR + 4] = (float (*)[CV_SR + 4])(s_acc + CV_BM);            // gathered inputs, channel-major
  float (*s_b)[CV_BN] = (float (*)[CV_BN])(s_a + CV_BK);                      // weight chunk
  int* s_in = (int*)(s_b + CV_BK);                                            // compacted: input row
  int* s_m = s_in + CV_BM;                                                    // compacted: local output row
  __shared__ int s_cnt;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                  // 16 x 16 thread grid, 4x4 outputs each
  const int64_t row0 = (int64_t)blockIdx.x * CV_BM;
  const int n0 = blockIdx.y * CV_BN;

  for (int i = tid; i < CV_BM * (CV_BN + 4); i += CV_THREADS) (&s_acc[0][0])[i] = 0.f;
  static_assert(CV_BM <= CV_THREADS, "one thread per output row in the compaction step");

  for (int k = 0; k < K; ++k) {
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (tid < CV_BM) {
      int64_t o = row0 + tid;
      int src = (o < n_out) ? __ldg(nbr + (int64_t)k * n_out + o) : -1;
      if (src >= 0) {
        int slot = atomicAdd(&s_cnt, 1);   // order within an offset is irrelevant: every output
        s_in[slot] = src;                  // row receives at most one contribution per offset
        s_m[slot] = tid;
      }
    }
    __syncthreads();
    const int v = s_cnt;
    if (v == 0) continue;
    for (int sub = 0; sub < v; sub += CV_SR) {
      const int rows = min(CV_SR, v - sub);
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int kc = 0; kc < c_in; kc += CV_BK) {
        // gather A: CV_SR rows x CV_BK channels, one float4 (4 channels) per thread-iteration
        for (int t = tid; t < CV_SR * (CV_BK / 4); t += CV_THREADS) {
          int r = t / (CV_BK / 4);
          int kk = (t % (CV_BK / 4)) * 4;
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < rows) {
            const float* src = in + (int64_t)s_in[sub + r] * c_in + kc + kk;
            if (kc + kk + 3 < c_in && (c_in & 3) == 0) {
              val = __ldg((const float4*)src);
            } else {
              if (kc + kk + 0 < c_in) val.x = __ldg(src + 0);
              if (kc + kk + 1 < c_in) val.y = __ldg(src + 1);
              if (kc + kk + 2 < c_in) val.z = __ldg(src + 2);
              if (kc + kk + 3 < c_in) val.w = __ldg(src + 3);
            }
          }
          s_a[kk + 0][r] = val.x; s_a[kk + 1][r] = val.y;
          s_a[kk + 2][r] = val.z; s_a[kk + 3][r] = val.w;
        }
        // weight chunk W[k][kc:kc+BK][n0:n0+BN]
        for (int t = tid; t < CV_BK * CV_BN; t += CV_THREADS) {
          int kk = t / CV_BN, nn = t % CV_BN;
          float val = 0.f;
          if (kc + kk < c_in && n0 + nn < c_out)
            val = __ldg(w + ((int64_t)k * c_in + kc + kk) * c_out + n0 + nn);
          s_b[kk][nn] = val;
        }
        __syncthreads();
        if (ty * 4 < rows) {
#pragma unroll 8
          for (int kk = 0; kk < CV_BK; ++kk) {
            float4 a = *(const float4*)&s_a[kk][ty * 4];
            float4 b = *(const float4*)&s_b[kk][tx * 4];
            acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
            acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
            acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
            acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
          }
        }
        __syncthreads();
      }
      // fold the sub-tile into the output tile (rows are distinct within one offset)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = ty * 4 + i;
        if (r < rows) {
          float* dst = &s_acc[s_m[sub + r]][tx * 4];
          dst[0] += acc[i][0]; dst[1] += acc[i][1]; dst[2] += acc[i][2]; dst[3] += acc[i][3];
        }
      }
    }
  }
  __syncthreads();
  for (int t = tid; t < CV_BM * CV_BN; t += CV_THREADS) {
    int r = t / CV_BN, nn = t % CV_BN;
    int64_t o = row0 + r;
    if (o < n_out && n0 + nn < c_out) {
      float v = s_acc[r][nn];
      if (ep.d_scale) v *= __ldg(ep.d_scale + n0 + nn);
      if (ep.d_shift) v += __ldg(ep.d_shift + n0 + nn);
      if (ep.d_residual) v += __ldg(ep.d_residual + o * c_out + n0 + nn);
      if (ep.relu) v = fmaxf(v, 0.f);
      out[o * c_out + n0 + nn] = v;
    }
  }
}

extern "C" int lk_conv_fwd(const float* d_in, const float* d_w, const int32_t* d_nbr, int64_t n_out,
                           int k, int c_in, int c_out, const float* d_bias, float* d_out,
                           lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, d_bias, nullptr, 0, 0};
  return lk_conv_fwd_ex(d_in, d_w, d_nbr, n_out, k, c_in, c_out, &ep, d_out, s);
}

extern "C" int lk_conv_fwd_ex(const float* d_in, const float* d_w, const int32_t* d_nbr,
                              int64_t n_out, int k, int c_in, int c_out,
                              const lk_conv_epilogue_t* epp, float* d_out, lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, nullptr, nullptr, 0, 0};
  if (epp) ep = *epp;
  LK_REQUIRE(n_out >= 0 && k > 0 && c_in > 0 && c_out > 0, "lk_conv_fwd: bad sizes");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in && d_w && d_nbr && d_out, "lk_conv_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_in % 16 == 0, "lk_conv_fwd: input features must be 16-byte aligned");
  dim3 grid((unsigned)((n_out + CV_BM - 1) / CV_BM), (unsigned)((c_out + CV_BN - 1) / CV_BN));
  static bool attr_set = false;
  if (!attr_set) {
    LK_CUDA(cudaFuncSetAttribute(conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)CV_SMEM_BYTES));
    attr_set = true;
  }
  conv_fwd_kernel<<<grid, CV_THREADS, CV_SMEM_BYTES, (cudaStream_t)s>>>(d_in, d_w, d_nbr, n_out, k, c_in,
                                                           c_out, ep, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

// grad_w[k] = sum over pairs (i -> o) of offset k of in[i]^T (x) grad_out[o].
// grid = (K, splits): CTA (k, s) walks its slice of output rows, keeps its share of the
// [c_in, c_out] tile in registers and flushes once with vector reductions.
#define CW_THREADS 256
#define CW_MAX_ITEMS 16
// grad_w[k] = sum over the pairs (i -> o) of offset k of  in[i]^T gout[o].
// One CTA owns offset k and a chunk of output rows.  Per 256 candidate rows the present pairs are
// compacted (ballot + prefix) into a shared list, then processed 32 pairs at a time: the 32 input
// rows and 32 grad rows are staged in shared memory with coalesced 128-bit loads and every thread
// accumulates its [1 x 4] patches of the C_in x C_out product from there (the `a` operand is a
// warp broadcast, the grad vectors are consecutive: no bank conflicts).  The previous version
// walked the rows one by one with dependent global loads (7.7 ms per conv at N = 108k, C = 64-128).
#define CW_TR 32
__global__ void __launch_bounds__(CW_THREADS) conv_bwd_weight_kernel(
    const float* __restrict__ in, const float* __restrict__ gout, const int* __restrict__ nbr,
    int64_t n_out, int c_in, int c_out, float* gw) {
  extern __shared__ float4 cw_smem[];
  float* a_s = (float*)cw_smem;                      // [CW_TR][c_in]
  float* g_s = a_s + CW_TR * c_in;                   // [CW_TR][c_out]
  __shared__ int list_i[CW_THREADS], list_o[CW_THREADS];
  __shared__ int wcnt[CW_THREADS / 32];
  const int k = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int vpr = c_out >> 2;                       // float4 per weight row
  const int items = (c_in * vpr + CW_THREADS - 1) / CW_THREADS;
  float4 acc[CW_MAX_ITEMS];
#pragma unroll
  for (int q = 0; q < CW_MAX_ITEMS; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t per = (n_out + gridDim.y - 1) / gridDim.y;
  int64_t o0 = (int64_t)blockIdx.y * per, o1 = min(n_out, o0 + per);
  const int va = c_in >> 2;                          // float4 per input row
  const bool quad = (c_in & 3) == 0;                 // 4 x 4 register patches (else scalar input channels)
  const int ca = c_in >> 2;
  const int items4 = (ca * vpr + CW_THREADS - 1) / CW_THREADS;
  for (int64_t base = o0; base < o1; base += CW_THREADS) {
    // ---- compact the present pairs of 256 candidate rows ----
    const int64_t o = base + tid;
    const int i = (o < o1) ? __ldg(nbr + (int64_t)k * n_out + o) : -1;
    const unsigned bal = __ballot_sync(0xffffffffu, i >= 0);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    int pos = __popc(bal & ((1u << lane) - 1u)), total = 0;
#pragma unroll
    for (int w = 0; w < CW_THREADS / 32; ++w) {
      pos += (w < warp) ? wcnt[w] : 0;
      total += wcnt[w];
    }
    if (i >= 0) { list_i[pos] = i; list_o[pos] = (int)(o - o0); }
    __syncthreads();
    // ---- 32 pairs at a time ----
    for (int p0 = 0; p0 < total; p0 += CW_TR) {
      const int np = min(CW_TR, total - p0);
      if ((c_in & 3) == 0) {
        for (int t = tid; t < CW_TR * va; t += CW_THREADS) {
          const int r = t / va, v = t - r * va;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < np) x = __ldg((const float4*)(in + (int64_t)list_i[p0 + r] * c_in) + v);
          ((float4*)a_s)[r * va + v] = x;
        }
      } else {                                         // odd channel counts (the 5-channel detection stem)
        for (int t = tid; t < CW_TR * c_in; t += CW_THREADS) {
          const int r = t / c_in, cch = t - r * c_in;
          a_s[t] = r < np ? __ldg(in + (int64_t)list_i[p0 + r] * c_in + cch) : 0.f;
        }
      }
      for (int t = tid; t < CW_TR * vpr; t += CW_THREADS) {
        const int r = t / vpr, v = t - r * vpr;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < np) x = __ldg((const float4*)(gout + (o0 + list_o[p0 + r]) * c_out) + v);
        ((float4*)g_s)[r * vpr + v] = x;
      }
      __syncthreads();
      if (quad) {
        // 4 x 4 register patches: one 128-bit read of 4 input channels and one of 4 grad channels
        // feed 16 FMAs (the scalar path below needs a shared-memory read per 4 FMAs)
#pragma unroll
        for (int q = 0; q < CW_MAX_ITEMS / 4; ++q) {
          if (q < items4) {
            const int e = q * CW_THREADS + tid;
            if (e < ca * vpr) {
              const int cg = e / vpr, cv = e % vpr;
#pragma unroll 8
              for (int r = 0; r < CW_TR; ++r) {
                const float4 a4 = ((const float4*)a_s)[r * ca + cg];
                const float4 g = ((const float4*)g_s)[r * vpr + cv];
                float4& c0 = acc[4 * q + 0]; float4& c1 = acc[4 * q + 1];
                float4& c2 = acc[4 * q + 2]; float4& c3 = acc[4 * q + 3];
                c0.x = fmaf(a4.x, g.x, c0.x); c0.y = fmaf(a4.x, g.y, c0.y); c0.z = fmaf(a4.x, g.z, c0.z); c0.w = fmaf(a4.x, g.w, c0.w);
                c1.x = fmaf(a4.y, g.x, c1.x); c1.y = fmaf(a4.y, g.y, c1.y); c1.z = fmaf(a4.y, g.z, c1.z); c1.w = fmaf(a4.y, g.w, c1.w);
                c2.x = fmaf(a4.z, g.x, c2.x); c2.y = fmaf(a4.z, g.y, c2.y); c2.z = fmaf(a4.z, g.z, c2.z); c2.w = fmaf(a4.z, g.w, c2.w);
                c3.x = fmaf(a4.w, g.x, c3.x); c3.y = fmaf(a4.w, g.y, c3.y); c3.z = fmaf(a4.w, g.z, c3.z); c3.w = fmaf(a4.w, g.w, c3.w);
              }
            }
          }
        }
      } else {
#pragma unroll
      for (int q = 0; q < CW_MAX_ITEMS; ++q) {
        if (q < items) {
          const int e = q * CW_THREADS + tid;
          if (e < c_in * vpr) {
            const int ci = e / vpr, cv = e % vpr;
#pragma unroll 8
            for (int r = 0; r < CW_TR; ++r) {
              const float a = a_s[r * c_in + ci];
              const float4 g = ((const float4*)g_s)[r * vpr + cv];
              acc[q].x = fmaf(a, g.x, acc[q].x); acc[q].y = fmaf(a, g.y, acc[q].y);
              acc[q].z = fmaf(a, g.z, acc[q].z); acc[q].w = fmaf(a, g.w, acc[q].w);
            }
          }
        }
      }
      }
      __syncthreads();
    }
  }
  if (quad) {
#pragma unroll
    for (int q = 0; q < CW_MAX_ITEMS / 4; ++q) {
      if (q < items4) {
        const int e = q * CW_THREADS + threadIdx.x;
        if (e < ca * vpr) {
          const int cg = e / vpr, cv = (e % vpr) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            lk_red_add_v4(gw + ((int64_t)k * c_in + 4 * cg + i) * c_out + cv, acc[4 * q + i]);
        }
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < CW_MAX_ITEMS; ++q) {
      if (q < items) {
        int e = q * CW_THREADS + threadIdx.x;
        if (e < c_in * vpr) {
          int ci = e / vpr, cv = (e % vpr) * 4;
          lk_red_add_v4(gw + ((int64_t)k * c_in + ci) * c_out + cv, acc[q]);
        }
      }
    }
  }
}

extern "C" int lk_conv_bwd_weight(const float* d_in, const float* d_gout, const int32_t* d_nbr,
                                  int64_t n_out, int k, int c_in, int c_out, float* d_gw,
                                  lk_stream_t s) {
  LK_REQUIRE(n_out >= 0 && k > 0 && c_in > 0 && c_out > 0, "lk_conv_bwd_weight: bad sizes");
  LK_REQUIRE(c_out % 4 == 0 && c_in * (c_out / 4) <= CW_THREADS * CW_MAX_ITEMS,
             "lk_conv_bwd_weight: c_out must be a multiple of 4 and c_in*c_out <= 16384");
  LK_REQUIRE(d_gw, "lk_conv_bwd_weight: null output");
  cudaStream_t st = (cudaStream_t)s;
  LK_CUDA(cudaMemsetAsync(d_gw, 0, (size_t)k * c_in * c_out * sizeof(float), st));
  lk_count_launch();
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in && d_gout && d_nbr, "lk_conv_bwd_weight: null pointer");
  int splits = (4 * LK_SM_COUNT + k - 1) / k;
  if (splits < 1) splits = 1;
  int64_t max_splits = (n_out + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  dim3 grid((unsigned)k, (unsigned)splits);
  const size_t smem = (size_t)CW_TR * (c_in + c_out) * sizeof(float);      // <= 32 KB
  conv_bwd_weight_kernel<<<grid, CW_THREADS, smem, st>>>(d_in, d_gout, d_nbr, n_out, c_in, c_out, d_gw);
  LK_LAUNCHED();
  return LK_OK;
}
