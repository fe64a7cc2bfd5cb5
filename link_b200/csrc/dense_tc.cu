// pre_mix on the 5th-generation tensor cores: out = LN(x @ W^T) with tcgen05.mma kind::tf32
// (3xTF32 split, fp32-level accuracy), accumulators in TMEM, LayerNorm in the epilogue.
// This is the "dense 1x1 channel-mixing" GEMM of the LinK block (linkencoder.py:112-115): a
// genuine [N,C]x[C,C] GEMM, so it goes to tcgen05; everything index-/gather-shaped stays SIMT.
//
// One CTA (128 threads) owns 128-row tiles: rows are staged (coalesced 128-bit loads) into the
// canonical K-major SWIZZLE_128B shared layout as tf32 hi/lo planes, one thread issues
// 3 * C/8 tcgen05.mma (M=128, N=C, K=8) into a C-column TMEM tile, tcgen05.commit signals an
// mbarrier, and in the epilogue thread t reads row t (tcgen05.ld 32x32b) so the LayerNorm is
// thread-local (no shuffles).  W (hi/lo) is loaded once per CTA; CTAs are persistent over tiles,
// two per SM so that one CTA's loads overlap the other's MMA + epilogue.
#include "common.cuh"
#include "tc.cuh"

#define DT_ROWS 128
#define DT_THREADS 128

template <int C>
__global__ void __launch_bounds__(DT_THREADS) linear_ln_tc_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int64_t n, float* __restrict__ out) {
  lk_pdl_enter();
  constexpr int KB = C / 32;                         // 128-byte K-blocks per row
  constexpr uint32_t A_BLK = DT_ROWS * 128;          // bytes of one [128 x 32] K-block
  constexpr uint32_t B_BLK = C * 128;                // bytes of one [C x 32] K-block
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles must start on a 1024-byte boundary (the launch adds 1 KB of slack)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + KB * A_BLK;
  uint8_t* b_hi = a_lo + KB * A_BLK;
  uint8_t* b_lo = b_hi + KB * B_BLK;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, C < 32 ? 32 : C);
  if (tid == 0) { tc::mbar_init(&mma_bar, 1); tc::fence_mbar_init(); }
  // W [C out][C in] is already "N rows x K contiguous" = K-major B.  8 lanes cover one 128-byte
  // K-block of one row per load instruction (fully coalesced).
  {
    constexpr int NW = C * KB * 8 / DT_THREADS;      // all weight loads in flight before the first use
    float4 wv[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      int t = tid + i * DT_THREADS;
      int chunk = t & 7, kb = (t >> 3) % KB, row = t / (8 * KB);
      wv[i] = __ldg((const float4*)(w + (int64_t)row * C + kb * 32 + chunk * 4));
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      int t = tid + i * DT_THREADS;
      int chunk = t & 7, kb = (t >> 3) % KB, row = t / (8 * KB);
      float4 hi, lo;
      tc::split_tf32(wv[i], hi, lo);
      uint32_t off = kb * B_BLK + tc::sw128_offset(row, chunk);
      *(float4*)(b_hi + off) = hi;
      *(float4*)(b_lo + off) = lo;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = tc::idesc_tf32(128, C);
  const uint32_t a_hi_u = tc::smem_u32(a_hi), a_lo_u = tc::smem_u32(a_lo);
  const uint32_t b_hi_u = tc::smem_u32(b_hi), b_lo_u = tc::smem_u32(b_lo);
  uint32_t parity = 0;

  const int64_t tiles = (n + DT_ROWS - 1) / DT_ROWS;
  // Software pipeline: the global loads of tile i+1 are issued right after tile i's MMAs and stay
  // in flight (registers) during its MMA wait and LayerNorm epilogue, so a tile costs
  // max(load latency, MMA + epilogue) instead of their sum.
  constexpr int NLD = DT_ROWS * KB * 8 / DT_THREADS;
  float4 ld[NLD];
  auto load_tile = [&](int64_t tile) {
    const int64_t row0 = tile * DT_ROWS;
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      int t = tid + i * DT_THREADS;
      int chunk = t & 7, kb = (t >> 3) % KB, row = t / (8 * KB);
      ld[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tile < tiles && row0 + row < n)
        ld[i] = lk_ldg_stream((const float4*)(x + (row0 + row) * C + kb * 32 + chunk * 4));
    }
  };
  load_tile(blockIdx.x);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * DT_ROWS;
    // ---- stage A (hi/lo): split + swizzled stores of the rows loaded one iteration ago ----
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      int t = tid + i * DT_THREADS;
      int chunk = t & 7, kb = (t >> 3) % KB, row = t / (8 * KB);
      float4 hi, lo;
      tc::split_tf32(ld[i], hi, lo);
      uint32_t off = kb * A_BLK + tc::sw128_offset(row, chunk);
      *(float4*)(a_hi + off) = hi;
      *(float4*)(a_lo + off) = lo;
    }
    tc::fence_proxy_async();
    __syncthreads();
    // ---- MMA: one thread, 3 * C/8 instructions ----
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      uint32_t acc = 0;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t ao = kb * A_BLK + ks * 32, bo = kb * B_BLK + ks * 32;
          uint64_t ah = tc::smem_desc_sw128(a_hi_u + ao), al = tc::smem_desc_sw128(a_lo_u + ao);
          uint64_t bh = tc::smem_desc_sw128(b_hi_u + bo), bl = tc::smem_desc_sw128(b_lo_u + bo);
          tc::mma_tf32(tmem_base, al, bh, idesc, acc);   // small terms first
          tc::mma_tf32(tmem_base, ah, bl, idesc, 1);
          tc::mma_tf32(tmem_base, ah, bh, idesc, 1);
          acc = 1;
        }
      }
      tc::mma_commit(&mma_bar);
    }
    load_tile(tile + gridDim.x);          // next tile's rows: in flight during the wait + epilogue
    tc::mbar_wait(&mma_bar, parity);
    parity ^= 1;
    tc::fence_after_sync();
    // ---- epilogue: thread t owns row t ----
    float v[C];
#pragma unroll
    for (int c0 = 0; c0 < C; c0 += 16)
      tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v + c0);
    // LayerNorm with 8 independent partial sums (a 64-term dependent chain costs 64 x 4 cycles on a
    // kernel that runs only 2 warps per scheduler)
    float ps[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ps[k] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) ps[c & 7] += v[c];
    const float s = ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
    const float mean = s / (float)C;
#pragma unroll
    for (int k = 0; k < 8; ++k) ps[k] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { v[c] -= mean; ps[c & 7] = fmaf(v[c], v[c], ps[c & 7]); }
    const float q = ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
    float rstd = 1.0f / sqrtf(q / (float)C + eps);
    int64_t r = row0 + tid;
    if (r < n) {
      float* dst = out + r * C;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        float4 g = __ldg((const float4*)(gamma + c)), b = __ldg((const float4*)(beta + c));
        lk_stg_stream((float4*)(dst + c), make_float4(v[c] * rstd * g.x + b.x, v[c + 1] * rstd * g.y + b.y,
                                                      v[c + 2] * rstd * g.z + b.z, v[c + 3] * rstd * g.w + b.w));
      }
    }
    tc::fence_before_sync();
    __syncthreads();     // TMEM tile and the A planes may be overwritten by the next tile
    tc::fence_after_sync();
  }
  if (warp == 0) tc::tmem_dealloc(tmem_base, C < 32 ? 32 : C);
}

template <int C>
static int launch_tc(const float* x, const float* w, const float* g, const float* b, float eps,
                     int64_t n, float* out, cudaStream_t st) {
  size_t smem = (size_t)2 * (C / 32) * DT_ROWS * 128 + (size_t)2 * (C / 32) * C * 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    LK_CUDA(cudaFuncSetAttribute(linear_ln_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int64_t tiles = (n + DT_ROWS - 1) / DT_ROWS;
  int ctas_per_sm = (smem <= 100 * 1024) ? 2 : 1;
  int grid = (int)(tiles < (int64_t)LK_SM_COUNT * ctas_per_sm ? tiles : (int64_t)LK_SM_COUNT * ctas_per_sm);
  LK_PDL_LAUNCH(linear_ln_tc_kernel<C>, grid, DT_THREADS, smem, st, x, w, g, b, eps, n, out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_linear_ln_tc_fwd(const float* d_x, const float* d_w, const float* d_gamma,
                                   const float* d_beta, float eps, int64_t n, int c, float* d_out,
                                   lk_stream_t s) {
  // C = 128 would need 256 KB of operand tiles (A and W, hi + lo); it stays on the FFMA kernel
  LK_REQUIRE(n >= 0 && (c == 32 || c == 64), "lk_linear_ln_tc_fwd: C must be 32 or 64");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_x && d_w && d_gamma && d_beta && d_out, "lk_linear_ln_tc_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_x % 16 == 0 && (uintptr_t)d_out % 16 == 0 && (uintptr_t)d_w % 16 == 0,
             "lk_linear_ln_tc_fwd: alignment");
  cudaStream_t st = (cudaStream_t)s;
  if (c == 32) return launch_tc<32>(d_x, d_w, d_gamma, d_beta, eps, n, d_out, st);
  return launch_tc<64>(d_x, d_w, d_gamma, d_beta, eps, n, d_out, st);
}
