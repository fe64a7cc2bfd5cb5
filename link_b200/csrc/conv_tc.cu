// Sparse 3D convolution on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
// Weight-stationary, TMEM-resident formulation (replaces the reference's host loop of
// gather -> cuBLAS mm -> scatter per offset, convolution_cuda.cu:101-164):
//
//   * a persistent CTA owns up to T = 512/C_out consecutive 128-row output tiles; their fp32
//     accumulators [128 x C_out] live in TMEM for the whole kernel (all 512 columns = 256 KB/SM),
//     so partial sums never touch registers, shared memory or HBM;
//   * outer loop over the K kernel offsets: W[k] (tf32 hi/lo planes, K-major SWIZZLE_128B) is
//     staged in shared memory ONCE per offset and reused by all T tiles of the CTA;
//   * inner loop over the tiles: the 128 input rows nbr[k, tile rows] are gathered with coalesced
//     128-bit loads (missing neighbours -> zero rows), split into tf32 hi/lo planes and written in
//     the canonical K-major SWIZZLE_128B layout; one thread issues 3*C_in/8 tcgen05.mma
//     (M=128, N=C_out, K=8) accumulating into that tile's TMEM columns; tcgen05.commit -> mbarrier
//     releases the operand stage.  Two operand stages: the gather of step j+1 overlaps the MMAs
//     of step j.  (tile, offset) steps with no neighbour at all are skipped.
//   * epilogue: tcgen05.ld 32x32b (thread = output row) -> + bias -> one streaming store per row.
//
// No atomics, no temporaries, deterministic, output written exactly once.
#include "common.cuh"
#include "tc.cuh"

#define CT_ROWS 128
#define CT_THREADS 256

template <int CIN, int COUT>
struct ConvTcCfg {
  static constexpr int KB = CIN / 32;                       // 128-byte K-blocks per row
  static constexpr uint32_t A_BLK = CT_ROWS * 128;          // one [128 x 32] K-block
  static constexpr uint32_t B_BLK = COUT * 128;             // one [COUT x 32] K-block
  static constexpr uint32_t A_PLANE = KB * A_BLK;           // hi (or lo) plane of one stage
  static constexpr uint32_t B_PLANE = KB * B_BLK;
  static constexpr uint32_t A_STAGE = 2 * A_PLANE;          // hi + lo
  static constexpr uint32_t B_STAGE = 2 * B_PLANE;
  static constexpr uint32_t SMEM = 2 * A_STAGE + 2 * B_STAGE + 1024;   // + alignment slack
  static constexpr int MAX_TILES = 512 / COUT;
  static constexpr int ITEMS_A = CT_ROWS * KB * 8 / CT_THREADS;        // float4 per thread per step
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(CT_THREADS, 1) conv_tc_kernel(
    const float* __restrict__ in, const float* __restrict__ wt /*[K][COUT][CIN]*/,
    const int* __restrict__ nbr, int64_t n_out, int K, int tiles_per_cta,
    const float* __restrict__ bias, float* __restrict__ out) {
  using Cfg = ConvTcCfg<CIN, COUT>;
  constexpr int KB = Cfg::KB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_stage[2] = {smem, smem + Cfg::A_STAGE};
  uint8_t* b_stage[2] = {smem + 2 * Cfg::A_STAGE, smem + 2 * Cfg::A_STAGE + Cfg::B_STAGE};
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t total_tiles = (n_out + CT_ROWS - 1) / CT_ROWS;
  const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int ntiles = (int)min((int64_t)tiles_per_cta, total_tiles - tile0);
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(tiles_per_cta * COUT)) ncols <<= 1;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    tc::mbar_init(&mma_bar[0], 1);
    tc::mbar_init(&mma_bar[1], 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = tc::idesc_tf32(128, COUT);

  uint32_t commits[2] = {0, 0};      // CTA-uniform: commits issued so far on each operand stage
  uint32_t touched = 0;              // CTA-uniform bitmask: tiles that received at least one MMA
  uint32_t step = 0;                 // CTA-uniform: operand stage = step & 1

  // row / chunk owned by this thread for item i of a step (8 lanes = one 128-byte K-block of a row)
  // item t = tid + i*256: chunk = t & 7, kb = (t >> 3) % KB, row = t / (8*KB)
  for (int k = 0; k < K; ++k) {
    // ---- stage W[k]: wait until every MMA that may still read this B buffer has completed ----
    if (commits[0]) tc::mbar_wait(&mma_bar[0], (commits[0] - 1) & 1);
    if (commits[1]) tc::mbar_wait(&mma_bar[1], (commits[1] - 1) & 1);
    {
      uint8_t* bh = b_stage[k & 1];
      uint8_t* bl = bh + Cfg::B_PLANE;
      const float* wk = wt + (int64_t)k * COUT * CIN;
      for (int t = tid; t < COUT * KB * 8; t += CT_THREADS) {
        int chunk = t & 7, kb = (t >> 3) % KB, row = t / (8 * KB);
        float4 v = __ldg((const float4*)(wk + row * CIN + kb * 32 + chunk * 4)), hi, lo;
        tc::split_tf32(v, hi, lo);
        uint32_t off = kb * Cfg::B_BLK + tc::sw128_offset(row, chunk);
        *(float4*)(bh + off) = hi;
        *(float4*)(bl + off) = lo;
      }
    }
    for (int t = 0; t < ntiles; ++t) {
      const int64_t row0 = (tile0 + t) * CT_ROWS;
      // ---- neighbour rows of this (tile, offset) ----
      int src[Cfg::ITEMS_A];
      int any = 0;
#pragma unroll
      for (int i = 0; i < Cfg::ITEMS_A; ++i) {
        int row = (tid + i * CT_THREADS) / (8 * KB);
        int64_t o = row0 + row;
        src[i] = (o < n_out) ? __ldg(nbr + (int64_t)k * n_out + o) : -1;
        any |= (src[i] >= 0);
      }
      if (!__syncthreads_or(any)) continue;          // nothing feeds this tile through offset k
      const int stage = step & 1;
      if (commits[stage]) tc::mbar_wait(&mma_bar[stage], (commits[stage] - 1) & 1);
      // ---- gather + tf32 split into the swizzled operand planes ----
      uint8_t* ah = a_stage[stage];
      uint8_t* al = ah + Cfg::A_PLANE;
      float4 v[Cfg::ITEMS_A];
#pragma unroll
      for (int i = 0; i < Cfg::ITEMS_A; ++i) {       // all loads first (memory-level parallelism)
        int tt = tid + i * CT_THREADS;
        int chunk = tt & 7, kb = (tt >> 3) % KB;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src[i] >= 0) v[i] = __ldg((const float4*)(in + (int64_t)src[i] * CIN + kb * 32 + chunk * 4));
      }
#pragma unroll
      for (int i = 0; i < Cfg::ITEMS_A; ++i) {
        int tt = tid + i * CT_THREADS;
        int chunk = tt & 7, kb = (tt >> 3) % KB, row = tt / (8 * KB);
        float4 hi, lo;
        tc::split_tf32(v[i], hi, lo);
        uint32_t off = kb * Cfg::A_BLK + tc::sw128_offset(row, chunk);
        *(float4*)(ah + off) = hi;
        *(float4*)(al + off) = lo;
      }
      tc::fence_proxy_async();
      __syncthreads();
      // ---- MMA issue: one thread ----
      if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)(t * COUT);
        const uint32_t ah_u = tc::smem_u32(ah), al_u = tc::smem_u32(al);
        const uint32_t bh_u = tc::smem_u32(b_stage[k & 1]), bl_u = bh_u + Cfg::B_PLANE;
        uint32_t acc = (touched >> t) & 1u;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t ao = kb * Cfg::A_BLK + ks * 32, bo = kb * Cfg::B_BLK + ks * 32;
            uint64_t dah = tc::smem_desc_sw128(ah_u + ao), dal = tc::smem_desc_sw128(al_u + ao);
            uint64_t dbh = tc::smem_desc_sw128(bh_u + bo), dbl = tc::smem_desc_sw128(bl_u + bo);
            tc::mma_tf32(d, dal, dbh, idesc, acc);
            tc::mma_tf32(d, dah, dbl, idesc, 1);
            tc::mma_tf32(d, dah, dbh, idesc, 1);
            acc = 1;
          }
        }
        tc::mma_commit(&mma_bar[stage]);
      }
      touched |= 1u << t;
      commits[stage]++;
      step++;
    }
  }
  // ---- drain, then epilogue ----
  if (commits[0]) tc::mbar_wait(&mma_bar[0], (commits[0] - 1) & 1);
  if (commits[1]) tc::mbar_wait(&mma_bar[1], (commits[1] - 1) & 1);
  tc::fence_after_sync();
  {
    constexpr int HALF = COUT / 2;                   // warps 0-3: columns [0,HALF), warps 4-7: rest
    const int q = warp & 3, c_base = (warp >> 2) * HALF;
    for (int t = 0; t < ntiles; ++t) {
      const int64_t o = (tile0 + t) * CT_ROWS + q * 32 + lane;
      const bool live = (touched >> t) & 1u;
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 16) {
        float acc[16];
        if (live) {
          tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * COUT + c_base + c0), acc);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        }
        if (o < n_out) {
          float* dst = out + o * COUT + c_base + c0;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            float4 b = bias ? __ldg((const float4*)(bias + c_base + c0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            lk_stg_stream((float4*)(dst + e),
                          make_float4(acc[e] + b.x, acc[e + 1] + b.y, acc[e + 2] + b.z, acc[e + 3] + b.w));
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, ncols);
}

template <int CIN, int COUT>
static int launch_conv_tc(const float* in, const float* wt, const int* nbr, int64_t n_out, int k,
                          const float* bias, float* out, cudaStream_t st) {
  using Cfg = ConvTcCfg<CIN, COUT>;
  static bool attr_set = false;
  if (!attr_set) {
    LK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg::SMEM));
    attr_set = true;
  }
  int64_t tiles = (n_out + CT_ROWS - 1) / CT_ROWS;
  int64_t tpc = (tiles + LK_SM_COUNT - 1) / LK_SM_COUNT;     // balance over the 148 SMs ...
  if (tpc > Cfg::MAX_TILES) tpc = Cfg::MAX_TILES;            // ... within the 512 TMEM columns
  if (tpc < 1) tpc = 1;
  int grid = (int)((tiles + tpc - 1) / tpc);
  conv_tc_kernel<CIN, COUT><<<grid, CT_THREADS, Cfg::SMEM, st>>>(in, wt, nbr, n_out, k, (int)tpc, bias, out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_conv_tc_supported(int c_in, int c_out) {
  return (c_in == 32 || c_in == 64) && (c_out == 32 || c_out == 64);
}

extern "C" int lk_conv_tc_fwd(const float* d_in, const float* d_wt, const int32_t* d_nbr,
                              int64_t n_out, int k, int c_in, int c_out, const float* d_bias,
                              float* d_out, lk_stream_t s) {
  LK_REQUIRE(n_out >= 0 && k > 0, "lk_conv_tc_fwd: bad sizes");
  LK_REQUIRE(lk_conv_tc_supported(c_in, c_out), "lk_conv_tc_fwd: channels must be 32 or 64");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in && d_wt && d_nbr && d_out, "lk_conv_tc_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_in % 16 == 0 && (uintptr_t)d_wt % 16 == 0 && (uintptr_t)d_out % 16 == 0,
             "lk_conv_tc_fwd: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  if (c_in == 32 && c_out == 32) return launch_conv_tc<32, 32>(d_in, d_wt, d_nbr, n_out, k, d_bias, d_out, st);
  if (c_in == 32 && c_out == 64) return launch_conv_tc<32, 64>(d_in, d_wt, d_nbr, n_out, k, d_bias, d_out, st);
  if (c_in == 64 && c_out == 32) return launch_conv_tc<64, 32>(d_in, d_wt, d_nbr, n_out, k, d_bias, d_out, st);
  return launch_conv_tc<64, 64>(d_in, d_wt, d_nbr, n_out, k, d_bias, d_out, st);
}
