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
//     releases the operand stage;
//   * warp-specialised: 8 producer warps (gather/split/store, then the epilogue) and 1 MMA-issuer
//     warp talk only through full/empty mbarriers (no __syncthreads in the main loop); two operand
//     stages in shared memory plus a register stage (rows of step j+1 and indices of step j+2
//     are in flight while step j is being written);
//   * epilogue: tcgen05.ld 32x32b (thread = output row) -> + bias -> one streaming store per row.
//
// No atomics, no temporaries, deterministic, output written exactly once.
#include "common.cuh"
#include "tc.cuh"

#define CT_ROWS 128
#define CT_THREADS 512      // producer threads (16 warps); one more warp issues the MMAs

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
__global__ void __launch_bounds__(CT_THREADS + 32, 1) conv_tc_kernel(
    const float* __restrict__ in, const float* __restrict__ wt /*[K][COUT][CIN]*/,
    const int* __restrict__ nbr, int64_t n_out, int K, int tiles_per_cta,
    lk_conv_epilogue_t ep, float* __restrict__ out) {
  using Cfg = ConvTcCfg<CIN, COUT>;
  constexpr int KB = Cfg::KB;
  constexpr int NI = Cfg::ITEMS_A;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* const a_base = smem;                          // 2 A stages, then 2 B stages
  uint8_t* const b_base = smem + 2 * Cfg::A_STAGE;
  __shared__ uint64_t full_bar[2];    // operand stage written   (one arrival per producer warp)
  __shared__ uint64_t empty_bar[2];   // operand stage consumed  (tcgen05.commit)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t total_tiles = (n_out + CT_ROWS - 1) / CT_ROWS;
  const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int ntiles = (int)min((int64_t)tiles_per_cta, total_tiles - tile0);
  const int total_steps = K * ntiles;            // step j = (offset k = j / ntiles, tile t = j % ntiles)
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(tiles_per_cta * COUT)) ncols <<= 1;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    tc::mbar_init(&full_bar[0], CT_THREADS / 32);      // one arrival per producer warp
    tc::mbar_init(&full_bar[1], CT_THREADS / 32);
    tc::mbar_init(&empty_bar[0], 1);
    tc::mbar_init(&empty_bar[1], 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == CT_THREADS / 32) {
    // ================= MMA issuer warp =================
    const uint32_t idesc = tc::idesc_tf32(128, COUT);
    // descriptors of stage 0; the start-address field counts 16-byte units, so other stages,
    // planes and k-slices are plain additions
    const uint64_t da0 = tc::smem_desc_sw128(tc::smem_u32(a_base));
    const uint64_t db0 = tc::smem_desc_sw128(tc::smem_u32(b_base));
    int k = 0, t = 0;
    for (int j = 0; j < total_steps; ++j) {
      const int stage = j & 1;
      tc::mbar_wait(&full_bar[stage], (uint32_t)(j >> 1) & 1u);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(t * COUT);
        const uint64_t da_hi = da0 + (uint64_t)((stage * Cfg::A_STAGE) >> 4);
        const uint64_t da_lo = da_hi + (uint64_t)(Cfg::A_PLANE >> 4);
        const uint64_t db_hi = db0 + (uint64_t)(((k & 1) * Cfg::B_STAGE) >> 4);
        const uint64_t db_lo = db_hi + (uint64_t)(Cfg::B_PLANE >> 4);
        uint32_t acc = k > 0 ? 1u : 0u;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // descriptor start-address field is in 16-byte units: advance inside the swizzle atom
            const uint64_t ao = (uint64_t)((kb * Cfg::A_BLK + ks * 32) >> 4);
            const uint64_t bo = (uint64_t)((kb * Cfg::B_BLK + ks * 32) >> 4);
            tc::mma_tf32(d, da_lo + ao, db_hi + bo, idesc, acc);
            tc::mma_tf32(d, da_hi + ao, db_lo + bo, idesc, 1);
            tc::mma_tf32(d, da_hi + ao, db_hi + bo, idesc, 1);
            acc = 1;
          }
        }
        tc::mma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++t == ntiles) { t = 0; ++k; }
    }
  } else {
    // ================= producer warps (gather + tf32 split), later the epilogue =================
    // Thread -> (row, quarter): TPR = CT_THREADS/128 threads share one gathered row; thread q of a
    // row owns the 16-byte chunks {q, q + TPR, q + 2 TPR, ...} of that row, so the TPR lanes of a
    // row read TPR*16 contiguous bytes per load instruction (full sectors) and ONE neighbour index
    // / validity test / dirty bit per thread and step covers all of its NI chunks.
    constexpr int TPR = CT_THREADS / CT_ROWS;              // 4
    static_assert(NI * TPR == KB * 8, "chunks of a row must tile over its threads");
    const int prow = tid / TPR, pq = tid % TPR;
    uint32_t soff[NI];
    int icol[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int cch = pq + i * TPR;                        // chunk index within the row (0 .. 8*KB)
      const int kb = cch >> 3, chunk = cch & 7;
      icol[i] = cch * 4;
      soff[i] = kb * Cfg::A_BLK + tc::sw128_offset(prow, chunk);
    }
    const int64_t row_base = tile0 * CT_ROWS;
    // neighbour index of my row for step (k2, t2); no integer division in the steady state
    auto load_idx = [&](int k2, int t2) -> int {
      const int64_t o = row_base + (int64_t)t2 * CT_ROWS + prow;
      return (k2 < K && o < n_out) ? __ldg(nbr + (int64_t)k2 * n_out + o) : -1;
    };
    auto load_rows = [&](int src, float4* v) {
      if (src >= 0) {
        const float* rp = in + (int64_t)src * CIN;
#pragma unroll
        for (int i = 0; i < NI; ++i) v[i] = __ldg((const float4*)(rp + icol[i]));
      }
    };
    float4 v[NI], v_next[NI];
    // bit `stage`: my slots of that operand stage currently hold non-zero data.  ~70 % of the
    // gathered rows are missing neighbours (zero rows): slots that are already zero are not
    // rewritten, which halves the split + store work.
    uint32_t dirty = 3u;                      // shared memory starts uninitialised
    int k = 0, t = 0;                         // step j
    int kc = 0, tc2 = 0;                      // step j + 2
    auto advance = [&](int& kk, int& tt) { if (++tt == ntiles) { tt = 0; ++kk; } };
    int src_a = load_idx(0, 0);
    load_rows(src_a, v);
    advance(kc, tc2);
    int src_b = load_idx(kc, tc2);
    advance(kc, tc2);
    for (int j = 0; j < total_steps; ++j) {
      const int stage = j & 1;
      const int src_c = load_idx(kc, tc2);    // index prefetch distance 2
      load_rows(src_b, v_next);               // row prefetch distance 1 (in flight during the stores)
      // split BEFORE waiting for the stage: after the wake-up only the stores remain on the
      // critical path  commit(j-2) -> stores -> full(j) -> MMA(j)
      float4 hi[NI], lo[NI];
      if (src_a >= 0) {
#pragma unroll
        for (int i = 0; i < NI; ++i) tc::split_tf32(v[i], hi[i], lo[i]);
      }
      if (j >= 2) tc::mbar_wait(&empty_bar[stage], (uint32_t)((j >> 1) - 1) & 1u);
      if (t == 0) {
        // stage W[k]: every MMA of offset k-2 (last reader of this buffer) precedes step j-2's
        // commit, which the wait above has just observed
        uint8_t* bh = b_base + (k & 1) * Cfg::B_STAGE;
        uint8_t* bl = bh + Cfg::B_PLANE;
        const float* wk = wt + (int64_t)k * COUT * CIN;
        for (int tt = tid; tt < COUT * KB * 8; tt += CT_THREADS) {
          int chunk = tt & 7, kb = (tt >> 3) % KB, row = tt / (8 * KB);
          float4 w4 = __ldg((const float4*)(wk + row * CIN + kb * 32 + chunk * 4)), whi, wlo;
          tc::split_tf32(w4, whi, wlo);
          uint32_t off = kb * Cfg::B_BLK + tc::sw128_offset(row, chunk);
          *(float4*)(bh + off) = whi;
          *(float4*)(bl + off) = wlo;
        }
      }
      uint8_t* ah = a_base + stage * Cfg::A_STAGE;
      uint8_t* al = ah + Cfg::A_PLANE;
      const uint32_t bit = 1u << stage;
      if (src_a >= 0) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          *(float4*)(ah + soff[i]) = hi[i];
          *(float4*)(al + soff[i]) = lo[i];
        }
        dirty |= bit;
      } else if (dirty & bit) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          *(float4*)(ah + soff[i]) = z;
          *(float4*)(al + soff[i]) = z;
        }
        dirty &= ~bit;
      }
      tc::fence_proxy_async();                // my generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[stage]);   // 16 arrivals instead of 512 serialised ones
#pragma unroll
      for (int i = 0; i < NI; ++i) v[i] = v_next[i];
      src_a = src_b;
      src_b = src_c;
      advance(k, t);
      advance(kc, tc2);
    }
    // ---- drain: the last commit on each stage ----
    const int c0 = (total_steps + 1) >> 1, c1 = total_steps >> 1;
    if (c0) tc::mbar_wait(&empty_bar[0], (uint32_t)(c0 - 1) & 1u);
    if (c1) tc::mbar_wait(&empty_bar[1], (uint32_t)(c1 - 1) & 1u);
    tc::fence_after_sync();
    // ---- epilogue: thread = output row (TMEM lane quarter q = warp & 3), 16 columns per warp ----
    constexpr int NSLICE = COUT / 16;
    const int q = warp & 3, slice = warp >> 2;
    if (slice < NSLICE) {
      const int c_base = slice * 16;
      for (int t = 0; t < ntiles; ++t) {
        const int64_t o = (tile0 + t) * CT_ROWS + q * 32 + lane;
        float acc[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * COUT + c_base), acc);
        if (o < n_out) {
          float* dst = out + o * COUT + c_base;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            float4 y = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
            if (ep.d_scale) {
              float4 sc = __ldg((const float4*)(ep.d_scale + c_base + e));
              y.x *= sc.x; y.y *= sc.y; y.z *= sc.z; y.w *= sc.w;
            }
            if (ep.d_shift) {
              float4 sh = __ldg((const float4*)(ep.d_shift + c_base + e));
              y.x += sh.x; y.y += sh.y; y.z += sh.z; y.w += sh.w;
            }
            if (ep.d_residual) {
              float4 rr = lk_ldg_stream((const float4*)(ep.d_residual + o * COUT + c_base + e));
              y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
            }
            if (ep.relu) {
              y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
            }
            lk_stg_stream((float4*)(dst + e), y);
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
                          const lk_conv_epilogue_t& ep, float* out, cudaStream_t st) {
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
  conv_tc_kernel<CIN, COUT><<<grid, CT_THREADS + 32, Cfg::SMEM, st>>>(in, wt, nbr, n_out, k, (int)tpc, ep, out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_conv_tc_supported(int c_in, int c_out) {
  return (c_in == 32 || c_in == 64) && (c_out == 32 || c_out == 64);
}

extern "C" int lk_conv_tc_fwd(const float* d_in, const float* d_wt, const int32_t* d_nbr,
                              int64_t n_out, int k, int c_in, int c_out, const float* d_bias,
                              float* d_out, lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, d_bias, nullptr, 0, 0};
  return lk_conv_tc_fwd_ex(d_in, d_wt, d_nbr, n_out, k, c_in, c_out, &ep, d_out, s);
}

extern "C" int lk_conv_tc_fwd_ex(const float* d_in, const float* d_wt, const int32_t* d_nbr,
                                 int64_t n_out, int k, int c_in, int c_out,
                                 const lk_conv_epilogue_t* epp, float* d_out, lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, nullptr, nullptr, 0, 0};
  if (epp) ep = *epp;
  LK_REQUIRE(n_out >= 0 && k > 0, "lk_conv_tc_fwd: bad sizes");
  LK_REQUIRE(lk_conv_tc_supported(c_in, c_out), "lk_conv_tc_fwd: channels must be 32 or 64");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in && d_wt && d_nbr && d_out, "lk_conv_tc_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_in % 16 == 0 && (uintptr_t)d_wt % 16 == 0 && (uintptr_t)d_out % 16 == 0,
             "lk_conv_tc_fwd: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  if (c_in == 32 && c_out == 32) return launch_conv_tc<32, 32>(d_in, d_wt, d_nbr, n_out, k, ep, d_out, st);
  if (c_in == 32 && c_out == 64) return launch_conv_tc<32, 64>(d_in, d_wt, d_nbr, n_out, k, ep, d_out, st);
  if (c_in == 64 && c_out == 32) return launch_conv_tc<64, 32>(d_in, d_wt, d_nbr, n_out, k, ep, d_out, st);
  return launch_conv_tc<64, 64>(d_in, d_wt, d_nbr, n_out, k, ep, d_out, st);
}
