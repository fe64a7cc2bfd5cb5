// Weight gradient of the sparse convolution on the 5th-generation tensor cores (tcgen05, kind::tf32,
// 3xTF32 split):   grad_w[k] = sum over the pairs (i -> o) of offset k of  in[i]^T (x) grad_out[o]
// (reference: the per-offset  gather -> torch::mm(in^T, grad_out)  loop of
// convolution_backward_cuda, convolution_cuda.cu:217-278).
//
// Formulation.  Per offset k this is a genuine GEMM  D_k [C_in x C_out] = A_k^T [C_in x P] . G [P x C_out]
// whose contraction index is the PAIR.  Both operands arrive as rows of channels (a gathered input
// row, a grad_out row), i.e. with the M / N index contiguous and the contraction index strided:
// "MN-major" operands.  tcgen05 reads those straight from shared memory (instruction-descriptor
// bits 15 / 16; for 32-bit elements the canonical layout is SWIZZLE_128B with a 32-byte swizzle
// base: 4 pair-rows of 128 bytes form one 512-byte atom), so no transposition pass exists anywhere:
//
//   * the relation is walked OUTPUT-stationary in tiles of 64 rows (in the order of the forward
//     tile-skipping plan when there is one): the grad_out tile is staged ONCE per tile (tf32 hi / lo
//     planes, double buffered) and reused by every offset, only the input rows are gathered per
//     offset; (offset, tile) steps without a single pair are skipped (per-tile offset masks from a
//     pre-pass over the kernel map, shared by all convs on the map);
//   * the M = 128 rows of one MMA hold 128 / C_in offsets side by side ("slot"): A = [in_k1 | in_k2 ..]^T,
//     so C_in = 32 / 64 use the full tensor-core tile; accumulators [128 x C_out] per slot stay in
//     TMEM (up to 512 / C_out slots) for the whole kernel, each CTA owns one group of offsets and
//     an interleaved share of the tiles and flushes its partial sums once with vector reductions;
//   * 16 producer warps fetch rows with coalesced 128-bit loads (indices two work items ahead, rows
//     one item ahead, in registers), split them into tf32 hi / lo and store both planes; one elected
//     lane of a 17th warp issues 3 x 8 tcgen05.mma (M=128, N=C_out, K=8 pairs) per step;
//     tcgen05.commit -> mbarriers release the A stage / the grad_out buffer; no __syncthreads in
//     the main loop.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

#define WG_ROWS 64              // pairs (relation rows) per tile = per step
#define WG_THREADS 512          // producer threads; one more warp issues the MMAs
#define WG_CHUNK_TILES 256      // tiles whose work items are listed at once
#define WG_MAX_SLOTS 16
#define WG_ITEM_CAP (WG_CHUNK_TILES * (1 + WG_MAX_SLOTS))
#define WG_A_PLANE 32768u       // 64 pairs x 128 fp32
#define WG_A_STAGE 65536u       // hi + lo
#define WG_G_REGION 65536u      // grad_out buffers (2 x (hi + lo) at C_out <= 64, 1 at 128)

// Shared-memory layout of an MN-major operand plane [64 pairs][chunks16 16-byte chunks]: the canonical
// SWIZZLE_128B_BASE32B form for 32-bit MN-major operands (descriptor layout type 1): 32-channel
// blocks 8 KB apart (LBO), 4-pair groups 512 B apart (SBO), pair rows of 128 B, the four 32-byte
// units of a row XOR-swizzled by the row's index in its 512-byte atom.  (Probed on B200 with
// scripts/wgrad_probe.py at the time: SWIZZLE_128B with 16-byte units and the unswizzled
// core-matrix form both return garbage for MN-major tf32; this one is exact.)
__device__ __forceinline__ uint32_t wg_off(int m16, int p) {
  const int mb = m16 >> 3, j = m16 & 7;
  return (uint32_t)(mb * 8192 + (p >> 2) * 512 + (p & 3) * 128 + ((((j >> 1) ^ (p & 3))) << 5) + (j & 1) * 16);
}
__device__ __forceinline__ uint64_t wg_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;       // LBO: next 32-channel block
  d |= (uint64_t)(512 >> 4) << 32;        // SBO: next group of 4 pairs
  d |= (uint64_t)1 << 46;                 // descriptor version 1
  d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
  return d;
}
#define WG_KSTEP 1024u                    // bytes per MMA k-step (8 pairs = two 4-pair groups)

// One warp per 64-row tile: offsets with at least one pair -> masks[tile]; with a plan order
// (perm) the relation is also copied in that order (nbrp[k][pos] = nbr[k][perm[pos]]) so that the
// main kernel's index chain is one load deep.
__global__ void __launch_bounds__(256) wgrad_prepass_kernel(const int* __restrict__ nbr, const int* __restrict__ perm,
                                                            int64_t n, int K, int* __restrict__ nbrp,
                                                            unsigned* __restrict__ masks) {
  const int lane = threadIdx.x & 31;
  const int64_t tiles = (n + WG_ROWS - 1) / WG_ROWS;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t tile = warp0; tile < tiles; tile += nwarps) {
    const int64_t p0 = tile * WG_ROWS + lane, p1 = p0 + 32;
    const int o0 = p0 < n ? (perm ? __ldg(perm + p0) : (int)p0) : -1;
    const int o1 = p1 < n ? (perm ? __ldg(perm + p1) : (int)p1) : -1;
    unsigned mask = 0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const int v0 = o0 >= 0 ? __ldg(nbr + (int64_t)k * n + o0) : -1;
      const int v1 = o1 >= 0 ? __ldg(nbr + (int64_t)k * n + o1) : -1;
      if (nbrp) {
        if (p0 < n) nbrp[(int64_t)k * n + p0] = v0;
        if (p1 < n) nbrp[(int64_t)k * n + p1] = v1;
      }
      if (__ballot_sync(0xffffffffu, v0 >= 0 || v1 >= 0)) mask |= 1u << k;
    }
    if (lane == 0) masks[tile] = mask;
  }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(WG_THREADS + 32, 1) conv_wgrad_tc_kernel(
    const float* __restrict__ in, const float* __restrict__ gout,
    const int* __restrict__ nbrp /*[K][n] relation in tile order*/,
    const int* __restrict__ perm /*[n] tile position -> grad_out row, or NULL = identity*/,
    const unsigned* __restrict__ masks /*[tiles] offsets present per 64-row tile*/, int64_t n, int K,
    int opg /*offsets per group = per CTA*/, float* __restrict__ gw) {
  constexpr int OPS = 128 / CIN;                      // offsets per slot (side by side in M)
  constexpr int CHG = COUT / 4;                       // 16-byte chunks per grad_out row
  constexpr int NCH_G = WG_ROWS * CHG / WG_THREADS;   // chunks per thread of a grad_out tile (COUT / 32)
  constexpr int RPP_G = WG_THREADS / CHG;             // grad_out rows per fetch pass
  constexpr int NG = COUT <= 64 ? 2 : 1;              // grad_out buffers
  constexpr uint32_t G_PLANE = WG_ROWS * COUT * 4;
  constexpr uint32_t G_BUF = 2 * G_PLANE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* const a_base = smem;                       // 2 stages x (hi, lo)
  uint8_t* const g_base = smem + 2 * WG_A_STAGE;      // NG buffers x (hi, lo)
  __shared__ uint64_t full_bar[2], empty_bar[2], gfull_bar[2], gempty_bar[2], done_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t items_s[WG_ITEM_CAP];           // tile << 5 | (31 = grad_out tile, else slot)
  __shared__ int wsum_s[WG_CHUNK_TILES / 32];
  __shared__ int nitems_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (n + WG_ROWS - 1) / WG_ROWS;
  const int k0 = blockIdx.y * opg;
  const int kend = min(K, k0 + opg);
  const int nslots = (kend - k0 + OPS - 1) / OPS;
  const uint32_t gmask = (kend - k0) >= 32 ? 0xFFFFFFFFu : ((1u << (kend - k0)) - 1u);
  const int splits = gridDim.x;
  const int64_t nt = tiles > blockIdx.x ? (tiles - blockIdx.x + splits - 1) / splits : 0;   // my tiles

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&full_bar[b], WG_THREADS / 32);
      tc::mbar_init(&empty_bar[b], 1);
      tc::mbar_init(&gfull_bar[b], WG_THREADS / 32);
      tc::mbar_init(&gempty_bar[b], 1);
    }
    tc::mbar_init(&done_bar, 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  int jcount = 0, gcount = 0, cur_gb = 0;     // A steps / grad_out tiles so far: stage + phase bookkeeping
  uint32_t touched = 0;                       // slots that have received an MMA
  for (int64_t j0 = 0; j0 < nt; j0 += WG_CHUNK_TILES) {
    // ---- work items of the next <= 256 tiles: [grad_out tile, active slots ...] per tile ----
    int cnt = 0, inc = 0;
    uint32_t sm = 0, tile32 = 0;
    if (tid < WG_CHUNK_TILES) {
      const int64_t j = j0 + tid;
      if (j < nt) {
        const int64_t tile = blockIdx.x + j * splits;
        tile32 = (uint32_t)tile;
        const uint32_t m = (__ldg(masks + tile) >> k0) & gmask;
        for (int sl = 0; sl < nslots; ++sl)
          if ((m >> (sl * OPS)) & ((1u << OPS) - 1u)) sm |= 1u << sl;
        cnt = sm ? 1 + __popc(sm) : 0;
      }
      inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) wsum_s[warp] = inc;
    }
    __syncthreads();
    if (tid < WG_CHUNK_TILES) {
      int base = 0;
#pragma unroll
      for (int w = 0; w < WG_CHUNK_TILES / 32; ++w) base += (w < warp) ? wsum_s[w] : 0;
      int pos = base + inc - cnt;
      if (cnt) {
        items_s[pos++] = (tile32 << 5) | 31u;
        for (int sl = 0; sl < nslots; ++sl)
          if ((sm >> sl) & 1u) items_s[pos++] = (tile32 << 5) | (uint32_t)sl;
      }
      if (tid == WG_CHUNK_TILES - 1) nitems_s = base + inc;
    }
    __syncthreads();
    const int nitems = nitems_s;

    if (warp < WG_THREADS / 32) {
      // ================= producers =================
      // A step: warp w fetches pair rows w, w+16, w+32, w+48; its 32 lanes are the 32 16-byte chunks of
      // the M = 128 channel row (OPS offsets side by side).  grad_out tile: CHG lanes per row.
      const int j16 = lane;
      const int hA = (j16 * 4) / CIN, cA = (j16 * 4) % CIN;
      const int pA0 = warp;
      const int jG = tid % CHG, pG0 = tid / CHG;
      auto get_idx = [&](int it, int* src) {
#pragma unroll
        for (int i = 0; i < 4; ++i) src[i] = -1;
        if (it < nitems) {
          const uint32_t e = items_s[it];
          const int code = (int)(e & 31u);
          const int64_t pos0 = (int64_t)(e >> 5) * WG_ROWS;
          if (code == 31) {
#pragma unroll
            for (int i = 0; i < NCH_G; ++i) {
              const int64_t pos = pos0 + pG0 + i * RPP_G;
              if (pos < n) src[i] = perm ? __ldg(perm + pos) : (int)pos;
            }
          } else {
            const int k = k0 + code * OPS + hA;
            if (k < kend) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int64_t pos = pos0 + pA0 + 16 * i;
                if (pos < n) src[i] = __ldg(nbrp + (int64_t)k * n + pos);
              }
            }
          }
        }
      };
      auto get_rows = [&](int it, const int* src, float4* v) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < nitems) {
          if ((items_s[it] & 31u) == 31u) {
#pragma unroll
            for (int i = 0; i < NCH_G; ++i)
              if (src[i] >= 0) v[i] = __ldg((const float4*)(gout + (int64_t)src[i] * COUT) + jG);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (src[i] >= 0) v[i] = __ldg((const float4*)(in + (int64_t)src[i] * CIN + cA));
          }
        }
      };
      auto process = [&](int it, const float4* v) {
        const int code = (int)(items_s[it] & 31u);
        if (code == 31) {
          const int gb = gcount % NG;
          if (gcount >= NG) tc::mbar_wait(&gempty_bar[gb], (uint32_t)((gcount / NG) - 1) & 1u);
          tc::fence_after_sync();
          uint8_t* hi_p = g_base + gb * G_BUF;
          uint8_t* lo_p = hi_p + G_PLANE;
#pragma unroll
          for (int i = 0; i < NCH_G; ++i) {
            float4 h4, l4;
            tc::split_tf32(v[i], h4, l4);
            const uint32_t off = wg_off(jG, pG0 + i * RPP_G);
            *(float4*)(hi_p + off) = h4;
            *(float4*)(lo_p + off) = l4;
          }
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&gfull_bar[gb]);
          ++gcount;
        } else {
          const int st = jcount & 1;
          if (jcount >= 2) tc::mbar_wait(&empty_bar[st], (uint32_t)((jcount >> 1) - 1) & 1u);
          tc::fence_after_sync();
          uint8_t* hi_p = a_base + st * WG_A_STAGE;
          uint8_t* lo_p = hi_p + WG_A_PLANE;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 h4, l4;
            tc::split_tf32(v[i], h4, l4);
            const uint32_t off = wg_off(j16, pA0 + 16 * i);
            *(float4*)(hi_p + off) = h4;
            *(float4*)(lo_p + off) = l4;
          }
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full_bar[st]);
          touched |= 1u << code;
          ++jcount;
        }
      };
      // software pipeline: indices of items it+1, it+2 and rows of item it are in registers
      int i1[4], i2[4], i3[4];
      float4 vcur[4], vnext[4];
      get_idx(0, i3);
      get_idx(1, i1);
      get_idx(2, i2);
      get_rows(0, i3, vcur);
      for (int it = 0; it < nitems; ++it) {
        get_rows(it + 1, i1, vnext);
        get_idx(it + 3, i3);
        process(it, vcur);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vcur[i] = vnext[i];
          i1[i] = i2[i];
          i2[i] = i3[i];
        }
      }
    } else {
      // ================= MMA issuer warp =================
      const uint32_t idesc = tc::idesc_tf32(128, COUT) | (1u << 15) | (1u << 16);   // A and B MN-major
      for (int it = 0; it < nitems; ++it) {
        const int code = (int)(items_s[it] & 31u);
        if (code == 31) {
          cur_gb = gcount % NG;
          tc::mbar_wait(&gfull_bar[cur_gb], (uint32_t)(gcount / NG) & 1u);
          ++gcount;
          continue;
        }
        const int st = jcount & 1;
        tc::mbar_wait(&full_bar[st], (uint32_t)(jcount >> 1) & 1u);
        tc::fence_after_sync();
        const bool last = (it + 1 == nitems) || ((items_s[it + 1] & 31u) == 31u);   // last step of its tile
        if (tc::elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(code * COUT);
          const uint32_t a_hi = tc::smem_u32(a_base + st * WG_A_STAGE), a_lo = a_hi + WG_A_PLANE;
          const uint32_t b_hi = tc::smem_u32(g_base + cur_gb * G_BUF), b_lo = b_hi + G_PLANE;
          uint32_t acc = (touched >> code) & 1u;
#pragma unroll
          for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
            const uint64_t dah = wg_desc(a_hi + ks * WG_KSTEP), dal = wg_desc(a_lo + ks * WG_KSTEP);
            const uint64_t dbh = wg_desc(b_hi + ks * WG_KSTEP), dbl = wg_desc(b_lo + ks * WG_KSTEP);
            tc::mma_tf32(d, dal, dbh, idesc, acc);      // small terms first
            tc::mma_tf32(d, dah, dbl, idesc, 1);
            tc::mma_tf32(d, dah, dbh, idesc, 1);
            acc = 1;
          }
          tc::mma_commit(&empty_bar[st]);
          if (last) tc::mma_commit(&gempty_bar[cur_gb]);
        }
        __syncwarp();
        touched |= 1u << code;
        ++jcount;
      }
    }
    __syncthreads();      // everybody is done READING the item list (MMAs may still be in flight)
  }

  // ---- all MMAs complete -> flush the partial sums of the slots this CTA touched ----
  if (warp == WG_THREADS / 32) {
    if (tc::elect_one()) {
      if (touched) tc::mma_commit(&done_bar);
      else tc::mbar_arrive(&done_bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&done_bar, 0);
  tc::fence_after_sync();
  if (warp < WG_THREADS / 32) {
    const int q = warp & 3, cs = warp >> 2;
    const int m = q * 32 + lane;                        // TMEM lane = row of D
    const int h = m / CIN, ci = m % CIN;
    for (int sl = 0; sl < nslots; ++sl) {
      if (!((touched >> sl) & 1u)) continue;
      const int k = k0 + sl * OPS + h;
      for (int s = cs; s < COUT / 16; s += 4) {
        float acc[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl * COUT + s * 16), acc);
        if (k < kend) {
          float* dst = gw + ((int64_t)k * CIN + ci) * COUT + s * 16;
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            lk_red_add_v4(dst + e, make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]));
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

extern "C" int lk_conv_wgrad_tc_supported(int c_in, int c_out) {
  return (c_in == 32 || c_in == 64 || c_in == 128) && (c_out == 32 || c_out == 64 || c_out == 128);
}

extern "C" int lk_conv_wgrad_prepass(const int32_t* d_nbr, const int32_t* d_perm, int64_t n, int k,
                                     int32_t* d_nbrp, uint32_t* d_masks, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && k > 0 && k <= 32, "lk_conv_wgrad_prepass: bad sizes (1 <= K <= 32)");
  LK_REQUIRE((d_perm == nullptr) == (d_nbrp == nullptr),
             "lk_conv_wgrad_prepass: the permuted copy goes with a plan order");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_nbr && d_masks, "lk_conv_wgrad_prepass: null pointer");
  const int64_t tiles = (n + WG_ROWS - 1) / WG_ROWS;
  wgrad_prepass_kernel<<<lk_grid(tiles * 32, 256, 8), 256, 0, (cudaStream_t)s>>>(d_nbr, d_perm, n, k, d_nbrp,
                                                                              d_masks);
  LK_LAUNCHED();
  return LK_OK;
}

template <int CIN, int COUT>
static int launch_wgrad(const float* in, const float* gout, const int* nbrp, const int* perm,
                        const unsigned* masks, int64_t n, int k, int slots, float* gw,
                        cudaStream_t st) {
  constexpr int OPS = 128 / CIN;
  constexpr uint32_t SMEM = 2 * WG_A_STAGE + WG_G_REGION + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    LK_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SMEM));
    attr_set = true;
  }
  int max_slots = 512 / COUT;
  if (max_slots > WG_MAX_SLOTS) max_slots = WG_MAX_SLOTS;
  if (slots <= 0 || slots > max_slots) slots = max_slots;
  const int opg_max = slots * OPS;
  const int groups = (k + opg_max - 1) / opg_max;
  int opg = (k + groups - 1) / groups;                 // balanced groups, whole slots
  opg = (opg + OPS - 1) / OPS * OPS;
  const int64_t tiles = (n + WG_ROWS - 1) / WG_ROWS;
  int64_t splits = LK_SM_COUNT / groups;
  if (splits < 1) splits = 1;
  if (splits > tiles) splits = tiles;
  dim3 grid((unsigned)splits, (unsigned)((k + opg - 1) / opg));
  conv_wgrad_tc_kernel<CIN, COUT><<<grid, WG_THREADS + 32, SMEM, st>>>(in, gout, nbrp, perm, masks, n, k, opg,
                                                                    gw);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_conv_wgrad_tc(const float* d_in, const float* d_gout, const int32_t* d_nbrp,
                                const int32_t* d_perm, const uint32_t* d_masks, int64_t n, int k,
                                int c_in, int c_out, float* d_gw, int slots_per_cta, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && k > 0 && k <= 32 && n * k < (1LL << 31) && n < (1LL << 32) * WG_ROWS / 32,
             "lk_conv_wgrad_tc: bad sizes (1 <= K <= 32, K n < 2^31)");
  LK_REQUIRE(lk_conv_wgrad_tc_supported(c_in, c_out), "lk_conv_wgrad_tc: channels must be 32, 64 or 128");
  LK_REQUIRE(d_gw, "lk_conv_wgrad_tc: null output");
  cudaStream_t st = (cudaStream_t)s;
  LK_CUDA(cudaMemsetAsync(d_gw, 0, (size_t)k * c_in * c_out * sizeof(float), st));
  lk_count_launch();
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_in && d_gout && d_nbrp && d_masks, "lk_conv_wgrad_tc: null pointer");
  LK_REQUIRE((uintptr_t)d_in % 16 == 0 && (uintptr_t)d_gout % 16 == 0 && (uintptr_t)d_gw % 16 == 0,
             "lk_conv_wgrad_tc: buffers must be 16-byte aligned");
#define LK_WGRAD_CASE(CI, CO) \
  if (c_in == CI && c_out == CO) \
    return launch_wgrad<CI, CO>(d_in, d_gout, d_nbrp, d_perm, d_masks, n, k, slots_per_cta, d_gw, st)
  LK_WGRAD_CASE(32, 32);
  LK_WGRAD_CASE(32, 64);
  LK_WGRAD_CASE(32, 128);
  LK_WGRAD_CASE(64, 32);
  LK_WGRAD_CASE(64, 64);
  LK_WGRAD_CASE(64, 128);
  LK_WGRAD_CASE(128, 32);
  LK_WGRAD_CASE(128, 64);
  LK_WGRAD_CASE(128, 128);
#undef LK_WGRAD_CASE
  lk_set_error("lk_conv_wgrad_tc: unsupported channel combination");
  return LK_EINVAL;
}
