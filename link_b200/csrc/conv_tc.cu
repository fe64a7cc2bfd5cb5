// Sparse 3D convolution on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
// Weight-stationary, TMEM-resident formulation (replaces the reference's host loop of
// gather -> cuBLAS mm -> scatter per offset, convolution_cuda.cu:101-164):
//
//   * one persistent CTA per SM works in rounds of up to T = 256/C_out 128-row output tiles (tile
//     slots interleaved over the CTAs); their fp32 accumulators [128 x C_out] live in TMEM columns
//     [0, 256) for the whole round, so partial sums never touch registers, shared memory or HBM;
//   * (offset, tile) steps in which no row of the tile has a neighbour are skipped altogether:
//     lk_conv_plan (conv_plan.cu) orders the output rows so that tiles are homogeneous in the
//     offsets they use and hands the kernel one offset mask per tile -- on LiDAR scans ~27 % of
//     the dense steps remain;
//   * the gathered A operand ALSO lives in TMEM (columns [256, 512): two stages of tf32 hi/lo
//     planes): producers write the rows they gathered straight from registers with tcgen05.st
//     (lane = row, column = input channel) and the MMA reads A from TMEM ("TS" form).  With A in
//     shared memory an M=128, N=64, K=8 tf32 MMA has to fetch 4 KB of A + 2 KB of B per 32-cycle
//     instruction (192 B/clk > the 128 B/clk shared-memory port) and the kernel was operand-fetch
//     bound at ~45 % tensor-pipe utilisation; from TMEM only B crosses the shared-memory port;
//   * outer loop over the K kernel offsets: W[k] (tf32 hi/lo planes, K-major SWIZZLE_128B in shared
//     memory, double buffered) is staged ONCE per offset and reused by all T tiles of the CTA;
//   * inner loop over the tiles: 16 producer warps gather the 128 input rows nbr[k, tile rows]
//     (missing neighbours -> zero rows; indices prefetched two steps ahead, rows one step ahead),
//     split them into tf32 hi/lo and tcgen05.st them; one elected lane of a 17th warp issues
//     3*C_in/8 tcgen05.mma (M=128, N=C_out, K=8; A_lo.B_hi + A_hi.B_lo + A_hi.B_hi) into the tile's
//     accumulator columns; tcgen05.commit -> "empty" mbarrier releases the A stage, producers
//     signal "full" mbarriers (one arrival per warp) -- no __syncthreads in the main loop;
//   * epilogue: tcgen05.ld 32x32b (thread = output row) -> scale/shift (folded BatchNorm / bias)
//     -> + residual -> ReLU -> one streaming store per row.
//
// No atomics, no temporaries, deterministic, output written exactly once.
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

#define CT_ROWS 128
#ifndef CT_NWB
#define CT_NWB 3          // weight ring depth (tuning knob; 3 and 4 measure the same, 6 is slower)
#endif
#ifndef CT_GST
#define CT_GST 3          // gather ring depth at 64 channels per step (3 x 32 KB)
#endif
#define CT_THREADS 512      // producer threads (16 warps); one more warp issues the MMAs
// TMEM: columns [0, 512 - 128 NST) hold the accumulators, the rest NST A stages of 128 columns

template <int CIN, int COUT, int NST>
struct ConvTcCfg {
  // C_in = 128 is processed as two K-halves of 64 channels: every real offset k becomes KH = 2
  // "virtual offsets" (k, h) with their own 64-channel W image and A stage, so the TMEM budget of
  // an A stage (hi + lo planes = 128 columns) and the step structure stay those of C_in = 64.
  static constexpr int KH = CIN > 64 ? 2 : 1;
  static constexpr int CE = CIN / KH;                       // channels per step (32 or 64)
  static constexpr int KB = CE / 32;                        // 128-byte K-blocks per weight row
  static constexpr uint32_t B_BLK = COUT * 128;             // one [COUT x 32] K-block
  static constexpr uint32_t B_PLANE = KB * B_BLK;           // hi (or lo) plane
  static constexpr uint32_t B_STAGE = 2 * B_PLANE;          // hi + lo of one virtual offset
  static constexpr int NWB = B_STAGE > 32768 ? 3 : CT_NWB;  // W ring: prefetched NWB-NST offsets ahead
  // Gather ring: the rows of the next GST-1 steps are fetched with cp.async into per-thread slots of
  // shared memory (no registers held while in flight).  0 = register ring of depth 2 (C_out = 128:
  // the 64 KB weight images leave no room).
  static constexpr uint32_t G_STAGE = CT_ROWS * CE * 4;     // raw fp32 rows of one step
  static constexpr int GST = B_STAGE > 32768 ? 0 : (CE == 64 ? CT_GST : 4);
  static constexpr uint32_t SMEM = NWB * B_STAGE + GST * G_STAGE + 1024;   // + alignment slack
  static constexpr int ACC_COLS = 512 - 128 * NST;
  static constexpr int MAX_TILES = ACC_COLS / COUT;
  static constexpr int A_STAGE_COLS = 128;                  // hi plane at +0, lo plane at +64
  static constexpr int CPT = CE / 4;                        // input channels per producer thread
};


// PREC = 0: 3xTF32 (fp32-level accuracy, three MMAs per k-slice on tf32 hi / lo planes);
// PREC = 1: single-pass TF32 (the tensor core truncates the fp32 operands to tf32: ~1e-3 relative,
//           what frameworks call "allow_tf32"): no split, one TMEM store, one MMA per k-slice.
// WS = true (the configurations with a shared-memory gather ring): the 16 producer warps are
// SPECIALISED instead of marching through every step in lock step -- warps 8..15 only gather
// (cp.async rows of step j+1, j+2 into the ring, completion reported through
// cp.async.mbarrier.arrive on a per-slot mbarrier), warps 0..7 only convert (ring slot -> tf32 hi / lo
// -> tcgen05.st), each side running ahead as far as the ring / the TMEM stages allow.  In the
// lock-step form every step paid  wait_group + a 512-thread bar.sync + LDGSTS issue + LDS + split +
// STTM  back to back on all warps, so the LSU, the shared-memory port and the TMEM store path were
// used one after the other (issue slots 26 % busy, no pipe above 35 %); here they overlap.
#define CT_CONV_WARPS 8
#define CT_GATHER_THREADS 256
// IO16 = true (warp-specialised configurations only, PREC = 1): the feature rows, the residual and the
// output are bf16.  A bf16 value is exact in tf32, so the A operand needs no lo plane and one MMA per
// k-slice remains (weights rounded to tf32); a gathered row is half the bytes -- half the LDGSTS per
// step and a ring twice as deep in the same shared memory -- and the accumulation stays fp32 in TMEM.
template <int CIN, int COUT, int NST, int PREC, bool WS, bool IO16>
__global__ void __launch_bounds__(CT_THREADS + 32, 1) conv_tc_kernel(
    const float* __restrict__ in, const float* __restrict__ wimg /*[K] packed B-operand images*/,
    const int* __restrict__ nbr /*[K][n_out]*/,
    const int* __restrict__ perm /*[n_out] plan position -> output row, or NULL = identity*/,
    const unsigned* __restrict__ tile_mask /*[tiles] offsets present in each tile, or NULL = all*/,
    int64_t n_out, int K, int tgroup /*adjacent tiles per slot group: 1, 2 or 4*/, int snake, lk_conv_epilogue_t ep,
    float* __restrict__ out) {
  using Cfg = ConvTcCfg<CIN, COUT, NST>;
  constexpr int CPT = Cfg::CPT;                 // 16 (CIN = 64) or 8 (CIN = 32)
  constexpr int MAXT = Cfg::MAX_TILES;
  constexpr int KH = Cfg::KH, CE = Cfg::CE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* const b_base = smem;
  __shared__ uint64_t full_bar[NST];  // A stage written   (one arrival per producer warp)
  __shared__ uint64_t empty_bar[NST]; // A stage consumed  (tcgen05.commit)
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t wfull_bar[Cfg::NWB];  // W[k] image landed (bulk copy, complete_tx)
  __shared__ uint64_t round_bar;            // every MMA of the round has completed (tcgen05.commit after the last step)
  __shared__ uint64_t slot_full[8];         // WS: ring slot filled   (cp.async.mbarrier.arrive of the gather threads)
  __shared__ uint64_t slot_empty[8];        // WS: ring slot consumed (one arrival per convert warp)
  __shared__ uint16_t steps_s[64 * MAXT];   // active (virtual offset, tile slot) steps: kv << 4 | t
  __shared__ uint8_t klist_s[64];           // distinct virtual offsets of the round, ascending
  __shared__ int orow_s[MAXT][CT_ROWS];     // output row of every tile-slot row (-1 past the end)
  __shared__ uint32_t tmask_s[MAXT];
  __shared__ int warp_cnt_s[16];
  __shared__ int nsteps_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t total_tiles = (n_out + CT_ROWS - 1) / CT_ROWS;
  const uint32_t kmask = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);
  const int n32 = (int)n_out;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NST; ++b) {
      tc::mbar_init(&full_bar[b], (WS && Cfg::GST > 0) ? CT_CONV_WARPS : CT_THREADS / 32);
      tc::mbar_init(&empty_bar[b], 1);
    }
    tc::mbar_init(&round_bar, 1);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      tc::mbar_init(&slot_full[b], CT_GATHER_THREADS);
      tc::mbar_init(&slot_empty[b], CT_CONV_WARPS);
    }
#pragma unroll
    for (int b = 0; b < Cfg::NWB; ++b) tc::mbar_init(&wfull_bar[b], 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  lk_pdl_enter();      // nothing above reads global memory: TMEM allocation and barrier set-up overlap the predecessor
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_a0 = tmem_base + Cfg::ACC_COLS;        // A stage s at + s * 128 columns

  // Persistent CTA: round r owns the tile slots t < MAXT with tile = blockIdx.x + (r MAXT + t) gridDim.x.
  // Interleaving spreads the heavy classes at the end of the plan order over all CTAs and keeps
  // the number of rounds per CTA minimal (a round costs ~5 us of set-up, drain and epilogue; work-
  // balanced CONTIGUOUS ranges share each W[k] between more steps but give the CTAs that own the
  // light classes 10+ rounds and were 25 % slower end to end).
  int jbase = 0;      // steps issued in earlier rounds: A-stage index and mbarrier phase continue
  int wcount = 0;     // offsets (W[k] images) consumed so far: weight ring index / phase continue
  for (int round = 0;; ++round) {
    // tile of slot t: groups of `tgroup` ADJACENT tiles, the groups interleaved over the CTAs.  Adjacent
    // tiles of the plan order share their offset class, so a round needs fewer distinct W[k] images
    // (each is a 32 KB bulk copy per CTA and round: at tgroup = 1 the weight traffic, 256 MB per conv at
    // N = 119k, equals the gather traffic), while interleaving the groups still spreads the heavy
    // classes at the end of the plan order over all CTAs.
    const int64_t round0 = (int64_t)round * MAXT * gridDim.x;
    // Stripes of gridDim.x groups alternate direction (CTA b takes group b of even stripes, group
    // grid-1-b of odd ones): the plan order is roughly ascending in cost, so a fixed direction hands the
    // high CTAs the heavier end of EVERY stripe.
    auto tile_of = [&](int t) -> int64_t {
      const int stripe = round * (MAXT / tgroup) + t / tgroup;
      const int bb = (snake && (stripe & 1)) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
      const int64_t tile = round0 + ((int64_t)(t / tgroup) * gridDim.x + bb) * tgroup + (t % tgroup);
      return tile < total_tiles ? tile : -1;
    };
    if (round0 >= total_tiles) break;
    const int ntiles = MAXT;                                              // slots; invalid ones have no mask and no rows
    // ---- active step list of the round, k-major (W[k] is staged once per offset and round) ----
    if (tid < MAXT)
      tmask_s[tid] = tile_of(tid) >= 0 ? (tile_mask ? tile_mask[tile_of(tid)] & kmask : kmask) : 0u;
    __syncthreads();
    if (tid < CT_THREADS) {
      const int kk = tid / MAXT, tt = tid % MAXT;            // kk = virtual offset (k, h) = k KH + h
      const bool on = kk < K * KH && ((tmask_s[tt] >> (kk / KH)) & 1u);
      const unsigned bal = __ballot_sync(0xffffffffu, on);
      if (lane == 0) warp_cnt_s[warp] = __popc(bal);
      __syncwarp();
      asm volatile("bar.sync 1, 512;" ::: "memory");          // the 16 list-building warps only
      int pos = __popc(bal & ((1u << lane) - 1u));
#pragma unroll
      for (int w = 0; w < 16; ++w) pos += (w < warp) ? warp_cnt_s[w] : 0;
      if (on) steps_s[pos] = (uint16_t)((kk << 4) | tt);
      if (tid == CT_THREADS - 1) nsteps_s = pos + (on ? 1 : 0);
    } else {
      uint32_t uni = 0;
#pragma unroll
      for (int t = 0; t < MAXT; ++t) uni |= tmask_s[t];
      if ((uni >> lane) & 1u) {
        const int base = __popc(uni & ((1u << lane) - 1u)) * KH;
#pragma unroll
        for (int h = 0; h < KH; ++h) klist_s[base + h] = (uint8_t)(lane * KH + h);
      }
    }
    for (int i = tid; i < MAXT * CT_ROWS; i += CT_THREADS + 32) {
      const int t = i / CT_ROWS, r = i % CT_ROWS;
      int o = -1;
      if (tile_of(t) >= 0) {
        const int64_t pos = tile_of(t) * CT_ROWS + r;
        if (pos < n_out) o = perm ? __ldg(perm + pos) : (int)pos;
      }
      orow_s[t][r] = o;
    }
    __syncthreads();
    const int nsteps = nsteps_s;
    uint32_t uni_mask = 0;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) uni_mask |= tmask_s[t];
    const int nk = __popc(uni_mask) * KH;       // distinct virtual offsets of this round
    // W[k] images arrive by bulk copy into a ring of NWB buffers, NWB-NST offsets ahead of their
    // use.  Offset number w (global count wcount + local index) lives in buffer w % NWB; the copy
    // for offset w + NWB - NST is issued when offset w begins and overwrites offset w - NST, all of
    // whose MMAs are complete: the NST - 1 offsets in between have >= 1 step each and the issuing
    // thread has just observed the commit of step j - NST (a round boundary drains everything).
    auto issue_w = [&](int local_idx) {
      if (local_idx < nk) {
        const int wg = wcount + local_idx;
        uint64_t* bar = &wfull_bar[wg % Cfg::NWB];
        const uint32_t dst = tc::smem_u32(b_base + (wg % Cfg::NWB) * Cfg::B_STAGE);
        const float* src = wimg + (int64_t)klist_s[local_idx] * (Cfg::B_STAGE / 4);
        constexpr uint32_t W_BYTES = PREC == 0 ? Cfg::B_STAGE : Cfg::B_PLANE;   // single-pass TF32 reads the hi plane only
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     ::"r"(tc::smem_u32(bar)), "r"(W_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(W_BYTES), "r"(tc::smem_u32(bar)) : "memory");
      }
    };
    if (tid == 0) {
#pragma unroll
      for (int w = 0; w < Cfg::NWB - NST; ++w) issue_w(w);
    }

    if (warp == CT_THREADS / 32) {
      // ================= MMA issuer warp =================
      const uint32_t idesc = tc::idesc_tf32(128, COUT);
      const uint64_t db0 = tc::smem_desc_sw128(tc::smem_u32(b_base));
      uint32_t touched = 0;
      int k_prev = -1, wc = wcount;
      for (int j = 0; j < nsteps; ++j) {
        const int jg = jbase + j;
        const int stage = jg % NST;
        const int k = steps_s[j] >> 4, t = steps_s[j] & 15;
        if (k != k_prev) {                                 // first step of an offset: its W image
          k_prev = k;
          tc::mbar_wait(&wfull_bar[wc % Cfg::NWB], (uint32_t)(wc / Cfg::NWB) & 1u);
          ++wc;                                            // wc - 1 = ring index of this offset
        }
        tc::mbar_wait(&full_bar[stage], (uint32_t)(jg / NST) & 1u);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(t * COUT);
          const uint32_t a_hi = tmem_a0 + (uint32_t)(stage * Cfg::A_STAGE_COLS);
          const uint32_t a_lo = a_hi + 64;
          const uint64_t db_hi = db0 + (uint64_t)((((wc - 1) % Cfg::NWB) * Cfg::B_STAGE) >> 4);
          const uint64_t db_lo = db_hi + (uint64_t)(Cfg::B_PLANE >> 4);
          uint32_t acc = (touched >> t) & 1u;
#ifndef CT_DEBUG_NO_MMA          // timing experiments only (results are wrong): see DESIGN.md 3.3
#pragma unroll
          for (int ks = 0; ks < CE / 8; ++ks) {
            // B descriptor start address is in 16-byte units: K-block ks/4, 32-byte slice ks%4
            const uint64_t bo = (uint64_t)(((ks >> 2) * Cfg::B_BLK + (ks & 3) * 32) >> 4);
            if (PREC == 0) {
              tc::mma_tf32_ts(d, a_lo + 8 * ks, db_hi + bo, idesc, acc);
              tc::mma_tf32_ts(d, a_hi + 8 * ks, db_lo + bo, idesc, 1);
              tc::mma_tf32_ts(d, a_hi + 8 * ks, db_hi + bo, idesc, 1);
            } else {
              tc::mma_tf32_ts(d, a_hi + 8 * ks, db_hi + bo, idesc, acc);
            }
            acc = 1;
          }
#endif
          tc::mma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        touched |= 1u << t;
      }
      // end of the round: one more commit, on a barrier of its own, that every producer warp waits for
      // before the epilogue.  (Waiting for the last commit of each A stage by parity is only safe for a
      // thread that has followed that barrier phase by phase; the gather warps have not.)
      if (tc::elect_one()) {
        if (nsteps > 0) tc::mma_commit(&round_bar);
        else tc::mbar_arrive(&round_bar);
      }
      __syncwarp();
    } else {
      // ================= producer warps (gather + tf32 split + tcgen05.st), later the epilogue ====
      // TMEM access rule: warp w touches lanes [32 (w%4), +32).  Thread (q = w%4, lane) owns tile row
      // 32q + lane; the 4 warps of a lane quarter split the C_in input channels into 4 slices.
      const int q = warp & 3, cs = warp >> 2;
      const int prow = q * 32 + lane;
      const int col0 = cs * CPT;                                // my channel slice [col0, col0 + CPT)
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      auto load_idx = [&](int jj) -> int {
        if (jj >= nsteps) return -1;
        const int kv2 = steps_s[jj] >> 4, t2 = steps_s[jj] & 15;
        const int o = orow_s[t2][prow];
        const int src = o >= 0 ? __ldg(nbr + (uint32_t)((kv2 / KH) * n32 + o)) : -1;   // K n_out < 2^31 (host check)
        // the K-half rides in bit 30 of the (non-negative) row index
        return (KH > 1 && src >= 0) ? (src | ((kv2 % KH) << 30)) : src;
      };
      auto load_rows = [&](int src, float4* v) {
#ifdef CT_DEBUG_NO_GATHER
        src = -1;
#endif
        if (src >= 0) {
          const int half = KH > 1 ? (src >> 30) : 0;
          if (KH > 1) src &= 0x3FFFFFFF;
          const float4* rp = (const float4*)(in + (int64_t)src * CIN + half * CE + col0);
#pragma unroll
          for (int i = 0; i < CPT / 4; ++i) v[i] = __ldg(rp + i);
        } else {
#pragma unroll
          for (int i = 0; i < CPT / 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      int k_prev = -1, wc = wcount;
      // everything of a step after its rows are in registers (hi/lo split done by the caller)
      auto publish = [&](int j, const float* hi, const float* lo) {
        const int jg = jbase + j;
        const int stage = jg % NST;
        const int k = steps_s[j] >> 4;
        if (jg >= NST) tc::mbar_wait(&empty_bar[stage], (uint32_t)((jg / NST) - 1) & 1u);
        tc::fence_after_sync();
        if (k != k_prev) {                      // an offset begins: prefetch the one after next
          k_prev = k;
          if (tid == 0) issue_w(wc - wcount + Cfg::NWB - NST);
          ++wc;
        }
        const uint32_t a_hi = tmem_a0 + (uint32_t)(stage * Cfg::A_STAGE_COLS) + lane_addr + (uint32_t)col0;
#ifndef CT_DEBUG_NO_ST
        if (CPT == 16) {
          tc::tmem_st16(a_hi, hi);
          if (PREC == 0) tc::tmem_st16(a_hi + 64, lo);
        } else {
          tc::tmem_st8(a_hi, hi);
          if (PREC == 0) tc::tmem_st8(a_hi + 64, lo);
        }
        tc::tmem_st_wait();
#endif
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
      };
      if constexpr (WS && Cfg::GST > 0) {
        constexpr int EB = IO16 ? 2 : 4;                // bytes per feature element
        constexpr int GST = IO16 ? 2 * Cfg::GST : Cfg::GST;   // same shared memory, rows half the size
        constexpr int CH_ROW = CE * EB / 16;            // 16-byte chunks per row
        constexpr uint32_t ROW_B = CE * EB;             // row pitch in a ring slot
        constexpr uint32_t SLOT_B = CT_ROWS * ROW_B;
        static_assert(!IO16 || PREC == 1, "bf16 rows run as single-pass TF32");
        if (warp >= CT_CONV_WARPS) {
          // ================= gather warps: rows of step j -> ring slot (jbase + j) % GST =================
          const int gt = tid - CT_CONV_WARPS * 32;
          constexpr int RPP = CT_GATHER_THREADS / CH_ROW;   // rows per pass (16 lanes fetch one 256-byte row)
          constexpr int NP = CT_ROWS / RPP;                 // passes per step
          const uint32_t g_base = tc::smem_u32(b_base + Cfg::NWB * Cfg::B_STAGE);
          const int f_chunk = gt % CH_ROW, f_row0 = gt / CH_ROW;
          uint32_t fetch_off[NP];
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const int fr = f_row0 + i * RPP;
            fetch_off[i] = (uint32_t)(fr * ROW_B + ((f_chunk ^ (fr & (CH_ROW - 1))) << 4));
          }
          auto load_idx = [&](int jj, int* src) {
#pragma unroll
            for (int i = 0; i < NP; ++i) src[i] = -1;
            if (jj < nsteps) {
              const int kv2 = steps_s[jj] >> 4, t2 = steps_s[jj] & 15;
#pragma unroll
              for (int i = 0; i < NP; ++i) {
                const int o = orow_s[t2][f_row0 + i * RPP];
                const int sr = o >= 0 ? __ldg(nbr + (uint32_t)((kv2 / KH) * n32 + o)) : -1;
                src[i] = (KH > 1 && sr >= 0) ? (sr | ((kv2 % KH) << 30)) : sr;
              }
            }
          };
          auto gstep = [&](int j, int* src) {
            const int jg = jbase + j;
            const int slot = jg % GST, use = jg / GST;
            if (use >= 1) tc::mbar_wait(&slot_empty[slot], (uint32_t)(use - 1) & 1u);
            const uint32_t dst = g_base + (uint32_t)slot * SLOT_B;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const int sv = src[i];
              const int half = (KH > 1 && sv >= 0) ? (sv >> 30) : 0;
              const int row = sv >= 0 ? (KH > 1 ? (sv & 0x3FFFFFFF) : sv) : 0;
              const uint8_t* rp = (const uint8_t*)in + ((int64_t)row * CIN + half * CE) * EB + 16 * f_chunk;
              const int nbytes = sv >= 0 ? 16 : 0;      // 0: zero-fill (missing neighbour)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                           ::"r"(dst + fetch_off[i]), "l"(rp), "r"(nbytes) : "memory");
            }
            // this thread's arrival on the slot's mbarrier fires when its copies above have landed
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];"
                         ::"r"(tc::smem_u32(&slot_full[slot])) : "memory");
            load_idx(j + 2, src);                        // indices two steps ahead of their use
          };
          int idx_a[NP], idx_b[NP];
          load_idx(0, idx_a);
          load_idx(1, idx_b);
          for (int j = 0; j < nsteps; j += 2) {
            gstep(j, idx_a);
            if (j + 1 < nsteps) gstep(j + 1, idx_b);
          }
        } else {
          // ================= convert warps: ring slot -> tf32 hi / lo -> TMEM A stage =================
          const int cq = warp & 3, ccs = warp >> 2;       // TMEM lane quarter, channel half
          const int crow = cq * 32 + lane;
          constexpr int CPC = CE / 2;                     // channels per convert thread (32 or 16)
          constexpr int NCC = CPC * EB / 16;              // 16-byte chunks per thread and step
          const int ccol0 = ccs * CPC;
          const uint32_t clane = (uint32_t)(cq * 32) << 16;
          uint32_t read_off[NCC];
#pragma unroll
          for (int i = 0; i < NCC; ++i)
            read_off[i] = (uint32_t)(crow * ROW_B + (((ccs * NCC + i) ^ (crow & (CH_ROW - 1))) << 4));
          int k_prev = -1, wc = wcount;
          for (int j = 0; j < nsteps; ++j) {
            const int jg = jbase + j;
            const int slot = jg % GST, stage = jg % NST;
            const int k = steps_s[j] >> 4;
            tc::mbar_wait(&slot_full[slot], (uint32_t)(jg / GST) & 1u);
            const uint8_t* sp = b_base + Cfg::NWB * Cfg::B_STAGE + (uint32_t)slot * SLOT_B;
            const uint32_t a_hi = tmem_a0 + (uint32_t)(stage * Cfg::A_STAGE_COLS) + clane + (uint32_t)ccol0;
            bool waited = false;
#pragma unroll
            for (int h = 0; h < CPC / 16; ++h) {          // 16 channels at a time
              float hi[16], lo[16];
              if constexpr (IO16) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {             // 8 bf16 per chunk: fp32 = bits << 16 (exact, and exact in tf32)
                  const uint4 v = *(const uint4*)(sp + read_off[2 * h + i]);
                  const unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    hi[8 * i + 2 * e] = __uint_as_float(w4[e] << 16);
                    hi[8 * i + 2 * e + 1] = __uint_as_float(w4[e] & 0xFFFF0000u);
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 v4 = *(const float4*)(sp + read_off[4 * h + i]);
                  float4 h4 = v4, l4 = v4;
                  if (PREC == 0) tc::split_tf32(v4, h4, l4);
                  hi[4 * i] = h4.x; hi[4 * i + 1] = h4.y; hi[4 * i + 2] = h4.z; hi[4 * i + 3] = h4.w;
                  lo[4 * i] = l4.x; lo[4 * i + 1] = l4.y; lo[4 * i + 2] = l4.z; lo[4 * i + 3] = l4.w;
                }
              }
              if (h == CPC / 16 - 1) {                    // every value of the slot is in registers
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&slot_empty[slot]);
              }
              if (!waited) {
                waited = true;
                if (jg >= NST) tc::mbar_wait(&empty_bar[stage], (uint32_t)((jg / NST) - 1) & 1u);
                tc::fence_after_sync();
                if (k != k_prev) {                        // an offset begins: prefetch the W image after next
                  k_prev = k;
                  if (tid == 0) issue_w(wc - wcount + Cfg::NWB - NST);
                  ++wc;
                }
              }
              tc::tmem_st16(a_hi + 16 * h, hi);
              if (PREC == 0) tc::tmem_st16(a_hi + 64 + 16 * h, lo);
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
          }
        }
      } else if constexpr (Cfg::GST > 0) {
        // ---- shared-memory gather ring with COALESCED row fetches.  Measured: when
        // every thread fetches its own row quarter, a warp-level 16-byte load touches 32 different
        // rows = 32 L1TEX tag cycles, 2 048 cycles per step and SM -- the kernel was bound by that,
        // not by latency or bandwidth.  Here the CE/4 lanes that are adjacent in a warp fetch the
        // CE/4 consecutive 16-byte chunks of ONE row (cp.async, 2-4 rows = 4 cache lines per
        // instruction, 8x fewer tag cycles), steps j+1 .. j+GST-1 stay in flight in the ring, and
        // after a producer-wide named barrier every thread reads the row quarter it has to store
        // to TMEM (lane = row).  Missing neighbours are zero-filled copies (src-size 0); chunks are
        // XOR-swizzled by the row so that both access patterns spread over all banks. ----
        constexpr int GST = Cfg::GST, D = GST - 1;
        constexpr int NCH = CPT / 4;                    // 16-byte chunks per thread and step
        constexpr int CH_ROW = CE / 4;                  // chunks per row = lanes per fetched row
        constexpr int RPP = CT_THREADS / CH_ROW;        // rows fetched per pass by the 512 threads
        static_assert(RPP * NCH == CT_ROWS, "fetch passes must cover the tile");
        const uint32_t g_base = tc::smem_u32(b_base + Cfg::NWB * Cfg::B_STAGE);
        const int f_chunk = tid % CH_ROW, f_row0 = tid / CH_ROW;       // fetch role
        uint32_t read_off[NCH], fetch_off[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          read_off[i] = (uint32_t)(prow * (CE * 4) + (((cs * NCH + i) ^ (prow & (CH_ROW - 1))) << 4));
          const int fr = f_row0 + i * RPP;
          fetch_off[i] = (uint32_t)(fr * (CE * 4) + ((f_chunk ^ (fr & (CH_ROW - 1))) << 4));
        }
        auto load_idx4 = [&](int jj, int* src) {        // neighbour rows of the NCH rows I fetch in step jj
#pragma unroll
          for (int i = 0; i < NCH; ++i) src[i] = -1;
          if (jj < nsteps) {
            const int kv2 = steps_s[jj] >> 4, t2 = steps_s[jj] & 15;
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
              const int o = orow_s[t2][f_row0 + i * RPP];
              const int sr = o >= 0 ? __ldg(nbr + (uint32_t)((kv2 / KH) * n32 + o)) : -1;   // K n_out < 2^31 (host check)
              src[i] = (KH > 1 && sr >= 0) ? (sr | ((kv2 % KH) << 30)) : sr;
            }
          }
        };
        int w_slot = 0, r_slot = 0;                     // ring positions of the next fetch / next read
        auto issue_gather = [&](int jj, const int* src) {   // rows of step jj -> slot jj % GST
          const uint32_t dst = g_base + (uint32_t)w_slot * Cfg::G_STAGE;
          w_slot = (w_slot + 1 == GST) ? 0 : w_slot + 1;
          if (jj < nsteps) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
#ifdef CT_DEBUG_NO_GATHER
              const int sv = -1;
#else
              const int sv = src[i];
#endif
              const int half = (KH > 1 && sv >= 0) ? (sv >> 30) : 0;
              const int row = sv >= 0 ? (KH > 1 ? (sv & 0x3FFFFFFF) : sv) : 0;
              const float* rp = in + (int64_t)row * CIN + half * CE + 4 * f_chunk;
              const int nbytes = sv >= 0 ? 16 : 0;      // 0: zero-fill (missing neighbour)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                           ::"r"(dst + fetch_off[i]), "l"(rp), "r"(nbytes) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");   // one group per step, even if empty
        };
        int idx_a[NCH], idx_b[NCH];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          load_idx4(d, idx_a);
          issue_gather(d, idx_a);
        }
        load_idx4(D, idx_a);                            // indices of steps j+D, two iterations ahead
        load_idx4(D + 1, idx_b);
        auto produce = [&](int j, int* idx_cur) {
          // my copies of step j (and older) have landed; after the barrier so have everybody's,
          // and every thread has finished reading step j-1, whose slot the next fetch overwrites
          asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
          asm volatile("bar.sync 2, 512;" ::: "memory");
          issue_gather(j + D, idx_cur);
          load_idx4(j + D + 2, idx_cur);
          const uint8_t* slot = b_base + Cfg::NWB * Cfg::B_STAGE + (uint32_t)r_slot * Cfg::G_STAGE;
          r_slot = (r_slot + 1 == GST) ? 0 : r_slot + 1;
          float hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            const float4 v4 = *(const float4*)(slot + read_off[i]);
            float4 h4 = v4, l4 = v4;
            if (PREC == 0) tc::split_tf32(v4, h4, l4);
            hi[4 * i] = h4.x; hi[4 * i + 1] = h4.y; hi[4 * i + 2] = h4.z; hi[4 * i + 3] = h4.w;
            lo[4 * i] = l4.x; lo[4 * i + 1] = l4.y; lo[4 * i + 2] = l4.z; lo[4 * i + 3] = l4.w;
          }
          publish(j, hi, lo);
        };
        for (int j = 0; j < nsteps; j += 2) {
          produce(j, idx_a);
          if (j + 1 < nsteps) produce(j + 1, idx_b);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        // ---- register ring of depth 2: the rows of step j are fetched while steps j-2 and j-1 are
        // being processed, their indices two steps before that ----
        float4 va[CPT / 4], vb[CPT / 4];
        load_rows(load_idx(0), va);
        load_rows(load_idx(1), vb);
        int idx_a = load_idx(2), idx_b = load_idx(3);     // rows of steps 2 / 3, fetched at steps 0 / 1
        auto produce = [&](int j, float4* cur, int& idx_cur) {
          // split BEFORE waiting for the stage: after the wake-up only the TMEM stores remain on
          // the critical path  commit(j-2) -> tcgen05.st -> full(j) -> MMA(j)
          float hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < CPT / 4; ++i) {
            float4 h4 = cur[i], l4 = cur[i];
            if (PREC == 0) tc::split_tf32(cur[i], h4, l4);
            hi[4 * i] = h4.x; hi[4 * i + 1] = h4.y; hi[4 * i + 2] = h4.z; hi[4 * i + 3] = h4.w;
            lo[4 * i] = l4.x; lo[4 * i + 1] = l4.y; lo[4 * i + 2] = l4.z; lo[4 * i + 3] = l4.w;
          }
          load_rows(idx_cur, cur);                // refill this buffer: rows of step j + 2
          idx_cur = load_idx(j + 4);              // ... and the index for its next refill
          publish(j, hi, lo);
        };
        for (int j = 0; j < nsteps; j += 2) {
          produce(j, va, idx_a);
          if (j + 1 < nsteps) produce(j + 1, vb, idx_b);
        }
      }
      // ---- drain: every MMA of the round has completed ----
      tc::mbar_wait(&round_bar, (uint32_t)round & 1u);
      tc::fence_after_sync();
      // ---- epilogue: thread = output row (TMEM lane quarter q), 16 columns per warp ----
      constexpr int NSLICE = COUT / 16;
      for (int sl = cs; sl < NSLICE; sl += 4) {
        const int c_base = sl * 16;
        for (int tt = 0; tt < ntiles; ++tt) {
          const int orow = orow_s[tt][prow];
          float acc[16];
          if (tmask_s[tt]) {
            tc::tmem_ld16(tmem_base + lane_addr + (uint32_t)(tt * COUT + c_base), acc);
          } else {                              // tile without a single pair: the sum is empty
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
          }
          if (orow >= 0 && IO16) {
            // bf16 rows: 16 channels = 32 bytes = two 128-bit stores; residual rows are bf16 as well
            const int64_t o = orow;
            float y[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              y[e] = acc[e];
              if (ep.d_scale) y[e] *= __ldg(ep.d_scale + c_base + e);
              if (ep.d_shift) y[e] += __ldg(ep.d_shift + c_base + e);
            }
            if (ep.d_residual) {
              const uint4* rp = (const uint4*)((const __nv_bfloat16*)ep.d_residual + o * COUT + c_base);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint4 v = __ldg(rp + i);
                const unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  y[8 * i + 2 * e] += __uint_as_float(w4[e] << 16);
                  y[8 * i + 2 * e + 1] += __uint_as_float(w4[e] & 0xFFFF0000u);
                }
              }
            }
            uint4* dst16 = (uint4*)((__nv_bfloat16*)out + o * COUT + c_base);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              unsigned w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a0 = y[8 * i + 2 * e], a1 = y[8 * i + 2 * e + 1];
                if (ep.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                const __nv_bfloat162 pk = __floats2bfloat162_rn(a0, a1);      // .x = low half = the even channel
                w4[e] = *(const unsigned*)&pk;
              }
              dst16[i] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
          } else if (orow >= 0) {
            const int64_t o = orow;
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
    // round boundary: every MMA has completed (drain) and every accumulator has been read before
    // the next round's first MMA overwrites it
    wcount += nk;
    jbase += nsteps;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

template <int CIN, int COUT, int NST, int PREC, bool WS, bool IO16 = false>
static int launch_conv_tc_n(const float* in, const float* wimg, const int* nbr, const int* perm,
                          const unsigned* tile_mask, int64_t n_out, int k,
                          const lk_conv_epilogue_t& ep, float* out, cudaStream_t st) {
  using Cfg = ConvTcCfg<CIN, COUT, NST>;
  static bool attr_set = false;
  if (!attr_set) {
    LK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CIN, COUT, NST, PREC, WS, IO16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg::SMEM));
    attr_set = true;
  }
  int64_t tiles = (n_out + CT_ROWS - 1) / CT_ROWS;
  int grid = (int)(tiles < LK_SM_COUNT ? tiles : LK_SM_COUNT);   // persistent: one CTA per SM
  // adjacent-tile groups only pay with a plan order (classes are contiguous there) and when every CTA
  // still gets a full group; LINKB200_CONV_TGROUP overrides (A/B measurements)
  static int tg_env = -1;
  if (tg_env < 0) {
    const char* e = getenv("LINKB200_CONV_TGROUP");
    tg_env = (e && (e[0] == '1' || e[0] == '2' || e[0] == '4')) ? e[0] - '0' : 0;
  }
  int tgroup = 1;                                      // measured: 2 / 4 adjacent tiles per group are 23 % / 70 % slower
  if (tg_env) tgroup = tg_env;                         // (the CTAs that draw the heavy classes finish last)
  while (tgroup > 1 && (Cfg::MAX_TILES % tgroup)) tgroup >>= 1;
  static int snake = -1;
  if (snake < 0) {
    const char* e = getenv("LINKB200_CONV_SNAKE");
    snake = (e && e[0] == '1') ? 1 : 0;        // measured slower (77 vs 69 us): the plain interleave already pairs the light end of
                                                  // every stripe with the extra heavy tile of the last, partial stripe
  }
  LK_PDL_LAUNCH((conv_tc_kernel<CIN, COUT, NST, PREC, WS, IO16>), grid, CT_THREADS + 32, Cfg::SMEM, st, in, wimg, nbr, perm,
                tile_mask, n_out, k, tgroup, snake, ep, out);
  LK_LAUNCHED();
  return LK_OK;
}

// number of TMEM A stages (tuning knob, default 2): 3 stages hide more of the producer <-> MMA
// handshake latency but leave room for only two accumulator tiles per round at C_out = 64, i.e.
// twice the rounds; measured slower
static int conv_tc_stages() {
  static int nst = 0;
  if (!nst) {
    const char* e = getenv("LINKB200_CONV_STAGES");
    nst = (e && e[0] == '3') ? 3 : 2;
  }
  return nst;
}

template <int CIN, int COUT>
static int launch_conv_tc(const float* in, const float* wimg, const int* nbr, const int* perm,
                          const unsigned* tile_mask, int64_t n_out, int k,
                          const lk_conv_epilogue_t& ep, float* out, cudaStream_t st) {
  (void)conv_tc_stages;
  // warp-specialised producers wherever the configuration has a shared-memory gather ring
  // (LINKB200_CONV_LOCKSTEP=1 selects the lock-step producers for A/B measurements)
  static int lockstep = -1;
  if (lockstep < 0) {
    const char* e = getenv("LINKB200_CONV_LOCKSTEP");
    lockstep = (e && e[0] == '1') ? 1 : 0;
  }
  constexpr bool has_ring = ConvTcCfg<CIN, COUT, 2>::GST > 0;
  if (has_ring && !lockstep) {
    if (ep.precision == LK_PREC_TF32)
      return launch_conv_tc_n<CIN, COUT, 2, 1, true>(in, wimg, nbr, perm, tile_mask, n_out, k, ep, out, st);
    return launch_conv_tc_n<CIN, COUT, 2, 0, true>(in, wimg, nbr, perm, tile_mask, n_out, k, ep, out, st);
  }
  if (ep.precision == LK_PREC_TF32)
    return launch_conv_tc_n<CIN, COUT, 2, 1, false>(in, wimg, nbr, perm, tile_mask, n_out, k, ep, out, st);
  return launch_conv_tc_n<CIN, COUT, 2, 0, false>(in, wimg, nbr, perm, tile_mask, n_out, k, ep, out, st);
}

extern "C" int lk_conv_tc_supported(int c_in, int c_out) {
  return (c_in == 32 || c_in == 64 || c_in == 128) && (c_out == 32 || c_out == 64 || c_out == 128);
}

// W[k] [COUT][CIN] -> the kernel's shared-memory images: one per virtual offset (k, h) (h = K-half
// of 64 channels when CIN = 128, else a single half): tf32 hi plane, then lo plane, each CE/32
// K-blocks of [COUT rows x 128 B] in the SWIZZLE_128B pattern (one thread per 16-byte chunk)
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ wt, int K, int cin,
                                                           int cout, float* __restrict__ img) {
  const int kh = cin > 64 ? 2 : 1, ce = cin / kh;
  const int kb_n = ce / 32;
  const int64_t per_v = (int64_t)cout * kb_n * 8;          // 16-byte chunks per plane and virtual offset
  const int64_t total = (int64_t)K * kh * per_v;
  const int64_t stage_floats = (int64_t)2 * ce * cout;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int kv = (int)(t / per_v);
    const int r = (int)(t - (int64_t)kv * per_v);
    const int k = kv / kh, h = kv % kh;
    const int chunk = r & 7, kb = (r >> 3) % kb_n, row = r / (8 * kb_n);
    float4 w4 = __ldg((const float4*)(wt + ((int64_t)k * cout + row) * cin + h * ce + kb * 32 + chunk * 4)), hi, lo;
    tc::split_tf32(w4, hi, lo);
    const uint32_t off = (uint32_t)kb * (uint32_t)(cout * 128) + tc::sw128_offset(row, chunk);
    float* base = img + (int64_t)kv * stage_floats;
    *(float4*)((uint8_t*)base + off) = hi;
    *(float4*)((uint8_t*)base + (size_t)ce * cout * 4 + off) = lo;
  }
}

// General form: the source may be the module parameter itself ([K][src_cin][src_cout], transposed on the
// fly: layout 1) or its per-offset transpose ([K][src_cout][src_cin]: layout 0), narrower than the image
// (zero padded to cin x cout) and read with the offsets reversed (flip_k: the input gradient of a
// submanifold conv) -- no transposed / padded / flipped copy is made on the way.
__global__ void __launch_bounds__(256) pack_weights_ex_kernel(const float* __restrict__ w, int K, int cin, int cout,
                                                              int src_cin, int src_cout, int layout, int flip_k,
                                                              float* __restrict__ img) {
  const int kh = cin > 64 ? 2 : 1, ce = cin / kh;
  const int kb_n = ce / 32;
  const int64_t per_v = (int64_t)cout * kb_n * 8;
  const int64_t total = (int64_t)K * kh * per_v;
  const int64_t stage_floats = (int64_t)2 * ce * cout;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int kv = (int)(t / per_v);
    const int r = (int)(t - (int64_t)kv * per_v);
    const int k = kv / kh, h = kv % kh;
    const int chunk = r & 7, kb = (r >> 3) % kb_n, row = r / (8 * kb_n);       // row = output channel
    const int ks = flip_k ? K - 1 - k : k;
    const int ci0 = h * ce + kb * 32 + chunk * 4;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ci = ci0 + e;
      float x = 0.f;
      if (row < src_cout && ci < src_cin)
        x = layout ? __ldg(w + ((int64_t)ks * src_cin + ci) * src_cout + row)
                   : __ldg(w + ((int64_t)ks * src_cout + row) * src_cin + ci);
      v[e] = x;
    }
    float4 w4 = make_float4(v[0], v[1], v[2], v[3]), hi, lo;
    tc::split_tf32(w4, hi, lo);
    const uint32_t off = (uint32_t)kb * (uint32_t)(cout * 128) + tc::sw128_offset(row, chunk);
    float* base = img + (int64_t)kv * stage_floats;
    *(float4*)((uint8_t*)base + off) = hi;
    *(float4*)((uint8_t*)base + (size_t)ce * cout * 4 + off) = lo;
  }
}

extern "C" int lk_conv_tc_pack_weights_ex(const float* d_w, int k, int c_in, int c_out, int src_c_in, int src_c_out,
                                          int layout, int flip_k, float* d_img, lk_stream_t s) {
  LK_REQUIRE(k > 0 && lk_conv_tc_supported(c_in, c_out) && src_c_in > 0 && src_c_in <= c_in && src_c_out > 0 &&
                 src_c_out <= c_out && (layout == 0 || layout == 1),
             "lk_conv_tc_pack_weights_ex: image channels must be 32, 64 or 128 and cover the source");
  LK_REQUIRE(d_w && d_img && (uintptr_t)d_img % 128 == 0, "lk_conv_tc_pack_weights_ex: null or misaligned pointer");
  pack_weights_ex_kernel<<<lk_grid((int64_t)k * c_out * (c_in / 32) * 8, 256, 4), 256, 0, (cudaStream_t)s>>>(
      d_w, k, c_in, c_out, src_c_in, src_c_out, layout, flip_k, d_img);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_conv_tc_pack_weights(const float* d_wt, int k, int c_in, int c_out, float* d_img,
                                       lk_stream_t s) {
  LK_REQUIRE(k > 0 && lk_conv_tc_supported(c_in, c_out), "lk_conv_tc_pack_weights: channels must be 32, 64 or 128");
  LK_REQUIRE(d_wt && d_img && (uintptr_t)d_wt % 16 == 0 && (uintptr_t)d_img % 128 == 0,
             "lk_conv_tc_pack_weights: null or misaligned pointer");
  pack_weights_kernel<<<lk_grid((int64_t)k * c_out * (c_in / 32) * 8, 256, 4), 256, 0, (cudaStream_t)s>>>(
      d_wt, k, c_in, c_out, d_img);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_conv_tc_fwd(const float* d_in, const float* d_wt, const int32_t* d_nbr,
                              int64_t n_out, int k, int c_in, int c_out, const float* d_bias,
                              float* d_out, lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, d_bias, nullptr, 0, 0};
  return lk_conv_tc_fwd_plan(d_in, d_wt, d_nbr, nullptr, nullptr, n_out, k, c_in, c_out, &ep, d_out, s);
}

extern "C" int lk_conv_tc_fwd_ex(const float* d_in, const float* d_wt, const int32_t* d_nbr,
                                 int64_t n_out, int k, int c_in, int c_out,
                                 const lk_conv_epilogue_t* epp, float* d_out, lk_stream_t s) {
  return lk_conv_tc_fwd_plan(d_in, d_wt, d_nbr, nullptr, nullptr, n_out, k, c_in, c_out, epp, d_out, s);
}

extern "C" int lk_conv_tc_fwd_plan(const float* d_in, const float* d_wt, const int32_t* d_nbr,
                                   const int32_t* d_perm, const uint32_t* d_tile_mask,
                                   int64_t n_out, int k, int c_in, int c_out,
                                   const lk_conv_epilogue_t* epp, float* d_out, lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, nullptr, nullptr, 0, 0};
  if (epp) ep = *epp;
  LK_REQUIRE(n_out >= 0 && k > 0 && k <= 32 && n_out * k < (1LL << 31), "lk_conv_tc_fwd: bad sizes (1 <= K <= 32, K n_out < 2^31)");
  LK_REQUIRE(lk_conv_tc_supported(c_in, c_out), "lk_conv_tc_fwd: channels must be 32, 64 or 128");
  LK_REQUIRE((d_perm == nullptr) == (d_tile_mask == nullptr),
             "lk_conv_tc_fwd: perm and tile_mask come together (lk_conv_plan)");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in && d_wt && d_nbr && d_out, "lk_conv_tc_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_in % 16 == 0 && (uintptr_t)d_wt % 16 == 0 && (uintptr_t)d_out % 16 == 0,
             "lk_conv_tc_fwd: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
#define LK_CONV_TC_CASE(CI, CO) \
  if (c_in == CI && c_out == CO) return launch_conv_tc<CI, CO>(d_in, d_wt, d_nbr, d_perm, d_tile_mask, n_out, k, ep, d_out, st)
  LK_CONV_TC_CASE(32, 32);
  LK_CONV_TC_CASE(32, 64);
  LK_CONV_TC_CASE(32, 128);
  LK_CONV_TC_CASE(64, 32);
  LK_CONV_TC_CASE(64, 64);
  LK_CONV_TC_CASE(64, 128);
  LK_CONV_TC_CASE(128, 32);
  LK_CONV_TC_CASE(128, 64);
  LK_CONV_TC_CASE(128, 128);
#undef LK_CONV_TC_CASE
  lk_set_error("lk_conv_tc_fwd: unsupported channel combination");
  return LK_EINVAL;
}

// bf16 feature rows in, bf16 rows out (fp32 accumulation in TMEM, weights rounded to tf32, the epilogue's
// residual rows bf16 as well): the configurations with a shared-memory gather ring.
extern "C" int lk_conv_tc_bf16_supported(int c_in, int c_out) {
  return (c_in == 32 || c_in == 64 || c_in == 128) && (c_out == 32 || c_out == 64);
}

extern "C" int lk_conv_tc_fwd_bf16(const void* d_in, const float* d_wimg, const int32_t* d_nbr,
                                   const int32_t* d_perm, const uint32_t* d_tile_mask, int64_t n_out, int k,
                                   int c_in, int c_out, const lk_conv_epilogue_t* epp, void* d_out,
                                   lk_stream_t s) {
  lk_conv_epilogue_t ep = {nullptr, nullptr, nullptr, 0, 0};
  if (epp) ep = *epp;
  ep.precision = LK_PREC_TF32;
  LK_REQUIRE(n_out >= 0 && k > 0 && k <= 32 && n_out * k < (1LL << 31), "lk_conv_tc_fwd_bf16: bad sizes (1 <= K <= 32, K n_out < 2^31)");
  LK_REQUIRE(lk_conv_tc_bf16_supported(c_in, c_out), "lk_conv_tc_fwd_bf16: c_in in {32,64,128}, c_out in {32,64}");
  LK_REQUIRE((d_perm == nullptr) == (d_tile_mask == nullptr), "lk_conv_tc_fwd_bf16: perm and tile_mask come together");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in && d_wimg && d_nbr && d_out, "lk_conv_tc_fwd_bf16: null pointer");
  LK_REQUIRE((uintptr_t)d_in % 16 == 0 && (uintptr_t)d_wimg % 16 == 0 && (uintptr_t)d_out % 16 == 0 &&
                 (uintptr_t)ep.d_residual % 16 == 0, "lk_conv_tc_fwd_bf16: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
#define LK_CONV_BF16_CASE(CI, CO)                                                                          \
  if (c_in == CI && c_out == CO)                                                                           \
    return launch_conv_tc_n<CI, CO, 2, 1, true, true>((const float*)d_in, d_wimg, d_nbr, d_perm, d_tile_mask, n_out, k, \
                                                      ep, (float*)d_out, st)
  LK_CONV_BF16_CASE(32, 32);
  LK_CONV_BF16_CASE(32, 64);
  LK_CONV_BF16_CASE(64, 32);
  LK_CONV_BF16_CASE(64, 64);
  LK_CONV_BF16_CASE(128, 32);
  LK_CONV_BF16_CASE(128, 64);
#undef LK_CONV_BF16_CASE
  lk_set_error("lk_conv_tc_fwd_bf16: unsupported channel combination");
  return LK_EINVAL;
}

