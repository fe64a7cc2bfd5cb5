// The LinK block's linear-kernel path as three kernels (reference: ELKBlock.forward,
// segmentation/core/models/semantic_kitti/linkencoder.py:124-185; voxel_to_aux / aux_to_voxel,
// segmentation/core/models/utils.py:44-84):
//
//   preagg      : S[b]  += F_in[i] * {cos,sin,(lin)}(p_i)      for every voxel i of block b
//   window_mean : A[b]   = (sum_{b' in N(b)} S[b']) / (sum_{b' in N(b)} n[b'])
//   apply       : out[i] = A_c[b(i)] cos p_i + A_s[b(i)] sin p_i (+ A_l - F_in p)  (+ fused norms)
//
// The reference materialises [N,kC] weighted planes (cat), scatter-means them with one CTA per
// voxel and float atomics per element, multiplies back by the counts, gathers r^3 neighbours with
// a global read-modify-write per neighbour, divides, gathers [N,kC] back to voxels and combines in
// five more elementwise kernels.  Here the phase p = W.x is recomputed from the integer coordinate
// in both passes, so the only HBM streams are: read F_in once, read coords + block index twice,
// write out once; block sums / means are L2 resident.
//
// Thread mapping.  A row of C floats is owned by LPR lanes; lane j of the group owns the VPL
// float4 vectors {i * LPR + j : i < VPL} (channels 4 (i LPR + j) .. +3), so every load instruction
// of a lane group covers 16 LPR contiguous bytes and a row is fetched with VPL fully coalesced
// instructions.  For C = 16, 32, 64, 128:  VPL = 4 (16 channels per lane), LPR = C / 16 -- one warp
// step covers 32 / LPR rows.  (Any other C % 4 == 0: VPL = 1, LPR = pow2 >= C / 4.)  Compared with a
// 4-channel-per-lane layout this cuts the per-row bookkeeping (index loads, address arithmetic,
// run-boundary tests) and the LayerNorm shuffle trees 4x, and channel groups
// (pos.repeat([1, groups]), linkencoder.py:152) fall INSIDE a lane: vectors i and i + IB share
// their phases, so a lane evaluates only 4 IB sincos per row and reuses them.
#include <stdlib.h>

#include "common.cuh"

#include "link_common.cuh"


// ------------------------------------------------------------------ pass 1: block sums
// One kernel, two visiting orders:
//   * segmented (order != NULL; what the block executor uses): voxels are visited in BLOCK order
//     through the sort permutation -- position i of the sorted sequence is voxel row order[i] in
//     block row rank[i] (both from lk_sort_unique_ex).  A warp loads 32 PL consecutive pairs with
//     coalesced loads; every lane group walks RPG consecutive sorted positions (about one whole
//     block at (3x7)^3) with the row loads of PRE_U voxels in flight, reduces runs of equal block
//     row in registers -- a warp-level segmented reduction over the variable-length voxel lists
//     of the blocks -- and issues one vector reduction per run.  Feature rows are still fetched as
//     whole 4C-byte rows.
//   * storage order (order == NULL, rank = voxel -> block row): same code with the identity
//     permutation; LiDAR scans arrive in column order, so runs are ~1 voxel long and the kernel
//     is bound by the L2 reduction rate (REDG) instead of HBM -- kept for callers without a sort.
template <int LPR, int VPL, int IB, int OP>
__global__ void __launch_bounds__(128, VPL >= 4 ? 3 : 5) link_preagg_kernel(
    const float* __restrict__ fin, const int4* __restrict__ coords, const int* __restrict__ order,
    const int* __restrict__ rank, int64_t n, GenDev g, float* sums) {
  constexpr int G = 32 / LPR;                    // row groups per warp
  constexpr int RPG = LPR > 8 ? LPR : 8;         // sorted positions walked by one lane group
  constexpr int PL = RPG / LPR;                  // positions held per lane
  constexpr int K = (OP == LK_OP_COSX) ? 3 : 2;
  constexpr bool COSX = (OP == LK_OP_COSX);
  constexpr int NP = 4 * IB;
  constexpr int PRE_U = (COSX && VPL > 1) ? 2 : 4;   // rows in flight per lane group (register budget)
  static_assert(RPG % PRE_U == 0, "batching must divide the span");
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;
  const int j = lane % LPR;
  const bool active = 4 * j < g.c;               // VPL == 1 only: C/4 need not be a power of two
  const int kc = K * g.c;
  LaneGen<NP> lg;
  load_lane_gen<LPR, IB>(g, j, active, lg);

  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (G * RPG); base < n;
       base += warps_total * (G * RPG)) {
    int ord[PL], rk[PL];
#pragma unroll
    for (int q = 0; q < PL; ++q) {
      const int64_t pos = base + (int64_t)lane * PL + q;
      ord[q] = pos < n ? (order ? __ldg(order + pos) : (int)pos) : 0;
      rk[q] = pos < n ? __ldg(rank + pos) : -1;
    }
    float acc[K][VPL][4];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[q][i][e] = 0.f;
    int cur = -1;
#pragma unroll
    for (int i0 = 0; i0 < RPG; i0 += PRE_U) {
      int b[PRE_U], cx[PRE_U], cy[PRE_U], cz[PRE_U];
      float4 f[PRE_U][VPL];
#pragma unroll
      for (int u = 0; u < PRE_U; ++u) {          // all row loads of the batch first (MLP)
        const int ii = i0 + u;
        const int src = grp * LPR + ii / PL;
        const int r = __shfl_sync(0xffffffffu, ord[ii % PL], src);
        b[u] = __shfl_sync(0xffffffffu, rk[ii % PL], src);
        const bool ok = b[u] >= 0;
        int4 cc = ok ? __ldg(coords + r) : make_int4(0, 0, 0, 0);
        cx[u] = cc.x; cy[u] = cc.y; cz[u] = cc.z;
        const float4* row = (const float4*)(fin + (int64_t)r * g.c) + j;
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          f[u][i] = (ok && active) ? lk_ldg_stream(row + i * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < PRE_U; ++u) {
        if (b[u] >= 0) {                         // else: past the end / unmapped voxel
          float p[NP], sn[NP], cs[NP];
          lane_trig<NP, COSX>(g, lg, cx[u], cy[u], cz[u], p, sn, cs);
          if (b[u] != cur) {                     // run boundary: flush the finished block
            if (cur >= 0 && active) {
              float* dst = sums + (int64_t)cur * kc + 4 * j;
#pragma unroll
              for (int q = 0; q < K; ++q)
#pragma unroll
                for (int i = 0; i < VPL; ++i)
                  lk_red_add_v4(dst + q * g.c + 4 * i * LPR,
                                make_float4(acc[q][i][0], acc[q][i][1], acc[q][i][2], acc[q][i][3]));
            }
#pragma unroll
            for (int q = 0; q < K; ++q)
#pragma unroll
              for (int i = 0; i < VPL; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[q][i][e] = 0.f;
            cur = b[u];
          }
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const float fv[4] = {f[u][i].x, f[u][i].y, f[u][i].z, f[u][i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int q = (i % IB) * 4 + e;
              // plane order: cos -> [cos, sin]; sin -> [sin, cos]; cos_x -> [cos, sin, lin]
              acc[0][i][e] = fmaf(fv[e], (OP == LK_OP_SIN ? sn[q] : cs[q]), acc[0][i][e]);
              acc[1][i][e] = fmaf(fv[e], (OP == LK_OP_SIN ? cs[q] : sn[q]), acc[1][i][e]);
              if (COSX) acc[K - 1][i][e] = fmaf(fv[e], p[q], acc[K - 1][i][e]);
            }
          }
        }
      }
    }
    if (cur >= 0 && active) {
      float* dst = sums + (int64_t)cur * kc + 4 * j;
#pragma unroll
      for (int q = 0; q < K; ++q)
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          lk_red_add_v4(dst + q * g.c + 4 * i * LPR,
                        make_float4(acc[q][i][0], acc[q][i][1], acc[q][i][2], acc[q][i][3]));
    }
  }
}


// ------------------------------------------------------------------ pass 1, shared-memory staged
// C in {16, 32, 64, 128}.  Same visiting orders as link_preagg_kernel, but the gathered feature
// rows never sit in registers while they are in flight: a warp issues ALL its row fetches
// (G x 8 rows = 8 KB, 16 cp.async of 16 bytes per lane, plus the rows' coordinates) straight into
// its private shared-memory stage right after it has read its 32 (voxel row, block row) pairs, so
//   * one dependent memory round trip (pairs -> rows) instead of one per register batch,
//   * ~60 registers per thread -> 24 resident warps per SM x 8 KB = 190 KB of HBM reads in flight
//     per SM, against the ~45 KB that saturate HBM3e (Little: 6.5 TB/s x ~1 us / 148 SMs).
// The accumulation then runs out of shared memory (conflict-free 128-bit reads: the LPR lanes of a
// row read consecutive 16-byte vectors).
#define PS_RPG 8                         // sorted positions walked by one lane group
#define PS_WARPS 4
template <int LPR, int IB, int OP>
__global__ void __launch_bounds__(PS_WARPS * 32, (IB == 1 && OP != LK_OP_COSX) ? 6 : 4) link_preagg_smem_kernel(
    const float* __restrict__ fin, const int4* __restrict__ coords, const int* __restrict__ order,
    const int* __restrict__ rank, int64_t n, GenDev g, float* sums) {
  constexpr int VPL = 2;
  constexpr int G = 32 / LPR;                    // row groups per warp
  constexpr int ROWS = G * PS_RPG;               // rows per warp step (8 KB of features)
  constexpr int PL = ROWS > 32 ? ROWS / 32 : 1;  // (voxel row, block row) pairs held per lane
  constexpr int K = (OP == LK_OP_COSX) ? 3 : 2;
  constexpr bool COSX = (OP == LK_OP_COSX);
  constexpr int NP = 4 * IB;
  constexpr int ROW_BYTES = 16 * VPL * LPR;      // = 4 C
  __shared__ __align__(16) uint8_t stage_s[PS_WARPS][ROWS * ROW_BYTES];
  __shared__ __align__(16) int4 coord_s[PS_WARPS][ROWS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / LPR;
  const int j = lane % LPR;
  const int kc = K * g.c;
  LaneGen<NP> lg;
  load_lane_gen<LPR, IB>(g, j, true, lg);
  uint8_t* const stage = stage_s[wib];
  const uint32_t stage_u = (uint32_t)__cvta_generic_to_shared(stage);
  const uint32_t coord_u = (uint32_t)__cvta_generic_to_shared(coord_s[wib]);

  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ROWS; base < n;
       base += warps_total * ROWS) {
    int ord[PL], rk[PL];
#pragma unroll
    for (int q = 0; q < PL; ++q) {
      const int64_t pos = base + (int64_t)lane * PL + q;
      const bool ok = pos < n && lane * PL + q < ROWS;
      ord[q] = ok ? (order ? __ldg(order + pos) : (int)pos) : 0;
      rk[q] = ok ? __ldg(rank + pos) : -1;
    }
    // ---- every fetch of the step in flight at once ----
    int bb[PS_RPG];                                  // block row of each of the group's positions
#pragma unroll
    for (int u = 0; u < PS_RPG; ++u) {
      const int p = grp * PS_RPG + u;                // position inside the warp step
      const int r = __shfl_sync(0xffffffffu, ord[p % PL], p / PL);
      bb[u] = __shfl_sync(0xffffffffu, rk[p % PL], p / PL);
      if (bb[u] >= 0) {
        const float* src = fin + (int64_t)r * g.c + 4 * j;
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                       ::"r"(stage_u + p * ROW_BYTES + (i * LPR + j) * 16), "l"(src + 4 * i * LPR) : "memory");
        if (j == 0)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
                       ::"r"(coord_u + p * 16), "l"(coords + r) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- segmented reduction over the group's rows, out of shared memory ----
    float acc[K][VPL][4];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[q][i][e] = 0.f;
    int cur = bb[0];                                 // positions past the end (b < 0) only trail
#pragma unroll
    for (int u = 0; u < PS_RPG; ++u) {
      const int p = grp * PS_RPG + u;
      const int b = bb[u];
      if (b != cur) {                                // run boundary: flush the finished block
        if (cur >= 0) {
          float* dst = sums + (int64_t)cur * kc + 4 * j;
#pragma unroll
          for (int q = 0; q < K; ++q)
#pragma unroll
            for (int i = 0; i < VPL; ++i)
              lk_red_add_v4(dst + q * g.c + 4 * i * LPR,
                            make_float4(acc[q][i][0], acc[q][i][1], acc[q][i][2], acc[q][i][3]));
        }
#pragma unroll
        for (int q = 0; q < K; ++q)
#pragma unroll
          for (int i = 0; i < VPL; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[q][i][e] = 0.f;
        cur = b;
      }
      // rows past the end were never fetched: their stage bytes are stale, so they are masked
      const int4 cc = *(const int4*)((const uint8_t*)coord_s[wib] + p * 16);
      float ph[NP], sn[NP], cs[NP];
      lane_trig<NP, COSX>(g, lg, b >= 0 ? cc.x : 0, b >= 0 ? cc.y : 0, b >= 0 ? cc.z : 0, ph, sn, cs);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 f4 = *(const float4*)(stage + p * ROW_BYTES + (i * LPR + j) * 16);
        const float fv[4] = {b >= 0 ? f4.x : 0.f, b >= 0 ? f4.y : 0.f, b >= 0 ? f4.z : 0.f, b >= 0 ? f4.w : 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = (i % IB) * 4 + e;
          acc[0][i][e] = fmaf(fv[e], (OP == LK_OP_SIN ? sn[q] : cs[q]), acc[0][i][e]);
          acc[1][i][e] = fmaf(fv[e], (OP == LK_OP_SIN ? cs[q] : sn[q]), acc[1][i][e]);
          if (COSX) acc[K - 1][i][e] = fmaf(fv[e], ph[q], acc[K - 1][i][e]);
        }
      }
    }
    if (cur >= 0) {
      float* dst = sums + (int64_t)cur * kc + 4 * j;
#pragma unroll
      for (int q = 0; q < K; ++q)
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          lk_red_add_v4(dst + q * g.c + 4 * i * LPR,
                        make_float4(acc[q][i][0], acc[q][i][1], acc[q][i][2], acc[q][i][3]));
    }
    __syncwarp();                                // the stage is rewritten by the next step
  }
}

// ------------------------------------------------------------------ pass 1, balanced ring
// Same arithmetic and visiting order as link_preagg_smem_kernel, scheduled differently:
//   * every lane group owns ONE contiguous range of q sorted positions, q = ceil(n / resident lane
//     groups), so the whole grid is resident in a single wave and all warps finish together for any
//     n (the one-step-per-warp kernel above runs 934 CTAs on 888 slots at n = 119k: a second wave
//     of 46 CTAs that costs a full latency chain);
//   * a warp streams its range through a ring of PR_STAGES chunks (PR_RC rows per lane group and
//     chunk): the row fetches of the next chunk(s) are in flight while chunk c is reduced, and the
//     (voxel row, block row) pairs are prefetched one more chunk ahead, so the pair -> row
//     dependency is paid once per warp, not once per step;
//   * leaner inner loop: SFU trig without the per-phase accuracy switch, block rows handed over
//     through shared memory, rows past the end skipped, not masked.
// Measured (B200, n = 119 325, C = 64, cos, CUDA-graph replay of 56 launches over 8 rotating
// feature buffers; scripts/preagg_ab.py, gpurun_out r01p3/r01p4): one-step-per-warp kernel 11.4 us;
// ring 3 stages x 4 rows 10.0 us, 2 x 4 9.2, 4 x 4 10.4, 2 x 8 10.3, 3 x 2 9.0, 2 x 2 8.9-9.3
// (default), 2 x 1 9.6; 24 instead of 16 warps per SM +0.3 us.  Small chunks win: the tail after
// the last fetch lands is one chunk of arithmetic.  Timing experiments (results wrong by
// construction): without the trig 9.5 us, without the reductions ~10 us (unchanged), without the
// feature fetch 7.4 us, without fetch and coordinates 6.4 us, without any of them 6.3 us -- the
// skeleton (launch, pair loads, ring hand-over) is 2/3 of the kernel; the 30.5 MB feature gather
// adds ~3 us.
#ifndef PR_STAGES
#define PR_STAGES 2
#endif
#ifndef PR_WARPS
#define PR_WARPS 4
#endif
#ifndef PR_RC
#define PR_RC 2                            // rows per lane group and chunk
#endif
#ifndef PR_MINB
#define PR_MINB 4                          // resident CTAs per SM the register budget is capped for
#endif
template <int LPR>
struct RingCfg {
  static constexpr int G = 32 / LPR;                 // lane groups (rows side by side) per warp
  static constexpr int RC = LPR < PR_RC ? LPR : PR_RC;   // rows per lane group and chunk
  static constexpr int CR = G * RC;                  // rows per warp chunk (<= 32: one pair per lane)
  static constexpr int ROW_BYTES = 32 * LPR;         // = 4 C
  static constexpr int SLOT = CR * ROW_BYTES + CR * 16 + CR * 4;   // rows | coords | block rows
  static constexpr int SMEM = PR_WARPS * PR_STAGES * SLOT;
};

#ifdef PR_DEBUG_NORED                                // timing experiments only (results are wrong)
#define PR_RED(p, v) do { if ((v).x == 123.456f) lk_red_add_v4(p, v); } while (0)
#else
#define PR_RED(p, v) lk_red_add_v4(p, v)
#endif
template <int LPR, int IB, int OP>
__global__ void __launch_bounds__(PR_WARPS * 32, PR_MINB) link_preagg_ring_kernel(
    const float* __restrict__ fin, const int4* __restrict__ coords, const int* __restrict__ order,
    const int* __restrict__ rank, int64_t n, int q, GenDev g, float* sums) {
  lk_pdl_enter();
  using Cfg = RingCfg<LPR>;
  constexpr int VPL = 2;
  constexpr int G = Cfg::G, RC = Cfg::RC, CR = Cfg::CR, ROW_BYTES = Cfg::ROW_BYTES, SLOT = Cfg::SLOT;
  constexpr int K = (OP == LK_OP_COSX) ? 3 : 2;
  constexpr bool COSX = (OP == LK_OP_COSX);
  constexpr int NP = 4 * IB;
  extern __shared__ __align__(16) uint8_t ring_s[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / LPR;
  const int j = lane % LPR;
  const int kc = K * g.c;
  uint8_t* const wbase = ring_s + (size_t)wib * (PR_STAGES * SLOT);
  const uint32_t wbase_u = (uint32_t)__cvta_generic_to_shared(wbase);
  const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_g * G * (int64_t)q >= n) return;          // whole warp past the end (no CTA-wide barriers below)
  LaneGen<NP> lg;
  load_lane_gen<LPR, IB>(g, j, true, lg);
  const int nchunks = (q + RC - 1) / RC;
  // the pair this lane fetches for a chunk: lane -> (group lane / RC, row lane % RC of the chunk)
  const int pg = lane / RC, pu = lane % RC;
  const int64_t pbase = (warp_g * G + pg) * (int64_t)q + pu;

  auto load_pair = [&](int c, int& o, int& r) {
    const int off = c * RC + pu;
    const int64_t pos = pbase + (int64_t)c * RC;
    const bool ok = lane < CR && off < q && pos < n;
    o = ok ? (order ? __ldg(order + pos) : (int)pos) : 0;
    r = ok ? __ldg(rank + pos) : -1;
  };
  auto issue = [&](int c, int o, int r) {
    const uint32_t slot_u = wbase_u + (c % PR_STAGES) * SLOT;
#pragma unroll
    for (int u = 0; u < RC; ++u) {
      const int p = grp * RC + u;                    // row of the warp chunk
      const int ro = __shfl_sync(0xffffffffu, o, p);
      const int rb = __shfl_sync(0xffffffffu, r, p);
      if (rb >= 0) {
        const float* src = fin + (int64_t)ro * g.c + 4 * j;
#ifndef PR_DEBUG_NOFETCH                             // timing experiments only (results are wrong)
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                       ::"r"(slot_u + p * ROW_BYTES + (i * LPR + j) * 16), "l"(src + 4 * i * LPR) : "memory");
#endif
#ifndef PR_DEBUG_NOCOORD
        if (j == 0)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
                       ::"r"(slot_u + CR * ROW_BYTES + p * 16), "l"(coords + ro) : "memory");
#endif
      }
    }
    if (lane < CR)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot_u + CR * ROW_BYTES + CR * 16 + lane * 4), "r"(r) : "memory");
  };

  // ---- prologue: pairs of the first PR_STAGES chunks at once, rows of the first PR_STAGES - 1 ----
  int po[PR_STAGES], pr[PR_STAGES];
#pragma unroll
  for (int c = 0; c < PR_STAGES; ++c) load_pair(c, po[c], pr[c]);
#pragma unroll
  for (int c = 0; c < PR_STAGES - 1; ++c) {
    if (c < nchunks) issue(c, po[c], pr[c]);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int o_n = po[PR_STAGES - 1], r_n = pr[PR_STAGES - 1];

  float acc[K][VPL][4];
#pragma unroll
  for (int qq = 0; qq < K; ++qq)
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[qq][i][e] = 0.f;
  int cur = -1;

  for (int c = 0; c < nchunks; ++c) {
    const int cn = c + PR_STAGES - 1;
    if (cn < nchunks) issue(cn, o_n, r_n);
    asm volatile("cp.async.commit_group;" ::: "memory");
    load_pair(cn + 1, o_n, r_n);                     // consumed by the next iteration's issue
    asm volatile("cp.async.wait_group %0;" ::"n"(PR_STAGES - 1) : "memory");
    __syncwarp();
    const uint8_t* slot = wbase + (c % PR_STAGES) * SLOT;
#pragma unroll
    for (int u = 0; u < RC; ++u) {
      const int p = grp * RC + u;
      const int b = *(const int*)(slot + CR * ROW_BYTES + CR * 16 + p * 4);
      if (b >= 0) {                                  // else: past the end of the range / of the input
        if (b != cur) {                              // run boundary: flush the finished block
          if (cur >= 0) {
            float* dst = sums + (int64_t)cur * kc + 4 * j;
#pragma unroll
            for (int qq = 0; qq < K; ++qq)
#pragma unroll
              for (int i = 0; i < VPL; ++i)
                PR_RED(dst + qq * g.c + 4 * i * LPR,
                              make_float4(acc[qq][i][0], acc[qq][i][1], acc[qq][i][2], acc[qq][i][3]));
          }
#pragma unroll
          for (int qq = 0; qq < K; ++qq)
#pragma unroll
            for (int i = 0; i < VPL; ++i)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[qq][i][e] = 0.f;
          cur = b;
        }
        const int4 cc = *(const int4*)(slot + CR * ROW_BYTES + p * 16);
        float ph[NP], sn[NP], cs[NP];
        lane_trig_sfu<NP, COSX>(g, lg, cc.x, cc.y, cc.z, ph, sn, cs);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const float4 f4 = *(const float4*)(slot + p * ROW_BYTES + (i * LPR + j) * 16);
          const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int qq = (i % IB) * 4 + e;
            acc[0][i][e] = fmaf(fv[e], (OP == LK_OP_SIN ? sn[qq] : cs[qq]), acc[0][i][e]);
            acc[1][i][e] = fmaf(fv[e], (OP == LK_OP_SIN ? cs[qq] : sn[qq]), acc[1][i][e]);
            if (COSX) acc[K - 1][i][e] = fmaf(fv[e], ph[qq], acc[K - 1][i][e]);
          }
        }
      }
    }
    __syncwarp();                                    // the slot is refilled by the next iteration
  }
  if (cur >= 0) {
    float* dst = sums + (int64_t)cur * kc + 4 * j;
#pragma unroll
    for (int qq = 0; qq < K; ++qq)
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        PR_RED(dst + qq * g.c + 4 * i * LPR,
                      make_float4(acc[qq][i][0], acc[qq][i][1], acc[qq][i][2], acc[qq][i][3]));
  }
}

// zero the first *d_num rows of a [capacity, row_floats] buffer (block sums are allocated for the
// worst case M = N but only the M live rows are ever touched)
__global__ void __launch_bounds__(256) zero_rows_kernel(float4* __restrict__ p,
                                                        const int* __restrict__ d_num,
                                                        int64_t capacity, int row_vec) {
  lk_pdl_enter();
  int64_t m = *d_num;
  if (m > capacity) m = capacity;
  int64_t total = m * row_vec;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x)
    p[t] = make_float4(0.f, 0.f, 0.f, 0.f);
}

extern "C" int lk_zero_rows(float* d_buf, const int32_t* d_num, int64_t capacity, int row_floats,
                            lk_stream_t s) {
  LK_REQUIRE(capacity >= 0 && row_floats > 0 && row_floats % 4 == 0, "lk_zero_rows: bad sizes");
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(d_buf && d_num && (uintptr_t)d_buf % 16 == 0, "lk_zero_rows: bad pointer");
  LK_PDL_LAUNCH_LINK(zero_rows_kernel, lk_grid(capacity * (row_floats / 4), 256, 8), 256, 0, (cudaStream_t)s,
                (float4*)d_buf, d_num, capacity, row_floats / 4);
  LK_LAUNCHED();
  return LK_OK;
}

// ------------------------------------------------------------------ pass 2a: window means
// One warp per block row.  The R neighbour indices are loaded by the first R lanes, the present
// ones are compacted with a ballot, and only those rows are fetched (a LiDAR block has ~9-12 of
// its 27 neighbours), four at a time so the L2 loads overlap.
// SEG: `counts` holds the segment starts of the sorted voxel sequence (n[b] = seg[b+1] - seg[b])
template <bool SEG>
__global__ void __launch_bounds__(256) link_window_mean_kernel(const float* __restrict__ sums,
                                                               const int* __restrict__ counts,
                                                               const int* __restrict__ nbr,
                                                               const int* __restrict__ d_num,
                                                               int64_t capacity, int R, int kc,
                                                               float* __restrict__ mean,
                                                               float* __restrict__ tot_out) {
  lk_pdl_enter();
  int64_t m = *d_num;
  if (m > capacity) m = capacity;
  const int lane = threadIdx.x & 31;
  const int vpr = kc >> 2;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < m; b += warps_total) {
    int src = -1, cnt = 0;
    if (lane < R) {                              // R <= 32
      src = __ldg(nbr + b * R + lane);
      if (src >= 0) cnt = SEG ? __ldg(counts + src + 1) - __ldg(counts + src) : __ldg(counts + src);
    }
    const unsigned present = __ballot_sync(0xffffffffu, src >= 0);
    int tot_i = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot_i += __shfl_xor_sync(0xffffffffu, tot_i, o);
    const float tot = (float)tot_i;
    if (tot_out && lane == 0) tot_out[b] = tot;
    for (int v0 = 0; v0 < vpr; v0 += 32) {
      const int v = v0 + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned rest = present;
      while (rest) {                             // warp-uniform loop over the present neighbours
        int s4[4];
        bool on[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          on[t] = rest != 0;
          const int l = on[t] ? __ffs(rest) - 1 : 0;
          rest &= rest - 1;
          s4[t] = __shfl_sync(0xffffffffu, src, l);
        }
        float4 r4[4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
          r4[t] = (v < vpr && on[t]) ? __ldg((const float4*)(sums + (int64_t)s4[t] * kc) + v)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc.x += r4[t].x; acc.y += r4[t].y; acc.z += r4[t].z; acc.w += r4[t].w;
        }
      }
      if (v < vpr) {
        acc.x /= tot; acc.y /= tot; acc.z /= tot; acc.w /= tot;
        *((float4*)(mean + b * kc) + v) = acc;
      }
    }
  }
}

// ------------------------------------------------------------------ pass 2b: combine (+ norms)

// (AP_MINB, AP_U) = (4, 1): 128 registers, 16 warps / SM, 36.0 us for the block chain vs 38.7 us at the
// round-1 setting (3, 2) (168 registers); (4, 2), (5, 2), (6, 1) spill and measure 42-48 us.
// Every lane group handles U rows per step with all of their loads in flight together: first the
// independent ones (block index, coordinate, local_mix row), then the mean rows that depend on
// the block index.
#ifndef AP_MINB
#define AP_MINB 4                          // resident CTAs per SM the register budget is capped for
#endif
#ifndef AP_U
#define AP_U 1                             // rows per lane group and step
#endif
template <int LPR, int VPL, int IB, int OP, bool NORM>
__global__ void __launch_bounds__(128, VPL >= 4 ? AP_MINB : 5) link_apply_kernel(
    const float* __restrict__ mean, const float* __restrict__ fin, const int4* __restrict__ coords,
    const int* __restrict__ blk, int64_t n, GenDev g, const float* __restrict__ local,
    const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
    const float* __restrict__ b2, float* __restrict__ out) {
  lk_pdl_enter();
  constexpr int G = 32 / LPR;
  constexpr int K = (OP == LK_OP_COSX) ? 3 : 2;
  constexpr bool COSX = (OP == LK_OP_COSX);
  constexpr int NP = 4 * IB;
  constexpr int U = (COSX && VPL > 1) ? 1 : AP_U;     // rows per lane group per step (register budget)
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;
  const int j = lane % LPR;
  const bool active = 4 * j < g.c;
  const int kc = K * g.c;
  LaneGen<NP> lg;
  load_lane_gen<LPR, IB>(g, j, active, lg);
  const float inv_c = 1.0f / (float)g.c;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t steps = (n + G * U - 1) / (G * U);
  for (int64_t step = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; step < steps;
       step += warps_total) {
    int b[U];
    int4 cc[U];
    float4 lv[U][VPL], m0[U][VPL], m1[U][VPL];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {          // independent loads of all U rows
      const int64_t r = (step * U + u) * G + grp;   // NB: whole groups go inactive together;
      ok[u] = r < n;                                //     shuffles stay inside a group
      b[u] = ok[u] ? __ldg(blk + r) : -1;
      cc[u] = ok[u] ? __ldg(coords + r) : make_int4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        lv[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NORM && ok[u] && active) lv[u][i] = lk_ldg_stream((const float4*)(local + r * g.c) + i * LPR + j);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {          // dependent loads: the blocks' window means (L2)
      const bool live = ok[u] && active && b[u] >= 0;
      const float4* mrow = (const float4*)(mean + (int64_t)(live ? b[u] : 0) * kc) + j;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        m0[u][i] = live ? __ldg(mrow + i * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
        m1[u][i] = live ? __ldg(mrow + (g.c >> 2) + i * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = (step * U + u) * G + grp;
      const bool live = ok[u] && active && b[u] >= 0;
      float p[NP], sn[NP], cs[NP];
      lane_trig<NP, COSX>(g, lg, cc[u].x, cc[u].y, cc[u].z, p, sn, cs);
      float v[VPL][4];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float a0[4] = {m0[u][i].x, m0[u][i].y, m0[u][i].z, m0[u][i].w};
        const float a1[4] = {m1[u][i].x, m1[u][i].y, m1[u][i].z, m1[u][i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = (i % IB) * 4 + e;
          // planes [sin, cos]: F[:, :C]*cos - F[:, C:]*sin;  planes [cos, sin]: ... + ...
          v[i][e] = (OP == LK_OP_SIN) ? a0[e] * cs[q] - a1[e] * sn[q] : a0[e] * cs[q] + a1[e] * sn[q];
        }
        if (COSX) {                        // + (mean(F*pos) - F*pos), linkencoder.py:176
          float4 m2 = make_float4(0.f, 0.f, 0.f, 0.f), f = m2;
          if (live) {
            m2 = __ldg((const float4*)(mean + (int64_t)b[u] * kc + 2 * g.c) + i * LPR + j);
            f = lk_ldg_stream((const float4*)(fin + r * g.c) + i * LPR + j);
          }
          v[i][0] += m2.x - f.x * p[(i % IB) * 4 + 0]; v[i][1] += m2.y - f.y * p[(i % IB) * 4 + 1];
          v[i][2] += m2.z - f.z * p[(i % IB) * 4 + 2]; v[i][3] += m2.w - f.w * p[(i % IB) * 4 + 3];
        }
      }
      if (NORM) {
        float l[VPL][4];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          l[i][0] = lv[u][i].x; l[i][1] = lv[u][i].y; l[i][2] = lv[u][i].z; l[i][3] = lv[u][i].w;
        }
        group_layernorm<LPR, VPL>(v, active, inv_c, g1, b1, j);
        group_layernorm<LPR, VPL>(l, active, inv_c, g2, b2, j);
#pragma unroll
        for (int i = 0; i < VPL; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) v[i][e] = fmaxf(v[i][e] + l[i][e], 0.f);
      }
      if (ok[u] && active) {
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          lk_stg_stream((float4*)(out + r * g.c) + i * LPR + j,
                        make_float4(v[i][0], v[i][1], v[i][2], v[i][3]));
      }
    }
  }
}


// ------------------------------------------------------------------ host side

// Lane layout of a C-channel row: (LPR lanes, VPL vectors per lane, IB distinct phase blocks per
// lane).  IB < VPL when the channel groups alias inside a lane: wrows a multiple of 4 LPR.
struct RowLayout { int lpr, vpl, ib; };

// Balanced ring kernel: the grid is sized to the resident capacity (occupancy query, cached per
// instantiation) and every lane group gets q = ceil(n / lane groups) consecutive sorted positions.
template <int LPR, int IB, int OP>
static int launch_ring(const float* d_fin, const int32_t* d_coords, const int32_t* d_order,
                       const int32_t* d_rank, int64_t n, const GenDev& g, float* d_sums,
                       cudaStream_t st) {
  using Cfg = RingCfg<LPR>;
  auto kern = link_preagg_ring_kernel<LPR, IB, OP>;
  // per-device cache (function attributes and occupancy belong to the device's context);
  // 0: not configured yet, < 0: configuration failed.  Racing first calls compute the same value.
  static int ctas_cache[64];
  int dev = 0;
  LK_CUDA(cudaGetDevice(&dev));
  LK_REQUIRE(dev >= 0 && dev < 64, "lk_link_preagg: device ordinal %d out of range", dev);
  if (ctas_cache[dev] == 0) {
    int occ = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PR_WARPS * 32, Cfg::SMEM) != cudaSuccess ||
        occ < 1) {
      ctas_cache[dev] = -1;
      (void)cudaGetLastError();
    } else {
      ctas_cache[dev] = occ;
    }
  }
  const int ctas_per_sm = ctas_cache[dev];
  LK_REQUIRE(ctas_per_sm > 0, "lk_link_preagg: cannot configure the ring kernel (shared memory %d bytes)",
             Cfg::SMEM);
  static const int q_env = [] {
    const char* e = getenv("LINKB200_PREAGG_Q");     // tuning knob: positions per lane group
    return e ? atoi(e) : 0;
  }();
  const int64_t groups_cap = (int64_t)LK_SM_COUNT * ctas_per_sm * PR_WARPS * Cfg::G;
  int64_t q = (n + groups_cap - 1) / groups_cap;
  if (q < 8) q = 8;
  if (q_env > 0 && q_env >= q) q = q_env;
  const int64_t groups = (n + q - 1) / q;
  const int grid = (int)((groups + (int64_t)Cfg::G * PR_WARPS - 1) / ((int64_t)Cfg::G * PR_WARPS));
  LK_PDL_LAUNCH_LINK(kern, grid, PR_WARPS * 32, Cfg::SMEM, st, d_fin, (const int4*)d_coords, d_order, d_rank, n, (int)q, g, d_sums);
  LK_LAUNCHED();
  return LK_OK;
}

static int launch_preagg(const float* d_fin, const int32_t* d_coords, const int32_t* d_order,
                         const int32_t* d_rank, int64_t n, const GenDev& g, float* d_sums,
                         cudaStream_t st) {
  if (g.c == 16 || g.c == 32 || g.c == 64 || g.c == 128) {
    // shared-memory staged kernels: 2 vectors per lane, LPR = C / 8 lanes per row
    const int lpr = g.c / 8, span = 4 * lpr;
    const int ib = (g.wrows % span == 0 && g.wrows / span == 1) ? 1 : 2;
    static const int use_ring = [] {
      const char* e = getenv("LINKB200_PREAGG");     // tuning knob: "smem" = one step per warp
      return !(e && e[0] == 's');
    }();
    if (use_ring && !g.accurate) {
      int rc = LK_EINVAL;
#define RING_IB(LPRV, IBV)                                                                        \
  do {                                                                                            \
    if (g.op == LK_OP_COS) rc = launch_ring<LPRV, IBV, LK_OP_COS>(d_fin, d_coords, d_order, d_rank, n, g, d_sums, st);      \
    else if (g.op == LK_OP_SIN) rc = launch_ring<LPRV, IBV, LK_OP_SIN>(d_fin, d_coords, d_order, d_rank, n, g, d_sums, st); \
    else rc = launch_ring<LPRV, IBV, LK_OP_COSX>(d_fin, d_coords, d_order, d_rank, n, g, d_sums, st);                       \
  } while (0)
#define RING(LPRV)                   \
  do {                               \
    if (ib == 1) RING_IB(LPRV, 1);   \
    else RING_IB(LPRV, 2);           \
  } while (0)
      if (lpr == 2) RING(2);
      else if (lpr == 4) RING(4);
      else if (lpr == 8) RING(8);
      else RING(16);
#undef RING
#undef RING_IB
      return rc;
    }
    const int64_t rows_per_warp = (int64_t)(32 / lpr) * PS_RPG;
    const int64_t warps = (n + rows_per_warp - 1) / rows_per_warp;
    const int grid = (int)((warps + PS_WARPS - 1) / PS_WARPS);
#define LAUNCH_PS_OP(LPRV, IBV, O)                                                              \
  link_preagg_smem_kernel<LPRV, IBV, O><<<grid, PS_WARPS * 32, 0, st>>>(                        \
      d_fin, (const int4*)d_coords, d_order, d_rank, n, g, d_sums)
#define LAUNCH_PS_IB(LPRV, IBV)                                      \
  do {                                                               \
    if (g.op == LK_OP_COS) LAUNCH_PS_OP(LPRV, IBV, LK_OP_COS);       \
    else if (g.op == LK_OP_SIN) LAUNCH_PS_OP(LPRV, IBV, LK_OP_SIN);  \
    else LAUNCH_PS_OP(LPRV, IBV, LK_OP_COSX);                        \
  } while (0)
#define LAUNCH_PS(LPRV)                       \
  do {                                        \
    if (ib == 1) LAUNCH_PS_IB(LPRV, 1);       \
    else LAUNCH_PS_IB(LPRV, 2);               \
  } while (0)
    if (lpr == 2) LAUNCH_PS(2);
    else if (lpr == 4) LAUNCH_PS(4);
    else if (lpr == 8) LAUNCH_PS(8);
    else LAUNCH_PS(16);
#undef LAUNCH_PS
#undef LAUNCH_PS_IB
#undef LAUNCH_PS_OP
    LK_LAUNCHED();
    return LK_OK;
  }
  RowLayout L;                                   // any other C % 4 == 0: register-staged kernel
  L.vpl = 1; L.ib = 1; L.lpr = 1;
  while (L.lpr < g.c / 4) L.lpr <<= 1;
  const int rpg = L.lpr > 8 ? L.lpr : 8;
  const int64_t per_warp = (int64_t)(32 / L.lpr) * rpg;
  const int64_t warps = (n + per_warp - 1) / per_warp;
  const int grid = (int)((warps + 3) / 4);
#define LAUNCH_PRE_OP(LPRV, O)                                                               \
  link_preagg_kernel<LPRV, 1, 1, O><<<grid, 128, 0, st>>>(d_fin, (const int4*)d_coords,       \
                                                          d_order, d_rank, n, g, d_sums)
#define LAUNCH_PRE(LPRV)                                      \
  do {                                                        \
    if (g.op == LK_OP_COS) LAUNCH_PRE_OP(LPRV, LK_OP_COS);    \
    else if (g.op == LK_OP_SIN) LAUNCH_PRE_OP(LPRV, LK_OP_SIN); \
    else LAUNCH_PRE_OP(LPRV, LK_OP_COSX);                     \
  } while (0)
  switch (L.lpr) {
    case 1: LAUNCH_PRE(1); break;
    case 2: LAUNCH_PRE(2); break;
    case 4: LAUNCH_PRE(4); break;
    case 8: LAUNCH_PRE(8); break;
    case 16: LAUNCH_PRE(16); break;
    default: LAUNCH_PRE(32); break;
  }
#undef LAUNCH_PRE
#undef LAUNCH_PRE_OP
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_link_preagg_fwd(const float* d_fin, const int32_t* d_coords, const int32_t* d_blk,
                                  int64_t n, const lk_kernelgen_t* gen, float* d_sums,
                                  lk_stream_t s) {
  GenDev g;
  int rc = check_gen(gen, &g, "lk_link_preagg_fwd");
  if (rc) return rc;
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_fin && d_coords && d_blk && d_sums && n > 0, "lk_link_preagg_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_fin % 16 == 0 && (uintptr_t)d_sums % 16 == 0,
             "lk_link_preagg_fwd: feature buffers must be 16-byte aligned");
  return launch_preagg(d_fin, d_coords, nullptr, d_blk, n, g, d_sums, (cudaStream_t)s);
}

extern "C" int lk_link_preagg_seg_fwd(const float* d_fin, const int32_t* d_coords,
                                      const int32_t* d_order, const int32_t* d_sorted_rank,
                                      int64_t n, const lk_kernelgen_t* gen, float* d_sums,
                                      lk_stream_t s) {
  GenDev g;
  int rc = check_gen(gen, &g, "lk_link_preagg_seg_fwd");
  if (rc) return rc;
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_fin && d_coords && d_order && d_sorted_rank && d_sums && n > 0,
             "lk_link_preagg_seg_fwd: null pointer");
  LK_REQUIRE((uintptr_t)d_fin % 16 == 0 && (uintptr_t)d_sums % 16 == 0,
             "lk_link_preagg_seg_fwd: feature buffers must be 16-byte aligned");
  return launch_preagg(d_fin, d_coords, d_order, d_sorted_rank, n, g, d_sums, (cudaStream_t)s);
}

extern "C" int lk_link_window_mean(const float* d_sums, const int32_t* d_counts,
                                   const int32_t* d_nbr, const int32_t* d_num, int64_t capacity,
                                   int r3, int kc, float* d_mean, lk_stream_t s) {
  LK_REQUIRE(capacity >= 0 && r3 > 0 && r3 <= 32 && kc > 0 && kc % 4 == 0,
             "lk_link_window_mean: bad sizes (needs r^3 <= 32)");
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(d_sums && d_counts && d_nbr && d_num && d_mean, "lk_link_window_mean: null pointer");
  LK_PDL_LAUNCH_LINK(link_window_mean_kernel<false>, lk_grid(capacity * 32, 256, 8), 256, 0, (cudaStream_t)s,
                d_sums, d_counts, d_nbr, d_num, capacity, r3, kc, d_mean, (float*)nullptr);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_link_window_mean_tot(const float* d_sums, const int32_t* d_seg, const int32_t* d_nbr,
                                       const int32_t* d_num, int64_t capacity, int r3, int kc,
                                       float* d_mean, float* d_tot, lk_stream_t s) {
  LK_REQUIRE(capacity >= 0 && r3 > 0 && r3 <= 32 && kc > 0 && kc % 4 == 0,
             "lk_link_window_mean_seg: bad sizes (needs r^3 <= 32)");
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(d_sums && d_seg && d_nbr && d_num && d_mean, "lk_link_window_mean_seg: null pointer");
  LK_PDL_LAUNCH_LINK(link_window_mean_kernel<true>, lk_grid(capacity * 32, 256, 8), 256, 0, (cudaStream_t)s,
                d_sums, d_seg, d_nbr, d_num, capacity, r3, kc, d_mean, d_tot);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_link_window_mean_seg(const float* d_sums, const int32_t* d_seg, const int32_t* d_nbr,
                                       const int32_t* d_num, int64_t capacity, int r3, int kc,
                                       float* d_mean, lk_stream_t s) {
  return lk_link_window_mean_tot(d_sums, d_seg, d_nbr, d_num, capacity, r3, kc, d_mean, nullptr, s);
}

extern "C" int lk_link_apply_fwd(const float* d_mean, const float* d_fin, const int32_t* d_coords,
                                 const int32_t* d_blk, int64_t n, const lk_kernelgen_t* gen,
                                 int fuse_norm, const float* d_local, const float* d_g1,
                                 const float* d_b1, const float* d_g2, const float* d_b2,
                                 float* d_out, lk_stream_t s) {
  GenDev g;
  int rc = check_gen(gen, &g, "lk_link_apply_fwd");
  if (rc) return rc;
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_mean && d_coords && d_blk && d_out && n > 0, "lk_link_apply_fwd: null pointer");
  LK_REQUIRE(g.op != LK_OP_COSX || d_fin, "lk_link_apply_fwd: cos_x needs the input features");
  LK_REQUIRE(!fuse_norm || (d_local && d_g1 && d_b1 && d_g2 && d_b2),
             "lk_link_apply_fwd: fused norms need local features and both LayerNorm parameters");
  cudaStream_t st = (cudaStream_t)s;
  if (g.c == 16 || g.c == 32 || g.c == 64 || g.c == 128) {
    // 4 vectors (16 channels) per lane, LPR = C / 16 lanes per row: fewest instructions per row
    // (LayerNorm trees of log2(LPR) steps); a shared-memory staged variant with 2 vectors per lane
    // measured no faster (the kernel is issue-bound, ~2 000 thread instructions per row)
    const int lpr = g.c / 16, span = 4 * lpr;
    const int ibr = (g.wrows % span == 0) ? g.wrows / span : 4;
    const int ib = (ibr == 1 || ibr == 2) ? ibr : 4;
    const int rows_per_step = (32 / lpr) * (g.op == LK_OP_COSX ? 1 : AP_U);
    const int64_t steps = (n + rows_per_step - 1) / rows_per_step;
    const int grid = lk_grid(steps * 32, 128, AP_MINB);
#define LAUNCH_A4_ON(LPRV, IBV, O, NRM)                                                        \
  LK_PDL_LAUNCH_LINK((link_apply_kernel<LPRV, 4, IBV, O, NRM>), grid, 128, 0, st,                      \
                d_mean, d_fin, (const int4*)d_coords, d_blk, n, g, d_local, d_g1, d_b1, d_g2, d_b2, d_out)
#define LAUNCH_A4_O(LPRV, IBV, O)                           \
  do {                                                      \
    if (fuse_norm) LAUNCH_A4_ON(LPRV, IBV, O, true);        \
    else LAUNCH_A4_ON(LPRV, IBV, O, false);                 \
  } while (0)
#define LAUNCH_A4_IB(LPRV, IBV)                                       \
  do {                                                                \
    if (g.op == LK_OP_COS) LAUNCH_A4_O(LPRV, IBV, LK_OP_COS);         \
    else if (g.op == LK_OP_SIN) LAUNCH_A4_O(LPRV, IBV, LK_OP_SIN);    \
    else LAUNCH_A4_O(LPRV, IBV, LK_OP_COSX);                          \
  } while (0)
#define LAUNCH_A4(LPRV)                      \
  do {                                       \
    if (ib == 1) LAUNCH_A4_IB(LPRV, 1);      \
    else if (ib == 2) LAUNCH_A4_IB(LPRV, 2); \
    else LAUNCH_A4_IB(LPRV, 4);              \
  } while (0)
    if (lpr == 1) LAUNCH_A4(1);
    else if (lpr == 2) LAUNCH_A4(2);
    else if (lpr == 4) LAUNCH_A4(4);
    else LAUNCH_A4(8);
#undef LAUNCH_A4
#undef LAUNCH_A4_IB
#undef LAUNCH_A4_O
#undef LAUNCH_A4_ON
    LK_LAUNCHED();
    return LK_OK;
  }
  RowLayout L;                                   // any other C % 4 == 0: register-staged kernel
  L.vpl = 1; L.ib = 1; L.lpr = 1;
  while (L.lpr < g.c / 4) L.lpr <<= 1;
  const int rows_per_step = (32 / L.lpr) * 2;
  const int64_t steps = (n + rows_per_step - 1) / rows_per_step;
  const int grid = lk_grid(steps * 32, 128, 5);
#define LAUNCH_APPLY_ON(LPRV, O, NRM)                                                          \
  LK_PDL_LAUNCH_LINK((link_apply_kernel<LPRV, 1, 1, O, NRM>), grid, 128, 0, st,                        \
                d_mean, d_fin, (const int4*)d_coords, d_blk, n, g, d_local, d_g1, d_b1, d_g2, d_b2, d_out)
#define LAUNCH_APPLY_O(LPRV, O)                         \
  do {                                                  \
    if (fuse_norm) LAUNCH_APPLY_ON(LPRV, O, true);      \
    else LAUNCH_APPLY_ON(LPRV, O, false);               \
  } while (0)
#define LAUNCH_APPLY(LPRV)                                      \
  do {                                                          \
    if (g.op == LK_OP_COS) LAUNCH_APPLY_O(LPRV, LK_OP_COS);     \
    else if (g.op == LK_OP_SIN) LAUNCH_APPLY_O(LPRV, LK_OP_SIN); \
    else LAUNCH_APPLY_O(LPRV, LK_OP_COSX);                      \
  } while (0)
  switch (L.lpr) {
    case 1: LAUNCH_APPLY(1); break;
    case 2: LAUNCH_APPLY(2); break;
    case 4: LAUNCH_APPLY(4); break;
    case 8: LAUNCH_APPLY(8); break;
    case 16: LAUNCH_APPLY(16); break;
    default: LAUNCH_APPLY(32); break;
  }
#undef LAUNCH_APPLY
#undef LAUNCH_APPLY_O
#undef LAUNCH_APPLY_ON
  LK_LAUNCHED();
  return LK_OK;
}
