// Hand-written backward of the linear-kernel path with the fused norms (reference:
// autograd through index / devoxelize_backward (devoxelize_cuda.cu:38-59, float atomics) / the [N,kC]
// cat temporaries / voxelize_backward (voxelize_cuda.cu:28-42); glue devoxelize.py:75-98).
//
//   forward (link.cu)  : S[b] = sum_{i in b} F_i {u_0, u_1}(p_i);  A[b] = (sum_{b' in N(b)} S[b']) / T[b];
//                        y_i = A_0 w_0(p_i) + A_1 w_1(p_i);  out_i = relu(LN(y_i) + LN(local_i))
//   link_bwd_norm      : per voxel i   recompute y_i, both LayerNorms, the ReLU mask ->
//                                      dy_i, dlocal_i, d(gamma, beta) of both norms
//                        per block b   G[b] = (sum_{i in b} [dy_i w_0(p_i), dy_i w_1(p_i)]) / T[b]
//   link_bwd_apply     : per block b'  dS[b'] = sum_{b in N^T(b')} G[b]
//                        per voxel j   dF_j = dS_0 u_0(p_j) + dS_1 u_1(p_j)
//                                      dp_j = F_j (dS_0 u_0'(p_j) + dS_1 u_1'(p_j)) + dy_j (A_0 w_0'(p_j) + A_1 w_1'(p_j))
//                                      dW  += dp_j x_j^T   (summed over the channel groups)
//
// One warp owns one block at a time (blocks are visited round-robin, M is read from the device).
// The window row (2C floats) is reduced by the whole warp -- lane v owns float4 v of the row, so a
// neighbour row is one fully coalesced 128-bit load per lane -- and handed to the voxel phase
// through shared memory; the voxels of a block are contiguous in the sorted sequence (seg / order
// from lk_sort_unique_ex), LPR = C/8 lanes own a row (two float4 each), 32/LPR voxels side by side.
// Because a block belongs to exactly one warp the backward block sums need no atomics and are
// deterministic.  (A forward kernel of the same shape -- window mean + apply fused, block- and
// chunk-centric -- was measured and dropped: 32 / 49 us against 6.5 + 9.1 / 6.5 + 19.2 us for the
// two-kernel form at N = 119k, C = 64; these kernels are issue-bound, not bound by the [M,kC] round
// trip the fusion saves.  profiles/r02_path_fused_ab.txt.)
#include <stdlib.h>

#include "link_common.cuh"

#define WA_WARPS 8

// float4 helpers
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// Window row of block b: acc[w] = sum over the present neighbours (ascending offset index) of
// float4 (lane + 32 w) of their rows in `rows` [*, KC4 float4].  `src` = this lane's neighbour
// (lane < R), `present` = ballot(src >= 0).
template <int KC4, int NV, int CH>
__device__ __forceinline__ void window_row(const float4* __restrict__ rows, int src, unsigned present,
                                           int lane, float4 acc[NV]) {
#pragma unroll
  for (int w = 0; w < NV; ++w) acc[w] = f4zero();
  unsigned rest = present;
  while (rest) {                                   // warp-uniform
    int s_[CH];
    bool on[CH];
#pragma unroll
    for (int t = 0; t < CH; ++t) {
      on[t] = rest != 0;
      const int l = on[t] ? __ffs(rest) - 1 : 0;
      rest &= rest - 1;
      s_[t] = __shfl_sync(0xffffffffu, src, l);
    }
    float4 r_[CH][NV];
#pragma unroll
    for (int t = 0; t < CH; ++t)
#pragma unroll
      for (int w = 0; w < NV; ++w) {
        const int v = lane + 32 * w;
        r_[t][w] = (on[t] && v < KC4) ? __ldg(rows + (int64_t)s_[t] * KC4 + v) : f4zero();
      }
#pragma unroll
    for (int t = 0; t < CH; ++t)
#pragma unroll
      for (int w = 0; w < NV; ++w) f4add(acc[w], r_[t][w]);
  }
}

// ------------------------------------------------------------------ backward
// plane weights of the two-plane ops (forward: g_q = F u_q(p), y = sum_q A_q w_q(p)):
//   cos: u = (cos, sin), w = (cos,  sin)      sin: u = (sin, cos), w = (cos, -sin)
template <int OP>
struct PlaneW {
  static __device__ __forceinline__ float u0(float sn, float cs) { return OP == LK_OP_SIN ? sn : cs; }
  static __device__ __forceinline__ float u1(float sn, float cs) { return OP == LK_OP_SIN ? cs : sn; }
  static __device__ __forceinline__ float du0(float sn, float cs) { return OP == LK_OP_SIN ? cs : -sn; }
  static __device__ __forceinline__ float du1(float sn, float cs) { return OP == LK_OP_SIN ? -sn : cs; }
  static __device__ __forceinline__ float w0(float sn, float cs) { return cs; }
  static __device__ __forceinline__ float w1(float sn, float cs) { return OP == LK_OP_SIN ? -sn : sn; }
  static __device__ __forceinline__ float dw0(float sn, float cs) { return -sn; }
  static __device__ __forceinline__ float dw1(float sn, float cs) { return OP == LK_OP_SIN ? -cs : cs; }
};

__device__ __forceinline__ void f4load(float v[4], const float4& f) { v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
// L1-cached 128-bit load the compiler may not hoist out of a loop (keeps loop-invariant LayerNorm
// parameters out of the register file: 32 registers in the backward kernel)
__device__ __forceinline__ void f4load_pinned(float v[4], const float* p) {
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}

// LayerNorm forward of a row spread over LPR lanes x 2 vectors that keeps what the backward needs:
// x <- xhat (normalised, before the affine), returns rstd
template <int LPR>
__device__ __forceinline__ float group_ln_stats(float x[2][4], float inv_c) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) s += (x[i][0] + x[i][1]) + (x[i][2] + x[i][3]);
  const float mean = group_sum<LPR>(s) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) { x[i][e] -= mean; q = fmaf(x[i][e], x[i][e], q); }
  const float rstd = rsqrtf(group_sum<LPR>(q) * inv_c + 1e-6f);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) x[i][e] *= rstd;
  return rstd;
}

// LayerNorm backward: d <- dx given d = dL/d(LN output), xhat, rstd, gamma
template <int LPR>
__device__ __forceinline__ void group_ln_bwd(float d[2][4], const float xhat[2][4], float rstd,
                                             const float* gam_lane /* gamma + 4 j */, float inv_c) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float gam[4];
    f4load_pinned(gam, gam_lane + 4 * i * LPR);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      d[i][e] *= gam[e];
      s1 += d[i][e];
      s2 = fmaf(d[i][e], xhat[i][e], s2);
    }
  }
  const float m1 = group_sum<LPR>(s1) * inv_c, m2 = group_sum<LPR>(s2) * inv_c;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) d[i][e] = rstd * (d[i][e] - m1 - xhat[i][e] * m2);
}

struct VoxB1 {
  int r, cx, cy, cz;
  float4 lv[2], dv[2];
};

// Backward, kernel 1 (per block, per voxel): LayerNorm / ReLU backward with everything recomputed from
// the saved window means, then the block sums of the weighted incoming gradient, scaled by 1/T[b].
//   dy, dlocal [n,C]; gsum [cap, 2C] = G[b]; dparam [4,C] += (dgamma1, dbeta1, dgamma2, dbeta2)
template <int LPR, int IB, int OP>
__global__ void __launch_bounds__(WA_WARPS * 32, 2) link_bwd_norm_kernel(
    const float4* __restrict__ mean, const float* __restrict__ tot, const int* __restrict__ seg,
    const int* __restrict__ order, const int* __restrict__ d_num, int64_t capacity,
    const int4* __restrict__ coords, GenDev g, const float* __restrict__ local,
    const float* __restrict__ dout, const float* __restrict__ g1, const float* __restrict__ b1,
    const float* __restrict__ g2, const float* __restrict__ b2, float* __restrict__ dy,
    float* __restrict__ dlocal, float4* __restrict__ gsum, float* __restrict__ dparam) {
  constexpr int K = 2;
  constexpr int C = 8 * LPR, C4 = C / 4, KC4 = K * C4;
  constexpr int G = 32 / LPR;
  constexpr int NP = 4 * IB;
  // per (warp, lane group) partial sums of the four affine-parameter gradients: [4][C] floats
  __shared__ float4 par_s[WA_WARPS][G][3][C4];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / LPR, j = lane % LPR;
  for (int t = threadIdx.x; t < WA_WARPS * G * 3 * C4; t += blockDim.x) (&par_s[0][0][0][0])[t] = f4zero();
  __syncthreads();
  int64_t m = *d_num;
  if (m > capacity) m = capacity;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  LaneGen<NP> lg;
  load_lane_gen<LPR, IB>(g, j, true, lg);
  const float inv_c = 1.0f / (float)C;

  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < m; b += warps_total) {
    const int s0 = __ldg(seg + b), s1 = __ldg(seg + b + 1);
    const int nvox = s1 - s0;
    const float inv_t = 1.0f / __ldg(tot + b);
    const float* arow = (const float*)(mean + b * KC4) + 4 * j;    // re-read per voxel (L1 resident)
    int myord = s0 + lane < s1 ? __ldg(order + s0 + lane) : -1;
    VoxB1 cur, nxt;
    auto load_batch = [&](VoxB1& d, int t0) {
      d.r = __shfl_sync(0xffffffffu, myord, (t0 + grp) & 31);
      const int4 c4 = d.r >= 0 ? __ldg(coords + d.r) : make_int4(0, 0, 0, 0);
      d.cx = c4.x; d.cy = c4.y; d.cz = c4.z;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        d.lv[i] = d.r >= 0 ? lk_ldg_stream((const float4*)(local + (int64_t)d.r * C) + i * LPR + j) : f4zero();
        d.dv[i] = d.r >= 0 ? lk_ldg_stream((const float4*)(dout + (int64_t)d.r * C) + i * LPR + j) : f4zero();
      }
    };
    load_batch(cur, 0);
    nxt = cur;
    float dA[K][2][4];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) dA[q][i][e] = 0.f;

    for (int t = 0; t < nvox; t += G) {
      const int tn = t + G;
      if (tn < nvox) {
        if ((tn & 31) == 0) myord = s0 + tn + lane < s1 ? __ldg(order + s0 + tn + lane) : -1;
        load_batch(nxt, tn & 31);
      }
      float p[NP], sn[NP], cs[NP];
      lane_trig<NP, false>(g, lg, cur.cx, cur.cy, cur.cz, p, sn, cs);
      float y[2][4], l[2][4], d1[2][4], d2[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float a0[4], a1[4];
        f4load_pinned(a0, arow + 4 * i * LPR); f4load_pinned(a1, arow + C + 4 * i * LPR);
        f4load(l[i], cur.lv[i]);
        f4load(d1[i], cur.dv[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = (i % IB) * 4 + e;
          y[i][e] = a0[e] * PlaneW<OP>::w0(sn[q], cs[q]) + a1[e] * PlaneW<OP>::w1(sn[q], cs[q]);
        }
      }
      const float rstd1 = group_ln_stats<LPR>(y, inv_c);      // y, l <- xhat
      const float rstd2 = group_ln_stats<LPR>(l, inv_c);
      const bool live = cur.r >= 0;
      float4 pg1[2], pb[2], pg2[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float gam1[4], bet1[4], gam2[4], bet2[4];
        f4load_pinned(gam1, g1 + 4 * (i * LPR + j)); f4load_pinned(bet1, b1 + 4 * (i * LPR + j));
        f4load_pinned(gam2, g2 + 4 * (i * LPR + j)); f4load_pinned(bet2, b2 + 4 * (i * LPR + j));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pre = fmaf(y[i][e], gam1[e], bet1[e]) + fmaf(l[i][e], gam2[e], bet2[e]);
          const float d = (live && pre > 0.f) ? d1[i][e] : 0.f;     // ReLU mask
          d1[i][e] = d;
          d2[i][e] = d;
        }
        pg1[i] = make_float4(d1[i][0] * y[i][0], d1[i][1] * y[i][1], d1[i][2] * y[i][2], d1[i][3] * y[i][3]);
        pg2[i] = make_float4(d1[i][0] * l[i][0], d1[i][1] * l[i][1], d1[i][2] * l[i][2], d1[i][3] * l[i][3]);
        pb[i] = make_float4(d1[i][0], d1[i][1], d1[i][2], d1[i][3]);
      }
      if (live) {                                  // every lane owns its own shared-memory cells
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float4* cell = &par_s[wib][grp][0][i * LPR + j];
          float4 a;
          a = cell[0];      f4add(a, pg1[i]); cell[0] = a;
          a = cell[C4];     f4add(a, pb[i]);  cell[C4] = a;
          a = cell[2 * C4]; f4add(a, pg2[i]); cell[2 * C4] = a;
        }
      }
      group_ln_bwd<LPR>(d1, y, rstd1, g1 + 4 * j, inv_c);      // d1 <- dy, d2 <- dlocal
      group_ln_bwd<LPR>(d2, l, rstd2, g2 + 4 * j, inv_c);
      if (live) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          lk_stg_stream((float4*)(dy + (int64_t)cur.r * C) + i * LPR + j, make_float4(d1[i][0], d1[i][1], d1[i][2], d1[i][3]));
          lk_stg_stream((float4*)(dlocal + (int64_t)cur.r * C) + i * LPR + j, make_float4(d2[i][0], d2[i][1], d2[i][2], d2[i][3]));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int q = (i % IB) * 4 + e;
            dA[0][i][e] = fmaf(d1[i][e], PlaneW<OP>::w0(sn[q], cs[q]), dA[0][i][e]);
            dA[1][i][e] = fmaf(d1[i][e], PlaneW<OP>::w1(sn[q], cs[q]), dA[1][i][e]);
          }
        }
      }
      cur = nxt;
    }
    // join the G lane groups (fixed order: deterministic) and store G[b] = dA / T
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = dA[q][i][e];
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          dA[q][i][e] = v * inv_t;
        }
    if (grp == 0) {
#pragma unroll
      for (int q = 0; q < K; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
          gsum[b * KC4 + q * C4 + i * LPR + j] = make_float4(dA[q][i][0], dA[q][i][1], dA[q][i][2], dA[q][i][3]);
    }
  }
  // ---- CTA reduction of the affine-parameter gradients, one atomic per CTA and element ----
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * C; t += blockDim.x) {
    const int par = t / C, ch = t % C;
    float acc = 0.f;
    for (int w = 0; w < WA_WARPS; ++w)
      for (int gq = 0; gq < G; ++gq) acc += ((const float*)&par_s[w][gq][par][0])[ch];
    // par_s planes: 0 dgamma1, 1 dbeta (shared by both norms), 2 dgamma2
    if (par == 0) atomicAdd(dparam + ch, acc);
    else if (par == 1) { atomicAdd(dparam + C + ch, acc); atomicAdd(dparam + 3 * C + ch, acc); }
    else atomicAdd(dparam + 2 * C + ch, acc);
  }
}

struct VoxB2 {
  int r, cx, cy, cz;
  float4 fv[2], dv[2];
};

// Backward, kernel 2 (per block, per voxel): dS[b'] = sum over the transposed neighbourhood of G, then
// dF and the phase gradient -> d(pos_weight).  dfin [n,C]; dw [wrows,3] += .
template <int LPR, int IB, int OP>
__global__ void __launch_bounds__(WA_WARPS * 32, 2) link_bwd_apply_kernel(
    const float4* __restrict__ gsum, const float4* __restrict__ mean, const int* __restrict__ nbr_t,
    const int* __restrict__ seg, const int* __restrict__ order, const int* __restrict__ d_num,
    int64_t capacity, int R, const int4* __restrict__ coords, GenDev g, const float* __restrict__ fin,
    const float* __restrict__ dy, float* __restrict__ dfin, float* __restrict__ dw) {
  constexpr int K = 2;
  constexpr int C = 8 * LPR, C4 = C / 4, KC4 = K * C4;
  constexpr int NV = (KC4 + 31) / 32;
  constexpr int CH = NV == 1 ? 8 : 4;
  constexpr int G = 32 / LPR;
  constexpr int NP = 4 * IB;
  __shared__ float4 arow_s[WA_WARPS][KC4];
  // per (warp, lane group, lane of the group) partial sums of dp x^T: [3 axes][NP phases]
  __shared__ float dw_s[WA_WARPS][32][3][NP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane / LPR, j = lane % LPR;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int q = 0; q < NP; ++q) dw_s[wib][lane][a][q] = 0.f;
  int64_t m = *d_num;
  if (m > capacity) m = capacity;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  LaneGen<NP> lg;
  load_lane_gen<LPR, IB>(g, j, true, lg);

  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < m; b += warps_total) {
    const int src = lane < R ? __ldg(nbr_t + b * R + lane) : -1;
    const int s0 = __ldg(seg + b), s1 = __ldg(seg + b + 1);
    const int nvox = s1 - s0;
    const unsigned present = __ballot_sync(0xffffffffu, src >= 0);
    int myord = s0 + lane < s1 ? __ldg(order + s0 + lane) : -1;
    VoxB2 cur, nxt;
    auto load_batch = [&](VoxB2& d, int t0) {
      d.r = __shfl_sync(0xffffffffu, myord, (t0 + grp) & 31);
      const int4 c4 = d.r >= 0 ? __ldg(coords + d.r) : make_int4(0, 0, 0, 0);
      d.cx = c4.x; d.cy = c4.y; d.cz = c4.z;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        d.fv[i] = d.r >= 0 ? lk_ldg_stream((const float4*)(fin + (int64_t)d.r * C) + i * LPR + j) : f4zero();
        d.dv[i] = d.r >= 0 ? lk_ldg_stream((const float4*)(dy + (int64_t)d.r * C) + i * LPR + j) : f4zero();
      }
    };
    load_batch(cur, 0);
    nxt = cur;
    float4 acc[NV];
    window_row<KC4, NV, CH>(gsum, src, present, lane, acc);
#pragma unroll
    for (int w = 0; w < NV; ++w) {
      const int v = lane + 32 * w;
      if (v < KC4) arow_s[wib][v] = acc[w];
    }
    __syncwarp();
    float4 dS[K][2], A[K][2];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        dS[q][i] = arow_s[wib][q * C4 + i * LPR + j];
        A[q][i] = __ldg(mean + b * KC4 + q * C4 + i * LPR + j);
      }
    float wx[NP], wy[NP], wz[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) wx[q] = wy[q] = wz[q] = 0.f;

    for (int t = 0; t < nvox; t += G) {
      const int tn = t + G;
      if (tn < nvox) {
        if ((tn & 31) == 0) myord = s0 + tn + lane < s1 ? __ldg(order + s0 + tn + lane) : -1;
        load_batch(nxt, tn & 31);
      }
      float p[NP], sn[NP], cs[NP], dp[NP];
      lane_trig<NP, false>(g, lg, cur.cx, cur.cy, cur.cz, p, sn, cs);
#pragma unroll
      for (int q = 0; q < NP; ++q) dp[q] = 0.f;
      float df[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float s0v[4], s1v[4], a0[4], a1[4], f[4], d[4];
        f4load(s0v, dS[0][i]); f4load(s1v, dS[1][i]);
        f4load(a0, A[0][i]); f4load(a1, A[1][i]);
        f4load(f, cur.fv[i]); f4load(d, cur.dv[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = (i % IB) * 4 + e;
          df[i][e] = s0v[e] * PlaneW<OP>::u0(sn[q], cs[q]) + s1v[e] * PlaneW<OP>::u1(sn[q], cs[q]);
          dp[q] += f[e] * (s0v[e] * PlaneW<OP>::du0(sn[q], cs[q]) + s1v[e] * PlaneW<OP>::du1(sn[q], cs[q])) +
                   d[e] * (a0[e] * PlaneW<OP>::dw0(sn[q], cs[q]) + a1[e] * PlaneW<OP>::dw1(sn[q], cs[q]));
        }
      }
      if (cur.r >= 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          lk_stg_stream((float4*)(dfin + (int64_t)cur.r * C) + i * LPR + j, make_float4(df[i][0], df[i][1], df[i][2], df[i][3]));
        const float x = (float)cur.cx, yv = (float)cur.cy, z = (float)cur.cz;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          wx[q] = fmaf(dp[q], x, wx[q]); wy[q] = fmaf(dp[q], yv, wy[q]); wz[q] = fmaf(dp[q], z, wz[q]);
        }
      }
      cur = nxt;
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      dw_s[wib][lane][0][q] += wx[q]; dw_s[wib][lane][1][q] += wy[q]; dw_s[wib][lane][2][q] += wz[q];
    }
    __syncwarp();                                  // arow_s is rewritten by the next block
  }
  // ---- CTA reduction: phase slot (ib, e) of lane j <-> channel 4 (ib LPR + j) + e <-> W row (channel % wrows) ----
  __syncthreads();
  for (int t = threadIdx.x; t < LPR * NP * 3; t += blockDim.x) {
    const int a = t % 3, q = (t / 3) % NP, jj = t / (3 * NP);
    float acc = 0.f;
    for (int w = 0; w < WA_WARPS; ++w)
      for (int gq = 0; gq < G; ++gq) acc += dw_s[w][gq * LPR + jj][a][q];
    const int ch = 4 * ((q / 4) * LPR + jj) + (q % 4);
    atomicAdd(dw + (ch % g.wrows) * 3 + a, acc);
  }
}

// persistent grid: resident CTAs per SM x SM count (cached per device), never more than the blocks
template <typename Kern>
static int fused_grid(Kern kern, int64_t capacity, int* cache) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cache[dev] == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WA_WARPS * 32, 0) != cudaSuccess || occ < 1) {
      (void)cudaGetLastError();
      occ = 1;
    }
    cache[dev] = occ;
  }
  int64_t grid = (capacity + WA_WARPS - 1) / WA_WARPS;
  const int64_t cap = (int64_t)LK_SM_COUNT * cache[dev];
  if (grid > cap) grid = cap;
  return (int)(grid < 1 ? 1 : grid);
}

extern "C" int lk_link_bwd_supported(int c) { return c == 16 || c == 32 || c == 64 || c == 128; }

// ------------------------------------------------------------------ backward entry points
template <int LPR, int IB, int OP>
static int launch_bwd_norm(const float* d_mean, const float* d_tot, const int32_t* d_seg, const int32_t* d_order,
                           const int32_t* d_num, int64_t capacity, const int32_t* d_coords, const GenDev& g,
                           const float* d_local, const float* d_dout, const float* g1, const float* b1,
                           const float* g2, const float* b2, float* d_dy, float* d_dlocal, float* d_gsum,
                           float* d_dparam, cudaStream_t st) {
  static int cache[64];
  auto kern = link_bwd_norm_kernel<LPR, IB, OP>;
  const int grid = fused_grid(kern, capacity, cache);
  LK_REQUIRE(grid > 0, "lk_link_bwd_norm: cannot query the device");
  kern<<<grid, WA_WARPS * 32, 0, st>>>((const float4*)d_mean, d_tot, d_seg, d_order, d_num, capacity,
                                       (const int4*)d_coords, g, d_local, d_dout, g1, b1, g2, b2, d_dy, d_dlocal,
                                       (float4*)d_gsum, d_dparam);
  LK_LAUNCHED();
  return LK_OK;
}

template <int LPR, int IB, int OP>
static int launch_bwd_apply(const float* d_gsum, const float* d_mean, const int32_t* d_nbr_t, const int32_t* d_seg,
                            const int32_t* d_order, const int32_t* d_num, int64_t capacity, int r3,
                            const int32_t* d_coords, const GenDev& g, const float* d_fin, const float* d_dy,
                            float* d_dfin, float* d_dw, cudaStream_t st) {
  static int cache[64];
  auto kern = link_bwd_apply_kernel<LPR, IB, OP>;
  const int grid = fused_grid(kern, capacity, cache);
  LK_REQUIRE(grid > 0, "lk_link_bwd_apply: cannot query the device");
  kern<<<grid, WA_WARPS * 32, 0, st>>>((const float4*)d_gsum, (const float4*)d_mean, d_nbr_t, d_seg, d_order, d_num,
                                       capacity, r3, (const int4*)d_coords, g, d_fin, d_dy, d_dfin, d_dw);
  LK_LAUNCHED();
  return LK_OK;
}

#define BW_DISPATCH(FN, ...)                                                                     \
  do {                                                                                           \
    const int lpr = g.c / 8, span = 4 * lpr;                                                     \
    const int ib = (g.wrows % span == 0 && g.wrows / span == 1) ? 1 : 2;                         \
    const bool sin_op = g.op == LK_OP_SIN;                                                       \
    switch (lpr * 4 + (ib - 1) * 2 + (sin_op ? 1 : 0)) {                                         \
      case 2 * 4 + 0: return FN<2, 1, LK_OP_COS>(__VA_ARGS__);                                   \
      case 2 * 4 + 1: return FN<2, 1, LK_OP_SIN>(__VA_ARGS__);                                   \
      case 2 * 4 + 2: return FN<2, 2, LK_OP_COS>(__VA_ARGS__);                                   \
      case 2 * 4 + 3: return FN<2, 2, LK_OP_SIN>(__VA_ARGS__);                                   \
      case 4 * 4 + 0: return FN<4, 1, LK_OP_COS>(__VA_ARGS__);                                   \
      case 4 * 4 + 1: return FN<4, 1, LK_OP_SIN>(__VA_ARGS__);                                   \
      case 4 * 4 + 2: return FN<4, 2, LK_OP_COS>(__VA_ARGS__);                                   \
      case 4 * 4 + 3: return FN<4, 2, LK_OP_SIN>(__VA_ARGS__);                                   \
      case 8 * 4 + 0: return FN<8, 1, LK_OP_COS>(__VA_ARGS__);                                   \
      case 8 * 4 + 1: return FN<8, 1, LK_OP_SIN>(__VA_ARGS__);                                   \
      case 8 * 4 + 2: return FN<8, 2, LK_OP_COS>(__VA_ARGS__);                                   \
      case 8 * 4 + 3: return FN<8, 2, LK_OP_SIN>(__VA_ARGS__);                                   \
      case 16 * 4 + 0: return FN<16, 1, LK_OP_COS>(__VA_ARGS__);                                 \
      case 16 * 4 + 1: return FN<16, 1, LK_OP_SIN>(__VA_ARGS__);                                 \
      case 16 * 4 + 2: return FN<16, 2, LK_OP_COS>(__VA_ARGS__);                                 \
      default: return FN<16, 2, LK_OP_SIN>(__VA_ARGS__);                                         \
    }                                                                                            \
  } while (0)

static int check_bwd_gen(const lk_kernelgen_t* gen, GenDev* g, const char* who) {
  int rc = check_gen(gen, g, who);
  if (rc) return rc;
  if (!lk_link_bwd_supported(g->c) || g->op == LK_OP_COSX) {
    lk_set_error("%s: the fused backward serves C in {16, 32, 64, 128} and the ops cos / sin", who);
    return LK_EINVAL;
  }
  return LK_OK;
}

extern "C" int lk_link_bwd_norm(const float* d_mean, const float* d_tot, const int32_t* d_seg,
                                const int32_t* d_order, const int32_t* d_num, int64_t capacity,
                                const int32_t* d_coords, const lk_kernelgen_t* gen, const float* d_local,
                                const float* d_dout, const float* d_g1, const float* d_b1, const float* d_g2,
                                const float* d_b2, float* d_dy, float* d_dlocal, float* d_gsum,
                                float* d_dparam, lk_stream_t s) {
  GenDev g;
  int rc = check_bwd_gen(gen, &g, "lk_link_bwd_norm");
  if (rc) return rc;
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(capacity > 0 && d_mean && d_tot && d_seg && d_order && d_num && d_coords && d_local && d_dout &&
                 d_g1 && d_b1 && d_g2 && d_b2 && d_dy && d_dlocal && d_gsum && d_dparam,
             "lk_link_bwd_norm: null pointer");
  BW_DISPATCH(launch_bwd_norm, d_mean, d_tot, d_seg, d_order, d_num, capacity, d_coords, g, d_local, d_dout, d_g1,
              d_b1, d_g2, d_b2, d_dy, d_dlocal, d_gsum, d_dparam, (cudaStream_t)s);
}

extern "C" int lk_link_bwd_apply(const float* d_gsum, const float* d_mean, const int32_t* d_nbr_t,
                                 const int32_t* d_seg, const int32_t* d_order, const int32_t* d_num,
                                 int64_t capacity, int r3, const int32_t* d_coords, const lk_kernelgen_t* gen,
                                 const float* d_fin, const float* d_dy, float* d_dfin, float* d_dw,
                                 lk_stream_t s) {
  GenDev g;
  int rc = check_bwd_gen(gen, &g, "lk_link_bwd_apply");
  if (rc) return rc;
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(capacity > 0 && r3 > 0 && r3 <= 32 && d_gsum && d_mean && d_nbr_t && d_seg && d_order && d_num &&
                 d_coords && d_fin && d_dy && d_dfin && d_dw,
             "lk_link_bwd_apply: null pointer / bad sizes");
  BW_DISPATCH(launch_bwd_apply, d_gsum, d_mean, d_nbr_t, d_seg, d_order, d_num, capacity, r3, d_coords, g, d_fin,
              d_dy, d_dfin, d_dw, (cudaStream_t)s);
}
