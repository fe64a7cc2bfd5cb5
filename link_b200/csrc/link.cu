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
// in both passes (12 FMAs + one sincos per 4 channels), so the only HBM streams are: read F_in
// once, read coords + block index twice, write out once; block sums / means are L2 resident.
//
// Thread mapping: a row of C floats is owned by a group of LPR = pow2 >= C/4 lanes, 4 channels
// (one 128-bit transaction) per lane, so a warp handles G = 32/LPR rows per step.
#include "common.cuh"

struct GenDev {
  int op, c, wrows, accurate;
  float coord_scale;
  const float* pw;
  const float* alpha;
};

struct LaneGen {       // per-lane kernel-generator state for its 4 channels
  float w0[4], w1[4], w2[4], al[4];
};

// sin and cos of one fp32 phase: two-term Cody-Waite reduction to [-pi, pi] (exact for the
// |p| < ~1e5 rad that voxel grids produce: fl(2pi) = 2pi + 1.7484555e-7) followed by the SFU
// approximations (sin.approx / cos.approx, max abs error 2^-20.9 on [-pi, pi]).  ~9 instructions
// instead of ~20 for sincosf; the 5e-7 absolute error is two orders below the parity tolerance.
__device__ __forceinline__ void fast_sincos(float p, bool accurate, float& s, float& c) {
  if (accurate) { sincosf(p, &s, &c); return; }      // warp-uniform switch
  float k = rintf(p * 0.15915494309189535f);
  float r = fmaf(k, -6.2831855f, p);
  r = fmaf(k, 1.7484555e-7f, r);
  s = __sinf(r);
  c = __cosf(r);
}

// Phase, sin and cos of the lane's 4 channels for voxel (x,y,z); the phase follows the operation
// order of nn.Linear(3, .) [ (x*w0 + y*w1) + z*w2 ] followed by "* alpha" (linkencoder.py:151,165).
// With channel groups (pos.repeat([1, groups]), linkencoder.py:152) the SH = C/wrows lanes
// {l0 + j*S} of a row own IDENTICAL phases: each evaluates only N = 4/SH of them (its weights for
// exactly those were loaded by load_lane_gen_owned) and the rest arrive by shuffle, so the
// transcendental work per row drops SH-fold.  Must be called by all 32 lanes.
template <int LPR, int SH, bool COSX>
__device__ __forceinline__ void lane_trig(const GenDev& g, const LaneGen& lg, int4 c, int lane,
                                          float p[4], float sn[4], float cs[4]) {
  constexpr int N = 4 / SH;            // phases evaluated by this lane
  float x = (float)c.x, y = (float)c.y, z = (float)c.z;
  if (COSX) { x = x / g.coord_scale; y = y / g.coord_scale; z = z / g.coord_scale; }
  float mp[N], ms[N], mc[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    float v = fmaf(z, lg.w2[q], fmaf(y, lg.w1[q], x * lg.w0[q]));
    mp[q] = COSX ? v * lg.al[q] : v;
    fast_sincos(mp[q], g.accurate != 0, ms[q], mc[q]);
  }
  if (SH == 1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { p[e] = mp[e % N]; sn[e] = ms[e % N]; cs[e] = mc[e % N]; }
  } else {
    constexpr int S = LPR / SH;        // lane distance between the copies
    const int li = lane % LPR, base = lane - li + (li % S);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int src = base + (e / N) * S;
      sn[e] = __shfl_sync(0xffffffffu, ms[e % N], src);
      cs[e] = __shfl_sync(0xffffffffu, mc[e % N], src);
      if (COSX) p[e] = __shfl_sync(0xffffffffu, mp[e % N], src);
    }
  }
}

// weights of the phases this lane evaluates: element q <-> channel ch + j*N + q, j = copy index
template <int LPR, int SH>
__device__ __forceinline__ void load_lane_gen_owned(const GenDev& g, int lane, int ch, bool active,
                                                    LaneGen& lg) {
  constexpr int N = 4 / SH;
  const int j = (SH == 1) ? 0 : (lane % LPR) / (LPR / SH);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int row = (active && q < N) ? (ch + j * N + q) % g.wrows : 0;
    bool on = active && q < N;
    lg.w0[q] = on ? g.pw[row * 3 + 0] : 0.f;
    lg.w1[q] = on ? g.pw[row * 3 + 1] : 0.f;
    lg.w2[q] = on ? g.pw[row * 3 + 2] : 0.f;
    lg.al[q] = (on && g.alpha) ? g.alpha[row] : 1.f;
  }
}

#define PREAGG_ROWS_PER_GROUP 8
#define PREAGG_UNROLL 2

// ------------------------------------------------------------------ pass 1: block sums
template <int LPR, int OP, int SH>
__global__ void __launch_bounds__(256, 4) link_preagg_kernel(const float* __restrict__ fin,
                                                          const int4* __restrict__ coords,
                                                          const int* __restrict__ blk, int64_t n,
                                                          GenDev g, float* sums) {
  constexpr int G = 32 / LPR;
  constexpr int K = (OP == LK_OP_COSX) ? 3 : 2;
  constexpr bool COSX = (OP == LK_OP_COSX);
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;
  const int ch = (lane % LPR) * 4;
  const bool active = ch < g.c;
  const int kc = K * g.c;
  LaneGen lg;
  load_lane_gen_owned<LPR, SH>(g, lane, ch, active, lg);

  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows_per_warp = (int64_t)PREAGG_ROWS_PER_GROUP * G;
  for (int64_t wbase = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * rows_per_warp;
       wbase < n; wbase += warps_total * rows_per_warp) {
    int64_t r0 = wbase + (int64_t)grp * PREAGG_ROWS_PER_GROUP;
    int64_t r1 = r0 + PREAGG_ROWS_PER_GROUP;
    if (r1 > n) r1 = n;
    float acc[K][4];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
    int cur = -1;
    // NB: trip count is warp-uniform (r1 - r0 may be shorter for the last group; `ok` masks it)
    for (int64_t r = r0; r < r0 + PREAGG_ROWS_PER_GROUP; r += PREAGG_UNROLL) {
      int b[PREAGG_UNROLL];
      int4 cc[PREAGG_UNROLL];
      float4 f[PREAGG_UNROLL];
#pragma unroll
      for (int u = 0; u < PREAGG_UNROLL; ++u) {   // all loads of the batch first (MLP)
        bool ok = r + u < r1;
        b[u] = ok ? __ldg(blk + r + u) : -1;
        cc[u] = ok ? __ldg(coords + r + u) : make_int4(0, 0, 0, 0);
        f[u] = (ok && active) ? lk_ldg_stream((const float4*)(fin + (r + u) * g.c + ch))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < PREAGG_UNROLL; ++u) {
        float p[4], sn[4], cs[4];
        lane_trig<LPR, SH, COSX>(g, lg, cc[u], lane, p, sn, cs);   // all lanes: full-mask shuffles
        if (b[u] < 0) continue;                    // tail of the batch (or unmapped voxel)
        if (b[u] != cur) {                         // run boundary: flush the finished block
          if (cur >= 0 && active) {
#pragma unroll
            for (int q = 0; q < K; ++q)
              lk_red_add_v4(sums + (int64_t)cur * kc + q * g.c + ch,
                            make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
          }
#pragma unroll
          for (int q = 0; q < K; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
          cur = b[u];
        }
        const float fv[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // plane order: cos -> [cos, sin]; sin -> [sin, cos]; cos_x -> [cos, sin, lin]
          acc[0][e] += fv[e] * (OP == LK_OP_SIN ? sn[e] : cs[e]);
          acc[1][e] += fv[e] * (OP == LK_OP_SIN ? cs[e] : sn[e]);
          if (COSX) acc[K - 1][e] += fv[e] * p[e];
        }
      }
    }
    if (cur >= 0 && active) {
#pragma unroll
      for (int q = 0; q < K; ++q)
        lk_red_add_v4(sums + (int64_t)cur * kc + q * g.c + ch,
                      make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
    }
  }
}

// zero the first *d_num rows of a [capacity, row_floats] buffer (block sums are allocated for the
// worst case M = N but only the M live rows are ever touched)
__global__ void __launch_bounds__(256) zero_rows_kernel(float4* __restrict__ p,
                                                        const int* __restrict__ d_num,
                                                        int64_t capacity, int row_vec) {
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
  zero_rows_kernel<<<lk_grid(capacity * (row_floats / 4), 256, 8), 256, 0, (cudaStream_t)s>>>(
      (float4*)d_buf, d_num, capacity, row_floats / 4);
  LK_LAUNCHED();
  return LK_OK;
}

// ------------------------------------------------------------------ pass 2a: window means
__global__ void __launch_bounds__(256) link_window_mean_kernel(const float* __restrict__ sums,
                                                               const int* __restrict__ counts,
                                                               const int* __restrict__ nbr,
                                                               const int* __restrict__ d_num,
                                                               int64_t capacity, int R, int kc,
                                                               float* __restrict__ mean) {
  int64_t m = *d_num;
  if (m > capacity) m = capacity;
  int vpr = kc >> 2;
  int64_t total = m * vpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / vpr;
    int j = (int)(t - b * vpr) * 4;
    const int* nb = nbr + b * R;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float tot = 0.f;
#pragma unroll 9
    for (int k = 0; k < R; ++k) {            // branch-free so the R row loads can be in flight
      int src = __ldg(nb + k);
      float wgt = src >= 0 ? 1.f : 0.f;
      src = max(src, 0);
      float4 v = __ldg((const float4*)(sums + (int64_t)src * kc + j));
      acc.x += wgt * v.x; acc.y += wgt * v.y; acc.z += wgt * v.z; acc.w += wgt * v.w;
      tot += wgt * (float)__ldg(counts + src);
    }
    acc.x /= tot; acc.y /= tot; acc.z /= tot; acc.w /= tot;
    *(float4*)(mean + b * kc + j) = acc;
  }
}

// ------------------------------------------------------------------ pass 2b: combine (+ norms)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR>
__device__ __forceinline__ void group_layernorm(float v[4], bool active, float inv_c, const float* gam,
                                                const float* bet, int ch) {
  float s = active ? (v[0] + v[1] + v[2] + v[3]) : 0.f;
  float mean = group_sum<LPR>(s) * inv_c;
  float d[4], q = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) { d[e] = v[e] - mean; q += d[e] * d[e]; }
  float var = group_sum<LPR>(active ? q : 0.f) * inv_c;
  float rstd = rsqrtf(var + 1e-6f);
  if (active) {
    float4 gg = __ldg((const float4*)(gam + ch));
    float4 bb = __ldg((const float4*)(bet + ch));
    v[0] = d[0] * rstd * gg.x + bb.x; v[1] = d[1] * rstd * gg.y + bb.y;
    v[2] = d[2] * rstd * gg.z + bb.z; v[3] = d[3] * rstd * gg.w + bb.w;
  }
}

template <int LPR, int OP, bool NORM, int SH>
__global__ void __launch_bounds__(256, 4) link_apply_kernel(
    const float* __restrict__ mean, const float* __restrict__ fin, const int4* __restrict__ coords,
    const int* __restrict__ blk, int64_t n, GenDev g, const float* __restrict__ local,
    const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
    const float* __restrict__ b2, float* __restrict__ out) {
  constexpr int G = 32 / LPR;
  constexpr int K = (OP == LK_OP_COSX) ? 3 : 2;
  constexpr bool COSX = (OP == LK_OP_COSX);
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;
  const int ch = (lane % LPR) * 4;
  const bool active = ch < g.c;
  const int kc = K * g.c;
  LaneGen lg;
  load_lane_gen_owned<LPR, SH>(g, lane, ch, active, lg);
  const float inv_c = 1.0f / (float)g.c;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t steps = (n + G - 1) / G;
  // software pipeline: block index, coordinate and local_mix row of the NEXT step are loaded
  // while the current step computes (the mean-row loads depend on the block index, so without
  // the prefetch every step pays two dependent memory latencies back to back)
  int64_t step = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int b_n = -1;
  int4 cc_n = make_int4(0, 0, 0, 0);
  float4 lv_n = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int64_t st) {
    int64_t r = st * G + grp;
    bool ok = st < steps && r < n;
    b_n = ok ? __ldg(blk + r) : -1;
    cc_n = ok ? __ldg(coords + r) : make_int4(0, 0, 0, 0);
    lv_n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NORM && ok && active) lv_n = lk_ldg_stream((const float4*)(local + r * g.c + ch));
  };
  prefetch(step);
  for (; step < steps; step += warps_total) {
    const int64_t r = step * G + grp;
    const bool ok = r < n;               // NB: whole groups go inactive together; shuffles stay
    const int b = b_n;                   // inside a group, so no divergence hazard
    const int4 cc = cc_n;
    const float4 lv = lv_n;
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0, m2 = m0, f = m0;
    const bool live = ok && active && b >= 0;
    if (live) {                          // issue the dependent loads first ...
      const float* mrow = mean + (int64_t)b * kc + ch;
      m0 = __ldg((const float4*)mrow);
      m1 = __ldg((const float4*)(mrow + g.c));
      if (COSX) {
        m2 = __ldg((const float4*)(mrow + 2 * g.c));
        f = lk_ldg_stream((const float4*)(fin + r * g.c + ch));
      }
    }
    prefetch(step + warps_total);        // ... then the next step's independent ones
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    float p[4], sn[4], cs[4];
    lane_trig<LPR, SH, COSX>(g, lg, cc, lane, p, sn, cs);   // all lanes: full-mask shuffles
    if (live) {
      const float a0[4] = {m0.x, m0.y, m0.z, m0.w};
      const float a1[4] = {m1.x, m1.y, m1.z, m1.w};
      if (OP == LK_OP_SIN) {             // planes [sin, cos]: F[:, :C]*cos - F[:, C:]*sin
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = a0[e] * cs[e] - a1[e] * sn[e];
      } else {                           // planes [cos, sin]: F[:, :C]*cos + F[:, C:]*sin
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = a0[e] * cs[e] + a1[e] * sn[e];
      }
      if (COSX) {                        // + (mean(F*pos) - F*pos), linkencoder.py:176
        v[0] += m2.x - f.x * p[0]; v[1] += m2.y - f.y * p[1];
        v[2] += m2.z - f.z * p[2]; v[3] += m2.w - f.w * p[3];
      }
    }
    if (NORM) {
      float l[4] = {lv.x, lv.y, lv.z, lv.w};
      group_layernorm<LPR>(v, active, inv_c, g1, b1, ch);
      group_layernorm<LPR>(l, active, inv_c, g2, b2, ch);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e] + l[e], 0.f);
    }
    if (ok && active)
      lk_stg_stream((float4*)(out + r * g.c + ch), make_float4(v[0], v[1], v[2], v[3]));
  }
}

// ------------------------------------------------------------------ host side
static int check_gen(const lk_kernelgen_t* gen, GenDev* g, const char* who) {
  if (!gen || !gen->d_pos_weight || gen->c <= 0 || gen->c % 4 != 0 || gen->c > 128 ||
      gen->wrows <= 0 || gen->op < 0 || gen->op > 2 || !(gen->coord_scale > 0.f)) {
    lk_set_error("%s: invalid kernel generator (C must be a multiple of 4, <= 128)", who);
    return LK_EINVAL;
  }
  g->op = gen->op; g->c = gen->c; g->wrows = gen->wrows; g->coord_scale = gen->coord_scale;
  g->accurate = gen->accurate_trig;
  g->pw = gen->d_pos_weight; g->alpha = gen->d_alpha;
  return LK_OK;
}

// number of lanes of a row that own identical phases (channel groups), if the layout allows the
// shuffle exchange: C/4 a power of two, wrows a multiple of 4, C/wrows in {2,4}
static int share_of(const GenDev& g, int lpr) {
  if (lpr * 4 != g.c || g.wrows % 4 != 0 || g.c % g.wrows != 0) return 1;
  int sh = g.c / g.wrows;
  return (sh == 2 || sh == 4) ? sh : 1;
}

static int lpr_of(int c) {
  int v = c / 4, l = 1;
  while (l < v) l <<= 1;
  return l;
}

#define DISPATCH_LPR(LPRV, MACRO) \
  switch (LPRV) {                 \
    case 1: MACRO(1); break;      \
    case 2: MACRO(2); break;      \
    case 4: MACRO(4); break;      \
    case 8: MACRO(8); break;      \
    case 16: MACRO(16); break;    \
    default: MACRO(32); break;    \
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
  int lpr = lpr_of(g.c);
  int64_t rows_per_warp = (int64_t)PREAGG_ROWS_PER_GROUP * (32 / lpr);
  int64_t warps = (n + rows_per_warp - 1) / rows_per_warp;
  int grid = lk_grid(warps * 32, 256, 8);
  cudaStream_t st = (cudaStream_t)s;
  const int sh = share_of(g, lpr);
#define LAUNCH_PRE2(L, O, S) \
  link_preagg_kernel<L, O, S><<<grid, 256, 0, st>>>(d_fin, (const int4*)d_coords, d_blk, n, g, d_sums)
#define LAUNCH_PRE1(L, O)                                           \
  do {                                                              \
    if (sh == 2 && L >= 2) LAUNCH_PRE2(L, O, (L >= 2 ? 2 : 1));     \
    else if (sh == 4 && L >= 4) LAUNCH_PRE2(L, O, (L >= 4 ? 4 : 1)); \
    else LAUNCH_PRE2(L, O, 1);                                      \
  } while (0)
#define LAUNCH_PRE(L)                                   \
  do {                                                  \
    if (g.op == LK_OP_COS) LAUNCH_PRE1(L, LK_OP_COS);   \
    else if (g.op == LK_OP_SIN) LAUNCH_PRE1(L, LK_OP_SIN); \
    else LAUNCH_PRE1(L, LK_OP_COSX);                    \
  } while (0)
  DISPATCH_LPR(lpr, LAUNCH_PRE);
#undef LAUNCH_PRE
#undef LAUNCH_PRE1
#undef LAUNCH_PRE2
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_link_window_mean(const float* d_sums, const int32_t* d_counts,
                                   const int32_t* d_nbr, const int32_t* d_num, int64_t capacity,
                                   int r3, int kc, float* d_mean, lk_stream_t s) {
  LK_REQUIRE(capacity >= 0 && r3 > 0 && kc > 0 && kc % 4 == 0, "lk_link_window_mean: bad sizes");
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(d_sums && d_counts && d_nbr && d_num && d_mean, "lk_link_window_mean: null pointer");
  link_window_mean_kernel<<<lk_grid(capacity * (kc / 4), 256, 8), 256, 0, (cudaStream_t)s>>>(
      d_sums, d_counts, d_nbr, d_num, capacity, r3, kc, d_mean);
  LK_LAUNCHED();
  return LK_OK;
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
  int lpr = lpr_of(g.c);
  int64_t steps = (n + (32 / lpr) - 1) / (32 / lpr);
  int grid = lk_grid(steps * 32, 256, 8);
  cudaStream_t st = (cudaStream_t)s;
  const int sh = share_of(g, lpr);
#define LAUNCH_APPLY3(L, O, NRM, S)                                                              \
  link_apply_kernel<L, O, NRM, S><<<grid, 256, 0, st>>>(d_mean, d_fin, (const int4*)d_coords,    \
                                                        d_blk, n, g, d_local, d_g1, d_b1, d_g2,  \
                                                        d_b2, d_out)
#define LAUNCH_APPLY2N(L, O, NRM)                                       \
  do {                                                                  \
    if (sh == 2 && L >= 2) LAUNCH_APPLY3(L, O, NRM, (L >= 2 ? 2 : 1));  \
    else if (sh == 4 && L >= 4) LAUNCH_APPLY3(L, O, NRM, (L >= 4 ? 4 : 1)); \
    else LAUNCH_APPLY3(L, O, NRM, 1);                                   \
  } while (0)
#define LAUNCH_APPLY2(L, O)                      \
  do {                                           \
    if (fuse_norm) LAUNCH_APPLY2N(L, O, true);   \
    else LAUNCH_APPLY2N(L, O, false);            \
  } while (0)
#define LAUNCH_APPLY(L)                                         \
  do {                                                          \
    if (g.op == LK_OP_COS) LAUNCH_APPLY2(L, LK_OP_COS);         \
    else if (g.op == LK_OP_SIN) LAUNCH_APPLY2(L, LK_OP_SIN);    \
    else LAUNCH_APPLY2(L, LK_OP_COSX);                          \
  } while (0)
  DISPATCH_LPR(lpr, LAUNCH_APPLY);
#undef LAUNCH_APPLY
#undef LAUNCH_APPLY2
#undef LAUNCH_APPLY2N
#undef LAUNCH_APPLY3
  LK_LAUNCHED();
  return LK_OK;
}
