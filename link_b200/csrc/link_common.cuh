// Device helpers shared by the LinK kernels (link.cu, link_fused.cu): kernel-generator state,
// phase / sin / cos evaluation, LayerNorm over a lane group.
#pragma once
#include "common.cuh"

struct GenDev {
  int op, c, wrows, accurate;
  float coord_scale;
  const float* pw;
  const float* alpha;
};

// libdevice sincosf kept out of line: inlined into the unrolled row loops its slow path
// (Payne-Hanek) multiplies the code size ~10x and the kernels become instruction-fetch bound.
static __device__ __noinline__ float2 accurate_sincos(float p) {
  float s, c;
  sincosf(p, &s, &c);
  return make_float2(s, c);
}

// per-lane kernel-generator state: weights of the NP = 4 IB distinct phases the lane evaluates
template <int NP>
struct LaneGen {
  float w0[NP], w1[NP], w2[NP], al[NP];
};

// phase slot q = ib * 4 + e  <->  channel 4 (ib * LPR + j) + e  (and every channel congruent to
// it modulo wrows that the lane owns)
template <int LPR, int IB>
__device__ __forceinline__ void load_lane_gen(const GenDev& g, int j, bool active, LaneGen<4 * IB>& lg) {
#pragma unroll
  for (int ib = 0; ib < IB; ++ib)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = ib * 4 + e;
      const int row = active ? (4 * (ib * LPR + j) + e) % g.wrows : 0;
      lg.w0[q] = active ? __ldg(g.pw + row * 3 + 0) : 0.f;
      lg.w1[q] = active ? __ldg(g.pw + row * 3 + 1) : 0.f;
      lg.w2[q] = active ? __ldg(g.pw + row * 3 + 2) : 0.f;
      lg.al[q] = (active && g.alpha) ? __ldg(g.alpha + row) : 1.f;
    }
}

// Phase, sin and cos of the lane's 4 IB distinct phases, SFU form (no accuracy switch inside the
// phase loop).  Per phase: two-term Cody-Waite reduction to [-pi, pi] (exact for the |p| < ~1e5 rad
// that voxel grids produce: fl(2pi) = 2pi + 1.7484555e-7), the quotient rounded to nearest by the
// 1.5 * 2^23 trick (FFMA + FADD on the FMA pipe instead of FMUL + FRND on the conversion pipe), then
// sin.approx / cos.approx (max abs error 2^-20.9 on [-pi, pi]).  ~10 instructions instead of ~20
// for sincosf; the 5e-7 absolute error is two orders below the parity tolerance.
template <int NP, bool COSX>
__device__ __forceinline__ void lane_trig_sfu(const GenDev& g, const LaneGen<NP>& lg, int cx, int cy, int cz,
                                              float p[NP], float sn[NP], float cs[NP]) {
  float x = (float)cx, y = (float)cy, z = (float)cz;
  if (COSX) { x = x / g.coord_scale; y = y / g.coord_scale; z = z / g.coord_scale; }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    float v = fmaf(z, lg.w2[q], fmaf(y, lg.w1[q], x * lg.w0[q]));
    p[q] = COSX ? v * lg.al[q] : v;
    const float t = fmaf(p[q], 0.15915494309189535f, 12582912.f);   // 1.5 * 2^23: rounds to an integer
    const float k = t - 12582912.f;
    float r = fmaf(k, -6.2831855f, p[q]);
    r = fmaf(k, 1.7484555e-7f, r);
#ifdef PR_DEBUG_NOTRIG                               // timing experiments only (results are wrong)
    sn[q] = r; cs[q] = k;
#else
    sn[q] = __sinf(r);
    cs[q] = __cosf(r);
#endif
  }
}

// Phase, sin and cos of the lane's 4 IB distinct phases for voxel (x,y,z); the phase follows the
// operation order of nn.Linear(3, .) [ (x*w0 + y*w1) + z*w2 ] followed by "* alpha"
// (linkencoder.py:151,165).
template <int NP, bool COSX>
__device__ __forceinline__ void lane_trig(const GenDev& g, const LaneGen<NP>& lg, int cx, int cy, int cz,
                                          float p[NP], float sn[NP], float cs[NP]) {
  if (!g.accurate) {                             // warp-uniform: one test per row, not one per phase
    lane_trig_sfu<NP, COSX>(g, lg, cx, cy, cz, p, sn, cs);
    return;
  }
  float x = (float)cx, y = (float)cy, z = (float)cz;
  if (COSX) { x = x / g.coord_scale; y = y / g.coord_scale; z = z / g.coord_scale; }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    float v = fmaf(z, lg.w2[q], fmaf(y, lg.w1[q], x * lg.w0[q]));
    p[q] = COSX ? v * lg.al[q] : v;
    const float2 sc = accurate_sincos(p[q]);
    sn[q] = sc.x; cs[q] = sc.y;
  }
}

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm of one row spread over LPR lanes x VPL vectors (eps 1e-6, biased variance)
template <int LPR, int VPL>
__device__ __forceinline__ void group_layernorm(float v[VPL][4], bool active, float inv_c,
                                                const float* __restrict__ gam,
                                                const float* __restrict__ bet, int j) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  const float mean = group_sum<LPR>(active ? s : 0.f) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[i][e] -= mean; q = fmaf(v[i][e], v[i][e], q); }
  const float var = group_sum<LPR>(active ? q : 0.f) * inv_c;
  const float rstd = rsqrtf(var + 1e-6f);
  if (active) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 gg = __ldg((const float4*)gam + i * LPR + j);
      const float4 bb = __ldg((const float4*)bet + i * LPR + j);
      v[i][0] = fmaf(v[i][0] * rstd, gg.x, bb.x); v[i][1] = fmaf(v[i][1] * rstd, gg.y, bb.y);
      v[i][2] = fmaf(v[i][2] * rstd, gg.z, bb.z); v[i][3] = fmaf(v[i][3] * rstd, gg.w, bb.w);
    }
  }
}

static inline int check_gen(const lk_kernelgen_t* gen, GenDev* g, const char* who) {
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
