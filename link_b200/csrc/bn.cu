// Training-mode BatchNorm over sparse feature rows, fused with what follows it in the LinK models:
//
//     y = relu?( (x - mean) * invstd * gamma + beta  [+ residual] )        x, y: [n, c] fp32 rows
//
// (reference: spnn.BatchNorm = nn.BatchNorm1d over `.feats`, torchsparse/nn/modules/norm.py:10-13,
// followed by spnn.ReLU / the shortcut add of ResidualBlock, linkencoder.py:26-37, 64-91; in the
// detection backbone nn.BatchNorm1d + nn.ReLU on `.features`, det3d/models/backbones/scn.py:64-107.)
// The reference runs BatchNorm, the add and the ReLU as three library ops with their own autograd
// nodes (statistics, transform, add, threshold forward; threshold, two-pass BatchNorm backward): 7
// passes over the [n, c] activation going forward and back.  Here each direction is TWO bandwidth-bound
// kernels behind one C call:
//
//   forward : bn_stats   per-channel sum / sum of squares     (reads x once)
//             bn_apply   normalise + affine [+ residual] [+ ReLU]; the first CTA also updates
//                        running_mean / running_var / num_batches_tracked and saves mean / invstd
//   backward: bn_bwd_reduce  g = dy * (y > 0);  sum g,  sum g * xhat          (dbeta, dgamma)
//             bn_bwd_apply   dx = gamma invstd (g - dbeta / n - xhat dgamma / n)  [, dresidual = g]
//
// Layout: a thread owns one float4 column group and strides over the rows (coalesced 16-byte loads,
// the per-channel constants live in registers).  Per-thread partial sums are kept in double and the
// CTAs join theirs with double atomics: with fp32 partials over 1e5+ rows the variance E[x^2] - mean^2
// loses digits that PyTorch's Welford keeps; in double the combination is exact to fp32 round-off and
// the order of the atomics does not show in the fp32 results.
#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;

struct BnGeom {
  int c4;    // float4 column groups per row
  int rpp;   // rows per CTA pass
};

__device__ __forceinline__ BnGeom bn_geom(int c) {
  BnGeom g;
  g.c4 = c >> 2;
  g.rpp = BN_THREADS / g.c4;
  return g;
}

// join the per-thread double partials of a CTA (8 per thread: 4 channels x 2 sums) and add them to
// the global accumulators acc[0:c] / acc[c:2c]
__device__ __forceinline__ void bn_block_join(const double (&a)[4], const double (&b)[4], int tx, int ty, int c,
                                              const BnGeom& g, double* __restrict__ acc) {
  __shared__ double sh[BN_THREADS * 8];
  const int active = g.rpp * g.c4;
  if (ty < g.rpp) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sh[(e * g.rpp + ty) * g.c4 + tx] = a[e];
      sh[((4 + e) * g.rpp + ty) * g.c4 + tx] = b[e];
    }
  }
  __syncthreads();
  // 8 c4 output sums; thread t < 8 c4 adds its column over the rpp rows
  for (int o = threadIdx.x; o < 8 * g.c4; o += BN_THREADS) {
    const int e = o / g.c4, x = o % g.c4;
    double s = 0.0;
    for (int r = 0; r < g.rpp; ++r) s += sh[(e * g.rpp + r) * g.c4 + x];
    const int ch = 4 * x + (e & 3);
    atomicAdd(acc + (e < 4 ? ch : c + ch), s);
  }
  (void)active;
}

__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float4* __restrict__ x, int64_t n, int c,
                                                              double* __restrict__ acc) {
  const BnGeom g = bn_geom(c);
  const int tx = threadIdx.x % g.c4, ty = threadIdx.x / g.c4;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (ty < g.rpp) {
    for (int64_t r = (int64_t)blockIdx.x * g.rpp + ty; r < n; r += (int64_t)gridDim.x * g.rpp) {
      const float4 v = lk_ldg_stream(x + r * g.c4 + tx);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += (double)v.x * v.x; q[1] += (double)v.y * v.y; q[2] += (double)v.z * v.z; q[3] += (double)v.w * v.w;
    }
  }
  bn_block_join(s, q, tx, ty, c, g, acc);
}

__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(
    const float4* __restrict__ x, const float4* __restrict__ res, int64_t n, int c, const double* __restrict__ acc,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum, int relu,
    float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt,
    float* __restrict__ save_mean, float* __restrict__ save_invstd, float4* __restrict__ y) {
  const BnGeom g = bn_geom(c);
  const int tx = threadIdx.x % g.c4, ty = threadIdx.x / g.c4;
  if (ty >= g.rpp) return;
  float sc[4], sh[4];
  const double inv_n = 1.0 / (double)n;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int ch = 4 * tx + e;
    const double mean = acc[ch] * inv_n;
    double var = acc[c + ch] * inv_n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float gm = gamma ? gamma[ch] : 1.f, bt = beta ? beta[ch] : 0.f;
    sc[e] = invstd * gm;
    sh[e] = bt - (float)mean * sc[e];
    if (blockIdx.x == 0 && ty == 0) {
      save_mean[ch] = (float)mean;
      save_invstd[ch] = invstd;
      if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
      if (running_var) {
        const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
  for (int64_t r = (int64_t)blockIdx.x * g.rpp + ty; r < n; r += (int64_t)gridDim.x * g.rpp) {
    const int64_t i = r * g.c4 + tx;
    const float4 v = lk_ldg_stream(x + i);
    float4 o = make_float4(fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3]));
    if (res) {
      const float4 rv = lk_ldg_stream(res + i);
      o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    y[i] = o;     // read again by the next layer: ordinary store
  }
}

__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(
    const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ y /*or NULL: no ReLU*/,
    int64_t n, int c, const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ acc) {
  const BnGeom g = bn_geom(c);
  const int tx = threadIdx.x % g.c4, ty = threadIdx.x / g.c4;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (ty < g.rpp) {
    float mu[4], is[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { mu[e] = mean[4 * tx + e]; is[e] = invstd[4 * tx + e]; }
    for (int64_t r = (int64_t)blockIdx.x * g.rpp + ty; r < n; r += (int64_t)gridDim.x * g.rpp) {
      const int64_t i = r * g.c4 + tx;
      float4 gd = lk_ldg_stream(dy + i);
      const float4 v = lk_ldg_stream(x + i);
      if (y) {
        const float4 o = lk_ldg_stream(y + i);
        gd.x = o.x > 0.f ? gd.x : 0.f; gd.y = o.y > 0.f ? gd.y : 0.f;
        gd.z = o.z > 0.f ? gd.z : 0.f; gd.w = o.w > 0.f ? gd.w : 0.f;
      }
      s[0] += gd.x; s[1] += gd.y; s[2] += gd.z; s[3] += gd.w;
      q[0] += (double)gd.x * ((v.x - mu[0]) * is[0]); q[1] += (double)gd.y * ((v.y - mu[1]) * is[1]);
      q[2] += (double)gd.z * ((v.z - mu[2]) * is[2]); q[3] += (double)gd.w * ((v.w - mu[3]) * is[3]);
    }
  }
  bn_block_join(s, q, tx, ty, c, g, acc);
}

__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(
    const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ y, int64_t n, int c,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const double* __restrict__ acc, float4* __restrict__ dx, float4* __restrict__ dres /*or NULL*/,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const BnGeom g = bn_geom(c);
  const int tx = threadIdx.x % g.c4, ty = threadIdx.x / g.c4;
  if (ty >= g.rpp) return;
  float mu[4], is[4], k[4], a[4], b[4];
  const double inv_n = 1.0 / (double)n;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int ch = 4 * tx + e;
    mu[e] = mean[ch];
    is[e] = invstd[ch];
    k[e] = (gamma ? gamma[ch] : 1.f) * is[e];
    a[e] = (float)(acc[ch] * inv_n);          // dbeta / n
    b[e] = (float)(acc[c + ch] * inv_n);      // dgamma / n
    if (blockIdx.x == 0 && ty == 0) {
      if (dbeta) dbeta[ch] = (float)acc[ch];
      if (dgamma) dgamma[ch] = (float)acc[c + ch];
    }
  }
  for (int64_t r = (int64_t)blockIdx.x * g.rpp + ty; r < n; r += (int64_t)gridDim.x * g.rpp) {
    const int64_t i = r * g.c4 + tx;
    float4 gd = lk_ldg_stream(dy + i);
    const float4 v = lk_ldg_stream(x + i);
    if (y) {
      const float4 o = lk_ldg_stream(y + i);
      gd.x = o.x > 0.f ? gd.x : 0.f; gd.y = o.y > 0.f ? gd.y : 0.f;
      gd.z = o.z > 0.f ? gd.z : 0.f; gd.w = o.w > 0.f ? gd.w : 0.f;
    }
    if (dres) dres[i] = gd;
    float4 o;
    o.x = k[0] * (gd.x - a[0] - (v.x - mu[0]) * is[0] * b[0]);
    o.y = k[1] * (gd.y - a[1] - (v.y - mu[1]) * is[1] * b[1]);
    o.z = k[2] * (gd.z - a[2] - (v.z - mu[2]) * is[2] * b[2]);
    o.w = k[3] * (gd.w - a[3] - (v.w - mu[3]) * is[3] * b[3]);
    dx[i] = o;
  }
}

inline int bn_grid(int64_t n, int c) {
  const int rpp = BN_THREADS / (c / 4);
  return lk_grid(n, rpp * 8, 4);      // >= 8 rows per thread before a second CTA is worth its join
}

}  // namespace

extern "C" int lk_bn_supported(int c) { return c >= 4 && c % 4 == 0 && c <= 4 * BN_THREADS; }

extern "C" int64_t lk_bn_ws_bytes(int c) { return (int64_t)2 * c * sizeof(double); }

extern "C" int lk_bn_train_fwd(const float* d_x, const float* d_residual, int64_t n, int c, const float* d_gamma,
                               const float* d_beta, float eps, float momentum, int relu, float* d_running_mean,
                               float* d_running_var, int64_t* d_num_batches_tracked, float* d_save_mean,
                               float* d_save_invstd, float* d_y, void* d_ws, int64_t ws_bytes, lk_stream_t s) {
  LK_REQUIRE(lk_bn_supported(c), "lk_bn_train_fwd: c = %d (need a multiple of 4, <= %d)", c, 4 * BN_THREADS);
  LK_REQUIRE(n >= 1 && d_x && d_y && d_save_mean && d_save_invstd && d_ws, "lk_bn_train_fwd: bad arguments");
  LK_REQUIRE(ws_bytes >= lk_bn_ws_bytes(c), "lk_bn_train_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)s;
  double* acc = (double*)d_ws;
  LK_CUDA(cudaMemsetAsync(acc, 0, lk_bn_ws_bytes(c), st));
  lk_count_launch();
  const int grid = bn_grid(n, c);
  bn_stats_kernel<<<grid, BN_THREADS, 0, st>>>((const float4*)d_x, n, c, acc);
  LK_LAUNCHED();
  bn_apply_kernel<<<grid, BN_THREADS, 0, st>>>((const float4*)d_x, (const float4*)d_residual, n, c, acc, d_gamma, d_beta,
                                               eps, momentum, relu, d_running_mean, d_running_var,
                                               (long long*)d_num_batches_tracked, d_save_mean, d_save_invstd,
                                               (float4*)d_y);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_bn_train_bwd(const float* d_dy, const float* d_x, const float* d_y, int64_t n, int c,
                               const float* d_save_mean, const float* d_save_invstd, const float* d_gamma,
                               float* d_dx, float* d_dresidual, float* d_dgamma, float* d_dbeta, void* d_ws,
                               int64_t ws_bytes, lk_stream_t s) {
  LK_REQUIRE(lk_bn_supported(c), "lk_bn_train_bwd: c = %d (need a multiple of 4, <= %d)", c, 4 * BN_THREADS);
  LK_REQUIRE(n >= 1 && d_dy && d_x && d_dx && d_save_mean && d_save_invstd && d_ws, "lk_bn_train_bwd: bad arguments");
  LK_REQUIRE(ws_bytes >= lk_bn_ws_bytes(c), "lk_bn_train_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)s;
  double* acc = (double*)d_ws;
  LK_CUDA(cudaMemsetAsync(acc, 0, lk_bn_ws_bytes(c), st));
  lk_count_launch();
  const int grid = bn_grid(n, c);
  bn_bwd_reduce_kernel<<<grid, BN_THREADS, 0, st>>>((const float4*)d_dy, (const float4*)d_x, (const float4*)d_y, n, c,
                                                    d_save_mean, d_save_invstd, acc);
  LK_LAUNCHED();
  bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, st>>>((const float4*)d_dy, (const float4*)d_x, (const float4*)d_y, n, c,
                                                   d_save_mean, d_save_invstd, d_gamma, acc, (float4*)d_dx,
                                                   (float4*)d_dresidual, d_dgamma, d_dbeta);
  LK_LAUNCHED();
  return LK_OK;
}
