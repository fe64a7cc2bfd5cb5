// Rotated bird's-eye-view IoU of 3D boxes, one thread per (a, b) pair (reference:
// detection/det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-300 boxes_iou_bev / nms kernels and
// their CPU twin src/iou3d_cpu.cpp:59-252).  Boxes are (x, y, z, dx, dy, dz, heading).
//
// The intersection polygon of two rotated rectangles is the set of proper edge crossings (<= 16,
// in fact <= 8) plus the corners of either box inside the other (margin 1e-2), ordered by angle
// around their mean and measured with the shoelace sum -- the reference's formulation, so results
// agree to float rounding.  Everything lives in registers / local arrays of 24 points; the pairwise
// matrix is written once (N*M*4 bytes), the boxes are read through L1/L2 (28 bytes each).
//
// lk_iou_pair is __host__ __device__: lk_boxes_iou_bev_hostcheck runs the SAME arithmetic on the
// host so that the CPU test suite can pin it to the reference-generated fixture; it is a test hook,
// not a fallback (python never calls it outside tests).
#include <math.h>

#include "common.cuh"

struct P2 { float x, y; };

__host__ __device__ __forceinline__ float lk_cross3(P2 p1, P2 p2, P2 p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

__host__ __device__ __forceinline__ void lk_corners(const float* b, P2 c[4]) {
  const float hx = b[3] * 0.5f, hy = b[4] * 0.5f;
  const float cs = cosf(b[6]), sn = sinf(b[6]);
  const float sx[4] = {-hx, hx, hx, -hx}, sy[4] = {-hy, -hy, hy, hy};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c[k].x = sx[k] * cs - sy[k] * sn + b[0];
    c[k].y = sx[k] * sn + sy[k] * cs + b[1];
  }
}

__host__ __device__ __forceinline__ bool lk_in_box(const float* b, P2 p) {
  const float cs = cosf(-b[6]), sn = sinf(-b[6]);
  const float dx = p.x - b[0], dy = p.y - b[1];
  const float rx = dx * cs - dy * sn, ry = dx * sn + dy * cs;
  return fabsf(rx) < b[3] * 0.5f + 1e-2f && fabsf(ry) < b[4] * 0.5f + 1e-2f;
}

__host__ __device__ inline float lk_iou_pair(const float* a, const float* b) {
  P2 ca[4], cb[4], pts[24];
  lk_corners(a, ca);
  lk_corners(b, cb);
  int cnt = 0;
  float sxm = 0.f, sym = 0.f;
  for (int i = 0; i < 4; ++i) {
    const P2 p0 = ca[i], p1 = ca[(i + 1) & 3];
    for (int j = 0; j < 4; ++j) {
      const P2 q0 = cb[j], q1 = cb[(j + 1) & 3];
      const float s1 = lk_cross3(q0, p1, p0), s2 = lk_cross3(p1, q1, p0);
      const float s3 = lk_cross3(p0, q1, q0), s4 = lk_cross3(q1, p1, q0);
      if (!(s1 * s2 > 0.f && s3 * s4 > 0.f)) continue;           // proper crossing only
      const float s5 = lk_cross3(q1, p1, p0);
      P2 r;
      if (fabsf(s5 - s1) > 1e-8f) {
        r.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        r.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
      } else {                                                    // nearly parallel supporting lines
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float d = a0 * b1 - a1 * b0;
        r.x = (b0 * c1 - b1 * c0) / d;
        r.y = (a1 * c0 - a0 * c1) / d;
      }
      pts[cnt++] = r; sxm += r.x; sym += r.y;
    }
  }
  for (int k = 0; k < 4; ++k) {
    if (lk_in_box(a, cb[k])) { pts[cnt++] = cb[k]; sxm += cb[k].x; sym += cb[k].y; }
    if (lk_in_box(b, ca[k])) { pts[cnt++] = ca[k]; sxm += ca[k].x; sym += ca[k].y; }
  }
  float inter = 0.f;
  if (cnt > 0) {
    const float mx = sxm / cnt, my = sym / cnt;
    float ang[24];
    for (int k = 0; k < cnt; ++k) ang[k] = atan2f(pts[k].y - my, pts[k].x - mx);
    for (int k = 1; k < cnt; ++k) {                               // insertion sort by angle (stable, <= 24 points)
      const P2 p = pts[k];
      const float t = ang[k];
      int m = k - 1;
      while (m >= 0 && ang[m] > t) { pts[m + 1] = pts[m]; ang[m + 1] = ang[m]; --m; }
      pts[m + 1] = p; ang[m + 1] = t;
    }
    float area = 0.f;
    for (int k = 0; k + 1 < cnt; ++k) {
      const float ux = pts[k].x - pts[0].x, uy = pts[k].y - pts[0].y;
      const float vx = pts[k + 1].x - pts[0].x, vy = pts[k + 1].y - pts[0].y;
      area += ux * vy - uy * vx;
    }
    inter = fabsf(area) * 0.5f;
  }
  const float uni = a[3] * a[4] + b[3] * b[4] - inter;
  return inter / fmaxf(uni, 1e-8f);
}

__global__ void __launch_bounds__(256) boxes_iou_bev_kernel(const float* __restrict__ a, int n,
                                                            const float* __restrict__ b, int m,
                                                            float* __restrict__ out) {
  const int64_t total = (int64_t)n * m;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / m), j = (int)(t % m);                 // consecutive threads: consecutive b, same a
    float ba[7], bb[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { ba[k] = __ldg(a + (int64_t)i * 7 + k); bb[k] = __ldg(b + (int64_t)j * 7 + k); }
    out[t] = lk_iou_pair(ba, bb);
  }
}

extern "C" int lk_boxes_iou_bev(const float* d_a, int64_t n, const float* d_b, int64_t m, float* d_out,
                                lk_stream_t s) {
  LK_REQUIRE(n >= 0 && m >= 0 && n < (1LL << 31) && m < (1LL << 31), "lk_boxes_iou_bev: bad sizes");
  if (n == 0 || m == 0) return LK_OK;
  LK_REQUIRE(d_a && d_b && d_out, "lk_boxes_iou_bev: null pointer");
  boxes_iou_bev_kernel<<<lk_grid(n * m, 256, 8), 256, 0, (cudaStream_t)s>>>(d_a, (int)n, d_b, (int)m, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

// ---- rotated NMS: 64 x 64 overlap bitmasks + a greedy scan that stays on the device -----------------
// (reference: nms_kernel, iou3d_nms_kernel.cu:328-414, builds the same bitmask matrix, then copies it to
// the HOST and runs the greedy loop there, iou3d_nms.cpp nms_gpu.)  Boxes arrive sorted by descending
// score.  mask[i][w] bit b = IoU(box i, box 64 w + b) > thresh, for 64 w + b > i.
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, int n, float thresh,
                                                      int words, unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;                                   // only the upper triangle is ever read
  __shared__ float cbox[64][7];
  const int ncol = min(n - cb * 64, 64);
  if ((int)threadIdx.x < ncol) {
#pragma unroll
    for (int k = 0; k < 7; ++k) cbox[threadIdx.x][k] = __ldg(boxes + (int64_t)(cb * 64 + threadIdx.x) * 7 + k);
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= n) return;
  float bi[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) bi[k] = __ldg(boxes + (int64_t)i * 7 + k);
  unsigned long long bits = 0;
  for (int j = (rb == cb) ? (int)threadIdx.x + 1 : 0; j < ncol; ++j)
    if (lk_iou_pair(bi, cbox[j]) > thresh) bits |= 1ULL << j;
  mask[(int64_t)i * words + cb] = bits;
}

// one warp: keep[i] = box i is not suppressed by an earlier kept box; removed bits live in shared memory
__global__ void __launch_bounds__(32) nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int words,
                                                      uint8_t* __restrict__ keep) {
  extern __shared__ unsigned long long remv[];
  for (int w = threadIdx.x; w < words; w += 32) remv[w] = 0ULL;
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    const int wi = i >> 6;
    const bool dead = (remv[wi] >> (i & 63)) & 1ULL;     // same word for every lane: a broadcast read
    if (!dead) {
      for (int w = wi + (int)threadIdx.x; w < words; w += 32) remv[w] |= __ldg(mask + (int64_t)i * words + w);
    }
    if (threadIdx.x == 0) keep[i] = dead ? 0 : 1;
    __syncwarp();
  }
}

// centre-distance ("circle") NMS of CenterPoint (det3d/core/utils/circle_nms_jit.py:4-28): the same greedy
// rule with  suppress(i, j) = |c_i - c_j|^2 <= thresh  on the (x, y) centres; same bitmask + scan kernels
__global__ void __launch_bounds__(64) circle_mask_kernel(const float2* __restrict__ centers, int n, float thresh,
                                                         int words, unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;
  __shared__ float2 cc[64];
  const int ncol = min(n - cb * 64, 64);
  if ((int)threadIdx.x < ncol) cc[threadIdx.x] = __ldg(centers + cb * 64 + threadIdx.x);
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= n) return;
  const float2 ci = __ldg(centers + i);
  unsigned long long bits = 0;
  for (int j = (rb == cb) ? (int)threadIdx.x + 1 : 0; j < ncol; ++j) {
    const float dx = ci.x - cc[j].x, dy = ci.y - cc[j].y;
    if (dx * dx + dy * dy <= thresh) bits |= 1ULL << j;
  }
  mask[(int64_t)i * words + cb] = bits;
}

extern "C" int64_t lk_nms_bev_ws_bytes(int64_t n) { return n * ((n + 63) / 64) * 8 + 256; }

extern "C" int lk_nms_circle(const float* d_centers_sorted, int64_t n, float thresh, void* d_ws, int64_t ws_bytes,
                             uint8_t* d_keep, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && n <= 65536, "lk_nms_circle: at most 65536 boxes");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_centers_sorted && d_ws && d_keep && (uintptr_t)d_ws % 8 == 0 && (uintptr_t)d_centers_sorted % 8 == 0,
             "lk_nms_circle: null or misaligned pointer");
  const int words = (int)((n + 63) / 64);
  if (ws_bytes < lk_nms_bev_ws_bytes(n)) {
    lk_set_error("lk_nms_circle: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)lk_nms_bev_ws_bytes(n));
    return LK_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)s;
  dim3 grid((unsigned)words, (unsigned)words);
  circle_mask_kernel<<<grid, 64, 0, st>>>((const float2*)d_centers_sorted, (int)n, thresh, words, (unsigned long long*)d_ws);
  LK_LAUNCHED();
  nms_scan_kernel<<<1, 32, (size_t)words * 8, st>>>((const unsigned long long*)d_ws, (int)n, words, d_keep);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_nms_bev(const float* d_boxes_sorted, int64_t n, float thresh, void* d_ws, int64_t ws_bytes,
                          uint8_t* d_keep, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && n <= 65536, "lk_nms_bev: at most 65536 boxes");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_boxes_sorted && d_ws && d_keep && (uintptr_t)d_ws % 8 == 0, "lk_nms_bev: null or misaligned pointer");
  const int words = (int)((n + 63) / 64);
  if (ws_bytes < lk_nms_bev_ws_bytes(n)) {
    lk_set_error("lk_nms_bev: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)lk_nms_bev_ws_bytes(n));
    return LK_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)s;
  dim3 grid((unsigned)words, (unsigned)words);
  nms_mask_kernel<<<grid, 64, 0, st>>>(d_boxes_sorted, (int)n, thresh, words, (unsigned long long*)d_ws);
  LK_LAUNCHED();
  nms_scan_kernel<<<1, 32, (size_t)words * 8, st>>>((const unsigned long long*)d_ws, (int)n, words, d_keep);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_boxes_iou_bev_hostcheck(const float* a, int64_t n, const float* b, int64_t m, float* out) {
  LK_REQUIRE(n >= 0 && m >= 0 && (n == 0 || m == 0 || (a && b && out)), "lk_boxes_iou_bev_hostcheck: bad arguments");
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j) out[i * m + j] = lk_iou_pair(a + i * 7, b + j * 7);
  return LK_OK;
}
