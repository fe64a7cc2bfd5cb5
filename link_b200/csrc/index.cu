// Composite index-build entry points: one C call enqueues what the Python layer otherwise issues
// as 4-7 separate library calls (each ~15-40 us of interpreter + FFI time, which made the encoder
// forward host-bound).  No new kernels; nothing here synchronises or allocates.
//
//   lk_kmap_build : hash(input coords) -> table build -> kernel-map query [-> tile-skipping plan]
//                   (reference: the kmap branch of F.conv3d, nn/functional/conv.py:103-121)
//   lk_downsample : packed (b,x,y,z) keys of floor(coords / stride) -> radix sort -> unique ->
//                   unpack, count left on the device
//                   (reference: F.spdownsample, nn/functional/downsample.py:11-51)
#include "common.cuh"
#if defined(__SSE4_1__)
#include <smmintrin.h>
#endif

static inline int64_t ix_al(int64_t x) { return (x + 255) & ~(int64_t)255; }

#define LK_TRY(call)                \
  do {                              \
    int rc__ = (call);              \
    if (rc__ != LK_OK) return rc__; \
  } while (0)

extern "C" int64_t lk_kmap_build_ws_bytes(int64_t n_in, int64_t n_out) {
  return ix_al(n_in * 8) + ix_al(lk_conv_plan_ws_bytes(n_out));
}

/* d_table: caller-owned, lk_table_capacity(n_in) * 16 bytes (kept: other maps of the same level
 * reuse it with build_table = 0). */
extern "C" int lk_kmap_build(const int32_t* d_in_coords, int64_t n_in, const int32_t* d_out_coords,
                             int64_t n_out, const int32_t* d_offsets, int k, int subm,
                             void* d_table, int64_t capacity, int build_table, int32_t* d_nbr,
                             int32_t* d_plan_perm, uint32_t* d_plan_mask, void* d_ws,
                             int64_t ws_bytes, lk_stream_t s) {
  LK_REQUIRE(n_in >= 0 && n_out >= 0 && k > 0, "lk_kmap_build: bad sizes");
  if (n_out == 0) return LK_OK;
  LK_REQUIRE(d_in_coords && d_out_coords && d_offsets && d_table && d_nbr && d_ws,
             "lk_kmap_build: null pointer");
  if (ws_bytes < lk_kmap_build_ws_bytes(n_in, n_out)) {
    lk_set_error("lk_kmap_build: workspace %lld < %lld bytes", (long long)ws_bytes,
                 (long long)lk_kmap_build_ws_bytes(n_in, n_out));
    return LK_ENOSPC;
  }
  char* ws = (char*)d_ws;
  if (build_table) LK_TRY(lk_table_build_coords(d_in_coords, n_in, d_table, capacity, s));
  if (subm)
    LK_TRY(lk_kmap_query_subm(d_out_coords, n_out, d_offsets, k, d_table, capacity, d_nbr, s));
  else
    LK_TRY(lk_kmap_query(d_out_coords, n_out, d_offsets, k, d_table, capacity, d_nbr, s));
  if (d_plan_perm && d_plan_mask && k <= 32)
    LK_TRY(lk_conv_plan(d_nbr, n_out, k, d_offsets, d_plan_perm, d_plan_mask, ws + ix_al(n_in * 8),
                        lk_conv_plan_ws_bytes(n_out), s));
  return LK_OK;
}

extern "C" int64_t lk_downsample_ws_bytes(int64_t n) {
  return 2 * ix_al(n * 8) + ix_al(lk_sort_unique_ws_bytes(n));
}

/* d_out_coords [n,4] (first *d_num rows valid, ascending (b,x,y,z) == torch.unique(dim=0) order of the
 * reference), d_num device scalar. */
extern "C" int lk_downsample(const int32_t* d_coords, int64_t n, const lk_keyspec_t* spec, int key_bits,
                             int32_t* d_out_coords, int32_t* d_num, void* d_ws, int64_t ws_bytes,
                             lk_stream_t s) {
  LK_REQUIRE(n >= 0 && spec && d_num, "lk_downsample: bad arguments");
  if (n == 0) return lk_sort_unique(nullptr, 0, key_bits, nullptr, nullptr, nullptr, nullptr, nullptr, d_num,
                                    nullptr, 0, s);
  LK_REQUIRE(d_coords && d_out_coords && d_ws, "lk_downsample: null pointer");
  if (ws_bytes < lk_downsample_ws_bytes(n)) {
    lk_set_error("lk_downsample: workspace %lld < %lld bytes", (long long)ws_bytes,
                 (long long)lk_downsample_ws_bytes(n));
    return LK_ENOSPC;
  }
  char* ws = (char*)d_ws;
  uint64_t* keys = (uint64_t*)ws;
  uint64_t* uniq = (uint64_t*)(ws + ix_al(n * 8));
  void* sws = ws + 2 * ix_al(n * 8);
  LK_TRY(lk_pack_keys(d_coords, n, spec, keys, s));
  LK_TRY(lk_sort_unique(keys, n, key_bits, uniq, nullptr, nullptr, nullptr, nullptr, d_num, sws,
                        lk_sort_unique_ws_bytes(n), s));
  LK_TRY(lk_unpack_keys(uniq, d_num, n, spec, d_out_coords, s));
  return LK_OK;
}

/* Host helper: per-column minimum and maximum of a HOST int32 [n, 4] coordinate array (the bounds
 * that size the packed sort keys), one pass, while the scan's upload is in flight. */
extern "C" int lk_host_coord_bounds(const int32_t* h_coords, int64_t n, int32_t* lo4, int32_t* hi4) {
  LK_REQUIRE(n >= 0 && lo4 && hi4 && (n == 0 || h_coords), "lk_host_coord_bounds: null pointer");
  int32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
  if (n > 0) {
    for (int a = 0; a < 4; ++a) lo[a] = hi[a] = h_coords[a];
    int64_t i = 1;
#if defined(__SSE4_1__)
    // one 128-bit min / max per row, four independent accumulator pairs (the loop is bound by the
    // 1-cycle latency of pminsd / pmaxsd otherwise)
    __m128i l0 = _mm_loadu_si128((const __m128i*)h_coords), l1 = l0, l2 = l0, l3 = l0;
    __m128i h0 = l0, h1 = l0, h2 = l0, h3 = l0;
    for (; i + 4 <= n; i += 4) {
      const __m128i r0 = _mm_loadu_si128((const __m128i*)(h_coords + 4 * i));
      const __m128i r1 = _mm_loadu_si128((const __m128i*)(h_coords + 4 * i + 4));
      const __m128i r2 = _mm_loadu_si128((const __m128i*)(h_coords + 4 * i + 8));
      const __m128i r3 = _mm_loadu_si128((const __m128i*)(h_coords + 4 * i + 12));
      l0 = _mm_min_epi32(l0, r0); h0 = _mm_max_epi32(h0, r0);
      l1 = _mm_min_epi32(l1, r1); h1 = _mm_max_epi32(h1, r1);
      l2 = _mm_min_epi32(l2, r2); h2 = _mm_max_epi32(h2, r2);
      l3 = _mm_min_epi32(l3, r3); h3 = _mm_max_epi32(h3, r3);
    }
    l0 = _mm_min_epi32(_mm_min_epi32(l0, l1), _mm_min_epi32(l2, l3));
    h0 = _mm_max_epi32(_mm_max_epi32(h0, h1), _mm_max_epi32(h2, h3));
    _mm_storeu_si128((__m128i*)lo, l0);
    _mm_storeu_si128((__m128i*)hi, h0);
#endif
    for (; i < n; ++i) {
      const int32_t* r = h_coords + 4 * i;
      for (int a = 0; a < 4; ++a) {
        lo[a] = r[a] < lo[a] ? r[a] : lo[a];
        hi[a] = r[a] > hi[a] ? r[a] : hi[a];
      }
    }
  }
  for (int a = 0; a < 4; ++a) { lo4[a] = lo[a]; hi4[a] = hi[a]; }
  return LK_OK;
}

/* Reference-layout kernel map -> the output-stationary map the conv kernels consume.
 * The reference keeps a kernel map as a pair list (nn/functional/conv.py:114-121): neighbor_map [P, 2]
 * = (input row, output row) ordered by offset, neighbor_offset [K] = pairs per offset (on the HOST,
 * convolution_cuda.cu:53-57).  d_nbr [K, n_rows] gets, for every offset k and row r, the partner row of
 * the pair whose column `row_col` equals r, or -1.  row_col = 1: rows are output rows (forward,
 * weight gradient); row_col = 0: rows are input rows (transposed conv, input gradient).
 * identity_mid = 1 reproduces the reference's shortcut for the centre offset of an odd kernel with
 * n_in == n_out (convolution_cuda.cu:74-88: W[mid] is applied to every row, its pairs are not read). */
struct PairPrefix { int32_t start[129]; };

__global__ void __launch_bounds__(256) kmap_from_pairs_kernel(const int32_t* __restrict__ pairs, PairPrefix pre, int k,
                                                              int64_t n_rows, int row_col, int mid,
                                                              int32_t* __restrict__ nbr) {
  const int64_t total = pre.start[k];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = k;                          // offset of pair p: last start <= p
    while (hi - lo > 1) {
      const int m = (lo + hi) >> 1;
      if (pre.start[m] <= p) lo = m; else hi = m;
    }
    if (lo == mid) continue;
    const int32_t row = __ldg(pairs + 2 * p + row_col), other = __ldg(pairs + 2 * p + (1 - row_col));
    if (row >= 0 && row < n_rows) nbr[(int64_t)lo * n_rows + row] = other;
  }
  if (mid >= 0)
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
      nbr[(int64_t)mid * n_rows + r] = (int32_t)r;
}

extern "C" int lk_kmap_from_pairs(const int32_t* d_nbmaps, const int32_t* h_nbsizes, int k, int64_t n_rows,
                                  int row_col, int identity_mid, int32_t* d_nbr, lk_stream_t s) {
  LK_REQUIRE(k > 0 && k <= 128 && n_rows >= 0 && (row_col == 0 || row_col == 1) && h_nbsizes,
             "lk_kmap_from_pairs: bad arguments (1 <= K <= 128)");
  if (n_rows == 0) return LK_OK;
  LK_REQUIRE(d_nbr, "lk_kmap_from_pairs: null output");
  PairPrefix pre;
  int64_t total = 0;
  for (int i = 0; i < k; ++i) {
    LK_REQUIRE(h_nbsizes[i] >= 0, "lk_kmap_from_pairs: negative pair count");
    pre.start[i] = (int32_t)total;
    total += h_nbsizes[i];
  }
  LK_REQUIRE(total < (1LL << 31), "lk_kmap_from_pairs: more than 2^31 pairs");
  pre.start[k] = (int32_t)total;
  LK_REQUIRE(total == 0 || d_nbmaps, "lk_kmap_from_pairs: null pair list");
  cudaStream_t st = (cudaStream_t)s;
  LK_CUDA(cudaMemsetAsync(d_nbr, 0xFF, (size_t)k * n_rows * sizeof(int32_t), st));
  lk_count_launch();
  const int mid = identity_mid ? k / 2 : -1;
  const int64_t work = total > n_rows ? total : n_rows;
  kmap_from_pairs_kernel<<<lk_grid(work, 256, 8), 256, 0, st>>>(d_nbmaps, pre, k, n_rows, row_col, mid, d_nbr);
  LK_LAUNCHED();
  return LK_OK;
}

/* Output sites of a strided / padded sparse conv that creates new active sites (spconv SparseConv3d,
 * detection/det3d/models/backbones/scn.py:494-566): input site i reaches output o through tap kappa when
 * i + p - kappa = s * o with 0 <= o < out_shape.  Per axis only the taps kappa = (i + p) mod s + s * t
 * (t < ceil(k / s)) can divide, so a voxel has at most cap = prod ceil(k_a / s_a) candidates (8 for
 * k = 3, s = 2) instead of the K = 27 the tensor-op formulation expands, filters and compacts with a
 * host round trip.  d_indices [n,4] = (batch, z, y, x); d_cand [n * cap, 4] = (x, y, z, batch) of every
 * candidate, invalid slots = (0, 0, 0, batch_size) -- one past the last batch, so that after a sort by
 * (batch, z, y, x) they collapse into ONE trailing key -- and *d_any_invalid is set if there is one. */
struct StridedSpec { int k[3], s[3], p[3], out[3], cap[3]; };

__global__ void __launch_bounds__(256) strided_candidates_kernel(const int4* __restrict__ indices, int64_t n,
                                                                 StridedSpec sp, int batch_size,
                                                                 int4* __restrict__ cand, int* __restrict__ any_invalid) {
  const int cap = sp.cap[0] * sp.cap[1] * sp.cap[2];
  const int64_t total = n * cap;
  bool bad = false;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / cap;
    int slot = (int)(t - i * cap);
    const int4 v = __ldg(indices + i);                 // (b, z, y, x)
    const int pos[3] = {v.y, v.z, v.w};
    int o[3];
    bool ok = true;
#pragma unroll
    for (int a = 2; a >= 0; --a) {
      const int ta = slot % sp.cap[a];
      slot /= sp.cap[a];
      const int num = pos[a] + sp.p[a];
      const int kappa = (num % sp.s[a]) + sp.s[a] * ta;           // num >= 0
      o[a] = (num - kappa) / sp.s[a];
      ok = ok && kappa < sp.k[a] && num >= kappa && o[a] < sp.out[a];
    }
    cand[t] = ok ? make_int4(o[2], o[1], o[0], v.x) : make_int4(0, 0, 0, batch_size);
    bad = bad || !ok;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *any_invalid = 1;
}

extern "C" int lk_strided_candidates(const int32_t* d_indices, int64_t n, const int32_t* kernel3,
                                     const int32_t* stride3, const int32_t* padding3, const int32_t* out_shape3,
                                     int batch_size, int32_t* d_cand, int32_t* d_any_invalid, lk_stream_t s) {
  LK_REQUIRE(n >= 0 && kernel3 && stride3 && padding3 && out_shape3 && batch_size > 0,
             "lk_strided_candidates: bad arguments");
  StridedSpec sp;
  for (int a = 0; a < 3; ++a) {
    LK_REQUIRE(kernel3[a] > 0 && stride3[a] > 0 && padding3[a] >= 0 && out_shape3[a] > 0,
               "lk_strided_candidates: kernel / stride / shape must be positive");
    sp.k[a] = kernel3[a]; sp.s[a] = stride3[a]; sp.p[a] = padding3[a]; sp.out[a] = out_shape3[a];
    sp.cap[a] = (kernel3[a] + stride3[a] - 1) / stride3[a];
  }
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_indices && d_cand && d_any_invalid && (uintptr_t)d_indices % 16 == 0 && (uintptr_t)d_cand % 16 == 0,
             "lk_strided_candidates: null or misaligned pointer");
  cudaStream_t st = (cudaStream_t)s;
  LK_CUDA(cudaMemsetAsync(d_any_invalid, 0, sizeof(int32_t), st));
  lk_count_launch();
  const int cap = sp.cap[0] * sp.cap[1] * sp.cap[2];
  strided_candidates_kernel<<<lk_grid(n * cap, 256, 8), 256, 0, st>>>((const int4*)d_indices, n, sp, batch_size,
                                                                     (int4*)d_cand, d_any_invalid);
  LK_LAUNCHED();
  return LK_OK;
}
