// Key packing, stable LSD radix sort and unique: the device-side replacement for the
// torch.unique(dim=0) calls on the reference hot path (segmentation/core/models/utils.py:47,
// nn/functional/downsample.py:48-50).  Hand-written (no CUB/Thrust): 8-bit digits, one
// histogram / scan / scatter triple per pass, number of passes = ceil(key_bits / 8) where
// key_bits comes from the host-known coordinate bounds, so block keys (~20 bits) need 3
// passes instead of 8.
#include "common.cuh"

#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

// ---------------------------------------------------------------- key packing
struct KeySpecDev {
  int div[3], mul[3], order[4], lo[4], shift[4];
  unsigned long long mask[4];
};

static int make_spec(const lk_keyspec_t* spec, KeySpecDev* d) {
  int total = 0;
  for (int f = 0; f < 4; ++f) {
    if (spec->bits[f] < 0 || spec->bits[f] > 32) return -1;
    total += spec->bits[f];
  }
  if (total > 64) return -1;
  bool seen[4] = {false, false, false, false};
  for (int f = 0; f < 4; ++f) {
    int a = spec->order[f];
    if (a < 0 || a > 3 || seen[a]) return -1;
    seen[a] = true;
  }
  for (int a = 0; a < 3; ++a) {
    if (spec->div[a] < 1) return -1;
    d->div[a] = spec->div[a];
    d->mul[a] = spec->mul[a];
  }
  int sh = total;
  for (int f = 0; f < 4; ++f) {  // most significant field first
    int a = spec->order[f];
    sh -= spec->bits[a];
    d->order[f] = a;
    d->shift[a] = sh;
    d->lo[a] = spec->lo[a];
    d->mask[a] = spec->bits[a] >= 64 ? ~0ULL : ((1ULL << spec->bits[a]) - 1ULL);
  }
  return total;
}

__device__ __forceinline__ unsigned long long pack_fields(const KeySpecDev& sp, int q0, int q1,
                                                         int q2, int q3) {
  unsigned long long k = 0;
  k |= ((unsigned long long)(unsigned)(q0 - sp.lo[0]) & sp.mask[0]) << sp.shift[0];
  k |= ((unsigned long long)(unsigned)(q1 - sp.lo[1]) & sp.mask[1]) << sp.shift[1];
  k |= ((unsigned long long)(unsigned)(q2 - sp.lo[2]) & sp.mask[2]) << sp.shift[2];
  k |= ((unsigned long long)(unsigned)(q3 - sp.lo[3]) & sp.mask[3]) << sp.shift[3];
  return k;
}

__global__ void __launch_bounds__(256) pack_keys_kernel(const int4* __restrict__ coords, int64_t n,
                                                        KeySpecDev sp,
                                                        unsigned long long* __restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
    keys[i] = pack_fields(sp, lk_floordiv(c.x, sp.div[0]), lk_floordiv(c.y, sp.div[1]),
                          lk_floordiv(c.z, sp.div[2]), c.w);
  }
}

__global__ void __launch_bounds__(256) unpack_keys_kernel(const unsigned long long* __restrict__ keys,
                                                          const int* __restrict__ d_count,
                                                          int64_t n, KeySpecDev sp,
                                                          int4* __restrict__ coords) {
  int64_t m = d_count ? (int64_t)*d_count : n;
  if (m > n) m = n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long k = keys[i];
    int4 c;
    c.x = ((int)((k >> sp.shift[0]) & sp.mask[0]) + sp.lo[0]) * sp.mul[0];
    c.y = ((int)((k >> sp.shift[1]) & sp.mask[1]) + sp.lo[1]) * sp.mul[1];
    c.z = ((int)((k >> sp.shift[2]) & sp.mask[2]) + sp.lo[2]) * sp.mul[2];
    c.w = (int)((k >> sp.shift[3]) & sp.mask[3]) + sp.lo[3];
    coords[i] = c;
  }
}

extern "C" int lk_pack_keys(const int32_t* d_coords, int64_t n, const lk_keyspec_t* spec,
                            uint64_t* d_keys, lk_stream_t s) {
  KeySpecDev sp;
  LK_REQUIRE(spec && make_spec(spec, &sp) >= 0, "lk_pack_keys: invalid key spec");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_coords && d_keys && n > 0, "lk_pack_keys: bad arguments");
  pack_keys_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>(
      (const int4*)d_coords, n, sp, (unsigned long long*)d_keys);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_unpack_keys(const uint64_t* d_keys, const int32_t* d_count, int64_t n,
                              const lk_keyspec_t* spec, int32_t* d_coords, lk_stream_t s) {
  KeySpecDev sp;
  LK_REQUIRE(spec && make_spec(spec, &sp) >= 0, "lk_unpack_keys: invalid key spec");
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_keys && d_coords && n > 0, "lk_unpack_keys: bad arguments");
  unpack_keys_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>(
      (const unsigned long long*)d_keys, d_count, n, sp, (int4*)d_coords);
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- block-neighbour table
// Sorted unique keys double as the dictionary: a neighbour block is found by re-packing the
// offset coordinate and binary-searching the (L1/L2 resident) key array.
__global__ void __launch_bounds__(256) block_neighbors_kernel(
    const unsigned long long* __restrict__ uniq, const int* __restrict__ d_num, int64_t capacity,
    KeySpecDev sp, const int* __restrict__ offsets, int R, int* __restrict__ nbr,
    float4* __restrict__ zero_buf, int zero_row_vec) {
  lk_pdl_enter();
  int64_t m = *d_num;
  if (m > capacity) m = capacity;
  int64_t total = m * R;
  if (zero_buf) {            // optional: clear the first M rows of the block-sum buffer in the same launch
    const int64_t zt = m * zero_row_vec;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < zt;
         t += (int64_t)gridDim.x * blockDim.x)
      zero_buf[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / R;
    int k = (int)(t - b * R);
    unsigned long long key = uniq[b];
    long long q[3];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      q[a] = (long long)((key >> sp.shift[a]) & sp.mask[a]) + offsets[3 * k + a];
      ok = ok && q[a] >= 0 && (unsigned long long)q[a] <= sp.mask[a];
    }
    int res = -1;
    if (ok) {
      unsigned long long want = key;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        want &= ~(sp.mask[a] << sp.shift[a]);
        want |= ((unsigned long long)q[a]) << sp.shift[a];
      }
      int64_t lo = 0, hi = m;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (uniq[mid] < want) lo = mid + 1; else hi = mid;
      }
      if (lo < m && uniq[lo] == want) res = (int)lo;
    }
    nbr[t] = res;
  }
}

extern "C" int lk_block_neighbors(const uint64_t* d_unique, const int32_t* d_num, int64_t capacity,
                                  const lk_keyspec_t* spec, const int32_t* d_offsets, int r3,
                                  int32_t* d_nbr, lk_stream_t s) {
  KeySpecDev sp;
  LK_REQUIRE(spec && make_spec(spec, &sp) >= 0, "lk_block_neighbors: invalid key spec");
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(d_unique && d_num && d_offsets && d_nbr && r3 > 0, "lk_block_neighbors: bad arguments");
  block_neighbors_kernel<<<lk_grid(capacity * r3, 256, 8), 256, 0, (cudaStream_t)s>>>(
      (const unsigned long long*)d_unique, d_num, capacity, sp, d_offsets, r3, d_nbr, nullptr, 0);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_block_neighbors_zero(const uint64_t* d_unique, const int32_t* d_num, int64_t capacity,
                                       const lk_keyspec_t* spec, const int32_t* d_offsets, int r3,
                                       int32_t* d_nbr, float* d_zero, int row_floats, lk_stream_t s) {
  KeySpecDev sp;
  LK_REQUIRE(spec && make_spec(spec, &sp) >= 0, "lk_block_neighbors_zero: invalid key spec");
  if (capacity == 0) return LK_OK;
  LK_REQUIRE(d_unique && d_num && d_offsets && d_nbr && r3 > 0 && d_zero && row_floats > 0 &&
                 row_floats % 4 == 0 && (uintptr_t)d_zero % 16 == 0,
             "lk_block_neighbors_zero: bad arguments");
  LK_PDL_LAUNCH(block_neighbors_kernel, lk_grid(capacity * r3, 256, 8), 256, 0, (cudaStream_t)s,
                (const unsigned long long*)d_unique, d_num, capacity, sp, d_offsets, r3, d_nbr, (float4*)d_zero,
                row_floats / 4);
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- radix sort
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(
    const unsigned long long* __restrict__ keys, int64_t n, int shift, unsigned* __restrict__ hist,
    int T) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = base + j * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * T + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `len` unsigned counters by one 1024-thread CTA; optionally stores the
// total to *d_total (int) and total2[total] = tail_value (used for the segment sentinel).
__global__ void __launch_bounds__(1024) scan_single_cta_kernel(unsigned* data, int64_t len,
                                                               int* d_total, int* d_tail_array,
                                                               int tail_value) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry_s;
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < len; base += 1024 * 4) {
    int64_t i0 = base + (int64_t)tid * 4;
    unsigned v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (i0 + j < len) ? data[i0 + j] : 0u;
    unsigned tsum = v[0] + v[1] + v[2] + v[3];
    unsigned incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_sums[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    unsigned carry = carry_s;
    unsigned excl = carry + warp_sums[warp] + (incl - tsum);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < len) data[i0 + j] = excl;
      excl += v[j];
    }
    __syncthreads();
    if (tid == 1023) carry_s = excl;  // excl now = inclusive total through this chunk
    __syncthreads();
  }
  if (tid == 0) {
    unsigned total = carry_s;
    if (d_total) *d_total = (int)total;
    if (d_tail_array) d_tail_array[total] = tail_value;
  }
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(
    const unsigned long long* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
    unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out, int64_t n,
    int shift, const unsigned* __restrict__ offs, int T) {
  __shared__ unsigned cnt[RS_WARPS][257];
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < RS_WARPS * 257; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();

  int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
  unsigned long long key[RS_ITEMS];
  unsigned val[RS_ITEMS];
  unsigned short dig[RS_ITEMS];
  unsigned rank[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = wbase + j * 32 + lane;
    bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0ULL;
    val[j] = valid ? (vals_in ? vals_in[i] : (unsigned)i) : 0u;
    dig[j] = valid ? (unsigned short)((unsigned)(key[j] >> shift) & 255u) : (unsigned short)256;
  }
  // stable rank of each item among equal digits of this warp (match-any multi-split)
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    unsigned d = dig[j];
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    unsigned before = __popc(peers & ((1u << lane) - 1u));
    unsigned base = 0;
    if (lane == leader) {
      base = cnt[warp][d];
      cnt[warp][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = base + before;
    __syncwarp();
  }
  __syncthreads();
  // digit `tid`: global base of this tile, then running offset over the warps
  {
    unsigned run = offs[(int64_t)tid * T + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      unsigned c = cnt[w][tid];
      cnt[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    if (dig[j] < 256) {
      unsigned pos = cnt[warp][dig[j]] + rank[j];
      keys_out[pos] = key[j];
      vals_out[pos] = val[j];
    }
  }
}

// ---------------------------------------------------------------- unique
// blocked arrangement: thread t owns items [t*8, t*8+8) of its tile
__global__ void __launch_bounds__(RS_THREADS) uniq_count_kernel(
    const unsigned long long* __restrict__ keys, int64_t n, unsigned* __restrict__ tile_heads) {
  lk_pdl_enter();
  __shared__ unsigned wsum[RS_WARPS];
  int64_t i0 = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
  unsigned c = 0;
  unsigned long long prev = (i0 > 0 && i0 - 1 < n) ? keys[i0 - 1] : 0ULL;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = i0 + j;
    if (i < n) {
      unsigned long long k = keys[i];
      c += (i == 0 || k != prev) ? 1u : 0u;
      prev = k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) t += wsum[w];
    tile_heads[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(RS_THREADS) uniq_write_kernel(
    const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, int64_t n,
    const unsigned* __restrict__ tile_base, unsigned long long* __restrict__ uniq,
    int* __restrict__ inverse, int* __restrict__ order, int* __restrict__ seg,
    int* __restrict__ rank_sorted) {
  __shared__ unsigned wsum[RS_WARPS];
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t i0 = (int64_t)blockIdx.x * RS_TILE + (int64_t)tid * RS_ITEMS;
  unsigned long long k[RS_ITEMS];
  unsigned head[RS_ITEMS];
  unsigned c = 0;
  unsigned long long prev = (i0 > 0 && i0 - 1 < n) ? keys[i0 - 1] : 0ULL;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = i0 + j;
    k[j] = (i < n) ? keys[i] : 0ULL;
    head[j] = (i < n && (i == 0 || k[j] != prev)) ? 1u : 0u;
    prev = k[j];
    c += head[j];
  }
  unsigned incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  unsigned wbase = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) wbase += (w < warp) ? wsum[w] : 0u;
  // rank of the LAST head at or before item j = (heads so far) - 1
  unsigned running = tile_base[blockIdx.x] + wbase + (incl - c);
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = i0 + j;
    if (i < n) {
      running += head[j];
      unsigned r = running - 1u;
      unsigned v = vals[i];
      if (head[j]) {
        if (uniq) uniq[r] = k[j];
        if (seg) seg[r] = (int)i;
      }
      if (inverse) inverse[v] = (int)r;
      if (order) order[i] = (int)v;
      if (rank_sorted) rank_sorted[i] = (int)r;
    }
  }
}

__global__ void __launch_bounds__(256) seg_counts_kernel(const int* __restrict__ seg,
                                                         const int* __restrict__ d_num, int64_t cap,
                                                         int* __restrict__ counts) {
  int64_t m = *d_num;
  if (m > cap) m = cap;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m;
       i += (int64_t)gridDim.x * blockDim.x)
    counts[i] = seg[i + 1] - seg[i];
}

// ---------------------------------------------------------------- fused variants (T <= RS_MAX_FUSED_TILES)
// The single-CTA scan launches are folded into their consumers: every scatter CTA derives its own
// digit bases from the (L2-resident) [T][256] tile-histogram matrix, and accumulates the NEXT pass's
// tile histograms with atomics on the destination tile of each element, so a P-pass sort is
// 1 histogram + P scatter launches instead of 3P.
#define RS_MAX_FUSED_TILES 1024

__global__ void __launch_bounds__(RS_THREADS) radix_hist_tm_kernel(
    const unsigned long long* __restrict__ keys, int64_t n, int shift, unsigned* __restrict__ hist) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = base + j * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];   // tile-major
}

// first pass of a sort that starts from coordinates: pack the keys (stored for the scatter passes)
// and build the tile histograms of digit 0 in one launch
__global__ void __launch_bounds__(RS_THREADS) radix_pack_hist_tm_kernel(
    const int4* __restrict__ coords, int64_t n, KeySpecDev sp, unsigned long long* __restrict__ keys,
    unsigned* __restrict__ hist) {
  lk_pdl_enter();
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = base + j * RS_THREADS + threadIdx.x;
    if (i < n) {
      int4 c = coords[i];
      unsigned long long k = pack_fields(sp, lk_floordiv(c.x, sp.div[0]), lk_floordiv(c.y, sp.div[1]),
                                         lk_floordiv(c.z, sp.div[2]), c.w);
      keys[i] = k;
      atomicAdd(&h[(unsigned)k & 255u], 1u);
    }
  }
  __syncthreads();
  hist[(int64_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];   // tile-major
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter_fused_kernel(
    const unsigned long long* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
    unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out, int64_t n,
    int shift, const unsigned* __restrict__ hist /*[T][256]*/, unsigned* hist_next /*or NULL*/,
    int T) {
  lk_pdl_enter();
  __shared__ unsigned cnt[RS_WARPS][257];
  __shared__ unsigned wsum[RS_WARPS];
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < RS_WARPS * 257; i += RS_THREADS) (&cnt[0][0])[i] = 0;

  // this thread's keys first: their loads overlap the histogram walk below
  int64_t wbase_i = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
  unsigned long long key[RS_ITEMS];
  unsigned val[RS_ITEMS];
  unsigned short dig[RS_ITEMS];
  unsigned rank[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = wbase_i + j * 32 + lane;
    bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0ULL;
    val[j] = valid ? (vals_in ? vals_in[i] : (unsigned)i) : 0u;
    dig[j] = valid ? (unsigned short)((unsigned)(key[j] >> shift) & 255u) : (unsigned short)256;
  }

  // ---- digit `tid`: elements of this digit in earlier tiles, and in all tiles ----
  unsigned below = 0, total = 0;
  // 16 independent loads per round trip to L2: this column walk is a pure latency chain (T = 59
  // tiles at the bench size) and sits on the critical path of every pass
  for (int t = 0; t < T; t += 16) {
    unsigned a[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) a[u] = (t + u < T) ? __ldg(hist + (int64_t)(t + u) * 256 + tid) : 0u;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      total += a[u];
      below += (t + u < (int)blockIdx.x) ? a[u] : 0u;
    }
  }
  // exclusive scan of `total` over the 256 digits
  unsigned incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  unsigned wbase = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) wbase += (w < warp) ? wsum[w] : 0u;
  const unsigned my_base = wbase + (incl - total) + below;      // global base of (digit tid, this tile)

#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    unsigned d = dig[j];
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    unsigned before = __popc(peers & ((1u << lane) - 1u));
    unsigned base = 0;
    if (lane == leader) {
      base = cnt[warp][d];
      cnt[warp][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = base + before;
    __syncwarp();
  }
  __syncthreads();
  {
    unsigned run = my_base;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      unsigned c = cnt[w][tid];
      cnt[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    if (dig[j] < 256) {
      unsigned pos = cnt[warp][dig[j]] + rank[j];
      keys_out[pos] = key[j];
      vals_out[pos] = val[j];
      if (hist_next)
        atomicAdd(hist_next + (int64_t)(pos / RS_TILE) * 256 + ((unsigned)(key[j] >> (shift + 8)) & 255u), 1u);
    }
  }
}

// unique: write pass with the tile prefix computed in-kernel; CTA 0 publishes M and the sentinel
__global__ void __launch_bounds__(RS_THREADS) uniq_write_fused_kernel(
    const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, int64_t n,
    const unsigned* __restrict__ tile_heads, int T, unsigned long long* __restrict__ uniq,
    int* __restrict__ inverse, int* __restrict__ order, int* __restrict__ seg, int* __restrict__ d_num,
    int* __restrict__ rank_sorted) {
  lk_pdl_enter();
  __shared__ unsigned wsum[RS_WARPS];
  __shared__ unsigned red[2][RS_WARPS];
  __shared__ unsigned tile_base_s;
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned below = 0, total = 0;
  for (int t = tid; t < T; t += RS_THREADS) {
    unsigned a = __ldg(tile_heads + t);
    total += a;
    below += t < (int)blockIdx.x ? a : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    below += __shfl_xor_sync(0xffffffffu, below, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  if (lane == 0) { red[0][warp] = below; red[1][warp] = total; }
  __syncthreads();
  if (tid == 0) {
    unsigned b = 0, tt = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { b += red[0][w]; tt += red[1][w]; }
    tile_base_s = b;
    if (blockIdx.x == 0) {
      if (d_num) *d_num = (int)tt;
      if (seg) seg[tt] = (int)n;
    }
  }
  __syncthreads();
  const unsigned tile_base = tile_base_s;

  int64_t i0 = (int64_t)blockIdx.x * RS_TILE + (int64_t)tid * RS_ITEMS;
  unsigned long long k[RS_ITEMS];
  unsigned head[RS_ITEMS];
  unsigned c = 0;
  unsigned long long prev = (i0 > 0 && i0 - 1 < n) ? keys[i0 - 1] : 0ULL;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = i0 + j;
    k[j] = (i < n) ? keys[i] : 0ULL;
    head[j] = (i < n && (i == 0 || k[j] != prev)) ? 1u : 0u;
    prev = k[j];
    c += head[j];
  }
  unsigned incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  unsigned wbase = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) wbase += (w < warp) ? wsum[w] : 0u;
  unsigned running = tile_base + wbase + (incl - c);
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = i0 + j;
    if (i < n) {
      running += head[j];
      unsigned r = running - 1u;
      unsigned v = vals[i];
      if (head[j]) {
        if (uniq) uniq[r] = k[j];
        if (seg) seg[r] = (int)i;
      }
      if (inverse) inverse[v] = (int)r;
      if (order) order[i] = (int)v;
      if (rank_sorted) rank_sorted[i] = (int)r;
    }
  }
}

static inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t lk_sort_unique_ws_bytes(int64_t n) {
  int64_t T = (n + RS_TILE - 1) / RS_TILE;
  if (T < 1) T = 1;
  return 2 * align256(n * 8) + 2 * align256(n * 4) + align256(8 * 256 * T * 4) +
         align256((T + 1) * 4) + align256((n + 1) * 4) + align256(4);
}

static int sort_unique_impl(const uint64_t* d_keys, const int32_t* d_coords, const lk_keyspec_t* spec,
                            int64_t n, int key_bits, uint64_t* d_unique, int32_t* d_inverse,
                            int32_t* d_order, int32_t* d_seg, int32_t* d_counts, int32_t* d_num,
                            int32_t* d_sorted_rank, void* d_ws, int64_t ws_bytes, lk_stream_t s);

extern "C" int lk_sort_unique_ex(const uint64_t* d_keys, int64_t n, int key_bits,
                                 uint64_t* d_unique, int32_t* d_inverse, int32_t* d_order,
                                 int32_t* d_seg, int32_t* d_counts, int32_t* d_num,
                                 int32_t* d_sorted_rank, void* d_ws, int64_t ws_bytes,
                                 lk_stream_t s) {
  return sort_unique_impl(d_keys, nullptr, nullptr, n, key_bits, d_unique, d_inverse, d_order, d_seg,
                          d_counts, d_num, d_sorted_rank, d_ws, ws_bytes, s);
}

extern "C" int lk_sort_unique_coords(const int32_t* d_coords, const lk_keyspec_t* spec, int64_t n,
                                     int key_bits, uint64_t* d_unique, int32_t* d_inverse,
                                     int32_t* d_order, int32_t* d_seg, int32_t* d_counts,
                                     int32_t* d_num, int32_t* d_sorted_rank, void* d_ws,
                                     int64_t ws_bytes, lk_stream_t s) {
  LK_REQUIRE(spec && (d_coords || n == 0), "lk_sort_unique_coords: null coords/spec");
  return sort_unique_impl(nullptr, d_coords, spec, n, key_bits, d_unique, d_inverse, d_order, d_seg,
                          d_counts, d_num, d_sorted_rank, d_ws, ws_bytes, s);
}

static int sort_unique_impl(const uint64_t* d_keys, const int32_t* d_coords, const lk_keyspec_t* spec,
                            int64_t n, int key_bits, uint64_t* d_unique, int32_t* d_inverse,
                            int32_t* d_order, int32_t* d_seg, int32_t* d_counts, int32_t* d_num,
                            int32_t* d_sorted_rank, void* d_ws, int64_t ws_bytes, lk_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  LK_REQUIRE(n >= 0 && n < (1LL << 31), "lk_sort_unique: n out of range");
  LK_REQUIRE(key_bits >= 0 && key_bits <= 64, "lk_sort_unique: key_bits out of range");
  if (n == 0) {
    if (d_num) { LK_CUDA(cudaMemsetAsync(d_num, 0, 4, st)); lk_count_launch(); }
    if (d_seg) { LK_CUDA(cudaMemsetAsync(d_seg, 0, 4, st)); lk_count_launch(); }
    return LK_OK;
  }
  LK_REQUIRE((d_keys || d_coords) && d_ws, "lk_sort_unique: null keys/workspace");
  KeySpecDev spd;
  if (d_coords) LK_REQUIRE(make_spec(spec, &spd) >= 0, "lk_sort_unique_coords: invalid key spec");
  if (ws_bytes < lk_sort_unique_ws_bytes(n)) {
    lk_set_error("lk_sort_unique: workspace %lld < %lld bytes", (long long)ws_bytes,
                 (long long)lk_sort_unique_ws_bytes(n));
    return LK_ENOSPC;
  }
  int T = (int)((n + RS_TILE - 1) / RS_TILE);
  char* p = (char*)d_ws;
  unsigned long long* ka = (unsigned long long*)p; p += align256(n * 8);
  unsigned long long* kb = (unsigned long long*)p; p += align256(n * 8);
  unsigned* va = (unsigned*)p; p += align256(n * 4);
  unsigned* vb = (unsigned*)p; p += align256(n * 4);
  unsigned* hist = (unsigned*)p; p += align256(8 * 256 * (int64_t)T * 4);
  unsigned* tile_heads = (unsigned*)p; p += align256(((int64_t)T + 1) * 4);
  int* seg_tmp = (int*)p; p += align256((n + 1) * 4);
  int* num_tmp = (int*)p;

  int passes = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  const unsigned long long* kin = (const unsigned long long*)d_keys;
  const unsigned* vin = nullptr;
  unsigned long long* kout = ka;
  unsigned* vout = va;
  const bool fused = T <= RS_MAX_FUSED_TILES;
  if (d_coords) {            // keys are packed into kb by the first kernel (fused with the histogram
    kin = kb;                // when the fused path applies); the first scatter then reads kb, writes ka
    if (!fused) {
      pack_keys_kernel<<<lk_grid(n, 256, 8), 256, 0, st>>>((const int4*)d_coords, n, spd, kb);
      LK_LAUNCHED();
    }
  }
  if (fused) {
    const int64_t hsz = 256 * (int64_t)T;          // one [T][256] matrix per pass
    if (passes > 1) {
      LK_CUDA(cudaMemsetAsync(hist + hsz, 0, (size_t)(passes - 1) * hsz * sizeof(unsigned), st));
      lk_count_launch();
    }
    if (d_coords)
      LK_PDL_LAUNCH(radix_pack_hist_tm_kernel, T, RS_THREADS, 0, st, (const int4*)d_coords, n, spd, kb, hist);
    else
      radix_hist_tm_kernel<<<T, RS_THREADS, 0, st>>>(kin, n, 0, hist);
    LK_LAUNCHED();
    for (int pss = 0; pss < passes; ++pss) {
      LK_PDL_LAUNCH(radix_scatter_fused_kernel, T, RS_THREADS, 0, st,
                    kin, vin, kout, vout, n, 8 * pss, hist + pss * hsz,
                    pss + 1 < passes ? hist + (pss + 1) * hsz : nullptr, T);
      LK_LAUNCHED();
      kin = kout; vin = vout;
      kout = (kout == ka) ? kb : ka;
      vout = (vout == va) ? vb : va;
    }
  }
  for (int pss = 0; pss < (fused ? 0 : passes); ++pss) {
    int shift = 8 * pss;
    radix_hist_kernel<<<T, RS_THREADS, 0, st>>>(kin, n, shift, hist, T);
    LK_LAUNCHED();
    scan_single_cta_kernel<<<1, 1024, 0, st>>>(hist, 256 * (int64_t)T, nullptr, nullptr, 0);
    LK_LAUNCHED();
    radix_scatter_kernel<<<T, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist, T);
    LK_LAUNCHED();
    kin = kout; vin = vout;
    kout = (kout == ka) ? kb : ka;
    vout = (vout == va) ? vb : va;
  }
  int* seg = d_seg ? d_seg : seg_tmp;
  int* num = d_num ? d_num : num_tmp;
  LK_PDL_LAUNCH(uniq_count_kernel, T, RS_THREADS, 0, st, kin, n, tile_heads);
  LK_LAUNCHED();
  if (fused) {
    LK_PDL_LAUNCH(uniq_write_fused_kernel, T, RS_THREADS, 0, st, kin, vin, n, tile_heads, T,
                  (unsigned long long*)d_unique, d_inverse, d_order, seg, num, d_sorted_rank);
    LK_LAUNCHED();
  } else {
    scan_single_cta_kernel<<<1, 1024, 0, st>>>(tile_heads, T, num, seg, (int)n);
    LK_LAUNCHED();
    uniq_write_kernel<<<T, RS_THREADS, 0, st>>>(kin, vin, n, tile_heads,
                                                (unsigned long long*)d_unique, d_inverse, d_order, seg,
                                                d_sorted_rank);
    LK_LAUNCHED();
  }
  if (d_counts) {
    seg_counts_kernel<<<lk_grid(n, 256, 8), 256, 0, st>>>(seg, num, n, d_counts);
    LK_LAUNCHED();
  }
  return LK_OK;
}

extern "C" int lk_sort_unique(const uint64_t* d_keys, int64_t n, int key_bits, uint64_t* d_unique,
                              int32_t* d_inverse, int32_t* d_order, int32_t* d_seg,
                              int32_t* d_counts, int32_t* d_num, void* d_ws, int64_t ws_bytes,
                              lk_stream_t s) {
  return lk_sort_unique_ex(d_keys, n, key_bits, d_unique, d_inverse, d_order, d_seg, d_counts, d_num,
                           nullptr, d_ws, ws_bytes, s);
}
