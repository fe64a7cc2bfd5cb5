// Coordinate hashing, the open-addressing hash table and the index histogram.
// Replaces backend/hash/hash_cuda.cu, backend/hashmap/hashmap_cuda.cu(+.cuh),
// backend/others/query_cuda.cu and backend/others/count_cuda.cu of the reference.
#include "common.cuh"

// ---------------------------------------------------------------- hashing
__global__ void __launch_bounds__(256) hash_kernel(const int4* __restrict__ coords, int64_t n,
                                                   int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
    out[i] = lk_fnv4(c.x, c.y, c.z, c.w);
  }
}

// one thread per voxel, all K offsets in registers' reach: the coordinate row is read once
// and each of the K output rows is written coalesced (layout [K, N]).
__global__ void __launch_bounds__(256) kernel_hash_kernel(const int4* __restrict__ coords,
                                                          int64_t n,
                                                          const int* __restrict__ offsets, int K,
                                                          int64_t* __restrict__ out) {
  extern __shared__ int s_off[];
  for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
    for (int k = 0; k < K; ++k)
      out[(int64_t)k * n + i] =
          lk_fnv4(c.x + s_off[3 * k], c.y + s_off[3 * k + 1], c.z + s_off[3 * k + 2], c.w);
  }
}

extern "C" int lk_hash(const int32_t* d_coords, int64_t n, int64_t* d_out, lk_stream_t s) {
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_coords && d_out && n > 0, "lk_hash: null pointer or negative n");
  hash_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>((const int4*)d_coords, n, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_kernel_hash(const int32_t* d_coords, int64_t n, const int32_t* d_offsets, int k,
                              int64_t* d_out, lk_stream_t s) {
  if (n == 0 || k == 0) return LK_OK;
  LK_REQUIRE(d_coords && d_out && d_offsets && n > 0 && k > 0 && k <= 4096,
             "lk_kernel_hash: bad arguments");
  kernel_hash_kernel<<<lk_grid(n, 256, 8), 256, k * 3 * sizeof(int), (cudaStream_t)s>>>(
      (const int4*)d_coords, n, d_offsets, k, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- hash table
// 16-byte slots so that one 128-bit load fetches key and value of a probe.
struct __align__(16) Slot {
  unsigned long long key;
  unsigned int val;  // row index; 0xFFFFFFFF after the 0xFF memset, so atomicMin works unsigned
  int pad;
};
#define LK_EMPTY 0xFFFFFFFFFFFFFFFFULL

__device__ __forceinline__ uint64_t slot_of(uint64_t key, uint64_t mask) {
  key ^= key >> 33; key *= 0xff51afd7ed558ccdULL;  // murmur3 finaliser
  key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ULL;
  key ^= key >> 33;
  return key & mask;
}

// One 128-bit compare-and-swap claims an empty slot AND stores the row index (sm_90+: atom.cas.b128):
// a slot is either all ones or a complete {key, row} record, so an insert costs ONE L2 atomic round
// trip instead of a 64-bit CAS on the key followed by a dependent atomicMin on the value (the insert
// kernel is pure atomic latency: 80 % long-scoreboard stalls, 18 us for 119k voxels before).
// Returns the key found in the slot before the operation (LK_EMPTY: the slot was claimed).
__device__ __forceinline__ unsigned long long slot_claim(Slot* slot, unsigned long long key, unsigned int row) {
  unsigned long long old_lo, old_hi;    // old_hi (row | pad of the previous record) is not needed
  const unsigned long long ones = LK_EMPTY, new_hi = (unsigned long long)row;    // pad = 0
  asm volatile(
      "{\n\t.reg .b128 cmp, swp, old;\n\t"
      "mov.b128 cmp, {%3, %3};\n\t"
      "mov.b128 swp, {%4, %5};\n\t"
      "atom.global.cas.b128 old, [%2], cmp, swp;\n\t"
      "mov.b128 {%0, %1}, old;\n\t}"
      : "=l"(old_lo), "=l"(old_hi) : "l"(slot), "l"(ones), "l"(key), "l"(new_hi) : "memory");
  (void)old_hi;
  return old_lo;
}

__global__ void __launch_bounds__(256) table_insert_kernel(const int64_t* __restrict__ keys,
                                                           int64_t n, Slot* table, uint64_t mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = (unsigned long long)keys[i];
    uint64_t s = slot_of(key, mask);
    while (true) {
      const unsigned long long prev = slot_claim(&table[s], key, (unsigned int)i);
      if (prev == LK_EMPTY) break;
      if (prev == key) {
        atomicMin(&table[s].val, (unsigned int)i);  // duplicates: lowest row wins
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

__device__ __forceinline__ int table_find(const Slot* __restrict__ table, uint64_t mask,
                                          unsigned long long key) {
  uint64_t s = slot_of(key, mask);
  while (true) {
    int4 raw = __ldg((const int4*)&table[s]);
    unsigned long long k = ((unsigned long long)(unsigned)raw.y << 32) | (unsigned)raw.x;
    if (k == key) return raw.z;
    if (k == LK_EMPTY) return -1;
    s = (s + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) table_query_kernel(const int64_t* __restrict__ q, int64_t nq,
                                                          const Slot* __restrict__ table,
                                                          uint64_t mask, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = table_find(table, mask, (unsigned long long)q[i]);
}

// same table, keys hashed on the fly from the coordinates (saves the hash array round trip and a launch)
__global__ void __launch_bounds__(256) table_insert_coords_kernel(const int4* __restrict__ coords,
                                                                  int64_t n, Slot* table, uint64_t mask) {
  lk_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = coords[i];
    unsigned long long key = (unsigned long long)lk_fnv4(c.x, c.y, c.z, c.w);
    uint64_t s = slot_of(key, mask);
    while (true) {
      const unsigned long long prev = slot_claim(&table[s], key, (unsigned int)i);
      if (prev == LK_EMPTY) break;
      if (prev == key) {
        atomicMin(&table[s].val, (unsigned int)i);
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

extern "C" int lk_table_build_coords(const int32_t* d_coords, int64_t n, void* d_table, int64_t capacity,
                                     lk_stream_t s) {
  LK_REQUIRE(d_table && capacity >= 2 * n && (capacity & (capacity - 1)) == 0 && n >= 0,
             "lk_table_build_coords: capacity must be a power of two >= 2n");
  LK_CUDA(cudaMemsetAsync(d_table, 0xFF, (size_t)capacity * sizeof(Slot), (cudaStream_t)s));
  lk_count_launch();
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_coords, "lk_table_build_coords: null coords");
  LK_PDL_LAUNCH(table_insert_coords_kernel, lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s, (const int4*)d_coords, n,
                (Slot*)d_table, (uint64_t)capacity - 1);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int64_t lk_table_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

extern "C" int lk_table_build(const int64_t* d_keys, int64_t n, void* d_table, int64_t capacity,
                              lk_stream_t s) {
  LK_REQUIRE(d_table && capacity >= 2 * n && (capacity & (capacity - 1)) == 0 && n >= 0,
             "lk_table_build: capacity must be a power of two >= 2n");
  // key = all ones (empty marker), value = 0xFFFFFFFF (unsigned max, for atomicMin)
  LK_CUDA(cudaMemsetAsync(d_table, 0xFF, (size_t)capacity * sizeof(Slot), (cudaStream_t)s));
  lk_count_launch();
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_keys, "lk_table_build: null keys");
  table_insert_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>(d_keys, n, (Slot*)d_table,
                                                                      (uint64_t)capacity - 1);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_table_query(const int64_t* d_queries, int64_t nq, const void* d_table,
                              int64_t capacity, int64_t* d_out, lk_stream_t s) {
  if (nq == 0) return LK_OK;
  LK_REQUIRE(d_queries && d_table && d_out && nq > 0, "lk_table_query: bad arguments");
  table_query_kernel<<<lk_grid(nq, 256, 8), 256, 0, (cudaStream_t)s>>>(
      d_queries, nq, (const Slot*)d_table, (uint64_t)capacity - 1, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- kernel-map query
// One thread per output voxel; the K neighbour hashes are formed in registers and probed
// directly, so the [K, N] int64 hash matrix of the reference (conv.py:113) never exists.
__global__ void __launch_bounds__(256) kmap_query_kernel(const int4* __restrict__ coords,
                                                         int64_t n, const int* __restrict__ offsets,
                                                         int K, const Slot* __restrict__ table,
                                                         uint64_t mask, int* __restrict__ nbr) {
  extern __shared__ int s_off[];
  for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
#pragma unroll 3
    for (int k = 0; k < K; ++k) {
      int64_t h = lk_fnv4(c.x + s_off[3 * k], c.y + s_off[3 * k + 1], c.z + s_off[3 * k + 2], c.w);
      nbr[(int64_t)k * n + i] = table_find(table, mask, (unsigned long long)h);
    }
  }
}

extern "C" int lk_kmap_query(const int32_t* d_out_coords, int64_t n_out, const int32_t* d_offsets,
                             int k, const void* d_table, int64_t capacity, int32_t* d_nbr,
                             lk_stream_t s) {
  if (n_out == 0 || k == 0) return LK_OK;
  LK_REQUIRE(d_out_coords && d_offsets && d_table && d_nbr && k > 0 && k <= 4096,
             "lk_kmap_query: bad arguments");
  kmap_query_kernel<<<lk_grid(n_out, 256, 8), 256, k * 3 * sizeof(int), (cudaStream_t)s>>>(
      (const int4*)d_out_coords, n_out, d_offsets, k, (const Slot*)d_table, (uint64_t)capacity - 1,
      d_nbr);
  LK_LAUNCHED();
  return LK_OK;
}

// Submanifold maps are symmetric: (i --k--> j)  <=>  (j --(K-1-k)--> i).  One thread per
// (voxel, k < K/2) pair: half the probes of the generic kernel and K/2-fold more parallelism.
__global__ void __launch_bounds__(256) kmap_query_subm_kernel(const int4* __restrict__ coords,
                                                              int64_t n, const int* __restrict__ offsets,
                                                              int K, const Slot* __restrict__ table,
                                                              uint64_t mask, unsigned* nbr) {
  lk_pdl_enter();
  // blockIdx.y = offset k <= K/2, blockIdx.x strides over the voxels: no (voxel, offset) index to take
  // apart -- the 64-bit division of the flat form was half of this kernel's ~220 instructions per probe
  // (issue 56 % busy in the ncu capture of the flat kernel).
  const int half = K / 2;
  const int k = blockIdx.y;
  if (k == half) {                                    // centre offset: identity
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      nbr[(int64_t)half * n + i] = (unsigned)i;
    return;
  }
  const int ox = __ldg(offsets + 3 * k), oy = __ldg(offsets + 3 * k + 1), oz = __ldg(offsets + 3 * k + 2);
  unsigned* const fwd = nbr + (int64_t)k * n;
  unsigned* const bwd = nbr + (int64_t)(K - 1 - k) * n;
  // (A variant with four probes in flight per thread measured SLOWER, 64 vs 55 us for the whole map build:
  // the extra registers halve the resident warps, and the probes are bound by the random 32-byte sector
  // reads of the 4 MB table, not by issue.)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = coords[i];
    const int64_t h = lk_fnv4(c.x + ox, c.y + oy, c.z + oz, c.w);
    const int j = table_find(table, mask, (unsigned long long)h);
    if (j >= 0) {
      fwd[i] = (unsigned)j;
      atomicMin(&bwd[j], (unsigned)i);                // 0xFFFFFFFF == -1 prefill
    }
  }
}

extern "C" int lk_kmap_query_subm(const int32_t* d_coords, int64_t n, const int32_t* d_offsets, int k,
                                  const void* d_table, int64_t capacity, int32_t* d_nbr,
                                  lk_stream_t s) {
  return lk_kmap_query_subm_ev(d_coords, n, d_offsets, k, d_table, capacity, d_nbr, nullptr, s);
}

// table_ready: cudaEvent_t recorded after the table build on ANOTHER stream, or NULL (table built on s).
// The map is cleared first, so that part overlaps a table that is still being built.
extern "C" int lk_kmap_query_subm_ev(const int32_t* d_coords, int64_t n, const int32_t* d_offsets, int k,
                                     const void* d_table, int64_t capacity, int32_t* d_nbr,
                                     void* table_ready, lk_stream_t s) {
  if (n == 0 || k == 0) return LK_OK;
  LK_REQUIRE(d_coords && d_offsets && d_table && d_nbr && k > 0 && (k & 1) == 1,
             "lk_kmap_query_subm: needs an odd kernel volume");
  cudaStream_t st = (cudaStream_t)s;
  LK_CUDA(cudaMemsetAsync(d_nbr, 0xFF, (size_t)k * n * sizeof(int), st));
  lk_count_launch();
  if (table_ready) LK_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)table_ready, 0));
  {
    // (k/2 + 1) offset rows x enough CTAs per row to fill the machine once (8 resident CTAs per SM)
    const int rows = k / 2 + 1;
    int per_row = (LK_SM_COUNT * 8 + rows - 1) / rows;
    const int64_t need = (n + 255) / 256;
    if (per_row > need) per_row = (int)need;
    if (per_row < 1) per_row = 1;
    const dim3 grid((unsigned)per_row, (unsigned)rows);
    LK_CUDA(lk_launch_pdl(kmap_query_subm_kernel, grid, dim3(256), 0, st, (const int4*)d_coords, n, d_offsets, k,
                          (const Slot*)d_table, (uint64_t)capacity - 1, (unsigned*)d_nbr));
  }
  LK_LAUNCHED();
  return LK_OK;
}

__global__ void __launch_bounds__(256) kmap_invert_kernel(const int* __restrict__ nbr,
                                                          int64_t n_out, int K, int64_t n_in,
                                                          int* __restrict__ inv) {
  for (int k = blockIdx.y; k < K; k += gridDim.y) {           // offset rows on grid.y: no 64-bit division per entry
    const int* row = nbr + (int64_t)k * n_out;
    int* dst = inv + (int64_t)k * n_in;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_out; o += (int64_t)gridDim.x * blockDim.x) {
      const int i = row[o];
      if (i >= 0) dst[i] = (int)o;
    }
  }
}

extern "C" int lk_kmap_invert(const int32_t* d_nbr, int64_t n_out, int k, int64_t n_in,
                              int32_t* d_inv, lk_stream_t s) {
  LK_REQUIRE(d_inv && k > 0 && n_in >= 0, "lk_kmap_invert: bad arguments");
  if (n_in > 0) {
    LK_CUDA(cudaMemsetAsync(d_inv, 0xFF, (size_t)k * n_in * sizeof(int), (cudaStream_t)s));
    lk_count_launch();
  }
  if (n_out == 0 || n_in == 0) return LK_OK;
  {
    int per_row = (LK_SM_COUNT * 8 + k - 1) / k;
    const int64_t need = (n_out + 255) / 256;
    if (per_row > need) per_row = (int)need;
    if (per_row < 1) per_row = 1;
    kmap_invert_kernel<<<dim3((unsigned)per_row, (unsigned)(k < 65535 ? k : 65535)), 256, 0, (cudaStream_t)s>>>(
        d_nbr, n_out, k, n_in, d_inv);
  }
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- coarse-parent lookup
// upsample_voxel (segmentation/core/models/utils.py:327-340) hashes floor(coord / stride) of the
// coarse and of the fine level and queries one against the other; the floor-division is folded
// into the hash / probe kernels so no intermediate coordinate or hash tensor is materialised.
__global__ void __launch_bounds__(256) hash_div_kernel(const int4* __restrict__ coords, int64_t n,
                                                       int div, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
    out[i] = lk_fnv4(lk_floordiv(c.x, div), lk_floordiv(c.y, div), lk_floordiv(c.z, div), c.w);
  }
}

__global__ void __launch_bounds__(256) table_query_div_kernel(const int4* __restrict__ coords,
                                                              int64_t n, int div,
                                                              const Slot* __restrict__ table,
                                                              uint64_t mask, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
    int64_t h = lk_fnv4(lk_floordiv(c.x, div), lk_floordiv(c.y, div), lk_floordiv(c.z, div), c.w);
    out[i] = table_find(table, mask, (unsigned long long)h);
  }
}

extern "C" int lk_hash_div(const int32_t* d_coords, int64_t n, int div, int64_t* d_out,
                           lk_stream_t s) {
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_coords && d_out && n > 0 && div >= 1, "lk_hash_div: bad arguments");
  hash_div_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>((const int4*)d_coords, n, div, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

extern "C" int lk_table_query_div(const int32_t* d_coords, int64_t n, int div, const void* d_table,
                                  int64_t capacity, int64_t* d_out, lk_stream_t s) {
  if (n == 0) return LK_OK;
  LK_REQUIRE(d_coords && d_table && d_out && n > 0 && div >= 1, "lk_table_query_div: bad arguments");
  table_query_div_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>(
      (const int4*)d_coords, n, div, (const Slot*)d_table, (uint64_t)capacity - 1, d_out);
  LK_LAUNCHED();
  return LK_OK;
}

// ---------------------------------------------------------------- count
__global__ void __launch_bounds__(256) count_kernel(const int* __restrict__ idx, int64_t n,
                                                    int* out, int64_t num) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int v = idx[i];
    if (v >= 0 && v < num) atomicAdd(&out[v], 1);
  }
}

extern "C" int lk_count(const int32_t* d_idx, int64_t n, int32_t* d_out, int64_t num,
                        lk_stream_t s) {
  LK_REQUIRE(num >= 0 && n >= 0, "lk_count: negative size");
  if (num > 0) {
    LK_REQUIRE(d_out, "lk_count: null output");
    LK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)num * sizeof(int), (cudaStream_t)s));
    lk_count_launch();
  }
  if (n == 0 || num == 0) return LK_OK;
  count_kernel<<<lk_grid(n, 256, 8), 256, 0, (cudaStream_t)s>>>(d_idx, n, d_out, num);
  LK_LAUNCHED();
  return LK_OK;
}
