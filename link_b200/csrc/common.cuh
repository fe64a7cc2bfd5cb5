// Shared helpers for liblinkb200 (sm_100a).  No torch headers: pure CUDA runtime + C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/linkb200.h"

#define LK_SM_COUNT 148  // B200: 2 dies x 74 SMs; persistent grids are sized from this

void lk_set_error(const char* fmt, ...);
void lk_count_launch(int n = 1);

#define LK_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      lk_set_error(__VA_ARGS__);         \
      return LK_EINVAL;                  \
    }                                    \
  } while (0)

#define LK_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      lk_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,                  \
                   cudaGetErrorString(e__));                                   \
      return LK_ECUDA;                                                         \
    }                                                                          \
  } while (0)

// after a <<<>>> launch
#define LK_LAUNCHED()              \
  do {                             \
    lk_count_launch();             \
    LK_CUDA(cudaGetLastError());   \
  } while (0)

static inline int lk_blocks(int64_t work, int per_block) {
  int64_t b = (work + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : b);
}

// grid for grid-stride kernels: enough CTAs to cover `work`, capped at a multiple of the
// SM count so large inputs run as a few full waves of resident CTAs.
static inline int lk_grid(int64_t work, int per_block, int ctas_per_sm) {
  int64_t b = (work + per_block - 1) / per_block;
  int64_t cap = (int64_t)LK_SM_COUNT * ctas_per_sm;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// 64-bit FNV-1a over four int32 fields, folded to 60 bits.
// Bit-exact with the reference hash (backend/hash/hash_cuda.cu:14-21).
__host__ __device__ __forceinline__ int64_t lk_fnv4(int x, int y, int z, int b) {
  unsigned long long h = 14695981039346656037ULL;
  h ^= (unsigned int)x; h *= 1099511628211ULL;
  h ^= (unsigned int)y; h *= 1099511628211ULL;
  h ^= (unsigned int)z; h *= 1099511628211ULL;
  h ^= (unsigned int)b; h *= 1099511628211ULL;
  h = (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFFULL);
  return (int64_t)h;
}

__device__ __forceinline__ int lk_floordiv(int a, int d) {  // d > 0
  int q = a / d;
  return (a % d != 0 && a < 0) ? q - 1 : q;
}

// streaming 128-bit loads/stores that do not pollute L1
__device__ __forceinline__ float4 lk_ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void lk_stg_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// vector reduction into global memory (sm_90+): one L2 atomic transaction for 4 floats
__device__ __forceinline__ void lk_red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
