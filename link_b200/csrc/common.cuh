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

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A chain of short dependent kernels pays launch latency + CTA scheduling + prologue at every
// boundary (~2-4 us each on B200; the LinK block has ~10 boundaries on its critical chain).  Kernels
// launched through lk_launch_pdl carry cudaLaunchAttributeProgrammaticStreamSerialization: the grid may
// be scheduled while its predecessor in the stream is still running, and every thread executes
// lk_pdl_enter() -- griddepcontrol.wait: block until the predecessor grid has completed and its writes
// are visible -- before it touches global memory, then griddepcontrol.launch_dependents, which lets the
// NEXT kernel of the stream be scheduled early in turn.  Because every such kernel waits before it
// does anything else, completion stays transitive along the chain (kernel C may rely on the output of
// A through B).  Without the attribute (LINKB200_PDL=0, or an event / memset in between) both
// instructions are no-ops and the launch is an ordinary one.
__device__ __forceinline__ void lk_pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool lk_pdl_enabled();
bool lk_pdl_enabled_link();   // the four kernels of the pre-aggregation path (link.cu): measured SLOWER with PDL
                              // (graph replay of the path 27.3 -> 29.5 us: the early-resident CTAs of the next
                              // kernel take shared memory and issue slots from the tail of a bandwidth-bound
                              // one), so they keep ordinary launches unless LINKB200_PDL_LINK=1
template <typename... KArgs, typename... Args>
static inline cudaError_t lk_launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t st, Args... args);
template <typename... KArgs, typename... Args>
static inline cudaError_t lk_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args... args) {
  return lk_launch_pdl_if(lk_pdl_enabled(), kernel, grid, block, smem, st, args...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t lk_launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define LK_PDL_LAUNCH(kernel, grid, block, smem, st, ...) \
  LK_CUDA(lk_launch_pdl(kernel, dim3((unsigned)(grid)), dim3((unsigned)(block)), (size_t)(smem), st, __VA_ARGS__))
#define LK_PDL_LAUNCH_LINK(kernel, grid, block, smem, st, ...)                                          \
  LK_CUDA(lk_launch_pdl_if(lk_pdl_enabled_link(), kernel, dim3((unsigned)(grid)), dim3((unsigned)(block)), \
                           (size_t)(smem), st, __VA_ARGS__))
