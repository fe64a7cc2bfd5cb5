// Probe for cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (no torch, no libcuda link):
//   (1) semantics: destination layout under SWIZZLE_128B, out-of-bounds rows (-1 and N) -> zeros;
//   (2) throughput: one producer warp per CTA streaming [128 rows x 64 ch x 2 planes] stages of
//       gathered rows (75 % missing) through a ring of shared-memory stages, one CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_gather4_probe tma_gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

// ---- (1) semantics ----
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, const int* idx, int rows, int col, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) ((float*)smem)[i] = -777.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect(&bar, rows * 128);
    for (int g = 0; g < rows / 4; ++g)
      gather4(smem + g * 512, &tm, col, idx[4 * g], idx[4 * g + 1], idx[4 * g + 2], idx[4 * g + 3], &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) out[i] = ((float*)smem)[i];
}

// ---- (2) throughput ----
template <int STAGES>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const int* nbr /*[steps][128]*/,
                                                       int steps, int chunks /*128-byte column chunks per row, all planes*/,
                                                       unsigned long long* cycles, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const uint32_t stage_bytes = 128u * 128u * chunks;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int* my = nbr + (size_t)blockIdx.x * steps * 128;
  long long t0 = clock64();
  if (warp == 0) {           // producer: lane l owns rows 4l..4l+3 of the tile
    for (int j = 0; j < steps; ++j) {
      const int s = j % STAGES;
      if (j >= STAGES) mbar_wait(&empty[s], ((j / STAGES) - 1) & 1);
      int4 r = *(const int4*)(my + (size_t)j * 128 + 4 * lane);
      if (lane == 0) mbar_expect(&full[s], stage_bytes);
      __syncwarp();
      uint8_t* dst = smem + (size_t)s * stage_bytes + lane * 512;
      for (int c = 0; c < chunks; ++c)
        gather4(dst + (size_t)c * 128 * 128, &tm, c * 32, r.x, r.y, r.z, r.w, &full[s]);
    }
  } else {                   // consumer: touch one word per stage, release it
    float acc = 0.f;
    for (int j = 0; j < steps; ++j) {
      const int s = j % STAGES;
      mbar_wait(&full[s], (j / STAGES) & 1);
      acc += ((float*)(smem + (size_t)s * stage_bytes))[lane * 33];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 123.456f) sink[0] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

int main() {
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int N = 120000, C = 128;   // [N, 2 planes x 64 ch] fp32
  std::vector<float> hx((size_t)N * C);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)(i % 100003) * 0.25f + 1.0f;
  float* dx; CK(cudaMalloc(&dx, hx.size() * 4)); CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {32, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;

  // ---- semantics ----
  const int rows = 16;
  int hidx[rows] = {5, -1, 7, N, 119999, 0, -1, -1, 3, 3, 100, N + 5, -2, 42, 77, 1};
  int* didx; CK(cudaMalloc(&didx, sizeof(hidx))); CK(cudaMemcpy(didx, hidx, sizeof(hidx), cudaMemcpyHostToDevice));
  float* dout; CK(cudaMalloc(&dout, rows * 32 * 4));
  for (int col = 0; col <= 96; col += 32) {
    probe_kernel<<<1, 128, rows * 128 + 1024>>>(tm, didx, rows, col, dout);
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(rows * 32);
    CK(cudaMemcpy(ho.data(), dout, rows * 32 * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int rr = 0; rr < rows; ++rr)
      for (int ch = 0; ch < 8; ++ch)
        for (int e = 0; e < 4; ++e) {
          float got = ho[rr * 32 + ((ch ^ (rr & 7)) << 2) + e];
          int src = hidx[rr];
          float want = (src >= 0 && src < N) ? hx[(size_t)src * C + col + ch * 4 + e] : 0.f;
          if (got != want) { if (bad < 5) printf("  mismatch col %d row %d (src %d) ch %d: got %g want %g\n", col, rr, src, ch * 4 + e, got, want); ++bad; }
        }
    printf("semantics col=%d: %s (%d mismatches)\n", col, bad ? "FAIL" : "OK", bad);
  }

  // ---- throughput ----
  int sms = 148;
  const int steps = 189;
  std::vector<int> hn((size_t)sms * steps * 128);
  for (int pres = 25; pres <= 100; pres += 75) {
    srand(1);
    for (size_t i = 0; i < hn.size(); ++i) hn[i] = (rand() % 100 < pres) ? (int)((i * 7919u) % N) : -1;
    int* dn; CK(cudaMalloc(&dn, hn.size() * 4)); CK(cudaMemcpy(dn, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice));
    unsigned long long* dc; CK(cudaMalloc(&dc, sms * 8));
    float* sink; CK(cudaMalloc(&sink, 4));
    for (int chunks = 2; chunks <= 4; chunks += 2) {
      const int stages = chunks == 4 ? 3 : 6;
      size_t smem = (size_t)stages * 128 * 128 * chunks + 1024;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 3; ++rep) {
        if (chunks == 4) {
          CK(cudaFuncSetAttribute(stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          cudaEventRecord(e0);
          stream_kernel<3><<<sms, 64, smem>>>(tm, dn, steps, chunks, dc, sink);
          cudaEventRecord(e1);
        } else {
          CK(cudaFuncSetAttribute(stream_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          cudaEventRecord(e0);
          stream_kernel<6><<<sms, 64, smem>>>(tm, dn, steps, chunks, dc, sink);
          cudaEventRecord(e1);
        }
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> hc(sms);
        CK(cudaMemcpy(hc.data(), dc, sms * 8, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0; for (auto v : hc) mx = v > mx ? v : mx;
        printf("stream present=%d%% chunks=%d stages=%d: %.1f us total, %.0f cycles/step (max CTA), %.1f B/clk/SM smem fill, %.2f TB/s L2-side\n",
               pres, chunks, stages, ms * 1e3, (double)mx / steps, 128.0 * 128 * chunks / ((double)mx / steps),
               (double)sms * steps * 128 * 128 * chunks * pres / 100.0 / (ms * 1e-3) / 1e12);
      }
    }
    cudaFree(dn);
  }
  return 0;
}
