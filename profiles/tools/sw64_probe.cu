// Probe of a small K-major fp16 UMMA whose operands are written to shared memory by ordinary threads (no TMA) in the SWIZZLE_64B
// layout: D[128 a, 16 b] = sum_{k<32} A[a][k] * B[b][k]  (the in-kernel query projection of decoder_persist.cu).
// Rows of 64 bytes (32 fp16), 8-row groups of 512 B, 16-byte chunk c of row r stored at chunk c ^ ((r >> 1) & 3).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a sw64_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int swz64(int r, int k) { return r * 32 + ((((k >> 3) ^ (r >> 1)) & 3) << 3) + (k & 7); }   // element index

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, uint32_t sbo, uint32_t lt) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __half* As = (__half*)smem;                 // [128][32]
  __half* Bs = (__half*)(smem + 8192);        // [16][32]
  uint64_t* done = (uint64_t*)(smem + 8192 + 1024);
  uint32_t* holder = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128 * 32; i += 128) As[swz64(i >> 5, i & 31)] = A[i];
  for (int i = threadIdx.x; i < 16 * 32; i += 128) Bs[swz64(i >> 5, i & 31)] = B[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *holder;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    auto desc = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(sbo >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)lt << 61;
      return d;
    };
    for (uint32_t k = 0; k < 2; ++k) {
      const uint64_t ad = desc(smem_u32(As)) + 2 * k, bd = desc(smem_u32(Bs)) + 2 * k;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(k) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(done)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  std::vector<__half> hA(128 * 32), hB(16 * 32);
  std::vector<float> fA(128 * 32), fB(16 * 32), hD(128 * 16), ref(128 * 16, 0.f);
  srand(2);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 7 - 3); hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = __float2half(fB[i]); }
  for (int a = 0; a < 128; ++a) for (int b = 0; b < 16; ++b) for (int k = 0; k < 32; ++k) ref[a * 16 + b] += fA[a * 32 + k] * fB[b * 32 + k];
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const uint32_t sbos[] = {512, 1024, 256, 64};
  const uint32_t lts[] = {4, 2, 6, 0};
  for (uint32_t lt : lts) for (uint32_t sbo : sbos) {
    cudaMemset(dD, 0xff, hD.size() * 4);
    probe<<<1, 128, 16384>>>(dA, dB, dD, sbo, lt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lt=%u sbo=%u: CUDA error %s\n", lt, sbo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 128 * 16; ++i) bad += hD[i] != ref[i];
    printf("layout_type=%u sbo=%4u: mismatches %4d   D[0][0..3]= %g %g %g %g  ref %g %g %g %g\n", lt, sbo, bad, hD[0], hD[1], hD[2], hD[3],
           ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
