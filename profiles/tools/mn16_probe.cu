// (fp16 variant, generated from mn_probe.cu) Probe of the UMMA shared-memory descriptor for MN-major fp16 operands staged by TMA (SWIZZLE_128B, box {32 floats, 32 rows}).
// One CTA computes D[128,128] = A^T B for A[32 k][128 m], B[32 k][128 n] with (LBO, SBO, K-advance) given on the command line
// loop, and reports which encoding reproduces the CPU result.  Build: nvcc -gencode arch=compute_100a,code=sm_100a mn_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { uint32_t lbo, sbo, kadv, amn, bmn, nk, kmaj_load, lt; };

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, Cfg c) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 32768);
  uint64_t* done = bar + 1;
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *holder;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(c.kmaj_load ? 32768u : 16384u) : "memory");
    if (c.kmaj_load) {      // control: K-major tiles [128 rows][32 k] in one box each
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem)), "l"(&tmA), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + 16384)), "l"(&tmB), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
    } else
    for (int g = 0; g < 2; ++g) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + g * 4096)), "l"(&tmA), "r"(64 * g), "r"(0), "r"(smem_u32(bar)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + 16384 + g * 4096)), "l"(&tmB), "r"(64 * g), "r"(0), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (c.amn << 15) | (c.bmn << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    auto desc = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)(c.lbo >> 4) << 16;
      d |= (uint64_t)(c.sbo >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(c.lt ? c.lt : 2) << 61;
      return d;
    };
    for (uint32_t k = 0; k < c.nk; ++k) {
      const uint64_t ad = desc(smem_u32(smem) + k * c.kadv), bd = desc(smem_u32(smem + 16384) + k * c.kadv);
      const uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(done)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int K = 32, M = 128, N = 128;
  std::vector<__half> hA(K * M), hB(K * N);
  std::vector<float> fA(K * M), fB(K * N), hD(M * N + 128), ref(M * N, 0.f);
  srand(1);
  for (int i = 0; i < K * M; ++i) { fA[i] = (float)(rand() % 7 - 3); hA[i] = __float2half(fA[i]); }
  for (int i = 0; i < K * N; ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = __float2half(fB[i]); }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaDriverEntryPointQueryResult q;
  void* fp = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  cuuint64_t dims[2] = {128, 32};
  cuuint64_t strides[1] = {128 * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000);
  const CUtensorMapSwizzle sws[2] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
  for (int swi = 0; swi < 2; ++swi) {
    CUtensorMap tmA, tmB;
    CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sws[swi],
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sws[swi],
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle mode %d encode %d %d\n", swi, (int)r1, (int)r2);
    const uint32_t lbos[] = {4096, 1024, 2048, 128};
    const uint32_t sbos[] = {1024, 512, 2048, 4096};
    const uint32_t kadvs[] = {2048, 1024, 512};
    for (int nk = 1; nk <= 2; ++nk) {
      for (auto& x : ref) x = 0.f;
      for (int k = 0; k < 16 * nk; ++k) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) ref[m * N + n] += fA[k * M + m] * fB[k * N + n];
      for (uint32_t lt = 1; lt <= 2; ++lt) for (uint32_t lbo : lbos) for (uint32_t sbo : sbos) for (uint32_t kadv : kadvs) {
        if (nk == 1 && kadv != 2048) continue;
        Cfg c = {lbo, sbo, kadv, 1, 1, (uint32_t)nk, 0, lt};
        cudaMemset(dD, 0xff, hD.size() * 4);
        probe<<<1, 128, 36000>>>(tmA, tmB, dD, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, zeros = 0;
        for (int i = 0; i < M * N; ++i) { bad += hD[i] != ref[i]; zeros += hD[i] == 0.f; }
        if (bad == 0 || (lbo == 4096 && sbo == 1024))
          printf("sw=%d lt=%u nk=%d lbo=%5u sbo=%5u kadv=%5u: mismatches %5d zeros %5d  D[0][0..3]= %g %g %g %g  ref %g %g %g %g\n", swi, lt, nk, lbo, sbo,
                 kadv, bad, zeros, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
      }
    }
  }
  return 0;
}
