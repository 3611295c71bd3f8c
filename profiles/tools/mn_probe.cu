// Probe of the UMMA shared-memory descriptor for MN-major tf32 operands staged by TMA (SWIZZLE_128B, box {32 floats, 32 rows}).
// One CTA computes D[128,128] = A^T B for A[32 k][128 m], B[32 k][128 n] with (LBO, SBO, K-advance) given on the command line
// loop, and reports which encoding reproduces the CPU result.  Build: nvcc -gencode arch=compute_100a,code=sm_100a mn_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
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
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(32768u) : "memory");
    if (c.kmaj_load) {      // control: K-major tiles [128 rows][32 k] in one box each
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem)), "l"(&tmA), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + 16384)), "l"(&tmB), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
    } else
    for (int g = 0; g < 4; ++g) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + g * 4096)), "l"(&tmA), "r"(32 * g), "r"(0), "r"(smem_u32(bar)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + 16384 + g * 4096)), "l"(&tmB), "r"(32 * g), "r"(0), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int i = 0; i < 64; ++i) D[128 * 128 + i] = ((float*)smem)[i];              // what TMA delivered (A tile head)
    for (int i = 0; i < 64; ++i) D[128 * 128 + 64 + i] = ((float*)(smem + 16384))[i];
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (c.amn << 15) | (c.bmn << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
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
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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
  std::vector<float> hA(K * M), hB(K * N), hD(M * N + 128), ref(M * N, 0.f);
  srand(1);
  for (auto& x : hA) x = (float)(rand() % 7 - 3);
  for (auto& x : hB) x = (float)(rand() % 5 - 2);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  cudaDriverEntryPointQueryResult q;
  void* fp = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  CUtensorMap tmA, tmB;
  cuuint64_t dims[2] = {128, 32};
  cuuint64_t strides[1] = {128 * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", (int)r1, (int)r2);
  // ---- control: the same product with K-major operands At[m][k], Bt[n][k] (what gemm_tc_kernel does)
  {
    std::vector<float> hAt(M * K), hBt(N * K);
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) hAt[m * K + k] = hA[k * M + m];
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) hBt[n * K + k] = hB[k * N + n];
    float *dAt, *dBt;
    cudaMalloc(&dAt, hAt.size() * 4); cudaMalloc(&dBt, hBt.size() * 4);
    cudaMemcpy(dAt, hAt.data(), hAt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBt, hBt.data(), hBt.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    cuuint64_t d2[2] = {32, 128};
    cuuint64_t s2[1] = {32 * 4};
    cuuint32_t b2[2] = {32, 128};
    enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dAt, d2, s2, b2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dBt, d2, s2, b2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000);
    float* dDc; cudaMalloc(&dDc, hD.size() * 4);
    Cfg c = {16, 1024, 32, 0, 0, 4, 1, 0};
    probe<<<1, 128, 36000>>>(tA, tB, dDc, c);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD.data(), dDc, hD.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<float> r(M * N, 0.f);
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) r[m * N + n] += hA[k * M + m] * hB[k * N + n];
    int bad = 0; for (int i = 0; i < M * N; ++i) bad += hD[i] != r[i];
    printf("CONTROL K-major: err=%s mismatches %d  D[0][0..3]= %g %g %g %g ref %g %g %g %g\n", cudaGetErrorString(e), bad, hD[0], hD[1], hD[2], hD[3], r[0], r[1], r[2], r[3]);
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000);
  const uint32_t lbos[] = {4096, 1024, 128, 16, 8192, 0};
  const uint32_t sbos[] = {1024, 4096, 128, 256, 0};
  const uint32_t kadvs[] = {1024, 128, 32, 4096, 256};
  // ---- what TMA delivered for the MN-major boxes + single-operand toggles
  for (int t = 0; t < 4; ++t) {
    Cfg c = {4096, 1024, 1024, (uint32_t)(t & 1), (uint32_t)(t >> 1), 1, 0, 0};
    cudaMemset(dD, 0xff, hD.size() * 4);
    probe<<<1, 128, 36000>>>(tmA, tmB, dD, c);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int zeros = 0; for (int i = 0; i < M * N; ++i) zeros += hD[i] == 0.f;
    printf("toggle amn=%d bmn=%d: err=%s zeros %d  D[0][0..7]= %g %g %g %g %g %g %g %g\n", t & 1, t >> 1, cudaGetErrorString(e), zeros,
           hD[0], hD[1], hD[2], hD[3], hD[4], hD[5], hD[6], hD[7]);
    if (t == 0) {
      printf("smem A head:"); for (int i = 0; i < 64; ++i) printf(" %g", hD[M * N + i]); printf("\n");
      printf("glob A head:"); for (int i = 0; i < 64; ++i) printf(" %g", hA[i]); printf("\n");
      printf("smem B head:"); for (int i = 0; i < 64; ++i) printf(" %g", hD[M * N + 64 + i]); printf("\n");
    }
  }
  {
    CUtensorMap tA32, tB32;
    CUresult q1 = enc(&tA32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult q2 = enc(&tB32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode ATOM_32B %d %d\n", (int)q1, (int)q2);
    const uint32_t lbos2[] = {4096, 512, 1024, 128};
    const uint32_t sbos2[] = {512, 1024, 256, 128, 4096};
    const uint32_t kadvs2[] = {1024, 512, 256};
    for (int nk = 1; nk <= 4; nk += 3) {
      for (auto& x : ref) x = 0.f;
      for (int k = 0; k < 8 * nk; ++k) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) ref[m * N + n] += hA[k * M + m] * hB[k * N + n];
      for (uint32_t lbo : lbos2) for (uint32_t sbo : sbos2) for (uint32_t kadv : kadvs2) {
        if (nk == 1 && kadv != 1024) continue;
        Cfg c = {lbo, sbo, kadv, 1, 1, (uint32_t)nk, 0, 1};
        cudaMemset(dD, 0xff, hD.size() * 4);
        probe<<<1, 128, 36000>>>(tA32, tB32, dD, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("BASE32B nk=%d lbo=%u sbo=%u: CUDA error %s\n", nk, lbo, sbo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, zeros = 0, bad32 = 0;
        for (int i = 0; i < M * N; ++i) { bad += hD[i] != ref[i]; zeros += hD[i] == 0.f; }
        for (int m = 0; m < 32; ++m) for (int n = 0; n < 32; ++n) bad32 += hD[m * N + n] != ref[m * N + n];
        printf("BASE32B nk=%d lbo=%5u sbo=%5u kadv=%5u: mismatches %5d (first 32x32: %4d) zeros %5d  D[0][0..3]= %g %g %g %g  ref %g %g %g %g\n",
               nk, lbo, sbo, kadv, bad, bad32, zeros, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
      }
    }
    return 0;
  }
  for (int nk = 1; nk <= 4; nk += 3) {
    for (auto& x : ref) x = 0.f;
    for (int k = 0; k < 8 * nk; ++k)
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) ref[m * N + n] += hA[k * M + m] * hB[k * N + n];
    for (uint32_t lbo : lbos) for (uint32_t sbo : sbos) for (uint32_t kadv : kadvs) {
      if (nk == 1 && kadv != 1024) continue;
      Cfg c = {lbo, sbo, kadv, 1, 1, (uint32_t)nk, 0, 0};
      cudaMemset(dD, 0xff, hD.size() * 4);
      probe<<<1, 128, 36000>>>(tmA, tmB, dD, c);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("nk=%d lbo=%u sbo=%u kadv=%u: CUDA error %s\n", nk, lbo, sbo, kadv, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, zeros = 0, bad32 = 0;
      for (int i = 0; i < M * N; ++i) { bad += hD[i] != ref[i]; zeros += hD[i] == 0.f; }
      for (int m = 0; m < 32; ++m) for (int n = 0; n < 32; ++n) bad32 += hD[m * N + n] != ref[m * N + n];
      printf("nk=%d lbo=%5u sbo=%5u kadv=%5u: mismatches %5d (first 32x32 block: %4d) zeros %5d   D[0][0..3]= %g %g %g %g  ref %g %g %g %g\n",
             nk, lbo, sbo, kadv, bad, bad32, zeros, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    }
  }
  return 0;
}
