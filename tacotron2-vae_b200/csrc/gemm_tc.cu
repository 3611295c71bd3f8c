// tcgen05 GEMM for sm_100a:  D[M,N] (fp32) = alpha * A[M,K] * B[N,K]^T (+bias[n]) , both operands K-major.
//
//   * operands staged global->shared by TMA (cp.async.bulk.tensor.2d, 128B swizzle) through a STAGES-deep
//     mbarrier ring; one elected thread issues tcgen05.mma (cta_group::1, M=128, N=BN, 32 bytes of K per
//     instruction); the fp32 accumulator lives in TMEM and is read back with tcgen05.ld by 4 epilogue warps.
//   * element type: bf16 (kind::f16) or fp32-storage/tf32-math (kind::tf32) -- the kernel is byte-generic.
//   * "taps": the K loop can walk several row-shifted views of A (k-iteration (tap, chunk) reads A rows
//     m0+tap*rowshift..), which turns a zero-padded channels-last Conv1d (Postnet / Encoder, reference
//     model.py:105-177) and the STFT framing (stft.py:91-95) into plain GEMMs without an im2col buffer.
//   * split-K over blockIdx.z: partial tiles are either stored to D + z*split_stride or atomically added.
#include "t2v_common.cuh"
#include "gemm_tc.h"
#include <stdlib.h>

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// same with an L2 eviction-priority hint (weights that are re-read every decoder step: evict-last)
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <bool kTF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 B apart (SBO); LBO unused (=1)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version for sm_100
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int kBM = 128;

template <int BN, int ESIZE, int STAGES, int BM = 128>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmTcParams p) {
  constexpr int BK = 128 / ESIZE;           // elements per 128-byte swizzle row
  // BM = 256: two 128-row accumulators (TMEM columns [0, BN) and [BN, 2 BN)) share every B tile -- the large GEMMs are bound by the
  // L2 -> shared-memory operand stream, and a 256 x 256 tile moves a third fewer bytes per FLOP than 128 x 256
  constexpr int MB = (BM == 256) ? 2 : 1;
  constexpr int MMA_M = (BM == 64) ? 64 : 128;
  constexpr int A_BYTES = BM * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int it0 = blockIdx.z * p.iters_per_split;
  const int n_it = p.iters_per_split;
  t2v_pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"((uint32_t)(BN * MB))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // split product: iteration g belongs to term g / iters_per_term (0: A,B  1: A2,B  2: A,B2)
      auto load_b = [&](int it) {
        const int s = it % STAGES;
        int g = it0 + it;
        const int term = p.iters_per_term > 0 ? g / p.iters_per_term : 0;
        g -= term * p.iters_per_term;
        const int tap = g / p.chunks_per_tap;
        const int chunk = g - tap * p.chunks_per_tap;
        uint8_t* sa = smem + s * STAGE_BYTES;
        const CUtensorMap* mb = (term == 2) ? &tmB2 : &tmB;
        if (p.b_evict_last)
          tma_load_2d_hint(sa + A_BYTES, mb, p.b_k0 + tap * p.b_tap_stride + chunk * BK, p.b_row0 + n0, &full[s],
                           0x14F0000000000000ull);           // L2 evict_last
        else
          tma_load_2d(sa + A_BYTES, mb, p.b_k0 + tap * p.b_tap_stride + chunk * BK, p.b_row0 + n0, &full[s]);
      };
      auto load_a = [&](int it) {
        const int s = it % STAGES;
        int g = it0 + it;
        const int term = p.iters_per_term > 0 ? g / p.iters_per_term : 0;
        g -= term * p.iters_per_term;
        const int tap = g / p.chunks_per_tap;
        const int chunk = g - tap * p.chunks_per_tap;
        const CUtensorMap* ma = (term == 1) ? &tmA2 : &tmA;
        tma_load_2d(smem + s * STAGE_BYTES, ma, p.a_k0 + chunk * BK, p.a_row0 + m0 + tap * p.a_tap_rowshift, &full[s]);
      };
      // the B operand (weights) never depends on the previous kernel: with PDL its first ring-full of tiles streams in
      // while the prerequisite grid is still finishing; the A operand (activations) is loaded after the dependency wait
      const int pre = p.b_independent ? min(n_it, STAGES) : 0;
      for (int it = 0; it < pre; ++it) {
        mbar_expect_tx(&full[it % STAGES], STAGE_BYTES);
        load_b(it);
      }
      t2v_pdl_wait();
      for (int it = 0; it < pre; ++it) load_a(it);
      for (int it = pre; it < n_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        load_a(it);
        load_b(it);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A/B format, K-major both, N>>3, M>>4
      const uint32_t fmt = (ESIZE == 4) ? 2u : (p.fmt16_fp16 ? 0u : 1u);   // TF32 : FP16 / BF16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(MMA_M >> 4) << 24);
      t2v_pdl_wait();                  // block in hardware (not in the mbarrier spin) while the prerequisite grid runs
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = make_kmajor_sw128_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x 32 bytes of K per 128-byte row
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)   // the second 128-row block of A starts 128 rows x 128 B further
            tc_mma<ESIZE == 4>(tmem_base + (uint32_t)(mb * BN), adesc + (uint64_t)(2 * k + mb * ((128 * 128) >> 4)),
                               bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);          // frees the smem slot when these MMAs retire
      }
      tc_commit(tmem_full);            // accumulator complete
    }
  } else {
    // epilogue warps 2..5: warp w may touch TMEM lanes 32*(w%4) .. +31
    t2v_pdl_wait();
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    // accumulator row held by this thread: M=128 -> TMEM lane == row; M=64 -> rows 16q..16q+15 sit in lanes 32q..32q+15
    const bool vec_ok = ((p.ldd & 3) == 0) && ((((uintptr_t)p.D) & 15) == 0) && ((p.split_stride & 3) == 0);
#pragma unroll 1
    for (int mb = 0; mb < MB; ++mb) {
    const int row = (BM != 64) ? (m0 + mb * 128 + q * 32 + lane) : (m0 + q * 16 + lane);
    const bool lane_has_row = (BM != 64) || (lane < 16);
    float* drow = p.D + (long long)blockIdx.z * p.split_stride + (long long)row * p.ldd;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * BN + c0), v);
      if (row < p.M && lane_has_row) {
        const int col0 = n0 + c0;
        if (p.epi_atomic != 1 && vec_ok && col0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j + 0]) * p.alpha;
            o.y = __uint_as_float(v[j + 1]) * p.alpha;
            o.z = __uint_as_float(v[j + 2]) * p.alpha;
            o.w = __uint_as_float(v[j + 3]) * p.alpha;
            if (p.bias && blockIdx.z == 0) {
              o.x += p.bias[col0 + j + 0]; o.y += p.bias[col0 + j + 1];
              o.z += p.bias[col0 + j + 2]; o.w += p.bias[col0 + j + 3];
            }
            if (p.epi_atomic == 2) {          // D += tile (single writer per element: splits == 1)
              const float4 old = *reinterpret_cast<const float4*>(drow + col0 + j);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(drow + col0 + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < p.N) {
              float o = __uint_as_float(v[j]) * p.alpha;
              if (p.bias && blockIdx.z == 0) o += p.bias[col];
              if (p.epi_atomic == 1) atomicAdd(drow + col, o);
              else if (p.epi_atomic == 2) drow[col] += o;
              else drow[col] = o;
            }
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(BN * MB)) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-reduction GEMM with MN-major operands:  D[i, j] (+)= alpha * sum_r A[a_row0 + r, m0 + i] * B[b_row0 + r, n0 + j].
// Both operands are read in their natural row-major layout (rows = the reduction index), so the weight-gradient GEMMs
// dW = dY^T X (reference: autograd of nn.Linear / Conv1d, model.py:91-148) need no transposed copies.  A TMA box is
// {32 floats of MN (128 B), BKR reduction rows} and lands as BKR 128-byte rows.  For 4-byte (tf32) MN-major operands the tensor core
// accepts exactly one shared-memory layout: 128-byte swizzle with a 32-BYTE atom (UMMA layout type 1, "SWIZZLE_128B_BASE32B";
// TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units, i.e. 4 reduction
// rows x 128 B per 512-byte swizzle atom, LBO = next group of 32 MN elements (= one TMA box), SBO = next group of 4 rows = 512 B.
// One tf32 MMA consumes 8 reduction rows (two atoms), so the K advance is +1024 B.  The plain SWIZZLE_128B layout with the MN-major
// bits set is silently computed as zeros by the hardware (profiles/r02_mn_major_probe.txt, tools/mn_probe.cu).
constexpr int BKR = 32;                      // reduction rows per stage
// one 16-byte reduction into global memory (sm_90+): a thread's 32 accumulator columns are contiguous in D, so the split-K epilogue
// issues 8 of these per 32 columns instead of 32 scalar atomics
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((BKR * 128) >> 4) << 16;   // LBO: next group of 32 MN elements = next box
  d |= (uint64_t)(512 >> 4) << 32;           // SBO: next group of 4 reduction rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
  return d;
}

// MB = number of 128-row blocks of D a CTA owns (1 or 2): with MB = 2 the B tile of a stage feeds two accumulators (TMEM columns
// [0, BN) and [BN, 2 BN)), which cuts the L2 -> shared-memory operand traffic per FLOP by a third -- these GEMMs are bound by it
// (fp32 operands: 42 FLOP per operand byte for a 128 x 256 tile, 64 for 256 x 256).
template <int BN, int STAGES, int MB>
__global__ void __launch_bounds__(192, 1)
gemm_tc_mn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcParams p) {
  constexpr int BM = 128 * MB;
  constexpr int A_BYTES = BM * BKR * 4;
  constexpr int B_BYTES = BN * BKR * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int BOX = 32 * BKR * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // Conv1d weight gradients (b_tap_stride = Ci > 0): D is [n_a, taps * Ci]; the column block of tap t reads the SAME B columns from
  // rows shifted by t, so one launch covers all taps (BN divides Ci: a tile never straddles two taps)
  const int tap = p.b_tap_stride > 0 ? n0 / p.b_tap_stride : 0;
  // batched form (batch_a_rows > 0): blockIdx.z is a batch index instead of a split -- A rows start at z * batch_a_rows, B columns
  // at z * batch_b_cols, D at z * split_stride; the whole reduction belongs to this CTA
  const bool batched = p.batch_a_rows > 0;
  const int nb0 = n0 - tap * p.b_tap_stride + (batched ? (int)blockIdx.z * p.batch_b_cols : 0), brow0 = p.b_row0 + tap;
  const int arow0 = p.a_row0 + (batched ? (int)blockIdx.z * p.batch_a_rows : 0);
  const int it0 = batched ? 0 : blockIdx.z * p.iters_per_split;
  const int n_it = min(p.iters_per_split, p.chunks_per_tap - it0);   // chunks_per_tap = total 32-row iterations; the last split may be short

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"((uint32_t)(BN * MB)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // boxes that start beyond the MN extent of an operand are skipped by TMA (fully out of bounds -> zero fill)
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const int r = (it0 + it) * BKR;
        uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
        for (int g = 0; g < BM / 32; ++g) tma_load_2d(sa + g * BOX, &tmA, m0 + 32 * g, arow0 + r, &full[s]);
#pragma unroll
        for (int g = 0; g < BN / 32; ++g) tma_load_2d(sa + A_BYTES + g * BOX, &tmB, nb0 + 32 * g, brow0 + r, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D=f32, A=B=tf32, both MN-major (bits 15 / 16)
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_mnmajor_sw128_desc(sa);
        const uint64_t bdesc = make_mnmajor_sw128_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BKR / 8; ++k)      // 8 reduction rows (1024 B of every MN group) per instruction
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)      // the second 128-row block of A starts 4 boxes (4 * BOX bytes) further
            tc_mma<true>(tmem_base + (uint32_t)(mb * BN), adesc + (uint64_t)(64 * k + mb * ((4 * BOX) >> 4)), bdesc + (uint64_t)(64 * k),
                         idesc, (it > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  } else if (n_it > 0) {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const bool vec4 = ((((uintptr_t)p.D) & 15) == 0) && (p.ldd % 4 == 0) && (p.split_stride % 4 == 0);
#pragma unroll 1
    for (int mb = 0; mb < MB; ++mb) {
    const int row = m0 + mb * 128 + q * 32 + lane;
    float* drow = p.D + (long long)blockIdx.z * p.split_stride + (long long)row * p.ldd;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * BN + c0), v);
      if (row < p.M) {
        if (p.epi_atomic == 1 && vec4 && n0 + c0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(drow + n0 + c0 + j, __uint_as_float(v[j]) * p.alpha, __uint_as_float(v[j + 1]) * p.alpha,
                       __uint_as_float(v[j + 2]) * p.alpha, __uint_as_float(v[j + 3]) * p.alpha);
        } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          if (col < p.N) {
            const float o = __uint_as_float(v[j]) * p.alpha;
            if (p.epi_atomic == 1) atomicAdd(drow + col, o);
            else if (p.epi_atomic == 2) drow[col] += o;
            else drow[col] = o;
          }
        }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(BN * MB)) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same row reduction over 16-bit operands (fp16 / bf16, kind::f16).  For 2-byte MN-major operands the plain 128-byte swizzle
// is the accepted layout (UMMA layout type 2, TMA SWIZZLE_128B): a box is {64 elements of MN (128 B), BKR16 reduction rows}, LBO = one box,
// SBO = 1024 B (8 rows), one K=16 instruction = two atoms -> +2048 B per instruction (profiles/tools/mn16_probe.cu).
constexpr int BKR16 = 64;
__device__ __forceinline__ uint64_t make_mnmajor16_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((BKR16 * 128) >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int BN, int STAGES, int MB>
__global__ void __launch_bounds__(192, 1)
gemm_tc_mn16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcParams p, int fmt_bf16) {
  constexpr int BM = 128 * MB;
  constexpr int BOX = 64 * BKR16 * 2;                    // 8 KB
  constexpr int A_BYTES = (BM / 64) * BOX;
  constexpr int B_BYTES = (BN / 64) * BOX;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = (uint32_t*)(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tap = p.b_tap_stride > 0 ? n0 / p.b_tap_stride : 0;      // see gemm_tc_mn_kernel: one launch for all taps of a Conv1d dW
  const int nb0 = n0 - tap * p.b_tap_stride, brow0 = p.b_row0 + tap;
  const int it0 = blockIdx.z * p.iters_per_split;
  const int n_it = min(p.iters_per_split, p.chunks_per_tap - it0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"((uint32_t)(BN * MB)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const int r = (it0 + it) * BKR16;
        uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
        for (int g = 0; g < BM / 64; ++g) tma_load_2d(sa + g * BOX, &tmA, m0 + 64 * g, p.a_row0 + r, &full[s]);
#pragma unroll
        for (int g = 0; g < BN / 64; ++g) tma_load_2d(sa + A_BYTES + g * BOX, &tmB, nb0 + 64 * g, brow0 + r, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t fmt = fmt_bf16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_mnmajor16_desc(sa);
        const uint64_t bdesc = make_mnmajor16_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BKR16 / 16; ++k)
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)      // the second 128-row block of A starts 2 boxes further
            tc_mma<false>(tmem_base + (uint32_t)(mb * BN), adesc + (uint64_t)(128 * k + mb * ((2 * BOX) >> 4)), bdesc + (uint64_t)(128 * k),
                          idesc, (it > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  } else if (n_it > 0) {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const float alpha = p.alpha * (p.alpha_dev ? *p.alpha_dev : 1.f);
    const int q = warp & 3;
    // column split (n_split > 0): columns [n_split, N) of the product belong to a second matrix D2 (tiles never straddle the split):
    // dW = [dW_ih | dW_hh] of an LSTM cell goes straight into the two parameters' gradient tensors
    const bool second = p.n_split > 0 && n0 >= p.n_split;
    float* const Dm = second ? p.D2 : p.D;
    const long long ldm = second ? p.ldd2 : p.ldd;
    const int cshift = second ? p.n_split : 0;
    const bool vec4 = ((((uintptr_t)Dm) & 15) == 0) && (ldm % 4 == 0);
#pragma unroll 1
    for (int mb = 0; mb < MB; ++mb) {
      const int row = m0 + mb * 128 + q * 32 + lane;
      float* drow = Dm + (long long)row * ldm - cshift;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * BN + c0), v);
        if (row < p.M) {
          if (p.epi_atomic == 1 && vec4 && n0 + c0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(drow + n0 + c0 + j, __uint_as_float(v[j]) * alpha, __uint_as_float(v[j + 1]) * alpha,
                         __uint_as_float(v[j + 2]) * alpha, __uint_as_float(v[j + 3]) * alpha);
          } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = n0 + c0 + j;
            if (col < p.N) {
              const float o = __uint_as_float(v[j]) * alpha;
              if (p.epi_atomic == 1) atomicAdd(drow + col, o);
              else if (p.epi_atomic == 2) drow[col] += o;
              else drow[col] = o;
            }
          }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(BN * MB)) : "memory");
  }
}

template <int BN, int ESIZE>
constexpr int stages_for() { return (BN == 256) ? 4 : (BN == 128 ? 6 : 8); }

template <int BN, int ESIZE, int BM = 128, int STG = 0>
int launch_gemm_tc(const T2VGemmTcPlan* plan, const GemmTcParams& p, int splits, cudaStream_t st) {
  const CUtensorMap& tmA = plan->tmA;
  const CUtensorMap& tmB = plan->tmB;
  const CUtensorMap& tmA2 = p.iters_per_term > 0 ? plan->tmA2 : plan->tmA;
  const CUtensorMap& tmB2 = p.iters_per_term > 0 ? plan->tmB2 : plan->tmB;
  constexpr int STAGES = (STG > 0) ? STG : ((BM == 64) ? 8 : (BM == 256 ? 3 : stages_for<BN, ESIZE>()));
  constexpr int smem = STAGES * (BM * 128 + BN * 128) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, ESIZE, STAGES, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(t2v_ceil_div(p.M, BM), t2v_ceil_div(p.N, BN), splits);
  T2V_CUDA_CHECK(t2v_launch(gemm_tc_kernel<BN, ESIZE, STAGES, BM>, grid, dim3(192), (size_t)smem, st, p.pdl != 0, 1, tmA, tmB, tmA2,
                            tmB2, p));
  T2V_COUNT_LAUNCH();
  return 0;
}

int encode_2d(CUtensorMap* map, const void* base, int esize, long long inner, long long rows, long long row_stride_elems,
              int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { t2v_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -2; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(row_stride_elems * esize)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = (esize == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    t2v_set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld rows=%lld stride=%lld esize=%d)", (int)r,
                  inner, rows, row_stride_elems, esize);
    return -3;
  }
  return 0;
}

}  // namespace

int t2v_encode_tmap_2d(CUtensorMap* map, const void* base, int esize, long long inner, long long rows,
                       long long row_stride_elems, int box_rows) {
  return encode_2d(map, base, esize, inner, rows, row_stride_elems, box_rows);
}

// D[M,N] = alpha * sum_{tap,k} A[a_row0 + m + tap*a_tap_rowshift, a_k0 + k] * B[b_row0 + n, b_k0 + tap*b_tap_stride + k] (+ bias[n])
//   A: `a_rows` x `a_inner` elements, row stride lda; B: `b_rows` rows of `b_inner` elements, row stride ldb.
//   esize 4 => fp32 storage / tf32 math, esize 2 => bf16.  splits>1: partial sums go to D + z*split_stride
//   (epi_atomic=0) or are atomically added into D (epi_atomic=1; caller pre-initialises D).
int t2v_gemm_tc_plan(T2VGemmTcPlan* plan, const void* A, long long lda, long long a_rows, long long a_inner, const void* B,
                     long long ldb, long long b_rows, long long b_inner, long long ldd, int M, int N, int k_sub, int taps,
                     int a_tap_rowshift, int b_tap_stride, int a_k0, int b_k0, int esize, int splits,
                     long long split_stride, int epi_atomic, float alpha, int bn_hint) {
  T2V_ARG_CHECK(esize == 4 || esize == 2, "esize must be 4 (tf32) or 2 (bf16)");
  T2V_ARG_CHECK(M > 0 && N > 0 && k_sub > 0 && taps >= 1 && splits >= 1, "shape");
  T2V_ARG_CHECK((((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0, "operand base must be 16-byte aligned");
  T2V_ARG_CHECK((lda * esize) % 16 == 0 && (ldb * esize) % 16 == 0, "row strides must be multiples of 16 bytes");
  const int BK = 128 / esize;
  const int cpt = t2v_ceil_div(k_sub, BK);
  const int total = cpt * taps;
  T2V_ARG_CHECK(total % splits == 0, "split count must divide the K iteration count");
  T2V_ARG_CHECK(splits == 1 || epi_atomic || split_stride > 0, "split_stride required for partial stores");
  int BN = bn_hint;
  if (BN != 64 && BN != 128 && BN != 256) BN = (N <= 64) ? 64 : 128;
  static const bool allow_m64 = !(getenv("T2V_GEMM_M64") && getenv("T2V_GEMM_M64")[0] == '0');
  static const bool allow_m256 = !(getenv("T2V_GEMM_M256") && getenv("T2V_GEMM_M256")[0] == '0');
  // 256 x 256 tiles for the large GEMMs: at least ~2 waves of 148 CTAs must remain
  const bool big = allow_m256 && BN == 256 && splits == 1 && (long long)t2v_ceil_div(M, 256) * t2v_ceil_div(N, 256) >= 296;
  const int BM = big ? 256 : ((M <= 64 && BN == 128 && esize == 4 && allow_m64) ? 64 : 128);
  plan->BM = BM;
  int r = encode_2d(&plan->tmA, A, esize, a_inner, a_rows, lda, BM);
  if (r) return r;
  r = encode_2d(&plan->tmB, B, esize, b_inner, b_rows, ldb, BN);
  if (r) return r;
  GemmTcParams& p = plan->p;
  p.D = nullptr; p.ldd = ldd; p.split_stride = split_stride; p.bias = nullptr; p.M = M; p.N = N;
  p.iters_per_split = total / splits; p.chunks_per_tap = cpt; p.a_tap_rowshift = a_tap_rowshift;
  p.b_tap_stride = b_tap_stride; p.epi_atomic = epi_atomic; p.alpha = alpha;
  p.a_row0 = 0; p.b_row0 = 0; p.a_k0 = a_k0; p.b_k0 = b_k0; p.b_evict_last = 0; p.b_independent = 0; p.pdl = 0;
  p.iters_per_term = 0; p.fmt16_fp16 = 0; p.batch_a_rows = 0; p.batch_b_cols = 0; p.D2 = nullptr; p.ldd2 = 0; p.n_split = 0;
  plan->BN = BN; plan->esize = esize; plan->splits = splits;
  return 0;
}

int t2v_gemm_tc_run(const T2VGemmTcPlan* plan, int a_row0, int b_row0, float* D, const float* bias, cudaStream_t stream) {
  GemmTcParams p = plan->p;
  p.a_row0 = a_row0; p.b_row0 = b_row0; p.D = D; p.bias = bias;
  const int BN = plan->BN, splits = plan->splits;
  if (plan->esize == 4) {
    if (BN == 128 && plan->BM == 64) {
      // 4 stages = 96 KB of shared memory: two CTAs fit on an SM, so the step GEMMs of the two decoder chains co-reside
      static const bool deep = getenv("T2V_M64_STAGES") && getenv("T2V_M64_STAGES")[0] == '8';
      return deep ? launch_gemm_tc<128, 4, 64, 8>(plan, p, splits, stream)
                  : launch_gemm_tc<128, 4, 64, 4>(plan, p, splits, stream);
    }
    if (BN == 64) return launch_gemm_tc<64, 4>(plan, p, splits, stream);
    if (BN == 128) return launch_gemm_tc<128, 4>(plan, p, splits, stream);
    if (plan->BM == 256) return launch_gemm_tc<256, 4, 256>(plan, p, splits, stream);
    return launch_gemm_tc<256, 4>(plan, p, splits, stream);
  } else {
    if (BN == 64) return launch_gemm_tc<64, 2>(plan, p, splits, stream);
    if (BN == 128) return launch_gemm_tc<128, 2>(plan, p, splits, stream);
    if (plan->BM == 256) return launch_gemm_tc<256, 2, 256>(plan, p, splits, stream);
    return launch_gemm_tc<256, 2>(plan, p, splits, stream);
  }
}

T2V_API int t2v_gemm_tc(const void* A, long long lda, long long a_rows, long long a_inner, const void* B, long long ldb,
                        long long b_rows, long long b_inner, float* D, long long ldd, const float* bias, int M, int N,
                        int k_sub, int taps, int a_tap_rowshift, int b_tap_stride, int a_k0, int b_k0, int esize,
                        int splits, long long split_stride, int epi_atomic, float alpha, int bn_hint,
                        cudaStream_t stream) {
  T2VGemmTcPlan plan;
  int r = t2v_gemm_tc_plan(&plan, A, lda, a_rows, a_inner, B, ldb, b_rows, b_inner, ldd, M, N, k_sub, taps, a_tap_rowshift,
                           b_tap_stride, a_k0, b_k0, esize, splits, split_stride, epi_atomic, alpha, bn_hint);
  if (r) return r;
  return t2v_gemm_tc_run(&plan, 0, 0, D, bias, stream);
}

// Split (error-compensated) form of t2v_gemm_tc for fp32 operands given as hi + lo parts on the tf32 grid:
//   D = alpha * (A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T) (+ bias)   -- the three terms run through ONE accumulator (3x the K
// loop, no read-modify-write of D), which recovers fp32-level accuracy from tf32 tensor-core products.  Same addressing as
// t2v_gemm_tc (taps included); A_lo has the layout of A_hi, B_lo the layout of B_hi.
T2V_API int t2v_gemm_tc_split3(const float* A_hi, const float* A_lo, long long lda, long long a_rows, long long a_inner,
                               const float* B_hi, const float* B_lo, long long ldb, long long b_rows, long long b_inner, float* D,
                               long long ldd, const float* bias, int M, int N, int k_sub, int taps, int a_tap_rowshift,
                               int b_tap_stride, int a_k0, int b_k0, float alpha, int bn_hint, cudaStream_t stream) {
  T2V_ARG_CHECK(A_hi && A_lo && B_hi && B_lo && D, "null operand");
  T2V_ARG_CHECK((((uintptr_t)A_lo) & 15) == 0 && (((uintptr_t)B_lo) & 15) == 0, "operand base must be 16-byte aligned");
  T2VGemmTcPlan plan;
  int r = t2v_gemm_tc_plan(&plan, A_hi, lda, a_rows, a_inner, B_hi, ldb, b_rows, b_inner, ldd, M, N, k_sub, taps, a_tap_rowshift,
                           b_tap_stride, a_k0, b_k0, 4, 1, 0, 0, alpha, bn_hint);
  if (r) return r;
  r = encode_2d(&plan.tmA2, A_lo, 4, a_inner, a_rows, lda, plan.BM);
  if (r) return r;
  r = encode_2d(&plan.tmB2, B_lo, 4, b_inner, b_rows, ldb, plan.BN);
  if (r) return r;
  plan.p.iters_per_term = plan.p.iters_per_split;
  plan.p.iters_per_split *= 3;
  return t2v_gemm_tc_run(&plan, 0, 0, D, bias, stream);
}

T2V_API int t2v_gemm_tc_split3_16(const void* A_hi, const void* A_lo, long long lda, long long a_rows, long long a_inner,
                                  const void* B_hi, const void* B_lo, long long ldb, long long b_rows, long long b_inner, float* D,
                                  long long ldd, const float* bias, int M, int N, int k_sub, int taps, int a_tap_rowshift,
                                  int b_tap_stride, int a_k0, int b_k0, float alpha, int bn_hint, cudaStream_t stream) {
  T2V_ARG_CHECK(A_hi && A_lo && B_hi && B_lo && D, "null operand");
  T2V_ARG_CHECK((((uintptr_t)A_lo) & 15) == 0 && (((uintptr_t)B_lo) & 15) == 0, "operand base must be 16-byte aligned");
  T2VGemmTcPlan plan;
  int r = t2v_gemm_tc_plan(&plan, A_hi, lda, a_rows, a_inner, B_hi, ldb, b_rows, b_inner, ldd, M, N, k_sub, taps, a_tap_rowshift,
                           b_tap_stride, a_k0, b_k0, 2, 1, 0, 0, alpha, bn_hint);
  if (r) return r;
  r = encode_2d(&plan.tmA2, A_lo, 2, a_inner, a_rows, lda, plan.BM);
  if (r) return r;
  r = encode_2d(&plan.tmB2, B_lo, 2, b_inner, b_rows, ldb, plan.BN);
  if (r) return r;
  plan.p.fmt16_fp16 = 1;
  plan.p.iters_per_term = plan.p.iters_per_split;
  plan.p.iters_per_split *= 3;
  return t2v_gemm_tc_run(&plan, 0, 0, D, bias, stream);
}

namespace {
int encode_mn(CUtensorMap* map, const void* base, long long cols, long long rows, long long ld) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { t2v_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -2; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * 4)};
  cuuint32_t box[2] = {32u, (cuuint32_t)BKR};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    t2v_set_error("cuTensorMapEncodeTiled (MN-major) failed with CUresult %d (cols=%lld rows=%lld ld=%lld)", (int)r, cols, rows, ld);
    return -3;
  }
  return 0;
}
template <int BN, int STAGES, int MB = 1>
int launch_gemm_tc_mn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcParams& p, int splits, cudaStream_t st) {
  constexpr int smem = STAGES * (128 * MB + BN) * BKR * 4 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_mn_kernel<BN, STAGES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(t2v_ceil_div(p.M, 128 * MB), t2v_ceil_div(p.N, BN), splits);
  gemm_tc_mn_kernel<BN, STAGES, MB><<<grid, 192, smem, st>>>(tmA, tmB, p);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
}  // namespace

// D[n_a, n_b] (+)= alpha * sum_{r < rows} A[a_row0 + r, i] * B[b_row0 + r, j]   (fp32 storage, tf32 math; operands must already be on
// the tf32 grid).  A: [>= a_row0 + rows, n_a] row-major with row stride lda, B likewise.  epi: 0 store (splits == 1 or split_stride),
// 1 atomicAdd into a pre-initialised D, 2 non-atomic D += (splits == 1).  The reduction is cut into `splits` equal ranges of 32-row
// iterations (splits must divide ceil(rows / 32)); rows past the end are zero-filled by TMA.
// taps > 1 (Conv1d weight gradient in tap-major form): D is [n_a, taps * n_b] and column block t is the reduction against B rows
// [b_row0 + t, b_row0 + t + rows) -- all taps in one launch, so the split count (and the atomic passes over D) stay small.
T2V_API int t2v_gemm_tc_rowred(const float* A, long long lda, int n_a, long long a_row0, const float* B, long long ldb, int n_b,
                               long long b_row0, float* D, long long ldd, long long rows, int splits, long long split_stride,
                               int epi, float alpha, int taps, cudaStream_t stream) {
  T2V_ARG_CHECK(A && B && D && n_a > 0 && n_b > 0 && rows > 0 && splits >= 1 && taps >= 1, "shape");
  T2V_ARG_CHECK(taps == 1 || n_b % 256 == 0 || n_b == 128 || n_b == 64, "taps > 1: the N tile (256 / 128 / 64) must divide n_b");
  T2V_ARG_CHECK((((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0, "operand base must be 16-byte aligned");
  T2V_ARG_CHECK((lda * 4) % 16 == 0 && (ldb * 4) % 16 == 0, "row strides must be multiples of 16 bytes");
  const int iters = t2v_ceil_div(rows, BKR);
  T2V_ARG_CHECK(splits <= iters, "more splits than 32-row iterations");
  T2V_ARG_CHECK(splits == 1 || epi == 1, "split-K partial tiles are accumulated with atomics (epi 1)");
  CUtensorMap tmA, tmB;
  int r = encode_mn(&tmA, A, n_a, a_row0 + rows, lda);       // row extent = end of the reduction range: the tail is zero-filled
  if (r) return r;
  r = encode_mn(&tmB, B, n_b, b_row0 + rows + (taps - 1), ldb);   // tap t reads B rows [b_row0 + t, b_row0 + t + rows): must exist
  if (r) return r;
  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.ldd = ldd; p.split_stride = split_stride; p.M = n_a; p.N = n_b * taps; p.iters_per_split = t2v_ceil_div(iters, splits); p.chunks_per_tap = iters;
  p.epi_atomic = epi; p.alpha = alpha; p.a_row0 = (int)a_row0; p.b_row0 = (int)b_row0; p.b_tap_stride = taps > 1 ? n_b : 0;
  if (n_b > 128 && n_b % 256 == 0 && n_a >= 512) return launch_gemm_tc_mn<256, 3, 2>(tmA, tmB, p, splits, stream);   // 256 x 256 tiles
  if (n_b > 128 && n_b % 256 == 0) return launch_gemm_tc_mn<256, 4>(tmA, tmB, p, splits, stream);
  if (n_b > 64) return launch_gemm_tc_mn<128, 6>(tmA, tmB, p, splits, stream);
  return launch_gemm_tc_mn<64, 8>(tmA, tmB, p, splits, stream);
}

// batched row reduction: D[z][i, j] = alpha * sum_{r < rows} A[z * a_batch_rows + r, i] * B[r, z * b_batch_cols + j], z < batch.
// A: [batch * a_batch_rows, n_a] row-major (row stride lda), B: [rows, batch * b_batch_cols] (row stride ldb), D: batch matrices
// [n_a, n_b] (row stride ldd, d_batch_stride floats apart), plain stores.  The reduction tail past `rows` is zero-filled on the B side
// (A's tail rows belong to the next batch: finite values times zero).  Use: d(memory)[b] = alignments[b]^T dctx[:, b, :]
// (autograd of attention_context = bmm(attention_weights, memory), model.py:84-85) over the saved alignments, tf32 operands.
T2V_API int t2v_gemm_tc_rowred_batched(const float* A, long long lda, int n_a, long long a_batch_rows, const float* B, long long ldb,
                                       int n_b, long long b_batch_cols, float* D, long long ldd, long long d_batch_stride,
                                       long long rows, int batch, float alpha, cudaStream_t stream) {
  T2V_ARG_CHECK(A && B && D && n_a > 0 && n_a <= 128 && n_b > 0 && rows > 0 && batch >= 1 && batch <= 65535, "shape (n_a <= 128)");
  T2V_ARG_CHECK((((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0, "operand base must be 16-byte aligned");
  T2V_ARG_CHECK((lda * 4) % 16 == 0 && (ldb * 4) % 16 == 0, "row strides must be multiples of 16 bytes");
  T2V_ARG_CHECK(n_b % 256 == 0 && b_batch_cols >= n_b && a_batch_rows >= rows, "n_b must be a multiple of the 256-column tile");
  const int iters = t2v_ceil_div(rows, BKR);
  CUtensorMap tmA, tmB;
  int r = encode_mn(&tmA, A, n_a, (long long)batch * a_batch_rows, lda);
  if (r) return r;
  r = encode_mn(&tmB, B, (long long)(batch - 1) * b_batch_cols + n_b, rows, ldb);
  if (r) return r;
  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.ldd = ldd; p.split_stride = d_batch_stride; p.M = n_a; p.N = n_b; p.iters_per_split = iters; p.chunks_per_tap = iters;
  p.epi_atomic = 0; p.alpha = alpha; p.batch_a_rows = (int)a_batch_rows; p.batch_b_cols = (int)b_batch_cols;
  return launch_gemm_tc_mn<256, 4>(tmA, tmB, p, batch, stream);
}

namespace {
int encode_mn16(CUtensorMap* map, const void* base, long long cols, long long rows, long long ld) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { t2v_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -2; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {64u, (cuuint32_t)BKR16};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    t2v_set_error("cuTensorMapEncodeTiled (MN-major, 16-bit) failed with CUresult %d (cols=%lld rows=%lld ld=%lld)", (int)r, cols, rows, ld);
    return -3;
  }
  return 0;
}
}  // namespace

// D[n_a, n_b] (+)= alpha * (*alpha_dev) * sum_{r < rows} A[a_row0 + r, i] * B[b_row0 + r, j] over 16-bit operands (fmt 1 fp16, 2 bf16).
// n_split > 0: product columns [0, n_split) go to D (row stride ldd), columns [n_split, n_b) to D2 (row stride ldd2).
T2V_API int t2v_gemm_tc_rowred16(const void* A, long long lda, int n_a, long long a_row0, const void* B, long long ldb, int n_b,
                                 long long b_row0, float* D, long long ldd, long long rows, int splits, int epi, float alpha,
                                 const float* alpha_dev, int fmt, int taps, float* D2, long long ldd2, int n_split,
                                 cudaStream_t stream) {
  T2V_ARG_CHECK(n_split == 0 || (D2 && n_split % 256 == 0 && n_split < n_b && taps == 1), "column split: D2 and a multiple of 256");
  T2V_ARG_CHECK(A && B && D && n_a > 0 && n_b > 0 && rows > 0 && splits >= 1 && (fmt == 1 || fmt == 2) && taps >= 1, "shape / fmt");
  T2V_ARG_CHECK((((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0, "operand base must be 16-byte aligned");
  T2V_ARG_CHECK((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0, "row strides must be multiples of 16 bytes");
  T2V_ARG_CHECK(n_b % 256 == 0 && n_a >= 256, "16-bit row reduction is built for the large weight gradients (n_b % 256 == 0)");
  const int iters = t2v_ceil_div(rows, BKR16);
  T2V_ARG_CHECK(splits <= iters, "more splits than 64-row iterations");
  T2V_ARG_CHECK(splits == 1 || epi == 1, "split-K partial tiles are accumulated with atomics (epi 1)");
  CUtensorMap tmA, tmB;
  int r = encode_mn16(&tmA, A, n_a, a_row0 + rows, lda);
  if (r) return r;
  r = encode_mn16(&tmB, B, n_b, b_row0 + rows + (taps - 1), ldb);
  if (r) return r;
  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.ldd = ldd; p.M = n_a; p.N = n_b * taps; p.iters_per_split = t2v_ceil_div(iters, splits); p.chunks_per_tap = iters;
  p.epi_atomic = epi; p.alpha = alpha; p.alpha_dev = alpha_dev; p.a_row0 = (int)a_row0; p.b_row0 = (int)b_row0;
  p.b_tap_stride = taps > 1 ? n_b : 0;
  p.D2 = D2; p.ldd2 = ldd2; p.n_split = n_split;
  constexpr int STAGES = 3, MB = 2, BN = 256;
  constexpr int smem = STAGES * (128 * MB + BN) * BKR16 * 2 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_mn16_kernel<BN, STAGES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(t2v_ceil_div(n_a, 128 * MB), t2v_ceil_div(n_b * taps, BN), splits);
  gemm_tc_mn16_kernel<BN, STAGES, MB><<<grid, 192, smem, stream>>>(tmA, tmB, p, fmt == 2 ? 1 : 0);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
