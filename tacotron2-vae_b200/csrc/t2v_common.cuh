// Shared helpers for the t2v_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define T2V_API extern "C" __attribute__((visibility("default")))

// ---- error reporting: 0 ok, <0 argument error (nothing launched), >0 cudaError_t passthrough ----
void t2v_set_error(const char* fmt, ...);
#define T2V_ARG_CHECK(cond, msg)                                                    \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      t2v_set_error("%s:%d: argument check failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
      return -1;                                                                    \
    }                                                                               \
  } while (0)
#define T2V_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      t2v_set_error("%s:%d: CUDA error %d: %s", __FILE__, __LINE__, (int)e__, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)
#define T2V_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      t2v_set_error("%s:%d: %s -> CUDA error %d: %s", __FILE__, __LINE__, #expr, (int)e__, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

// counts kernel launches made by this library (bench.py reports it as gpu_launches)
extern unsigned long long g_t2v_launches;
#define T2V_COUNT_LAUNCH() (++g_t2v_launches)

// Programmatic dependent launch (PDL): kernels of the decoder time loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that a kernel's launch + prologue (barrier init, TMEM allocation,
// weight staging -- nothing that depends on the previous kernel) overlaps the previous kernel's tail.  t2v_pdl_wait()
// blocks until every prerequisite grid has completed and flushed; it is a no-op for ordinary launches.
// T2V_PDL_LATE=1: a kernel releases its dependent only after its own wait returned, so at most ONE kernel of the chain
// is pre-launched (with the trigger at the top, the whole chain cascades onto the SMs and the waiting CTAs take shared
// memory / warp slots from the kernel that is actually running).
#ifndef T2V_PDL_LATE
#define T2V_PDL_LATE 1
#endif
__device__ __forceinline__ void t2v_pdl_trigger() {
#if !T2V_PDL_LATE
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void t2v_pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if T2V_PDL_LATE
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
bool t2v_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t t2v_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl && t2v_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline int t2v_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- counter-based dropout RNG: keep(seed, site, idx) is a pure function, so forward, backward and
// the mask-materialisation kernel used by the tests all see the same bits ----
__host__ __device__ __forceinline__ uint32_t t2v_hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)(x >> 11);
}
// uniform in [0,1) with 24 bits
__host__ __device__ __forceinline__ float t2v_uniform(uint64_t seed, uint32_t site, uint64_t idx) {
  uint64_t key = seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)site << 40) + idx;
  return (float)(t2v_hash32(key) & 0xFFFFFFu) * (1.0f / 16777216.0f);
}

// A dropout site: either an explicit keep-mask (float 0/1, element strides given by the kernel) or the RNG.
struct T2VDrop {
  const float* mask;   // nullable
  uint64_t seed;
  uint32_t site;
  float p;             // drop probability; p<=0 => identity
};
// A seed argument is either a value or, with bit 63 set, a device pointer to the value: CUDA-graph replays bake kernel
// arguments, so the per-iteration seed has to live in device memory.
__device__ __forceinline__ uint64_t t2v_resolve_seed(uint64_t seed) {
  return (seed >> 63) ? *reinterpret_cast<const uint64_t*>(seed & 0x7FFFFFFFFFFFFFFFULL) : seed;
}
__device__ __forceinline__ float t2v_keep_scale(const T2VDrop& d, uint64_t logical_idx) {
  if (d.p <= 0.f) return 1.f;
  float keep = d.mask ? d.mask[logical_idx]
                      : (t2v_uniform(t2v_resolve_seed(d.seed), d.site, logical_idx) >= d.p ? 1.f : 0.f);
  return keep * (1.f / (1.f - d.p));
}

// round-to-nearest to the tf32 grid (10-bit mantissa): tcgen05 kind::tf32 TRUNCATES fp32 operands, which biases every
// dot product by ~-1e-3; kernels that produce GEMM operands for the tensor-core path round them on store instead.
__device__ __forceinline__ float t2v_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// 16-bit operand formats of the tensor-core path (kind::f16): the fp32 copy of an operand is rounded onto the same grid as
// its 16-bit copy, so a tf32 GEMM over the fp32 copy and an f16 GEMM over the 16-bit copy see the same numbers.
//   fp16 has the significand of tf32 (11 bits) with a narrower exponent (|x| <= 65504; below 6.1e-5 the spacing is 6e-8):
//   used for forward activations and weights.  bf16 = 8-bit significand, fp32 exponent.
__device__ __forceinline__ uint16_t t2v_f16_bits(float x) {
  uint16_t h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ uint16_t t2v_bf16_bits(float x) {
  uint16_t h;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ float t2v_f16_to_f32(uint16_t h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}
__device__ __forceinline__ float t2v_bf16_to_f32(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }
// operand rounding flag of the ABI (`rnd`): 0 none, 1 tf32 grid, 2 fp16 grid, 3 bf16 grid
__device__ __forceinline__ float t2v_rnd(float x, int flag) {
  if (flag == 0) return x;
  if (flag == 1) return t2v_tf32(x);
  if (flag == 2) return t2v_f16_to_f32(t2v_f16_bits(x));
  return t2v_bf16_to_f32(t2v_bf16_bits(x));
}
// 16-bit encoding of an operand for fmt 1 (fp16) / 2 (bf16)
__device__ __forceinline__ uint16_t t2v_enc16(float x, int fmt) { return fmt == 2 ? t2v_bf16_bits(x) : t2v_f16_bits(x); }

__device__ __forceinline__ float t2v_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// tanh through one MUFU exp + one fast divide: absolute error ~1e-7 (fp32 rounding level), ~4x fewer instructions than
// tanhf; used inside the latency-critical decoder-step kernels.
__device__ __forceinline__ float t2v_tanh(float x) {
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ float t2v_sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` must hold >= 32 floats; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}
