// Internal C++ interface of the tcgen05 GEMM (gemm_tc.cu): a "plan" holds the two TMA tensor maps and the launch
// parameters so that a time loop can re-launch the same GEMM on successive row blocks without re-encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

struct GemmTcParams {
  float* D;
  long long ldd;
  long long split_stride;
  const float* bias;
  int M, N;
  int iters_per_split;
  int chunks_per_tap;
  int a_tap_rowshift;
  int b_tap_stride;
  int epi_atomic;
  float alpha;
  int a_row0, b_row0, a_k0, b_k0;
  int b_independent;     // B does not depend on the preceding kernel: prefetch it before the PDL dependency wait
  int pdl;               // launch with programmatic stream serialization
  int b_evict_last;      // load the B operand (weights re-read every time step) with the L2 evict_last policy
  const float* alpha_dev; // nullable device scalar multiplied onto alpha (row-reduction kernels)
  float* D2; long long ldd2; int n_split;   // 16-bit row reduction: product columns >= n_split are written to D2 (row stride ldd2)
  int fmt16_fp16;        // esize 2: 1 = the 16-bit operands are fp16 (default bf16)
  int batch_a_rows, batch_b_cols;   // row-reduction kernels, > 0: blockIdx.z is a batch index (A row / B column offsets per batch)
  int iters_per_term;    // > 0: split (error-compensated) product, the K loop walks 3 terms of iters_per_term iterations each:
                         // (A, B), (A2, B), (A, B2) -- x_hi W_hi + x_lo W_hi + x_hi W_lo in ONE accumulator
};
struct T2VGemmTcPlan {
  CUtensorMap tmA, tmB, tmA2, tmB2;
  GemmTcParams p;
  int BN, esize, splits, BM;
};
int t2v_gemm_tc_plan(T2VGemmTcPlan* plan, const void* A, long long lda, long long a_rows, long long a_inner, const void* B,
                     long long ldb, long long b_rows, long long b_inner, long long ldd, int M, int N, int k_sub, int taps,
                     int a_tap_rowshift, int b_tap_stride, int a_k0, int b_k0, int esize, int splits,
                     long long split_stride, int epi_atomic, float alpha, int bn_hint);
int t2v_gemm_tc_run(const T2VGemmTcPlan* plan, int a_row0, int b_row0, float* D, const float* bias, cudaStream_t stream);
