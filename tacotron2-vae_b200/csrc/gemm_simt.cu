// Exact-fp32 reference GEMM on the CUDA cores (FFMA), fully strided:
//   C[m,n] = alpha * sum_k A[m*a_rs + k*a_cs] * B[n*b_rs + k*b_cs]  + beta*C[m,n] + bias[n]
// Any transpose / overlapping-row view (conv-as-GEMM over the padded channels-last layout, STFT frames)
// is expressed through the strides.  This is the "fp32" precision mode used for tight parity checks and for
// the small GEMMs; the large ones go through gemm_tc.cu (tcgen05).
#include "t2v_common.cuh"

namespace {

constexpr int BK = 16;
// 256 threads; tile BM x BN with TM x TN outputs per thread: 128x128 (8x8) for wide outputs, 128x64 (8x4) and 256x32 (8x4) for the
// narrow ones (reference-encoder convolutions with 32 / 64 filters, VAE head): no FMAs wasted on columns that do not exist
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long a_rs, long long a_cs, const float* __restrict__ B,
                 long long b_rs, long long b_cs, float* __restrict__ C, long long c_rs, int M, int N, int K,
                 float alpha, float beta, const float* __restrict__ bias, long long a_bs, long long b_bs,
                 long long c_bs) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  A += (long long)blockIdx.z * a_bs;
  B += (long long)blockIdx.z * b_bs;
  C += (long long)blockIdx.z * c_bs;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  constexpr int NTX = BN / TN;
  static_assert((BM / TM) * NTX == 256, "tile shape");
  const int tx = tid % NTX, ty = tid / NTX;   // thread computes rows ty*TM.., cols tx*TN..
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (a_cs == 1);
  const bool b_kfast = (b_cs == 1);
  for (int k0 = 0; k0 < K; k0 += BK) {
    // stage A tile (BM x BK) and B tile (BN x BK), transposed into [k][m]
#pragma unroll
    for (int i = tid; i < BM * BK; i += 256) {
      int mm, kk;
      if (a_kfast) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
    }
#pragma unroll
    for (int i = tid; i < BN * BK; i += 256) {
      int nn, kk;
      if (b_kfast) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? B[gn * b_rs + gk * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[gn];
      float* c = C + gm * c_rs + gn;
      if (beta != 0.f) v += beta * (*c);
      *c = v;
    }
  }
}

}  // namespace

T2V_API int t2v_gemm_f32(const float* A, long long a_rs, long long a_cs, const float* B, long long b_rs, long long b_cs,
                         float* C, long long c_rs, int M, int N, int K, float alpha, float beta, const float* bias,
                         int batch, long long a_bs, long long b_bs, long long c_bs, cudaStream_t stream) {
  T2V_ARG_CHECK(A && B && C, "null operand");
  T2V_ARG_CHECK(M > 0 && N > 0 && K > 0 && batch >= 1 && batch <= 65535, "shape");
  const int BM = (N <= 32) ? 256 : 128, BN = (N <= 32) ? 32 : (N <= 64 ? 64 : 128);
  dim3 grid(t2v_ceil_div(N, BN), t2v_ceil_div(M, BM), batch);
  T2V_ARG_CHECK(grid.y <= 65535, "M too large for this launch geometry");
  if (BN == 32)
    gemm_simt_kernel<256, 32, 8, 4><<<grid, 256, 0, stream>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, c_rs, M, N, K, alpha, beta, bias,
                                                             a_bs, b_bs, c_bs);
  else if (BN == 64)
    gemm_simt_kernel<128, 64, 8, 4><<<grid, 256, 0, stream>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, c_rs, M, N, K, alpha, beta, bias,
                                                             a_bs, b_bs, c_bs);
  else
    gemm_simt_kernel<128, 128, 8, 8><<<grid, 256, 0, stream>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, c_rs, M, N, K, alpha, beta, bias,
                                                              a_bs, b_bs, c_bs);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
