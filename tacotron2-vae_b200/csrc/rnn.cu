// Encoder BiLSTM recurrence (reference model.py:171-190: nn.LSTM, bidirectional, packed by input_lengths), one launch
// per time step for BOTH directions with the recurrent matvec and the cell pointwise math fused:
//   forward  : gates = GX[t] (hoisted input projection) + h_{prev} W_hh^T + b_hh -> i,f,g,o -> c,h ; one warp owns one
//              hidden unit (all four gates, all batch rows), so the cell update happens in registers
//   backward : dh_rec = dgates_{next} W_hh (same warp-per-unit mapping over the transposed weights), then the cell
//              backward for the unit -> dgates_t.
// The batch is processed in tiles of 64 rows (lane -> rows lane, lane+32); exact fp32 FFMA.
#include "t2v_common.cuh"

namespace {

constexpr int UPC = 8;          // hidden units (= warps) per CTA
constexpr int MT = 64;          // batch rows per tile
constexpr int MS = MT + 1;      // (unused) transposed-tile stride

struct BiFwdArgs {
  const float* gx[2]; long long gx_bs;     // per dir: pointer to row (b=0, t) of the pre-gates; batch stride
  const float* w_hh[2];                    // [4H,H]
  const float* b_hh[2];                    // [4H]
  const float* h_prev[2]; float* h_next[2];// [B,H] ping-pong
  float* c_state[2];                       // [B,H] in place
  float* seq_out[2]; long long seq_bs;     // per dir: pointer to out[b=0, t, dir*H]; batch stride
  float* gates_save[2];                    // [B,4H]
  float* c_save[2];                        // [B,H]
  const long long* lens; int t[2];
  int B, H;
};

__global__ void __launch_bounds__(UPC * 32) bilstm_step_fwd_kernel(BiFwdArgs p) {
  extern __shared__ __align__(16) float smf[];
  const int H = p.H, dir = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UPC, u = u0 + warp;
  const int HS = H + 1;                  // padded row stride: lane-varying rows hit distinct banks
  float* hT = smf;                       // [MT][H+1] h_prev tile, row major
  float* ws = hT + ((MT * HS + 3) & ~3); // [H][UPC][4] recurrent weights of this CTA's units, k-major
  const float* W = p.w_hh[dir];
  for (int i = threadIdx.x; i < H * UPC * 4; i += UPC * 32) {
    const int k = i % H, g = (i / H) % 4, uu = i / (4 * H);
    ws[(k * UPC + uu) * 4 + g] = W[((long long)g * H + u0 + uu) * H + k];
  }
  const int t = p.t[dir];
  for (int m0 = 0; m0 < p.B; m0 += MT) {
    __syncthreads();
    {   // stage the h_prev tile: float4 global loads, 8 in flight per thread
      const int nv = MT * H / 4;
      for (int i0 = threadIdx.x; i0 < nv; i0 += UPC * 32 * 16) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = i0 + j * UPC * 32;
          const int m = (i * 4) / H, k = (i * 4) % H;
          v[j] = (i < nv && m0 + m < p.B) ? *reinterpret_cast<const float4*>(p.h_prev[dir] + (long long)(m0 + m) * H + k)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = i0 + j * UPC * 32;
          if (i < nv) {
            const int m = (i * 4) / H, k = (i * 4) % H;
            float* d = hT + m * HS + k;
            d[0] = v[j].x; d[1] = v[j].y; d[2] = v[j].z; d[3] = v[j].w;
          }
        }
      }
    }
    __syncthreads();
    float acc[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float a0 = hT[lane * HS + k], a1 = hT[(32 + lane) * HS + k];
      const float4 w = *reinterpret_cast<const float4*>(ws + (k * UPC + warp) * 4);
      acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
      acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
      acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
      acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = m0 + r * 32 + lane;
      if (b >= p.B) continue;
      const bool live = (p.lens == nullptr) || (t < p.lens[b]);
      float* gs = p.gates_save[dir] + (long long)b * 4 * H + u;
      if (!live) {          // dead step of a packed row: state untouched, zero output (pad_packed_sequence)
        p.seq_out[dir][b * p.seq_bs + u] = 0.f;
        gs[0] = 0.f; gs[H] = 0.f; gs[2 * H] = 0.f; gs[3 * H] = 0.f;
        p.c_save[dir][(long long)b * H + u] = 0.f;
        p.h_next[dir][(long long)b * H + u] = p.h_prev[dir][(long long)b * H + u];
        continue;
      }
      const float* gx = p.gx[dir] + b * p.gx_bs;
      const float* bh = p.b_hh[dir];
      const float ig = t2v_sigmoid(acc[r][0] + gx[u] + bh[u]);
      const float fg = t2v_sigmoid(acc[r][1] + gx[H + u] + bh[H + u]);
      const float gg = tanhf(acc[r][2] + gx[2 * H + u] + bh[2 * H + u]);
      const float og = t2v_sigmoid(acc[r][3] + gx[3 * H + u] + bh[3 * H + u]);
      float* cs = p.c_state[dir] + (long long)b * H + u;
      const float c2 = fg * (*cs) + ig * gg;
      const float h2 = og * tanhf(c2);
      *cs = c2;
      p.h_next[dir][(long long)b * H + u] = h2;
      p.seq_out[dir][b * p.seq_bs + u] = h2;
      gs[0] = ig; gs[H] = fg; gs[2 * H] = gg; gs[3 * H] = og;
      p.c_save[dir][(long long)b * H + u] = c2;
    }
  }
}

struct BiBwdArgs {
  const float* dg_next[2]; long long dg_bs;   // dgates of the previously processed step (row b=0), batch stride; nullable
  const float* w_hhT[2];                      // [H,4H] transposed recurrent weights
  const float* dout[2]; long long dout_bs;    // grad wrt the layer output at (t, dir half); batch stride
  float* dc[2];                               // [B,H] running cell-state gradient (in place)
  const float* gates_save[2]; const float* c_save[2];     // of step t
  const float* c_prev[2];                     // [B,H] cell after the previous step of this direction
  float* dg_out[2];                           // dgates of step t (row b=0), batch stride dg_bs
  const long long* lens; int t[2];
  int B, H;
};

__global__ void __launch_bounds__(UPC * 32) bilstm_step_bwd_kernel(BiBwdArgs p) {
  extern __shared__ __align__(16) float smb[];
  const int H = p.H, K = 4 * H, dir = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UPC, u = u0 + warp;
  constexpr int KC = 256;
  constexpr int DS = KC + 1;
  float* dT = smb;                       // [MT][KC+1] dgates chunk, row major
  float* ws = dT + MT * DS;              // [KC][UPC]
  const int t = p.t[dir];
  for (int m0 = 0; m0 < p.B; m0 += MT) {
    float acc[2] = {0.f, 0.f};
    if (p.dg_next[dir]) {
      for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        {
          const int nv = MT * KC / 4;
          for (int i0 = threadIdx.x; i0 < nv; i0 += UPC * 32 * 16) {
            float4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int i = i0 + j * UPC * 32;
              const int m = (i * 4) / KC, k = (i * 4) % KC;
              v[j] = (i < nv && m0 + m < p.B) ? *reinterpret_cast<const float4*>(p.dg_next[dir] + (m0 + m) * p.dg_bs + k0 + k)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int i = i0 + j * UPC * 32;
              if (i < nv) {
                const int m = (i * 4) / KC, k = (i * 4) % KC;
                float* d = dT + m * DS + k;
                d[0] = v[j].x; d[1] = v[j].y; d[2] = v[j].z; d[3] = v[j].w;
              }
            }
          }
        }
        for (int i = threadIdx.x; i < KC * UPC; i += UPC * 32) {
          const int k = i % KC, uu = i / KC;
          ws[k * UPC + uu] = p.w_hhT[dir][(long long)(u0 + uu) * K + k0 + k];
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < KC; ++k) {
          const float w = ws[k * UPC + warp];
          acc[0] = fmaf(dT[lane * DS + k], w, acc[0]);
          acc[1] = fmaf(dT[(32 + lane) * DS + k], w, acc[1]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = m0 + r * 32 + lane;
      if (b >= p.B) continue;
      float* dg = p.dg_out[dir] + b * p.dg_bs + u;
      const bool live = (p.lens == nullptr) || (t < p.lens[b]);
      if (!live) { dg[0] = 0.f; dg[H] = 0.f; dg[2 * H] = 0.f; dg[3 * H] = 0.f; continue; }
      const float dh = acc[r] + p.dout[dir][b * p.dout_bs + u];
      const float* gs = p.gates_save[dir] + (long long)b * 4 * H + u;
      const float ig = gs[0], fg = gs[H], gg = gs[2 * H], og = gs[3 * H];
      const float c2 = p.c_save[dir][(long long)b * H + u];
      const float tc = tanhf(c2);
      float* dcp = p.dc[dir] + (long long)b * H + u;
      const float dc = *dcp + dh * og * (1.f - tc * tc);
      const float cp = p.c_prev[dir][(long long)b * H + u];
      dg[0] = dc * gg * ig * (1.f - ig);
      dg[H] = dc * cp * fg * (1.f - fg);
      dg[2 * H] = dc * ig * (1.f - gg * gg);
      dg[3 * H] = dh * tc * og * (1.f - og);
      *dcp = dc * fg;
    }
  }
}

}  // namespace

// one time step of both directions; arrays are indexed by direction (0 forward, 1 reverse)
T2V_API int t2v_bilstm_step_fwd(const float* gx0, const float* gx1, long long gx_bs, const float* whh0, const float* whh1,
                                const float* bhh0, const float* bhh1, const float* hprev0, const float* hprev1, float* hnext0,
                                float* hnext1, float* c0, float* c1, float* seq0, float* seq1, long long seq_bs, float* gs0,
                                float* gs1, float* cs0, float* cs1, const long long* lens, int t0, int t1, int B, int H,
                                cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && H % UPC == 0 && H <= 512 && H % 4 == 0, "shape");
  BiFwdArgs a;
  a.gx[0] = gx0; a.gx[1] = gx1; a.gx_bs = gx_bs; a.w_hh[0] = whh0; a.w_hh[1] = whh1; a.b_hh[0] = bhh0; a.b_hh[1] = bhh1;
  a.h_prev[0] = hprev0; a.h_prev[1] = hprev1; a.h_next[0] = hnext0; a.h_next[1] = hnext1; a.c_state[0] = c0; a.c_state[1] = c1;
  a.seq_out[0] = seq0; a.seq_out[1] = seq1; a.seq_bs = seq_bs; a.gates_save[0] = gs0; a.gates_save[1] = gs1;
  a.c_save[0] = cs0; a.c_save[1] = cs1; a.lens = lens; a.t[0] = t0; a.t[1] = t1; a.B = B; a.H = H;
  const size_t smem = sizeof(float) * (size_t)(((MT * (H + 1) + 3) & ~3) + H * UPC * 4);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(bilstm_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  bilstm_step_fwd_kernel<<<dim3(H / UPC, 2), UPC * 32, smem, st>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}

T2V_API int t2v_bilstm_step_bwd(const float* dgn0, const float* dgn1, long long dg_bs, const float* whhT0, const float* whhT1,
                                const float* dout0, const float* dout1, long long dout_bs, float* dc0, float* dc1,
                                const float* gs0, const float* gs1, const float* cs0, const float* cs1, const float* cp0,
                                const float* cp1, float* dgo0, float* dgo1, const long long* lens, int t0, int t1, int B, int H,
                                cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && H % UPC == 0 && (4 * H) % 256 == 0, "shape");
  BiBwdArgs a;
  a.dg_next[0] = dgn0; a.dg_next[1] = dgn1; a.dg_bs = dg_bs; a.w_hhT[0] = whhT0; a.w_hhT[1] = whhT1; a.dout[0] = dout0;
  a.dout[1] = dout1; a.dout_bs = dout_bs; a.dc[0] = dc0; a.dc[1] = dc1; a.gates_save[0] = gs0; a.gates_save[1] = gs1;
  a.c_save[0] = cs0; a.c_save[1] = cs1; a.c_prev[0] = cp0; a.c_prev[1] = cp1; a.dg_out[0] = dgo0; a.dg_out[1] = dgo1;
  a.lens = lens; a.t[0] = t0; a.t[1] = t1; a.B = B; a.H = H;
  const size_t smem = sizeof(float) * (size_t)(MT * 257 + 256 * UPC);
  static bool set = false;
  if (!set) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(bilstm_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  bilstm_step_bwd_kernel<<<dim3(H / UPC, 2), UPC * 32, smem, st>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
