// Location-sensitive attention, one decoder step, fused per utterance (reference model.py:31-88 called from
// Decoder.decode, model.py:366-372): location conv (2->32, k=31) + location dense (32->128) + query add +
// processed-memory add + tanh + v-projection + length mask + softmax + context reduction + cumulative-weights
// update in ONE kernel (one CTA per batch row; everything between the query GEMM and the decoder-LSTM GEMM).
// The backward kernel is the exact adjoint, accumulating d(memory), d(processed_memory) and per-row partial
// weight gradients in place across the reverse time loop.
#include "t2v_common.cuh"

namespace {

constexpr int NF = 32;     // attention_location_n_filters
constexpr int KS = 31;     // attention_location_kernel_size
constexpr int AD = 128;    // attention_dim
constexpr int ED = 512;    // encoder_embedding_dim
constexpr int HALO = (KS - 1) / 2;

struct AttnFwdArgs {
  const float* qparts; int n_qparts; long long qpart_stride;   // [parts][B][AD]
  const float* w_prev;      // [B,Ti] (nullptr => zeros; step 0)
  long long wprev_rs;
  const float* cum_in;      // [B,Ti]
  float* cum_out;           // [B,Ti]
  const float* pmem;        // [B,Ti,AD]
  const float* mem;         // [B,Ti,ED]
  const float* w_conv;      // [NF,2,KS]
  const float* w_loc;       // [AD,NF]
  const float* v;           // [AD]
  const long long* lens;    // [B] nullable (inference: no mask)
  float mask_value;         // score_mask_value (-inf, or finfo(fp16).min in the reference's fp16 mode)
  float* w_out; long long wout_rs;      // [B,Ti] row stride (alignments[b, t, :])
  float* ctx_out1; long long ctx1_rs;   // [B,ED] destinations of the context
  float* ctx_out2; long long ctx2_rs;
  float* a_save;            // [B,Ti,AD] tanh activations (nullable)
  int B, Ti;
  int rnd;                  // round the context to tf32 on store (tensor-core GEMM operand)
};

__global__ void __launch_bounds__(256) attn_step_fwd_kernel(AttnFwdArgs p) {
  extern __shared__ float sm[];
  const int Ti = p.Ti, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Wp = Ti + 2 * HALO;
  float* wcat = sm;                       // [2][Wp]
  float* f = wcat + 2 * Wp;               // [Ti][NF]
  float* e = f + Ti * NF;                 // [Ti]
  float* q = e + Ti;                      // [AD]
  float* wconv = q + AD;                  // [2][KS][NF]  (c fastest)
  float* wloc = wconv + 2 * KS * NF;      // [NF][AD]     (d fastest)
  float* vv = wloc + NF * AD;             // [AD]
  float* red = vv + AD;                   // [32]

  for (int i = tid; i < 2 * Wp; i += 256) wcat[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < Ti; i += 256) {
    wcat[HALO + i] = p.w_prev ? p.w_prev[b * p.wprev_rs + i] : 0.f;
    wcat[Wp + HALO + i] = p.cum_in[(long long)b * Ti + i];
  }
  for (int i = tid; i < AD; i += 256) {
    float a = 0.f;
    for (int s = 0; s < p.n_qparts; ++s) a += p.qparts[s * p.qpart_stride + (long long)b * AD + i];
    q[i] = a;
    vv[i] = p.v[i];
  }
  for (int i = tid; i < NF * 2 * KS; i += 256) {   // src [c][ch][k] -> dst [ch][k][c]
    const int k = i % KS, ch = (i / KS) % 2, c = i / (2 * KS);
    wconv[(ch * KS + k) * NF + c] = p.w_conv[i];
  }
  for (int i = tid; i < AD * NF; i += 256) {        // src [d][c] -> dst [c][d]
    const int c = i % NF, d = i / NF;
    wloc[c * AD + d] = p.w_loc[i];
  }
  __syncthreads();
  // location conv: f[ti][c]
  for (int i = tid; i < Ti * NF; i += 256) {
    const int c = i % NF, ti = i / NF;
    float a = 0.f;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* wrow = wcat + ch * Wp + ti;
      const float* wk = wconv + ch * KS * NF + c;
#pragma unroll
      for (int k = 0; k < KS; ++k) a = fmaf(wrow[k], wk[k * NF], a);
    }
    f[i] = a;
  }
  __syncthreads();
  // energies: one warp per text position, lanes over the 128 attention dims (4 each)
  const long long len = p.lens ? p.lens[b] : Ti;
  for (int ti = warp; ti < Ti; ti += 8) {
    const float* frow = f + ti * NF;
    const float* pm = p.pmem + ((long long)b * Ti + ti) * AD;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int d = lane + 32 * u;
      float s = q[d] + pm[d];
#pragma unroll
      for (int c = 0; c < NF; ++c) s = fmaf(frow[c], wloc[c * AD + d], s);
      const float a = tanhf(s);
      if (p.a_save) p.a_save[((long long)b * Ti + ti) * AD + d] = a;
      acc = fmaf(vv[d], a, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) e[ti] = (ti < len) ? acc : p.mask_value;
  }
  __syncthreads();
  // softmax over Ti
  float m = -INFINITY;
  for (int i = tid; i < Ti; i += 256) m = fmaxf(m, e[i]);
  m = block_max(m, red);
  float s = 0.f;
  for (int i = tid; i < Ti; i += 256) {
    const float x = expf(e[i] - m);
    e[i] = x;
    s += x;
  }
  s = block_sum(s, red);
  const float inv = 1.f / s;
  __syncthreads();
  for (int i = tid; i < Ti; i += 256) {
    const float w = e[i] * inv;
    e[i] = w;
    p.w_out[b * p.wout_rs + i] = w;
    p.cum_out[(long long)b * Ti + i] = wcat[Wp + HALO + i] + w;
  }
  __syncthreads();
  // context: 512 channels, 2 per thread
  float c0 = 0.f, c1 = 0.f;
  const float* mrow = p.mem + (long long)b * Ti * ED;
  for (int ti = 0; ti < Ti; ++ti) {
    const float w = e[ti];
    if (w != 0.f) {
      c0 = fmaf(w, mrow[(long long)ti * ED + tid], c0);
      c1 = fmaf(w, mrow[(long long)ti * ED + 256 + tid], c1);
    }
  }
  c0 = t2v_rnd(c0, p.rnd); c1 = t2v_rnd(c1, p.rnd);
  if (p.ctx_out1) { p.ctx_out1[b * p.ctx1_rs + tid] = c0; p.ctx_out1[b * p.ctx1_rs + 256 + tid] = c1; }
  if (p.ctx_out2) { p.ctx_out2[b * p.ctx2_rs + tid] = c0; p.ctx_out2[b * p.ctx2_rs + 256 + tid] = c1; }
}

struct AttnBwdArgs {
  const float* dctx1; long long dctx1_rs;   // grads wrt ctx_t (nullable sources, summed)
  const float* dctx2; long long dctx2_rs;
  const float* dctx3; long long dctx3_rs;
  const float* dw_in;       // [B,Ti] grad wrt w_t through step t+1's "previous weights" channel (nullable)
  float* dw_out;            // [B,Ti] grad wrt w_{t-1} through this step's "previous weights" channel
  float* gcum;              // [B,Ti] in: sum of grads wrt cum inputs of steps > t ; out: += this step's
  const float* w;  long long w_rs;          // w_t   [B,Ti]
  const float* w_prev; long long wprev_rs;  // w_{t-1} (nullable => zeros)
  const float* cum_in;      // cum_{t-1} [B,Ti]
  const float* a_save;      // [B,Ti,AD]
  const float* mem;         // [B,Ti,ED]
  const float* w_conv; const float* w_loc; const float* v;
  const long long* lens;
  float* dmem;              // [B,Ti,ED]  +=
  float* dpmem;             // [B,Ti,AD]  +=
  float* dq;                // [B,AD]     =
  float* dv_part;           // [B,AD]     +=
  float* dwloc_part;        // [B,AD,NF]  +=
  float* dwconv_part;       // [B,NF,2,KS] +=
  int B, Ti;
  int rnd;
};

__global__ void __launch_bounds__(256) attn_step_bwd_kernel(AttnBwdArgs p) {
  extern __shared__ float sm[];
  const int Ti = p.Ti, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Wp = Ti + 2 * HALO;
  float* wcat = sm;                       // [2][Wp]
  float* f = wcat + 2 * Wp;               // [Ti][NF]
  float* df = f + Ti * NF;                // [Ti + 2*HALO][NF] zero-haloed (for the adjoint conv)
  float* de = df + (Ti + 2 * HALO) * NF;  // [Ti]
  float* wt = de + Ti;                    // [Ti]  w_t
  float* dctx = wt + Ti;                  // [ED]
  float* wconv = dctx + ED;               // [2][KS][NF]
  float* wloc = wconv + 2 * KS * NF;      // [NF][AD]
  float* vv = wloc + NF * AD;             // [AD]
  float* dpre = vv + AD;                  // [32][AD] tile
  float* dqs = dpre + 32 * AD;            // [AD]
  float* dvs = dqs + AD;                  // [AD]
  float* red = dvs + AD;                  // [32]

  for (int i = tid; i < 2 * Wp; i += 256) wcat[i] = 0.f;
  for (int i = tid; i < (Ti + 2 * HALO) * NF; i += 256) df[i] = 0.f;
  for (int i = tid; i < AD; i += 256) { dqs[i] = 0.f; dvs[i] = 0.f; vv[i] = p.v[i]; }
  __syncthreads();
  for (int i = tid; i < Ti; i += 256) {
    wcat[HALO + i] = p.w_prev ? p.w_prev[b * p.wprev_rs + i] : 0.f;
    wcat[Wp + HALO + i] = p.cum_in[(long long)b * Ti + i];
    wt[i] = p.w[b * p.w_rs + i];
  }
  for (int i = tid; i < ED; i += 256) {
    float a = 0.f;
    if (p.dctx1) a += p.dctx1[b * p.dctx1_rs + i];
    if (p.dctx2) a += p.dctx2[b * p.dctx2_rs + i];
    if (p.dctx3) a += p.dctx3[b * p.dctx3_rs + i];
    dctx[i] = a;
  }
  for (int i = tid; i < NF * 2 * KS; i += 256) {
    const int k = i % KS, ch = (i / KS) % 2, c = i / (2 * KS);
    wconv[(ch * KS + k) * NF + c] = p.w_conv[i];
  }
  for (int i = tid; i < AD * NF; i += 256) {
    const int c = i % NF, d = i / NF;
    wloc[c * AD + d] = p.w_loc[i];
  }
  __syncthreads();
  const long long len = p.lens ? p.lens[b] : Ti;
  // (1) dw[ti] = <dctx, mem[ti]> + future grads ; dmem[ti] += w[ti]*dctx      (warp per ti)
  for (int ti = warp; ti < Ti; ti += 8) {
    float acc = 0.f;
    if (ti < len) {
      const float* mrow = p.mem + ((long long)b * Ti + ti) * ED;
      float* dmrow = p.dmem + ((long long)b * Ti + ti) * ED;
      const float w = wt[ti];
#pragma unroll 4
      for (int c = lane; c < ED; c += 32) {
        acc = fmaf(dctx[c], mrow[c], acc);
        dmrow[c] = fmaf(w, dctx[c], dmrow[c]);
      }
      acc = warp_sum(acc);
    }
    if (lane == 0) {
      float extra = p.gcum[(long long)b * Ti + ti];
      if (p.dw_in) extra += p.dw_in[(long long)b * Ti + ti];
      de[ti] = acc + extra;     // holds dw for now
    }
  }
  __syncthreads();
  // (2) softmax backward: de = w*(dw - sum_j w_j dw_j)
  float s = 0.f;
  for (int i = tid; i < Ti; i += 256) s += wt[i] * de[i];
  s = block_sum(s, red);
  __syncthreads();
  for (int i = tid; i < Ti; i += 256) de[i] = wt[i] * (de[i] - s);
  // (3) recompute the location features f[ti][c]
  for (int i = tid; i < Ti * NF; i += 256) {
    const int c = i % NF, ti = i / NF;
    float a = 0.f;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* wrow = wcat + ch * Wp + ti;
      const float* wk = wconv + ch * KS * NF + c;
#pragma unroll
      for (int k = 0; k < KS; ++k) a = fmaf(wrow[k], wk[k * NF], a);
    }
    f[i] = a;
  }
  __syncthreads();
  // (4) tiles of 32 text positions: dpre, dpmem, dq, dv, df, dWloc
  float dwl[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dwl[i] = 0.f;
  // thread owns dWloc entries (d = tid/2, c = (tid%2)*16 .. +15)
  const int own_d = tid >> 1, own_c0 = (tid & 1) * 16;
  float dq_loc[4] = {0.f, 0.f, 0.f, 0.f}, dv_loc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t0 = 0; t0 < Ti; t0 += 32) {
    const int nt = min(32, Ti - t0);
    for (int tt = warp; tt < 32; tt += 8) {
      const int ti = t0 + tt;
      if (tt < nt) {
        const float g = de[ti];
        const float* arow = p.a_save + ((long long)b * Ti + ti) * AD;
        float* dprow = p.dpmem + ((long long)b * Ti + ti) * AD;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int d = lane + 32 * u;
          const float a = arow[d];
          const float dp = g * vv[d] * (1.f - a * a);
          dpre[tt * AD + d] = dp;
          if (g != 0.f) dprow[d] += dp;
          dq_loc[u] += dp;
          dv_loc[u] = fmaf(g, a, dv_loc[u]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) dpre[tt * AD + lane + 32 * u] = 0.f;
      }
    }
    __syncthreads();
    // df[ti][c] = sum_d dpre[ti][d] * Wloc[d][c]
    for (int i = tid; i < 32 * NF; i += 256) {
      const int c = i % NF, tt = i / NF;
      if (tt < nt) {
        float a = 0.f;
        const float* dp = dpre + tt * AD;
#pragma unroll 8
        for (int d = 0; d < AD; ++d) a = fmaf(dp[d], wloc[c * AD + d], a);
        df[(HALO + t0 + tt) * NF + c] = a;
      }
    }
    // dWloc[d][c] += sum_ti dpre[ti][d] * f[ti][c]
    for (int tt = 0; tt < nt; ++tt) {
      const float dp = dpre[tt * AD + own_d];
      const float* frow = f + (t0 + tt) * NF + own_c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) dwl[j] = fmaf(dp, frow[j], dwl[j]);
    }
    __syncthreads();
  }
  // reduce dq / dv across the 8 warps (each lane holds d = lane+32u)
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    atomicAdd(&dqs[lane + 32 * u], dq_loc[u]);
    atomicAdd(&dvs[lane + 32 * u], dv_loc[u]);
  }
  {
    float* o = p.dwloc_part + ((long long)b * AD + own_d) * NF + own_c0;
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] += dwl[j];
  }
  __syncthreads();
  for (int i = tid; i < AD; i += 256) {
    p.dq[(long long)b * AD + i] = t2v_rnd(dqs[i], p.rnd);
    p.dv_part[(long long)b * AD + i] += dvs[i];
  }
  // (5) dWconv[c][ch][k] += sum_ti df[ti][c] * wcat[ch][ti+k]
  for (int i = tid; i < NF * 2 * KS; i += 256) {
    const int k = i % KS, ch = (i / KS) % 2, c = i / (2 * KS);
    float a = 0.f;
    const float* wrow = wcat + ch * Wp + k;
    for (int ti = 0; ti < Ti; ++ti) a = fmaf(df[(HALO + ti) * NF + c], wrow[ti], a);
    p.dwconv_part[(long long)b * NF * 2 * KS + i] += a;
  }
  // (6) adjoint conv: dwcat[ch][s] = sum_{c,k} df[s-k+HALO][c] * Wconv[c][ch][k]
  for (int i = tid; i < 2 * Ti; i += 256) {
    const int ch = i / Ti, sidx = i % Ti;
    float a = 0.f;
    for (int k = 0; k < KS; ++k) {
      const float* dfrow = df + (sidx - k + 2 * HALO) * NF;   // (s-k+HALO) + HALO halo offset
      const float* wk = wconv + (ch * KS + k) * NF;
#pragma unroll
      for (int c = 0; c < NF; ++c) a = fmaf(dfrow[c], wk[c], a);
    }
    if (ch == 0) p.dw_out[(long long)b * Ti + sidx] = a;
    else p.gcum[(long long)b * Ti + sidx] += a;
  }
}

size_t attn_fwd_smem(int Ti) {
  return sizeof(float) * (size_t)(2 * (Ti + 2 * HALO) + Ti * NF + Ti + AD + 2 * KS * NF + NF * AD + AD + 32);
}
size_t attn_bwd_smem(int Ti) {
  return sizeof(float) * (size_t)(2 * (Ti + 2 * HALO) + Ti * NF + (Ti + 2 * HALO) * NF + 2 * Ti + ED + 2 * KS * NF +
                                  NF * AD + AD + 32 * AD + 2 * AD + 32);
}

}  // namespace

T2V_API int t2v_attn_step_fwd(const float* qparts, int n_qparts, long long qpart_stride, const float* w_prev,
                              long long wprev_rs, const float* cum_in, float* cum_out, const float* pmem,
                              const float* mem, const float* w_conv, const float* w_loc, const float* v,
                              const long long* lens, float mask_value, float* w_out, long long wout_rs, float* ctx_out1,
                              long long ctx1_rs, float* ctx_out2, long long ctx2_rs, float* a_save, int B, int Ti,
                              int rnd, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0, "shape");
  const size_t smem = attn_fwd_smem(Ti);
  T2V_ARG_CHECK(smem <= 220 * 1024, "Ti too large for the fused attention kernel");
  static size_t cur = 0;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(attn_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  AttnFwdArgs a;
  a.qparts = qparts; a.n_qparts = n_qparts; a.qpart_stride = qpart_stride; a.w_prev = w_prev; a.wprev_rs = wprev_rs;
  a.cum_in = cum_in; a.cum_out = cum_out; a.pmem = pmem; a.mem = mem; a.w_conv = w_conv; a.w_loc = w_loc; a.v = v;
  a.lens = lens; a.mask_value = mask_value; a.w_out = w_out; a.wout_rs = wout_rs; a.ctx_out1 = ctx_out1;
  a.ctx1_rs = ctx1_rs; a.ctx_out2 = ctx_out2; a.ctx2_rs = ctx2_rs; a.a_save = a_save; a.B = B; a.Ti = Ti; a.rnd = rnd;
  attn_step_fwd_kernel<<<B, 256, smem, st>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}

T2V_API int t2v_attn_step_bwd(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs,
                              const float* dctx3, long long dctx3_rs, const float* dw_in, float* dw_out, float* gcum,
                              const float* w, long long w_rs, const float* w_prev, long long wprev_rs,
                              const float* cum_in, const float* a_save, const float* mem, const float* w_conv,
                              const float* w_loc, const float* v, const long long* lens, float* dmem, float* dpmem,
                              float* dq, float* dv_part, float* dwloc_part, float* dwconv_part, int B, int Ti,
                              int rnd, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0, "shape");
  const size_t smem = attn_bwd_smem(Ti);
  T2V_ARG_CHECK(smem <= 220 * 1024, "Ti too large for the fused attention backward kernel");
  static size_t cur = 0;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(attn_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  AttnBwdArgs a;
  a.dctx1 = dctx1; a.dctx1_rs = dctx1_rs; a.dctx2 = dctx2; a.dctx2_rs = dctx2_rs; a.dctx3 = dctx3; a.dctx3_rs = dctx3_rs;
  a.dw_in = dw_in; a.dw_out = dw_out; a.gcum = gcum; a.w = w; a.w_rs = w_rs; a.w_prev = w_prev; a.wprev_rs = wprev_rs;
  a.cum_in = cum_in; a.a_save = a_save; a.mem = mem; a.w_conv = w_conv; a.w_loc = w_loc; a.v = v; a.lens = lens;
  a.dmem = dmem; a.dpmem = dpmem; a.dq = dq; a.dv_part = dv_part; a.dwloc_part = dwloc_part;
  a.dwconv_part = dwconv_part; a.B = B; a.Ti = Ti; a.rnd = rnd;
  attn_step_bwd_kernel<<<B, 256, smem, st>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
