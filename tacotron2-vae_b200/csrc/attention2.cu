// Location-sensitive attention (reference model.py:12-88, called from Decoder.decode model.py:366-374), version 2:
// the per-step work is spread over all SMs instead of one CTA per utterance.
//
//   forward   attn2_energy_kernel   grid (text-chunk, b): location conv (2->32, k31) on a register-resident sliding
//                                   window, location dense (32->128) with W_loc rows in registers, + query +
//                                   processed memory, tanh, v-projection -> masked energies e[b,ti]
//             attn2_context_kernel  grid (channel-chunk, b): softmax over Ti (recomputed per CTA, it is tiny), context
//                                   reduction over Ti with all row loads in flight, cumulative-weights update
//   backward  attn2_bwd_ctx_kernel  grid (channel-chunk, b): dctx_t = sum of its three sources (stored for the batched
//                                   d(memory) GEMM after the loop) and the partial <dctx, memory[ti]> products
//             attn2_bwd_energy_kernel grid (text-chunk, b): softmax backward, tanh/v backward, d(processed memory) +=,
//                                   dq, location dense/conv backward incl. the adjoint conv scattered into the
//                                   next step's "previous weights"/"cumulative weights" gradient buffers.
// Weight-gradient partials (v, W_loc, W_conv) are accumulated per CTA slot across the time loop (no atomics) and
// reduced once after it.
#include "t2v_common.cuh"

namespace {

constexpr int NF = 32, KS = 31, AD = 128, ED = 512, HALO = 15;
constexpr int TC = 32;                 // text positions per CTA
constexpr int WIN = TC + 2 * HALO;     // 62
constexpr int NCH = 4;                 // channel chunks of 128 for the context kernels
constexpr int TPB = 8;                 // text positions per thread in the conv stage (TC / 4 thread groups)
constexpr int DPS = TC + 4;            // row stride of the transposed dpre tile (bank spread, keeps float4 alignment)

struct E2Args {
  const float* qparts; int n_qparts; long long qpart_stride;
  const float* w_prev; long long wprev_rs;     // nullable
  const float* cum_in;                          // [B,Ti]
  const float* pmem;                            // [B,Ti,AD]
  const float* w_conv; const float* w_loc; const float* v;
  const long long* lens; float mask_value;
  float* e_out;                                 // [B,Ti]
  float* a_save;                                // [B,Ti,AD] nullable
  int B, Ti;
};

// f[ti][c] for ti in the CTA's chunk: shared by forward and backward (recompute)
__device__ __forceinline__ void conv_stage(const float* __restrict__ w_prev, long long wprev_rs, const float* __restrict__ cum_in,
                                           const float* __restrict__ w_conv, int b, int t0, int Ti, float* win /*[2][WIN]*/,
                                           float* wcT /*[2*KS][NF]*/, float* f /*[TC][NF+1]*/) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * WIN; i += 128) {
    const int ch = i / WIN, j = i % WIN;
    const int s = t0 - HALO + j;
    float v = 0.f;
    if (s >= 0 && s < Ti) v = (ch == 0) ? (w_prev ? w_prev[b * wprev_rs + s] : 0.f) : cum_in[(long long)b * Ti + s];
    win[i] = v;
  }
  for (int i = tid; i < NF * 2 * KS; i += 128) {          // src [c][ch][k] -> dst [ch*KS+k][c]
    const int k = i % KS, ch = (i / KS) % 2, c = i / (2 * KS);
    wcT[(ch * KS + k) * NF + c] = w_conv[i];
  }
  __syncthreads();
  const int c = tid & 31, g = tid >> 5;
  float acc[TPB];
#pragma unroll
  for (int j = 0; j < TPB; ++j) acc[j] = 0.f;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    float xr[TPB + KS - 1];
#pragma unroll
    for (int j = 0; j < TPB + KS - 1; ++j) xr[j] = win[ch * WIN + g * TPB + j];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const float w = wcT[(ch * KS + k) * NF + c];
#pragma unroll
      for (int j = 0; j < TPB; ++j) acc[j] = fmaf(w, xr[j + k], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < TPB; ++j) f[(g * TPB + j) * (NF + 1) + c] = acc[j];
  __syncthreads();
}

__global__ void __launch_bounds__(128) attn2_energy_kernel(E2Args p) {
  __shared__ float win[2 * WIN];
  __shared__ float wcT[2 * KS * NF];
  __shared__ float f[TC * (NF + 1)];
  __shared__ float red[4][TC];
  const int b = blockIdx.y, t0 = blockIdx.x * TC, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Ti = p.Ti;
  // thread d: W_loc row, query and the chunk's processed-memory values in registers -- every global load of the kernel
  // is in flight before the first use
  const int d = tid;
  const int nt = min(TC, Ti - t0);
  float wl[NF];
#pragma unroll
  for (int c = 0; c < NF; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p.w_loc + d * NF + c);
    wl[c] = t.x; wl[c + 1] = t.y; wl[c + 2] = t.z; wl[c + 3] = t.w;
  }
  float pmv[TC];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) pmv[tt] = (tt < nt) ? p.pmem[((long long)b * Ti + t0 + tt) * AD + d] : 0.f;
  float q = 0.f;
  for (int s = 0; s < p.n_qparts; ++s) q += p.qparts[s * p.qpart_stride + (long long)b * AD + d];
  const float vd = p.v[d];
  conv_stage(p.w_prev, p.wprev_rs, p.cum_in, p.w_conv, b, t0, Ti, win, wcT, f);
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) {
    if (tt < nt) {
      const long long row = (long long)b * Ti + t0 + tt;
      float s = q + pmv[tt];
      const float* fr = f + tt * (NF + 1);
#pragma unroll
      for (int c = 0; c < NF; ++c) s = fmaf(fr[c], wl[c], s);
      const float a = tanhf(s);
      if (p.a_save) p.a_save[row * AD + d] = a;
      const float part = warp_sum(vd * a);
      if (lane == 0) red[warp][tt] = part;
    }
  }
  __syncthreads();
  if (tid < nt) {
    const float e = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    const long long len = p.lens ? p.lens[b] : Ti;
    p.e_out[(long long)b * Ti + t0 + tid] = (t0 + tid < len) ? e : p.mask_value;
  }
}

struct C2Args {
  const float* e;            // [B,Ti]
  const float* cum_in; float* cum_out;
  const float* mem;          // [B,Ti,ED]
  float* w_out; long long wout_rs;
  float* ctx_out1; long long ctx1_rs; float* ctx_out2; long long ctx2_rs;
  int B, Ti, rnd;
};
__global__ void __launch_bounds__(128) attn2_context_kernel(C2Args p) {
  extern __shared__ __align__(16) float sm2[];
  float* w = sm2;                 // [Ti]
  float* part = w + ((p.Ti + 3) & ~3);   // [4][128] (16-byte aligned)
  float* red = part + 4 * 128;    // [32]
  const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Ti = p.Ti;
  float m = -INFINITY;
  for (int i = tid; i < Ti; i += 128) {
    const float x = p.e[(long long)b * Ti + i];
    w[i] = x;
    m = fmaxf(m, x);
  }
  m = block_max(m, red);
  float s = 0.f;
  for (int i = tid; i < Ti; i += 128) {
    const float x = expf(w[i] - m);
    w[i] = x;
    s += x;
  }
  s = block_sum(s, red);
  const float inv = 1.f / s;
  __syncthreads();
  for (int i = tid; i < Ti; i += 128) {
    const float x = w[i] * inv;
    w[i] = x;
    if (ch == 0) {
      p.w_out[b * p.wout_rs + i] = x;
      p.cum_out[(long long)b * Ti + i] = p.cum_in[(long long)b * Ti + i] + x;
    }
  }
  __syncthreads();
  // warp handles ti = warp, warp+4, ...; lane handles 4 channels
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* mbase = p.mem + (long long)b * Ti * ED + ch * 128 + lane * 4;
  constexpr int NB = 16;                       // memory rows in flight per lane
  for (int base = warp; base < Ti; base += 4 * NB) {
    float4 mv[NB];
    float wi[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int ti = base + 4 * j;
      wi[j] = (ti < Ti) ? w[ti] : 0.f;
      mv[j] = (wi[j] != 0.f) ? *reinterpret_cast<const float4*>(mbase + (long long)ti * ED) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      acc.x = fmaf(wi[j], mv[j].x, acc.x); acc.y = fmaf(wi[j], mv[j].y, acc.y);
      acc.z = fmaf(wi[j], mv[j].z, acc.z); acc.w = fmaf(wi[j], mv[j].w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(part + warp * 128 + lane * 4) = acc;
  __syncthreads();
  const float c = t2v_rnd(part[tid] + part[128 + tid] + part[256 + tid] + part[384 + tid], p.rnd);
  const int col = ch * 128 + tid;
  if (p.ctx_out1) p.ctx_out1[b * p.ctx1_rs + col] = c;
  if (p.ctx_out2) p.ctx_out2[b * p.ctx2_rs + col] = c;
}

// ------------------------------------------------------------------------------------------------ backward
struct B1Args {
  const float* dctx1; long long dctx1_rs; const float* dctx2; long long dctx2_rs; const float* dctx3; long long dctx3_rs;
  float* dctx_out;           // [B,ED] total gradient wrt ctx_t (kept for the batched d(memory) GEMM)
  const float* mem;          // [B,Ti,ED]
  const long long* lens;
  float* dw_part;            // [NCH][B][Ti] partial <dctx, mem[ti]>
  float* dw_next_zero;       // [B,Ti] buffer to clear for this step's scatter target (nullable)
  const float* gcum_prev; float* gcum_next;   // gcum_next = gcum_prev (copied here; the energy kernel adds into it)
  int B, Ti;
};
__global__ void __launch_bounds__(128) attn2_bwd_ctx_kernel(B1Args p) {
  __shared__ __align__(16) float dsh[128];
  const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Ti = p.Ti;
  const int col = ch * 128 + tid;
  float d = 0.f;
  if (p.dctx1) d += p.dctx1[b * p.dctx1_rs + col];
  if (p.dctx2) d += p.dctx2[b * p.dctx2_rs + col];
  if (p.dctx3) d += p.dctx3[b * p.dctx3_rs + col];
  dsh[tid] = d;
  p.dctx_out[(long long)b * ED + col] = d;
  if (ch == 0) {
    for (int i = tid; i < Ti; i += 128) {
      if (p.dw_next_zero) p.dw_next_zero[(long long)b * Ti + i] = 0.f;
      p.gcum_next[(long long)b * Ti + i] = p.gcum_prev[(long long)b * Ti + i];
    }
  }
  __syncthreads();
  const float4 dv = *reinterpret_cast<const float4*>(dsh + lane * 4);
  const long long len = p.lens ? p.lens[b] : Ti;
  const float* mbase = p.mem + (long long)b * Ti * ED + ch * 128 + lane * 4;
  float* out = p.dw_part + ((long long)ch * p.B + b) * Ti;
  constexpr int NB = 16;
  for (int base = warp; base < Ti; base += 4 * NB) {
    float4 mv[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int ti = base + 4 * j;
      mv[j] = (ti < len) ? *reinterpret_cast<const float4*>(mbase + (long long)ti * ED) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int ti = base + 4 * j;
      float acc = dv.x * mv[j].x + dv.y * mv[j].y + dv.z * mv[j].z + dv.w * mv[j].w;
      acc = warp_sum(acc);
      if (lane == 0 && ti < Ti) out[ti] = acc;
    }
  }
}

struct B2Args {
  const float* dw_part;      // [NCH][B][Ti]
  const float* dw_in;        // [B,Ti] nullable
  const float* gcum_prev;    // [B,Ti]
  float* gcum_next;          // [B,Ti] (+= this step's cumulative-channel gradient)
  float* dw_out;             // [B,Ti] (+= this step's previous-weights-channel gradient; pre-zeroed)
  const float* w; long long w_rs;
  const float* w_prev; long long wprev_rs;
  const float* cum_in;
  const float* a_save;       // [B,Ti,AD]
  const float* w_conv; const float* w_loc; const float* v;
  float* dpmem;              // [B,Ti,AD] +=
  float* dq;                 // [B,AD] atomicAdd (pre-zeroed)
  float* dv_part;            // [B*nchunk,AD] +=
  float* dwloc_part;         // [B*nchunk,AD*NF] +=
  float* dwconv_part;        // [B*nchunk,NF*2*KS] +=
  int B, Ti;
};
__global__ void __launch_bounds__(128) attn2_bwd_energy_kernel(B2Args p) {
  extern __shared__ __align__(16) float sm3[];
  const int Ti = p.Ti;
  float* dwv = sm3;                         // [Ti] dw then scratch
  float* win = dwv + ((Ti + 3) & ~3);       // [2][WIN]
  float* wcT = win + 2 * WIN;               // [2*KS][NF]
  float* f = wcT + 2 * KS * NF;             // [TC][NF+1]
  float* dpT = f + TC * (NF + 1);           // [AD][DPS] (dpre transposed: [d][ti], row stride DPS)
  float* wlT = dpT + AD * DPS;               // [AD][NF+1]... W_loc [d][c] padded
  float* df = wlT + AD * (NF + 1);          // [TC][NF+1]
  float* de = df + TC * (NF + 1);           // [TC]
  float* scat = de + TC;                    // [2][WIN]
  float* red = scat + 2 * WIN;              // [32]
  const int b = blockIdx.y, chunk = blockIdx.x, t0 = chunk * TC, tid = threadIdx.x;
  const int nchunk = gridDim.x;
  const int nt = min(TC, Ti - t0);
  float av[TC];                              // saved tanh activations of this chunk (thread = attention dim)
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) av[tt] = (tt < nt) ? p.a_save[((long long)b * Ti + t0 + tt) * AD + tid] : 0.f;
  // (1) dw over the whole row, s = <w, dw>, de for this chunk
  float part = 0.f;
  for (int i = tid; i < Ti; i += 128) {
    float dw = p.gcum_prev[(long long)b * Ti + i];
    if (p.dw_in) dw += p.dw_in[(long long)b * Ti + i];
#pragma unroll
    for (int c = 0; c < NCH; ++c) dw += p.dw_part[((long long)c * p.B + b) * Ti + i];
    dwv[i] = dw;
    part = fmaf(p.w[b * p.w_rs + i], dw, part);
  }
  const float s = block_sum(part, red);
  __syncthreads();
  if (tid < TC) de[tid] = (tid < nt) ? p.w[b * p.w_rs + t0 + tid] * (dwv[t0 + tid] - s) : 0.f;
  // (2) recompute the location features of this chunk
  conv_stage(p.w_prev, p.wprev_rs, p.cum_in, p.w_conv, b, t0, Ti, win, wcT, f);
  for (int i = tid; i < AD * NF; i += 128) wlT[(i / NF) * (NF + 1) + (i % NF)] = p.w_loc[i];
  for (int i = tid; i < 2 * WIN; i += 128) scat[i] = 0.f;
  // (3) thread d: tanh/v backward over the chunk; dWloc row d in registers
  const int d = tid;
  const float vd = p.v[d];
  const long long slot = (long long)b * nchunk + chunk;
  float* wl_part = p.dwloc_part + slot * (AD * NF) + d * NF;
  float gwl[NF];
#pragma unroll
  for (int c = 0; c < NF; c += 4) {           // start from the running partial: loads issued before the compute
    const float4 t = *reinterpret_cast<const float4*>(wl_part + c);
    gwl[c] = t.x; gwl[c + 1] = t.y; gwl[c + 2] = t.z; gwl[c + 3] = t.w;
  }
  float dq_acc = 0.f, dv_acc = p.dv_part[slot * AD + d];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) {
    float dp = 0.f;
    if (tt < nt) {
      const float g = de[tt];
      const float a = av[tt];
      dp = g * vd * (1.f - a * a);
      if (g != 0.f) atomicAdd(p.dpmem + ((long long)b * Ti + t0 + tt) * AD + d, dp);   // sole writer: compiles to RED, no round trip
      dq_acc += dp;
      dv_acc = fmaf(g, a, dv_acc);
      const float* fr = f + tt * (NF + 1);
#pragma unroll
      for (int c = 0; c < NF; ++c) gwl[c] = fmaf(dp, fr[c], gwl[c]);
    }
    dpT[d * DPS + tt] = dp;
  }
#pragma unroll
  for (int c = 0; c < NF; c += 4) *reinterpret_cast<float4*>(wl_part + c) = make_float4(gwl[c], gwl[c + 1], gwl[c + 2], gwl[c + 3]);
  p.dv_part[slot * AD + d] = dv_acc;
  atomicAdd(p.dq + (long long)b * AD + d, dq_acc);
  __syncthreads();
  // (4) df[ti][c] = sum_d dpre[ti][d] * Wloc[d][c]; thread (c, group of 8 ti)
  {
    const int c = tid & 31, g = tid >> 5;
    float acc[TPB];
#pragma unroll
    for (int j = 0; j < TPB; ++j) acc[j] = 0.f;
    for (int dd = 0; dd < AD; ++dd) {
      const float w = wlT[dd * (NF + 1) + c];
      const float4 x0 = *reinterpret_cast<const float4*>(dpT + dd * DPS + g * TPB);
      const float4 x1 = *reinterpret_cast<const float4*>(dpT + dd * DPS + g * TPB + 4);
      acc[0] = fmaf(w, x0.x, acc[0]); acc[1] = fmaf(w, x0.y, acc[1]); acc[2] = fmaf(w, x0.z, acc[2]); acc[3] = fmaf(w, x0.w, acc[3]);
      acc[4] = fmaf(w, x1.x, acc[4]); acc[5] = fmaf(w, x1.y, acc[5]); acc[6] = fmaf(w, x1.z, acc[6]); acc[7] = fmaf(w, x1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < TPB; ++j) df[(g * TPB + j) * (NF + 1) + c] = acc[j];
  }
  __syncthreads();
  // (5) dWconv[c][ch][k] += sum_ti df[ti][c] * win[ch][ti+k]
  {
    for (int i = tid; i < NF * 2 * KS; i += 128) {
      const int k = i % KS, ch = (i / KS) % 2, c = i / (2 * KS);
      float a = 0.f;
      const float* x = win + ch * WIN + k;
#pragma unroll 8
      for (int tt = 0; tt < TC; ++tt) a = fmaf(df[tt * (NF + 1) + c], x[tt], a);
      p.dwconv_part[slot * (NF * 2 * KS) + i] += a;
    }
  }
  // (6) adjoint conv, scattered: contribution of this chunk's df to dwcat[ch][t0-HALO+j], j in [0,WIN)
  for (int i = tid; i < 2 * WIN; i += 128) {
    const int ch = i / WIN, j = i % WIN;
    float a = 0.f;
    // s = t0 - HALO + j ; ti = s - k + HALO  =>  tt = j - k, 0 <= tt < TC
    const int k_lo = max(0, j - (TC - 1)), k_hi = min(KS - 1, j);
    for (int k = k_lo; k <= k_hi; ++k) {
      const float* dfr = df + (j - k) * (NF + 1);
      const float* wk = wcT + (ch * KS + k) * NF;
#pragma unroll 8
      for (int c = 0; c < NF; ++c) a = fmaf(dfr[c], wk[c], a);
    }
    const int sidx = t0 - HALO + j;
    if (sidx >= 0 && sidx < Ti && a != 0.f) {
      if (ch == 0) atomicAdd(p.dw_out + (long long)b * Ti + sidx, a);
      else atomicAdd(p.gcum_next + (long long)b * Ti + sidx, a);
    }
  }
}

size_t bwd_energy_smem(int Ti) {
  return sizeof(float) * (size_t)(((Ti + 3) & ~3) + 2 * WIN + 2 * KS * NF + TC * (NF + 1) + AD * DPS + AD * (NF + 1) + TC * (NF + 1) + TC +
                                  2 * WIN + 32);
}

}  // namespace

#define LAUNCH_END() do { T2V_COUNT_LAUNCH(); T2V_LAUNCH_CHECK(); return 0; } while (0)

T2V_API int t2v_attn2_chunks(int Ti) { return (Ti + TC - 1) / TC; }

// forward: energies (e_buf [B,Ti] scratch) then softmax + context
T2V_API int t2v_attn2_fwd(const float* qparts, int n_qparts, long long qpart_stride, const float* w_prev, long long wprev_rs,
                          const float* cum_in, float* cum_out, const float* pmem, const float* mem, const float* w_conv,
                          const float* w_loc, const float* v, const long long* lens, float mask_value, float* e_buf,
                          float* w_out, long long wout_rs, float* ctx_out1, long long ctx1_rs, float* ctx_out2,
                          long long ctx2_rs, float* a_save, int B, int Ti, int rnd, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  E2Args e;
  e.qparts = qparts; e.n_qparts = n_qparts; e.qpart_stride = qpart_stride; e.w_prev = w_prev; e.wprev_rs = wprev_rs;
  e.cum_in = cum_in; e.pmem = pmem; e.w_conv = w_conv; e.w_loc = w_loc; e.v = v; e.lens = lens; e.mask_value = mask_value;
  e.e_out = e_buf; e.a_save = a_save; e.B = B; e.Ti = Ti;
  dim3 g1((Ti + TC - 1) / TC, B);
  attn2_energy_kernel<<<g1, 128, 0, st>>>(e);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  C2Args c;
  c.e = e_buf; c.cum_in = cum_in; c.cum_out = cum_out; c.mem = mem; c.w_out = w_out; c.wout_rs = wout_rs;
  c.ctx_out1 = ctx_out1; c.ctx1_rs = ctx1_rs; c.ctx_out2 = ctx_out2; c.ctx2_rs = ctx2_rs; c.B = B; c.Ti = Ti; c.rnd = rnd;
  dim3 g2(NCH, B);
  const size_t smem = sizeof(float) * (size_t)(((Ti + 3) & ~3) + 4 * 128 + 32);
  attn2_context_kernel<<<g2, 128, smem, st>>>(c);
  LAUNCH_END();
}

T2V_API int t2v_attn2_bwd(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs, const float* dctx3,
                          long long dctx3_rs, float* dctx_out, const float* dw_in, float* dw_out, const float* gcum_prev,
                          float* gcum_next, float* dw_part, const float* w, long long w_rs, const float* w_prev,
                          long long wprev_rs, const float* cum_in, const float* a_save, const float* mem, const float* w_conv,
                          const float* w_loc, const float* v, const long long* lens, float* dpmem, float* dq, float* dv_part,
                          float* dwloc_part, float* dwconv_part, int B, int Ti, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  B1Args a;
  a.dctx1 = dctx1; a.dctx1_rs = dctx1_rs; a.dctx2 = dctx2; a.dctx2_rs = dctx2_rs; a.dctx3 = dctx3; a.dctx3_rs = dctx3_rs;
  a.dctx_out = dctx_out; a.mem = mem; a.lens = lens; a.dw_part = dw_part; a.dw_next_zero = dw_out; a.gcum_prev = gcum_prev;
  a.gcum_next = gcum_next; a.B = B; a.Ti = Ti;
  attn2_bwd_ctx_kernel<<<dim3(NCH, B), 128, 0, st>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  B2Args e;
  e.dw_part = dw_part; e.dw_in = dw_in; e.gcum_prev = gcum_prev; e.gcum_next = gcum_next; e.dw_out = dw_out; e.w = w;
  e.w_rs = w_rs; e.w_prev = w_prev; e.wprev_rs = wprev_rs; e.cum_in = cum_in; e.a_save = a_save; e.w_conv = w_conv;
  e.w_loc = w_loc; e.v = v; e.dpmem = dpmem; e.dq = dq; e.dv_part = dv_part; e.dwloc_part = dwloc_part;
  e.dwconv_part = dwconv_part; e.B = B; e.Ti = Ti;
  const size_t smem = bwd_energy_smem(Ti);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(attn2_bwd_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  attn2_bwd_energy_kernel<<<dim3((Ti + TC - 1) / TC, B), 128, smem, st>>>(e);
  LAUNCH_END();
}
