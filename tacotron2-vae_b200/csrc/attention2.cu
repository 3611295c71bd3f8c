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
//             attn2_bwd_dq_kernel   grid (text-chunk, b): softmax backward, tanh/v backward, d(processed memory) +=, dq
//             attn2_bwd_loc_kernel  grid (text-chunk, b): location dense/conv backward incl. the adjoint conv scattered
//                                   into the next step's "previous weights"/"cumulative weights" gradient buffers
//                                   (off the recurrence: runs on a side stream in the decoder loop).
// Weight-gradient partials (v, W_loc, W_conv) are accumulated per CTA slot across the time loop (no atomics) and
// reduced once after it.
#include "t2v_common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace {

constexpr int NF = 32, KS = 31, AD = 128, ED = 512, HALO = 15;
#ifndef T2V_ATTN_TC
#define T2V_ATTN_TC 32
#endif
constexpr int TC = T2V_ATTN_TC;        // text positions per CTA (16 was tried: twice the CTAs but the per-CTA weight staging dominates -> slower)
constexpr int WIN = TC + 2 * HALO;     // 62
constexpr int NCH = 4;                 // channel chunks of 128 for the context kernels
constexpr int TPB = TC / 4;             // text positions per thread in the conv stage (4 thread groups)
static_assert(TPB == 4 || TPB == 8, "TC must be 16 or 32");
constexpr int DPS = TC + 4;            // row stride of the transposed dpre tile (bank spread, keeps float4 alignment)

struct E2Args {
  const float* qparts; int n_qparts; long long qpart_stride;
  const float* w_prev; long long wprev_rs;     // nullable
  const float* cum_in;                          // [B,Ti]
  const float* pmem;                            // [B,Ti,AD]
  const float* w_conv; const float* w_loc; const float* v;   // w_conv: transposed [2*KS][NF]
  const long long* lens; float mask_value;
  float* e_out;                                 // [B,Ti]
  float* a_save;                                // [B,Ti,AD] nullable
  int B, Ti;
};

// f[ti][c] for ti in the CTA's chunk: shared by forward and backward (recompute)
__device__ __forceinline__ void conv_stage(const float* __restrict__ w_prev, long long wprev_rs, const float* __restrict__ cum_in,
                                           const float* __restrict__ w_convT, int b, int t0, int Ti, float* win /*[2][WIN]*/,
                                           float* wcT /*[2*KS][NF]*/, float* f /*[TC][NF+1]*/) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * WIN; i += 128) {
    const int ch = i / WIN, j = i % WIN;
    const int s = t0 - HALO + j;
    float v = 0.f;
    if (s >= 0 && s < Ti) v = (ch == 0) ? (w_prev ? w_prev[b * wprev_rs + s] : 0.f) : cum_in[(long long)b * Ti + s];
    win[i] = v;
  }
  {   // conv weights are pre-transposed by the host ([ch*KS+k][c]): coalesced loads, conflict-free stores
    constexpr int NL = (NF * 2 * KS + 127) / 128;
    float tmp[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) { const int i = tid + 128 * j; tmp[j] = (i < NF * 2 * KS) ? w_convT[i] : 0.f; }
#pragma unroll
    for (int j = 0; j < NL; ++j) { const int i = tid + 128 * j; if (i < NF * 2 * KS) wcT[i] = tmp[j]; }
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
  t2v_pdl_trigger();
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
  const float vd = p.v[d];
  t2v_pdl_wait();                                   // everything above is weights only
  float pmv[TC];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) pmv[tt] = (tt < nt) ? p.pmem[((long long)b * Ti + t0 + tt) * AD + d] : 0.f;
  float q = 0.f;
  for (int s = 0; s < p.n_qparts; ++s) q += p.qparts[s * p.qpart_stride + (long long)b * AD + d];
  conv_stage(p.w_prev, p.wprev_rs, p.cum_in, p.w_conv, b, t0, Ti, win, wcT, f);
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) {
    if (tt < nt) {
      const long long row = (long long)b * Ti + t0 + tt;
      float s = q + pmv[tt];
      const float* fr = f + tt * (NF + 1);
      float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < NF; c += 4) {
        s = fmaf(fr[c], wl[c], s); s1 = fmaf(fr[c + 1], wl[c + 1], s1);
        s2 = fmaf(fr[c + 2], wl[c + 2], s2); s3 = fmaf(fr[c + 3], wl[c + 3], s3);
      }
      s = (s + s1) + (s2 + s3);
      const float a = t2v_tanh(s);
      if (p.a_save) __stcs(p.a_save + row * AD + d, a);
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
  t2v_pdl_trigger();
  t2v_pdl_wait();
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

// ---- fused forward: one thread-block cluster per utterance (CS CTAs).  CTA r computes the energies of text chunk r
// into its shared memory; after a cluster barrier every CTA gathers all chunks through distributed shared memory,
// does the softmax and reduces the context for its 512/CS channel slice.  One launch per decoder step.
struct F2Args {
  E2Args e;
  const float* cum_in; float* cum_out; const float* mem;
  float* w_out; long long wout_rs;
  float* ctx_out1; long long ctx1_rs; float* ctx_out2; long long ctx2_rs;
  int rnd, tcx;              // tcx = text positions per CTA (<= TC)
};
__global__ void __launch_bounds__(128) attn2_fused_kernel(F2Args p) {
  __shared__ float win[2 * WIN];
  __shared__ float wcT[2 * KS * NF];
  __shared__ float f[TC * (NF + 1)];
  __shared__ float red[4][TC];
  __shared__ float e_loc[TC];
  __shared__ __align__(16) float wfull[8 * TC];
  __shared__ __align__(16) float part[4 * 128];
  __shared__ float red2[32];
  t2v_pdl_trigger();
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Ti = p.e.Ti;
  const int t0 = rank * p.tcx;
  const int nt = max(0, min(p.tcx, Ti - t0));
  {
    const E2Args& q = p.e;
    const int d = tid;
    float wl[NF];
#pragma unroll
    for (int c = 0; c < NF; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(q.w_loc + d * NF + c);
      wl[c] = t.x; wl[c + 1] = t.y; wl[c + 2] = t.z; wl[c + 3] = t.w;
    }
    const float vd = q.v[d];
    t2v_pdl_wait();                                 // everything above is weights only
    float pmv[TC];
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) pmv[tt] = (tt < nt) ? q.pmem[((long long)b * Ti + t0 + tt) * AD + d] : 0.f;
    float qv = 0.f;
    for (int s = 0; s < q.n_qparts; ++s) qv += q.qparts[s * q.qpart_stride + (long long)b * AD + d];
    conv_stage(q.w_prev, q.wprev_rs, q.cum_in, q.w_conv, b, t0, Ti, win, wcT, f);
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) {
      if (tt < nt) {
        const long long row = (long long)b * Ti + t0 + tt;
        float s = qv + pmv[tt];
        const float* fr = f + tt * (NF + 1);
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < NF; c += 4) {
          s = fmaf(fr[c], wl[c], s); s1 = fmaf(fr[c + 1], wl[c + 1], s1);
          s2 = fmaf(fr[c + 2], wl[c + 2], s2); s3 = fmaf(fr[c + 3], wl[c + 3], s3);
        }
        s = (s + s1) + (s2 + s3);
        const float a = t2v_tanh(s);
        if (q.a_save) __stcs(q.a_save + row * AD + d, a);
        const float pr = warp_sum(vd * a);
        if (lane == 0) red[warp][tt] = pr;
      }
    }
    __syncthreads();
    if (tid < TC) {
      float e = -INFINITY;
      if (tid < nt) {
        e = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
        const long long len = q.lens ? q.lens[b] : Ti;
        if (t0 + tid >= len) e = q.mask_value;
      }
      e_loc[tid] = e;
    }
  }
  cluster.sync();
  // gather the energies of every chunk (distributed shared memory) and softmax
  for (int i = tid; i < CS * TC; i += 128) {
    const int r = i / TC, j = i % TC;
    const float* remote = cluster.map_shared_rank(e_loc, r);
    wfull[i] = (j < p.tcx) ? remote[j] : -INFINITY;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int i = tid; i < CS * TC; i += 128) m = fmaxf(m, wfull[i]);
  m = block_max(m, red2);
  float ssum = 0.f;
  for (int i = tid; i < CS * TC; i += 128) {
    const float x = expf(wfull[i] - m);       // padding slots hold -inf -> 0
    wfull[i] = x;
    ssum += x;
  }
  ssum = block_sum(ssum, red2);
  const float inv = 1.f / ssum;
  __syncthreads();
  for (int i = tid; i < CS * TC; i += 128) {
    const int r = i / TC, j = i % TC, ti = r * p.tcx + j;
    const float x = wfull[i] * inv;
    wfull[i] = x;
    if (r == rank && j < nt) {               // each CTA publishes its own chunk
      p.w_out[b * p.wout_rs + ti] = x;
      p.cum_out[(long long)b * Ti + ti] = p.cum_in[(long long)b * Ti + ti] + x;
    }
  }
  __syncthreads();
  cluster.sync();                            // nobody may exit while its e_loc can still be read remotely
  // context for channels [rank*cpc, rank*cpc + cpc), cpc = 512/CS: warp -> text positions, lane -> 4 channels
  const int cpc = ED / CS;                   // 128 (CS=4) or 64 (CS=8)
  const int lanes_used = cpc / 4;            // 32 or 16
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes_used) {
    const float* mbase = p.mem + (long long)b * Ti * ED + rank * cpc + lane * 4;
    constexpr int NB = 16;
    for (int base = warp; base < Ti; base += 4 * NB) {
      float4 mv[NB];
      float wi[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int ti = base + 4 * j;
        wi[j] = (ti < Ti) ? wfull[(ti / p.tcx) * TC + (ti % p.tcx)] : 0.f;
        mv[j] = (wi[j] != 0.f) ? *reinterpret_cast<const float4*>(mbase + (long long)ti * ED) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        acc.x = fmaf(wi[j], mv[j].x, acc.x); acc.y = fmaf(wi[j], mv[j].y, acc.y);
        acc.z = fmaf(wi[j], mv[j].z, acc.z); acc.w = fmaf(wi[j], mv[j].w, acc.w);
      }
    }
  }
  *reinterpret_cast<float4*>(part + warp * 128 + lane * 4) = acc;
  __syncthreads();
  if (tid < cpc) {
    const float c = t2v_rnd(part[tid] + part[128 + tid] + part[256 + tid] + part[384 + tid], p.rnd);
    const int col = rank * cpc + tid;
    if (p.ctx_out1) p.ctx_out1[b * p.ctx1_rs + col] = c;
    if (p.ctx_out2) p.ctx_out2[b * p.ctx2_rs + col] = c;
  }
}

// ---- fused forward, one CTA per utterance (the north-star "attention energy + softmax + context reduction in one
// kernel"): 512 threads = 4 groups of 128; group g runs the chunk pipeline above on text chunk g (g+4, ... for long
// texts) concurrently, the energies meet in shared memory, all 16 warps do the softmax and the context reduction
// (warp -> text positions, lane -> 16 channels as 4 float4, 16 rows in flight per warp).
struct R3Args {
  E2Args e;
  const float* cum_in; float* cum_out; const float* mem;
  float* w_out; long long wout_rs;
  float* ctx_out1; long long ctx1_rs; float* ctx_out2; long long ctx2_rs;
  int rnd;
};
__global__ void __launch_bounds__(512) attn3_row_kernel(R3Args p) {
  extern __shared__ __align__(16) float sm4[];
  t2v_pdl_trigger();
  const E2Args& q = p.e;
  const int Ti = q.Ti, Tia = (Ti + 3) & ~3;
  float* wcT = sm4;                              // [2*KS][NF]
  float* winA = wcT + 2 * KS * NF;               // [4][2*WIN]
  float* fA = winA + 4 * 2 * WIN;                // [4][TC*(NF+1)]
  float* redA = fA + 4 * TC * (NF + 1);          // [4][4][TC]
  float* ew = redA + 4 * 4 * TC;                 // [Tia] energies -> weights
  float* part = ew + Tia;                        // [16][512]
  float* red2 = part + 16 * 512;                 // [32]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid >> 7, gt = tid & 127, gwarp = warp & 3;
  float* win = winA + grp * 2 * WIN;
  float* f = fA + grp * TC * (NF + 1);
  float* red = redA + grp * 4 * TC;
  // ---- weights (independent of the previous kernel)
  const int d = gt;
  float wl[NF];
#pragma unroll
  for (int c = 0; c < NF; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(q.w_loc + d * NF + c);
    wl[c] = t.x; wl[c + 1] = t.y; wl[c + 2] = t.z; wl[c + 3] = t.w;
  }
  const float vd = q.v[d];
  {
    constexpr int NL = (NF * 2 * KS + 511) / 512;
    float tmp[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) { const int i = tid + 512 * j; tmp[j] = (i < NF * 2 * KS) ? q.w_conv[i] : 0.f; }
#pragma unroll
    for (int j = 0; j < NL; ++j) { const int i = tid + 512 * j; if (i < NF * 2 * KS) wcT[i] = tmp[j]; }
  }
  t2v_pdl_wait();
  float qv = 0.f;
  for (int s = 0; s < q.n_qparts; ++s) qv += q.qparts[s * q.qpart_stride + (long long)b * AD + d];
  const long long len = q.lens ? q.lens[b] : Ti;
  const int nchunk = (Ti + TC - 1) / TC;
  for (int pass = 0; pass * 4 < nchunk; ++pass) {
    const int t0 = (pass * 4 + grp) * TC;
    const int nt = max(0, min(TC, Ti - t0));
    float pmv[TC];
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) pmv[tt] = (tt < nt) ? q.pmem[((long long)b * Ti + t0 + tt) * AD + d] : 0.f;
    for (int i = gt; i < 2 * WIN; i += 128) {
      const int ch = i / WIN, j = i % WIN;
      const int sidx = t0 - HALO + j;
      float v = 0.f;
      if (sidx >= 0 && sidx < Ti) v = (ch == 0) ? (q.w_prev ? q.w_prev[b * q.wprev_rs + sidx] : 0.f) : q.cum_in[(long long)b * Ti + sidx];
      win[i] = v;
    }
    __syncthreads();
    {
      const int c = gt & 31, g = gt >> 5;
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
    }
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) {
      if (tt < nt) {
        const long long row = (long long)b * Ti + t0 + tt;
        float s = qv + pmv[tt], s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const float* fr = f + tt * (NF + 1);
#pragma unroll
        for (int c = 0; c < NF; c += 4) {
          s = fmaf(fr[c], wl[c], s); s1 = fmaf(fr[c + 1], wl[c + 1], s1);
          s2 = fmaf(fr[c + 2], wl[c + 2], s2); s3 = fmaf(fr[c + 3], wl[c + 3], s3);
        }
        s = (s + s1) + (s2 + s3);
        const float a = t2v_tanh(s);
        if (q.a_save) __stcs(q.a_save + row * AD + d, a);
        const float pr = warp_sum(vd * a);
        if (lane == 0) red[gwarp * TC + tt] = pr;
      }
    }
    __syncthreads();
    if (gt < nt) {
      const float e = red[gt] + red[TC + gt] + red[2 * TC + gt] + red[3 * TC + gt];
      ew[t0 + gt] = (t0 + gt < len) ? e : q.mask_value;
    }
  }
  __syncthreads();
  // ---- softmax over the text positions
  float m = -INFINITY;
  for (int i = tid; i < Ti; i += 512) m = fmaxf(m, ew[i]);
  m = block_max(m, red2);
  float ssum = 0.f;
  for (int i = tid; i < Ti; i += 512) {
    const float x = expf(ew[i] - m);
    ew[i] = x;
    ssum += x;
  }
  ssum = block_sum(ssum, red2);
  const float inv = 1.f / ssum;
  __syncthreads();
  for (int i = tid; i < Ti; i += 512) {
    const float x = ew[i] * inv;
    ew[i] = x;
    p.w_out[b * p.wout_rs + i] = x;
    p.cum_out[(long long)b * Ti + i] = p.cum_in[(long long)b * Ti + i] + x;
  }
  __syncthreads();
  // ---- context: warp -> ti = warp, warp+16, ... ; lane -> channels lane*4 + 128*j
  float4 acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* mbase = p.mem + (long long)b * Ti * ED + lane * 4;
  for (int base = warp; base < Ti; base += 16 * 4) {
    float4 mv[4][4];
    float wi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ti = base + 16 * r;
      wi[r] = (ti < Ti) ? ew[ti] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mv[r][j] = (wi[r] != 0.f) ? *reinterpret_cast<const float4*>(mbase + (long long)ti * ED + 128 * j)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j].x = fmaf(wi[r], mv[r][j].x, acc[j].x); acc[j].y = fmaf(wi[r], mv[r][j].y, acc[j].y);
        acc[j].z = fmaf(wi[r], mv[r][j].z, acc[j].z); acc[j].w = fmaf(wi[r], mv[r][j].w, acc[j].w);
      }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(part + warp * 512 + 128 * j + lane * 4) = acc[j];
  __syncthreads();
  float c = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) c += part[w * 512 + tid];
  c = t2v_rnd(c, p.rnd);
  if (p.ctx_out1) p.ctx_out1[b * p.ctx1_rs + tid] = c;
  if (p.ctx_out2) p.ctx_out2[b * p.ctx2_rs + tid] = c;
}

// ---- forward split along the recurrence: the location term of step t+1 depends on the alignments of step t but NOT on
// the query of step t+1, so it is computed off the critical path (side stream, under the attention_rnn GEMM / cell / query
// GEMM of step t+1) by attn3_loc_kernel as  pre[b,ti,:] = processed_memory[b,ti,:] + W_loc conv([w_prev; w_cum])[ti] ;
// attn3_rowq_kernel, one CTA per utterance, then only does  tanh(q + pre) . v -> mask -> softmax -> context.
struct L3Args {
  const float* w_prev; long long wprev_rs;     // nullable (first step)
  const float* cum_in; const float* pmem; const float* w_conv; const float* w_loc;
  float* pre;                                   // [B,Ti,AD]
  int B, Ti;
};
__global__ void __launch_bounds__(128) attn3_loc_kernel(L3Args p) {
  t2v_pdl_trigger();
  __shared__ float win[2 * WIN];
  __shared__ float wcT[2 * KS * NF];
  __shared__ float f[TC * (NF + 1)];
  const int b = blockIdx.y, t0 = blockIdx.x * TC, tid = threadIdx.x;
  const int Ti = p.Ti, d = tid, nt = min(TC, Ti - t0);
  float wl[NF];
#pragma unroll
  for (int c = 0; c < NF; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p.w_loc + d * NF + c);
    wl[c] = t.x; wl[c + 1] = t.y; wl[c + 2] = t.z; wl[c + 3] = t.w;
  }
  float pmv[TC];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) pmv[tt] = (tt < nt) ? p.pmem[((long long)b * Ti + t0 + tt) * AD + d] : 0.f;
  t2v_pdl_wait();
  conv_stage(p.w_prev, p.wprev_rs, p.cum_in, p.w_conv, b, t0, Ti, win, wcT, f);
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) {
    if (tt < nt) {
      float s = pmv[tt], s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* fr = f + tt * (NF + 1);
#pragma unroll
      for (int c = 0; c < NF; c += 4) {
        s = fmaf(fr[c], wl[c], s); s1 = fmaf(fr[c + 1], wl[c + 1], s1);
        s2 = fmaf(fr[c + 2], wl[c + 2], s2); s3 = fmaf(fr[c + 3], wl[c + 3], s3);
      }
      p.pre[((long long)b * Ti + t0 + tt) * AD + d] = (s + s1) + (s2 + s3);
    }
  }
}

struct Q3Args {
  const float* qparts; int n_qparts; long long qpart_stride;
  const float* pre;                             // [B,Ti,AD]
  const float* v; const long long* lens; float mask_value;
  const float* cum_in; float* cum_out; const float* mem;
  float* w_out; long long wout_rs;
  float* ctx_out1; long long ctx1_rs; float* ctx_out2; long long ctx2_rs;
  float* a_save;                                // [B,Ti,AD] nullable
  int B, Ti, rnd;
};
__global__ void __launch_bounds__(512) attn3_rowq_kernel(Q3Args p) {
  extern __shared__ __align__(16) float sm5[];
  t2v_pdl_trigger();
  const int Ti = p.Ti, Tia = (Ti + 3) & ~3;
  float* redA = sm5;                             // [4 groups][4 warps][TC]
  float* ew = redA + 4 * 4 * TC;                 // [Tia] energies -> weights
  float* part = ew + Tia;                        // [16][512]
  float* red2 = part + 16 * 512;                 // [32]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid >> 7, d = tid & 127, gwarp = warp & 3;
  float* red = redA + grp * 4 * TC;
  const float vd = p.v[d];
  const long long len = p.lens ? p.lens[b] : Ti;
  t2v_pdl_wait();
  float qv = 0.f;
  for (int s = 0; s < p.n_qparts; ++s) qv += p.qparts[s * p.qpart_stride + (long long)b * AD + d];
  const int nchunk = (Ti + TC - 1) / TC;
  for (int pass = 0; pass * 4 < nchunk; ++pass) {
    const int t0 = (pass * 4 + grp) * TC;
    const int nt = max(0, min(TC, Ti - t0));
    float x[TC];
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) x[tt] = (tt < nt) ? __ldcs(p.pre + ((long long)b * Ti + t0 + tt) * AD + d) : 0.f;
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) {
      const float a = t2v_tanh(qv + x[tt]);
      if (tt < nt && p.a_save) __stcs(p.a_save + ((long long)b * Ti + t0 + tt) * AD + d, a);
      x[tt] = (tt < nt) ? vd * a : 0.f;
    }
    // transposing warp reduction: 31 shuffles leave the sum over the warp's 32 attention dims of position tt in lane tt
#pragma unroll
    for (int sft = TC / 2; sft >= 1; sft >>= 1) {
      const bool up = (lane & sft) != 0;
#pragma unroll
      for (int i = 0; i < sft; ++i) {
        const float send = up ? x[i] : x[i + sft];
        const float keep = up ? x[i + sft] : x[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
      }
    }
    red[gwarp * TC + lane] = x[0];
    __syncthreads();
    if (d < nt) {
      const float e = red[d] + red[TC + d] + red[2 * TC + d] + red[3 * TC + d];
      ew[t0 + d] = (t0 + d < len) ? e : p.mask_value;
    }
    __syncthreads();
  }
  // ---- softmax over the text positions
  float m = -INFINITY;
  for (int i = tid; i < Ti; i += 512) m = fmaxf(m, ew[i]);
  m = block_max(m, red2);
  float ssum = 0.f;
  for (int i = tid; i < Ti; i += 512) {
    const float xx = expf(ew[i] - m);
    ew[i] = xx;
    ssum += xx;
  }
  ssum = block_sum(ssum, red2);
  const float inv = 1.f / ssum;
  __syncthreads();
  for (int i = tid; i < Ti; i += 512) {
    const float xx = ew[i] * inv;
    ew[i] = xx;
    p.w_out[b * p.wout_rs + i] = xx;
    p.cum_out[(long long)b * Ti + i] = p.cum_in[(long long)b * Ti + i] + xx;
  }
  __syncthreads();
  // ---- context: warp -> ti = warp, warp+16, ... ; lane -> channels lane*4 + 128*j
  float4 acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* mbase = p.mem + (long long)b * Ti * ED + lane * 4;
  for (int base = warp; base < Ti; base += 16 * 4) {
    float4 mv[4][4];
    float wi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ti = base + 16 * r;
      wi[r] = (ti < Ti) ? ew[ti] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mv[r][j] = (wi[r] != 0.f) ? *reinterpret_cast<const float4*>(mbase + (long long)ti * ED + 128 * j)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j].x = fmaf(wi[r], mv[r][j].x, acc[j].x); acc[j].y = fmaf(wi[r], mv[r][j].y, acc[j].y);
        acc[j].z = fmaf(wi[r], mv[r][j].z, acc[j].z); acc[j].w = fmaf(wi[r], mv[r][j].w, acc[j].w);
      }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(part + warp * 512 + 128 * j + lane * 4) = acc[j];
  __syncthreads();
  float c = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) c += part[w * 512 + tid];
  c = t2v_rnd(c, p.rnd);
  if (p.ctx_out1) p.ctx_out1[b * p.ctx1_rs + tid] = c;
  if (p.ctx_out2) p.ctx_out2[b * p.ctx2_rs + tid] = c;
}

// ------------------------------------------------------------------------------------------------ backward
struct B1Args {
  const float* dctx1; long long dctx1_rs; const float* dctx2; long long dctx2_rs; const float* dctx3; long long dctx3_rs;
  float* dctx_out;           // [B,ED] total gradient wrt ctx_t (kept for the batched d(memory) GEMM)
  const float* mem;          // [B,Ti,ED]
  const long long* lens;
  float* dw_part;            // [NCH][B][Ti] partial <dctx, mem[ti]>
  int B, Ti;
};
__global__ void __launch_bounds__(128) attn2_bwd_ctx_kernel(B1Args p) {
  t2v_pdl_trigger();
  t2v_pdl_wait();
  __shared__ __align__(16) float dsh[128];
  const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Ti = p.Ti;
  const int col = ch * 128 + tid;
  float d = 0.f;
  if (p.dctx1) d += p.dctx1[b * p.dctx1_rs + col];
  if (p.dctx2) d += p.dctx2[b * p.dctx2_rs + col];
  if (p.dctx3) d += p.dctx3[b * p.dctx3_rs + col];
  dsh[tid] = d;
  p.dctx_out[(long long)b * ED + col] = d;
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
  float* gcum_next;          // [B,Ti] (= gcum_prev + this step's cumulative-channel gradient)
  float* dw_out;             // [B,Ti] (this step's previous-weights-channel gradient)
  float* de_buf;             // [B,Ti] softmax-backward energies gradient, handed from the dq kernel to the loc kernel
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

// The energy backward is two kernels because only part of it sits on the recurrence of the backward time loop:
//   attn2_bwd_dq_kernel   softmax backward, tanh / v backward -> dq (feeds the attention_rnn cell backward of the SAME
//                         step: critical path), d(processed memory), dv.  Also initialises this step's scatter targets.
//   attn2_bwd_loc_kernel  location dense / conv backward: dW_loc, dW_conv and the adjoint conv scattered into the
//                         "previous weights" / "cumulative weights" gradients that only the NEXT step's dq kernel reads;
//                         the decoder loop runs it on a side stream under the rest of the step.
__global__ void __launch_bounds__(128) attn2_bwd_dq_kernel(B2Args p) {
  t2v_pdl_trigger();
  extern __shared__ __align__(16) float sm3[];
  const int Ti = p.Ti;
  float* dwv = sm3;                         // [Ti]
  float* de = dwv + ((Ti + 3) & ~3);        // [TC]
  float* red = de + TC;                     // [32]
  const int b = blockIdx.y, chunk = blockIdx.x, t0 = chunk * TC, tid = threadIdx.x;
  const int nchunk = gridDim.x;
  const int nt = min(TC, Ti - t0);
  // saved tanh activations of this chunk (thread = attention dim): forward data, loaded before the dependency wait
  float av[TC];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) av[tt] = (tt < nt) ? __ldcs(p.a_save + ((long long)b * Ti + t0 + tt) * AD + tid) : 0.f;
  const int d = tid;
  const float vd = p.v[d];
  const long long slot = (long long)b * nchunk + chunk;
  t2v_pdl_wait();
  // (1) dw over the whole row, s = <w, dw>, de for this chunk
  float part = 0.f;
  for (int i = tid; i < Ti; i += 128) {
    const float gc = p.gcum_prev[(long long)b * Ti + i];
    float dw = gc;
    if (p.dw_in) dw += p.dw_in[(long long)b * Ti + i];
#pragma unroll
    for (int c = 0; c < NCH; ++c) dw += p.dw_part[((long long)c * p.B + b) * Ti + i];
    dwv[i] = dw;
    part = fmaf(p.w[b * p.w_rs + i], dw, part);
    if (chunk == 0) {                        // scatter targets of this step's loc kernel
      p.gcum_next[(long long)b * Ti + i] = gc;
      p.dw_out[(long long)b * Ti + i] = 0.f;
    }
  }
  const float s = block_sum(part, red);
  __syncthreads();
  if (tid < TC) {
    const float g = (tid < nt) ? p.w[b * p.w_rs + t0 + tid] * (dwv[t0 + tid] - s) : 0.f;
    de[tid] = g;
    if (tid < nt) p.de_buf[(long long)b * Ti + t0 + tid] = g;
  }
  __syncthreads();
  // (2) thread d: tanh/v backward over the chunk
  float dq_acc = 0.f, dv_acc = p.dv_part[slot * AD + d];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) {
    if (tt < nt) {
      const float g = de[tt];
      const float a = av[tt];
      const float dp = g * vd * (1.f - a * a);
      if (g != 0.f) atomicAdd(p.dpmem + ((long long)b * Ti + t0 + tt) * AD + d, dp);   // sole writer: compiles to RED, no round trip
      dq_acc += dp;
      dv_acc = fmaf(g, a, dv_acc);
    }
  }
  p.dv_part[slot * AD + d] = dv_acc;
  atomicAdd(p.dq + (long long)b * AD + d, dq_acc);
}

__global__ void __launch_bounds__(128) attn2_bwd_loc_kernel(B2Args p) {
  t2v_pdl_trigger();
  extern __shared__ __align__(16) float sm3[];
  const int Ti = p.Ti;
  float* win = sm3;                         // [2][WIN]
  float* wcT = win + 2 * WIN;               // [2*KS][NF]
  float* f = wcT + 2 * KS * NF;             // [TC][NF+1]
  float* dpT = f + TC * (NF + 1);           // [AD][DPS] (dpre transposed: [d][ti], row stride DPS)
  float* wlT = dpT + AD * DPS;              // [AD][NF+1]... W_loc [d][c] padded
  float* df = wlT + AD * (NF + 1);          // [TC][NF+1]
  float* de = df + TC * (NF + 1);           // [TC]
  const int b = blockIdx.y, chunk = blockIdx.x, t0 = chunk * TC, tid = threadIdx.x;
  const int nchunk = gridDim.x;
  const int nt = min(TC, Ti - t0);
  // ---- everything up to the wait depends only on tensors saved by the FORWARD pass (alignments, cumulative weights,
  // tanh activations) and on weights
  float av[TC];
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) av[tt] = (tt < nt) ? __ldcs(p.a_save + ((long long)b * Ti + t0 + tt) * AD + tid) : 0.f;
  conv_stage(p.w_prev, p.wprev_rs, p.cum_in, p.w_conv, b, t0, Ti, win, wcT, f);      // recompute the location features
  {
    float tmp[AD * NF / 128];
#pragma unroll
    for (int j = 0; j < AD * NF / 128; ++j) tmp[j] = p.w_loc[tid + 128 * j];
#pragma unroll
    for (int j = 0; j < AD * NF / 128; ++j) { const int i = tid + 128 * j; wlT[(i / NF) * (NF + 1) + (i % NF)] = tmp[j]; }
  }
  const int d = tid;
  const float vd = p.v[d];
  const long long slot = (long long)b * nchunk + chunk;
  float* wl_part = p.dwloc_part + slot * (AD * NF) + d * NF;
  float gwl[NF];
#pragma unroll
  for (int c = 0; c < NF; c += 4) {           // start from the running partial (only this kernel's earlier steps wrote it)
    const float4 t = *reinterpret_cast<const float4*>(wl_part + c);
    gwl[c] = t.x; gwl[c + 1] = t.y; gwl[c + 2] = t.z; gwl[c + 3] = t.w;
  }
  t2v_pdl_wait();
  if (tid < TC) de[tid] = (tid < nt) ? p.de_buf[(long long)b * Ti + t0 + tid] : 0.f;
  __syncthreads();
  // (3) dpre (recomputed from de and the saved activations), dWloc row d in registers
#pragma unroll
  for (int tt = 0; tt < TC; ++tt) {
    float dp = 0.f;
    if (tt < nt) {
      const float a = av[tt];
      dp = de[tt] * vd * (1.f - a * a);
      const float* fr = f + tt * (NF + 1);
#pragma unroll
      for (int c = 0; c < NF; ++c) gwl[c] = fmaf(dp, fr[c], gwl[c]);
    }
    dpT[d * DPS + tt] = dp;
  }
#pragma unroll
  for (int c = 0; c < NF; c += 4) *reinterpret_cast<float4*>(wl_part + c) = make_float4(gwl[c], gwl[c + 1], gwl[c + 2], gwl[c + 3]);
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
      acc[0] = fmaf(w, x0.x, acc[0]); acc[1] = fmaf(w, x0.y, acc[1]); acc[2] = fmaf(w, x0.z, acc[2]); acc[3] = fmaf(w, x0.w, acc[3]);
      if (TPB == 8) {
        const float4 x1 = *reinterpret_cast<const float4*>(dpT + dd * DPS + g * TPB + 4);
        acc[TPB - 4] = fmaf(w, x1.x, acc[TPB - 4]); acc[TPB - 3] = fmaf(w, x1.y, acc[TPB - 3]);
        acc[TPB - 2] = fmaf(w, x1.z, acc[TPB - 2]); acc[TPB - 1] = fmaf(w, x1.w, acc[TPB - 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < TPB; ++j) df[(g * TPB + j) * (NF + 1) + c] = acc[j];
  }
  __syncthreads();
  // (5) dWconv[ch][k][c] += sum_ti df[ti][c] * win[ch][ti+k]   (partials kept in the transposed [ch*KS+k][c] layout)
  {
    constexpr int NL = (NF * 2 * KS + 127) / 128;
    float* part = p.dwconv_part + slot * (NF * 2 * KS);
    float old[NL], acc[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) { const int i = tid + 128 * j; old[j] = (i < NF * 2 * KS) ? part[i] : 0.f; }
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int i = tid + 128 * j;
      float a = 0.f;
      if (i < NF * 2 * KS) {
        const int c = i % NF, chk = i / NF;                 // chk = ch*KS + k
        const float* x = win + (chk / KS) * WIN + (chk % KS);
        float b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
        for (int tt = 0; tt < TC; tt += 4) {
          a = fmaf(df[tt * (NF + 1) + c], x[tt], a); b1 = fmaf(df[(tt + 1) * (NF + 1) + c], x[tt + 1], b1);
          b2 = fmaf(df[(tt + 2) * (NF + 1) + c], x[tt + 2], b2); b3 = fmaf(df[(tt + 3) * (NF + 1) + c], x[tt + 3], b3);
        }
        a = (a + b1) + (b2 + b3);
      }
      acc[j] = a;
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) { const int i = tid + 128 * j; if (i < NF * 2 * KS) part[i] = old[j] + acc[j]; }
  }
  // (6) adjoint conv, scattered: contribution of this chunk's df to dwcat[ch][t0-HALO+j], j in [0,WIN)
  for (int i = tid; i < 2 * WIN; i += 128) {
    const int ch = i / WIN, j = i % WIN;
    // s = t0 - HALO + j ; ti = s - k + HALO  =>  tt = j - k, 0 <= tt < TC   (k uniform across the warp: broadcast reads);
    // four independent accumulators break the 992-long dependent FMA chain
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
    for (int k = 0; k < KS; ++k) {
      const int tt = j - k;
      if (tt >= 0 && tt < TC) {
        const float* dfr = df + tt * (NF + 1);
        const float* wk = wcT + (ch * KS + k) * NF;
#pragma unroll
        for (int c = 0; c < NF; c += 4) {
          a0 = fmaf(dfr[c], wk[c], a0); a1 = fmaf(dfr[c + 1], wk[c + 1], a1);
          a2 = fmaf(dfr[c + 2], wk[c + 2], a2); a3 = fmaf(dfr[c + 3], wk[c + 3], a3);
        }
      }
    }
    const float a = (a0 + a1) + (a2 + a3);
    const int sidx = t0 - HALO + j;
    if (sidx >= 0 && sidx < Ti && a != 0.f) {
      if (ch == 0) atomicAdd(p.dw_out + (long long)b * Ti + sidx, a);
      else atomicAdd(p.gcum_next + (long long)b * Ti + sidx, a);
    }
  }
}

size_t bwd_dq_smem(int Ti) { return sizeof(float) * (size_t)(((Ti + 3) & ~3) + TC + 32); }
size_t bwd_loc_smem() {
  return sizeof(float) * (size_t)(2 * WIN + 2 * KS * NF + TC * (NF + 1) + AD * DPS + AD * (NF + 1) + TC * (NF + 1) + TC);
}

}  // namespace

#define LAUNCH_END() do { T2V_COUNT_LAUNCH(); T2V_LAUNCH_CHECK(); return 0; } while (0)

T2V_API int t2v_attn2_chunks(int Ti) { return (Ti + TC - 1) / TC; }

// forward in two launches (see attn3_loc_kernel): _loc for step t+1 may run on a side stream as soon as step t's
// alignments exist; _row is the part on the recurrence
T2V_API int t2v_attn3_loc_fwd(const float* w_prev, long long wprev_rs, const float* cum_in, const float* pmem,
                              const float* w_conv, const float* w_loc, float* pre, int B, int Ti, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  L3Args a;
  a.w_prev = w_prev; a.wprev_rs = wprev_rs; a.cum_in = cum_in; a.pmem = pmem; a.w_conv = w_conv; a.w_loc = w_loc; a.pre = pre;
  a.B = B; a.Ti = Ti;
  T2V_CUDA_CHECK(t2v_launch(attn3_loc_kernel, dim3((Ti + TC - 1) / TC, B), dim3(128), 0, st, true, 1, a));
  LAUNCH_END();
}
T2V_API int t2v_attn3_row_fwd(const float* qparts, int n_qparts, long long qpart_stride, const float* pre, const float* cum_in,
                              float* cum_out, const float* mem, const float* v, const long long* lens, float mask_value,
                              float* w_out, long long wout_rs, float* ctx_out1, long long ctx1_rs, float* ctx_out2,
                              long long ctx2_rs, float* a_save, int B, int Ti, int rnd, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  Q3Args a;
  a.qparts = qparts; a.n_qparts = n_qparts; a.qpart_stride = qpart_stride; a.pre = pre; a.v = v; a.lens = lens;
  a.mask_value = mask_value; a.cum_in = cum_in; a.cum_out = cum_out; a.mem = mem; a.w_out = w_out; a.wout_rs = wout_rs;
  a.ctx_out1 = ctx_out1; a.ctx1_rs = ctx1_rs; a.ctx_out2 = ctx_out2; a.ctx2_rs = ctx2_rs; a.a_save = a_save; a.B = B;
  a.Ti = Ti; a.rnd = rnd;
  const size_t smem = sizeof(float) * (size_t)(4 * 4 * TC + ((Ti + 3) & ~3) + 16 * 512 + 32);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(attn3_rowq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(t2v_launch(attn3_rowq_kernel, dim3(B), dim3(512), smem, st, true, 1, a));
  LAUNCH_END();
}

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
  // modes: "row" (default) one CTA per utterance, everything in one launch; "split" energy + context kernels;
  // "cluster" one 4/8-CTA cluster per utterance (measured slower than "split" on the C3 shape: 28 vs 17 us)
  static const char* mode_env = getenv("T2V_ATTN_MODE");
  static const int mode = (mode_env && mode_env[0] == 's') ? 1 : ((mode_env && mode_env[0] == 'c') ? 2 : 0);
  const bool use_cluster = (mode == 2);
  if (mode == 0) {
    R3Args ra;
    ra.e = e; ra.cum_in = cum_in; ra.cum_out = cum_out; ra.mem = mem; ra.w_out = w_out; ra.wout_rs = wout_rs;
    ra.ctx_out1 = ctx_out1; ra.ctx1_rs = ctx1_rs; ra.ctx_out2 = ctx_out2; ra.ctx2_rs = ctx2_rs; ra.rnd = rnd;
    const size_t smem = sizeof(float) * (size_t)(2 * KS * NF + 4 * 2 * WIN + 4 * TC * (NF + 1) + 4 * 4 * TC + ((Ti + 3) & ~3) +
                                                 16 * 512 + 32);
    static size_t cur = 48 * 1024;
    if (smem > cur) {
      T2V_CUDA_CHECK(cudaFuncSetAttribute(attn3_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cur = smem;
    }
    T2V_CUDA_CHECK(t2v_launch(attn3_row_kernel, dim3(B), dim3(512), smem, st, true, 1, ra));
    T2V_COUNT_LAUNCH();
    return 0;
  }
  if (Ti <= 8 * TC && use_cluster) {          // one cluster of 4 or 8 CTAs per utterance
    const int CS = (Ti <= 4 * TC) ? 4 : 8;
    F2Args fa;
    fa.e = e; fa.cum_in = cum_in; fa.cum_out = cum_out; fa.mem = mem; fa.w_out = w_out; fa.wout_rs = wout_rs;
    fa.ctx_out1 = ctx_out1; fa.ctx1_rs = ctx1_rs; fa.ctx_out2 = ctx_out2; fa.ctx2_rs = ctx2_rs; fa.rnd = rnd;
    fa.tcx = (Ti + CS - 1) / CS;
    T2V_CUDA_CHECK(t2v_launch(attn2_fused_kernel, dim3(CS, B, 1), dim3(128), 0, st, true, CS, fa));
    T2V_COUNT_LAUNCH();
    return 0;
  }
  dim3 g1((Ti + TC - 1) / TC, B);
  T2V_CUDA_CHECK(t2v_launch(attn2_energy_kernel, g1, dim3(128), 0, st, true, 1, e));
  T2V_COUNT_LAUNCH();
  C2Args c;
  c.e = e_buf; c.cum_in = cum_in; c.cum_out = cum_out; c.mem = mem; c.w_out = w_out; c.wout_rs = wout_rs;
  c.ctx_out1 = ctx_out1; c.ctx1_rs = ctx1_rs; c.ctx_out2 = ctx_out2; c.ctx2_rs = ctx2_rs; c.B = B; c.Ti = Ti; c.rnd = rnd;
  dim3 g2(NCH, B);
  const size_t smem = sizeof(float) * (size_t)(((Ti + 3) & ~3) + 4 * 128 + 32);
  T2V_CUDA_CHECK(t2v_launch(attn2_context_kernel, g2, dim3(128), smem, st, true, 1, c));
  T2V_COUNT_LAUNCH();
  return 0;
}

// backward, three launches.  _ctx and _dq belong to the recurrence of the time loop; _loc only has to finish before the
// NEXT step's _dq (it produces that step's dw_in / gcum_prev), so the decoder loop puts it on a side stream.
T2V_API int t2v_attn2_bwd_ctx(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs,
                              const float* dctx3, long long dctx3_rs, float* dctx_out, float* dw_part, const float* mem,
                              const long long* lens, int B, int Ti, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  B1Args a;
  a.dctx1 = dctx1; a.dctx1_rs = dctx1_rs; a.dctx2 = dctx2; a.dctx2_rs = dctx2_rs; a.dctx3 = dctx3; a.dctx3_rs = dctx3_rs;
  a.dctx_out = dctx_out; a.mem = mem; a.lens = lens; a.dw_part = dw_part; a.B = B; a.Ti = Ti;
  T2V_CUDA_CHECK(t2v_launch(attn2_bwd_ctx_kernel, dim3(NCH, B), dim3(128), 0, st, true, 1, a));
  LAUNCH_END();
}

static B2Args make_b2(const float* dw_in, float* dw_out, const float* gcum_prev, float* gcum_next, const float* dw_part,
                      float* de_buf, const float* w, long long w_rs, const float* w_prev, long long wprev_rs,
                      const float* cum_in, const float* a_save, const float* w_conv, const float* w_loc, const float* v,
                      float* dpmem, float* dq, float* dv_part, float* dwloc_part, float* dwconv_part, int B, int Ti) {
  B2Args e;
  e.dw_part = dw_part; e.dw_in = dw_in; e.gcum_prev = gcum_prev; e.gcum_next = gcum_next; e.dw_out = dw_out;
  e.de_buf = de_buf; e.w = w; e.w_rs = w_rs; e.w_prev = w_prev; e.wprev_rs = wprev_rs; e.cum_in = cum_in; e.a_save = a_save;
  e.w_conv = w_conv; e.w_loc = w_loc; e.v = v; e.dpmem = dpmem; e.dq = dq; e.dv_part = dv_part; e.dwloc_part = dwloc_part;
  e.dwconv_part = dwconv_part; e.B = B; e.Ti = Ti;
  return e;
}

T2V_API int t2v_attn2_bwd_dq(const float* dw_in, float* dw_out, const float* gcum_prev, float* gcum_next,
                             const float* dw_part, float* de_buf, const float* w, long long w_rs, const float* a_save,
                             const float* v, float* dpmem, float* dq, float* dv_part, int B, int Ti, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  T2V_ARG_CHECK(dw_out && gcum_next && de_buf, "scatter targets");
  B2Args e = make_b2(dw_in, dw_out, gcum_prev, gcum_next, dw_part, de_buf, w, w_rs, nullptr, 0, nullptr, a_save, nullptr,
                     nullptr, v, dpmem, dq, dv_part, nullptr, nullptr, B, Ti);
  const size_t smem = bwd_dq_smem(Ti);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(attn2_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(t2v_launch(attn2_bwd_dq_kernel, dim3((Ti + TC - 1) / TC, B), dim3(128), smem, st, true, 1, e));
  LAUNCH_END();
}

T2V_API int t2v_attn2_bwd_loc(float* dw_out, float* gcum_next, const float* de_buf, const float* w_prev, long long wprev_rs,
                              const float* cum_in, const float* a_save, const float* w_conv, const float* w_loc,
                              const float* v, float* dwloc_part, float* dwconv_part, int B, int Ti, cudaStream_t st) {
  T2V_ARG_CHECK(B > 0 && Ti > 0 && Ti <= 8192, "shape");
  B2Args e = make_b2(nullptr, dw_out, nullptr, gcum_next, nullptr, const_cast<float*>(de_buf), nullptr, 0, w_prev, wprev_rs,
                     cum_in, a_save, w_conv, w_loc, v, nullptr, nullptr, nullptr, dwloc_part, dwconv_part, B, Ti);
  const size_t smem = bwd_loc_smem();
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(attn2_bwd_loc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(t2v_launch(attn2_bwd_loc_kernel, dim3((Ti + TC - 1) / TC, B), dim3(128), smem, st, true, 1, e));
  LAUNCH_END();
}

// the three launches in stream order (what the unit tests and a single-stream caller use)
T2V_API int t2v_attn2_bwd(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs, const float* dctx3,
                          long long dctx3_rs, float* dctx_out, const float* dw_in, float* dw_out, const float* gcum_prev,
                          float* gcum_next, float* dw_part, float* de_buf, const float* w, long long w_rs,
                          const float* w_prev, long long wprev_rs, const float* cum_in, const float* a_save, const float* mem,
                          const float* w_conv, const float* w_loc, const float* v, const long long* lens, float* dpmem,
                          float* dq, float* dv_part, float* dwloc_part, float* dwconv_part, int B, int Ti, cudaStream_t st) {
  int r = t2v_attn2_bwd_ctx(dctx1, dctx1_rs, dctx2, dctx2_rs, dctx3, dctx3_rs, dctx_out, dw_part, mem, lens, B, Ti, st);
  if (r) return r;
  r = t2v_attn2_bwd_dq(dw_in, dw_out, gcum_prev, gcum_next, dw_part, de_buf, w, w_rs, a_save, v, dpmem, dq, dv_part, B, Ti, st);
  if (r) return r;
  return t2v_attn2_bwd_loc(dw_out, gcum_next, de_buf, w_prev, wprev_rs, cum_in, a_save, w_conv, w_loc, v, dwloc_part,
                           dwconv_part, B, Ti, st);
}
