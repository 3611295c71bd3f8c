// Encoder BiLSTM (reference model.py:171-190: nn.LSTM, bidirectional, packed by input_lengths) as ONE persistent kernel per pass:
// 64 CTAs (32 per direction, 8 hidden units each, warp = unit) stay resident for all Ti steps.  The recurrent weights of a CTA's
// units are staged into shared memory once (the per-step launches of rnn.cu re-staged them 120 times), the cell / cell-gradient
// state lives in registers, and the only per-step exchange is the new h (forward) or the new gate gradients (backward) through
// L2, ordered by one monotonic device-wide counter per direction (red.release / ld.acquire, bounded waits).
// Same arithmetic as bilstm_step_fwd / bilstm_step_bwd (exact fp32 FFMA): the two paths are compared bit for bit in the tests.
#include "t2v_common.cuh"
#include <stdlib.h>

namespace {

constexpr int UPC = 8;          // hidden units (= warps) per CTA
constexpr int MT = 64;          // batch rows (lane -> rows lane, lane + 32)
#ifndef T2V_WAIT_LIMIT
#define T2V_WAIT_LIMIT 4000000000LL
#endif
constexpr long long WAIT_LIMIT = T2V_WAIT_LIMIT;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
  if (ld_acquire_u32(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_u32(p) < target) {
    if (clock64() - t0 > WAIT_LIMIT) __trap();
  }
}
__device__ __forceinline__ void signal_counter(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

struct SeqFwd {
  const float* gx[2];            // [B*Tp, 4H] hoisted input projections (padded rows: row b*Tp + 2 + t)
  const float* w_hh[2];          // [4H, H]
  const float* b_hh[2];          // [4H]
  float* seq;                    // [B*Tp, 2H] layer output (direction d writes columns d*H ..)
  float* gates;                  // [2, Ti, B, 4H] saved gate activations
  float* cells;                  // [2, Ti + 2, B, H] saved cell states, slot t + 1 = cell after time t
  float* hbuf;                   // [2 dirs][2][B][H] h exchange (ping-pong)
  unsigned* counters;            // [2][32] one line per direction, zero-initialised
  const long long* lens;         // [B] or nullptr (Encoder.inference: no packing)
  int B, H, Ti, Tp;
};

__global__ void __launch_bounds__(UPC * 32, 1) bilstm_seq_fwd_kernel(SeqFwd p) {
  extern __shared__ __align__(16) float smf[];
  const int H = p.H, dir = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UPC, u = u0 + warp;
  const int HS = H + 1;
  const int ncta = gridDim.x;
  float* hT = smf;                       // [MT][H+1] h_prev tile
  float* ws = hT + ((MT * HS + 3) & ~3); // [H][UPC][4] recurrent weights of this CTA's units, k-major
  const float* W = p.w_hh[dir];
  for (int i = threadIdx.x; i < H * UPC * 4; i += UPC * 32) {
    const int k = i % H, g = (i / H) % 4, uu = i / (4 * H);
    ws[(k * UPC + uu) * 4 + g] = W[((long long)g * H + u0 + uu) * H + k];
  }
  const float* bh = p.b_hh[dir];
  const float b_i = bh[u], b_f = bh[H + u], b_g = bh[2 * H + u], b_o = bh[3 * H + u];
  unsigned* cnt = p.counters + 32 * dir;
  float cst[2] = {0.f, 0.f};
  long long len[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int b = r * 32 + lane;
    len[r] = (b < p.B) ? (p.lens ? p.lens[b] : (long long)p.Ti) : 0;
  }
  for (int s = 0; s < p.Ti; ++s) {
    const int t = dir ? p.Ti - 1 - s : s;
    const float* hprev = p.hbuf + ((long long)(dir * 2 + (s & 1)) * p.B) * H;
    float* hnext = p.hbuf + ((long long)(dir * 2 + ((s + 1) & 1)) * p.B) * H;
    // gate pre-activations of this step: independent of the recurrence, in flight while the counter is polled
    float gxv[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b < p.B) {
        const float* gx = p.gx[dir] + ((long long)b * p.Tp + 2 + t) * 4 * H;
        gxv[r][0] = gx[u]; gxv[r][1] = gx[H + u]; gxv[r][2] = gx[2 * H + u]; gxv[r][3] = gx[3 * H + u];
      } else {
        gxv[r][0] = gxv[r][1] = gxv[r][2] = gxv[r][3] = 0.f;
      }
    }
    if (s > 0) {
      if (threadIdx.x == 0) wait_counter(cnt, (unsigned)(ncta * s));
      __syncthreads();
      const int nv = MT * H / 4;
      for (int i0 = threadIdx.x; i0 < nv; i0 += UPC * 32 * 16) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = i0 + j * UPC * 32;
          const int m = (i * 4) / H, k = (i * 4) % H;
          v[j] = (i < nv && m < p.B) ? __ldcg(reinterpret_cast<const float4*>(hprev + (long long)m * H + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    } else {
      for (int i = threadIdx.x; i < MT * HS; i += UPC * 32) hT[i] = 0.f;
    }
    __syncthreads();
    float acc[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
    if (s > 0) {
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float a0 = hT[lane * HS + k], a1 = hT[(32 + lane) * HS + k];
        const float4 w = *reinterpret_cast<const float4*>(ws + (k * UPC + warp) * 4);
        acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
        acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
        acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
        acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b >= p.B) continue;
      float* gs = p.gates + (((long long)dir * p.Ti + t) * p.B + b) * 4 * H + u;
      float* cs = p.cells + (((long long)dir * (p.Ti + 2) + t + 1) * p.B + b) * H + u;
      float* so = p.seq + ((long long)b * p.Tp + 2 + t) * 2 * H + dir * H + u;
      if (t >= len[r]) {      // dead step of a packed row: state untouched, zero output (pad_packed_sequence)
        *so = 0.f;
        gs[0] = 0.f; gs[H] = 0.f; gs[2 * H] = 0.f; gs[3 * H] = 0.f;
        *cs = 0.f;
        hnext[(long long)b * H + u] = hT[b * HS + u];
        continue;
      }
      const float ig = t2v_sigmoid(acc[r][0] + gxv[r][0] + b_i);
      const float fg = t2v_sigmoid(acc[r][1] + gxv[r][1] + b_f);
      const float gg = tanhf(acc[r][2] + gxv[r][2] + b_g);
      const float og = t2v_sigmoid(acc[r][3] + gxv[r][3] + b_o);
      const float c2 = fg * cst[r] + ig * gg;
      const float h2 = og * tanhf(c2);
      cst[r] = c2;
      hnext[(long long)b * H + u] = h2;
      *so = h2;
      gs[0] = ig; gs[H] = fg; gs[2 * H] = gg; gs[3 * H] = og;
      *cs = c2;
    }
    __syncthreads();
    if (threadIdx.x == 0) signal_counter(cnt);
  }
}

struct SeqBwd {
  const float* w_hhT[2];         // [H, 4H] transposed recurrent weights
  const float* dout;             // [B, Ti, 2H] gradient wrt the layer output
  const float* gates;            // [2, Ti, B, 4H]
  const float* cells;            // [2, Ti + 2, B, H]
  float* dg[2];                  // [B*Tp, 4H] gate gradients per direction (padded rows: row b*Tp + 2 + t)
  unsigned* counters;
  const long long* lens;
  int B, H, Ti, Tp;
};

__global__ void __launch_bounds__(UPC * 32, 1) bilstm_seq_bwd_kernel(SeqBwd p) {
  extern __shared__ __align__(16) float smb[];
  const int H = p.H, K = 4 * H, dir = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UPC, u = u0 + warp;
  const int ncta = gridDim.x;
  constexpr int KC = 256, DS = KC + 1;
  float* dT = smb;                       // [MT][KC+1] gate-gradient chunk of the previously processed step
  float* ws = dT + MT * DS;              // [K][UPC] this CTA's columns of W_hh^T, staged once
  float* red = ws + (size_t)K * UPC;     // [8 warps][UPC][MT] partial sums
  for (int i = threadIdx.x; i < K * UPC; i += UPC * 32) {
    const int k = i % K, uu = i / K;
    ws[k * UPC + uu] = p.w_hhT[dir][(long long)(u0 + uu) * K + k];
  }
  unsigned* cnt = p.counters + 32 * dir;
  float dcs[2] = {0.f, 0.f};
  long long len[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int b = r * 32 + lane;
    len[r] = (b < p.B) ? (p.lens ? p.lens[b] : (long long)p.Ti) : 0;
  }
  __syncthreads();
  for (int s = 0; s < p.Ti; ++s) {
    const int t = dir ? s : p.Ti - 1 - s;             // each direction walks its own forward order backwards
    const int tn = dir ? t - 1 : t + 1;               // the step processed in the previous iteration
    const int tp = dir ? t + 1 : t - 1;               // the step whose cell state entered step t (slot tp + 1)
    // saved activations of this step: independent of the recurrence
    float sg[2][4], sc2[2], scp[2], dov[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b < p.B) {
        const float* gs = p.gates + (((long long)dir * p.Ti + t) * p.B + b) * 4 * H + u;
        sg[r][0] = gs[0]; sg[r][1] = gs[H]; sg[r][2] = gs[2 * H]; sg[r][3] = gs[3 * H];
        sc2[r] = p.cells[(((long long)dir * (p.Ti + 2) + t + 1) * p.B + b) * H + u];
        scp[r] = p.cells[(((long long)dir * (p.Ti + 2) + tp + 1) * p.B + b) * H + u];
        dov[r] = p.dout[((long long)b * p.Ti + t) * 2 * H + dir * H + u];
      } else {
        sg[r][0] = sg[r][1] = sg[r][2] = sg[r][3] = 0.f; sc2[r] = scp[r] = dov[r] = 0.f;
      }
    }
    float acc[2] = {0.f, 0.f};
    if (s > 0) {
      // dh_prev[b][u] = sum_k dg_next[b][k] W_hh[k][u] over K = 4H gate gradients.  Register-blocked: warp w takes k = 32 w .. 32 w + 31
      // of every 256-wide chunk for ALL 64 rows x 8 units (lane -> 2 rows x 8 units = 16 accumulators: 16 FMAs per 4 shared-memory
      // loads; warp = unit with one accumulator per row was bound by shared-memory bandwidth at 2 FMAs per 3 loads), then the 8
      // partial sums of every (row, unit) are reduced through shared memory.  The global loads of chunk c + 1 are in flight while
      // chunk c is multiplied.
      if (threadIdx.x == 0) wait_counter(cnt, (unsigned)(ncta * s));
      __syncthreads();
      const float* dgn = p.dg[dir] + ((long long)2 + tn) * K;       // row (b = 0, tn); batch stride Tp * K
      float4 v[16];
      auto load_chunk = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = threadIdx.x + j * UPC * 32;
          const int m = (i * 4) / KC, k = (i * 4) % KC;
          v[j] = (m < p.B) ? __ldcg(reinterpret_cast<const float4*>(dgn + (long long)m * p.Tp * K + k0 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float a8[2][UPC];
#pragma unroll
      for (int uu = 0; uu < UPC; ++uu) a8[0][uu] = a8[1][uu] = 0.f;
      load_chunk(0);
      for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = threadIdx.x + j * UPC * 32;
          const int m = (i * 4) / KC, k = (i * 4) % KC;
          float* d = dT + m * DS + k;
          d[0] = v[j].x; d[1] = v[j].y; d[2] = v[j].z; d[3] = v[j].w;
        }
        __syncthreads();
        if (k0 + KC < K) load_chunk(k0 + KC);
        const float* wk = ws + (size_t)(k0 + 32 * warp) * UPC;
        const float* d0 = dT + lane * DS + 32 * warp;
        const float* d1 = dT + (32 + lane) * DS + 32 * warp;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
          const float x0 = d0[k], x1 = d1[k];
          const float4 wa = *reinterpret_cast<const float4*>(wk + k * UPC), wb = *reinterpret_cast<const float4*>(wk + k * UPC + 4);
          a8[0][0] = fmaf(x0, wa.x, a8[0][0]); a8[0][1] = fmaf(x0, wa.y, a8[0][1]); a8[0][2] = fmaf(x0, wa.z, a8[0][2]); a8[0][3] = fmaf(x0, wa.w, a8[0][3]);
          a8[0][4] = fmaf(x0, wb.x, a8[0][4]); a8[0][5] = fmaf(x0, wb.y, a8[0][5]); a8[0][6] = fmaf(x0, wb.z, a8[0][6]); a8[0][7] = fmaf(x0, wb.w, a8[0][7]);
          a8[1][0] = fmaf(x1, wa.x, a8[1][0]); a8[1][1] = fmaf(x1, wa.y, a8[1][1]); a8[1][2] = fmaf(x1, wa.z, a8[1][2]); a8[1][3] = fmaf(x1, wa.w, a8[1][3]);
          a8[1][4] = fmaf(x1, wb.x, a8[1][4]); a8[1][5] = fmaf(x1, wb.y, a8[1][5]); a8[1][6] = fmaf(x1, wb.z, a8[1][6]); a8[1][7] = fmaf(x1, wb.w, a8[1][7]);
        }
      }
#pragma unroll
      for (int uu = 0; uu < UPC; ++uu) {
        red[(warp * UPC + uu) * MT + lane] = a8[0][uu];
        red[(warp * UPC + uu) * MT + 32 + lane] = a8[1][uu];
      }
      __syncthreads();
#pragma unroll
      for (int w2 = 0; w2 < UPC; ++w2) {       // 8 warps
        acc[0] += red[(w2 * UPC + warp) * MT + lane];
        acc[1] += red[(w2 * UPC + warp) * MT + 32 + lane];
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b >= p.B) continue;
      float* dg = p.dg[dir] + ((long long)b * p.Tp + 2 + t) * K + u;
      if (t >= len[r]) { dg[0] = 0.f; dg[H] = 0.f; dg[2 * H] = 0.f; dg[3 * H] = 0.f; continue; }
      const float dh = acc[r] + dov[r];
      const float ig = sg[r][0], fg = sg[r][1], gg = sg[r][2], og = sg[r][3];
      const float tc = tanhf(sc2[r]);
      const float dc = dcs[r] + dh * og * (1.f - tc * tc);
      dg[0] = dc * gg * ig * (1.f - ig);
      dg[H] = dc * scp[r] * fg * (1.f - fg);
      dg[2 * H] = dc * ig * (1.f - gg * gg);
      dg[3 * H] = dh * tc * og * (1.f - og);
      dcs[r] = dc * fg;
    }
    __syncthreads();
    if (threadIdx.x == 0) signal_counter(cnt);
  }
}


// ---- reference-encoder GRU (reference modules.py:60-80: nn.GRU(128 * n_mel / 64, 256, batch_first), last hidden state) ---------
// Same scheme, one direction: H / 8 CTAs, warp = hidden unit, lane -> batch rows lane, lane + 32; W_hh rows (r, z, n) of the CTA's
// units stay in shared memory for all Tq steps.  Arithmetic identical to gru_pointwise_fwd / _bwd + exact fp32 GEMMs.
struct GruFwd {
  const float* gi; long long gi_bs;   // x W_ih^T rows: (b, t) at gi + b * gi_bs + t * 3H
  const float* w_hh;                  // [3H, H]
  const float* b_ih; const float* b_hh;
  float* hs;                          // [Tq + 1, B, H], slot 0 = h0 (zero), slot t + 1 = h after step t
  float* save;                        // [Tq, B, 4H]: r, z, n, gh_n
  unsigned* counter;
  int B, H, Tq;
};

__global__ void __launch_bounds__(UPC * 32, 1) gru_seq_fwd_kernel(GruFwd p) {
  extern __shared__ __align__(16) float smg[];
  const int H = p.H, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UPC, u = u0 + warp, HS = H + 1, ncta = gridDim.x;
  float* hT = smg;                       // [MT][H+1]
  float* ws = hT + ((MT * HS + 3) & ~3); // [H][UPC][4] (r, z, n, -)
  for (int i = threadIdx.x; i < H * UPC * 3; i += UPC * 32) {
    const int k = i % H, g = (i / H) % 3, uu = i / (3 * H);
    ws[(k * UPC + uu) * 4 + g] = p.w_hh[((long long)g * H + u0 + uu) * H + k];
  }
  const float br = p.b_ih[u], bz = p.b_ih[H + u], bn = p.b_ih[2 * H + u];
  const float cr = p.b_hh[u], cz = p.b_hh[H + u], cn = p.b_hh[2 * H + u];
  for (int s = 0; s < p.Tq; ++s) {
    float giv[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b < p.B) {
        const float* g = p.gi + (long long)b * p.gi_bs + (long long)s * 3 * H;
        giv[r][0] = g[u]; giv[r][1] = g[H + u]; giv[r][2] = g[2 * H + u];
      } else {
        giv[r][0] = giv[r][1] = giv[r][2] = 0.f;
      }
    }
    const float* hprev = p.hs + (long long)s * p.B * H;
    if (s > 0) {
      if (threadIdx.x == 0) wait_counter(p.counter, (unsigned)(ncta * s));
      __syncthreads();
      const int nv = MT * H / 4;
      for (int i0 = threadIdx.x; i0 < nv; i0 += UPC * 32 * 16) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = i0 + j * UPC * 32;
          const int m = (i * 4) / H, k = (i * 4) % H;
          v[j] = (i < nv && m < p.B) ? __ldcg(reinterpret_cast<const float4*>(hprev + (long long)m * H + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    } else {
      for (int i = threadIdx.x; i < MT * HS; i += UPC * 32) hT[i] = 0.f;
    }
    __syncthreads();
    float acc[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 0.f;
    if (s > 0) {
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float a0 = hT[lane * HS + k], a1 = hT[(32 + lane) * HS + k];
        const float4 w = *reinterpret_cast<const float4*>(ws + (k * UPC + warp) * 4);
        acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]); acc[0][2] = fmaf(a0, w.z, acc[0][2]);
        acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]); acc[1][2] = fmaf(a1, w.z, acc[1][2]);
      }
    }
    float* hnext = p.hs + (long long)(s + 1) * p.B * H;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b >= p.B) continue;
      const float rg = t2v_sigmoid(giv[r][0] + br + acc[r][0] + cr);
      const float zg = t2v_sigmoid(giv[r][1] + bz + acc[r][1] + cz);
      const float ghn = acc[r][2] + cn;
      const float ng = tanhf(giv[r][2] + bn + rg * ghn);
      const float hp = hT[b * HS + u];
      hnext[(long long)b * H + u] = (1.f - zg) * ng + zg * hp;
      float* sv = p.save + ((long long)s * p.B + b) * 4 * H + u;
      sv[0] = rg; sv[H] = zg; sv[2 * H] = ng; sv[3 * H] = ghn;
    }
    __syncthreads();
    if (threadIdx.x == 0) signal_counter(p.counter);
  }
}

struct GruBwd {
  const float* w_hh;                  // [3H, H]
  const float* dh_last;               // [B, H]
  const float* save; const float* hs;
  float* dgi; long long dgi_bs;       // gradient wrt x W_ih^T: row (b, t) at dgi + b * dgi_bs + t * 3H
  float* dgh;                         // [Tq, B, 3H] gradient wrt h W_hh^T (+ b_hh), kept for the batched dW_hh / db_hh
  unsigned* counter;
  int B, H, Tq;
};

__global__ void __launch_bounds__(UPC * 32, 1) gru_seq_bwd_kernel(GruBwd p) {
  extern __shared__ __align__(16) float smg[];
  const int H = p.H, K = 3 * H, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UPC, u = u0 + warp, ncta = gridDim.x;
  constexpr int KC = 256, DS = KC + 1;
  float* dT = smg;                       // [MT][KC+1]
  float* ws = dT + MT * DS;              // [K][UPC]: W_hh[k][u]
  for (int i = threadIdx.x; i < K * UPC; i += UPC * 32) {
    const int uu = i % UPC, k = i / UPC;
    ws[k * UPC + uu] = p.w_hh[(long long)k * H + u0 + uu];
  }
  float carry[2] = {0.f, 0.f};
  __syncthreads();
  for (int s = 0; s < p.Tq; ++s) {
    const int t = p.Tq - 1 - s;
    float sr[2], sz[2], sn[2], sg[2], hp[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b < p.B) {
        const float* sv = p.save + ((long long)t * p.B + b) * 4 * H + u;
        sr[r] = sv[0]; sz[r] = sv[H]; sn[r] = sv[2 * H]; sg[r] = sv[3 * H];
        hp[r] = p.hs[((long long)t * p.B + b) * H + u];
      } else {
        sr[r] = sz[r] = sn[r] = sg[r] = hp[r] = 0.f;
      }
    }
    float acc[2] = {0.f, 0.f};
    if (s > 0) {
      if (threadIdx.x == 0) wait_counter(p.counter, (unsigned)(ncta * s));
      const float* dgn = p.dgh + (long long)(t + 1) * p.B * K;
      for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        const int nv = MT * KC / 4;
        for (int i0 = threadIdx.x; i0 < nv; i0 += UPC * 32 * 16) {
          float4 v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int i = i0 + j * UPC * 32;
            const int m = (i * 4) / KC, k = (i * 4) % KC;
            v[j] = (i < nv && m < p.B) ? __ldcg(reinterpret_cast<const float4*>(dgn + (long long)m * K + k0 + k))
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
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < KC; ++k) {
          const float w = ws[(k0 + k) * UPC + warp];
          acc[0] = fmaf(dT[lane * DS + k], w, acc[0]);
          acc[1] = fmaf(dT[(32 + lane) * DS + k], w, acc[1]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = r * 32 + lane;
      if (b >= p.B) continue;
      const float d = (s == 0) ? p.dh_last[(long long)b * H + u] : carry[r] + acc[r];
      const float dn = d * (1.f - sz[r]) * (1.f - sn[r] * sn[r]);
      const float dz = d * (hp[r] - sn[r]) * sz[r] * (1.f - sz[r]);
      const float dr = dn * sg[r] * sr[r] * (1.f - sr[r]);
      float* a = p.dgi + (long long)b * p.dgi_bs + (long long)t * K + u;
      a[0] = dr; a[H] = dz; a[2 * H] = dn;
      float* c = p.dgh + ((long long)t * p.B + b) * K + u;
      c[0] = dr; c[H] = dz; c[2 * H] = dn * sr[r];
      carry[r] = d * sz[r];
    }
    __syncthreads();
    if (threadIdx.x == 0) signal_counter(p.counter);
  }
}

// All CTAs of these kernels (64 / 32) wait on each other, so all of them must become resident.  Plain launch by default: their CTAs
// take SMs one by one as the kernels around them (weight-gradient GEMMs on lower-priority streams, whose CTAs always run to completion)
// free them -- at most 96 of the 148 SMs are needed when the BiLSTM and GRU kernels overlap, and the two decoder loops (the other
// resident kernels of the step) are ordered before / after them by the step's dependencies.  A cooperative launch (T2V_RNN_COOP=1)
// is placed as a whole or not at all: behind a 900-CTA GEMM it was observed to wait for that whole kernel to drain (0.9 ms on the
// critical path of the backward tail).  Bounded waits (trap) stay as the guard against a grid that never becomes resident.
template <typename Args>
cudaError_t launch_coop(void (*kernel)(Args), dim3 grid, int block, size_t smem, cudaStream_t st, Args a) {
  static const bool coop = getenv("T2V_RNN_COOP") && getenv("T2V_RNN_COOP")[0] == '1';
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

}  // namespace

// Whole-sequence forward of both directions.  gx0 / gx1: [B*Tp, 4H] input projections (+ b_ih) on the padded rows (Tp = Ti + 4, valid
// rows 2..Ti+1); seq: [B*Tp, 2H] output rows; gates [2,Ti,B,4H], cells [2,Ti+2,B,H] (zero-initialised: slots 0 and Ti+1 stay zero);
// hbuf: scratch [2,2,B,H]; counters: 64 unsigned, zeroed by this call.  B <= 64, H <= 512 (larger batches: the per-step launches
// t2v_bilstm_step_fwd / _bwd).
T2V_API int t2v_bilstm_seq_fwd(const float* gx0, const float* gx1, const float* whh0, const float* whh1, const float* bhh0,
                               const float* bhh1, float* seq, float* gates, float* cells, float* hbuf, unsigned int* counters,
                               const long long* lens, int B, int H, int Ti, cudaStream_t st) {
  T2V_ARG_CHECK(gx0 && gx1 && whh0 && whh1 && seq && gates && cells && hbuf && counters && B > 0 && Ti > 0, "null / shape");
  T2V_ARG_CHECK(B <= MT && H % UPC == 0 && H <= 512 && H % 4 == 0, "B <= 64, H <= 512");
  SeqFwd a;
  a.gx[0] = gx0; a.gx[1] = gx1; a.w_hh[0] = whh0; a.w_hh[1] = whh1; a.b_hh[0] = bhh0; a.b_hh[1] = bhh1; a.seq = seq;
  a.gates = gates; a.cells = cells; a.hbuf = hbuf; a.counters = counters; a.lens = lens; a.B = B; a.H = H; a.Ti = Ti; a.Tp = Ti + 4;
  const size_t smem = sizeof(float) * (size_t)(((MT * (H + 1) + 3) & ~3) + H * UPC * 4);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(bilstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64 * sizeof(unsigned), st));
  T2V_CUDA_CHECK(launch_coop(bilstm_seq_fwd_kernel, dim3(H / UPC, 2), UPC * 32, smem, st, a));
  T2V_COUNT_LAUNCH();
  return 0;
}

// Whole-sequence backward: dout [B,Ti,2H] -> dg0 / dg1 [B*Tp, 4H] (gate gradients on the padded rows; pad rows untouched).
T2V_API int t2v_bilstm_seq_bwd(const float* whhT0, const float* whhT1, const float* dout, const float* gates, const float* cells,
                               float* dg0, float* dg1, unsigned int* counters, const long long* lens, int B, int H, int Ti,
                               cudaStream_t st) {
  T2V_ARG_CHECK(whhT0 && whhT1 && dout && gates && cells && dg0 && dg1 && counters && B > 0 && Ti > 0, "null / shape");
  T2V_ARG_CHECK(B <= MT && H % UPC == 0 && (4 * H) % 256 == 0 && H <= 512, "B <= 64, H <= 512");
  SeqBwd a;
  a.w_hhT[0] = whhT0; a.w_hhT[1] = whhT1; a.dout = dout; a.gates = gates; a.cells = cells; a.dg[0] = dg0; a.dg[1] = dg1;
  a.counters = counters; a.lens = lens; a.B = B; a.H = H; a.Ti = Ti; a.Tp = Ti + 4;
  const size_t smem = sizeof(float) * (size_t)(MT * 257 + 4 * H * UPC + UPC * UPC * MT);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(bilstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64 * sizeof(unsigned), st));
  T2V_CUDA_CHECK(launch_coop(bilstm_seq_bwd_kernel, dim3(H / UPC, 2), UPC * 32, smem, st, a));
  T2V_COUNT_LAUNCH();
  return 0;
}

// Reference-encoder GRU, all Tq steps in one launch.  gi: rows (b, t) at gi + b * gi_bs + t * 3H (x W_ih^T without bias);
// hs [Tq+1, B, H] with slot 0 zero (h0); save [Tq, B, 4H]; counter: 32 unsigned, zeroed by this call.  B <= 64, H <= 512.
T2V_API int t2v_gru_seq_fwd(const float* gi, long long gi_bs, const float* w_hh, const float* b_ih, const float* b_hh, float* hs,
                            float* save, unsigned int* counter, int B, int H, int Tq, cudaStream_t st) {
  T2V_ARG_CHECK(gi && w_hh && b_ih && b_hh && hs && save && counter && B > 0 && Tq > 0, "null / shape");
  T2V_ARG_CHECK(B <= MT && H % UPC == 0 && H <= 512 && H % 4 == 0, "B <= 64, H <= 512");
  GruFwd a;
  a.gi = gi; a.gi_bs = gi_bs; a.w_hh = w_hh; a.b_ih = b_ih; a.b_hh = b_hh; a.hs = hs; a.save = save; a.counter = counter;
  a.B = B; a.H = H; a.Tq = Tq;
  const size_t smem = sizeof(float) * (size_t)(((MT * (H + 1) + 3) & ~3) + H * UPC * 4);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(gru_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(cudaMemsetAsync(counter, 0, 32 * sizeof(unsigned), st));
  T2V_CUDA_CHECK(launch_coop(gru_seq_fwd_kernel, dim3(H / UPC), UPC * 32, smem, st, a));
  T2V_COUNT_LAUNCH();
  return 0;
}

// Backward through time: dh_last [B,H] -> dgi rows (b, t) at dgi + b * dgi_bs + t * 3H and dgh [Tq, B, 3H] (the caller forms
// dW_hh = sum_t dgh[t]^T hs[t] and db_hh = colsum(dgh) as two batched reductions).
T2V_API int t2v_gru_seq_bwd(const float* w_hh, const float* dh_last, const float* save, const float* hs, float* dgi,
                            long long dgi_bs, float* dgh, unsigned int* counter, int B, int H, int Tq, cudaStream_t st) {
  T2V_ARG_CHECK(w_hh && dh_last && save && hs && dgi && dgh && counter && B > 0 && Tq > 0, "null / shape");
  T2V_ARG_CHECK(B <= MT && H % UPC == 0 && (3 * H) % 256 == 0 && H <= 512, "B <= 64, 3H % 256 == 0, H <= 512");
  GruBwd a;
  a.w_hh = w_hh; a.dh_last = dh_last; a.save = save; a.hs = hs; a.dgi = dgi; a.dgi_bs = dgi_bs; a.dgh = dgh; a.counter = counter;
  a.B = B; a.H = H; a.Tq = Tq;
  const size_t smem = sizeof(float) * (size_t)(MT * 257 + 3 * H * UPC);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(gru_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  T2V_CUDA_CHECK(cudaMemsetAsync(counter, 0, 32 * sizeof(unsigned), st));
  T2V_CUDA_CHECK(launch_coop(gru_seq_bwd_kernel, dim3(H / UPC), UPC * 32, smem, st, a));
  T2V_COUNT_LAUNCH();
  return 0;
}
