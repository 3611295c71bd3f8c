// Memory-bound kernels of the hot path: embedding gather, BatchNorm (1d over padded channels-last rows and 2d
// over NHWC rows -- same kernels), activation+dropout, weight (re)packing, transposes, LSTM / GRU cell pointwise
// math (forward + backward), broadcast add / reductions.  All activations are fp32, channels-last.
//
// Layout convention for Conv1d stacks (Encoder model.py:159-177, Postnet model.py:105-148): a tensor the
// reference holds as [B, C, T] is stored as rows [B*(T+4), C] with two zero rows on each side of every
// utterance ("padded channels-last"), so that a k=5/p=2 convolution is a GEMM over overlapping row windows.
#include "t2v_common.cuh"

namespace {

// valid-row predicate of a padded row space: row r is valid iff lo <= (r % period) < hi
struct RowSpace {
  int period, lo, hi;
  __device__ __forceinline__ bool valid(long long r) const {
    int x = (int)(r % period);
    return x >= lo && x < hi;
  }
};

// ------------------------------------------------------------------------------------------ embedding
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                     float* __restrict__ out, int B, int T, int C, int n_symbols, int rnd) {
  // one block per (b,t); out row = b*(T+4)+2+t
  const int bt = blockIdx.x;
  const int b = bt / T, t = bt % T;
  long long id = ids[bt];
  if (id < 0 || id >= n_symbols) __trap();   // nn.Embedding raises on out-of-range ids
  const float4* src = reinterpret_cast<const float4*>(table + id * C);
  float4* dst = reinterpret_cast<float4*>(out + ((long long)b * (T + 4) + 2 + t) * C);
  for (int i = threadIdx.x; i < C / 4; i += blockDim.x) {
    float4 v = src[i];
    v.x = t2v_rnd(v.x, rnd); v.y = t2v_rnd(v.y, rnd); v.z = t2v_rnd(v.z, rnd); v.w = t2v_rnd(v.w, rnd);
    dst[i] = v;
  }
}
__global__ void embedding_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dout,
                                     float* __restrict__ dtable, int B, int T, int C) {
  const int bt = blockIdx.x;
  const int b = bt / T, t = bt % T;
  long long id = ids[bt];
  const float* src = dout + ((long long)b * (T + 4) + 2 + t) * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dtable + id * C + i, src[i]);
}

// ------------------------------------------------------------------------------------------ column reductions
// MODE 0: sum x, sum x^2          (BatchNorm statistics)
// MODE 1: sum g, sum g*xhat       (BatchNorm backward: dbeta, dgamma), g = act'/dropout-scaled upstream grad
// MODE 2: sum x only              (bias gradients)
struct BnCtx {
  const float* y;        // pre-BN tensor (conv output), rows x C
  const float* mean;     // [C]
  const float* invstd;   // [C]
  const float* gamma;
  const float* beta;
  int act;               // 0 none, 1 relu, 2 tanh
  T2VDrop drop;
  int T;                 // frames per utterance for the dropout index ((b*C+c)*T+t); rows b*period+lo+t
};

__device__ __forceinline__ float act_fwd(float v, int act) {
  return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? tanhf(v) : v);
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) { const float t = t2v_tanh(pre); return 1.f - t * t; }    // exp-based tanh: absolute error ~2e-7 on a factor in [0, 1]
  return 1.f;
}
__device__ __forceinline__ uint64_t drop_index(const RowSpace& rs, long long r, int c, int C, int T) {
  long long b = r / rs.period;
  int t = (int)(r % rs.period) - rs.lo;
  return ((uint64_t)b * C + c) * (uint64_t)T + t;
}

template <int MODE>
__global__ void __launch_bounds__(256)
colreduce_kernel(const float* __restrict__ x, long long rows, int C, RowSpace rs, int rows_per_block, BnCtx ctx,
                 double* __restrict__ out0, double* __restrict__ out1) {
  // block (32, 8): x -> channel, y -> row phase
  __shared__ float s0[8][33], s1[8][33];
  const int c = blockIdx.y * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    float mean = 0.f, invstd = 0.f, gamma = 0.f, beta = 0.f;
    if (MODE == 1) { mean = ctx.mean[c]; invstd = ctx.invstd[c]; gamma = ctx.gamma[c]; beta = ctx.beta[c]; }
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      if (!rs.valid(r)) continue;
      const float v = x[r * C + c];
      if (MODE == 0) { a0 += v; a1 += v * v; }
      else if (MODE == 2) { a0 += v; }
      else {
        const float xhat = (ctx.y[r * C + c] - mean) * invstd;
        const float pre = gamma * xhat + beta;
        const float g = v * act_grad(pre, ctx.act) * t2v_keep_scale(ctx.drop, drop_index(rs, r, c, C, ctx.T));
        a0 += g; a1 += g * xhat;
      }
    }
  }
  s0[threadIdx.y][threadIdx.x] = a0;
  s1[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += s0[i][threadIdx.x]; t1 += s1[i][threadIdx.x]; }
    atomicAdd(out0 + c, (double)t0);
    if (MODE != 2) atomicAdd(out1 + c, (double)t1);
  }
}

// ---- shared pieces of the vectorised BatchNorm kernels.  Thread layout: a block of 256 threads is LX lanes (LX = 2^lx_log2 <= 32, four
// consecutive channels each) x 256 / LX row phases, so narrow tensors (the reference encoder's C = 32 .. 128 NHWC rows) keep every
// lane busy.  Row arithmetic is 32-bit (rows < 2^31 is checked on the host): the 64-bit divisions of the scalar kernels cost more
// instructions than the rest of the element's math.
__device__ __forceinline__ uint16_t f16_sat_bits(float x) {      // round to nearest, clamp to +-65504 instead of producing inf
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
struct RowPos { unsigned b; int t; bool ok; };
__device__ __forceinline__ RowPos row_pos(const RowSpace& rs, unsigned r) {
  RowPos p;
  p.b = r / (unsigned)rs.period;
  const int x = (int)(r - p.b * (unsigned)rs.period);
  p.t = x - rs.lo;
  p.ok = x >= rs.lo && x < rs.hi;
  return p;
}
// dropout keep-scale with the seed dereference and the key multiply hoisted out of the element loop: same bits as t2v_keep_scale
struct DropFast {
  const float* mask; uint64_t key0; float p, scale;
  __device__ __forceinline__ void init(const T2VDrop& d) {
    mask = d.mask; p = d.p; scale = d.p > 0.f ? 1.f / (1.f - d.p) : 1.f;
    key0 = (d.p > 0.f && !d.mask) ? t2v_resolve_seed(d.seed) * 0x9E3779B97F4A7C15ULL + ((uint64_t)d.site << 40) : 0ull;
  }
  __device__ __forceinline__ float keep(uint64_t idx) const {
    if (p <= 0.f) return 1.f;
    const float k = mask ? mask[idx] : (((float)(t2v_hash32(key0 + idx) & 0xFFFFFFu) * (1.0f / 16777216.0f)) >= p ? 1.f : 0.f);
    return k * scale;
  }
};
// Vectorised column reduction (C % 4 == 0): these reductions are HBM streams (one read of x, MODE 1: of x and y); four rows in
// flight per thread.
template <int MODE>
__global__ void __launch_bounds__(256)
colreduce4_kernel(const float* __restrict__ x, long long rows, int C, RowSpace rs, int rows_per_block, BnCtx ctx,
                  double* __restrict__ out0, double* __restrict__ out1, int lx_log2) {
  __shared__ float4 s0[256], s1[256];
  const int LX = 1 << lx_log2, LY = 256 >> lx_log2;
  const int tx = threadIdx.x & (LX - 1), ty = threadIdx.x >> lx_log2;
  const int c = (blockIdx.y * LX + tx) * 4;
  const unsigned r0 = blockIdx.x * (unsigned)rows_per_block;
  unsigned r1 = r0 + (unsigned)rows_per_block;
  if (r1 > (unsigned)rows) r1 = (unsigned)rows;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (c < C) {
    float mm[4] = {0.f, 0.f, 0.f, 0.f}, is[4] = {0.f, 0.f, 0.f, 0.f}, gm[4] = {0.f, 0.f, 0.f, 0.f}, bt[4] = {0.f, 0.f, 0.f, 0.f};
    DropFast drop;
    drop.init(ctx.drop);
    if (MODE == 1) {      // scalar loads: parameters may be views into the optimizer's flat buffer (4-byte aligned only)
#pragma unroll
      for (int j = 0; j < 4; ++j) { mm[j] = ctx.mean[c + j]; is[j] = ctx.invstd[c + j]; gm[j] = ctx.gamma[c + j]; bt[j] = ctx.beta[c + j]; }
    }
    for (unsigned rb = r0 + ty; rb < r1; rb += 4 * LY) {
      float4 v[4], y[4];
      RowPos rp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned r = rb + LY * i;
        rp[i] = row_pos(rs, r);
        rp[i].ok = rp[i].ok && r < r1;
        const long long off = (long long)r * C + c;
        v[i] = rp[i].ok ? __ldcs(reinterpret_cast<const float4*>(x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 1) y[i] = rp[i].ok ? __ldcs(reinterpret_cast<const float4*>(ctx.y + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!rp[i].ok) continue;
        if (MODE == 0) {
          a0.x += v[i].x; a0.y += v[i].y; a0.z += v[i].z; a0.w += v[i].w;
          a1.x += v[i].x * v[i].x; a1.y += v[i].y * v[i].y; a1.z += v[i].z * v[i].z; a1.w += v[i].w * v[i].w;
        } else if (MODE == 2) {
          a0.x += v[i].x; a0.y += v[i].y; a0.z += v[i].z; a0.w += v[i].w;
        } else {
          const float vv[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, yy[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
          const uint64_t base = ((uint64_t)rp[i].b * C + c) * (uint64_t)ctx.T + rp[i].t;      // dropout index of (b, c, t)
          float g[4], xh[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xh[j] = (yy[j] - mm[j]) * is[j];
            const float pre = gm[j] * xh[j] + bt[j];
            g[j] = vv[j] * act_grad(pre, ctx.act) * drop.keep(base + (uint64_t)j * ctx.T);
          }
          a0.x += g[0]; a0.y += g[1]; a0.z += g[2]; a0.w += g[3];
          a1.x += g[0] * xh[0]; a1.y += g[1] * xh[1]; a1.z += g[2] * xh[2]; a1.w += g[3] * xh[3];
        }
      }
    }
  }
  s0[threadIdx.x] = a0;
  s1[threadIdx.x] = a1;
  __syncthreads();
  if (ty == 0 && c < C) {
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    for (int i = 0; i < LY; ++i) {
      const float4 p0 = s0[i * LX + tx], p1 = s1[i * LX + tx];
      t0.x += p0.x; t0.y += p0.y; t0.z += p0.z; t0.w += p0.w;
      t1.x += p1.x; t1.y += p1.y; t1.z += p1.z; t1.w += p1.w;
    }
    atomicAdd(out0 + c, (double)t0.x); atomicAdd(out0 + c + 1, (double)t0.y);
    atomicAdd(out0 + c + 2, (double)t0.z); atomicAdd(out0 + c + 3, (double)t0.w);
    if (MODE != 2) {
      atomicAdd(out1 + c, (double)t1.x); atomicAdd(out1 + c + 1, (double)t1.y);
      atomicAdd(out1 + c + 2, (double)t1.z); atomicAdd(out1 + c + 3, (double)t1.w);
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double n, int C,
                                   float eps, float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const double m = sum[c] / n;
    double var = sumsq[c] / n - m * m;
    if (var < 0) var = 0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = var * (n / (n > 1 ? n - 1 : 1));
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
}
__global__ void bn_eval_prepare_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                       int C, float eps, float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = running_mean[c]; invstd[c] = rsqrtf(running_var[c] + eps); }
}

// out = dropout(act(gamma*(y-mean)*invstd+beta)) on valid rows, 0 on pad rows
// out_lo (optional): the residual x - round(x) on the tf32 grid, for the split (error-compensated) tensor-core GEMMs
// Thread layout as in colreduce4_kernel; a thread keeps the parameters of its four channels in registers and walks rows.
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ y, float* __restrict__ out, float* __restrict__ out_lo, long long rows,
                  int C, RowSpace rs, BnCtx ctx, int rnd, int rows_per_block, int lx_log2, uint16_t* __restrict__ hi16,
                  uint16_t* __restrict__ lo16) {
  const int LX = 1 << lx_log2, LY = 256 >> lx_log2;
  const int tx = threadIdx.x & (LX - 1), ty = threadIdx.x >> lx_log2;
  const int c = (blockIdx.y * LX + tx) * 4;
  if (c >= C) return;
  const unsigned r0 = blockIdx.x * (unsigned)rows_per_block;
  unsigned r1 = r0 + (unsigned)rows_per_block;
  if (r1 > (unsigned)rows) r1 = (unsigned)rows;
  float sc[4], sh[4], mm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { mm[j] = ctx.mean[c + j]; sc[j] = ctx.invstd[c + j]; sh[j] = ctx.beta[c + j]; }
  float gm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) gm[j] = ctx.gamma[c + j];
  DropFast drop;
  drop.init(ctx.drop);
  for (unsigned r = r0 + ty; r < r1; r += LY) {
    const RowPos rp = row_pos(rs, r);
    const long long off = (long long)r * C + c;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), ol = o;
    if (rp.ok) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(y + off));
      const float in[4] = {v.x, v.y, v.z, v.w};
      const uint64_t base = ((uint64_t)rp.b * C + c) * (uint64_t)ctx.T + rp.t;
      float res[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pre = gm[j] * ((in[j] - mm[j]) * sc[j]) + sh[j];
        const float x = act_fwd(pre, ctx.act) * drop.keep(base + (uint64_t)j * ctx.T);
        res[j] = t2v_rnd(x, rnd);
        lo[j] = hi16 ? (x - res[j]) : t2v_tf32(x - res[j]);      // fp16 split: the exact residual feeds the 16-bit encoding below
      }
      o = make_float4(res[0], res[1], res[2], res[3]);
      ol = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    *reinterpret_cast<float4*>(out + off) = o;
    if (out_lo) *reinterpret_cast<float4*>(out_lo + off) = ol;
    if (hi16) {      // fp16 split operand x = hi + lo (rnd == 2: `out` already sits on the fp16 grid, so hi == out exactly)
      const float xs[4] = {o.x + ol.x, o.y + ol.y, o.z + ol.z, o.w + ol.w};
      uint16_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { h[j] = f16_sat_bits(xs[j]); l[j] = t2v_f16_bits(xs[j] - t2v_f16_to_f32(h[j])); }
      *reinterpret_cast<uint2*>(hi16 + off) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
      *reinterpret_cast<uint2*>(lo16 + off) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
  }
}

// x * scale -> fp16 hi + fp16 lo (x * scale = hi + lo up to 2^-22 relative): operands of the 16-bit split tensor-core GEMMs
__global__ void split16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, long long n4, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const float xs[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
    uint16_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] = f16_sat_bits(xs[j]); l[j] = t2v_f16_bits(xs[j] - t2v_f16_to_f32(h[j])); }
    hi[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    lo[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  }
}

// training-mode BN backward (given dgamma/dbeta sums): dy = gamma*invstd*(g - dbeta/n - xhat*dgamma/n); 0 on pad rows
// eval-mode (use_batch_stats=0): dy = gamma*invstd*g
__global__ void __launch_bounds__(256)
bn_act_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dy, long long rows, int C,
                  RowSpace rs, BnCtx ctx, const double* __restrict__ dbeta_sum,
                  const double* __restrict__ dgamma_sum, double n, int use_batch_stats, int rnd, int rows_per_block, int lx_log2) {
  const int LX = 1 << lx_log2, LY = 256 >> lx_log2;
  const int tx = threadIdx.x & (LX - 1), ty = threadIdx.x >> lx_log2;
  const int c = (blockIdx.y * LX + tx) * 4;
  if (c >= C) return;
  const unsigned r0 = blockIdx.x * (unsigned)rows_per_block;
  unsigned r1 = r0 + (unsigned)rows_per_block;
  if (r1 > (unsigned)rows) r1 = (unsigned)rows;
  float mm[4], is[4], gm[4], bt[4], mg[4], mgx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mm[j] = ctx.mean[c + j]; is[j] = ctx.invstd[c + j]; gm[j] = ctx.gamma[c + j]; bt[j] = ctx.beta[c + j];
    mg[j] = use_batch_stats ? (float)(dbeta_sum[c + j] / n) : 0.f;       // the two per-channel means, once per thread
    mgx[j] = use_batch_stats ? (float)(dgamma_sum[c + j] / n) : 0.f;
  }
  DropFast drop;
  drop.init(ctx.drop);
  for (unsigned rb = r0 + ty; rb < r1; rb += 2 * LY) {
    float4 gv[2], yv[2];
    RowPos rp[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned r = rb + LY * i;
      rp[i] = row_pos(rs, r);
      rp[i].ok = rp[i].ok && r < r1;
      const long long off = (long long)r * C + c;
      if (rp[i].ok) { gv[i] = __ldcs(reinterpret_cast<const float4*>(dout + off)); yv[i] = __ldcs(reinterpret_cast<const float4*>(ctx.y + off)); }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned r = rb + LY * i;
      if (r >= r1) continue;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rp[i].ok) {
        const float gin[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w}, yin[4] = {yv[i].x, yv[i].y, yv[i].z, yv[i].w};
        const uint64_t base = ((uint64_t)rp[i].b * C + c) * (uint64_t)ctx.T + rp[i].t;
        float res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xhat = (yin[j] - mm[j]) * is[j];
          const float pre = gm[j] * xhat + bt[j];
          const float g = gin[j] * act_grad(pre, ctx.act) * drop.keep(base + (uint64_t)j * ctx.T);
          res[j] = t2v_rnd(use_batch_stats ? gm[j] * is[j] * (g - mg[j] - xhat * mgx[j]) : gm[j] * is[j] * g, rnd);
        }
        o = make_float4(res[0], res[1], res[2], res[3]);
      }
      *reinterpret_cast<float4*>(dy + (long long)r * C + c) = o;
    }
  }
}

__global__ void double_to_float_acc_kernel(const double* __restrict__ src, float* __restrict__ dst, int n, float beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (beta != 0.f ? beta * dst[i] : 0.f) + (float)src[i];
}

// ------------------------------------------------------------------------------------------ layout helpers
// generic strided 2-level copy: dst[r*d_rs + c] = src[r*s_rs + c*s_cs] (+ optional accumulate)
__global__ void copy2d_kernel(const float* __restrict__ src, long long s_rs, long long s_cs, float* __restrict__ dst,
                              long long d_rs, long long rows, int cols, float beta, int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i % cols);
  const float v = t2v_rnd(src[r * s_rs + c * s_cs], rnd);
  float* d = dst + r * d_rs + c;
  *d = (beta != 0.f) ? beta * (*d) + v : v;
}
// tiled transpose: out[c*o_ld + r] = in[r*i_ld + c]
__global__ void transpose_kernel(const float* __restrict__ in, long long i_ld, float* __restrict__ out, long long o_ld,
                                 long long rows, int cols, int rnd) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long r = r0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[r * i_ld + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j;
    const long long r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(long long)c * o_ld + r] = t2v_rnd(tile[threadIdx.x][j], rnd);
  }
}
// Conv1d weight [Co,Ci,K] -> tap-major [Co, K*Ci] (flip=0) or dgrad form [Ci, K*Co] with taps reversed (flip=1)
__global__ void conv1d_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int Co, int Ci, int K, int flip,
                                   int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Co * Ci * K) return;
  const int k = (int)(i % K);
  const int ci = (int)((i / K) % Ci);
  const int co = (int)(i / ((long long)K * Ci));
  const float v = t2v_rnd(w[i], rnd);
  if (!flip) out[((long long)co * K + k) * Ci + ci] = v;
  else out[((long long)ci * K + (K - 1 - k)) * Co + co] = v;
}
// gradient in tap-major form [Co, K*Ci] -> accumulate into [Co,Ci,K]
__global__ void conv1d_unpack_grad_kernel(const float* __restrict__ gk, float* __restrict__ gw, int Co, int Ci, int K,
                                          float beta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Co * Ci * K) return;
  const int k = (int)(i % K);
  const int ci = (int)((i / K) % Ci);
  const int co = (int)(i / ((long long)K * Ci));
  const float v = gk[((long long)co * K + k) * Ci + ci];
  gw[i] = (beta != 0.f ? beta * gw[i] : 0.f) + v;
}

__global__ void axpby_kernel(const float* __restrict__ x, float a, float* __restrict__ y, float b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + (b != 0.f ? b * y[i] : 0.f);
}
// y[r, c] += v[b(r), c] where b(r) = r / rows_per_batch (style broadcast-add, model.py:536-537); valid rows only
__global__ void bcast_add_rows_kernel(float* __restrict__ y, const float* __restrict__ v, long long rows, int C,
                                      int rows_per_batch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i % C);
  y[i] += v[(r / rows_per_batch) * C + c];
}
// out[b, c] = sum_{r in batch b} x[r, c]
__global__ void sum_rows_per_batch_kernel(const float* __restrict__ x, float* __restrict__ out, int rows_per_batch,
                                          int C, float beta) {
  const int b = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;       // column blocks across blockIdx.y: B = 1 calls still fill the GPU
  if (c >= C) return;
  const float* p = x + (long long)b * rows_per_batch * C + c;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;              // four independent chains: the loads of a column stay in flight
  int r = 0;
  for (; r + 4 <= rows_per_batch; r += 4) {
    a0 += p[(long long)r * C]; a1 += p[(long long)(r + 1) * C]; a2 += p[(long long)(r + 2) * C]; a3 += p[(long long)(r + 3) * C];
  }
  for (; r < rows_per_batch; ++r) a0 += p[(long long)r * C];
  const float a = (a0 + a1) + (a2 + a3);
  float* o = out + (long long)b * C + c;
  *o = (beta != 0.f ? beta * (*o) : 0.f) + a;
}
// sums `n_parts` partial buffers (stride part_stride) into dst (+ optional second/third plain sources)
__global__ void sum_parts_kernel(const float* __restrict__ parts, int n_parts, long long part_stride,
                                 float* __restrict__ dst, long long n) {
  t2v_pdl_trigger();
  t2v_pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int p = 0; p < n_parts; ++p) a += parts[p * part_stride + i];
  dst[i] = a;
}
// relu + dropout pointwise (Prenet, model.py:101): out = relu(x) * keep/(1-p); logical idx = row*C+c + idx_base
__global__ void relu_drop_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long long o_rs,
                                     long long rows, int C, T2VDrop drop, unsigned long long idx_base, int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i % C);
  out[r * o_rs + c] = t2v_rnd(fmaxf(x[i], 0.f) * t2v_keep_scale(drop, idx_base + (uint64_t)i), rnd);
}
// dx = dout * 1[x>0] * keep/(1-p)
__global__ void relu_drop_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, long long do_rs,
                                     float* __restrict__ dx, long long rows, int C, T2VDrop drop,
                                     unsigned long long idx_base, int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i % C);
  dx[i] = (x[i] > 0.f) ? t2v_rnd(dout[r * do_rs + c] * t2v_keep_scale(drop, idx_base + (uint64_t)i), rnd) : 0.f;
}
__global__ void materialize_mask_kernel(float* __restrict__ out, long long n, T2VDrop drop, unsigned long long idx_base) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (t2v_uniform(t2v_resolve_seed(drop.seed), drop.site, idx_base + (uint64_t)i) >= drop.p) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------ LSTM / GRU cells
struct LstmFwdArgs {
  const float* parts; int n_parts; long long part_stride; long long parts_rs;   // gate pre-activations [B,4H] x n_parts
  const float* pre; long long pre_rs;                                           // optional extra [B,4H]
  const float* b1; const float* b2;                                             // [4H] optional
  const float* c_prev; long long cprev_rs;
  float* h_out; long long hout_rs;        // post-dropout h (nullable)
  float* h_out2; long long hout2_rs;      // second destination of h (nullable)
  float* c_out; long long cout_rs;        // post-dropout c
  float* gates_save;                      // [B,4H] activated i,f,g,o (nullable)
  float* cpre_save;                       // [B,H] c' before dropout (nullable)
  float* seq_out; long long seq_rs;       // packed mode: out[b] row pointer base (h or 0)
  T2VDrop drop_h, drop_c; unsigned long long drop_base;
  const long long* lens; int t;           // packed mode (nullable lens => every row live)
  int B, H;
  int rnd;                                // round h to tf32 on store (it is a tensor-core GEMM operand)
};
__global__ void lstm_pointwise_fwd_kernel(LstmFwdArgs a) {
  t2v_pdl_trigger();
  t2v_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.H) return;
  const int b = i / a.H, j = i % a.H;
  const bool live = (a.lens == nullptr) || (a.t < a.lens[b]);
  if (!live) {
    if (a.seq_out) a.seq_out[b * a.seq_rs + j] = 0.f;
    if (a.gates_save) {
#pragma unroll
      for (int g = 0; g < 4; ++g) a.gates_save[(long long)b * 4 * a.H + g * a.H + j] = 0.f;
    }
    if (a.cpre_save) a.cpre_save[(long long)b * a.H + j] = 0.f;
    return;
  }
  float g4[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = g * a.H + j;
    float v = 0.f;
    for (int p = 0; p < a.n_parts; ++p) v += a.parts[p * a.part_stride + b * a.parts_rs + col];
    if (a.pre) v += a.pre[b * a.pre_rs + col];
    if (a.b1) v += a.b1[col];
    if (a.b2) v += a.b2[col];
    g4[g] = v;
  }
  const float ig = t2v_sigmoid_fast(g4[0]), fg = t2v_sigmoid_fast(g4[1]), gg = t2v_tanh(g4[2]), og = t2v_sigmoid_fast(g4[3]);
  const float cp = a.c_prev[b * a.cprev_rs + j];
  const float c2 = fg * cp + ig * gg;
  const float h2 = og * t2v_tanh(c2);
  const uint64_t idx = a.drop_base + (uint64_t)b * a.H + j;
  const float hd = t2v_rnd(h2 * t2v_keep_scale(a.drop_h, idx), a.rnd);
  const float cd = c2 * t2v_keep_scale(a.drop_c, idx);
  if (a.h_out) a.h_out[b * a.hout_rs + j] = hd;
  if (a.h_out2) a.h_out2[b * a.hout2_rs + j] = hd;
  a.c_out[b * a.cout_rs + j] = cd;
  if (a.seq_out) a.seq_out[b * a.seq_rs + j] = hd;
  if (a.gates_save) {
    float* gs = a.gates_save + (long long)b * 4 * a.H + j;
    __stcs(gs, ig); __stcs(gs + a.H, fg); __stcs(gs + 2 * a.H, gg); __stcs(gs + 3 * a.H, og);
  }
  if (a.cpre_save) __stcs(a.cpre_save + (long long)b * a.H + j, c2);
}

struct LstmBwdArgs {
  const float* dh1; long long dh1_rs;     // grads wrt post-dropout h (up to 3 sources, nullable)
  int dh1_parts; long long dh1_pstride;   // dh1 may be a sum of split-K partial buffers
  const float* dh2; long long dh2_rs;
  const float* dh3; long long dh3_rs;
  float* dc;                              // [B,H] in: grad wrt post-dropout c ; out: grad wrt c_prev (in place)
  const float* gates_save; const float* cpre_save;
  const float* c_prev; long long cprev_rs;
  float* dgates; long long dg_rs;         // [B,4H] pre-activation grads
  T2VDrop drop_h, drop_c; unsigned long long drop_base;
  const long long* lens; int t;
  int B, H;
  int rnd;
};
__global__ void lstm_pointwise_bwd_kernel(LstmBwdArgs a) {
  t2v_pdl_trigger();
  t2v_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.H) return;
  const int b = i / a.H, j = i % a.H;
  float* dg = a.dgates + b * a.dg_rs + j;
  const bool live = (a.lens == nullptr) || (a.t < a.lens[b]);
  if (!live) {   // dead step of a packed sequence: no dependence on anything
    dg[0] = 0.f; dg[a.H] = 0.f; dg[2 * a.H] = 0.f; dg[3 * a.H] = 0.f;
    return;      // dc (state gradient) passes through unchanged
  }
  float dh = 0.f;
  if (a.dh1) {
    for (int pp = 0; pp < a.dh1_parts; ++pp) dh += a.dh1[pp * a.dh1_pstride + b * a.dh1_rs + j];
  }
  if (a.dh2) dh += a.dh2[b * a.dh2_rs + j];
  if (a.dh3) dh += a.dh3[b * a.dh3_rs + j];
  const uint64_t idx = a.drop_base + (uint64_t)b * a.H + j;
  dh *= t2v_keep_scale(a.drop_h, idx);
  const float* gs = a.gates_save + (long long)b * 4 * a.H + j;
  const float ig = gs[0], fg = gs[a.H], gg = gs[2 * a.H], og = gs[3 * a.H];
  const float c2 = a.cpre_save[(long long)b * a.H + j];
  const float tc = t2v_tanh(c2);
  float dc = a.dc[(long long)b * a.H + j] * t2v_keep_scale(a.drop_c, idx) + dh * og * (1.f - tc * tc);
  const float cp = a.c_prev[b * a.cprev_rs + j];
  dg[0] = t2v_rnd(dc * gg * ig * (1.f - ig), a.rnd);
  dg[a.H] = t2v_rnd(dc * cp * fg * (1.f - fg), a.rnd);
  dg[2 * a.H] = t2v_rnd(dc * ig * (1.f - gg * gg), a.rnd);
  dg[3 * a.H] = t2v_rnd(dh * tc * og * (1.f - og), a.rnd);
  a.dc[(long long)b * a.H + j] = dc * fg;
}

// GRU cell (modules.py:60-62 / nn.GRU): gi = x W_ih^T (+b_ih), gh = h W_hh^T (+b_hh) given as [B,3H] each
__global__ void gru_pointwise_fwd_kernel(const float* __restrict__ gi, long long gi_rs, const float* __restrict__ gh,
                                         const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                         const float* __restrict__ h_prev, float* __restrict__ h_out,
                                         float* __restrict__ save /* [B,4H]: r,z,n,ghn */, int B, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i % H;
  const float* gib = gi + (long long)b * gi_rs;
  const float* ghb = gh + (long long)b * 3 * H;
  const float r = t2v_sigmoid(gib[j] + b_ih[j] + ghb[j] + b_hh[j]);
  const float z = t2v_sigmoid(gib[H + j] + b_ih[H + j] + ghb[H + j] + b_hh[H + j]);
  const float ghn = ghb[2 * H + j] + b_hh[2 * H + j];
  const float n = tanhf(gib[2 * H + j] + b_ih[2 * H + j] + r * ghn);
  const float hp = h_prev[i];
  h_out[i] = (1.f - z) * n + z * hp;
  float* s = save + (long long)b * 4 * H + j;
  s[0] = r; s[H] = z; s[2 * H] = n; s[3 * H] = ghn;
}
// in: dh [B,H] (grad wrt h_out).  out: dgi [B,3H], dgh [B,3H], dh_prev_direct [B,H] (= dh*z; caller adds dgh*W_hh)
__global__ void gru_pointwise_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ save,
                                         const float* __restrict__ h_prev, float* __restrict__ dgi, long long dgi_rs,
                                         float* __restrict__ dgh, float* __restrict__ dh_prev, int B, int H, int rnd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i % H;
  const float* s = save + (long long)b * 4 * H + j;
  const float r = s[0], z = s[H], n = s[2 * H], ghn = s[3 * H];
  const float d = dh[i];
  const float hp = h_prev[i];
  const float dn = d * (1.f - z) * (1.f - n * n);
  const float dz = d * (hp - n) * z * (1.f - z);
  const float dr = dn * ghn * r * (1.f - r);
  float* a = dgi + (long long)b * dgi_rs + j;
  float* c = dgh + (long long)b * 3 * H + j;
  a[0] = t2v_rnd(dr, rnd); a[H] = t2v_rnd(dz, rnd); a[2 * H] = t2v_rnd(dn, rnd);
  c[0] = dr; c[H] = dz; c[2 * H] = dn * r;
  dh_prev[i] = d * z;
}

// VAE head (modules.py:16-22): z = mu + eps*exp(.5*logvar) (training) ; mulv = [mu | logvar] rows of 64
__global__ void vae_reparam_fwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                       float* __restrict__ z, int B, int Z, int training) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Z) return;
  const int b = i / Z, j = i % Z;
  const float mu = mulv[b * 2 * Z + j], lv = mulv[b * 2 * Z + Z + j];
  z[i] = training ? (eps ? eps[i] : 0.f) * expf(0.5f * lv) + mu : mu;
}
// dmulv = [dmu_ext + dz | dlogvar_ext + dz*eps*.5*exp(.5 lv)]
__global__ void vae_reparam_bwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                       const float* __restrict__ dz, const float* __restrict__ dmu_ext,
                                       const float* __restrict__ dlv_ext, float* __restrict__ dmulv, int B, int Z,
                                       int training) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Z) return;
  const int b = i / Z, j = i % Z;
  const float lv = mulv[b * 2 * Z + Z + j];
  const float e = (training && eps) ? eps[i] : 0.f;
  dmulv[b * 2 * Z + j] = dz[i] + (dmu_ext ? dmu_ext[i] : 0.f);
  dmulv[b * 2 * Z + Z + j] = (training ? dz[i] * e * 0.5f * expf(0.5f * lv) : 0.f) + (dlv_ext ? dlv_ext[i] : 0.f);
}

// lane layout of the vectorised BatchNorm kernels: LX = 2^lx lanes of four channels (<= 32), rows per block of the reductions
inline int lanes_log2(int C) { int lx = 0; while ((4 << lx) < C && lx < 5) ++lx; return lx; }
inline int rpb4(int lx) { return (256 >> lx) * 32; }
inline RowSpace mk_rs(int period, int lo, int hi) { RowSpace r; r.period = period; r.lo = lo; r.hi = hi; return r; }
inline T2VDrop mk_drop(const float* mask, unsigned long long seed, unsigned int site, float p) {
  T2VDrop d; d.mask = mask; d.seed = seed; d.site = site; d.p = p; return d;
}
inline BnCtx mk_ctx(const float* y, const float* mean, const float* invstd, const float* gamma, const float* beta,
                    int act, T2VDrop d, int T) {
  BnCtx c; c.y = y; c.mean = mean; c.invstd = invstd; c.gamma = gamma; c.beta = beta; c.act = act; c.drop = d; c.T = T;
  return c;
}
inline unsigned grid1d(long long n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

#define LAUNCH_END() do { T2V_COUNT_LAUNCH(); T2V_LAUNCH_CHECK(); return 0; } while (0)

T2V_API int t2v_embedding_fwd(const long long* ids, const float* table, float* out, int B, int T, int C, int n_symbols,
                              int rnd, cudaStream_t st) {
  T2V_ARG_CHECK(C % 4 == 0 && B > 0 && T > 0, "shape");
  embedding_fwd_kernel<<<B * T, 128, 0, st>>>(ids, table, out, B, T, C, n_symbols, rnd);
  LAUNCH_END();
}
T2V_API int t2v_embedding_bwd(const long long* ids, const float* dout, float* dtable, int B, int T, int C, cudaStream_t st) {
  embedding_bwd_kernel<<<B * T, 128, 0, st>>>(ids, dout, dtable, B, T, C);
  LAUNCH_END();
}

// column sums over valid rows: mode 0 -> (sum, sumsq) ; mode 2 -> (sum).  out buffers are double[C], pre-zeroed by caller.
T2V_API int t2v_col_stats(const float* x, long long rows, int C, int period, int lo, int hi, int mode, double* out0,
                          double* out1, cudaStream_t st) {
  T2V_ARG_CHECK(mode == 0 || mode == 2, "mode");
  const int rpb = 256;
  dim3 grid(t2v_ceil_div(rows, rpb), t2v_ceil_div(C, 32)), block(32, 8);
  BnCtx ctx; memset(&ctx, 0, sizeof(ctx));
  if (C % 4 == 0 && (((uintptr_t)x) & 15) == 0 && rows < (1LL << 31)) {
    const int lx = lanes_log2(C);
    dim3 grid4(t2v_ceil_div(rows, rpb4(lx)), t2v_ceil_div(C, 4 << lx));
    if (mode == 0) colreduce4_kernel<0><<<grid4, 256, 0, st>>>(x, rows, C, mk_rs(period, lo, hi), rpb4(lx), ctx, out0, out1, lx);
    else colreduce4_kernel<2><<<grid4, 256, 0, st>>>(x, rows, C, mk_rs(period, lo, hi), rpb4(lx), ctx, out0, out1, lx);
    LAUNCH_END();
  }
  if (mode == 0) colreduce_kernel<0><<<grid, block, 0, st>>>(x, rows, C, mk_rs(period, lo, hi), rpb, ctx, out0, out1);
  else colreduce_kernel<2><<<grid, block, 0, st>>>(x, rows, C, mk_rs(period, lo, hi), rpb, ctx, out0, out1);
  LAUNCH_END();
}
T2V_API int t2v_bn_finalize(const double* sum, const double* sumsq, double n, int C, float eps, float momentum,
                            float* mean, float* invstd, float* running_mean, float* running_var,
                            long long* num_batches_tracked, cudaStream_t st) {
  bn_finalize_kernel<<<t2v_ceil_div(C, 128), 128, 0, st>>>(sum, sumsq, n, C, eps, momentum, mean, invstd, running_mean,
                                                         running_var, num_batches_tracked);
  LAUNCH_END();
}
T2V_API int t2v_bn_eval_prepare(const float* running_mean, const float* running_var, int C, float eps, float* mean,
                                float* invstd, cudaStream_t st) {
  bn_eval_prepare_kernel<<<t2v_ceil_div(C, 128), 128, 0, st>>>(running_mean, running_var, C, eps, mean, invstd);
  LAUNCH_END();
}
T2V_API int t2v_bn_act_fwd(const float* y, float* out, float* out_lo, long long rows, int C, int period, int lo, int hi,
                           const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                           const float* drop_mask, unsigned long long seed, unsigned int site, float p, int T,
                           int rnd, void* out_hi16, void* out_lo16, cudaStream_t st) {
  T2V_ARG_CHECK(C % 4 == 0 && rows < (1LL << 31), "C must be a multiple of 4, rows < 2^31");
  T2V_ARG_CHECK((out_hi16 == nullptr) == (out_lo16 == nullptr), "fp16 split outputs come as a pair");
  T2V_ARG_CHECK(!out_hi16 || ((((uintptr_t)out_hi16) & 7) == 0 && (((uintptr_t)out_lo16) & 7) == 0), "fp16 outputs: 8-byte aligned");
  BnCtx ctx = mk_ctx(y, mean, invstd, gamma, beta, act, mk_drop(drop_mask, seed, site, p), T);
  const int lx = lanes_log2(C), rpb = (256 >> lx) * 8;
  dim3 grid(t2v_ceil_div(rows, rpb), t2v_ceil_div(C, 4 << lx));
  bn_act_fwd_kernel<<<grid, 256, 0, st>>>(y, out, out_lo, rows, C, mk_rs(period, lo, hi), ctx, rnd, rpb, lx,
                                          reinterpret_cast<uint16_t*>(out_hi16), reinterpret_cast<uint16_t*>(out_lo16));
  LAUNCH_END();
}
T2V_API int t2v_split16(const float* x, void* hi, void* lo, long long n, float scale, cudaStream_t st) {
  T2V_ARG_CHECK(x && hi && lo && n > 0 && n % 4 == 0, "n must be a multiple of 4");
  T2V_ARG_CHECK((((uintptr_t)x) & 15) == 0 && (((uintptr_t)hi) & 7) == 0 && (((uintptr_t)lo) & 7) == 0, "alignment");
  const long long n4 = n / 4;
  const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  split16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), n4, scale);
  LAUNCH_END();
}
// pass 1 of BN backward: dbeta_sum / dgamma_sum (double[C], pre-zeroed)
T2V_API int t2v_bn_act_bwd_reduce(const float* dout, const float* y, long long rows, int C, int period, int lo, int hi,
                                  const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                                  const float* drop_mask, unsigned long long seed, unsigned int site, float p, int T,
                                  double* dbeta_sum, double* dgamma_sum, cudaStream_t st) {
  const int rpb = 256;
  dim3 grid(t2v_ceil_div(rows, rpb), t2v_ceil_div(C, 32)), block(32, 8);
  BnCtx ctx = mk_ctx(y, mean, invstd, gamma, beta, act, mk_drop(drop_mask, seed, site, p), T);
  if (C % 4 == 0 && (((uintptr_t)dout) & 15) == 0 && (((uintptr_t)y) & 15) == 0 && rows < (1LL << 31)) {
    const int lx = lanes_log2(C);
    dim3 grid4(t2v_ceil_div(rows, rpb4(lx)), t2v_ceil_div(C, 4 << lx));
    colreduce4_kernel<1><<<grid4, 256, 0, st>>>(dout, rows, C, mk_rs(period, lo, hi), rpb4(lx), ctx, dbeta_sum, dgamma_sum, lx);
    LAUNCH_END();
  }
  colreduce_kernel<1><<<grid, block, 0, st>>>(dout, rows, C, mk_rs(period, lo, hi), rpb, ctx, dbeta_sum, dgamma_sum);
  LAUNCH_END();
}
T2V_API int t2v_bn_act_bwd_apply(const float* dout, const float* y, float* dy, long long rows, int C, int period, int lo,
                                 int hi, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                 int act, const float* drop_mask, unsigned long long seed, unsigned int site, float p,
                                 int T, const double* dbeta_sum, const double* dgamma_sum, double n,
                                 int use_batch_stats, int rnd, cudaStream_t st) {
  T2V_ARG_CHECK(C % 4 == 0 && rows < (1LL << 31), "C must be a multiple of 4, rows < 2^31");
  BnCtx ctx = mk_ctx(y, mean, invstd, gamma, beta, act, mk_drop(drop_mask, seed, site, p), T);
  const int lx = lanes_log2(C), rpb = (256 >> lx) * 8;
  dim3 grid(t2v_ceil_div(rows, rpb), t2v_ceil_div(C, 4 << lx));
  bn_act_bwd_kernel<<<grid, 256, 0, st>>>(dout, dy, rows, C, mk_rs(period, lo, hi), ctx, dbeta_sum, dgamma_sum, n, use_batch_stats,
                                          rnd, rpb, lx);
  LAUNCH_END();
}
T2V_API int t2v_double_to_float(const double* src, float* dst, int n, float beta, cudaStream_t st) {
  double_to_float_acc_kernel<<<t2v_ceil_div(n, 128), 128, 0, st>>>(src, dst, n, beta);
  LAUNCH_END();
}
T2V_API int t2v_copy2d(const float* src, long long s_rs, long long s_cs, float* dst, long long d_rs, long long rows,
                       int cols, float beta, int rnd, cudaStream_t st) {
  copy2d_kernel<<<grid1d(rows * cols, 256), 256, 0, st>>>(src, s_rs, s_cs, dst, d_rs, rows, cols, beta, rnd);
  LAUNCH_END();
}
T2V_API int t2v_transpose(const float* in, long long i_ld, float* out, long long o_ld, long long rows, int cols,
                          int rnd, cudaStream_t st) {
  dim3 grid(t2v_ceil_div(rows, 32), t2v_ceil_div(cols, 32)), block(32, 8);
  T2V_ARG_CHECK(grid.y <= 65535, "cols too large");
  transpose_kernel<<<grid, block, 0, st>>>(in, i_ld, out, o_ld, rows, cols, rnd);
  LAUNCH_END();
}
T2V_API int t2v_conv1d_pack(const float* w, float* out, int Co, int Ci, int K, int flip, int rnd, cudaStream_t st) {
  conv1d_pack_kernel<<<grid1d((long long)Co * Ci * K, 256), 256, 0, st>>>(w, out, Co, Ci, K, flip, rnd);
  LAUNCH_END();
}
T2V_API int t2v_conv1d_unpack_grad(const float* gk, float* gw, int Co, int Ci, int K, float beta, cudaStream_t st) {
  conv1d_unpack_grad_kernel<<<grid1d((long long)Co * Ci * K, 256), 256, 0, st>>>(gk, gw, Co, Ci, K, beta);
  LAUNCH_END();
}
T2V_API int t2v_axpby(const float* x, float a, float* y, float b, long long n, cudaStream_t st) {
  axpby_kernel<<<grid1d(n, 256), 256, 0, st>>>(x, a, y, b, n);
  LAUNCH_END();
}
T2V_API int t2v_bcast_add_rows(float* y, const float* v, long long rows, int C, int rows_per_batch, cudaStream_t st) {
  bcast_add_rows_kernel<<<grid1d(rows * C, 256), 256, 0, st>>>(y, v, rows, C, rows_per_batch);
  LAUNCH_END();
}
T2V_API int t2v_sum_rows_per_batch(const float* x, float* out, int B, int rows_per_batch, int C, float beta,
                                   cudaStream_t st) {
  sum_rows_per_batch_kernel<<<dim3(B, t2v_ceil_div(C, 128)), 128, 0, st>>>(x, out, rows_per_batch, C, beta);
  LAUNCH_END();
}
T2V_API int t2v_sum_parts(const float* parts, int n_parts, long long part_stride, float* dst, long long n, cudaStream_t st) {
  T2V_CUDA_CHECK(t2v_launch(sum_parts_kernel, dim3(grid1d(n, 256)), dim3(256), 0, st, true, 1, parts, n_parts, part_stride, dst, n));
  T2V_COUNT_LAUNCH();
  return 0;
}
T2V_API int t2v_relu_drop_fwd(const float* x, float* out, long long o_rs, long long rows, int C, const float* mask,
                              unsigned long long seed, unsigned int site, float p, unsigned long long idx_base,
                              int rnd, cudaStream_t st) {
  relu_drop_fwd_kernel<<<grid1d(rows * C, 256), 256, 0, st>>>(x, out, o_rs, rows, C, mk_drop(mask, seed, site, p), idx_base, rnd);
  LAUNCH_END();
}
T2V_API int t2v_relu_drop_bwd(const float* x, const float* dout, long long do_rs, float* dx, long long rows, int C,
                              const float* mask, unsigned long long seed, unsigned int site, float p,
                              unsigned long long idx_base, int rnd, cudaStream_t st) {
  relu_drop_bwd_kernel<<<grid1d(rows * C, 256), 256, 0, st>>>(x, dout, do_rs, dx, rows, C, mk_drop(mask, seed, site, p),
                                                             idx_base, rnd);
  LAUNCH_END();
}
T2V_API int t2v_materialize_mask(float* out, long long n, unsigned long long seed, unsigned int site, float p,
                                 unsigned long long idx_base, cudaStream_t st) {
  materialize_mask_kernel<<<grid1d(n, 256), 256, 0, st>>>(out, n, mk_drop(nullptr, seed, site, p), idx_base);
  LAUNCH_END();
}

T2V_API int t2v_lstm_pointwise_fwd(const float* parts, int n_parts, long long part_stride, long long parts_rs,
                                   const float* pre, long long pre_rs, const float* b1, const float* b2,
                                   const float* c_prev, long long cprev_rs, float* h_out, long long hout_rs,
                                   float* h_out2, long long hout2_rs, float* c_out, long long cout_rs,
                                   float* gates_save, float* cpre_save, float* seq_out, long long seq_rs,
                                   const float* mask_h, const float* mask_c, unsigned long long seed,
                                   unsigned int site_h, unsigned int site_c, float p, unsigned long long drop_base,
                                   const long long* lens, int t, int B, int H, int rnd, cudaStream_t st) {
  LstmFwdArgs a;
  a.rnd = rnd;
  a.parts = parts; a.n_parts = n_parts; a.part_stride = part_stride; a.parts_rs = parts_rs;
  a.pre = pre; a.pre_rs = pre_rs; a.b1 = b1; a.b2 = b2; a.c_prev = c_prev; a.cprev_rs = cprev_rs;
  a.h_out = h_out; a.hout_rs = hout_rs; a.h_out2 = h_out2; a.hout2_rs = hout2_rs; a.c_out = c_out; a.cout_rs = cout_rs;
  a.gates_save = gates_save; a.cpre_save = cpre_save; a.seq_out = seq_out; a.seq_rs = seq_rs;
  a.drop_h = mk_drop(mask_h, seed, site_h, p); a.drop_c = mk_drop(mask_c, seed, site_c, p); a.drop_base = drop_base;
  a.lens = lens; a.t = t; a.B = B; a.H = H;
  T2V_CUDA_CHECK(t2v_launch(lstm_pointwise_fwd_kernel, dim3(grid1d((long long)B * H, 256)), dim3(256), 0, st, true, 1, a));
  T2V_COUNT_LAUNCH();
  return 0;
}
T2V_API int t2v_lstm_pointwise_bwd_parts(const float* dh1, long long dh1_rs, int dh1_parts, long long dh1_pstride,
                                         const float* dh2, long long dh2_rs, const float* dh3, long long dh3_rs, float* dc,
                                         const float* gates_save, const float* cpre_save, const float* c_prev,
                                         long long cprev_rs, float* dgates, long long dg_rs, const float* mask_h,
                                         const float* mask_c, unsigned long long seed, unsigned int site_h,
                                         unsigned int site_c, float p, unsigned long long drop_base, const long long* lens,
                                         int t, int B, int H, int rnd, cudaStream_t st) {
  LstmBwdArgs a;
  a.rnd = rnd;
  a.dh1 = dh1; a.dh1_rs = dh1_rs; a.dh1_parts = dh1_parts; a.dh1_pstride = dh1_pstride; a.dh2 = dh2; a.dh2_rs = dh2_rs;
  a.dh3 = dh3; a.dh3_rs = dh3_rs; a.dc = dc;
  a.gates_save = gates_save; a.cpre_save = cpre_save; a.c_prev = c_prev; a.cprev_rs = cprev_rs; a.dgates = dgates;
  a.dg_rs = dg_rs; a.drop_h = mk_drop(mask_h, seed, site_h, p); a.drop_c = mk_drop(mask_c, seed, site_c, p);
  a.drop_base = drop_base; a.lens = lens; a.t = t; a.B = B; a.H = H;
  T2V_CUDA_CHECK(t2v_launch(lstm_pointwise_bwd_kernel, dim3(grid1d((long long)B * H, 256)), dim3(256), 0, st, true, 1, a));
  T2V_COUNT_LAUNCH();
  return 0;
}
T2V_API int t2v_lstm_pointwise_bwd(const float* dh1, long long dh1_rs, const float* dh2, long long dh2_rs,
                                   const float* dh3, long long dh3_rs, float* dc, const float* gates_save,
                                   const float* cpre_save, const float* c_prev, long long cprev_rs, float* dgates,
                                   long long dg_rs, const float* mask_h, const float* mask_c, unsigned long long seed,
                                   unsigned int site_h, unsigned int site_c, float p, unsigned long long drop_base,
                                   const long long* lens, int t, int B, int H, int rnd, cudaStream_t st) {
  LstmBwdArgs a;
  a.rnd = rnd;
  a.dh1 = dh1; a.dh1_rs = dh1_rs; a.dh1_parts = 1; a.dh1_pstride = 0; a.dh2 = dh2; a.dh2_rs = dh2_rs; a.dh3 = dh3; a.dh3_rs = dh3_rs; a.dc = dc;
  a.gates_save = gates_save; a.cpre_save = cpre_save; a.c_prev = c_prev; a.cprev_rs = cprev_rs; a.dgates = dgates;
  a.dg_rs = dg_rs; a.drop_h = mk_drop(mask_h, seed, site_h, p); a.drop_c = mk_drop(mask_c, seed, site_c, p);
  a.drop_base = drop_base; a.lens = lens; a.t = t; a.B = B; a.H = H;
  T2V_CUDA_CHECK(t2v_launch(lstm_pointwise_bwd_kernel, dim3(grid1d((long long)B * H, 256)), dim3(256), 0, st, true, 1, a));
  T2V_COUNT_LAUNCH();
  return 0;
}
T2V_API int t2v_gru_pointwise_fwd(const float* gi, long long gi_rs, const float* gh, const float* b_ih, const float* b_hh,
                                  const float* h_prev, float* h_out, float* save, int B, int H, cudaStream_t st) {
  gru_pointwise_fwd_kernel<<<grid1d((long long)B * H, 256), 256, 0, st>>>(gi, gi_rs, gh, b_ih, b_hh, h_prev, h_out, save, B, H);
  LAUNCH_END();
}
T2V_API int t2v_gru_pointwise_bwd(const float* dh, const float* save, const float* h_prev, float* dgi, long long dgi_rs,
                                  float* dgh, float* dh_prev, int B, int H, int rnd, cudaStream_t st) {
  gru_pointwise_bwd_kernel<<<grid1d((long long)B * H, 256), 256, 0, st>>>(dh, save, h_prev, dgi, dgi_rs, dgh, dh_prev, B, H, rnd);
  LAUNCH_END();
}
T2V_API int t2v_vae_reparam_fwd(const float* mulv, const float* eps, float* z, int B, int Z, int training, cudaStream_t st) {
  vae_reparam_fwd_kernel<<<grid1d((long long)B * Z, 128), 128, 0, st>>>(mulv, eps, z, B, Z, training);
  LAUNCH_END();
}
T2V_API int t2v_vae_reparam_bwd(const float* mulv, const float* eps, const float* dz, const float* dmu_ext,
                                const float* dlv_ext, float* dmulv, int B, int Z, int training, cudaStream_t st) {
  vae_reparam_bwd_kernel<<<grid1d((long long)B * Z, 128), 128, 0, st>>>(mulv, eps, dz, dmu_ext, dlv_ext, dmulv, B, Z, training);
  LAUNCH_END();
}
