// Remaining memory-bound pieces of the hot path: layout shuttles between the reference's [B,C,T] tensors and the
// internal padded channels-last / time-major rows, output masking (parse_output, model.py:509-520), the
// reference-encoder im2col/col2im with on-the-fly CoordConv channels (CoordConv.py:37-74, modules.py:65-80),
// the VAE loss (loss_function.py:27-45), STFT/mel epilogues (stft.py:97-101, layers.py:88-91,
// audio_processing.py:77-83) and the fused clip+Adam step (train.py:171-172,226-229).
#include "t2v_common.cuh"

namespace {
inline unsigned grid1d(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// out[b, 2+t, c] (=|+=) in[b, c, t]        ([B,C,T] -> padded channels-last)
__global__ void bct_to_padded_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int T, float beta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int c = (int)(i % C);
  const int t = (int)((i / C) % T);
  const int b = (int)(i / ((long long)C * T));
  const float v = in[((long long)b * C + c) * T + t];
  float* o = out + ((long long)b * (T + 4) + 2 + t) * C + c;
  *o = (beta != 0.f ? beta * (*o) : 0.f) + v;
}
// out[b,c,t] = t < len[b] ? in1[b,2+t,c] (+ in2[b,2+t,c]) : fill      (padded channels-last -> [B,C,T], masked)
__global__ void padded_to_bct_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                                     float* __restrict__ out, int B, int C, int T, const long long* __restrict__ lens,
                                     float fill) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int t = (int)(i % T);
  const int c = (int)((i / T) % C);
  const int b = (int)(i / ((long long)C * T));
  float v = fill;
  if (!lens || t < lens[b]) {
    const long long src = ((long long)b * (T + 4) + 2 + t) * C + c;
    v = in1[src] + (in2 ? in2[src] : 0.f);
  }
  out[i] = v;
}
// rows (t*B+b) x ld  ->  padded channels-last [B,T+4,C]  (first C columns)
__global__ void rows_tb_to_padded_kernel(const float* __restrict__ rows, long long ld, float* __restrict__ out, int B,
                                         int C, int T, int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int c = (int)(i % C);
  const int b = (int)((i / C) % B);
  const int t = (int)(i / ((long long)C * B));
  out[((long long)b * (T + 4) + 2 + t) * C + c] = t2v_rnd(rows[((long long)t * B + b) * ld + c], rnd);
}
// drows[(t*B+b)*ld + c] = dpad1[b,2+t,c] + dpad2[b,2+t,c]   (c < C) ; column C = dgate[b,t] ; columns C+1..ld-1 = 0
__global__ void padded_to_rows_tb_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                         const float* __restrict__ dgate, float* __restrict__ rows, long long ld, int B,
                                         int C, int T, int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * ld) return;
  const int c = (int)(i % ld);
  const int b = (int)((i / ld) % B);
  const int t = (int)(i / (ld * B));
  float v = 0.f;
  if (c < C) {
    const long long src = ((long long)b * (T + 4) + 2 + t) * C + c;
    v = p1[src] + (p2 ? p2[src] : 0.f);
  } else if (c == C && dgate) {
    v = dgate[(long long)b * T + t];
  }
  rows[i] = t2v_rnd(v, rnd);
}
// teacher-forcing frames: rows[((t+1)*B+b)*C + c] = tgt[b,c,t] ; rows[0..B) = 0 (go frame, model.py:406-408)
__global__ void bct_to_rows_tb_shift_kernel(const float* __restrict__ tgt, float* __restrict__ rows, int B, int C, int T,
                                            int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(T + 1) * B * C) return;
  const int c = (int)(i % C);
  const int b = (int)((i / C) % B);
  const int t = (int)(i / ((long long)C * B));
  rows[i] = (t == 0) ? 0.f : t2v_rnd(tgt[((long long)b * C + c) * T + (t - 1)], rnd);
}
// gate[b,t] = t < len[b] ? rows[(t*B+b)*ld + col] : fill
__global__ void gate_from_rows_kernel(const float* __restrict__ rows, long long ld, int col, float* __restrict__ gate,
                                      int B, int T, const long long* __restrict__ lens, float fill) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T) return;
  const int t = (int)(i % T), b = (int)(i / T);
  gate[i] = (!lens || t < lens[b]) ? rows[((long long)t * B + b) * ld + col] : fill;
}
// zero rows t >= len[b] of a padded channels-last tensor (the in-place .data mask that makes Postnet conv-0's saved
// input the masked mel, quirk Q10)
__global__ void mask_padded_rows_kernel(float* __restrict__ x, int B, int C, int T, const long long* __restrict__ lens) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int c = (int)(i % C);
  const int t = (int)((i / C) % T);
  const int b = (int)(i / ((long long)C * T));
  if (t >= lens[b]) x[((long long)b * (T + 4) + 2 + t) * C + c] = 0.f;
}

// out[b,t,c] = in[b,2+t,c] + add_vec[b,c]   (padded channels-last -> compact, + per-utterance vector: the style
// broadcast-add of model.py:536-537; add_vec may be NULL)
__global__ void unpad_add_kernel(const float* __restrict__ in, const float* __restrict__ addv, float* __restrict__ out,
                                 int B, int T, int C, int rnd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * C) return;
  const int c = (int)(i % C);
  const int t = (int)((i / C) % T);
  const int b = (int)(i / ((long long)C * T));
  out[i] = t2v_rnd(in[((long long)b * (T + 4) + 2 + t) * C + c] + (addv ? addv[(long long)b * C + c] : 0.f), rnd);
}

// ---------------------------------------------------------------------------------------- reference encoder
// im2col for 3x3 / stride 2 / pad 1 over NHWC; col[(n,ho,wo), (kh*3+kw)*Ct + c].  coord=1: the input has one real
// channel and channels 1..3 are the CoordConv xx/yy/rr planes generated on the fly.
__global__ void im2col_3x3s2_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int H, int W, int Ci,
                                    int Ho, int Wo, int coord, int rnd) {
  const int Ct = coord ? 4 : Ci;
  const long long total = (long long)N * Ho * Wo * 9 * Ct;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Ct);
  const int kk = (int)((i / Ct) % 9);
  const long long pix = i / (9 * Ct);
  const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
  const int h = ho * 2 - 1 + kk / 3, w = wo * 2 - 1 + kk % 3;
  float v = 0.f;
  if (h >= 0 && h < H && w >= 0 && w < W) {
    if (!coord) v = x[(((long long)n * H + h) * W + w) * Ci + c];
    else if (c == 0) v = x[((long long)n * H + h) * W + w];
    else {
      const float xx = ((float)h / (float)(H - 1)) * 2.f - 1.f;
      const float yy = ((float)w / (float)(W - 1)) * 2.f - 1.f;
      v = (c == 1) ? xx : (c == 2 ? yy : sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f)));
    }
  }
  col[i] = t2v_rnd(v, rnd);
}
// Vectorised forms (32-bit index arithmetic, 16-byte accesses): thread = (output pixel, tap, 4 channels).  c4_shift = log2(Ci / 4);
// the CoordConv layer (1 input channel + 3 generated coordinate planes) is the c4_shift = 0 case with one float4 per (pixel, tap).
__global__ void im2col_3x3s2_v4_kernel(const float* __restrict__ x, float4* __restrict__ col, unsigned total, int H, int W, int Ci,
                                       int Ho, int Wo, int c4_shift, int coord, int rnd) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned c4 = i & ((1u << c4_shift) - 1u);
  const unsigned g = i >> c4_shift;
  const unsigned pix = g / 9u, kk = g - pix * 9u;
  const unsigned wo = pix % (unsigned)Wo, t = pix / (unsigned)Wo;
  const unsigned ho = t % (unsigned)Ho, n = t / (unsigned)Ho;
  const int h = (int)ho * 2 - 1 + (int)(kk / 3u), w = (int)wo * 2 - 1 + (int)(kk % 3u);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (h >= 0 && h < H && w >= 0 && w < W) {
    if (!coord) {
      v = __ldg(reinterpret_cast<const float4*>(x + (((long long)n * H + h) * W + w) * Ci) + c4);
    } else {
      const float xx = ((float)h / (float)(H - 1)) * 2.f - 1.f;
      const float yy = ((float)w / (float)(W - 1)) * 2.f - 1.f;
      v = make_float4(__ldg(x + ((long long)n * H + h) * W + w), xx, yy, sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f)));
    }
  }
  col[i] = make_float4(t2v_rnd(v.x, rnd), t2v_rnd(v.y, rnd), t2v_rnd(v.z, rnd), t2v_rnd(v.w, rnd));
}
__global__ void col2im_3x3s2_v4_kernel(const float4* __restrict__ dcol, float4* __restrict__ dx, unsigned total, int H, int W,
                                       int Ho, int Wo, int c4_shift) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned c4 = i & ((1u << c4_shift) - 1u);
  const unsigned g = i >> c4_shift;
  const unsigned w = g % (unsigned)W, t = g / (unsigned)W;
  const unsigned h = t % (unsigned)H, n = t / (unsigned)H;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hh = (int)h + 1 - kh;
    if (hh < 0 || (hh & 1)) continue;
    const int ho = hh >> 1;
    if (ho >= Ho) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ww = (int)w + 1 - kw;
      if (ww < 0 || (ww & 1)) continue;
      const int wo = ww >> 1;
      if (wo >= Wo) continue;
      const float4 d = __ldg(dcol + ((((((long long)n * Ho + ho) * Wo + wo) * 9 + kh * 3 + kw)) << c4_shift) + c4);
      a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
    }
  }
  dx[i] = a;
}
// adjoint: dx[n,h,w,c] = sum over (kh,kw) with matching output pixel of dcol
__global__ void col2im_3x3s2_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int N, int H, int W, int Ci,
                                    int Ho, int Wo) {
  const long long total = (long long)N * H * W * Ci;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Ci);
  const int w = (int)((i / Ci) % W), h = (int)((i / ((long long)Ci * W)) % H), n = (int)(i / ((long long)Ci * W * H));
  float a = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hh = h + 1 - kh;
    if (hh < 0 || (hh & 1)) continue;
    const int ho = hh >> 1;
    if (ho >= Ho) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ww = w + 1 - kw;
      if (ww < 0 || (ww & 1)) continue;
      const int wo = ww >> 1;
      if (wo >= Wo) continue;
      a += dcol[((((long long)n * Ho + ho) * Wo + wo) * 9 + kh * 3 + kw) * Ci + c];
    }
  }
  dx[i] = a;
}

// ---------------------------------------------------------------------------------------- loss
// acc[0] += sum (mel-tgt)^2 ; acc[1] += sum (post-tgt)^2 ; acc[2] += sum bce(gate,gtgt) ; acc[3] += KL sum
__global__ void loss_fwd_kernel(const float* __restrict__ mel, const float* __restrict__ post,
                                const float* __restrict__ tgt, long long n_mel, const float* __restrict__ gate,
                                const float* __restrict__ gtgt, long long n_gate, const float* __restrict__ mu,
                                const float* __restrict__ logvar, long long n_z, double* __restrict__ acc) {
  __shared__ float red[32];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long i = i0; i < n_mel; i += stride) {
    const float t = tgt[i];
    const float d0 = mel[i] - t, d1 = post[i] - t;
    a0 = fmaf(d0, d0, a0);
    a1 = fmaf(d1, d1, a1);
  }
  for (long long i = i0; i < n_gate; i += stride) {
    const float x = gate[i], y = gtgt[i];
    a2 += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));     // BCEWithLogits, numerically stable form
  }
  for (long long i = i0; i < n_z; i += stride) {
    const float m = mu[i], lv = logvar[i];
    a3 += -0.5f * (1.f + lv - m * m - expf(lv));
  }
  a0 = block_sum(a0, red); a1 = block_sum(a1, red); a2 = block_sum(a2, red); a3 = block_sum(a3, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, (double)a0); atomicAdd(acc + 1, (double)a1);
    atomicAdd(acc + 2, (double)a2); atomicAdd(acc + 3, (double)a3);
  }
}
// out = [total, recon, kl]
__global__ void loss_finalize_kernel(const double* __restrict__ acc, double n_mel, double n_gate, float kl_weight,
                                     float* __restrict__ out) {
  const double recon = acc[0] / n_mel + acc[1] / n_mel + acc[2] / n_gate;
  out[0] = (float)(recon + (double)kl_weight * acc[3]);
  out[1] = (float)recon;
  out[2] = (float)acc[3];
}
__global__ void loss_bwd_kernel(const float* __restrict__ mel, const float* __restrict__ post,
                                const float* __restrict__ tgt, long long n_mel, const float* __restrict__ gate,
                                const float* __restrict__ gtgt, long long n_gate, const float* __restrict__ mu,
                                const float* __restrict__ logvar, long long n_z, float kl_weight,
                                const float* __restrict__ gout, float* __restrict__ dmel, float* __restrict__ dpost,
                                float* __restrict__ dgate, float* __restrict__ dmu, float* __restrict__ dlogvar) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float g = gout ? gout[0] : 1.f;
  const float sm = g * 2.f / (float)n_mel, sg = g / (float)n_gate;
  for (long long i = i0; i < n_mel; i += stride) {
    const float t = tgt[i];
    dmel[i] = sm * (mel[i] - t);
    dpost[i] = sm * (post[i] - t);
  }
  for (long long i = i0; i < n_gate; i += stride) dgate[i] = sg * (t2v_sigmoid(gate[i]) - gtgt[i]);
  for (long long i = i0; i < n_z; i += stride) {
    dmu[i] = g * kl_weight * mu[i];
    dlogvar[i] = g * kl_weight * 0.5f * (expf(logvar[i]) - 1.f);
  }
}

// ---------------------------------------------------------------------------------------- STFT / mel
// reflect-pad (stft.py:86-90): out[b, i] = wav[b, reflect(i - pad)], i < S+2*pad ; zeros up to ld
__global__ void reflect_pad_kernel(const float* __restrict__ wav, float* __restrict__ out, int B, int S, int pad, long long ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * ld) return;
  const int b = (int)(i / ld);
  const long long j = i % ld;
  float v = 0.f;
  if (j < S + 2 * pad) {
    long long s = j - pad;
    if (s < 0) s = -s;
    if (s >= S) s = 2 * ((long long)S - 1) - s;
    v = wav[(long long)b * S + s];
  }
  out[i] = v;
}
// mag[r, k] = sqrt(re^2 + im^2), ft rows hold [re(0..nb-1) | im(0..nb-1)]
__global__ void stft_mag_kernel(const float* __restrict__ ft, long long ft_ld, float* __restrict__ mag, long long mag_ld,
                                long long rows, int nb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * mag_ld) return;
  const long long r = i / mag_ld;
  const int k = (int)(i % mag_ld);
  float v = 0.f;
  if (k < nb) {
    const float re = ft[r * ft_ld + k], im = ft[r * ft_ld + nb + k];
    v = sqrtf(re * re + im * im);
  }
  mag[i] = v;
}
// out[b, m, f] = log(max(mel[(b*rows_per_batch + f), m], clip))
__global__ void mel_log_kernel(const float* __restrict__ mel, long long mel_ld, float* __restrict__ out, int B, int n_mel,
                               int n_frames, long long rows_per_batch, float clip) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * n_mel * n_frames) return;
  const int f = (int)(i % n_frames);
  const int m = (int)((i / n_frames) % n_mel);
  const int b = (int)(i / ((long long)n_frames * n_mel));
  out[i] = logf(fmaxf(mel[((long long)b * rows_per_batch + f) * mel_ld + m], clip));
}

// ---------------------------------------------------------------------------------------- optimizer
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float scale, double* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] * scale;
    a = fmaf(v, v, a);
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) atomicAdd(out, (double)a);
}
// grad <- grad*gscale ; clip to max_norm using sumsq (of the scaled grads) ; torch.optim.Adam with L2 weight decay
__global__ void adam_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const double* __restrict__ sumsq, float gscale,
                                 float max_norm, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                                 float bc2, float* __restrict__ norm_out) {
  const float total = (float)sqrt(*sumsq);
  float coef = max_norm / (total + 1e-6f);
  if (coef > 1.f) coef = 1.f;
  if (max_norm <= 0.f) coef = 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = total;
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gr = g[i] * gscale * coef;
    g[i] = gr;
    gr = fmaf(wd, p[i], gr);
    const float mi = beta1 * m[i] + (1.f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
  }
}
// standard normal samples (Box-Muller over the counter-based uniforms): the VAE eps of modules.py:19
__global__ void randn_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned int site) {
  const uint64_t sd = t2v_resolve_seed(seed);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u1 = fmaxf(t2v_uniform(sd, site, 2 * (uint64_t)i), 5.9604645e-8f);
    const float u2 = t2v_uniform(sd, site, 2 * (uint64_t)i + 1);
    out[i] = sqrtf(-2.f * logf(u1)) * cosf(6.2831853071795864f * u2);
  }
}
// dst[off_i : off_i + n_i] = src_i  for a table of segments (gradient packing into one flat buffer)
__global__ void round_tf32_kernel(float* __restrict__ x, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = t2v_tf32(x[i]);
}
__global__ void fill_kernel(float* __restrict__ x, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
}  // namespace

#define LAUNCH_END() do { T2V_COUNT_LAUNCH(); T2V_LAUNCH_CHECK(); return 0; } while (0)

T2V_API int t2v_bct_to_padded(const float* in, float* out, int B, int C, int T, float beta, cudaStream_t st) {
  bct_to_padded_kernel<<<grid1d((long long)B * C * T, 256), 256, 0, st>>>(in, out, B, C, T, beta);
  LAUNCH_END();
}
T2V_API int t2v_padded_to_bct(const float* in1, const float* in2, float* out, int B, int C, int T, const long long* lens,
                              float fill, cudaStream_t st) {
  padded_to_bct_kernel<<<grid1d((long long)B * C * T, 256), 256, 0, st>>>(in1, in2, out, B, C, T, lens, fill);
  LAUNCH_END();
}
T2V_API int t2v_rows_tb_to_padded(const float* rows, long long ld, float* out, int B, int C, int T, int rnd, cudaStream_t st) {
  rows_tb_to_padded_kernel<<<grid1d((long long)B * C * T, 256), 256, 0, st>>>(rows, ld, out, B, C, T, rnd);
  LAUNCH_END();
}
T2V_API int t2v_padded_to_rows_tb(const float* p1, const float* p2, const float* dgate, float* rows, long long ld, int B,
                                  int C, int T, int rnd, cudaStream_t st) {
  padded_to_rows_tb_kernel<<<grid1d((long long)B * T * ld, 256), 256, 0, st>>>(p1, p2, dgate, rows, ld, B, C, T, rnd);
  LAUNCH_END();
}
T2V_API int t2v_bct_to_rows_tb_shift(const float* tgt, float* rows, int B, int C, int T, int rnd, cudaStream_t st) {
  bct_to_rows_tb_shift_kernel<<<grid1d((long long)(T + 1) * B * C, 256), 256, 0, st>>>(tgt, rows, B, C, T, rnd);
  LAUNCH_END();
}
T2V_API int t2v_gate_from_rows(const float* rows, long long ld, int col, float* gate, int B, int T, const long long* lens,
                               float fill, cudaStream_t st) {
  gate_from_rows_kernel<<<grid1d((long long)B * T, 256), 256, 0, st>>>(rows, ld, col, gate, B, T, lens, fill);
  LAUNCH_END();
}
T2V_API int t2v_mask_padded_rows(float* x, int B, int C, int T, const long long* lens, cudaStream_t st) {
  mask_padded_rows_kernel<<<grid1d((long long)B * C * T, 256), 256, 0, st>>>(x, B, C, T, lens);
  LAUNCH_END();
}
T2V_API int t2v_unpad_add(const float* in_padded, const float* add_vec, float* out, int B, int T, int C, int rnd,
                          cudaStream_t st) {
  unpad_add_kernel<<<grid1d((long long)B * T * C, 256), 256, 0, st>>>(in_padded, add_vec, out, B, T, C, rnd);
  LAUNCH_END();
}
T2V_API int t2v_im2col_3x3s2(const float* x, float* col, int N, int H, int W, int Ci, int coord, int rnd, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, Ct = coord ? 4 : Ci;
  {
    const long long tot4 = (long long)N * Ho * Wo * 9 * (Ct / 4);
    int sh = 0;
    while ((4 << sh) < Ct) ++sh;
    const bool pow2 = (4 << sh) == Ct && (coord || Ci % 4 == 0);
    if (pow2 && tot4 < (1LL << 32) - 256 && (((uintptr_t)x | (uintptr_t)col) & 15) == 0) {
      im2col_3x3s2_v4_kernel<<<grid1d(tot4, 256), 256, 0, st>>>(x, reinterpret_cast<float4*>(col), (unsigned)tot4, H, W, Ci, Ho, Wo,
                                                                sh, coord, rnd);
      LAUNCH_END();
    }
  }
  im2col_3x3s2_kernel<<<grid1d((long long)N * Ho * Wo * 9 * Ct, 256), 256, 0, st>>>(x, col, N, H, W, Ci, Ho, Wo, coord, rnd);
  LAUNCH_END();
}
T2V_API int t2v_col2im_3x3s2(const float* dcol, float* dx, int N, int H, int W, int Ci, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  {
    const long long tot4 = (long long)N * H * W * (Ci / 4);
    int sh = 0;
    while ((4 << sh) < Ci) ++sh;
    if ((4 << sh) == Ci && tot4 < (1LL << 32) - 256 && (((uintptr_t)dcol | (uintptr_t)dx) & 15) == 0) {
      col2im_3x3s2_v4_kernel<<<grid1d(tot4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(dcol), reinterpret_cast<float4*>(dx),
                                                                (unsigned)tot4, H, W, Ho, Wo, sh);
      LAUNCH_END();
    }
  }
  col2im_3x3s2_kernel<<<grid1d((long long)N * H * W * Ci, 256), 256, 0, st>>>(dcol, dx, N, H, W, Ci, Ho, Wo);
  LAUNCH_END();
}
// acc: double[4] scratch (zeroed here); out: float[3] = total, recon, kl
T2V_API int t2v_loss_fwd(const float* mel, const float* post, const float* tgt, long long n_mel, const float* gate,
                         const float* gtgt, long long n_gate, const float* mu, const float* logvar, long long n_z,
                         float kl_weight, double* acc, float* out, cudaStream_t st) {
  T2V_CUDA_CHECK(cudaMemsetAsync(acc, 0, 4 * sizeof(double), st));
  loss_fwd_kernel<<<592, 256, 0, st>>>(mel, post, tgt, n_mel, gate, gtgt, n_gate, mu, logvar, n_z, acc);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 1, 0, st>>>(acc, (double)n_mel, (double)n_gate, kl_weight, out);
  LAUNCH_END();
}
T2V_API int t2v_loss_bwd(const float* mel, const float* post, const float* tgt, long long n_mel, const float* gate,
                         const float* gtgt, long long n_gate, const float* mu, const float* logvar, long long n_z,
                         float kl_weight, const float* gout, float* dmel, float* dpost, float* dgate, float* dmu,
                         float* dlogvar, cudaStream_t st) {
  loss_bwd_kernel<<<592, 256, 0, st>>>(mel, post, tgt, n_mel, gate, gtgt, n_gate, mu, logvar, n_z, kl_weight, gout, dmel,
                                       dpost, dgate, dmu, dlogvar);
  LAUNCH_END();
}
T2V_API int t2v_reflect_pad(const float* wav, float* out, int B, int S, int pad, long long ld, cudaStream_t st) {
  T2V_ARG_CHECK(S > pad, "signal shorter than the reflect pad");
  reflect_pad_kernel<<<grid1d((long long)B * ld, 256), 256, 0, st>>>(wav, out, B, S, pad, ld);
  LAUNCH_END();
}
T2V_API int t2v_stft_mag(const float* ft, long long ft_ld, float* mag, long long mag_ld, long long rows, int nb, cudaStream_t st) {
  stft_mag_kernel<<<grid1d(rows * mag_ld, 256), 256, 0, st>>>(ft, ft_ld, mag, mag_ld, rows, nb);
  LAUNCH_END();
}
T2V_API int t2v_mel_log(const float* mel, long long mel_ld, float* out, int B, int n_mel, int n_frames,
                        long long rows_per_batch, float clip, cudaStream_t st) {
  mel_log_kernel<<<grid1d((long long)B * n_mel * n_frames, 256), 256, 0, st>>>(mel, mel_ld, out, B, n_mel, n_frames,
                                                                              rows_per_batch, clip);
  LAUNCH_END();
}
// sumsq: double[1], zeroed here, receives sum (g*gscale)^2
T2V_API int t2v_grad_sumsq(const float* g, long long n, float gscale, double* sumsq, cudaStream_t st) {
  T2V_CUDA_CHECK(cudaMemsetAsync(sumsq, 0, sizeof(double), st));
  sumsq_kernel<<<592, 256, 0, st>>>(g, n, gscale, sumsq);
  LAUNCH_END();
}
T2V_API int t2v_adam_clip_step(float* p, float* g, float* m, float* v, long long n, const double* sumsq, float gscale,
                               float max_norm, float lr, float beta1, float beta2, float eps, float wd, int step,
                               float* norm_out, cudaStream_t st) {
  T2V_ARG_CHECK(step >= 1, "step counts from 1");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_clip_kernel<<<1184, 256, 0, st>>>(p, g, m, v, n, sumsq, gscale, max_norm, lr, beta1, beta2, eps, wd, bc1, bc2, norm_out);
  LAUNCH_END();
}
T2V_API int t2v_randn(float* out, long long n, unsigned long long seed, unsigned int site, cudaStream_t st) {
  randn_kernel<<<t2v_ceil_div(n, 256) < 1184 ? t2v_ceil_div(n, 256) : 1184, 256, 0, st>>>(out, n, seed, site);
  LAUNCH_END();
}
T2V_API int t2v_round_tf32(float* x, long long n, cudaStream_t st) {
  round_tf32_kernel<<<1184, 256, 0, st>>>(x, n);
  LAUNCH_END();
}
T2V_API int t2v_fill(float* x, long long n, float v, cudaStream_t st) {
  fill_kernel<<<592, 256, 0, st>>>(x, n, v);
  LAUNCH_END();
}
