// Fused STFT -> magnitude -> mel filterbank -> log for the reference front-end (TacotronSTFT.mel_spectrogram, layers.py:75-92;
// STFT.transform, stft.py:77-105; dynamic_range_compression, audio_processing.py:77-83), ONE kernel, no intermediate in HBM:
//   wav [B,S] --reflect pad 512 (index arithmetic, no padded copy)--> 1024-sample frames every 256 samples --hann (periodic)-->
//   1024-point FFT (TWO real frames ride one complex transform) --> |X[0..512]| --> mel_basis [n_mel, 513] (each filter is a short
//   contiguous band: only its non-zero range is walked) --> log(max(., clip)) --> out [B, n_mel, T].
// The reference computes the same thing as a conv1d with a [1026, 1, 1024] windowed Fourier basis (2.18 MFLOP / frame); the FFT
// needs ~0.03 MFLOP / frame, and the HBM traffic is the algorithmic minimum: 256 new samples (1 KB) in and n_mel floats (320 B) out
// per frame (neighbouring frames re-read their overlap from L1 / L2).
//
// FFT organisation (round 2, second version): ONE WARP per complex transform, 1024 = 32 x 32 (Cooley-Tukey, n = l + 32 m,
// k = k1 + 32 k2).  Lane l holds the 32 samples x[l + 32 m] in registers and runs a 32-point FFT over m entirely in registers
// (radix-2, compile-time twiddles), multiplies by W_1024^(l k1), the warp transposes the 32 x 32 tile through its private shared-memory
// pad (conflict-free), and lane k1 runs the second 32-point FFT over l: X[k1 + 32 k2] for all k2.  Two warp-level exchanges instead of
// the five block-wide barrier-separated radix-4 passes of the first version (which was bound by them: 0.024 of the HBM roofline).
#include "t2v_common.cuh"

namespace {

constexpr int NFFT = 1024, HOP = 256, NB = NFFT / 2 + 1;
constexpr int WARPS = 8, FPW = 8, FPC = WARPS * FPW;       // frames per warp (4 pairs) / per CTA (64 consecutive frames of one utterance)
constexpr int MAX_MEL = 128;
constexpr int PAD = 33;                                    // row pitch of the per-warp transpose tile
constexpr int OPAD = FPC + 1;                              // row pitch of the CTA's output tile

struct StftArgs {
  const float* wav; int B, S;
  const float* window;         // [1024] hann, periodic
  const float2* twiddle;       // [1024] exp(-2 pi i m / 1024)
  const float* mel_basis;      // [n_mel, 513]
  const int* band_lo;          // [n_mel] first / one-past-last non-zero bin of each filter
  const int* band_hi;
  float* out;                  // [B, n_mel, n_frames]
  int n_mel, n_frames;
  float clip;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ constexpr int bitrev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// one radix-2 decimation-in-frequency stage of a 32-point FFT held in registers (half-size H); twiddle exponents are compile-time
template <int H>
__device__ __forceinline__ void fft32_stage(float2 (&v)[32]) {
  constexpr float C[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 6.123233996e-17f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
  constexpr float S[16] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f, 1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};
#pragma unroll
  for (int b = 0; b < 32; b += 2 * H) {
#pragma unroll
    for (int j = 0; j < H; ++j) {
      constexpr int step = 16 / H;
      const int tw = j * step;                    // exponent of W_32 = exp(-2 pi i / 32)
      const float2 u = v[b + j], t = v[b + j + H];
      v[b + j] = make_float2(u.x + t.x, u.y + t.y);
      const float2 d = make_float2(u.x - t.x, u.y - t.y);
      if (tw == 0) v[b + j + H] = d;
      else if (tw == 8) v[b + j + H] = make_float2(d.y, -d.x);                 // * (-i)
      else v[b + j + H] = make_float2(d.x * C[tw] + d.y * S[tw], d.y * C[tw] - d.x * S[tw]);     // * (cos - i sin)
    }
  }
}
// in-register 32-point forward FFT; register i ends up holding X[bitrev5(i)]
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
  fft32_stage<16>(v); fft32_stage<8>(v); fft32_stage<4>(v); fft32_stage<2>(v); fft32_stage<1>(v);
}

__global__ void __launch_bounds__(WARPS * 32) stft_mel_fused_kernel(StftArgs p) {
  extern __shared__ __align__(16) float smem[];
  float2* tw2 = reinterpret_cast<float2*>(smem);                     // [32 k1][32 l] = W_1024^(l k1)
  float* win = smem + 2 * 32 * 32;                                    // [1024]
  float* outs = win + NFFT;                                           // [MAX_MEL][OPAD] log-mel of this CTA's frames
  float* tiles = outs + MAX_MEL * OPAD;                                // [WARPS][32 * PAD] transpose tile / magnitudes of a warp
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b = blockIdx.y;
  const int t0 = blockIdx.x * FPC;
  for (int i = tid; i < NFFT; i += WARPS * 32) {
    win[i] = p.window[i];
    const int k1 = i >> 5, l = i & 31;
    tw2[i] = p.twiddle[(l * k1) & (NFFT - 1)];
  }
  const float* w = p.wav + (long long)b * p.S;
  float* tile = tiles + warp * (32 * PAD);
  __syncthreads();
#pragma unroll 1
  for (int pr = 0; pr < FPW / 2; ++pr) {
    const int ta = t0 + warp * FPW + 2 * pr, tb = ta + 1;            // frame tb rides the imaginary part
    if (ta >= p.n_frames) break;                                      // warp-uniform
    const bool has_b = tb < p.n_frames;
    // ---- windowed samples x[lane + 32 m], reflect padding by index (reference: F.pad(..., mode='reflect') by n_fft / 2 on both sides)
    // frame tb starts one hop = 8 rows of 32 samples later: 40 loads per lane cover both frames
    float2 v[32];
    {
      float r[40];
#pragma unroll
      for (int m = 0; m < 40; ++m) {
        int ia = ta * HOP + lane + 32 * m - NFFT / 2;
        ia = ia < 0 ? -ia : (ia >= p.S ? 2 * (p.S - 1) - ia : ia);
        r[m] = (m < 32 || has_b) ? __ldg(w + ia) : 0.f;
      }
#pragma unroll
      for (int m = 0; m < 32; ++m) {
        const float wj = win[lane + 32 * m];
        v[m] = make_float2(r[m] * wj, r[m + 8] * wj);
      }
    }
    // ---- first 32-point FFT over m, twiddle W_1024^(lane k1)
    fft32(v);
    __syncwarp();                                                     // the previous pair's magnitudes have been consumed
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k1 = bitrev5(i);
      v[i] = cmul(v[i], tw2[k1 * 32 + lane]);
    }
    // ---- 32 x 32 transpose through the warp's tile: real parts, then imaginary parts (pitch 33: conflict-free both ways)
    float re[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tile[lane * PAD + bitrev5(i)] = v[i].x;
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 32; ++l) re[l] = tile[l * PAD + lane];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) tile[lane * PAD + bitrev5(i)] = v[i].y;
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 32; ++l) v[l] = make_float2(re[l], tile[l * PAD + lane]);
    __syncwarp();
    // ---- second 32-point FFT over l: register i now holds Z[lane + 32 * bitrev5(i)]
    fft32(v);
    // ---- separate the two real transforms: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / (2i); magnitudes.
    // k = lane + 32 k2 (k2 <= 16); N - k sits in lane (32 - lane) & 31 at k2' = 31 - k2 (lane 0: k2' = (32 - k2) & 31)
    float* mag_a = tile;                // [513] + [513] <= 32 * PAD floats
    float* mag_b = tile + NB + 3;
    const int src_lane = (32 - lane) & 31;
#pragma unroll
    for (int k2 = 0; k2 <= 16; ++k2) {
      const float2 z = v[bitrev5(k2)];
      const float2 other = v[bitrev5(31 - k2)];
      float2 zc = make_float2(__shfl_sync(0xffffffffu, other.x, src_lane), __shfl_sync(0xffffffffu, other.y, src_lane));
      if (lane == 0) zc = v[bitrev5((32 - k2) & 31)];
      const int k = lane + 32 * k2;
      if (k <= NFFT / 2) {
        const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
        const float br = 0.5f * (z.y + zc.y), bi = 0.5f * (zc.x - z.x);
        mag_a[k] = sqrtf(ar * ar + ai * ai);
        mag_b[k] = sqrtf(br * br + bi * bi);
      }
    }
    __syncwarp();
    // ---- mel filterbank + log compression: a lane takes filters lane, 63 - lane (narrow + wide bands balance) and 64 + lane, ...
    const int fa = warp * FPW + 2 * pr;                               // column of frame ta in outs
    for (int q = 0; q * 32 < p.n_mel; ++q) {
      int m = (q == 1) ? (63 - lane) : (q * 32 + lane);
      if (q == 1 && p.n_mel < 64) m = 32 + lane;
      if (m < p.n_mel && m >= 0) {
        const float* mb = p.mel_basis + (long long)m * NB;
        const int lo = p.band_lo[m], hi = p.band_hi[m];
        // four independent accumulator pairs: the loads of a band overlap instead of forming one latency chain
        float aa[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
        int k = lo;
        for (; k + 4 <= hi; k += 4) {
          float wk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) wk[j] = __ldg(mb + k + j);
#pragma unroll
          for (int j = 0; j < 4; ++j) { aa[j] = fmaf(wk[j], mag_a[k + j], aa[j]); ab[j] = fmaf(wk[j], mag_b[k + j], ab[j]); }
        }
        for (; k < hi; ++k) {
          const float wk = __ldg(mb + k);
          aa[0] = fmaf(wk, mag_a[k], aa[0]);
          ab[0] = fmaf(wk, mag_b[k], ab[0]);
        }
        outs[m * OPAD + fa] = logf(fmaxf((aa[0] + aa[1]) + (aa[2] + aa[3]), p.clip));
        outs[m * OPAD + fa + 1] = logf(fmaxf((ab[0] + ab[1]) + (ab[2] + ab[3]), p.clip));
      }
    }
  }
  __syncthreads();
  // ---- out[b][m][t0 + f]: FPC consecutive frames per mel row
  for (int i = tid; i < p.n_mel * FPC; i += WARPS * 32) {
    const int m = i / FPC, f = i - m * FPC;
    if (t0 + f < p.n_frames) p.out[((long long)b * p.n_mel + m) * p.n_frames + t0 + f] = outs[m * OPAD + f];
  }
}

constexpr int STFT_SMEM = (2 * 32 * 32 + NFFT + MAX_MEL * OPAD + WARPS * 32 * PAD) * 4;

}  // namespace

// wav [B,S] fp32 in [-1,1] -> out [B,n_mel,n_frames], n_frames = S / 256 + 1 (filter_length 1024, hop 256, win 1024: the reference
// recipe, hparams.py:33-37).  window [1024], twiddle [1024] complex (exp(-2 pi i m / 1024)), mel_basis [n_mel,513] with the
// non-zero band [band_lo[m], band_hi[m]) of every filter.  S >= 513 (reflect padding).
T2V_API int t2v_stft_mel_fused(const float* wav, int B, int S, const float* window, const float* twiddle, const float* mel_basis,
                               const int* band_lo, const int* band_hi, float* out, int n_mel, int n_frames, float clip,
                               cudaStream_t stream) {
  T2V_ARG_CHECK(wav && window && twiddle && mel_basis && band_lo && band_hi && out, "null pointer");
  T2V_ARG_CHECK(B > 0 && B <= 65535 && S > NFFT / 2 && n_mel > 0 && n_mel <= MAX_MEL, "shape (S must exceed the reflect pad of 512)");
  T2V_ARG_CHECK(n_frames == S / HOP + 1, "n_frames must be S / 256 + 1");
  StftArgs a;
  a.wav = wav; a.B = B; a.S = S; a.window = window; a.twiddle = reinterpret_cast<const float2*>(twiddle); a.mel_basis = mel_basis;
  a.band_lo = band_lo; a.band_hi = band_hi; a.out = out; a.n_mel = n_mel; a.n_frames = n_frames; a.clip = clip;
  static bool attr_set = false;
  if (!attr_set) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STFT_SMEM));
    attr_set = true;
  }
  dim3 grid(t2v_ceil_div(n_frames, FPC), B);
  stft_mel_fused_kernel<<<grid, WARPS * 32, STFT_SMEM, stream>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
