// Fused STFT -> magnitude -> mel filterbank -> log for the reference front-end (TacotronSTFT.mel_spectrogram, layers.py:75-92;
// STFT.transform, stft.py:77-105; dynamic_range_compression, audio_processing.py:77-83), ONE kernel, no intermediate in HBM:
//   wav [B,S] --reflect pad 512 (index arithmetic, no padded copy)--> 1024-sample frames every 256 samples --hann (periodic)-->
//   1024-point FFT in shared memory (radix-4 Stockham, 5 passes; TWO real frames ride one complex transform) --> |X[0..512]|
//   --> mel_basis [n_mel, 513] (each filter is a short contiguous band: only its non-zero range is walked) --> log(max(., clip))
//   --> out [B, n_mel, T].
// The reference computes the same thing as a conv1d with a [1026, 1, 1024] windowed Fourier basis (2.18 MFLOP / frame); the FFT
// needs ~0.03 MFLOP / frame, and the HBM traffic is the algorithmic minimum: 256 new samples (1 KB) in and n_mel floats (320 B) out
// per frame (neighbouring frames re-read their overlap from L1 / L2).
#include "t2v_common.cuh"

namespace {

constexpr int NFFT = 1024, HOP = 256, NB = NFFT / 2 + 1, FPC = 8;     // frames per CTA (4 pairs), one utterance per blockIdx.y
constexpr int MAX_MEL = 128;

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

__global__ void __launch_bounds__(256) stft_mel_fused_kernel(StftArgs p) {
  __shared__ float2 buf[2][NFFT];
  __shared__ float2 tw[NFFT];
  __shared__ float win[NFFT];
  __shared__ float mag[2][NB + 3];
  __shared__ float outs[MAX_MEL][FPC + 1];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int t0 = blockIdx.x * FPC;
  for (int i = tid; i < NFFT; i += 256) { tw[i] = p.twiddle[i]; win[i] = p.window[i]; }
  const float* w = p.wav + (long long)b * p.S;
  __syncthreads();
  for (int pr = 0; pr < FPC / 2; ++pr) {
    const int ta = t0 + 2 * pr, tb = ta + 1;             // frame tb rides the imaginary part
    if (ta >= p.n_frames) break;
    // ---- windowed frames, reflect padding by index (reference: F.pad(..., mode='reflect') by n_fft / 2 on both sides)
    for (int j = tid; j < NFFT; j += 256) {
      int ia = ta * HOP + j - NFFT / 2, ib = ia + HOP;
      ia = ia < 0 ? -ia : (ia >= p.S ? 2 * (p.S - 1) - ia : ia);
      ib = ib < 0 ? -ib : (ib >= p.S ? 2 * (p.S - 1) - ib : ib);
      const float xa = w[ia] * win[j];
      const float xb = (tb < p.n_frames) ? w[ib] * win[j] : 0.f;
      buf[0][j] = make_float2(xa, xb);
    }
    __syncthreads();
    // ---- 1024-point complex FFT: radix-4 Stockham autosort, natural order out, ping-pong between the two buffers
    int src = 0;
#pragma unroll
    for (int Ns = 1; Ns < NFFT; Ns *= 4) {
      const int j = tid;                                  // 256 butterflies per pass, one per thread
      const int k = j & (Ns - 1);
      float2 x0 = buf[src][j], x1 = buf[src][j + 256], x2 = buf[src][j + 512], x3 = buf[src][j + 768];
      if (Ns > 1) {
        const int step = NFFT / (4 * Ns);                 // twiddle exp(-2 pi i r k / (4 Ns))
        x1 = cmul(x1, tw[(k * step) & (NFFT - 1)]);
        x2 = cmul(x2, tw[(2 * k * step) & (NFFT - 1)]);
        x3 = cmul(x3, tw[(3 * k * step) & (NFFT - 1)]);
      }
      const float2 s02 = make_float2(x0.x + x2.x, x0.y + x2.y), d02 = make_float2(x0.x - x2.x, x0.y - x2.y);
      const float2 s13 = make_float2(x1.x + x3.x, x1.y + x3.y), d13 = make_float2(x1.x - x3.x, x1.y - x3.y);
      const int j0 = ((j - k) << 2) + k;
      float2* dst = buf[src ^ 1];
      dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
      dst[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);            // x0 - i x1 - x2 + i x3
      dst[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
      dst[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);        // x0 + i x1 - x2 - i x3
      src ^= 1;
      __syncthreads();
    }
    // ---- separate the two real transforms: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / (2i); magnitudes
    for (int k = tid; k < NB; k += 256) {
      const float2 z = buf[src][k], zc = buf[src][(NFFT - k) & (NFFT - 1)];
      const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
      const float br = 0.5f * (z.y + zc.y), bi = 0.5f * (zc.x - z.x);
      mag[0][k] = sqrtf(ar * ar + ai * ai);
      mag[1][k] = sqrtf(br * br + bi * bi);
    }
    __syncthreads();
    // ---- mel filterbank + log compression: threads 0..n_mel-1 -> frame ta, threads 128.. -> frame tb
    {
      const int f = tid >> 7, m = tid & 127;
      if (m < p.n_mel) {
        const float* mb = p.mel_basis + (long long)m * NB;
        const int lo = p.band_lo[m], hi = p.band_hi[m];
        float acc = 0.f;
        for (int k = lo; k < hi; ++k) acc = fmaf(mb[k], mag[f][k], acc);
        outs[m][2 * pr + f] = logf(fmaxf(acc, p.clip));
      }
    }
    __syncthreads();
  }
  // ---- out[b][m][t0 + f]
  for (int i = tid; i < p.n_mel * FPC; i += 256) {
    const int m = i / FPC, f = i - m * FPC;
    if (t0 + f < p.n_frames) p.out[((long long)b * p.n_mel + m) * p.n_frames + t0 + f] = outs[m][f];
  }
}

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
  dim3 grid(t2v_ceil_div(n_frames, FPC), B);
  stft_mel_fused_kernel<<<grid, 256, 0, stream>>>(a);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
