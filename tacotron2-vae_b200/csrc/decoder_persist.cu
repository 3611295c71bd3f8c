// Persistent teacher-forced decoder loop (reference Decoder.forward / Decoder.decode, model.py:346-426) as ONE kernel:
// 128 CTAs (32 clusters of 4, one CTA per SM) stay resident for all To steps; nothing is launched per step.
//
//   * weights never depend on the recurrence, so a producer thread per CTA streams this CTA's slice of
//     [W_attention_rnn | W_decoder_rnn] (fp32 storage, tf32 math) through a TMA ring continuously, running ahead of the
//     recurrence by the ring depth; the activation operand follows through a second ring as soon as the device-wide
//     counters say that h_att / ctx / h_dec of the step exist.
//   * GEMM decomposition: cluster c owns hidden units [32c, 32c+32) (the 128 gate rows i,f,g,o of those units = the UMMA
//     M dimension, weights are the A operand), the batch (<= 64) is the UMMA N dimension, and the 4 CTAs of the cluster
//     split K.  The four partial accumulators (TMEM) are exchanged through distributed shared memory (each CTA ends up
//     with 16 batch rows x 32 units x 4 gates), summed, and the LSTM cell (+ dropout on h and c, model.py:361-364) runs on
//     them directly: gates never touch global memory except as the saved activations of the backward pass.
//   * the query projection is folded into the cell epilogue (per-cluster partial sums over its 32 units, summed by the
//     consumer), the location term of the attention (conv 2->32 k31 + dense 32->128 + processed memory) is computed
//     while the GEMM of the same step runs, and energies / softmax / context follow as soon as h_att is complete:
//     two CTAs per utterance, energies exchanged through DSMEM, context columns split between them.
//   * the decoder_rnn GEMM of step t-1 is interleaved behind the attention_rnn GEMM of step t on the same tensor pipe
//     (it feeds nothing back under teacher forcing), so it overlaps the attention phase.
//   * synchronisation: three monotonic device-wide counters (h_att, ctx, h_dec complete for step t) with release /
//     acquire semantics; mbarriers inside the CTA / cluster.  Every wait is bounded (trap instead of hanging the GPU).
#include "t2v_common.cuh"
#include <stdlib.h>
#include <type_traits>
#include "persist_common.cuh"
#include "gemm_tc.h"
#include "../../include/t2v_b200.h"

namespace {

constexpr int H = 1024, XA_W = 1792, XD_W = 2560, AD = 128, ED = 512, PD = 256;
constexpr int NF = 32, KS = 31, HALO = 15;
constexpr unsigned SITE_ATT_H = 10, SITE_ATT_C = 11, SITE_DEC_H = 12, SITE_DEC_C = 13;
constexpr int CL = 4, NCLUSTER = 32, NCTA = CL * NCLUSTER;
constexpr int NTHREADS = 512;
constexpr int W_STAGE = 128 * 128, A_STAGE = 64 * 128;   // bytes of one ring stage: weights [128 gate rows][128 B of K], activations [64][128 B]
// K chunks per CTA (one 128-byte swizzle row of K per chunk): fp32 storage = 32 columns -> (64 + 128 + 256) / 32 = 14 and
// (256 + 128 + 256) / 32 = 20 chunks; 16-bit operands = 64 columns -> 7 and 10 chunks of the same 16 KB + 8 KB
template <int OP> struct Chunks { static constexpr int ATT = OP ? 7 : 14, DEC = OP ? 10 : 20, CK = OP ? 64 : 32; };
constexpr int ATT_CHUNKS = 14, DEC_CHUNKS = 20;
constexpr int SLOT = 4 * 16 * 32;              // floats of one exchange slot: [gate][batch row of the owner][unit]
constexpr int TH_MAX = 64;                     // text positions per CTA (two CTAs per utterance) -> Ti <= 128
constexpr int PADW = 160;                      // alignment / cumulative-alignment rows with a 15-wide zero halo
constexpr int BAR_EPI = 1, BAR_ATT = 2;
constexpr int FS = 36;                         // row stride of the fp32 conv output (16-byte aligned rows)

// ---- shared memory carve-up (bytes from the 1024-aligned base), per operand mode.  op16: a stage holds twice the K (64 columns), so
// three stages buffer more of the weight stream than four did; the 24 KB pay for the 16-bit operands of the two small in-kernel UMMAs
// (query projection, location dense) whose fp32 FFMA versions were the busiest shared-memory loops of the step.
template <int OP> struct SM {
  static constexpr int NS = OP ? 3 : 4;                             // ring depth: a stage = weights (16 KB) + activations (8 KB)
  static constexpr int OFF_WRING = 0;
  static constexpr int OFF_ARING = OFF_WRING + NS * W_STAGE;
  static constexpr int OFF_RECV = OFF_ARING + NS * A_STAGE;         // [2 gemms][4 sources][SLOT]
  static constexpr int OFF_S = OFF_RECV + 2 * 4 * SLOT * 4;         // [TH_MAX][128] location term + processed memory
  static constexpr int OFF_F = OFF_S + TH_MAX * AD * 4;             // fp32: [TH_MAX][FS] conv output; aliased: ctx partials [4][256],
  static constexpr int F_BYTES = OP ? 2304 * 4 : TH_MAX * FS * 4;   //       inference scratch (op16: only those: ctx partials [8][256] + [256])
  static constexpr int OFF_WCT = OFF_F + F_BYTES;                   // [62][32]
  static constexpr int OFF_HQ = OFF_WCT + 2 * KS * NF * 4;          // [16][32] h_att of this CTA's rows (query partials)
  static constexpr int OFF_WPAD = OFF_HQ + 16 * 32 * 4;
  static constexpr int OFF_CPAD = OFF_WPAD + PADW * 4;
  static constexpr int OFF_E = OFF_CPAD + PADW * 4;                 // [128] energies
  static constexpr int OFF_Q = OFF_E + 128 * 4;                     // [2][128]
  static constexpr int OFF_BARS = OFF_Q + 256 * 4;
  static constexpr int N_BARS = 2 * NS + 10;
  static constexpr int OFF_TMEM = OFF_BARS + N_BARS * 8;
  // op16 only: K-major 16-bit tiles with rows of 64 bytes in the SWIZZLE_64B layout, written by ordinary threads:
  static constexpr int OFF_WQ16 = (OFF_TMEM + 16 + 1023) / 1024 * 1024;   // query_layer columns of this cluster's units [128 a][32 u]
  static constexpr int OFF_HQ16 = OFF_WQ16 + 128 * 64;                    // this CTA's h_att rows [16 b][32 u]
  static constexpr int OFF_WL16 = OFF_HQ16 + 16 * 64;                     // location_dense weight [128 a][32 c]
  static constexpr int OFF_F16 = OFF_WL16 + 128 * 64;                     // location conv output of this CTA's rows [64 r][32 c]
  static constexpr int SMEM_BYTES = (OP ? OFF_F16 + 64 * 64 : OFF_TMEM + 16) + 1024;
};
static_assert(SM<0>::SMEM_BYTES <= 232448 && SM<1>::SMEM_BYTES <= 232448, "shared memory budget");
struct PersistParams {
  T2VDecoderSeq s;
  int t_begin, t_end;
  unsigned* counters;      // [0] h_att complete, [32] ctx complete, [64] h_dec complete (one 128-byte line each)
  float* qpart;            // [2][NCLUSTER][B][128] per-cluster partial query projections
  int packed;              // tmWa / tmWd describe the re-tiled copies (one contiguous 16 KB box per chunk)
  int wa_hint, wd_hint, mem_hint;   // L2 eviction priority of the weight streams / the encoder memory: 0 normal, 1 evict_last, 2 evict_first
  long long* trace;        // T2V_PERSIST_TRACE: [2 CTAs][TRACE_STEPS][32] clock64 stamps, else nullptr
  long long* gtrace;       // T2V_PERSIST_TRACE: [NCTA][8] %globaltimer stamps (ns, comparable across CTAs) of step TRACE_T0
  // ---- free-running inference (Decoder.inference, model.py:428-464) only
  const float *Wp1, *Wp2, *Wpg, *bpg;   // prenet [256,80], [256,256]; [linear_projection ; gate_layer] [81,1536], bias [81]
  const float* prenet_masks;            // [n,2,B,256] or nullptr (counter RNG)
  float* O;                             // [n,B,84] mel | gate rows
  float* opart;                         // [2][NCLUSTER][B][84] per-cluster partial projections of h_dec
  float* ocpart;                        // [2][B][2][84] per-CTA partial projections of ctx
  float gate_threshold;
  int* n_frames;                        // [B] or nullptr
};
constexpr unsigned SITE_PRENET0 = 3, SITE_PRENET1 = 4;
constexpr unsigned Q_SENTINEL = 0x7FC0DEADu;       // "no value yet" in the query-partial slots: a NaN payload no partial sum can produce
__device__ __forceinline__ float ld_relaxed_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__global__ void fill_u32_kernel(unsigned* __restrict__ x, long long n, unsigned v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
constexpr int TRACE_T0 = 100, TRACE_STEPS = 4, TRACE_CTA_B = 77;

// K offset (columns of the weight matrix = columns of the activation row) of chunk j of this CTA's K slice.
//   attention_rnn row XA[t] = [prenet_t (256) | ctx_{t-1} (512) | h_att_{t-1} (1024)]: prenet chunks first (known long
//   before), then h_att (complete one attention phase earlier), then ctx (the last thing to become ready).
__host__ __device__ __forceinline__ int att_kofs(int j, int rank) {
  if (j < 2) return 64 * rank + 32 * j;
  if (j < 10) return PD + ED + 256 * rank + 32 * (j - 2);
  return PD + 128 * rank + 32 * (j - 10);
}
// 16-bit operands: the same K slices in 64-column chunks: prenet (1), h_att (4), ctx (2)
__host__ __device__ __forceinline__ int att_kofs16(int j, int rank) {
  if (j < 1) return 64 * rank;
  if (j < 5) return PD + ED + 256 * rank + 64 * (j - 1);
  return PD + 128 * rank + 64 * (j - 5);
}
//   decoder_rnn row XD[t] = [h_att_t (1024) | ctx_t (512) | h_dec_{t-1} (1024)]
__host__ __device__ __forceinline__ int dec_kofs(int j, int rank) {
  if (j < 8) return 256 * rank + 32 * j;
  if (j < 12) return H + 128 * rank + 32 * (j - 8);
  return H + ED + 256 * rank + 32 * (j - 12);
}
// 16-bit operands: h_att (4), ctx (2), h_dec (4)
__host__ __device__ __forceinline__ int dec_kofs16(int j, int rank) {
  if (j < 4) return 256 * rank + 64 * j;
  if (j < 6) return H + 128 * rank + 64 * (j - 4);
  return H + ED + 256 * rank + 64 * (j - 6);
}
template <int OP> __device__ __forceinline__ int att_kofs_t(int j, int rank) { return OP ? att_kofs16(j, rank) : att_kofs(j, rank); }
template <int OP> __device__ __forceinline__ int dec_kofs_t(int j, int rank) { return OP ? dec_kofs16(j, rank) : dec_kofs(j, rank); }

// Free-running inference: the prenet of step t needs the mel frame of step t-1, so its chunks come LAST in the attention_rnn GEMM
// (h_att_{t-1}, ctx_{t-1} are older), and the decoder_rnn GEMM of step t runs inside step t: h_dec_{t-1} first, ctx_t last.
// The functions return the chunk index of the teacher-forcing order (tile index of the packed weights); kofs follows from it.
template <int OP> __device__ __forceinline__ int att_chunk_infer(int j) {      // h(8) ctx(4) prenet(2)  |  16-bit: h(4) ctx(2) prenet(1)
  return OP ? (j < 6 ? j + 1 : 0) : (j < 12 ? j + 2 : j - 12);
}
template <int OP> __device__ __forceinline__ int dec_chunk_infer(int j) {      // h_dec(8) h_att(8) ctx(4)  |  16-bit: h_dec(4) h_att(4) ctx(2)
  return OP ? (j < 4 ? j + 6 : j - 4) : (j < 8 ? j + 12 : j - 8);
}

template <bool INFER, int OP>
__global__ void __launch_bounds__(NTHREADS, 1)
dec_persist_fwd_kernel(const __grid_constant__ CUtensorMap tmWa, const __grid_constant__ CUtensorMap tmWd,
                       const __grid_constant__ CUtensorMap tmXA, const __grid_constant__ CUtensorMap tmXD,
                       const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  using L = SM<OP>;
  constexpr int NS = L::NS;
  uint8_t* wring = smem + L::OFF_WRING;
  uint8_t* aring = smem + L::OFF_ARING;
  float* recv = (float*)(smem + L::OFF_RECV);
  float* S = (float*)(smem + L::OFF_S);
  float* fbuf = (float*)(smem + L::OFF_F);
  float* wcT = (float*)(smem + L::OFF_WCT);
  float* hq = (float*)(smem + L::OFF_HQ);
  float* wpad = (float*)(smem + L::OFF_WPAD);
  float* cpad = (float*)(smem + L::OFF_CPAD);
  float* e_s = (float*)(smem + L::OFF_E);
  float* q_s = (float*)(smem + L::OFF_Q);
  uint64_t* bars = (uint64_t*)(smem + L::OFF_BARS);
  uint64_t* full = bars;                  // [NS] both producers arrive (count 2) with their byte counts: ONE wait per chunk
  uint64_t* empty = full + NS;            // [NS] freed by the MMA commit; both producers wait on it
  uint64_t* acc_full = empty + NS;        // [2]
  uint64_t* acc_free = acc_full + 2;      // [2]
  uint64_t* recv_full = acc_free + 2;     // [2]
  uint64_t* e_full = recv_full + 2;       // [1]
  uint64_t* s_full = e_full + 1;          // [1] processed-memory tile landed in S
  uint64_t* q_full = s_full + 1;          // [1] op16: the query-projection MMA of this step has retired
  uint64_t* l_full = q_full + 1;          // [1] op16: the location-dense MMA of this step has retired
  uint16_t* wq16 = (uint16_t*)(smem + L::OFF_WQ16);
  uint16_t* hq16 = (uint16_t*)(smem + L::OFF_HQ16);
  uint16_t* wl16 = (uint16_t*)(smem + L::OFF_WL16);
  uint16_t* f16s = (uint16_t*)(smem + L::OFF_F16);
  uint32_t* tmem_holder = (uint32_t*)(smem + L::OFF_TMEM);

  constexpr int ATT_CHUNKS = Chunks<OP>::ATT, DEC_CHUNKS = Chunks<OP>::DEC;     // (shadow the fp32 counts of the file scope)
  constexpr int JA_H = OP ? 1 : 2, JA_C = OP ? 5 : 10;          // attention_rnn GEMM: first chunk that needs h_att / ctx
  constexpr int JD_C = OP ? 4 : 8, JD_D = OP ? 6 : 12;          // decoder_rnn GEMM: first chunk that needs ctx / h_dec
  const T2VDecoderSeq& s = p.s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int B = s.B, Ti = s.Ti, To = s.To;
  const int tb = p.t_begin, te = p.t_end;
  // rounding grid of the operand copies: tf32 for fp32 storage, the 16-bit format otherwise (ABI `rnd` codes, t2v_common.cuh)
  const int opfmt = OP ? s.op16 : 0;
  const int rnd_op = OP ? (s.op16 == 2 ? 3 : 2) : s.use_tc;
  uint16_t* const xa16 = reinterpret_cast<uint16_t*>(s.XA16);
  uint16_t* const xd16 = reinterpret_cast<uint16_t*>(s.XD16);
  unsigned* cnt_h = p.counters;
  unsigned* cnt_c = p.counters + 32;
  unsigned* cnt_d = p.counters + 64;
  unsigned* cnt_p = p.counters + 96;      // inference: prenet rows of step t complete
  // debug time stamps of a few mid-sequence steps (two CTAs); compiled in, one predictable branch per event when off
  const int trace_slot = (blockIdx.x == 0) ? 0 : ((blockIdx.x == TRACE_CTA_B) ? 1 : -1);
  auto GT = [&](unsigned n, int ev) {
    if (p.gtrace && n == (unsigned)TRACE_T0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      p.gtrace[blockIdx.x * 8 + ev] = (long long)ns;
    }
  };
  auto TR = [&](unsigned n, int ev) {
    if (p.trace && trace_slot >= 0 && n >= (unsigned)TRACE_T0 && n < (unsigned)(TRACE_T0 + TRACE_STEPS))
      p.trace[((long long)trace_slot * TRACE_STEPS + (n - TRACE_T0)) * 32 + ev] = clock64();
  };

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 128);
      mbar_init(&recv_full[i], 1);          // one expect_tx arrive per phase; the data arrives as st.async bytes
    }
    mbar_init(e_full, 1);
    mbar_init(s_full, 1);
    mbar_init(q_full, 1);
    mbar_init(l_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(OP ? 256u : 128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  cluster_sync_all();            // every CTA's barriers exist before any remote arrive / store

  if (warp == 0) {
    // =========================================================================== weight producer
    if (lane == 0) {
      uint32_t iw = 0;
      const uint64_t pol_a = l2_policy(p.wa_hint), pol_d = l2_policy(p.wd_hint);
      auto load_w = [&](const CUtensorMap* tm, int kofs, int tile, int hint, uint64_t pol) {
        const int st = iw % NS;
        const uint32_t ph = (iw / NS) & 1u;
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_expect_tx(&full[st], W_STAGE);
        uint8_t* dst = wring + st * W_STAGE;
        if (p.packed) {                 // tile-contiguous copy: the whole [4 gates x 32 units][32] tile is one box
          if (hint) tma_load_2d_hint(dst, tm, 0, tile * 128, &full[st], pol);
          else tma_load_2d(dst, tm, 0, tile * 128, &full[st]);
          ++iw;
          return;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // gate rows i,f,g,o of this cluster's 32 hidden units -> 128 consecutive tile rows
          if (hint) tma_load_2d_hint(dst + g * 4096, tm, kofs, g * H + 32 * cid, &full[st], pol);
          else tma_load_2d(dst + g * 4096, tm, kofs, g * H + 32 * cid, &full[st]);
        }
        ++iw;
      };
      if (INFER) {
        for (int t = tb; t < te; ++t) {
          for (int j = 0; j < ATT_CHUNKS; ++j) { const int c = att_chunk_infer<OP>(j); load_w(&tmWa, att_kofs_t<OP>(c, rank), (cid * CL + rank) * ATT_CHUNKS + c, p.wa_hint, pol_a); }
          for (int j = 0; j < DEC_CHUNKS; ++j) { const int c = dec_chunk_infer<OP>(j); load_w(&tmWd, dec_kofs_t<OP>(c, rank), (cid * CL + rank) * DEC_CHUNKS + c, p.wd_hint, pol_d); }
        }
      } else {
        for (int t = tb; t <= te; ++t) {
          if (t < te) for (int j = 0; j < ATT_CHUNKS; ++j) { load_w(&tmWa, att_kofs_t<OP>(j, rank), (cid * CL + rank) * ATT_CHUNKS + j, p.wa_hint, pol_a); if (j == 0) TR(t - tb, 24); }
          if (t > tb) for (int j = 0; j < DEC_CHUNKS; ++j) load_w(&tmWd, dec_kofs_t<OP>(j, rank), (cid * CL + rank) * DEC_CHUNKS + j, p.wd_hint, pol_d);
          TR(t - tb, 25);
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== activation producer
    if (lane == 0) {
      uint32_t ia = 0;
      unsigned seen_h = 0, seen_c = 0, seen_d = 0;
      auto need = [&](const unsigned* cnt, unsigned& seen, unsigned target) {
        if (seen >= target) return;
        wait_counter(cnt, target);
        seen = target;
        fence_proxy_async();          // the rows were written with generic-proxy stores, TMA reads them through the async proxy
      };
      auto load_a = [&](const CUtensorMap* tm, int kofs, int row0) {
        const int st = ia % NS;
        const uint32_t ph = (ia / NS) & 1u;
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_expect_tx(&full[st], A_STAGE);
        tma_load_2d(aring + st * A_STAGE, tm, kofs, row0, &full[st]);
        ++ia;
      };
      unsigned seen_p = 0;
      if (INFER) {
        for (int t = tb; t < te; ++t) {
          const unsigned n = (unsigned)(t - tb);
          for (int j = 0; j < ATT_CHUNKS; ++j) {
            if (j < JD_C) need(cnt_h, seen_h, NCTA * n);               // (h: 8 | 4 chunks, ctx: 4 | 2, prenet: 2 | 1)
            else if (j < JD_D) need(cnt_c, seen_c, NCTA * n);
            else need(cnt_p, seen_p, NCTA * (n + 1));
            load_a(&tmXA, att_kofs_t<OP>(att_chunk_infer<OP>(j), rank), t * B);
          }
          for (int j = 0; j < DEC_CHUNKS; ++j) {
            if (j < JD_C) need(cnt_d, seen_d, NCTA * n);               // (h_dec: 8 | 4 chunks, h_att: 8 | 4, ctx: 4 | 2)
            else if (j < 2 * JD_C) need(cnt_h, seen_h, NCTA * (n + 1));
            else need(cnt_c, seen_c, NCTA * (n + 1));
            load_a(&tmXD, dec_kofs_t<OP>(dec_chunk_infer<OP>(j), rank), t * B);
          }
        }
      } else
      for (int t = tb; t <= te; ++t) {
        const unsigned n = (unsigned)(t - tb);
        if (t < te) {
          for (int j = 0; j < ATT_CHUNKS; ++j) {
            if (j >= JA_C) need(cnt_c, seen_c, NCTA * n);
            else if (j >= JA_H) need(cnt_h, seen_h, NCTA * n);
            if (j == JA_H) TR(n, 0);
            if (j == JA_C) { TR(n, 1); GT(n, 2); }
            load_a(&tmXA, att_kofs_t<OP>(j, rank), t * B);
          }
          TR(n, 2);
        }
        if (t > tb) {
          for (int j = 0; j < DEC_CHUNKS; ++j) {
            if (j < JD_C) need(cnt_h, seen_h, NCTA * n);
            else if (j < JD_D) need(cnt_c, seen_c, NCTA * n);
            else need(cnt_d, seen_d, NCTA * (n - 1));
            load_a(&tmXD, dec_kofs_t<OP>(j, rank), (t - 1) * B);
          }
          TR(n, 3);
        }
      }
    }
  } else if (warp == 2) {
    // =========================================================================== MMA issuer (whole warp in the loop, one
    // elected lane issues: keeps every tcgen05 operand in uniform registers)
    {
      // instruction descriptor: D=f32, A=B=tf32, K-major both, N=64 (batch), M=128 (gate rows)
      // (16-bit operands: format 0 = fp16, 1 = bf16; one instruction consumes 32 bytes of K per row either way)
      const uint32_t fmt = OP ? (opfmt == 2 ? 1u : 0u) : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // running ring position; the descriptors of stage s are the stage-0 descriptors + s * (stage bytes >> 4)
      int st = 0;
      uint32_t ph = 0;
      const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(wring));
      const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(aring));
      bool ready = false;                    // phase test of the current stage, started while the previous chunk was issued
      auto gemm = [&](int which, int nch, unsigned idx) {
        mbar_wait(&acc_free[which], (idx & 1u) ^ 1u);      // the epilogue has drained the previous accumulator
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(which * 64);
        for (int j = 0; j < nch; ++j) {
          if (!ready) mbar_wait(&full[st], ph);
          tc_fence_after();
          const int sn = (st + 1 == NS) ? 0 : st + 1;
          const uint32_t pn = (st + 1 == NS) ? (ph ^ 1u) : ph;
          ready = mbar_test_wait(&full[sn], pn);           // overlaps the issue below
          if (elect_one()) {
            if (which == 0 && j == 0) TR(idx, 4);
            if (which == 0 && j == JA_C) TR(idx, 5);
            const uint64_t adesc = adesc0 + (uint64_t)(st * (W_STAGE >> 4));
            const uint64_t bdesc = bdesc0 + (uint64_t)(st * (A_STAGE >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (OP) tc_mma_f16(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
              else tc_mma_tf32(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&empty[st]);
            if (j == nch - 1) {
              tc_commit(&acc_full[which]);
              TR(which ? idx + 1 : idx, which ? 7 : 6);
            }
          }
          __syncwarp();
          st = sn; ph = pn;
        }
      };
      if (INFER) {
        for (int t = tb; t < te; ++t) {
          gemm(0, ATT_CHUNKS, (unsigned)(t - tb));
          gemm(1, DEC_CHUNKS, (unsigned)(t - tb));
        }
      } else {
        for (int t = tb; t <= te; ++t) {
          const unsigned n = (unsigned)(t - tb);
          if (t < te) gemm(0, ATT_CHUNKS, n);
          if (t > tb) gemm(1, DEC_CHUNKS, n - 1);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================================================================== epilogue: exchange + LSTM cells
    const int etid = tid - 128;
    const int q = warp & 3;                 // TMEM lane quarter = gate index of the rows this thread drains
    const int u = lane;                     // hidden unit within the cluster's 32
    const int jg = 32 * cid + u;            // global hidden unit
    const int blq = etid >> 5;              // cell phase: this thread owns batch rows bl = blq + 4*j (j<4) of unit u
    const int rnd = rnd_op;
    uint32_t recv_remote[4], full_remote[2][4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      recv_remote[d] = mapa(smem_u32(recv), (uint32_t)d);
      full_remote[0][d] = mapa(smem_u32(&recv_full[0]), (uint32_t)d);
      full_remote[1][d] = mapa(smem_u32(&recv_full[1]), (uint32_t)d);
    }
    float wq[32];                           // query_layer weight [a = etid][this cluster's 32 units]
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(s.Wq + (long long)etid * H + 32 * cid + i);
      wq[i] = w4.x; wq[i + 1] = w4.y; wq[i + 2] = w4.z; wq[i + 3] = w4.w;
    }
    if (OP) {      // op16: the same weights as the A operand of the per-step query MMA (M = 128 attention dims, K = 32 units)
#pragma unroll
      for (int i = 0; i < 32; ++i) wq16[swz64(etid, i)] = t2v_enc16(wq[i], opfmt);
      fence_proxy_async();
      named_bar(BAR_EPI, 128);
    }
    float bias_a[4], bias_d[4], c_att[4], c_dec[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      bias_a[g] = s.ba1[g * H + jg] + s.ba2[g * H + jg];
      bias_d[g] = s.bd1[g * H + jg] + s.bd2[g * H + jg];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = 16 * rank + blq + 4 * j;
      c_att[j] = (b < B) ? s.CA[((long long)tb * B + b) * H + jg] : 0.f;
      c_dec[j] = (b < B) ? s.CD[((long long)tb * B + b) * H + jg] : 0.f;
    }
    const uint64_t seed = (s.training && !s.drop_masks) ? t2v_resolve_seed(s.seed) : 0ull;
    const float p_att = s.training ? s.p_att : 0.f, p_dec = s.training ? s.p_dec : 0.f;

    auto epilogue = [&](auto which_c, const int ts, const unsigned idx) {
      constexpr int which = decltype(which_c)::value;      // compile-time: keeps the per-GEMM register arrays in registers
      // dropout keep scales of this thread's four (unit, batch row) pairs: independent of the recurrence, computed while the
      // accumulator is still being produced
      const float pdrop = which ? p_dec : p_att;
      const float kscale = 1.f / (1.f - pdrop);
      const float* mk = s.drop_masks ? s.drop_masks + (long long)ts * 4 * B * H + (which ? 2LL * B * H : 0) : nullptr;
      float kh4[4], kc4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = 16 * rank + blq + 4 * j;
        kh4[j] = 1.f; kc4[j] = 1.f;
        if (pdrop > 0.f && b < B) {
          const uint64_t li = (uint64_t)b * H + jg;
          if (mk) { kh4[j] = mk[li] * kscale; kc4[j] = mk[(long long)B * H + li] * kscale; }
          else {
            const uint64_t di = (uint64_t)ts * B * H + li;
            kh4[j] = (t2v_uniform(seed, which ? SITE_DEC_H : SITE_ATT_H, di) >= pdrop ? 1.f : 0.f) * kscale;
            kc4[j] = (t2v_uniform(seed, which ? SITE_DEC_C : SITE_ATT_C, di) >= pdrop ? 1.f : 0.f) * kscale;
          }
        }
      }
      mbar_wait(&acc_full[which], idx & 1u);
      tc_fence_after();
      const unsigned tn = which ? idx + 1 : idx;       // trace row = the step during which this epilogue runs
      if (etid == 0) {
        mbar_expect_tx(&recv_full[which], 4 * SLOT * 4);   // four 8 KB slots (own included) land as st.async bytes
        TR(tn, which ? 14 : 8);
        if (which == 0) GT(tn, 4);
      }
      // ---- drain TMEM: this thread holds D[gate q, unit u][batch 0..63]; batch rows 16d..16d+15 go to cluster rank d.
      // A 4x4 transpose inside every lane quad turns "one unit, 4 batch rows" into "one batch row, 4 units", so the
      // exchange is 16-byte DSMEM stores that cover whole 128-byte rows of the [gate][batch row][unit] slot.
      {
        const int r4 = lane & 3, m4 = lane >> 2;
        const uint32_t slot_off = (uint32_t)(((which * 4 + rank) * 4 + q) * 16 * 32 + 4 * m4) * 4u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(which * 64 + half * 32), v);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float a0 = __uint_as_float(v[4 * k]), a1 = __uint_as_float(v[4 * k + 1]);
            float a2 = __uint_as_float(v[4 * k + 2]), a3 = __uint_as_float(v[4 * k + 3]);
            {
              const bool odd = (r4 & 1) != 0;
              const float y0 = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
              const float y1 = __shfl_xor_sync(0xffffffffu, odd ? a2 : a3, 1);
              if (odd) { a0 = y0; a2 = y1; } else { a1 = y0; a3 = y1; }
            }
            {
              const bool hi = (r4 & 2) != 0;
              const float y0 = __shfl_xor_sync(0xffffffffu, hi ? a0 : a2, 2);
              const float y1 = __shfl_xor_sync(0xffffffffu, hi ? a1 : a3, 2);
              if (hi) { a0 = y0; a1 = y1; } else { a2 = y0; a3 = y1; }
            }
            // now a_j = D[gate q, unit 4*m4 + j][batch half*32 + 4k + r4]
            const int d = half * 2 + (k >> 2);             // destination rank (compile-time), batch row 4*(k&3) + r4 of its 16
            st_async_v4(recv_remote[d] + slot_off + (uint32_t)((4 * (k & 3) + r4) * 32) * 4u, full_remote[which][d], a0, a1, a2, a3);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free[which]);
      if (etid == 0 && which == 0) TR(tn, 9);
      mbar_wait_cluster(&recv_full[which], idx & 1u);
      if (etid == 0 && which == 0) TR(tn, 10);
      // ---- LSTM cell (model.py:357-364 / 375-381) on (unit u, batch rows blq + 4j)
      const float* bias = which ? bias_d : bias_a;
      float* cst = which ? c_dec : c_att;
      const long long r0 = (long long)ts * B, r1 = (long long)(ts + 1) * B;
      float sv_i[4], sv_f[4], sv_g[4], sv_o[4], sv_c2[4];     // saved activations: written after the signal
      float sv_lo[4] = {0.f, 0.f, 0.f, 0.f};                  // h_dec - (h_dec on the operand grid): the split projection's low part
      float sv_hi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bl = blq + 4 * j;
        const int b = 16 * rank + bl;
        float g4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float* rp = recv + (which * 4 * 4 + g) * 16 * 32 + bl * 32 + u;
          g4[g] = ((rp[0] + rp[SLOT]) + (rp[2 * SLOT] + rp[3 * SLOT])) + bias[g];
        }
        const float ig = t2v_sigmoid_fast(g4[0]), fg = t2v_sigmoid_fast(g4[1]), gg = t2v_tanh(g4[2]), og = t2v_sigmoid_fast(g4[3]);
        const float c2 = fg * cst[j] + ig * gg;
        const float h2 = og * t2v_tanh(c2);
        const float kh = kh4[j], kc = kc4[j];
        const float hx = h2 * kh;               // exact h (after dropout); hd = its copy on the operand grid
        const float hd = t2v_rnd(hx, rnd);
        cst[j] = c2 * kc;
        sv_i[j] = ig; sv_f[j] = fg; sv_g[j] = gg; sv_o[j] = og; sv_c2[j] = c2;
        // (inference: the mel / gate projection inside the kernel takes the exact h_dec, like the reference's fp32 Linear)
        if (which == 0 || INFER) hq[bl * 32 + u] = which ? hx : hd;
        if (OP && which == 0) hq16[swz64(bl, u)] = t2v_enc16(hx, opfmt);
        if (b < B) {       // only what other CTAs wait for goes out before the signal
          if (which == 0) {
            s.XA[(r1 + b) * XA_W + (PD + ED) + jg] = hd;       // h_att -> next step's recurrent input
            s.XD[(r0 + b) * XD_W + jg] = hd;                   // h_att -> decoder_rnn input / deferred dW operand
            if (OP) {
              const uint16_t h16 = t2v_enc16(hx, opfmt);
              xa16[(r1 + b) * XA_W + (PD + ED) + jg] = h16;
              xd16[(r0 + b) * XD_W + jg] = h16;
            }
          } else {
            s.XD[(r1 + b) * XD_W + (H + ED) + jg] = hd;        // h_dec -> next step's recurrent input
            if (OP) xd16[(r1 + b) * XD_W + (H + ED) + jg] = t2v_enc16(hx, opfmt);
            sv_lo[j] = hx - hd;
            sv_hi[j] = hd;
          }
        }
      }
      if (etid == 0 && which == 0) TR(tn, 11);
      if (OP && which == 0) {
        // ---- partial query projection over this cluster's 32 units for this CTA's 16 batch rows on the tensor core:
        // D[128 a, 16 b] = Wq16[a][u] * hq16[b][u]  (two K=16 instructions) into TMEM columns 128..143; thread = attention dim a
        fence_proxy_async();                  // the h rows were written with ordinary stores, the MMA reads them through the async proxy
        tc_fence_before();
        named_bar(BAR_EPI, 128);
        tc_fence_after();
        if (warp == 4 && elect_one()) {
          const uint32_t fq = (opfmt == 2) ? 1u : 0u;
          const uint32_t idq = (1u << 4) | (fq << 7) | (fq << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
          const uint64_t ad = make_kmajor_sw64_desc(smem_u32(wq16)), bd = make_kmajor_sw64_desc(smem_u32(hq16));
          tc_mma_f16(tmem_base + 128u, ad, bd, idq, 0u);
          tc_mma_f16(tmem_base + 128u, ad + 2, bd + 2, idq, 1u);
          tc_commit(q_full);
        }
        __syncwarp();
        mbar_wait(q_full, idx & 1u);
        tc_fence_after();
        uint32_t qv[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + 128u, qv);
        float* qp = p.qpart + ((long long)((ts & 1) * NCLUSTER + cid) * B) * AD + etid;
#pragma unroll
        for (int bl = 0; bl < 16; ++bl) {
          const int b = 16 * rank + bl;
          if (b < B) qp[(long long)b * AD] = __uint_as_float(qv[bl]);
        }
        tc_fence_before();
      } else if (which == 0) {
        // ---- partial query projection over this cluster's 32 units for this CTA's 16 batch rows: thread = attention dim a
        named_bar(BAR_EPI, 128);
        float* qp = p.qpart + ((long long)((ts & 1) * NCLUSTER + cid) * B) * AD + etid;
#pragma unroll 4
        for (int bl = 0; bl < 16; ++bl) {
          const float4* hp = reinterpret_cast<const float4*>(hq + bl * 32);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 h4 = hp[i];
            a0 = fmaf(wq[4 * i], h4.x, a0); a1 = fmaf(wq[4 * i + 1], h4.y, a1);
            a2 = fmaf(wq[4 * i + 2], h4.z, a2); a3 = fmaf(wq[4 * i + 3], h4.w, a3);
          }
          const int b = 16 * rank + bl;
          if (b < B) qp[(long long)b * AD] = (a0 + a1) + (a2 + a3);
        }
      }
      if (INFER && which == 1) {
        // ---- inference: partial mel / gate projection of h_dec over this cluster's 32 units for this CTA's 16 batch rows
        // (linear_projection + gate_layer, model.py:383-388): thread = output o (81 of the 128 threads)
        float wp[32];
        if (etid < 81) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.Wpg + (long long)etid * (H + ED) + 32 * cid + i));
            wp[i] = w4.x; wp[i + 1] = w4.y; wp[i + 2] = w4.z; wp[i + 3] = w4.w;
          }
        }
        named_bar(BAR_EPI, 128);
        if (etid < 81) {
          float* op = p.opart + ((long long)((ts & 1) * NCLUSTER + cid) * B) * 84 + etid;
#pragma unroll 4
          for (int bl = 0; bl < 16; ++bl) {
            const float4* hp = reinterpret_cast<const float4*>(hq + bl * 32);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 h4 = hp[i];
              a0 = fmaf(wp[4 * i], h4.x, a0); a1 = fmaf(wp[4 * i + 1], h4.y, a1);
              a2 = fmaf(wp[4 * i + 2], h4.z, a2); a3 = fmaf(wp[4 * i + 3], h4.w, a3);
            }
            const int b = 16 * rank + bl;
            if (b < B) op[(long long)b * 84] = (a0 + a1) + (a2 + a3);
          }
        }
      }
      named_bar(BAR_EPI, 128);
      if (etid == 0 && which == 0) TR(tn, 12);
      if (etid == 0 && which == 0) GT(tn, 3);
      if (etid == 0) signal_counter(which ? cnt_d : cnt_h);
      if (etid == 0) TR(tn, which ? 15 : 13);
      // ---- saved activations of the backward pass + the cell state sequence (nobody inside this kernel reads them)
      {
        float* Gs = which ? s.GD : s.GA;
        float* CPs = which ? s.CPD : s.CPA;
        float* Cs = which ? s.CD : s.CA;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = 16 * rank + blq + 4 * j;
          if (b < B) {
            Cs[(r1 + b) * H + jg] = cst[j];
            if (Gs) {
              float* gs = Gs + (r0 + b) * 4 * H + jg;
              __stcs(gs, sv_i[j]); __stcs(gs + H, sv_f[j]); __stcs(gs + 2 * H, sv_g[j]); __stcs(gs + 3 * H, sv_o[j]);
            }
            if (CPs) __stcs(CPs + (r0 + b) * H + jg, sv_c2[j]);
            if (which == 1 && s.HCLO) {
              __stcs(s.HCHI + (r0 + b) * (H + ED) + jg, sv_hi[j]);
              __stcs(s.HCLO + (r0 + b) * (H + ED) + jg, t2v_tf32(sv_lo[j]));
            }
          }
        }
      }
    };
    if (INFER) {
      for (int t = tb; t < te; ++t) {
        epilogue(std::integral_constant<int, 0>{}, t, (unsigned)(t - tb));
        epilogue(std::integral_constant<int, 1>{}, t, (unsigned)(t - tb));
      }
    } else {
      for (int t = tb; t <= te; ++t) {
        const unsigned n = (unsigned)(t - tb);
        if (t < te) epilogue(std::integral_constant<int, 0>{}, t, n);
        if (t > tb) epilogue(std::integral_constant<int, 1>{}, t - 1, n - 1);
      }
    }
  } else if (warp >= 8) {
    // =========================================================================== attention (two CTAs per utterance)
    const int atid = tid - 256, aw = warp - 8;
    const int b = 2 * cid + (rank >> 1), hh = rank & 1;
    const bool active = b < B;
    const int Th = (Ti + 1) >> 1;
    const int i0 = hh * Th;
    const int nrow = hh ? (Ti - Th) : Th;
    const int rnd = rnd_op;
    int len = Ti;
    if (active && s.in_lens) { const long long l = s.in_lens[b]; len = l < Ti ? (int)l : Ti; }
    const int a = atid & 127;
    float vreg[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) vreg[k] = s.v[lane + 32 * k];
    for (int i = atid; i < 2 * KS * NF; i += 256) wcT[i] = s.Wconv[i];
    if (OP) {      // op16: location_dense weight as the A operand of the per-step dense MMA (M = 128 attention dims, K = 32 filters)
      for (int i = atid; i < AD * NF; i += 256) wl16[swz64(i >> 5, i & 31)] = t2v_enc16(s.Wloc[i], opfmt);
      for (int i = atid; i < 64 * 32; i += 256) f16s[i] = 0;
      fence_proxy_async();
    }
    for (int i = atid; i < PADW; i += 256) {
      const int ti = i - HALO;
      float w0 = 0.f, c0 = 0.f;
      if (active && ti >= 0 && ti < Ti) {
        if (tb > 0) w0 = s.align[((long long)b * To + (tb - 1)) * Ti + ti];
        c0 = s.CUM[((long long)tb * B + b) * Ti + ti];
      }
      wpad[i] = w0; cpad[i] = c0;
    }
    const uint32_t partner_e = mapa(smem_u32(e_s), (uint32_t)(rank ^ 1));
    const uint32_t partner_full = mapa(smem_u32(e_full), (uint32_t)(rank ^ 1));
    named_bar(BAR_ATT, 256);
    const uint64_t pol_m = l2_policy(p.mem_hint);

    // ---- inference only: mel / gate frame f = sum of the per-cluster h_dec partials + the two ctx partials + bias -> O[f],
    // stop bookkeeping (Decoder.inference, model.py:449-459); leaves the frame in mel_s for the prenet
    float* mel_s = fbuf;                      // [84]   (fbuf is free between the location phase and the context phase)
    float* p1_s = fbuf + 128;                 // [256]
    float* ctx_s = fbuf + (OP ? 2048 : 1024); // [256]  (next to the context partial rows)
    auto frame_from_partials = [&](const int f) {
      if (atid < 81) {
        const int o = atid;
        float acc = p.bpg[o];
        const float* op = p.opart + ((long long)((f & 1) * NCLUSTER) * B + b) * 84 + o;
        float pv[NCLUSTER];                   // all 32 partials in flight at once (one L2 round trip)
#pragma unroll
        for (int c = 0; c < NCLUSTER; ++c) pv[c] = __ldcg(op + (long long)c * B * 84);
        const float* oc = p.ocpart + ((long long)((f & 1) * B + b) * 2) * 84 + o;
        const float oc0 = __ldcg(oc), oc1 = __ldcg(oc + 84);
#pragma unroll
        for (int c = 0; c < NCLUSTER; ++c) acc += pv[c];
        acc += oc0 + oc1;
        mel_s[o] = acc;
        if (hh == 0) {
          p.O[((long long)f * B + b) * 84 + o] = acc;
          if (o == 80 && p.n_frames && p.n_frames[b] < 0 && 1.f / (1.f + expf(-acc)) > p.gate_threshold) p.n_frames[b] = f + 1;
        }
      }
    };
    const uint64_t pseed = (INFER && !p.prenet_masks) ? t2v_resolve_seed(s.seed) : 0ull;
    const uint64_t pol_keep = l2_policy(1);

    float qv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = tb; t < te; ++t) {
      const unsigned n = (unsigned)(t - tb);
      if (active) {
        // the partner's energies of this step arrive as st.async bytes on e_full
        if (atid == 0) mbar_expect_tx(e_full, (uint32_t)(hh ? Th : Ti - Th) * 4u);
        // ---- processed-memory rows of this CTA -> S (bulk copy, lands while the conv runs); every reader of the previous
        // step's S passed the barrier at the end of that step
        if (atid == 0 && nrow > 0) {
          mbar_expect_tx(s_full, (uint32_t)nrow * AD * 4u);
          bulk_load_1d(S, s.pmem + ((long long)b * Ti + i0) * AD, (uint32_t)nrow * AD * 4u, s_full);
        }
        float wl[NF];                        // location_dense weight row of attention dim a (re-read per step: the
        if (!OP) {                           // registers are needed by the context phase); op16: the dense runs on the tensor core
#pragma unroll
          for (int c = 0; c < NF; c += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(s.Wloc + a * NF + c));
            wl[c] = t4.x; wl[c + 1] = t4.y; wl[c + 2] = t4.z; wl[c + 3] = t4.w;
          }
        }
        // ---- location conv (2 -> 32, k = 31, zero padding) on this CTA's rows: thread = (filter c, 8 consecutive rows)
        {
          const int c = atid & 31, r0 = (atid >> 5) * 8;
          if (r0 < nrow) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const float* src = (ch ? cpad : wpad) + i0 + r0;
              float xr[8 + KS - 1];
#pragma unroll
              for (int j = 0; j < 8 + KS - 1; ++j) xr[j] = src[j];
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                const float w = wcT[(ch * KS + k) * NF + c];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, xr[j + k], acc[j]);
              }
            }
            if (OP) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f16s[swz64(r0 + j, c)] = t2v_enc16(acc[j], opfmt);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) fbuf[(r0 + j) * FS + c] = acc[j];
            }
          }
        }
        if (OP) {
          // ---- S += location dense on the tensor core: D[128 a, 64 rows] = Wloc16[a][c] * f16s[row][c] (two K=16 instructions)
          // into TMEM columns 144..207; warp w reads attention dims 32 (w & 3) .. +31 of rows 32 (w >> 2) .. +31
          fence_proxy_async();
          tc_fence_before();
          named_bar(BAR_ATT, 256);
          tc_fence_after();
          if (aw == 0 && elect_one()) {
            const uint32_t fq = (opfmt == 2) ? 1u : 0u;
            const uint32_t idl = (1u << 4) | (fq << 7) | (fq << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t ad = make_kmajor_sw64_desc(smem_u32(wl16)), bd = make_kmajor_sw64_desc(smem_u32(f16s));
            tc_mma_f16(tmem_base + 144u, ad, bd, idl, 0u);
            tc_mma_f16(tmem_base + 144u, ad + 2, bd + 2, idl, 1u);
            tc_commit(l_full);
          }
          __syncwarp();
          if (nrow > 0) mbar_wait(s_full, n & 1u);
          mbar_wait(l_full, n & 1u);
          tc_fence_after();
          uint32_t lv[32];
          tmem_ld32(tmem_base + ((uint32_t)((aw & 3) * 32) << 16) + 144u + (uint32_t)((aw >> 2) * 32), lv);
          const int a2 = 32 * (aw & 3) + lane, rb = 32 * (aw >> 2);
#pragma unroll
          for (int rr = 0; rr < 32; ++rr)
            if (rb + rr < nrow) S[(rb + rr) * AD + a2] += __uint_as_float(lv[rr]);
          tc_fence_before();
        } else {
        named_bar(BAR_ATT, 256);
        // ---- S += location dense: thread = (attention dim a, 32 rows)
        if (nrow > 0) mbar_wait(s_full, n & 1u);
        {
          const int rbeg = (atid >> 7) * 32;
#pragma unroll 4
          for (int rr = rbeg; rr < rbeg + 32; ++rr) {
            if (rr < nrow) {
              float s0 = S[rr * AD + a], s1 = 0.f, s2 = 0.f, s3 = 0.f;
              const float4* fr = reinterpret_cast<const float4*>(fbuf + rr * FS);     // broadcast 16-byte loads
#pragma unroll
              for (int c = 0; c < NF; c += 4) {
                const float4 f4 = fr[c >> 2];
                s0 = fmaf(f4.x, wl[c], s0); s1 = fmaf(f4.y, wl[c + 1], s1);
                s2 = fmaf(f4.z, wl[c + 2], s2); s3 = fmaf(f4.w, wl[c + 3], s3);
              }
              S[rr * AD + a] = (s0 + s1) + (s2 + s3);
            }
          }
        }
        }
        named_bar(BAR_ATT, 256);
      }
      if (INFER) {
        // ---- prenet of step t (model.py:91-102; dropout always on) from the mel frame of step t-1: needs every cluster's
        // projection partial (cnt_d) and the partner CTA's ctx partial (cnt_c) of that step
        if (atid == 0 && n > 0) { TR(n, 0); wait_counter(cnt_d, NCTA * n); wait_counter(cnt_c, NCTA * n); TR(n, 1); }
        named_bar(BAR_ATT, 256);
        if (active) {
          if (n > 0) frame_from_partials(t - 1);
          else if (t > 0 && atid < 81) mel_s[atid] = p.O[((long long)(t - 1) * B + b) * 84 + atid];
          named_bar(BAR_ATT, 256);
          float* xa = s.XA + ((long long)t * B + b) * XA_W;
          if (t == 0) {                       // go frame = zeros (model.py:241-247): both prenet layers give exactly 0
            if (atid < 128) { xa[128 * hh + atid] = 0.f; if (OP) xa16[((long long)t * B + b) * XA_W + 128 * hh + atid] = 0; }
          } else {
            const float* pm0 = p.prenet_masks ? p.prenet_masks + ((long long)t * 2 * B + b) * PD : nullptr;
            const float* pm1 = pm0 ? pm0 + (long long)B * PD : nullptr;
            const uint64_t pidx = ((uint64_t)t * B + b) * PD;
            // layer 1 (80 -> 256, both CTAs of the pair): warp per output, lanes 0..19 hold four inputs each
            const float4 m4 = (lane < 20) ? *reinterpret_cast<const float4*>(mel_s + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
              const int j = aw + 8 * k;
              float acc = 0.f;
              if (lane < 20) {
                const float4 w4 = ldg_v4_hint(p.Wp1 + j * 80 + 4 * lane, pol_keep);     // stays in L2 next to the weight streams
                acc = w4.x * m4.x + w4.y * m4.y + w4.z * m4.z + w4.w * m4.w;
              }
              acc = warp_sum(acc);
              if (lane == 0) {
                const float keep = pm0 ? pm0[j] : (t2v_uniform(pseed, SITE_PRENET0, pidx + j) >= 0.5f ? 1.f : 0.f);
                p1_s[j] = fmaxf(acc, 0.f) * keep * 2.f;       // (FFMA mat-vecs: the prenet stays exact fp32)
              }
            }
            named_bar(BAR_ATT, 256);
            // layer 2 (256 -> 256): this CTA's 128 outputs, warp per output, each lane 8 inputs
            const float4 x0 = *reinterpret_cast<const float4*>(p1_s + 4 * lane);
            const float4 x1 = *reinterpret_cast<const float4*>(p1_s + 128 + 4 * lane);
#pragma unroll 8
            for (int k = 0; k < 16; ++k) {
              const int j = 128 * hh + aw + 8 * k;
              const float4 w0 = ldg_v4_hint(p.Wp2 + j * PD + 4 * lane, pol_keep);
              const float4 w1 = ldg_v4_hint(p.Wp2 + j * PD + 128 + 4 * lane, pol_keep);
              float acc = (w0.x * x0.x + w0.y * x0.y + w0.z * x0.z + w0.w * x0.w) + (w1.x * x1.x + w1.y * x1.y + w1.z * x1.z + w1.w * x1.w);
              acc = warp_sum(acc);
              if (lane == 0) {
                const float keep = pm1 ? pm1[j] : (t2v_uniform(pseed, SITE_PRENET1, pidx + j) >= 0.5f ? 1.f : 0.f);
                const float pv = fmaxf(acc, 0.f) * keep * 2.f;
                xa[j] = t2v_rnd(pv, rnd);
                if (OP) xa16[((long long)t * B + b) * XA_W + j] = t2v_enc16(pv, opfmt);
              }
            }
          }
        }
        named_bar(BAR_ATT, 256);
        if (atid == 0) { TR(n, 2); signal_counter(cnt_p); TR(n, 3); }
      }
      // ---- the query of step t.  Flag-in-data: the 32 per-cluster partials of this utterance are polled directly -- every slot
      // holds a sentinel (a NaN pattern no partial can take) until its producer's store lands -- instead of waiting for the
      // device-wide h counter and THEN fetching them (one L2 round trip less on the critical chain, no exposure to the slowest of
      // the 128 CTAs).  Slots are re-armed one step later by the hh == 0 CTA: by then both CTAs of the pair have consumed them
      // (they exchanged the energies computed from this query), and the next write to the same parity comes after this CTA's
      // ctx signal of the current step.  CTAs without an utterance keep the counter wait (it keeps their signals in step).
      if (!active) {
        if (atid == 0) wait_counter(cnt_h, NCTA * (n + 1));
        named_bar(BAR_ATT, 256);
      }
      if (active) {
        if (atid == 0) TR(n, 16);
        {
          const int half = atid >> 7;
          if (hh == 0 && n > 0) {      // re-arm the slots of step t - 1
            float* qo = p.qpart + ((long long)(((t - 1) & 1) * NCLUSTER + half * 16) * B + b) * AD + a;
#pragma unroll
            for (int k = 0; k < 16; ++k) qo[(long long)k * B * AD] = __uint_as_float(Q_SENTINEL);
          }
          const float* qp = p.qpart + ((long long)((t & 1) * NCLUSTER + half * 16) * B + b) * AD + a;
          float v[16];
          bool ok = false;
          const long long t0 = clock64();
          // (while nothing has arrived only ONE slot per thread is polled: 32 sectors per CTA and poll instead of 512)
          while (__float_as_uint(ld_relaxed_f32(qp)) == Q_SENTINEL) { if (clock64() - t0 > WAIT_LIMIT) __trap(); }
          while (!ok) {
            ok = true;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              v[k] = ld_relaxed_f32(qp + (long long)k * B * AD);
              ok = ok && (__float_as_uint(v[k]) != Q_SENTINEL);
            }
            if (!ok && clock64() - t0 > WAIT_LIMIT) __trap();
          }
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int k = 0; k < 16; k += 2) { acc0 += v[k]; acc1 += v[k + 1]; }
          q_s[half * 128 + a] = acc0 + acc1;
        }
        if (atid == 0) { TR(n, 17); GT(n, 0); }
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(n, 18);
#pragma unroll
        for (int k = 0; k < 4; ++k) qv[k] = q_s[lane + 32 * k] + q_s[128 + lane + 32 * k];
        // ---- energies of this CTA's rows: warp aw owns rows aw + 8k (k < 8), all in flight at once; a transposing
        // reduction (9 shuffles) leaves row aw + 8*(lane & 7) in every lane, lanes 0..7 mirror them into the partner CTA
        {
          float x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = aw + 8 * k;
            float acc = 0.f;
            if (rr < nrow) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const float av = t2v_tanh(qv[kk] + S[rr * AD + lane + 32 * kk]);
                acc = fmaf(vreg[kk], av, acc);
              }
            }
            x[k] = acc;
          }
#pragma unroll
          for (int sft = 4; sft >= 1; sft >>= 1) {
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int i = 0; i < sft; ++i) {
              const float send = up ? x[i] : x[i + sft];
              const float keep = up ? x[i + sft] : x[i];
              x[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
            }
          }
          float e = x[0];
          e += __shfl_xor_sync(0xffffffffu, e, 8);
          e += __shfl_xor_sync(0xffffffffu, e, 16);
          const int rr = aw + 8 * (lane & 7);
          if (lane < 8 && rr < nrow) {
            e = (i0 + rr < len) ? e : s.mask_value;
            e_s[i0 + rr] = e;
            st_async_f32(partner_e + (uint32_t)(i0 + rr) * 4u, partner_full, e);
          }
        }
        if (atid == 0) TR(n, 19);
        mbar_wait_cluster(e_full, n & 1u);
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(n, 20);
        // ---- softmax over all Ti positions (one warp; Ti <= 128), cumulative weights
        if (aw == 0) {
          float ev[4], m = -INFINITY;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = lane + 32 * k;
            ev[k] = (i < Ti) ? e_s[i] : -INFINITY;
            m = fmaxf(m, ev[k]);
          }
          m = warp_max(m);
          float ssum = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = lane + 32 * k;
            ev[k] = (i < Ti) ? expf(ev[k] - m) : 0.f;
            ssum += ev[k];
          }
          ssum = warp_sum(ssum);
          const float inv = 1.f / ssum;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = lane + 32 * k;
            if (i < Ti) {
              const float w = ev[k] * inv;
              const float cn = cpad[HALO + i] + w;
              wpad[HALO + i] = w;
              cpad[HALO + i] = cn;
              if (hh == 0) {
                s.align[((long long)b * To + t) * Ti + i] = w;
                s.CUM[((long long)(t + 1) * B + b) * Ti + i] = cn;
              }
            }
          }
        }
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(n, 21);
        // ---- context columns [256 hh, 256 hh + 256)
        const bool ctx16 = OP && s.mem16 != nullptr;
        if (ctx16) {
          // fp16 copy of the encoder memory: thread = (8 columns, every 8th text position) -> all Ti <= 128 rows of the column group
          // in flight at once (ONE L2 round trip instead of two, half the bytes)
          const int cg = atid & 31, rg = atid >> 5;
          const uint16_t* mb = reinterpret_cast<const uint16_t*>(s.mem16) + (long long)b * Ti * ED + 256 * hh + 8 * cg;
          uint4 mv[16];
          float wi[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const int ti = rg + 8 * r;
            wi[r] = (ti < Ti) ? wpad[HALO + ti] : 0.f;
            mv[r] = (wi[r] == 0.f) ? make_uint4(0u, 0u, 0u, 0u) : ldg_u4_hint(mb + (long long)ti * ED, pol_m);
          }
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const uint32_t w4[4] = {mv[r].x, mv[r].y, mv[r].z, mv[r].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[2 * i] = fmaf(wi[r], t2v_f16_to_f32((uint16_t)(w4[i] & 0xFFFFu)), acc[2 * i]);
              acc[2 * i + 1] = fmaf(wi[r], t2v_f16_to_f32((uint16_t)(w4[i] >> 16)), acc[2 * i + 1]);
            }
          }
          *reinterpret_cast<float4*>(fbuf + rg * 256 + 8 * cg) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(fbuf + rg * 256 + 8 * cg + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
          const int cg = atid & 63, rg = atid >> 6;
          const float* mb = s.mem + (long long)b * Ti * ED + 256 * hh + 4 * cg;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int base = rg; base < Ti; base += 64) {
            float4 mv[16];
            float wi[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const int ti = base + 4 * r;
              wi[r] = (ti < Ti) ? wpad[HALO + ti] : 0.f;
              mv[r] = (wi[r] == 0.f) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                     : (p.mem_hint ? ldg_v4_hint(mb + (long long)ti * ED, pol_m)
                                                   : __ldg(reinterpret_cast<const float4*>(mb + (long long)ti * ED)));
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              acc.x = fmaf(wi[r], mv[r].x, acc.x); acc.y = fmaf(wi[r], mv[r].y, acc.y);
              acc.z = fmaf(wi[r], mv[r].z, acc.z); acc.w = fmaf(wi[r], mv[r].w, acc.w);
            }
          }
          *reinterpret_cast<float4*>(fbuf + rg * 256 + 4 * cg) = acc;
        }
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(n, 22);
        {
          float cx = (fbuf[atid] + fbuf[256 + atid]) + (fbuf[512 + atid] + fbuf[768 + atid]);
          if (ctx16) cx += (fbuf[1024 + atid] + fbuf[1280 + atid]) + (fbuf[1536 + atid] + fbuf[1792 + atid]);
          const float c = t2v_rnd(cx, rnd);
          const int col = 256 * hh + atid;
          s.XD[((long long)t * B + b) * XD_W + H + col] = c;              // ctx_t -> decoder_rnn input
          s.XA[((long long)(t + 1) * B + b) * XA_W + PD + col] = c;      // ctx_t -> next attention_rnn input
          if (OP) {
            const uint16_t c16 = t2v_enc16(cx, opfmt);
            xd16[((long long)t * B + b) * XD_W + H + col] = c16;
            xa16[((long long)(t + 1) * B + b) * XA_W + PD + col] = c16;
          }
          if (s.HCLO) {
            __stcs(s.HCHI + ((long long)t * B + b) * (H + ED) + H + col, c);
            __stcs(s.HCLO + ((long long)t * B + b) * (H + ED) + H + col, t2v_tf32(cx - c));
          }
          if (INFER) ctx_s[atid] = cx;                                    // exact ctx for the in-kernel projection
        }
        if (INFER) {
          // ---- partial mel / gate projection of this CTA's 256 ctx columns (the ctx half of [h_dec | ctx] W_pg^T): warp per output
          named_bar(BAR_ATT, 256);
          const float4 c0 = *reinterpret_cast<const float4*>(ctx_s + 4 * lane);
          const float4 c1 = *reinterpret_cast<const float4*>(ctx_s + 128 + 4 * lane);
          float* oc = p.ocpart + ((long long)(((t & 1) * B + b) * 2 + hh)) * 84;
#pragma unroll 4
          for (int o = aw; o < 81; o += 8) {
            const float* wr = p.Wpg + (long long)o * (H + ED) + H + 256 * hh;
            const float4 w0 = ldg_v4_hint(wr + 4 * lane, pol_keep);
            const float4 w1 = ldg_v4_hint(wr + 128 + 4 * lane, pol_keep);
            float acc = (w0.x * c0.x + w0.y * c0.y + w0.z * c0.z + w0.w * c0.w) + (w1.x * c1.x + w1.y * c1.y + w1.z * c1.z + w1.w * c1.w);
            acc = warp_sum(acc);
            if (lane == 0) oc[o] = acc;
          }
        }
      }
      named_bar(BAR_ATT, 256);
      if (atid == 0) { GT(n, 1); signal_counter(cnt_c); TR(n, 23); GT(n, 5); }
      // ---- saved tanh activations of the backward pass, recomputed off the critical path (S and the query are still here;
      // storing them inside the energy loop would put 30 KB of global stores in front of the exchange)
      if (active && s.ASAVE) {
        float* asave = s.ASAVE + (((long long)t * B + b) * Ti + i0) * AD;
        for (int rr = aw; rr < nrow; rr += 8) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            __stcs(asave + (long long)rr * AD + lane + 32 * kk, t2v_tanh(qv[kk] + S[rr * AD + lane + 32 * kk]));
        }
        named_bar(BAR_ATT, 256);           // S is overwritten by the next step's bulk copy
      }
    }
  }
  if (INFER && warp >= 8) {
    // ---- the last frame of the range (the loop publishes frame t-1 at the start of step t)
    const int atid = tid - 256;
    const int b = 2 * cid + (rank >> 1), hh = rank & 1;
    float* fb = (float*)(smem + SM<OP>::OFF_F);
    if (atid == 0) { wait_counter(cnt_d, NCTA * (unsigned)(te - tb)); wait_counter(cnt_c, NCTA * (unsigned)(te - tb)); }
    named_bar(BAR_ATT, 256);
    if (b < B && atid < 81) {
      const int f = te - 1, o = atid;
      float acc = p.bpg[o];
      const float* op = p.opart + ((long long)((f & 1) * NCLUSTER) * B + b) * 84 + o;
      for (int c = 0; c < NCLUSTER; ++c) acc += __ldcg(op + (long long)c * B * 84);
      const float* oc = p.ocpart + ((long long)((f & 1) * B + b) * 2) * 84 + o;
      acc += __ldcg(oc) + __ldcg(oc + 84);
      fb[o] = acc;
      if (hh == 0) {
        p.O[((long long)f * B + b) * 84 + o] = acc;
        if (o == 80 && p.n_frames && p.n_frames[b] < 0 && 1.f / (1.f + expf(-acc)) > p.gate_threshold) p.n_frames[b] = f + 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();            // no CTA of the cluster exits while a peer may still address its shared memory
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(OP ? 256u : 128u) : "memory");
  }
}

// out tile order: mode 0/1 (forward): [cluster][rank][chunk][gate*32 + unit][32 k] ; mode 2/3 (backward, from the transposed
// weights): [cluster][rank][chunk][output column of the cluster][32 gate rows]
__global__ void pack_step_tiles_kernel(const float* __restrict__ W, int mode, float* __restrict__ out, long long n4) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const int kk = (int)(i4 & 7) * 4;
  long long row = i4 >> 3;                       // tile * rows_per_tile + m
  long long src;
  if (mode < 2) {
    const int nch = mode ? DEC_CHUNKS : ATT_CHUNKS, ld = mode ? XD_W : XA_W;
    const int m = (int)(row & 127);
    const long long tile = row >> 7;
    const int j = (int)(tile % nch), cr = (int)(tile / nch), rank = cr & 3, c = cr >> 2;
    const int kofs = mode ? dec_kofs(j, rank) : att_kofs(j, rank);
    src = (long long)((m >> 5) * H + 32 * c + (m & 31)) * ld + kofs + kk;
  } else {
    const int rows = (mode == 3) ? XD_W / NCLUSTER : XA_W / NCLUSTER;
    const int m = (int)(row % rows);
    const long long tile = row / rows;
    const int j = (int)(tile & 31), cr = (int)(tile >> 5), rank = cr & 3, c = cr >> 2;
    src = (long long)(rows * c + m) * (4 * H) + 1024 * rank + 32 * j + kk;
  }
  *reinterpret_cast<float4*>(out + i4 * 4) = *reinterpret_cast<const float4*>(W + src);
}

// 16-bit tiles (modes 0 / 1): [cluster][rank][chunk][gate*32 + unit][64 k] as fp16 / bf16, one contiguous 16 KB TMA box per chunk
__global__ void pack_step_tiles16_kernel(const float* __restrict__ W, int mode, uint16_t* __restrict__ out, int fmt, long long n4) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const int kk = (int)(i4 & 15) * 4;
  const long long row = i4 >> 4;                 // tile * rows_per_tile + m
  long long src;
  if (mode < 2) {
    const int nch = mode ? Chunks<1>::DEC : Chunks<1>::ATT, ld = mode ? XD_W : XA_W;
    const int m = (int)(row & 127);
    const long long tile = row >> 7;
    const int j = (int)(tile % nch), cr = (int)(tile / nch), rank = cr & 3, c = cr >> 2;
    const int kofs = mode ? dec_kofs16(j, rank) : att_kofs16(j, rank);
    src = (long long)((m >> 5) * H + 32 * c + (m & 31)) * ld + kofs + kk;
  } else {       // backward tiles from W^T [XA_W | XD_W, 4096]: [cluster][rank][16 chunks][output column of the cluster][64 gate rows]
    const int rows = (mode == 3) ? XD_W / NCLUSTER : XA_W / NCLUSTER;
    const int m = (int)(row % rows);
    const long long tile = row / rows;
    const int j = (int)(tile & 15), cr = (int)(tile >> 4), rank = cr & 3, c = cr >> 2;
    src = (long long)(rows * c + m) * (4 * H) + 1024 * rank + 64 * j + kk;
  }
  const float4 v = *reinterpret_cast<const float4*>(W + src);
  uint2 o;
  o.x = (uint32_t)t2v_enc16(v.x, fmt) | ((uint32_t)t2v_enc16(v.y, fmt) << 16);
  o.y = (uint32_t)t2v_enc16(v.z, fmt) | ((uint32_t)t2v_enc16(v.w, fmt) << 16);
  *reinterpret_cast<uint2*>(out + i4 * 4) = o;
}
__global__ void cvt16_2d_kernel(const float* __restrict__ src, long long s_ld, uint16_t* __restrict__ dst, long long d_ld,
                                long long rows, int cols, int fmt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i % cols);
  dst[r * d_ld + c] = t2v_enc16(src[r * s_ld + c], fmt);
}

// contiguous fp32 -> 16-bit, 8 elements per thread, optional device scale (fp16: saturating)
__global__ void cvt16_flat8_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, long long n8, int fmt,
                                   const float* __restrict__ scale_dev) {
  const float sc = scale_dev ? *scale_dev : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
    const float v[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint16_t lo, hi;
      if (fmt == 1) {
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(lo) : "f"(v[2 * j]));
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hi) : "f"(v[2 * j + 1]));
      } else {
        lo = t2v_bf16_bits(v[2 * j]); hi = t2v_bf16_bits(v[2 * j + 1]);
      }
      w[j] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// max |x| -> power-of-two scale (t2v_grad_scale)
__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  if ((((uintptr_t)x) & 15) == 0) {
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(x4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(x[i]));
  } else
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));     // non-negative floats order like their bit patterns
}
__global__ void scale_from_max_kernel(float* out, int target_log2) {
  const float m = out[2];
  int ex = 0;
  (void)frexpf(m, &ex);                          // m in [2^(ex-1), 2^ex)
  const float sc = (m > 0.f && isfinite(m)) ? exp2f((float)(target_log2 - ex)) : 1.f;
  out[0] = sc;
  out[1] = 1.f / sc;
}

bool persist_enabled() {      // read per call: the tests flip T2V_PERSIST to compare against the per-step launches
  const char* e = getenv("T2V_PERSIST");
  return !(e && e[0] == '0');
}

}  // namespace

int t2v_encode_tmap_2d(CUtensorMap* map, const void* base, int esize, long long inner, long long rows,
                       long long row_stride_elems, int box_rows);

// Returns 0 when the loop was enqueued, 1 when the persistent kernel does not apply to this problem (the caller then
// uses the per-step launches), anything else = error.  inf != nullptr: free-running inference (prenet, mel / gate projection
// and stop bookkeeping inside the kernel).
static int launch_persist(const T2VDecoderSeq* s, const T2VDecoderInfer* inf, int t_begin, int t_end, cudaStream_t stream) {
  if (!persist_enabled()) return 1;
  if (!s->use_tc || s->B > 64 || s->Ti > 2 * TH_MAX || s->Ti < 1 || t_end - t_begin < 2) return 1;
  if (!s->parts || !s->ebuf) return 1;
  if ((s->HCHI != nullptr) != (s->HCLO != nullptr)) { t2v_set_error("HCHI and HCLO come as a pair"); return -1; }
  if (inf && (!inf->Wp1 || !inf->Wp2 || !inf->Wpg || !inf->bpg || !inf->O)) return 1;
  const int op = s->op16 ? 1 : 0;
  if (op && (!s->XA16 || !s->XD16 || !s->WaP16 || !s->WdP16 || s->op16 > 2)) {
    t2v_set_error("op16 needs XA16 / XD16 / WaP16 / WdP16");
    return -1;
  }
  void (*kernel)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, PersistParams) =
      inf ? (op ? dec_persist_fwd_kernel<true, 1> : dec_persist_fwd_kernel<true, 0>)
          : (op ? dec_persist_fwd_kernel<false, 1> : dec_persist_fwd_kernel<false, 0>);
  static int max_clusters_dev[16][4];
  static bool mc_init = false;
  if (!mc_init) { for (auto& row : max_clusters_dev) for (int& v : row) v = -1; mc_init = true; }
  int* max_clusters = max_clusters_dev[t2v_device_slot()];
  static bool attr_set[4] = {false, false, false, false};
  const int ki = (inf ? 1 : 0) + 2 * op;
  const int smem_bytes = op ? SM<1>::SMEM_BYTES : SM<0>::SMEM_BYTES;
  if (!attr_set[ki]) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[ki] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(NCTA); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (max_clusters[ki] < 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    max_clusters[ki] = n;
  }
  if (max_clusters[ki] < NCLUSTER) return 1;       // all 128 CTAs must be co-resident (they wait on each other)

  PersistParams p;
  memset(&p, 0, sizeof(p));
  p.s = *s;
  p.t_begin = t_begin; p.t_end = t_end;
  p.counters = reinterpret_cast<unsigned*>(s->ebuf);
  p.qpart = s->parts;
  p.trace = nullptr;
  p.gtrace = nullptr;
  if (inf) {
    p.Wp1 = inf->Wp1; p.Wp2 = inf->Wp2; p.Wpg = inf->Wpg; p.bpg = inf->bpg; p.prenet_masks = inf->prenet_masks;
    p.O = inf->O; p.gate_threshold = inf->gate_threshold; p.n_frames = inf->n_frames;
    p.opart = s->parts + 8192LL * s->B;
    p.ocpart = s->parts + 16384LL * s->B;
  }
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  p.wa_hint = env_int("T2V_PERSIST_WA_HINT", 1);
  // fp32 operands: the 42 MB of decoder_rnn weights stream from HBM every step and must not evict W_a / the encoder memory
  // (evict_first; inference: 50.8 vs 51.4 us with evict_last).  16-bit operands: all 36 MB of weights stay L2-resident next to the
  // memory (evict_last: 13.90 -> 13.75 us per training step, 40.2 -> 39.6 us per inference step)
  p.wd_hint = env_int("T2V_PERSIST_WD_HINT", op ? 1 : 2);
  p.mem_hint = env_int("T2V_PERSIST_MEM_HINT", 1);
  const bool trace = getenv("T2V_PERSIST_TRACE") != nullptr && (t_end - t_begin) >= TRACE_T0 + TRACE_STEPS + 2;
  const size_t trace_bytes = 2 * TRACE_STEPS * 32 * sizeof(long long);
  if (trace) {
    T2V_CUDA_CHECK(cudaMalloc(&p.trace, trace_bytes));
    T2V_CUDA_CHECK(cudaMemsetAsync(p.trace, 0, trace_bytes, stream));
    T2V_CUDA_CHECK(cudaMalloc(&p.gtrace, NCTA * 8 * sizeof(long long)));
    T2V_CUDA_CHECK(cudaMemsetAsync(p.gtrace, 0, NCTA * 8 * sizeof(long long), stream));
  }
  CUtensorMap tmWa, tmWd, tmXA, tmXD;
  const long long rows = (long long)(s->To + 1) * s->B;
  int r;
  p.packed = (s->WaP && s->WdP && env_int("T2V_PERSIST_PACKED", 1)) ? 1 : 0;
  if (op) {
    p.packed = 1;
    if ((r = t2v_encode_tmap_2d(&tmWa, s->WaP16, 2, 64, (long long)NCTA * Chunks<1>::ATT * 128, 64, 128))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd, s->WdP16, 2, 64, (long long)NCTA * Chunks<1>::DEC * 128, 64, 128))) return r;
  } else if (p.packed) {
    if ((r = t2v_encode_tmap_2d(&tmWa, s->WaP, 4, 32, (long long)NCTA * ATT_CHUNKS * 128, 32, 128))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd, s->WdP, 4, 32, (long long)NCTA * DEC_CHUNKS * 128, 32, 128))) return r;
  } else {
    if ((r = t2v_encode_tmap_2d(&tmWa, s->Wa, 4, XA_W, 4 * H, XA_W, 32))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd, s->Wd, 4, XD_W, 4 * H, XD_W, 32))) return r;
  }
  if (op) {
    if ((r = t2v_encode_tmap_2d(&tmXA, s->XA16, 2, XA_W, rows, XA_W, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmXD, s->XD16, 2, XD_W, rows, XD_W, 64))) return r;
  } else {
    if ((r = t2v_encode_tmap_2d(&tmXA, s->XA, 4, XA_W, rows, XA_W, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmXD, s->XD, 4, XD_W, rows, XD_W, 64))) return r;
  }
  T2V_CUDA_CHECK(cudaMemsetAsync(p.counters, 0, 128 * sizeof(unsigned), stream));
  fill_u32_kernel<<<64, 256, 0, stream>>>(reinterpret_cast<unsigned*>(p.qpart), 2LL * NCLUSTER * s->B * AD, Q_SENTINEL);   // arm the query slots
  cfg.numAttrs = t2v_coop_enabled() ? 2 : 1;
  T2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmWa, tmWd, tmXA, tmXD, p));
  T2V_COUNT_LAUNCH();
  if (trace) {      // debugging aid: not usable under stream capture
    static const char* names[26] = {"A:h ready", "A:ctx ready", "A:att loads issued", "A:dec loads issued", "M:att first chunk",
                                    "M:att ctx chunk", "M:att committed", "M:dec committed", "E:att acc full", "E:att pushed",
                                    "E:att recv full", "E:att cell done", "E:att q partial done", "E:h signalled", "E:dec acc full",
                                    "E:dec signalled", "T:loc done, wait h", "T:h seen", "T:query summed", "T:energies done",
                                    "T:energies exchanged", "T:softmax done", "T:context partials", "T:ctx signalled",
                                    "W:att first load", "W:step loads issued"};
    long long h[2 * TRACE_STEPS * 32];
    T2V_CUDA_CHECK(cudaStreamSynchronize(stream));
    T2V_CUDA_CHECK(cudaMemcpy(h, p.trace, trace_bytes, cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    {      // cross-CTA view of one step (global timer, ns)
      static long long g[NCTA * 8];
      T2V_CUDA_CHECK(cudaMemcpy(g, p.gtrace, sizeof(g), cudaMemcpyDeviceToHost));
      cudaFree(p.gtrace);
      static const char* gn[6] = {"T:query partials seen", "T:ctx signal begin", "A:ctx ready (this step's GEMM)", "E:h signal begin", "E:att acc full",
                                  "T:ctx signal done"};
      long long base = 0;
      for (int c = 0; c < NCTA; ++c) if (g[c * 8 + 4] && (!base || g[c * 8 + 4] < base)) base = g[c * 8 + 4];
      fprintf(stderr, "[t2v persist gtrace] step %d, ns relative to the earliest 'E:att acc full'; per event: min / median / p90 / max (CTA of max)\n", TRACE_T0);
      for (int ev = 0; ev < 6; ++ev) {
        long long v[NCTA]; int arg = 0;
        for (int c = 0; c < NCTA; ++c) { v[c] = g[c * 8 + ev] - base; if (v[c] > v[arg]) arg = c; }
        long long srt[NCTA];
        memcpy(srt, v, sizeof(srt));
        for (int i = 1; i < NCTA; ++i) { long long x = srt[i]; int j = i - 1; while (j >= 0 && srt[j] > x) { srt[j + 1] = srt[j]; --j; } srt[j + 1] = x; }
        fprintf(stderr, "  %-32s %7lld %7lld %7lld %7lld  (CTA %d)\n", gn[ev], srt[0], srt[NCTA / 2], srt[NCTA * 9 / 10], srt[NCTA - 1], arg);
      }
      fprintf(stderr, "  late CTAs at 'T:ctx signal begin' (> median + 500 ns):");
      { long long v[NCTA], srt[NCTA]; for (int c = 0; c < NCTA; ++c) v[c] = g[c * 8 + 1] - base; memcpy(srt, v, sizeof(srt));
        for (int i = 1; i < NCTA; ++i) { long long x = srt[i]; int j = i - 1; while (j >= 0 && srt[j] > x) { srt[j + 1] = srt[j]; --j; } srt[j + 1] = x; }
        for (int c = 0; c < NCTA; ++c) if (v[c] > srt[NCTA / 2] + 500) fprintf(stderr, " %d(+%lld)", c, v[c] - srt[NCTA / 2]); }
      fprintf(stderr, "\n");
    }
    for (int c = 0; c < 2; ++c) {
      const long long base = h[(c * TRACE_STEPS) * 32 + 1];     // "ctx ready" of the first traced step
      fprintf(stderr, "[t2v persist trace] CTA %d: SM clocks relative to 'A:ctx ready' of step %d\n", c ? TRACE_CTA_B : 0, TRACE_T0);
      for (int ev = 0; ev < 26; ++ev) {
        fprintf(stderr, "  %-24s", names[ev]);
        for (int n = 0; n < TRACE_STEPS; ++n) fprintf(stderr, " %8lld", h[(c * TRACE_STEPS + n) * 32 + ev] ? h[(c * TRACE_STEPS + n) * 32 + ev] - base : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return 0;
}

int t2v_decoder_fwd_persist(const T2VDecoderSeq* s, int t_begin, int t_end, cudaStream_t stream) {
  return launch_persist(s, nullptr, t_begin, t_end, stream);
}
int t2v_decoder_infer_persist(const T2VDecoderInfer* d, int t_begin, int t_end, cudaStream_t stream) {
  return launch_persist(&d->f, d, t_begin, t_end, stream);
}

T2V_API int t2v_pack_step_tiles16(const float* W, int mode, void* out, int fmt, cudaStream_t stream) {
  T2V_ARG_CHECK(W && out && mode >= 0 && mode <= 3 && (fmt == 1 || fmt == 2), "mode 0..3, fmt 1 (fp16) / 2 (bf16)");
  const long long n4 = (long long)4 * H * ((mode == 0 || mode == 2) ? XA_W : XD_W) / 4;
  pack_step_tiles16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(W, mode, reinterpret_cast<uint16_t*>(out), fmt, n4);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
T2V_API int t2v_grad_scale(const float* x, long long n, int target_log2, float* out, cudaStream_t stream) {
  T2V_ARG_CHECK(x && out && n > 0 && target_log2 >= -8 && target_log2 <= 15, "shape / target");
  T2V_CUDA_CHECK(cudaMemsetAsync(out, 0, 4 * sizeof(float), stream));
  absmax_kernel<<<296, 256, 0, stream>>>(x, n, reinterpret_cast<unsigned*>(out + 2));
  T2V_COUNT_LAUNCH();
  scale_from_max_kernel<<<1, 1, 0, stream>>>(out, target_log2);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
T2V_API int t2v_cvt16_2d(const float* src, long long s_ld, void* dst, long long d_ld, long long rows, int cols, int fmt,
                         cudaStream_t stream) {
  T2V_ARG_CHECK(src && dst && rows > 0 && cols > 0 && (fmt == 1 || fmt == 2), "fmt 1 (fp16) / 2 (bf16)");
  const long long n = rows * cols;
  cvt16_2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, s_ld, reinterpret_cast<uint16_t*>(dst), d_ld, rows, cols, fmt);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
T2V_API int t2v_cvt16_scaled(const float* src, void* dst, long long n, int fmt, const float* scale_dev, cudaStream_t stream) {
  T2V_ARG_CHECK(src && dst && n > 0 && n % 8 == 0 && (fmt == 1 || fmt == 2), "n % 8 == 0, fmt 1 (fp16) / 2 (bf16)");
  T2V_ARG_CHECK(((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0, "16-byte aligned buffers");
  const long long n8 = n / 8;
  const unsigned grid = (unsigned)((n8 + 255) / 256 < 148 * 16 ? (n8 + 255) / 256 : 148 * 16);
  cvt16_flat8_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst), n8, fmt, scale_dev);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
T2V_API int t2v_sizeof_decoder_structs(int which) {
  return which == 0 ? (int)sizeof(T2VDecoderSeq) : (which == 1 ? (int)sizeof(T2VDecoderBwd) : (int)sizeof(T2VDecoderInfer));
}

T2V_API int t2v_pack_step_tiles(const float* W, int mode, float* out, cudaStream_t stream) {
  T2V_ARG_CHECK(W && out && mode >= 0 && mode <= 3, "mode 0..3");
  const long long n4 = (long long)4 * H * ((mode == 0 || mode == 2) ? XA_W : XD_W) / 4;
  pack_step_tiles_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(W, mode, out, n4);
  T2V_COUNT_LAUNCH();
  T2V_LAUNCH_CHECK();
  return 0;
}
