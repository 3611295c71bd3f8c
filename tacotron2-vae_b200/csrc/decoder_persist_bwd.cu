// Persistent backward of the teacher-forced decoder loop (the reverse-time recurrence of Decoder.decode, model.py:346-389)
// as ONE kernel: 128 CTAs (32 clusters of 4) stay resident for all To steps, like decoder_persist.cu for the forward.
//
// Per step t (descending) the loop has two chains:
//   decoder_rnn chain  (one step ahead, feeds nothing back into the attention chain of earlier steps except dXD[t]):
//       D1  cell backward   dh_dec = DHC[t][:1024] + dXD[t+1][1536:]  -> DGD[t]                       (epilogue warps)
//       D2  dXD[t] = DGD[t] W_d   (tcgen05, split-K over the cluster, DSMEM exchange)               (MMA + epilogue warps)
//   attention chain (the critical recurrence):
//       A1  attention backward per utterance (two CTAs, rows split): dctx -> dw -> softmax backward -> dpre -> dq,
//           then (off the critical path) dW_loc, dW_conv, dv accumulated in registers for the whole sequence and the
//           adjoint location conv that carries the gradient to the previous step's alignments     (attention warps)
//       A2  dHq = dq W_q, attention_rnn cell backward -> DGA[t]                                     (epilogue warps)
//       A3  DXA[t] = DGA[t] W_a                                                                      (MMA + epilogue warps)
//   GEMM decomposition: cluster c owns output columns [56c, 56c+56) of DXA (UMMA M=64) and [80c, 80c+80) of dXD (M=128),
//   the batch is N, the 4 CTAs split K = 4096 gate rows; W^T tiles stream through a TMA ring that runs ahead of the
//   recurrence, the gate gradients follow as soon as the device-wide counters say they are complete.
// Five monotonic device-wide counters order the phases (release / acquire); every wait is bounded.
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
constexpr int CL = 4, NCLUSTER = 32, NCTA = CL * NCLUSTER, NTHREADS = 512;
constexpr int XA_CPC = XA_W / NCLUSTER, XD_CPC = XD_W / NCLUSTER;   // 56 / 80 output columns per cluster
// ring depth: a stage = one K chunk: W^T tile (<= 80 rows, 10 KB) + gate gradients (8 KB)
#ifndef T2V_BWD_NS
#define T2V_BWD_NS 5
#endif
constexpr int W_PART = 80 * 128, A_STAGE = 64 * 128, STAGE = W_PART + A_STAGE;
constexpr int KCH32 = 32;                      // fp32 storage: 32-wide K chunks per CTA and GEMM (K slice = 4096 / 4)
// op16: the W^T tiles and the gate gradients stream as fp16 copies (kind::f16), 64 K columns per 128-byte row -> 16 chunks of the
// same bytes.  The gate gradients are ~1e-8: they are multiplied by a power-of-two scale (device scalar, derived from max |dO| by
// t2v_grad_scale) before the saturating fp16 conversion and the dX rows are unscaled in fp32, so the fp16 mode keeps tf32's 11-bit
// significand on both operands and only the RANGE is managed.
template <int OP> struct KC { static constexpr int N = OP ? 16 : 32, W = OP ? 64 : 32; };
constexpr int RA_P = 64, RD_P = 80;            // column pitch of the exchange slots [src][batch row 16][cols]
constexpr int TH_MAX = 64, FS = 36;
constexpr int BAR_EPI = 1, BAR_ATT = 2;
constexpr int US_P = 33;                       // row pitch of the staged adjoint-conv products (op16)

// small arrays (floats): wpad[160] cpad[160] wt[64] dwv[64] de[64] Ps[64] Gs[64] adj[2][64] halo[2][2][64] spart[2][2] q_s[2][128]
// tmax[8]
constexpr int SM_WPAD = 0, SM_CPAD = 160, SM_WT = 320, SM_DWV = 384, SM_DE = 448, SM_P = 512, SM_G = 576, SM_ADJ = 640,
              SM_HALO = 768, SM_SPART = 1024, SM_QS = 1028, SM_TMAX = 1284, SM_TOTAL = 1292;
constexpr int N_BARS_MAX = 2 * 5 + 12;

// Shared-memory map.  op16 trades one ring stage and the fp32 location-dense weights for the fp16 operands of the attention-tail
// UMMAs (dW_loc, df, adjoint-conv products; see the attention warps).
template <int OP> struct LY {
  static constexpr int NS = OP ? (T2V_BWD_NS > 4 ? 4 : T2V_BWD_NS) : T2V_BWD_NS;
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_RECVA = OFF_RING + NS * STAGE;
  static constexpr int OFF_RECVD = OFF_RECVA + 4 * 16 * RA_P * 4;
  static constexpr int OFF_ABUF = OFF_RECVD + 4 * 16 * RD_P * 4;            // [TH_MAX][128] saved tanh -> dpre (in place)
  static constexpr int OFF_F = OFF_ABUF + TH_MAX * AD * 4;                   // [TH_MAX][FS] conv output f -> df (in place)
  static constexpr int OFF_WCT = OFF_F + TH_MAX * FS * 4;                    // [62][32]
  static constexpr int OFF_WLOC = OFF_WCT + 2 * KS * NF * 4;                 // [128][32] fp32 (FFMA path only)
  // dHq = dq W_q on the tensor core (per CTA: 32 units x 16 batch rows, K = 128 attention dims): fp16 operands, K-major rows of 128
  // bytes (SWIZZLE_128B), two K blocks of 64.  A = W_q^T slice [64 rows: 32 units + 32 zero rows][128 a] (static), B = dq rows of this
  // CTA [16 b][128 a], scaled by a per-CTA power of two so that gradient magnitudes sit inside the fp16 range; D [unit][b] in TMEM.
  static constexpr int OFF_WQ16 = (OFF_WLOC + (OP ? 0 : AD * NF * 4) + 1023) / 1024 * 1024;   // 2 blocks x [64][64] fp16 = 16 KB
  static constexpr int OFF_DQ16 = OFF_WQ16 + 2 * 64 * 128;                   // 2 blocks x [16][64] fp16 = 4 KB
  // op16 attention tail: dpre [64 rows][128 a] (2 SW128 blocks), f^T [32 filters][64 rows] (SW128), W_loc^T [32][128 a] (2 SW128
  // blocks), df [64 rows][32 filters] (SW64), W_conv [64 = 2 x 32 taps][32 filters] (SW64)
  static constexpr int OFF_DP16 = OFF_DQ16 + 2 * 16 * 128;
  static constexpr int OFF_F16T = OFF_DP16 + (OP ? 2 * 64 * 128 : 0);
  static constexpr int OFF_WL16T = OFF_F16T + (OP ? 32 * 128 : 0);
  static constexpr int OFF_DF16 = OFF_WL16T + (OP ? 2 * 32 * 128 : 0);
  static constexpr int OFF_WC16 = OFF_DF16 + (OP ? 64 * 64 : 0);
  static constexpr int OFF_DHQ = OFF_WC16 + (OP ? 64 * 64 : 0);              // [16][32] fp32 result tile + [4] partial maxima
  static constexpr int OFF_DCTX = OFF_DHQ + 16 * 32 * 4 + 16;                // [512]
  static constexpr int OFF_BIAS = OFF_DCTX + ED * 4;                         // [4 gates][4 warps][32 units] staging of the bias-gradient terms
  static constexpr int OFF_SMALL = OFF_BIAS + 4 * 4 * 32 * 4;
  static constexpr int OFF_BARS = OFF_SMALL + ((SM_TOTAL * 4 + 15) / 16) * 16;
  static constexpr int OFF_TMEM = OFF_BARS + N_BARS_MAX * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(2 * TH_MAX * US_P * 4 <= TH_MAX * AD * 4, "adjoint products are staged in the tanh tile");
};

__device__ __forceinline__ uint16_t f16_sat(float x) {      // round to nearest, clamp to +-65504 instead of producing inf
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}

struct BwdParams {
  T2VDecoderBwd d;
  unsigned* counters;      // [0] DGD[t] complete, [32] dXD[t], [64] dq[t], [96] DGA[t], [128] DXA[t]
  int two_issuers;         // warps 2 and 3 both issue MMAs (even / odd chunks)
  int dbg_skip;            // timing experiments only (results are garbage): 1 = no weight loads, 2 = no gate-gradient loads
  int rotate;              // rotate the chunk order per cluster
  int packed;              // the weight tensor maps describe the re-tiled copies (contiguous boxes)
  int wa_hint, wd_hint;
  const float* scale;      // op16: [2] = {s, 1 / s}
  long long* trace;        // T2V_PERSIST_TRACE: [2 CTAs][TRACE_STEPS][32] clock64 stamps, else nullptr
  long long* gtrace;       // T2V_PERSIST_TRACE: [NCTA][8] %globaltimer stamps (ns, comparable ACROSS CTAs) of iteration TRACE_I0
};
constexpr int TRACE_I0 = 100, TRACE_STEPS = 4, TRACE_CTA_B = 77;

template <int OP>
__global__ void __launch_bounds__(NTHREADS, 1)
dec_persist_bwd_kernel(const __grid_constant__ CUtensorMap tmWa, const __grid_constant__ CUtensorMap tmWd64,
                       const __grid_constant__ CUtensorMap tmWd16, const __grid_constant__ CUtensorMap tmGA,
                       const __grid_constant__ CUtensorMap tmGD, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  using L = LY<OP>;
  constexpr int NS = L::NS;
  uint8_t* ring = smem + L::OFF_RING;     // stage s: [W^T tile 10 KB | gate gradients 8 KB]; the M=128 descriptor of the dXD
                                          // GEMM reads 48 rows past the 80 loaded ones (into the gradient part): those
                                          // accumulator rows are never drained
  float* recv_a = (float*)(smem + L::OFF_RECVA);
  float* recv_d = (float*)(smem + L::OFF_RECVD);
  float* Abuf = (float*)(smem + L::OFF_ABUF);
  float* f_s = (float*)(smem + L::OFF_F);
  float* wcT = (float*)(smem + L::OFF_WCT);
  float* WlocS = (float*)(smem + L::OFF_WLOC);
  uint16_t* wq16 = (uint16_t*)(smem + L::OFF_WQ16);
  uint16_t* dq16 = (uint16_t*)(smem + L::OFF_DQ16);
  uint16_t* dp16 = (uint16_t*)(smem + L::OFF_DP16);
  uint16_t* f16T = (uint16_t*)(smem + L::OFF_F16T);
  uint16_t* wl16T = (uint16_t*)(smem + L::OFF_WL16T);
  uint16_t* df16 = (uint16_t*)(smem + L::OFF_DF16);
  uint16_t* wc16 = (uint16_t*)(smem + L::OFF_WC16);
  float* dhq_s = (float*)(smem + L::OFF_DHQ);
  float* qmax_s = dhq_s + 16 * 32;
  float* dctx_s = (float*)(smem + L::OFF_DCTX);
  float* bias_s = (float*)(smem + L::OFF_BIAS);
  float* small = (float*)(smem + L::OFF_SMALL);
  uint64_t* bars = (uint64_t*)(smem + L::OFF_BARS);
  uint64_t* full = bars;                  // [NS] both producers arrive (count 2) with their byte counts: ONE wait per chunk
  uint64_t* empty = full + NS;            // [NS]
  uint64_t* acc_full = empty + NS;        // [2]: 0 = DXA GEMM, 1 = dXD GEMM
  uint64_t* acc_free = acc_full + 2;      // [2]
  uint64_t* recv_full = acc_free + 2;     // [2]
  uint64_t* at_full = recv_full + 2;      // saved tanh tile landed
  uint64_t* x_full = at_full + 1;         // partner's softmax-backward partial sum landed
  uint64_t* h_full = x_full + 1;          // partner's adjoint-conv halo landed
  uint64_t* dq_full = h_full + 1;         // the dHq MMA of this iteration has retired
  uint64_t* t1_full = dq_full + 1;        // op16 attention tail: the dW_loc / df MMAs have retired
  uint64_t* t2_full = t1_full + 1;        // op16 attention tail: the adjoint-conv product MMAs have retired
  uint32_t* tmem_holder = (uint32_t*)(smem + L::OFF_TMEM);

  constexpr int KCH = KC<OP>::N, CKW = KC<OP>::W;
  const T2VDecoderBwd& d = p.d;
  const T2VDecoderSeq& s = d.f;
  const float g_scale = OP ? p.scale[0] : 1.f, g_inv = OP ? p.scale[1] : 1.f;
  uint16_t* const dga16 = reinterpret_cast<uint16_t*>(d.DGA16);
  uint16_t* const dgd16 = reinterpret_cast<uint16_t*>(d.DGD16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int B = s.B, Ti = s.Ti, To = s.To;
  unsigned* cnt_gd = p.counters;
  unsigned* cnt_xd = p.counters + 32;
  unsigned* cnt_q = p.counters + 64;
  unsigned* cnt_ga = p.counters + 96;
  unsigned* cnt_xa = p.counters + 128;
  // the 32 clusters walk the K chunks in rotated order: at any moment the CTAs that share a K slice (same rank) read 32
  // DIFFERENT gate-gradient tiles instead of all hitting the same 64 L2 lines
  const int rot = p.rotate ? cid : 0;
  const int trace_slot = (blockIdx.x == 0) ? 0 : ((blockIdx.x == TRACE_CTA_B) ? 1 : -1);
  auto GT = [&](int it, int ev) {
    if (p.gtrace && it == TRACE_I0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      p.gtrace[blockIdx.x * 8 + ev] = (long long)ns;
    }
  };
  auto TR = [&](int it, int ev) {
    if (p.trace && trace_slot >= 0 && it >= TRACE_I0 && it < TRACE_I0 + TRACE_STEPS)
      p.trace[((long long)trace_slot * TRACE_STEPS + (it - TRACE_I0)) * 32 + ev] = clock64();
  };

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 128);
      mbar_init(&recv_full[i], 1);
    }
    mbar_init(at_full, 1);
    mbar_init(x_full, 1);
    mbar_init(h_full, 1);
    mbar_init(dq_full, 1);
    mbar_init(t1_full, 1);
    mbar_init(t2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  cluster_sync_all();

  // iteration `it` = -1 .. To-1: attention-chain step t = To-1-it (it >= 0), decoder_rnn-chain step td = t-1 (td >= 0)
  if (warp == 0) {
    // =========================================================================== weight producer
    if (lane == 0) {
      uint32_t iw = 0;
      const uint64_t pol_a = l2_policy(p.wa_hint), pol_d = l2_policy(p.wd_hint);
      auto tma_w = [&](void* dst, const CUtensorMap* tm, int kofs, int row, uint64_t* bar, int hint, uint64_t pol) {
        if (hint) tma_load_2d_hint(dst, tm, kofs, row, bar, pol);
        else tma_load_2d(dst, tm, kofs, row, bar);
      };
      for (int it = -1; it < To; ++it) {
        const int td = To - 2 - it;
        if (td >= 0) {
          for (int jj = 0; jj < KCH; ++jj, ++iw) {    // 80 rows of W_d^T (two boxes: 64 + 16 rows)
            const int j = (jj + rot) & (KCH - 1);
            const int st = iw % NS;
            mbar_wait(&empty[st], ((iw / NS) & 1u) ^ 1u);
            const int k0 = p.packed ? 0 : 1024 * rank + 32 * j;
            const int r0 = p.packed ? ((cid * CL + rank) * KCH + j) * XD_CPC : XD_CPC * cid;
            if (p.dbg_skip & 1) { mbar_arrive(&full[st]); continue; }
            mbar_expect_tx(&full[st], XD_CPC * 128);
            tma_w(ring + st * STAGE, &tmWd64, k0, r0, &full[st], p.wd_hint, pol_d);
            tma_w(ring + st * STAGE + 64 * 128, &tmWd16, k0, r0 + 64, &full[st], p.wd_hint, pol_d);
          }
        }
        if (it >= 0) {
          for (int jj = 0; jj < KCH; ++jj, ++iw) {
            const int j = (jj + rot) & (KCH - 1);
            const int st = iw % NS;
            mbar_wait(&empty[st], ((iw / NS) & 1u) ^ 1u);
            const int k0 = p.packed ? 0 : 1024 * rank + 32 * j;
            const int r0 = p.packed ? ((cid * CL + rank) * KCH + j) * XA_CPC : XA_CPC * cid;
            if (p.dbg_skip & 1) { mbar_arrive(&full[st]); continue; }
            mbar_expect_tx(&full[st], XA_CPC * 128);
            tma_w(ring + st * STAGE, &tmWa, k0, r0, &full[st], p.wa_hint, pol_a);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== gate-gradient producer
    if (lane == 0) {
      uint32_t ia = 0;
      unsigned seen_gd = 0, seen_ga = 0;
      auto need = [&](const unsigned* cnt, unsigned& seen, unsigned target) {
        if (seen >= target) return;
        wait_counter(cnt, target);
        seen = target;
        fence_proxy_async();
      };
      auto load_a = [&](const CUtensorMap* tm, int kofs, int row0) {
        const int st = ia % NS;
        const uint32_t ph = (ia / NS) & 1u;
        mbar_wait(&empty[st], ph ^ 1u);
        if (p.dbg_skip & 2) { mbar_arrive(&full[st]); ++ia; return; }
        mbar_expect_tx(&full[st], A_STAGE);
        tma_load_2d(ring + st * STAGE + W_PART, tm, kofs, row0, &full[st]);
        ++ia;
      };
      for (int it = -1; it < To; ++it) {
        const int t = To - 1 - it, td = t - 1;
        if (td >= 0) {
          need(cnt_gd, seen_gd, NCTA * (unsigned)(it + 2));
          for (int j = 0; j < KCH; ++j) load_a(&tmGD, 1024 * rank + CKW * ((j + rot) & (KCH - 1)), td * B);
        }
        if (it >= 0) {
          need(cnt_ga, seen_ga, NCTA * (unsigned)(it + 1));
          for (int j = 0; j < KCH; ++j) load_a(&tmGA, 1024 * rank + CKW * ((j + rot) & (KCH - 1)), t * B);
        }
      }
    }
  } else if (warp == 2) {
    // =========================================================================== MMA issuer (whole warp in the loop, one
    // elected lane issues).  A second issuing warp (even / odd chunks into the same accumulator) was 2-8 % faster but made
    // about one bench run in four die with "unspecified launch failure": tcgen05.mma from two threads into the same TMEM
    // columns is not ordered by anything, so this stays a single issuer.
    {
      constexpr uint32_t fmt = OP ? 0u : 2u;       // fp16 : tf32
      constexpr uint32_t idesc64 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      constexpr uint32_t idesc128 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int st = 0;
      uint32_t ph = 0;
      const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(ring));
      const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(ring + W_PART));
      bool ready = false;                    // phase test of the current stage, started while the previous chunk was issued
      // which 1: dXD[td] = DGD[td] W_d, M = 128 (80 live rows), accumulator columns 64..127
      // which 0: DXA[t]  = DGA[t]  W_a, M = 64 (56 live rows), accumulator columns 0..63
      auto gemm = [&](const int which, const unsigned idx, const int it) {
        mbar_wait(&acc_free[which], (idx & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(which * 64);
        const uint32_t idesc = which ? idesc128 : idesc64;
        for (int j = 0; j < KCH; ++j) {
          if (!ready) mbar_wait(&full[st], ph);
          tc_fence_after();
          const int sn = (st + 1 == NS) ? 0 : st + 1;
          const uint32_t pn = (st + 1 == NS) ? (ph ^ 1u) : ph;
          ready = mbar_test_wait(&full[sn], pn);           // overlaps the issue below
          if (elect_one()) {
            if (j == 0) TR(it, which ? 27 : 29);
            const uint64_t adesc = adesc0 + (uint64_t)(st * (STAGE >> 4));
            const uint64_t bdesc = bdesc0 + (uint64_t)(st * (STAGE >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (OP) tc_mma_f16(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
              else tc_mma_tf32(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
            if (p.dbg_skip & 4) {        // timing experiment: twice the tensor work per chunk
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma_tf32(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
            }
            if (p.dbg_skip & 8) mbar_arrive(&empty[st]);     // timing experiment: free the stage at issue, not at MMA completion
            else tc_commit(&empty[st]);
            if (j == KCH - 1) {
              tc_commit(&acc_full[which]);
              TR(it, which ? 28 : 30);
            }
          }
          __syncwarp();
          st = sn; ph = pn;
        }
      };
      for (int it = -1; it < To; ++it) {
        const int td = To - 2 - it;
        if (td >= 0) gemm(1, (unsigned)(it + 1), it);
        if (it >= 0) gemm(0, (unsigned)it, it);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================================================================== cells + GEMM epilogues
    const int etid = tid - 128;
    const int q = warp & 3;
    const int u = lane, jg = 32 * cid + u, blq = etid >> 5;
    const int rnd = s.use_tc;
    uint32_t recv_remote[2][4], full_remote[2][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      recv_remote[0][r] = mapa(smem_u32(recv_a), (uint32_t)r);
      recv_remote[1][r] = mapa(smem_u32(recv_d), (uint32_t)r);
      full_remote[0][r] = mapa(smem_u32(&recv_full[0]), (uint32_t)r);
      full_remote[1][r] = mapa(smem_u32(&recv_full[1]), (uint32_t)r);
    }
    // W_q^T slice as the A operand: element (unit u, a) at K block a / 64, row u, 16-byte chunk ((a % 64) / 8) ^ (u & 7)
    for (int i = etid; i < 64 * AD; i += 128) {
      const int u = i >> 7, a = i & 127, kk = a & 63;
      const float w = (u < 32) ? s.Wq[(long long)a * H + 32 * cid + u] : 0.f;
      wq16[(a >> 6) * (64 * 64) + u * 64 + ((((kk >> 3) ^ (u & 7)) << 3) | (kk & 7))] = t2v_f16_bits(w);
    }
    fence_proxy_async();
    float dca[4], dcd[4];                    // running cell-state gradients of this thread's (unit, batch row) pairs
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = 16 * rank + blq + 4 * j;
      dca[j] = (b < B) ? d.dCa[(long long)b * H + jg] : 0.f;
      dcd[j] = (b < B) ? d.dCd[(long long)b * H + jg] : 0.f;
    }
    const uint64_t seed = (s.training && !s.drop_masks) ? t2v_resolve_seed(s.seed) : 0ull;
    const float p_att = s.training ? s.p_att : 0.f, p_dec = s.training ? s.p_dec : 0.f;
    unsigned seen_xd = 0, seen_xa = 0, seen_q = 0;
    int cur_it = -1;
    // bias gradients (sum of the gate gradients over batch and time, autograd of nn.LSTMCell's bias_ih / bias_hh, model.py:363-380):
    // accumulated here instead of a column reduction over the 1.7 GB of DGA / DGD after the loop.  Every thread stages the sums of
    // its four batch rows, after the cell's barrier epilogue warp g adds up gate g: one register per cell and thread, no atomics.
    float gb_acc[2] = {0.f, 0.f};
    named_bar(BAR_EPI, 128);

    // ---- LSTM cell backward (the math of lstm_pointwise_bwd_kernel) for step ts; which: 0 attention_rnn, 1 decoder_rnn
    auto cell_bwd = [&](auto which_c, const int ts, const unsigned n_xd, const unsigned n_xa, const unsigned n_q) {
      constexpr int which = decltype(which_c)::value;
      const long long r0 = (long long)ts * B, r1 = (long long)(ts + 1) * B;
      const bool has_next = ts + 1 < To;
      const float* Gs = which ? s.GD : s.GA;
      const float* CPs = which ? s.CPD : s.CPA;
      const float* Cs = which ? s.CD : s.CA;
      float* DG = which ? d.DGD : d.DGA;
      float* dcs = which ? dcd : dca;
      const float pdrop = which ? p_dec : p_att;
      const float kscale = 1.f / (1.f - pdrop);
      const float* mk = s.drop_masks ? s.drop_masks + (long long)ts * 4 * B * H + (which ? 2LL * B * H : 0) : nullptr;
      // the same lines of the NEXT iteration (step ts - 1) are pulled from HBM into L2 now: the loads below then cost an L2 hit
      // instead of a DRAM round trip inside the epilogue warps' serial program (the slowest CTA sets the pace of every step)
      if (ts > 0) {
        const long long rp = (long long)(ts - 1) * B;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = 16 * rank + blq + 4 * j;
          if (b < B) {
            const float* gs = Gs + (rp + b) * 4 * H + jg;
            prefetch_l2(gs); prefetch_l2(gs + H); prefetch_l2(gs + 2 * H); prefetch_l2(gs + 3 * H);
            prefetch_l2(CPs + (rp + b) * H + jg);
            prefetch_l2(Cs + (rp + b) * H + jg);
            if (which) prefetch_l2(d.DHC + (rp + b) * (H + ED) + jg);
          }
        }
      }
      // saved forward activations: independent of the recurrence, in flight while the counters are polled
      float sg[4][4], sc2[4], scp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = 16 * rank + blq + 4 * j;
        if (b < B) {
          const float* gs = Gs + (r0 + b) * 4 * H + jg;
          sg[j][0] = __ldcs(gs); sg[j][1] = __ldcs(gs + H); sg[j][2] = __ldcs(gs + 2 * H); sg[j][3] = __ldcs(gs + 3 * H);
          sc2[j] = __ldcs(CPs + (r0 + b) * H + jg);
          scp[j] = __ldcs(Cs + (r0 + b) * H + jg);
        } else {
          sg[j][0] = sg[j][1] = sg[j][2] = sg[j][3] = 0.f; sc2[j] = 0.f; scp[j] = 0.f;
        }
      }
      // dropout keep-scales of this step (counter RNG: no memory dependence) while the loads above are in flight
      float kh4[4], kc4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = 16 * rank + blq + 4 * j;
        float kh = 1.f, kc = 1.f;
        if (pdrop > 0.f && b < B) {
          const uint64_t li = (uint64_t)b * H + jg;
          if (mk) { kh = mk[li] * kscale; kc = mk[(long long)B * H + li] * kscale; }
          else {
            const uint64_t di = (uint64_t)ts * B * H + li;
            kh = (t2v_uniform(seed, which ? SITE_DEC_H : SITE_ATT_H, di) >= pdrop ? 1.f : 0.f) * kscale;
            kc = (t2v_uniform(seed, which ? SITE_DEC_C : SITE_ATT_C, di) >= pdrop ? 1.f : 0.f) * kscale;
          }
        }
        kh4[j] = kh; kc4[j] = kc;
      }
      if (etid == 0) {
        TR(cur_it, which ? 0 : 15);
        if (seen_xd < n_xd) { wait_counter(cnt_xd, n_xd); seen_xd = n_xd; }
        if (seen_xa < n_xa) { wait_counter(cnt_xa, n_xa); seen_xa = n_xa; }
        if (which) { TR(cur_it, 1); GT(cur_it, 1); }
      }
      named_bar(BAR_EPI, 128);
      float dh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = 16 * rank + blq + 4 * j;
        float v = 0.f;
        if (b < B) {
          if (which == 0) {
            v = __ldcg(d.DXD + (r0 + b) * XD_W + jg);
            if (has_next) v += __ldcg(d.DXA + (r1 + b) * XA_W + (PD + ED) + jg);
          } else {
            v = __ldcg(d.DHC + (r0 + b) * (H + ED) + jg);
            if (has_next) v += __ldcg(d.DXD + (r1 + b) * XD_W + (H + ED) + jg);
          }
        }
        dh[j] = v;
      }
      // everything of the cell that does not need dq is evaluated before the wait for it: tanh of the saved cell state
      float tc4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tc4[j] = t2v_tanh(sc2[j]);
      if (which == 0) {
        // the attention chain's critical wait: dq[t] of the whole batch (the loads and the RNG above ran in its shadow)
        if (etid == 0) {
          if (seen_q < n_q) { wait_counter(cnt_q, n_q); seen_q = n_q; }
          TR(cur_it, 3);
          GT(cur_it, 3);
        }
        named_bar(BAR_EPI, 128);
        // ---- dHq = dq W_q for this CTA's 16 batch rows x 32 units (dq of the whole batch is complete: cnt_q) on the tensor core.
        // thread -> (row etid / 8, attention dims 16 (etid % 8) .. +16): straight from L2 (the rows were accumulated with atomics)
        const int qr = etid >> 3, qa = (etid & 7) * 16;
        float4 dqv[4];
        {
          const int bq = 16 * rank + qr;
          const float4* src = reinterpret_cast<const float4*>(d.DQ + (r0 + bq) * AD + qa);
#pragma unroll
          for (int i = 0; i < 4; ++i) dqv[i] = (bq < B) ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(dqv[i].x), fabsf(dqv[i].y)), fmaxf(fabsf(dqv[i].z), fabsf(dqv[i].w))));
        mx = warp_max(mx);
        if (lane == 0) qmax_s[q] = mx;
        named_bar(BAR_EPI, 128);
        mx = fmaxf(fmaxf(qmax_s[0], qmax_s[1]), fmaxf(qmax_s[2], qmax_s[3]));
        // power-of-two scale that puts the largest |dq| of the tile into [2^13, 2^14): gradients are far below the fp16 range
        int ex = 0;
        (void)frexpf(mx, &ex);
        const float scale = (mx > 0.f) ? exp2f((float)(14 - ex)) : 1.f;
        {
          const int kb = qa >> 6, kk = qa & 63;                  // 16 consecutive a = two 16-byte chunks of row qr
          uint16_t* row = dq16 + kb * (16 * 64) + qr * 64;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4 lo = dqv[2 * c], hi = dqv[2 * c + 1];
            uint4 pk;
            pk.x = (uint32_t)t2v_f16_bits(lo.x * scale) | ((uint32_t)t2v_f16_bits(lo.y * scale) << 16);
            pk.y = (uint32_t)t2v_f16_bits(lo.z * scale) | ((uint32_t)t2v_f16_bits(lo.w * scale) << 16);
            pk.z = (uint32_t)t2v_f16_bits(hi.x * scale) | ((uint32_t)t2v_f16_bits(hi.y * scale) << 16);
            pk.w = (uint32_t)t2v_f16_bits(hi.z * scale) | ((uint32_t)t2v_f16_bits(hi.w * scale) << 16);
            *reinterpret_cast<uint4*>(row + ((((kk >> 3) + c) ^ (qr & 7)) << 3)) = pk;
          }
        }
        fence_proxy_async();
        tc_fence_before();
        named_bar(BAR_EPI, 128);
        tc_fence_after();
        if (etid == 0) TR(cur_it, 4);
        if (warp == 4 && elect_one()) {
          constexpr uint32_t idq = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);      // D f32, A = B = f16
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t ad = make_kmajor_sw128_desc(smem_u32(wq16) + kb * (64 * 128));
            const uint64_t bd = make_kmajor_sw128_desc(smem_u32(dq16) + kb * (16 * 128));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(tmem_base + 128u, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), idq, (kb | k4) ? 1u : 0u);
          }
          tc_commit(dq_full);
        }
        __syncwarp();
        mbar_wait(dq_full, (unsigned)cur_it & 1u);
        tc_fence_after();
        {
          // M = 64: accumulator row (unit) 16 q' + l sits in TMEM lane 32 q' + l (l < 16); units 0..31 -> warps q = 0, 1
          uint32_t hv[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + 128u, hv);
          if (q < 2 && lane < 16) {
            const float inv = 1.f / scale;
#pragma unroll
            for (int bl = 0; bl < 16; ++bl) dhq_s[bl * 32 + 16 * q + lane] = __uint_as_float(hv[bl]) * inv;
          }
        }
        tc_fence_before();
        named_bar(BAR_EPI, 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) dh[j] += dhq_s[(blq + 4 * j) * 32 + u];
        if (etid == 0) TR(cur_it, 5);
      }
      float bs0 = 0.f, bs1 = 0.f, bs2 = 0.f, bs3 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = 16 * rank + blq + 4 * j;
        const float kh = kh4[j], kc = kc4[j];
        const float ig = sg[j][0], fg = sg[j][1], gg = sg[j][2], og = sg[j][3];
        const float dhh = dh[j] * kh;
        const float tc = tc4[j];
        const float dc = dcs[j] * kc + dhh * og * (1.f - tc * tc);
        dcs[j] = dc * fg;
        if (b < B) {
          float* dg = DG + (r0 + b) * 4 * H + jg;
          const float d0 = dc * gg * ig * (1.f - ig), d1 = dc * scp[j] * fg * (1.f - fg);
          const float d2 = dc * ig * (1.f - gg * gg), d3 = dhh * tc * og * (1.f - og);
          dg[0] = t2v_rnd(d0, rnd); dg[H] = t2v_rnd(d1, rnd); dg[2 * H] = t2v_rnd(d2, rnd); dg[3 * H] = t2v_rnd(d3, rnd);
          bs0 += d0; bs1 += d1; bs2 += d2; bs3 += d3;
          if (OP) {
            uint16_t* dg16 = (which ? dgd16 : dga16) + (r0 + b) * 4 * H + jg;
            dg16[0] = f16_sat(d0 * g_scale); dg16[H] = f16_sat(d1 * g_scale);
            dg16[2 * H] = f16_sat(d2 * g_scale); dg16[3 * H] = f16_sat(d3 * g_scale);
          }
        }
      }
      {
        float* st = bias_s + blq * 32 + u;           // [gate][warp][unit]; the previous cell's readers passed a BAR_EPI barrier since
        st[0] = bs0; st[128] = bs1; st[256] = bs2; st[384] = bs3;
      }
      named_bar(BAR_EPI, 128);
      if (etid == 0) { GT(cur_it, which ? 5 : 6); signal_counter(which ? cnt_gd : cnt_ga); TR(cur_it, which ? 2 : 6); }
      {
        const float* st = bias_s + blq * 128 + u;    // warp blq owns gate blq
        gb_acc[which] += (st[0] + st[32]) + (st[64] + st[96]);
      }
    };

    // ---- GEMM epilogue: drain TMEM, exchange the split-K partials (16 batch rows per rank), sum, write dX rows
    auto gemm_epi = [&](auto which_c, const int ts, const unsigned idx) {
      constexpr int which = decltype(which_c)::value;
      constexpr int CP = which ? RD_P : RA_P;                  // slot column pitch
      constexpr int NC = which ? XD_CPC : XA_CPC;              // live columns
      if (etid == 0) mbar_expect_tx(&recv_full[which], 4u * 16u * NC * 4u);
      mbar_wait(&acc_full[which], idx & 1u);
      tc_fence_after();
      if (etid == 0) TR(cur_it, which ? 7 : 11);
      {
        const int r4 = lane & 3, m4 = lane >> 2;
        // accumulator row (= output column) held by this lane: M=64 keeps rows 16q..16q+15 in lanes 0..15 of quarter q
        const int col0 = which ? (32 * q + 4 * m4) : (16 * q + 4 * m4);
        const bool live = which ? (col0 < NC) : (lane < 16 && col0 < NC);
        const uint32_t slot_off = (uint32_t)((rank * 16) * CP + col0) * 4u;
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
            // a_j = D[column col0 + j][batch half*32 + 4k + r4]
            const int dst = half * 2 + (k >> 2);
            if (live)
              st_async_v4(recv_remote[which][dst] + slot_off + (uint32_t)((4 * (k & 3) + r4) * CP) * 4u, full_remote[which][dst],
                          a0, a1, a2, a3);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free[which]);
      if (etid == 0) TR(cur_it, which ? 8 : 12);
      mbar_wait_cluster(&recv_full[which], idx & 1u);
      if (etid == 0) TR(cur_it, which ? 9 : 13);
      {
        const float* rv = which ? recv_d : recv_a;
        float* out = which ? d.DXD : d.DXA;
        const int ld = which ? XD_W : XA_W;
        constexpr int NC4 = NC / 4;
        for (int task = etid; task < 16 * NC4; task += 128) {
          const int bl = task / NC4, c4 = task - bl * NC4;
          const int b = 16 * rank + bl;
          const float* rp = rv + bl * CP + 4 * c4;
          const float4 x0 = *reinterpret_cast<const float4*>(rp);
          const float4 x1 = *reinterpret_cast<const float4*>(rp + 16 * CP);
          const float4 x2 = *reinterpret_cast<const float4*>(rp + 2 * 16 * CP);
          const float4 x3 = *reinterpret_cast<const float4*>(rp + 3 * 16 * CP);
          float4 o;
          o.x = (x0.x + x1.x) + (x2.x + x3.x); o.y = (x0.y + x1.y) + (x2.y + x3.y);
          o.z = (x0.z + x1.z) + (x2.z + x3.z); o.w = (x0.w + x1.w) + (x2.w + x3.w);
          if (OP) { o.x *= g_inv; o.y *= g_inv; o.z *= g_inv; o.w *= g_inv; }
          if (b < B) *reinterpret_cast<float4*>(out + ((long long)ts * B + b) * ld + NC * cid + 4 * c4) = o;
        }
      }
      named_bar(BAR_EPI, 128);
      if (etid == 0) { signal_counter(which ? cnt_xd : cnt_xa); TR(cur_it, which ? 10 : 14); }
    };

    for (int it = -1; it < To; ++it) {
      const int t = To - 1 - it, td = t - 1;
      cur_it = it;
      // D1(td): needs dXD[td+1] (the dXD epilogue of the previous iteration)
      if (td >= 0) cell_bwd(std::integral_constant<int, 1>{}, td, NCTA * (unsigned)(it + 1), 0u, 0u);
      // A2(t): needs dXD[t], DXA[t+1] and dq[t]
      if (it >= 0) cell_bwd(std::integral_constant<int, 0>{}, t, NCTA * (unsigned)(it + 1), NCTA * (unsigned)it, NCTA * (unsigned)(it + 1));
      if (td >= 0) gemm_epi(std::integral_constant<int, 1>{}, td, (unsigned)(it + 1));
      if (it >= 0) gemm_epi(std::integral_constant<int, 0>{}, t, (unsigned)it);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = 16 * rank + blq + 4 * j;
      if (b < B) { d.dCa[(long long)b * H + jg] = dca[j]; d.dCd[(long long)b * H + jg] = dcd[j]; }
    }
    if (d.gb_att && d.gb_dec) {                      // this CTA's 16 batch rows -> the [4096] bias gradients (4 ranks add up)
      atomicAdd(d.gb_att + blq * H + jg, gb_acc[0]);
      atomicAdd(d.gb_dec + blq * H + jg, gb_acc[1]);
    }
  } else if (warp >= 8) {
    // =========================================================================== attention backward (two CTAs per utterance)
    const int atid = tid - 256, aw = warp - 8;
    const int b = 2 * cid + (rank >> 1), hh = rank & 1;
    const bool active = b < B;
    const int Th = (Ti + 1) >> 1;
    const int i0 = hh * Th;
    const int nrow = hh ? (Ti - Th) : Th;
    // partner (rows of the other half) sends its adjoint-conv outputs that fall into my rows: if I own [0,Th) these are
    // s in [Th-15, Th) clipped at 0 -> my local rows [Th-cnt, Th), cnt = min(15, Th); if I own [Th,Ti): s in [Th, Th+15)
    // clipped at Ti -> my local rows [0, cnt), cnt = min(15, Ti-Th)
    const int halo_cnt = hh ? min(HALO, Ti - Th) : min(HALO, Th);
    int len = Ti;
    if (active && s.in_lens) { const long long l = s.in_lens[b]; len = l < Ti ? (int)l : Ti; }
    const int a = atid & 127;
    float* wpad = small + SM_WPAD; float* cpad = small + SM_CPAD; float* wt_s = small + SM_WT; float* dwv_s = small + SM_DWV;
    float* de_s = small + SM_DE; float* P_s = small + SM_P; float* G_s = small + SM_G; float* adj_s = small + SM_ADJ;
    float* halo_s = small + SM_HALO; float* spart = small + SM_SPART; float* q_s = small + SM_QS;
    const float va = s.v[a];
    for (int i = atid; i < 2 * KS * NF; i += 256) wcT[i] = s.Wconv[i];
    if constexpr (OP) {
      // W_loc^T as the B operand of df = dpre W_loc: element (filter c, a) at K block a / 64, row c, chunk ((a % 64) / 8) ^ (c & 7);
      // W_conv as the B operand of the adjoint products: row n = 32 ch + tap (tap 31 = zero row), 32 filters = 64-byte rows (SW64)
      for (int i = atid; i < AD * NF; i += 256) {
        const int aa = i >> 5, c = i & 31, kk = aa & 63;
        wl16T[(aa >> 6) * (32 * 64) + c * 64 + ((((kk >> 3) ^ (c & 7)) << 3) | (kk & 7))] = t2v_f16_bits(s.Wloc[i]);
      }
      for (int i = atid; i < 64 * NF; i += 256) {
        const int n = i >> 5, c = i & 31, ch = n >> 5, k = n & 31;
        wc16[swz64(n, c)] = (k < KS) ? t2v_f16_bits(s.Wconv[(ch * KS + k) * NF + c]) : (uint16_t)0;
      }
      for (int i = atid; i < 2 * 64 * 64; i += 256) dp16[i] = 0;
      for (int i = atid; i < 32 * 64; i += 256) f16T[i] = 0;
      for (int i = atid; i < 64 * 32; i += 256) df16[i] = 0;
      fence_proxy_async();
    } else {
      for (int i = atid; i < AD * NF; i += 256) WlocS[i] = s.Wloc[i];
    }
    for (int i = atid; i < SM_TOTAL; i += 256) small[i] = 0.f;
    const uint32_t partner_spart = mapa(smem_u32(spart), (uint32_t)(rank ^ 1));
    const uint32_t partner_x = mapa(smem_u32(x_full), (uint32_t)(rank ^ 1));
    const uint32_t partner_halo = mapa(smem_u32(halo_s), (uint32_t)(rank ^ 1));
    const uint32_t partner_h = mapa(smem_u32(h_full), (uint32_t)(rank ^ 1));
    const uint64_t pol_m = l2_policy(1);
    float gwl[16];                           // dW_loc[a][16*(atid>>7) .. +16) over this CTA's rows, all steps
#pragma unroll
    for (int c = 0; c < 16; ++c) gwl[c] = 0.f;
    float wcacc[8];                          // dW_conv[(channel atid>>7, taps 8*((atid>>5)&3) .. +8)][filter atid&31], all steps
#pragma unroll
    for (int m = 0; m < 8; ++m) wcacc[m] = 0.f;
    float dv_acc = 0.f;                      // dv[a] over rows of row group (atid>>7), all steps
    float dpm_acc[32];                       // d(processed memory)[row group rows][a], all steps (was 1 M atomics per step)
#pragma unroll
    for (int k = 0; k < 32; ++k) dpm_acc[k] = 0.f;
    unsigned seen_xd = 0, seen_xa = 0;
    named_bar(BAR_ATT, 256);

    for (int it = 0; it < To; ++it) {
      const int t = To - 1 - it;
      const bool has_next = t + 1 < To;
      const int par = it & 1;
      if (active) {
        // ---- forward data of this step (independent of the recurrence): alignments, cumulative weights, saved tanh tile
        if (atid == 0) {
          mbar_expect_tx(x_full, 4u);
          mbar_expect_tx(h_full, (uint32_t)halo_cnt * 2u * 4u);
          if (nrow > 0) {
            mbar_expect_tx(at_full, (uint32_t)nrow * AD * 4u);
            bulk_load_1d(Abuf, s.ASAVE + (((long long)t * B + b) * Ti + i0) * AD, (uint32_t)nrow * AD * 4u, at_full);
          }
        }
        if (t > 0 && atid < ED / 32) prefetch_l2(d.DHC + ((long long)(t - 1) * B + b) * (H + ED) + H + 32 * atid);   // next step's dctx term
        for (int i = atid; i < Ti; i += 256) {
          wpad[HALO + i] = (t > 0) ? __ldg(s.align + ((long long)b * To + (t - 1)) * Ti + i) : 0.f;
          cpad[HALO + i] = __ldg(s.CUM + ((long long)t * B + b) * Ti + i);
        }
        if (atid < nrow) wt_s[atid] = __ldg(s.align + ((long long)b * To + t) * Ti + i0 + atid);
        named_bar(BAR_ATT, 256);
        // ---- location conv recomputed on this CTA's rows (needed by dW_loc / dW_conv below)
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
            if constexpr (OP) {     // f^T [filter c][rows]: 8 consecutive rows = one 16-byte chunk of the K-major operand
              uint4 pk;
              pk.x = (uint32_t)t2v_f16_bits(acc[0]) | ((uint32_t)t2v_f16_bits(acc[1]) << 16);
              pk.y = (uint32_t)t2v_f16_bits(acc[2]) | ((uint32_t)t2v_f16_bits(acc[3]) << 16);
              pk.z = (uint32_t)t2v_f16_bits(acc[4]) | ((uint32_t)t2v_f16_bits(acc[5]) << 16);
              pk.w = (uint32_t)t2v_f16_bits(acc[6]) | ((uint32_t)t2v_f16_bits(acc[7]) << 16);
              *reinterpret_cast<uint4*>(f16T + c * 64 + ((((r0 >> 3) ^ (c & 7)) & 7) << 3)) = pk;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) f_s[(r0 + j) * FS + c] = acc[j];
            }
          }
        }
      }
      // ---- dXD[t] and DXA[t+1] complete device-wide
      if (atid == 0) {
        const unsigned n_xd = NCTA * (unsigned)(it + 1), n_xa = NCTA * (unsigned)it;
        TR(it, 16);
        if (seen_xd < n_xd) { wait_counter(cnt_xd, n_xd); seen_xd = n_xd; }
        if (seen_xa < n_xa) { wait_counter(cnt_xa, n_xa); seen_xa = n_xa; }
        TR(it, 17);
        GT(it, 0);
      }
      named_bar(BAR_ATT, 256);
      if (active) {
        // ---- total gradient wrt ctx_t (three consumers of ctx_t, model.py:357,375,383)
        for (int col = atid; col < ED; col += 256) {
          float v = __ldcg(d.DXD + ((long long)t * B + b) * XD_W + H + col) + __ldcg(d.DHC + ((long long)t * B + b) * (H + ED) + H + col);
          if (has_next) v += __ldcg(d.DXA + ((long long)(t + 1) * B + b) * XA_W + PD + col);
          dctx_s[col] = v;
          if (hh == 0) d.DCTX[((long long)t * B + b) * ED + col] = v;
        }
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(it, 18);
        // ---- dw_i = <dctx, memory_i> + (gradient wrt w_t from step t+1's location conv) + (gradient wrt cum_{t+1})
        if (OP && s.mem16 != nullptr) {
          // fp16 copy of the encoder memory (exact: the memory sits on the tf32 grid): half the L2 bytes of this reduction, which is
          // on the critical path of the step.  lane -> columns [8 lane, 8 lane + 8) and [256 + 8 lane, ...); warp -> rows aw + 8 k
          float dc[2][8];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float4 lo = *reinterpret_cast<const float4*>(dctx_s + 256 * j + lane * 8);
            const float4 hi = *reinterpret_cast<const float4*>(dctx_s + 256 * j + lane * 8 + 4);
            dc[j][0] = lo.x; dc[j][1] = lo.y; dc[j][2] = lo.z; dc[j][3] = lo.w;
            dc[j][4] = hi.x; dc[j][5] = hi.y; dc[j][6] = hi.z; dc[j][7] = hi.w;
          }
          const uint16_t* mb = reinterpret_cast<const uint16_t*>(s.mem16) + ((long long)b * Ti + i0) * ED + lane * 8;
          for (int rb = 0; rb < nrow; rb += 32) {
            uint4 mv[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rr = rb + aw + 8 * k;
              const bool ld = rr < nrow && (i0 + rr) < len;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                mv[k][j] = ld ? ldg_u4_hint(mb + (long long)rr * ED + 256 * j, pol_m) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rr = rb + aw + 8 * k;
              float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t w4[4] = {mv[k][j].x, mv[k][j].y, mv[k][j].z, mv[k][j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  acc0 = fmaf(dc[j][2 * i], t2v_f16_to_f32((uint16_t)(w4[i] & 0xFFFFu)), acc0);
                  acc1 = fmaf(dc[j][2 * i + 1], t2v_f16_to_f32((uint16_t)(w4[i] >> 16)), acc1);
                }
              }
              const float acc = warp_sum(acc0 + acc1);
              if (lane == 0 && rr < nrow) dwv_s[rr] = acc + G_s[rr] + P_s[rr];
            }
          }
        } else {
          float4 dc4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) dc4[j] = *reinterpret_cast<const float4*>(dctx_s + lane * 4 + 128 * j);
          const float* mb = s.mem + ((long long)b * Ti + i0) * ED + lane * 4;
          for (int rb = 0; rb < nrow; rb += 16) {
            float4 mv[2][4];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int rr = rb + aw + 8 * k;
              const bool ld = rr < nrow && (i0 + rr) < len;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                mv[k][j] = ld ? ldg_v4_hint(mb + (long long)rr * ED + 128 * j, pol_m) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int rr = rb + aw + 8 * k;
              float acc = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                acc += dc4[j].x * mv[k][j].x + dc4[j].y * mv[k][j].y + dc4[j].z * mv[k][j].z + dc4[j].w * mv[k][j].w;
              acc = warp_sum(acc);
              if (lane == 0 && rr < nrow) dwv_s[rr] = acc + G_s[rr] + P_s[rr];
            }
          }
        }
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(it, 19);
        // ---- softmax backward: s = sum_i w_i dw_i over both halves, de_i = w_i (dw_i - s)
        if (aw == 0) {
          float part = 0.f;
          for (int rr = lane; rr < nrow; rr += 32) part = fmaf(wt_s[rr], dwv_s[rr], part);
          part = warp_sum(part);
          if (lane == 0) {
            spart[par * 2 + hh] = part;
            st_async_f32(partner_spart + (uint32_t)(par * 2 + hh) * 4u, partner_x, part);
          }
        }
        mbar_wait_cluster(x_full, (unsigned)it & 1u);
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(it, 20);
        {
          const float ssum = spart[par * 2] + spart[par * 2 + 1];
          if (atid < nrow) de_s[atid] = wt_s[atid] * (dwv_s[atid] - ssum);
        }
        if (nrow > 0) mbar_wait(at_full, (unsigned)it & 1u);
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(it, 21);
        // ---- tanh / v backward on this CTA's rows: dpre (kept in place of the saved tanh), dq, dv, d(processed memory)
        {
          const int rbeg = (atid >> 7) * 32;
          float dq_acc = 0.f;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int rr = rbeg + k;
            if (rr < nrow) {
              const float g = de_s[rr];
              const float av = Abuf[rr * AD + a];
              const float dp = g * va * (1.f - av * av);
              Abuf[rr * AD + a] = dp;
              dpm_acc[k] += dp;
              dq_acc += dp;
              dv_acc = fmaf(g, av, dv_acc);
            }
          }
          q_s[(atid >> 7) * 128 + a] = dq_acc;
        }
        named_bar(BAR_ATT, 256);
        if (atid == 0) TR(it, 22);
        if (atid < AD) atomicAdd(d.DQ + ((long long)t * B + b) * AD + atid, q_s[atid] + q_s[128 + atid]);
      }
      named_bar(BAR_ATT, 256);
      if (atid == 0) { GT(it, 2); signal_counter(cnt_q); TR(it, 23); GT(it, 4); }
      // ================= off the critical path: weight gradients of the location layer, adjoint conv for step t-1
      if (active) {
        float t_inv = 1.f;
        if constexpr (OP) {
          // ---- op16: dW_loc and df on the tensor core.  dpre of this step (fp32, in the tanh tile) -> fp16 copy scaled by a power of
          // two that puts the largest |dpre| of the tile into [2^9, 2^10) (gradients are ~1e-8; df = dpre W_loc then stays below the
          // fp16 maximum), rows >= nrow zero.  Thread -> 16-byte chunks (row, 8 attention dims): chunk index atid + 256 i.
          float mx = 0.f;
          float4 dv4[4][2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ci = atid + 256 * i, row = ci >> 4, ch8 = ci & 15;
            const float4* src = reinterpret_cast<const float4*>(Abuf + row * AD + ch8 * 8);
            const bool live = row < nrow;
            dv4[i][0] = live ? src[0] : make_float4(0.f, 0.f, 0.f, 0.f);
            dv4[i][1] = live ? src[1] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
              mx = fmaxf(mx, fmaxf(fmaxf(fabsf(dv4[i][h2].x), fabsf(dv4[i][h2].y)), fmaxf(fabsf(dv4[i][h2].z), fabsf(dv4[i][h2].w))));
          }
          mx = warp_max(mx);
          float* tmax = small + SM_TMAX;
          if (lane == 0) tmax[aw] = mx;
          named_bar(BAR_ATT, 256);
          mx = fmaxf(fmaxf(fmaxf(tmax[0], tmax[1]), fmaxf(tmax[2], tmax[3])), fmaxf(fmaxf(tmax[4], tmax[5]), fmaxf(tmax[6], tmax[7])));
          int ex = 0;
          (void)frexpf(mx, &ex);
          const float t_scale = (mx > 0.f) ? exp2f((float)(10 - ex)) : 1.f;
          t_inv = 1.f / t_scale;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ci = atid + 256 * i, row = ci >> 4, ch8 = ci & 15;
            const float4 lo = dv4[i][0], hi = dv4[i][1];
            uint4 pk;
            pk.x = (uint32_t)t2v_f16_bits(lo.x * t_scale) | ((uint32_t)t2v_f16_bits(lo.y * t_scale) << 16);
            pk.y = (uint32_t)t2v_f16_bits(lo.z * t_scale) | ((uint32_t)t2v_f16_bits(lo.w * t_scale) << 16);
            pk.z = (uint32_t)t2v_f16_bits(hi.x * t_scale) | ((uint32_t)t2v_f16_bits(hi.y * t_scale) << 16);
            pk.w = (uint32_t)t2v_f16_bits(hi.z * t_scale) | ((uint32_t)t2v_f16_bits(hi.w * t_scale) << 16);
            *reinterpret_cast<uint4*>(dp16 + (ch8 >> 3) * (64 * 64) + row * 64 + (((ch8 & 7) ^ (row & 7)) << 3)) = pk;
          }
          fence_proxy_async();
          tc_fence_before();
          named_bar(BAR_ATT, 256);
          tc_fence_after();
          if (aw == 0 && elect_one()) {
            // df [64 rows, 32 filters] = dpre [rows][a] (K-major A) x W_loc^T [filters][a] (K-major B), K = 128 -> TMEM columns 176..207
            constexpr uint32_t id_df = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t ad = make_kmajor_sw128_desc(smem_u32(dp16) + kb * (64 * 128));
              const uint64_t bd = make_kmajor_sw128_desc(smem_u32(wl16T) + kb * (32 * 128));
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                tc_mma_f16(tmem_base + 176u, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), id_df, (kb | k4) ? 1u : 0u);
            }
            // dW_loc [128 a, 32 filters] = dpre^T (the SAME tile read MN-major: a contiguous, rows = K) x f^T [filters][rows]
            // (K-major B), K = 64 rows -> TMEM columns 144..175
            constexpr uint32_t id_wl = (1u << 4) | (1u << 15) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t am = make_mnmajor16_sw128_desc(smem_u32(dp16), 64 * 128);
            const uint64_t bf = make_kmajor_sw128_desc(smem_u32(f16T));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16(tmem_base + 144u, am + (uint64_t)(128 * k), bf + (uint64_t)(2 * k), id_wl, k ? 1u : 0u);
            tc_commit(t1_full);
          }
          __syncwarp();
          mbar_wait(t1_full, (unsigned)it & 1u);
          tc_fence_after();
          {
            const int q4 = aw & 3, chalf = aw >> 2;
            uint32_t wv[16];
            tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + 144u + (uint32_t)(16 * chalf), wv);
#pragma unroll
            for (int c = 0; c < 16; ++c) gwl[c] = fmaf(__uint_as_float(wv[c]), t_inv, gwl[c]);
            // M = 64: accumulator row 16 q + l sits in TMEM lane 32 q + l (l < 16)
            uint32_t fv[16];
            tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + 176u + (uint32_t)(16 * chalf), fv);
            if (lane < 16) {
              const int row = 16 * q4 + lane, c0 = 16 * chalf;
              float4* fo = reinterpret_cast<float4*>(f_s + row * FS + c0);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4)
                fo[c4] = make_float4(__uint_as_float(fv[4 * c4]) * t_inv, __uint_as_float(fv[4 * c4 + 1]) * t_inv,
                                     __uint_as_float(fv[4 * c4 + 2]) * t_inv, __uint_as_float(fv[4 * c4 + 3]) * t_inv);
#pragma unroll
              for (int c8 = 0; c8 < 2; ++c8) {
                uint4 pk;
                pk.x = (uint32_t)f16_sat(__uint_as_float(fv[8 * c8])) | ((uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 1])) << 16);
                pk.y = (uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 2])) | ((uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 3])) << 16);
                pk.z = (uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 4])) | ((uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 5])) << 16);
                pk.w = (uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 6])) | ((uint32_t)f16_sat(__uint_as_float(fv[8 * c8 + 7])) << 16);
                *reinterpret_cast<uint4*>(df16 + swz64(row, c0 + 8 * c8)) = pk;
              }
            }
          }
          fence_proxy_async();
          tc_fence_before();
          named_bar(BAR_ATT, 256);
          tc_fence_after();
          if (atid == 0) TR(it, 24);
          if (aw == 0 && elect_one()) {
            // adjoint-conv products U [64 rows, 2 x 32 taps] = df [rows][filters] x W_conv [(ch, tap)][filters], K = 32 (SW64 operands)
            // -> TMEM columns 176..239 (df has been drained)
            constexpr uint32_t id_u = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            const uint64_t ad = make_kmajor_sw64_desc(smem_u32(df16)), bd = make_kmajor_sw64_desc(smem_u32(wc16));
            tc_mma_f16(tmem_base + 176u, ad, bd, id_u, 0u);
            tc_mma_f16(tmem_base + 176u, ad + 2, bd + 2, id_u, 1u);
            tc_commit(t2_full);
          }
          __syncwarp();
        } else {
          // ---- dW_loc[a][c] += sum_rows dpre[row][a] f[row][c]   (thread: a, 16 of the 32 filters, all rows)
          {
            const int cb = (atid >> 7) * 16;
            for (int rr = 0; rr < nrow; ++rr) {
              const float dp = Abuf[rr * AD + a];
              const float4* fr = reinterpret_cast<const float4*>(f_s + rr * FS + cb);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const float4 f4 = fr[c4];
                gwl[4 * c4] = fmaf(dp, f4.x, gwl[4 * c4]); gwl[4 * c4 + 1] = fmaf(dp, f4.y, gwl[4 * c4 + 1]);
                gwl[4 * c4 + 2] = fmaf(dp, f4.z, gwl[4 * c4 + 2]); gwl[4 * c4 + 3] = fmaf(dp, f4.w, gwl[4 * c4 + 3]);
              }
            }
          }
          // ---- df[row][c] = sum_a dpre[row][a] W_loc[a][c]   (thread: filter c, 8 rows)
          float dfr[8];
          {
            const int c = atid & 31, r0 = (atid >> 5) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) dfr[j] = 0.f;
            if (r0 < nrow) {
#pragma unroll 2
              for (int a0 = 0; a0 < AD; a0 += 4) {
                const float w0 = WlocS[a0 * NF + c], w1 = WlocS[(a0 + 1) * NF + c], w2 = WlocS[(a0 + 2) * NF + c], w3 = WlocS[(a0 + 3) * NF + c];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 d4 = *reinterpret_cast<const float4*>(Abuf + (r0 + j) * AD + a0);
                  dfr[j] = fmaf(d4.x, w0, dfr[j]); dfr[j] = fmaf(d4.y, w1, dfr[j]);
                  dfr[j] = fmaf(d4.z, w2, dfr[j]); dfr[j] = fmaf(d4.w, w3, dfr[j]);
                }
              }
            }
          }
          named_bar(BAR_ATT, 256);             // every reader of f is done: f_s becomes df
          if (atid == 0) TR(it, 24);
          {
            const int c = atid & 31, r0 = (atid >> 5) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) f_s[(r0 + j) * FS + c] = (r0 + j < nrow) ? dfr[j] : 0.f;
          }
          named_bar(BAR_ATT, 256);
        }
        // ---- dW_conv[(ch,k)][c] += sum_rows df[row][c] x_ch[row + k - 15]: thread = (filter c, channel, 8 consecutive taps),
        // the taps share one sliding window of x (3 shared-memory loads per row instead of 16)
        {
          const int c = atid & 31, g = atid >> 5;
          const int kk0 = 8 * (g & 3);
          const float* x = ((g >> 2) ? cpad : wpad) + i0 + kk0;
          for (int rr0 = 0; rr0 < nrow; rr0 += 8) {
            float xw[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) xw[i] = x[rr0 + i];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float dfv = f_s[(rr0 + r) * FS + c];          // rows >= nrow hold zeros
#pragma unroll
              for (int e = 0; e < 8; ++e) wcacc[e] = fmaf(dfv, xw[r + e], wcacc[e]);
            }
          }
        }
        if (atid == 0) TR(it, 25);
        // ---- adjoint conv: dx[ch][s] = sum_{k,c} df[s-k+15][c] W_conv[c][ch][k], s in [i0-15, i0+nrow+15);
        // own rows stay here, the rows of the other half go to the partner CTA
        float* Us = Abuf;                    // op16: [2 channels][64 rows][US_P] products, staged in the (now dead) tanh tile
        if constexpr (OP) {
          mbar_wait(t2_full, (unsigned)it & 1u);
          tc_fence_after();
          if (t > 0) {
            uint32_t uv[32];
            tmem_ld32(tmem_base + ((uint32_t)((aw & 3) * 32) << 16) + 176u + (uint32_t)(32 * (aw >> 2)), uv);
            if (lane < 16) {
              float* dst = Us + ((aw >> 2) * TH_MAX + 16 * (aw & 3) + lane) * US_P;
#pragma unroll
              for (int k = 0; k < KS; ++k) dst[k] = __uint_as_float(uv[k]);
            }
          }
          tc_fence_before();
          named_bar(BAR_ATT, 256);
        }
        if (t > 0) {
          const int span = nrow + 2 * HALO;
          if (atid < 2 * span) {
            const int ch = atid / span, j = atid - ch * span;
            const int sidx = i0 - HALO + j;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if constexpr (OP) {
              // dx[ch][s] = sum_k U[s - k + 15][ch][k]: one anti-diagonal of the product tile (pitch 33: conflict-free)
              const float* u0 = Us + ch * TH_MAX * US_P;
              for (int k = 0; k < KS; ++k) {
                const int tt = j - k;
                if (tt >= 0 && tt < nrow) a0 += u0[tt * US_P + k];
              }
              a0 *= t_inv;
            } else {
              for (int k = 0; k < KS; ++k) {
                const int tt = j - k;
                if (tt >= 0 && tt < nrow) {
                  const float4* dfp = reinterpret_cast<const float4*>(f_s + tt * FS);
                  const float4* wk = reinterpret_cast<const float4*>(wcT + (ch * KS + k) * NF);
#pragma unroll
                  for (int c4 = 0; c4 < NF / 4; ++c4) {
                    const float4 x4 = dfp[c4], w4 = wk[c4];
                    a0 = fmaf(x4.x, w4.x, a0); a1 = fmaf(x4.y, w4.y, a1); a2 = fmaf(x4.z, w4.z, a2); a3 = fmaf(x4.w, w4.w, a3);
                  }
                }
              }
            }
            const float av = (a0 + a1) + (a2 + a3);
            if (sidx >= 0 && sidx < Ti) {
              if (sidx >= i0 && sidx < i0 + nrow) adj_s[ch * 64 + (sidx - i0)] = av;
              else st_async_f32(partner_halo + (uint32_t)((par * 2 + ch) * 64 + (hh ? sidx : sidx - Th)) * 4u, partner_h, av);
            }
          }
          mbar_wait_cluster(h_full, (unsigned)it & 1u);
          named_bar(BAR_ATT, 256);
          if (atid < nrow) {
            // rows that received a halo contribution from the partner: the last `halo_cnt` rows of the first half /
            // the first `halo_cnt` rows of the second half
            const bool got = hh ? (atid < halo_cnt) : (atid >= nrow - halo_cnt);
            const float h0 = got ? halo_s[(par * 2 + 0) * 64 + atid] : 0.f;
            const float h1 = got ? halo_s[(par * 2 + 1) * 64 + atid] : 0.f;
            P_s[atid] = adj_s[atid] + h0;
            G_s[atid] += adj_s[64 + atid] + h1;
          }
        }
      }
      if constexpr (OP) fence_proxy_async();     // generic writes into the tanh tile (products) before the next bulk load into it
      named_bar(BAR_ATT, 256);
      if (atid == 0) TR(it, 26);
    }
    // ---- sequence-long accumulators -> the partial buffers the engine reduces over (slot 0 of the utterance)
    if (active) {
      const int nck = (Ti + 31) / 32;
      const long long slot = (long long)b * nck;
      atomicAdd(d.dv_part + slot * AD + a, dv_acc);
#pragma unroll
      for (int k = 0; k < 32; ++k) {         // single writer per element; the caller zero-initialised the buffer
        const int rr = (atid >> 7) * 32 + k;
        if (rr < nrow) d.dpmem[((long long)b * Ti + i0 + rr) * AD + a] = dpm_acc[k];
      }
      float* wl = d.dwloc_part + slot * (AD * NF) + a * NF + (atid >> 7) * 16;
#pragma unroll
      for (int c = 0; c < 16; ++c) atomicAdd(wl + c, gwl[c]);
      {
        const int c = atid & 31, g = atid >> 5;
        const int kk0 = 8 * (g & 3);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (kk0 + e < KS) atomicAdd(d.dwconv_part + slot * (2 * KS * NF) + ((g >> 2) * KS + kk0 + e) * NF + c, wcacc[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

bool persist_bwd_enabled() {
  const char* e = getenv("T2V_PERSIST_BWD");
  return !(e && e[0] == '0');
}

}  // namespace

int t2v_encode_tmap_2d(CUtensorMap* map, const void* base, int esize, long long inner, long long rows,
                       long long row_stride_elems, int box_rows);

// Returns 0 when the loop was enqueued, 1 when the persistent kernel does not apply (the caller uses the per-step launches).
int t2v_decoder_bwd_persist(const T2VDecoderBwd* d, int t_hi, int t_lo, cudaStream_t stream) {
  const T2VDecoderSeq* s = &d->f;
  if (!persist_bwd_enabled()) return 1;
  if (!s->use_tc || s->B > 64 || s->Ti > 2 * TH_MAX || s->Ti < 1 || t_hi != s->To || t_lo != 0 || s->To < 2) return 1;
  if (!s->GA || !s->GD || !s->CPA || !s->CPD || !s->ASAVE || !d->dHq) return 1;
  const int op = d->op16 ? 1 : 0;
  if (op && (!d->DGA16 || !d->DGD16 || !d->WaTP16 || !d->WdTP16 || !d->dg_scale)) {
    t2v_set_error("op16 backward needs DGA16 / DGD16 / WaTP16 / WdTP16 / dg_scale");
    return -1;
  }
  void (*kernel)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, BwdParams) =
      op ? dec_persist_bwd_kernel<1> : dec_persist_bwd_kernel<0>;
  static int max_clusters_dev[16][2];
  static bool mc_init = false;
  if (!mc_init) { for (auto& row : max_clusters_dev) for (int& v : row) v = -1; mc_init = true; }
  int& max_clusters = max_clusters_dev[t2v_device_slot()][op];
  static bool attr_set[2] = {false, false};
  const int smem_bytes = op ? LY<1>::SMEM_BYTES : LY<0>::SMEM_BYTES;
  if (!attr_set[op]) {
    T2V_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[op] = true;
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
  if (max_clusters < 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    max_clusters = n;
  }
  if (max_clusters < NCLUSTER) return 1;

  BwdParams p;
  p.d = *d;
  p.counters = reinterpret_cast<unsigned*>(d->dHq);          // scratch [B,1024] floats, unused by this path otherwise
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  p.rotate = env_int("T2V_PERSIST_ROTATE", 1);
  p.dbg_skip = env_int("T2V_PERSIST_DBG_SKIP", 0);
  p.two_issuers = 0;       // removed (see the MMA issuer comment in the kernel)
  p.wa_hint = env_int("T2V_PERSIST_BWD_WA_HINT", 1);
  p.wd_hint = env_int("T2V_PERSIST_BWD_WD_HINT", 2);
  p.scale = d->dg_scale;
  p.trace = nullptr;
  p.gtrace = nullptr;
  const bool trace = getenv("T2V_PERSIST_TRACE") != nullptr && s->To >= TRACE_I0 + TRACE_STEPS + 2;
  const size_t trace_bytes = 2 * TRACE_STEPS * 32 * sizeof(long long);
  if (trace) {
    T2V_CUDA_CHECK(cudaMalloc(&p.trace, trace_bytes));
    T2V_CUDA_CHECK(cudaMemsetAsync(p.trace, 0, trace_bytes, stream));
    T2V_CUDA_CHECK(cudaMalloc(&p.gtrace, NCTA * 8 * sizeof(long long)));
    T2V_CUDA_CHECK(cudaMemsetAsync(p.gtrace, 0, NCTA * 8 * sizeof(long long), stream));
  }
  CUtensorMap tmWa, tmWd64, tmWd16, tmGA, tmGD;
  const long long rows = (long long)s->To * s->B;
  int r;
  p.packed = (d->WaTP && d->WdTP && env_int("T2V_PERSIST_PACKED", 1)) ? 1 : 0;
  if (op) {
    p.packed = 1;
    if ((r = t2v_encode_tmap_2d(&tmWa, d->WaTP16, 2, 64, (long long)NCTA * KC<1>::N * XA_CPC, 64, XA_CPC))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd64, d->WdTP16, 2, 64, (long long)NCTA * KC<1>::N * XD_CPC, 64, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd16, d->WdTP16, 2, 64, (long long)NCTA * KC<1>::N * XD_CPC, 64, 16))) return r;
  } else if (p.packed) {
    if ((r = t2v_encode_tmap_2d(&tmWa, d->WaTP, 4, 32, (long long)NCTA * KCH32 * XA_CPC, 32, XA_CPC))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd64, d->WdTP, 4, 32, (long long)NCTA * KCH32 * XD_CPC, 32, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd16, d->WdTP, 4, 32, (long long)NCTA * KCH32 * XD_CPC, 32, 16))) return r;
  } else {
    if ((r = t2v_encode_tmap_2d(&tmWa, d->WaT, 4, 4 * H, XA_W, 4 * H, XA_CPC))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd64, d->WdT, 4, 4 * H, XD_W, 4 * H, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmWd16, d->WdT, 4, 4 * H, XD_W, 4 * H, 16))) return r;
  }
  if (op) {
    if ((r = t2v_encode_tmap_2d(&tmGA, d->DGA16, 2, 4 * H, rows, 4 * H, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmGD, d->DGD16, 2, 4 * H, rows, 4 * H, 64))) return r;
  } else {
    if ((r = t2v_encode_tmap_2d(&tmGA, d->DGA, 4, 4 * H, rows, 4 * H, 64))) return r;
    if ((r = t2v_encode_tmap_2d(&tmGD, d->DGD, 4, 4 * H, rows, 4 * H, 64))) return r;
  }
  T2V_CUDA_CHECK(cudaMemsetAsync(p.counters, 0, 160 * sizeof(unsigned), stream));
  cfg.numAttrs = t2v_coop_enabled() ? 2 : 1;
  T2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmWa, tmWd64, tmWd16, tmGA, tmGD, p));
  T2V_COUNT_LAUNCH();
  if (trace) {      // debugging aid: not usable under stream capture
    static const char* names[31] = {"E:D1 start", "E:D1 inputs seen", "E:D1 DGD signalled", "E:A2 dq seen", "E:A2 dq staged",
                                    "E:A2 dHq done", "E:A2 DGA signalled", "E:D2 acc full", "E:D2 pushed", "E:D2 recv full",
                                    "E:D2 dXD signalled", "E:A3 acc full", "E:A3 pushed", "E:A3 recv full", "E:A3 DXA signalled",
                                    "E:A2 start", "T:pre done, wait", "T:inputs seen", "T:dctx done", "T:dw done", "T:s exchanged",
                                    "T:de + tanh tile", "T:dpre done", "T:dq signalled", "T:dWloc + df done", "T:dWconv done",
                                    "T:adjoint + halo done", "M:D2 first chunk", "M:D2 committed", "M:A3 first chunk", "M:A3 committed"};
    long long h[2 * TRACE_STEPS * 32];
    T2V_CUDA_CHECK(cudaStreamSynchronize(stream));
    T2V_CUDA_CHECK(cudaMemcpy(h, p.trace, trace_bytes, cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    {      // cross-CTA view of one iteration (global timer, ns): who signals late, who sees the counters late
      static long long g[NCTA * 8];
      T2V_CUDA_CHECK(cudaMemcpy(g, p.gtrace, sizeof(g), cudaMemcpyDeviceToHost));
      cudaFree(p.gtrace);
      static const char* gn[8] = {"T:inputs seen", "E:D1 inputs seen", "T:dq signal begin", "E:A2 inputs (dq) seen", "T:dq signal done",
                                  "E:D1 DGD signal begin", "E:A2 DGA signal begin", ""};
      long long base = 0;
      for (int c = 0; c < NCTA; ++c) if (g[c * 8 + 0] && (!base || g[c * 8 + 0] < base)) base = g[c * 8 + 0];
      fprintf(stderr, "[t2v persist bwd gtrace] iteration %d, ns relative to the earliest 'T:inputs seen'; per event: min / median / max (CTA of max)\n", TRACE_I0);
      for (int ev = 0; ev < 7; ++ev) {
        long long v[NCTA]; int arg = 0;
        for (int c = 0; c < NCTA; ++c) { v[c] = g[c * 8 + ev] - base; if (v[c] > v[arg]) arg = c; }
        long long srt[NCTA];
        memcpy(srt, v, sizeof(srt));
        for (int i = 1; i < NCTA; ++i) { long long x = srt[i]; int j = i - 1; while (j >= 0 && srt[j] > x) { srt[j + 1] = srt[j]; --j; } srt[j + 1] = x; }
        fprintf(stderr, "  %-26s %7lld %7lld %7lld  (CTA %d)   p90 %lld\n", gn[ev], srt[0], srt[NCTA / 2], srt[NCTA - 1], arg, srt[NCTA * 9 / 10]);
      }
    }
    for (int c = 0; c < 2; ++c) {
      const long long base = h[(c * TRACE_STEPS) * 32 + 17];     // "inputs seen" of the first traced iteration
      fprintf(stderr, "[t2v persist bwd trace] CTA %d: SM clocks relative to 'T:inputs seen' of iteration %d\n", c ? TRACE_CTA_B : 0, TRACE_I0);
      for (int ev = 0; ev < 31; ++ev) {
        fprintf(stderr, "  %-24s", names[ev]);
        for (int n = 0; n < TRACE_STEPS; ++n) fprintf(stderr, " %8lld", h[(c * TRACE_STEPS + n) * 32 + ev] ? h[(c * TRACE_STEPS + n) * 32 + ev] - base : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return 0;
}
