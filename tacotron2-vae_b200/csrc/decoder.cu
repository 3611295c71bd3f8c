// The decoder time loop (reference Decoder.decode / forward / inference, model.py:346-464) driven from C: per step
//   gates_a = XA[t] Wa^T -> LSTM cell (+dropout on h and c) -> q = h_att Wq^T -> fused attention -> gates_d = XD[t] Wd^T
//   -> LSTM cell, with every step's GEMM operand rows living in two time-major sequence buffers
//     XA[t] = [prenet_t | ctx_{t-1} | h_att_{t-1}]   (1792 wide = the attention_rnn input ++ hidden, model.py:357-359)
//     XD[t] = [h_att_t  | ctx_t     | h_dec_{t-1}]   (2560 wide = the decoder_rnn input ++ hidden,  model.py:375-377)
// so the concatenations of the reference become "write the result where the next GEMM reads it", and the same
// buffers are the saved activations of the batched weight-gradient GEMMs.  linear_projection / gate_layer are
// deferred to one GEMM over all steps under teacher forcing (nothing in the loop consumes them).
#include "t2v_common.cuh"
#include <stdlib.h>
#include "gemm_tc.h"
#include "../../include/t2v_b200.h"

int t2v_decoder_fwd_persist(const T2VDecoderSeq* s, int t_begin, int t_end, cudaStream_t stream);   // decoder_persist.cu
int t2v_decoder_bwd_persist(const T2VDecoderBwd* d, int t_hi, int t_lo, cudaStream_t stream);         // decoder_persist_bwd.cu
int t2v_decoder_infer_persist(const T2VDecoderInfer* d, int t_begin, int t_end, cudaStream_t stream);  // decoder_persist.cu

namespace {

constexpr int H = 1024, XA_W = 1792, XD_W = 2560, AD = 128, ED = 512, PD = 256;
constexpr unsigned SITE_ATT_H = 10, SITE_ATT_C = 11, SITE_DEC_H = 12, SITE_DEC_C = 13, SITE_PRENET0 = 3, SITE_PRENET1 = 4;

struct StepGemm {
  bool tc;
  T2VGemmTcPlan plan;
  const float* A; long long lda; int a_k0;
  const float* W; long long ldw;
  int M, N, K, splits;
  long long split_stride, ldd;
};

// D[parts][M][N] = A[a_row0 + m, a_k0 + k] * W[n, k]
int setup_gemm(StepGemm* g, bool tc, const float* A, long long lda, long long a_rows, int a_k0, const float* W,
               long long ldw, int M, int N, int K, int splits, long long ldd = 0) {
  if (ldd == 0) ldd = N;
  g->ldd = ldd;
  g->tc = tc; g->A = A; g->lda = lda; g->a_k0 = a_k0; g->W = W; g->ldw = ldw; g->M = M; g->N = N; g->K = K;
  g->split_stride = (long long)M * ldd;
  if (tc) {
    const int total = t2v_ceil_div(K, 32);
    while (splits > 1 && total % splits) --splits;
    g->splits = splits;
    // inner extents are the true K so that TMA zero-fills the tail chunk on both operands
    int r = t2v_gemm_tc_plan(&g->plan, A, lda, a_rows, (long long)a_k0 + K, W, ldw, N, K, ldd, M, N, K, 1, 0, 0, a_k0, 0, 4,
                             splits, g->split_stride, 0, 1.f, 128);
    static const bool hint = !(getenv("T2V_L2_HINT") && getenv("T2V_L2_HINT")[0] == '0');
    g->plan.p.b_evict_last = hint ? 1 : 0;     // step weights are re-read 800 times: keep them in L2
    g->plan.p.b_independent = 1;               // weights: prefetch before the dependency wait
    g->plan.p.pdl = 1;
    return r;
  }
  g->splits = 1;
  return 0;
}
int run_gemm(const StepGemm* g, long long a_row0, float* D, cudaStream_t st) {
  if (g->tc) return t2v_gemm_tc_run(&g->plan, (int)a_row0, 0, D, nullptr, st);
  return t2v_gemm_f32(g->A + a_row0 * g->lda + g->a_k0, g->lda, 1, g->W, g->ldw, 1, D, g->ldd, g->M, g->N, g->K, 1.f,
                      0.f, nullptr, 1, 0, 0, 0, st);
}
#define CHK(expr) do { int r__ = (expr); if (r__) return r__; } while (0)

// split-K factors of the four big step GEMMs (att fwd, dec fwd, dec bwd, att bwd); T2V_SPLITS="a,b,c,d" overrides for tuning
int step_splits(int which) {
  static int v[4] = {4, 4, 16, 16};
  static bool init = false;
  if (!init) {
    init = true;
    if (const char* e = getenv("T2V_SPLITS")) sscanf(e, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]);
    for (int& x : v) x = x < 1 ? 1 : (x > 16 ? 16 : x);
  }
  return v[which];
}

// T2V_STEP_PROFILE=1: CUDA events after every launch of a few mid-sequence steps; averages printed to stderr.
struct StepProfiler {
  static constexpr int MAXE = 16, NSTEPS = 16;
  bool on = false; int t_first = 0, slot = 0, step = -1, nslots = 0;
  cudaEvent_t ev[NSTEPS][MAXE + 1];
  const char* names[MAXE];
  void begin(int t_begin, int t_end) {
    on = getenv("T2V_STEP_PROFILE") != nullptr && (t_end - t_begin) >= NSTEPS + 8;
    if (!on) return;
    nslots = 0;
    t_first = t_begin + 4;
    for (int i = 0; i < NSTEPS; ++i) for (int j = 0; j <= MAXE; ++j) cudaEventCreate(&ev[i][j]);
  }
  bool active(int idx) const { return on && idx >= 0 && idx < NSTEPS; }
  void step_begin(int idx, cudaStream_t st) { step = idx; slot = 0; if (active(step)) cudaEventRecord(ev[step][0], st); }
  void mark(const char* name, cudaStream_t st) {
    if (!active(step) || slot >= MAXE) return;
    names[slot] = name; ++slot; cudaEventRecord(ev[step][slot], st); if (slot > nslots) nslots = slot;
  }
  void end(const char* title, cudaStream_t st) {
    if (!on) return;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[t2v step profile] %s (avg over %d steps, event-to-event us)\n", title, NSTEPS);
    float total = 0;
    for (int j = 0; j < nslots; ++j) {
      float acc = 0;
      for (int i = 0; i < NSTEPS; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i][j], ev[i][j + 1]); acc += ms; }
      fprintf(stderr, "   %-22s %8.2f\n", names[j], acc * 1e3f / NSTEPS);
      total += acc * 1e3f / NSTEPS;
    }
    fprintf(stderr, "   %-22s %8.2f\n", "sum", total);
    for (int i = 0; i < NSTEPS; ++i) for (int j = 0; j <= MAXE; ++j) cudaEventDestroy(ev[i][j]);
  }
};
static StepProfiler g_prof;
#define PMARK(name) g_prof.mark(name, st)

struct FwdPlans { StepGemm ga, gq, gd; };

// The attention chain (attention_rnn gates -> cell -> query -> attention) is the only true recurrence of the decoder:
// the decoder_rnn chain (gates -> cell) of step t consumes h_att_t / ctx_t but feeds nothing back into the attention
// chain under teacher forcing (its output only reaches the deferred mel/gate projection), so it runs on a second
// stream, one step behind, overlapped with the attention chain of step t+1.  Events order the hand-offs; the same code
// is captured into the train-step CUDA graph as two parallel branches.
struct TwoChains {
  cudaStream_t s2 = nullptr, s3 = nullptr;      // s3: side stream of the backward loop (location backward of the attention)
  cudaEvent_t ev[9] = {};
  int init() {
    if (s2) return 0;
    T2V_CUDA_CHECK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    T2V_CUDA_CHECK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
    for (int i = 0; i < 9; ++i) T2V_CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    return 0;
  }
  int fork(cudaStream_t st) {      // s2 joins after everything already enqueued on st
    T2V_CUDA_CHECK(cudaEventRecord(ev[2], st));
    T2V_CUDA_CHECK(cudaStreamWaitEvent(s2, ev[2], 0));
    return 0;
  }
  int join(cudaStream_t st) {
    T2V_CUDA_CHECK(cudaEventRecord(ev[3], s2));
    T2V_CUDA_CHECK(cudaStreamWaitEvent(st, ev[3], 0));
    return 0;
  }
};
static TwoChains g_chains;
static bool two_chains_enabled() {
  static const bool on = !(getenv("T2V_TWO_CHAINS") && getenv("T2V_TWO_CHAINS")[0] == '0');
  return on;
}

int make_fwd_plans(const T2VDecoderSeq* s, FwdPlans* P) {
  const bool tc = s->use_tc != 0;
  const long long rows = (long long)(s->To + 1) * s->B;
  CHK(setup_gemm(&P->ga, tc, s->XA, XA_W, rows, 0, s->Wa, XA_W, s->B, 4 * H, XA_W, step_splits(0)));
  CHK(setup_gemm(&P->gq, tc, s->XD, XD_W, rows, 0, s->Wq, H, s->B, AD, H, 8));
  CHK(setup_gemm(&P->gd, tc, s->XD, XD_W, rows, 0, s->Wd, XD_W, s->B, 4 * H, XD_W, step_splits(1)));
  return 0;
}

// attention chain of step t: attention_rnn gates + cell, query projection, fused attention (Decoder.decode, model.py:357-374)
// location term of step t's attention (needs only step t-1's alignments): off the recurrence
int fwd_step_loc(const T2VDecoderSeq* s, int t, cudaStream_t st) {
  const int B = s->B, Ti = s->Ti;
  return t2v_attn3_loc_fwd(t > 0 ? s->align + (long long)(t - 1) * Ti : nullptr, (long long)s->To * Ti,
                           s->CUM + (long long)t * B * Ti, s->pmem, s->Wconv, s->Wloc, s->ebuf + (long long)B * Ti, B, Ti, st);
}
static bool attn_legacy_mode() {
  static const bool on = getenv("T2V_ATTN_MODE") != nullptr;     // "row"/"split"/"cluster": the single-stream kernels
  return on;
}
// loc_done: event to wait for before the attention (the side stream's location kernel), or nullptr = run it in order
int fwd_step_att(const T2VDecoderSeq* s, const FwdPlans* P, int t, float* parts, cudaStream_t st, cudaEvent_t loc_done = nullptr) {
  const int B = s->B, Ti = s->Ti;
  if (!attn_legacy_mode() && !loc_done) CHK(fwd_step_loc(s, t, st));
  const long long r0 = (long long)t * B, r1 = (long long)(t + 1) * B;
  const float p_att = s->training ? s->p_att : 0.f;
  const float* mk = s->drop_masks ? s->drop_masks + (long long)t * 4 * B * H : nullptr;
  const unsigned long long dbase = s->drop_masks ? 0ull : (unsigned long long)t * B * H;
  CHK(run_gemm(&P->ga, r0, parts, st));
  PMARK("gemm_att");
  CHK(t2v_lstm_pointwise_fwd(parts, P->ga.splits, P->ga.split_stride, 4 * H, nullptr, 0, s->ba1, s->ba2,
                             s->CA + r0 * H, H,
                             s->XA + r1 * XA_W + (PD + ED), XA_W,          // h_att -> next step's recurrent input
                             s->XD + r0 * XD_W, XD_W,                      // h_att -> decoder_rnn input / query
                             s->CA + r1 * H, H,
                             s->GA ? s->GA + r0 * 4 * H : nullptr, s->CPA ? s->CPA + r0 * H : nullptr, nullptr, 0,
                             mk, mk ? mk + (long long)B * H : nullptr, s->seed, SITE_ATT_H, SITE_ATT_C, p_att, dbase,
                             nullptr, 0, B, H, s->use_tc, st));
  PMARK("cell_att");
  CHK(run_gemm(&P->gq, r0, s->qparts, st));
  PMARK("gemm_q");
  if (attn_legacy_mode()) {
    CHK(t2v_attn2_fwd(s->qparts, P->gq.splits, P->gq.split_stride, t > 0 ? s->align + (long long)(t - 1) * Ti : nullptr,
                      (long long)s->To * Ti, s->CUM + r0 * Ti, s->CUM + r1 * Ti, s->pmem, s->mem, s->Wconv, s->Wloc, s->v,
                      s->in_lens, s->mask_value, s->ebuf, s->align + (long long)t * Ti, (long long)s->To * Ti,
                      s->XD + r0 * XD_W + H, XD_W, s->XA + r1 * XA_W + PD, XA_W,
                      s->ASAVE ? s->ASAVE + r0 * Ti * AD : nullptr, B, Ti, s->use_tc, st));
  } else {
    if (loc_done) T2V_CUDA_CHECK(cudaStreamWaitEvent(st, loc_done, 0));
    CHK(t2v_attn3_row_fwd(s->qparts, P->gq.splits, P->gq.split_stride, s->ebuf + (long long)B * Ti, s->CUM + r0 * Ti,
                          s->CUM + r1 * Ti, s->mem, s->v, s->in_lens, s->mask_value, s->align + (long long)t * Ti,
                          (long long)s->To * Ti,
                          s->XD + r0 * XD_W + H, XD_W,                       // ctx_t -> decoder_rnn input
                          s->XA + r1 * XA_W + PD, XA_W,                      // ctx_t -> next attention_rnn input
                          s->ASAVE ? s->ASAVE + r0 * Ti * AD : nullptr, B, Ti, s->use_tc, st));
  }
  PMARK("attention");
  return 0;
}
// decoder_rnn chain of step t (model.py:375-381)
int fwd_step_dec(const T2VDecoderSeq* s, const FwdPlans* P, int t, float* parts, cudaStream_t st) {
  const int B = s->B;
  const long long r0 = (long long)t * B, r1 = (long long)(t + 1) * B;
  const float p_dec = s->training ? s->p_dec : 0.f;
  const float* mk = s->drop_masks ? s->drop_masks + (long long)t * 4 * B * H : nullptr;
  const unsigned long long dbase = s->drop_masks ? 0ull : (unsigned long long)t * B * H;
  CHK(run_gemm(&P->gd, r0, parts, st));
  PMARK("gemm_dec");
  CHK(t2v_lstm_pointwise_fwd(parts, P->gd.splits, P->gd.split_stride, 4 * H, nullptr, 0, s->bd1, s->bd2,
                             s->CD + r0 * H, H,
                             s->XD + r1 * XD_W + (H + ED), XD_W,           // h_dec -> next step's recurrent input
                             nullptr, 0, s->CD + r1 * H, H,
                             s->GD ? s->GD + r0 * 4 * H : nullptr, s->CPD ? s->CPD + r0 * H : nullptr, nullptr, 0,
                             mk ? mk + 2LL * B * H : nullptr, mk ? mk + 3LL * B * H : nullptr, s->seed, SITE_DEC_H,
                             SITE_DEC_C, p_dec, dbase, nullptr, 0, B, H, s->use_tc, st));
  PMARK("cell_dec");
  return 0;
}
// one full decoder step on one stream (free-running inference: the next prenet input depends on h_dec, no overlap)
int fwd_step(const T2VDecoderSeq* s, const FwdPlans* P, int t, cudaStream_t st) {
  CHK(fwd_step_att(s, P, t, s->parts, st));
  CHK(fwd_step_dec(s, P, t, s->parts, st));
  return 0;
}

__global__ void stop_flags_kernel(const float* __restrict__ O, long long ld, int col, float thr, int t, int B,
                                  int* __restrict__ n_frames) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float g = O[(long long)b * ld + col];
  if (n_frames[b] < 0 && 1.f / (1.f + expf(-g)) > thr) n_frames[b] = t + 1;
}

}  // namespace

// 1 when the last t2v_decoder_fwd_steps call of this thread enqueued the persistent loop kernel (which also emits HCHI / HCLO),
// 0 when it issued the per-step launches
static thread_local int g_last_path = 0;
T2V_API int t2v_decoder_last_path(void) { return g_last_path; }
static thread_local int g_last_bwd_path = 0;
T2V_API int t2v_decoder_last_bwd_path(void) { return g_last_bwd_path; }     // same for t2v_decoder_bwd_steps (DGA16 / DGD16 valid)

T2V_API int t2v_decoder_fwd_steps(const T2VDecoderSeq* s, int t_begin, int t_end, cudaStream_t stream) {
  T2V_ARG_CHECK(s && s->B > 0 && s->Ti > 0 && s->To > 0, "shape");
  T2V_ARG_CHECK(t_begin >= 0 && t_end <= s->To && t_begin <= t_end, "step range");
  g_last_path = 0;
  if (!getenv("T2V_STEP_PROFILE")) {
    // the whole loop as one persistent kernel (decoder_persist.cu) when the problem fits it; 1 = not applicable
    const int r = t2v_decoder_fwd_persist(s, t_begin, t_end, stream);
    if (r == 0) g_last_path = 1;
    if (r != 1) return r;
  }
  FwdPlans P;
  CHK(make_fwd_plans(s, &P));
  g_prof.begin(t_begin, t_end);
  const bool two = two_chains_enabled() && !g_prof.on && (t_end - t_begin) > 1;
  if (!two) {
    for (int t = t_begin; t < t_end; ++t) {
      g_prof.step_begin(t - g_prof.t_first, stream);
      CHK(fwd_step(s, &P, t, stream));
    }
    g_prof.end("decoder forward step", stream);
    return 0;
  }
  CHK(g_chains.init());
  CHK(g_chains.fork(stream));
  float* parts_dec = s->parts + 16LL * s->B * 4 * H;         // second half of the split-K workspace
  const bool side = !attn_legacy_mode();
  cudaStream_t sl = g_chains.s3;
  if (side) {                                                 // location term of the first step
    T2V_CUDA_CHECK(cudaStreamWaitEvent(sl, g_chains.ev[2], 0));
    CHK(fwd_step_loc(s, t_begin, sl));
    T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[5 + (t_begin & 1)], sl));
  }
  for (int t = t_begin; t < t_end; ++t) {
    CHK(fwd_step_att(s, &P, t, s->parts, stream, side ? g_chains.ev[5 + (t & 1)] : nullptr));
    T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[t & 1], stream));
    T2V_CUDA_CHECK(cudaStreamWaitEvent(g_chains.s2, g_chains.ev[t & 1], 0));
    CHK(fwd_step_dec(s, &P, t, parts_dec, g_chains.s2));
    if (side && t + 1 < t_end) {                              // next step's location term, under its GEMM / cell / query
      T2V_CUDA_CHECK(cudaStreamWaitEvent(sl, g_chains.ev[t & 1], 0));
      CHK(fwd_step_loc(s, t + 1, sl));
      T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[5 + ((t + 1) & 1)], sl));
    }
  }
  CHK(g_chains.join(stream));
  if (side) {
    T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[4], sl));
    T2V_CUDA_CHECK(cudaStreamWaitEvent(stream, g_chains.ev[4], 0));
  }
  return 0;
}

T2V_API int t2v_decoder_bwd_steps(const T2VDecoderBwd* d, int t_hi, int t_lo, cudaStream_t st) {
  const T2VDecoderSeq* s = &d->f;
  T2V_ARG_CHECK(s->B > 0 && s->Ti > 0 && s->To > 0, "shape");
  T2V_ARG_CHECK(t_lo >= 0 && t_hi <= s->To && t_lo <= t_hi, "step range");
  g_last_bwd_path = 0;
  if (!getenv("T2V_STEP_PROFILE")) {
    const int r = t2v_decoder_bwd_persist(d, t_hi, t_lo, st);      // the whole reverse loop as one persistent kernel; 1 = n/a
    if (r == 0) g_last_bwd_path = 1;
    if (r != 1) return r;
  }
  const bool tc = s->use_tc != 0;
  const int B = s->B, Ti = s->Ti, To = s->To;
  const float p_att = s->training ? s->p_att : 0.f, p_dec = s->training ? s->p_dec : 0.f;
  StepGemm gxd, gxa, ghq;
  const long long rows = (long long)To * B;
  CHK(setup_gemm(&gxd, tc, d->DGD, 4 * H, rows, 0, d->WdT, 4 * H, B, XD_W, 4 * H, step_splits(2)));
  CHK(setup_gemm(&gxa, tc, d->DGA, 4 * H, rows, 0, d->WaT, 4 * H, B, XA_W, 4 * H, step_splits(3)));
  CHK(setup_gemm(&ghq, tc, d->DQ, AD, rows, 0, d->WqT, AD, B, H, AD, 1));
  g_prof.begin(t_lo, t_hi);
  const bool two = two_chains_enabled() && !g_prof.on && (t_hi - t_lo) > 1;
  cudaStream_t sd = st;                                       // stream of the decoder_rnn chain
  float* parts_dec = s->parts;
  cudaStream_t sl = st;                                       // stream of the attention location backward
  if (two) {
    CHK(g_chains.init());
    CHK(g_chains.fork(st));
    sd = g_chains.s2;
    sl = g_chains.s3;
    T2V_CUDA_CHECK(cudaStreamWaitEvent(sl, g_chains.ev[2], 0));     // the fork event
    parts_dec = s->parts + 16LL * B * 4 * H;
  }
  float* de_buf = d->dw_part + 4LL * B * Ti;
  for (int t = t_hi - 1; t >= t_lo; --t) {
    g_prof.step_begin(t - g_prof.t_first, st);
    const long long r0 = (long long)t * B, r1 = (long long)(t + 1) * B;
    const bool has_next = (t + 1 < To);
    float* dxd = d->DXD + r0 * XD_W;                          // full sequence [To,B,2560]: the decoder_rnn chain may run ahead
    const float* dxd_next = d->DXD + r1 * XD_W;
    const float* mk = s->drop_masks ? s->drop_masks + (long long)t * 4 * B * H : nullptr;
    const unsigned long long dbase = s->drop_masks ? 0ull : (unsigned long long)t * B * H;
    // ---- decoder_rnn chain: depends only on DHC[t] and on its own previous step (h_dec part of dXD[t+1])
    CHK(t2v_lstm_pointwise_bwd(d->DHC + r0 * (H + ED), H + ED, has_next ? dxd_next + (H + ED) : nullptr, XD_W, nullptr, 0,
                               d->dCd, s->GD + r0 * 4 * H, s->CPD + r0 * H, s->CD + r0 * H, H, d->DGD + r0 * 4 * H, 4 * H,
                               mk ? mk + 2LL * B * H : nullptr, mk ? mk + 3LL * B * H : nullptr, s->seed, SITE_DEC_H,
                               SITE_DEC_C, p_dec, dbase, nullptr, 0, B, H, s->use_tc, sd));
    PMARK("cell_dec_bwd");
    CHK(run_gemm(&gxd, r0, parts_dec, sd));
    PMARK("gemm_dXD");
    CHK(t2v_sum_parts(parts_dec, gxd.splits, gxd.split_stride, dxd, (long long)B * XD_W, sd));
    PMARK("sum_parts");
    if (two) {
      T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[t & 1], sd));
      T2V_CUDA_CHECK(cudaStreamWaitEvent(st, g_chains.ev[t & 1], 0));
    }
    // ---- attention chain: attention backward, query backward, attention_rnn cell + gates backward
    float* dw_out = d->dwprev + (long long)(t & 1) * B * Ti;
    float* gcum_next = d->gcum + (long long)(t & 1) * B * Ti;
    const float* a_save = s->ASAVE + r0 * Ti * AD;
    CHK(t2v_attn2_bwd_ctx(dxd + H, XD_W, d->DHC + r0 * (H + ED) + H, H + ED, has_next ? d->DXA + r1 * XA_W + PD : nullptr,
                          XA_W, d->DCTX + r0 * ED, d->dw_part, s->mem, s->in_lens, B, Ti, st));
    PMARK("attention_bwd_ctx");
    // the previous step's location backward produced this step's dw_in / gcum_prev (and read de_buf)
    if (two && t + 1 < t_hi) T2V_CUDA_CHECK(cudaStreamWaitEvent(st, g_chains.ev[5 + ((t + 1) & 1)], 0));
    CHK(t2v_attn2_bwd_dq(has_next ? d->dwprev + (long long)((t + 1) & 1) * B * Ti : nullptr, dw_out,
                         d->gcum + (long long)((t + 1) & 1) * B * Ti, gcum_next, d->dw_part, de_buf,
                         s->align + (long long)t * Ti, (long long)To * Ti, a_save, s->v, d->dpmem, d->DQ + r0 * AD, d->dv_part,
                         B, Ti, st));
    PMARK("attention_bwd_dq");
    if (two) {
      T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[7 + (t & 1)], st));
      T2V_CUDA_CHECK(cudaStreamWaitEvent(sl, g_chains.ev[7 + (t & 1)], 0));
    }
    CHK(t2v_attn2_bwd_loc(dw_out, gcum_next, de_buf, t > 0 ? s->align + (long long)(t - 1) * Ti : nullptr, (long long)To * Ti,
                          s->CUM + r0 * Ti, a_save, s->Wconv, s->Wloc, s->v, d->dwloc_part, d->dwconv_part, B, Ti, sl));
    if (two) T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[5 + (t & 1)], sl));
    PMARK("attention_bwd_loc");
    CHK(run_gemm(&ghq, r0, d->dHq, st));
    PMARK("gemm_dHq");
    CHK(t2v_lstm_pointwise_bwd(dxd, XD_W, has_next ? d->DXA + r1 * XA_W + (PD + ED) : nullptr, XA_W, d->dHq, H, d->dCa,
                               s->GA + r0 * 4 * H, s->CPA + r0 * H, s->CA + r0 * H, H, d->DGA + r0 * 4 * H, 4 * H, mk,
                               mk ? mk + (long long)B * H : nullptr, s->seed, SITE_ATT_H, SITE_ATT_C, p_att, dbase,
                               nullptr, 0, B, H, s->use_tc, st));
    PMARK("cell_att_bwd");
    CHK(run_gemm(&gxa, r0, s->parts, st));
    PMARK("gemm_dXA");
    CHK(t2v_sum_parts(s->parts, gxa.splits, gxa.split_stride, d->DXA + r0 * XA_W, (long long)B * XA_W, st));
    PMARK("sum_parts2");
  }
  if (two) {
    CHK(g_chains.join(st));
    T2V_CUDA_CHECK(cudaEventRecord(g_chains.ev[4], sl));
    T2V_CUDA_CHECK(cudaStreamWaitEvent(st, g_chains.ev[4], 0));
  }
  g_prof.end("decoder backward step", st);
  return 0;
}

// free-running decode (Decoder.inference, model.py:428-464; the notebook / synthesizer.py:139-154 loop), batched,
// with per-row stop bookkeeping on the device instead of a host sync per step
T2V_API int t2v_decoder_infer_steps(const T2VDecoderInfer* d, int t_begin, int t_end, cudaStream_t st) {
  const T2VDecoderSeq* s = &d->f;
  T2V_ARG_CHECK(s->B > 0 && s->Ti > 0 && s->To > 0, "shape");
  T2V_ARG_CHECK(t_begin >= 0 && t_end <= s->To && t_begin <= t_end, "step range");
  if (!getenv("T2V_STEP_PROFILE")) {
    const int r = t2v_decoder_infer_persist(d, t_begin, t_end, st);     // the whole free-running loop as one persistent kernel; 1 = n/a
    if (r != 1) return r;
  }
  const bool tc = s->use_tc != 0;
  const int B = s->B;
  FwdPlans P;
  CHK(make_fwd_plans(s, &P));
  const long long rows = (long long)(s->To + 1) * B;
  StepGemm gp1, gp2, gph, gpc;
  // prenet layer 1 reads the previous step's mel from O (84-wide rows), layer 2 reads P1
  CHK(setup_gemm(&gp1, tc, d->O, 84, (long long)s->To * B, 0, d->Wp1, 80, B, PD, 80, 1));
  CHK(setup_gemm(&gp2, tc, d->P1 + (long long)B * PD, PD, B, 0, d->Wp2, PD, B, PD, PD, 1));
  CHK(setup_gemm(&gph, tc, s->XD, XD_W, rows, H + ED, d->Wpg, H + ED, B, 81, H, 1, 84));
  CHK(setup_gemm(&gpc, tc, s->XD, XD_W, rows, H, d->Wpg + H, H + ED, B, 81, ED, 1, 84));
  float* p1 = d->P1;                       // [B,256] pre-activation scratch
  float* p1a = d->P1 + (long long)B * PD;  // [B,256] layer-1 output
  for (int t = t_begin; t < t_end; ++t) {
    const long long r0 = (long long)t * B, r1 = (long long)(t + 1) * B;
    float* xa_pre = s->XA + r0 * XA_W;
    const float* pm = d->prenet_masks ? d->prenet_masks + (long long)t * 2 * B * PD : nullptr;
    const unsigned long long pbase = pm ? 0ull : (unsigned long long)t * B * PD;
    if (t == 0) {
      // go frame is all zeros (model.py:241-247): relu(0 W) = 0 -> layer 1 output is 0
      CHK(t2v_fill(p1a, (long long)B * PD, 0.f, st));
    } else {
      CHK(run_gemm(&gp1, r0 - B, p1, st));
      CHK(t2v_relu_drop_fwd(p1, p1a, PD, B, PD, pm, s->seed, SITE_PRENET0, 0.5f, pbase, s->use_tc, st));
    }
    CHK(run_gemm(&gp2, 0, p1, st));
    CHK(t2v_relu_drop_fwd(p1, xa_pre, XA_W, B, PD, pm ? pm + (long long)B * PD : nullptr, s->seed, SITE_PRENET1, 0.5f,
                          pbase, s->use_tc, st));
    CHK(fwd_step(s, &P, t, st));
    // mel/gate projection of [h_dec_t | ctx_t]
    float* o = d->O + r0 * 84;
    if (tc) {
      CHK(t2v_gemm_tc_run(&gph.plan, (int)r1, 0, o, d->bpg, st));
      T2VGemmTcPlan acc = gpc.plan;
      acc.p.epi_atomic = 1;
      CHK(t2v_gemm_tc_run(&acc, (int)r0, 0, o, nullptr, st));
    } else {
      CHK(t2v_gemm_f32(s->XD + r1 * XD_W + (H + ED), XD_W, 1, d->Wpg, H + ED, 1, o, 84, B, 81, H, 1.f, 0.f, d->bpg, 1, 0,
                       0, 0, st));
      CHK(t2v_gemm_f32(s->XD + r0 * XD_W + H, XD_W, 1, d->Wpg + H, H + ED, 1, o, 84, B, 81, ED, 1.f, 1.f, nullptr, 1, 0,
                       0, 0, st));
    }
    if (d->n_frames) {
      stop_flags_kernel<<<t2v_ceil_div(B, 128), 128, 0, st>>>(o, 84, 80, d->gate_threshold, t, B, d->n_frames);
      T2V_COUNT_LAUNCH();
      T2V_LAUNCH_CHECK();
    }
  }
  return 0;
}
