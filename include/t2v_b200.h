/* t2v_b200 -- C ABI of the B200-native Tacotron2-VAE hot path (libt2v_b200.so).
 *
 * The reference (jinhan/tacotron2-vae) has no FFI layer: its hot path is PyTorch modules (model.py, modules.py,
 * layers.py, stft.py, loss_function.py, distributed.py).  This header is the boundary a maintainer binds instead
 * (ctypes stub in INTEGRATION.md): plain device pointers, sizes and a cudaStream_t -- no torch types.  Each entry
 * names the reference code it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (or int64 where stated) owned by the caller; the library never
 *     allocates, frees or retains device memory;
 *   - every call only enqueues work on `stream` (no host synchronisation) and returns
 *       0  ok | <0 argument error, nothing launched | >0 cudaError_t ;  t2v_last_error() has the message;
 *   - "padded channels-last" = a reference [B,C,T] tensor stored as rows [B*(T+4), C], two zero rows before and
 *     after each utterance;  "tb rows" = time-major rows (t*B + b);
 *   - dropout sites take an explicit keep-mask (float 0/1) or, when it is NULL, a counter-based RNG keyed by
 *     (seed, site, logical index) that t2v_materialize_mask() reproduces for the test oracle.
 */
#ifndef T2V_B200_H
#define T2V_B200_H
#include <cuda_runtime_api.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ---------------------------------------------------------------------------------------- */
const char* t2v_last_error(void);
int t2v_version(void);
unsigned long long t2v_launch_count(void);      /* kernels launched by this library since the last reset */
void t2v_reset_launch_count(void);
void t2v_add_launch_count(unsigned long long n);   /* account for launches replayed from a captured CUDA graph */

/* ---- GEMMs: every nn.Linear / Conv1d / LSTM-cell matmul of model.py goes through one of these two ------------ */
/* exact fp32 FFMA, fully strided: C[m,n] = alpha*sum_k A[m*a_rs+k*a_cs]*B[n*b_rs+k*b_cs] + beta*C + bias[n] */
int t2v_gemm_f32(const float* A, long long a_rs, long long a_cs, const float* B, long long b_rs, long long b_cs,
                 float* C, long long c_rs, int M, int N, int K, float alpha, float beta, const float* bias, int batch,
                 long long a_bs, long long b_bs, long long c_bs, cudaStream_t stream);
/* tcgen05 (TMA + TMEM) GEMM, both operands K-major; see csrc/gemm_tc.cu for the "taps" (row-shifted K segments). */
int t2v_gemm_tc(const void* A, long long lda, long long a_rows, long long a_inner, const void* B, long long ldb,
                long long b_rows, long long b_inner, float* D, long long ldd, const float* bias, int M, int N, int k_sub,
                int taps, int a_tap_rowshift, int b_tap_stride, int a_k0, int b_k0, int esize, int splits,
                long long split_stride, int epi_atomic, float alpha, int bn_hint, cudaStream_t stream);

/* split (error-compensated) product for fp32 operands given as hi + lo parts on the tf32 grid (x = hi + lo):
   D = alpha * (A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T) (+ bias) through one accumulator -- fp32-level accuracy from tf32 tensor-core
   products at 3x the K loop.  Used where tf32 rounding dominated the error of the outputs: the deferred mel / gate projection
   (model.py:383-388) and the Postnet forward (model.py:105-148). */
/* the same three-term product over fp16 hi / lo pairs (kind::f16: 64 K-columns per 128-byte row, half the operand bytes of the tf32
   form, which is bound by the L2 -> shared-memory operand stream); element counts / strides in fp16 elements */
int t2v_gemm_tc_split3_16(const void* A_hi, const void* A_lo, long long lda, long long a_rows, long long a_inner, const void* B_hi,
                          const void* B_lo, long long ldb, long long b_rows, long long b_inner, float* D, long long ldd,
                          const float* bias, int M, int N, int k_sub, int taps, int a_tap_rowshift, int b_tap_stride, int a_k0,
                          int b_k0, float alpha, int bn_hint, cudaStream_t stream);
int t2v_gemm_tc_split3(const float* A_hi, const float* A_lo, long long lda, long long a_rows, long long a_inner, const float* B_hi,
                       const float* B_lo, long long ldb, long long b_rows, long long b_inner, float* D, long long ldd,
                       const float* bias, int M, int N, int k_sub, int taps, int a_tap_rowshift, int b_tap_stride, int a_k0,
                       int b_k0, float alpha, int bn_hint, cudaStream_t stream);

/* row-reduction form with MN-major operands (no transposed copies): D[n_a,n_b] (+)= alpha * sum_{r<rows} A[a_row0+r, i] * B[b_row0+r, j];
   the weight-gradient GEMMs dW = dY^T X of every nn.Linear / Conv1d (autograd of model.py:91-148).  epi: 0 store, 1 atomicAdd,
   2 non-atomic += (splits == 1); splits > 1 needs epi 1 (D pre-initialised).  taps > 1: the Conv1d weight gradient in tap-major form,
   D[n_a, taps*n_b], column block t reduces against B rows [b_row0 + t, b_row0 + t + rows) (model.py:105-177: k = 5 convolutions). */
int t2v_gemm_tc_rowred(const float* A, long long lda, int n_a, long long a_row0, const float* B, long long ldb, int n_b,
                       long long b_row0, float* D, long long ldd, long long rows, int splits, long long split_stride,
                       int epi, float alpha, int taps, cudaStream_t stream);

/* batched form: D[z][i,j] = alpha * sum_{r<rows} A[z*a_batch_rows + r, i] * B[r, z*b_batch_cols + j] (plain stores, n_a <= 128,
   n_b % 256 == 0): d(memory)[b] = alignments[b]^T dctx[:, b, :], the backward of attention_context = bmm(attention_weights, memory)
   (model.py:84-85) summed over the decoder steps. */
int t2v_gemm_tc_rowred_batched(const float* A, long long lda, int n_a, long long a_batch_rows, const float* B, long long ldb,
                               int n_b, long long b_batch_cols, float* D, long long ldd, long long d_batch_stride,
                               long long rows, int batch, float alpha, cudaStream_t stream);

/* the same row reduction over 16-bit operands (fp16 / bf16 copies, kind::f16, 64 reduction rows per stage): the decoder's big weight
   gradients dW = DG^T X from the fp16 copies the persistent loops leave behind (XA16 / XD16, DGA16 / DGD16).  alpha_dev (nullable):
   device scalar multiplied onto alpha (the inverse gradient scale).  taps: as t2v_gemm_tc_rowred.  n_split > 0 (multiple of 256):
   product columns [n_split, n_b) are written to D2 (row stride ldd2) instead of D: [dW_ih | dW_hh] of an LSTMCell (model.py:363-380)
   land in the two parameters' own gradient tensors. */
int t2v_gemm_tc_rowred16(const void* A, long long lda, int n_a, long long a_row0, const void* B, long long ldb, int n_b,
                         long long b_row0, float* D, long long ldd, long long rows, int splits, int epi, float alpha,
                         const float* alpha_dev, int fmt, int taps, float* D2, long long ldd2, int n_split, cudaStream_t stream);

/* ---- text embedding (model.py:474,528) -- integer gather, bit exact ------------------------------------------- */
int t2v_embedding_fwd(const long long* ids, const float* table, float* out_padded, int B, int T, int C, int n_symbols,
                      int rnd, cudaStream_t stream);
int t2v_embedding_bwd(const long long* ids, const float* dout_padded, float* dtable, int B, int T, int C,
                      cudaStream_t stream);

/* ---- BatchNorm1d/2d + activation + dropout (model.py:143-146,176-177; modules.py:69-71) ------------------------ */
int t2v_col_stats(const float* x, long long rows, int C, int period, int lo, int hi, int mode, double* out0,
                  double* out1, cudaStream_t stream);
int t2v_bn_finalize(const double* sum, const double* sumsq, double n, int C, float eps, float momentum, float* mean,
                    float* invstd, float* running_mean, float* running_var, long long* num_batches_tracked,
                    cudaStream_t stream);
int t2v_bn_eval_prepare(const float* running_mean, const float* running_var, int C, float eps, float* mean,
                        float* invstd, cudaStream_t stream);
/* out_lo (nullable): tf32(x - out) -- the low part of the split operand x = out + out_lo of an error-compensated tensor-core GEMM.
   out_hi16 / out_lo16 (nullable pair): the same split as two fp16 arrays (x = hi + lo, 22 significant bits) for t2v_gemm_tc_split3_16 */
int t2v_bn_act_fwd(const float* y, float* out, float* out_lo, long long rows, int C, int period, int lo, int hi, const float* mean,
                   const float* invstd, const float* gamma, const float* beta, int act, const float* drop_mask,
                   unsigned long long seed, unsigned int site, float p, int T, int rnd, void* out_hi16, void* out_lo16,
                   cudaStream_t stream);
/* x * scale -> fp16 hi + fp16 lo (n % 4 == 0): weights (scale = a power of two that lifts the lo part out of the fp16 subnormals) and
   first-layer inputs of the 16-bit split GEMMs */
int t2v_split16(const float* x, void* hi, void* lo, long long n, float scale, cudaStream_t stream);
int t2v_bn_act_bwd_reduce(const float* dout, const float* y, long long rows, int C, int period, int lo, int hi,
                          const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                          const float* drop_mask, unsigned long long seed, unsigned int site, float p, int T,
                          double* dbeta_sum, double* dgamma_sum, cudaStream_t stream);
int t2v_bn_act_bwd_apply(const float* dout, const float* y, float* dy, long long rows, int C, int period, int lo, int hi,
                         const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                         const float* drop_mask, unsigned long long seed, unsigned int site, float p, int T,
                         const double* dbeta_sum, const double* dgamma_sum, double n, int use_batch_stats,
                         int rnd, cudaStream_t stream);
int t2v_double_to_float(const double* src, float* dst, int n, float beta, cudaStream_t stream);

/* ---- layout / packing helpers ------------------------------------------------------------------------------------ */
int t2v_copy2d(const float* src, long long s_rs, long long s_cs, float* dst, long long d_rs, long long rows, int cols,
               float beta, int rnd, cudaStream_t stream);
int t2v_transpose(const float* in, long long i_ld, float* out, long long o_ld, long long rows, int cols, int rnd,
                  cudaStream_t stream);
int t2v_conv1d_pack(const float* w, float* out, int Co, int Ci, int K, int flip, int rnd, cudaStream_t stream);
int t2v_conv1d_unpack_grad(const float* gk, float* gw, int Co, int Ci, int K, float beta, cudaStream_t stream);
int t2v_axpby(const float* x, float a, float* y, float b, long long n, cudaStream_t stream);
int t2v_bcast_add_rows(float* y, const float* v, long long rows, int C, int rows_per_batch, cudaStream_t stream);
int t2v_sum_rows_per_batch(const float* x, float* out, int B, int rows_per_batch, int C, float beta, cudaStream_t stream);
int t2v_sum_parts(const float* parts, int n_parts, long long part_stride, float* dst, long long n, cudaStream_t stream);
int t2v_fill(float* x, long long n, float v, cudaStream_t stream);
int t2v_bct_to_padded(const float* in, float* out, int B, int C, int T, float beta, cudaStream_t stream);
int t2v_padded_to_bct(const float* in1, const float* in2, float* out, int B, int C, int T, const long long* lens,
                      float fill, cudaStream_t stream);
int t2v_rows_tb_to_padded(const float* rows, long long ld, float* out, int B, int C, int T, int rnd, cudaStream_t stream);
int t2v_padded_to_rows_tb(const float* p1, const float* p2, const float* dgate, float* rows, long long ld, int B, int C,
                          int T, int rnd, cudaStream_t stream);
int t2v_bct_to_rows_tb_shift(const float* tgt, float* rows, int B, int C, int T, int rnd, cudaStream_t stream);
int t2v_gate_from_rows(const float* rows, long long ld, int col, float* gate, int B, int T, const long long* lens,
                       float fill, cudaStream_t stream);
int t2v_mask_padded_rows(float* x, int B, int C, int T, const long long* lens, cudaStream_t stream); /* model.py:515 */
int t2v_unpad_add(const float* in_padded, const float* add_vec, float* out, int B, int T, int C, int rnd,
                  cudaStream_t stream);
int t2v_round_tf32(float* x, long long n, cudaStream_t stream);
/* standard-normal samples from the counter-based RNG (the VAE eps, modules.py:19).  Every `seed` argument of this ABI is a
   value or, with bit 63 set, a device pointer to the value (so CUDA-graph replays can change it). */
int t2v_randn(float* out, long long n, unsigned long long seed, unsigned int site, cudaStream_t stream);   /* in-place round-to-nearest onto the tf32 grid */

/* ---- Prenet pointwise (model.py:91-102) and dropout-mask materialisation for the oracle ------------------------- */
int t2v_relu_drop_fwd(const float* x, float* out, long long o_rs, long long rows, int C, const float* mask,
                      unsigned long long seed, unsigned int site, float p, unsigned long long idx_base,
                      int rnd, cudaStream_t stream);
int t2v_relu_drop_bwd(const float* x, const float* dout, long long do_rs, float* dx, long long rows, int C,
                      const float* mask, unsigned long long seed, unsigned int site, float p,
                      unsigned long long idx_base, int rnd, cudaStream_t stream);
int t2v_materialize_mask(float* out, long long n, unsigned long long seed, unsigned int site, float p,
                         unsigned long long idx_base, cudaStream_t stream);

/* ---- recurrent cells: nn.LSTM / nn.LSTMCell / nn.GRU pointwise (model.py:171-190,357-381; modules.py:60-78) ----- */
int t2v_lstm_pointwise_fwd(const float* parts, int n_parts, long long part_stride, long long parts_rs, const float* pre,
                           long long pre_rs, const float* b1, const float* b2, const float* c_prev, long long cprev_rs,
                           float* h_out, long long hout_rs, float* h_out2, long long hout2_rs, float* c_out,
                           long long cout_rs, float* gates_save, float* cpre_save, float* seq_out, long long seq_rs,
                           const float* mask_h, const float* mask_c, unsigned long long seed, unsigned int site_h,
                           unsigned int site_c, float p, unsigned long long drop_base, const long long* lens, int t,
                           int B, int H, int rnd, cudaStream_t stream);
int t2v_lstm_pointwise_bwd(const float* dh1, long long dh1_rs, const float* dh2, long long dh2_rs, const float* dh3,
                           long long dh3_rs, float* dc, const float* gates_save, const float* cpre_save,
                           const float* c_prev, long long cprev_rs, float* dgates, long long dg_rs, const float* mask_h,
                           const float* mask_c, unsigned long long seed, unsigned int site_h, unsigned int site_c,
                           float p, unsigned long long drop_base, const long long* lens, int t, int B, int H,
                           int rnd, cudaStream_t stream);
/* encoder BiLSTM (model.py:171-190): one fused launch per time step for both directions (recurrent matvec + cell) */
int t2v_bilstm_step_fwd(const float* gx0, const float* gx1, long long gx_bs, const float* whh0, const float* whh1,
                        const float* bhh0, const float* bhh1, const float* hprev0, const float* hprev1, float* hnext0,
                        float* hnext1, float* c0, float* c1, float* seq0, float* seq1, long long seq_bs, float* gs0, float* gs1,
                        float* cs0, float* cs1, const long long* lens, int t0, int t1, int B, int H, cudaStream_t stream);
int t2v_bilstm_step_bwd(const float* dgn0, const float* dgn1, long long dg_bs, const float* whhT0, const float* whhT1,
                        const float* dout0, const float* dout1, long long dout_bs, float* dc0, float* dc1, const float* gs0,
                        const float* gs1, const float* cs0, const float* cs1, const float* cp0, const float* cp1, float* dgo0,
                        float* dgo1, const long long* lens, int t0, int t1, int B, int H, cudaStream_t stream);
/* whole-sequence persistent versions (ONE launch for all Ti steps of both directions, B <= 64): gx0 / gx1 [B*(Ti+4), 4H] hoisted input
   projections on the padded rows, seq [B*(Ti+4), 2H] output rows, gates [2,Ti,B,4H], cells [2,Ti+2,B,H] (zero-initialised), hbuf scratch
   [2,2,B,H], counters 64 x u32 (zeroed by the call); backward: dout [B,Ti,2H] -> dg0 / dg1 [B*(Ti+4), 4H] */
int t2v_bilstm_seq_fwd(const float* gx0, const float* gx1, const float* whh0, const float* whh1, const float* bhh0,
                       const float* bhh1, float* seq, float* gates, float* cells, float* hbuf, unsigned int* counters,
                       const long long* lens, int B, int H, int Ti, cudaStream_t stream);
int t2v_bilstm_seq_bwd(const float* whhT0, const float* whhT1, const float* dout, const float* gates, const float* cells,
                       float* dg0, float* dg1, unsigned int* counters, const long long* lens, int B, int H, int Ti,
                       cudaStream_t stream);
/* reference-encoder GRU (reference modules.py:60-80, nn.GRU batch_first, h0 = 0) as one persistent launch per pass, B <= 64:
   gi rows (b, t) at gi + b * gi_bs + t * 3H (x W_ih^T, no bias), hs [Tq+1,B,H] (slot 0 zero), save [Tq,B,4H] (r, z, n, gh_n),
   counter 32 x u32 (zeroed by the call); backward: dh_last [B,H] -> dgi (same row layout as gi) and dgh [Tq,B,3H] */
int t2v_gru_seq_fwd(const float* gi, long long gi_bs, const float* w_hh, const float* b_ih, const float* b_hh, float* hs,
                    float* save, unsigned int* counter, int B, int H, int Tq, cudaStream_t stream);
int t2v_gru_seq_bwd(const float* w_hh, const float* dh_last, const float* save, const float* hs, float* dgi, long long dgi_bs,
                    float* dgh, unsigned int* counter, int B, int H, int Tq, cudaStream_t stream);
/* same, with dh1 given as `dh1_parts` split-K partial buffers (stride dh1_pstride) that are summed on the fly */
int t2v_lstm_pointwise_bwd_parts(const float* dh1, long long dh1_rs, int dh1_parts, long long dh1_pstride, const float* dh2,
                                 long long dh2_rs, const float* dh3, long long dh3_rs, float* dc, const float* gates_save,
                                 const float* cpre_save, const float* c_prev, long long cprev_rs, float* dgates, long long dg_rs,
                                 const float* mask_h, const float* mask_c, unsigned long long seed, unsigned int site_h,
                                 unsigned int site_c, float p, unsigned long long drop_base, const long long* lens, int t,
                                 int B, int H, int rnd, cudaStream_t stream);
int t2v_gru_pointwise_fwd(const float* gi, long long gi_rs, const float* gh, const float* b_ih, const float* b_hh, const float* h_prev,
                          float* h_out, float* save, int B, int H, cudaStream_t stream);
int t2v_gru_pointwise_bwd(const float* dh, const float* save, const float* h_prev, float* dgi, long long dgi_rs, float* dgh,
                          float* dh_prev, int B, int H, int rnd, cudaStream_t stream);
int t2v_vae_reparam_fwd(const float* mulv, const float* eps, float* z, int B, int Z, int training, cudaStream_t stream);
int t2v_vae_reparam_bwd(const float* mulv, const float* eps, const float* dz, const float* dmu_ext, const float* dlv_ext,
                        float* dmulv, int B, int Z, int training, cudaStream_t stream);

/* ---- reference encoder im2col (CoordConv.py:37-74 + modules.py:45-71) ------------------------------------------- */
int t2v_im2col_3x3s2(const float* x, float* col, int N, int H, int W, int Ci, int coord, int rnd, cudaStream_t stream);
int t2v_col2im_3x3s2(const float* dcol, float* dx, int N, int H, int W, int Ci, cudaStream_t stream);

/* ---- fused location-sensitive attention step (model.py:31-88, 366-374) ------------------------------------------ */
int t2v_attn_step_fwd(const float* qparts, int n_qparts, long long qpart_stride, const float* w_prev, long long wprev_rs,
                      const float* cum_in, float* cum_out, const float* pmem, const float* mem, const float* w_conv,
                      const float* w_loc, const float* v, const long long* lens, float mask_value, float* w_out,
                      long long wout_rs, float* ctx_out1, long long ctx1_rs, float* ctx_out2, long long ctx2_rs,
                      float* a_save, int B, int Ti, int rnd, cudaStream_t stream);
int t2v_attn_step_bwd(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs, const float* dctx3,
                      long long dctx3_rs, const float* dw_in, float* dw_out, float* gcum, const float* w, long long w_rs,
                      const float* w_prev, long long wprev_rs, const float* cum_in, const float* a_save, const float* mem,
                      const float* w_conv, const float* w_loc, const float* v, const long long* lens, float* dmem,
                      float* dpmem, float* dq, float* dv_part, float* dwloc_part, float* dwconv_part, int B, int Ti,
                      int rnd, cudaStream_t stream);

/* version 2 of the attention step: work spread over (utterance x text-chunk) and (utterance x channel-chunk) CTAs */
int t2v_attn2_chunks(int Ti);
int t2v_attn2_fwd(const float* qparts, int n_qparts, long long qpart_stride, const float* w_prev, long long wprev_rs,
                  const float* cum_in, float* cum_out, const float* pmem, const float* mem, const float* w_conv,
                  const float* w_loc, const float* v, const long long* lens, float mask_value, float* e_buf, float* w_out,
                  long long wout_rs, float* ctx_out1, long long ctx1_rs, float* ctx_out2, long long ctx2_rs, float* a_save,
                  int B, int Ti, int rnd, cudaStream_t stream);
/* forward of the step in two launches: _loc computes pre[b,ti,:] = processed_memory + W_loc conv([w_prev; w_cum]) for a step
 * from the PREVIOUS step's alignments only (no query), so a caller may run it on a side stream under the attention_rnn
 * GEMM / cell / query GEMM; _row = tanh(q + pre).v -> mask -> softmax -> context, cumulative weights, saved activations. */
int t2v_attn3_loc_fwd(const float* w_prev, long long wprev_rs, const float* cum_in, const float* pmem, const float* w_conv,
                      const float* w_loc, float* pre, int B, int Ti, cudaStream_t stream);
int t2v_attn3_row_fwd(const float* qparts, int n_qparts, long long qpart_stride, const float* pre, const float* cum_in,
                      float* cum_out, const float* mem, const float* v, const long long* lens, float mask_value,
                      float* w_out, long long wout_rs, float* ctx_out1, long long ctx1_rs, float* ctx_out2,
                      long long ctx2_rs, float* a_save, int B, int Ti, int rnd, cudaStream_t stream);
/* backward of the step in three launches: _ctx (context reduction backward) and _dq (softmax / tanh backward -> dq,
 * d processed-memory, dv; also initialises dw_out = 0 and gcum_next = gcum_prev) are on the recurrence of the backward time
 * loop; _loc (location dense / conv backward, adjoint conv scattered into dw_out / gcum_next, dW_loc / dW_conv partials)
 * only has to finish before the NEXT step's _dq, so a caller may run it on a side stream.  de_buf [B,Ti] carries the
 * energies gradient from _dq to _loc.  t2v_attn2_bwd = the three in stream order. */
int t2v_attn2_bwd_ctx(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs, const float* dctx3,
                      long long dctx3_rs, float* dctx_out, float* dw_part, const float* mem, const long long* lens, int B,
                      int Ti, cudaStream_t stream);
int t2v_attn2_bwd_dq(const float* dw_in, float* dw_out, const float* gcum_prev, float* gcum_next, const float* dw_part,
                     float* de_buf, const float* w, long long w_rs, const float* a_save, const float* v, float* dpmem,
                     float* dq, float* dv_part, int B, int Ti, cudaStream_t stream);
int t2v_attn2_bwd_loc(float* dw_out, float* gcum_next, const float* de_buf, const float* w_prev, long long wprev_rs,
                      const float* cum_in, const float* a_save, const float* w_conv, const float* w_loc, const float* v,
                      float* dwloc_part, float* dwconv_part, int B, int Ti, cudaStream_t stream);
int t2v_attn2_bwd(const float* dctx1, long long dctx1_rs, const float* dctx2, long long dctx2_rs, const float* dctx3,
                  long long dctx3_rs, float* dctx_out, const float* dw_in, float* dw_out, const float* gcum_prev,
                  float* gcum_next, float* dw_part, float* de_buf, const float* w, long long w_rs, const float* w_prev,
                  long long wprev_rs, const float* cum_in, const float* a_save, const float* mem, const float* w_conv,
                  const float* w_loc, const float* v, const long long* lens, float* dpmem, float* dq, float* dv_part,
                  float* dwloc_part, float* dwconv_part, int B, int Ti, cudaStream_t stream);

/* ---- the decoder time loop: Decoder.decode / Decoder.forward / Decoder.inference (model.py:346-464) ------------- */
typedef struct T2VDecoderSeq {
  int B, Ti, To;                 /* batch, padded text length, number of decoder steps the buffers hold */
  int use_tc;                    /* 0: exact fp32 FFMA GEMMs ; 1: tcgen05 tf32 GEMMs */
  int training;                  /* dropout p_att/p_dec on h AND c of both cells (model.py:361-364,378-381) */
  float p_att, p_dec;
  unsigned long long seed;       /* RNG seed when drop_masks == NULL */
  const float* drop_masks;       /* [To,4,B,1024] keep masks (att_h, att_c, dec_h, dec_c) or NULL */
  float mask_value;              /* attention score_mask_value */
  const long long* in_lens;      /* [B] text lengths (attention mask) or NULL */
  const float *Wa, *ba1, *ba2;   /* attention_rnn: [4096,1792] = [weight_ih | weight_hh], bias_ih, bias_hh */
  const float *Wd, *bd1, *bd2;   /* decoder_rnn:   [4096,2560] */
  const float *Wq, *Wconv, *Wloc, *v;   /* query_layer [128,1024], location conv [32,2,31], dense [128,32], v [128] */
  const float *mem, *pmem;       /* encoder outputs [B,Ti,512], processed memory [B,Ti,128] */
  float *XA, *XD;                /* [(To+1),B,1792] = [prenet_t | ctx_{t-1} | h_att_{t-1}] ; [(To+1),B,2560] = [h_att_t | ctx_t | h_dec_{t-1}] */
  float *CA, *CD;                /* [(To+1),B,1024] cell states, slot t = state before step t */
  float *CUM;                    /* [(To+1),B,Ti] cumulative attention weights before step t */
  float *align;                  /* [B,To,Ti] */
  float *GA, *GD, *CPA, *CPD;    /* saved gates [To,B,4096] / pre-dropout cells [To,B,1024]; NULL at inference */
  float *ASAVE;                  /* [To,B,Ti,128] tanh activations; NULL at inference */
  float *parts, *qparts;         /* split-K workspaces: >= 32*B*4096 (two halves of 16 parts, one per chain) and 8*B*128 floats */
  float *ebuf;                   /* [B,Ti,129] scratch: [B,Ti] energies, then [B,Ti,128] location term + processed memory */
  const float *WaP, *WdP;        /* optional (NULL = unused): Wa / Wd re-tiled by t2v_pack_step_tiles (modes 0 / 1) for the persistent
                                    loop kernel: every TMA box is one contiguous 16 KB block instead of 128 strided 128-byte rows */
  float *HCHI, *HCLO;            /* optional [To,B,1536] each: [h_dec_t | ctx_t] on the operand grid (hi) and the low-order residual
                                    (x = x_hi + x_lo, both on the tf32 grid): the operands of the deferred mel / gate projection as ONE
                                    split (error-compensated) tensor-core GEMM.  Written by the persistent loop kernel only */
  int op16;                      /* 0: fp32 storage / tf32 math in the persistent loops; 1: fp16, 2: bf16 operand copies (kind::f16) */
  void *XA16, *XD16;             /* op16: 16-bit copies of XA / XD (same shapes), the tensor-core operands; zero-initialise */
  const void *WaP16, *WdP16;     /* op16: 16-bit re-tiled weights (t2v_pack_step_tiles16 modes 0 / 1) */
  const void *mem16;             /* op16, optional: fp16 copy of `mem` (exact for values on the tf32 grid) read by the context reduction:
                                    half the bytes and all Ti <= 128 rows of a column group in flight at once */
} T2VDecoderSeq;
int t2v_sizeof_decoder_structs(int which);   /* sizeof(T2VDecoderSeq / T2VDecoderBwd / T2VDecoderInfer) for which = 0 / 1 / 2: binding check */
int t2v_decoder_fwd_steps(const T2VDecoderSeq* s, int t_begin, int t_end, cudaStream_t stream);
int t2v_decoder_last_path(void);   /* 1: the last t2v_decoder_fwd_steps of this thread enqueued the persistent kernel (HCHI / HCLO valid) */
/* re-tile a decoder-step weight matrix into the order the persistent loop kernels stream it (same number of floats):
   mode 0: Wa [4096,1792] -> WaP ; 1: Wd [4096,2560] -> WdP ; 2: WaT [1792,4096] -> WaTP ; 3: WdT [2560,4096] -> WdTP */
int t2v_pack_step_tiles(const float* W, int mode, float* out, cudaStream_t stream);
/* power-of-two scale that maps max |x| to 2^target_log2: out[0] = s, out[1] = 1 / s (out[2..3] scratch); keeps fp16 copies of
   gradients (~1e-8) inside the fp16 range without touching their significands */
int t2v_grad_scale(const float* x, long long n, int target_log2, float* out, cudaStream_t stream);
/* 16-bit variant for op16 (fmt 1 = fp16, 2 = bf16): K chunks of 64 columns; modes 0 / 1 forward tiles, 2 / 3 backward tiles (W^T) */
int t2v_pack_step_tiles16(const float* W, int mode, void* out, int fmt, cudaStream_t stream);
/* contiguous fp32 -> 16-bit conversion of n elements (n % 8 == 0), multiplied by *scale_dev when given (fp16: saturating): the fp16
   operand copies of the convolution weight-gradient GEMMs */
int t2v_cvt16_scaled(const float* src, void* dst, long long n, int fmt, const float* scale_dev, cudaStream_t stream);
/* strided fp32 -> 16-bit conversion (fmt 1 = fp16, 2 = bf16), e.g. the prenet columns of XA into XA16 */
int t2v_cvt16_2d(const float* src, long long s_ld, void* dst, long long d_ld, long long rows, int cols, int fmt,
                 cudaStream_t stream);

typedef struct T2VDecoderBwd {
  T2VDecoderSeq f;               /* the forward description (same buffers) */
  const float *WaT, *WdT, *WqT;  /* transposed weights [1792,4096], [2560,4096], [1024,128] */
  const float *DHC;              /* [To,B,1536] grad wrt [h_dec_t | ctx_t] from linear_projection/gate_layer */
  float *DGA, *DGD;              /* out: pre-activation gate grads [To,B,4096] */
  float *DXA;                    /* out: grad wrt XA rows [To,B,1792] */
  float *DXD;                    /* [To,B,2560] grad wrt the XD rows (the decoder_rnn chain runs ahead of the attention chain) */
  float *dCa, *dCd;              /* [B,1024] running cell-state grads (zero-initialised by caller) */
  float *dwprev;                 /* ring [2,B,Ti] */
  float *gcum;                   /* ring [2,B,Ti] zero-initialised */
  float *dpmem;                  /* [B,Ti,128] accumulator (zero-initialised) */
  float *DCTX;                   /* out [To,B,512] total grad wrt ctx_t (d(memory) is one batched GEMM after the loop) */
  float *dw_part;                /* scratch [5,B,Ti]: 4 context partials + the energies gradient (de_buf) */
  float *DQ;                     /* out [To,B,128], zero-initialised (accumulated with atomics) */
  float *dHq;                    /* scratch [B,1024] */
  float *dv_part, *dwloc_part, *dwconv_part;   /* [B*nchunk,128], [B*nchunk,128*32], [B*nchunk,32*2*31] accumulators (zero-init), nchunk = t2v_attn2_chunks(Ti) */
  const float *WaTP, *WdTP;      /* optional (NULL = unused): WaT / WdT re-tiled by t2v_pack_step_tiles (modes 2 / 3) */
  int op16;                      /* 1: the two dX GEMMs of the loop run on fp16 copies (kind::f16): */
  void *DGA16, *DGD16;           /*   out: gate gradients x dg_scale[0] as saturating fp16, [To,B,4096] each (also the operands of the
                                      16-bit weight-gradient GEMMs after the loop) */
  const void *WaTP16, *WdTP16;   /*   fp16 re-tiled W^T (t2v_pack_step_tiles16 modes 2 / 3) */
  const float* dg_scale;         /*   device [2] = {s, 1/s}: power-of-two gradient scale (t2v_grad_scale) */
  float *gb_att, *gb_dec;        /* optional (NULL = unused), zero-initialised [4096] each: bias gradients of attention_rnn / decoder_rnn
                                    (column sums of DGA / DGD) accumulated by the persistent kernel; valid when
                                    t2v_decoder_last_bwd_path() == 1, the per-step path leaves them untouched */
} T2VDecoderBwd;
int t2v_decoder_bwd_steps(const T2VDecoderBwd* s, int t_hi, int t_lo, cudaStream_t stream);  /* t = t_hi-1 .. t_lo */
int t2v_decoder_last_bwd_path(void);   /* 1: the last t2v_decoder_bwd_steps of this thread enqueued the persistent kernel (DGA16 / DGD16 valid) */

typedef struct T2VDecoderInfer {
  T2VDecoderSeq f;
  const float *Wp1, *Wp2;        /* prenet [256,80], [256,256] */
  const float *Wpg, *bpg;        /* [81,1536] = [linear_projection ; gate_layer], bias [81] */
  const float *prenet_masks;     /* [n,2,B,256] or NULL (RNG; prenet dropout is always on, model.py:101) */
  float *O;                      /* [n,B,84] mel(80)+gate(1) rows, time-major */
  float *P1;                     /* scratch [B,256] x2 */
  float gate_threshold;
  int *n_frames;                 /* [B] first step whose sigmoid(gate) > threshold (+1), or n if never */
} T2VDecoderInfer;
/* free-running decode of steps [t_begin, t_end): one persistent kernel (prenet, both cells, attention, mel / gate projection, stop
   bookkeeping inside) when use_tc, B <= 64, Ti <= 128 and the range has >= 2 steps; otherwise the per-step launch sequence */
int t2v_decoder_infer_steps(const T2VDecoderInfer* s, int t_begin, int t_end, cudaStream_t stream);

/* ---- loss (loss_function.py:27-45) -------------------------------------------------------------------------------- */
int t2v_loss_fwd(const float* mel, const float* post, const float* tgt, long long n_mel, const float* gate,
                 const float* gtgt, long long n_gate, const float* mu, const float* logvar, long long n_z,
                 float kl_weight, double* acc, float* out, cudaStream_t stream);
int t2v_loss_bwd(const float* mel, const float* post, const float* tgt, long long n_mel, const float* gate,
                 const float* gtgt, long long n_gate, const float* mu, const float* logvar, long long n_z,
                 float kl_weight, const float* gout, float* dmel, float* dpost, float* dgate, float* dmu, float* dlogvar,
                 cudaStream_t stream);

/* ---- STFT / mel front-end pieces (stft.py:77-105, layers.py:75-92, audio_processing.py:77-83) ------------------- */
/* the whole front-end as ONE kernel for the reference recipe (filter 1024, hop 256, win 1024): reflect pad -> hann -> 1024-point FFT
   in shared memory (two real frames per complex transform) -> |X| -> mel filterbank -> log(max(., clip)); wav [B,S] -> out [B,n_mel,S/256+1].
   window [1024], twiddle [1024] complex, mel_basis [n_mel,513], band_lo / band_hi [n_mel] = non-zero bin range of every filter */
int t2v_stft_mel_fused(const float* wav, int B, int S, const float* window, const float* twiddle, const float* mel_basis,
                       const int* band_lo, const int* band_hi, float* out, int n_mel, int n_frames, float clip,
                       cudaStream_t stream);
int t2v_reflect_pad(const float* wav, float* out, int B, int S, int pad, long long ld, cudaStream_t stream);
int t2v_stft_mag(const float* ft, long long ft_ld, float* mag, long long mag_ld, long long rows, int nb, cudaStream_t stream);
int t2v_mel_log(const float* mel, long long mel_ld, float* out, int B, int n_mel, int n_frames, long long rows_per_batch,
                float clip, cudaStream_t stream);

/* ---- optimiser: clip_grad_norm_ + Adam(L2 weight decay) fused (train.py:171-172,226-229) ------------------------ */
int t2v_grad_sumsq(const float* g, long long n, float gscale, double* sumsq, cudaStream_t stream);
int t2v_adam_clip_step(float* p, float* g, float* m, float* v, long long n, const double* sumsq, float gscale,
                       float max_norm, float lr, float beta1, float beta2, float eps, float wd, int step, float* norm_out,
                       cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
