/* srnn_b200.h - C ABI of the B200 (sm_100a) SampleRNN training-step kernels.
 *
 * The reference (AlomdaElmasry/samplernn_pase) has no FFI: its hot path is Python that calls
 * torch.  This header is the boundary a maintainer of the reference would bind (ctypes stub in
 * INTEGRATION.md); each entry cites the reference code it replaces (file:line into the
 * reference repository).
 *
 * Conventions:
 *   - every entry returns 0 on success, a positive cudaError_t, or a negative SRNN_ERR_* code;
 *     srnn_last_error() returns a thread-local message for the last failure;
 *   - all pointers are raw DEVICE pointers unless the name ends in _host; the caller owns all
 *     memory including workspaces; nothing is allocated or freed by the library;
 *   - kernels are enqueued on the passed stream (a cudaStream_t cast to void*) and the library
 *     never synchronises;
 *   - "bf16" buffers are __nv_bfloat16; matrices are row-major with an explicit leading
 *     dimension in ELEMENTS; every bf16 leading dimension / batch stride must be a multiple of
 *     8 elements (16 bytes) and every bf16 base pointer 16-byte aligned (TMA requirement).
 */
#ifndef SRNN_B200_H
#define SRNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRNN_OK 0
#define SRNN_ERR_ARG (-1)     /* bad argument (shape, alignment, null pointer) */
#define SRNN_ERR_DEVICE (-2)  /* not an sm_100 device / driver entry point missing */

#define SRNN_ABI_VERSION 6

typedef void* srnn_stream_t; /* cudaStream_t */

const char* srnn_last_error(void);
int srnn_abi_version(void);
/* sm count and compute capability of the current device.  Host-side state of the library (SM count, function
 * attributes, occupancy answers) is cached per CUDA device and may be used from several host threads. */
int srnn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* how many clusters of `cluster_size` CTAs (dynamic shared memory and threads per CTA given) can be co-resident on the
 * current device (decides which decompositions the generation kernels may use) */
int srnn_probe_clusters(int32_t cluster_size, int32_t smem_bytes, int32_t threads, int32_t* max_clusters);

/* ---------------------------------------------------------------------------------------------
 * Quantiser - replaces SampleRNNQuantizer (utils.py:25-73)
 * ------------------------------------------------------------------------------------------- */
/* quantize_ulaw (utils.py:59-65), bit-exact with the reference op chain executed by torch on
 * CUDA.  Writes any of: int64 indices (the reference's API dtype), uint8 indices (internal,
 * saturated at 255) ; *overflow_count is incremented for every element that maps to index >=
 * q_levels (the reference would raise an index error downstream; SURVEY trap 3). */
int srnn_quantize_ulaw(const float* x, int64_t n, int64_t* idx_i64, uint8_t* idx_u8, int32_t* overflow_count,
                       srnn_stream_t stream);
/* quantize_linear (utils.py:48-54) with per-row min/max; rows x cols input; idx = long(norm * (q_levels - 1e-2) +
 * 0.005) with every step rounded to fp32 like the reference.  A constant row (0/0 in the reference) maps to 0. */
int srnn_quantize_linear(const float* x, int64_t rows, int64_t cols, int32_t q_levels, int64_t* idx_i64,
                         uint8_t* idx_u8, srnn_stream_t stream);
/* dequantize (utils.py:56-57,67-73) through a 256(+1)-entry table: out[i] = lut[idx[i]].
 * Exactly one of idx_i64 / idx_u8 is non-null; exactly one of out_f32 / out_bf16 is non-null. */
int srnn_dequantize_lut(const int64_t* idx_i64, const uint8_t* idx_u8, int64_t n, const float* lut, float* out_f32,
                        void* out_bf16, srnn_stream_t stream);
/* One-hot rows for the sample-level contraction (model.py:192-193 restated as a one-hot x table
 * product): onehot[i, :] = e_{idx[i]} as bf16, row length q (=256). */
int srnn_onehot_rows(const uint8_t* idx_u8, int64_t n, int32_t q, void* onehot_bf16, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / parameter preparation
 * ------------------------------------------------------------------------------------------- */
/* weight_norm forward (torch.nn.utils.weight_norm dim=0; model.py:135-138,183-186):
 * v is (R, A, B) contiguous fp32, g is (R); w[r,a,b] = g[r] * v[r,a,b] / ||v[r]||.
 * Writes bf16 copies of w in up to two GEMM layouts: element (r,a,b) goes to
 * out + r*s[0] + a*s[1] + b*s[2].  inv_norm (R) receives 1/||v[r]|| for the backward.
 * If g is NULL the kernel is a plain cast/permute (w = v). */
int srnn_weight_prep(const float* v, const float* g, int32_t R, int32_t A, int32_t B, void* out1_bf16,
                     const int64_t* s1, void* out2_bf16, const int64_t* s2, float* inv_norm, srnn_stream_t stream);
/* weight_norm backward: dw is read at dw + r*s[0] + a*s[1] + b*s[2] (fp32, GEMM layout);
 * dv (R,A,B) and dg (R) are written (not accumulated). If g is NULL: dv = permuted dw. */
int srnn_weight_prep_bwd(const float* dw, const int64_t* s, const float* v, const float* g, const float* inv_norm,
                         int32_t R, int32_t A, int32_t B, float* dv, float* dg, srnn_stream_t stream);
/* fp32 (rows, cols) with leading dim ld_in -> bf16 (rows, cols_pad) zero padded, leading dim ld_out */
int srnn_pad_cast_bf16(const float* in, int64_t rows, int32_t cols, int64_t ld_in, void* out_bf16, int32_t cols_pad,
                       int64_t ld_out, srnn_stream_t stream);
/* fp32 (rows, cols) -> two bf16 matrices with in = hi + lo up to 2^-17 relative (hi = bf16(in), lo = bf16(in - hi)),
 * both (rows, cols_pad) zero padded with leading dim ld_out.  Used where an fp32 intermediate feeds another tensor-core
 * contraction and a single bf16 rounding would be the dominant error (the folded embedding-table gradients). */
int srnn_split_bf16(const float* in, int64_t rows, int32_t cols, int64_t ld_in, void* hi_bf16, void* lo_bf16,
                    int32_t cols_pad, int64_t ld_out, srnn_stream_t stream);
/* bf16 (rows, cols) -> fp32, out[r, c] (+)= in[r, c]  (accumulate != 0 adds) */
int srnn_bf16_to_f32(const void* in_bf16, int64_t rows, int32_t cols, int64_t ld_in, float* out, int64_t ld_out,
                     int32_t accumulate, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Conditioning mixer input (model.py:60-72): row (b,l) = [speaker_emb[spk[b]] (S) | utt[b,l] (U) | 0 pad]
 * ------------------------------------------------------------------------------------------- */
int srnn_mixer_input(const float* utt, const float* spk_table, const int32_t* spk_ids, int32_t batch, int32_t frames,
                     int32_t U, int32_t S, void* out_bf16, int32_t k_pad, srnn_stream_t stream);
/* backward of the speaker part: d_table[spk[b], s] += sum_l d_in[(b,l), s]  (d_in bf16, ld k_pad) */
int srnn_mixer_input_bwd(const void* d_in_bf16, const int32_t* spk_ids, int32_t batch, int32_t frames, int32_t S,
                         int32_t k_pad, float* d_table, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Frame tier input assembly (model.py:142-147,268-271): row (b,t) =
 *   [ lut[xq[b, x_off + t*fs + i]] for i<fs | conds[b, t / rep, :C] | 0 pad ]   as bf16.
 * conds is fp32 (B, L, C); rep = T / L.  If `frames` (fp32 (B,T,fs), already dequantised - the
 * FrameLevelLayer.forward calling convention, model.py:140) is non-null it replaces the xq/lut path.
 * ------------------------------------------------------------------------------------------- */
int srnn_tier_input(const uint8_t* xq, int64_t xq_ld, int32_t x_off, const float* lut, const float* frames,
                    const float* conds, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C, void* out_bf16,
                    int32_t k_pad, srnn_stream_t stream);
/* dconds[b, l, c] += sum_{t in frame l} d_in[(b,t), fs + c]   (d_in bf16 with ld k_pad) */
int srnn_tier_input_bwd(const void* d_in_bf16, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C,
                        int32_t k_pad, float* dconds, srnn_stream_t stream);

/* Row repeat (model.py:189-191): out[(b, l*rep + i), :] = in[(b,l), :] for i < rep, bf16, and its
 * adjoint (sum over the rep rows, fp32 accumulate into out). */
int srnn_repeat_rows(const void* in_bf16, int64_t rows, int32_t cols, int64_t ld_in, int32_t rep, void* out_bf16,
                     int64_t ld_out, srnn_stream_t stream);
int srnn_repeat_rows_bwd(const void* d_out_bf16, int64_t rows, int32_t cols, int64_t ld_dout, int32_t rep,
                         void* d_in_bf16, int64_t ld_din, srnn_stream_t stream);

/* column sums of a bf16 matrix: out[c] = sum_r in[r, c]  (bias gradients); out is overwritten */
int srnn_colsum(const void* in_bf16, int64_t rows, int32_t cols, int64_t ld, float* out, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM - replaces the cuDNN/cuBLAS calls behind model.py:48,108-109,112,146-147,153-155,
 * 168-172,193-202 and their autograd backward.
 * ------------------------------------------------------------------------------------------- */
typedef struct srnn_gemm_args {
  /* op = 0 ("NT"): for each batch i < batch:  C_i[m,n] = epi( A_i[m,k] . B[n,k]^T )
   *                A_i = a + i*a_batch_stride (lda), C_i = c + i*c_batch_stride (ldc), B shared.
   * op = 1 ("TN"): C[m,n] += sum_i A_i[k,m]^T . B_i[k,n]   (fp32 atomic accumulate, split-K)
   *                A_i = a + i*a_batch_stride + a_row_offset*lda, B_i likewise. */
  int32_t op;
  int32_t m, n, k, batch;
  const void* a; int64_t lda; int64_t a_batch_stride; int32_t a_row_offset;
  const void* b; int64_t ldb; int64_t b_batch_stride; int32_t b_row_offset;
  void* c;       int64_t ldc; int64_t c_batch_stride;
  int32_t c_dtype;            /* 0 = bf16, 1 = fp32 (op 1 is always fp32) */
  int32_t n_fold;             /* op 0: if >0, element (row j, col n) is stored at row j*(N/n_fold) + n/n_fold,
                                 col n % n_fold  (learned-upsampling layout, model.py:153-155) */
  const float* bias;          /* [n] or NULL, added before the activation */
  const void* aux; int64_t ldaux; int64_t aux_batch_stride;   /* bf16 [m,n] per batch or NULL */
  int32_t aux_mode;           /* 0 none, 1 add (before activation), 2 gate: C = acc * (aux > 0) */
  int32_t relu;               /* apply max(.,0) last */
  int32_t aux_row_div;        /* 0/1: aux row = output row; d > 1: aux row = row / d (a term that is constant
                                 over groups of d rows, e.g. the conditioning of the d samples of one frame) */
  int32_t max_ctas;           /* 0: one persistent CTA per SM; n > 0: at most n CTAs (leaves SMs to a kernel that
                                 runs concurrently on another stream, e.g. the persistent recurrence) */
  float* colsum;              /* NT only, or NULL: colsum[j] += sum over all rows and batches of the fp32 epilogue
                                 result C[., j] (a bias gradient without a second pass over C; model.py:201 etc.
                                 backward); the caller zeroes it */
  /* NT only, optional second A operand: A_i = [ A_i[m, k1] | A2_i[m, k - k1] ] concatenated along K without being
   * materialised (comb_layer's input [one-hot windows | upper conditioning], model.py:196-199).  k1 % 64 == 0. */
  const void* a2; int64_t lda2; int64_t a2_batch_stride; int32_t k1;
  /* NT only, ReLU as a bit mask: with relu != 0 and relu_mask non-null the epilogue also writes bit (row, col) = result > 0
   * (one uint32 per row and 32 columns, row = batch * m + row in batch, ldmask words per row); a later GEMM given the same
   * words as gate_mask multiplies its result by that bit - the ReLU gradient (model.py:195,201 backward) from 4 bytes
   * per row and 32 columns instead of re-reading the 64 bytes of the saved activation (aux_mode 2). */
  uint32_t* relu_mask; const uint32_t* gate_mask; int64_t ldmask;
} srnn_gemm_args;

int srnn_gemm_bf16(const srnn_gemm_args* args, srnn_stream_t stream);

/* Last linear layer fused with log-softmax + NLL (model.py:202-203, runner.py:52): logits =
 * A[m,k] . W[256,k]^T + bias never leave TMEM.
 *   mode 0: lse[m], logp_target[m] (= logit[target] - lse) written.
 *   mode 1: additionally the full log-probabilities logp[m,256] (fp32) are written.
 *   mode 2: backward of mode 0: dlogits[m,n] = row_grad[m] * (onehot(target)[n] - softmax[n])  (bf16)
 *   mode 3: backward of mode 1: dlogits[m,n] = g[m,n] - softmax[n] * sum_n g[m,n]              (bf16)
 */
typedef struct srnn_nll_args {
  int32_t mode;
  int32_t m, k;                 /* n is fixed at 256 */
  const void* a; int64_t lda;   /* bf16 [m,k] */
  const void* w; int64_t ldw;   /* bf16 [256,k] */
  const float* bias;            /* [256] */
  const uint8_t* target;        /* [m] */
  float* lse; float* logp_target;        /* [m]: written by modes 0,1.  Modes 2,3: if lse is non-null it must hold the
                                            forward's values and the backward makes one pass over the row instead of
                                            three (max, sum of exp, gradient) */
  float* logp; int64_t ldlogp;           /* mode 1 */
  const float* row_grad;                 /* [m] mode 2 */
  const float* g; int64_t ldg;           /* [m,256] mode 3 */
  void* dlogits; int64_t lddlogits;      /* bf16 [m,256] modes 2,3 */
} srnn_nll_args;

int srnn_gemm_nll(const srnn_nll_args* args, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Persistent recurrent kernels - replace torch.nn.GRU (model.py:110,152) and its backward.
 * One cooperative launch runs all `steps` timesteps; W_hh stays resident in shared memory.
 * ------------------------------------------------------------------------------------------- */
typedef struct srnn_gru_args {
  int32_t batch, steps, hidden;   /* batch <= 512 per launch: rows are processed in groups of 64 (one MMA tile) inside
                                     every timestep with one grid handshake per timestep; with more than one group the
                                     fp32 state of a row is kept in h_state / c_state (backward: dh0 / dc0) between
                                     timesteps instead of in registers; hidden % 8 == 0 */
  int32_t ext_batch;     /* rows per time slot of the TIME-major buffers (>= batch; a launch may cover a
                            sub-range of a larger batch: pass pointers offset to its first row) */
  const void* gi;        /* bf16 [batch*steps, 3H] = W_ih u_t + b_ih, batch-major: row (b,t) = b*steps + t */
  const void* w_hh;      /* fwd: bf16 [3H, H];  bwd: bf16 [H, 3H] (= W_hh^T) */
  const float* b_hh;     /* [3H] (fwd) */
  void* h_ext;           /* bf16 [steps+1, ext_batch, H] TIME-major exchange buffer: slot 0 holds h_init
                            (input), slot t+1 receives h_t.  One timestep is one dense block, which is what
                            every CTA re-reads through TMA each step. */
  void* hall;            /* fwd out (nullable): bf16 [batch*steps, H] batch-major copy of h_t (GEMM operand) */
  float* h_state;        /* fwd: fp32 [batch, H], in = h_init, out = h_T */
  void* gates;           /* bf16 [batch*steps, 4H] batch-major: r, z, n, (W_hn h + b_hn); fwd writes, bwd reads */
  /* backward only */
  const void* dh_out;    /* bf16 [batch*steps, H] batch-major: dL/dh_t from the layers above */
  void* dgi;             /* bf16 [batch*steps, 3H] batch-major, out */
  void* dgh;             /* bf16 [steps, ext_batch, 3H] TIME-major, out (also the per-step exchange buffer) */
  float* dh0;            /* fp32 [batch, H] out: dL/dh_init */
  uint32_t* sync;        /* >= 32 KB (8192 uint32), zeroed by the caller before every launch: [0] grid-wide arrival
                            counter, [64 + 32 j] release flag of CTA j, [4160 + 32 g] arrival counter of row group g; on return [32] holds the number of exchange
                            attempts the kernel rejected and repeated (see below) */
  int32_t tuning_flags;  /* 0 = defaults.  Every documented bit leaves the results unchanged:
                            2 = force a cooperative launch for steps == 1, 16 = strict exchange protocol (release
                            increment + acquire fence instead of relaxed increment + validated read), 32 = land the
                            per-step operand as one TMA box per K block instead of ONE box, 64 = two MMA-issuing warps, bits 12-13 = polls of the arrival
                            counter kept in flight (0: 1, 1: 2, 2: 4), bits 14-15 = their spacing ((n+1)*64 cycles),
                            bits 16-19 = n*32 cycles to hold the TMA read back after the wait, bits 20-23 = cycles before the first
                            poll of a wait (0: 256, n: (n-1)*128), 1 << 26 = forward multi-group launches (batch > 64) use ONE arrival counter per timestep instead of one
                            per 64-row group (default forward: per group - a group's handshake then overlaps the other groups' work;
                            default backward: per timestep, 1 << 28 = per group there too), 1 << 27 =
                            32-row groups for 32 < batch <= 64 (measured slower), 1 << 25 = speculative landing (the first attempt of a timestep
                            skips the counter and relies on the validation; measured no faster), 1 << 24 = the last arriver releases the others
                            through per-CTA flag lines (default: every CTA polls the arrival counter; measured faster), 128 = debug_ts receives the global
                            timer of every CTA at timestep 24, bits 8-11 = force a cluster size.  Bits 1 and 4 exist only
                            in instrumented (-DSRNN_DEBUG) builds and are rejected with SRNN_ERR_ARG otherwise. */
  uint64_t* debug_ts;    /* NULL, or [256][8] clock64 stamps of CTA 0's pipeline events (profiling aid) */
  /* LSTM extension (cell = 1; no reference counterpart, torch.nn.LSTM semantics, gates i,f,g,o): every
   * "3H" above becomes 4H, `gates` is [batch*steps, 5H] (i, f, g, o, c_t). */
  int32_t cell;          /* 0 = GRU (reference), 1 = LSTM */
  float* c_state;        /* LSTM fwd: fp32 [batch, H] cell state, in = c_init, out = c_T */
  const float* c_init;   /* LSTM bwd: the cell state the forward started from */
  float* dc0;            /* LSTM bwd out: dL/dc_init */
  int32_t units_per_cta; /* 0 or 8: H/8 CTAs (fastest step); 16: H/16 CTAs, leaving SMs free for kernels that run
                            concurrently on another stream (falls back to 8 if H % 32 != 0) */
  float* db_ih;          /* bwd, nullable: fp32 [3H] (4H LSTM); db_ih[j] += sum over rows and timesteps of dgi[., j]
                            (the gradient of b_ih) - accumulated in registers during the loop, no extra pass */
  float* db_hh;          /* bwd, nullable: same for dgh (the gradient of b_hh) */
} srnn_gru_args;

/* Exchange protocol.  Every timestep each CTA publishes its slice of h_t (backward: of the gate gradients) to the
 * time-major buffer and increments the arrival counter; every CTA then reads the whole matrix back through TMA.  By
 * default the increment is RELAXED (no gpu-scope fence between the data stores and the counter): the entry points first
 * fill the slots the kernel is going to write with bf16 NaN (0xFFFF) - forward: slots 1..steps of h_ext, backward: all
 * of dgh - and the kernel validates what it read (a missing element leaves NaN in every accumulator of its batch row)
 * and repeats the read if needed.  Results are bit-identical to the strict protocol (tuning flag 16: release increment
 * + acquire fence, no sentinel fill), which costs ~1 us more per timestep.  A state that genuinely contains NaN is
 * accepted after 256 repeats. */
int srnn_gru_forward(const srnn_gru_args* args, srnn_stream_t stream);
int srnn_gru_backward(const srnn_gru_args* args, srnn_stream_t stream);

/* Initial-state assembly (model.py:149-151, 239-243): h_init[b] = use_carry[b] ? carried[b] : h0;
 * writes fp32 h_state [B,H] and, if h_ext is non-null, the bf16 slot 0 of h_ext (row stride
 * ext_ld = (steps+1)*H). */
int srnn_state_select(const float* carried, const float* h0, const uint8_t* use_carry, int32_t batch, int32_t hidden,
                      float* h_state, void* h_ext_bf16, int64_t ext_ld, srnn_stream_t stream);
/* d_h0[j] = sum_{b: !use_carry[b]} dh_init[b, j] */
int srnn_state_select_bwd(const float* dh_init, const uint8_t* use_carry, int32_t batch, int32_t hidden, float* d_h0,
                          srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Loss reduction (runner.py:52 + model.py:283-284): out[0] = -sum_{valid rows} logp_target / count,
 * out[1] = count, where row m is valid iff row_valid[m / rows_per_slot] != 0.
 * ------------------------------------------------------------------------------------------- */
int srnn_masked_nll_mean(const float* logp_target, const uint8_t* slot_valid, int64_t rows, int32_t rows_per_slot,
                         float* out2, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * AdamClipped (optimizer.py:6-14): g = clamp(g,-1,1); Adam update, one fused pass over a flat
 * fp32 parameter buffer.  step is 1-based; grad_scale multiplies g before the clamp (data-parallel
 * averaging).
 * ------------------------------------------------------------------------------------------- */
int srnn_adam_clipped(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                      double beta1, double beta2, double eps, int32_t step, double grad_scale, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Autoregressive generation (SampleRNNModel.test, model.py:289-351), per-sample kernels.
 * ------------------------------------------------------------------------------------------- */
/* Process-wide switch: launch the per-sample kernels (srnn_embed_sum, the small-M path of srnn_gemm_bf16,
 * srnn_sample_categorical) with programmatic dependent launch, so that the launch and set-up of each overlaps the
 * tail of its predecessor in the stream (every such kernel waits for its prerequisites before touching global
 * memory).  Off by default. */
int srnn_set_pdl(int32_t on);
/* Embedding side of the sample-level layer for ONE time step (model.py:192-200): the embedding, the
 * conv1d over the last r0 samples and the embedding block of comb_layer are linear in the one-hot codes,
 * so with fixed weights they are r0 tables of q rows (table bf16 (r0*q, hidden), row k*q+code):
 *   out[b, :] = act( sum_{k<r0} table[k*q + idx[b*idx_ld + k], :] + pre[b*pre_ld + :] )
 * pre (bf16, nullable) carries the conditioning / upper-tier blocks and the bias; relu != 0 applies ReLU. */
int srnn_embed_sum(const void* table_bf16, const uint8_t* idx, int64_t idx_ld, int32_t batch, int32_t r0, int32_t q,
                   int32_t hidden, const void* pre_bf16, int64_t pre_ld, int32_t relu, void* out_bf16,
                   int64_t out_ld, srnn_stream_t stream);
/* log-softmax + the draw (model.py:203, 346-348).  in[b*ld + :q] holds log-probabilities, or raw logits when
 * normalise != 0 (then the log-softmax is taken here); if logp_out is non-null the log-probabilities are
 * written to logp_out[b*ld_out + :q].  pick[b] is sampled from them by inverse CDF with the uniform u[b] in
 * [0,1) (multinomial of the softmax).  When u is null and rng_state is not, the uniform is drawn ON THE DEVICE:
 * rng_state is a device array of 3 uint64 {seed, step, 0}; utterance b uses Philox4x32-10(seed; step, b) and the launch
 * advances `step` by one, so a captured CUDA graph of step programs draws fresh numbers at every replay (the reference
 * calls torch.multinomial per sample, model.py:346-348).  With both null, pick[b] is the arg-max.  If win is non-null, the utterance's
 * window of its last win_len codes (u8 (batch, win_len)) is shifted left by one and pick[b] appended; if out is
 * non-null, out[b*out_ld] = pick[b]. */
int srnn_sample_categorical(const float* in, int64_t ld, int32_t batch, int32_t q, int32_t normalise, float* logp_out,
                            int64_t ld_out, const float* u, uint64_t* rng_state, uint8_t* win, int32_t win_len,
                            uint8_t* out, int64_t out_ld, srnn_stream_t stream);

/* srnn_sample_categorical followed, in the same launch, by srnn_embed_sum for the NEXT sample step: after the draw has
 * been appended to the window, h1_next[b, :] = relu(sum_k table[k*q + win[b, win_len - r0 + k], :] + pre_next[b, :]).
 * Usable whenever the frame-constant term of the next step is already known (every step that is not followed by a
 * frame-tier step); saves one launch of the per-sample latency chain. */
int srnn_sample_embed(const float* in, int64_t ld, int32_t batch, int32_t q, int32_t normalise, float* logp_out,
                      int64_t ld_out, const float* u, uint64_t* rng_state, uint8_t* win, int32_t win_len, uint8_t* out,
                      int64_t out_ld, const void* table_bf16, int32_t r0, int32_t hidden, const void* pre_next_bf16,
                      int64_t pre_ld, void* h1_next_bf16, int64_t h1_ld, srnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * fp32-tolerance arithmetic mode (SampleRNNModel(precision='fp32')).  The reference computes in fp32 end to end
 * (model.py:146-155,192-203; no autocast); this mode reproduces its results to fp32-level tolerances (loss
 * rel <= 1e-5, per-tensor gradient rel-L2 <= 3e-3; measured 6e-8 / 9.3e-4 at H = 1024) while still contracting on the
 * tcgen05 GEMM: every activation and gradient is an fp32 tensor, and a product enters srnn_gemm_bf16 on SPLIT operands
 * concatenated along K - three products of two bf16 pieces (backward) or six products of three pieces (forward), ONE
 * bf16 GEMM with a 3x / 6x longer K.  The entries below are what the mode needs besides the GEMM.  It is ~5.7x slower
 * than the bf16 path.
 * ------------------------------------------------------------------------------------------- */
/* fp32 (rows, cols) -> bf16 (rows, n_seg * cols_pad), every segment zero padded to cols_pad columns.
 * Two pieces (x ~ hi + lo to 2^-18; three products): role 0: [lo | hi | hi] (first operand of a product), role 1:
 * [hi | lo | hi] (second operand), role 2: [lo | hi] (against an operand that is exact in bf16, e.g. one-hot rows).
 * Three pieces (x = p0 + p1 + p2 to 2^-24, i.e. all of fp32; the six products down to order 2^-16): role 3:
 * [p2 | p1 | p0 | p1 | p0 | p0] (first operand), role 4: [q0 | q1 | q2 | q0 | q1 | q0] (second operand).  The SMALLEST
 * products come first along K: the tensor core's fp32 accumulation truncates (~3e-8 relative to the accumulator per
 * K=16 update, measured), so with the leading product last the accumulator is small during all the other updates.
 * A TN product uses the three column segments of roles 0 / 1 in three accumulating calls. */
int srnn_split3_bf16(const float* in, int64_t rows, int32_t cols, int64_t ld_in, void* out_bf16, int32_t cols_pad,
                     int64_t ld_out, int32_t role, srnn_stream_t stream);
/* fp32-output variants of srnn_mixer_input / srnn_tier_input / srnn_weight_prep and fp32-input variants of their
 * adjoints (same argument meaning; leading dimensions in elements) */
int srnn_mixer_input_f32(const float* utt, const float* spk_table, const int32_t* spk_ids, int32_t batch, int32_t frames,
                         int32_t U, int32_t S, float* out, int32_t k_pad, srnn_stream_t stream);
int srnn_mixer_input_bwd_f32(const float* d_in, const int32_t* spk_ids, int32_t batch, int32_t frames, int32_t S,
                             int32_t k_pad, float* d_table, srnn_stream_t stream);
int srnn_tier_input_f32(const uint8_t* xq, int64_t xq_ld, int32_t x_off, const float* lut, const float* frames,
                        const float* conds, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C, float* out,
                        int32_t k_pad, srnn_stream_t stream);
int srnn_tier_input_bwd_f32(const float* d_in, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C, int64_t ld,
                            float* dconds, srnn_stream_t stream);
int srnn_weight_prep_f32(const float* v, const float* g, int32_t R, int32_t A, int32_t B, float* out1, const int64_t* s1,
                         float* out2, const int64_t* s2, float* inv_norm, srnn_stream_t stream);
/* out[(b, j), :] = sum_{k < r0} table[k*q + idx[b*idx_ld + j + k], :]: embedding + conv1d + the embedding block of
 * comb_layer (model.py:192-200) as a gather-sum over the folded fp32 table (r0*q, hidden) - no rounding beyond fp32 adds */
int srnn_embed_gather_f32(const float* table, const uint8_t* idx, int64_t idx_ld, int32_t batch, int32_t rows_per_slot,
                          int32_t r0, int32_t q, int32_t hidden, float* out, srnn_stream_t stream);
/* in place: x[r, c] = act(x[r, c] + aux[r / aux_row_div, c] + aux2[r, c]) (aux, aux2 nullable; relu != 0: max(., 0));
 * mask (nullable) receives bit (r, c) = result > 0 in the word layout of srnn_gemm_args.relu_mask */
int srnn_bias_act_f32(float* x, int64_t rows, int32_t cols, int64_t ld, const float* aux, int64_t ldaux,
                      int32_t aux_row_div, const float* aux2, int64_t ldaux2, int32_t relu, uint32_t* mask,
                      int64_t ldmask, srnn_stream_t stream);
/* d_in[r, :] = sum_{i < rep} d_out[r * rep + i, :]  (adjoint of the row repeat, model.py:189-191) */
int srnn_segment_sum_f32(const float* d_out, int64_t rows, int32_t cols, int64_t ld_dout, int32_t rep, float* d_in,
                         int64_t ld_din, srnn_stream_t stream);
int srnn_colsum_f32(const float* in, int64_t rows, int32_t cols, int64_t ld, float* out, srnn_stream_t stream);
/* log_softmax (model.py:203) in place over rows of q <= 256 fp32 logits + the target pick of nll_loss (runner.py:52):
 * lse[m] and logp_target[m] (nullable) are written */
int srnn_logsoftmax_nll_f32(float* logits_inout, int64_t ld, int64_t m, int32_t q, const uint8_t* target, float* lse,
                            float* logp_target, srnn_stream_t stream);
/* its backward from the saved log-probabilities: g == NULL: dlogits = row_grad[m] * (onehot(target) - softmax);
 * else dlogits = g - softmax * sum_n g */
int srnn_logsoftmax_nll_bwd_f32(const float* logp, int64_t ld, int64_t m, int32_t q, const uint8_t* target,
                                const float* row_grad, const float* g, int64_t ldg, float* dlogits, int64_t lddlogits,
                                srnn_stream_t stream);
/* torch.nn.GRU (model.py:110,152) in the fp32 mode: a host loop over the timesteps, each one split-operand GEMM
 * (h_{t-1} . W_hh^T, resp. dgh_t . W_hh) and one fp32 cell kernel; batch-major fp32 buffers throughout. */
typedef struct srnn_gru_f32_args {
  int32_t batch, steps, hidden;  /* hidden % 8 == 0 */
  const float* gi;       /* fwd: [batch*steps, 3H] = W_ih x_t + b_ih, row (b,t) = b*steps + t */
  const void* w3;        /* fwd: srnn_split3_bf16 role 4 of W_hh [3H, H] -> bf16 [3H, 6H]; bwd: role 1 of W_hh^T [H, 3H]
                            -> bf16 [H, 9H] */
  const float* b_hh;     /* fwd: [3H] */
  float* h_state;        /* fwd: [batch, H] in = h_init, out = h_T */
  float* hall;           /* [batch*steps, H]: fwd out, bwd in */
  const float* h_init;   /* bwd: [batch, H] the state the forward started from */
  float* gates;          /* [batch*steps, 4H]: r, z, n, W_hn h + b_hn; fwd out, bwd in */
  void* a3;              /* workspace, bf16: fwd [batch, 6H], bwd [batch, 9H] */
  float* ws;             /* workspace: fwd [batch, 3H], bwd [batch, H] */
  const float* dh_out;   /* bwd: [batch*steps, H] */
  float* dgi;            /* bwd out: [batch*steps, 3H] */
  float* dgh;            /* bwd out: [batch*steps, 3H] (batch-major, unlike the bf16 kernels) */
  float* dh0;            /* bwd out: [batch, H] */
  float* carry;          /* bwd workspace: [batch, H] */
} srnn_gru_f32_args;
int srnn_gru_forward_f32(const srnn_gru_f32_args* args, srnn_stream_t stream);
int srnn_gru_backward_f32(const srnn_gru_f32_args* args, srnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SRNN_B200_H */
