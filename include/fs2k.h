/* libfs2k — C ABI of the B200 (sm_100a) kernels behind the FastSpeech2 acoustic-model hot path.
 *
 * The reference (EveryVoiceTTS/FastSpeech2_lightning) has no FFI of its own: the seam is plain
 * Python `nn.Module` classes.  Each entry point below replaces the PyTorch library
 * calls at the cited reference lines; the Python mirror package `fastspeech2_lightning_b200.fs2`
 * binds them with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - activations are channels-last fp32 `[B, L, C]`, contiguous; lengths/durations int32;
 *     bucket ids int64 (to match torch.bucketize); boolean masks uint8
 *   - no allocation, no synchronisation, no host<->device copy inside; work is enqueued on `stream`
 *   - returns FS2K_OK (0) or a negative FS2K_ERR_* code; never aborts; fs2k_strerror() explains
 *   - `[B,L]` pairs describe the row space: row m = b·L + l; convolution taps never cross b
 */
#ifndef FS2K_H_
#define FS2K_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* fs2k_stream_t; /* a cudaStream_t */

enum {
    FS2K_OK = 0,
    FS2K_ERR_BAD_SHAPE = -1,
    FS2K_ERR_UNSUPPORTED = -2,
    FS2K_ERR_WORKSPACE = -3,
    FS2K_ERR_NULL = -4,
    FS2K_ERR_CUDA = -5,
    FS2K_ERR_ARCH = -6
};

enum { FS2K_ACT_NONE = 0, FS2K_ACT_RELU = 1, FS2K_ACT_SILU = 2, FS2K_ACT_TANH = 3 };

const char* fs2k_strerror(int code);
int fs2k_version(void);
int fs2k_check_device(void); /* FS2K_OK iff the current device is compute capability 10.x */
/* Programmatic dependent launch for every kernel of the library (default on): the next kernel of a stream / graph is
 * scheduled while the previous one drains and waits in griddepcontrol.wait before touching memory.  0 = plain launches. */
int fs2k_set_pdl(int enabled);

/* ---- monotonic alignment search -------------------------------------------------------------
 * replaces VarianceAdaptor.binarize_attention (fs2/variance_adaptor.py:160-181) and
 * mas_width1 / b_mas (fs2/attn/alignment.py:48-85).
 * attn [B,F,T]: log-probabilities, or probabilities when take_log != 0 (the `torch.log` of :168 fused).
 * out: path [B,F] int32 (text index of every mel frame, -1 on padded frames), durations [B,T] int32
 * (= attn_hard.sum(2), :267-268), hard [B,1,F,T] fp32 0/1 (may be NULL).  T <= 4096. */
/* T <= 1024 runs the skewed-block kernel (warps work one block of 8 frames apart, so the block barrier is met once per 8 frames
 * instead of every frame); fs2k_mas_set_wavefront(0) selects the one-barrier-per-frame kernel (same results). */
int fs2k_mas_set_wavefront(int enabled);
size_t fs2k_mas_workspace_bytes(int B, int F, int T);
int fs2k_mas_fwd(const float* attn, int take_log, const int* in_lens, const int* out_lens, int B, int F, int T,
                 int* path, int* durations, float* hard, void* workspace, size_t workspace_bytes,
                 fs2k_stream_t stream);

/* ---- LengthRegulator (fs2/variance_adaptor.py:65-81) ------------------------------------------
 * scan:   cum[b,t] = inclusive cumsum of max(dur,0); total[b] = cum[b,T-1]
 * gather: out[b,f,:] = f < total[b] ? x[b, #{t: cum[b,t] <= f}, :] : 0 ; mask[b,f] = f < total[b];
 *         out_pos (optional) = out + PositionalEmbedding(f)·mask  (fs2/model.py:233-241 fused);
 *         idx_out (optional) = the gather index, -1 on padding.  F_out is the caller's
 *         min(max_b total, max_length). */
int fs2k_lr_scan(const int* durations, int B, int T, int* cum, int* total, fs2k_stream_t stream);
int fs2k_lr_gather(const float* x, const int* cum, const int* total, int B, int T, int D, int F_out, float* out,
                   float* out_pos, const float* inv_freq, uint8_t* mask, int* idx_out, fs2k_stream_t stream);

/* ---- variance embedding (fs2/variance_adaptor.py:183-205,:322,:343) ---------------------------
 * ids = torch.bucketize(v·scale, bins) (right=False; NaN -> n_bins); y = x + table[ids].
 * v_scaled (optional) receives v·scale (the `prediction * control` of :202). y may alias x. */
int fs2k_bucketize_embed_add(const float* v, float scale, float* v_scaled, const float* bins, int n_bins,
                             const float* table, const float* x, float* y, long long* ids, long N, int D,
                             fs2k_stream_t stream);
int fs2k_bucketize(const float* v, const float* bins, int n_bins, long long* ids, long N, fs2k_stream_t stream);

/* average_variance (fs2/variance_adaptor.py:207-222): per phone mean of the non-zero frame values
 * inside its duration span; cum is the inclusive cumsum from fs2k_lr_scan. */
int fs2k_average_variance(const float* var, const int* cum, int B, int F, int T, float* out, fs2k_stream_t stream);

/* inference duration rounding (fs2/variance_adaptor.py:359-366):
 * dur = int(clamp(rint(exp(log_dur) - 1) * control, min=0)) */
int fs2k_round_durations(const float* log_dur, float control, long N, int* dur, fs2k_stream_t stream);

/* ---- embeddings (fs2/model.py:183-213, fs2/layers.py:132-140) ---------------------------------
 * emb = table[text] (optional output: the aligner's keys); x = emb + PositionalEmbedding(t)·[t < lens[b]].
 * err_flag (optional) is set to 1 when an id is outside [0, n_sym). */
int fs2k_embed_posenc(const int* text, const float* table, int n_sym, const float* inv_freq, const int* lens, int B,
                      int T, int D, float* emb, float* x, int* err_flag, fs2k_stream_t stream);
int fs2k_add_posenc(const float* x_in, const float* inv_freq, const int* lens, int B, int L, int D, float* x,
                    fs2k_stream_t stream);
/* y[b,l,:] = x[b,l,:] + sum_k rows_k[ids_k ? ids_k[b] : b, :]  (GST / speaker / language rows) */
int fs2k_add_rows(const float* x, float* y, int B, int L, int D, const float* rows0, const int* ids0,
                  const float* rows1, const int* ids1, const float* rows2, const int* ids2, fs2k_stream_t stream);
/* mask[b,l] = l < lens[b]  (mask_from_lens, fs2/utils/heavy.py:11-15) */
int fs2k_lens_mask(const int* lens, int B, int L, uint8_t* mask, fs2k_stream_t stream);
/* lens[b] = count_nonzero(mask[b,:])  (fs2/model.py:226-230) */
int fs2k_mask_lens(const uint8_t* mask, int B, int L, int* lens, fs2k_stream_t stream);

/* ---- normalisation ------------------------------------------------------------------------------
 * LayerNorm over the last dim (eps 1e-5): torchaudio conformer.py:41,103,151,165; fs2/layers.py:42.
 * mean_out / rstd_out (optional, [M]) are saved for the backward pass. */
int fs2k_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, long M, int D, float dropout_p,
                       long seed, float* y, float* mean_out, float* rstd_out, fs2k_stream_t stream);
/* BatchNorm1d (conformer.py:62-64, fs2/layers.py:168-202): colstats = per-channel sum / sum-of-squares of
 * z[M,C] in fp64 (sums[2C], zeroed here); bn_finalize turns them (training) or the running stats (eval)
 * into scale/shift and, in training, updates running_mean/var (momentum, unbiased var) in place. */
int fs2k_colstats(const float* z, long M, int C, double* sums, fs2k_stream_t stream);
int fs2k_bn_finalize(const double* sums, long M, int C, const float* gamma, const float* beta, float eps,
                     float momentum, int training, float* running_mean, float* running_var,
                     long long* num_batches_tracked, float* scale, float* shift, float* save_mean, float* save_rstd,
                     fs2k_stream_t stream);
/* y = dropout(act(z*scale[c] + shift[c])) (+ residual); scale == NULL: plain activation; dropout_p == 0: none */
int fs2k_affine_act(const float* z, const float* scale, const float* shift, int act, const float* residual, long M,
                    int C, float dropout_p, long seed, float* y, fs2k_stream_t stream);

/* ---- dense contractions ---------------------------------------------------------------------------
 * C[(b,l),n] = (act((sum_tap sum_k A[b,l+tap-pad,k] W[tap][n][k] + bias[n]) * scale[n] + shift[n]) * alpha
 *               + residual[(b,l),n]) * row_mask[(b,l)]
 * = nn.Linear / Conv1d(k) on channels-last data with the following elementwise ops fused.
 * W is [taps][N][K] (fs2k_repack_conv_weight converts PyTorch's [N][K][taps]).
 * fs2k_gemm_f32: exact fp32 FFMA path.  fs2k_gemm_tc (gemm_tc.cu): tcgen05 tensor-core path. */
int fs2k_gemm_f32(const float* A, int lda, int B, int L, int K, const float* W, int N, int taps, int pad,
                  const float* bias, const float* scale, const float* shift, int act, float alpha,
                  const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc, fs2k_stream_t stream);
/* Tensor-core path (tcgen05.mma kind::tf32, TMEM accumulators, TMA operand boxes; gemm_tc.cu): same
 * contract as fs2k_gemm_f32; passes = 1 (single TF32) or 3 (3xTF32 split accumulation, fp32-level accuracy).
 * Optional fused LayerNorm(s) of the finished row when one tile spans the row (N <= 256):
 * ln_out = LN(C; ln_gamma, ln_beta), ln2_out = LN(ln_out; ln2_gamma, ln2_beta); C may be NULL if only
 * the normalised row is wanted.  dropout_p > 0: the value is dropped (counter hash of (seed, m*N + n), scaled by
 * 1/(1-p)) before the residual is added — `residual + Dropout(Linear(x))*alpha` in one launch.
 * W_small (optional, passes = 3): x - tf32(x) of every weight, same layout as W (fs2k_split_small) — the kernel then loads
 * it with TMA instead of splitting the weight tile in shared memory at every k-block.
 * fs2k_gemm_tc_supported: K % 4 == 0, lda % 4 == 0, N % 16 == 0 (N <= 256)
 * or N % 128 == 0. */
int fs2k_gemm_tc_supported(int K, int N, int lda, int taps);
int fs2k_gemm_tc(const float* A, int lda, int B, int L, int K, const float* W, int N, int taps, int pad,
                 const float* bias, const float* scale, const float* shift, int act, float alpha,
                 const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc, const float* ln_gamma,
                 const float* ln_beta, float ln_eps, float* ln_out, const float* ln2_gamma, const float* ln2_beta,
                 float* ln2_out, float dropout_p, long seed, int passes, const float* W_small, fs2k_stream_t stream);
/* bf16 arithmetic mode (BASELINE configs[2]; gemm_bf16.cu): tcgen05.mma kind::f16 — operands rounded to bf16,
 * fp32 accumulation in TMEM, fp32 epilogue.  Same contraction contract as fs2k_gemm_f32 with
 *   A: fp32 (a_is_bf16 = 0; rounded in the kernel, no cast pass) or bf16 (a_is_bf16 = 1; TMA), row stride lda elements
 *   W_hi: bf16 weights; W_lo (optional, fp32 A only): bf16(W - W_hi) — the kernel then also splits A and accumulates
 *         hi*hi + hi*lo + lo*hi (fp32-level accuracy)
 *   w_mn = 0: W is [taps][N][K];  w_mn = 1: W is [taps][K][N] read MN-major with the taps visited in reverse — the
 *         data-gradient GEMM dX = G * W of a layer whose forward weights are that same array (no transposed copy)
 *   outputs (each optional, at least one): C fp32 / C16 bf16 = the final value; P32 / P16 (stride ldp) = acc + bias,
 *         the pre-activation saved for the backward.
 * block_n_hint: 0 = automatic; 256 = wide tile for compute-bound shapes; < 0 = never the row-panel kernel.
 * fp32 A with K <= 256 (K % 64 == 0), one tap, N % 128 == 0 runs the row-panel kernel (gemm_bf16_panel.cu): A panel
 * read once and resident in shared memory, two TMEM accumulators, epilogue overlapped with the next tile's MMAs.
 * fs2k_gemm_bf16_supported: K % 8 == 0 (w_mn = 0) / N % 8 == 0 (w_mn = 1), lda % 4 == 0 (fp32 A) or % 8 (bf16 A),
 * N % 16 == 0 (N <= 256) or N % 128 == 0. */
int fs2k_gemm_bf16_supported(int K, int N, int lda, int taps, int a_is_bf16, int w_mn);
int fs2k_gemm_bf16(const void* A, int a_is_bf16, int lda, int B, int L, int K, const void* W_hi, const void* W_lo,
                   int w_mn, int N, int taps, int pad, const float* bias, const float* scale, const float* shift,
                   int act, float alpha, const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc,
                   void* C16, int ldc16, float* P32, void* P16, int ldp, float dropout_p, long seed, int block_n_hint,
                   fs2k_stream_t stream);
/* Masked multi-head self-attention on tcgen05 tensor cores, bf16 mode (attention_tc.cu; replaces
 * nn.MultiheadAttention at torchaudio conformer.py:151-153,193-202): qkv_bf16 [B,L,3*H*128] bf16 (the in-projection
 * GEMM's C16 output), keys >= lens[b] masked, fp32 softmax statistics, out [B,L,H*128] fp32, lse [B,H,L] (= m + ln l of
 * the scaled scores, saved for the backward).  head_dim must be 128.  Attention-probability dropout by counter hash.
 * Backward: delta [B,H,L] is workspace (rowsum(dO*O)), dout / dout_bf16 the same gradient in fp32 and bf16,
 * dqkv [B,L,3*H*128] fp32.  Deterministic (no atomics). */
int fs2k_attention_bf16(const void* qkv_bf16, const int* lens, int B, int L, int H, int head_dim, float dropout_p,
                        long seed, float* out, float* lse_out, const int* order, fs2k_stream_t stream);
int fs2k_attention_bwd_bf16(const void* qkv_bf16, const float* out, const float* lse, const float* dout,
                            const void* dout_bf16, const int* lens, int B, int L, int H, int head_dim, float dropout_p,
                            long seed, float* delta, float* dqkv, const int* order, fs2k_stream_t stream);
/* Split-K reduction of the single-tap weight gradients (both wgrad entry points).  1 (default): partial tiles are added into
   dW by 16-byte reductions resolved in L2 — no workspace traffic, no reduce launch, summation order not fixed; 0: workspace +
   deterministic reduce kernel.  Multi-tap (conv) gradients always use the workspace. */
int fs2k_wgrad_set_atomic(int enabled);
/* fs2k_gemm_wgrad_bf16 (below) with either operand already bf16 in HBM (read by TMA instead of the converting producer warps);
   ld % 8 == 0 for those.  Replaces the weight-gradient half of autograd through nn.Linear / nn.Conv1d in the Conformer stacks
   and the PostNet (torchaudio conformer.py:103-108,151-153; fs2/layers.py:143-212) in the bf16 mode. */
int fs2k_gemm_wgrad_bf16_ex(const void* G, int g_is_bf16, int ldg, const void* X, int x_is_bf16, int ldx, int B, int L,
                            int N, int K, int taps, int pad, void* workspace, size_t workspace_bytes,
                            float* dW_param_layout, int accumulate, fs2k_stream_t stream);
/* LayerNorm backward with the residual branch's gradient added in the same launch: dx = LN_bwd(g) + g_add */
int fs2k_layernorm_bwd_add(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma, long M,
                           int D, const float* g_add, float* dx, float* dgamma, float* dbeta, int accumulate,
                           fs2k_stream_t stream);
/* Data-gradient GEMM of the bf16 mode fused with the derivative of the activation (and dropout mask) that followed the forward
 * layer: out = (G * W) .* act'(pre) .* keep * alpha, pre_bf16 [M,N] = the forward GEMM's saved pre-activation.  fp32 G [M,K],
 * K <= 256, K % 64 == 0, N % 128 == 0; w_mn as in fs2k_gemm_bf16; C (fp32) and / or C16 (bf16) outputs, row stride N. */
int fs2k_gemm_bf16_dact(const float* G, int ldg, long M, int K, const void* W_hi, int w_mn, int N, const void* pre_bf16, int act,
                        float alpha, float dropout_p, long seed, float* C, void* C16, fs2k_stream_t stream);
/* fs2k_colsum (bias gradient) of a bf16 matrix */
int fs2k_colsum_bf16(const void* z_bf16, long M, int C, float* out, int accumulate, fs2k_stream_t stream);
/* bf16-output variants of fs2k_affine_act / fs2k_bn_act_bwd: the result only feeds tensor-core contractions of the bf16 mode
 * (its TMA reads bf16 tiles), so it is rounded once where it is produced and the fp32 copy is never written. */
int fs2k_affine_act_bf16(const float* z, const float* scale, const float* shift, int act, const float* residual, long M,
                         int C, float dropout_p, long seed, void* y_bf16, fs2k_stream_t stream);
int fs2k_bn_act_bwd_bf16(const float* g, const float* z, const float* scale, const float* shift, const float* mean,
                         const float* rstd, int act, int training, long M, int C, float dropout_p, long seed,
                         double* sums, void* gz_bf16, float* dgamma, float* dbeta, int accumulate, fs2k_stream_t stream);
/* hi[i] = bf16(x[i]); lo[i] = bf16(x[i] - hi[i]) when lo != NULL */
int fs2k_cast_bf16(const float* x, long n, void* hi, void* lo, fs2k_stream_t stream);
/* y = fp32(x): widens a bf16 buffer (gradient exchange in bf16: the reference trains under Lightning DDP, whose
   `bf16_compress_hook` does the same narrowing around the all-reduce). */
int fs2k_cast_f32(const void* x_bf16, long n, float* y, fs2k_stream_t stream);
/* Weight gradient in the bf16 mode (gemm_wgrad_bf16.cu): G, X fp32 in HBM, rounded to bf16 in the kernel, both read as
 * MN-major tcgen05 operands, fp32 accumulation; same outputs as fs2k_gemm_wgrad_tc.  N % 4 == 0, K % 16 == 0,
 * K <= 256 or K % 256 == 0. */
int fs2k_gemm_wgrad_bf16_supported(int N, int K, int ldg, int ldx);
size_t fs2k_gemm_wgrad_bf16_workspace_bytes(int B, int L, int N, int K, int taps);
int fs2k_gemm_wgrad_bf16(const float* G, int ldg, const float* X, int ldx, int B, int L, int N, int K, int taps,
                         int pad, void* workspace, size_t workspace_bytes, float* dW_param_layout, int accumulate,
                         fs2k_stream_t stream);
/* out[i] = x[i] - tf32_truncate(x[i]): the "small" operand of the 3xTF32 scheme */
int fs2k_split_small(const float* x, long n, float* out, fs2k_stream_t stream);
int fs2k_rowdot(const float* x, const float* w, const float* b, const uint8_t* mask, long M, int D, float* y,
                fs2k_stream_t stream);
int fs2k_repack_conv_weight(const float* w, int N, int K, int taps, float* out, fs2k_stream_t stream);

/* ---- attention (torchaudio conformer.py:151-153,193-202) ------------------------------------------
 * qkv [B,L,3·H·hd] packed in_proj output; out [B,L,H·hd]; keys >= lens[b] masked; lse_out optional [B,H,L];
 * dropout_p > 0 (training): dropout on the attention probabilities, mask = counter hash of (seed, b, h, q, k). */
/* order (optional): utterance indices, longest first (fs2k_attention_order fills it from lens) — the attention kernels
 * then dispatch the long utterances first so the ragged tail of the last wave is filled by short ones. */
int fs2k_attention_order(const int* lens, int B, int* order_ws, fs2k_stream_t stream);
int fs2k_attention_f32(const float* qkv, const int* lens, int B, int L, int H, int head_dim, float dropout_p, long seed,
                       float* out, float* lse_out, const int* order, fs2k_stream_t stream);

/* ---- depthwise conv (conformer.py:50-65 with GLU/BatchNorm/SiLU fused; fs2/blocks.py:8-15) ---------
 * x [B,L,ldx] (glu: value c, gate c+C); w [C][K]; scale/shift non-NULL: y = silu((conv+bias)*scale+shift). */
int fs2k_dwconv_fwd(const float* x, int ldx, int B, int L, int C, const float* w, int K, const float* bias, int glu,
                    const float* scale, const float* shift, float* y, fs2k_stream_t stream);

/* ---- aligner scores (fs2/attn/attention.py:238-251) -------------------------------------------------
 * q [B,F,C], k [B,T,C] projected queries/keys; prior [B,F,T] or NULL; logprob, soft [B,F,T]. */
int fs2k_aligner_fwd(const float* q, const float* k, const float* prior, const int* key_lens, int B, int F, int T,
                     int C, float* logprob, float* soft, fs2k_stream_t stream);

/* ---- losses (fs2/loss.py:44-106, fs2/attn/attention_loss.py:65-73) ----------------------------------
 * masked_loss: loss = weight · mean_{all M·C elements} f(pred·m − target·m), f = square (kind 0) or abs (1),
 * m = row_mask[M]; target_is_int_log1p: target is int32 and enters as log(target + 1) (duration loss).
 * scratch: one double (fwd).  bwd: dpred = gout · weight/N · f'(...) · m. */
int fs2k_masked_loss_fwd(const float* pred, const void* target, int target_is_int_log1p, const uint8_t* row_mask,
                         long M, int C, int kind, float weight, double* scratch, float* loss, fs2k_stream_t stream);
int fs2k_masked_loss_bwd(const float* pred, const void* target, int target_is_int_log1p, const uint8_t* row_mask,
                         long M, int C, int kind, float weight, const float* gout, float* dpred, fs2k_stream_t stream);
/* bin loss = −Σ_{hard==1} log(max(soft, eps)) / Σ hard ; sums[2] doubles are kept for the backward */
int fs2k_bin_loss_fwd(const float* hard, const float* soft, long N, float eps, double* sums, float* loss,
                      fs2k_stream_t stream);
int fs2k_bin_loss_bwd(const float* hard, const float* soft, long N, float eps, const double* sums, const float* gout,
                      float* dsoft, fs2k_stream_t stream);
/* AttentionCTCLoss / ForwardSumLoss (fs2/attn/attention_loss.py:22-62), fused: blank column (log-prob
 * blank_logprob) + key mask + log_softmax over the key_len+1 classes + CTC α recursion with targets 1..key_len,
 * input length query_len, reduction "mean" (nll/key_len averaged over B), zero_infinity.
 * attn_logprob [B,F,T] (the [B,1,F,T] aligner output); lse [B,F], log_alpha [B,F,2T+1] (fs2k_ctc_alpha_elems
 * elements) and nll [B] are fp64 (the recursion runs in fp64) and are kept for the backward; loss: one float.  The backward writes the gradient with
 * respect to attn_logprob itself (through the log_softmax and the padding), scaled by the device scalar gout. */
size_t fs2k_ctc_alpha_elems(int B, int F, int T);
int fs2k_ctc_forward_sum_fwd(const float* attn_logprob, const int* key_lens, const int* query_lens, int B, int F,
                             int T, float blank_logprob, double* lse, double* log_alpha, double* nll, float* loss,
                             fs2k_stream_t stream);
int fs2k_ctc_forward_sum_bwd(const float* attn_logprob, const double* lse, const double* log_alpha, const double* nll,
                             const int* key_lens, const int* query_lens, const float* gout, int B, int F, int T,
                             float blank_logprob, float* d_attn_logprob, fs2k_stream_t stream);

/* ---- backward kernels (training step: fs2/model.py:384-390 + Lightning's backward) ----------------------
 * Every forward entry point above has its gradient here; torch.autograd only sequences the calls.
 * act_bwd:  gz = g · row_mask · alpha · act'(·); mode 0 none, 1 relu (aux = output), 2 silu (aux = pre-activation),
 *           3 tanh (aux = output).           colsum: out[c] = Σ_m z[m,c] (bias gradients).
 * repack_weight_t: [taps][N][K] → [taps][K][N] with reversed taps (weights of the transposed conv, so dx is a
 *           forward fs2k_gemm_* call on gz).  unpack_conv_weight: [taps][N][K] → PyTorch [N][K][taps].
 * gemm_wgrad: dW[tap][n][k] = Σ_(b,l) G[b,l,n] · X[b,l+tap-pad,k]  (dW zeroed here, split over rows + atomics). */
int fs2k_act_bwd(const float* g, const float* aux, int mode, float alpha, const uint8_t* row_mask, long M, int C,
                 float dropout_p, long seed, float* gz, fs2k_stream_t stream);
int fs2k_colsum(const float* z, long M, int C, float* out, int accumulate, fs2k_stream_t stream);
int fs2k_repack_weight_t(const float* w, int N, int K, int taps, float* out, fs2k_stream_t stream);
int fs2k_unpack_conv_weight(const float* w, int N, int K, int taps, float* out, fs2k_stream_t stream);
int fs2k_gemm_wgrad(const float* G, int ldg, const float* X, int ldx, int B, int L, int N, int K, int taps, int pad,
                    float* dW, fs2k_stream_t stream);
/* Tensor-core weight gradient (gemm_wgrad_tc.cu): both operands MN-major for tcgen05.mma (the contraction runs over
 * rows), row range split over CTAs, partial tiles summed in a fixed order (deterministic); the result is written
 * in the PARAMETER layout: [N][K][taps] (Conv1d) — identical to [N][K] when taps == 1.  Needs N % 128 == 0 and
 * K % 32 == 0 (K % 256 == 0 when K >= 256); other shapes use fs2k_gemm_wgrad. */
int fs2k_gemm_wgrad_tc_supported(int N, int K, int ldg, int ldx);
size_t fs2k_gemm_wgrad_tc_workspace_bytes(int B, int L, int N, int K, int taps);
int fs2k_gemm_wgrad_tc(const float* G, int ldg, const float* X, int ldx, int B, int L, int N, int K, int taps, int pad,
                       int passes, void* workspace, size_t workspace_bytes, float* dW_param_layout, int accumulate,
                       fs2k_stream_t stream);
/* LayerNorm backward (dgamma/dbeta zeroed here unless `accumulate`, accumulated with atomics).  `accumulate` != 0 in this
 * and in bn_act_bwd / dwconv_bwd adds the parameter gradients onto the given buffers (e.g. the flat .grad) instead. */
int fs2k_layernorm_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                       long M, int D, float dropout_p, long seed, float* dx, float* dgamma, float* dbeta, int accumulate,
        fs2k_stream_t stream);
/* BatchNorm1d (+ activation) backward: y = act(z*scale + shift), zhat = (z - mean)*rstd.
 * training: gz = scale*(gu - mean(gu) - zhat*mean(gu*zhat)); eval: gz = gu*scale; dgamma = Σ gu*zhat, dbeta = Σ gu.
 * sums: 2C doubles of scratch. */
int fs2k_bn_act_bwd(const float* g, const float* z, const float* scale, const float* shift, const float* mean,
                    const float* rstd, int act, int training, long M, int C, float dropout_p, long seed, double* sums,
                    float* gz, float* dgamma, float* dbeta, int accumulate,
        fs2k_stream_t stream);
/* attention backward (flash style, recomputes P from lse); delta: [B,H,L] scratch; dqkv [B,L,3·H·hd] */
int fs2k_attention_bwd_f32(const float* qkv, const float* out, const float* lse, const float* dout, const int* lens,
                           int B, int L, int H, int head_dim, float dropout_p, long seed, float* delta, float* dqkv,
                           const int* order,
                           fs2k_stream_t stream);
/* depthwise conv backward (glu: x = (value, gate) and dx has the same 2C layout); dw/dbias zeroed here */
int fs2k_dwconv_bwd(const float* gz, const float* x, int ldx, int B, int L, int C, const float* w, int K, int glu,
                    float* dx, float* dw, float* dbias, int accumulate,
        fs2k_stream_t stream);
int fs2k_rowdot_bwd(const float* g, const float* x, const float* w, const uint8_t* mask, long M, int D, float* dx,
                    float* dw, float* db, fs2k_stream_t stream);
/* LengthRegulator backward: contiguous segment sums of the two output gradients (either may be NULL) */
int fs2k_lr_bwd(const float* g_out, const float* g_out_pos, const int* cum, int B, int T, int D, int F, float* dx,
                fs2k_stream_t stream);
/* dtable[ids[n],:] += g1[n,:] (+ g2[n,:]), rows with id == skip_id skipped (padding_idx); dtable NOT zeroed */
int fs2k_scatter_add_rows(const float* g1, const float* g2, const void* ids, int ids_are_int64, long N, int D,
                          long skip_id, float* dtable, fs2k_stream_t stream);
/* drows[ids ? ids[b] : b, :] += Σ_l g[b,l,:]  (speaker / language / style rows); drows NOT zeroed */
int fs2k_rows_sum_scatter(const float* g, const int* ids, int B, int L, int D, float* drows, fs2k_stream_t stream);
/* aligner backward: gradients w.r.t. the projected queries/keys from g_soft and/or g_logprob; dd [B,F,T] scratch */
int fs2k_aligner_bwd(const float* g_soft, const float* g_logprob, const float* soft, const float* logprob,
                     const float* prior, const int* key_lens, const float* q, const float* k, int B, int F, int T,
                     int C, float* dd, float* dq, float* dk, fs2k_stream_t stream);

/* ---- global style tokens, inference (fs2/gst/model.py:103-257) ------------------------------
 * conv2d_s2_bn_relu: one ReferenceEncoder block — Conv2d(3x3, stride 2, pad 1, no bias) + folded BatchNorm2d + ReLU on
 *   channels-last x [B,H,W,Ci]; w re-packed to [3][3][Ci][Co]; y [B,Ho,Wo,Co], or [B,Ho,Co,Wo] when cw_layout != 0
 *   (the feature order `hs.transpose(1,2).view(B,T,C*W)` of :195-197 feeds the GRU with).
 * gru_gate: torch.nn.GRU cell update from the two projections (both with their biases), gate order r,z,n (:198).
 * gst_token_attention: multi-head attention of one query per batch row over T <= 32 tokens (gst/attn.py:172-194). */
int fs2k_conv2d_s2_bn_relu(const float* x, const float* w_khwcico, const float* scale, const float* shift, int B, int H,
                           int W, int Ci, int Co, int cw_layout, float* y, fs2k_stream_t stream);
/* training through the reference encoder: scale == shift == NULL in conv2d_s2_bn_relu gives the raw (pre-BatchNorm) conv;
 * its data / weight gradients (dw zeroed here, [3][3][Ci][Co]); GRU cell backward (d_xproj / d_hproj [B,3U], dh_prev = dh*z;
 * the W_hh term is the caller's GEMM); token-attention backward (dk / dv zeroed here, summed over the batch). */
int fs2k_conv2d_s2_dgrad(const float* gz, const float* w_khwcico, int B, int H, int W, int Ci, int Co, float* dx,
                         fs2k_stream_t stream);
int fs2k_conv2d_s2_wgrad(const float* x, const float* gz, int B, int H, int W, int Ci, int Co, float* dw_khwcico,
                         fs2k_stream_t stream);
int fs2k_gru_gate_bwd(const float* xproj, long xproj_row_stride, const float* hproj, const float* h, const float* dh, int B,
                      int U, float* dxproj, long dxproj_row_stride, float* dhproj, float* dh_prev, fs2k_stream_t stream);
int fs2k_gst_token_attention_bwd(const float* q, const float* k, const float* v, const float* dout, int B, int T, int heads,
                                 int dk, float* dq, float* dk_out, float* dv_out, fs2k_stream_t stream);
int fs2k_gru_gate(const float* xproj, long xproj_row_stride, const float* hproj, const float* h, int B, int U, float* h_out,
                  fs2k_stream_t stream);
int fs2k_gst_token_attention(const float* q, const float* k, const float* v, int B, int T, int heads, int dk, float* out,
                             fs2k_stream_t stream);

/* ---- batch builder / prediction trimming (fs2/dataset.py:257-293, fs2/prediction_writing_callback.py:255-262; 8f rank 4)
 * unpack_ragged: item b = rows[b] x cols[b] 4-byte words at packed + word_offsets[b] (valid values only, packed back to
 *   back by the host into one staging buffer) -> out[B][Rmax][Cmax] words, zero padded (pad_sequence and the two-sided
 *   padding of the [F,T] attention prior, done on the device).
 * trim_transpose: out[out_offsets[b] + c*len_b + t] = mel[b,t,c] for t < len_b = min(lens[b], F): every utterance's valid
 *   frames as [n_mels, T_b], packed for one D2H copy. */
int fs2k_unpack_ragged(const void* packed_words, const long long* word_offsets, const int* rows, const int* cols, int B,
                       int Rmax, int Cmax, void* out_words, fs2k_stream_t stream);
int fs2k_trim_transpose(const float* mel, const int* lens, const long long* out_offsets, int B, int F, int C, float* out,
                        fs2k_stream_t stream);

/* ---- optimizer over one flat buffer (torch.optim.AdamW at fs2/model.py:530-537; clip 1.0 at fs2/cli/train.py:38) ---
 * sumsq: out[0] = Σ g² (fp64).  adamw_step: g' = g·grad_scale·min(1, max_norm/(‖g·grad_scale‖+1e-6)) when sumsq is given
 * (grad_scale = 1/world_size folds the data-parallel mean), then AdamW with decoupled weight decay and bias correction
 * for `step` (1-based).  p_bf16 (optional): bf16 shadow of the updated parameters, written by the same launch — the
 * weight operands of the bf16 arithmetic mode. */
int fs2k_sumsq(const float* g, long N, double* out, fs2k_stream_t stream);
int fs2k_adamw_step(float* p, const float* g, float* m, float* v, long N, float lr, float beta1, float beta2, float eps,
                    float weight_decay, long step, float max_norm, float grad_scale, const double* sumsq,
                    void* p_bf16, fs2k_stream_t stream);
/* CUDA-graph replay of a whole training step: per-step scalars live in device memory so that a captured launch
 * sees fresh values.  step_state = 4 floats {lr, 1−β1^t, sqrt(1−β2^t), unused}; set_step_state writes them and the
 * dropout seed base in one tiny launch (by-value kernel arguments, so no host staging buffer can race).
 * set_dropout_seed_base registers the device counter every dropout kernel adds to its seed (NULL = none). */
int fs2k_adamw_step_dev(float* p, const float* g, float* m, float* v, long N, const float* step_state, float beta1,
                        float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                        const double* sumsq, void* p_bf16, fs2k_stream_t stream);
int fs2k_set_step_state(float* step_state, unsigned long long* seed_base, float lr, float bias_correction1,
                        float bias_correction2_sqrt, long seed_base_value, fs2k_stream_t stream);
int fs2k_set_dropout_seed_base(const unsigned long long* device_counter);

/* ---- small elementwise helpers ------------------------------------------------------------------------ */
int fs2k_axpby(const float* a, float alpha, const float* b, float beta, long N, float* out, fs2k_stream_t stream);
int fs2k_gather_rows(const float* table, const long long* ids, long R, int D, float* out, fs2k_stream_t stream);
int fs2k_tanh(const float* x, long N, float* y, fs2k_stream_t stream);
/* dropout: y = x·keep/(1-p) (+ residual), keep from a counter hash of (seed, index); the backward calls it again on
 * the gradient.  The same mask can be fused into fs2k_affine_act / fs2k_layernorm_fwd (forward) and
 * fs2k_act_bwd / fs2k_bn_act_bwd / fs2k_layernorm_bwd (backward) through their (dropout_p, seed) arguments. */
int fs2k_dropout(const float* x, const float* residual, float p, long seed, long N, float* y, fs2k_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FS2K_H_ */
