/*
 * xmodal_b200.h -- C ABI of the B200-native hot path of bacon205/Multimodal_eeg_fmri
 * (paired EEG/fMRI cross-modal training step).
 *
 * The reference has no FFI layer: its hot path is `torch.nn` calls inside the module
 * `forward`s.  Each entry point below replaces the ATen/oneDNN/cuDNN call
 * the cited reference line makes; the Python modules in multimodal_eeg_fmri_b200/ (same class
 * names, forward signatures and state_dict keys as the reference) bind these with ctypes --
 * see INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (or int64 / int32 where stated), row-major;
 *     the caller owns and allocates every buffer, including workspaces; nothing here
 *     allocates, frees or synchronises; all work is enqueued on `stream` (a cudaStream_t
 *     passed as void*, e.g. torch.cuda.current_stream().cuda_stream).
 *   - return value: XM_OK (0) or a negative XM_ERR_* code; xm_strerror() names it.
 *   - "ld*" arguments are leading dimensions (row pitches) in ELEMENTS.  Tensors read through
 *     TMA need 16-byte aligned bases and row pitches that are multiples of 4 elements.
 *   - GEMM-shaped work runs on tcgen05 tensor cores in TF32 with fp32 accumulation;
 *     reductions, norms, activations and the spectral path are fp32 (fp64 for the final
 *     combination of batch statistics).
 */
#ifndef XMODAL_B200_H_
#define XMODAL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XM_ABI_VERSION 1

#define XM_OK 0
#define XM_ERR_INVALID -1     /* bad argument (shape, alignment, null pointer) */
#define XM_ERR_UNSUPPORTED -2 /* shape outside what the kernels implement */
#define XM_ERR_LAUNCH -3      /* CUDA launch / driver error; see xm_last_cuda_error() */
#define XM_ERR_NO_DRIVER -4   /* cuTensorMapEncodeTiled could not be resolved (no GPU driver) */

#define XM_ACT_NONE 0
#define XM_ACT_RELU 1
#define XM_ACT_GELU 2 /* exact erf GELU, as nn.GELU() */
#define XM_ACT_TANH 3
#define XM_ACT_SIGMOID 4

int xm_abi_version(void);
const char* xm_strerror(int code);
int xm_last_cuda_error(void);

/* ------------------------------------------------------------------ EEG preprocessing
 * No reference implementation exists: vocabulary from EEG_CODE/config.py:34-36,
 * PW layout from EEG_CODE/CrossModal_EEG_scr.ipynb cell 7:45-47.  Oracle: oracle/spectral.py. */

/* Window index generation for `n_rec` recordings of `n_samples` samples each:
 * n_win = (n_samples - win)/hop + 1 per recording; window g = r*n_win + w has
 * starts[g] = w*hop, rec_ids[g] = r, labels[g] = rec_labels[r], subjects[g] = rec_subjects[r].
 * All arrays int64 on the device.  rec_labels / rec_subjects may be NULL (outputs skipped). */
int xm_window_index_i64(int64_t n_rec, int64_t n_samples, int64_t win, int64_t hop, const int64_t* rec_labels,
                        const int64_t* rec_subjects, int64_t* starts, int64_t* rec_ids, int64_t* labels,
                        int64_t* subjects, void* stream);

/* Gather windows from rec (n_rec, C, n_samples):
 *   channels_last == 0: out (n_rec*n_win, C, ld_out >= win)   -- the reference's (C, T) sample layout
 *   channels_last != 0: out (n_rec*n_win, win, ld_out >= C)   -- the layout the conv GEMMs consume
 * round_tf32 == 1 rounds values to tf32 (removes the truncation bias of the first conv); round_tf32 == 2
 * (channels_last only, ld_out >= 2*C) writes the tf32 split [hi | lo] along the channel axis, the operand of a 3-pass
 * (fp32-accurate) first conv (xm_conv1d_fwd_stats_f32 with x_cols = 2*C).
 * With n_win == 1 (win == n_samples) this is the (B, C, T) -> (B, T, C) layout change. */
int xm_window_gather_f32(const float* rec, int64_t n_rec, int64_t C, int64_t n_samples, int64_t win, int64_t hop,
                         float* out, int64_t ld_out, int channels_last, int round_tf32, void* stream);

/* Fused window gather + taper + real FFT (nfft = power of two in [64, 2048], >= win, zero padded) +
 * one-sided PSD + band-power reduction.  power[g, c, b] = sum_{k in [band_bins[2b], band_bins[2b+1])}
 * |X_k|^2 * s_k / (fs * sum(taper^2)) * (fs / nfft),  s_k = 2 except k = 0 and k = nfft/2.
 * taper: (win) fp32 window coefficients (device); band_bins: (2*n_bands) int32 (device). */
int xm_bandpower_f32(const float* rec, int64_t n_rec, int64_t C, int64_t n_samples, int64_t win, int64_t hop,
                     int64_t nfft, float fs, const float* taper, float taper_sumsq, const int32_t* band_bins,
                     int n_bands, float* power, void* stream);

/* The same band powers as a tensor-core DFT (csrc/bandpower_dft.cu): only the bins inside the bands are computed,
 * X = x . [taper cos | taper sin] as a 3-pass tf32 product with the split samples in tensor memory.  Eligible when
 * xm_bandpower_dft_supported(...) != 0: total_bins = sum of the band widths <= 32, n_bands <= 8, win <= 1024,
 * n_samples % 4 == 0, hop % 4 == 0 (nfft: any integer >= win, not only powers of two).  total_bins is the caller's
 * (host) knowledge of band_bins; workspace: xm_bandpower_dft_workspace_floats(win) floats, 16-B aligned. */
int64_t xm_bandpower_dft_workspace_floats(int64_t win);
int xm_bandpower_dft_supported(int64_t C, int64_t n_samples, int64_t win, int64_t hop, int64_t nfft, int n_bands,
                               int total_bins);
int xm_bandpower_dft_f32(const float* rec, int64_t n_rec, int64_t C, int64_t n_samples, int64_t win, int64_t hop,
                         int64_t nfft, float fs, const float* taper, float taper_sumsq, const int32_t* band_bins,
                         int n_bands, int total_bins, float* workspace, float* power, void* stream);

/* normalize_modality (EEG_CODE/run_training_lite.py:48-51): out = (x - mean)/(std + eps) per item,
 * population std, over `item_len` contiguous elements. */
int xm_zscore_f32(const float* x, int64_t n_items, int64_t item_len, float eps, float* out, void* stream);

/* fMRI ROI aggregation (fMRI_CODE/fmri_utils.py:140-147, agg_method='both'): x (B, TR, ROI) ->
 * out (B, 2*ROI) = concat(mean over TR, population std over TR); NaN inputs count as 0. */
int xm_roi_meanstd_f32(const float* x, int64_t B, int64_t TR, int64_t ROI, float* out, void* stream);

/* Functional connectivity of the ROI series (csrc/connectivity.cu): out (B, ROI*ROI) = per-sample numpy.corrcoef of the
 * ROI columns over TR, flattened row-major, after nan_to_num -- the `connectivity` input of fMRIFusionNet
 * (fMRI_CODE/fmri_utils.py:90-103; the reference reads such matrices from CSV, :161-198; SURVEY.md section 8d defines
 * the synthetic connectivity input this way).  fp32, 1e-5.  A constant column yields NaN in its row and column.
 * split3 != 0: out is (3*B, ROI*ROI), the row-stacked tf32 split [hi; hi; lo] that xm_linear_fwd_stacked3_f32 and the
 * weight gradient of the connectivity projection consume (what xm_split3_f32(which 1, axis 0) would make of it).
 * xm_roi_corrcoef_supported: the (TR, ROI) series of one sample fits shared memory. */
int xm_roi_corrcoef_supported(int64_t TR, int64_t ROI);
int xm_roi_corrcoef_f32(const float* x, int64_t B, int64_t TR, int64_t ROI, float* out, int split3, void* stream);

/* ------------------------------------------------------------------ dense projections (nn.Linear)
 * fMRI_CODE/fmri_utils.py:27,31,45,49,66 ; bridge_utils.py:35,41,61,65 ; enhanced_models_v4.py:164 */

/* y (M,N) = act(x (M,K) @ w (N,K)^T + bias).  splits > 1 splits K across CTAs through
 * `workspace` (splits*M*N floats).  flags: XM_LINEAR_ROUND_TF32 rounds y to tf32; XM_LINEAR_FP32_ACCUM
 * limits the tensor core's own (truncating) accumulation to 256 contraction elements at a time and adds
 * those chunks in fp32 round-to-nearest -- for the 3-pass fp32-accurate projections over long K. */
#define XM_LINEAR_ROUND_TF32 1
#define XM_LINEAR_FP32_ACCUM 2
int xm_linear_fwd_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int64_t N, int64_t K,
                      int64_t ldx, int64_t ldw, int64_t ldy, int act, int flags, int splits, float* workspace,
                      void* stream);
/* The same product for operands that arrive as three ROW-stacked tf32 split blocks: x3 (3M, K) = [hi; hi; lo],
 * w3 (3N, K) = [hi; lo; hi] (xm_split3_f32 along axis 0) -> y = x w^T to fp32 accuracy.  The row-stacked split of
 * x is exactly the operand of the weight gradient (xm_linear_wgrad_f32 over 3M rows), so one split of a large
 * input (the 40 000-d connectivity features) serves both the forward and the backward. */
int xm_linear_fwd_stacked3_f32(const float* x3, const float* w3, const float* bias, float* y, int64_t M, int64_t N, int64_t K,
                               int64_t ldx, int64_t ldw, int64_t ldy, int act, int flags, int splits, float* workspace,
                               void* stream);
/* dx (M,K) = dy (M,N) @ w (N,K) */
int xm_linear_dgrad_f32(const float* dy, const float* w, float* dx, int64_t M, int64_t N, int64_t K, int64_t lddy,
                        int64_t ldw, int64_t lddx, int round_out, void* stream);
/* dw (N,K) = dy (M,N)^T @ x (M,K); db (N) = column sums of dy (db may be NULL).
 * workspace: splits*N*K floats when splits > 1; db_workspace: xm_colsum_nsplit(M, N)*N floats when db != NULL. */
int xm_linear_wgrad_f32(const float* dy, const float* x, float* dw, float* db, int64_t M, int64_t N, int64_t K,
                        int64_t lddy, int64_t ldx, int64_t lddw, int splits, float* workspace, float* db_workspace,
                        void* stream);

/* ------------------------------------------------------------------ dense Conv1d ("same" padding, stride 1)
 * EEG_CODE/enhanced_models_v4.py:128-144 ; EEG_CODE/crossmodal_v4_enhancements.py:822-834,854-866
 * Activations are CHANNELS-LAST on the device: x (B, T, C) with row pitch ld (elements, % 4 == 0);
 * one row per time step.  (TMA inner coordinates must be 16-byte aligned, so a conv tap has to
 * be a shift of the row coordinate; this is also the (B, L, D) layout the transformer tail wants.) */

/* Repack w (Cout, Cin, taps) into wk (taps, Cout, ldk) and wt (taps, Cin, ldt), tf32-rounded,
 * zero padded; ldk >= Cin, ldt >= Cout, both multiples of 4. */
int xm_conv1d_pack_weight_f32(const float* w, int64_t Cout, int64_t Cin, int64_t taps, float* wk, int64_t ldk,
                              float* wt, int64_t ldt, void* stream);
/* y (B,T,Cout) = conv1d(x (B,T,Cin), w) + bias, pad = taps/2. */
int xm_conv1d_fwd_f32(const float* x, const float* wk, const float* bias, float* y, int64_t B, int64_t Cin,
                      int64_t Cout, int64_t T, int64_t taps, int64_t ldx, int64_t ldk, int64_t ldy, int round_out,
                      void* stream);
/* The same convolution that also emits the BatchNorm batch statistics of its output: stat_part
 * (xm_conv1d_fwd_stat_rows(), Cout, 2) doubles = per-(CTA, lane quadrant) column sums and sums of squares of y over
 * (B, T), the `partials` argument of xm_bn_finalize_stats -- accumulated in the epilogue warps (fp32 butterfly per
 * 32-row block, fp64 across blocks), so xm_bn_partial_stats_f32's pass over y is not needed
 * (EEG_CODE/enhanced_models_v4.py:128-144: Conv1d -> BatchNorm1d).  Cout <= the N tile (256), y 16-B aligned, ldy % 4 == 0;
 * XM_ERR_UNSUPPORTED otherwise (callers then use the two-kernel path).  stat_part may be NULL (plain convolution).
 * x_cols (0 = Cin): the number of channels x physically stores when that is fewer than the Cin the contraction walks:
 * channel coordinates >= x_cols wrap back by x_cols, so the 3-pass convolution over [hi | lo | hi] x [wh | wh | wl]
 * (Cin = 3 C) reads a tensor that holds [hi | lo] (x_cols = 2 C, a multiple of 32, taps > 1). */
int xm_conv1d_fwd_stat_rows(void);
int xm_conv1d_fwd_stats_f32(const float* x, const float* wk, const float* bias, float* y, double* stat_part, int64_t B,
                            int64_t Cin, int64_t Cout, int64_t T, int64_t taps, int64_t ldx, int64_t ldk, int64_t ldy,
                            int round_out, int64_t x_cols, void* stream);
/* dx (B,T,Cin) from dy (B,T,Cout) */
int xm_conv1d_dgrad_f32(const float* dy, const float* wt, float* dx, int64_t B, int64_t Cin, int64_t Cout, int64_t T,
                        int64_t taps, int64_t lddy, int64_t ldt, int64_t lddx, int round_out, void* stream);
/* dw (Cout,Cin,taps) [reference layout], db (Cout) (db may be NULL).
 * workspace: xm_conv1d_wgrad_workspace() floats; db_workspace: xm_colsum_nsplit(B*T, Cout)*Cout floats when db != NULL. */
int64_t xm_conv1d_wgrad_workspace(int64_t B, int64_t Cin, int64_t Cout, int64_t taps);
int xm_conv1d_wgrad_f32(const float* dy, const float* x, float* dw, float* db, int64_t B, int64_t Cin, int64_t Cout,
                        int64_t T, int64_t taps, int64_t lddy, int64_t ldx, float* workspace, float* db_workspace,
                        void* stream);

/* ------------------------------------------------------------------ normalisation + activation (+pool, +dropout)
 * nn.BatchNorm1d (train mode) + nn.GELU/nn.ReLU + nn.MaxPool1d(2) + nn.Dropout chains of the encoders,
 * on channels-last activations y (B, T, C) with row pitch ldy (nn.Linear outputs: T = 1). */

/* Per-split {sum, sumsq} of every channel over R = B*T rows: partials (nsplit, C, 2) doubles,
 * nsplit = xm_bn_nsplit(R, C).  (Under data parallelism the partials are what gets all-reduced.) */
int xm_bn_nsplit(int64_t R, int64_t C);
int xm_bn_partial_stats_f32(const float* y, int64_t R, int64_t C, int64_t ldy, double* partials, void* stream);
/* mean / invstd from partials; total_count = rows behind them.  running_mean/var (may be NULL) are
 * updated with `momentum` and the unbiased variance, as torch does. */
int xm_bn_finalize_stats(const double* partials, int nsplit, int64_t C, double total_count, float eps, float* mean,
                         float* invstd, float* running_mean, float* running_var, float momentum, void* stream);
/* out = drop(pool?(act(gamma*(y-mean)*invstd + beta))) ; pool: 0 none, 2 = MaxPool1d(2) over T
 * (out has B*(T/2) rows).  drop_p in [0,1): keep with prob 1-p, scale 1/(1-p); mask from
 * (seed, element index).  drop_before_pool selects Conv-BN-GELU-Drop-Pool (Lite) vs
 * Conv-BN-GELU-Pool-Drop (v4) ordering.  round_out: 0 fp32, 1 rounded to tf32, 2 = the tf32 split [hi | lo] along the
 * channel axis (ldo >= 2*C), the operand of a following 3-pass conv (xm_conv1d_fwd_stats_f32, x_cols = 2*C). */
int xm_bn_act_fwd_f32(const float* y, const float* mean, const float* invstd, const float* gamma, const float* beta,
                      float* out, int64_t B, int64_t T, int64_t C, int64_t ldy, int64_t ldo, int act, int pool,
                      float drop_p, uint64_t seed, int drop_before_pool, int round_out, void* stream);
/* Backward, pass 1: per-channel partial sums of dz and dz*xhat (dz = grad wrt the BN output);
 * partials (xm_bn_nsplit(B*T, C), C, 2) doubles.  Masks / argmax are recomputed, not stored. */
int xm_bn_act_bwd_reduce_f32(const float* dout, const float* y, const float* mean, const float* invstd,
                             const float* gamma, const float* beta, int64_t B, int64_t T, int64_t C, int64_t ldy,
                             int64_t ldo, int act, int pool, float drop_p, uint64_t seed, int drop_before_pool,
                             double* partials, void* stream);
/* dbeta (C) = sum dz, dgamma (C) = sum dz*xhat */
int xm_bn_bwd_finalize(const double* partials, int nsplit, int64_t C, float* dbeta, float* dgamma, void* stream);
/* Backward, pass 2: dy = gamma*invstd*(dz - dbeta/count - xhat*dgamma/count) */
int xm_bn_act_bwd_apply_f32(const float* dout, const float* y, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, const float* dbeta, const float* dgamma,
                            double total_count, float* dy, int64_t B, int64_t T, int64_t C, int64_t ldy, int64_t ldo,
                            int act, int pool, float drop_p, uint64_t seed, int drop_before_pool, int round_out,
                            void* stream);
/* AdaptiveAvgPool1d(1): x (B, T, C) pitch ldx -> out (B, C); and its backward (broadcast / T). */
int xm_seqmean_f32(const float* x, int64_t B, int64_t T, int64_t C, int64_t ldx, float* out, void* stream);
int xm_seqmean_bwd_f32(const float* dout, int64_t B, int64_t T, int64_t C, int64_t lddx, float* dx, void* stream);

/* LayerNorm over the last dim + activation + dropout on (M, D) rows (bridge_utils.py:36-38,42-44,62-64).
 * Saves mean/rstd (M each) for the backward. */
int xm_ln_act_fwd_f32(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd,
                      int64_t M, int64_t D, float eps, int act, float drop_p, uint64_t seed, void* stream);
/* dx (M,D); dgamma/dbeta partials (nblk, D) reduced by xm_colsum_f32. nblk = xm_ln_nblk(M). */
int xm_ln_nblk(int64_t M);
int xm_ln_act_bwd_f32(const float* dout, const float* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, float* dx, float* dgamma_part, float* dbeta_part, int64_t M, int64_t D,
                      int act, float drop_p, uint64_t seed, void* stream);

/* out = drop(act(x)) over n contiguous elements, and its backward dx = dout * mask * act'(x)
 * (nn.GELU/ReLU/Tanh/Sigmoid + nn.Dropout after a projection). */
int xm_act_fwd_f32(const float* x, float* out, int64_t n, int act, float drop_p, uint64_t seed, int round_out,
                   void* stream);
int xm_act_bwd_f32(const float* dout, const float* x, float* dx, int64_t n, int act, float drop_p, uint64_t seed,
                   int round_out, void* stream);

/* xm_act_bwd_f32 over (M, C) rows that also returns per-block column sums of dx in colsum_part
 * (xm_act_bwd_colsum_nblk(M, C), C) -- the bias gradient of the Linear in front of the activation, without a
 * second pass over dx (reduce the partials with xm_colsum_f32).  C % 4 == 0. */
int xm_act_bwd_colsum_nblk(int64_t M, int64_t C);
int xm_act_bwd_colsum_f32(const float* dout, const float* x, float* dx, int64_t M, int64_t C, int act, float drop_p,
                          uint64_t seed, int round_out, float* colsum_part, void* stream);

/* out = x rounded to nearest tf32 (10-bit mantissa) in an fp32 container: operands handed to the tensor
 * cores are rounded at their producer so the contraction sees no truncation bias (may run in place). */
int xm_round_tf32_f32(const float* x, float* out, int64_t n, void* stream);

/* 3-way tf32 split of x (rows, cols) concatenated along `axis` (0: out (3*rows, cols), 1: out (rows, 3*cols)):
 * which == 0 -> [hi | lo | hi], which == 1 -> [hi | hi | lo], hi = tf32(x), lo = tf32(x - hi).  A which-0
 * operand contracted with a which-1 operand over the tripled axis yields an fp32-accurate product from
 * three tf32 tensor-core passes ("precise" mode of the small projections: fMRI MLPs, bridge, heads). */
int xm_split3_f32(const float* x, float* out, int64_t rows, int64_t cols, int which, int axis, void* stream);

/* out (N) = column sums of x (M, N) (bias gradients, partial reductions), deterministic two-stage
 * reduction; workspace: xm_colsum_nsplit(M, N) * N floats (may be NULL when nsplit == 1). */
int xm_colsum_nsplit(int64_t M, int64_t N);
int xm_colsum_f32(const float* x, int64_t M, int64_t N, int64_t ldx, float* out, float* workspace, void* stream);
/* ------------------------------------------------------------------ similarity + symmetric InfoNCE
 * No reference implementation; embeddings are the outputs of
 * bridge_utils.py:71-72.  Oracle: oracle/infonce.py. */

/* xn = x / max(||x||_2, eps) per row (F.normalize), tf32-rounded; inv_norm (M) saved. */
int xm_l2norm_fwd_f32(const float* x, float* xn, float* inv_norm, int64_t M, int64_t D, float eps, void* stream);
/* Same normalisation kept in full fp32 (xn) plus xs (M, 3D), the 3-way tf32 split of xn laid out along the
 * contraction axis: which == 0 -> [hi | lo | hi], which == 1 -> [hi | hi | lo].  The dot product of a
 * which-0 row with a which-1 row is hi*hi + lo*hi + hi*lo: an fp32-accurate similarity on the tf32
 * tensor cores (S is divided by the temperature 0.07 before the softmax, which would otherwise
 * amplify tf32 rounding into ~4e-3 relative error of every probability and gradient). */
int xm_l2norm_split_fwd_f32(const float* x, float* xn, float* xs, float* inv_norm, int64_t M, int64_t D, float eps,
                            int which, void* stream);
/* dx = (dxn - xn * <xn, dxn>) * inv_norm */
int xm_l2norm_bwd_f32(const float* dxn, const float* xn, const float* inv_norm, float* dx, int64_t M, int64_t D,
                      void* stream);
/* S (Ml, Ng) = a (Ml, D) @ b (Ng, D)^T * inv_tau, materialised (bridge_utils.similarity_matrix). */
int xm_similarity_f32(const float* a, const float* b, float* S, int64_t Ml, int64_t Ng, int64_t D, float inv_tau,
                      void* stream);
/* Row-wise logsumexp of S = a @ b^T * inv_tau without materialising S:
 * lse (Ml) ; diag (Ml) = S[i, i + diag_off].  Rows of a and b must be unit-norm (|S| <= inv_tau).
 * workspace: Ml * ceil(Ng/ xm_infonce_tile_n()) floats. */
int xm_infonce_tile_n(void);
int xm_infonce_lse_f32(const float* a, const float* b, float* lse, float* diag, int64_t Ml, int64_t Ng, int64_t D,
                       float inv_tau, int64_t diag_off, float* workspace, void* stream);
/* G (Ml, Ng) = coef * (exp(S - lse_row[i]) + exp(S - lse_col[j]) - 2*[j == i + diag_off]); round_out: rounded to tf32 (feeds a single-pass product) */
int xm_infonce_grad_f32(const float* a, const float* b, const float* lse_row, const float* lse_col, float* G,
                        int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off, float coef, int round_out,
                        void* stream);

/* dx (Ml, D) = G @ f_n in the fp32-accurate 3-pass mode: g3 (Ml, 3*Ng) = xm_split3_f32(G, which, axis 1) and
 * f3 (Ng, 3*D) = the xm_l2norm_split_fwd_f32 split of the unit vectors with the COMPLEMENTARY `which`
 * (the rows of G sum to ~0, so this product cancels heavily: single-pass tf32 is not enough).  Ng % 4 == 0. */
int xm_infonce_dgrad_f32(const float* g3, const float* f3, float* dx, int64_t Ml, int64_t Ng, int64_t D, void* stream);

/* Fused backward of the symmetric InfoNCE loss (csrc/infonce_fused.cu): both softmax-gradient blocks of
 * xm_infonce_grad_f32 are formed tile by tile in tensor memory and contracted in place,
 *   de (Ml, D) = G1 f_n,  G1 = coef * (exp(S - lse_ef[i]) + exp(S - lse_fe_all[j]) - 2 [j == i + diag_off]),  S = e_n f_n^T * inv_tau
 *   df (Ml, D) = G2 e_n,  G2 = the same with e and f exchanged (lse_fe[i], lse_ef_all[j]),
 * without writing anything of size (Ml, Ng).  e3 / f3 (Ml, 3D): this rank's xm_l2norm_split_fwd_f32 splits (which = 0 / 1),
 * e3_all / f3_all (Ng, 3D): the splits of the global batch (all ranks' rows, this rank's at row diag_off).
 * precise != 0: G is split hi/lo on the fly and the contraction runs in the 3-pass mode (fp32-accurate, as
 * xm_infonce_dgrad_f32); 0: one tf32 pass.  The sum over the global batch is accumulated in fp32 registers per
 * 128-column chunk.  D == 128, Ml % 128 == Ng % 128 == diag_off % 128 == 0 (xm_infonce_bwd_fused_supported).
 * workspace: xm_infonce_bwd_fused_workspace(Ml, Ng, D) floats, 32-B aligned (column scales, the transposed unit vectors,
 * per-unit partial outputs when a row tile is split over more than two units).  Results are bit-reproducible. */
int xm_infonce_bwd_fused_supported(int64_t Ml, int64_t Ng, int64_t D, int64_t diag_off);
/* Forward of the same kernel family: lse_ef[i] = logsumexp_j S[i, j] (my e x all f), lse_fe[i] = logsumexp_j S'[i, j] (my f x
 * all e), diag[i] = S[i, i + diag_off], scores in the 3-pass mode, nothing of size (Ml, Ng) written; same shape rules.
 * workspace: xm_infonce_lse_fused_workspace(Ml, Ng) floats (per-thread partial sums, combined in a fixed order). */
int64_t xm_infonce_lse_fused_workspace(int64_t Ml, int64_t Ng);
int xm_infonce_lse_fused_f32(const float* e3, const float* f3, const float* e3_all, const float* f3_all, float* lse_ef,
                             float* lse_fe, float* diag, int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off,
                             float* workspace, void* stream);
int64_t xm_infonce_bwd_fused_workspace(int64_t Ml, int64_t Ng, int64_t D);
int xm_infonce_bwd_fused_f32(const float* e3, const float* f3, const float* e3_all, const float* f3_all, const float* lse_ef,
                             const float* lse_fe, const float* lse_ef_all, const float* lse_fe_all, float* de, float* df,
                             int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off, float coef, int precise,
                             float* workspace, void* stream);

/* Fused all-gather + contraction over NVLink peer memory (data-parallel global negatives).  The second
 * operand is ROW-SHARDED: rank r holds rows [r*rows_per_peer, (r+1)*rows_per_peer) in its own buffer and
 * b_peers[r] (a HOST array of n_peers <= 8 device pointers) is that buffer mapped into this process
 * (torch symmetric memory / cudaIpc / cuMem peer mapping).  The TMA producer of the GEMM reads every tile
 * straight from the owner's HBM through NVLink / NVSwitch: no gathered copy is ever materialised.
 * rows_per_peer % 128 == 0.  The caller orders the kernels after the peers' writes (device barrier). */
int xm_infonce_lse_peers_f32(const float* a, const void* const* b_peers, int n_peers, int64_t rows_per_peer, float* lse,
                             float* diag, int64_t Ml, int64_t D, float inv_tau, int64_t diag_off, float* workspace,
                             void* stream);
int xm_infonce_grad_peers_f32(const float* a, const void* const* b_peers, int n_peers, int64_t rows_per_peer,
                              const float* lse_row, const float* lse_col, float* G, int64_t Ml, int64_t D, float inv_tau,
                              int64_t diag_off, float coef, int round_out, void* stream);
/* All-gather through the same peer mappings: dst (n_peers * elems_per_peer) <- concatenation of the ranks'
 * shards, one launch of 128-bit peer loads.  Used instead of the in-GEMM peer reads when a rank has many
 * row tiles (every row tile would otherwise re-fetch every remote tile across NVLink). */
int xm_peer_gather_f32(const void* const* src_peers, int n_peers, int64_t elems_per_peer, float* dst, void* stream);
/* Small all-reduce (SUM) of n doubles over peer memory (csrc/peer_exchange.cu; SyncBN statistics of the data-parallel
 * step): this rank pushes x into row `rank` of the current slot of every peer's symmetric buffer (data_dst[p], p <
 * n_peers <= 8, self included), publishes `seq` at flag_dst[p], waits until slot_flags[0..n_peers) of ITS OWN buffer
 * all equal seq, and writes out[i] = sum over r of slot_data[r * row_stride + i] (rank order: identical on all ranks).
 * The caller rotates >= 2 slots per channel and increases seq by one per call of the channel (seq > 0; flags start
 * at 0); calls of one channel are issued in the same order on every rank.  A peer that never arrives poisons the
 * result with NaN after ~10 s instead of hanging the device.  The number actually published is seq + (seed epoch << 32)
 * (xm_seed_epoch_*): a call captured in a CUDA graph stays distinct from replay to replay; seq itself must be < 2^32. */
int xm_peer_allreduce_f64(const double* x, double* out, int64_t n, const void* const* data_dst, const void* const* flag_dst,
                          int n_peers, const double* slot_data, const uint64_t* slot_flags, int64_t row_stride, uint64_t seq,
                          void* stream);
/* dx (M, K) = dy (M, n_peers*rows_per_peer) @ w, w row-sharded across peers (pitch ldw). */
int xm_linear_dgrad_peers_f32(const float* dy, const void* const* w_peers, int n_peers, int64_t rows_per_peer, float* dx,
                              int64_t M, int64_t K, int64_t lddy, int64_t ldw, int64_t lddx, int round_out, void* stream);

/* ------------------------------------------------------------------ multi-head self-attention core
 * nn.MultiheadAttention inside TemporalTransformerBlock (EEG_CODE/enhanced_models_v4.py:71-73,98; torch computes
 * softmax(q k^T / sqrt(dh)), dropout on the weights, times v).  qkv (B, L, 3*H*dh) is the packed in_proj output
 * [q | k | v], head h at columns h*dh inside each third; out (B, L, H*dh); lse (B*H, L).  No kernel stores the L x L
 * probabilities (the reference's need_weights path materialises them only to discard them). */

/* Fused tcgen05 kernels (head dim 32, L <= 512): the L x L probability / score-gradient matrices stay in tensor
 * memory -- the forward keeps only out and lse (B*H, L); the backward regenerates the probabilities from q, k and lse
 * and needs the forward's `out` (delta = dO . O) plus a (B*H, L) float workspace `delta`.
 * The dropout mask is a pure function of (seed, slab, query, key), exported by xm_attn_fused_mask_u8
 * (mask (B*H, L, L), 1 = kept) so tests can replay it. */
int xm_attn_fused_fwd_f32(const float* qkv, float* out, float* lse, int64_t B, int64_t L, int64_t H, int64_t dh, float scale,
                          float drop_p, uint64_t seed, int round_out, void* stream);
/* dbias_part (may be NULL): (xm_attn_fused_bwd_nblk(), 3*H*dh) partial column sums of dqkv (zeroed by the call; reduce
 * with xm_colsum_f32): the bias gradient of the in-projection, accumulated where dq / dk / dv leave tensor memory. */
int xm_attn_fused_bwd_nblk(void);
int xm_attn_fused_bwd_f32(const float* dout, const float* qkv, const float* out, const float* lse, float* dqkv, float* delta,
                          float* dbias_part, int64_t B, int64_t L, int64_t H, int64_t dh, float scale, float drop_p,
                          uint64_t seed, int round_out, void* stream);
int xm_attn_fused_mask_u8(uint8_t* mask, int64_t B, int64_t L, int64_t H, float drop_p, uint64_t seed, void* stream);
/* Shape-general variant (csrc/attention_general.cu; SIMT fp32): every shape the fused kernel does not cover --
 * head dim <= 256, any L with xm_attn_general_supported(L, dh) != 0 -- plus nn.MultiheadAttention's attn_mask
 * (EEG_CODE/enhanced_models_v4.py:89,98 passes `mask` through): `mask` is ADDITIVE fp32 (-inf = not allowed to
 * attend), shape (L, L) when mask_per_head == 0 or (B*H, L, L) when 1, or NULL.  A fully masked row yields NaN, as
 * torch's softmax does.  lse / delta: (B*H, L).  Dropout mask = hash(seed, (bh*L + query)*L + key). */
int xm_attn_general_supported(int64_t L, int64_t dh);
int xm_attn_general_fwd_f32(const float* qkv, const float* mask, int mask_per_head, float* out, float* lse, int64_t B,
                            int64_t L, int64_t H, int64_t dh, float scale, float drop_p, uint64_t seed, int round_out,
                            void* stream);
int xm_attn_general_bwd_f32(const float* dout, const float* qkv, const float* mask, int mask_per_head, const float* lse,
                            float* dqkv, float* delta, int64_t B, int64_t L, int64_t H, int64_t dh, float scale,
                            float drop_p, uint64_t seed, int round_out, void* stream);
/* Debug: when non-NULL, CTA 0 of the fused attention kernels appends clock64() stamps at its phase boundaries
 * (3 roles x 4096 slots of int64). */
int xm_debug_set_attn_trace(int64_t* device_buffer);

/* ------------------------------------------------------------------ residual stream of the pre-norm transformer block
 * EEG_CODE/enhanced_models_v4.py:89-107 (x + Dropout(branch), LayerNorm) and :44-55 (x + pe, Dropout), fused:
 *   s = x + Dropout(a)            [a may be NULL]      or      s = Dropout(x + pe[row % L])   [pe may be NULL]
 *   h = tf32(LayerNorm(s) * gamma + beta);   mean / rstd (M) saved.
 * s_out may be NULL when a == pe == NULL (s == x).  D % 128 == 0, 128 <= D <= 512 (xm_resid_ln_supported). */
int xm_resid_ln_supported(int64_t D);
int xm_resid_ln_fwd_f32(const float* x, const float* a, const float* pe, int64_t L, const float* gamma, const float* beta,
                        float* s_out, float* h, float* mean, float* rstd, int64_t M, int64_t D, float eps, float drop_p,
                        uint64_t seed, void* stream);
/* ds = dres + LayerNormBackward(dh) [dres may be NULL];  dx = ds;  da = tf32(mask * ds / (1-p)) [da may be NULL];
 * dgamma_part / dbeta_part / dabias_part: (xm_resid_ln_nblk(M), D) per-block partial sums (reduce with
 * xm_colsum_f32); dabias_part (may be NULL) = column sums of da, i.e. the bias gradient of the Linear that
 * produced the branch a -- accumulated here so that no separate pass over da is needed. */
int xm_resid_ln_nblk(int64_t M);
int xm_resid_ln_bwd_f32(const float* dh, const float* dres, const float* s, const float* gamma, const float* mean,
                        const float* rstd, float* dx, float* da, float* dgamma_part, float* dbeta_part, float* dabias_part,
                        int64_t M, int64_t D, float drop_p, uint64_t seed, void* stream);
/* out (B, D) = mean over T of x + Dropout(a)  (last residual add + AdaptiveAvgPool1d(1), :161-163), and its
 * backward: dx = dout / T broadcast, da = tf32(mask * dx / (1-p)); either output may be NULL. */
int xm_resid_seqmean_fwd_f32(const float* x, const float* a, int64_t B, int64_t T, int64_t D, float* out, float drop_p,
                             uint64_t seed, void* stream);
int xm_resid_seqmean_bwd_f32(const float* dout, int64_t B, int64_t T, int64_t D, float* dx, float* da, float drop_p,
                             uint64_t seed, void* stream);

/* ------------------------------------------------------------------ fused feed-forward branch of the transformer block
 * EEG_CODE/enhanced_models_v4.py:79-80,102-105: linear2(Dropout(act(linear1(x)))) with the (M, hidden) intermediate kept
 * on chip (csrc/ffn_fused.cu).  xm_ffn_fused_supported: D == 128, hidden % 128 == 0, hidden <= 1024, act GELU | RELU.
 * x (M, D), w1 (hidden, D), w2 (D, hidden): tf32-rounded by their producers; b1 (hidden), b2 (D); y / a / dh / dx 32-B aligned.
 *   y = tf32(Dropout(act(x w1^T + b1))) w2^T + b2 */
int xm_ffn_fused_supported(int64_t D, int64_t hidden, int act);
int xm_ffn_fused_fwd_f32(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y, int64_t M,
                         int64_t D, int64_t hidden, int act, float drop_p, uint64_t seed, void* stream);
/* Data-gradient half of the backward: recomputes the hidden activations from x and emits, in ONE pass,
 *   a  (M, hidden) = Dropout(act(x w1^T + b1))           operand of dw2 = dy^T a   (xm_linear_wgrad_f32)
 *   dh (M, hidden) = dy w2 * act'(.) * mask / (1 - p)    operand of dw1 = dh^T x
 *                    (both pre-scaled by 1 + 0.7213 * 2^-11, so that the tensor core's truncation to tf32 is zero-mean)
 *   dx (M, D)      = dh w1
 *   db1_part (xm_ffn_fused_nblk(M), hidden): partial column sums of dh (reduce with xm_colsum_f32 -> db1); may be NULL.
 * w2t = w2^T (hidden, D) and w1t = w1^T (D, hidden) are transposed tf32 copies, so every weight operand is K-major. */
int xm_ffn_fused_nblk(int64_t M);
int xm_ffn_fused_dgrad_f32(const float* x, const float* dy, const float* w1, const float* b1, const float* w2t, const float* w1t,
                           float* a, float* dh, float* dx, float* db1_part, int64_t M, int64_t D, int64_t hidden, int act,
                           float drop_p, uint64_t seed, void* stream);
/* mask (M, hidden), 1 = kept: the dropout mask both kernels generate for (drop_p, seed), so tests can replay it. */
int xm_ffn_fused_mask_u8(uint8_t* mask, int64_t M, int64_t hidden, float drop_p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------ bridge head: cross-attention over two tokens
 * bridge_utils.py:74-83 (nn.MultiheadAttention(query = eeg token, key = value = [eeg, fmri])): per (sample, head) two
 * scores, a two-way softmax, dropout on the weights, a weighted sum of two value vectors (csrc/bridge_head.cu).
 * q (B, H*dh); kv (2B, 2*H*dh): row t*B + b = token t of sample b, columns [0, d) keys, [d, 2d) values (the packed k / v
 * in-projection of the stacked tokens).  out (B, H*dh); att (B, H, 2) = the dropped, rescaled weights (what
 * need_weights returns before the head average).  The backward recomputes the softmax from q and k and regenerates
 * the mask from (seed, sample, head, token): dq (B, d), dkv (2B, 2d). */
int xm_cross2_attn_fwd_f32(const float* q, const float* kv, float* out, float* att, int64_t B, int64_t H, int64_t dh, float drop_p,
                           uint64_t seed, void* stream);
int xm_cross2_attn_bwd_f32(const float* dout, const float* q, const float* kv, float* dq, float* dkv, int64_t B, int64_t H, int64_t dh,
                           float drop_p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------ clip_grad_norm_ + AdamW over one flat bucket
 * The step recipe of _test_bridge.py:775-788,869 (run_fmri_v11.py:430-450, run_training_lite.py:478-489):
 * clip_grad_norm_(max_norm) then torch.optim.AdamW.step, as two launches over flat fp32 buffers (csrc/optimizer.cu).
 * partials: xm_sumsq_nblk(n) doubles.  xm_clip_adamw_f32: coef = min(1, max_norm / (sqrt(sum partials) + 1e-6))
 * (max_norm <= 0: no clipping); g <- coef * g; AdamW update of p, m, v for 1-based `step`; norm_out (1 float, may be
 * NULL) receives the pre-clip total norm. */
int xm_sumsq_nblk(int64_t n);
int xm_sumsq_partials_f32(const float* g, int64_t n, double* partials, void* stream);
int xm_clip_adamw_f32(float* p, float* g, float* m, float* v, int64_t n, const double* partials, int nblk, float max_norm,
                      float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float* norm_out,
                      void* stream);
/* The same update with the 1-based step count (int64) and the learning rate (float) read from DEVICE memory, so a step
 * captured in a CUDA graph replays with the values the host (or an increment node of the graph) left there. */
int xm_clip_adamw_dev_f32(float* p, float* g, float* m, float* v, int64_t n, const double* partials, int nblk, float max_norm,
                          const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, const int64_t* step_dev,
                          float* norm_out, void* stream);

/* ------------------------------------------------------------------ seed epoch (CUDA-graph replay of the step)
 * Every dropout mask of this library is hash(seed, element); a captured launch freezes `seed`.  The library therefore
 * folds a device-resident 64-bit epoch into every hash: masks = f(seed + epoch * odd constant, element).  The epoch is 0
 * (seeds used as passed) until these entry points change it.
 * xm_seed_epoch_init: allocates the counter on the current device (call once, outside stream capture).
 * xm_seed_epoch_advance: epoch += 1 on `stream` -- one kernel node + seven 8-byte device-to-device copies, capturable:
 * placed first in a captured step, every replay draws fresh masks, identical in its forward and backward.
 * xm_seed_epoch_set / _get: set (on `stream`) / read back (synchronous) the epoch, for tests and eager replays. */
int xm_seed_epoch_init(void);
int xm_seed_epoch_advance(void* stream);
int xm_seed_epoch_set(uint64_t value, void* stream);
int xm_seed_epoch_get(uint64_t* value_out);

/* dst[dst_off[t] .. + numel[t]) = src[t][0 .. numel[t]) for t < n_tensors, one launch per 96 tensors: gathers the
 * per-parameter gradient tensors autograd produces into the flat bucket.  src / dst_off / numel are HOST arrays. */
int xm_gather_flat_f32(const void* const* src, const int64_t* dst_off, const int64_t* numel, int n_tensors, float* dst,
                       void* stream);

/* ------------------------------------------------------------------ diagnostics (not on the product path)
 * Dump the raw shared-memory image of one TMA box {32,32} loaded at (c0, c1) from a (rows, cols)
 * fp32 matrix; swizzle_atom32 selects SWIZZLE_128B_ATOM_32B instead of SWIZZLE_128B. */
/* A/B switch for the conv-wgrad operand staging (1: one halo tile per k-block serves every tap through
 * descriptor start-address shifts; 0: one shifted TMA copy per tap).  Returns the previous setting. */
int xm_debug_set_conv_halo(int on);
/* When non-NULL (3 * 8192 int64), the next xm_ffn_fused_fwd_f32 launches run the tracing instance of the kernel:
 * CTA 0 logs, per MMA-schedule step, (kind, chunk, start clock, end clock, cycles waited on the weight ring, cycles
 * waited on the other roles) for the TMA producer [0, 8192) and the MMA issuer [8192, 16384), and per transformed
 * chunk (4, chunk, wait start, accumulator ready, handed back) for one warp of each transform group [16384 + g*4096). */
int xm_debug_set_ffn_trace(int64_t* device_buffer);
/* Timing experiments of the tracing instance (its results are then wrong): bit 0 = transform warps skip their
 * arithmetic, bit 1 = the weight ring is loaded once and only re-signalled. */
int xm_debug_set_ffn_flags(int flags);
int xm_debug_tma_probe(const float* src, int64_t rows, int64_t cols, int64_t ld, int c0, int c1, int swizzle_atom32,
                       float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMODAL_B200_H_ */
