/*
 * signaltrain_b200.h -- C ABI of the B200-native SignalTrain train-step library.
 *
 * The reference has no FFI for this path: its boundary is a Python class API.
 * Each entry point below replaces the reference call named beside it; the Python mirror in
 * signaltrain_b200/ binds them with ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; st_last_error() gives the message.
 *     Nothing throws across the ABI.
 *   - all pointers are DEVICE pointers to float32 unless the name ends in _host.  Buffers are
 *     borrowed for the duration of the call (asynchronous on `stream`); no ownership moves.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - one handle per device; a handle is not thread-safe (the reference drives the device from a
 *     single host thread, signaltrain/train.py:104-151).
 *   - `params`, `grads`, `exp_avg`, `exp_avg_sq` are host arrays of ST_NUM_PARAMS device pointers in
 *     the reference's state_dict order (st_param_name / st_param_numel describe each slot):
 *       0..3   mpaec.dft_analysis.conv_analysis_{real,imag}.weight, mpaec.dft_synthesis.conv_synthesis_{real,imag}.weight  (N,1,N)
 *       4..21  mpaec.aenc.{fnn_enc,fnn_enc2,fnn_enc3,fnn_enc4,fnn_addknobs,fnn_dec4,fnn_dec3,fnn_dec2,fnn_dec}.{weight,bias}
 *       22..39 mpaec.phs_aenc.<same nine layers>.{weight,bias}
 */
#ifndef SIGNALTRAIN_B200_H
#define SIGNALTRAIN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ST_NUM_PARAMS 40
#define ST_NUM_ACTS 30
#define ST_ABI_VERSION 1

typedef struct st_handle st_handle;

/* Geometry of one model instance.  Mirrors st_model.__init__ (signaltrain/nn_proc.py:348-385). */
typedef struct st_config {
    int chunk;       /* C : input samples per window            nn_proc.py:357 */
    int ft;          /* N : DFT size / taps (1024)              nn_proc.py:370 */
    int hop;         /* H : hop (384)                           nn_proc.py:371 */
    int frames_in;   /* T : expected_time_frames                nn_proc.py:378 */
    int frames_out;  /* OT: output_time_frames                  nn_proc.py:379 */
    int knobs;       /* K : number of knobs                     nn_proc.py:363 */
    int rank;        /* R : decomposition rank (64)             nn_proc.py:279 */
} st_config;

/* Adam hyper-parameters (torch.optim.Adam defaults used at signaltrain/train.py:228). */
typedef struct st_adam {
    float lr;          /* learning rate for THIS step (train.py:150 lag is the caller's business) */
    float beta1, beta2, eps;
    int step;          /* 1-based step count t, for the bias corrections */
    float grad_scale;  /* multiply every gradient by this first (1/world_size after an allreduce-sum) */
    float max_norm;    /* L1 clip threshold over the four DFT tensors (nn_proc.py:299-302); <=0 disables */
} st_adam;

int st_abi_version(void);

/* st_model(...) constructor: nn_proc.py:348-385.  Allocates the device workspace lazily per batch. */
int st_create(const st_config* cfg, int device, st_handle** out);
void st_destroy(st_handle* h);
const char* st_last_error(const st_handle* h);   /* h may be NULL: last error of a failed st_create */

/* state_dict slot description. */
const char* st_param_name(const st_handle* h, int idx);
long st_param_numel(const st_handle* h, int idx);
int st_out_samples(const st_handle* h);           /* L = (OT-1)*hop - ft   nn_proc.py:380 */
int st_bins(const st_handle* h);                  /* F = ft/2+1            cls_fe_dft.py:24 */

/* Analysis / Synthesis .initialize(): cls_fe_dft.py:36-48 and :87-100 (incl. GLA window :133-163).
 * Writes the four (N,N) front-end matrices into params[0..3]. */
int st_init_frontend(st_handle* h, float* const* params, void* stream);

/* Analysis.forward(wave_form) alone (cls_fe_dft.py:50-58): x (B,C) -> an_real, an_imag (B,T,F) with the (N,1,N)
 * weights w_real / w_imag.  Synthesis.forward(real, imag) alone (cls_fe_dft.py:102-115): (B,OT,F) x2 -> wave (B,L).
 * Forward only; they use the handle's workspace (a following st_backward needs a fresh st_forward). */
int st_analysis(st_handle* h, const float* x, const float* w_real, const float* w_imag, int batch, float* an_real,
                float* an_imag, void* stream);
int st_synthesis(st_handle* h, const float* real, const float* imag, const float* w_real, const float* w_imag, int batch,
                 float* wave, void* stream);

/* DCT / MDCT front-end variant, signaltrain/cls_fe_dct_bases.py (the reference ships it but wires it to nothing):
 *   Analysis.forward (:128-135): Conv1d(1 -> ft_size, kernel w_size, stride hop, padding ft_size, bias) then transpose:
 *     x (B, chunk) -> out (B, frames, ft_size), frames = (chunk + 2 ft_size - w_size) / hop + 1;  w (ft_size, 1, w_size), bias (ft_size)
 *   Synthesis.forward (:173-179): ConvTranspose1d(ft_size -> 1, kernel w_size, stride hop) trimmed by ft_size on both sides:
 *     x_ft (B, frames, ft_size) -> wave (B, (frames - 1) hop + w_size - 2 ft_size);  w (ft_size, 1, w_size)
 *   (tied_transform :36-54 is st_dct_synthesis with the analysis weights.)  Forward only; geometry is per call, the handle
 *   supplies the device, the GEMM engine and a grow-only workspace. */
int st_dct_analysis(st_handle* h, const float* x, const float* w, const float* bias, int batch, int chunk, int ft_size,
                    int w_size, int hop, float* out, void* stream);
int st_dct_synthesis(st_handle* h, const float* x_ft, const float* w, int batch, int frames, int ft_size, int w_size,
                     int hop, float* wave, void* stream);

/* Data step in front of the train step (the reference does it on CPU workers).
 *   st_compressor_4c: audio.compressor_4controls (audio.py:380-426) as Compressor_4c.go_wc applies it (:497-498), batched:
 *     x (batch, n) float32, knobs_wc (batch, 4) float64 world coordinates [threshold dB, ratio, attack s, release s]
 *     -> y (batch, n) float32 (the reference's float64 result after the caller's .float(), train.py:120).
 *   st_crop_windows: AudioFileDataSet.get_single_chunk's crop (datasets.py:236-241) and do_augment's polarity flip (:21-30)
 *     from a device-resident corpus: window b is corpus_x[offsets[b] : +chunk] and the last y_size samples of the same
 *     span of corpus_y, both times signs[b] (NULL: no flip).  offsets is a HOST array (the host RNG picks them). */
int st_compressor_4c(st_handle* h, const float* x, const double* knobs_wc, int batch, int n, double sr, float* y, void* stream);
int st_crop_windows(st_handle* h, const float* corpus_x, const float* corpus_y, long corpus_len, const long* offsets_host,
                    const float* signs, int batch, int chunk, int y_size, float* x, float* y, void* stream);

/* st_model.forward(x, knobs, return_acts): nn_proc.py:392 -> AsymMPAEC.forward :305-340.
 *   x (B,C)  knobs (B,K)  ->  y_hat (B,L) [= 2*y_hat of :340]  mag (B,T,F)  mag_hat (B,OT,F)
 *   acts: NULL, or ST_NUM_ACTS device pointers receiving the reference's layer_acts (:311-335), each
 *   contiguous in the reference's logical shape. */
int st_forward(st_handle* h, const float* x, const float* knobs, int batch,
               const float* const* params, float* y_hat, float* mag, float* mag_hat,
               float* const* acts, void* stream);

/* model.train() / model.eval() + torch.no_grad(): when training (the default) st_forward also saves the autoencoders'
 * hidden activations (~1.3 KB per (window, bin)) for the following st_backward; switch it off for validation / inference
 * (train.py:39-48, utils/predict_long.py:66) to skip those writes.  st_backward after a non-training forward still works
 * (it recomputes the chain with the slower SIMT kernels). */
int st_set_training(st_handle* h, int on);

/* Arithmetic of the contractions (the reference's `apex_opt` argument of train.train(), train.py:167-170,232-236, is its
 * only precision knob: "O0" fp32 ... "O2" mixed).  Storage, accumulation, loss, atan2 and the optimiser are fp32 in both modes.
 *   ST_PRECISION_FP32 (default): fp32-faithful products -- 3xTF32 tensor-core GEMMs on exact (hi, lo) operand pairs, exact
 *     fp32 FMA chains in the autoencoders; waveforms within 1e-5 of the reference.
 *   ST_PRECISION_TF32: every product of the five DFT contractions and of the autoencoder layers is ONE tensor-core TF32
 *     multiply (operands rounded to 10 mantissa bits, fp32 accumulate) -- the mixed-precision class of BASELINE configs 3 and 5
 *     (apex O1/O2 in the reference); waveforms within ~2e-3 of the fp32 path (tests/test_gpu_precision.py states the bound).
 * Takes effect at the next st_forward; a backward must follow a forward of the same mode. */
#define ST_PRECISION_FP32 0
#define ST_PRECISION_TF32 1
int st_set_precision(st_handle* h, int mode);
int st_get_precision(const st_handle* h);

/* loss_functions.calc_loss(y_hat, y, mag_hat, scale_by_freq=..., l1_lambda): loss_functions.py:26-43,
 * branches :34 (scale_by_freq NULL: l1_coef*mean|mag_hat|) and :36 (l1_coef*mean|mag_hat*s|; the caller
 * passes l1_coef = l1_lambda/10 as the reference does).  scale_by_freq is F floats (the reference
 * expands exp(7f/F) over (B,OT,F), train.py:115-117).
 * Writes the scalar loss to loss[0] and, if non-NULL, dLoss/dy_hat (B,L) and dLoss/dmag_hat (B,OT,F). */
int st_loss(st_handle* h, const float* y_hat, const float* y, const float* mag_hat,
            const float* scale_by_freq, float l1_coef, int batch,
            float* loss, float* g_y_hat, float* g_mag_hat, void* stream);

/* The same loss for tensors that do not come from this handle's model: wave tensors (batch, n_wave), spectra
 * (batch, n_frames, n_bins) -- e.g. utils/lr_finder.py:38 passes the INPUT magnitude (B, T, F) as mag_hat.
 * n_wave must be a multiple of 4.  Nothing but the three sizes is taken from the call; the handle only lends its
 * reduction scratch and device. */
int st_loss_shaped(st_handle* h, const float* y_hat, const float* y, const float* mag_hat,
                   const float* scale_by_freq, float l1_coef, int batch, int n_wave, int n_frames, int n_bins,
                   float* loss, float* g_y_hat, float* g_mag_hat, void* stream);

/* loss_functions.mae (loss_functions.py:22-23) -- validation metric, train.py:58. */
int st_mae(st_handle* h, const float* a, const float* b, long n, float* out, void* stream);

/* loss.backward() through st_model.forward (train.py:138).  Must follow the st_forward of the same
 * batch on the same handle (activations live in the handle's workspace).  g_mag / g_mag_hat may be NULL.
 * Writes (overwrites) all ST_NUM_PARAMS gradients. */
int st_backward(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat,
                int batch, const float* const* params, float* const* grads, void* stream);

/* The same backward in two calls, for data-parallel callers: after st_backward_begin the synthesis gradients (grads[2],
 * grads[3]) are final, so their allreduce can run while st_backward_finish computes the autoencoder and analysis gradients.
 * begin + finish == st_backward, bit for bit. */
int st_backward_begin(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat,
                      int batch, const float* const* params, float* const* grads, void* stream);
int st_backward_finish(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat,
                       int batch, const float* const* params, float* const* grads, void* stream);

/* model.clip_grad_norm_() (nn_proc.py:299-302, torch clip_grad_norm_ max_norm=1 norm_type=1 over the
 * four DFT tensors): scales grads[0..3] in place, writes the total L1 norm to total_norm[0] (device). */
int st_clip_grad_norm(st_handle* h, float* const* grads, float max_norm, float* total_norm, void* stream);

/* optimizer.step() of torch.optim.Adam (train.py:147), all tensors in one launch; optional fused
 * gradient scaling + L1 clip (hp->max_norm > 0) so the unfused clip pass is not needed. */
int st_adam_step(st_handle* h, float* const* params, const float* const* grads,
                 float* const* exp_avg, float* const* exp_avg_sq, const st_adam* hp, void* stream);

/* One whole iteration of the loop body train.py:112-147: forward, calc_loss, backward, clip, Adam.
 * y is (B,L) float32.  loss[0] (device) receives the scalar loss.  If allreduce-before-update is
 * needed (data parallel), call st_forward/st_loss/st_backward, reduce, then st_adam_step instead.
 * On a stream that can be captured (not the legacy default stream) the second call with the same tensor tables, batch size and
 * stream captures the step into a CUDA graph and later calls replay it (x, y, knobs, loss and the Adam scalars are refreshed in
 * the graph on every call, so the batch may live anywhere); ST_CUDA_GRAPH=0 disables that.  st_debug_graph_replays counts replays. */
long st_debug_graph_replays(const st_handle* h);
int st_train_step(st_handle* h, const float* x, const float* y, const float* knobs, int batch,
                  float* const* params, float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const float* scale_by_freq, float l1_coef,
                  const st_adam* hp, float* loss, void* stream);

/* The same iteration WITHOUT the update, for data-parallel callers: forward, calc_loss and backward (train.py:112-138) with
 * the fused forward tail of st_train_step; all ST_NUM_PARAMS gradients are written and loss[0] receives the scalar loss.  The
 * caller then sums the gradients over its ranks and calls st_adam_step with grad_scale = 1/world_size and max_norm = 1
 * (the L1 clip of nn_proc.py:299-302 is a function of the AVERAGED gradients, so it must follow the exchange). */
int st_grad_step(st_handle* h, const float* x, const float* y, const float* knobs, int batch,
                 float* const* params, float* const* grads, const float* scale_by_freq, float l1_coef,
                 float* loss, void* stream);

/* Data-parallel exchange payload.  Of the 16.8 MB of gradients only 8.5 MB carry information: analysis rows >= F never
 * receive gradient (cls_fe_dft.py:55-56) and the synthesis gradients are Hermitian (cls_fe_dft.py:109-110).
 * st_pack_grads gathers [Wr rows < F | Wi rows < F | Sr rows < F | Si rows < F | the 36 autoencoder tensors] into one
 * contiguous buffer of st_packed_grad_floats() floats (ONE collective covers it); st_unpack_grads scatters the reduced
 * buffer back and restores the mirrored synthesis rows, so all 40 gradient tensors are exactly what reducing them in
 * full would have given. */
long st_packed_grad_floats(const st_handle* h);
int st_pack_grads(st_handle* h, float* const* grads, float* packed, void* stream);
int st_unpack_grads(st_handle* h, const float* packed, float* const* grads, void* stream);
/* The data-parallel step in three calls around ONE collective: st_grad_step_packed = st_grad_step whose gradients leave directly
 * as the packed payload (no 40-tensor materialisation, no pack pass); [allreduce-sum of `packed`]; st_unpack_clip = st_unpack_grads
 * fused with the L1 norm of the four restored DFT tensors times grad_scale and the clip coefficient min(1, max_norm / (norm + 1e-6))
 * (nn_proc.py:299-302; total_norm may be NULL); st_adam_step_clipped = st_adam_step that uses that coefficient and skips the
 * analysis rows >= F, whose gradient is identically zero (cls_fe_dft.py:55-56). */
int st_grad_step_packed(st_handle* h, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                        float* packed, const float* scale_by_freq, float l1_coef, float* loss, void* stream);
int st_unpack_clip(st_handle* h, const float* packed, float* const* grads, float grad_scale, float max_norm, float* total_norm,
                   void* stream);
int st_adam_step_clipped(st_handle* h, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, const st_adam* hp, void* stream);

/* Measurement support (bench.py): number of kernels / device copies this handle has launched, and per-stage
 * device time bracketed with CUDA events on the launching stream.  st_profile_read synchronises the device,
 * fills ms[i] / calls[i] for i < st_profile_stage_count() with the totals since the previous read, and resets. */
long st_launch_count(const st_handle* h);
int st_profile_stage_count(void);
const char* st_profile_stage_name(int i);
int st_profile(st_handle* h, int enable);
int st_profile_read(st_handle* h, float* ms, long* calls);

/* Number of calls this handle served with a SIMT fallback kernel instead of the tensor-core kernels since creation
 * (a GEMM shape or autoencoder geometry they do not cover; forwards with return_acts count as well).  The benchmarked
 * path must read 0: bench.py prints it. */
long st_debug_fallbacks(const st_handle* h);

/* Diagnostic: cycle counters of the regions of the tensor-core autoencoder backward (64 values; reading resets). */
int st_debug_ae_timing(st_handle* h, int on, long long* out_host);

/* Test/diagnostic access to workspace buffers by name ("spec", "ri", "frames_out", "g_ri", "g_spec",
 * "wcat", "sfold").  Copies up to n floats to a HOST buffer, synchronising the device. */
int st_debug_read(st_handle* h, const char* name, float* dst_host, long n);
/* Test hook: C[M,N] = A*B on (hi, lo) tf32-pair operands through the tcgen05 (use_tc=1; 2 = promoted accumulation;
 * 3 = single-pass TF32 on the hi planes only) or FFMA (use_tc=0) kernel.
 * a_mn/b_mn: 0 = K-major [row][k], 1 = MN-major [k][row]; rows may overlap (ld < row length).  Split-K planes are
 * written M*ldc apart; returns their number or -1 if the shape is not covered. */
int st_debug_gemm(st_handle* h, int use_tc, int a_mn, int b_mn, const float* a_hi, const float* a_lo, long a_ld,
                  const float* b_hi, const float* b_lo, long b_ld, float* C, long ldc, int M, int N, int K, int splits,
                  void* stream);
long st_debug_numel(st_handle* h, const char* name);
/* Host-only (no device work): launch plan of the five front-end contractions (analysis, synthesis, synthesis dgrad, synthesis
 * wgrad, analysis wgrad) for this geometry and batch on a chip of sm_count SMs (<= 0: 148): out[5][4] =
 * {1 = cta_group::2 kernel | 0 = 1-CTA kernel, tile width BN, split-K planes, grid size}. */
int st_debug_gemm_plan(const st_config* cfg, int batch, int sm_count, int* out);

#ifdef __cplusplus
}
#endif
#endif /* SIGNALTRAIN_B200_H */
