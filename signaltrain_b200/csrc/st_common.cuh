// Shared declarations for the signaltrain_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ST_NUM_PARAMS 40
#define ST_NUM_ACTS 30
#define ST_AE_LAYERS 9
#define ST_AE_ROWS 64          // (batch, bin) rows per CTA tile in the autoencoder kernels
#define ST_AE_RS 68            // smem row stride (floats) of a [feature][row] activation plane
#define ST_AE_THREADS 256
#define ST_MAX_TFRAMES 64      // AE kernels keep ceil(T/16) register tiles; T, OT <= 64

// Geometry, device-visible.  Names follow the reference.
// per-device bookkeeping of one-time kernel configuration (cudaFuncSetAttribute applies to the current device only)
constexpr int ST_MAX_DEVICES = 64;
inline int st_current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < ST_MAX_DEVICES ? dev : 0;
}

struct StDims {
    int C;    // chunk
    int N;    // ft size (taps)
    int H;    // hop
    int F;    // kept bins N/2+1
    int Fp;   // bins padded to a multiple of 8: row stride of one (re|im) half of a spectrum row
    int T;    // analysis frames
    int OT;   // output frames
    int L;    // output samples (OT-1)*H - N
    int K;    // knobs
    int R;    // AE rank (64)
    int Cp;   // padded input row  C + 2N
    int Lp;   // padded output-gradient row  L + 2N = (OT-1)*H + N
    // Frame rows are addressed with a UNIFORM stride so that "all frames of all windows" is one 2-D tensor (TMA map with
    // overlapping rows, row stride H): window b starts at b*Sx in the padded-input buffer, Sx = Tp*H >= Cp, and frame
    // (b, t) is row b*Tp + t.  Rows t in [T, Tp) are dummies: computed where cheaper than masking, never consumed, and
    // kept exactly zero in every buffer that feeds a reduction over rows.
    int Tp;   // ceil(Cp / H)
    int OTp;  // ceil(Lp / H)
    int Sx;   // Tp * H   (stride between windows in the padded-input buffer)
    int Sg;   // OTp * H  (same for the padded output-gradient buffer)
};

// One autoencoder's nine Linear layers, raw reference layout W[out][in] row-major, b[out].
struct AeParams {
    const float* W[ST_AE_LAYERS];
    const float* b[ST_AE_LAYERS];
};
struct AeGrads {
    float* W[ST_AE_LAYERS];
    float* b[ST_AE_LAYERS];
};

// Layer geometry computed on the host once (st_api.cu) and passed by value.
struct AeGeom {
    int in[ST_AE_LAYERS];      // IN_l  (layer 5 includes the K knob inputs)
    int out[ST_AE_LAYERS];     // OUT_l
    int inp[ST_AE_LAYERS];     // IN_l rounded up to a multiple of 4   (row stride of W[o][i] in smem)
    int outp[ST_AE_LAYERS];    // OUT_l rounded up to 16*OPW            (row stride of Wt[i][o] in smem)
    int opw[ST_AE_LAYERS];     // outputs per half-warp in the forward mapping (1,2,4)
    int off_wt[ST_AE_LAYERS];  // float offsets inside the packed smem weight block
    int off_w[ST_AE_LAYERS];
    int off_b[ST_AE_LAYERS];
    int wt_floats;             // size of [Wt..., b...] block
    int w_floats;              // size of the W[o][i] block (backward only)
    int flat_off[ST_AE_LAYERS];   // offset of layer l's weight in the flat [W1,b1,W2,b2,...] gradient vector
    int flat_total;            // total floats in that vector
    int opw_T;                 // outputs per half-warp for the T-wide data-gradient of layer 1
};

#define ST_CUDA_OK(call)                                                         \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) return st_fail_cuda(h, e__, #call, __FILE__, __LINE__); \
    } while (0)

struct st_handle;
int st_fail_cuda(st_handle* h, cudaError_t e, const char* what, const char* file, int line);
int st_fail_msg(st_handle* h, const char* fmt, ...);

static inline int st_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// Exact two-term tf32 split: hi = rna_tf32(x) (low 13 mantissa bits zero), lo = rna_tf32(x - hi).
__device__ __forceinline__ void st_split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}
#endif

// ---- launchers implemented in the .cu files (all asynchronous on `s`) ----------------------
// st_frontend.cu
// Every buffer that is a GEMM operand is stored as an exact tf32 pair (hi, lo): x = hi + lo up to 2^-22 relative.
void st_launch_pad_split(const float* src, float* dst_hi, float* dst_lo, int rows, int len, int pad, int dst_stride, float scale,
                         cudaStream_t s);
void st_launch_pack_analysis(const StDims& d, const float* Wr, const float* Wi, float* wcat_hi, float* wcat_lo, cudaStream_t s);
void st_launch_fold_synthesis(const StDims& d, const float* Sr, const float* Si, float* sfold_hi, float* sfold_lo, cudaStream_t s);
void st_launch_overlap_add(const StDims& d, const float* frames_out, const float* x, int B,
                           float* y_hat, float* x_fwdsyn, float* y_half, cudaStream_t s);
void st_launch_finalize_dft_grads(const StDims& d, const float* part_a, const float* part_s, int splits_a,
                                  int splits_s, float* gWr, float* gWi, float* gSr, float* gSi, int which /*1: analysis, 2: synthesis*/,
                                  cudaStream_t s);
void st_launch_unpack_spec(const StDims& d, const float* spec, int B, float* re, float* im, cudaStream_t s);
void st_launch_dct_bias_unpack(const float* tmp, const float* bias, int B, int nf, int Tp, int sz, float* out, cudaStream_t s);
void st_launch_dct_overlap_add(const float* fo, int B, int nf, int sz, int wsz, int hop, int C, float* wave, cudaStream_t s);
void st_launch_pack_ri(const StDims& d, const float* re, const float* im, int B, float* ri_hi, float* ri_lo, cudaStream_t s);
void st_launch_init_frontend(const StDims& d, float* Wr, float* Wi, float* Sr, float* Si, float* scratch, cudaStream_t s);

// st_gemm_simt.cu  C[M,N] (+split partials) = op(A) * op(B)
struct GemmOperand {
    const float* ptr;   // hi part (or the plain value when lo == nullptr)
    const float* lo;    // lo part, added on load
    long ld;            // leading dimension in floats; rows may overlap (frames: ld = hop)
};
// returns the number of split-K planes actually written (<= splits)
int st_launch_gemm(bool a_kcontig, bool b_kcontig, const GemmOperand& A, const GemmOperand& B, float* C, long ldc,
                   int M, int N, int K, int splits, long split_stride, cudaStream_t s);

// st_gemm_tc.cu  tcgen05 / TMA path.  Operand = exact (hi, lo) tf32 pair, 2-D row-major view (rows may overlap).
struct TcOperand {
    const float* hi;
    const float* lo;
    long rows, cols, ld;
};
int st_tc_pick_bn(int n);
// host-only launch plan of a shape: out = {pair kernel?, BN, split-K planes, grid}; -1 when not covered (no CUDA call)
int st_tc_plan(bool b_mn_major, int M, int N, int K, int splits, int sm_count, int out[4]);
// returns split planes written, or -1 when the shape is not covered (caller falls back to st_launch_gemm)
// promote: start a fresh TMEM accumulator every k-block and sum the partials in fp32 registers (forward GEMMs)
// passes: 3 = exact (hi, lo) operands, 3xTF32 (fp32 fidelity);  1 = hi planes only, single-pass TF32 (reduced precision mode)
int st_launch_gemm_tc(bool a_mn_major, bool b_mn_major, const TcOperand& A, const TcOperand& B, float* C, long ldc, int M,
                      int N, int K, int splits, long split_stride, bool promote, int sm_count, cudaStream_t s, int passes = 3);

// st_ae.cu
size_t st_ae_fwd_smem(const StDims& d, const AeGeom& g);
size_t st_ae_bwd_smem(const StDims& d, const AeGeom& g);
int st_ae_configure(st_handle* h, const StDims& d, const AeGeom& g);
void st_launch_ae_forward(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                          const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri_hi, float* ri_lo,
                          float* const* acts_dev, int grid, cudaStream_t s);
void st_launch_ae_backward(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                           const float* knobs, int B, const float* mag_hat, const float* phs_hat, const float* g_ri,
                           const float* g_mag_hat, const float* g_mag, float* g_spec_hi, float* g_spec_lo, float* partials,
                           int grid, cudaStream_t s);
void st_launch_ae_grad_reduce(const AeGeom& g, const float* partials, int ncta, const AeGrads& gm, const AeGrads& gp,
                              cudaStream_t s);

// st_ae_mma.cu (tensor-core autoencoders; return false if the geometry is not covered)
int st_ae_mma_record_floats(const StDims& d);      // floats per row of the saved-activation record
bool st_launch_ae_forward_mma(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                              const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri_hi, float* ri_lo,
                              float* save_m, float* save_p, int sm_count, cudaStream_t s, int passes = 3);

int st_launch_ae_backward_mma(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, int B,
                              const float* save_m, const float* save_p, const float* mag_hat, const float* phs_hat,
                              const float* g_ri, const float* g_mag_hat, const float* g_mag, float* tail_ws, float* g_spec_hi,
                              float* g_spec_lo, float* partials, long long* timing /*nullable: 16 counters*/, int sm_count, cudaStream_t s,
                              int passes = 3);

// st_ae_tm.cu: tcgen05 autoencoders with TMEM-resident activations (production path; T <= 64, OT <= 16, K <= 8).
// Forward: both autoencoders of a 128-row tile in one CTA, no saved activations.  Returns false when the geometry is not covered.
// wpack: workspace of st_ae_tm_pack_floats() floats holding the shared-memory image of the weights; rebuilt on s_pack when pack is set.
// dbg (nullable, tests only): [2][9][128][64] layer outputs of tile 0.
long st_ae_tm_pack_floats();
bool st_launch_ae_forward_tm(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                             const float* knobs, int B, float* mag, float* trk /*nullable: [2][B][T][F] tracks for the backward*/,
                             float* mag_hat, float* phs_hat, float* ri, float* ri_lo,
                             float* wpack, float* dbg, long long* timing /*nullable: 256 counters*/, int sm_count,
                             bool pack, cudaStream_t s_pack, cudaStream_t s);

long st_ae_tm_bwd_pack_floats();
// Backward with in-kernel recompute.  Returns the number of per-CTA partial-gradient vectors per autoencoder (0: not covered).
int st_launch_ae_backward_tm(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                             const float* trk /*the forward kernel's tracks*/, const float* knobs, int B, const float* mag_hat, const float* phs_hat, const float* g_ri,
                             const float* g_mag_hat, const float* g_mag, float* g_track, float* g_spec, float* g_spec_lo,
                             float* partials, float* wpack, float* dbg, long long* timing /*nullable: 64 counters*/, int sm_count, bool pack,
                             cudaStream_t s_pack, cudaStream_t s);

// st_data.cu
void st_launch_compressor_4c(const float* x, const double* knobs_wc, int B, int n, double sr, float* scratch, float* y, cudaStream_t s);
void st_launch_crop_windows(const float* cx, const float* cy, const long* off, const float* sign, int B, int C, int L, float* ox,
                            float* oy, cudaStream_t s);

// st_loss_opt.cu
void st_launch_loss(const StDims& d, const float* y_hat, const float* y, const float* mag_hat, const float* sbf,
                    float l1_coef, int B, float* loss, float* g_y_hat, float* g_mag_hat, float* scratch,
                    unsigned* counter, cudaStream_t s);
// fused forward tail of st_train_step: overlap-add + residual + loss + both loss gradients + padded (hi, lo) 2*dL/dy_hat
void st_launch_ola_loss(const StDims& d, const float* fo, const float* x, const float* y, const float* mag_hat, const float* sbf,
                        float l1_coef, int B, float* loss, float* gwave_hi, float* gwave_lo, float* g_mag_hat, float* scratch,
                        unsigned* counter, cudaStream_t s);
// fused split-K sum + un-fold of the DFT gradients and their L1 norm / clip coefficient (st_train_step)
void st_launch_finalize_norm(const StDims& d, const float* pa, const float* ps, int sa, int ss, float* gWr, float* gWi, float* gSr,
                             float* gSi, float grad_scale, float max_norm, float* norm_out, float* coef_out, float* scratch,
                             unsigned* counter, cudaStream_t s);
void st_launch_mae(const float* a, const float* b, long n, float* out, float* scratch, unsigned* counter, cudaStream_t s);
void st_launch_l1_norm4(const float* const g[4], long n_each, long live_rows_a, long row_len, float grad_scale,
                        float max_norm, float* norm_out, float* coef_out, float* scratch, unsigned* counter, cudaStream_t s);
void st_launch_scale4(float* const g[4], long n_each, const float* coef, cudaStream_t s);

// data-parallel exchange payload (st_pack_grads / st_unpack_grads)
struct GradPack {
    float* g[ST_NUM_PARAMS];
    float* packed;
    long live;                   // F * N
    int N;
    int ae_off[ST_NUM_PARAMS - 4];
    int ae_n[ST_NUM_PARAMS - 4];
};
void st_launch_pack_grads(const GradPack& gp, int dir, cudaStream_t s);
void st_launch_unpack_clip(const GradPack& gp, float grad_scale, float max_norm, float* norm_out, float* coef_out, float* scratch,
                           unsigned* counter, cudaStream_t s);
void st_launch_finalize_packed(const StDims& d, const float* pa, const float* ps, int sa, int ss, float* packed, cudaStream_t s);

struct AdamTensors {
    float* p[ST_NUM_PARAMS];
    const float* g[ST_NUM_PARAMS];
    float* m[ST_NUM_PARAMS];
    float* v[ST_NUM_PARAMS];
    long n[ST_NUM_PARAMS];       // live elements (DFT analysis tensors: F*N, rows >= F never change)
};
struct AdamScalars {
    float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps, grad_scale;
};
void st_launch_adam(const AdamTensors& t, const int2* chunk_map, int nchunks, const AdamScalars& sc,
                    const float* clip_coef, cudaStream_t s);
const void* st_adam_kernel_fn();
const void* st_ola_loss_kernel_fn();     // kernel addresses: a captured step finds the nodes whose inputs change per step
const void* st_pad_split_kernel_fn();
#define ST_ADAM_CHUNK 4096
