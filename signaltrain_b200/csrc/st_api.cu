// C ABI of signaltrain_b200 (include/signaltrain_b200.h): handle, workspace, and the kernel sequences of
// forward / loss / backward / clip / Adam.  Host code only orchestrates; all arithmetic is in the kernels.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <utility>
#include <vector>

#include "../../include/signaltrain_b200.h"
#include "st_common.cuh"

struct st_handle {
    StDims d;
    AeGeom g;
    int device = 0;
    int sm_count = 148;
    int maxB = 0;
    int fwdB = 0;                 // batch of the last st_forward (st_backward must match)
    int bwd_ss = -1;              // split-K planes of the synthesis weight gradient written by st_backward_begin
    int ae_grid = 0;
    // workspace (device)
    // GEMM operands are kept as exact tf32 (hi, lo) pairs: *_lo is the residual of the buffer of the same name
    float *xpad = nullptr, *xpad_lo = nullptr, *wcat = nullptr, *wcat_lo = nullptr, *sfold = nullptr, *sfold_lo = nullptr;
    float *spec = nullptr, *ri = nullptr, *ri_lo = nullptr, *fo = nullptr;
    float *mag_hat_ws = nullptr, *phs_hat_ws = nullptr, *gwave = nullptr, *gwave_lo = nullptr, *g_ri = nullptr;
    float *g_spec = nullptr, *g_spec_lo = nullptr;
    int passes = 3;               // st_set_precision: 3 = fp32 fidelity (3xTF32 GEMMs, exact-fp32 FFMA2 autoencoders);
                                  // 1 = reduced precision (single-pass TF32 products in the GEMMs and the autoencoder chains)
    bool fuse_tail = true;        // st_train_step: fused overlap-add + loss + padded gradient kernel and fused finalize + L1 norm
                                  // (ST_DISABLE_FUSED_TAIL=1 -> the separate kernels the piecewise entry points use)
    bool use_tc = true;           // tcgen05/TMA GEMMs (falls back to the FFMA GEMM per call when a shape is not covered)
    float *part_a = nullptr, *part_s = nullptr, *ae_part = nullptr;
    float *yhat_ws = nullptr, *gy_ws = nullptr, *gmh_ws = nullptr;   // fused train step only
    float* knobs_ws = nullptr;    // copy of the forward's knobs (the SIMT backward recomputes the AE chain)
    float *ae_save_m = nullptr, *ae_save_p = nullptr;   // per-row activation records written by the tensor-core forward
    // side stream: independent small kernels run beside the main stream's work (fork / join with events; off while profiling
    // stages, whose event pairs live on the main stream)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pack = nullptr;
    float* dct_ws = nullptr;      // workspace of the DCT / MDCT front-end variant (st_dct_analysis / st_dct_synthesis)
    long dct_ws_floats = 0;
    float* trk_ws = nullptr;      // magnitude | phase tracks written by the TMEM forward kernel, read by the TMEM backward kernel
    float* gtrack_ws = nullptr;   // track gradients of the two autoencoders (FFMA2 backward -> ae_input_grad_kernel)
    float* tail_ws = nullptr;     // skip / residual gradient scratch of the tensor-core backward
    bool training = true;         // st_set_training: save activations in st_forward for a following st_backward
    bool have_saves = false;      // the last forward wrote ae_save_*
    bool use_mma_bwd = true;
    bool use_tm = true;           // tcgen05 autoencoders with TMEM-resident activations (st_ae_tm.cu); ST_DISABLE_TMEM_AE=1 -> the
                                  // mma.sync kernels with saved activation records (the path geometries with T > 32 train on)
    bool tm_fwd = false;          // the last forward ran the TMEM kernel (its backward recomputes: nothing was saved)
    bool tm_bwd_image = false;    // ae_wpack_bwd holds the backward weight image of the current parameters
    float *ae_wpack = nullptr, *ae_wpack_bwd = nullptr;   // shared-memory images of the autoencoder weights (st_ae_tm.cu)
    long long* ae_timing = nullptr;   // device: 16 region counters of the tensor-core AE backward (st_debug_ae_timing)
    float* small = nullptr;       // reduction scratch + scalar outputs
    unsigned* counters = nullptr;
    double* win = nullptr;        // [hamming | GLA] in double, for st_init_frontend
    int2 *map_live = nullptr, *map_full = nullptr;
    int n_live = 0, n_full = 0;
    long numel[ST_NUM_PARAMS];
    std::string names[ST_NUM_PARAMS];
    // per-stage profiling (st_profile / st_profile_read): CUDA events on the launching stream
    bool prof_on = false;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_events;
    long launches = 0;            // kernels + device copies launched by this handle since creation
    // st_train_step captured into a CUDA graph per (buffers, batch, stream): see StepGraph below
    bool use_graph = true;
    std::vector<struct StepGraph*> graphs;
    long graph_replays = 0;
    long simt_fallbacks = 0;      // calls served by a SIMT fallback kernel (GEMM shape or autoencoder geometry not covered by the
                                  // tensor-core kernels; return_acts forwards count too): st_debug_fallbacks
    char err[1024];
};

enum Stage { SG_PAD_X, SG_PACK_W, SG_GEMM_ANALYSIS, SG_AE_FWD, SG_GEMM_SYNTH, SG_OLA, SG_LOSS, SG_PAD_G, SG_GEMM_SYNTH_DGRAD,
             SG_GEMM_SYNTH_WGRAD, SG_AE_BWD, SG_AE_REDUCE, SG_GEMM_ANALYSIS_WGRAD, SG_FINALIZE, SG_L1NORM, SG_ADAM, SG_COUNT };
static const char* kStageNames[SG_COUNT] = {"pad_x", "pack_weights", "gemm_analysis", "ae_forward", "gemm_synthesis",
    "overlap_add", "loss", "pad_grad", "gemm_synthesis_dgrad", "gemm_synthesis_wgrad", "ae_backward", "ae_grad_reduce",
    "gemm_analysis_wgrad", "finalize_dft_grads", "l1_norm", "adam"};

// RAII scope around one stage: counts its launches and, when profiling is on, brackets it with events.
struct StageScope {
    st_handle* h; int stage; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr;
    StageScope(st_handle* h_, int stage_, int nlaunch, cudaStream_t s_) : h(h_), stage(stage_), s(s_) {
        h->launches += nlaunch;
        if (h->prof_on) {
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a, s);
        }
    }
    ~StageScope() {
        if (a) {
            cudaEventRecord(b, s);
            h->prof_events.push_back({stage, {a, b}});
        }
    }
};

static thread_local char g_create_err[1024] = "";

// Every entry point runs on the handle's device and puts the caller's current device back on return (a process may hold
// engines on several GPUs, and torch allocates on the CURRENT device).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        else prev = -1;                                  // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ST_ON_DEVICE(h)                                                                   \
    DeviceGuard guard__((h)->device);                                                     \
    if (!guard__.ok) return st_fail_msg(h, "cudaSetDevice(%d) failed", (h)->device)

// offsets inside h->small (floats)
enum { SM_LOSS = 0, SM_NORM = 2400, SM_MAE = 3008, SM_TOTAL_NORM = 4200, SM_COEF = 4201, SM_FLOATS = 4352 };
enum { CT_LOSS = 0, CT_NORM = 1, CT_MAE = 2, CT_COUNT = 4 };
static const int kMaxSplits = 8;

int st_fail_msg(st_handle* h, const char* fmt, ...) {
    char* dst = h ? h->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 1024, fmt, ap);
    va_end(ap);
    return 1;
}
int st_fail_cuda(st_handle* h, cudaError_t e, const char* what, const char* file, int line) {
    return st_fail_msg(h, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
}

#define ST_LAUNCH_OK(h)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) return st_fail_cuda(h, e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

static int opw_for(int n) { return n <= 16 ? 1 : (n <= 32 ? 2 : 4); }
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static void build_geom(const StDims& d, AeGeom& g) {
    const int in[ST_AE_LAYERS] = {d.T, 64, 32, 16, 16 + d.K, 16, 16, 32, 64};
    const int out[ST_AE_LAYERS] = {64, 32, 16, 16, 16, 16, 32, 64, d.OT};
    int owt = 0, ow = 0, flat = 0;
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        g.in[l] = in[l];
        g.out[l] = out[l];
        g.opw[l] = opw_for(out[l]);
        g.outp[l] = 16 * g.opw[l];
        const int back_out = (l == 4) ? 16 : in[l];       // data-gradient width (knobs carry none)
        g.inp[l] = std::max(round_up(in[l], 4), 16 * opw_for(back_out));
        g.off_wt[l] = owt;
        owt += in[l] * g.outp[l];
        g.off_w[l] = ow;
        ow += out[l] * g.inp[l];
        g.flat_off[l] = flat;
        flat += out[l] * in[l] + out[l];
    }
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        g.off_b[l] = owt;
        owt += g.outp[l];
    }
    g.wt_floats = round_up(owt, 4);
    g.w_floats = round_up(ow, 4);
    g.flat_total = flat;
    g.opw_T = opw_for(d.T);
}

static std::vector<double> hamming_sym(int n) {
    std::vector<double> w(n);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < n; ++i) w[i] = 0.54 - 0.46 * std::cos(2.0 * pi * i / (n - 1));
    return w;
}

extern "C" int st_abi_version(void) { return ST_ABI_VERSION; }

extern "C" int st_create(const st_config* cfg, int device, st_handle** out) {
    if (!cfg || !out) return st_fail_msg(nullptr, "st_create: null argument");
    *out = nullptr;
    const int C = cfg->chunk, N = cfg->ft, H = cfg->hop, T = cfg->frames_in, OT = cfg->frames_out, K = cfg->knobs;
    if (cfg->rank != 64) return st_fail_msg(nullptr, "st_create: decomposition rank %d unsupported (reference hard-codes 64)", cfg->rank);
    if (N < 64 || (N & 7) || (H & 3) || (C & 3) || H <= 0 || C <= 0)
        return st_fail_msg(nullptr, "st_create: need ft %% 8 == 0, hop %% 4 == 0, chunk %% 4 == 0 (got ft=%d hop=%d chunk=%d)", N, H, C);
    if (K < 0 || K > 16) return st_fail_msg(nullptr, "st_create: knobs=%d outside [0,16]", K);
    if ((C + N) / H + 1 != T)
        return st_fail_msg(nullptr, "st_create: conv output frames %d != expected_time_frames %d (the reference would fail at nn_proc.py:81)",
                           (C + N) / H + 1, T);
    if (OT > T || OT < 2 || T > ST_MAX_TFRAMES)
        return st_fail_msg(nullptr, "st_create: need 2 <= frames_out <= frames_in <= %d (got %d, %d)", ST_MAX_TFRAMES, OT, T);
    const int L = (OT - 1) * H - N;
    if (L <= 0 || L > C || (L & 3)) return st_fail_msg(nullptr, "st_create: invalid output size L=%d", L);

    DeviceGuard guard(device);
    if (!guard.ok) return st_fail_msg(nullptr, "st_create: cudaSetDevice(%d) failed", device);
    cudaError_t e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return st_fail_cuda(nullptr, e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    if (prop.major != 10)
        return st_fail_msg(nullptr, "st_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);

    st_handle* h = new st_handle();
    h->err[0] = 0;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    StDims& d = h->d;
    d.C = C; d.N = N; d.H = H; d.F = N / 2 + 1; d.T = T; d.OT = OT; d.L = L; d.K = K; d.R = 64;
    // 2*Fp is a GEMM N (analysis, synthesis dgrad) and K (synthesis) extent: a multiple of 32 with a divisor in [128, 256]
    // that is a multiple of 16, so the tcgen05 tiles cover it exactly
    d.Fp = round_up(d.F, 16);
    while (st_tc_pick_bn(2 * d.Fp) < 128) d.Fp += 16;
    d.Cp = C + 2 * N; d.Lp = L + 2 * N;
    d.Tp = (d.Cp + H - 1) / H; d.OTp = (d.Lp + H - 1) / H;
    d.Sx = d.Tp * H; d.Sg = d.OTp * H;
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        h->side = nullptr;                       // no overlap: everything stays on the caller's stream
    }
    if (const char* e = getenv("ST_DISABLE_TCGEN05")) h->use_tc = !(e[0] == '1');
    if (const char* e = getenv("ST_DISABLE_MMA_BACKWARD")) h->use_mma_bwd = !(e[0] == '1');
    if (const char* e = getenv("ST_DISABLE_TMEM_AE")) h->use_tm = !(e[0] == '1');
    if (const char* e = getenv("ST_CUDA_GRAPH")) h->use_graph = !(e[0] == '0');
    if (const char* e = getenv("ST_DISABLE_FUSED_TAIL")) h->fuse_tail = !(e[0] == '1');
    build_geom(d, h->g);
    if (st_ae_configure(h, d, h->g)) {
        snprintf(g_create_err, sizeof(g_create_err), "%s", h->err);
        delete h;
        return 1;
    }
    // parameter table
    const char* dft[4] = {"mpaec.dft_analysis.conv_analysis_real.weight", "mpaec.dft_analysis.conv_analysis_imag.weight",
                          "mpaec.dft_synthesis.conv_synthesis_real.weight", "mpaec.dft_synthesis.conv_synthesis_imag.weight"};
    const char* layers[ST_AE_LAYERS] = {"fnn_enc", "fnn_enc2", "fnn_enc3", "fnn_enc4", "fnn_addknobs",
                                        "fnn_dec4", "fnn_dec3", "fnn_dec2", "fnn_dec"};
    for (int i = 0; i < 4; ++i) { h->names[i] = dft[i]; h->numel[i] = (long)N * N; }
    for (int a = 0; a < 2; ++a)
        for (int l = 0; l < ST_AE_LAYERS; ++l) {
            const int i = 4 + a * 18 + 2 * l;
            h->names[i] = std::string("mpaec.") + (a ? "phs_aenc." : "aenc.") + layers[l] + ".weight";
            h->names[i + 1] = std::string("mpaec.") + (a ? "phs_aenc." : "aenc.") + layers[l] + ".bias";
            h->numel[i] = (long)h->g.out[l] * h->g.in[l];
            h->numel[i + 1] = h->g.out[l];
        }
    // persistent small buffers
    auto fail = [&](cudaError_t ce, const char* what) {
        st_fail_cuda(nullptr, ce, what, __FILE__, __LINE__);
        st_destroy(h);
        return 1;
    };
    if ((e = cudaMalloc(&h->small, SM_FLOATS * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(small)");
    if ((e = cudaMalloc(&h->counters, CT_COUNT * sizeof(unsigned))) != cudaSuccess) return fail(e, "cudaMalloc(counters)");
    if ((e = cudaMemset(h->counters, 0, CT_COUNT * sizeof(unsigned))) != cudaSuccess) return fail(e, "cudaMemset(counters)");
    if ((e = cudaMalloc(&h->wcat, 2L * d.Fp * N * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(wcat)");
    if ((e = cudaMalloc(&h->sfold, 2L * d.Fp * N * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(sfold)");
    if ((e = cudaMalloc(&h->wcat_lo, 2L * d.Fp * N * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(wcat_lo)");
    if ((e = cudaMalloc(&h->sfold_lo, 2L * d.Fp * N * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(sfold_lo)");
    if ((e = cudaMalloc(&h->part_a, (long)kMaxSplits * 2 * d.Fp * N * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(part_a)");
    if ((e = cudaMalloc(&h->part_s, (long)kMaxSplits * 2 * d.Fp * N * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(part_s)");
    if ((e = cudaMalloc(&h->ae_wpack, st_ae_tm_pack_floats() * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(ae_wpack)");
    if ((e = cudaMalloc(&h->ae_wpack_bwd, st_ae_tm_bwd_pack_floats() * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(ae_wpack_bwd)");
    // windows for st_init_frontend: hamming (cls_fe_dft.py:38) and the Griffin-Lim LSEE window (:133-163)
    {
        std::vector<double> w = hamming_sym(N), win(2 * N), env(N, 0.0);
        const int red = N / H;
        for (int k = -red; k <= red; ++k) {
            const int lo = std::max(0, H * k), hi = std::min(N, N + H * k);
            for (int j = lo; j < hi; ++j) env[j] += w[j - H * k] * w[j - H * k];
        }
        for (int i = 0; i < N; ++i) { win[i] = w[i]; win[N + i] = w[i] / env[i]; }
        if ((e = cudaMalloc(&h->win, 2 * N * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc(win)");
        if ((e = cudaMemcpy(h->win, win.data(), 2 * N * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy(win)");
    }
    // Adam chunk maps: live-only (analysis rows >= F never receive gradient) and full
    {
        std::vector<int2> live, full;
        for (int i = 0; i < ST_NUM_PARAMS; ++i) {
            const long nf = h->numel[i], nl = (i < 2) ? (long)d.F * N : nf;
            for (long c = 0; c * ST_ADAM_CHUNK < nf; ++c) full.push_back(make_int2(i, (int)c));
            for (long c = 0; c * ST_ADAM_CHUNK < nl; ++c) live.push_back(make_int2(i, (int)c));
        }
        h->n_live = (int)live.size();
        h->n_full = (int)full.size();
        if ((e = cudaMalloc(&h->map_live, live.size() * sizeof(int2))) != cudaSuccess) return fail(e, "cudaMalloc(map_live)");
        if ((e = cudaMalloc(&h->map_full, full.size() * sizeof(int2))) != cudaSuccess) return fail(e, "cudaMalloc(map_full)");
        cudaMemcpy(h->map_live, live.data(), live.size() * sizeof(int2), cudaMemcpyHostToDevice);
        cudaMemcpy(h->map_full, full.data(), full.size() * sizeof(int2), cudaMemcpyHostToDevice);
    }
    *out = h;
    return 0;
}

static void free_batch_buffers(st_handle* h) {
    float** bufs[] = {&h->xpad, &h->xpad_lo, &h->spec, &h->ri, &h->ri_lo, &h->fo, &h->mag_hat_ws, &h->phs_hat_ws, &h->gwave,
                      &h->gwave_lo, &h->g_ri, &h->g_spec, &h->g_spec_lo, &h->ae_part, &h->yhat_ws, &h->gy_ws, &h->gmh_ws,
                      &h->knobs_ws, &h->ae_save_m, &h->ae_save_p, &h->tail_ws, &h->gtrack_ws, &h->trk_ws};
    for (float** b : bufs) {
        if (*b) cudaFree(*b);
        *b = nullptr;
    }
    h->maxB = 0;
}

static void drop_step_graphs(st_handle* h);

extern "C" void st_destroy(st_handle* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    drop_step_graphs(h);
    free_batch_buffers(h);
    if (h->dct_ws) cudaFree(h->dct_ws);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_pack) cudaEventDestroy(h->ev_pack);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->small) cudaFree(h->small);
    if (h->counters) cudaFree(h->counters);
    if (h->wcat) cudaFree(h->wcat);
    if (h->sfold) cudaFree(h->sfold);
    if (h->wcat_lo) cudaFree(h->wcat_lo);
    if (h->sfold_lo) cudaFree(h->sfold_lo);
    if (h->part_a) cudaFree(h->part_a);
    if (h->part_s) cudaFree(h->part_s);
    if (h->ae_wpack) cudaFree(h->ae_wpack);
    if (h->ae_wpack_bwd) cudaFree(h->ae_wpack_bwd);
    if (h->win) cudaFree(h->win);
    if (h->map_live) cudaFree(h->map_live);
    if (h->map_full) cudaFree(h->map_full);
    delete h;
}

extern "C" const char* st_last_error(const st_handle* h) { return h ? h->err : g_create_err; }
extern "C" const char* st_param_name(const st_handle* h, int idx) {
    return (h && idx >= 0 && idx < ST_NUM_PARAMS) ? h->names[idx].c_str() : "";
}
extern "C" long st_param_numel(const st_handle* h, int idx) { return (h && idx >= 0 && idx < ST_NUM_PARAMS) ? h->numel[idx] : -1; }
extern "C" int st_out_samples(const st_handle* h) { return h ? h->d.L : -1; }
extern "C" int st_bins(const st_handle* h) { return h ? h->d.F : -1; }

// (Re)allocate the batch-sized workspace.  Grows only; a steady-state training loop never allocates.
static void drop_step_graphs(st_handle* h);
static int ensure_workspace(st_handle* h, int B) {
    if (B <= 0) return st_fail_msg(h, "batch must be positive (got %d)", B);
    ST_CUDA_OK(cudaSetDevice(h->device));
    if (B <= h->maxB) return 0;
    ST_CUDA_OK(cudaDeviceSynchronize());
    drop_step_graphs(h);                                       // captured steps hold the addresses of the buffers freed below
    free_batch_buffers(h);
    const StDims& d = h->d;
    const long BT = (long)B * d.Tp, BO = (long)B * d.OTp;      // frame rows incl. the dummy rows of the uniform-stride view
    const long tiles = ((long)B * d.F + ST_AE_ROWS - 1) / ST_AE_ROWS;
    h->ae_grid = (int)std::min<long>(tiles, h->sm_count);
    struct { float** p; long n; bool zero; } req[] = {
        {&h->xpad, (long)B * d.Sx + d.N, true},    {&h->xpad_lo, (long)B * d.Sx + d.N, true},
        {&h->spec, BT * 2 * d.Fp, true},           {&h->ri, BO * 2 * d.Fp, true},           {&h->ri_lo, BO * 2 * d.Fp, true},
        {&h->fo, BO * d.N, false},                 {&h->mag_hat_ws, (long)B * d.OT * d.F, false},
        {&h->phs_hat_ws, (long)B * d.OT * d.F, false},
        {&h->gwave, (long)B * d.Sg + d.N, true},   {&h->gwave_lo, (long)B * d.Sg + d.N, true},
        {&h->g_ri, BO * 2 * d.Fp, true},           {&h->g_spec, BT * 2 * d.Fp, true},       {&h->g_spec_lo, BT * 2 * d.Fp, true},
        {&h->ae_part, (long)h->sm_count * 2 * h->g.flat_total, true},
        {&h->yhat_ws, (long)B * d.L, false},     {&h->gy_ws, (long)B * d.L, false}, {&h->gmh_ws, (long)B * d.OT * d.F, false},
        {&h->knobs_ws, (long)B * std::max(d.K, 1), false},
        {&h->ae_save_m, (long)B * d.F * st_ae_mma_record_floats(d), false},
        {&h->ae_save_p, (long)B * d.F * st_ae_mma_record_floats(d), false},
        {&h->tail_ws, (long)B * d.OT * d.F, false},
        {&h->gtrack_ws, 2L * B * d.T * d.F, false}, {&h->trk_ws, 2L * B * d.T * d.F, false},
    };
    for (auto& r : req) {
        ST_CUDA_OK(cudaMalloc(r.p, r.n * sizeof(float)));
        if (r.zero) ST_CUDA_OK(cudaMemset(*r.p, 0, r.n * sizeof(float)));   // padding columns must stay exactly zero
    }
    ST_CUDA_OK(cudaDeviceSynchronize());
    h->maxB = B;
    return 0;
}

static void split_params(const float* const* params, AeParams& pm, AeParams& pp) {
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        pm.W[l] = params[4 + 2 * l];  pm.b[l] = params[5 + 2 * l];
        pp.W[l] = params[22 + 2 * l]; pp.b[l] = params[23 + 2 * l];
    }
}

static int check_ptrs(st_handle* h, const void* const* p, int n, const char* what) {
    if (!p) return st_fail_msg(h, "%s: null pointer table", what);
    for (int i = 0; i < n; ++i)
        if (!p[i]) return st_fail_msg(h, "%s: entry %d is null", what, i);
    return 0;
}

extern "C" int st_init_frontend(st_handle* h, float* const* params, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (check_ptrs(h, (const void* const*)params, 4, "st_init_frontend(params)")) return 1;
    ST_CUDA_OK(cudaSetDevice(h->device));
    st_launch_init_frontend(h->d, params[0], params[1], params[2], params[3], reinterpret_cast<float*>(h->win), (cudaStream_t)stream);
    ST_LAUNCH_OK(h);
    return 0;
}

// Analysis.forward (cls_fe_dft.py:50-58) alone: x (B, C) -> an_real, an_imag (B, T, F).  `scale` = 1 here (the model
// halves its input, nn_proc.py:307; a bare Analysis does not).
static int analysis_only(st_handle* h, const float* x, const float* Wr, const float* Wi, int B, float* re, float* im, cudaStream_t s) {
    const StDims& d = h->d;
    if (ensure_workspace(h, B)) return 1;
    st_launch_pad_split(x, h->xpad, h->xpad_lo, B, d.C, d.N, d.Sx, 1.0f, s);
    st_launch_pack_analysis(d, Wr, Wi, h->wcat, h->wcat_lo, s);
    const int MT = B * d.Tp, F2 = 2 * d.Fp;
    int r = -1;
    if (h->use_tc) {
        TcOperand A{h->xpad, h->xpad_lo, MT, d.N, d.H}, W{h->wcat, h->wcat_lo, F2, d.N, d.N};
        r = st_launch_gemm_tc(false, false, A, W, h->spec, F2, MT, F2, d.N, 1, 0, h->passes == 3, h->sm_count, s, h->passes);
    }
    if (r < 0) {
        GemmOperand A{h->xpad, h->xpad_lo, d.H}, W{h->wcat, h->wcat_lo, d.N};
        ++h->simt_fallbacks;
        st_launch_gemm(true, true, A, W, h->spec, F2, MT, F2, d.N, 1, 0, s);
    }
    st_launch_unpack_spec(d, h->spec, B, re, im, s);
    h->launches += 4;
    ST_LAUNCH_OK(h);
    h->fwdB = 0;      // the workspace no longer holds a model forward
    return 0;
}

// Synthesis.forward (cls_fe_dft.py:102-115) alone: real, imag (B, OT, F) -> wave (B, L).
static int synthesis_only(st_handle* h, const float* re, const float* im, const float* Sr, const float* Si, int B, float* wave,
                          cudaStream_t s) {
    const StDims& d = h->d;
    if (ensure_workspace(h, B)) return 1;
    st_launch_pack_ri(d, re, im, B, h->ri, h->ri_lo, s);
    st_launch_fold_synthesis(d, Sr, Si, h->sfold, h->sfold_lo, s);
    const int MO = B * d.OTp, F2 = 2 * d.Fp;
    int r = -1;
    if (h->use_tc) {
        TcOperand R{h->ri, h->ri_lo, MO, F2, F2}, S{h->sfold, h->sfold_lo, F2, d.N, d.N};
        r = st_launch_gemm_tc(false, true, R, S, h->fo, d.N, MO, d.N, F2, 1, 0, h->passes == 3, h->sm_count, s, h->passes);
    }
    if (r < 0) {
        GemmOperand R{h->ri, h->ri_lo, F2}, S{h->sfold, h->sfold_lo, d.N};
        ++h->simt_fallbacks;
        st_launch_gemm(true, false, R, S, h->fo, d.N, MO, d.N, F2, 1, 0, s);
    }
    st_launch_overlap_add(d, h->fo, nullptr, B, wave, nullptr, nullptr, s);
    h->launches += 4;
    ST_LAUNCH_OK(h);
    h->fwdB = 0;
    return 0;
}

static int forward_impl(st_handle* h, const float* x, const float* knobs, int B, const float* const* params, float* y_hat,
                        float* mag, float* mag_hat_user, float* const* acts, cudaStream_t s) {
    const StDims& d = h->d;
    if (ensure_workspace(h, B)) return 1;
    // packing the weights does not depend on the batch: it runs on the side stream, beside the input padding
    const bool beside = h->side && !h->prof_on;
    AeParams pm, pp;
    split_params(params, pm, pp);
    // Autoencoders: the tcgen05 / TMEM kernels (st_ae_tm.cu) wherever they cover the geometry.  Their backward recomputes
    // the chain, so nothing is saved; it covers T <= 32, which is why a TRAINING forward with 32 < T <= 64 still takes the
    // mma.sync kernels with saved records (st_ae_mma.cu).  return_acts is served by the SIMT kernel.
    const bool tm_geom = h->use_tm && !acts && d.T <= 64 && d.OT <= 16 && d.K <= 8;
    const bool tm_bwd_geom = tm_geom && d.T <= 32;
    const bool tm_fwd = tm_geom && (tm_bwd_geom || !h->training);
    h->tm_bwd_image = false;
    {
        cudaStream_t sp = beside ? h->side : s;
        StageScope sc(h, SG_PACK_W, 2 + tm_fwd + (tm_fwd && tm_bwd_geom && h->training), s);
        if (beside) {
            ST_CUDA_OK(cudaEventRecord(h->ev_fork, s));          // behind whatever wrote the weights (the previous step's Adam)
            ST_CUDA_OK(cudaStreamWaitEvent(sp, h->ev_fork, 0));
        }
        st_launch_pack_analysis(d, params[0], params[1], h->wcat, h->wcat_lo, sp);
        if (beside) ST_CUDA_OK(cudaEventRecord(h->ev_pack, sp));     // all the analysis GEMM needs; the rest hides behind that GEMM
        st_launch_fold_synthesis(d, params[2], params[3], h->sfold, h->sfold_lo, sp);
        if (tm_fwd)       // shared-memory images of the autoencoder weights (B = 0: pack only)
            st_launch_ae_forward_tm(d, h->g, pm, pp, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, h->ae_wpack,
                                    nullptr, nullptr, h->sm_count, true, sp, sp);
        if (tm_fwd && tm_bwd_geom && h->training) {
            st_launch_ae_backward_tm(d, h->g, pm, pp, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, nullptr, h->ae_wpack_bwd, nullptr, nullptr, h->sm_count, true, sp, sp);
            h->tm_bwd_image = true;
        }
        if (beside) ST_CUDA_OK(cudaEventRecord(h->ev_join, sp));
    }
    {
        // the knobs are kept for the backward kernels that recompute the chain (TMEM and SIMT); the record-based mma.sync
        // backward reads them from the saved activations
        const bool keep_knobs = d.K > 0;
        StageScope sc(h, SG_PAD_X, 1 + keep_knobs, s);
        // x/2 with the conv padding, as (hi, lo), window stride Sx = Tp*H (frame (b,t) = row b*Tp+t of a stride-H view)
        st_launch_pad_split(x, h->xpad, h->xpad_lo, B, d.C, d.N, d.Sx, 0.5f, s);
        if (keep_knobs) ST_CUDA_OK(cudaMemcpyAsync(h->knobs_ws, knobs, (long)B * d.K * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (beside) ST_CUDA_OK(cudaStreamWaitEvent(s, h->ev_pack, 0));
    ST_LAUNCH_OK(h);
    const int MT = B * d.Tp, MO = B * d.OTp, F2 = 2 * d.Fp;
    {   // analysis: spec[(b,t), (re|im) k] = sum_n frame[(b,t), n] * wcat[k, n]
        StageScope sc(h, SG_GEMM_ANALYSIS, 1, s);
        int r = -1;
        if (h->use_tc) {
            TcOperand A{h->xpad, h->xpad_lo, MT, d.N, d.H}, W{h->wcat, h->wcat_lo, F2, d.N, d.N};
            r = st_launch_gemm_tc(false, false, A, W, h->spec, F2, MT, F2, d.N, 1, 0, /*promote=*/h->passes == 3, h->sm_count, s, h->passes);
        }
        if (r < 0) {
            GemmOperand A{h->xpad, h->xpad_lo, d.H}, W{h->wcat, h->wcat_lo, d.N};
            ++h->simt_fallbacks;
            st_launch_gemm(true, true, A, W, h->spec, F2, MT, F2, d.N, 1, 0, s);
        }
    }
    ST_LAUNCH_OK(h);
    if (beside) ST_CUDA_OK(cudaStreamWaitEvent(s, h->ev_join, 0));        // folded synthesis weights, autoencoder weight images
    {
        StageScope sc(h, SG_AE_FWD, ((acts || tm_fwd) ? 1 : 2) + (mag_hat_user != nullptr), s);
        // the handle's copy of the knobs (made above), so that no kernel of a captured step holds the caller's pointer
        const float* knobs_in = d.K > 0 ? h->knobs_ws : knobs;
        const bool save = h->training && h->use_mma_bwd;
        h->have_saves = false;
        h->tm_fwd = false;
        bool done = false;
        if (tm_fwd) {
            done = st_launch_ae_forward_tm(d, h->g, pm, pp, h->spec, knobs_in, B, mag, h->trk_ws, h->mag_hat_ws, h->phs_hat_ws, h->ri, h->ri_lo, h->ae_wpack,
                                           nullptr, nullptr, h->sm_count, false, s, s);
            h->tm_fwd = done;
        }
        if (!acts && !done) {
            done = st_launch_ae_forward_mma(d, h->g, pm, pp, h->spec, knobs_in, B, mag, h->mag_hat_ws, h->phs_hat_ws, h->ri, h->ri_lo,
                                            save ? h->ae_save_m : nullptr, save ? h->ae_save_p : nullptr, h->sm_count, s, h->passes);
            if (done) h->have_saves = save;
        }
        if (!done) {
            ++h->simt_fallbacks;
            st_launch_ae_forward(d, h->g, pm, pp, h->spec, knobs_in, B, mag, h->mag_hat_ws, h->phs_hat_ws, h->ri, h->ri_lo, acts,
                                 h->ae_grid, s);
        }
        if (mag_hat_user)
            ST_CUDA_OK(cudaMemcpyAsync(mag_hat_user, h->mag_hat_ws, (long)B * d.OT * d.F * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    ST_LAUNCH_OK(h);
    {   // synthesis: frames_out[(b,t), n] = sum_k ri[(b,t), k] * sfold[k, n]
        StageScope sc(h, SG_GEMM_SYNTH, 1, s);
        int r = -1;
        if (h->use_tc) {
            TcOperand R{h->ri, h->ri_lo, MO, F2, F2}, S{h->sfold, h->sfold_lo, F2, d.N, d.N};
            r = st_launch_gemm_tc(false, true, R, S, h->fo, d.N, MO, d.N, F2, 1, 0, /*promote=*/h->passes == 3, h->sm_count, s, h->passes);
        }
        if (r < 0) {
            GemmOperand R{h->ri, h->ri_lo, F2}, S{h->sfold, h->sfold_lo, d.N};
            ++h->simt_fallbacks;
            st_launch_gemm(true, false, R, S, h->fo, d.N, MO, d.N, F2, 1, 0, s);
        }
    }
    ST_LAUNCH_OK(h);
    if (y_hat) {       // st_train_step passes NULL: its fused tail kernel does the overlap-add together with the loss
        StageScope sc(h, SG_OLA, 1, s);
        st_launch_overlap_add(d, h->fo, x, B, y_hat, acts ? acts[28] : nullptr, acts ? acts[29] : nullptr, s);
    }
    ST_LAUNCH_OK(h);
    h->fwdB = B;
    return 0;
}

extern "C" int st_forward(st_handle* h, const float* x, const float* knobs, int batch, const float* const* params,
                          float* y_hat, float* mag, float* mag_hat, float* const* acts, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !knobs || !y_hat) return st_fail_msg(h, "st_forward: null x / knobs / y_hat");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_forward(params)")) return 1;
    if (acts && check_ptrs(h, (const void* const*)acts, ST_NUM_ACTS, "st_forward(acts)")) return 1;
    return forward_impl(h, x, knobs, batch, params, y_hat, mag, mag_hat, acts, (cudaStream_t)stream);
}

extern "C" int st_loss(st_handle* h, const float* y_hat, const float* y, const float* mag_hat, const float* sbf,
                       float l1_coef, int batch, float* loss, float* g_y_hat, float* g_mag_hat, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!y_hat || !y || !mag_hat || !loss) return st_fail_msg(h, "st_loss: null argument");
    if (batch <= 0) return st_fail_msg(h, "st_loss: batch must be positive");
    ST_CUDA_OK(cudaSetDevice(h->device));
    {
        StageScope sc(h, SG_LOSS, 1, (cudaStream_t)stream);
        st_launch_loss(h->d, y_hat, y, mag_hat, sbf, l1_coef, batch, loss, g_y_hat, g_mag_hat, h->small + SM_LOSS,
                       h->counters + CT_LOSS, (cudaStream_t)stream);
    }
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_loss_shaped(st_handle* h, const float* y_hat, const float* y, const float* mag_hat, const float* sbf,
                              float l1_coef, int batch, int n_wave, int n_frames, int n_bins, float* loss, float* g_y_hat,
                              float* g_mag_hat, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!y_hat || !y || !mag_hat || !loss) return st_fail_msg(h, "st_loss_shaped: null argument");
    if (batch <= 0 || n_wave <= 0 || n_frames <= 0 || n_bins <= 0 || (n_wave & 3))
        return st_fail_msg(h, "st_loss_shaped: need batch, n_frames, n_bins > 0 and n_wave a positive multiple of 4 (got %d, %d, %d, %d)",
                           batch, n_frames, n_bins, n_wave);
    StDims d = h->d;
    d.L = n_wave; d.OT = n_frames; d.F = n_bins;
    {
        StageScope sc(h, SG_LOSS, 1, (cudaStream_t)stream);
        st_launch_loss(d, y_hat, y, mag_hat, sbf, l1_coef, batch, loss, g_y_hat, g_mag_hat, h->small + SM_LOSS,
                       h->counters + CT_LOSS, (cudaStream_t)stream);
    }
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_mae(st_handle* h, const float* a, const float* b, long n, float* out, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!a || !b || !out || n <= 0) return st_fail_msg(h, "st_mae: bad argument");
    ST_CUDA_OK(cudaSetDevice(h->device));
    st_launch_mae(a, b, n, out, h->small + SM_MAE, h->counters + CT_MAE, (cudaStream_t)stream);
    ST_LAUNCH_OK(h);
    return 0;
}

// phase: 0 = the whole backward; 1 = "begin": up to and including the FINAL synthesis gradients (grads[2], grads[3]), so a
// data-parallel caller can start their allreduce; 2 = "finish": the autoencoders and the analysis gradients.
// gwave_ready (st_train_step / st_grad_step): gwave already holds the padded 2*dL/dy_hat, written by the fused forward tail.
// fused_clip (st_train_step only): the final DFT-gradient pass also produces the L1 norm / clip coefficient for the Adam launch
// that follows.
static long packed_grad_floats(const st_handle* h, int* ae_off);

// packed (st_grad_step_packed): the gradients leave as the data-parallel exchange payload instead of the 40 tensors -- the
// final DFT pass writes the live rows only, the autoencoder partial sums land in their payload slots; `grads` is not touched.
static int backward_impl(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat, int B,
                         const float* const* params, float* const* grads, cudaStream_t s, int phase = 0,
                         bool gwave_ready = false, const st_adam* fused_clip = nullptr, float* packed = nullptr) {
    const StDims& d = h->d;
    if (B != h->fwdB || B > h->maxB)
        return st_fail_msg(h, "st_backward: batch %d does not match the preceding st_forward (%d)", B, h->fwdB);
    ST_CUDA_OK(cudaSetDevice(h->device));
    const int MT = B * d.Tp, MO = B * d.OTp, F2 = 2 * d.Fp;
    const long plane = (long)F2 * d.N;
    int ss = -1, sa = -1;
    if (phase != 2) {
    // adjoint of (*2, trim [N:-N]): zero-padded 2*g
    if (!gwave_ready) {
        StageScope sc(h, SG_PAD_G, 1, s);
        // adjoint of (*2, trim [N:-N]): zero-padded 2*g, window stride Sg = OTp*H
        st_launch_pad_split(g_y_hat, h->gwave, h->gwave_lo, B, d.L, d.N, d.Sg, 2.0f, s);
    }
    ST_LAUNCH_OK(h);
    {   // synthesis data gradient: g_ri[(b,t), k] = sum_n gframe[(b,t), n] * sfold[k, n]   (gframe = adjoint of overlap-add)
        StageScope sc(h, SG_GEMM_SYNTH_DGRAD, 1, s);
        int r = -1;
        if (h->use_tc) {
            TcOperand G{h->gwave, h->gwave_lo, MO, d.N, d.H}, S{h->sfold, h->sfold_lo, F2, d.N, d.N};
            r = st_launch_gemm_tc(false, false, G, S, h->g_ri, F2, MO, F2, d.N, 1, 0, /*promote=*/false, h->sm_count, s, h->passes);
        }
        if (r < 0) {
            GemmOperand G{h->gwave, h->gwave_lo, d.H}, S{h->sfold, h->sfold_lo, d.N};
            ++h->simt_fallbacks;
            st_launch_gemm(true, true, G, S, h->g_ri, F2, MO, F2, d.N, 1, 0, s);
        }
    }
    ST_LAUNCH_OK(h);
    {   // synthesis weight gradient (folded): G[k, n] = sum_(b,t) ri[(b,t), k] * gframe[(b,t), n]
        StageScope sc(h, SG_GEMM_SYNTH_WGRAD, 1, s);
        if (h->use_tc) {
            TcOperand R{h->ri, h->ri_lo, MO, F2, F2}, G{h->gwave, h->gwave_lo, MO, d.N, d.H};
            ss = st_launch_gemm_tc(true, true, R, G, h->part_s, d.N, F2, d.N, MO, std::min(kMaxSplits, std::max(1, MO / 512)), plane,
                                   /*promote=*/false, h->sm_count, s, h->passes);
        }
        if (ss < 0) {
            GemmOperand R{h->ri, h->ri_lo, F2}, G{h->gwave, h->gwave_lo, d.H};
            ++h->simt_fallbacks;
            ss = st_launch_gemm(false, false, R, G, h->part_s, d.N, F2, d.N, MO, std::min(kMaxSplits, std::max(1, MO / 256)), plane, s);
        }
    }
    ST_LAUNCH_OK(h);
    h->bwd_ss = ss;
    if (phase == 1) {       // the synthesis pair is complete after its split-K planes are summed and un-folded
        StageScope sc(h, SG_FINALIZE, 1, s);
        st_launch_finalize_dft_grads(d, h->part_a, h->part_s, 0, ss, grads[0], grads[1], grads[2], grads[3], 2, s);
        ST_LAUNCH_OK(h);
        return 0;
    }
    }   // phase != 2
    if (phase == 2) ss = h->bwd_ss;
    int part_ctas = h->ae_grid;
    {   // both autoencoders: back-propagate, dL/d(re|im) -> g_spec, per-CTA weight-gradient partials.  Tensor-core path
        // from the saved activations when the forward wrote them; otherwise the SIMT kernel recomputes the chain.
        StageScope sc(h, SG_AE_BWD, 2 + (h->tm_fwd && !h->tm_bwd_image), s);
        AeParams pm, pp;
        split_params(params, pm, pp);
        int gr = 0;
        if (h->tm_fwd && h->use_tm)       // recompute in tensor memory; the weight image is packed here when the forward was an eval one
            gr = st_launch_ae_backward_tm(d, h->g, pm, pp, h->spec, h->trk_ws, h->knobs_ws, B, h->mag_hat_ws, h->phs_hat_ws, h->g_ri, g_mag_hat, g_mag,
                                          h->gtrack_ws, h->g_spec, h->g_spec_lo, h->ae_part, h->ae_wpack_bwd, nullptr, h->ae_timing, h->sm_count,
                                          !h->tm_bwd_image, s, s);
        if (gr > 0) h->tm_bwd_image = true;
        if (gr == 0 && h->have_saves)
            gr = st_launch_ae_backward_mma(d, h->g, pm, pp, h->spec, B, h->ae_save_m, h->ae_save_p, h->mag_hat_ws, h->phs_hat_ws,
                                           h->g_ri, g_mag_hat, g_mag, h->tail_ws, h->g_spec, h->g_spec_lo, h->ae_part, h->ae_timing, h->sm_count, s, h->passes);
        if (gr > 0)
            part_ctas = gr;
        else {
            ++h->simt_fallbacks;
            st_launch_ae_backward(d, h->g, pm, pp, h->spec, h->knobs_ws, B, h->mag_hat_ws, h->phs_hat_ws, h->g_ri, g_mag_hat,
                                  g_mag, h->g_spec, h->g_spec_lo, h->ae_part, h->ae_grid, s);
        }
    }
    ST_LAUNCH_OK(h);
    // the per-CTA autoencoder partials are summed on the side stream, beside the analysis weight-gradient GEMM
    const bool beside = h->side && !h->prof_on;
    {
        cudaStream_t sr = beside ? h->side : s;
        StageScope sc(h, SG_AE_REDUCE, 1, s);
        AeGrads gm, gp;
        if (packed) {
            int ae_off[ST_NUM_PARAMS - 4];
            packed_grad_floats(h, ae_off);
            float* base = packed + 4L * d.F * d.N;
            for (int l = 0; l < ST_AE_LAYERS; ++l) {
                gm.W[l] = base + ae_off[2 * l];       gm.b[l] = base + ae_off[2 * l + 1];
                gp.W[l] = base + ae_off[18 + 2 * l];  gp.b[l] = base + ae_off[19 + 2 * l];
            }
        } else {
            for (int l = 0; l < ST_AE_LAYERS; ++l) {
                gm.W[l] = grads[4 + 2 * l];  gm.b[l] = grads[5 + 2 * l];
                gp.W[l] = grads[22 + 2 * l]; gp.b[l] = grads[23 + 2 * l];
            }
        }
        if (beside) {
            ST_CUDA_OK(cudaEventRecord(h->ev_fork, s));
            ST_CUDA_OK(cudaStreamWaitEvent(sr, h->ev_fork, 0));
        }
        st_launch_ae_grad_reduce(h->g, h->ae_part, part_ctas, gm, gp, sr);
        if (beside) ST_CUDA_OK(cudaEventRecord(h->ev_join, sr));
    }
    ST_LAUNCH_OK(h);
    {   // analysis weight gradient: G[(re|im) k, n] = sum_(b,t) g_spec[(b,t), k] * frame[(b,t), n]
        StageScope sc(h, SG_GEMM_ANALYSIS_WGRAD, 1, s);
        if (h->use_tc) {
            TcOperand Gs{h->g_spec, h->g_spec_lo, MT, F2, F2}, X{h->xpad, h->xpad_lo, MT, d.N, d.H};
            sa = st_launch_gemm_tc(true, true, Gs, X, h->part_a, d.N, F2, d.N, MT, std::min(kMaxSplits, std::max(1, MT / 512)), plane,
                                   /*promote=*/false, h->sm_count, s, h->passes);
        }
        if (sa < 0) {
            GemmOperand Gs{h->g_spec, h->g_spec_lo, F2}, X{h->xpad, h->xpad_lo, d.H};
            ++h->simt_fallbacks;
            sa = st_launch_gemm(false, false, Gs, X, h->part_a, d.N, F2, d.N, MT, std::min(kMaxSplits, std::max(1, MT / 256)), plane, s);
        }
    }
    ST_LAUNCH_OK(h);
    {
        StageScope sc(h, SG_FINALIZE, 1, s);
        if (packed)
            st_launch_finalize_packed(d, h->part_a, h->part_s, sa, ss, packed, s);
        else if (fused_clip && phase == 0)
            st_launch_finalize_norm(d, h->part_a, h->part_s, sa, ss, grads[0], grads[1], grads[2], grads[3], fused_clip->grad_scale,
                                    fused_clip->max_norm, h->small + SM_TOTAL_NORM, h->small + SM_COEF, h->small + SM_NORM,
                                    h->counters + CT_NORM, s);
        else
            st_launch_finalize_dft_grads(d, h->part_a, h->part_s, sa, ss, grads[0], grads[1], grads[2], grads[3], phase == 2 ? 1 : 3, s);
    }
    if (beside) ST_CUDA_OK(cudaStreamWaitEvent(s, h->ev_join, 0));        // all 40 gradients are complete on the caller's stream
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_backward(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat, int batch,
                           const float* const* params, float* const* grads, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!g_y_hat) return st_fail_msg(h, "st_backward: null g_y_hat");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_backward(params)")) return 1;
    if (check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_backward(grads)")) return 1;
    return backward_impl(h, g_y_hat, g_mag, g_mag_hat, batch, params, grads, (cudaStream_t)stream);
}

extern "C" int st_backward_begin(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat, int batch,
                                 const float* const* params, float* const* grads, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!g_y_hat) return st_fail_msg(h, "st_backward_begin: null g_y_hat");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_backward_begin(params)")) return 1;
    if (check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_backward_begin(grads)")) return 1;
    h->bwd_ss = -1;
    return backward_impl(h, g_y_hat, g_mag, g_mag_hat, batch, params, grads, (cudaStream_t)stream, 1);
}

extern "C" int st_backward_finish(st_handle* h, const float* g_y_hat, const float* g_mag, const float* g_mag_hat, int batch,
                                  const float* const* params, float* const* grads, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (h->bwd_ss < 0) return st_fail_msg(h, "st_backward_finish: no st_backward_begin before it");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_backward_finish(params)")) return 1;
    if (check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_backward_finish(grads)")) return 1;
    const int rc = backward_impl(h, g_y_hat, g_mag, g_mag_hat, batch, params, grads, (cudaStream_t)stream, 2);
    h->bwd_ss = -1;
    return rc;
}

extern "C" int st_clip_grad_norm(st_handle* h, float* const* grads, float max_norm, float* total_norm, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (check_ptrs(h, (const void* const*)grads, 4, "st_clip_grad_norm(grads)")) return 1;
    ST_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const float* g[4] = {grads[0], grads[1], grads[2], grads[3]};
    {
        StageScope sc(h, SG_L1NORM, 2, s);
        st_launch_l1_norm4(g, (long)h->d.N * h->d.N, 0, 0, 1.f, max_norm, total_norm ? total_norm : h->small + SM_TOTAL_NORM,
                           h->small + SM_COEF, h->small + SM_NORM, h->counters + CT_NORM, s);
        float* gw[4] = {grads[0], grads[1], grads[2], grads[3]};
        st_launch_scale4(gw, (long)h->d.N * h->d.N, h->small + SM_COEF, s);   // torch multiplies even when coef == 1
    }
    ST_LAUNCH_OK(h);
    return 0;
}

// Kernel arguments of the Adam launch: tensor table (live rows only for the analysis pair when asked) and the per-step scalars.
static void adam_args(const st_handle* h, float* const* params, const float* const* grads, float* const* m, float* const* v,
                      const st_adam* hp, bool live_only, AdamTensors& t, AdamScalars& sc) {
    for (int i = 0; i < ST_NUM_PARAMS; ++i) {
        t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = m[i]; t.v[i] = v[i];
        t.n[i] = (live_only && i < 2) ? (long)h->d.F * h->d.N : h->numel[i];
    }
    // bias corrections in double, as torch does on the host for non-capturable Adam
    const double bc1 = 1.0 - std::pow((double)hp->beta1, (double)hp->step);
    const double bc2 = 1.0 - std::pow((double)hp->beta2, (double)hp->step);
    sc.lr_over_bc1 = (float)((double)hp->lr / bc1);
    sc.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
    sc.beta1 = hp->beta1; sc.beta2 = hp->beta2; sc.eps = hp->eps; sc.grad_scale = hp->grad_scale;
}

static int adam_impl(st_handle* h, float* const* params, const float* const* grads, float* const* m, float* const* v,
                     const st_adam* hp, bool live_only, cudaStream_t s, bool coef_ready = false) {
    if (hp->step < 1) return st_fail_msg(h, "st_adam_step: step must be >= 1 (got %d)", hp->step);
    ST_CUDA_OK(cudaSetDevice(h->device));
    const float* coef = nullptr;
    if (hp->max_norm > 0.f && coef_ready) coef = h->small + SM_COEF;      // written by the fused finalize pass of st_train_step
    if (hp->max_norm > 0.f && !coef_ready) {
        const float* g[4] = {grads[0], grads[1], grads[2], grads[3]};
        {
            StageScope sc(h, SG_L1NORM, 1, s);
            st_launch_l1_norm4(g, (long)h->d.N * h->d.N, 0, 0, hp->grad_scale, hp->max_norm, h->small + SM_TOTAL_NORM,
                               h->small + SM_COEF, h->small + SM_NORM, h->counters + CT_NORM, s);
        }
        ST_LAUNCH_OK(h);
        coef = h->small + SM_COEF;
    }
    AdamTensors t;
    AdamScalars sc;
    adam_args(h, params, grads, m, v, hp, live_only, t, sc);
    {
        StageScope scope(h, SG_ADAM, 1, s);
        st_launch_adam(t, live_only ? h->map_live : h->map_full, live_only ? h->n_live : h->n_full, sc, coef, s);
    }
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_adam_step(st_handle* h, float* const* params, const float* const* grads, float* const* exp_avg,
                            float* const* exp_avg_sq, const st_adam* hp, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!hp) return st_fail_msg(h, "st_adam_step: null hyper-parameters");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_adam_step(params)") ||
        check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_adam_step(grads)") ||
        check_ptrs(h, (const void* const*)exp_avg, ST_NUM_PARAMS, "st_adam_step(exp_avg)") ||
        check_ptrs(h, (const void* const*)exp_avg_sq, ST_NUM_PARAMS, "st_adam_step(exp_avg_sq)"))
        return 1;
    return adam_impl(h, params, grads, exp_avg, exp_avg_sq, hp, /*live_only=*/false, (cudaStream_t)stream);
}

// forward without its overlap-add; one kernel then does overlap-add + residual + loss + both loss gradients and writes the
// padded (hi, lo) 2*dL/dy_hat operand; then the whole backward.  With fused_clip the last DFT-gradient pass also yields the
// clip coefficient (single-GPU step); without it the gradients are left for the caller's allreduce.
static int grad_step_impl(st_handle* h, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                          float* const* grads, const float* sbf, float l1_coef, float* loss, cudaStream_t s, const st_adam* fused_clip,
                          float* packed = nullptr) {
    if (forward_impl(h, x, knobs, batch, params, nullptr, nullptr, nullptr, nullptr, s)) return 1;
    {
        StageScope sc(h, SG_LOSS, 1, s);
        st_launch_ola_loss(h->d, h->fo, x, y, h->mag_hat_ws, sbf, l1_coef, batch, loss, h->gwave, h->gwave_lo, h->gmh_ws,
                           h->small + SM_LOSS, h->counters + CT_LOSS, s);
    }
    ST_LAUNCH_OK(h);
    return backward_impl(h, h->gy_ws, nullptr, h->gmh_ws, batch, params, grads, s, 0, /*gwave_ready=*/true, fused_clip, packed);
}

extern "C" int st_grad_step(st_handle* h, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                            float* const* grads, const float* sbf, float l1_coef, float* loss, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !y || !knobs || !loss) return st_fail_msg(h, "st_grad_step: null argument");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_grad_step(params)") ||
        check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_grad_step(grads)"))
        return 1;
    if (ensure_workspace(h, batch)) return 1;
    return grad_step_impl(h, x, y, knobs, batch, params, grads, sbf, l1_coef, loss, (cudaStream_t)stream, nullptr);
}

static long packed_grad_floats(const st_handle* h, int* ae_off) {
    long n = 4L * h->d.F * h->d.N;
    long o = 0;
    for (int t = 4; t < ST_NUM_PARAMS; ++t) {
        if (ae_off) ae_off[t - 4] = (int)o;
        o += (h->numel[t] + 3) / 4 * 4;
    }
    return n + o;
}

extern "C" long st_packed_grad_floats(const st_handle* h) { return h ? packed_grad_floats(h, nullptr) : -1; }

static int pack_impl(st_handle* h, float* const* grads, float* packed, int dir, cudaStream_t s) {
    if (check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_pack_grads(grads)")) return 1;
    if (!packed) return st_fail_msg(h, "st_pack_grads: null packed buffer");
    GradPack gp;
    for (int t = 0; t < ST_NUM_PARAMS; ++t) gp.g[t] = grads[t];
    gp.packed = packed;
    gp.live = (long)h->d.F * h->d.N;
    gp.N = h->d.N;
    packed_grad_floats(h, gp.ae_off);
    for (int t = 4; t < ST_NUM_PARAMS; ++t) gp.ae_n[t - 4] = (int)h->numel[t];
    st_launch_pack_grads(gp, dir, s);
    h->launches += 1;
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_pack_grads(st_handle* h, float* const* grads, float* packed, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    return pack_impl(h, grads, packed, 0, (cudaStream_t)stream);
}

extern "C" int st_unpack_grads(st_handle* h, const float* packed, float* const* grads, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    return pack_impl(h, grads, const_cast<float*>(packed), 1, (cudaStream_t)stream);
}

static int step_with_graph(st_handle* h, int kind, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                           float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const float* sbf, float l1_coef,
                           const st_adam* hp, float* loss, float* packed, cudaStream_t s);

extern "C" int st_grad_step_packed(st_handle* h, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                                   float* packed, const float* sbf, float l1_coef, float* loss, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !y || !knobs || !loss || !packed) return st_fail_msg(h, "st_grad_step_packed: null argument");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_grad_step_packed(params)")) return 1;
    return step_with_graph(h, 1, x, y, knobs, batch, params, nullptr, nullptr, nullptr, sbf, l1_coef, nullptr, loss, packed,
                           (cudaStream_t)stream);
}

extern "C" int st_unpack_clip(st_handle* h, const float* packed, float* const* grads, float grad_scale, float max_norm,
                              float* total_norm, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_unpack_clip(grads)")) return 1;
    if (!packed) return st_fail_msg(h, "st_unpack_clip: null packed buffer");
    GradPack gp;
    for (int t = 0; t < ST_NUM_PARAMS; ++t) gp.g[t] = grads[t];
    gp.packed = const_cast<float*>(packed);
    gp.live = (long)h->d.F * h->d.N;
    gp.N = h->d.N;
    packed_grad_floats(h, gp.ae_off);
    for (int t = 4; t < ST_NUM_PARAMS; ++t) gp.ae_n[t - 4] = (int)h->numel[t];
    cudaStream_t s = (cudaStream_t)stream;
    {
        StageScope sc(h, SG_L1NORM, 1, s);
        st_launch_unpack_clip(gp, grad_scale, max_norm, total_norm ? total_norm : h->small + SM_TOTAL_NORM, h->small + SM_COEF,
                              h->small + SM_NORM, h->counters + CT_NORM, s);
    }
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_adam_step_clipped(st_handle* h, float* const* params, const float* const* grads, float* const* exp_avg,
                                    float* const* exp_avg_sq, const st_adam* hp, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!hp) return st_fail_msg(h, "st_adam_step_clipped: null hyper-parameters");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_adam_step_clipped(params)") ||
        check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_adam_step_clipped(grads)") ||
        check_ptrs(h, (const void* const*)exp_avg, ST_NUM_PARAMS, "st_adam_step_clipped(exp_avg)") ||
        check_ptrs(h, (const void* const*)exp_avg_sq, ST_NUM_PARAMS, "st_adam_step_clipped(exp_avg_sq)"))
        return 1;
    // the clip coefficient is the one st_unpack_clip left in the handle; rows >= F of the analysis tensors never change
    return adam_impl(h, params, grads, exp_avg, exp_avg_sq, hp, /*live_only=*/true, (cudaStream_t)stream, /*coef_ready=*/true);
}

static int train_step_body(st_handle* h, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                           float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const float* sbf, float l1_coef,
                           const st_adam* hp, float* loss, cudaStream_t s) {
    if (h->fuse_tail) {
        if (grad_step_impl(h, x, y, knobs, batch, params, grads, sbf, l1_coef, loss, s, hp)) return 1;
        return adam_impl(h, params, grads, exp_avg, exp_avg_sq, hp, /*live_only=*/true, s, /*coef_ready=*/true);
    }
    if (forward_impl(h, x, knobs, batch, params, h->yhat_ws, nullptr, nullptr, nullptr, s)) return 1;
    {
        StageScope sc(h, SG_LOSS, 1, s);
        st_launch_loss(h->d, h->yhat_ws, y, h->mag_hat_ws, sbf, l1_coef, batch, loss, h->gy_ws, h->gmh_ws, h->small + SM_LOSS,
                       h->counters + CT_LOSS, s);
    }
    ST_LAUNCH_OK(h);
    if (backward_impl(h, h->gy_ws, nullptr, h->gmh_ws, batch, params, grads, s)) return 1;
    // gradients of the dead analysis rows are exactly zero here, so Adam may skip them
    return adam_impl(h, params, grads, exp_avg, exp_avg_sq, hp, /*live_only=*/true, s);
}

// One captured train step.  The step is a fixed sequence of ~17 dependent launches over two streams whose only per-step inputs
// are the Adam scalars (learning rate, bias corrections), so the second call with the same buffers, batch size and stream is
// captured (cudaStreamBeginCapture; the side stream joins through the fork / join events) and every later one is a single
// cudaGraphLaunch after the Adam node's scalars have been refreshed.  Not on the legacy default stream (it cannot be captured),
// not while stage profiling is on, ST_CUDA_GRAPH=0 turns it off; any capture error falls back to the plain launches for good.
struct StepGraph {
    const void *sbf, *g0 = nullptr, *m0 = nullptr, *v0 = nullptr, *packed = nullptr;
    const void* ptab[ST_NUM_PARAMS];     // the parameter table the step was captured with
    int kind = 0, batch, fuse, training, precision;
    float l1, beta1 = 0.f, beta2 = 0.f, eps = 0.f, grad_scale = 0.f, max_norm = 0.f;
    cudaStream_t s;
    int seen = 0;
    bool bad = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    // the nodes whose arguments change from step to step: Adam (scalars), input padding (x), fused loss tail (x, y), knob copy
    cudaGraphNode_t adam = nullptr, pad = nullptr, ola = nullptr, kcopy = nullptr;
    cudaKernelNodeParams kp{}, kp_pad{}, kp_ola{};
    void* knobs_dst = nullptr;
    size_t knobs_bytes = 0;
    long launches = 0;
};

static void drop_step_graphs(st_handle* h) {
    for (StepGraph* g : h->graphs) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
    }
    h->graphs.clear();
}

extern "C" long st_debug_graph_replays(const st_handle* h) { return h ? h->graph_replays : -1; }

// kind 0: st_train_step (forward .. Adam); kind 1: st_grad_step_packed (forward .. packed gradient payload, no update)
static int step_body(st_handle* h, int kind, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                     float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const float* sbf, float l1_coef,
                     const st_adam* hp, float* loss, float* packed, cudaStream_t s) {
    if (kind == 0) return train_step_body(h, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, s);
    return grad_step_impl(h, x, y, knobs, batch, params, nullptr, sbf, l1_coef, loss, s, nullptr, packed);
}

static int step_with_graph(st_handle* h, int kind, const float* x, const float* y, const float* knobs, int batch, float* const* params,
                           float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const float* sbf, float l1_coef,
                           const st_adam* hp, float* loss, float* packed, cudaStream_t s) {
    if (ensure_workspace(h, batch)) return 1;
    const bool graphable = h->use_graph && h->fuse_tail && !h->prof_on && h->side && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
    if (!graphable) return step_body(h, kind, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, packed, s);

    StepGraph* g = nullptr;
    for (StepGraph* c : h->graphs) {
        bool same = c->kind == kind && c->packed == (const void*)packed && c->sbf == sbf && c->batch == batch && c->s == s &&
                    c->l1 == l1_coef && c->fuse == (int)h->fuse_tail && c->training == (int)h->training && c->precision == h->passes;
        if (same && kind == 0)
            same = c->beta1 == hp->beta1 && c->beta2 == hp->beta2 && c->eps == hp->eps && c->grad_scale == hp->grad_scale &&
                   c->max_norm == hp->max_norm;
        // the 40-entry tables are compared entry by entry: a caller may re-point a single tensor
        for (int i = 0; same && i < ST_NUM_PARAMS; ++i)
            same = c->ptab[i] == (const void*)params[i];
        if (same && kind == 0) same = c->g0 == (const void*)grads[0] && c->m0 == (const void*)exp_avg[0] && c->v0 == (const void*)exp_avg_sq[0];
        if (same) { g = c; break; }
    }
    if (!g) {
        if (h->graphs.size() >= 8) drop_step_graphs(h);         // a caller that keeps changing buffers: start over
        g = new StepGraph();
        for (int i = 0; i < ST_NUM_PARAMS; ++i) g->ptab[i] = params[i];
        g->kind = kind; g->packed = packed; g->sbf = sbf; g->batch = batch; g->s = s; g->l1 = l1_coef;
        if (kind == 0) {
            g->g0 = grads[0]; g->m0 = exp_avg[0]; g->v0 = exp_avg_sq[0];
            g->beta1 = hp->beta1; g->beta2 = hp->beta2; g->eps = hp->eps; g->grad_scale = hp->grad_scale; g->max_norm = hp->max_norm;
        }
        g->fuse = (int)h->fuse_tail; g->training = (int)h->training;
        g->precision = h->passes;
        h->graphs.push_back(g);
    }
    if (g->exec) {                                              // replay: refresh the per-step arguments, one launch
        bool ok = true;
        if (kind == 0) {
            AdamTensors t;
            AdamScalars sc;
            adam_args(h, params, grads, exp_avg, exp_avg_sq, hp, /*live_only=*/true, t, sc);
            void** args = g->kp.kernelParams;                   // [tensors, chunk map, scalars, clip coefficient]
            void* mine[4] = {args[0], args[1], &sc, args[3]};
            cudaKernelNodeParams kp = g->kp;
            kp.kernelParams = mine;
            ok = cudaGraphExecKernelNodeSetParams(g->exec, g->adam, &kp) == cudaSuccess;
        }
        if (ok) {                                               // this step's inputs (the batch may live anywhere)
            void* pa[8];
            for (int i = 0; i < 8; ++i) pa[i] = g->kp_pad.kernelParams[i];
            pa[0] = &x;
            cudaKernelNodeParams k2 = g->kp_pad;
            k2.kernelParams = pa;
            ok = cudaGraphExecKernelNodeSetParams(g->exec, g->pad, &k2) == cudaSuccess;
            void* oa[14];
            for (int i = 0; i < 14; ++i) oa[i] = g->kp_ola.kernelParams[i];
            oa[2] = &x; oa[3] = &y; oa[8] = &loss;
            cudaKernelNodeParams k3 = g->kp_ola;
            k3.kernelParams = oa;
            ok = ok && cudaGraphExecKernelNodeSetParams(g->exec, g->ola, &k3) == cudaSuccess;
            if (ok && g->kcopy)
                ok = cudaGraphExecMemcpyNodeSetParams1D(g->exec, g->kcopy, g->knobs_dst, knobs, g->knobs_bytes, cudaMemcpyDeviceToDevice) == cudaSuccess;
        }
        if (ok) ok = cudaGraphLaunch(g->exec, s) == cudaSuccess;
        if (ok) {
            h->launches += g->launches;
            ++h->graph_replays;
            return 0;
        }
        // a refused update or launch (nothing of this step has run): give the graph up for these buffers, plain launches from here on
        cudaGetLastError();
        cudaGraphExecDestroy(g->exec);
        g->exec = nullptr;
        g->bad = true;
        return step_body(h, kind, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, packed, s);
    }
    if (g->bad || g->seen++ < 1)                                // first sight of these buffers (or capture refused earlier): plain launches
        return step_body(h, kind, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, packed, s);

    // ---- capture
    const long launches0 = h->launches;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        g->bad = true;
        return step_body(h, kind, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, packed, s);
    }
    const int rc = step_body(h, kind, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, packed, s);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    bool ok = rc == 0 && e == cudaSuccess && graph != nullptr;
    if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
    if (ok) {                                                   // the Adam node: the only kernel node whose function is adam_kernel
        size_t n = 0;
        ok = cudaGraphGetNodes(graph, nullptr, &n) == cudaSuccess && n > 0;
        std::vector<cudaGraphNode_t> nodes(n);
        if (ok) ok = cudaGraphGetNodes(graph, nodes.data(), &n) == cudaSuccess;
        const void *f_adam = st_adam_kernel_fn(), *f_pad = st_pad_split_kernel_fn(), *f_ola = st_ola_loss_kernel_fn();
        int n_pad = 0, n_ola = 0, n_adam = 0, n_copy = 0;
        for (size_t i = 0; ok && i < n; ++i) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess) continue;
            if (ty == cudaGraphNodeTypeMemcpy) { g->kcopy = nodes[i]; ++n_copy; continue; }
            if (ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp{};
            if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) continue;
            if (kp.func == f_adam) { g->adam = nodes[i]; g->kp = kp; ++n_adam; }
            else if (kp.func == f_pad) { g->pad = nodes[i]; g->kp_pad = kp; ++n_pad; }
            else if (kp.func == f_ola) { g->ola = nodes[i]; g->kp_ola = kp; ++n_ola; }
        }
        const bool keep_knobs = h->d.K > 0;
        ok = ok && n_adam == (kind == 0 ? 1 : 0) && n_pad == 1 && n_ola == 1 && n_copy == (keep_knobs ? 1 : 0);   // exactly the nodes this code knows
        g->knobs_dst = h->knobs_ws;
        g->knobs_bytes = (size_t)batch * h->d.K * sizeof(float);
    }
    if (!ok) {
        cudaGetLastError();
        if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
        if (graph) cudaGraphDestroy(graph);
        g->bad = true;
        h->launches = launches0;
        return step_body(h, kind, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, packed, s);
    }
    g->graph = graph;                                           // kept: the node handle above belongs to it
    g->launches = h->launches - launches0;
    ST_CUDA_OK(cudaGraphLaunch(g->exec, s));
    ++h->graph_replays;
    return 0;
}

extern "C" int st_train_step(st_handle* h, const float* x, const float* y, const float* knobs, int batch,
                             float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                             const float* sbf, float l1_coef, const st_adam* hp, float* loss, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !y || !knobs || !hp || !loss) return st_fail_msg(h, "st_train_step: null argument");
    if (check_ptrs(h, (const void* const*)params, ST_NUM_PARAMS, "st_train_step(params)") ||
        check_ptrs(h, (const void* const*)grads, ST_NUM_PARAMS, "st_train_step(grads)") ||
        check_ptrs(h, (const void* const*)exp_avg, ST_NUM_PARAMS, "st_train_step(exp_avg)") ||
        check_ptrs(h, (const void* const*)exp_avg_sq, ST_NUM_PARAMS, "st_train_step(exp_avg_sq)"))
        return 1;
    if (hp->step < 1) return st_fail_msg(h, "st_adam_step: step must be >= 1 (got %d)", hp->step);
    return step_with_graph(h, 0, x, y, knobs, batch, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss, nullptr, (cudaStream_t)stream);
}

// Host-only: the launch plan of the five front-end contractions of THIS geometry at a batch size (no device work, so the CPU
// test suite can check the tile / split-K heuristic): out[5][4] = {pair kernel?, BN, split-K planes, grid} for
// analysis, synthesis, synthesis dgrad, synthesis wgrad, analysis wgrad.  sm_count <= 0: the handle's device.
extern "C" int st_debug_gemm_plan(const st_config* cfg, int batch, int sm_count, int* out) {
    if (!cfg || !out || batch <= 0) return 1;
    const int N = cfg->ft, H = cfg->hop, F = N / 2 + 1;
    int Fp = round_up(F, 16);
    while (st_tc_pick_bn(2 * Fp) < 128) Fp += 16;
    const int L = (cfg->frames_out - 1) * H - N;
    const int Tp = (cfg->chunk + 2 * N + H - 1) / H, OTp = (L + 2 * N + H - 1) / H;
    const int MT = batch * Tp, MO = batch * OTp, F2 = 2 * Fp;
    if (sm_count <= 0) sm_count = 148;
    int rc = 0;
    rc |= st_tc_plan(false, MT, F2, N, 1, sm_count, out + 0);
    rc |= st_tc_plan(true, MO, N, F2, 1, sm_count, out + 4);
    rc |= st_tc_plan(false, MO, F2, N, 1, sm_count, out + 8);
    rc |= st_tc_plan(true, F2, N, MO, std::min(kMaxSplits, std::max(1, MO / 512)), sm_count, out + 12);
    rc |= st_tc_plan(true, F2, N, MT, std::min(kMaxSplits, std::max(1, MT / 512)), sm_count, out + 16);
    return rc ? 1 : 0;
}

extern "C" long st_debug_numel(st_handle* h, const char* name) {
    if (!h || !name) return -1;
    const StDims& d = h->d;
    const long B = h->fwdB;
    if (!strcmp(name, "spec") || !strcmp(name, "g_spec")) return B * d.Tp * 2 * d.Fp;
    if (!strcmp(name, "ri") || !strcmp(name, "g_ri")) return B * d.OTp * 2 * d.Fp;
    if (!strcmp(name, "frames_out")) return B * d.OTp * d.N;
    if (!strcmp(name, "wcat") || !strcmp(name, "sfold")) return 2L * d.Fp * d.N;
    if (!strcmp(name, "xpad")) return B * d.Sx;
    if (!strcmp(name, "phs_hat")) return B * d.OT * d.F;
    return -1;
}

extern "C" int st_debug_read(st_handle* h, const char* name, float* dst, long n) {
    if (!h || !name || !dst) return 1;
    const long have = st_debug_numel(h, name);
    if (have < 0) return st_fail_msg(h, "st_debug_read: unknown buffer '%s'", name);
    const float* src = !strcmp(name, "spec") ? h->spec : !strcmp(name, "g_spec") ? h->g_spec : !strcmp(name, "ri") ? h->ri
                     : !strcmp(name, "g_ri") ? h->g_ri : !strcmp(name, "frames_out") ? h->fo : !strcmp(name, "wcat") ? h->wcat
                     : !strcmp(name, "sfold") ? h->sfold : !strcmp(name, "xpad") ? h->xpad : h->phs_hat_ws;
    if (!src) return st_fail_msg(h, "st_debug_read: buffer '%s' not allocated yet", name);
    ST_CUDA_OK(cudaSetDevice(h->device));
    ST_CUDA_OK(cudaDeviceSynchronize());
    ST_CUDA_OK(cudaMemcpy(dst, src, std::min(n, have) * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int st_set_precision(st_handle* h, int mode) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (mode != ST_PRECISION_FP32 && mode != ST_PRECISION_TF32)
        return st_fail_msg(h, "st_set_precision: unknown mode %d (0 = fp32, 1 = tf32)", mode);
    if (mode == ST_PRECISION_TF32 && !h->use_tc)
        return st_fail_msg(h, "st_set_precision: the reduced-precision mode needs the tcgen05 GEMMs (ST_DISABLE_TCGEN05 is set)");
    h->passes = mode == ST_PRECISION_TF32 ? 1 : 3;
    h->fwdB = 0;                  // a backward must not mix modes with the forward that saved its activations
    return 0;
}
extern "C" int st_get_precision(const st_handle* h) { return h ? (h->passes == 1 ? ST_PRECISION_TF32 : ST_PRECISION_FP32) : -1; }

extern "C" int st_set_training(st_handle* h, int on) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    h->training = on != 0;
    return 0;
}

extern "C" long st_launch_count(const st_handle* h) { return h ? h->launches : -1; }
extern "C" int st_profile_stage_count(void) { return SG_COUNT; }
extern "C" const char* st_profile_stage_name(int i) { return (i >= 0 && i < SG_COUNT) ? kStageNames[i] : ""; }

extern "C" int st_profile(st_handle* h, int enable) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    h->prof_on = enable != 0;
    return 0;
}

// Synchronises the device, adds up the event-timed milliseconds and call counts per stage since the last
// read, and clears the record.  ms / calls: arrays of st_profile_stage_count() entries.
extern "C" int st_profile_read(st_handle* h, float* ms, long* calls) {
    if (!h || !ms || !calls) return 1;
    ST_CUDA_OK(cudaSetDevice(h->device));
    ST_CUDA_OK(cudaDeviceSynchronize());
    for (int i = 0; i < SG_COUNT; ++i) { ms[i] = 0.f; calls[i] = 0; }
    for (auto& e : h->prof_events) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e.second.first, e.second.second);
        ms[e.first] += t;
        calls[e.first] += 1;
        cudaEventDestroy(e.second.first);
        cudaEventDestroy(e.second.second);
    }
    h->prof_events.clear();
    return 0;
}

// Test hook: one GEMM on caller-provided (hi, lo) operands through either implementation.
//   a_mn / b_mn: 0 = K-major ([rows = M or N][K], leading dim ld), 1 = MN-major ([rows = K][M or N]).
//   use_tc: 1 = tcgen05/TMA kernel, 0 = FFMA kernel.  Returns the number of split planes written into C
//   (plane stride = M * ldc), or -1.
extern "C" int st_debug_gemm(st_handle* h, int use_tc /*0 FFMA, 1 tcgen05, 2 tcgen05 promoted, 3 tcgen05 single-pass TF32 (hi planes only)*/, int a_mn, int b_mn, const float* a_hi, const float* a_lo, long a_ld,
                             const float* b_hi, const float* b_lo, long b_ld, float* C, long ldc, int M, int N, int K,
                             int splits, void* stream) {
    if (!h) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaSetDevice(h->device) != cudaSuccess) return -1;
    int r;
    if (use_tc) {
        TcOperand A{a_hi, a_lo, a_mn ? K : M, a_mn ? M : K, a_ld}, B{b_hi, b_lo, b_mn ? K : N, b_mn ? N : K, b_ld};
        r = st_launch_gemm_tc(a_mn != 0, b_mn != 0, A, B, C, ldc, M, N, K, splits, (long)M * ldc, use_tc == 2, h->sm_count, s, use_tc == 3 ? 1 : 3);
    } else {
        GemmOperand A{a_hi, a_lo, a_ld}, B{b_hi, b_lo, b_ld};
        r = st_launch_gemm(!a_mn, !b_mn, A, B, C, ldc, M, N, K, splits, (long)M * ldc, s);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return r;
}

// Diagnostic: enable (on=1) region timing of the tensor-core autoencoder backward and read the 16 cycle counters
// (backward: 8 regions x {magnitude, phase}; then tcgen05 forward: 4 regions x 2; 24 values).  Reading resets them.  out may be NULL.
extern "C" long st_debug_fallbacks(const st_handle* h) { return h ? h->simt_fallbacks : -1; }

extern "C" int st_debug_ae_timing(st_handle* h, int on, long long* out_host) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    ST_CUDA_OK(cudaSetDevice(h->device));
    ST_CUDA_OK(cudaDeviceSynchronize());
    if (on && !h->ae_timing) {
        ST_CUDA_OK(cudaMalloc(&h->ae_timing, 64 * sizeof(long long)));
        ST_CUDA_OK(cudaMemset(h->ae_timing, 0, 64 * sizeof(long long)));
    }
    if (out_host && h->ae_timing) {
        ST_CUDA_OK(cudaMemcpy(out_host, h->ae_timing, 24 * sizeof(long long), cudaMemcpyDeviceToHost));
        ST_CUDA_OK(cudaMemset(h->ae_timing, 0, 64 * sizeof(long long)));
    }
    if (!on && h->ae_timing) { cudaFree(h->ae_timing); h->ae_timing = nullptr; }
    return 0;
}

// ---- DCT / MDCT front-end variant (cls_fe_dct_bases.py): same contraction + overlap-add machinery, other sizes ------
static int dct_workspace(st_handle* h, long floats) {
    ST_CUDA_OK(cudaSetDevice(h->device));
    if (floats <= h->dct_ws_floats) return 0;
    ST_CUDA_OK(cudaDeviceSynchronize());
    if (h->dct_ws) cudaFree(h->dct_ws);
    h->dct_ws = nullptr;
    h->dct_ws_floats = 0;
    ST_CUDA_OK(cudaMalloc(&h->dct_ws, floats * sizeof(float)));
    h->dct_ws_floats = floats;
    return 0;
}
static long up4(long n) { return (n + 3) / 4 * 4; }

extern "C" int st_dct_analysis(st_handle* h, const float* x, const float* w, const float* bias, int batch, int chunk, int ft_size,
                               int w_size, int hop, float* out, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !w || !bias || !out) return st_fail_msg(h, "st_dct_analysis: null argument");
    if (batch <= 0 || chunk <= 0 || ft_size <= 0 || w_size <= 0 || hop <= 0 || (chunk & 3) || (ft_size & 3) || (w_size & 31) || (hop & 3))
        return st_fail_msg(h, "st_dct_analysis: sizes must be positive, chunk / ft_size / hop multiples of 4, w_size of 32");
    if (chunk + 2 * ft_size < w_size) return st_fail_msg(h, "st_dct_analysis: chunk %d too short for a %d-tap window", chunk, w_size);
    cudaStream_t s = (cudaStream_t)stream;
    const int B = batch, C = chunk, sz = ft_size, wsz = w_size;
    const int nf = (C + 2 * sz - wsz) / hop + 1;                 // Conv1d output length (padding = sz, cls_fe_dct_bases.py:116)
    const int Tp = (C + 2 * sz + hop - 1) / hop;                 // uniform-stride rows per window (>= nf; extras are dummies)
    const long Sx = (long)Tp * hop, MT = (long)B * Tp;
    const long n_x = up4(B * Sx + wsz), n_w = (long)sz * wsz, n_tmp = MT * sz;
    if (dct_workspace(h, 2 * n_x + 2 * n_w + n_tmp)) return 1;
    float *xh = h->dct_ws, *xl = xh + n_x, *wh = xl + n_x, *wl = wh + n_w, *tmp = wl + n_w;
    ST_CUDA_OK(cudaMemsetAsync(xh, 0, 2 * n_x * sizeof(float), s));
    st_launch_pad_split(x, xh, xl, B, C, sz, (int)Sx, 1.0f, s);
    st_launch_pad_split(w, wh, wl, sz, wsz, 0, wsz, 1.0f, s);
    int r = -1;
    if (h->use_tc) {
        TcOperand A{xh, xl, MT, wsz, hop}, W{wh, wl, sz, wsz, wsz};
        r = st_launch_gemm_tc(false, false, A, W, tmp, sz, (int)MT, sz, wsz, 1, 0, h->passes == 3, h->sm_count, s, h->passes);
    }
    if (r < 0) {
        GemmOperand A{xh, xl, hop}, W{wh, wl, wsz};
        ++h->simt_fallbacks;
        st_launch_gemm(true, true, A, W, tmp, sz, (int)MT, sz, wsz, 1, 0, s);
    }
    st_launch_dct_bias_unpack(tmp, bias, B, nf, Tp, sz, out, s);
    h->launches += 4;
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_dct_synthesis(st_handle* h, const float* x_ft, const float* w, int batch, int frames, int ft_size, int w_size,
                                int hop, float* wave, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x_ft || !w || !wave) return st_fail_msg(h, "st_dct_synthesis: null argument");
    if (batch <= 0 || frames <= 0 || ft_size <= 0 || w_size <= 0 || hop <= 0 || (ft_size & 31) || (w_size & 3) || (hop & 3))
        return st_fail_msg(h, "st_dct_synthesis: sizes must be positive, ft_size a multiple of 32, w_size / hop of 4");
    const int B = batch, nf = frames, sz = ft_size, wsz = w_size;
    const int C = (nf - 1) * hop + wsz - 2 * sz;                 // ConvTranspose1d length minus the trim (cls_fe_dct_bases.py:174-176)
    if (C <= 0) return st_fail_msg(h, "st_dct_synthesis: %d frames are too few for ft_size=%d", nf, sz);
    cudaStream_t s = (cudaStream_t)stream;
    const long M = (long)B * nf, n_a = M * sz, n_w = (long)sz * wsz, n_fo = M * wsz;
    if (dct_workspace(h, 2 * n_a + 2 * n_w + n_fo)) return 1;
    float *ah = h->dct_ws, *al = ah + n_a, *wh = al + n_a, *wl = wh + n_w, *fo = wl + n_w;
    st_launch_pad_split(x_ft, ah, al, (int)M, sz, 0, sz, 1.0f, s);
    st_launch_pad_split(w, wh, wl, sz, wsz, 0, wsz, 1.0f, s);
    int r = -1;
    if (h->use_tc) {
        TcOperand A{ah, al, M, sz, sz}, S{wh, wl, sz, wsz, wsz};
        r = st_launch_gemm_tc(false, true, A, S, fo, wsz, (int)M, wsz, sz, 1, 0, h->passes == 3, h->sm_count, s, h->passes);
    }
    if (r < 0) {
        GemmOperand A{ah, al, sz}, S{wh, wl, wsz};
        ++h->simt_fallbacks;
        st_launch_gemm(true, false, A, S, fo, wsz, (int)M, wsz, sz, 1, 0, s);
    }
    st_launch_dct_overlap_add(fo, B, nf, sz, wsz, hop, C, wave, s);
    h->launches += 4;
    ST_LAUNCH_OK(h);
    return 0;
}

// ---- data step in front of the path -------------------------------------------------------
extern "C" int st_compressor_4c(st_handle* h, const float* x, const double* knobs_wc, int batch, int n, double sr, float* y, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !knobs_wc || !y) return st_fail_msg(h, "st_compressor_4c: null argument");
    if (batch <= 0 || n <= 0 || !(sr > 0)) return st_fail_msg(h, "st_compressor_4c: batch, n and sr must be positive");
    if (dct_workspace(h, (long)batch * n)) return 1;
    st_launch_compressor_4c(x, knobs_wc, batch, n, sr, h->dct_ws, y, (cudaStream_t)stream);
    h->launches += 3;
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_crop_windows(st_handle* h, const float* corpus_x, const float* corpus_y, long corpus_len, const long* offsets_host,
                               const float* signs, int batch, int chunk, int y_size, float* x, float* y, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!corpus_x || !corpus_y || !offsets_host || !x || !y) return st_fail_msg(h, "st_crop_windows: null argument");
    if (batch <= 0 || chunk <= 0 || y_size <= 0 || y_size > chunk) return st_fail_msg(h, "st_crop_windows: need 0 < y_size <= chunk");
    for (int b = 0; b < batch; ++b)
        if (offsets_host[b] < 0 || offsets_host[b] + chunk > corpus_len)
            return st_fail_msg(h, "st_crop_windows: window %d (offset %ld, chunk %d) leaves the corpus of %ld samples", b, offsets_host[b],
                               chunk, corpus_len);
    ST_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long need = ((long)batch * sizeof(long) + sizeof(float) - 1) / sizeof(float);
    if (dct_workspace(h, need)) return 1;
    long* off_dev = reinterpret_cast<long*>(h->dct_ws);
    ST_CUDA_OK(cudaMemcpyAsync(off_dev, offsets_host, (size_t)batch * sizeof(long), cudaMemcpyHostToDevice, s));
    st_launch_crop_windows(corpus_x, corpus_y, off_dev, signs, batch, chunk, y_size, x, y, s);
    h->launches += 1;
    ST_LAUNCH_OK(h);
    return 0;
}

extern "C" int st_analysis(st_handle* h, const float* x, const float* w_real, const float* w_imag, int batch, float* an_real,
                           float* an_imag, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!x || !w_real || !w_imag || !an_real || !an_imag) return st_fail_msg(h, "st_analysis: null argument");
    return analysis_only(h, x, w_real, w_imag, batch, an_real, an_imag, (cudaStream_t)stream);
}

extern "C" int st_synthesis(st_handle* h, const float* real, const float* imag, const float* w_real, const float* w_imag, int batch,
                            float* wave, void* stream) {
    if (!h) return 1;
    ST_ON_DEVICE(h);
    if (!real || !imag || !w_real || !w_imag || !wave) return st_fail_msg(h, "st_synthesis: null argument");
    return synthesis_only(h, real, imag, w_real, w_imag, batch, wave, (cudaStream_t)stream);
}
