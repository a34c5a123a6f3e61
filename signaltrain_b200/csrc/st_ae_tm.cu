// Autoencoders on tcgen05 with TMEM-resident activations  --  the production path for T <= 64, OT <= 16.
//
// Reference semantics: AsymAutoEncoder.forward (nn_proc.py:77-126) for both autoencoders, the prologue (mag / phase,
// nn_proc.py:309-310) and the epilogue (skip-filter :115, phase residual :322, polar->rect :325-326) of AsymMPAEC.forward.
//
// One (batch, bin) row is one TMEM lane.  A layer  h_l = ELU(h_{l-1} W_l^T + b_l)  is ONE group of tcgen05.mma.kind::tf32
// instructions with M = 128 rows, N = layer width, the activations h_{l-1} as the A operand READ FROM TENSOR MEMORY and the
// weights (all nine layers, staged once per CTA as exact (hi, lo) tf32 pairs, K-major SWIZZLE_128B) as the B operand from
// shared memory.  Three MMAs per 8-wide k-step (a_lo*w_hi, a_hi*w_lo, a_hi*w_hi) give fp32-class products.  The epilogue
// threads (thread = row) pull the accumulator with tcgen05.ld, add the bias, apply ELU, split into (hi, lo) and write the
// next layer's A operand straight back into TMEM with tcgen05.st: no activation ever touches shared or global memory.
//
// FORWARD (ae_fwd_tm_kernel): a CTA works on one 128-row tile at a time and runs BOTH autoencoders of that tile
// concurrently (two independent chains = two sets of TMEM columns), so the tensor pipe always has the other chain's layer
// to run while one chain is in its CUDA-core epilogue.  Each chain is served by two warpgroups that split the layer's
// columns (thread = (row, column half)), which halves the epilogue latency of the chain.  The magnitude chain hands
// mag_hat to the phase chain through shared memory, so polar->rect is fused and (re, im) leave as the (hi, lo) operand of
// the synthesis GEMM.  The layer table is a compile-time constant (template on the padded widths of the two layers whose
// input depends on the geometry), so the issuing warp's descriptors are base + immediate.
#include <algorithm>

#include "st_common.cuh"
#include "st_tc_prims.cuh"

namespace {

using namespace st_tc;

constexpr int NL = ST_AE_LAYERS;
constexpr int TILE = 128;                       // rows per tile = TMEM lanes
constexpr int NSPLIT = 2;                       // warpgroups per chain (column halves)
constexpr int FWD_CHAIN_WARPS = 2 * NSPLIT * 4; // two chains
constexpr int FWD_THREADS = (FWD_CHAIN_WARPS + 1) * 32;   // + the MMA issuer warp (last warp)
constexpr int XCH_J = 16;                       // mag_hat hand-over: OT <= 16

// Compile-time layer table.  KP1: fnn_enc input (T) padded to 32 or 64; KP5: fnn_addknobs input (16 + K) padded to 16 or 24.
template <int KP1_, int KP5_>
struct Tab {
    static constexpr int KP1 = KP1_, KP5 = KP5_;
    __host__ __device__ static constexpr int n(int l) { constexpr int t[NL] = {64, 32, 16, 16, 16, 16, 32, 64, 16}; return t[l]; }   // UMMA N (OT padded to 16)
    __host__ __device__ static constexpr int kp(int l) { constexpr int t[NL] = {KP1_, 64, 32, 16, KP5_, 16, 16, 32, 64}; return t[l]; } // A columns read
    __host__ __device__ static constexpr int kslab(int l) { return (kp(l) + 31) / 32 * 32; }
    __host__ __device__ static constexpr int woff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += n(i) * kslab(i); return o; }
    __host__ __device__ static constexpr int boff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += n(i); return o; }
    static constexpr int wfloats = woff(NL);     // floats of one forward weight plane (every slab: multiple of 8 rows x 128 B)
    static constexpr int bfloats = boff(NL);
    // ---- backward.  act[l] = input of layer l (act[0] = the track, act[l] = ELU output of layer l-1); gz[l] = dLoss/d(pre-activation l)
    // data gradient of layer l:  gh[l] = gz[l] . W_l  (dn(l) columns; the knob columns of layer 4 carry none)
    __host__ __device__ static constexpr int dn(int l) { constexpr int t[NL] = {KP1_, 64, 32, 16, 16, 16, 16, 32, 64}; return t[l]; }
    __host__ __device__ static constexpr int tslab(int l) { return (n(l) + 31) / 32 * 32; }      // K extent (layer outputs) of the W^T slabs
    __host__ __device__ static constexpr int toff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += dn(i) * tslab(i); return o; }
    static constexpr int tfloats = toff(NL);
    // weight gradient of layer l on the tensor core: D_w[128 x nf] = Mop^T-tile (live rows at lane `moff`) x Nop, reduction over
    // the tile's rows.  Mop / Nop = gz[l] / act[l] in the orientation that balances the flush registers (see wg_* below).
    __host__ __device__ static constexpr int act_w(int l) { constexpr int t[NL] = {KP1_, 64, 32, 16, KP5_, 16, 16, 32, 64}; return t[l]; }   // staged act[l] features
    __host__ __device__ static constexpr bool wg_m_is_gz(int l) { constexpr bool t[NL] = {true, false, false, true, false, true, true, true, false}; return t[l]; }
    __host__ __device__ static constexpr int wg_mf(int l) { return wg_m_is_gz(l) ? n(l) : act_w(l); }
    __host__ __device__ static constexpr int wg_nf(int l) { return wg_m_is_gz(l) ? (act_w(l) + 15) / 16 * 16 : n(l); }
    // One MMA per 8-row k-step covers every product the split needs, because an SS-mode MMA of this size is bound by the 128-row
    // A tile it reads, not by N: the N operand is the stacked [N_hi ; N_lo] planes (2 NF columns: D[:, 0:NF] = m . n_hi,
    // D[:, NF:2NF] = m . n_lo), and where 2 MF <= 64 the A tile is the stacked [M_hi ; M_lo] planes as well (wg_one: the lo
    // feature rows come out MF lanes below the hi rows and are folded in when the kernel ends); the wide layers run a second
    // MMA (A = M_lo, B = N_hi) into the first NF columns instead.
    __host__ __device__ static constexpr bool wg_one(int l) { constexpr bool t[NL] = {false, false, true, true, true, true, true, false, false}; return t[l]; }
    __host__ __device__ static constexpr int wg_moff(int l) { constexpr int t[NL] = {0, 0, 64, 0, 0, 32, 64, 64, 64}; return t[l]; }   // first live TMEM lane
    __host__ __device__ static constexpr int wg_reg(int l) { constexpr int t[NL] = {0, 32, 48, 64, 80, 64, 64, 0, 32}; return t[l]; }  // first flush register
    static constexpr int wg_regs = 96;
};

__device__ __forceinline__ float elu_f(float z) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 1.4426950408889634f));
    return z > 0.f ? z : e - 1.f;
}
__device__ __forceinline__ float elu_grad(float h) { return h > 0.f ? 1.f : h + 1.f; }   // dELU/dz through the output h

// atan2 for the phase track (nn_proc.py:310), ~25 instructions, inlined: one division by folding the second range reduction
// into it (|t| <= tan(pi/8) after it), then the 4-term minimax polynomial of Cephes atanf (2 ulp).  atan2(0, 0) = 0 like
// torch / libdevice; the model never feeds infinities or NaNs here that it does not also produce in the reference.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const bool mid = mn > 0.41421356237f * mx;                 // atan(a) = pi/4 + atan((a - 1) / (a + 1))
    const float num = mid ? mn - mx : mn, den = mid ? mn + mx : mx;
    const float t = den > 0.f ? __fdividef(num, den) : 0.f;
    const float z = t * t;
    float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    float r = fmaf(t * z, p, t) + (mid ? 0.78539816339744831f : 0.f);
    if (ay > ax) r = 1.5707963267948966f - r;
    if (x < 0.f) r = 3.14159265358979323846f - r;
    return copysignf(r, y);
}
__device__ __noinline__ float2 sincos_ni(float x) {
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

// Float offset of element (row, col) of a K-major SWIZZLE_128B operand: K-block kb = col/32 is a [rows][32] slab with
// 128-byte rows; inside a row the 16-byte chunk index is XORed with (row & 7).
__device__ __forceinline__ int sw128_off(int row, int col, int slab_rows) {
    const int kb = col >> 5, c = (col >> 2) & 7, e = col & 3;
    return kb * slab_rows * 32 + row * 32 + ((c ^ (row & 7)) << 2) + e;
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "tmem_ld width");
    if constexpr (N == 8) tmem_ld8(taddr, r);
    else if constexpr (N == 16) tmem_ld16(taddr, r);
    else tmem_ld32(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "tmem_st width");
    if constexpr (N == 8) tmem_st8(taddr, r);
    else if constexpr (N == 16) tmem_st16(taddr, r);
    else tmem_st32(taddr, r);
}

// Two-term tf32 split on the integer / fp32 pipes (cvt.rna.tf32 issues on the quarter-rate conversion pipe, which the ELU's
// ex2 already loads): hi = x rounded to 10 mantissa bits, nearest with ties away, exactly as cvt.rna does it on the magnitude
// (add half an ulp to the bit pattern, clear the 13 low bits); lo = x - hi is exact in fp32 and is handed over unrounded --
// the tensor core reads its upper 19 bits, i.e. truncates it, a 2^-22 relative effect on x.
__device__ __forceinline__ void split_tf32_alu(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// Write NV values of this thread's row as a tf32 pair into the (hi, lo) A-operand columns.
template <int NV>
__device__ __forceinline__ void store_pair(uint32_t t_hi, uint32_t t_lo, const float (&v)[NV]) {
    uint32_t hi[NV], lo[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) split_tf32_alu(v[i], hi[i], lo[i]);
    tmem_st<NV>(t_hi, hi);
    tmem_st<NV>(t_lo, lo);
}

// Image of one autoencoder's operands as the kernels want them in shared memory: [hi plane][lo plane][bias].  Built once per
// step by ae_pack_kernel (the weights change every step), copied into each CTA with one-dimensional bulk copies.
template <class TB>
__host__ __device__ constexpr int image_floats() { return 2 * TB::wfloats + (TB::bfloats + 255) / 256 * 256; }   // keeps the next image 1024-byte aligned
// backward image: [W hi][W lo][W^T hi][W^T lo][bias]
template <class TB>
__host__ __device__ constexpr int bwd_image_floats() { return 2 * TB::wfloats + 2 * TB::tfloats + (TB::bfloats + 255) / 256 * 256; }

// B operand of the forward layers = W[n = out][k = in], K-major, zero padded to (n(l), 32-float K-blocks), as (hi, lo).
// BWD: the backward image (adds the data-gradient operand W^T[n = in][k = out], K-major over the layer's outputs).
template <class TB, bool BWD>
__global__ void ae_pack_kernel(AeGeom g, AeParams pm, AeParams pp, float* __restrict__ image) {
    const AeParams& p = blockIdx.y ? pp : pm;
    float* img = image + (long)blockIdx.y * (BWD ? bwd_image_floats<TB>() : image_floats<TB>());
    const int l = blockIdx.x;
    const int IN = g.in[l], OUT = g.out[l], NP = TB::n(l), KP = TB::kslab(l);
    float* whi = img + TB::woff(l);
    float* wlo = whi + TB::wfloats;
    for (int idx = threadIdx.x; idx < NP * KP; idx += blockDim.x) {
        const int o = idx / KP, i = idx - o * KP;
        const float w = (o < OUT && i < IN) ? __ldg(p.W[l] + o * IN + i) : 0.f;
        float hi, lo;
        st_split_tf32(w, hi, lo);
        const int off = sw128_off(o, i, NP);
        whi[off] = hi;
        wlo[off] = lo;
    }
    float* bias = img + 2 * TB::wfloats + (BWD ? 2 * TB::tfloats : 0) + TB::boff(l);
    for (int o = threadIdx.x; o < NP; o += blockDim.x) bias[o] = (o < OUT) ? __ldg(p.b[l] + o) : 0.f;
    if (BWD) {
        const int DN = TB::dn(l), KS = TB::tslab(l), BACK = (l == 4) ? 16 : IN;      // knob inputs carry no data gradient
        float* thi = img + 2 * TB::wfloats + TB::toff(l);
        float* tlo = thi + TB::tfloats;
        for (int idx = threadIdx.x; idx < DN * KS; idx += blockDim.x) {
            const int i = idx / KS, o = idx - i * KS;
            const float w = (o < OUT && i < BACK) ? __ldg(p.W[l] + o * IN + i) : 0.f;
            float hi, lo;
            st_split_tf32(w, hi, lo);
            const int off = sw128_off(i, o, DN);
            thi[off] = hi;
            tlo[off] = lo;
        }
    }
}

// global -> shared bulk copy (bytes: multiple of 16), completion on an mbarrier of this CTA
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// tcgen05.mma, A from TMEM, issued by a CONVERGED warp: every lane executes this, elect.sync inside picks the issuing lane.
// With compile-time TMEM addresses (the CTA owns all 512 columns, so its allocation starts at column 0) and the descriptor
// given as (uniform base + immediate), ptxas keeps every operand in uniform registers: measured 10.9 / 17.6 / 33.5 clk per
// MMA at N = 16 / 32 / 64 (scripts/ubench/umma_small_ubench.cu) -- the hardware floor N/2 -- against 48 clk for ANY N <= 64
// when a single diverged thread issues (operands broadcast lane -> uniform register per instruction).
__device__ __forceinline__ void umma_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 bd;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B (bits 32..63)
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }

// One forward layer L of chain AE: D[128 x n] = A[128 x kp] (TMEM: hi, lo) * W^T (smem: hi, lo), fully unrolled.
// dlo_img: low descriptor word of the chain's weight image (hi plane at +0, lo plane at +wfloats).
template <class TB, int AE, int L>
__device__ __forceinline__ void issue_fwd_layer(uint32_t dlo_img) {
    constexpr int n = TB::n(L), ksteps = TB::kp(L) / 8;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    constexpr uint32_t t_hi = 192u * AE, t_lo = t_hi + 64, t_d = t_hi + 128;
    constexpr uint32_t hi16 = (4u * TB::woff(L)) >> 4, lo16 = (4u * (TB::wfloats + TB::woff(L))) >> 4;   // 16-byte units
#pragma unroll
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t bo = (uint32_t)((ks >> 2) * (n * 128) + (ks & 3) * 32) >> 4;
        umma_ts_lohi(t_d, t_lo + 8 * ks, dlo_img + hi16 + bo, DESC_HI_SW128, idesc, ks > 0 ? 1u : 0u);
        umma_ts_lohi(t_d, t_hi + 8 * ks, dlo_img + lo16 + bo, DESC_HI_SW128, idesc, 1u);
        umma_ts_lohi(t_d, t_hi + 8 * ks, dlo_img + hi16 + bo, DESC_HI_SW128, idesc, 1u);
    }
}

// layer widths / bias offsets are the same for every geometry (only the K extents of layers 1 and 5 and the LIVE part of fnn_dec vary)
__constant__ int c_n[NL] = {64, 32, 16, 16, 16, 16, 32, 64, 16};
__constant__ int c_boff[NL] = {0, 64, 96, 112, 128, 144, 160, 192, 256};

// Hidden-layer epilogue of one thread: NLOC accumulator columns -> bias + ELU -> (hi, lo) A columns of the next layer.
// SAVE (backward kernel): the raw layer output is also kept in TMEM at t_save (ELU' and the weight gradients need it).
template <int NLOC, bool SAVE = false>
__device__ __forceinline__ void epi_hidden(uint32_t t_d, uint32_t t_hi, uint32_t t_lo, const float* __restrict__ bias, float* dbg,
                                           uint32_t t_save = 0, long long* probe = nullptr) {
    uint32_t r[NLOC];
    tmem_ld<NLOC>(t_d, r);
    tmem_wait_ld();
    if (probe) probe[0] = clock64();
    constexpr int CH = NLOC < 16 ? NLOC : 16;
#pragma unroll
    for (int c0 = 0; c0 < NLOC; c0 += CH) {
        float h[CH];
#pragma unroll
        for (int c4 = 0; c4 < CH; c4 += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + c4);
            h[c4 + 0] = elu_f(__uint_as_float(r[c0 + c4 + 0]) + b4.x);
            h[c4 + 1] = elu_f(__uint_as_float(r[c0 + c4 + 1]) + b4.y);
            h[c4 + 2] = elu_f(__uint_as_float(r[c0 + c4 + 2]) + b4.z);
            h[c4 + 3] = elu_f(__uint_as_float(r[c0 + c4 + 3]) + b4.w);
        }
        if (dbg) {
#pragma unroll
            for (int c = 0; c < CH; ++c) dbg[c0 + c] = h[c];
        }
        if (SAVE) {
            uint32_t raw[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) raw[c] = __float_as_uint(h[c]);
            tmem_st<CH>(t_save + c0, raw);
        }
        store_pair<CH>(t_hi + c0, t_lo + c0, h);
    }
    if (probe) probe[1] = clock64();
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
// TMEM: the CTA allocates all 512 columns, so its allocation starts at column 0 (checked): every TMEM address the issuing
// warp uses is a compile-time constant.  Chain `ae` (0: magnitude, 1: phase): A_hi at 192*ae, A_lo at +64, accumulator at +128.
//
// Per tile:  prologue (all 16 chain warps: thread = (row, quarter of the input frames) computes magnitude AND phase and feeds
// both chains; the last OT frames are kept in shared memory for the skip / residual connections)  ->  nine layers per chain
// ->  hand-over of ELU(dec) * mag_tail and ELU(dec) + phase_tail through shared memory  ->  all 16 warps share the output
// frames for sincos, the (hi, lo) split and the stores.
template <class TB>
__global__ void __launch_bounds__(FWD_THREADS, 1)
ae_fwd_tm_kernel(StDims d, const float* __restrict__ image, const float* __restrict__ spec,
                 const float* __restrict__ knobs, int B, float* __restrict__ mag_out, float* __restrict__ trk_out, float* __restrict__ mag_hat,
                 float* __restrict__ phs_hat, float* __restrict__ ri, float* __restrict__ ri_lo, float* __restrict__ dbg,
                 long long* __restrict__ timing) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];       // SWIZZLE_128B operands need 1024-byte alignment (checked below)
#ifdef ST_AE_TM_TIMING
    long long tclk = 0, treg[4] = {0, 0, 0, 0};
#endif
#ifdef ST_AE_TM_TIMING
#define ST_T0() if (timing) tclk = clock64();
#define ST_T(i) if (timing) { const long long n_ = clock64(); treg[i] += n_ - tclk; tclk = n_; }
#define ST_TL(idx) if (timing && blockIdx.x == 0 && tile == (int)gridDim.x && lane == 0) timing[idx] = clock64();
#else
#define ST_T0()
#define ST_T(i)
#define ST_TL(idx)
#endif
    constexpr int KP1 = TB::KP1;
    constexpr int IMG = image_floats<TB>();
    float* wbase = reinterpret_cast<float*>(smem_raw);     // [ae][hi plane | lo plane | bias]
    // [2: mag | phase][XCH_J][TILE]: the last OT input frames (skip / residual), updated IN PLACE by the fnn_dec stage to
    // mag_hat | phs_hat for the output stage
    float* tails = wbase + 2 * IMG;
    float* xch = tails;
    // [re | im][KP1][TILE]: the NEXT tile's input frames, fetched with cp.async while this tile runs (each thread fetches and
    // later reads only its own elements, so no barrier is involved); only when it fits beside the weights (T <= 32)
    constexpr bool PRE = KP1 == 32;
    float* pre = tails + 2 * XCH_J * TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + (PRE ? 2 * KP1 * TILE : 0));
    uint64_t* a_ready = bars;                              // [ae] epilogue -> issuer: the layer's A operand is in TMEM
    uint64_t* d_ready = bars + 2;                          // [ae] issuer -> epilogue: the accumulator is complete
    uint64_t* w_ready = bars + 4;                          // weight image landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform by construction
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int a = 0; a < 2; ++a) { mbar_init(&a_ready[a], 4 * NSPLIT); mbar_init(&d_ready[a], 1); }
        mbar_init(w_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(w_ready, 2u * IMG * 4u);
        constexpr uint32_t CHUNK = 32768;
        for (uint32_t o = 0; o < 2u * IMG * 4u; o += CHUNK)
            bulk_g2s(reinterpret_cast<uint8_t*>(wbase) + o, reinterpret_cast<const uint8_t*>(image) + o,
                     (2u * IMG * 4u - o) < CHUNK ? (2u * IMG * 4u - o) : CHUNK, w_ready);
    }
    if (warp == FWD_CHAIN_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0 || (smem_u32(smem_raw) & 1023u) != 0) __trap();   // a 512-column allocation starts at column 0

    const int BF = B * d.F;
    const int ntiles = (BF + TILE - 1) / TILE;

    if (warp == FWD_CHAIN_WARPS) {
        // ======================= MMA issuer (converged warp, fixed order: chain 0 layer l, chain 1 layer l, ...) =============
        mbar_wait_spin(w_ready, 0);                        // the async-proxy copy is visible to the async-proxy MMAs
        const uint32_t dlo0 = desc_lo_sw128(smem_u32(wbase)), dlo1 = desc_lo_sw128(smem_u32(wbase + IMG));
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#define ST_ISSUE(L)                                                        \
            mbar_wait_spin(&a_ready[0], ph);                               \
            tc_fence_after();                                              \
            ST_TL(128 + (L) * 2)                                           \
            issue_fwd_layer<TB, 0, L>(dlo0);                               \
            umma_commit_elect(&d_ready[0]);                                \
            ST_TL(128 + (L) * 2 + 1)                                       \
            mbar_wait_spin(&a_ready[1], ph);                               \
            tc_fence_after();                                              \
            ST_TL(160 + (L) * 2)                                           \
            issue_fwd_layer<TB, 1, L>(dlo1);                               \
            umma_commit_elect(&d_ready[1]);                                \
            ST_TL(160 + (L) * 2 + 1)                                       \
            ph ^= 1;
            ST_ISSUE(0) ST_ISSUE(1) ST_ISSUE(2) ST_ISSUE(3) ST_ISSUE(4) ST_ISSUE(5) ST_ISSUE(6) ST_ISSUE(7) ST_ISSUE(8)
#undef ST_ISSUE
        }
    } else {
        // ======================= chain warps =======================
        const int grp = warp >> 2;                           // 0..3: (chain, column half) in the layers, frame quarter in the prologue
        const int ae = grp / NSPLIT, half = grp % NSPLIT;
        const int q = warp & 3;                              // TMEM lane quadrant this warp may access
        const int row = 32 * q + lane;
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        const uint32_t t_hi = t_lane + 192 * ae, t_lo = t_hi + 64, t_d = t_hi + 128;
        const float* mybias = wbase + ae * IMG + 2 * TB::wfloats;
        const int tail0 = d.T - d.OT, rowstride = 2 * d.Fp;
        uint32_t ph = 0;
        constexpr int NLOC0 = KP1 / 4;
        const int pc0 = grp * NLOC0;
        auto prefetch = [&](int tile) {
            const int R = tile * TILE + row;
            const bool ok = tile < ntiles && R < BF;
            const int b = ok ? R / d.F : 0, f = ok ? R - b * d.F : 0;
            const float* sp = spec + (long)b * d.Tp * rowstride + f;
#pragma unroll
            for (int e = 0; e < NLOC0; ++e) {
                const int t = pc0 + e;
                const int bytes = (ok && t < d.T) ? 4 : 0;          // src-size 0: zero fill
                const float* src = sp + (long)(bytes ? t : 0) * rowstride;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(pre + t * TILE + row)), "l"(src), "r"(bytes) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(pre + (KP1 + t) * TILE + row)), "l"(src + d.Fp), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (PRE) prefetch(blockIdx.x);
        mbar_wait_spin(w_ready, 0);                          // biases
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int R = tile * TILE + row;
            const bool ok = R < BF;
            const int b = ok ? R / d.F : 0, f = ok ? R - b * d.F : 0;
            float* mydbg = (dbg && tile == 0) ? dbg + ((long)(ae * NL) * TILE + row) * 64 : nullptr;
#ifdef ST_AE_TM_TIMING
            const bool tl = timing && blockIdx.x == 0 && tile == (int)gridDim.x && half == 0 && q == 0 && lane == 0;
#else
            constexpr bool tl = false;
#endif
            ST_T0()
            // ---- input frames [grp * KP1/4, +KP1/4) of this row: magnitude and phase (nn_proc.py:309-310) for BOTH chains
            {
                constexpr int NLOC = KP1 / 4;
                const int c0 = grp * NLOC;
                const float* sp = spec + (long)b * d.Tp * rowstride + f;
                float re[NLOC], im[NLOC], vm[NLOC], vp[NLOC];
                if (PRE) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < NLOC; ++e) {
                        re[e] = pre[(c0 + e) * TILE + row];
                        im[e] = pre[(KP1 + c0 + e) * TILE + row];
                    }
                    prefetch(tile + gridDim.x);
                } else {
#pragma unroll
                    for (int e = 0; e < NLOC; ++e) {
                        const bool in = ok && c0 + e < d.T;
                        re[e] = in ? __ldg(sp + (long)(c0 + e) * rowstride) : 0.f;
                        im[e] = in ? __ldg(sp + (long)(c0 + e) * rowstride + d.Fp) : 0.f;
                    }
                }
#pragma unroll
                for (int e = 0; e < NLOC; ++e) {
                    const int t = c0 + e;
                    const bool in = ok && t < d.T;
                    vm[e] = sqrtf(re[e] * re[e] + im[e] * im[e]);
                    vp[e] = in ? atan2_fast(im[e], re[e] + 1e-7f) : 0.f;
                    if (in && mag_out) mag_out[((long)b * d.T + t) * d.F + f] = vm[e];
                    if (in && trk_out) {                       // both tracks, for the backward kernel's prologue
                        trk_out[((long)b * d.T + t) * d.F + f] = vm[e];
                        trk_out[(long)B * d.T * d.F + ((long)b * d.T + t) * d.F + f] = vp[e];
                    }
                    if (t >= tail0 && t < d.T) {
                        tails[(t - tail0) * TILE + row] = vm[e];
                        tails[(XCH_J + t - tail0) * TILE + row] = vp[e];
                    }
                }
                store_pair<NLOC>(t_lane + c0, t_lane + 64 + c0, vm);
                store_pair<NLOC>(t_lane + 192 + c0, t_lane + 256 + c0, vp);
            }
            tmem_wait_st();
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(FWD_CHAIN_WARPS * 32) : "memory");   // both chains' inputs complete (all quarters)
            if (lane == 0) mbar_arrive(&a_ready[ae]);
            if (tl) timing[64 + ae * 32] = clock64();
            ST_T(0)

            // ---- hidden layers: thread = (row, column half) of chain `ae`
#pragma unroll 1
            for (int l = 0; l < NL - 1; ++l) {
                mbar_wait_spin(&d_ready[ae], ph);
                ph ^= 1;
                tc_fence_after();
                if (tl) timing[64 + ae * 32 + 2 * l + 1] = clock64();
                ST_T(1)
                const int nloc = c_n[l] / NSPLIT, c0 = half * nloc;
                float* ld = mydbg ? mydbg + (long)l * TILE * 64 + c0 : nullptr;
                const float* bl = mybias + c_boff[l] + c0;
                long long* probe = (tl && ae == 0) ? timing + 192 + 4 * l : nullptr;
                if (nloc == 32) epi_hidden<32>(t_d + c0, t_hi + c0, t_lo + c0, bl, ld, 0u, probe);
                else if (nloc == 16) epi_hidden<16>(t_d + c0, t_hi + c0, t_lo + c0, bl, ld, 0u, probe);
                else epi_hidden<8>(t_d + c0, t_hi + c0, t_lo + c0, bl, ld, 0u, probe);
                if (l == 3 && half == NSPLIT - 1 && TB::KP5 > 16) {
                    // knob concat (torch.cat, nn_proc.py:95-96): columns 16..23 of fnn_addknobs' input, zero padded
                    float kv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) kv[e] = (ok && e < d.K) ? __ldg(knobs + (long)b * d.K + e) : 0.f;
                    store_pair<8>(t_hi + 16, t_lo + 16, kv);
                }
                tmem_wait_st();
                if (probe) probe[2] = clock64();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[ae]);
                if (tl) timing[64 + ae * 32 + 2 * l + 2] = clock64();
                ST_T(2)
            }

            // ---- fnn_dec: ELU(dec) * mag_tail ('sf' skip, nn_proc.py:115) / ELU(dec) + phase_tail (nn_proc.py:322) -> shared memory
            {
                constexpr int NJ = 16 / NSPLIT;
                const int j0 = half * NJ;
                mbar_wait_spin(&d_ready[ae], ph);
                ph ^= 1;
                tc_fence_after();
                if (tl) timing[64 + ae * 32 + 2 * (NL - 1) + 1] = clock64();
                ST_T(1)
                uint32_t rr[NJ];
                tmem_ld<NJ>(t_d + j0, rr);
                tmem_wait_ld();
                tc_fence_before();
                const float* bl = mybias + c_boff[NL - 1] + j0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float ev = elu_f(__uint_as_float(rr[j]) + bl[j]);
                    if (mydbg) mydbg[(long)(NL - 1) * TILE * 64 + j0 + j] = ev;
                    const float tv = tails[(ae * XCH_J + j0 + j) * TILE + row];
                    xch[(ae * XCH_J + j0 + j) * TILE + row] = ae == 0 ? ev * tv : ev + tv;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(FWD_CHAIN_WARPS * 32) : "memory");
            // ---- output stage: frames j = grp, grp + 4, ...: polar -> rect (nn_proc.py:325-326), stores coalesced along the bins
            if (ok) {
#pragma unroll 1
                for (int j = grp; j < d.OT; j += 4) {
                    const float mh = xch[j * TILE + row], phv = xch[(XCH_J + j) * TILE + row];
                    const float2 sc = sincos_ni(phv);
                    const long oo = ((long)b * d.OT + j) * d.F + f;
                    mag_hat[oo] = mh;
                    phs_hat[oo] = phv;
                    const long orr = ((long)b * d.OTp + j) * rowstride + f;
                    st_split_tf32(mh * sc.y, ri[orr], ri_lo[orr]);
                    st_split_tf32(mh * sc.x, ri[orr + d.Fp], ri_lo[orr + d.Fp]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(FWD_CHAIN_WARPS * 32) : "memory");   // xch / tails are free for the next tile
            ST_T(3)
        }
#ifdef ST_AE_TM_TIMING
        if (timing && (threadIdx.x & 127) == 0)
            for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(timing) + 8 * (1 + grp) + i, (unsigned long long)treg[i]);
#endif
    }
#undef ST_T0
#undef ST_T
#undef ST_TL
    tc_fence_before();
    __syncthreads();
    if (warp == FWD_CHAIN_WARPS) {
        tc_fence_after();
        tmem_dealloc(0u, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward (one autoencoder per CTA; recomputes the forward chain in tensor memory instead of reading saved activations)
// ---------------------------------------------------------------------------------------------------------------------
// Per 128-row tile (thread = (row, column half) in two chain warpgroups):
//   1. forward chain as above; every layer output h_l is ALSO kept raw in TMEM (256 columns);
//   2. output side (nn_proc.py:115, 322, 325-326 backwards) gives gz[8];  then for l = 8..0 one MMA group
//      gh[l] = gz[l] . W_l (A = gz[l] from TMEM, B = W_l^T from shared memory) and an epilogue gz[l-1] = gh[l] * ELU'(h_l);
//      gh[0] (+ skip / residual gradient) leaves as the track gradient;
//   3. weight gradients dW_l = gz[l]^T act[l] are a reduction over ROWS, so both operands go through shared memory: every chain
//      warp writes its 32 rows of (gz[l], act[l]) as (hi, lo) K-major SWIZZLE_128B slices ([feature][32 rows]: conflict-free
//      4-byte stores, one row per lane); a second issuing warp runs them as SS-mode MMAs (M = 128: the live feature rows sit at a
//      lane offset chosen per layer, the other lanes read neighbouring shared memory and are never looked at) into a 32-column
//      accumulator, which a third warpgroup folds into REGISTER accumulators after every layer (lane = feature, 80 registers);
//      the same warpgroup sums gz over the slice rows for the bias gradients.  Per-CTA partial gradients go to global memory
//      once, at the end (deterministic: no atomics).
constexpr int BWD_CHAIN_WARPS = 8;                  // warps 0-7: (half = w >> 2, quadrant = w & 3)
constexpr int BWD_FLUSH_WARP0 = 8;                  // warps 8-11: weight-gradient flush / bias gradient, quadrant = w & 3
constexpr int BWD_ISSUER = 12;                      // chain MMAs
constexpr int BWD_WISSUER = 13;                     // weight-gradient MMAs
constexpr int BWD_THREADS = 14 * 32;
// TMEM columns
constexpr uint32_t TC_AH = 256, TC_AL = 320, TC_D = 384, TC_DW = 448;
__host__ __device__ constexpr uint32_t tc_act(int l) { constexpr uint32_t t[NL] = {0, 0, 64, 96, 112, 128, 144, 160, 192}; return t[l]; }   // act[1..8]
// Accumulator columns of layer L's weight gradient (2 NF columns).  A layer may only use columns that are dead in EVERY quadrant
// when its first MMA (which overwrites all 128 lanes) is issued, i.e. after the issuer has seen all of layer L + 1 staged: the
// saved outputs act[j], j >= L + 1 (their last readers ran before the layer-(L + 1) slices were signalled), the spare columns, and
// accumulators of layers >= L + 3 (a flush warp lags the issuer by at most two layers).
__host__ __device__ constexpr uint32_t tc_dw(int l) { constexpr uint32_t t[NL] = {128, 192, 96, 448, 128, 480, 160, 192, 448}; return t[l]; }

// Shared-memory map of the backward kernel.  The staging area of the weight gradients is four 32 KB regions (one per quadrant
// of rows) laid over the forward weights AND the dedicated staging bytes behind them: nobody reads the forward weights during
// the backward half of a tile, so they are copied in again (one 80 KB bulk copy per tile, issued by the weight-gradient issuer
// once the tile's last slices are consumed) while the next tile's prologue runs.  The small arrays sit in front so that an A
// tile whose live rows start at a TMEM lane offset (its descriptor starts up to 8 KB before the plane) stays inside the window.
template <class TB>
struct BwdSmem {
    static constexpr uint32_t BIAS = 0, VTAIL = 4u * ((TB::bfloats + 255) / 256 * 256);
    static constexpr uint32_t W_HI = VTAIL + 4u * XCH_J * TILE, W_LO = W_HI + 4u * TB::wfloats, STAGE = W_LO + 4u * TB::wfloats;
    static constexpr uint32_t REGION0 = W_HI, REGION_BYTES = 32768;                 // quadrant q stages into [REGION0 + q * 32 KB, + 32 KB)
    static constexpr uint32_t WT_HI = REGION0 + 4u * REGION_BYTES, WT_LO = WT_HI + 4u * TB::tfloats;
    static constexpr uint32_t BARS = WT_LO + 4u * TB::tfloats;
    static constexpr uint32_t TOTAL = BARS + 256;
    static constexpr uint32_t RELOAD_BYTES = 8u * TB::wfloats;                      // the whole forward image
    static_assert(TOTAL <= 227u * 1024u, "backward tile does not fit shared memory");
    static_assert(STAGE <= WT_HI && REGION0 >= 64u * 128u, "staging regions must cover the forward weights and leave room below");
    static_assert(W_HI % 1024 == 0 && W_LO % 1024 == 0 && WT_HI % 1024 == 0 && WT_LO % 1024 == 0, "operand planes must be 1024-byte aligned");
};

__device__ __forceinline__ void umma_ss_lohi(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 ad, bd;\n\t"
        "mov.b64 ad, {%1, %3};\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// One layer's MMAs of the backward kernel's single chain: D[128 x n] = A (hi | lo columns of TMEM, `ksteps` 8-wide k-steps) x B
// (K-major SWIZZLE_128B slabs of 32 k, `n` rows each, hi plane at b_hi16, lo plane at b_lo16, in descriptor units of 16 bytes).
// The k-step loop is rolled (the kernel is bound by instruction fetch); its trip count and every stride are
// compile-time constants, the running addresses live in uniform registers.
template <int N, int KSTEPS>
__device__ __forceinline__ void issue_chain_layer(uint32_t b_hi16, uint32_t b_lo16) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    constexpr uint32_t slab16 = (uint32_t)(N * 128) >> 4;            // next 32-wide K slab
    uint32_t acc = 0u;
#pragma unroll 1
    for (int k0 = 0; k0 < KSTEPS; k0 += (KSTEPS < 4 ? KSTEPS : 4)) {
#pragma unroll
        for (int ks = 0; ks < (KSTEPS < 4 ? KSTEPS : 4); ++ks) {
            const uint32_t ta = (uint32_t)(8 * (k0 + ks));
            umma_ts_lohi(TC_D, TC_AL + ta, b_hi16 + 2 * ks, DESC_HI_SW128, idesc, ks > 0 ? 1u : acc);
            umma_ts_lohi(TC_D, TC_AH + ta, b_lo16 + 2 * ks, DESC_HI_SW128, idesc, 1u);
            umma_ts_lohi(TC_D, TC_AH + ta, b_hi16 + 2 * ks, DESC_HI_SW128, idesc, 1u);
        }
        b_hi16 += slab16;
        b_lo16 += slab16;
        acc = 1u;
    }
}
// forward layer L (recompute)
template <class TB, int L>
__device__ __forceinline__ void issue_refwd_layer(uint32_t dlo_base) {
    issue_chain_layer<TB::n(L), TB::kp(L) / 8>(dlo_base + ((BwdSmem<TB>::W_HI + 4u * TB::woff(L)) >> 4),
                                                dlo_base + ((BwdSmem<TB>::W_LO + 4u * TB::woff(L)) >> 4));
}
// data gradient of layer L:  gh[L] (dn columns) = gz[L] (n columns, TMEM) . W_L   (B = W_L^T [dn rows][n], K-major)
template <class TB, int L>
__device__ __forceinline__ void issue_dgrad_layer(uint32_t dlo_base) {
    issue_chain_layer<TB::dn(L), TB::n(L) / 8>(dlo_base + ((BwdSmem<TB>::WT_HI + 4u * TB::toff(L)) >> 4),
                                                dlo_base + ((BwdSmem<TB>::WT_LO + 4u * TB::toff(L)) >> 4));
}

// Staging geometry of layer L's weight-gradient slice (one quadrant = 32 rows) inside the quadrant's 32 KB region, in units of
// one feature = 256 bytes (a 128-byte K-major row of the hi plane + one of the lo plane; the two planes of an operand are
// stacked [hi ; lo], which is what the single-MMA formulation reads).  act[L] is staged ONE LAYER AHEAD (while layer L + 1's
// MMAs run), so its block must not touch layer L + 1's blocks: even layers keep act at the bottom of the region, odd layers at
// the top; gz[L] (written once layer L + 1 is consumed) only has to avoid act[L] and act[L - 1].
__host__ __device__ constexpr int stg_act_row(int l) { constexpr int t[NL] = {0, 64, 0, 112, 0, 112, 0, 96, 0}; return t[l]; }
__host__ __device__ constexpr int stg_gz_row(int l) { constexpr int t[NL] = {32, 32, 32, 32, 32, 32, 16, 32, 64}; return t[l]; }
template <class TB, int L>
struct Stg {
    static constexpr int MF = TB::wg_mf(L), NF = TB::wg_nf(L);
    static constexpr bool M_IS_GZ = TB::wg_m_is_gz(L);
    static constexpr int GZ_W = TB::n(L), ACT_W = M_IS_GZ ? NF : MF;
    static constexpr uint32_t GZ_HI = stg_gz_row(L) * 256u, GZ_LO = GZ_HI + GZ_W * 128u;
    static constexpr uint32_t ACT_HI = stg_act_row(L) * 256u, ACT_LO = ACT_HI + ACT_W * 128u;
    static constexpr uint32_t M_HI = M_IS_GZ ? GZ_HI : ACT_HI, M_LO = M_IS_GZ ? GZ_LO : ACT_LO;
    static constexpr uint32_t N_HI = M_IS_GZ ? ACT_HI : GZ_HI, N_LO = M_IS_GZ ? ACT_LO : GZ_LO;
    __host__ __device__ static constexpr uint32_t buf(int q) { return BwdSmem<TB>::REGION0 + (uint32_t)q * BwdSmem<TB>::REGION_BYTES; }
    static_assert(stg_act_row(L) + ACT_W <= 128 && stg_gz_row(L) + GZ_W <= 128, "block outside the region");
    static_assert(stg_act_row(L) + ACT_W <= stg_gz_row(L) || stg_gz_row(L) + GZ_W <= stg_act_row(L), "act[L] and gz[L] overlap");
    static_assert((M_IS_GZ ? MF : NF) == GZ_W, "gz width");
};
// act[L - 1] (staged ahead) must not touch layer L's blocks, and gz[L - 1] must not touch act[L - 2]
template <class TB, int L>
__host__ __device__ constexpr bool stg_ahead_ok() {
    if constexpr (L == 0) return true;
    else {
        using A = Stg<TB, L>;
        using P = Stg<TB, L - 1>;
        constexpr int a0 = stg_act_row(L - 1), a1 = a0 + P::ACT_W;
        constexpr bool vs_act = a1 <= stg_act_row(L) || stg_act_row(L) + A::ACT_W <= a0;
        constexpr bool vs_gz = a1 <= stg_gz_row(L) || stg_gz_row(L) + A::GZ_W <= a0;
        constexpr bool gz_vs_next_act = stg_gz_row(L) + A::GZ_W <= a0 || a1 <= stg_gz_row(L);
        return vs_act && vs_gz && gz_vs_next_act && stg_ahead_ok<TB, L - 1>();
    }
}

// weight-gradient MMAs of layer L for one 32-row slice (4 k-steps); `dlo_buf` = descriptor low word of the quadrant's buffer.
// The quadrant is a run-time value (the loop over quadrants stays rolled: the kernel's code must stay small, its instruction
// fetch is what bounds it), everything else is an immediate.
template <class TB, int L>
__device__ __forceinline__ void issue_wgrad_slice(uint32_t dlo_buf, uint32_t first) {
    using S = Stg<TB, L>;
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * S::NF >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(S::NF >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    constexpr uint32_t moff = (uint32_t)TB::wg_moff(L) * 128u;
    static_assert(S::buf(0) + S::M_HI + 0u >= moff + 0u, "the A tile must start inside shared memory");
    static_assert(S::M_LO == S::M_HI + S::MF * 128u && S::N_LO == S::N_HI + S::NF * 128u, "the (hi, lo) planes must be stacked");
    static_assert(S::MF % 8 == 0 && S::NF % 8 == 0, "planes must start on a swizzle period");
    const uint32_t a_hi = dlo_buf + (S::M_HI >> 4) - (moff >> 4), a_lo = dlo_buf + (S::M_LO >> 4) - (moff >> 4), b_hi = dlo_buf + (S::N_HI >> 4);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        // A = [M_hi ; M_lo ; ...] (the lo rows count only for wg_one layers), B = [N_hi ; N_lo]
        umma_ss_lohi(tc_dw(L), a_hi + 2 * ks, b_hi + 2 * ks, DESC_HI_SW128, idesc2, ks > 0 ? 1u : first);
        if constexpr (!TB::wg_one(L))
            umma_ss_lohi(tc_dw(L), a_lo + 2 * ks, b_hi + 2 * ks, DESC_HI_SW128, idesc1, 1u);
    }
}

// One value of this lane's row into a [feature][32 rows] K-major SWIZZLE_128B plane: feature row f is 128 bytes, the lane's
// 4-byte slot sits in 16-byte chunk (lane >> 2) ^ (f & 7).  lane_off = ((lane >> 2) << 4) | ((lane & 3) << 2).
__device__ __forceinline__ void stage_put(uint8_t* plane_hi, uint8_t* plane_lo, int f, uint32_t lane_off, uint32_t hi, uint32_t lo) {
    const uint32_t off = (uint32_t)f * 128u + (lane_off ^ ((uint32_t)(f & 7) << 4));
    *reinterpret_cast<uint32_t*>(plane_hi + off) = hi;
    *reinterpret_cast<uint32_t*>(plane_lo + off) = lo;
}

// gz[L] of this thread (its NLOC = n(L)/2 columns) is ready: (1) A operand of layer L's data-gradient MMAs -> TMEM, signal the
// chain issuer; (2) ELU'(act[L]) of the thread's own columns for the NEXT epilogue (gz[L-1] = gh[L] * ELU'), read here so
// that nobody touches act[L] once the layer-L slices have been signalled.
template <class TB, int L>
__device__ __forceinline__ void bwd_handoff(uint32_t t_lane, int half, int lane, const float (&gz)[TB::n(L) / 2], uint64_t* a_ready,
                                            float (&eg)[L > 0 ? TB::n(L > 0 ? L - 1 : 0) / 2 : 1], uint32_t (&ghi)[TB::n(L) / 2],
                                            uint32_t (&glo)[TB::n(L) / 2], long long* hb = nullptr) {
    long long hc = hb ? clock64() : 0;
#define HB(k) if (hb) { const long long n_ = clock64(); hb[k] += n_ - hc; hc = n_; }
    constexpr int NLOC = TB::n(L) / 2;
    const int c0 = half * NLOC;
#pragma unroll
    for (int c = 0; c < NLOC; ++c) split_tf32_alu(gz[c], ghi[c], glo[c]);
    tmem_st<NLOC>(t_lane + TC_AH + c0, ghi);
    tmem_st<NLOC>(t_lane + TC_AL + c0, glo);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready);
    HB(0)
    if constexpr (L > 0) {
        constexpr int AW = TB::n(L - 1) / 2;                 // this thread's share of act[L] (width n(L-1))
        uint32_t hv[AW];
        tmem_ld<AW>(t_lane + tc_act(L) + half * AW, hv);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < AW; ++c) eg[c] = elu_grad(__uint_as_float(hv[c]));
    }
    HB(3)
#undef HB
}

// gz[L] of this thread (already split) -> this quadrant's slice of layer L's weight gradient in shared memory, signal the
// weight-gradient issuer.  Follows the hand-over above (the data-gradient MMAs of layer L are in flight meanwhile).  act[L] is
// staged by the quadrant's flush warp, except the track (act[0]), which the chain warps hold in registers.
// (Staging one chain stage late, to give the buffer longer to drain, was measured slower: 372 vs 354 us.)
template <class TB, int L>
__device__ __forceinline__ void bwd_stage_gz(uint8_t* smem_raw, uint32_t lane_off, int half, int q, int it, int lane,
                                             const uint32_t (&ghi)[TB::n(L) / 2], const uint32_t (&glo)[TB::n(L) / 2], uint64_t* full,
                                             uint64_t* freeb, const float (&vkeep)[16], long long* hb = nullptr) {
    using S = Stg<TB, L>;
    long long hc = hb ? clock64() : 0;
#define HB(k) if (hb) { const long long n_ = clock64(); hb[k] += n_ - hc; hc = n_; }
    constexpr int NLOC = TB::n(L) / 2;
    const int c0 = half * NLOC;
    // the quadrant's own buffer: its previous slice must have been consumed (weight-gradient MMAs + bias-gradient pass)
    const int i = it * NL + (NL - 1 - L);                    // this quadrant's fill count = phase index of its barriers
    if (i > 0) mbar_wait_spin(&freeb[q], (uint32_t)((i - 1) & 1));
    HB(1)
    uint8_t* buf = smem_raw + BwdSmem<TB>::REGION0 + (uint32_t)q * BwdSmem<TB>::REGION_BYTES;
#pragma unroll
    for (int c = 0; c < NLOC; ++c) stage_put(buf + S::GZ_HI, buf + S::GZ_LO, c0 + c, lane_off, ghi[c], glo[c]);
    if constexpr (L == 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            uint32_t hi, lo;
            split_tf32_alu(vkeep[e], hi, lo);
            stage_put(buf + S::ACT_HI, buf + S::ACT_LO, 16 * half + e, lane_off, hi, lo);
        }
    }
    HB(2)
    fence_async_smem();                                      // generic-proxy writes -> the MMAs' async-proxy reads
    __syncwarp();
    if (lane == 0) mbar_arrive(&full[q]);
    HB(4)
#undef HB
}

// Sum over the 32 rows of a slice of feature f of a staged (hi, lo) plane pair (bias gradient; not inlined: code size).
__device__ __noinline__ float slice_row_sum(const uint8_t* plane_hi, const uint8_t* plane_lo, int f) {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t o = (uint32_t)f * 128u + ((uint32_t)(c ^ (f & 7)) << 4);
        const float4 h4 = *reinterpret_cast<const float4*>(plane_hi + o);
        const float4 l4 = *reinterpret_cast<const float4*>(plane_lo + o);
        sum += ((h4.x + l4.x) + (h4.y + l4.y)) + ((h4.z + l4.z) + (h4.w + l4.w));
    }
    return sum;
}
// Flush warp: this quadrant's 32 rows (lane = row) of `width` saved activations starting at TMEM address t_src -> (hi, lo)
// planes of a slice.  One rolled loop shared by every layer (not inlined: code size).
__device__ __noinline__ void stage_act_rows(uint8_t* plane_hi, uint8_t* plane_lo, uint32_t t_src, int width, uint32_t lane_off) {
#pragma unroll 1
    for (int c0 = 0; c0 < width; c0 += 16) {
        uint32_t hv[16];
        tmem_ld<16>(t_src + c0, hv);
        tmem_wait_ld();
        uint8_t* ph = plane_hi + c0 * 128;
        uint8_t* pl = plane_lo + c0 * 128;
#pragma unroll
        for (int c = 0; c < 16; ++c) {                       // (c0 + c) & 7 == c & 7: the swizzle term is an immediate
            uint32_t hi, lo;
            split_tf32_alu(__uint_as_float(hv[c]), hi, lo);
            stage_put(ph, pl, c, lane_off, hi, lo);
        }
    }
}
// act[L] (the saved output of layer L - 1; knobs appended for layer 4) -> layer L's slice
template <class TB, int L>
__device__ __forceinline__ void stage_act(uint8_t* buf, uint32_t t_lq, uint32_t lane_off, const float* knobs, long knob_off, int nk) {
    using S = Stg<TB, L>;
    static_assert(L > 0, "the track is staged by the chain warps");
    stage_act_rows(buf + S::ACT_HI, buf + S::ACT_LO, t_lq + tc_act(L), TB::n(L - 1), lane_off);
    if constexpr (L == 4 && TB::KP5 > 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float kv = e < nk ? __ldg(knobs + knob_off + e) : 0.f;
            uint32_t hi, lo;
            split_tf32_alu(kv, hi, lo);
            stage_put(buf + S::ACT_HI, buf + S::ACT_LO, 16 + e, lane_off, hi, lo);
        }
    }
}

struct BwdArgs {
    const float* image;      // [2 autoencoders] backward images
    const float* trk;        // [2][B][T][F] magnitude | phase tracks written by the forward kernel
    const float* knobs;
    const float* mag_hat;    // forward outputs
    const float* phs_hat;
    const float* g_ri;       // dLoss/d(re | im) of the synthesis input, [(b, j)][2 Fp]
    const float* g_mag_hat;  // nullable: direct gradient on mag_hat (the L1 term of the loss)
    float* g_track;          // [2][B][T][F] track gradients (magnitude | phase)
    float* partials;         // [(slot, autoencoder)][flat_total] per-CTA weight / bias gradient sums
    float* dbg;              // nullable: [18][128][64] layer outputs then gz of tile 0 (tests)
    long long* timing;       // nullable (ST_AE_TM_TIMING builds): clock buckets of the chain warps
    int B;
};

// DBG: the instantiation the tests use to dump tile 0's chain (kept out of the production kernel: code size)
template <class TB, bool DBG>
__global__ void __launch_bounds__(BWD_THREADS, 1)
ae_bwd_tm_kernel(StDims d, AeGeom g, BwdArgs a) {
    static_assert(TB::KP1 == 32, "backward kernel: T <= 32");
    static_assert(stg_ahead_ok<TB, NL - 1>(), "act[L - 1] is staged while layer L's slice is live: the blocks must be disjoint");
    using SM = BwdSmem<TB>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* bias = reinterpret_cast<float*>(smem_raw + SM::BIAS);
    float* vtail = reinterpret_cast<float*>(smem_raw + SM::VTAIL);          // [XCH_J][TILE]: tail of the track, then skip / residual gradient
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SM::BARS);
    uint64_t* a_ready = bars;            // chain warps -> chain issuer (count 8)
    uint64_t* d_ready = bars + 1;        // chain issuer -> chain warps
    uint64_t* w_ready = bars + 2;        // weight image landed
    uint64_t* full = bars + 3;           // [4] gz[L] of the quadrant's slice written (count 2: the two column halves)
    uint64_t* freeb = bars + 7;          // [4] slice consumed (count 2: weight-gradient MMAs + bias-gradient reader)
    uint64_t* flushed = bars + 11;       // flush warps -> chain warps: every accumulator of the tile has been read (count 4), once per tile
    uint64_t* fwd_done = bars + 12;      // chain warps -> flush warps: the tile's forward chain is complete (count 8), once per tile
    uint64_t* w_reload = bars + 13;      // the forward weights the staging regions overwrote are back (once per tile)
    uint64_t* dw_ready = bars + 16;      // [4] ring: weight-gradient issuer -> flush warps: the layer's accumulator is complete (the
                                         // issuer may run two layers ahead of a flush warp; a parity wait cannot tell phase k from k + 2)
    uint64_t* afull = bars + 20;         // [4][2] act[L] of the quadrant's slice written by its flush warp (count 1), ring over the
                                         // parity of L: the flush warp stages one layer ahead of the issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int ae = blockIdx.x & 1, slot = blockIdx.x >> 1, nslot = gridDim.x >> 1;
    constexpr int IMGF = bwd_image_floats<TB>();
    if (threadIdx.x == 0) {
        mbar_init(a_ready, BWD_CHAIN_WARPS); mbar_init(d_ready, 1); mbar_init(w_ready, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 2); mbar_init(&freeb[i], 2); }
        for (int i = 0; i < 8; ++i) mbar_init(&afull[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&dw_ready[i], 1);
        mbar_init(flushed, 4); mbar_init(w_reload, 1); mbar_init(fwd_done, BWD_CHAIN_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t* img = reinterpret_cast<const uint8_t*>(a.image + (long)ae * IMGF);
        constexpr uint32_t WB = 8u * TB::wfloats, TBY = 8u * TB::tfloats, BB = 4u * TB::bfloats;
        static_assert(BB % 16 == 0, "bulk copies move multiples of 16 bytes");
        mbar_expect_tx(w_ready, WB + TBY + BB);
        constexpr uint32_t CHUNK = 32768;
        for (uint32_t o = 0; o < WB; o += CHUNK) bulk_g2s(smem_raw + SM::W_HI + o, img + o, (WB - o) < CHUNK ? (WB - o) : CHUNK, w_ready);
        for (uint32_t o = 0; o < TBY; o += CHUNK) bulk_g2s(smem_raw + SM::WT_HI + o, img + WB + o, (TBY - o) < CHUNK ? (TBY - o) : CHUNK, w_ready);
        bulk_g2s(smem_raw + SM::BIAS, img + WB + TBY, BB, w_ready);
    }
    if (warp == BWD_ISSUER) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0 || (smem_u32(smem_raw) & 1023u) != 0) __trap();

    const int BF = a.B * d.F;
    const int ntiles = (BF + TILE - 1) / TILE;
    const uint32_t dlo_base = desc_lo_sw128(smem_u32(smem_raw));

    if (warp == BWD_ISSUER) {
        // ======================= chain MMAs: forward layers 0..8, then data gradients of layers 8..0 =======================
        mbar_wait_spin(w_ready, 0);
        uint32_t ph = 0;
        int it = 0;
        for (int tile = slot; tile < ntiles; tile += nslot, ++it) {
            if (it > 0) mbar_wait_spin(w_reload, (uint32_t)((it - 1) & 1));
#define ST_CH(CALL) mbar_wait_spin(a_ready, ph); ph ^= 1; tc_fence_after(); CALL; umma_commit_elect(d_ready);
            ST_CH((issue_refwd_layer<TB, 0>(dlo_base))) ST_CH((issue_refwd_layer<TB, 1>(dlo_base))) ST_CH((issue_refwd_layer<TB, 2>(dlo_base)))
            ST_CH((issue_refwd_layer<TB, 3>(dlo_base))) ST_CH((issue_refwd_layer<TB, 4>(dlo_base))) ST_CH((issue_refwd_layer<TB, 5>(dlo_base)))
            ST_CH((issue_refwd_layer<TB, 6>(dlo_base))) ST_CH((issue_refwd_layer<TB, 7>(dlo_base))) ST_CH((issue_refwd_layer<TB, 8>(dlo_base)))
            ST_CH((issue_dgrad_layer<TB, 8>(dlo_base))) ST_CH((issue_dgrad_layer<TB, 7>(dlo_base))) ST_CH((issue_dgrad_layer<TB, 6>(dlo_base)))
            ST_CH((issue_dgrad_layer<TB, 5>(dlo_base))) ST_CH((issue_dgrad_layer<TB, 4>(dlo_base))) ST_CH((issue_dgrad_layer<TB, 3>(dlo_base)))
            ST_CH((issue_dgrad_layer<TB, 2>(dlo_base))) ST_CH((issue_dgrad_layer<TB, 1>(dlo_base))) ST_CH((issue_dgrad_layer<TB, 0>(dlo_base)))
#undef ST_CH
        }
    } else if (warp == BWD_WISSUER) {
        // ======================= weight-gradient MMAs: per layer 8..0, the four 32-row slices into one accumulator ===========
        int nl = 0, it = 0;                                  // layers issued so far = phase index of full[] / free[] / dw_*
        for (int tile = slot; tile < ntiles; tile += nslot) {
#define ST_WL(L)                                                                                            \
            _Pragma("unroll 1")                                                                             \
            for (int qq = 0; qq < 4; ++qq) {                                                                \
                mbar_wait_spin(&afull[2 * qq + ((L) & 1)], (uint32_t)(((L) & 1 ? 4 * it + (7 - (L)) / 2 : 5 * it + (8 - (L)) / 2) & 1)); \
                mbar_wait_spin(&full[qq], (uint32_t)(nl & 1));                                              \
                tc_fence_after();                                                                           \
                issue_wgrad_slice<TB, L>(dlo_base + ((SM::REGION0 + (uint32_t)qq * SM::REGION_BYTES) >> 4), qq > 0 ? 1u : 0u); \
                umma_commit_elect(&freeb[qq]);                                                              \
            }                                                                                               \
            umma_commit_elect(&dw_ready[nl & 3]);                                                           \
            ++nl;
            ST_WL(8) ST_WL(7) ST_WL(6) ST_WL(5) ST_WL(4) ST_WL(3) ST_WL(2) ST_WL(1) ST_WL(0)
            if (tile + nslot < ntiles) {
                // the tile's last slices have been consumed: put the forward weights the staging regions overwrote back
                mbar_wait_spin(&freeb[0], (uint32_t)((nl - 1) & 1));
                mbar_wait_spin(&freeb[1], (uint32_t)((nl - 1) & 1));
                mbar_wait_spin(&freeb[2], (uint32_t)((nl - 1) & 1));
                fence_async_smem();
                if (lane == 0) {
                    const uint8_t* img = reinterpret_cast<const uint8_t*>(a.image + (long)ae * IMGF);
                    mbar_expect_tx(w_reload, SM::RELOAD_BYTES);
                    constexpr uint32_t CHUNK = 32768;
                    for (uint32_t o = 0; o < SM::RELOAD_BYTES; o += CHUNK)
                        bulk_g2s(smem_raw + SM::W_HI + o, img + o, (SM::RELOAD_BYTES - o) < CHUNK ? (SM::RELOAD_BYTES - o) : CHUNK, w_reload);
                }
                __syncwarp();
            }
            ++it;
#undef ST_WL
        }
    } else if (warp >= BWD_FLUSH_WARP0) {
        // ======================= flush warps: stage act[L], bias gradients from the staged gz, accumulate D_w into registers ====
        const int q = warp & 3;
        const int glane = 32 * q + lane;                     // TMEM lane = feature row of the M operand (+ its lane offset)
        const uint32_t t_lq = (uint32_t)(32 * q) << 16;
        const uint32_t lane_off = ((uint32_t)(lane >> 2) << 4) | ((uint32_t)(lane & 3) << 2);
        uint8_t* const qbuf = smem_raw + SM::REGION0 + (uint32_t)q * SM::REGION_BYTES;
        float acc[TB::wg_regs];
        float dbacc[11];                                     // bias-gradient sums: lane = feature, one register per 32 features of a layer
#pragma unroll
        for (int i = 0; i < TB::wg_regs; ++i) acc[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 11; ++i) dbacc[i] = 0.f;
        int nl = 0, it = 0;
        for (int tile = slot; tile < ntiles; tile += nslot, ++it) {
            const int R = tile * TILE + glane;
            const bool ok = R < BF;
            const long knob_off = (long)(ok ? R / d.F : 0) * d.K;
            const int nk = ok ? d.K : 0;
            // act[8] of this tile: the forward chain must be complete (which also means the forward weights that quadrants 2, 3
            // stage over are no longer being read), and the buffer's last slice of the previous tile consumed
            mbar_wait_spin(fwd_done, (uint32_t)(it & 1));
            tc_fence_after();
            if (nl > 0) mbar_wait_spin(&freeb[q], (uint32_t)((nl - 1) & 1));
            stage_act<TB, NL - 1>(qbuf, t_lq, lane_off, a.knobs, knob_off, nk);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[2 * q]);
#define ST_FL(L, DBSLOT)                                                                                              \
            {                                                                                                          \
                using S = Stg<TB, L>;                                                                                  \
                mbar_wait_spin(&full[q], (uint32_t)(nl & 1));                                                          \
                _Pragma("unroll")                                                                                      \
                for (int p = 0; p < (TB::n(L) + 31) / 32; ++p) {                                                       \
                    const int f = 32 * p + lane;                                                                       \
                    if (f < TB::n(L)) dbacc[DBSLOT + p] += slice_row_sum(qbuf + S::GZ_HI, qbuf + S::GZ_LO, f);         \
                }                                                                                                      \
                __syncwarp();                                                                                          \
                if (lane == 0) mbar_arrive(&freeb[q]);                                                                 \
                /* the next layer's act, one layer ahead: its block is disjoint from this layer's, and what it overwrites */ \
                /* (layer L + 1) was consumed before the chain warps could write the gz this warp has just read        */ \
                if constexpr ((L) > 1) {                                                                               \
                    stage_act<TB, ((L) > 1 ? (L) - 1 : 1)>(qbuf, t_lq, lane_off, a.knobs, knob_off, nk);               \
                    fence_async_smem();                                                                                \
                    __syncwarp();                                                                                      \
                }                                                                                                      \
                if constexpr ((L) > 0) { if (lane == 0) mbar_arrive(&afull[2 * q + (((L) - 1) & 1)]); }                \
                mbar_wait_spin(&dw_ready[nl & 3], (uint32_t)(nl >> 2) & 1u);                                           \
                tc_fence_after();                                                                                      \
                constexpr int MO = TB::wg_moff(L), MF = S::MF, NF = S::NF, RG = TB::wg_reg(L);                         \
                constexpr int LIVE = TB::wg_one(L) ? 2 * MF : MF;                                                      \
                if (32 * q < MO + LIVE && 32 * q + 32 > MO) {                                                          \
                    const bool hi_row = glane >= MO && glane < MO + MF;                                                \
                    const bool lo_row = TB::wg_one(L) && glane >= MO + MF && glane < MO + 2 * MF;                      \
                    _Pragma("unroll")                                                                                  \
                    for (int c0 = 0; c0 < NF; c0 += 8) {                                                               \
                        uint32_t v[8], w[8];                                                                           \
                        tmem_ld<8>(t_lq + tc_dw(L) + c0, v);                                                           \
                        tmem_ld<8>(t_lq + tc_dw(L) + NF + c0, w);                                                      \
                        tmem_wait_ld();                                                                                \
                        if (hi_row) {                                                                                  \
                            _Pragma("unroll")                                                                          \
                            for (int i = 0; i < 8; ++i) acc[RG + c0 + i] += __uint_as_float(v[i]) + __uint_as_float(w[i]); \
                        } else if (lo_row) {                                                                           \
                            _Pragma("unroll")                                                                          \
                            for (int i = 0; i < 8; ++i) acc[RG + c0 + i] += __uint_as_float(v[i]);                     \
                        }                                                                                              \
                    }                                                                                                  \
                }                                                                                                      \
                tc_fence_before();                                                                                     \
                ++nl;                                                                                                  \
            }
            ST_FL(8, 10) ST_FL(7, 8) ST_FL(6, 7) ST_FL(5, 6) ST_FL(4, 5) ST_FL(3, 4) ST_FL(2, 3) ST_FL(1, 2) ST_FL(0, 0)
#undef ST_FL
            __syncwarp();
            if (lane == 0) mbar_arrive(flushed);             // the next tile's forward may overwrite the accumulator columns
        }
        // ---- per-CTA partial gradients -> global memory (layout: [W_0, b_0, W_1, b_1, ...] like ae_grad_reduce_kernel reads it)
        float* part = a.partials + ((long)slot * 2 + ae) * g.flat_total;
#define ST_WR(L, LO)                                                                                                   \
        if (!(LO) || TB::wg_one(L)) {                                                                                  \
            constexpr int MF = TB::wg_mf(L), NF = TB::wg_nf(L), RG = TB::wg_reg(L);                                    \
            constexpr int MO = TB::wg_moff(L) + ((LO) ? MF : 0);                                                       \
            const int IN = g.in[L], OUT = g.out[L];                                                                    \
            float* W = part + g.flat_off[L];                                                                           \
            if (glane >= MO && glane < MO + MF) {                                                                      \
                const int m = glane - MO;                                                                              \
                _Pragma("unroll")                                                                                      \
                for (int i = 0; i < NF; ++i) {                                                                         \
                    const int o = TB::wg_m_is_gz(L) ? m : i, k = TB::wg_m_is_gz(L) ? i : m;                            \
                    if (o < OUT && k < IN) { if (LO) W[o * IN + k] += acc[RG + i]; else W[o * IN + k] = acc[RG + i]; } \
                }                                                                                                      \
            }                                                                                                          \
        }
        ST_WR(0, false) ST_WR(1, false) ST_WR(2, false) ST_WR(3, false) ST_WR(4, false) ST_WR(5, false) ST_WR(6, false) ST_WR(7, false) ST_WR(8, false)
        __threadfence_block();
        asm volatile("bar.sync 2, 128;" ::: "memory");                        // hi rows written; every flush warp is past its last slice
        // the (m_lo . n_hi) sums of the single-MMA layers sit MF lanes below their hi rows: fold them in
        ST_WR(2, true) ST_WR(3, true) ST_WR(4, true) ST_WR(5, true) ST_WR(6, true)
#undef ST_WR
        // bias gradients: the four flush warps hold the sums of their own quadrant's rows -> add through shared memory
        float* dbs = reinterpret_cast<float*>(smem_raw + SM::STAGE);          // staging is idle now: [4][11][32]
        static_assert(SM::STAGE + 4u * 11u * 32u * 4u <= SM::WT_HI, "bias-gradient exchange must fit the dedicated staging bytes");
#pragma unroll
        for (int i = 0; i < 11; ++i) dbs[(q * 11 + i) * 32 + lane] = dbacc[i];
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (q == 0) {
            constexpr int slot_of[NL] = {0, 2, 3, 4, 5, 6, 7, 8, 10};
#pragma unroll
            for (int L = 0; L < NL; ++L) {
                const int OUT = g.out[L], IN = g.in[L];
                for (int p = 0; p < (TB::n(L) + 31) / 32; ++p) {
                    const int f = 32 * p + lane, i = slot_of[L] + p;
                    if (f < OUT)
                        part[g.flat_off[L] + OUT * IN + f] = (dbs[(0 * 11 + i) * 32 + lane] + dbs[(1 * 11 + i) * 32 + lane]) +
                                                             (dbs[(2 * 11 + i) * 32 + lane] + dbs[(3 * 11 + i) * 32 + lane]);
                }
            }
        }
    } else {
        // ======================= chain warps: thread = (row, column half) =======================
        const int half = warp >> 2, q = warp & 3;
        const int row = 32 * q + lane;
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        const uint32_t lane_off = ((uint32_t)(lane >> 2) << 4) | ((uint32_t)(lane & 3) << 2);
        const int tail0 = d.T - d.OT, rowstride = 2 * d.Fp;
        const long ntrk = (long)a.B * d.T * d.F;
        uint32_t ph = 0;
        int it = 0;
#ifdef ST_AE_TM_TIMING
        long long tclk = 0, treg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long hbuf[5] = {0, 0, 0, 0, 0};
        long long* hb = a.timing ? hbuf : nullptr;
#define BT0() if (a.timing) tclk = clock64();
#define BT(i) if (a.timing) { const long long n_ = clock64(); treg[i] += n_ - tclk; tclk = n_; }
#else
#define BT0()
#define BT(i)
        long long* hb = nullptr;
#endif
        mbar_wait_spin(w_ready, 0);                          // biases
        for (int tile = slot; tile < ntiles; tile += nslot, ++it) {
            const int R = tile * TILE + row;
            const bool ok = R < BF;
            const int b = ok ? R / d.F : 0, f = ok ? R - b * d.F : 0;
            float* mydbg = (DBG && a.dbg && tile == 0 && ae == 0) ? a.dbg + (long)row * 64 : nullptr;
            float vkeep[16];
            BT0()
            // ---- input track (this thread: frames [16 half, +16)) -> A operand, kept in registers for layer 0's weight gradient.
            // The forward kernel stored both tracks (nn_proc.py:309-310), so nothing is re-derived from the spectrum here.
            {
                const int c0 = 16 * half;
                const float* tp = a.trk + (long)ae * ntrk + ((long)b * d.T) * d.F + f;
#pragma unroll
                for (int e = 0; e < 16; ++e) vkeep[e] = (ok && c0 + e < d.T) ? __ldg(tp + (long)(c0 + e) * d.F) : 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int t = c0 + e;
                    if (t >= tail0 && t < d.T) vtail[(t - tail0) * TILE + row] = vkeep[e];
                }
                store_pair<16>(t_lane + TC_AH + c0, t_lane + TC_AL + c0, vkeep);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            asm volatile("bar.sync 1, 256;" ::: "memory");   // vtail complete (both halves)
            if (it > 0) {                                    // the weight-gradient accumulators live in the saved-output columns
                mbar_wait_spin(flushed, (uint32_t)((it - 1) & 1));
                tc_fence_after();
            }
            BT(0)

            // ---- forward layers 0..7: h = ELU(D + b) -> raw copy (kept for the backward) + (hi, lo) A operand
#pragma unroll 1
            for (int l = 0; l < NL - 1; ++l) {
                mbar_wait_spin(d_ready, ph);
                ph ^= 1;
                tc_fence_after();
                BT(1)
                const int nloc = c_n[l] / 2, c0 = half * nloc;
                const float* bl = bias + c_boff[l] + c0;
                const uint32_t t_save = t_lane + tc_act(l + 1) + c0;
                float* ld = mydbg ? mydbg + (long)l * TILE * 64 + c0 : nullptr;
                if (nloc == 32) epi_hidden<32, true>(t_lane + TC_D + c0, t_lane + TC_AH + c0, t_lane + TC_AL + c0, bl, ld, t_save);
                else if (nloc == 16) epi_hidden<16, true>(t_lane + TC_D + c0, t_lane + TC_AH + c0, t_lane + TC_AL + c0, bl, ld, t_save);
                else epi_hidden<8, true>(t_lane + TC_D + c0, t_lane + TC_AH + c0, t_lane + TC_AL + c0, bl, ld, t_save);
                if (l == 3 && half == 1 && TB::KP5 > 16) {
                    float kv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) kv[e] = (ok && e < d.K) ? __ldg(a.knobs + (long)b * d.K + e) : 0.f;
                    store_pair<8>(t_lane + TC_AH + 16, t_lane + TC_AL + 16, kv);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
                BT(2)
            }

            // ---- fnn_dec + output side backwards: gz[8], and the skip / residual gradient (left in vtail for the end of the tile)
            float eg8[TB::n(7) / 2];                         // ELU'(act[8]), this thread's columns (from the layer-8 hand-over)
            {
                const int j0 = 8 * half;
                float gre[8], gim[8], phv[8], x3[8];
                const float* gri = a.g_ri + (long)b * d.OTp * rowstride + f;
                const long oo0 = (long)b * d.OT * d.F + f;
                const float* x3p = ae == 1 ? a.mag_hat : a.g_mag_hat;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool in = ok && j0 + j < d.OT;
                    gre[j] = in ? __ldg(gri + (long)(j0 + j) * rowstride) : 0.f;
                    gim[j] = in ? __ldg(gri + (long)(j0 + j) * rowstride + d.Fp) : 0.f;
                    phv[j] = in ? __ldg(a.phs_hat + oo0 + (long)(j0 + j) * d.F) : 0.f;
                    x3[j] = (in && x3p) ? __ldg(x3p + oo0 + (long)(j0 + j) * d.F) : 0.f;
                }
                mbar_wait_spin(d_ready, ph);
                ph ^= 1;
                tc_fence_after();
                __syncwarp();
                if (lane == 0) mbar_arrive(fwd_done);        // the flush warps may stage act[8] (and overwrite forward weights)
                BT(1)
                uint32_t rr[8];
                tmem_ld<8>(t_lane + TC_D + j0, rr);
                tmem_wait_ld();
                const float* bl = bias + c_boff[NL - 1] + j0;
                float gz[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float e9 = elu_f(__uint_as_float(rr[j]) + bl[j]);
                    if (mydbg) mydbg[(long)(NL - 1) * TILE * 64 + j0 + j] = e9;
                    // sin / cos of phs_hat for the polar->rect gradient: the fast approximations (abs error ~1e-6 for |x| of a few
                    // radians) are far inside the gradient tolerance and keep 8 x 40 instructions off every tile's critical path
                    float2 sc;
                    __sincosf(phv[j], &sc.x, &sc.y);
                    float tb;
                    if (ae == 0) {                  // an = mag_hat (cos, sin);  mag_hat = ELU(dec) * track tail
                        const float gm = gre[j] * sc.y + gim[j] * sc.x + x3[j];
                        gz[j] = gm * vtail[(j0 + j) * TILE + row] * elu_grad(e9);
                        tb = gm * e9;
                    } else {                        // phs_hat = ELU(dec) + track tail
                        const float gp = x3[j] * (gim[j] * sc.y - gre[j] * sc.x);
                        gz[j] = gp * elu_grad(e9);
                        tb = gp;
                    }
                    if (j0 + j >= d.OT) { gz[j] = 0.f; tb = 0.f; }
                    vtail[(j0 + j) * TILE + row] = tb;
                }
                if (mydbg) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) mydbg[(long)(NL + 8) * TILE * 64 + j0 + j] = gz[j];
                }
                BT(3)
                uint32_t ghi[8], glo[8];
                bwd_handoff<TB, 8>(t_lane, half, lane, gz, a_ready, eg8, ghi, glo, hb);
                bwd_stage_gz<TB, 8>(smem_raw, lane_off, half, q, it, lane, ghi, glo, full, freeb, vkeep, hb);
                BT(5)
            }
            // ---- data gradients, layers 8..1: gz[l-1] = gh[l] * ELU'(act[l]); the weight-gradient slice of layer l-1 follows
#define ST_BW(L, EG_IN, EG_OUT)                                                                                      \
            float EG_OUT[(L) > 1 ? TB::n((L) > 1 ? (L) - 2 : 0) / 2 : 1];                                            \
            {                                                                                                        \
                constexpr int NLOC = TB::dn(L) / 2;                                                                  \
                const int c0 = half * NLOC;                                                                          \
                mbar_wait_spin(d_ready, ph);                                                                         \
                ph ^= 1;                                                                                             \
                tc_fence_after();                                                                                    \
                BT(1)                                                                                                \
                uint32_t gh[NLOC];                                                                                   \
                tmem_ld<NLOC>(t_lane + TC_D + c0, gh);                                                               \
                tmem_wait_ld();                                                                                      \
                float gz[NLOC];                                                                                      \
                _Pragma("unroll")                                                                                    \
                for (int c = 0; c < NLOC; ++c) gz[c] = __uint_as_float(gh[c]) * EG_IN[c];                            \
                if (mydbg) {                                                                                         \
                    _Pragma("unroll")                                                                                \
                    for (int c = 0; c < NLOC; ++c) mydbg[(long)(NL + (L) - 1) * TILE * 64 + c0 + c] = gz[c];          \
                }                                                                                                    \
                BT(4)                                                                                                \
                uint32_t ghi[NLOC], glo[NLOC];                                                                       \
                bwd_handoff<TB, (L) - 1>(t_lane, half, lane, gz, a_ready, EG_OUT, ghi, glo, hb);                     \
                bwd_stage_gz<TB, (L) - 1>(smem_raw, lane_off, half, q, it, lane, ghi, glo, full, freeb, vkeep, hb);  \
                BT(5)                                                                                                \
            }
            ST_BW(8, eg8, eg7) ST_BW(7, eg7, eg6) ST_BW(6, eg6, eg5) ST_BW(5, eg5, eg4) ST_BW(4, eg4, eg3) ST_BW(3, eg3, eg2) ST_BW(2, eg2, eg1)
            ST_BW(1, eg1, eg0)
#undef ST_BW
            // ---- layer 0: gh[0] + skip / residual gradient = dLoss/d(track), stored lane <-> bin (coalesced)
            {
                const int c0 = 16 * half;
                mbar_wait_spin(d_ready, ph);
                ph ^= 1;
                tc_fence_after();
                BT(1)
                uint32_t gh[16];
                tmem_ld<16>(t_lane + TC_D + c0, gh);
                tmem_wait_ld();
                tc_fence_before();
                float* gt = a.g_track + (long)ae * ntrk + ((long)b * d.T) * d.F + f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int t = c0 + e;
                    if (ok && t < d.T) gt[(long)t * d.F] = __uint_as_float(gh[e]) + (t >= tail0 ? vtail[(t - tail0) * TILE + row] : 0.f);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // vtail is free for the next tile
            BT(6)
        }
#ifdef ST_AE_TM_TIMING
        if (a.timing && lane == 0)
        {
            for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.timing) + 8 * warp + i, (unsigned long long)treg[i]);
            for (int i = 0; i < 5; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.timing) + 64 + 8 * warp + i, (unsigned long long)hbuf[i]);
        }
#endif
#undef BT0
#undef BT
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BWD_ISSUER) {
        tc_fence_after();
        tmem_dealloc(0u, 512);
    }
}

template <class TB>
constexpr size_t fwd_smem_bytes() {
    return sizeof(float) * (2 * (size_t)image_floats<TB>() + 2 * XCH_J * TILE + (TB::KP1 == 32 ? 2 * TB::KP1 * TILE : 0)) + 5 * sizeof(uint64_t) + 16;
}

template <class TB>
bool launch_fwd(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, const float* knobs,
                int B, float* mag, float* trk, float* mag_hat, float* phs_hat, float* ri, float* ri_lo, float* wpack, float* dbg,
                long long* timing, int sm_count, bool pack, cudaStream_t s_pack, cudaStream_t s) {
    constexpr size_t smem = fwd_smem_bytes<TB>();
    static_assert(smem <= 227 * 1024, "forward tile does not fit shared memory");
    static bool configured_on[ST_MAX_DEVICES] = {};              // function attributes are per device (context), not per process
    bool& configured = configured_on[st_current_device_slot()];
    if (!configured) {
        if (cudaFuncSetAttribute(ae_fwd_tm_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        configured = true;
    }
    if (pack) ae_pack_kernel<TB, false><<<dim3(NL, 2), 256, 0, s_pack>>>(g, pm, pp, wpack);
    if (B > 0) {
        const long ntiles = ((long)B * d.F + TILE - 1) / TILE;
        const int grid = (int)std::min<long>(ntiles, sm_count);
        ae_fwd_tm_kernel<TB><<<grid, FWD_THREADS, smem, s>>>(d, wpack, spec, knobs, B, mag, trk, mag_hat, phs_hat, ri, ri_lo, dbg, timing);
    }
    return true;
}

// dL/d(re, im) from the two track gradients (magnitude and phase autoencoder), stored as the (hi, lo) tf32 pair the analysis
// weight-gradient GEMM consumes.  mag = sqrt(re^2 + im^2) with subgradient 0 at 0 (torch.norm backward, nn_proc.py:309);
// phs = atan2(im, re + 1e-7) (nn_proc.py:310).  g_mag: optional external gradient of the mag output.
__global__ void ae_track_to_spec_kernel(StDims d, int B, const float* __restrict__ spec, const float* __restrict__ gt_m,
                                        const float* __restrict__ gt_p, const float* __restrict__ g_mag, float* __restrict__ g_spec,
                                        float* __restrict__ g_spec_lo) {
    const long n = (long)B * d.T * d.F;
    const int rowstride = 2 * d.Fp;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long bt = idx / d.F;
        const int f = (int)(idx - bt * d.F);
        const int b = (int)(bt / d.T), tt = (int)(bt - (long)b * d.T);
        const long o = ((long)b * d.Tp + tt) * rowstride + f;
        const float re = __ldg(spec + o), im = __ldg(spec + o + d.Fp);
        float gm = __ldg(gt_m + idx);
        if (g_mag) gm += __ldg(g_mag + idx);
        const float gp = __ldg(gt_p + idx);
        const float m = sqrtf(re * re + im * im);
        const float scm = m > 0.f ? gm / m : 0.f;
        const float u = re + 1e-7f;
        const float den = u * u + im * im;
        const float scp = den > 0.f ? gp / den : 0.f;
        st_split_tf32(scm * re - scp * im, g_spec[o], g_spec_lo[o]);
        st_split_tf32(scm * im + scp * u, g_spec[o + d.Fp], g_spec_lo[o + d.Fp]);
    }
}

template <class TB>
int launch_bwd(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, const float* trk,
               const float* knobs, int B,
               const float* mag_hat, const float* phs_hat, const float* g_ri, const float* g_mag_hat, const float* g_mag,
               float* g_track, float* g_spec, float* g_spec_lo, float* partials, float* wpack, float* dbg, long long* timing, int sm_count, bool pack,
               cudaStream_t s_pack, cudaStream_t s) {
    constexpr size_t smem = BwdSmem<TB>::TOTAL;
    static_assert(smem <= 227 * 1024, "backward tile does not fit shared memory");
    static bool configured_on[ST_MAX_DEVICES] = {};              // function attributes are per device (context), not per process
    bool& configured = configured_on[st_current_device_slot()];
    if (!configured) {
        if (cudaFuncSetAttribute(ae_bwd_tm_kernel<TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        if (cudaFuncSetAttribute(ae_bwd_tm_kernel<TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        configured = true;
    }
    if (pack) ae_pack_kernel<TB, true><<<dim3(NL, 2), 256, 0, s_pack>>>(g, pm, pp, wpack);
    if (B <= 0) return 0;
    const long ntiles = ((long)B * d.F + TILE - 1) / TILE;
    const int nslot = (int)std::min<long>(ntiles, sm_count / 2);
    BwdArgs a;
    a.image = wpack; a.trk = trk; a.knobs = knobs; a.mag_hat = mag_hat; a.phs_hat = phs_hat; a.g_ri = g_ri; a.g_mag_hat = g_mag_hat;
    a.g_track = g_track; a.partials = partials; a.dbg = dbg; a.timing = timing; a.B = B;
    if (dbg) ae_bwd_tm_kernel<TB, true><<<2 * nslot, BWD_THREADS, smem, s>>>(d, g, a);
    else ae_bwd_tm_kernel<TB, false><<<2 * nslot, BWD_THREADS, smem, s>>>(d, g, a);
    if (g_spec) {
        const long ntrk = (long)B * d.T * d.F;
        ae_track_to_spec_kernel<<<(int)std::min<long>((ntrk + 255) / 256, 8L * sm_count), 256, 0, s>>>(d, B, spec, g_track, g_track + ntrk, g_mag,
                                                                                                    g_spec, g_spec_lo);
    }
    return nslot;
}

}  // namespace

long st_ae_tm_pack_floats() { return 2L * image_floats<Tab<64, 24>>(); }

// Covers T <= 64, OT <= 16, K <= 8.  Writes mag (optional), trk (optional: [2][B][T][F] magnitude | phase tracks for the backward
// kernel), mag_hat, phs_hat and the (hi, lo) polar->rect operand ri.
// wpack: workspace of st_ae_tm_pack_floats() floats.  pack: (re)build the weight image on `s_pack` first (the launching stream,
// or a side stream the caller joins before the forward); B = 0 packs only.
// dbg (nullable): [2 autoencoders][9 layers][128 rows][64] layer outputs of tile 0 (test harness only).
bool st_launch_ae_forward_tm(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                             const float* knobs, int B, float* mag, float* trk, float* mag_hat, float* phs_hat, float* ri, float* ri_lo,
                             float* wpack, float* dbg, long long* timing, int sm_count, bool pack, cudaStream_t s_pack, cudaStream_t s) {
    if (d.T > 64 || d.OT > 16 || d.K > 8) return false;
    const bool wide = d.T > 32, knob = d.K > 0;
#define ST_FWD(...) launch_fwd<__VA_ARGS__>(d, g, pm, pp, spec, knobs, B, mag, trk, mag_hat, phs_hat, ri, ri_lo, wpack, dbg, timing, sm_count, pack, s_pack, s)
    if (!wide && knob) return ST_FWD(Tab<32, 24>);
    if (!wide) return ST_FWD(Tab<32, 16>);
    if (knob) return ST_FWD(Tab<64, 24>);
    return ST_FWD(Tab<64, 16>);
#undef ST_FWD
}

long st_ae_tm_bwd_pack_floats() { return 2L * bwd_image_floats<Tab<32, 24>>(); }

// Backward of both autoencoders with in-kernel recompute (no saved activations): track gradients -> g_track (2 * B * T * F
// floats) and, when g_spec is non-null, dL/d(re, im) as the (hi, lo) operand of the analysis weight-gradient GEMM; per-CTA
// partial weight / bias gradients -> partials ([(slot, autoencoder)][flat_total]).  Returns the number of slots written per
// autoencoder (0: geometry not covered -- T <= 32, OT <= 16, K <= 8).  wpack: st_ae_tm_bwd_pack_floats() floats.
int st_launch_ae_backward_tm(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, const float* trk,
                             const float* knobs, int B, const float* mag_hat, const float* phs_hat, const float* g_ri,
                             const float* g_mag_hat, const float* g_mag, float* g_track, float* g_spec, float* g_spec_lo,
                             float* partials, float* wpack, float* dbg, long long* timing, int sm_count, bool pack, cudaStream_t s_pack, cudaStream_t s) {
    if (d.T > 32 || d.OT > 16 || d.K > 8) return 0;
    if (d.K > 0)
        return launch_bwd<Tab<32, 24>>(d, g, pm, pp, spec, trk, knobs, B, mag_hat, phs_hat, g_ri, g_mag_hat, g_mag, g_track, g_spec, g_spec_lo,
                                       partials, wpack, dbg, timing, sm_count, pack, s_pack, s);
    return launch_bwd<Tab<32, 16>>(d, g, pm, pp, spec, trk, knobs, B, mag_hat, phs_hat, g_ri, g_mag_hat, g_mag, g_track, g_spec, g_spec_lo,
                                   partials, wpack, dbg, timing, sm_count, pack, s_pack, s);
}
