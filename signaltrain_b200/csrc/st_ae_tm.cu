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
};

__device__ __forceinline__ float elu_f(float z) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 1.4426950408889634f));
    return z > 0.f ? z : e - 1.f;
}
__device__ __forceinline__ float elu_grad(float h) { return h > 0.f ? 1.f : h + 1.f; }   // dELU/dz through the output h

// One copy of the libdevice routines in the instruction stream (they are inlined per call site otherwise: ~25 call sites).
__device__ __noinline__ float atan2_ni(float y, float x) { return atan2f(y, x); }
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

// Write NV values of this thread's row as an exact tf32 pair into the (hi, lo) A-operand columns.
template <int NV>
__device__ __forceinline__ void store_pair(uint32_t t_hi, uint32_t t_lo, const float (&v)[NV]) {
    uint32_t hi[NV], lo[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float h, l;
        st_split_tf32(v[i], h, l);
        hi[i] = __float_as_uint(h);
        lo[i] = __float_as_uint(l);
    }
    tmem_st<NV>(t_hi, hi);
    tmem_st<NV>(t_lo, lo);
}

// Image of one autoencoder's operands as the kernels want them in shared memory: [hi plane][lo plane][bias].  Built once per
// step by ae_pack_kernel (the weights change every step), copied into each CTA with one-dimensional bulk copies.
template <class TB>
__host__ __device__ constexpr int image_floats() { return 2 * TB::wfloats + (TB::bfloats + 255) / 256 * 256; }   // keeps the next image 1024-byte aligned

// B operand of the forward layers = W[n = out][k = in], K-major, zero padded to (n(l), 32-float K-blocks), as (hi, lo).
template <class TB>
__global__ void ae_pack_kernel(AeGeom g, AeParams pm, AeParams pp, float* __restrict__ image) {
    const AeParams& p = blockIdx.y ? pp : pm;
    float* img = image + (long)blockIdx.y * image_floats<TB>();
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
    float* bias = img + 2 * TB::wfloats + TB::boff(l);
    for (int o = threadIdx.x; o < NP; o += blockDim.x) bias[o] = (o < OUT) ? __ldg(p.b[l] + o) : 0.f;
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
template <int NLOC>
__device__ __forceinline__ void epi_hidden(uint32_t t_d, uint32_t t_hi, uint32_t t_lo, const float* __restrict__ bias, float* dbg, long long* probe = nullptr) {
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
                 const float* __restrict__ knobs, int B, float* __restrict__ mag_out, float* __restrict__ mag_hat,
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
    uint64_t* bars = reinterpret_cast<uint64_t*>(tails + 2 * XCH_J * TILE);
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
#pragma unroll
                for (int e = 0; e < NLOC; ++e) {
                    const bool in = ok && c0 + e < d.T;
                    re[e] = in ? __ldg(sp + (long)(c0 + e) * rowstride) : 0.f;
                    im[e] = in ? __ldg(sp + (long)(c0 + e) * rowstride + d.Fp) : 0.f;
                }
#pragma unroll
                for (int e = 0; e < NLOC; ++e) {
                    const int t = c0 + e;
                    const bool in = ok && t < d.T;
                    vm[e] = sqrtf(re[e] * re[e] + im[e] * im[e]);
                    vp[e] = in ? atan2_ni(im[e], re[e] + 1e-7f) : 0.f;
                    if (in && mag_out) mag_out[((long)b * d.T + t) * d.F + f] = vm[e];
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
                if (nloc == 32) epi_hidden<32>(t_d + c0, t_hi + c0, t_lo + c0, bl, ld, probe);
                else if (nloc == 16) epi_hidden<16>(t_d + c0, t_hi + c0, t_lo + c0, bl, ld, probe);
                else epi_hidden<8>(t_d + c0, t_hi + c0, t_lo + c0, bl, ld, probe);
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

template <class TB>
constexpr size_t fwd_smem_bytes() {
    return sizeof(float) * (2 * (size_t)image_floats<TB>() + 2 * XCH_J * TILE) + 5 * sizeof(uint64_t) + 16;
}

template <class TB>
bool launch_fwd(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, const float* knobs,
                int B, float* mag, float* mag_hat, float* phs_hat, float* ri, float* ri_lo, float* wpack, float* dbg,
                long long* timing, int sm_count, bool pack, cudaStream_t s_pack, cudaStream_t s) {
    constexpr size_t smem = fwd_smem_bytes<TB>();
    static_assert(smem <= 227 * 1024, "forward tile does not fit shared memory");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(ae_fwd_tm_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        configured = true;
    }
    if (pack) ae_pack_kernel<TB><<<dim3(NL, 2), 256, 0, s_pack>>>(g, pm, pp, wpack);
    if (B > 0) {
        const long ntiles = ((long)B * d.F + TILE - 1) / TILE;
        const int grid = (int)std::min<long>(ntiles, sm_count);
        ae_fwd_tm_kernel<TB><<<grid, FWD_THREADS, smem, s>>>(d, wpack, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, dbg, timing);
    }
    return true;
}

}  // namespace

long st_ae_tm_pack_floats() { return 2L * image_floats<Tab<64, 24>>(); }

// Covers T <= 64, OT <= 16, K <= 8.  Writes mag (optional), mag_hat, phs_hat and the (hi, lo) polar->rect operand ri.
// wpack: workspace of st_ae_tm_pack_floats() floats.  pack: (re)build the weight image on `s_pack` first (the launching stream,
// or a side stream the caller joins before the forward); B = 0 packs only.
// dbg (nullable): [2 autoencoders][9 layers][128 rows][64] layer outputs of tile 0 (test harness only).
bool st_launch_ae_forward_tm(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                             const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri, float* ri_lo,
                             float* wpack, float* dbg, long long* timing, int sm_count, bool pack, cudaStream_t s_pack, cudaStream_t s) {
    if (d.T > 64 || d.OT > 16 || d.K > 8) return false;
    const bool wide = d.T > 32, knob = d.K > 0;
#define ST_FWD(...) launch_fwd<__VA_ARGS__>(d, g, pm, pp, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, wpack, dbg, timing, sm_count, pack, s_pack, s)
    if (!wide && knob) return ST_FWD(Tab<32, 24>);
    if (!wide) return ST_FWD(Tab<32, 16>);
    if (knob) return ST_FWD(Tab<64, 24>);
    return ST_FWD(Tab<64, 16>);
#undef ST_FWD
}
