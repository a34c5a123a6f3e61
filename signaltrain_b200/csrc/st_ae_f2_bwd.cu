// Autoencoder backward on the packed-FP32 pipe (FFMA2), warp-specialised -- the production path.
//
// Back-propagates AsymAutoEncoder.forward (nn_proc.py:77-126) for one autoencoder per launch, together with the output
// side of AsymMPAEC.forward (skip-filter :115, phase residual :322, polar->rect :325-326) and its input side
// (magnitude / phase, :309-310), from the activation records the forward saved (st_ae_f2.cu).
//
// Two kinds of work per (batch, bin) row, 8128 MAC each, with opposite natural layouts:
//   * data gradient   gz_{l-1} = (gz_l . W_l) * ELU'(h_{l-1})   -- a chain per row, weights shared by all rows:
//     LANE <-> ROW, gz in registers, W_l^T broadcast from shared memory (like the forward);
//   * weight gradient dW_l += gz_l^T h_{l-1}                    -- a reduction over rows:
//     LANE <-> (out block, in block) register tile of dW_l, rows streamed from shared memory, accumulators stay in
//     registers for the whole kernel (no atomics, deterministic).
// So the CTA is split into PRODUCER warps (data gradient of 32-row chunks) and CONSUMER warps (weight gradient), coupled
// by shared-memory slots and mbarriers.  The nine layers are grouped into seven UNITS; a slot holds one unit of one
// chunk: [32 rows][gz columns | h columns].  A producer owns three rotating slots; it fills the slot of unit u+1 while
// the consumers of unit u (one or two warps per unit, which own that unit's dW tiles) read theirs.  Exact fp32 (no tf32
// split), no block-wide barrier in the steady state.
//
//   unit   layers (0-based)   slot columns
//   u0     8 (fnn_dec)        gz9 @0 (16) | h8 @16 (64)
//   u1     7                  gz8 @0 (64) | h7 @64 (32)
//   u2     6, 5               gz7 @0 (32), gz6 @32 (16) | h6 @48 (16), h5 @64 (16)
//   u3     4 (fnn_addknobs)   gz5 @0 (16) | h4 ++ knobs @16 (32)
//   u4     3, 2               gz4 @0 (16), gz3 @16 (16) | h3 @32 (16), h2 @48 (32)
//   u5     1                  gz2 @0 (32) | h1 @32 (64)
//   u6     0 (fnn_enc)        gz1 @0 (64) | track v @64 (32)
#include <algorithm>
#include <cstdlib>

#include "st_common.cuh"
#include "st_tc_prims.cuh"

namespace {

using st_tc::mbar_arrive;
using st_tc::mbar_init;
// Waiting warps back off between polls: the consumers idle most of the time, and every mbarrier poll is a shared-memory
// operation competing with the producers' weight loads.  Bounded like st_tc::mbar_wait (trap, never a hung GPU).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (st_tc::mbar_try_wait(bar, parity)) return;
    long long spins = 0;
    while (!st_tc::mbar_try_wait(bar, parity)) {
        __nanosleep(128);
        if (++spins > 20000000LL) __trap();        // several seconds
    }
}
using st_tc::smem_u32;

constexpr int NL = ST_AE_LAYERS;
constexpr int NPROD = 4;                  // producer warps
constexpr int NSLOT = 3;                  // rotating slots per producer
constexpr int NUNIT = 7;
constexpr int NCONS = 10;                 // consumer warps
constexpr int WARPS = NPROD + NCONS;
constexpr int ROWS = 32;                  // rows per chunk: lane <-> row in the producers
constexpr int LD = 100;                   // slot row stride (floats): 4 (mod 32) -> conflict-free 16-byte row accesses
constexpr int TLD = 36;                   // per-producer side buffer row: [0,16) e9 -> skip gradient, [16,32) track tail
constexpr int SLOT_FLOATS = ROWS * LD;

struct BwdGeom {
    int woff[NL];         // float offset of W_l^T [i][KD_l] in the staged weight block
    int wfloats;
    int soff[NL];         // record offsets (shared with the forward kernels): h1..h8 (h4 slot 32 wide), e9
    int soff_v, ss;
};

__device__ __forceinline__ float elu_grad(float h) { return h > 0.f ? 1.f : h + 1.f; }   // dELU/dz through the output h

// acc += a * (wx, wy) on the packed-fp32 pipe; volatile keeps the interleaving of the independent accumulator chains
__device__ __forceinline__ void fma2v(float2& acc, const float2& a, float wx, float wy) {
    unsigned long long& c = reinterpret_cast<unsigned long long&>(acc);
    unsigned long long w;
    asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(wx), "f"(wy));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(w));
}
// acc(x, y) += (ax, ay) * (b, b)
__device__ __forceinline__ void fma2d(float2& acc, float ax, float ay, unsigned long long bb) {
    unsigned long long& c = reinterpret_cast<unsigned long long&>(acc);
    unsigned long long a;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(ax), "f"(ay));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(bb));
}
__device__ __forceinline__ unsigned long long dup2(float x) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool ok) {
    const int bytes = ok ? 16 : 0;        // src-size 0: zero fill (rows past the end of the batch)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool ok) {
    const int bytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}

// NF4 float4 of the lane's record -> the lane's slot row
template <int NF4>
__device__ __forceinline__ void prefetch_row(float* dst, const float* src, bool ok) {
#pragma unroll
    for (int q = 0; q < NF4; ++q) cp_async16(dst + 4 * q, src + 4 * q, ok);
}

template <int WIDTH>
__device__ __forceinline__ void reload_row(float2 (&gz)[16], const float* __restrict__ src) {
#pragma unroll
    for (int c = 0; c < WIDTH / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * c);
        gz[2 * c] = make_float2(v.x, v.y);
        gz[2 * c + 1] = make_float2(v.z, v.w);
    }
}

// Data gradient of one layer for the lane's row:  c[i] = sum_o gz[o] WT[i][o],  i in [0, NOUT), handed block-wise to `epi`.
// Each accumulator is an (even-o, odd-o) pair of partial sums; IBLK independent FFMA2 chains.  The weight loads of
// o-step s+1 are issued before the FMAs of step s (two register stages): a single warp per scheduler has nobody else to
// hide the shared-memory latency behind.  gz comes from registers (KD <= 32) or from the lane's own slot row (KD = 64).
template <int KD, int NOUT, int IBLK, class Epi>
__device__ __forceinline__ void dgrad_core(const float* __restrict__ WT, const float2 (&gz)[16], const float* __restrict__ gzrow,
                                           Epi epi) {
    constexpr bool GZREG = KD <= 32;
    constexpr int NST = KD / 4;
#pragma unroll(NOUT / IBLK <= 2 ? 2 : 1)
    for (int i = 0; i < NOUT; i += IBLK) {
        float2 acc[IBLK];
#pragma unroll
        for (int j = 0; j < IBLK; ++j) acc[j] = make_float2(0.f, 0.f);
        const float* w = WT + i * KD;
        float4 wv[2][IBLK], gq[2];
#pragma unroll
        for (int j = 0; j < IBLK; ++j) wv[0][j] = *reinterpret_cast<const float4*>(w + j * KD);
        if (!GZREG) gq[0] = *reinterpret_cast<const float4*>(gzrow);
#pragma unroll
        for (int st = 0; st < NST; ++st) {
            const int cur = st & 1;
            if (st + 1 < NST) {
#pragma unroll
                for (int j = 0; j < IBLK; ++j) wv[cur ^ 1][j] = *reinterpret_cast<const float4*>(w + j * KD + 4 * (st + 1));
                if (!GZREG) gq[cur ^ 1] = *reinterpret_cast<const float4*>(gzrow + 4 * (st + 1));
            }
            const float2 g0 = GZREG ? gz[(2 * st) % 16] : make_float2(gq[cur].x, gq[cur].y);
            const float2 g1 = GZREG ? gz[(2 * st + 1) % 16] : make_float2(gq[cur].z, gq[cur].w);
#pragma unroll
            for (int j = 0; j < IBLK; ++j) fma2v(acc[j], g0, wv[cur][j].x, wv[cur][j].y);
#pragma unroll
            for (int j = 0; j < IBLK; ++j) fma2v(acc[j], g1, wv[cur][j].z, wv[cur][j].w);
        }
        float c[IBLK];
#pragma unroll
        for (int j = 0; j < IBLK; ++j) c[j] = acc[j].x + acc[j].y;
        epi(i, c);
    }
}

// Hidden layers:  dst[i] = ELU'(hsrc[i]) * c[i]
template <int KD, int NOUT, int IBLK>
__device__ __forceinline__ void dgrad_layer(const float* __restrict__ WT, const float2 (&gz)[16], const float* __restrict__ gzrow,
                                            const float* __restrict__ hsrc, float* __restrict__ dst) {
    dgrad_core<KD, NOUT, IBLK>(WT, gz, gzrow, [&](int i, const float (&c)[IBLK]) {
#pragma unroll
        for (int j = 0; j < IBLK; j += 4) {
            const float4 hh = *reinterpret_cast<const float4*>(hsrc + i + j);
            *reinterpret_cast<float4*>(dst + i + j) =
                make_float4(c[j] * elu_grad(hh.x), c[j + 1] * elu_grad(hh.y), c[j + 2] * elu_grad(hh.z), c[j + 3] * elu_grad(hh.w));
        }
    });
}

// Weight-gradient tile of one lane: acc[o2][i] (+)= (gz[2 o2], gz[2 o2 + 1]) * h[i] over `nrows` (even) slot rows; the
// operands of row r+1 are loaded before the FMAs of row r.
template <int OB, int IB>
struct WgOperands {
    float g[OB], hv[IB];
    __device__ __forceinline__ void load(const float* __restrict__ gp, const float* __restrict__ hp) {
        if (OB == 2) {
            const float2 v = *reinterpret_cast<const float2*>(gp);
            g[0] = v.x; g[1] = v.y;
        } else {
#pragma unroll
            for (int q = 0; q < OB / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(gp + 4 * q);
                g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
            }
        }
#pragma unroll
        for (int q = 0; q < IB / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(hp + 4 * q);
            hv[4 * q] = v.x; hv[4 * q + 1] = v.y; hv[4 * q + 2] = v.z; hv[4 * q + 3] = v.w;
        }
    }
    __device__ __forceinline__ void fma(float2 (&acc)[OB / 2][IB]) const {
#pragma unroll
        for (int i = 0; i < IB; ++i) {
            const unsigned long long hh = dup2(hv[i]);
#pragma unroll
            for (int o2 = 0; o2 < OB / 2; ++o2) fma2d(acc[o2][i], g[2 * o2], g[2 * o2 + 1], hh);
        }
    }
};
template <int OB, int IB>
__device__ __forceinline__ void wgrad_rows(float2 (&acc)[OB / 2][IB], const float* __restrict__ gp, const float* __restrict__ hp,
                                           int nrows) {
    WgOperands<OB, IB> a, b;
    a.load(gp, hp);
#pragma unroll 1
    for (int r = 0; r < nrows; r += 2) {
        b.load(gp + (r + 1) * LD, hp + (r + 1) * LD);
        a.fma(acc);
        if (r + 2 < nrows) a.load(gp + (r + 2) * LD, hp + (r + 2) * LD);
        b.fma(acc);
    }
}

// Bias gradient partial: lane <-> (4-column block cb, row subset rs) of the slot's gz columns.
template <int CBP>
__device__ __forceinline__ void bias_rows(float4& b, const float* __restrict__ rows, int nrows, int lane) {
    constexpr int NRS = 32 / CBP;
    const int cb = lane % CBP, rs = lane / CBP;
#pragma unroll 2
    for (int r = rs; r < nrows; r += NRS) {
        const float4 v = *reinterpret_cast<const float4*>(rows + r * LD + 4 * cb);
        b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
}

template <int OB, int IB>
__device__ __forceinline__ void zero_acc(float2 (&acc)[OB / 2][IB]) {
#pragma unroll
    for (int a = 0; a < OB / 2; ++a)
#pragma unroll
        for (int i = 0; i < IB; ++i) acc[a][i] = make_float2(0.f, 0.f);
}

// Flush of one layer's tile: (optionally add the partner half's copy from scratch, or park the own copy there) and write
// dW[o][i] into the CTA's partial-gradient vector.  mode 0: write global, 1: park in scratch, 2: add scratch then write.
template <int OB, int IB>
__device__ __forceinline__ void flush_tile(const float2 (&acc)[OB / 2][IB], int mode, float* __restrict__ scratch, int lane,
                                           float* __restrict__ dst, int OUT, int IN, int ob, int ib) {
#pragma unroll
    for (int o2 = 0; o2 < OB / 2; ++o2)
#pragma unroll
        for (int i = 0; i < IB; ++i) {
            float2 v = acc[o2][i];
            const int e = o2 * IB + i;
            if (mode == 1) {
                scratch[(2 * e) * 32 + lane] = v.x;
                scratch[(2 * e + 1) * 32 + lane] = v.y;
                continue;
            }
            if (mode == 2) {
                v.x += scratch[(2 * e) * 32 + lane];
                v.y += scratch[(2 * e + 1) * 32 + lane];
            }
            const int o = ob * OB + 2 * o2, ii = ib * IB + i;
            if (ii < IN) {
                if (o < OUT) dst[o * IN + ii] = v.x;
                if (o + 1 < OUT) dst[(o + 1) * IN + ii] = v.y;
            }
        }
}

// One consumer warp: owns the dW tiles of unit U's layers; `half` of `nh` warps splitting the 32 rows of every chunk.
//   L1/L2: 0-based layers of the unit (L2 < 0: single layer); G*/H*: slot columns of gz / h; tile (OB x IB), NIB in-blocks.
template <int L1, int G1, int H1, int OB1, int IB1, int NIB1, int L2, int G2, int H2, int OB2, int IB2, int NIB2, int CBP>
__device__ __forceinline__ void consumer(int unit, int half, int nh, int lane, int nk, float* slots, uint64_t* full, uint64_t* empty,
                                         float* scratch, float* bscratch, int cw, const AeGeom& g, float* __restrict__ dst,
                                         long long* __restrict__ timing) {
    constexpr bool TWO = L2 >= 0;
    constexpr int OB2e = TWO ? OB2 : 2, IB2e = TWO ? IB2 : 4, NIB2e = TWO ? NIB2 : 1;
    float2 acc1[OB1 / 2][IB1];
    float2 acc2[OB2e / 2][IB2e];
    zero_acc<OB1, IB1>(acc1);
    zero_acc<OB2e, IB2e>(acc2);
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ob1 = lane / NIB1, ib1 = lane % NIB1, ob2 = lane / NIB2e, ib2 = lane % NIB2e;
    const int nrows = ROWS / nh, r0 = half * nrows;
    long long tw = 0, tr = 0, tc = timing ? clock64() : 0;
    for (int k = 0; k < nk; ++k) {
        // full barriers are per (producer, unit): this warp sees their phases strictly in order (a barrier shared by
        // several units would let the consumer of a later use slip through on the parity of an earlier phase)
        const int p = k % NPROD, kk = k / NPROD, j = kk * NUNIT + unit;
        const int s = p * NSLOT + j % NSLOT;
        mbar_wait(full + p * NUNIT + unit, (uint32_t)(kk & 1));
        if (timing) { const long long n_ = clock64(); tw += n_ - tc; tc = n_; }
        const float* rows = slots + (long)s * SLOT_FLOATS + r0 * LD;
        wgrad_rows<OB1, IB1>(acc1, rows + G1 + ob1 * OB1, rows + H1 + ib1 * IB1, nrows);
        if (TWO) wgrad_rows<OB2e, IB2e>(acc2, rows + G2 + ob2 * OB2e, rows + H2 + ib2 * IB2e, nrows);
        bias_rows<CBP>(bsum, rows, nrows, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive_n(empty + s, (uint32_t)(2 / nh));
        if (timing) { const long long n_ = clock64(); tr += n_ - tc; tc = n_; }
    }
    if (timing && cw == 0 && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(timing) + 6, (unsigned long long)tw);
        atomicAdd(reinterpret_cast<unsigned long long*>(timing) + 7, (unsigned long long)tr);
    }
    // ---- flush (the slots are free: every producer and consumer has passed the CTA barrier)
    __syncthreads();
    const int base = cw - half;                                  // first warp of the unit
    float* my_scr = scratch + (long)base * 64 * 32;
    *reinterpret_cast<float4*>(bscratch + ((long)cw * 32 + lane) * 4) = bsum;
    if (nh == 2 && half == 1) flush_tile<OB1, IB1>(acc1, 1, my_scr, lane, nullptr, 0, 0, 0, 0);
    __syncthreads();
    if (half == 0) {
        const int mode = nh == 2 ? 2 : 0;
        flush_tile<OB1, IB1>(acc1, mode, my_scr, lane, dst + g.flat_off[L1], g.out[L1], g.in[L1], ob1, ib1);
        if (TWO) flush_tile<OB2e, IB2e>(acc2, 0, nullptr, lane, dst + g.flat_off[TWO ? L2 : 0], g.out[TWO ? L2 : 0], g.in[TWO ? L2 : 0], ob2, ib2);
        // bias: lanes [0, CBP) combine the row subsets (and the partner half) of their 4-column block, in fixed order
        if (lane < CBP) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int hh = 0; hh < nh; ++hh)
                for (int rs = 0; rs < 32 / CBP; ++rs) {
                    const float4 v = *reinterpret_cast<const float4*>(bscratch + ((long)(base + hh) * 32 + rs * CBP + lane) * 4);
                    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
                }
            const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = 4 * lane + q;
                // gz column -> (layer, output)
                int L = -1, o = 0;
                if (col >= G1 && col < G1 + OB1 * (32 / NIB1)) { L = L1; o = col - G1; }
                if (TWO && col >= G2 && col < G2 + OB2e * (32 / NIB2e)) { L = L2; o = col - G2; }
                if (L >= 0 && o < g.out[L]) dst[g.flat_off[L] + g.out[L] * g.in[L] + o] = tv[q];
            }
        }
    }
}

// AE = 0: magnitude autoencoder, AE = 1: phase autoencoder.  IN0 / OUT8: T and OT rounded up to a multiple of four.
template <int AE, int IN0, int OUT8>
__global__ void __launch_bounds__(WARPS * 32, 1)
ae_bwd_f2_kernel(StDims d, AeGeom g, BwdGeom bg, AeParams p, int B, const float* __restrict__ save,
                 const float* __restrict__ mag_hat, const float* __restrict__ phs_hat, const float* __restrict__ g_ri,
                 const float* __restrict__ g_mag_hat, float* __restrict__ g_track, float* __restrict__ partials,
                 long long* __restrict__ timing) {
    extern __shared__ __align__(16) float smem[];
    // optional region timing (st_debug_ae_timing): cycles of producer warp 0 (regions 0-5) and of consumer warp 0
    // (6: waiting, 7: working), summed over CTAs
    long long tclk = 0, treg[6] = {0, 0, 0, 0, 0, 0};
#define ST_T0() if (timing) tclk = clock64();
#define ST_T(i) if (timing) { const long long n_ = clock64(); treg[i] += n_ - tclk; tclk = n_; }
    float* WT = smem;
    float* slots = WT + bg.wfloats;                                   // [NPROD * NSLOT][ROWS][LD]
    float* tails = slots + NPROD * NSLOT * SLOT_FLOATS;               // [NPROD][ROWS][TLD]
    uint64_t* full = reinterpret_cast<uint64_t*>(tails + NPROD * ROWS * TLD);
    uint64_t* empty = full + NPROD * NUNIT;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- stage W_l^T [i][KD_l] (zero padded); KD_l = padded width of gz_l, rows i = the inputs that carry gradient
    {
        const int kd[NL] = {64, 32, 16, 16, 16, 16, 32, 64, OUT8}, ni[NL] = {IN0, 64, 32, 16, 16, 16, 16, 32, 64};
        for (int l = 0; l < NL; ++l) {
            const int IN = g.in[l], OUT = g.out[l], KD = kd[l];
            for (int idx = threadIdx.x; idx < ni[l] * KD; idx += blockDim.x) {
                const int i = idx / KD, o = idx - i * KD;
                WT[bg.woff[l] + idx] = (o < OUT && i < IN) ? p.W[l][o * IN + i] : 0.f;
            }
        }
        if (threadIdx.x == 0)
        {
            for (int s = 0; s < NPROD * NUNIT; ++s) mbar_init(full + s, 1);
            for (int s = 0; s < NPROD * NSLOT; ++s) mbar_init(empty + s, 2);
        }
    }
    __syncthreads();

    const long BF = (long)B * d.F;
    const long nchunks = (BF + ROWS - 1) / ROWS;
    const int nk = (int)((nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x);     // chunks of this CTA: blockIdx.x + k * grid
    float* dst = partials + ((long)blockIdx.x * 2 + AE) * g.flat_total;
    float* scratch = slots;                                            // reused after the main loop: [NCONS][64][32]
    float* bscratch = slots + NCONS * 64 * 32;                         // [NCONS][32][4]

    if (warp >= NPROD) {
        const int cw = warp - NPROD;
        // consumer warps: 0,1 -> u1 (layer 7); 2,3 -> u5 (layer 1); 4,5 -> u6 (layer 0); 6 -> u0; 7 -> u2; 8 -> u3; 9 -> u4
        if (cw < 2)       consumer<7, 0, 64, 8, 8, 4, -1, 0, 0, 0, 0, 0, 16>(1, cw, 2, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        else if (cw < 4)  consumer<1, 0, 32, 8, 8, 8, -1, 0, 0, 0, 0, 0, 8>(5, cw - 2, 2, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        else if (cw < 6)  consumer<0, 0, 64, 8, 8, 4, -1, 0, 0, 0, 0, 0, 16>(6, cw - 4, 2, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        else if (cw == 6) consumer<8, 0, 16, 4, 8, 8, -1, 0, 0, 0, 0, 0, 4>(0, 0, 1, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        else if (cw == 7) consumer<6, 0, 48, 4, 4, 4, 5, 32, 64, 2, 4, 4, 16>(2, 0, 1, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        else if (cw == 8) consumer<4, 0, 16, 4, 4, 8, -1, 0, 0, 0, 0, 0, 4>(3, 0, 1, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        else              consumer<3, 0, 32, 2, 4, 4, 2, 16, 48, 4, 4, 8, 8>(4, 0, 1, lane, nk, slots, full, empty, scratch, bscratch, cw, g, dst, timing);
        return;
    }

    // ------------------------------------------------------------------------------------------------------------
    // producer warp: chunks k = warp, warp + NPROD, ... of this CTA
    // ------------------------------------------------------------------------------------------------------------
    {
        const int pw = warp;
        float* myslots = slots + (long)pw * NSLOT * SLOT_FLOATS;
        uint64_t* myfull = full + pw * NUNIT;
        uint64_t* myempty = empty + pw * NSLOT;
        float* tb = tails + ((long)pw * ROWS + lane) * TLD;
        const int tail0 = d.T - d.OT, rowstride = 2 * d.Fp;
        int j = 0;                                                    // slot-use counter of this producer: kk * NUNIT + unit
        // acquire the slot of use j (wait until its previous occupant was consumed) and return the lane's row in it
#define ST_ACQ(JJ) (mbar_wait(myempty + (JJ) % NSLOT, (uint32_t)((((JJ) / NSLOT) & 1) ^ 1)), myslots + ((JJ) % NSLOT) * SLOT_FLOATS + lane * LD)
#define ST_FULL(U) { ST_T(3) cp_async_wait_all(); ST_T(1) __syncwarp(); if (lane == 0) mbar_arrive(myfull + (U)); }
        for (int k = pw; k < nk; k += NPROD, j += NUNIT) {
            const long R = ((long)blockIdx.x + (long)k * gridDim.x) * ROWS + lane;
            const bool ok = R < BF;
            const int b = ok ? (int)(R / d.F) : 0, f = ok ? (int)(R - (long)b * d.F) : 0;
            const float* rec = save + (ok ? R : 0) * bg.ss;
            float2 gz[16];

            // ---- output side (nn_proc.py:115, 322, 325-326): gz9 and the skip / residual gradient.  Everything it reads
            // from global memory is fetched by ONE batch of asynchronous copies (a single exposed DRAM round trip per
            // chunk); the columns of the next slot that gz8 will occupy serve as the landing zone.
            ST_T0()
            float* s0 = ST_ACQ(j);
            float* s1 = ST_ACQ(j + 1);
            ST_T(0)
            prefetch_row<16>(s0 + 16, rec + bg.soff[7], ok);                                   // h8
            prefetch_row<4>(tb, rec + bg.soff[8], ok);                                          // e9
            prefetch_row<8>(s1 + 64, rec + bg.soff[6], ok);                                    // h7
            {
                const float* gri = g_ri + (long)b * d.OTp * rowstride + f;
                const long oo0 = (long)b * d.OT * d.F + f;
                const float* x3 = AE == 1 ? mag_hat : g_mag_hat;
                for (int jj = 0; jj < d.OT; ++jj) {
                    cp_async4(s1 + jj, gri + (long)jj * rowstride, ok);
                    cp_async4(s1 + 16 + jj, gri + (long)jj * rowstride + d.Fp, ok);
                    cp_async4(s1 + 32 + jj, phs_hat + oo0 + (long)jj * d.F, ok);
                    if (x3) cp_async4(s1 + 48 + jj, x3 + oo0 + (long)jj * d.F, ok);
                    if (AE == 0) cp_async4(tb + 16 + jj, rec + bg.soff_v + tail0 + jj, ok);
                }
            }
            cp_async_wait_all();
            ST_T(1)
#pragma unroll 3
            for (int jj = 0; jj < d.OT; ++jj) {
                float gzv = 0.f;
                if (ok) {
                    const float e9 = tb[jj];
                    const float gre = s1[jj], gim = s1[16 + jj];
                    float sn, cs;
                    sincosf(s1[32 + jj], &sn, &cs);
                    if (AE == 0) {   // an = mag_hat (cos, sin);  mag_hat = ELU(dec) * v_tail
                        float gm = gre * cs + gim * sn;
                        if (g_mag_hat) gm += s1[48 + jj];
                        gzv = gm * tb[16 + jj] * elu_grad(e9);
                        tb[jj] = gm * e9;
                    } else {         // phs_hat = ELU(dec) + phs_tail
                        const float gp = s1[48 + jj] * (gim * cs - gre * sn);
                        gzv = gp * elu_grad(e9);
                        tb[jj] = gp;
                    }
                }
                s0[jj] = gzv;
            }
            for (int jj = d.OT; jj < 16; ++jj) s0[jj] = 0.f;
            ST_T(2)
            ST_FULL(0)

            // ---- layer 9 (fnn_dec): gz9 -> gz8
            reload_row<16>(gz, s0);
            dgrad_layer<OUT8, 64, 8>(WT + bg.woff[8], gz, nullptr, s0 + 16, s1);
            ST_FULL(1)
            // ---- layer 8: gz8 -> gz7
            ST_T(3)
            float* s2 = ST_ACQ(j + 2);
            ST_T(0)
            prefetch_row<4>(s2 + 48, rec + bg.soff[5], ok);                                    // h6
            prefetch_row<4>(s2 + 64, rec + bg.soff[4], ok);                                    // h5
            dgrad_layer<64, 32, 8>(WT + bg.woff[7], gz, s1, s1 + 64, s2);
            ST_T(3)
            cp_async_wait_all();
            ST_T(1)
            // ---- layer 7: gz7 -> gz6
            reload_row<32>(gz, s2);
            dgrad_layer<32, 16, 8>(WT + bg.woff[6], gz, nullptr, s2 + 48, s2 + 32);
            ST_FULL(2)
            // ---- layer 6: gz6 -> gz5
            ST_T(3)
            float* s3 = ST_ACQ(j + 3);
            ST_T(0)
            prefetch_row<8>(s3 + 16, rec + bg.soff[3], ok);                                    // h4 ++ knobs
            reload_row<16>(gz, s2 + 32);
            dgrad_layer<16, 16, 8>(WT + bg.woff[5], gz, nullptr, s2 + 64, s3);
            ST_FULL(3)
            // ---- layer 5 (fnn_addknobs): gz5 -> gz4 (the knob inputs carry no gradient)
            ST_T(3)
            float* s4 = ST_ACQ(j + 4);
            ST_T(0)
            prefetch_row<4>(s4 + 32, rec + bg.soff[2], ok);                                    // h3
            prefetch_row<8>(s4 + 48, rec + bg.soff[1], ok);                                    // h2
            reload_row<16>(gz, s3);
            dgrad_layer<16, 16, 8>(WT + bg.woff[4], gz, nullptr, s3 + 16, s4);
            ST_T(3)
            cp_async_wait_all();
            ST_T(1)
            // ---- layer 4: gz4 -> gz3
            reload_row<16>(gz, s4);
            dgrad_layer<16, 16, 8>(WT + bg.woff[3], gz, nullptr, s4 + 32, s4 + 16);
            ST_FULL(4)
            // ---- layer 3: gz3 -> gz2
            ST_T(3)
            float* s5 = ST_ACQ(j + 5);
            ST_T(0)
            prefetch_row<16>(s5 + 32, rec + bg.soff[0], ok);                                   // h1
            reload_row<16>(gz, s4 + 16);
            dgrad_layer<16, 32, 8>(WT + bg.woff[2], gz, nullptr, s4 + 48, s5);
            ST_FULL(5)
            // ---- layer 2: gz2 -> gz1
            ST_T(3)
            float* s6 = ST_ACQ(j + 6);
            ST_T(0)
            prefetch_row<8>(s6 + 64, rec + bg.soff_v, ok);                                     // track v
            reload_row<32>(gz, s5);
            dgrad_layer<32, 64, 8>(WT + bg.woff[1], gz, nullptr, s5 + 32, s6);
            ST_FULL(6)
            // ---- layer 1 (fnn_enc): gz1 -> dL/d(track), plus the skip / residual gradient on the last OT frames.
            // Stored lane <-> bin (coalesced); ae_input_grad_kernel turns both autoencoders' track gradients into
            // dL/d(re, im) (nn_proc.py:309-310).
            ST_T(3)
            {
                float* gt = g_track + ((long)b * d.T) * d.F + f;
                dgrad_core<64, IN0, 4>(WT + bg.woff[0], gz, s6, [&](int i, const float (&c)[4]) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int tt = i + q;
                        if (ok && tt < d.T) gt[(long)tt * d.F] = c[q] + (tt >= tail0 ? tb[tt - tail0] : 0.f);
                    }
                });
            }
            ST_T(4)
        }
#undef ST_ACQ
#undef ST_FULL
        if (timing && pw == 0 && lane == 0)
            for (int i = 0; i < 6; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(timing) + i, (unsigned long long)treg[i]);
    }
#undef ST_T0
#undef ST_T
    // the consumers' flush phases
    __syncthreads();
    __syncthreads();
}

// dL/d(re, im) from the two track gradients (magnitude and phase autoencoder), stored as the (hi, lo) tf32 pair the
// analysis weight-gradient GEMM consumes.  mag = sqrt(re^2 + im^2) with subgradient 0 at 0 (torch.norm backward,
// nn_proc.py:309); phs = atan2(im, re + 1e-7) (nn_proc.py:310).  g_mag: optional external gradient of the mag output.
__global__ void ae_input_grad_kernel(StDims d, int B, const float* __restrict__ spec, const float* __restrict__ gt_m,
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
        const float gre = scm * re, gim = scm * im;
        st_split_tf32(gre - scp * im, g_spec[o], g_spec_lo[o]);
        st_split_tf32(gim + scp * u, g_spec[o + d.Fp], g_spec_lo[o + d.Fp]);
    }
}

BwdGeom build_bwd_geom(int in0, int out8) {
    BwdGeom bg;
    const int kd[NL] = {64, 32, 16, 16, 16, 16, 32, 64, out8}, ni[NL] = {in0, 64, 32, 16, 16, 16, 16, 32, 64};
    const int soff[NL] = {0, 64, 96, 112, 144, 160, 176, 208, 272};
    int off = 0;
    for (int l = 0; l < NL; ++l) {
        bg.woff[l] = off;
        off += kd[l] * ni[l];
        bg.soff[l] = soff[l];
    }
    bg.wfloats = off;
    bg.soff_v = 272 + 16;
    bg.ss = bg.soff_v + 32;
    return bg;
}

template <int IN0, int OUT8>
int launch_bwd_f2(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, int B,
                  const float* save_m, const float* save_p, const float* mag_hat, const float* phs_hat, const float* g_ri,
                  const float* g_mag_hat, const float* g_mag, float* g_track, float* g_spec, float* g_spec_lo, float* partials,
                  long long* timing, int sm_count, cudaStream_t s) {
    const BwdGeom bg = build_bwd_geom(IN0, OUT8);
    const size_t smem = sizeof(float) * ((size_t)bg.wfloats + (size_t)NPROD * NSLOT * SLOT_FLOATS + (size_t)NPROD * ROWS * TLD) +
                        sizeof(uint64_t) * NPROD * (NSLOT + NUNIT);
    if (smem > 227 * 1024) return 0;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(ae_bwd_f2_kernel<0, IN0, OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return 0;
        if (cudaFuncSetAttribute(ae_bwd_f2_kernel<1, IN0, OUT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return 0;
        configured = true;
    }
    const long nchunks = ((long)B * d.F + ROWS - 1) / ROWS;
    const long ntrk = (long)B * d.T * d.F;
    // The two autoencoders' kernels are independent (separate records, track-gradient halves and partial vectors) and each
    // CTA fills an SM: run them SIDE BY SIDE on half the SMs each instead of back to back on all of them.  Same work per SM,
    // but a CTA then walks twice as many chunks, so the round-up of chunks per producer warp (5.4 -> 6 at B = 200, i.e. 10 %
    // idle) shrinks to 10.8 -> 11 (1.5 %), and the two kernels' prologues and tails overlap.
    static cudaStream_t side = nullptr;
    static cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    static int side_state = 0;                      // 0: not tried, 1: ready, -1: unavailable / switched off
    if (side_state == 0) {
        const char* e = getenv("ST_AE_BWD_ONE_STREAM");
        side_state = (!(e && e[0] == '1') && cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess &&
                      cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming) == cudaSuccess &&
                      cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming) == cudaSuccess) ? 1 : -1;
    }
    const bool two = side_state == 1 && sm_count >= 2;
    const int grid = (int)std::min<long>(nchunks, two ? sm_count / 2 : sm_count);
    cudaStream_t s1 = s;
    if (two) {
        cudaEventRecord(fork_ev, s);
        cudaStreamWaitEvent(side, fork_ev, 0);
        s1 = side;
    }
    ae_bwd_f2_kernel<0, IN0, OUT8><<<grid, WARPS * 32, smem, s>>>(d, g, bg, pm, B, save_m, mag_hat, phs_hat, g_ri, g_mag_hat, g_track,
                                                                 partials, timing);
    ae_bwd_f2_kernel<1, IN0, OUT8><<<grid, WARPS * 32, smem, s1>>>(d, g, bg, pp, B, save_p, mag_hat, phs_hat, g_ri, g_mag_hat,
                                                                  g_track + ntrk, partials, timing ? timing + 8 : nullptr);
    if (two) {
        cudaEventRecord(join_ev, side);
        cudaStreamWaitEvent(s, join_ev, 0);
    }
    ae_input_grad_kernel<<<(int)std::min<long>((ntrk + 255) / 256, 8L * sm_count), 256, 0, s>>>(d, B, spec, g_track, g_track + ntrk, g_mag,
                                                                                              g_spec, g_spec_lo);
    return grid;
}

}  // namespace

// Same contract as st_launch_ae_backward_mma (records in the shared 320-float layout); covers T <= 32, OT <= 16, K <= 16.
// g_track: workspace of 2 * B * T * F floats (track gradients of the two autoencoders).
// Returns the number of per-CTA partial-gradient vectors written (0: geometry not covered).
int st_launch_ae_backward_f2(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, int B,
                             const float* save_m, const float* save_p, const float* mag_hat, const float* phs_hat,
                             const float* g_ri, const float* g_mag_hat, const float* g_mag, float* g_track, float* g_spec,
                             float* g_spec_lo, float* partials, long long* timing, int sm_count, cudaStream_t s) {
    if (d.T > 32 || d.OT > 16 || d.K > 16) return 0;
    if (st_ae_mma_record_floats(d) != 320) return 0;
    if (d.T <= 28 && d.OT <= 12)
        return launch_bwd_f2<28, 12>(d, g, pm, pp, spec, B, save_m, save_p, mag_hat, phs_hat, g_ri, g_mag_hat, g_mag, g_track, g_spec,
                                     g_spec_lo, partials, timing, sm_count, s);
    return launch_bwd_f2<32, 16>(d, g, pm, pp, spec, B, save_m, save_p, mag_hat, phs_hat, g_ri, g_mag_hat, g_mag, g_track, g_spec,
                                 g_spec_lo, partials, timing, sm_count, s);
}
