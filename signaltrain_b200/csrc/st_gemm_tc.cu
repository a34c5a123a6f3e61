// tcgen05 GEMM for the five front-end contractions (analysis, synthesis, and their data / weight gradients).
//
//   C[M,N] (+ split-K planes) = A * B    fp32 in, fp32 out, "3xTF32": every operand is carried as an exact pair
//   x = hi + lo of tf32-representable fp32 words (written by the producing kernel), and
//       A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi            (three kind::tf32 UMMAs per k-step, fp32 accumulate in TMEM)
//   which restores fp32-level accuracy (the dropped lo*lo term is 2^-22 relative) -- single-pass TF32 misses the 1e-5
//   waveform bar by 40x.
//
// Blackwell structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer : cp.async.bulk.tensor.2d tiles of A_hi/A_lo/B_hi/B_lo -> 128B-swizzled smem ring, mbarrier tx
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=8) from smem descriptors into a
//                              double-buffered TMEM accumulator; tcgen05.commit releases smem stages / publishes the tile
//   warps 2..9  epilogue     : tcgen05.ld (32 lanes x 32b) -> fp32 registers -> vectorised global stores.  TMEM accumulation
//                              rounds toward zero at every MMA step (measured: 7e-6 relative after K=1024), so for the two
//                              forward GEMMs (1e-5 waveform budget) the issuer starts a fresh TMEM accumulator every k-block
//                              and these warps add the 12-step partial sums in round-to-nearest fp32 registers
//                              ("promoted accumulation"); the gradient GEMMs accumulate all of K in TMEM.
// Two kernels share this structure: gemm_tc2_kernel (default) runs it on CTA PAIRS -- tcgen05.mma.cta_group::2, a 256 x BN tile
// per pair, each CTA staging its 128 rows of A and half of B (see the comment above that kernel) -- and gemm_tc_kernel is the
// 1-CTA form (128 x BN tile; ST_GEMM_PAIR=0, shapes with a single M-tile, or a refused cluster launch).  PASSES = 1 is the
// reduced-precision mode: hi planes only, one UMMA per k-step.
// Operands may be K-major ([row][k], one 128-row x 32-float box per stage) or MN-major ([k][row], 32x32 boxes); the frame
// gather of Conv1d / ConvTranspose1d (cls_fe_dft.py:28-31,78-82) is a 2-D tensor map whose row stride is the hop (rows
// overlap), so frames are never materialised.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "st_common.cuh"
#include "st_tc_prims.cuh"

namespace {

using namespace st_tc;

constexpr int BM = 128;
constexpr int BKF = 32;                 // floats per k-block = one 128-byte swizzle span
constexpr int NTHREADS = 320;           // 1 TMA warp + 1 MMA warp + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr uint32_t A_BYTES = BM * 128;  // one (hi or lo) A tile per stage
constexpr int TMEM_COLS = 512;
#ifndef ST_GEMM_PAIR_DEFAULT
#define ST_GEMM_PAIR_DEFAULT 1      // cta_group::2 tiles on by default (measured: the five GEMMs 0.267 -> 0.249 ms; ST_GEMM_PAIR=0 = 1-CTA kernel)
#endif

// Shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64).
//   K-major  tile [rows][32 floats], SWIZZLE_128B (type 2): 8-row x 128 B atoms, SBO = 1024 B, LBO unused (1)
//   MN-major tile, 32-column slabs of [32 k-rows][128 B].  32-bit MN-major operands only exist in the
//   SWIZZLE_128B_BASE32B layout (type 1: 32-byte swizzle atoms, 4-row x 128 B atoms; TMA: SWIZZLE_128B_ATOM_32B):
//   SBO = 512 B (next 4 k-rows), LBO = 4096 B (next 32-column slab)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, bool mn_major) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(mn_major ? (4096 >> 4) : 1) << 16;
    d |= (uint64_t)((mn_major ? 512 : 1024) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(mn_major ? 1 : 2) << 61;
    return d;
}

// ---- 2-CTA cluster variant: the pair works on two M-tiles of the same N-tile, each CTA fetches HALF of the B tile and
// multicasts it into both CTAs' shared memory (half the L2 reads for B; the GEMMs are operand-bandwidth bound because every
// k-block moves hi and lo planes of both operands).  The MMA stays cta_group::1; only the B loads, the stage-release barrier
// (both CTAs' MMAs must have retired before either overwrites a stage) and cluster syncs at both ends differ.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct TcParams {
    float* C;
    long ldc;
    int M, N, BN;
    int tiles_m, tiles_n, splits;
    int kb_total, kb_per_split;    // k-blocks of 32
    int kb_per_chunk;              // k-blocks accumulated in TMEM before promotion to fp32 registers
    long split_stride;
    int stages;
    int passes;                    // 3: exact (hi, lo) pairs, 3xTF32;  1: hi planes only, one kind::tf32 UMMA per k-step
};

template <bool A_MN, bool B_MN, bool MC, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    constexpr bool exact = PASSES == 3;        // single pass: a stage holds the hi planes only (half the bytes, more stages)
    const uint32_t stage_bytes = (exact ? 2u : 1u) * (A_BYTES + b_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                     // [stages]  TMA -> MMA
    uint64_t* empty = bars + p.stages;         // [stages]  MMA -> TMA
    uint64_t* tfull = bars + 2 * p.stages;     // [2]       MMA -> epilogue
    uint64_t* tempty = tfull + 2;              // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MC ? 2 : 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) cluster_sync_all();                // the peer's barriers are initialised before anything arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // work items: (M-tile, N-tile, split); in cluster mode the pair shares an item of (M-tile PAIR, N-tile, split) and CTA
    // `rank` takes M-tile 2 * pair + rank (a pair's second tile may lie past M: loads zero-fill, stores are guarded)
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    const int tiles_mw = MC ? (p.tiles_m + 1) / 2 : p.tiles_m;
    const int total_work = tiles_mw * p.tiles_n * p.splits;
    const int w_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
#define ST_TILE_M0(TT) (((MC ? 2 * ((TT) / p.tiles_n) + (int)rank : (TT) / p.tiles_n)) * BM)

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = w_first; w < total_work; w += w_step) {
                const int split = w % p.splits, tt = w / p.splits;
                const int m0 = ST_TILE_M0(tt), n0 = (tt % p.tiles_n) * p.BN;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA_hi = smem + (size_t)stage * stage_bytes;
                    uint8_t* sA_lo = sA_hi + A_BYTES;                         // (exact only)
                    uint8_t* sB_hi = sA_hi + (exact ? 2u : 1u) * A_BYTES;
                    uint8_t* sB_lo = sB_hi + b_bytes;                         // (exact only)
                    mbar_expect_tx(&full[stage], stage_bytes);
                    const int k0 = kb * BKF;
                    if (!A_MN) {
                        tma_load_2d(sA_hi, &tmAh, &full[stage], k0, m0);
                        if (exact) tma_load_2d(sA_lo, &tmAl, &full[stage], k0, m0);
                    } else {
#pragma unroll
                        for (int g = 0; g < BM / 32; ++g) {
                            tma_load_2d(sA_hi + g * 4096, &tmAh, &full[stage], m0 + 32 * g, k0);
                            if (exact) tma_load_2d(sA_lo + g * 4096, &tmAl, &full[stage], m0 + 32 * g, k0);
                        }
                    }
                    if (MC) {                 // my half of the B tile, into both CTAs (the peer sends the other half)
                        if (!B_MN) {
                            const int hr = p.BN >> 1;                       // rows per half (a multiple of the 8-row swizzle atom)
                            tma_load_2d_mc(sB_hi + rank * hr * 128, &tmBh, &full[stage], k0, n0 + (int)rank * hr, 0x3);
                            if (exact) tma_load_2d_mc(sB_lo + rank * hr * 128, &tmBl, &full[stage], k0, n0 + (int)rank * hr, 0x3);
                        } else {
                            for (int g = (int)rank; g < p.BN / 32; g += 2) {
                                tma_load_2d_mc(sB_hi + g * 4096, &tmBh, &full[stage], n0 + 32 * g, k0, 0x3);
                                if (exact) tma_load_2d_mc(sB_lo + g * 4096, &tmBl, &full[stage], n0 + 32 * g, k0, 0x3);
                            }
                        }
                    } else if (!B_MN) {
                        tma_load_2d(sB_hi, &tmBh, &full[stage], k0, n0);
                        if (exact) tma_load_2d(sB_lo, &tmBl, &full[stage], k0, n0);
                    } else {
                        for (int g = 0; g < p.BN / 32; ++g) {
                            tma_load_2d(sB_hi + g * 4096, &tmBh, &full[stage], n0 + 32 * g, k0);
                            if (exact) tma_load_2d(sB_lo + g * 4096, &tmBl, &full[stage], n0 + 32 * g, k0);
                        }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=TF32 [7,10),[10,13), majors [15],[16],
            // N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0, 0};
            for (int w = w_first; w < total_work; w += w_step) {
                const int split = w % p.splits;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kc = kb0; kc < kb1; kc += p.kb_per_chunk) {
                    const int kce = min(kb1, kc + p.kb_per_chunk);
                    mbar_wait(&tempty[acc], acc_phase[acc] ^ 1);          // epilogue has drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
                    uint32_t accumulate = 0;
                    for (int kb = kc; kb < kce; ++kb) {
                        mbar_wait(&full[stage], phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sA_hi = smem_u32(smem + (size_t)stage * stage_bytes);
                        const uint32_t sA_lo = sA_hi + A_BYTES, sB_hi = sA_hi + (exact ? 2u : 1u) * A_BYTES, sB_lo = sB_hi + b_bytes;
#pragma unroll
                        for (int ks = 0; ks < BKF / 8; ++ks) {
                            const uint32_t ao = A_MN ? ks * 1024 : ks * 32;   // next 8 k: 8 rows of 128 B | 32 B inside the span
                            const uint32_t bo = B_MN ? ks * 1024 : ks * 32;
                            const uint64_t ah = make_desc(sA_hi + ao, A_MN), al = make_desc(sA_lo + ao, A_MN);
                            const uint64_t bh = make_desc(sB_hi + bo, B_MN), bl = make_desc(sB_lo + bo, B_MN);
                            if (exact) {
                                umma_tf32(tmem_d, al, bh, idesc, accumulate);
                                umma_tf32(tmem_d, ah, bl, idesc, 1u);
                                umma_tf32(tmem_d, ah, bh, idesc, 1u);
                            } else {
                                umma_tf32(tmem_d, ah, bh, idesc, accumulate);
                            }
                            accumulate = 1u;
                        }
                        if (MC) umma_commit_mc(&empty[stage], 0x3);        // ... in BOTH CTAs: each also writes the other's stage
                        else umma_commit(&empty[stage]);                   // smem stage reusable once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);                              // partial accumulator complete -> epilogue
                    acc_phase[acc] ^= 1;
                    acc ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue ====================================
        const int q = warp & 3;                        // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;              // which half of the tile's columns this warp owns
        const int ncols = p.BN >> 1;                   // multiple of 8, <= 128
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        for (int w = w_first; w < total_work; w += w_step) {
            const int split = w % p.splits, tt = w / p.splits;
            const int m0 = ST_TILE_M0(tt), n0 = (tt % p.tiles_n) * p.BN + half * ncols;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
            float sum[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) sum[i] = 0.f;
            for (int kc = kb0; kc < kb1; kc += p.kb_per_chunk) {
                mbar_wait(&tfull[acc], acc_phase[acc]);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * p.BN + half * ncols);
#pragma unroll
                for (int c = 0; c < 128; c += 32) {
                    if (c < ncols) {
                        uint32_t r[4][8];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + 8 * u < ncols) tmem_ld8(taddr + c + 8 * u, r[u]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + 8 * u < ncols) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) sum[c + 8 * u + i] += __uint_as_float(r[u][i]);
                            }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                acc_phase[acc] ^= 1;
                acc ^= 1;
            }
            const int row = m0 + 32 * q + lane;
            if (row < p.M) {
                float* crow = p.C + (long)split * p.split_stride + (long)row * p.ldc + n0;
#pragma unroll
                for (int c = 0; c < 128; c += 4)
                    if (c < ncols && n0 + c < p.N)
                        *reinterpret_cast<float4*>(crow + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
            }
        }
    }
#undef ST_TILE_M0
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) cluster_sync_all();                // nobody leaves while the peer may still write its shared memory or barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- cta_group::2 variant ---------------------------------------------------------------------------
// A CTA pair (cluster of 2, one TPC) computes a 256 x BN tile: each CTA stages ITS 128 rows of A and ITS half of the B tile
// (BN/2 rows), the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256, and the tensor cores of both SMs
// read both CTAs' shared memory -- per CTA and k-block the operand bytes drop from A + B to A + B/2, which is what bounds
// these GEMMs (every k-block moves fp32 hi and lo planes of both operands).  Protocol differences to the 1-CTA kernel:
//   * both CTAs' TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2, barrier address
//     mapped to rank 0), whose expected byte count covers both halves;
//   * tcgen05.commit.cta_group::2 multicasts "stage free" and "accumulator ready" to the barriers of both CTAs;
//   * each CTA's epilogue drains its own 128 TMEM lanes and both arrive on the leader's "accumulator drained" barrier;
//   * TMEM is allocated / freed with cta_group::2 by the same warp of both CTAs.
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) __trap();
    }
}

template <bool A_MN, bool B_MN, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int hn = p.BN >> 1;                                  // B rows (N extent) staged by each CTA
    const uint32_t b_bytes = (uint32_t)hn * 128u;
    constexpr bool exact = PASSES == 3;
    constexpr uint32_t planes = exact ? 2u : 1u;
    const uint32_t stage_bytes = planes * (A_BYTES + b_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                     // [stages]  leader only: both CTAs' TMA bytes -> MMA
    uint64_t* empty = bars + p.stages;         // [stages]  per CTA: MMA (leader's commit, multicast) -> this CTA's TMA
    uint64_t* tfull = bars + 2 * p.stages;     // [2]       per CTA: MMA -> this CTA's epilogue
    uint64_t* tempty = tfull + 2;              // [2]       leader only: both CTAs' epilogues -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();                        // both CTAs' barriers exist before anything remote arrives on them
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // work items shared by the pair: (M-tile PAIR, N-tile, split); CTA `rank` owns M-tile 2 * pair + rank
    const int tiles_mw = (p.tiles_m + 1) / 2;
    const int total_work = tiles_mw * p.tiles_n * p.splits;
    const int w_first = (int)(blockIdx.x >> 1), w_step = (int)(gridDim.x >> 1);
#define ST_TILE_M0(TT) ((2 * ((TT) / p.tiles_n) + (int)rank) * BM)

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = w_first; w < total_work; w += w_step) {
                const int split = w % p.splits, tt = w / p.splits;
                const int m0 = ST_TILE_M0(tt), n0 = (tt % p.tiles_n) * p.BN + (int)rank * hn;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA_hi = smem + (size_t)stage * stage_bytes;
                    uint8_t* sA_lo = sA_hi + A_BYTES;
                    uint8_t* sB_hi = sA_hi + planes * A_BYTES;
                    uint8_t* sB_lo = sB_hi + b_bytes;
                    if (rank == 0) mbar_expect_tx(&full[stage], 2u * stage_bytes);      // my bytes + the peer's
                    const uint32_t fb = mapa_rank(smem_u32(&full[stage]), 0);
                    const int k0 = kb * BKF;
                    if (!A_MN) {
                        tma_load_2d_pair(sA_hi, &tmAh, fb, k0, m0);
                        if (exact) tma_load_2d_pair(sA_lo, &tmAl, fb, k0, m0);
                    } else {
#pragma unroll
                        for (int g = 0; g < BM / 32; ++g) {
                            tma_load_2d_pair(sA_hi + g * 4096, &tmAh, fb, m0 + 32 * g, k0);
                            if (exact) tma_load_2d_pair(sA_lo + g * 4096, &tmAl, fb, m0 + 32 * g, k0);
                        }
                    }
                    if (!B_MN) {
                        tma_load_2d_pair(sB_hi, &tmBh, fb, k0, n0);
                        if (exact) tma_load_2d_pair(sB_lo, &tmBl, fb, k0, n0);
                    } else {
                        for (int g = 0; g < hn / 32; ++g) {
                            tma_load_2d_pair(sB_hi + g * 4096, &tmBh, fb, n0 + 32 * g, k0);
                            if (exact) tma_load_2d_pair(sB_lo + g * 4096, &tmBl, fb, n0 + 32 * g, k0);
                        }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) =================
        if (lane == 0 && rank == 0) {
            // M = 256 over the pair (M>>4 at [24,29)), N = BN: each CTA supplies BN/2 rows of B
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0, 0};
            for (int w = w_first; w < total_work; w += w_step) {
                const int split = w % p.splits;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kc = kb0; kc < kb1; kc += p.kb_per_chunk) {
                    const int kce = min(kb1, kc + p.kb_per_chunk);
                    mbar_wait_cluster(&tempty[acc], acc_phase[acc] ^ 1);  // both CTAs' epilogues have drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
                    uint32_t accumulate = 0;
                    for (int kb = kc; kb < kce; ++kb) {
                        mbar_wait(&full[stage], phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sA_hi = smem_u32(smem + (size_t)stage * stage_bytes);
                        const uint32_t sA_lo = sA_hi + A_BYTES, sB_hi = sA_hi + planes * A_BYTES, sB_lo = sB_hi + b_bytes;
#pragma unroll
                        for (int ks = 0; ks < BKF / 8; ++ks) {
                            const uint32_t ao = A_MN ? ks * 1024 : ks * 32;
                            const uint32_t bo = B_MN ? ks * 1024 : ks * 32;
                            const uint64_t ah = make_desc(sA_hi + ao, A_MN), al = make_desc(sA_lo + ao, A_MN);
                            const uint64_t bh = make_desc(sB_hi + bo, B_MN), bl = make_desc(sB_lo + bo, B_MN);
                            if (exact) {
                                umma_tf32_pair(tmem_d, al, bh, idesc, accumulate);
                                umma_tf32_pair(tmem_d, ah, bl, idesc, 1u);
                                umma_tf32_pair(tmem_d, ah, bh, idesc, 1u);
                            } else {
                                umma_tf32_pair(tmem_d, ah, bh, idesc, accumulate);
                            }
                            accumulate = 1u;
                        }
                        umma_commit_pair(&empty[stage]);                   // both CTAs may refill this stage once the MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_pair(&tfull[acc]);                         // accumulator complete -> both epilogues
                    acc_phase[acc] ^= 1;
                    acc ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue (each CTA: its own 128 rows) =========
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int ncols = p.BN >> 1;
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        const uint32_t te0 = mapa_rank(smem_u32(&tempty[0]), 0), te1 = mapa_rank(smem_u32(&tempty[1]), 0);
        for (int w = w_first; w < total_work; w += w_step) {
            const int split = w % p.splits, tt = w / p.splits;
            const int m0 = ST_TILE_M0(tt), n0 = (tt % p.tiles_n) * p.BN + half * ncols;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
            float sum[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) sum[i] = 0.f;
            for (int kc = kb0; kc < kb1; kc += p.kb_per_chunk) {
                mbar_wait(&tfull[acc], acc_phase[acc]);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * p.BN + half * ncols);
#pragma unroll
                for (int c = 0; c < 128; c += 32) {
                    if (c < ncols) {
                        uint32_t r[4][8];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + 8 * u < ncols) tmem_ld8(taddr + c + 8 * u, r[u]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + 8 * u < ncols) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) sum[c + 8 * u + i] += __uint_as_float(r[u][i]);
                            }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc ? te1 : te0);
                acc_phase[acc] ^= 1;
                acc ^= 1;
            }
            const int row = m0 + 32 * q + lane;
            if (row < p.M) {
                float* crow = p.C + (long)split * p.split_stride + (long)row * p.ldc + n0;
#pragma unroll
                for (int c = 0; c < 128; c += 4)
                    if (c < ncols && n0 + c < p.N)
                        *reinterpret_cast<float4*>(crow + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
            }
        }
    }
#undef ST_TILE_M0
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                        // the peer's MMAs / barrier traffic that touch this CTA have all retired
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D fp32 row-major view: element (r, c) at base + r*ld + c; rows may overlap (ld < cols).  Box = {32 floats, box_rows}.
bool make_map(CUtensorMap* tm, const float* base, long rows, long cols, long ld, int box_rows, bool mn_major) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <bool A_MN, bool B_MN, bool MC, int PASSES>
cudaError_t launch_p(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcParams& p,
                   int grid, size_t smem, cudaStream_t s) {
    static bool configured_on[ST_MAX_DEVICES] = {};              // function attributes are per device (context), not per process
    bool& configured = configured_on[st_current_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, MC, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (!MC) {
        gemm_tc_kernel<A_MN, B_MN, MC, PASSES><<<grid, NTHREADS, smem, s>>>(ah, al, bh, bl, p);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<A_MN, B_MN, MC, PASSES>, ah, al, bh, bl, p);
}

template <bool A_MN, bool B_MN, int PASSES>
cudaError_t launch_pair_p(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcParams& p,
                        int grid, size_t smem, cudaStream_t s) {
    static bool configured_on[ST_MAX_DEVICES] = {};              // function attributes are per device (context), not per process
    bool& configured = configured_on[st_current_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<A_MN, B_MN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<A_MN, B_MN, PASSES>, ah, al, bh, bl, p);
}

template <bool A_MN, bool B_MN, bool MC>
cudaError_t launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcParams& p,
                   int grid, size_t smem, cudaStream_t s) {
    return p.passes == 3 ? launch_p<A_MN, B_MN, MC, 3>(ah, al, bh, bl, p, grid, smem, s)
                         : launch_p<A_MN, B_MN, MC, 1>(ah, al, bh, bl, p, grid, smem, s);
}
template <bool A_MN, bool B_MN>
cudaError_t launch_pair(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcParams& p,
                        int grid, size_t smem, cudaStream_t s) {
    return p.passes == 3 ? launch_pair_p<A_MN, B_MN, 3>(ah, al, bh, bl, p, grid, smem, s)
                         : launch_pair_p<A_MN, B_MN, 1>(ah, al, bh, bl, p, grid, smem, s);
}

}  // namespace

// Largest multiple of 16 in [64, 256] that divides n (0 if none).
int st_tc_pick_bn(int n) {
    for (int bn = 256; bn >= 64; bn -= 16)
        if (n % bn == 0) return bn;
    return 0;
}

// N-tile width for a given problem: among the multiples of 16 (32 for MN-major B) in [64, 256] that divide N, the one with
// the fewest rounds x width over the SMs (a persistent CTA per SM walks ceil(work / SMs) tiles whose duration scales with
// the width); ties go to the wider tile (fewer operand re-reads).  The synthesis GEMM (18 M-tiles x N = 1024) thus gets
// 144 tiles of 128 instead of 72 of 256, which left half the SMs idle.
static int pick_bn_for(int n, int tiles_m, int splits, bool b_mn, int sm_count, int mn_mult = 32) {
    int best = 0;
    long best_cost = 0;
    for (int bn = 256; bn >= 64; bn -= 16) {
        if (n % bn || (b_mn && (bn % mn_mult))) continue;
        const long work = (long)tiles_m * (n / bn) * splits;
        const long cost = (work + sm_count - 1) / sm_count * bn;
        if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
    }
    return best;
}

// cta_group::2 kernel: tile width and split-K count together -- minimise the operand bytes one CTA pulls in over its rounds of
// work items (A tile + half B tile per k-block, plus the tile's store), the quantity these GEMMs are bound by.  BN = 0: no width
// fits (N not divisible, or half the tile is not a whole number of swizzle atoms / 32-column slabs).
static void pick_pair_tile(int N, int pairs_m, int npairs, int kb_total, int splits, bool b_mn, int* bn_out, int* splits_out) {
    int BN2 = 0, SP2 = 1;
    double best = 0;
    for (int bn = 256; bn >= 64; bn -= 16) {
        if (N % bn || (b_mn ? (bn % 64) : ((bn / 2) % 8))) continue;
        for (int sp = splits; sp >= 1; --sp) {
            const int kbps = (kb_total + sp - 1) / sp;
            if ((kb_total + kbps - 1) / kbps != sp) continue;
            if (sp < splits && 2 * sp < splits) break;            // keep at least half the requested split-K parallelism
            const long work = (long)pairs_m * (N / bn) * sp;
            const long rounds = (work + npairs - 1) / npairs;
            const double cost = (double)rounds * ((double)kbps * (A_BYTES + bn / 2 * 128) + 0.5 * BM * bn * 4 + 8192.0);
            if (BN2 == 0 || cost < best) { BN2 = bn; SP2 = sp; best = cost; }
        }
    }
    *bn_out = BN2;
    *splits_out = SP2;
}

// Host-only view of the launch plan (no CUDA call): which kernel a shape gets, its tile width, split-K planes and grid.
// out = {pair (1 | 0), BN, splits, grid}; returns 0, or -1 when the tensor-core path does not cover the shape.
int st_tc_plan(bool b_mn, int M, int N, int K, int splits, int sm_count, int out[4]) {
    if (splits < 1) splits = 1;
    const int tiles_m = (M + BM - 1) / BM, kb_total = (K + BKF - 1) / BKF;
    if (ST_GEMM_PAIR_DEFAULT == 1 && sm_count >= 2 && tiles_m >= 2) {
        const int pairs_m = (tiles_m + 1) / 2, npairs = sm_count / 2;
        int bn = 0, sp = 1;
        pick_pair_tile(N, pairs_m, npairs, kb_total, splits, b_mn, &bn, &sp);
        if (bn > 0) {
            const long work = (long)pairs_m * (N / bn) * sp;
            out[0] = 1; out[1] = bn; out[2] = sp; out[3] = 2 * (int)std::min<long>(work, npairs);
            return 0;
        }
    }
    const int bn = pick_bn_for(N, tiles_m, splits, b_mn, sm_count > 0 ? sm_count : 148);
    if (bn == 0) return -1;
    const int kbps = (kb_total + splits - 1) / splits, sp = (kb_total + kbps - 1) / kbps;
    out[0] = 0; out[1] = bn; out[2] = sp; out[3] = (int)std::min<long>((long)tiles_m * (N / bn) * sp, sm_count);
    return 0;
}

// A: K-major -> A.rows = M, A.cols = K;  MN-major -> A.rows = K, A.cols = M.   Same for B with N.
// Returns the number of split-K planes written, or -1 if this shape cannot take the tensor-core path.
int st_launch_gemm_tc(bool a_mn, bool b_mn, const TcOperand& A, const TcOperand& B, float* C, long ldc, int M, int N, int K,
                      int splits, long split_stride, bool promote, int sm_count, cudaStream_t s, int passes) {
    if (passes != 1) passes = 3;
    if (splits < 1) splits = 1;
    // cta_group::2 tiles (256 x BN per CTA pair), ST_GEMM_PAIR=0 disables
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("ST_GEMM_PAIR"); pair_env = (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : ST_GEMM_PAIR_DEFAULT; }
    if (pair_env == 1 && sm_count >= 2 && (M + BM - 1) / BM >= 2 && !(A.ld & 3) && !(B.ld & 3) && !(ldc & 3) && !(N & 3) &&
        (a_mn || K % BKF == 0) && (b_mn || K % BKF == 0)) {
        const int tiles_m = (M + BM - 1) / BM, pairs_m = (tiles_m + 1) / 2, npairs = sm_count / 2;
        const int kb_total = (K + BKF - 1) / BKF;
        int BN2 = 0, SP2 = 1;
        pick_pair_tile(N, pairs_m, npairs, kb_total, splits, b_mn, &BN2, &SP2);
        if (BN2 > 0) {
            TcParams p;
            p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.BN = BN2;
            p.tiles_m = tiles_m;
            p.tiles_n = N / BN2;
            p.kb_total = kb_total;
            p.kb_per_split = (p.kb_total + SP2 - 1) / SP2;
            p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
            p.split_stride = split_stride;
            // promoted accumulation every TWO k-blocks (24 MMA steps): the pair's "accumulator drained" handshake crosses SMs
            p.kb_per_chunk = promote ? 2 : p.kb_per_split;
            p.passes = passes;
            const size_t sb = (passes == 3 ? 2 : 1) * ((size_t)A_BYTES + (size_t)(BN2 / 2) * 128);
            p.stages = (int)std::min<size_t>(6, (227 * 1024 - 2048) / sb);
            const size_t smem2 = (size_t)p.stages * sb + 1024 + 256;
            CUtensorMap ah, al, bh, bl;
            const int abox = a_mn ? 32 : BM, bbox = b_mn ? 32 : BN2 / 2;
            if (p.stages >= 2 && make_map(&ah, A.hi, A.rows, A.cols, A.ld, abox, a_mn) && make_map(&al, A.lo, A.rows, A.cols, A.ld, abox, a_mn) &&
                make_map(&bh, B.hi, B.rows, B.cols, B.ld, bbox, b_mn) && make_map(&bl, B.lo, B.rows, B.cols, B.ld, bbox, b_mn)) {
                const int work = pairs_m * p.tiles_n * p.splits;
                const int grid = 2 * std::min(work, npairs);
                cudaError_t e;
                if (!a_mn && !b_mn) e = launch_pair<false, false>(ah, al, bh, bl, p, grid, smem2, s);
                else if (!a_mn && b_mn) e = launch_pair<false, true>(ah, al, bh, bl, p, grid, smem2, s);
                else if (a_mn && !b_mn) e = launch_pair<true, false>(ah, al, bh, bl, p, grid, smem2, s);
                else e = launch_pair<true, true>(ah, al, bh, bl, p, grid, smem2, s);
                if (e == cudaSuccess) return p.splits;
                cudaGetLastError();          // a refused cluster launch is not fatal: the 1-CTA kernel below takes the shape
            }
        }
    }
    const int BN = pick_bn_for(N, (M + BM - 1) / BM, splits, b_mn, sm_count > 0 ? sm_count : 148);
    if (BN == 0) return -1;
    if ((A.ld & 3) || (B.ld & 3) || (ldc & 3) || (N & 3)) return -1;
    if (!a_mn && (K % BKF)) return -1;           // K-major operands are not zero-filled along K by a row bound
    if (!b_mn && (K % BKF)) return -1;
    TcParams p;
    p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.BN = BN;
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = N / BN;
    p.kb_total = (K + BKF - 1) / BKF;
    if (splits < 1) splits = 1;
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.split_stride = split_stride;
    p.kb_per_chunk = promote ? 1 : p.kb_per_split;
    p.passes = passes;
    const size_t stage_bytes = (passes == 3 ? 2 : 1) * ((size_t)A_BYTES + (size_t)BN * 128);
    p.stages = (int)std::min<size_t>(passes == 3 ? 4 : 5, (227 * 1024 - 2048) / stage_bytes);
    if (p.stages < 2) return -1;
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
    // cluster (B-multicast) mode, opt-in (ST_GEMM_MULTICAST=1): measured no gain on the K-major GEMMs and a 25-35 % loss on the
    // MN-major ones (the analysis GEMM takes 81 us with or without it): halving the L2 reads of B does not help because the
    // bytes landing in each SM's shared memory are unchanged -- the limit is per-SM operand ingress / TMA latency with two
    // 78 KB stages, which only cta_group::2 tiles (half of B per CTA) would lower.  Needs an even slab count for MN-major B.
    static int mc_env = -1;
    if (mc_env < 0) { const char* e = getenv("ST_GEMM_MULTICAST"); mc_env = (e && e[0] == '1') ? 1 : 0; }
    const bool mc = mc_env == 1 && sm_count >= 2 && p.tiles_m >= 2 && (b_mn ? ((BN / 32) % 2 == 0) : ((BN / 2) % 8 == 0));
    CUtensorMap ah, al, bh, bl;
    const int abox = a_mn ? 32 : BM, bbox = b_mn ? 32 : (mc ? BN / 2 : BN);
    if (!make_map(&ah, A.hi, A.rows, A.cols, A.ld, abox, a_mn) || !make_map(&al, A.lo, A.rows, A.cols, A.ld, abox, a_mn) ||
        !make_map(&bh, B.hi, B.rows, B.cols, B.ld, bbox, b_mn) || !make_map(&bl, B.lo, B.rows, B.cols, B.ld, bbox, b_mn))
        return -1;
    cudaError_t e;
    if (mc) {
        const int pairs = (p.tiles_m + 1) / 2 * p.tiles_n * p.splits;
        const int grid = 2 * std::min(pairs, sm_count / 2);
        if (!a_mn && !b_mn) e = launch<false, false, true>(ah, al, bh, bl, p, grid, smem, s);
        else if (!a_mn && b_mn) e = launch<false, true, true>(ah, al, bh, bl, p, grid, smem, s);
        else if (a_mn && !b_mn) e = launch<true, false, true>(ah, al, bh, bl, p, grid, smem, s);
        else e = launch<true, true, true>(ah, al, bh, bl, p, grid, smem, s);
    } else {
        const int work = p.tiles_m * p.tiles_n * p.splits;
        const int grid = std::min(work, sm_count);
        if (!a_mn && !b_mn) e = launch<false, false, false>(ah, al, bh, bl, p, grid, smem, s);
        else if (!a_mn && b_mn) e = launch<false, true, false>(ah, al, bh, bl, p, grid, smem, s);
        else if (a_mn && !b_mn) e = launch<true, false, false>(ah, al, bh, bl, p, grid, smem, s);
        else e = launch<true, true, false>(ah, al, bh, bl, p, grid, smem, s);
    }
    return e == cudaSuccess ? p.splits : -1;
}
