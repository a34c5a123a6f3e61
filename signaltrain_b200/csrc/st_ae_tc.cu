// Autoencoder forward on tcgen05 (5th-gen tensor cores + TMEM).  [first landing: forward pass, T <= 32, OT <= 16]
//
// A CTA keeps TWO independent streams of 128-row tiles in flight.  Per stream and per layer:
//   * the layer's input A (128 rows x K, as an exact tf32 (hi, lo) pair) sits in shared memory in the K-major
//     SWIZZLE_128B layout tcgen05 reads directly; the weights of all nine layers (hi, lo, same layout) are staged once;
//   * ONE thread (the MMA warp) issues the layer as tcgen05.mma.kind::tf32, M = 128, N = layer width (16..64), three
//     MMAs per 8-wide k-step (a_lo*w_hi, a_hi*w_lo, a_hi*w_hi), accumulator in TMEM (64 columns per stream);
//   * the stream's epilogue warpgroup (128 threads, thread = row = TMEM lane) pulls the row with tcgen05.ld, applies
//     bias + ELU, splits to (hi, lo) and writes the next layer's A straight back into the swizzled tile (and, when
//     training, the row's slot of the saved-activation record); fence.proxy.async + mbarrier hand it to the MMA warp.
// While one stream's epilogue warpgroup does its CUDA-core work the other stream's MMAs run, so the tensor pipe and
// the epilogue overlap without any block-wide barrier.  Compared with the warp-level mma.sync kernel this removes all
// fragment loads / per-fragment operand splitting from the CUDA cores (about 5x fewer issue slots per row).
//
// Reference semantics: AsymAutoEncoder.forward (nn_proc.py:77-126) for both autoencoders, plus the prologue
// (mag / phase, :309-310) and epilogue (skip-filter :115, phase residual :322, polar->rect :325-326) of AsymMPAEC.forward.
#include <algorithm>

#include "st_common.cuh"
#include "st_tc_prims.cuh"

namespace {

using namespace st_tc;

constexpr int NL = ST_AE_LAYERS;
constexpr int TILE = 128;                  // rows per tile = TMEM lanes
constexpr int NSTREAM = 2;
constexpr int THREADS = 32 + NSTREAM * 128;   // warp 0: MMA issuer; warps 1-4 / 5-8: epilogue warpgroups
constexpr int KB_FLOATS = TILE * 32;       // one 32-column K-block of an A tile (16 KB)
constexpr int TMEM_COLS_AE = 128;          // 64 accumulator columns per stream

struct TcAeGeom {
    int n[NL];        // layer width padded to a multiple of 16 (UMMA N)
    int kb[NL];       // 32-float K-blocks of the layer's input (1 or 2)
    int woff[NL];     // float offset of the layer's W_hi slab(s) inside the weight block (W_lo at + wfloats)
    int boff[NL];     // float offset of the bias
    int wfloats;      // floats of one (hi or lo) weight block
    int bfloats;
    int soff[NL];     // record offsets (same record as the mma.sync kernels: h1..h8, e9)
    int soff_v, ss;
};

__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : __expf(z) - 1.f; }

// Float offset of element (row, col) inside an A tile / weight slab set: K-block kb = col/32 is a [rows][32] slab with
// 128-byte rows; inside a row the 16-byte chunk index is XORed with (row & 7)  (SWIZZLE_128B).
__device__ __forceinline__ int sw128_off(int row, int col, int slab_rows) {
    const int kb = col >> 5, c = (col >> 2) & 7, e = col & 3;
    return kb * slab_rows * 32 + row * 32 + ((c ^ (row & 7)) << 2) + e;
}

// Stage W_l (hi, lo) for all layers: B operand = W[n = out][k = in], K-major, zero padded to (n[l], 32*kb[l]).
__device__ void stage_weights_tc(const TcAeGeom& tg, const AeGeom& g, const AeParams& p, float* whi, float* wlo, float* bias,
                                 int tid, int nthreads) {
    for (int l = 0; l < NL; ++l) {
        const int IN = g.in[l], OUT = g.out[l], NP = tg.n[l], KP = 32 * tg.kb[l];
        for (int idx = tid; idx < NP * KP; idx += nthreads) {
            const int o = idx / KP, i = idx - o * KP;
            const float w = (o < OUT && i < IN) ? p.W[l][o * IN + i] : 0.f;
            float hi, lo;
            st_split_tf32(w, hi, lo);
            const int off = tg.woff[l] + sw128_off(o, i, NP);
            whi[off] = hi;
            wlo[off] = lo;
        }
        for (int o = tid; o < NP; o += nthreads) bias[tg.boff[l] + o] = (o < OUT) ? p.b[l][o] : 0.f;
    }
}

// Write 4 consecutive features (col4 % 4 == 0) of one row into the stream's A tile as (hi, lo).
__device__ __forceinline__ void write_a4(float* ahi, float* alo, int row, int col4, float4 v) {
    const int off = sw128_off(row, col4, TILE);
    float4 h, l;
    st_split_tf32(v.x, h.x, l.x); st_split_tf32(v.y, h.y, l.y); st_split_tf32(v.z, h.z, l.z); st_split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(ahi + off) = h;
    *reinterpret_cast<float4*>(alo + off) = l;
}

// Copy `ncols` (16, 32 or 64; from column col0) of the stream's tile, as hi + lo, to the rows' saved records: each lane
// moves one 16-byte chunk, a warp instruction covers whole 64/128-byte row segments (coalesced), 128 threads cooperate.
__device__ __forceinline__ void save_from_tile(const float* __restrict__ ahi, const float* __restrict__ alo, float* __restrict__ save,
                                               long R0, long BF, int ss, int soff, int col0, int ncols, int tid128) {
    const int cpr = ncols >> 2;                               // chunks per row: 4, 8 or 16
    const int lg = cpr == 4 ? 2 : (cpr == 8 ? 3 : 4);
    for (int idx = tid128; idx < TILE * cpr; idx += 128) {
        const int r = idx >> lg, c = idx & (cpr - 1);
        if (R0 + r >= BF) continue;
        const int off = sw128_off(r, col0 + 4 * c, TILE);
        const float4 h = *reinterpret_cast<const float4*>(ahi + off), l = *reinterpret_cast<const float4*>(alo + off);
        *reinterpret_cast<float4*>(save + (R0 + r) * ss + soff + 4 * c) = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
    }
}

// AE = 0: magnitude autoencoder ('sf').  AE = 1: phase autoencoder + residual + polar->rect.
template <int AE>
// 9 warps are allocated as 12 (register file granule = 4 warps): 168 registers per thread is the ceiling
__global__ void __launch_bounds__(THREADS, 1)
ae_fwd_tc_kernel(StDims d, AeGeom g, TcAeGeom tg, AeParams p, const float* __restrict__ spec, const float* __restrict__ knobs,
                 int B, float* __restrict__ mag_out, float* __restrict__ mag_hat, float* __restrict__ phs_hat,
                 float* __restrict__ ri, float* __restrict__ ri_lo, float* __restrict__ save, long long* __restrict__ timing) {
    extern __shared__ uint8_t smem_raw[];
    long long tclk = 0, treg[4] = {0, 0, 0, 0};
#define ST_T0() if (timing) tclk = clock64();
#define ST_T(i) if (timing) { const long long n_ = clock64(); treg[i] += n_ - tclk; tclk = n_; }
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* whi = smem;
    float* wlo = whi + tg.wfloats;
    float* a_base = wlo + tg.wfloats;                       // [stream][hi|lo][2 K-blocks][128][32]
    float* bias = a_base + NSTREAM * 2 * 2 * KB_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias + tg.bfloats);
    uint64_t* a_ready = bars;                               // [stream] epilogue -> MMA : next layer's A is in smem
    uint64_t* d_ready = bars + NSTREAM;                     // [stream] MMA -> epilogue : accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTREAM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stage_weights_tc(tg, g, p, whi, wlo, bias, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTREAM; ++s) { mbar_init(&a_ready[s], 1); mbar_init(&d_ready[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS_AE);
    fence_async_smem();                                     // staged weights -> async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long BF = (long)B * d.F;
    const long ntiles = (BF + TILE - 1) / TILE;
    // tile t of stream s, round r:  t = (r * gridDim.x + blockIdx.x) * NSTREAM + s
    const long rounds = (ntiles + (long)gridDim.x * NSTREAM - 1) / ((long)gridDim.x * NSTREAM);

    if (warp == 0) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            uint32_t ph[NSTREAM] = {0, 0};
            for (long r = 0; r < rounds; ++r) {
                for (int l = 0; l < NL; ++l) {
                    for (int s = 0; s < NSTREAM; ++s) {
                        const long tile = (r * gridDim.x + blockIdx.x) * NSTREAM + s;
                        if (tile >= ntiles) continue;
                        mbar_wait(&a_ready[s], ph[s]);
                        ph[s] ^= 1;
                        tc_fence_after();
                        const uint32_t idesc = idesc_tf32(TILE, tg.n[l], false, false);
                        const uint32_t a_hi = smem_u32(a_base + (s * 2 + 0) * 2 * KB_FLOATS);
                        const uint32_t a_lo = smem_u32(a_base + (s * 2 + 1) * 2 * KB_FLOATS);
                        const uint32_t w_hi = smem_u32(whi + tg.woff[l]), w_lo = smem_u32(wlo + tg.woff[l]);
                        const uint32_t tmem_d = tmem_base + 64 * s;
                        uint32_t accumulate = 0;
                        for (int kb = 0; kb < tg.kb[l]; ++kb) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t ao = kb * (KB_FLOATS * 4) + ks * 32, bo = kb * (tg.n[l] * 128) + ks * 32;
                                const uint64_t dah = desc_kmajor_sw128(a_hi + ao), dal = desc_kmajor_sw128(a_lo + ao);
                                const uint64_t dbh = desc_kmajor_sw128(w_hi + bo), dbl = desc_kmajor_sw128(w_lo + bo);
                                umma_tf32(tmem_d, dal, dbh, idesc, accumulate);
                                umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                                umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                                accumulate = 1u;
                            }
                        }
                        umma_commit(&d_ready[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue warpgroups =======================
        const int s = (warp - 1) >> 2;                      // stream
        const int q = warp & 3;                             // TMEM lane quadrant of this warp
        const int row = 32 * q + lane;                      // row inside the tile = TMEM lane
        float* ahi = a_base + (s * 2 + 0) * 2 * KB_FLOATS;
        float* alo = a_base + (s * 2 + 1) * 2 * KB_FLOATS;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + 64 * s;
        const int bar_id = 1 + s;                           // named barrier of this warpgroup
        const int tail0 = d.T - d.OT, rowstride = 2 * d.Fp;
        uint32_t ph = 0;
        for (long r = 0; r < rounds; ++r) {
            const long tile = (r * gridDim.x + blockIdx.x) * NSTREAM + s;
            if (tile >= ntiles) break;
            const long R = tile * TILE + row;
            const bool ok = R < BF;
            const int b = ok ? (int)(R / d.F) : 0, f = ok ? (int)(R - (long)b * d.F) : 0;
            float* rec = save ? save + (ok ? R : 0) * tg.ss : nullptr;
            ST_T0()
            // ---- input track -> K-block 0 of A (columns >= T are zero).  All loads are issued before any use.
            {
                const float* sp = spec + (long)b * d.Tp * rowstride + f;
                float* mo = (AE == 0 && mag_out && ok) ? mag_out + (long)b * d.T * d.F + f : nullptr;
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 16) {          // two batches of 16 frames: 32 loads in flight per thread
                    float re[16], im[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const bool in = ok && c0 + e < d.T;
                        re[e] = in ? __ldg(sp + (long)(c0 + e) * rowstride) : 0.f;
                        im[e] = in ? __ldg(sp + (long)(c0 + e) * rowstride + d.Fp) : 0.f;
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4 += 4) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int tt = c0 + c4 + e;
                            if (AE == 0) {
                                v[e] = sqrtf(re[c4 + e] * re[c4 + e] + im[c4 + e] * im[c4 + e]);       // nn_proc.py:309
                                if (mo && tt < d.T) mo[(long)tt * d.F] = v[e];
                            } else {
                                v[e] = (ok && tt < d.T) ? atan2f(im[c4 + e], re[c4 + e] + 1e-7f) : 0.f; // nn_proc.py:310
                            }
                        }
                        write_a4(ahi, alo, row, c0 + c4, make_float4(v[0], v[1], v[2], v[3]));
                    }
                }
            }
            fence_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (threadIdx.x == 32 + 128 * s) mbar_arrive(&a_ready[s]);
            if (save) {
                save_from_tile(ahi, alo, save, tile * TILE, BF, tg.ss, tg.soff_v, 0, 32, threadIdx.x - 32 - 128 * s);
                // the copy reads other threads' rows: nobody may start overwriting the tile before everyone is done
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            }
            ST_T(0)
            // ---- layers
#pragma unroll 1
            for (int l = 0; l < NL; ++l) {
                mbar_wait(&d_ready[s], ph);
                ph ^= 1;
                tc_fence_after();
                ST_T(1)
                const int n = tg.n[l];
                const float* bl = bias + tg.boff[l];
                if (l < NL - 1) {
                    // hidden layer: bias + ELU -> next layer's A (hi, lo) and the record
#pragma unroll 1
                    for (int c0 = 0; c0 < n; c0 += 16) {
                        uint32_t rr[16];
                        tmem_ld16(taddr + c0, rr);
                        tmem_wait_ld();
#pragma unroll
                        for (int c4 = 0; c4 < 16; c4 += 4) {
                            float4 h;
                            h.x = elu_f(__uint_as_float(rr[c4 + 0]) + bl[c0 + c4 + 0]);
                            h.y = elu_f(__uint_as_float(rr[c4 + 1]) + bl[c0 + c4 + 1]);
                            h.z = elu_f(__uint_as_float(rr[c4 + 2]) + bl[c0 + c4 + 2]);
                            h.w = elu_f(__uint_as_float(rr[c4 + 3]) + bl[c0 + c4 + 3]);
                            write_a4(ahi, alo, row, c0 + c4, h);
                        }
                    }
                    if (l == 3) {
                        // knob concat (torch.cat, nn_proc.py:95-96): columns 16..31 of fnn_addknobs' input
                        const float* kp = knobs + (long)b * d.K;
#pragma unroll
                        for (int c4 = 0; c4 < 16; c4 += 4) {
                            float4 kv;
                            kv.x = (ok && c4 + 0 < d.K) ? __ldg(kp + c4 + 0) : 0.f;
                            kv.y = (ok && c4 + 1 < d.K) ? __ldg(kp + c4 + 1) : 0.f;
                            kv.z = (ok && c4 + 2 < d.K) ? __ldg(kp + c4 + 2) : 0.f;
                            kv.w = (ok && c4 + 3 < d.K) ? __ldg(kp + c4 + 3) : 0.f;
                            write_a4(ahi, alo, row, 16 + c4, kv);
                        }
                    } else if (n < 32 * tg.kb[l + 1]) {
                        // the next layer reads a full 32-column K-block: clear the columns this layer did not write
                        for (int c4 = n; c4 < 32 * tg.kb[l + 1]; c4 += 4) write_a4(ahi, alo, row, c4, make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                    tc_fence_before();
                    fence_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (threadIdx.x == 32 + 128 * s) mbar_arrive(&a_ready[s]);
                    // the record copy reads the tile the MMAs are reading: it overlaps them for free
                    if (save) {
                        save_from_tile(ahi, alo, save, tile * TILE, BF, tg.ss, tg.soff[l], 0, l == 3 ? 32 : n, threadIdx.x - 32 - 128 * s);
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    }
                    ST_T(2)
                } else {
                    // fnn_dec + output-side math (thread = row: coalesced along the bin axis)
                    uint32_t rr[16];
                    tmem_ld16(taddr, rr);
                    tmem_wait_ld();
                    tc_fence_before();
                    if (ok) {
                        const float* sp = spec + ((long)b * d.Tp + tail0) * rowstride + f;
                        float re[16], im[16], mh[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const bool in = j < d.OT;
                            re[j] = in ? __ldg(sp + (long)j * rowstride) : 0.f;
                            im[j] = in ? __ldg(sp + (long)j * rowstride + d.Fp) : 0.f;
                            mh[j] = (AE == 1 && in) ? mag_hat[((long)b * d.OT + j) * d.F + f] : 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (j < d.OT) {
                                const float ev = elu_f(__uint_as_float(rr[j]) + bl[j]);
                                const long oo = ((long)b * d.OT + j) * d.F + f;
                                if (rec) rec[tg.soff[8] + j] = ev;
                                if (AE == 0) {
                                    mag_hat[oo] = ev * sqrtf(re[j] * re[j] + im[j] * im[j]);     // 'sf', nn_proc.py:115
                                } else {
                                    const float phv = ev + atan2f(im[j], re[j] + 1e-7f);         // nn_proc.py:322
                                    float sn, cs;
                                    sincosf(phv, &sn, &cs);
                                    phs_hat[oo] = phv;
                                    const long orr = ((long)b * d.OTp + j) * rowstride + f;
                                    st_split_tf32(mh[j] * cs, ri[orr], ri_lo[orr]);             // nn_proc.py:325-326
                                    st_split_tf32(mh[j] * sn, ri[orr + d.Fp], ri_lo[orr + d.Fp]);
                                }
                            }
                        }
                    }
                    ST_T(3)
                }
            }
        }
        if (timing && threadIdx.x == 32)
            for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(timing) + i, (unsigned long long)treg[i]);
    }
#undef ST_T0
#undef ST_T
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS_AE);
    }
}

TcAeGeom build_tc_geom(const AeGeom& g) {
    TcAeGeom tg;
    const int soff[NL] = {0, 64, 96, 112, 144, 160, 176, 208, 272};
    int off = 0, boff = 0;
    for (int l = 0; l < NL; ++l) {
        tg.n[l] = (g.out[l] + 15) / 16 * 16;
        const int in = (l == 4) ? 32 : g.in[l];             // layer 5: 16 features + 16 knob slots
        tg.kb[l] = (in + 31) / 32;
        tg.woff[l] = off;
        off += tg.n[l] * 32 * tg.kb[l];
        off = (off + 255) / 256 * 256;                      // keep every slab 1024-byte aligned
        tg.boff[l] = boff;
        boff += tg.n[l];
        tg.soff[l] = soff[l];
    }
    tg.wfloats = off;
    tg.bfloats = (boff + 3) / 4 * 4;
    tg.soff_v = 272 + 16;
    tg.ss = tg.soff_v + 32;
    return tg;
}

}  // namespace

// Same contract as st_launch_ae_forward_mma (same saved-record layout); covers T <= 32, OT <= 16, K <= 16.
bool st_launch_ae_forward_tc(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                             const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri, float* ri_lo,
                             float* save_m, float* save_p, long long* timing, int sm_count, cudaStream_t s) {
    if (d.T > 32 || d.OT > 16 || d.K > 16) return false;
    const TcAeGeom tg = build_tc_geom(g);
    const size_t smem = 1024 + sizeof(float) * (2 * (size_t)tg.wfloats + NSTREAM * 2 * 2 * (size_t)KB_FLOATS + tg.bfloats) + 64;
    if (smem > 227 * 1024) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(ae_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(ae_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return false;
        configured = true;
    }
    const long ntiles = ((long)B * d.F + TILE - 1) / TILE;
    const int grid = (int)std::min<long>((ntiles + NSTREAM - 1) / NSTREAM, sm_count);
    ae_fwd_tc_kernel<0><<<grid, THREADS, smem, s>>>(d, g, tg, pm, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, save_m, timing);
    ae_fwd_tc_kernel<1><<<grid, THREADS, smem, s>>>(d, g, tg, pp, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, save_p, timing ? timing + 4 : nullptr);
    return true;
}
