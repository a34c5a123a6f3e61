// Autoencoders on the tensor cores (production path; st_ae.cu's SIMT kernels serve return_acts and act as cross-check).
//
// The nine Linear layers of AsymAutoEncoder (nn_proc.py:47-57,79-121) are dense contractions over rows = (batch, bin)
// pairs with tiny K/N (9..64).  Each WARP owns 32 rows (two m16 tiles) and walks the whole chain with warp-level
// mma.sync.m16n8k8 (TF32 operands, FP32 accumulate, 3xTF32 split for fp32 fidelity: a_lo*w_hi + a_hi*w_lo + a_hi*w_hi):
//   * activations live in a per-warp shared-memory tile [32 rows][feature] (row stride 68 -> conflict-free fragment
//     loads); a layer reads its A fragments from it, keeps ALL its outputs in accumulator registers, then overwrites
//     the tile in place -- no block-wide barrier anywhere in the chain, only __syncwarp;
//   * weights are staged once per CTA in their reference layout W[out][in] (row stride in+4 -> conflict-free B
//     fragments) and split into tf32 hi/lo on the fly;
//   * the k loop of every layer is a real loop (compact code: the first, fully unrolled version of this kernel was
//     I-cache bound, see profiles/), the n loop is unrolled over the accumulators.
// Forward: one launch per autoencoder (the phase launch consumes mag_hat and finishes polar->rect, nn_proc.py:322-326).
// When training, the forward also saves the eight hidden activations + the last ELU output per row (272+16 floats);
// the backward kernels read them back instead of recomputing the chain.
#include <algorithm>

#include "st_common.cuh"

namespace {

constexpr int RS = 68;                   // activation tile row stride (floats): 4*odd -> conflict-free A fragments
constexpr int ROWS_PER_WARP = 32;
constexpr int FWD_WARPS = 12;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
// PASSES = 3: exact two-term split (3xTF32).  PASSES = 1: reduced-precision mode (st_set_precision): hi only, one MMA per product.
template <int PASSES>
__device__ __forceinline__ void split_p(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = PASSES == 3 ? to_tf32(x - __uint_as_float(hi)) : 0u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// ELU(alpha=1).  exp through MUFU.EX2 (__expf): abs error <= ~3e-7 on outputs in (-1, 0], far inside the 1e-5 waveform
// budget ("precision"); expm1f would cost several times the instructions of the MMAs it sits between.
__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : __expf(z) - 1.f; }

// c[mt][n][.] = bias + sum_k act[row][k] * W[8n+g..][k]   for the warp's 32 rows; k loop rolled, n unrolled.
//   A fragment (m16 x k8): a0 (row g, k t), a1 (row g+8, k t), a2 (row g, k t+4), a3 (row g+8, k t+4)
//   B fragment (k8 x n8):  b0 (k t, n g), b1 (k t+4, n g)      with B[k][n] = W[n][k]
//   C fragment (m16 x n8): c0 (row g, n 2t), c1 (row g, n 2t+1), c2 (row g+8, n 2t), c3 (row g+8, n 2t+1)
template <int NT, int PASSES>
__device__ __forceinline__ void layer_mma(const float* __restrict__ act, int ksteps, const float* __restrict__ W, int ld,
                                          const float* __restrict__ bias, float (&c)[2][NT][4], int g, int t) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        float2 bb = make_float2(0.f, 0.f);
        if (bias) bb = *reinterpret_cast<const float2*>(bias + 8 * n + 2 * t);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { c[mt][n][0] = bb.x; c[mt][n][1] = bb.y; c[mt][n][2] = bb.x; c[mt][n][3] = bb.y; }
    }
#pragma unroll 1
    for (int j = 0; j < ksteps; ++j) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float* ap = act + (16 * mt + g) * RS + 8 * j + t;
            split_p<PASSES>(ap[0], ahi[mt][0], alo[mt][0]);
            split_p<PASSES>(ap[8 * RS], ahi[mt][1], alo[mt][1]);
            split_p<PASSES>(ap[4], ahi[mt][2], alo[mt][2]);
            split_p<PASSES>(ap[8 * RS + 4], ahi[mt][3], alo[mt][3]);
        }
        // pass-major over groups of n-tiles: consecutive MMAs hit different accumulators (the three passes of one
        // accumulator are 2*NG instructions apart), so the tensor pipe is not serialised on accumulator latency
        constexpr int NG = NT < 4 ? NT : 4;
#pragma unroll
        for (int n0 = 0; n0 < NT; n0 += NG) {
            uint32_t bh[NG][2], bl[NG][2];
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const float* wp = W + (8 * (n0 + q) + g) * ld + 8 * j + t;
                split_p<PASSES>(wp[0], bh[q][0], bl[q][0]);
                split_p<PASSES>(wp[4], bh[q][1], bl[q][1]);
            }
if (PASSES == 3) {
#pragma unroll
            for (int q = 0; q < NG; ++q)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][n0 + q], alo[mt], bh[q][0], bh[q][1]);
#pragma unroll
            for (int q = 0; q < NG; ++q)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][n0 + q], ahi[mt], bl[q][0], bl[q][1]);
            }
#pragma unroll
            for (int q = 0; q < NG; ++q)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][n0 + q], ahi[mt], bh[q][0], bh[q][1]);
        }
    }
}

// ELU the accumulators and overwrite the warp's activation tile (columns [0, 8*NT)).
template <int NT>
__device__ __forceinline__ void store_act(float* __restrict__ act, const float (&c)[2][NT][4], int g, int t) {
    __syncwarp();          // every lane has finished reading the previous contents
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float* p = act + (16 * mt + g) * RS + 8 * n + 2 * t;
            *reinterpret_cast<float2*>(p) = make_float2(elu_f(c[mt][n][0]), elu_f(c[mt][n][1]));
            *reinterpret_cast<float2*>(p + 8 * RS) = make_float2(elu_f(c[mt][n][2]), elu_f(c[mt][n][3]));
        }
    __syncwarp();
}

// Copy columns [0, width) of the warp's tile to the saved-activation record of each row (coalesced float4 rows).
template <int WIDTH>
__device__ __forceinline__ void save_tile(const float* __restrict__ act, float* __restrict__ save, long R0, long BF, int ss,
                                          int soff, int lane) {
    constexpr int W4 = WIDTH / 4;                 // float4 per row: 4, 8 or 16 -> 32 / W4 rows per warp instruction
    constexpr int RPI = 32 / W4;
    const int c4 = lane % W4, r0 = lane / W4;
    float* dst = save + (R0 + r0) * ss + soff + 4 * c4;
    const float* src = act + r0 * RS + 4 * c4;
#pragma unroll
    for (int i = 0; i < ROWS_PER_WARP / RPI; ++i)
        if (R0 + r0 + i * RPI < BF)
            *reinterpret_cast<float4*>(dst + (long)i * RPI * ss) = *reinterpret_cast<const float4*>(src + i * RPI * RS);
}

struct MmaGeom {
    int ks[ST_AE_LAYERS];      // k-steps of 8 (layer 5: 2 + 2 knob steps)
    int outp[ST_AE_LAYERS];    // outputs padded to a multiple of 8
    int ld[ST_AE_LAYERS];      // smem row stride of W_l: 8*ks + 4
    int off[ST_AE_LAYERS];     // float offset of W_l in the weight block
    int boff[ST_AE_LAYERS];    // float offset of bias_l in the bias block
    int soff[ST_AE_LAYERS];    // offset of layer l's output inside a saved-activation record
    int wfloats, bfloats;
    int soff_v;                // offset of the input track (mag or phase, 8*ks[0] columns) inside a record
    int ss;                    // saved-activation record length (floats)
};

// Stage one AE's weights (reference layout W[out][in], zero padded) and biases.
__device__ void stage_weights_mma(const MmaGeom& mg, const AeGeom& g, const AeParams& p, float* w, float* bias, int tid,
                                  int nthreads) {
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        const int IN = g.in[l], OUT = g.out[l], ld = mg.ld[l];
        const int total = mg.outp[l] * ld;
        for (int idx = tid; idx < total; idx += nthreads) {
            const int o = idx / ld, i = idx - o * ld;
            w[mg.off[l] + idx] = (o < OUT && i < IN) ? p.W[l][o * IN + i] : 0.f;
        }
        for (int o = tid; o < mg.outp[l]; o += nthreads) bias[mg.boff[l] + o] = (o < OUT) ? p.b[l][o] : 0.f;
    }
}

// AE = 0: magnitude autoencoder ('sf' skip-filter).  AE = 1: phase autoencoder + residual + polar->rect.
template <int NT9, int AE, int PASSES>
__global__ void __launch_bounds__(FWD_WARPS * 32, 1)
ae_fwd_mma_kernel(StDims d, AeGeom g, MmaGeom mg, AeParams p, const float* __restrict__ spec,
                  const float* __restrict__ knobs, int B, float* __restrict__ mag_out, float* __restrict__ mag_hat,
                  float* __restrict__ phs_hat, float* __restrict__ ri, float* __restrict__ ri_lo, float* __restrict__ save) {
    extern __shared__ __align__(16) float smem[];
    float* W = smem;
    float* bias = W + mg.wfloats;
    float* acts = bias + mg.bfloats;
    stage_weights_mma(mg, g, p, W, bias, threadIdx.x, blockDim.x);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, t = lane & 3;
    float* act = acts + warp * (ROWS_PER_WARP * RS);
    const long BF = (long)B * d.F;
    const long ntiles = (BF + ROWS_PER_WARP - 1) / ROWS_PER_WARP;
    const int tail0 = d.T - d.OT;
    const int rowstride = 2 * d.Fp;
    const int kcols = mg.ks[0] <= 4 ? 32 : 64;      // track columns held in the tile / record (zero beyond T)

    for (long tile = (long)blockIdx.x * FWD_WARPS + warp; tile < ntiles; tile += (long)gridDim.x * FWD_WARPS) {
        const long R0 = tile * ROWS_PER_WARP;
        // ---- input tracks: lane <-> row (coalesced along the bin axis), columns = time frames
        {
            const long R = R0 + lane;
            const bool ok = R < BF;
            const int b = ok ? (int)(R / d.F) : 0;
            const int f = ok ? (int)(R - (long)b * d.F) : 0;
            const float* sp = spec + (long)b * d.Tp * rowstride + f;
            float* mo = (AE == 0 && mag_out) ? mag_out + (long)b * d.T * d.F + f : nullptr;
            __syncwarp();
#pragma unroll 8
            for (int tt = 0; tt < kcols; ++tt) {
                float v = 0.f;
                if (ok && tt < d.T) {
                    const float re = __ldg(sp + (long)tt * rowstride), im = __ldg(sp + (long)tt * rowstride + d.Fp);
                    if (AE == 0) {
                        v = sqrtf(re * re + im * im);                                      // nn_proc.py:309
                        if (mo) mo[(long)tt * d.F] = v;
                    } else {
                        v = atan2f(im, re + 1e-7f);                                        // nn_proc.py:310
                    }
                }
                act[lane * RS + tt] = v;
            }
            __syncwarp();
            if (save) { if (kcols == 32) save_tile<32>(act, save, R0, BF, mg.ss, mg.soff_v, lane); else save_tile<64>(act, save, R0, BF, mg.ss, mg.soff_v, lane); }
        }
        // ---- fnn_enc .. fnn_enc4
        {
            float c[2][8][4];
            layer_mma<8, PASSES>(act, mg.ks[0], W + mg.off[0], mg.ld[0], bias + mg.boff[0], c, gq, t);
            store_act<8>(act, c, gq, t);
            if (save) save_tile<64>(act, save, R0, BF, mg.ss, mg.soff[0], lane);
        }
        {
            float c[2][4][4];
            layer_mma<4, PASSES>(act, 8, W + mg.off[1], mg.ld[1], bias + mg.boff[1], c, gq, t);
            store_act<4>(act, c, gq, t);
            if (save) save_tile<32>(act, save, R0, BF, mg.ss, mg.soff[1], lane);
        }
        {
            float c[2][2][4];
            layer_mma<2, PASSES>(act, 4, W + mg.off[2], mg.ld[2], bias + mg.boff[2], c, gq, t);
            store_act<2>(act, c, gq, t);
            if (save) save_tile<16>(act, save, R0, BF, mg.ss, mg.soff[2], lane);
        }
        {
            float c[2][2][4];
            layer_mma<2, PASSES>(act, 2, W + mg.off[3], mg.ld[3], bias + mg.boff[3], c, gq, t);
            store_act<2>(act, c, gq, t);
        }
        // ---- knob concat (torch.cat, nn_proc.py:95-96): columns 16..31 = knobs, zero padded
        {
            const long R = R0 + lane;
            const bool ok = R < BF;
            const float* kp = knobs + (ok ? R / d.F : 0) * d.K;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) act[lane * RS + 16 + kk] = (ok && kk < d.K) ? __ldg(kp + kk) : 0.f;
            __syncwarp();
            if (save) save_tile<32>(act, save, R0, BF, mg.ss, mg.soff[3], lane);     // h4 ++ knobs: fnn_addknobs' input
        }
        {
            float c[2][2][4];
            layer_mma<2, PASSES>(act, 4, W + mg.off[4], mg.ld[4], bias + mg.boff[4], c, gq, t);
            store_act<2>(act, c, gq, t);
            if (save) save_tile<16>(act, save, R0, BF, mg.ss, mg.soff[4], lane);
        }
        {
            float c[2][2][4];
            layer_mma<2, PASSES>(act, 2, W + mg.off[5], mg.ld[5], bias + mg.boff[5], c, gq, t);
            store_act<2>(act, c, gq, t);
            if (save) save_tile<16>(act, save, R0, BF, mg.ss, mg.soff[5], lane);
        }
        {
            float c[2][4][4];
            layer_mma<4, PASSES>(act, 2, W + mg.off[6], mg.ld[6], bias + mg.boff[6], c, gq, t);
            store_act<4>(act, c, gq, t);
            if (save) save_tile<32>(act, save, R0, BF, mg.ss, mg.soff[6], lane);
        }
        {
            float c[2][8][4];
            layer_mma<8, PASSES>(act, 4, W + mg.off[7], mg.ld[7], bias + mg.boff[7], c, gq, t);
            store_act<8>(act, c, gq, t);
            if (save) save_tile<64>(act, save, R0, BF, mg.ss, mg.soff[7], lane);
        }
        // ---- fnn_dec; ELU'd outputs go through the tile so the output-side math runs lane <-> row (coalesced, and one
        //      copy of the transcendental code instead of one per accumulator register)
        {
            float c9[2][NT9][4];
            layer_mma<NT9, PASSES>(act, 8, W + mg.off[8], mg.ld[8], bias + mg.boff[8], c9, gq, t);
            store_act<NT9>(act, c9, gq, t);
        }
        {
            const long R = R0 + lane;
            if (R < BF) {
                const int b = (int)(R / d.F), f = (int)(R - (long)b * d.F);
                const float* sp = spec + ((long)b * d.Tp + tail0) * rowstride + f;
#pragma unroll 1
                for (int j = 0; j < d.OT; ++j) {
                    const float ev = act[lane * RS + j];
                    const float re = __ldg(sp + (long)j * rowstride), im = __ldg(sp + (long)j * rowstride + d.Fp);
                    const long oo = ((long)b * d.OT + j) * d.F + f;
                    if (save) save[R * mg.ss + mg.soff[8] + j] = ev;
                    if (AE == 0) {
                        mag_hat[oo] = ev * sqrtf(re * re + im * im);                         // 'sf', nn_proc.py:115
                    } else {
                        const float ph = ev + atan2f(im, re + 1e-7f);                       // nn_proc.py:322
                        const float m = mag_hat[oo];
                        float sn, cs;
                        sincosf(ph, &sn, &cs);
                        phs_hat[oo] = ph;
                        const long orr = ((long)b * d.OTp + j) * rowstride + f;
                        st_split_tf32(m * cs, ri[orr], ri_lo[orr]);                         // nn_proc.py:325-326
                        st_split_tf32(m * sn, ri[orr + d.Fp], ri_lo[orr + d.Fp]);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = 8;
constexpr int CTA_ROWS = BWD_WARPS * ROWS_PER_WARP;     // 256 rows per CTA tile
constexpr int PS = 72;     // backward plane row stride: 8 (mod 32) -> the transposed (weight-gradient) fragment loads, which
                           // dominate the backward's shared-memory traffic, are conflict-free

__device__ __forceinline__ float elu_grad(float h) { return h > 0.f ? 1.f : h + 1.f; }   // dELU/dz through the output h

// Data gradient of one layer for the warp's own 32 rows:  c[row][i] = sum_o gz[row][o] * W[o][i]
//   B fragment (k8 x n8) with B[k = o][n = i] = W[o][i]:  b0 = W[(8j+t)*ld + 8n+g],  b1 = W[(8j+t+4)*ld + 8n+g]
template <int NT, int PASSES>
__device__ __forceinline__ void layer_mma_T(const float* __restrict__ gz, int ksteps, const float* __restrict__ W, int ld,
                                            float (&c)[2][NT][4], int g, int t) {
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) c[mt][n][0] = c[mt][n][1] = c[mt][n][2] = c[mt][n][3] = 0.f;
#pragma unroll 1
    for (int j = 0; j < ksteps; ++j) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float* ap = gz + (16 * mt + g) * PS + 8 * j + t;
            split_p<PASSES>(ap[0], ahi[mt][0], alo[mt][0]);
            split_p<PASSES>(ap[8 * PS], ahi[mt][1], alo[mt][1]);
            split_p<PASSES>(ap[4], ahi[mt][2], alo[mt][2]);
            split_p<PASSES>(ap[8 * PS + 4], ahi[mt][3], alo[mt][3]);
        }
        constexpr int NG = NT < 4 ? NT : 4;      // pass-major over groups of n-tiles (see layer_mma)
#pragma unroll
        for (int n0 = 0; n0 < NT; n0 += NG) {
            uint32_t bh[NG][2], bl[NG][2];
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const float* wp = W + (8 * j + t) * ld + 8 * (n0 + q) + g;
                split_p<PASSES>(wp[0], bh[q][0], bl[q][0]);
                split_p<PASSES>(wp[4 * ld], bh[q][1], bl[q][1]);
            }
if (PASSES == 3) {
#pragma unroll
            for (int q = 0; q < NG; ++q)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][n0 + q], alo[mt], bh[q][0], bh[q][1]);
#pragma unroll
            for (int q = 0; q < NG; ++q)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][n0 + q], ahi[mt], bl[q][0], bl[q][1]);
            }
#pragma unroll
            for (int q = 0; q < NG; ++q)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][n0 + q], ahi[mt], bh[q][0], bh[q][1]);
        }
    }
}

// gz_{l-1} = (data gradient) * ELU'(h_{l-1});  h_{l-1} (own rows) is read from the staged plane `hp`, the result
// overwrites it in place (the caller has synchronised the CTA: nobody else reads these rows any more).
template <int NT>
__device__ __forceinline__ void store_gz(float* __restrict__ hp, const float (&c)[2][NT][4], int g, int t) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float* p = hp + (16 * mt + g + 8 * h) * PS + 2 * t;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const float2 hh = *reinterpret_cast<const float2*>(p + 8 * n);
                *reinterpret_cast<float2*>(p + 8 * n) =
                    make_float2(c[mt][n][2 * h] * elu_grad(hh.x), c[mt][n][2 * h + 1] * elu_grad(hh.y));
            }
        }
}

// Asynchronously stage columns [hbase, hbase + 4*w4) of the CTA tile's saved records into a plane (cp.async, 16 B per
// request, rows past the end of the batch zero-filled).
__device__ __forceinline__ void stage_h_async(float* __restrict__ plane, const float* __restrict__ save, long R0c, long BF, int ss,
                                              int hbase, int lg_w4 /* log2(float4 per row): 2, 3 or 4 */, int tid) {
    const int w4 = 1 << lg_w4;
    for (int idx = tid; idx < CTA_ROWS * w4; idx += BWD_WARPS * 32) {
        const int r = idx >> lg_w4, c4 = idx & (w4 - 1);
        const bool ok = R0c + r < BF;
        const float* src = save + (ok ? (R0c + r) : 0) * ss + hbase + 4 * c4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(plane + r * PS + 4 * c4);
        const int bytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_h_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Weight gradient of one layer, this warp's (mo, ni) output tiles, reduction over the CTA tile's 256 rows:
//   dW[o][i] += sum_rows gz[row][o] * h[row][i]
//   A = gz^T from the gz plane (m = o, k = row):  a0 (m g, k t) = G[(8ks+t)*PS + 16mo+g], a1 (m g+8, k t),
//                                                 a2 (m g, k t+4), a3 (m g+8, k t+4)
//   B = h from the staged plane (k = row, n = i): b0 (k t, n g) = H[(8ks+t)*PS + 8ni+g], b1 (k t+4, n g)
// Pair p = warp + 8q -> (mo = p % MB, ni = p / MB); MB divides 8, so all of a warp's pairs share mo (one A fragment).
template <int NPW, int PASSES>
__device__ __forceinline__ void wgrad_mma(float (&acc)[NPW][4], const float* __restrict__ gzp, const float* __restrict__ hpl,
                                          int MB, int NB, int warp, int g, int t) {
    const int P = MB * NB;
    if (warp >= P) return;
    const int mo = warp % MB;
    // KG k-steps in flight with their own accumulators, MMAs issued pass-major: KG*NPW independent accumulators between the
    // three dependent passes of any one of them
    constexpr int KG = NPW == 1 ? 4 : 2;
    float tmp[KG][NPW][4];
#pragma unroll
    for (int k = 0; k < KG; ++k)
#pragma unroll
        for (int q = 0; q < NPW; ++q) tmp[k][q][0] = tmp[k][q][1] = tmp[k][q][2] = tmp[k][q][3] = 0.f;
    const float* pa = gzp + t * PS + 16 * mo + g;
    const float* pb = hpl + t * PS + g;
    int nib[NPW];
    bool on[NPW];
#pragma unroll
    for (int q = 0; q < NPW; ++q) {
        on[q] = warp + 8 * q < P;
        nib[q] = 8 * ((warp + 8 * q) / MB);
    }
#pragma unroll 1
    for (int ks0 = 0; ks0 < CTA_ROWS / 8; ks0 += KG) {
        uint32_t ahi[KG][4], alo[KG][4], bh[KG][NPW][2], bl[KG][NPW][2];
#pragma unroll
        for (int k = 0; k < KG; ++k) {
            const float* ap = pa + 8 * (ks0 + k) * PS;
            split_p<PASSES>(ap[0], ahi[k][0], alo[k][0]);
            split_p<PASSES>(ap[8], ahi[k][1], alo[k][1]);
            split_p<PASSES>(ap[4 * PS], ahi[k][2], alo[k][2]);
            split_p<PASSES>(ap[4 * PS + 8], ahi[k][3], alo[k][3]);
            const float* bp = pb + 8 * (ks0 + k) * PS;
#pragma unroll
            for (int q = 0; q < NPW; ++q) {
                split_p<PASSES>(on[q] ? bp[nib[q]] : 0.f, bh[k][q][0], bl[k][q][0]);
                split_p<PASSES>(on[q] ? bp[4 * PS + nib[q]] : 0.f, bh[k][q][1], bl[k][q][1]);
            }
        }
        if (PASSES == 3) {
#pragma unroll
        for (int k = 0; k < KG; ++k)
#pragma unroll
            for (int q = 0; q < NPW; ++q) mma_tf32(tmp[k][q], alo[k], bh[k][q][0], bh[k][q][1]);
#pragma unroll
        for (int k = 0; k < KG; ++k)
#pragma unroll
            for (int q = 0; q < NPW; ++q) mma_tf32(tmp[k][q], ahi[k], bl[k][q][0], bl[k][q][1]);
        }
#pragma unroll
        for (int k = 0; k < KG; ++k)
#pragma unroll
            for (int q = 0; q < NPW; ++q) mma_tf32(tmp[k][q], ahi[k], bh[k][q][0], bh[k][q][1]);
    }
#pragma unroll
    for (int q = 0; q < NPW; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < KG; ++k) sum += tmp[k][q][e];
            acc[q][e] += sum;
        }
}

template <int NPW>
__device__ __forceinline__ void wgrad_flush(const float (&acc)[NPW][4], float* __restrict__ dst, int MB, int NB, int OUT, int IN,
                                            int warp, int g, int t) {
#pragma unroll
    for (int q = 0; q < NPW; ++q) {
        const int pidx = warp + 8 * q;
        if (pidx >= MB * NB) continue;
        const int mo = pidx % MB, ni = pidx / MB;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = 16 * mo + g + 8 * (e >> 1), i = 8 * ni + 2 * t + (e & 1);
            if (o < OUT && i < IN) dst[o * IN + i] = acc[q][e];
        }
    }
}

// Bias gradient: column sums of the gz plane.  Each warp sums ITS 32 rows for all columns (lane <-> column: conflict-free)
// into bpart[warp][col]; after the CTA barrier that follows, bias_combine adds the 8 partials in fixed order.
__device__ __forceinline__ void bias_partial(float* __restrict__ bpart, const float* __restrict__ myplane, int outp, int warp,
                                             int lane) {
    for (int o = lane; o < outp; o += 32) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
        for (int r = 0; r < ROWS_PER_WARP; r += 2) {
            s0 += myplane[r * PS + o];
            s1 += myplane[(r + 1) * PS + o];
        }
        bpart[warp * 64 + o] = s0 + s1;
    }
}
__device__ __forceinline__ void bias_combine(float* __restrict__ db, const float* __restrict__ bpart, int outp, int tid) {
    if (tid < outp) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < BWD_WARPS; ++w) s += bpart[w * 64 + tid];
        db[tid] += s;
    }
}

// AE = 0: magnitude autoencoder, AE = 1: phase autoencoder.  NT1 = n-tiles of the input track (4: T <= 32, 8: T <= 64).
template <int NT1, int AE, int PASSES>
__global__ void __launch_bounds__(BWD_WARPS * 32, 1)
ae_bwd_mma_kernel(StDims d, AeGeom g, MmaGeom mg, AeParams p, const float* __restrict__ spec, int B,
                  const float* __restrict__ save, const float* __restrict__ mag_hat, const float* __restrict__ phs_hat,
                  const float* __restrict__ g_ri, const float* __restrict__ g_mag_hat, const float* __restrict__ g_mag,
                  float* __restrict__ tail_ws, float* __restrict__ g_spec, float* __restrict__ g_spec_lo,
                  float* __restrict__ partials, long long* __restrict__ timing) {
    extern __shared__ __align__(16) float smem[];
    // optional region timing (st_debug_ae_timing): cycles of warp 0 per region, summed over CTAs
    long long tclk = 0, treg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define ST_T0() if (timing) tclk = clock64();
#define ST_T(i) if (timing) { const long long n_ = clock64(); treg[i] += n_ - tclk; tclk = n_; }
    float* W = smem;
    float* bias = W + mg.wfloats;                      // staged but unused here (keeps one staging routine)
    float* plane0 = bias + mg.bfloats;
    float* plane1 = plane0 + CTA_ROWS * PS;
    float* dbias = plane1 + CTA_ROWS * PS;             // [9][64]
    float* bpart = dbias + ST_AE_LAYERS * 64;          // [8 warps][64] bias-gradient partials of the current layer
    stage_weights_mma(mg, g, p, W, bias, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < ST_AE_LAYERS * 64; i += blockDim.x) dbias[i] = 0.f;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, t = lane & 3;
    const long BF = (long)B * d.F;
    const long nct = (BF + CTA_ROWS - 1) / CTA_ROWS;
    const int tail0 = d.T - d.OT;
    const int rowstride = 2 * d.Fp;
    const int ss = mg.ss;
    const int ks1 = mg.ks[0];
    float* my0 = plane0 + warp * ROWS_PER_WARP * PS;
    float* my1 = plane1 + warp * ROWS_PER_WARP * PS;

    // persistent weight-gradient accumulators: [pairs of this warp][C fragment]
    float a1[4][4], a2[2][4], a3[1][4], a4[1][4], a5[1][4], a6[1][4], a7[1][4], a8[2][4], a9[1][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a1[q][e] = 0.f;
        a2[0][e] = a2[1][e] = a8[0][e] = a8[1][e] = 0.f;
        a3[0][e] = a4[0][e] = a5[0][e] = a6[0][e] = a7[0][e] = a9[0][e] = 0.f;
    }

    for (long ct = blockIdx.x; ct < nct; ct += gridDim.x) {
        const long R0c = ct * CTA_ROWS, R0 = R0c + warp * ROWS_PER_WARP;
        ST_T0()
        // ---- output side: gz9 into plane0 (own rows), skip/residual gradient to tail_ws.  lane <-> row.
        {
            const long R = R0 + lane;
            const bool ok = R < BF;
            const int b = ok ? (int)(R / d.F) : 0, f = ok ? (int)(R - (long)b * d.F) : 0;
            const float* rec = save + (ok ? R : 0) * ss;
            for (int j = d.OT; j < mg.outp[8]; ++j) my0[lane * PS + j] = 0.f;
#pragma unroll 3
            for (int j = 0; j < d.OT; ++j) {
                float gz = 0.f;
                if (ok) {
                    const float e9 = __ldg(rec + mg.soff[8] + j);
                    const long oo = ((long)b * d.OT + j) * d.F + f;
                    const long orr = ((long)b * d.OTp + j) * rowstride + f;
                    const float gre = __ldg(g_ri + orr), gim = __ldg(g_ri + orr + d.Fp);
                    float sn, cs;
                    sincosf(__ldg(phs_hat + oo), &sn, &cs);
                    if (AE == 0) {   // an = mag_hat (cos, sin);  mag_hat = ELU(d) * v_tail     (nn_proc.py:115, 325-326)
                        float gm = gre * cs + gim * sn;
                        if (g_mag_hat) gm += __ldg(g_mag_hat + oo);
                        gz = gm * __ldg(rec + mg.soff_v + tail0 + j) * elu_grad(e9);
                        tail_ws[oo] = gm * e9;
                    } else {         // phs_hat = ELU(d) + phs_tail                              (nn_proc.py:322)
                        const float gp = __ldg(mag_hat + oo) * (gim * cs - gre * sn);
                        gz = gp * elu_grad(e9);
                        tail_ws[oo] = gp;
                    }
                }
                my0[lane * PS + j] = gz;
            }
        }
        // One layer step:  gz_l lives in `cur` (all 256 rows).  (1) start staging h_{l-1} into `nxt` (cp.async);
        // (2) the data-gradient MMAs of the warp's own rows overlap that load; (3) wait + CTA barrier; (4) weight / bias
        // gradient over all rows (A: cur, B: nxt); (5) CTA barrier; (6) gz_{l-1} = c * ELU'(h_{l-1}) overwrites the
        // warp's own rows of `nxt`; (7) CTA barrier.  `cur` and `nxt` then swap roles.
#define ST_BWD_LAYER(L, NPW_, ACC, CUR, NXT, MYCUR, MYNXT, MB_, NB_, OUTP_, HBASE_, HW4_, NT_, KS_)                          \
        __syncthreads();                                                                                                    \
        ST_T(0)                                                                                                             \
        stage_h_async(NXT, save, R0c, BF, ss, HBASE_, HW4_, threadIdx.x);                                                   \
        {                                                                                                                   \
            float c[2][NT_][4];                                                                                             \
            ST_T(1)                                                                                                         \
            layer_mma_T<NT_, PASSES>(MYCUR, KS_, W + mg.off[L], mg.ld[L], c, gq, t);                                                \
            ST_T(2)                                                                                                         \
            bias_partial(bpart, MYCUR, OUTP_, warp, lane);                                                                  \
            ST_T(5)                                                                                                         \
            stage_h_wait();                                                                                                 \
            __syncthreads();                                                                                                \
            ST_T(3)                                                                                                         \
            wgrad_mma<NPW_, PASSES>(ACC, CUR, NXT, MB_, NB_, warp, gq, t);                                                          \
            bias_combine(dbias + L * 64, bpart, OUTP_, threadIdx.x);                                                        \
            ST_T(4)                                                                                                         \
            __syncthreads();                                                                                                \
            store_gz<NT_>(MYNXT, c, gq, t);                                                                                 \
            ST_T(6)                                                                                                         \
        }
        // layer 9 (fnn_dec):      OT(<=16) <- 64 ;  h8 is 64 wide
        ST_BWD_LAYER(8, 1, a9, plane0, plane1, my0, my1, 1, 8, mg.outp[8], mg.soff[7], 4, 8, mg.outp[8] / 8)
        // layer 8 (fnn_dec2):     64 <- 32
        ST_BWD_LAYER(7, 2, a8, plane1, plane0, my1, my0, 4, 4, 64, mg.soff[6], 3, 4, 8)
        // layer 7 (fnn_dec3):     32 <- 16
        ST_BWD_LAYER(6, 1, a7, plane0, plane1, my0, my1, 2, 2, 32, mg.soff[5], 2, 2, 4)
        // layer 6 (fnn_dec4):     16 <- 16
        ST_BWD_LAYER(5, 1, a6, plane1, plane0, my1, my0, 1, 2, 16, mg.soff[4], 2, 2, 2)
        // layer 5 (fnn_addknobs): 16 <- 16 + knobs (record slot of h4 is 32 wide: h4 ++ knobs); only h4 carries gradient
        ST_BWD_LAYER(4, 1, a5, plane0, plane1, my0, my1, 1, 4, 16, mg.soff[3], 3, 2, 2)
        // layer 4 (fnn_enc4):     16 <- 16
        ST_BWD_LAYER(3, 1, a4, plane1, plane0, my1, my0, 1, 2, 16, mg.soff[2], 2, 2, 2)
        // layer 3 (fnn_enc3):     16 <- 32
        ST_BWD_LAYER(2, 1, a3, plane0, plane1, my0, my1, 1, 4, 16, mg.soff[1], 3, 4, 2)
        // layer 2 (fnn_enc2):     32 <- 64
        ST_BWD_LAYER(1, 2, a2, plane1, plane0, my1, my0, 2, 8, 32, mg.soff[0], 4, 8, 4)
#undef ST_BWD_LAYER
        // ---- layer 1 (fnn_enc): 64 <- T.  Data gradient = dL/d(track), turned into dL/d(re, im).
        __syncthreads();
        ST_T(0)
        stage_h_async(plane1, save, R0c, BF, ss, mg.soff_v, ks1 <= 4 ? 3 : 4, threadIdx.x);   // the input tracks V (32|64 wide)
        {
            float c[2][NT1][4];
            ST_T(1)
            layer_mma_T<NT1, PASSES>(my0, 8, W + mg.off[0], mg.ld[0], c, gq, t);
            ST_T(2)
            bias_partial(bpart, my0, 64, warp, lane);
            ST_T(5)
            stage_h_wait();
            __syncthreads();
            ST_T(3)
            wgrad_mma<4, PASSES>(a1, plane0, plane1, 4, ks1, warp, gq, t);
            bias_combine(dbias + 0 * 64, bpart, 64, threadIdx.x);
            ST_T(4)
            __syncthreads();
            // dL/d(track) goes through the warp's own rows of plane0 so the output-side math runs lane <-> row
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n = 0; n < NT1; ++n) {
                    float* q = my0 + (16 * mt + gq) * PS + 8 * n + 2 * t;
                    *reinterpret_cast<float2*>(q) = make_float2(c[mt][n][0], c[mt][n][1]);
                    *reinterpret_cast<float2*>(q + 8 * PS) = make_float2(c[mt][n][2], c[mt][n][3]);
                }
            __syncwarp();
            const long R = R0 + lane;
            if (R < BF) {
                const int b = (int)(R / d.F), f = (int)(R - (long)b * d.F);
#pragma unroll 5
                for (int tt = 0; tt < d.T; ++tt) {
                    float gv = my0[lane * PS + tt];
                    if (tt >= tail0) gv += tail_ws[((long)b * d.OT + (tt - tail0)) * d.F + f];
                    const long o = ((long)b * d.Tp + tt) * rowstride + f;
                    const float re = __ldg(spec + o), im = __ldg(spec + o + d.Fp);
                    if (AE == 0) {          // mag = sqrt(re^2+im^2); subgradient 0 at 0 (torch.norm backward)
                        if (g_mag) gv += __ldg(g_mag + ((long)b * d.T + tt) * d.F + f);
                        const float m = my1[lane * PS + tt];
                        const float sc = m > 0.f ? gv / m : 0.f;
                        g_spec[o] = sc * re;
                        g_spec[o + d.Fp] = sc * im;
                    } else {                // phs = atan2(im, re + 1e-7); second pass: finish the sum, store (hi, lo)
                        const float u = re + 1e-7f;
                        const float den = u * u + im * im;
                        const float sc = den > 0.f ? gv / den : 0.f;
                        st_split_tf32(g_spec[o] - sc * im, g_spec[o], g_spec_lo[o]);
                        st_split_tf32(g_spec[o + d.Fp] + sc * u, g_spec[o + d.Fp], g_spec_lo[o + d.Fp]);
                    }
                }
            }
        }
        __syncthreads();
        ST_T(7)
    }
    if (timing && threadIdx.x == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(timing) + i, (unsigned long long)treg[i]);
#undef ST_T0
#undef ST_T
    // ---- flush this CTA's partial gradients (summed over CTAs in fixed order by ae_grad_reduce_kernel)
    float* dst = partials + ((long)blockIdx.x * 2 + AE) * g.flat_total;
    wgrad_flush<4>(a1, dst + g.flat_off[0], 4, ks1, 64, d.T, warp, gq, t);
    wgrad_flush<2>(a2, dst + g.flat_off[1], 2, 8, 32, 64, warp, gq, t);
    wgrad_flush<1>(a3, dst + g.flat_off[2], 1, 4, 16, 32, warp, gq, t);
    wgrad_flush<1>(a4, dst + g.flat_off[3], 1, 2, 16, 16, warp, gq, t);
    wgrad_flush<1>(a5, dst + g.flat_off[4], 1, 4, 16, 16 + d.K, warp, gq, t);
    wgrad_flush<1>(a6, dst + g.flat_off[5], 1, 2, 16, 16, warp, gq, t);
    wgrad_flush<1>(a7, dst + g.flat_off[6], 2, 2, 32, 16, warp, gq, t);
    wgrad_flush<2>(a8, dst + g.flat_off[7], 4, 4, 64, 32, warp, gq, t);
    wgrad_flush<1>(a9, dst + g.flat_off[8], 1, 8, d.OT, 64, warp, gq, t);
    __syncthreads();
    for (int i = threadIdx.x; i < ST_AE_LAYERS * 64; i += blockDim.x) {
        const int l = i >> 6, o = i & 63;
        if (o < g.out[l]) dst[g.flat_off[l] + g.out[l] * g.in[l] + o] = dbias[i];
    }
}


MmaGeom build_mma_geom(const AeGeom& g, int nt9) {
    MmaGeom mg;
    int off = 0, boff = 0;
    const int soff[ST_AE_LAYERS] = {0, 64, 96, 112, 144, 160, 176, 208, 272};   // h1..h8 (h4 slot 32 wide: ++knobs), e9
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        mg.ks[l] = (l == 4) ? 4 : (g.in[l] + 7) / 8;     // layer 5: 16 features + 16 knob slots
        mg.outp[l] = (l == 8) ? 8 * nt9 : (g.out[l] + 7) / 8 * 8;
        mg.ld[l] = 8 * mg.ks[l] + 4;
        mg.off[l] = off;
        off += mg.outp[l] * mg.ld[l];
        mg.boff[l] = boff;
        boff += mg.outp[l];
        mg.soff[l] = soff[l];
    }
    mg.wfloats = (off + 3) / 4 * 4;
    mg.bfloats = (boff + 3) / 4 * 4;
    mg.soff_v = 272 + 8 * nt9;
    mg.ss = mg.soff_v + (mg.ks[0] <= 4 ? 32 : 64);
    return mg;
}

template <int NT9, int PASSES>
void launch_fwd_pair(const StDims& d, const AeGeom& g, const MmaGeom& mg, const AeParams& pm, const AeParams& pp, const float* spec,
                     const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri, float* ri_lo, float* save_m,
                     float* save_p, int grid, size_t smem, cudaStream_t s) {
    static bool configured_on[ST_MAX_DEVICES] = {};              // function attributes are per device (context), not per process
    bool& configured = configured_on[st_current_device_slot()];
    if (!configured) {
        cudaFuncSetAttribute(ae_fwd_mma_kernel<NT9, 0, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(ae_fwd_mma_kernel<NT9, 1, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured = true;
    }
    ae_fwd_mma_kernel<NT9, 0, PASSES><<<grid, FWD_WARPS * 32, smem, s>>>(d, g, mg, pm, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, save_m);
    ae_fwd_mma_kernel<NT9, 1, PASSES><<<grid, FWD_WARPS * 32, smem, s>>>(d, g, mg, pp, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, save_p);
}

}  // namespace

int st_ae_mma_record_floats(const StDims& d) { return 272 + 8 * (d.OT <= 16 ? 2 : (d.OT <= 32 ? 4 : 8)) + (d.T <= 32 ? 32 : 64); }

// save_m / save_p: NULL (inference) or B*F records of st_ae_mma_record_floats() floats each (training).
// Returns false when the geometry is outside what the tensor-core kernels cover (caller uses the SIMT kernel).
bool st_launch_ae_forward_mma(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                              const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri, float* ri_lo,
                              float* save_m, float* save_p, int sm_count, cudaStream_t s, int passes) {
    if (d.T > 64 || d.OT > 64 || d.K > 16) return false;
    const int nt9 = d.OT <= 16 ? 2 : (d.OT <= 32 ? 4 : 8);
    const MmaGeom mg = build_mma_geom(g, nt9);
    const size_t smem = sizeof(float) * ((size_t)mg.wfloats + mg.bfloats + (size_t)FWD_WARPS * ROWS_PER_WARP * RS);
    if (smem > 227 * 1024) return false;
    const long tiles = ((long)B * d.F + ROWS_PER_WARP - 1) / ROWS_PER_WARP;
    const int grid = (int)std::min<long>((tiles + FWD_WARPS - 1) / FWD_WARPS, sm_count);
#define ST_FWD_PAIR(NT9_, P_) launch_fwd_pair<NT9_, P_>(d, g, mg, pm, pp, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, save_m, save_p, grid, smem, s)
    if (passes == 1) { if (nt9 == 2) ST_FWD_PAIR(2, 1); else if (nt9 == 4) ST_FWD_PAIR(4, 1); else ST_FWD_PAIR(8, 1); }
    else             { if (nt9 == 2) ST_FWD_PAIR(2, 3); else if (nt9 == 4) ST_FWD_PAIR(4, 3); else ST_FWD_PAIR(8, 3); }
#undef ST_FWD_PAIR
    return true;
}

namespace {
template <int NT1, int PASSES>
void launch_bwd_pair(const StDims& d, const AeGeom& g, const MmaGeom& mg, const AeParams& pm, const AeParams& pp, const float* spec,
                     int B, const float* save_m, const float* save_p, const float* mag_hat, const float* phs_hat, const float* g_ri,
                     const float* g_mag_hat, const float* g_mag, float* tail_ws, float* g_spec, float* g_spec_lo, float* partials,
                     long long* timing, int grid, size_t smem, cudaStream_t s) {
    static bool configured_on[ST_MAX_DEVICES] = {};              // function attributes are per device (context), not per process
    bool& configured = configured_on[st_current_device_slot()];
    if (!configured) {
        cudaFuncSetAttribute(ae_bwd_mma_kernel<NT1, 0, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(ae_bwd_mma_kernel<NT1, 1, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured = true;
    }
    ae_bwd_mma_kernel<NT1, 0, PASSES><<<grid, BWD_WARPS * 32, smem, s>>>(d, g, mg, pm, spec, B, save_m, mag_hat, phs_hat, g_ri, g_mag_hat,
                                                                 g_mag, tail_ws, g_spec, g_spec_lo, partials, timing);
    ae_bwd_mma_kernel<NT1, 1, PASSES><<<grid, BWD_WARPS * 32, smem, s>>>(d, g, mg, pp, spec, B, save_p, mag_hat, phs_hat, g_ri, g_mag_hat,
                                                                 g_mag, tail_ws, g_spec, g_spec_lo, partials, timing ? timing + 8 : nullptr);
}
}  // namespace

// Tensor-core backward of both autoencoders from the records saved by st_launch_ae_forward_mma.  Writes g_spec (hi, lo)
// and one partial-gradient vector per CTA into `partials` ([grid][2][flat_total]); returns the grid size (0 if the
// geometry is not covered and the caller must use the SIMT kernel).
int st_launch_ae_backward_mma(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec, int B,
                              const float* save_m, const float* save_p, const float* mag_hat, const float* phs_hat,
                              const float* g_ri, const float* g_mag_hat, const float* g_mag, float* tail_ws, float* g_spec,
                              float* g_spec_lo, float* partials, long long* timing, int sm_count, cudaStream_t s, int passes) {
    if (d.T > 64 || d.OT > 16 || d.K > 16) return 0;
    const MmaGeom mg = build_mma_geom(g, 2);
    const size_t smem = sizeof(float) * ((size_t)mg.wfloats + mg.bfloats + 2 * (size_t)CTA_ROWS * PS + ST_AE_LAYERS * 64 + BWD_WARPS * 64);
    if (smem > 227 * 1024) return 0;
    const long nct = ((long)B * d.F + CTA_ROWS - 1) / CTA_ROWS;
    const int grid = (int)std::min<long>(nct, sm_count);
#define ST_BWD_PAIR(NT1_, P_) launch_bwd_pair<NT1_, P_>(d, g, mg, pm, pp, spec, B, save_m, save_p, mag_hat, phs_hat, g_ri, g_mag_hat, g_mag, \
                                                         tail_ws, g_spec, g_spec_lo, partials, timing, grid, smem, s)
    if (passes == 1) { if (mg.ks[0] <= 4) ST_BWD_PAIR(4, 1); else ST_BWD_PAIR(8, 1); }
    else             { if (mg.ks[0] <= 4) ST_BWD_PAIR(4, 3); else ST_BWD_PAIR(8, 3); }
#undef ST_BWD_PAIR
    return grid;
}
