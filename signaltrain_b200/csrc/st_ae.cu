// Magnitude / phase autoencoders (AsymAutoEncoder, nn_proc.py:28-126) fused with their pro/epilogues:
//   forward : spectrum (re|im) -> mag, phase (nn_proc.py:309-310) -> 2 x nine Linear+ELU layers with the knob
//             concat (:79-121) -> skip-filter / phase residual (:115,:322) -> polar->rect (:325-326)
//   backward: recompute the chain per tile, back-propagate, accumulate all 36 weight/bias gradients in
//             registers across the CTA's tiles, emit dL/d(re|im).
//
// Work decomposition: a "row" is one (batch, bin) pair; its T-frame magnitude (phase) track is the input
// vector of the AE.  A CTA owns tiles of 64 consecutive rows.  Activations live in shared memory as
// [feature][row] planes (row stride 68 floats); every layer is a small in-smem GEMM:
//   half-warp hw (16 of them) owns OPW consecutive output features, lane li (0..15) owns 4 consecutive rows.
// Weights are staged once per CTA in shared memory, transposed to [in][out] so a half-warp's outputs are one
// broadcast vector load.  HBM traffic is the algorithmic minimum: the spectrum is read once per pass.
#include "st_common.cuh"

namespace {

constexpr int RS = ST_AE_RS;
constexpr int ROWS = ST_AE_ROWS;
constexpr int NT = ST_AE_THREADS;

struct ActsDev {
    float* p[ST_NUM_ACTS];
};

__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : expm1f(z); }
// d ELU / dz expressed through the OUTPUT h = ELU(z):  z>0 <=> h>0;  z<=0: exp(z) = h + 1
__device__ __forceinline__ float elu_grad(float h) { return h > 0.f ? 1.f : h + 1.f; }

// rows 4*li .. 4*li+3 of the current tile -> (batch, bin)
struct RowMap {
    int b[4], f[4];
    bool ok[4];
    long g0;
};
__device__ __forceinline__ RowMap make_rowmap(long row0, int li, long BF, int F) {
    RowMap m;
    m.g0 = row0 + 4 * li;
    int b = (int)(m.g0 / F), f = (int)(m.g0 - (long)b * F);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        m.b[r] = b; m.f[r] = f; m.ok[r] = (m.g0 + r) < BF;
        if (++f == F) { f = 0; ++b; }
    }
    return m;
}

__device__ __forceinline__ float f4get(const float4& v, int r) { return r == 0 ? v.x : (r == 1 ? v.y : (r == 2 ? v.z : v.w)); }

// One in-smem layer:  out[o][rows] = sum_i W[i][o] * Hin[i][rows]   for this half-warp's OPW outputs.
template <int OPW, class Epi>
__device__ __forceinline__ void ae_gemm(const float* __restrict__ Hin, int IN, const float* __restrict__ W, int ldw,
                                        int hw, int li, Epi&& epi) {
    float acc[OPW][4];
#pragma unroll
    for (int o = 0; o < OPW; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
    const float* hp = Hin + 4 * li;
    const float* wp = W + hw * OPW;
#pragma unroll 4
    for (int i = 0; i < IN; ++i) {
        const float4 h = *reinterpret_cast<const float4*>(hp + i * RS);
        float w[OPW];
        if (OPW == 4) {
            const float4 t = *reinterpret_cast<const float4*>(wp + i * ldw);
            w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
        } else if (OPW == 2) {
            const float2 t = *reinterpret_cast<const float2*>(wp + i * ldw);
            w[0] = t.x; w[1] = t.y;
        } else {
            w[0] = wp[i * ldw];
        }
#pragma unroll
        for (int o = 0; o < OPW; ++o) {
            acc[o][0] = fmaf(w[o], h.x, acc[o][0]);
            acc[o][1] = fmaf(w[o], h.y, acc[o][1]);
            acc[o][2] = fmaf(w[o], h.z, acc[o][2]);
            acc[o][3] = fmaf(w[o], h.w, acc[o][3]);
        }
    }
#pragma unroll
    for (int o = 0; o < OPW; ++o) epi(hw * OPW + o, make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]));
}

template <class Epi>
__device__ __forceinline__ void ae_gemm_dyn(int opw, const float* Hin, int IN, const float* W, int ldw, int hw, int li,
                                            Epi&& epi) {
    if (opw == 1) ae_gemm<1>(Hin, IN, W, ldw, hw, li, epi);
    else if (opw == 2) ae_gemm<2>(Hin, IN, W, ldw, hw, li, epi);
    else ae_gemm<4>(Hin, IN, W, ldw, hw, li, epi);
}

// Stage one AE's weights: Wt[l][i][o] (+ zero padding), biases; optionally the un-transposed W[l][o][i].
__device__ void stage_weights(const AeGeom& g, const AeParams& p, float* wt_block, float* w_block, int tid) {
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        const int IN = g.in[l], OUT = g.out[l], OP = g.outp[l], IP = g.inp[l];
        float* wt = wt_block + g.off_wt[l];
        for (int idx = tid; idx < IN * OP; idx += NT) {
            const int i = idx / OP, o = idx - i * OP;
            wt[idx] = (o < OUT) ? p.W[l][o * IN + i] : 0.f;
        }
        float* bb = wt_block + g.off_b[l];
        for (int o = tid; o < OP; o += NT) bb[o] = (o < OUT) ? p.b[l][o] : 0.f;
        if (w_block) {
            float* w = w_block + g.off_w[l];
            for (int idx = tid; idx < OUT * IP; idx += NT) {
                const int o = idx / IP, i = idx - o * IP;
                w[idx] = (i < IN) ? p.W[l][o * IN + i] : 0.f;
            }
        }
    }
}

// Load one tile's input tracks: V[t][r] = mag or phase of spec[b, t, f].  which: 0 = mag only, 1 = phase only,
// 2 = both (Vm and Vp).  Rows past the end of the batch are zero.
__device__ __forceinline__ void load_tracks(const StDims& d, const float* __restrict__ spec, long row0, long BF, int which,
                                            float* Vm, float* Vp, float* __restrict__ mag_out, const ActsDev* acts,
                                            int tid) {
    const int r = tid & (ROWS - 1), tq = tid >> 6;   // 256 threads: 64 rows x 4 frame phases
    const long gr = row0 + r;
    const bool ok = gr < BF;
    const int b = ok ? (int)(gr / d.F) : 0;
    const int f = ok ? (int)(gr - (long)b * d.F) : 0;
    for (int t = tq; t < d.T; t += NT / ROWS) {
        float mg = 0.f, ph = 0.f;
        if (ok) {
            const long o = ((long)b * d.Tp + t) * (2 * d.Fp) + f;
            const float re = spec[o], im = spec[o + d.Fp];
            if (which != 1) mg = sqrtf(re * re + im * im);                   // nn_proc.py:309
            if (which != 0) ph = atan2f(im, re + 1e-7f);                     // nn_proc.py:310
            const long oo = ((long)b * d.T + t) * d.F + f;
            if (mag_out) mag_out[oo] = mg;
            if (acts && acts->p[0]) { acts->p[0][oo] = re; acts->p[1][oo] = im; acts->p[2][oo] = mg; acts->p[3][oo] = ph; }
        }
        if (which != 1) Vm[t * RS + r] = mg;
        if (which != 0) Vp[t * RS + r] = ph;
    }
}

// Hidden-layer epilogue: bias + ELU -> next plane (+ optional activation dump, reference shape (B,F,width)).
struct HiddenEpi {
    float* Hout; const float* bias; int OUT; int li;
    float* act; int act_w; int act_off; long g0; long BF;
    __device__ __forceinline__ void operator()(int o, float4 z) const {
        if (o >= OUT) return;
        const float b = bias[o];
        const float4 h = make_float4(elu_f(z.x + b), elu_f(z.y + b), elu_f(z.z + b), elu_f(z.w + b));
        *reinterpret_cast<float4*>(Hout + o * RS + 4 * li) = h;
        if (act) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (g0 + r < BF) act[(g0 + r) * act_w + act_off + o] = f4get(h, r);
        }
    }
};

// Layers 1..8 of one AE for the current tile.  Planes: pl[0] = layer-1 output (64 rows) ... pl[7] = layer-8
// output.  In the forward kernel the planes alternate between two buffers; in the backward kernel each has
// its own storage (the backward pass needs them all).  The knob rows are appended to pl[3] (layer-4 output).
__device__ __forceinline__ void ae_hidden_chain(const StDims& d, const AeGeom& g, const float* wt, const float* V,
                                                float* const pl[8], const float* __restrict__ knobs, long row0, long BF,
                                                int hw, int li, int tid, float* const* act /*10 or null*/) {
    const long g0 = row0 + 4 * li;
    auto A = [&](int i) -> float* { return act ? act[i] : nullptr; };
    ae_gemm<4>(V, g.in[0], wt + g.off_wt[0], g.outp[0], hw, li,
               HiddenEpi{pl[0], wt + g.off_b[0], 64, li, A(0), 64, 0, g0, BF});
    __syncthreads();
    ae_gemm<2>(pl[0], 64, wt + g.off_wt[1], g.outp[1], hw, li, HiddenEpi{pl[1], wt + g.off_b[1], 32, li, A(1), 32, 0, g0, BF});
    __syncthreads();
    ae_gemm<1>(pl[1], 32, wt + g.off_wt[2], g.outp[2], hw, li, HiddenEpi{pl[2], wt + g.off_b[2], 16, li, A(2), 16, 0, g0, BF});
    __syncthreads();
    // layer 4 writes rows 0..15 of pl[3]; the knobs are rows 16..16+K-1 (torch.cat, nn_proc.py:95-96)
    ae_gemm<1>(pl[2], 16, wt + g.off_wt[3], g.outp[3], hw, li, HiddenEpi{pl[3], wt + g.off_b[3], 16, li, A(3), 16, 0, g0, BF});
    if (act && act[4]) {   // "catted" activation (B,F,16+K): the 16 features are written below after the sync
    }
    for (int idx = tid; idx < d.K * ROWS; idx += NT) {
        const int kk = idx / ROWS, r = idx - kk * ROWS;
        const long gr = row0 + r;
        float kv = 0.f;
        if (gr < BF) kv = knobs[(gr / d.F) * d.K + kk];
        pl[3][(16 + kk) * RS + r] = kv;
    }
    __syncthreads();
    if (act && act[4]) {
        const int w = 16 + d.K;
        for (int idx = tid; idx < w * ROWS; idx += NT) {
            const int c = idx / ROWS, r = idx - c * ROWS;
            if (row0 + r < BF) act[4][(row0 + r) * w + c] = pl[3][c * RS + r];
        }
    }
    ae_gemm<1>(pl[3], g.in[4], wt + g.off_wt[4], g.outp[4], hw, li, HiddenEpi{pl[4], wt + g.off_b[4], 16, li, A(5), 16, 0, g0, BF});
    __syncthreads();
    ae_gemm<1>(pl[4], 16, wt + g.off_wt[5], g.outp[5], hw, li, HiddenEpi{pl[5], wt + g.off_b[5], 16, li, A(6), 16, 0, g0, BF});
    __syncthreads();
    ae_gemm<2>(pl[5], 16, wt + g.off_wt[6], g.outp[6], hw, li, HiddenEpi{pl[6], wt + g.off_b[6], 32, li, A(7), 32, 0, g0, BF});
    __syncthreads();
    ae_gemm<4>(pl[6], 32, wt + g.off_wt[7], g.outp[7], hw, li, HiddenEpi{pl[7], wt + g.off_b[7], 64, li, A(8), 64, 0, g0, BF});
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward
// smem: [Wt mag | Wt phs | Vm T*RS | Vp T*RS | Ha 64*RS | Hb 64*RS | MH OT*RS]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
ae_forward_kernel(StDims d, AeGeom g, AeParams pm, AeParams pp, const float* __restrict__ spec,
                  const float* __restrict__ knobs, int B, float* __restrict__ mag, float* __restrict__ mag_hat,
                  float* __restrict__ phs_hat, float* __restrict__ ri, float* __restrict__ ri_lo, ActsDev acts) {
    extern __shared__ __align__(16) float smem[];
    float* wtm = smem;
    float* wtp = wtm + g.wt_floats;
    float* Vm = wtp + g.wt_floats;
    float* Vp = Vm + d.T * RS;
    float* Ha = Vp + d.T * RS;
    float* Hb = Ha + 64 * RS;
    float* MH = Hb + 64 * RS;
    const int tid = threadIdx.x, hw = tid >> 4, li = tid & 15;
    const long BF = (long)B * d.F;
    const long ntiles = (BF + ROWS - 1) / ROWS;
    const bool dump = acts.p[0] != nullptr;

    stage_weights(g, pm, wtm, nullptr, tid);
    stage_weights(g, pp, wtp, nullptr, tid);
    __syncthreads();

    float* const planes[8] = {Ha, Hb, Ha, Hb, Ha, Hb, Ha, Hb};
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long row0 = tile * ROWS;
        load_tracks(d, spec, row0, BF, 2, Vm, Vp, mag, dump ? &acts : nullptr, tid);
        __syncthreads();
        const RowMap rm = make_rowmap(row0, li, BF, d.F);
        const int tail0 = d.T - d.OT;

        // ---- magnitude AE, skip_connections='sf' (nn_proc.py:315, :114-115)
        ae_hidden_chain(d, g, wtm, Vm, planes, knobs, row0, BF, hw, li, tid, dump ? &acts.p[4] : nullptr);
        {
            const float* b9 = wtm + g.off_b[8];
            ae_gemm_dyn(g.opw[8], Hb, 64, wtm + g.off_wt[8], g.outp[8], hw, li, [&](int j, float4 z) {
                if (j >= d.OT) return;
                const float4 vt = *reinterpret_cast<const float4*>(Vm + (tail0 + j) * RS + 4 * li);
                const float bb = b9[j];
                const float4 o4 = make_float4(elu_f(z.x + bb) * vt.x, elu_f(z.y + bb) * vt.y, elu_f(z.z + bb) * vt.z,
                                              elu_f(z.w + bb) * vt.w);
                *reinterpret_cast<float4*>(MH + j * RS + 4 * li) = o4;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (!rm.ok[r]) continue;
                    const float v = f4get(o4, r);
                    const long oo = ((long)rm.b[r] * d.OT + j) * d.F + rm.f[r];
                    mag_hat[oo] = v;
                    if (dump) { acts.p[13][(rm.g0 + r) * d.OT + j] = v; acts.p[24][oo] = v; }
                }
            });
        }
        __syncthreads();
        // ---- phase AE, skip_connections='' (nn_proc.py:316, :117) + residual (:322) + polar->rect (:325-326)
        ae_hidden_chain(d, g, wtp, Vp, planes, knobs, row0, BF, hw, li, tid, dump ? &acts.p[14] : nullptr);
        {
            const float* b9 = wtp + g.off_b[8];
            ae_gemm_dyn(g.opw[8], Hb, 64, wtp + g.off_wt[8], g.outp[8], hw, li, [&](int j, float4 z) {
                if (j >= d.OT) return;
                const float4 pt = *reinterpret_cast<const float4*>(Vp + (tail0 + j) * RS + 4 * li);
                const float4 mh = *reinterpret_cast<const float4*>(MH + j * RS + 4 * li);
                const float bb = b9[j];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (!rm.ok[r]) continue;
                    const float e = elu_f(f4get(z, r) + bb);
                    const float ph = e + f4get(pt, r);
                    const float m = f4get(mh, r);
                    float sn, cs;
                    sincosf(ph, &sn, &cs);
                    const long oo = ((long)rm.b[r] * d.OT + j) * d.F + rm.f[r];
                    const long or_ = ((long)rm.b[r] * d.OTp + j) * (2 * d.Fp) + rm.f[r];
                    phs_hat[oo] = ph;
                    st_split_tf32(m * cs, ri[or_], ri_lo[or_]);
                    st_split_tf32(m * sn, ri[or_ + d.Fp], ri_lo[or_ + d.Fp]);
                    if (dump) {
                        acts.p[23][(rm.g0 + r) * d.OT + j] = e;
                        acts.p[25][oo] = ph; acts.p[26][oo] = m * cs; acts.p[27][oo] = m * sn;
                    }
                }
            });
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Weight-gradient tile: acc[ob][ip] += sum_rows G[hw*OB+ob][row] * Hprev[li+16*ip][row]
template <int OB, int IPL>
__device__ __forceinline__ void ae_wgrad(float (&acc)[OB][IPL], const float* __restrict__ G, const float* __restrict__ Hprev,
                                         int IN, int hw, int li) {
#pragma unroll 2
    for (int r4 = 0; r4 < ROWS / 4; ++r4) {
        float4 gz[OB], hh[IPL];
#pragma unroll
        for (int o = 0; o < OB; ++o) gz[o] = *reinterpret_cast<const float4*>(G + (hw * OB + o) * RS + 4 * r4);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
            const int i = li + 16 * j;
            hh[j] = (i < IN) ? *reinterpret_cast<const float4*>(Hprev + i * RS + 4 * r4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int o = 0; o < OB; ++o)
#pragma unroll
            for (int j = 0; j < IPL; ++j)
                acc[o][j] += gz[o].x * hh[j].x + gz[o].y * hh[j].y + gz[o].z * hh[j].z + gz[o].w * hh[j].w;
    }
}
__device__ __forceinline__ float ae_bgrad(const float* __restrict__ G, int o) {
    float s = 0.f;
#pragma unroll 4
    for (int r4 = 0; r4 < ROWS / 4; ++r4) {
        const float4 v = *reinterpret_cast<const float4*>(G + o * RS + 4 * r4);
        s += (v.x + v.y) + (v.z + v.w);
    }
    return s;
}
template <int OB, int IPL>
__device__ __forceinline__ void ae_wflush(const float (&acc)[OB][IPL], float* __restrict__ dst, int OUT, int IN, int hw, int li) {
#pragma unroll
    for (int o = 0; o < OB; ++o)
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
            const int oo = hw * OB + o, i = li + 16 * j;
            if (oo < OUT && i < IN) dst[oo * IN + i] = acc[o][j];
        }
}
template <int OB, int IPL>
__device__ __forceinline__ void ae_wzero(float (&acc)[OB][IPL]) {
#pragma unroll
    for (int o = 0; o < OB; ++o)
#pragma unroll
        for (int j = 0; j < IPL; ++j) acc[o][j] = 0.f;
}

// Data-gradient epilogue for hidden layers: gz_{l-1} = (W_l^T gz_l) * ELU'(h_{l-1})
struct BackEpi {
    float* Gout; const float* Hprev; int OUT; int li;
    __device__ __forceinline__ void operator()(int i, float4 gh) const {
        if (i >= OUT) return;
        const float4 h = *reinterpret_cast<const float4*>(Hprev + i * RS + 4 * li);
        *reinterpret_cast<float4*>(Gout + i * RS + 4 * li) =
            make_float4(gh.x * elu_grad(h.x), gh.y * elu_grad(h.y), gh.z * elu_grad(h.z), gh.w * elu_grad(h.w));
    }
};

// smem: [Wt | Wb | V T*RS | H1 64 | H2 32 | H3 16 | H4 16+K | H5 16 | H6 16 | H7 32 | H8 64 (each *RS) | G0 64*RS | G1 64*RS | TAIL OT*RS]
__global__ void __launch_bounds__(NT, 1)
ae_backward_kernel(StDims d, AeGeom g, AeParams pm, AeParams pp, const float* __restrict__ spec,
                   const float* __restrict__ knobs, int B, const float* __restrict__ mag_hat,
                   const float* __restrict__ phs_hat, const float* __restrict__ g_ri,
                   const float* __restrict__ g_mag_hat, const float* __restrict__ g_mag, float* __restrict__ g_spec,
                   float* __restrict__ g_spec_lo, float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    float* wt = smem;
    float* wb = wt + g.wt_floats;
    float* V = wb + g.w_floats;
    float* H[8];
    H[0] = V + d.T * RS;
    H[1] = H[0] + 64 * RS;
    H[2] = H[1] + 32 * RS;
    H[3] = H[2] + 16 * RS;
    H[4] = H[3] + (16 + d.K) * RS;
    H[5] = H[4] + 16 * RS;
    H[6] = H[5] + 16 * RS;
    H[7] = H[6] + 32 * RS;
    float* G0 = H[7] + 64 * RS;
    float* G1 = G0 + 64 * RS;
    float* TAIL = G1 + 64 * RS;
    const int tid = threadIdx.x, hw = tid >> 4, li = tid & 15;
    const long BF = (long)B * d.F;
    const long ntiles = (BF + ROWS - 1) / ROWS;
    const int tail0 = d.T - d.OT;

    // persistent gradient accumulators (see file header): [outputs per half-warp][inputs per lane]
    float a1[4][4], a2[2][4], a3[1][2], a4[1][1], a5[1][2], a6[1][1], a7[2][1], a8[4][2], a9[4][4];
    float bacc[ST_AE_LAYERS];

    for (int ae = 0; ae < 2; ++ae) {
        const AeParams& P = ae ? pp : pm;
        __syncthreads();
        stage_weights(g, P, wt, wb, tid);
        ae_wzero(a1); ae_wzero(a2); ae_wzero(a3); ae_wzero(a4); ae_wzero(a5); ae_wzero(a6); ae_wzero(a7); ae_wzero(a8); ae_wzero(a9);
#pragma unroll
        for (int l = 0; l < ST_AE_LAYERS; ++l) bacc[l] = 0.f;
        __syncthreads();

        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long row0 = tile * ROWS;
            load_tracks(d, spec, row0, BF, ae, V, V, nullptr, nullptr, tid);
            __syncthreads();
            const RowMap rm = make_rowmap(row0, li, BF, d.F);
            ae_hidden_chain(d, g, wt, V, H, knobs, row0, BF, hw, li, tid, nullptr);

            // ---- layer 9 forward + output-side gradient -> gz9 in G0, skip/residual gradient in TAIL
            {
                const float* b9 = wt + g.off_b[8];
                ae_gemm_dyn(g.opw[8], H[7], 64, wt + g.off_wt[8], g.outp[8], hw, li, [&](int j, float4 z) {
                    if (j >= d.OT) return;
                    const float4 vt = *reinterpret_cast<const float4*>(V + (tail0 + j) * RS + 4 * li);
                    const float bb = b9[j];
                    float gz[4], tl[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        gz[r] = 0.f; tl[r] = 0.f;
                        if (!rm.ok[r]) continue;
                        const float e = elu_f(f4get(z, r) + bb);
                        const long oo = ((long)rm.b[r] * d.OT + j) * d.F + rm.f[r];
                        const long or_ = ((long)rm.b[r] * d.OTp + j) * (2 * d.Fp) + rm.f[r];
                        const float gre = g_ri[or_], gim = g_ri[or_ + d.Fp];
                        float sn, cs;
                        sincosf(phs_hat[oo], &sn, &cs);
                        if (ae == 0) {   // an = mag_hat * (cos, sin);  mag_hat = ELU(d) * v_tail
                            float gm = gre * cs + gim * sn;
                            if (g_mag_hat) gm += g_mag_hat[oo];
                            gz[r] = gm * f4get(vt, r) * elu_grad(e);
                            tl[r] = gm * e;
                        } else {         // phs_hat = ELU(d) + phs_tail
                            const float gp = mag_hat[oo] * (gim * cs - gre * sn);
                            gz[r] = gp * elu_grad(e);
                            tl[r] = gp;
                        }
                    }
                    *reinterpret_cast<float4*>(G0 + j * RS + 4 * li) = make_float4(gz[0], gz[1], gz[2], gz[3]);
                    *reinterpret_cast<float4*>(TAIL + j * RS + 4 * li) = make_float4(tl[0], tl[1], tl[2], tl[3]);
                });
            }
            __syncthreads();
            // ---- reverse sweep.  G0/G1 ping-pong; weight gradients accumulate in registers.
            // layer 9: OUT=OT, IN=64
            ae_wgrad(a9, G0, H[7], 64, hw, li);
            if (tid < d.OT) bacc[8] += ae_bgrad(G0, tid);
            ae_gemm<4>(G0, d.OT, wb + g.off_w[8], g.inp[8], hw, li, BackEpi{G1, H[7], 64, li});
            __syncthreads();
            // layer 8: OUT=64, IN=32
            ae_wgrad(a8, G1, H[6], 32, hw, li);
            if (tid < 64) bacc[7] += ae_bgrad(G1, tid);
            ae_gemm<2>(G1, 64, wb + g.off_w[7], g.inp[7], hw, li, BackEpi{G0, H[6], 32, li});
            __syncthreads();
            // layer 7: OUT=32, IN=16
            ae_wgrad(a7, G0, H[5], 16, hw, li);
            if (tid < 32) bacc[6] += ae_bgrad(G0, tid);
            ae_gemm<1>(G0, 32, wb + g.off_w[6], g.inp[6], hw, li, BackEpi{G1, H[5], 16, li});
            __syncthreads();
            // layer 6: OUT=16, IN=16
            ae_wgrad(a6, G1, H[4], 16, hw, li);
            if (tid < 16) bacc[5] += ae_bgrad(G1, tid);
            ae_gemm<1>(G1, 16, wb + g.off_w[5], g.inp[5], hw, li, BackEpi{G0, H[4], 16, li});
            __syncthreads();
            // layer 5 (fnn_addknobs): OUT=16, IN=16+K; only the 16 non-knob inputs propagate
            ae_wgrad(a5, G0, H[3], 16 + d.K, hw, li);
            if (tid < 16) bacc[4] += ae_bgrad(G0, tid);
            ae_gemm<1>(G0, 16, wb + g.off_w[4], g.inp[4], hw, li, BackEpi{G1, H[3], 16, li});
            __syncthreads();
            // layer 4: OUT=16, IN=16
            ae_wgrad(a4, G1, H[2], 16, hw, li);
            if (tid < 16) bacc[3] += ae_bgrad(G1, tid);
            ae_gemm<1>(G1, 16, wb + g.off_w[3], g.inp[3], hw, li, BackEpi{G0, H[2], 16, li});
            __syncthreads();
            // layer 3: OUT=16, IN=32
            ae_wgrad(a3, G0, H[1], 32, hw, li);
            if (tid < 16) bacc[2] += ae_bgrad(G0, tid);
            ae_gemm<2>(G0, 16, wb + g.off_w[2], g.inp[2], hw, li, BackEpi{G1, H[1], 32, li});
            __syncthreads();
            // layer 2: OUT=32, IN=64
            ae_wgrad(a2, G1, H[0], 64, hw, li);
            if (tid < 32) bacc[1] += ae_bgrad(G1, tid);
            ae_gemm<4>(G1, 32, wb + g.off_w[1], g.inp[1], hw, li, BackEpi{G0, H[0], 64, li});
            __syncthreads();
            // layer 1: OUT=64, IN=T; the data gradient is dL/d(track) -> dL/d(re,im), accumulated into g_spec
            ae_wgrad(a1, G0, V, d.T, hw, li);
            if (tid < 64) bacc[0] += ae_bgrad(G0, tid);
            ae_gemm_dyn(g.opw_T, G0, 64, wb + g.off_w[0], g.inp[0], hw, li, [&](int t, float4 gv4) {
                if (t >= d.T) return;
                float4 tl = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t >= tail0) tl = *reinterpret_cast<const float4*>(TAIL + (t - tail0) * RS + 4 * li);
                const float4 vv = *reinterpret_cast<const float4*>(V + t * RS + 4 * li);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (!rm.ok[r]) continue;
                    float gv = f4get(gv4, r) + f4get(tl, r);
                    const long o = ((long)rm.b[r] * d.Tp + t) * (2 * d.Fp) + rm.f[r];
                    const float re = spec[o], im = spec[o + d.Fp];
                    if (ae == 0) {          // mag = sqrt(re^2+im^2); subgradient 0 at 0 (torch.norm backward)
                        if (g_mag) gv += g_mag[((long)rm.b[r] * d.T + t) * d.F + rm.f[r]];
                        const float m = f4get(vv, r);
                        const float s = m > 0.f ? gv / m : 0.f;
                        g_spec[o] = s * re;
                        g_spec[o + d.Fp] = s * im;
                    } else {                // phs = atan2(im, re + 1e-7)
                        const float u = re + 1e-7f;
                        const float den = u * u + im * im;
                        const float s = den > 0.f ? gv / den : 0.f;
                        // second pass over this element: finish the sum and store it as the wgrad GEMM's (hi, lo) operand
                        st_split_tf32(g_spec[o] - s * im, g_spec[o], g_spec_lo[o]);
                        st_split_tf32(g_spec[o + d.Fp] + s * u, g_spec[o + d.Fp], g_spec_lo[o + d.Fp]);
                    }
                }
            });
            __syncthreads();
        }
        // ---- flush this AE's partial gradients: partials[(cta*2+ae)*flat_total + ...]
        float* dst = partials + ((long)blockIdx.x * 2 + ae) * g.flat_total;
        ae_wflush(a1, dst + g.flat_off[0], 64, d.T, hw, li);
        ae_wflush(a2, dst + g.flat_off[1], 32, 64, hw, li);
        ae_wflush(a3, dst + g.flat_off[2], 16, 32, hw, li);
        ae_wflush(a4, dst + g.flat_off[3], 16, 16, hw, li);
        ae_wflush(a5, dst + g.flat_off[4], 16, 16 + d.K, hw, li);
        ae_wflush(a6, dst + g.flat_off[5], 16, 16, hw, li);
        ae_wflush(a7, dst + g.flat_off[6], 32, 16, hw, li);
        ae_wflush(a8, dst + g.flat_off[7], 64, 32, hw, li);
        ae_wflush(a9, dst + g.flat_off[8], d.OT, 64, hw, li);
#pragma unroll
        for (int l = 0; l < ST_AE_LAYERS; ++l)
            if (tid < g.out[l]) dst[g.flat_off[l] + g.out[l] * g.in[l] + tid] = bacc[l];
    }
}

// Sum the per-CTA partial gradient vectors (fixed order -> deterministic) and scatter to the 36 tensors.
__global__ void ae_grad_reduce_kernel(AeGeom g, const float* __restrict__ partials, int ncta, AeGrads gm, AeGrads gp) {
    const int ae = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.flat_total) return;
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += partials[((long)c * 2 + ae) * g.flat_total + e];
    int l = ST_AE_LAYERS - 1;
    while (l > 0 && e < g.flat_off[l]) --l;
    const int rel = e - g.flat_off[l];
    const int nw = g.out[l] * g.in[l];
    const AeGrads& G = ae ? gp : gm;
    if (rel < nw) G.W[l][rel] = s;
    else G.b[l][rel - nw] = s;
}

}  // namespace

size_t st_ae_fwd_smem(const StDims& d, const AeGeom& g) {
    return sizeof(float) * (2L * g.wt_floats + 2L * d.T * RS + 2L * 64 * RS + (long)d.OT * RS);
}
size_t st_ae_bwd_smem(const StDims& d, const AeGeom& g) {
    const long planes = 64 + 32 + 16 + (16 + d.K) + 16 + 16 + 32 + 64;
    return sizeof(float) * ((long)g.wt_floats + g.w_floats + (long)d.T * RS + planes * RS + 2L * 64 * RS + (long)d.OT * RS);
}

int st_ae_configure(st_handle* h, const StDims& d, const AeGeom& g) {
    const size_t sf = st_ae_fwd_smem(d, g), sb = st_ae_bwd_smem(d, g);
    if (sf > 227 * 1024 || sb > 227 * 1024)
        return st_fail_msg(h, "autoencoder tile does not fit shared memory (fwd %zu B, bwd %zu B > 227 KiB): T=%d OT=%d too large",
                           sf, sb, d.T, d.OT);
    // the attribute is per function (not per handle): always opt in to the full 227 KiB so handles with
    // different geometries can coexist in one process
    cudaError_t e = cudaFuncSetAttribute(ae_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return st_fail_cuda(h, e, "cudaFuncSetAttribute(ae_forward_kernel)", __FILE__, __LINE__);
    e = cudaFuncSetAttribute(ae_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return st_fail_cuda(h, e, "cudaFuncSetAttribute(ae_backward_kernel)", __FILE__, __LINE__);
    return 0;
}

void st_launch_ae_forward(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                          const float* knobs, int B, float* mag, float* mag_hat, float* phs_hat, float* ri, float* ri_lo,
                          float* const* acts_host, int grid, cudaStream_t s) {
    ActsDev acts;
    for (int i = 0; i < ST_NUM_ACTS; ++i) acts.p[i] = acts_host ? acts_host[i] : nullptr;
    ae_forward_kernel<<<grid, NT, st_ae_fwd_smem(d, g), s>>>(d, g, pm, pp, spec, knobs, B, mag, mag_hat, phs_hat, ri, ri_lo, acts);
}

void st_launch_ae_backward(const StDims& d, const AeGeom& g, const AeParams& pm, const AeParams& pp, const float* spec,
                           const float* knobs, int B, const float* mag_hat, const float* phs_hat, const float* g_ri,
                           const float* g_mag_hat, const float* g_mag, float* g_spec, float* g_spec_lo, float* partials,
                           int grid, cudaStream_t s) {
    ae_backward_kernel<<<grid, NT, st_ae_bwd_smem(d, g), s>>>(d, g, pm, pp, spec, knobs, B, mag_hat, phs_hat, g_ri,
                                                               g_mag_hat, g_mag, g_spec, g_spec_lo, partials);
}

void st_launch_ae_grad_reduce(const AeGeom& g, const float* partials, int ncta, const AeGrads& gm, const AeGrads& gp,
                              cudaStream_t s) {
    dim3 grid((g.flat_total + 255) / 256, 2);
    ae_grad_reduce_kernel<<<grid, 256, 0, s>>>(g, partials, ncta, gm, gp);
}
