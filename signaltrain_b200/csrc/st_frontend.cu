// Front-end glue kernels: padding, weight packing/folding, overlap-add, DFT-gradient finalisation.
// All are HBM-streaming kernels: float4 coalesced accesses, grids sized from the element count.
#include "st_common.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// dst[r, :] = [0]*pad ++ scale*src[r, :len] ++ [0]*(stride-pad-len), written as an exact tf32 pair (hi, lo).
// Used for x/2 (nn_proc.py:307 + Conv1d padding=N, cls_fe_dft.py:28) and for 2*dL/dy_hat (adjoint of the [N:-N] trim,
// cls_fe_dft.py:113, and of the final *2, nn_proc.py:340).  The row stride is Tp*H (resp. OTp*H) so that frame (b, t)
// starts at (b*Tp + t)*H: one uniform-stride 2-D view over all frames of all windows.
__global__ void pad_split_kernel(const float* __restrict__ src, float* __restrict__ dst_hi, float* __restrict__ dst_lo,
                                 int rows, int len, int pad, int stride, float scale) {
    const int s4 = stride >> 2;
    const long total = (long)rows * s4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / s4);
        const int c = (int)(i - (long)r * s4) << 2;
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
        if (c >= pad && c < pad + len) {
            const float4 v = ld4(src + (long)r * len + (c - pad));
            st_split_tf32(v.x * scale, hi.x, lo.x);
            st_split_tf32(v.y * scale, hi.y, lo.y);
            st_split_tf32(v.z * scale, hi.z, lo.z);
            st_split_tf32(v.w * scale, hi.w, lo.w);
        }
        st4(dst_hi + (long)r * stride + c, hi);
        st4(dst_lo + (long)r * stride + c, lo);
    }
}

// wcat[2Fp][N]: rows [0,F) = Wr[0:F], rows [Fp,Fp+F) = Wi[0:F], padding rows zero.
// (only bins [:F] of the conv output are kept, cls_fe_dft.py:55-56)
__global__ void pack_analysis_kernel(StDims d, const float* __restrict__ Wr, const float* __restrict__ Wi,
                                     float* __restrict__ wcat, float* __restrict__ wcat_lo) {
    const int n4 = d.N >> 2;
    const long total = 2L * d.Fp * n4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int row = (int)(i / n4);
        const int c = (int)(i - (long)row * n4) << 2;
        const int half = row >= d.Fp;
        const int k = row - half * d.Fp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < d.F) v = ld4((half ? Wi : Wr) + (long)k * d.N + c);
        float4 hi, lo;
        st_split_tf32(v.x, hi.x, lo.x); st_split_tf32(v.y, hi.y, lo.y); st_split_tf32(v.z, hi.z, lo.z); st_split_tf32(v.w, hi.w, lo.w);
        st4(wcat + (long)row * d.N + c, hi);
        st4(wcat_lo + (long)row * d.N + c, lo);
    }
}

// sfold[2Fp][N]: Hermitian mirror (cls_fe_dft.py:109-110) folded into the synthesis matrices:
//   rows [0,F):      Sr[k] + (1<=k<=F-2 ? Sr[N-k] : 0)
//   rows [Fp,Fp+F):  Si[k] - (1<=k<=F-2 ? Si[N-k] : 0)
__global__ void fold_synthesis_kernel(StDims d, const float* __restrict__ Sr, const float* __restrict__ Si,
                                      float* __restrict__ sfold, float* __restrict__ sfold_lo) {
    const int n4 = d.N >> 2;
    const long total = 2L * d.Fp * n4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int row = (int)(i / n4);
        const int c = (int)(i - (long)row * n4) << 2;
        const int half = row >= d.Fp;
        const int k = row - half * d.Fp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < d.F) {
            const float* S = half ? Si : Sr;
            v = ld4(S + (long)k * d.N + c);
            if (k >= 1 && k <= d.F - 2) {
                const float4 m = ld4(S + (long)(d.N - k) * d.N + c);
                const float sg = half ? -1.f : 1.f;
                v.x += sg * m.x; v.y += sg * m.y; v.z += sg * m.z; v.w += sg * m.w;
            }
        }
        float4 hi, lo;
        st_split_tf32(v.x, hi.x, lo.x); st_split_tf32(v.y, hi.y, lo.y); st_split_tf32(v.z, hi.z, lo.z); st_split_tf32(v.w, hi.w, lo.w);
        st4(sfold + (long)row * d.N + c, hi);
        st4(sfold_lo + (long)row * d.N + c, lo);
    }
}

// Overlap-add of the per-frame synthesis output (ConvTranspose1d stride H, cls_fe_dft.py:112), trim
// [N:-N] (:113), add the input residual and undo the /2 (nn_proc.py:332,340).
//   frames_out (B*OTp, N)   x (B, C)   ->   y_hat (B, L)
__global__ void overlap_add_kernel(StDims d, const float* __restrict__ fo, const float* __restrict__ x, int B,
                                   float* __restrict__ y_hat, float* __restrict__ x_fwdsyn, float* __restrict__ y_half) {
    const int l4 = d.L >> 2;
    const long total = (long)B * l4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / l4);
        const int j = (int)(i - (long)b * l4) << 2;
        // frame t covers wave samples [tH, tH+N); output sample j sits at wave index j+N, so frame t
        // contributes iff j < tH <= j+N.  j, H, N are multiples of 4, so the four lanes of this float4
        // always see the same frame set.
        const int t_lo = j / d.H + 1;
        int t_hi = (j + d.N) / d.H;
        if (t_hi > d.OT - 1) t_hi = d.OT - 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = t_lo; t <= t_hi; ++t) {
            const float4 v = ld4(fo + ((long)b * d.OTp + t) * d.N + (j + d.N - t * d.H));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (!x) {      // stand-alone Synthesis.forward (cls_fe_dft.py:102-115): no residual, no *2
            st4(y_hat + (long)b * d.L + j, acc);
            continue;
        }
        float4 xr = ld4(x + (long)b * d.C + (d.C - d.L) + j);
        xr.x *= 0.5f; xr.y *= 0.5f; xr.z *= 0.5f; xr.w *= 0.5f;
        if (x_fwdsyn) st4(x_fwdsyn + (long)b * d.L + j, acc);
        float4 yh = make_float4(acc.x + xr.x, acc.y + xr.y, acc.z + xr.z, acc.w + xr.w);
        if (y_half) st4(y_half + (long)b * d.L + j, yh);
        yh.x *= 2.f; yh.y *= 2.f; yh.z *= 2.f; yh.w *= 2.f;
        st4(y_hat + (long)b * d.L + j, yh);
    }
}

// Sum the split-K partials of the two weight-gradient GEMMs and scatter into the reference's four
// (N,1,N) gradient tensors:
//   analysis: rows >= F receive no gradient (sliced off at cls_fe_dft.py:55-56) -> written as zero
//   synthesis: un-fold: dS[k] = G[k] (k<=F-1), dSr[N-k] = G_r[k], dSi[N-k] = -G_i[k] (1<=k<=F-2)
// which: bit 0 = the analysis pair, bit 1 = the synthesis pair (data parallel finalizes the synthesis gradients early so
// their allreduce overlaps the rest of the backward)
__global__ void finalize_dft_grads_kernel(StDims d, const float* __restrict__ pa, const float* __restrict__ ps,
                                          int sa, int ss, float* __restrict__ gWr, float* __restrict__ gWi,
                                          float* __restrict__ gSr, float* __restrict__ gSi, int which) {
    const int n4 = d.N >> 2;
    const long plane = 2L * d.Fp * d.N;
    const long total = (long)d.N * n4;           // one thread per (row k in [0,N), 4 columns)
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i / n4);
        const int c = (int)(i - (long)k * n4) << 2;
        const long o = (long)k * d.N + c;
        if (k >= d.F) {   // dead analysis rows; synthesis rows >= F are written by their mirror partner below
            if (which & 1) {
                st4(gWr + o, make_float4(0.f, 0.f, 0.f, 0.f));
                st4(gWi + o, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            continue;
        }
        float4 ar = make_float4(0.f, 0.f, 0.f, 0.f), ai = ar, sr = ar, si = ar;
        for (int s = 0; s < ((which & 1) ? sa : 0); ++s) {
            const float4 a = ld4(pa + s * plane + (long)k * d.N + c);
            const float4 b = ld4(pa + s * plane + (long)(d.Fp + k) * d.N + c);
            ar.x += a.x; ar.y += a.y; ar.z += a.z; ar.w += a.w;
            ai.x += b.x; ai.y += b.y; ai.z += b.z; ai.w += b.w;
        }
        for (int s = 0; s < ((which & 2) ? ss : 0); ++s) {
            const float4 a = ld4(ps + s * plane + (long)k * d.N + c);
            const float4 b = ld4(ps + s * plane + (long)(d.Fp + k) * d.N + c);
            sr.x += a.x; sr.y += a.y; sr.z += a.z; sr.w += a.w;
            si.x += b.x; si.y += b.y; si.z += b.z; si.w += b.w;
        }
        if (which & 1) {
            st4(gWr + o, ar);
            st4(gWi + o, ai);
        }
        if (!(which & 2)) continue;
        st4(gSr + o, sr);
        st4(gSi + o, si);
        if (k >= 1 && k <= d.F - 2) {
            const long om = (long)(d.N - k) * d.N + c;
            st4(gSr + om, sr);
            st4(gSi + om, make_float4(-si.x, -si.y, -si.z, -si.w));
        }
    }
}

// Analysis/Synthesis.initialize(): ortho DFT rows x window, evaluated in double then rounded once,
// as numpy does (cls_fe_dft.py:36-41, 87-92).  win = [hamming(N) | GLA(N,H)] in double (2N values).
__global__ void init_frontend_kernel(StDims d, const double* __restrict__ win, float* __restrict__ Wr,
                                     float* __restrict__ Wi, float* __restrict__ Sr, float* __restrict__ Si) {
    const long total = (long)d.N * d.N;
    const double inv = 1.0 / sqrt((double)d.N);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i / d.N);
        const int n = (int)(i - (long)k * d.N);
        const long kn = ((long)k * n) % d.N;
        double sv, cv;
        sincospi(2.0 * (double)kn / (double)d.N, &sv, &cv);
        const double re = cv * inv, im = -sv * inv;
        Wr[i] = (float)(re * win[n]);
        Wi[i] = (float)(im * win[n]);
        Sr[i] = (float)(re * win[d.N + n]);
        Si[i] = (float)(im * win[d.N + n]);
    }
}

// spec[(b,t), (re|im)] (row stride 2Fp, Tp rows per window) -> re, im as contiguous (B, T, F)   [Analysis.forward output]
__global__ void unpack_spec_kernel(StDims d, const float* __restrict__ spec, int B, float* __restrict__ re, float* __restrict__ im) {
    const long total = (long)B * d.T * d.F;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int f = (int)(i % d.F);
        const long bt = i / d.F;
        const int t = (int)(bt % d.T), b = (int)(bt / d.T);
        const long o = ((long)b * d.Tp + t) * (2 * d.Fp) + f;
        re[i] = spec[o];
        im[i] = spec[o + d.Fp];
    }
}
// re, im (B, OT, F) -> ri[(b,t), (re|im)] as a tf32 (hi, lo) pair   [Synthesis.forward input]
__global__ void pack_ri_kernel(StDims d, const float* __restrict__ re, const float* __restrict__ im, int B, float* __restrict__ ri,
                               float* __restrict__ ri_lo) {
    const long total = (long)B * d.OT * d.F;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int f = (int)(i % d.F);
        const long bt = i / d.F;
        const int t = (int)(bt % d.OT), b = (int)(bt / d.OT);
        const long o = ((long)b * d.OTp + t) * (2 * d.Fp) + f;
        st_split_tf32(re[i], ri[o], ri_lo[o]);
        st_split_tf32(im[i], ri[o + d.Fp], ri_lo[o + d.Fp]);
    }
}

inline int grid_for(long items, int threads) {
    long g = (items + threads - 1) / threads;
    const long cap = 148L * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

void st_launch_pad_split(const float* src, float* dst_hi, float* dst_lo, int rows, int len, int pad, int stride, float scale,
                         cudaStream_t s) {
    const long items = (long)rows * (stride >> 2);
    pad_split_kernel<<<grid_for(items, 256), 256, 0, s>>>(src, dst_hi, dst_lo, rows, len, pad, stride, scale);
}
const void* st_pad_split_kernel_fn() { return reinterpret_cast<const void*>(&pad_split_kernel); }
void st_launch_pack_analysis(const StDims& d, const float* Wr, const float* Wi, float* wcat, float* wcat_lo, cudaStream_t s) {
    pack_analysis_kernel<<<grid_for(2L * d.Fp * (d.N >> 2), 256), 256, 0, s>>>(d, Wr, Wi, wcat, wcat_lo);
}
void st_launch_fold_synthesis(const StDims& d, const float* Sr, const float* Si, float* sfold, float* sfold_lo, cudaStream_t s) {
    fold_synthesis_kernel<<<grid_for(2L * d.Fp * (d.N >> 2), 256), 256, 0, s>>>(d, Sr, Si, sfold, sfold_lo);
}
void st_launch_overlap_add(const StDims& d, const float* fo, const float* x, int B, float* y_hat, float* x_fwdsyn,
                           float* y_half, cudaStream_t s) {
    overlap_add_kernel<<<grid_for((long)B * (d.L >> 2), 256), 256, 0, s>>>(d, fo, x, B, y_hat, x_fwdsyn, y_half);
}
void st_launch_finalize_dft_grads(const StDims& d, const float* pa, const float* ps, int sa, int ss, float* gWr,
                                  float* gWi, float* gSr, float* gSi, int which, cudaStream_t s) {
    finalize_dft_grads_kernel<<<grid_for((long)d.N * (d.N >> 2), 256), 256, 0, s>>>(d, pa, ps, sa, ss, gWr, gWi, gSr, gSi, which);
}
void st_launch_init_frontend(const StDims& d, float* Wr, float* Wi, float* Sr, float* Si, float* scratch, cudaStream_t s) {
    init_frontend_kernel<<<grid_for((long)d.N * d.N, 256), 256, 0, s>>>(d, reinterpret_cast<const double*>(scratch), Wr, Wi,
                                                                        Sr, Si);
}

// ---- DCT / MDCT front-end variant (cls_fe_dct_bases.py) -------------------------------------------------------------
// Analysis epilogue: out[b, t, k] = tmp[(b Tp + t), k] + bias[k], t < nf  (Conv1d bias, cls_fe_dct_bases.py:116-117, then
// the transpose of :134)
__global__ void dct_bias_unpack_kernel(const float* __restrict__ tmp, const float* __restrict__ bias, int B, int nf, int Tp, int sz,
                                       float* __restrict__ out) {
    const long total = (long)B * nf * (sz >> 2);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k4 = (int)(i % (sz >> 2)) << 2;
        const long bt = i / (sz >> 2);
        const int t = (int)(bt % nf), b = (int)(bt / nf);
        float4 v = ld4(tmp + ((long)b * Tp + t) * sz + k4);
        const float4 bb = ld4(bias + k4);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        st4(out + bt * sz + k4, v);
    }
}
// Synthesis overlap-add + trim: ConvTranspose1d(stride hop) output position p = m + sz receives frame t's tap p - t hop
// (cls_fe_dct_bases.py:156-157, trimmed by sz on both sides :174-176).  wave (B, C).
__global__ void dct_overlap_add_kernel(const float* __restrict__ fo, int B, int nf, int sz, int wsz, int hop, int C, float* __restrict__ wave) {
    const long total = (long)B * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int m = (int)(i % C), b = (int)(i / C);
        const int p = m + sz;
        const int t_hi = min(p / hop, nf - 1), t_lo = max(0, (p - wsz) / hop + ((p - wsz) >= 0 ? 1 : 0));
        float acc = 0.f;
        for (int t = t_lo; t <= t_hi; ++t) acc += fo[((long)b * nf + t) * wsz + (p - t * hop)];
        wave[i] = acc;
    }
}
void st_launch_dct_bias_unpack(const float* tmp, const float* bias, int B, int nf, int Tp, int sz, float* out, cudaStream_t s) {
    dct_bias_unpack_kernel<<<grid_for((long)B * nf * (sz >> 2), 256), 256, 0, s>>>(tmp, bias, B, nf, Tp, sz, out);
}
void st_launch_dct_overlap_add(const float* fo, int B, int nf, int sz, int wsz, int hop, int C, float* wave, cudaStream_t s) {
    dct_overlap_add_kernel<<<grid_for((long)B * C, 256), 256, 0, s>>>(fo, B, nf, sz, wsz, hop, C, wave);
}

void st_launch_unpack_spec(const StDims& d, const float* spec, int B, float* re, float* im, cudaStream_t s) {
    unpack_spec_kernel<<<grid_for((long)B * d.T * d.F, 256), 256, 0, s>>>(d, spec, B, re, im);
}
void st_launch_pack_ri(const StDims& d, const float* re, const float* im, int B, float* ri, float* ri_lo, cudaStream_t s) {
    pack_ri_kernel<<<grid_for((long)B * d.OT * d.F, 256), 256, 0, s>>>(d, re, im, B, ri, ri_lo);
}
