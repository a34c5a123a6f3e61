// Loss (loss_functions.py:9-10,22-23,26-43), L1 gradient clip (nn_proc.py:299-302) and Adam
// (torch.optim.Adam as constructed at train.py:228).  All HBM-streaming: float4 loads, warp-shuffle
// block reductions, and a deterministic last-block final reduction (fixed summation order, no float atomics).
#include <algorithm>

#include "st_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of up to NV values; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* sh /* NV*32 floats */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) sh[i * 32 + w] = v[i];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = lane < nw ? sh[i * 32 + lane] : 0.f;
            v[i] = warp_sum(t);
        }
    }
    __syncthreads();
}

// Last block to arrive sums the per-block partials (in double, fixed order) and calls fin(sums).
template <int NV, class Fin>
__device__ __forceinline__ void grid_finish(float (&v)[NV], float* scratch, unsigned* counter, Fin&& fin) {
    __shared__ float sh[NV * 32];
    __shared__ bool last;
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * gridDim.x + blockIdx.x] = v[i];
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double s[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            s[i] = 0.0;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) s[i] += (double)scratch[i * gridDim.x + b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
        }
        if (threadIdx.x == 0) {
            fin(s);
            *counter = 0;   // re-arm for the next launch on this stream
        }
    }
}

// log(cosh(d)) without overflow and without cancellation near 0:
//   |d| < 1 : log1p(2 sinh^2(d/2))       else: |d| + log1p(exp(-2|d|)) - ln 2
__device__ __forceinline__ float logcosh_f(float d) {
    const float a = fabsf(d);
    if (a < 1.f) {
        const float sh = sinhf(0.5f * a);
        return log1pf(2.f * sh * sh);
    }
    return a + log1pf(__expf(-2.f * a)) - 0.69314718055994531f;
}

__global__ void __launch_bounds__(256)
loss_kernel(StDims d, const float* __restrict__ y_hat, const float* __restrict__ y, const float* __restrict__ mag_hat,
            const float* __restrict__ sbf, float l1_coef, int B, float* __restrict__ loss, float* __restrict__ g_y,
            float* __restrict__ g_m, float* scratch, unsigned* counter) {
    const long n1 = (long)B * d.L, n2 = (long)B * d.OT * d.F;
    const float inv1 = 1.f / (float)n1, inv2 = 1.f / (float)n2;
    const long stride = (long)gridDim.x * blockDim.x;
    float acc[2] = {0.f, 0.f};
    for (long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4; i < n1; i += stride * 4) {   // L % 4 == 0
        const float4 a = *reinterpret_cast<const float4*>(y + i);
        const float4 b = *reinterpret_cast<const float4*>(y_hat + i);
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
        acc[0] += (logcosh_f(dx) + logcosh_f(dy)) + (logcosh_f(dz) + logcosh_f(dw));
        if (g_y)   // d/dy_hat mean(log cosh(y - y_hat)) = -tanh(y - y_hat)/n
            *reinterpret_cast<float4*>(g_y + i) =
                make_float4(-tanhf(dx) * inv1, -tanhf(dy) * inv1, -tanhf(dz) * inv1, -tanhf(dw) * inv1);
    }
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += stride) {
        const float s = sbf ? sbf[i % d.F] : 1.f;
        const float v = mag_hat[i] * s;
        acc[1] += fabsf(v);
        if (g_m) g_m[i] = l1_coef * inv2 * s * (float)((v > 0.f) - (v < 0.f));
    }
    grid_finish<2>(acc, scratch, counter, [&](const double* s) {
        loss[0] = (float)(s[0] / (double)n1 + (double)l1_coef * s[1] / (double)n2);
    });
}

// Fused tail of the forward for the whole-step entry point (st_train_step): overlap-add + trim + input residual + *2
// (overlap_add_kernel, st_frontend.cu), log-cosh / weighted-L1 loss with both gradients (loss_kernel above) and the zero-padded
// 2*dL/dy_hat operand of the synthesis gradient GEMMs (pad_split_kernel), in one pass: y_hat and dL/dy_hat never touch HBM.
// Same arithmetic per element as the three kernels it replaces; only the summation order of the loss differs.
__global__ void __launch_bounds__(256)
ola_loss_kernel(StDims d, const float* __restrict__ fo, const float* __restrict__ x, const float* __restrict__ y,
                const float* __restrict__ mag_hat, const float* __restrict__ sbf, float l1_coef, int B, float* __restrict__ loss,
                float* __restrict__ gwave_hi, float* __restrict__ gwave_lo, float* __restrict__ g_m, float* scratch, unsigned* counter) {
    const long n1 = (long)B * d.L, n2 = (long)B * d.OT * d.F;
    const float inv1 = 1.f / (float)n1, inv2 = 1.f / (float)n2;
    const long stride = (long)gridDim.x * blockDim.x;
    const int l4 = d.L >> 2;
    float acc[2] = {0.f, 0.f};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)B * l4; i += stride) {
        const int b = (int)(i / l4);
        const int j = (int)(i - (long)b * l4) << 2;
        const int t_lo = j / d.H + 1;                       // frames covering output sample j (see overlap_add_kernel)
        int t_hi = (j + d.N) / d.H;
        if (t_hi > d.OT - 1) t_hi = d.OT - 1;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = t_lo; t <= t_hi; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(fo + ((long)b * d.OTp + t) * d.N + (j + d.N - t * d.H));
            w.x += v.x; w.y += v.y; w.z += v.z; w.w += v.w;
        }
        const float4 xr = *reinterpret_cast<const float4*>(x + (long)b * d.C + (d.C - d.L) + j);
        const float4 yh = make_float4((w.x + xr.x * 0.5f) * 2.f, (w.y + xr.y * 0.5f) * 2.f, (w.z + xr.z * 0.5f) * 2.f,
                                      (w.w + xr.w * 0.5f) * 2.f);                        // nn_proc.py:332,340
        const float4 a = *reinterpret_cast<const float4*>(y + (long)b * d.L + j);
        const float dx = a.x - yh.x, dy = a.y - yh.y, dz = a.z - yh.z, dw = a.w - yh.w;
        acc[0] += (logcosh_f(dx) + logcosh_f(dy)) + (logcosh_f(dz) + logcosh_f(dw));
        float4 hi, lo;                                       // 2 * dL/dy_hat = 2 * (-tanh(y - y_hat) / n), as a tf32 pair
        st_split_tf32(-tanhf(dx) * inv1 * 2.f, hi.x, lo.x);
        st_split_tf32(-tanhf(dy) * inv1 * 2.f, hi.y, lo.y);
        st_split_tf32(-tanhf(dz) * inv1 * 2.f, hi.z, lo.z);
        st_split_tf32(-tanhf(dw) * inv1 * 2.f, hi.w, lo.w);
        const long o = (long)b * d.Sg + d.N + j;             // the pad columns around it stay zero (zeroed at allocation)
        *reinterpret_cast<float4*>(gwave_hi + o) = hi;
        *reinterpret_cast<float4*>(gwave_lo + o) = lo;
    }
    // weighted-L1 term, four bins per thread (the flat (b, t, f) index runs over f fastest; F is odd, so rows are not 16-byte
    // aligned but the flat array is)
    const long n2v = n2 >> 2;
    const float c2 = l1_coef * inv2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2v; i += stride) {
        const float4 mh = *reinterpret_cast<const float4*>(mag_hat + 4 * i);
        int f = (int)((4 * i) % d.F);
        const float m[4] = {mh.x, mh.y, mh.z, mh.w};
        float g[4], a = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float s = sbf ? __ldg(sbf + f) : 1.f;
            const float v = m[e] * s;
            a += fabsf(v);
            g[e] = c2 * s * (float)((v > 0.f) - (v < 0.f));
            if (++f == d.F) f = 0;
        }
        acc[1] += a;
        *reinterpret_cast<float4*>(g_m + 4 * i) = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n2 & 3)) {
        const long i = 4 * n2v + threadIdx.x;
        const float s = sbf ? sbf[i % d.F] : 1.f;
        const float v = mag_hat[i] * s;
        acc[1] += fabsf(v);
        g_m[i] = c2 * s * (float)((v > 0.f) - (v < 0.f));
    }
    grid_finish<2>(acc, scratch, counter, [&](const double* s) {
        loss[0] = (float)(s[0] / (double)n1 + (double)l1_coef * s[1] / (double)n2);
    });
}

// Fused split-K sum + Hermitian un-fold of the four DFT gradients (finalize_dft_grads_kernel, st_frontend.cu) and the L1
// norm / clip coefficient over them (l1_norm4_kernel below) for st_train_step: the gradients are summed while they are
// written instead of being read back by a second kernel.
__global__ void __launch_bounds__(256)
finalize_norm_kernel(StDims d, const float* __restrict__ pa, const float* __restrict__ ps, int sa, int ss, float* __restrict__ gWr,
                     float* __restrict__ gWi, float* __restrict__ gSr, float* __restrict__ gSi, float grad_scale, float max_norm,
                     float* __restrict__ norm_out, float* __restrict__ coef_out, float* scratch, unsigned* counter) {
    const int n4 = d.N >> 2;
    const long plane = 2L * d.Fp * d.N;
    const long total = (long)d.N * n4;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[1] = {0.f};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i / n4);
        const int c = (int)(i - (long)k * n4) << 2;
        const long o = (long)k * d.N + c;
        if (k >= d.F) {                                      // dead analysis rows (cls_fe_dft.py:55-56)
            *reinterpret_cast<float4*>(gWr + o) = z4;
            *reinterpret_cast<float4*>(gWi + o) = z4;
            continue;
        }
        float4 ar = z4, ai = z4, sr = z4, si = z4;
        for (int s = 0; s < sa; ++s) {
            const float4 a = *reinterpret_cast<const float4*>(pa + s * plane + (long)k * d.N + c);
            const float4 b = *reinterpret_cast<const float4*>(pa + s * plane + (long)(d.Fp + k) * d.N + c);
            ar.x += a.x; ar.y += a.y; ar.z += a.z; ar.w += a.w;
            ai.x += b.x; ai.y += b.y; ai.z += b.z; ai.w += b.w;
        }
        for (int s = 0; s < ss; ++s) {
            const float4 a = *reinterpret_cast<const float4*>(ps + s * plane + (long)k * d.N + c);
            const float4 b = *reinterpret_cast<const float4*>(ps + s * plane + (long)(d.Fp + k) * d.N + c);
            sr.x += a.x; sr.y += a.y; sr.z += a.z; sr.w += a.w;
            si.x += b.x; si.y += b.y; si.z += b.z; si.w += b.w;
        }
        *reinterpret_cast<float4*>(gWr + o) = ar;
        *reinterpret_cast<float4*>(gWi + o) = ai;
        *reinterpret_cast<float4*>(gSr + o) = sr;
        *reinterpret_cast<float4*>(gSi + o) = si;
        const float na = (fabsf(ar.x) + fabsf(ar.y)) + (fabsf(ar.z) + fabsf(ar.w)) + (fabsf(ai.x) + fabsf(ai.y)) + (fabsf(ai.z) + fabsf(ai.w));
        const float ns = (fabsf(sr.x) + fabsf(sr.y)) + (fabsf(sr.z) + fabsf(sr.w)) + (fabsf(si.x) + fabsf(si.y)) + (fabsf(si.z) + fabsf(si.w));
        float mult = 1.f;
        if (k >= 1 && k <= d.F - 2) {                        // mirrored synthesis rows (cls_fe_dft.py:109-110)
            const long om = (long)(d.N - k) * d.N + c;
            *reinterpret_cast<float4*>(gSr + om) = sr;
            *reinterpret_cast<float4*>(gSi + om) = make_float4(-si.x, -si.y, -si.z, -si.w);
            mult = 2.f;
        }
        acc[0] += na + mult * ns;
    }
    grid_finish<1>(acc, scratch, counter, [&](const double* s) {
        const double total_norm = s[0] * (double)grad_scale;
        if (norm_out) norm_out[0] = (float)total_norm;
        float cf = 1.f;
        if (max_norm > 0.f) {
            const double cc = (double)max_norm / (total_norm + 1e-6);
            cf = cc < 1.0 ? (float)cc : 1.f;
        }
        coef_out[0] = cf;
    });
}

__global__ void __launch_bounds__(256)
mae_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, float* __restrict__ out, float* scratch,
           unsigned* counter) {
    float acc[1] = {0.f};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        acc[0] += fabsf(a[i] - b[i]);
    grid_finish<1>(acc, scratch, counter, [&](const double* s) { out[0] = (float)(s[0] / (double)n); });
}

struct Ptr4 {
    const float* p[4];
};
struct MPtr4 {
    float* p[4];
};

// total = grad_scale * sum_{4 tensors} |g|   (torch clip_grad_norm_ norm_type=1: the L1 norm of the per-tensor
// L1 norms);  coef = min(1, max_norm / (total + 1e-6))
__global__ void __launch_bounds__(256)
l1_norm4_kernel(Ptr4 g, long n_each, float grad_scale, float max_norm, float* __restrict__ norm_out,
                float* __restrict__ coef_out, float* scratch, unsigned* counter) {
    float acc[1] = {0.f};
    const long n4 = n_each >> 2;
    for (int t = 0; t < 4; ++t) {
        const float4* p = reinterpret_cast<const float4*>(g.p[t]);
        float s0 = 0.f, s1 = 0.f;
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
            const float4 v = p[i];
            s0 += fabsf(v.x) + fabsf(v.y);
            s1 += fabsf(v.z) + fabsf(v.w);
        }
        acc[0] += s0 + s1;
    }
    grid_finish<1>(acc, scratch, counter, [&](const double* s) {
        const double total = s[0] * (double)grad_scale;
        if (norm_out) norm_out[0] = (float)total;
        float c = 1.f;
        if (max_norm > 0.f) {
            const double cc = (double)max_norm / (total + 1e-6);
            c = cc < 1.0 ? (float)cc : 1.f;
        }
        coef_out[0] = c;
    });
}

__global__ void __launch_bounds__(256) scale4_kernel(MPtr4 g, long n_each, const float* __restrict__ coef) {
    const float c = coef[0];
    const long n4 = n_each >> 2;
    float4* p = reinterpret_cast<float4*>(g.p[blockIdx.y]);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = p[i];
        v.x *= c; v.y *= c; v.z *= c; v.w *= c;
        p[i] = v;
    }
}

// torch.optim.Adam single step (amsgrad off, weight_decay 0, maximize off):
//   m = lerp(m, g, 1-b1);  v = b2*v + (1-b2) g^2;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& sc) {
    m = m + (g - m) * (1.f - sc.beta1);
    v = v * sc.beta2 + (1.f - sc.beta2) * g * g;
    const float denom = sqrtf(v) * sc.inv_sqrt_bc2 + sc.eps;
    p = p - sc.lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_kernel(const __grid_constant__ AdamTensors t, const int2* __restrict__ chunk_map, AdamScalars sc,
            const float* __restrict__ clip_coef) {
    const int2 cm = chunk_map[blockIdx.x];
    const int ti = cm.x;
    const long base = (long)cm.y * ST_ADAM_CHUNK;
    const long n = t.n[ti];
    const long end = base + ST_ADAM_CHUNK < n ? base + ST_ADAM_CHUNK : n;
    float* __restrict__ p = t.p[ti];
    const float* __restrict__ g = t.g[ti];
    float* __restrict__ m = t.m[ti];
    float* __restrict__ v = t.v[ti];
    float gs = sc.grad_scale;
    if (ti < 4 && clip_coef) gs *= clip_coef[0];     // the clip covers the four DFT tensors only
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15u) == 0;
    long i0 = base;
    if (aligned) {
        const long vend = base + ((end - base) & ~3L);
        for (long i = base + threadIdx.x * 4L; i < vend; i += 256 * 4) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            adam_one(pp.x, gg.x * gs, mm.x, vv.x, sc);
            adam_one(pp.y, gg.y * gs, mm.y, vv.y, sc);
            adam_one(pp.z, gg.z * gs, mm.z, vv.z, sc);
            adam_one(pp.w, gg.w * gs, mm.w, vv.w, sc);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        }
        i0 = vend;
    }
    for (long i = i0 + threadIdx.x; i < end; i += 256) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_one(pp, g[i] * gs, mm, vv, sc);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

inline int grid_for(long items, int threads, int cap) {
    long g = (items + threads - 1) / threads;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// scratch must hold >= 2 * 1184 floats; counter is a zero-initialised device word owned by the handle.
void st_launch_loss(const StDims& d, const float* y_hat, const float* y, const float* mag_hat, const float* sbf,
                    float l1_coef, int B, float* loss, float* g_y_hat, float* g_mag_hat, float* scratch, unsigned* counter,
                    cudaStream_t s) {
    const long n = (long)B * d.OT * d.F;
    loss_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, s>>>(d, y_hat, y, mag_hat, sbf, l1_coef, B, loss, g_y_hat, g_mag_hat,
                                                            scratch, counter);
}
void st_launch_ola_loss(const StDims& d, const float* fo, const float* x, const float* y, const float* mag_hat, const float* sbf,
                        float l1_coef, int B, float* loss, float* gwave_hi, float* gwave_lo, float* g_mag_hat, float* scratch,
                        unsigned* counter, cudaStream_t s) {
    const long n = std::max((long)B * d.OT * d.F / 4, (long)B * (d.L / 4));      // float4 items of the larger loop
    ola_loss_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, s>>>(d, fo, x, y, mag_hat, sbf, l1_coef, B, loss, gwave_hi, gwave_lo,
                                                                g_mag_hat, scratch, counter);
}
void st_launch_finalize_norm(const StDims& d, const float* pa, const float* ps, int sa, int ss, float* gWr, float* gWi, float* gSr,
                             float* gSi, float grad_scale, float max_norm, float* norm_out, float* coef_out, float* scratch,
                             unsigned* counter, cudaStream_t s) {
    finalize_norm_kernel<<<148 * 4, 256, 0, s>>>(d, pa, ps, sa, ss, gWr, gWi, gSr, gSi, grad_scale, max_norm, norm_out, coef_out,
                                                 scratch, counter);
}
void st_launch_mae(const float* a, const float* b, long n, float* out, float* scratch, unsigned* counter, cudaStream_t s) {
    mae_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, s>>>(a, b, n, out, scratch, counter);
}
void st_launch_l1_norm4(const float* const g[4], long n_each, long, long, float grad_scale, float max_norm, float* norm_out,
                        float* coef_out, float* scratch, unsigned* counter, cudaStream_t s) {
    Ptr4 p;
    for (int i = 0; i < 4; ++i) p.p[i] = g[i];
    l1_norm4_kernel<<<148 * 4, 256, 0, s>>>(p, n_each, grad_scale, max_norm, norm_out, coef_out, scratch, counter);
}
void st_launch_scale4(float* const g[4], long n_each, const float* coef, cudaStream_t s) {
    MPtr4 p;
    for (int i = 0; i < 4; ++i) p.p[i] = g[i];
    dim3 grid(148 * 2, 4);
    scale4_kernel<<<grid, 256, 0, s>>>(p, n_each, coef);
}
// Data-parallel exchange payload: only what carries information travels.  Packed layout:
//   [Wr rows 0..F-1 | Wi rows 0..F-1 | Sr rows 0..F-1 | Si rows 0..F-1 | the 36 autoencoder tensors, 4-float aligned]
// Analysis rows >= F never receive gradient (cls_fe_dft.py:55-56) and the synthesis gradients are Hermitian
// (g_r[N-k] = g_r[k], g_i[N-k] = -g_i[k], cls_fe_dft.py:109-110), so unpacking restores all four full tensors exactly.
__global__ void __launch_bounds__(256)
pack_grads_kernel(GradPack gp, int dir /*0: pack, 1: unpack*/) {
    const long live4 = gp.live / 4;                          // float4 items of one live block (F * N)
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < 4 * live4; i += stride) {
        const int t = (int)(i / live4);
        const long e = i - t * live4;
        float4* full = reinterpret_cast<float4*>(gp.g[t]) + e;
        float4* pk = reinterpret_cast<float4*>(gp.packed + t * gp.live) + e;
        if (dir == 0) {
            *pk = *full;
        } else {
            const float4 v = *pk;
            *full = v;
            if (t >= 2) {                                    // synthesis: mirror rows 1..N/2-1 onto rows N-1..N/2+1
                const long row = (e * 4) / gp.N, col4 = e - row * (gp.N / 4);
                if (row >= 1 && row < gp.N / 2) {
                    float4* m = reinterpret_cast<float4*>(gp.g[t]) + (gp.N - row) * (gp.N / 4) + col4;
                    *m = t == 2 ? v : make_float4(-v.x, -v.y, -v.z, -v.w);
                }
            }
        }
    }
    // autoencoder tensors: one block per tensor slice
    for (int t = 4 + blockIdx.x; t < ST_NUM_PARAMS; t += gridDim.x) {
        float* full = gp.g[t];
        float* pk = gp.packed + 4 * gp.live + gp.ae_off[t - 4];
        for (int e = threadIdx.x; e < gp.ae_n[t - 4]; e += blockDim.x) {
            if (dir == 0) pk[e] = full[e];
            else full[e] = pk[e];
        }
    }
}

void st_launch_pack_grads(const GradPack& gp, int dir, cudaStream_t s) { pack_grads_kernel<<<148 * 4, 256, 0, s>>>(gp, dir); }

// Unpack (as above, dir = 1) fused with the L1 norm of the four unpacked DFT tensors and the clip coefficient
// (nn_proc.py:299-302) for the Adam launch that follows: mirrored synthesis rows count twice, like the tensors they restore.
__global__ void __launch_bounds__(256)
unpack_clip_kernel(GradPack gp, float grad_scale, float max_norm, float* __restrict__ norm_out, float* __restrict__ coef_out,
                   float* scratch, unsigned* counter) {
    const long live4 = gp.live / 4;
    const long stride = (long)gridDim.x * blockDim.x;
    float acc[1] = {0.f};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < 4 * live4; i += stride) {
        const int t = (int)(i / live4);
        const long e = i - t * live4;
        const float4 v = *(reinterpret_cast<const float4*>(gp.packed + t * gp.live) + e);
        *(reinterpret_cast<float4*>(gp.g[t]) + e) = v;
        float mult = 1.f;
        if (t >= 2) {
            const long row = (e * 4) / gp.N, col4 = e - row * (gp.N / 4);
            if (row >= 1 && row < gp.N / 2) {
                float4* m = reinterpret_cast<float4*>(gp.g[t]) + (gp.N - row) * (gp.N / 4) + col4;
                *m = t == 2 ? v : make_float4(-v.x, -v.y, -v.z, -v.w);
                mult = 2.f;
            }
        }
        acc[0] += mult * ((fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w)));
    }
    for (int t = 4 + blockIdx.x; t < ST_NUM_PARAMS; t += gridDim.x) {
        float* full = gp.g[t];
        const float* pk = gp.packed + 4 * gp.live + gp.ae_off[t - 4];
        for (int e = threadIdx.x; e < gp.ae_n[t - 4]; e += blockDim.x) full[e] = pk[e];
    }
    grid_finish<1>(acc, scratch, counter, [&](const double* s) {
        const double total_norm = s[0] * (double)grad_scale;
        if (norm_out) norm_out[0] = (float)total_norm;
        float cf = 1.f;
        if (max_norm > 0.f) {
            const double cc = (double)max_norm / (total_norm + 1e-6);
            cf = cc < 1.0 ? (float)cc : 1.f;
        }
        coef_out[0] = cf;
    });
}
void st_launch_unpack_clip(const GradPack& gp, float grad_scale, float max_norm, float* norm_out, float* coef_out, float* scratch,
                           unsigned* counter, cudaStream_t s) {
    unpack_clip_kernel<<<148 * 4, 256, 0, s>>>(gp, grad_scale, max_norm, norm_out, coef_out, scratch, counter);
}

// Split-K sum of the DFT gradient planes straight into the packed exchange payload: rows k < F of the four tensors only (no
// dead analysis rows, no mirrored synthesis rows -- the unpack restores those after the allreduce).
__global__ void __launch_bounds__(256)
finalize_packed_kernel(StDims d, const float* __restrict__ pa, const float* __restrict__ ps, int sa, int ss, float* __restrict__ packed) {
    const int n4 = d.N >> 2;
    const long plane = 2L * d.Fp * d.N, live = (long)d.F * d.N;
    const long total = (long)d.F * n4;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i / n4);
        const int c = (int)(i - (long)k * n4) << 2;
        const long o = (long)k * d.N + c;
        float4 ar = z4, ai = z4, sr = z4, si = z4;
        for (int s = 0; s < sa; ++s) {
            const float4 a = *reinterpret_cast<const float4*>(pa + s * plane + (long)k * d.N + c);
            const float4 b = *reinterpret_cast<const float4*>(pa + s * plane + (long)(d.Fp + k) * d.N + c);
            ar.x += a.x; ar.y += a.y; ar.z += a.z; ar.w += a.w;
            ai.x += b.x; ai.y += b.y; ai.z += b.z; ai.w += b.w;
        }
        for (int s = 0; s < ss; ++s) {
            const float4 a = *reinterpret_cast<const float4*>(ps + s * plane + (long)k * d.N + c);
            const float4 b = *reinterpret_cast<const float4*>(ps + s * plane + (long)(d.Fp + k) * d.N + c);
            sr.x += a.x; sr.y += a.y; sr.z += a.z; sr.w += a.w;
            si.x += b.x; si.y += b.y; si.z += b.z; si.w += b.w;
        }
        *reinterpret_cast<float4*>(packed + o) = ar;
        *reinterpret_cast<float4*>(packed + live + o) = ai;
        *reinterpret_cast<float4*>(packed + 2 * live + o) = sr;
        *reinterpret_cast<float4*>(packed + 3 * live + o) = si;
    }
}
void st_launch_finalize_packed(const StDims& d, const float* pa, const float* ps, int sa, int ss, float* packed, cudaStream_t s) {
    finalize_packed_kernel<<<148 * 4, 256, 0, s>>>(d, pa, ps, sa, ss, packed);
}

void st_launch_adam(const AdamTensors& t, const int2* chunk_map, int nchunks, const AdamScalars& sc, const float* clip_coef,
                    cudaStream_t s) {
    adam_kernel<<<nchunks, 256, 0, s>>>(t, chunk_map, sc, clip_coef);
}
// The kernel's address, so that a captured step can find its Adam node and refresh the per-step scalars (st_api.cu).
const void* st_adam_kernel_fn() { return reinterpret_cast<const void*>(&adam_kernel); }
const void* st_ola_loss_kernel_fn() { return reinterpret_cast<const void*>(&ola_loss_kernel); }
