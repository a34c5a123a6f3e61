// Loss (loss_functions.py:9-10,22-23,26-43), L1 gradient clip (nn_proc.py:299-302) and Adam
// (torch.optim.Adam as constructed at train.py:228).  All HBM-streaming: float4 loads, warp-shuffle
// block reductions, and a deterministic last-block final reduction (fixed summation order, no float atomics).
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
void st_launch_adam(const AdamTensors& t, const int2* chunk_map, int nchunks, const AdamScalars& sc, const float* clip_coef,
                    cudaStream_t s) {
    adam_kernel<<<nchunks, 256, 0, s>>>(t, chunk_map, sc, clip_coef);
}
