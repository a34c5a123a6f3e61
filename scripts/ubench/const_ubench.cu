// Micro-benchmark: FFMA2 with weights from __constant__ memory (immediate-offset operands vs uniform-indexed loads).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float2 CW[4096];
template <int MODE>
__global__ void k(float* out, const float* in, int iters, long long* cyc) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 8]);
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(i, i + 1);
    float2 acc2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = make_float2(i, i + 2);
    long long t0 = clock64();
    if (MODE == 0) {          // fully static constant indices: 64 FFMA2 per iteration, weights c[bank][imm]
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(a[r], CW[r * 8 + i], acc[i]);
        }
    } else if (MODE == 1) {   // uniform (loop-variant) constant index: 8 outputs x 8 k-pairs per iteration
        for (int it = 0; it < iters; ++it) {
            const float2* w = CW + (it & 31) * 64;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(a[r], w[r * 8 + i], acc[i]);
        }
    } else if (MODE == 3 || MODE == 4) {   // LDCU.128: 2 FFMA2 per load (MODE 3) or 4 FFMA2 per load with two activation rows (MODE 4)
        const float4* cw4 = reinterpret_cast<const float4*>(CW);
        for (int it = 0; it < iters; ++it) {
            const float4* w = cw4 + (it & 31) * 32;
#pragma unroll
            for (int r = 0; r < (MODE == 3 ? 8 : 4); ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = w[r * 4 + i];
                if (MODE == 3) {
                    acc[2 * i] = __ffma2_rn(a[r], make_float2(v.x, v.y), acc[2 * i]);
                    acc[2 * i + 1] = __ffma2_rn(a[r], make_float2(v.z, v.w), acc[2 * i + 1]);
                } else {
                    acc[2 * i] = __ffma2_rn(a[r], make_float2(v.x, v.y), acc[2 * i]);
                    acc[2 * i + 1] = __ffma2_rn(a[r], make_float2(v.z, v.w), acc[2 * i + 1]);
                    acc2[2 * i] = __ffma2_rn(a[r + 4], make_float2(v.x, v.y), acc2[2 * i]);
                    acc2[2 * i + 1] = __ffma2_rn(a[r + 4], make_float2(v.z, v.w), acc2[2 * i + 1]);
                }
            }
        }
    } else if (MODE == 2) {   // scalar FFMA, uniform constant index
        const float* cw = reinterpret_cast<const float*>(CW);
        float s[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] = i;
        for (int it = 0; it < iters; ++it) {
            const float* w = cw + (it & 31) * 64;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fmaf(a[r].x, w[r * 16 + i], s[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = make_float2(s[2 * i], s[2 * i + 1]);
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; ++i) { s.x += acc[i].x + acc2[i].x; s.y += acc[i].y + acc2[i].y; }
    reinterpret_cast<float2*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, float* out, float* in, long long* cyc, double fma_per_inst) {
    const int iters = 4096;
    k<MODE><<<148, threads>>>(out, in, 16, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst_per_warp = 64.0 * iters;
    const int wps = threads / 32 / 4;
    printf("%-40s threads=%4d  cyc=%9lld  clk/arith-inst/SMSP=%.3f  FMA/clk/SMSP=%.1f  err=%s\n", name, threads, c,
           (double)c / (inst_per_warp * wps), fma_per_inst * 32 * inst_per_warp * wps / c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&in, 8192); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8192);
    for (int threads : {128, 256, 512}) {
        run<0>("FFMA2 const imm", threads, out, in, cyc, 2);
        run<1>("FFMA2 const uniform-index", threads, out, in, cyc, 2);
        run<2>("FFMA const uniform-index", threads, out, in, cyc, 1);
        run<3>("FFMA2 LDCU.128 2:1", threads, out, in, cyc, 2);
        run<4>("FFMA2 LDCU.128 4:1", threads, out, in, cyc, 2);
    }
    return 0;
}
