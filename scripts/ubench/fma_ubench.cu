// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue rates, alone and mixed with broadcast LDS.128, on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
template <int MODE>
__global__ void k(float* out, const float* in, int iters, long long* cyc) {
    __shared__ __align__(16) float w[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) w[i] = in[i];
    __syncthreads();
    float a0 = in[threadIdx.x], a1 = in[threadIdx.x + 1];
    long long t0 = clock64();
    if (MODE == 0) {          // scalar FFMA, 16 chains
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a0, a1);
        }
        float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 1) {   // FFMA2, 16 chains
        unsigned long long acc[16]; unsigned long long a = pack2(a0, a1), b = pack2(a1, a0);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = pack2(i, i + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) ffma2(acc[i], a, b);
        }
        unsigned long long s = 0; for (int i = 0; i < 16; ++i) s ^= acc[i];
        reinterpret_cast<unsigned long long*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 2 || MODE == 3) {   // FFMA2 with weights from broadcast LDS.128: MODE 2 -> 2 FFMA2 per LDS, MODE 3 -> 4 per LDS
        unsigned long long acc[16]; unsigned long long a = pack2(a0, a1), a2 = pack2(a1, a0);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = pack2(i, i + 1);
        for (int it = 0; it < iters; ++it) {
            const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(w) + (it & 7) * 32;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { ulonglong2 v = wp[r * 8 + i]; ffma2(acc[2 * i], a, v.x); ffma2(acc[2 * i + 1], a, v.y); }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ulonglong2 v = wp[r * 4 + i]; ffma2(acc[4 * i], a, v.x); ffma2(acc[4 * i + 1], a, v.y);
                                                  ffma2(acc[4 * i + 2], a2, v.x); ffma2(acc[4 * i + 3], a2, v.y); }
                }
            }
        }
        unsigned long long s = 0; for (int i = 0; i < 16; ++i) s ^= acc[i];
        reinterpret_cast<unsigned long long*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 4) {   // scalar FFMA with weights from broadcast LDS.128: 4 FFMA per LDS
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = i;
        for (int it = 0; it < iters; ++it) {
            const float4* wp = reinterpret_cast<const float4*>(w) + (it & 7) * 32;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) { float4 v = wp[r * 4 + i]; acc[4*i] = fmaf(a0, v.x, acc[4*i]); acc[4*i+1] = fmaf(a0, v.y, acc[4*i+1]);
                                          acc[4*i+2] = fmaf(a0, v.z, acc[4*i+2]); acc[4*i+3] = fmaf(a0, v.w, acc[4*i+3]); }
        }
        float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, float* out, float* in, long long* cyc, double fma_per_inst) {
    const int iters = 4096;
    k<MODE><<<148, threads>>>(out, in, 16, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst_per_warp = 64.0 * iters;   // arithmetic instructions per warp
    const int warps_per_smsp = threads / 32 / 4;
    printf("%-34s threads=%4d  cyc=%9lld  clk/arith-inst/SMSP=%.3f  FMA/clk/SMSP=%.1f  err=%s\n", name, threads, c,
           (double)c / (inst_per_warp * (warps_per_smsp > 0 ? warps_per_smsp : 1)), fma_per_inst * 32 * inst_per_warp * (warps_per_smsp > 0 ? warps_per_smsp : 1) / c,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&in, 8192); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8192);
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("FFMA scalar", threads, out, in, cyc, 1);
        run<1>("FFMA2", threads, out, in, cyc, 2);
        run<2>("FFMA2 + LDS.128 (2:1)", threads, out, in, cyc, 2);
        run<3>("FFMA2 + LDS.128 (4:1)", threads, out, in, cyc, 2);
        run<4>("FFMA + LDS.128 (4:1)", threads, out, in, cyc, 1);
    }
    return 0;
}
