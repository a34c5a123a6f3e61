// Micro-benchmark: FFMA2 dependent-issue latency -- throughput of one warp per SMSP with 1..16 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
template <int NCH>
__global__ void k(float* out, const float* in, int iters, long long* cyc) {
    float a0 = in[threadIdx.x], a1 = in[threadIdx.x + 1];
    unsigned long long acc[NCH]; unsigned long long a = pack2(a0, a1), b = pack2(a1, a0);
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = pack2(i, i + 1);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 64 / NCH; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) ffma2(acc[i], a, b);
    }
    long long t1 = clock64();
    unsigned long long s = 0; for (int i = 0; i < NCH; ++i) s ^= acc[i];
    reinterpret_cast<unsigned long long*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NCH> void run(int threads, float* out, float* in, long long* cyc) {
    const int iters = 4096;
    k<NCH><<<148, threads>>>(out, in, 16, cyc);
    cudaDeviceSynchronize();
    k<NCH><<<148, threads>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const int wps = threads / 128;
    printf("chains=%2d warps/SMSP=%d  clk per FFMA2 per SMSP = %.3f   (per warp: %.3f)\n", NCH, wps, (double)c / (64.0 * iters * wps), (double)c / (64.0 * iters));
}
int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&in, 8192); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 8192);
    for (int threads : {128, 256}) {
        run<1>(threads, out, in, cyc); run<2>(threads, out, in, cyc); run<4>(threads, out, in, cyc); run<8>(threads, out, in, cyc); run<16>(threads, out, in, cyc);
    }
    return 0;
}
