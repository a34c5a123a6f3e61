// Micro-benchmark: which warps of a CTA share a scheduler (SMSP)?  Three warps run an FFMA2 chain, the others idle
// (spinning on a shared flag with nanosleep, like the consumers of the autoencoder backward).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__global__ void k(float* out, const float* in, int iters, long long* cyc, unsigned mask, int spin) {
    __shared__ volatile int flag;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    float a0 = in[threadIdx.x], a1 = in[threadIdx.x + 1];
    unsigned long long acc[8]; unsigned long long a = pack2(a0, a1), b = pack2(a1, a0);
    for (int i = 0; i < 8; ++i) acc[i] = pack2(i, i + 1);
    if ((mask >> warp) & 1) {
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) ffma2(acc[i], a, b);
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
        __threadfence_block();
        if ((threadIdx.x & 31) == 0) atomicAdd((int*)&flag, 1);
    } else if (spin) {
        while (flag < __popc(mask)) __nanosleep(32);
    }
    unsigned long long s = 0; for (int i = 0; i < 8; ++i) s ^= acc[i];
    reinterpret_cast<unsigned long long*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&in, 8192); cudaMalloc(&cyc, 8 * 32);
    cudaMemset(in, 0, 8192);
    const int iters = 4096;
    struct { unsigned mask; const char* name; } cases[] = {{0x7, "warps 0,1,2"}, {0x111, "warps 0,4,8"}, {0x1, "warp 0"}, {0xf, "warps 0-3"}, {0x249, "warps 0,3,6,9"}};
    for (int threads : {384, 352}) for (int spin : {0, 1}) for (auto& c : cases) {
        cudaMemset(cyc, 0, 8 * 32);
        k<<<148, threads>>>(out, in, 16, cyc, c.mask, spin);
        cudaDeviceSynchronize();
        k<<<148, threads>>>(out, in, iters, cyc, c.mask, spin);
        cudaDeviceSynchronize();
        long long h[32]; cudaMemcpy(h, cyc, 8 * 32, cudaMemcpyDeviceToHost);
        printf("threads=%d spin=%d %-14s clk/FFMA2 per warp:", threads, spin, c.name);
        for (int w = 0; w < 12; ++w) if ((c.mask >> w) & 1) printf(" w%d=%.2f", w, (double)h[w] / (64.0 * iters));
        printf("  %s\n", cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
