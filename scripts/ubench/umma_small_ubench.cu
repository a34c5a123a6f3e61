// Cost of small tcgen05.mma instructions (M = 128, K = 8 tf32 / K = 16 f16, N = 16..64), A from tensor memory (TS) vs
// shared memory (SS): REPS back-to-back MMAs issued by one thread, then tcgen05.commit; clocks from first issue to the
// mbarrier completion.  Answers: is a TS-mode MMA with a tiny N bound by the A read from TMEM?   (B200; results in DESIGN.md)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../signaltrain_b200/csrc/st_tc_prims.cuh"

using namespace st_tc;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Converged-warp issue: all 32 lanes run the loop, elect.sync inside the asm picks the issuing lane; TMEM addresses are
// compile-time constants (a 512-column allocation starts at column 0) and the descriptor is (uniform base + immediate).
__device__ __forceinline__ void umma_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 bd;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int N, int REPS>
__global__ void __launch_bounds__(128, 1) bench_conv_kernel(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float* fs = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) fs[i] = 0.001f * (i % 97);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (slot != 0) __trap();
    {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(0.01f * (threadIdx.x + i));
        const uint32_t ta = ((uint32_t)(threadIdx.x & ~31) << 16);
        for (int c = 0; c < 128; c += 16) tmem_st16(ta + 256 + c, r);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t b_s = smem_u32(smem + 32768);
        const uint32_t dlo = ((b_s >> 4) & 0x3FFF) | (1u << 16);
        constexpr uint32_t dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < REPS; ++i)
            umma_ts_lohi(0u, 256u + 8u * (i & 7), dlo + 2u * (i & 3), dhi, idesc, i > 0 ? 1u : 0u);
        const long long t1 = clock64();
        umma_commit_elect(&bar);
        mbar_wait_spin(&bar, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(0u, 512); }
}
template <int N, int REPS>
void run_conv(const char* name) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench_conv_kernel<N, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int w = 0; w < 2; ++w) bench_conv_kernel<N, REPS><<<1, 128, 100 * 1024>>>(d);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    printf("%-28s N=%3d  %4d MMAs: issue %6lld clk (%.1f / MMA)   until complete %6lld clk (%.1f / MMA)\n", name, N, REPS, h[0],
           (double)h[0] / REPS, h[1], (double)h[1] / REPS);
    cudaFree(d);
}

// SS-mode, converged issue: A [M x 8] and B [N x 8] both from shared memory (the weight-gradient MMAs of the backward kernel).
__device__ __forceinline__ void umma_ss_lohi(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 ad, bd;\n\t"
        "mov.b64 ad, {%1, %3};\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int M, int N, int REPS>
__global__ void __launch_bounds__(128, 1) bench_conv_ss_kernel(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float* fs = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) fs[i] = 0.001f * (i % 97);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (slot != 0) __trap();
    if (warp == 0) {
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 32768);
        const uint32_t alo = ((a_s >> 4) & 0x3FFF) | (1u << 16), blo = ((b_s >> 4) & 0x3FFF) | (1u << 16);
        constexpr uint32_t dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < REPS; ++i)
            umma_ss_lohi(0u, alo + 2u * (i & 3) + 64u * ((i >> 2) & 3), blo + 2u * (i & 3), dhi, idesc, i > 0 ? 1u : 0u);
        const long long t1 = clock64();
        umma_commit_elect(&bar);
        mbar_wait_spin(&bar, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(0u, 512); }
}
template <int M, int N, int REPS>
void run_conv_ss(const char* name) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench_conv_ss_kernel<M, N, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int w = 0; w < 2; ++w) bench_conv_ss_kernel<M, N, REPS><<<1, 128, 100 * 1024>>>(d);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    printf("%-28s M=%3d N=%3d  %4d MMAs: issue %6lld clk (%.1f / MMA)   until complete %6lld clk (%.1f / MMA)\n", name, M, N, REPS, h[0],
           (double)h[0] / REPS, h[1], (double)h[1] / REPS);
    cudaFree(d);
}

// mode 0: TS tf32, 1: SS tf32, 2: TS f16 (fp16 operands, K = 16)
template <int MODE, int N, int REPS>
__global__ void __launch_bounds__(128, 1) bench_kernel(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float* fs = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) fs[i] = 0.001f * (i % 97);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    // fill the A columns of TMEM with something finite
    {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(0.01f * (threadIdx.x + i));
        const uint32_t ta = tb + ((uint32_t)(threadIdx.x & ~31) << 16);
        for (int c = 0; c < 128; c += 16) tmem_st16(ta + 256 + c, r);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 32768);
        uint32_t idesc;
        if (MODE == 2) idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
        else idesc = idesc_tf32(128, N, false, false);
        const long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < REPS; ++i) {
            const uint32_t ks = i & 3;
            const uint64_t bd = desc_kmajor_sw128(b_s + ks * 32);
            if (MODE == 0) umma_tf32_ts(tb, tb + 256 + 8 * (i & 7), bd, idesc, i > 0);
            else if (MODE == 1) umma_tf32(tb, desc_kmajor_sw128(a_s + ks * 32), bd, idesc, i > 0);
            else umma_f16_ts(tb, tb + 256 + 8 * (i & 7), bd, idesc, i > 0);
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int MODE, int N, int REPS>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench_kernel<MODE, N, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int w = 0; w < 2; ++w) bench_kernel<MODE, N, REPS><<<1, 128, 100 * 1024>>>(d);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    printf("%-28s N=%3d  %4d MMAs: issue %6lld clk (%.1f / MMA)   until complete %6lld clk (%.1f / MMA)\n", name, N, REPS, h[0],
           (double)h[0] / REPS, h[1], (double)h[1] / REPS);
    cudaFree(d);
}

int main() {
    run_conv<16, 64>("TS tf32 converged+const");
    run_conv<32, 64>("TS tf32 converged+const");
    run_conv<64, 64>("TS tf32 converged+const");
    run_conv<128, 64>("TS tf32 converged+const");
    run_conv_ss<128, 16, 64>("SS tf32 converged+const");
    run_conv_ss<128, 32, 64>("SS tf32 converged+const");
    run_conv_ss<128, 64, 64>("SS tf32 converged+const");
    run_conv_ss<128, 96, 64>("SS tf32 converged+const");
    run_conv_ss<128, 128, 64>("SS tf32 converged+const");
    run_conv_ss<64, 16, 64>("SS tf32 converged+const");
    run_conv_ss<64, 32, 64>("SS tf32 converged+const");
    run_conv_ss<64, 64, 64>("SS tf32 converged+const");
    run_conv_ss<64, 128, 64>("SS tf32 converged+const");
    run<0, 16, 64>("TS tf32 (A in TMEM)");
    run<0, 32, 64>("TS tf32 (A in TMEM)");
    run<0, 64, 64>("TS tf32 (A in TMEM)");
    run<0, 128, 64>("TS tf32 (A in TMEM)");
    run<0, 256, 64>("TS tf32 (A in TMEM)");
    run<0, 16, 256>("TS tf32 (A in TMEM)");
    run<1, 16, 64>("SS tf32 (A in smem)");
    run<1, 32, 64>("SS tf32 (A in smem)");
    run<1, 64, 64>("SS tf32 (A in smem)");
    run<1, 256, 64>("SS tf32 (A in smem)");
    run<2, 16, 64>("TS f16 K=16 (A in TMEM)");
    run<2, 64, 64>("TS f16 K=16 (A in TMEM)");
    return 0;
}
