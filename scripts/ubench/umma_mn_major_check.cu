// Correctness probe for MN-major SWIZZLE_128B tf32 operands of tcgen05.mma (both A and B from shared memory).
// The backward autoencoder kernel stages its weight-gradient operands as [row][feature] slices (row = reduction index K,
// feature = M or N index), which is the MN-major canonical layout: a block of 32 features x 32 rows is 4 KB, inside it a
// group of 8 rows is 1 KB (SBO), a row is 128 B, and the 16-byte chunk index of a row is XORed with (row & 7).
// Checks D = A^T-tile x B over K = 32 rows (4 k-steps) against the CPU for: plain, A start moved back by whole blocks
// (live features at a lane offset) and an N tile that spans two blocks.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "../../signaltrain_b200/csrc/st_tc_prims.cuh"

using namespace st_tc;

__host__ __device__ inline uint32_t mn_off(int f, int k) {       // byte offset of (feature f, row k) from the operand base
    const int blk = f >> 5, fi = f & 31, kg = k >> 3, r = k & 7;
    return (uint32_t)(blk * 4096 + kg * 1024 + r * 128 + ((((fi >> 2) ^ r) & 7) << 4) + ((fi & 3) << 2));
}
__host__ __device__ inline uint32_t k_off(int f, int k) {        // K-major SW128 reference layout: [feature][32 rows]
    return (uint32_t)(f * 128 + ((((k >> 2) ^ f) & 7) << 4) + ((k & 3) << 2));
}
__host__ __device__ inline float aval(int m, int k) { return (float)((m * 7 + k * 3) % 11 - 5); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k) % 7 - 3); }

__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}

// MF live A features placed at lane `moff` (multiple of 32), N columns
template <int N, bool MN>
__global__ void __launch_bounds__(128, 1) check_kernel(float* out, int MF, int moff) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    constexpr uint32_t A_OFF = 32768, B_OFF = 65536;          // A base leaves room for a negative block offset
    for (int i = threadIdx.x; i < 24576; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 777.f;   // garbage everywhere
    __syncthreads();
    // thread = (row k = lane, feature group = warp): what a chain warp does
    {
        const int k = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int f = w; f < MF; f += 4) *reinterpret_cast<float*>(smem + A_OFF + (MN ? mn_off(f, k) : k_off(f, k))) = aval(f, k);
        for (int f = w; f < N; f += 4) *reinterpret_cast<float*>(smem + B_OFF + (MN ? mn_off(f, k) : k_off(f, k))) = bval(f, k);
    }
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
        const uint32_t a_s = smem_u32(smem + A_OFF) - (MN ? (uint32_t)(moff >> 5) * 4096u : (uint32_t)moff * 128u), b_s = smem_u32(smem + B_OFF);
        // lo word: start address >> 4 | LBO (MN block stride 4 KB) >> 4 << 16;  hi word: SBO (8-row group stride 1 KB) >> 4 | version 1 | SWIZZLE_128B
        constexpr uint32_t dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (MN ? (1u << 15) | (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t kadv = MN ? 1024u : 32u, lbo = MN ? (4096u >> 4) : 1u;
            const uint64_t ad = ((uint64_t)dhi << 32) | ((((a_s + ks * kadv) >> 4) & 0x3FFFu) | (lbo << 16));
            const uint64_t bd = ((uint64_t)dhi << 32) | ((((b_s + ks * kadv) >> 4) & 0x3FFFu) | (lbo << 16));
            umma_ss(0u, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit_elect(&bar);
        mbar_wait_spin(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
        const uint32_t ta = ((uint32_t)(threadIdx.x & ~31) << 16);
        for (int c = 0; c < N; c += 16) {
            uint32_t r[16];
            tmem_ld16(ta + c, r);
            tmem_wait_ld();
            for (int i = 0; i < 16; ++i) out[threadIdx.x * N + c + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(0u, 512); }
}

template <int N, bool MN = true>
int run(int MF, int moff) {
    float* d;
    cudaMalloc(&d, 128 * N * sizeof(float));
    cudaFuncSetAttribute(check_kernel<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    check_kernel<N, MN><<<1, 128, 100 * 1024>>>(d, MF, moff);
    std::vector<float> h(128 * N);
    cudaError_t e = cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N=%d MF=%d moff=%d: CUDA error %s\n", N, MF, moff, cudaGetErrorString(e)); return 1; }
    int bad = 0;
    for (int m = 0; m < MF; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 32; ++k) ref += aval(m, k) * bval(n, k);
            const float got = h[(moff + m) * N + n];
            if (got != ref && bad++ < 5) printf("  mismatch m=%d n=%d got %g want %g\n", m, n, got, ref);
        }
    printf("%s-major SW128 tf32  N=%3d MF=%3d lane offset %3d: %s (%d wrong of %d)\n", MN ? "MN" : "K", N, MF, moff, bad ? "FAIL" : "ok", bad, MF * N);
    cudaFree(d);
    return bad != 0;
}

// Sweep kernel: every descriptor field at run time, to find which combinations the hardware accepts.
struct Var { int amn, bmn, lbo, sbo, kadv; };
__global__ void __launch_bounds__(128, 1) sweep_kernel(float* out, Var v) {
    constexpr int N = 32, MF = 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    constexpr uint32_t A_OFF = 32768, B_OFF = 65536;
    for (int i = threadIdx.x; i < 24576; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 777.f;
    __syncthreads();
    {
        const int k = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int f = w; f < MF; f += 4) *reinterpret_cast<float*>(smem + A_OFF + (v.amn ? mn_off(f, k) : k_off(f, k))) = aval(f, k);
        for (int f = w; f < N; f += 4) *reinterpret_cast<float*>(smem + B_OFF + (v.bmn ? mn_off(f, k) : k_off(f, k))) = bval(f, k);
    }
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
    {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(555.f);
        const uint32_t ta = ((uint32_t)(threadIdx.x & ~31) << 16);
        for (int c = 0; c < N; c += 16) tmem_st16(ta + c, r);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t a_s = smem_u32(smem + A_OFF), b_s = smem_u32(smem + B_OFF);
        const uint32_t dhi_k = (1024u >> 4) | (1u << 14) | (2u << 29), dhi_mn = ((uint32_t)v.sbo >> 4) | (1u << 14) | (2u << 29);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (v.amn ? (1u << 15) : 0u) | (v.bmn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = v.amn ? ((uint64_t)dhi_mn << 32) | ((((a_s + ks * v.kadv) >> 4) & 0x3FFFu) | (((uint32_t)v.lbo >> 4) << 16))
                                      : ((uint64_t)dhi_k << 32) | ((((a_s + ks * 32u) >> 4) & 0x3FFFu) | (1u << 16));
            const uint64_t bd = v.bmn ? ((uint64_t)dhi_mn << 32) | ((((b_s + ks * v.kadv) >> 4) & 0x3FFFu) | (((uint32_t)v.lbo >> 4) << 16))
                                      : ((uint64_t)dhi_k << 32) | ((((b_s + ks * 32u) >> 4) & 0x3FFFu) | (1u << 16));
            umma_ss(0u, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit_elect(&bar);
        mbar_wait_spin(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
        const uint32_t ta = ((uint32_t)(threadIdx.x & ~31) << 16);
        for (int c = 0; c < N; c += 16) {
            uint32_t r[16];
            tmem_ld16(ta + c, r);
            tmem_wait_ld();
            for (int i = 0; i < 16; ++i) out[threadIdx.x * N + c + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(0u, 512); }
}
void sweep(Var v) {
    constexpr int N = 32, MF = 128;
    float* d;
    cudaMalloc(&d, 128 * N * sizeof(float));
    cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    sweep_kernel<<<1, 128, 100 * 1024>>>(d, v);
    std::vector<float> h(128 * N);
    cudaError_t e = cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("sweep: CUDA error %s\n", cudaGetErrorString(e)); return; }
    int bad = 0;
    for (int m = 0; m < MF; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 32; ++k) ref += aval(m, k) * bval(n, k);
            if (h[m * N + n] != ref) ++bad;
        }
    printf("sweep A %s B %s lbo %5d sbo %5d kadv %5d: %4d wrong of %d   D[0][0..3] = %g %g %g %g  D[1][0..1] = %g %g (want %g %g %g %g ..)\n", v.amn ? "MN" : "K ",
           v.bmn ? "MN" : "K ", v.lbo, v.sbo, v.kadv, bad, MF * N, h[0], h[1], h[2], h[3], h[N], h[N + 1], -18.f, -12.f, 29.f, -35.f);
    cudaFree(d);
}

int main() {
    for (int am = 0; am < 2; ++am)
        for (int bm = 0; bm < 2; ++bm) {
            if (!am && !bm) { sweep({0, 0, 16, 1024, 32}); continue; }
            const int lbos[] = {4096, 1024, 128, 16}, sbos[] = {1024, 4096, 128};
            for (int lbo : lbos)
                for (int sbo : sbos) sweep({am, bm, lbo, sbo, 1024});
        }
    int rc = 0;
    rc |= run<32, false>(128, 0);
    rc |= run<64, false>(32, 64);
    rc |= run<32>(128, 0);
    rc |= run<32>(32, 0);
    rc |= run<64>(64, 0);
    rc |= run<64>(32, 64);
    rc |= run<128>(64, 64);
    rc |= run<16>(32, 96);
    printf(rc ? "MN-major check FAILED\n" : "MN-major check passed\n");
    return rc;
}
