// Micro-benchmark: the autoencoder backward's dgrad_run routine in isolation (N warps per CTA, no consumers, no barriers).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LD = 100, HALF = 32 * LD, TLD = 20;
struct TrackOut { float* gt[2]; const float* tail; long F; int T, tail0; };
__device__ __forceinline__ float elu_grad(float h) { return h > 0.f ? 1.f : h + 1.f; }
__device__ __forceinline__ void fma2v(float2& acc, const float2& a, float wx, float wy) {
    unsigned long long& c = reinterpret_cast<unsigned long long&>(acc);
    unsigned long long w;
    asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(wx), "f"(wy));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(w));
}
template <bool L1>
__device__ __noinline__ void dgrad_run(const float* __restrict__ w, int nst, int nblk, const float* __restrict__ gzrow,
                                       const float* __restrict__ hsrc, float* __restrict__ dst, const TrackOut* __restrict__ trk) {
    float2 acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = make_float2(0.f, 0.f);
    float4 wa[8], wb[8], ga[2], gb[2], hv[2][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) wa[j] = *reinterpret_cast<const float4*>(w + 4 * j);
    ga[0] = *reinterpret_cast<const float4*>(gzrow);
    ga[1] = *reinterpret_cast<const float4*>(gzrow + HALF);
    const int total = nst * nblk;
    int st = 0, i = 0;
#pragma unroll 1
    for (int sidx = 0; sidx < total; sidx += 2) {
        // stage s+1 -> B registers, then the FMAs of stage s
#pragma unroll
        for (int j = 0; j < 8; ++j) wb[j] = *reinterpret_cast<const float4*>(w + 32 + 4 * j);
        gb[0] = *reinterpret_cast<const float4*>(gzrow + 4 * (st + 1));
        gb[1] = *reinterpret_cast<const float4*>(gzrow + HALF + 4 * (st + 1));
        st += 2;
        const bool last = st == nst;
        if (last) st = 0;
        if (!L1 && last) {                  // this iteration completes an output block: fetch its ELU' operands now
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                hv[r][0] = *reinterpret_cast<const float4*>(hsrc + r * HALF + i);
                hv[r][1] = *reinterpret_cast<const float4*>(hsrc + r * HALF + i + 4);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) fma2v(acc[r][j], make_float2(ga[r].x, ga[r].y), wa[j].x, wa[j].y);
#pragma unroll
            for (int j = 0; j < 8; ++j) fma2v(acc[r][j], make_float2(ga[r].z, ga[r].w), wa[j].z, wa[j].w);
        }
        // stage s+2 -> A registers (past the end of the layer: harmless reads of the next layer's weights), FMAs of s+1
#pragma unroll
        for (int j = 0; j < 8; ++j) wa[j] = *reinterpret_cast<const float4*>(w + 64 + 4 * j);
        ga[0] = *reinterpret_cast<const float4*>(gzrow + 4 * st);
        ga[1] = *reinterpret_cast<const float4*>(gzrow + HALF + 4 * st);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) fma2v(acc[r][j], make_float2(gb[r].x, gb[r].y), wb[j].x, wb[j].y);
#pragma unroll
            for (int j = 0; j < 8; ++j) fma2v(acc[r][j], make_float2(gb[r].z, gb[r].w), wb[j].z, wb[j].w);
        }
        w += 64;
        if (last) {                         // an output block is complete
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float c[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    c[j] = acc[r][j].x + acc[r][j].y;
                    acc[r][j] = make_float2(0.f, 0.f);
                }
                if (!L1) {                  // c * ELU'(h),  ELU'(h) = min(h, 0) + 1
                    const float4 h0 = hv[r][0], h1 = hv[r][1];
                    float* dp = dst + r * HALF + i;
                    *reinterpret_cast<float4*>(dp) = make_float4(fmaf(c[0], fminf(h0.x, 0.f), c[0]), fmaf(c[1], fminf(h0.y, 0.f), c[1]),
                                                                 fmaf(c[2], fminf(h0.z, 0.f), c[2]), fmaf(c[3], fminf(h0.w, 0.f), c[3]));
                    *reinterpret_cast<float4*>(dp + 4) = make_float4(fmaf(c[4], fminf(h1.x, 0.f), c[4]), fmaf(c[5], fminf(h1.y, 0.f), c[5]),
                                                                     fmaf(c[6], fminf(h1.z, 0.f), c[6]), fmaf(c[7], fminf(h1.w, 0.f), c[7]));
                } else if (trk->gt[r]) {    // dL/d(track) + the skip / residual gradient on the last OT frames
                    const float* tl = trk->tail + r * 32 * TLD;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int tt = i + j;
                        if (tt < trk->T) trk->gt[r][(long)tt * trk->F] = c[j] + (tt >= trk->tail0 ? tl[tt - trk->tail0] : 0.f);
                    }
                }
            }
            i += 8;
        }
    }
}

__global__ void k(float* out, int iters, long long* cyc, int nwarps_active, int nst, int nblk) {
    extern __shared__ __align__(16) float smem[];
    float* W = smem;                       // 9216 floats of weights
    float* slots = W + 9216;               // per warp: 2 slots of 64 x LD
    for (int i = threadIdx.x; i < 9216 + 3 * 2 * 64 * LD; i += blockDim.x) smem[i] = 0.001f * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nwarps_active) return;
    float* s0 = slots + (warp * 2) * 64 * LD + lane * LD;
    float* s1 = s0 + 64 * LD;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) dgrad_run<false>(W, nst, nblk, s0, s0 + 64, s1, nullptr);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s1[0];
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8 * 32);
    const size_t smem = sizeof(float) * (9216 + 3 * 2 * 64 * LD);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int iters = 200;
    struct { int nst, nblk; const char* name; } shapes[] = {{16, 4, "KD=64 NI=32"}, {8, 8, "KD=32 NI=64"}, {4, 2, "KD=16 NI=16"}, {4, 8, "KD=16 NI=64"}};
    for (int threads : {384}) for (int nw : {3}) for (auto& sh : shapes) {
        k<<<148, threads, smem>>>(out, 2, cyc, nw, sh.nst, sh.nblk);
        cudaDeviceSynchronize();
        k<<<148, threads, smem>>>(out, iters, cyc, nw, sh.nst, sh.nblk);
        cudaDeviceSynchronize();
        long long h[32]; cudaMemcpy(h, cyc, 8 * 32, cudaMemcpyDeviceToHost);
        const double ffma2 = 32.0 * sh.nst * sh.nblk * iters;
        printf("threads=%3d active warps=%d %-12s clk per call=%8.0f  clk/FFMA2=%.2f  %s\n", threads, nw, sh.name, (double)h[0] / iters,
               (double)h[0] / ffma2, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
