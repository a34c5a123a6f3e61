// FP32 FFMA GEMM (fallback for shapes the tcgen05 path does not cover, and its cross-check in the tests).
//
//   C[M,N] = op(A) * op(B),   fp32 in / fp32 accumulate, 128x128x8 tiles, 256 threads, 8x8 per thread,
//   register-prefetched double-buffered shared memory, optional split-K into separate partial planes.
//
// An operand is either "k-contiguous"  X[row][k]  (row = m or n)  or "row-contiguous"  X[k][row].
// Rows of an operand may OVERLAP (leading dimension = hop < row length): frame (b, t) of the padded waveform is row
// b*Tp + t of a uniform-stride view, which is how Conv1d(stride=hop) (cls_fe_dft.py:28-31,55-56) and ConvTranspose1d(stride=hop)
// (cls_fe_dft.py:78-82,112) and their weight/data gradients become plain GEMMs without im2col.
#include "st_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, LDS = BM + 4;

__device__ __forceinline__ float4 op_load4(const GemmOperand& o, long off) {
    float4 v = *reinterpret_cast<const float4*>(o.ptr + off);
    if (o.lo) {
        const float4 l = *reinterpret_cast<const float4*>(o.lo + off);
        v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
    }
    return v;
}

// Fetch this thread's float4 of a 128(rows) x 8(k) operand tile.
template <bool KC>
__device__ __forceinline__ float4 tile_fetch(const GemmOperand& o, int row0, int rows, int k0, int kend, int tid) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KC) {
        const int r = row0 + (tid >> 1), k = k0 + ((tid & 1) << 2);
        if (r < rows && k < kend) v = op_load4(o, (long)r * o.ld + k);
    } else {
        const int k = k0 + (tid >> 5), r = row0 + ((tid & 31) << 2);
        if (k < kend && r < rows) v = op_load4(o, (long)k * o.ld + r);
    }
    return v;
}

template <bool KC>
__device__ __forceinline__ void tile_stash(float (*sm)[LDS], float4 v, int tid) {
    if (KC) {
        const int r = tid >> 1, k = (tid & 1) << 2;
        sm[k + 0][r] = v.x; sm[k + 1][r] = v.y; sm[k + 2][r] = v.z; sm[k + 3][r] = v.w;
    } else {
        *reinterpret_cast<float4*>(&sm[tid >> 5][(tid & 31) << 2]) = v;
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2)
gemm_simt_kernel(GemmOperand A, GemmOperand B, float* __restrict__ C, long ldc, int M, int N, int K, int k_per_split,
                 long split_stride) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    C += (long)blockIdx.z * split_stride;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int ktiles = (kend - kbeg + BK - 1) / BK;
    float4 ra = tile_fetch<A_KC>(A, m0, M, kbeg, kend, tid);
    float4 rb = tile_fetch<B_KC>(B, n0, N, kbeg, kend, tid);
    if (ktiles > 0) {
        tile_stash<A_KC>(As[0], ra, tid);
        tile_stash<B_KC>(Bs[0], rb, tid);
    }
    __syncthreads();
    for (int kt = 0; kt < ktiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) {
            ra = tile_fetch<A_KC>(A, m0, M, kbeg + (kt + 1) * BK, kend, tid);
            rb = tile_fetch<B_KC>(B, n0, N, kbeg + (kt + 1) * BK, kend, tid);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            tile_stash<A_KC>(As[buf ^ 1], ra, tid);
            tile_stash<B_KC>(Bs[buf ^ 1], rb, tid);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n < N)
                *reinterpret_cast<float4*>(C + (long)m * ldc + n) =
                    make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
        }
    }
}

}  // namespace

// K must be a multiple of 4 for k-contiguous operands; M (resp. N) a multiple of 4 for row-contiguous
// A (resp. B); N a multiple of 4 always (float4 stores).  Checked by the caller (st_api.cu).
int st_launch_gemm(bool a_kc, bool b_kc, const GemmOperand& A, const GemmOperand& B, float* C, long ldc, int M, int N,
                    int K, int splits, long split_stride, cudaStream_t s) {
    if (splits < 1) splits = 1;
    int kps = (K + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (K + kps - 1) / kps;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    if (a_kc && b_kc)
        gemm_simt_kernel<true, true><<<grid, 256, 0, s>>>(A, B, C, ldc, M, N, K, kps, split_stride);
    else if (a_kc && !b_kc)
        gemm_simt_kernel<true, false><<<grid, 256, 0, s>>>(A, B, C, ldc, M, N, K, kps, split_stride);
    else if (!a_kc && b_kc)
        gemm_simt_kernel<false, true><<<grid, 256, 0, s>>>(A, B, C, ldc, M, N, K, kps, split_stride);
    else
        gemm_simt_kernel<false, false><<<grid, 256, 0, s>>>(A, B, C, ldc, M, N, K, kps, split_stride);
    return splits;
}
