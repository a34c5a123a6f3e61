// Data step in front of the train step: the comp_4c target generator and the window cropper,
// on the device, so on-the-fly synthetic windows no longer come from one CPU core per ~1 k windows/s.
//
// Reference: audio.compressor_4controls (audio.py:380-426; feed-forward compressor with a branchy attack / release
// smoother of the gain change in dB) as Compressor_4c.go_wc calls it (:497-498) on float32 windows, and
// AudioFileDataSet.get_single_chunk's crop (datasets.py:236-241) + do_augment's polarity flip (:21-30).
// The arithmetic follows the dtypes of the numba-compiled reference: level detection and the static curve in float64,
// the gain change and its smoothed copy rounded to float32 at every sample, 10^(lin_A/20) * x in float64; the result is
// rounded to float32 (the caller's y.float(), train.py:120).
#include <algorithm>

#include "st_common.cuh"

namespace {

// gain change in dB (audio.py:398-410): g[n] = x_dB > thresh ? thresh + (x_dB - thresh)/ratio - x_dB : 0, rounded to float32
__global__ void comp_gain_kernel(const float* __restrict__ x, const double* __restrict__ knobs, int B, int n, float* __restrict__ g) {
    const long total = (long)B * n;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / n);
        const double thresh = knobs[4 * b], ratio = knobs[4 * b + 1];
        double x_dB = 20.0 * log10(fabs((double)x[i]) + 1e-8);
        x_dB = x_dB < -96.0 ? -96.0 : x_dB;
        g[i] = x_dB > thresh ? (float)(thresh + (x_dB - thresh) / ratio - x_dB) : 0.f;
    }
}

// the smoother (audio.py:412-418): sequential in n, one thread per window; lin_A overwrites g in place
__global__ void comp_smooth_kernel(const double* __restrict__ knobs, int B, int n, double sr, float* __restrict__ g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double aA = exp(-log(9.0) / (sr * knobs[4 * b + 2])), aR = exp(-log(9.0) / (sr * knobs[4 * b + 3]));
    float* row = g + (long)b * n;
    float prev = 0.f;
    row[0] = 0.f;
    for (int k0 = 1; k0 < n; k0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = k0 + q < n ? row[k0 + q] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double a = v[q] < prev ? aA : aR;
            prev = (float)((1.0 - a) * (double)v[q] + a * (double)prev);
            v[q] = prev;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (k0 + q < n) row[k0 + q] = v[q];
    }
}

// y = 10^(lin_A / 20) * x (audio.py:420-422), rounded to float32
__global__ void comp_apply_kernel(const float* __restrict__ x, const float* __restrict__ lin, long total, float* __restrict__ y) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        y[i] = (float)(pow(10.0, (double)lin[i] / 20.0) * (double)x[i]);
}

__global__ void crop_windows_kernel(const float* __restrict__ cx, const float* __restrict__ cy, const long* __restrict__ off,
                                    const float* __restrict__ sign, int B, int C, int L, float* __restrict__ ox, float* __restrict__ oy) {
    const long total = (long)B * (C + L);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / (C + L)), c = (int)(i - (long)b * (C + L));
        const float s = sign ? sign[b] : 1.f;
        if (c < C) ox[(long)b * C + c] = s * cx[off[b] + c];
        else oy[(long)b * L + (c - C)] = s * cy[off[b] + C - L + (c - C)];
    }
}

int grid_for(long items, int block) { return (int)std::min<long>((items + block - 1) / block, 148L * 16); }

}  // namespace

void st_launch_compressor_4c(const float* x, const double* knobs_wc, int B, int n, double sr, float* scratch, float* y, cudaStream_t s) {
    comp_gain_kernel<<<grid_for((long)B * n, 256), 256, 0, s>>>(x, knobs_wc, B, n, scratch);
    comp_smooth_kernel<<<(B + 31) / 32, 32, 0, s>>>(knobs_wc, B, n, sr, scratch);
    comp_apply_kernel<<<grid_for((long)B * n, 256), 256, 0, s>>>(x, scratch, (long)B * n, y);
}
void st_launch_crop_windows(const float* cx, const float* cy, const long* off, const float* sign, int B, int C, int L, float* ox,
                            float* oy, cudaStream_t s) {
    crop_windows_kernel<<<grid_for((long)B * (C + L), 256), 256, 0, s>>>(cx, cy, off, sign, B, C, L, ox, oy);
}
