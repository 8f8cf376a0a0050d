// Score post-processing on the device: what the reference's ood_detection.py:150-206 does with pandas
// on the result CSVs - per-t z-scores against the validation set, mean z per file, ROC-AUC - for score tensors that
// are already on the GPU ([n_t, n_images] fp32, the layout BatchReconstructor.score_batch returns).
#include "../../include/ddpm_ood_b200.h"

#include <cuda_runtime.h>

#include "conv_gemm.cuh"  // set_error

namespace ddpm {

// One CTA per t: mean and sample standard deviation (ddof = 1, pandas' default) over the validation images.
// Two passes in double precision (n_val is small; the reference computes in float64).
__global__ void __launch_bounds__(256) val_stats_kernel(const float* __restrict__ val, int n, float* __restrict__ mean,
                                                        float* __restrict__ std) {
    __shared__ double s_red[256];
    const float* row = val + static_cast<size_t>(blockIdx.x) * n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(row[i]);
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    const double m = s_red[0] / n;
    __syncthreads();
    acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = static_cast<double>(row[i]) - m;
        acc += d * d;
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        mean[blockIdx.x] = static_cast<float>(m);
        std[blockIdx.x] = static_cast<float>(sqrt(s_red[0] / (n - 1)));
    }
}

// out[i] = mean over t of (scores[t, i] - mean[t]) / std[t]
__global__ void mean_z_kernel(const float* __restrict__ scores, const float* __restrict__ mean,
                              const float* __restrict__ std, int n_t, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int t = 0; t < n_t; ++t)
        acc += (static_cast<double>(scores[static_cast<size_t>(t) * n + i]) - mean[t]) / static_cast<double>(std[t]);
    out[i] = static_cast<float>(acc / n_t);
}

// counts[0] += #{(o, i): out[o] > in[i]}, counts[1] += #{out[o] == in[i]}  (integer atomics: order-independent)
__global__ void __launch_bounds__(256) auc_counts_kernel(const float* __restrict__ in_s, int n_in,
                                                         const float* __restrict__ out_s, int n_out,
                                                         unsigned long long* __restrict__ counts) {
    __shared__ float s_in[1024];
    unsigned long long gt = 0, eq = 0;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const float v = o < n_out ? out_s[o] : 0.f;
    for (int base = 0; base < n_in; base += 1024) {
        const int m = min(1024, n_in - base);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) s_in[j] = in_s[base + j];
        __syncthreads();
        if (o < n_out) {
            for (int j = 0; j < m; ++j) {
                gt += v > s_in[j];
                eq += v == s_in[j];
            }
        }
    }
    if (gt) atomicAdd(&counts[0], gt);
    if (eq) atomicAdd(&counts[1], eq);
}

}  // namespace ddpm

extern "C" {

int ddpm_val_stats(const float* val, int n_t, int n_val, float* mean, float* std, void* stream) {
    if (!val || !mean || !std || n_t < 1 || n_val < 2) { ddpm::set_error("ddpm_val_stats: need n_t >= 1 and n_val >= 2"); return 2; }
    ddpm::val_stats_kernel<<<n_t, 256, 0, static_cast<cudaStream_t>(stream)>>>(val, n_val, mean, std);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("ddpm_val_stats: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

int ddpm_mean_z(const float* scores, const float* mean, const float* std, int n_t, int n, float* out, void* stream) {
    if (!scores || !mean || !std || !out || n_t < 1 || n < 1) { ddpm::set_error("ddpm_mean_z: bad argument"); return 2; }
    ddpm::mean_z_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, mean, std, n_t, n, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("ddpm_mean_z: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

int ddpm_auc_counts(const float* in_scores, int n_in, const float* out_scores, int n_out, unsigned long long* counts,
                    void* stream) {
    if (!in_scores || !out_scores || !counts || n_in < 1 || n_out < 1) { ddpm::set_error("ddpm_auc_counts: bad argument"); return 2; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), s);
    ddpm::auc_counts_kernel<<<(n_out + 255) / 256, 256, 0, s>>>(in_scores, n_in, out_scores, n_out, counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("ddpm_auc_counts: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ data ingest (f-4)
// Per-image intensity scaling to [0, 1] on the device: MONAI's ScaleIntensityd(minv=0, maxv=1) as the reference's
// loader applies it (src/data/get_train_and_val_dataloader.py:76), for a batch of raw images that was copied to the
// GPU as stored (uint8 or float32, [N, per_image] contiguous). A constant image maps to zeros.
namespace ddpm {

template <typename T>
__global__ void __launch_bounds__(256) scale_intensity_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                                              long long per_image) {
    __shared__ float s_min[8], s_max[8];
    const T* in = src + static_cast<size_t>(blockIdx.x) * per_image;
    float* out = dst + static_cast<size_t>(blockIdx.x) * per_image;
    float mn = INFINITY, mx = -INFINITY;
    for (long long i = threadIdx.x; i < per_image; i += blockDim.x) {
        const float v = static_cast<float>(in[i]);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = s_min[0]; mx = s_max[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, s_min[i]); mx = fmaxf(mx, s_max[i]); }
    const float range = mx - mn;
    for (long long i = threadIdx.x; i < per_image; i += blockDim.x) {
        const float v = static_cast<float>(in[i]) - mn;
        out[i] = range > 0.f ? v / range : v;
    }
}

}  // namespace ddpm

extern "C" int ddpm_scale_intensity(const void* src, int src_is_u8, float* dst, int N, long long per_image, void* stream) {
    if (!src || !dst || N < 1 || per_image < 1) { ddpm::set_error("ddpm_scale_intensity: bad argument"); return 2; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (src_is_u8)
        ddpm::scale_intensity_kernel<unsigned char><<<N, 256, 0, s>>>(static_cast<const unsigned char*>(src), dst, per_image);
    else
        ddpm::scale_intensity_kernel<float><<<N, 256, 0, s>>>(static_cast<const float*>(src), dst, per_image);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("ddpm_scale_intensity: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}
