// LPIPS (AlexNet, v0.1, linear heads, spatial average) on the device, fp32 throughout.
// Replaces `lpips.LPIPS.forward(in0, in1, normalize=...)` as called through the reference's wrapper
// src/losses/perceptual_loss.py:125,181-183 from src/trainers/reconstruct.py:170-187.
// The feature extractor is a few MFLOP per image (vs ~400 GFLOP for the reconstruction chain it scores): shared-memory
// tiled fp32 GEMM convolutions on CUDA cores (~0.5 ms per batch of 256 image pairs); the whole score stays on the device.
#include "lpips.cuh"

#include <string.h>

#include <cuda_fp16.h>
#include <stdint.h>

#include "conv_gemm.cuh"  // set_error, the tcgen05 implicit-GEMM conv

namespace ddpm {

int num_sms();  // api.cu

#define LP_CHECK(name)                                                             \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
            return 5;                                                              \
        }                                                                          \
    } while (0)

static const int kChn[5] = {64, 192, 384, 256, 256};
static const int kCin[5] = {3, 64, 192, 384, 256};
// kernel / stride / padding: 11/4/2, 5/1/2, 3/1/1 x3 (torchvision AlexNet features), template arguments of the conv kernel
static const int kK[5] = {11, 5, 3, 3, 3};

// in0/in1: [B, C, H, W] (C = 1 broadcasts to 3, like the reference's ScalingLayer does for grayscale); out: [2B,3,H,W]
__global__ void lpips_scale_kernel(const float* __restrict__ in0, const float* __restrict__ in1, float* __restrict__ out,
                                   int B, int C, int HW, int normalize, float3 shift, float3 scale) {
    const long long total = 2LL * B * 3 * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % HW);
        const int c = static_cast<int>((i / HW) % 3);
        const long long b2 = i / (3LL * HW);
        const float* src = b2 < B ? in0 : in1;
        const long long b = b2 < B ? b2 : b2 - B;
        float v = src[(b * C + (C == 1 ? 0 : c)) * HW + p];
        if (normalize) v = 2.0f * v - 1.0f;
        const float sh = c == 0 ? shift.x : (c == 1 ? shift.y : shift.z);
        const float sc = c == 0 ? scale.x : (c == 1 ? scale.y : scale.z);
        out[i] = (v - sh) / sc;
    }
}

// Convolution + bias + ReLU as a shared-memory tiled fp32 GEMM (NCHW in, NCHW out):
//   out[(n, p), co] = relu(bias[co] + sum_k patch[(n, p), k] * w[co, k]),   k = (ci, kh, kw)
// A CTA owns a 128 (image, pixel) x 64 output-channel tile; the patch matrix is gathered on the fly (zero padding), the
// filter rows are read once per 128 rows. 16 x 16 threads, 8 x 4 outputs each, K in steps of 16; the global loads of
// step i + 1 are in registers while step i's FMAs run. fp32 CUDA-core math on purpose: the reference computes LPIPS in
// fp32 (src/losses/perceptual_loss.py:107-108) and the score tolerance is 2e-4.
// CENTER: a 3x3 / pad 1 conv over a 1x1 map only ever sees its centre tap (AlexNet's last three convs on 32x32 images):
// the reduction runs over Cin with the centre weights, 1/9 of the work.
constexpr int kLpBM = 128, kLpBN = 64, kLpBK = 16;
template <int K, int STRIDE, int PAD, bool CENTER = false>
__global__ void __launch_bounds__(256) lpips_conv_relu_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int NB, int Cin, int H, int W, int Cout, int Ho, int Wo) {
    constexpr int BM = kLpBM, BN = kLpBN, BK = kLpBK;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int npix = Ho * Wo;
    const long long M = static_cast<long long>(NB) * npix;
    const int red = CENTER ? Cin : Cin * K * K;
    const long long m0 = static_cast<long long>(blockIdx.x) * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // loader mapping. A: row = tid % 128, k = 8 * (tid / 128) + j (j < 8). W: column = tid % 64, k = 4 * (tid / 64) + j
    const int arow = tid & 127, ak = (tid >> 7) * 8;
    const long long am = m0 + arow;
    const bool a_ok = am < M;
    const long long an = a_ok ? am / npix : 0;
    const int ap = a_ok ? static_cast<int>(am - an * npix) : 0;
    const int aho = ap / Wo, awo = ap - aho * Wo;
    const float* img = in + an * Cin * H * W;
    const int wcol = tid & 63, wk = (tid >> 6) * 4;
    const int wco = n0 + wcol;
    const bool w_ok = wco < Cout;
    const float* wrow = w + static_cast<long long>(w_ok ? wco : 0) * Cin * K * K;
    float ra[8], rw[4];
    auto gload = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + ak + j;
            float av = 0.f;
            if (k < red && a_ok) {
                if (CENTER) {
                    av = __ldg(img + k);
                } else {
                    const int ci = k / (K * K);
                    const int rem = k - ci * (K * K);
                    const int kh = rem / K, kw = rem - kh * K;
                    const int hi = aho * STRIDE + kh - PAD, wi = awo * STRIDE + kw - PAD;
                    if (hi >= 0 && hi < H && wi >= 0 && wi < W) av = __ldg(img + (ci * H + hi) * W + wi);
                }
            }
            ra[j] = av;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + wk + j;
            rw[j] = (k < red && w_ok) ? __ldg(CENTER ? wrow + k * (K * K) + (K * K) / 2 : wrow + k) : 0.f;
        }
    };
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    gload(0);
    for (int k0 = 0; k0 < red; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 8; ++j) As[ak + j][arow] = ra[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) Ws[wk + j][wcol] = rw[j];
        __syncthreads();
        if (k0 + BK < red) gload(k0 + BK);  // in flight during this step's FMAs
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= M) continue;
        const long long n = m / npix;
        const int pp = static_cast<int>(m - n * npix);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co < Cout) out[(n * Cout + co) * npix + pp] = fmaxf(acc[i][j] + __ldg(bias + co), 0.f);
        }
    }
}

__global__ void lpips_maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, long long NC, int H, int W,
                                     int Ho, int Wo) {
    const long long total = NC * Ho * Wo;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int wo = static_cast<int>(i % Wo);
        const int ho = static_cast<int>((i / Wo) % Ho);
        const long long nc = i / (static_cast<long long>(Wo) * Ho);
        float m = -INFINITY;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int hi = ho * 2 + kh, wi = wo * 2 + kw;
                if (hi < H && wi < W) m = fmaxf(m, in[(nc * H + hi) * W + wi]);
            }
        out[i] = m;
    }
}

struct LpipsFeat {
    const float* f[5];  // [2B, C_k, npix_k] fp32, or null when the layer's features are channels-last halves:
    const __half* fh[5];  // [2B, npix_k, C_k] fp16 hi
    const __half* fl[5];  // [2B, npix_k, C_k] fp16 lo (value = hi + lo)
    const float* lin[5];
    int npix[5];
    int cstride[5];       // channels per row of the halves (conv2's 192 channels live in rows of 256)
};

// ---- tensor-core path of conv2 (5x5, pad 2, 64 -> 192) and conv3 / conv4 / conv5 (3x3, pad 1, 192 -> 384 -> 256 -> 256)
// for large maps. conv2's 192 output channels are padded to 256 (zero weight rows, zero bias: relu(0) = 0), the second
// max-pool runs on the channels-last halves. conv1 (11x11 stride 4 over 3 channels, 10 % of the FLOPs) stays on the
// CUDA-core kernel.
// 2.5-D LPIPS of a 3-D volume pushes 2 x 128 slices of 128 x 128 through AlexNet per item (src/trainers/reconstruct.py:
// 181-187): ~50 GFLOP in these three layers, 2 ms on the CUDA-core kernel above. They run instead on the tcgen05
// implicit-GEMM conv (conv_gemm.cu) with SPLIT-PRECISION operands - activations and weights as fp16 hi + lo halves, K
// segments [a_hi | a_lo | a_hi] x [w_hi | w_hi | w_lo], fp32 accumulation - which keeps ~22 mantissa bits per product:
// the fp32 semantics of the reference at tensor-core speed. Features stay channels-last hi / lo halves for the distance
// kernel.
__global__ void lpips_split_nhwc_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, int NB,
                                        int C, int npix) {
    // in [NB][C][npix] fp32 -> hi / lo [NB][npix][C] fp16, via a 32 x 32 shared-memory transpose
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, pp = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && pp < npix) ? in[(static_cast<long long>(n) * C + c) * npix + pp] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int pp = p0 + i, c = c0 + threadIdx.x;
        if (pp < npix && c < C) {
            const float v = tile[threadIdx.x][i];
            const __half h = __float2half_rn(v);
            const long long o = (static_cast<long long>(n) * npix + pp) * C + c;
            hi[o] = h;
            lo[o] = __float2half_rn(v - __half2float(h));
        }
    }
}

// fp32 [Cout][Cin][taps] -> fp16 [Cout][3 * taps * Cin]: K = segment * taps Cin + tap * Cin + ci, segments hi | hi | lo
__global__ void lpips_pack_split_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, __half* __restrict__ dst) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    const long long ktot = static_cast<long long>(taps) * Cin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        const long long r = i / Cin;
        const int tap = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        const float v = w[(static_cast<long long>(co) * Cin + ci) * taps + tap];
        const __half h = __float2half_rn(v);
        const long long k = co * 3 * ktot + static_cast<long long>(tap) * Cin + ci;
        dst[k] = h;
        dst[k + ktot] = h;
        dst[k + 2 * ktot] = __float2half_rn(v - __half2float(h));
    }
}

// 3x3 / stride 2 max-pool on channels-last halves: in [NB][H][W][Cs] (C real channels of a row of Cs), out [NB][Ho][Wo][C].
// The maximum of hi + lo is one of the inputs: its two halves are copied, nothing is re-rounded.
__global__ void lpips_maxpool_halves_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, __half* __restrict__ ohi,
                                            __half* __restrict__ olo, int NB, int H, int W, int Cs, int C, int Ho, int Wo) {
    const long long total = static_cast<long long>(NB) * Ho * Wo * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long long r = i / C;
        const int wo = static_cast<int>(r % Wo); r /= Wo;
        const int ho = static_cast<int>(r % Ho);
        const long long n = r / Ho;
        float best = -INFINITY;
        __half bh = __float2half(0.f), bl = __float2half(0.f);
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                const int y = 2 * ho + dy, x = 2 * wo + dx;
                if (y >= H || x >= W) continue;
                const long long o = ((n * H + y) * W + x) * Cs + c;
                const __half a = hi[o], b = lo[o];
                const float v = __half2float(a) + __half2float(b);
                if (v > best) { best = v; bh = a; bl = b; }
            }
        ohi[i] = bh;
        olo[i] = bl;
    }
}

// One CTA per image pair: sum_k mean_p sum_c lin_k[c] * (f0/(|f0|+eps) - f1/(|f1|+eps))^2
__global__ void __launch_bounds__(256) lpips_distance_kernel(LpipsFeat F, int B, float* __restrict__ out) {
    __shared__ float s_red[8];
    __shared__ float s_total;
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_total = 0.f;
    __syncthreads();
    for (int k = 0; k < 5; ++k) {
        const int Ck = (k == 0) ? 64 : (k == 1 ? 192 : (k == 2 ? 384 : 256));
        const int np = F.npix[k];
        const bool halves = F.f[k] == nullptr;  // channels-last hi / lo halves (tensor-core path)
        const float* f0 = halves ? nullptr : F.f[k] + static_cast<long long>(b) * Ck * np;
        const float* f1 = halves ? nullptr : F.f[k] + static_cast<long long>(b + B) * Ck * np;
        const int Cs = F.cstride[k];
        const long long h0 = static_cast<long long>(b) * np * Cs, h1 = static_cast<long long>(b + B) * np * Cs;
        auto feat = [&](int which, int c, int p) -> float {
            if (!halves) return (which ? f1 : f0)[c * np + p];
            const long long o = (which ? h1 : h0) + static_cast<long long>(p) * Cs + c;
            return __half2float(F.fh[k][o]) + __half2float(F.fl[k][o]);
        };
        float layer = 0.f;  // per-warp partial over its pixels
        for (int p = warp; p < np; p += 8) {
            float n0 = 0.f, n1 = 0.f;
            for (int c = lane; c < Ck; c += 32) {
                const float a = feat(0, c, p), d = feat(1, c, p);
                n0 += a * a;
                n1 += d * d;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                n0 += __shfl_xor_sync(0xffffffffu, n0, o);
                n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            }
            const float i0 = 1.0f / (sqrtf(n0) + 1e-10f), i1 = 1.0f / (sqrtf(n1) + 1e-10f);
            float acc = 0.f;
            for (int c = lane; c < Ck; c += 32) {
                const float d = feat(0, c, p) * i0 - feat(1, c, p) * i1;
                acc += F.lin[k][c] * d * d;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            layer += acc;
        }
        if (lane == 0) s_red[warp] = layer;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; ++i) t += s_red[i];
            s_total += t / static_cast<float>(np);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[b] = s_total;
}

// ------------------------------------------------------------------------------------------------ host
Lpips::Lpips() {}
Lpips::~Lpips() {
    if (arena_) cudaFree(arena_);
    if (wsplit_) cudaFree(wsplit_);
}

// rows of the three 3x3 layers from which the tensor-core path is used (below, the CUDA-core kernel's few blocks are faster
// than three conv_prepare calls)
static bool lpips_use_tc(int NB, int h2, int w2) { return h2 * w2 > 1 && static_cast<long long>(NB) * h2 * w2 >= 4096; }

int Lpips::init() {
    size_t total = 0;
    for (int k = 0; k < 5; ++k) total += static_cast<size_t>(kChn[k]) * kCin[k] * kK[k] * kK[k] + kChn[k] + kChn[k];
    if (cudaMalloc(&arena_, total * sizeof(float)) != cudaSuccess) {
        set_error("lpips: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 6;
    }
    float* p = arena_;
    static const char* conv_names[5] = {"net.slice1.0", "net.slice2.3", "net.slice3.6", "net.slice4.8", "net.slice5.10"};
    for (int k = 0; k < 5; ++k) {
        w_[k] = p; p += static_cast<size_t>(kChn[k]) * kCin[k] * kK[k] * kK[k];
        b_[k] = p; p += kChn[k];
        lin_[k] = p; p += kChn[k];
        slots_[std::string(conv_names[k]) + ".weight"] = {w_[k], static_cast<long long>(kChn[k]) * kCin[k] * kK[k] * kK[k], false};
        slots_[std::string(conv_names[k]) + ".bias"] = {b_[k], kChn[k], false};
        slots_["lin" + std::to_string(k) + ".model.1.weight"] = {lin_[k], kChn[k], false};
    }
    return 0;
}

int Lpips::set_param(const char* name, const float* data, long long numel, cudaStream_t stream) {
    auto it = slots_.find(name);
    if (it == slots_.end()) { set_error("lpips: unexpected parameter '%s'", name); return 8; }
    if (it->second.numel != numel) { set_error("lpips: parameter '%s' has %lld elements, expected %lld", name, numel, it->second.numel); return 8; }
    cudaError_t e = cudaMemcpyAsync(it->second.dst, data, numel * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) { set_error("lpips: copy failed: %s", cudaGetErrorString(e)); return 5; }
    it->second.set = true;
    return 0;
}

int Lpips::finalize() {
    for (auto& kv : slots_)
        if (!kv.second.set) { set_error("lpips: parameter '%s' was never set", kv.first.c_str()); return 8; }
    ready_ = true;
    return 0;
}

static void lpips_dims(int H, int W, int (&h)[5], int (&w)[5], int (&ph)[2], int (&pw)[2]) {
    h[0] = (H + 4 - 11) / 4 + 1; w[0] = (W + 4 - 11) / 4 + 1;
    ph[0] = (h[0] - 3) / 2 + 1;  pw[0] = (w[0] - 3) / 2 + 1;   // maxpool
    h[1] = ph[0];                w[1] = pw[0];                 // 5x5 pad 2
    ph[1] = (h[1] - 3) / 2 + 1;  pw[1] = (w[1] - 3) / 2 + 1;
    h[2] = h[3] = h[4] = ph[1];  w[2] = w[3] = w[4] = pw[1];
}

size_t Lpips::workspace_bytes(int B, int H, int W) const {
    int h[5], w[5], ph[2], pw[2];
    lpips_dims(H, W, h, w, ph, pw);
    if (H < 11 - 4 || W < 11 - 4 || h[0] < 3 || w[0] < 3 || h[1] < 3 || w[1] < 3) return 0;
    size_t fl = static_cast<size_t>(2) * B * 3 * H * W;
    for (int k = 0; k < 5; ++k) fl += static_cast<size_t>(2) * B * kChn[k] * h[k] * w[k];
    fl += static_cast<size_t>(2) * B * 64 * ph[0] * pw[0] + static_cast<size_t>(2) * B * 192 * ph[1] * pw[1];
    size_t extra = 0;
    if (lpips_use_tc(2 * B, h[2], w[2]))  // hi / lo halves of the pooled input and of the three feature maps
        extra = (static_cast<size_t>(2) * B * h[2] * w[2] * (192 + 384 + 256 + 256) +
                 static_cast<size_t>(2) * B * h[1] * w[1] * (64 + 256)) * 2 * sizeof(__half) + 32 * 1024;
    return fl * sizeof(float) + 16 * 256 + extra;
}

int Lpips::forward(const float* in0, const float* in1, float* out, int B, int C, int H, int W, bool normalize, void* ws,
                   size_t ws_bytes, cudaStream_t stream) {
    if (!ready_) { set_error("lpips: forward before finalize()"); return 10; }
    if (C != 1 && C != 3) { set_error("lpips: %d input channels unsupported (1 or 3)", C); return 2; }
    const size_t need = workspace_bytes(B, H, W);
    if (!need) { set_error("lpips: %dx%d input is too small for AlexNet (the reference pads 28x28 to 32x32)", H, W); return 2; }
    if (need > ws_bytes) { set_error("lpips: workspace too small (%zu < %zu)", ws_bytes, need); return 9; }
    int h[5], w[5], ph[2], pw[2];
    lpips_dims(H, W, h, w, ph, pw);
    const int NB = 2 * B;
    float* p = static_cast<float*>(ws);
    auto take = [&](size_t n) { float* r = p; p += (n + 63) & ~size_t(63); return r; };
    float* x = take(static_cast<size_t>(NB) * 3 * H * W);
    float* f[5];
    for (int k = 0; k < 5; ++k) f[k] = take(static_cast<size_t>(NB) * kChn[k] * h[k] * w[k]);
    float* p0 = take(static_cast<size_t>(NB) * 64 * ph[0] * pw[0]);
    float* p1 = take(static_cast<size_t>(NB) * 192 * ph[1] * pw[1]);
    const bool use_tc = lpips_use_tc(NB, h[2], w[2]);
    constexpr int kC1Pad = 256;  // conv2's 192 output channels in rows of 256 (two 128-wide N tiles)
    __half *p0_hi = nullptr, *p0_lo = nullptr, *p1_hi = nullptr, *p1_lo = nullptr, *fh[5] = {}, *fl[5] = {};
    if (use_tc) {
        const size_t rows = static_cast<size_t>(NB) * h[2] * w[2];
        __half* hp = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
        auto take_h = [&](size_t n) { __half* r = hp; hp += (n + 511) & ~size_t(511); return r; };
        const size_t rows1 = static_cast<size_t>(NB) * h[1] * w[1];
        p0_hi = take_h(rows1 * 64); p0_lo = take_h(rows1 * 64);
        fh[1] = take_h(rows1 * kC1Pad); fl[1] = take_h(rows1 * kC1Pad);
        p1_hi = take_h(rows * 192); p1_lo = take_h(rows * 192);
        for (int k = 2; k < 5; ++k) { fh[k] = take_h(rows * kChn[k]); fl[k] = take_h(rows * kChn[k]); }
        if (!wsplit_ready_) {  // one-time: split-precision weight matrices of conv2 (rows padded to 256) .. conv5
            size_t total = static_cast<size_t>(kC1Pad) * 3 * 25 * kCin[1];
            for (int k = 2; k < 5; ++k) total += static_cast<size_t>(kChn[k]) * 27 * kCin[k];
            const size_t bytes = total * sizeof(__half) + kC1Pad * sizeof(float);
            if (!wsplit_ && cudaMalloc(&wsplit_, bytes) != cudaSuccess) {
                set_error("lpips: cudaMalloc of the split-precision weights failed");
                return 6;
            }
            cudaMemsetAsync(wsplit_, 0, bytes, stream);
            __half* wp = static_cast<__half*>(wsplit_);
            for (int k = 1; k < 5; ++k) {
                const int taps = kK[k] * kK[k];
                wsplit_k_[k] = wp;
                lpips_pack_split_kernel<<<592, 256, 0, stream>>>(w_[k], kChn[k], kCin[k], taps, wp);
                LP_CHECK("lpips_pack_split");
                wp += static_cast<size_t>(k == 1 ? kC1Pad : kChn[k]) * 3 * taps * kCin[k];
            }
            bias1_pad_ = reinterpret_cast<float*>(wp);
            cudaMemcpyAsync(bias1_pad_, b_[1], kChn[1] * sizeof(float), cudaMemcpyDeviceToDevice, stream);
            wsplit_ready_ = true;
        }
    }

    auto blocks_for = [](long long n) { long long b = (n + 255) / 256; return static_cast<int>(b > 148 * 16 ? 148 * 16 : b); };
    lpips_scale_kernel<<<blocks_for(static_cast<long long>(NB) * 3 * H * W), 256, 0, stream>>>(
        in0, in1, x, B, C, H * W, normalize ? 1 : 0, make_float3(-0.030f, -0.088f, -0.188f),
        make_float3(0.458f, 0.448f, 0.450f));
    LP_CHECK("lpips_scale");
    const float* cur = x;
    int ch = H, cw = W;
    for (int k = 0; k < 5; ++k) {
        if (k == 2 && use_tc) {  // second max-pool on the channels-last halves of conv2's output
            lpips_maxpool_halves_kernel<<<blocks_for(static_cast<long long>(NB) * ph[1] * pw[1] * 192), 256, 0, stream>>>(
                fh[1], fl[1], p1_hi, p1_lo, NB, h[1], w[1], kC1Pad, 192, ph[1], pw[1]);
            LP_CHECK("lpips_maxpool_halves");
            ch = ph[1]; cw = pw[1];
        } else if (k == 1 || k == 2) {
            float* pool = (k == 1) ? p0 : p1;
            const int oh = ph[k - 1], ow = pw[k - 1];
            lpips_maxpool_kernel<<<blocks_for(static_cast<long long>(NB) * kCin[k] * oh * ow), 256, 0, stream>>>(
                cur, pool, static_cast<long long>(NB) * kCin[k], ch, cw, oh, ow);
            LP_CHECK("lpips_maxpool");
            cur = pool; ch = oh; cw = ow;
        }
        const long long M = static_cast<long long>(NB) * h[k] * w[k];
        if (k >= 1 && use_tc) {
            const __half* in_hi = k == 1 ? p0_hi : (k == 2 ? p1_hi : fh[k - 1]);
            const __half* in_lo = k == 1 ? p0_lo : (k == 2 ? p1_lo : fl[k - 1]);
            if (k == 1) {  // the pooled conv1 features (NCHW fp32) -> channels-last halves
                dim3 g((h[1] * w[1] + 31) / 32, (64 + 31) / 32, NB);
                lpips_split_nhwc_kernel<<<g, dim3(32, 8), 0, stream>>>(cur, p0_hi, p0_lo, NB, 64, h[1] * w[1]);
                LP_CHECK("lpips_split_nhwc");
            }
            const int cout = k == 1 ? kC1Pad : kChn[k];
            ConvProblem q{};
            q.spatial_dims = 2; q.N = NB; q.D = 1; q.H = h[k]; q.W = w[k]; q.stride = 1;
            q.n_seg = 3;
            q.seg[0] = {in_hi, kCin[k], kK[k]};
            q.seg[1] = {in_lo, kCin[k], kK[k]};
            q.seg[2] = {in_hi, kCin[k], kK[k]};
            q.weights = wsplit_k_[k]; q.w_rows = cout; q.Cout = cout;
            q.mode = EPI_STORE; q.bias = k == 1 ? bias1_pad_ : b_[k]; q.relu = 1;
            q.out = fh[k]; q.out_lo = fl[k];
            ConvLaunch l;
            int rc = conv_prepare(q, num_sms(), &l);
            if (!rc) rc = conv_launch(l, stream);
            if (rc) return rc;
            cur = nullptr; ch = h[k]; cw = w[k];
            continue;
        }
        dim3 grid(static_cast<unsigned>((M + kLpBM - 1) / kLpBM), static_cast<unsigned>((kChn[k] + kLpBN - 1) / kLpBN));
        if (k == 0)
            lpips_conv_relu_kernel<11, 4, 2><<<grid, 256, 0, stream>>>(cur, w_[k], b_[k], f[k], NB, kCin[k], ch, cw, kChn[k],
                                                                      h[k], w[k]);
        else if (k == 1)
            lpips_conv_relu_kernel<5, 1, 2><<<grid, 256, 0, stream>>>(cur, w_[k], b_[k], f[k], NB, kCin[k], ch, cw, kChn[k],
                                                                     h[k], w[k]);
        else if (ch == 1 && cw == 1)
            lpips_conv_relu_kernel<3, 1, 1, true><<<grid, 256, 0, stream>>>(cur, w_[k], b_[k], f[k], NB, kCin[k], ch, cw,
                                                                           kChn[k], h[k], w[k]);
        else
            lpips_conv_relu_kernel<3, 1, 1><<<grid, 256, 0, stream>>>(cur, w_[k], b_[k], f[k], NB, kCin[k], ch, cw, kChn[k],
                                                                     h[k], w[k]);
        LP_CHECK("lpips_conv");
        cur = f[k]; ch = h[k]; cw = w[k];
    }
    LpipsFeat F;
    for (int k = 0; k < 5; ++k) {
        const bool halves = use_tc && k >= 1;
        F.f[k] = halves ? nullptr : f[k]; F.fh[k] = fh[k]; F.fl[k] = fl[k];
        F.cstride[k] = (use_tc && k == 1) ? kC1Pad : kChn[k];
        F.lin[k] = lin_[k]; F.npix[k] = h[k] * w[k];
    }
    lpips_distance_kernel<<<B, 256, 0, stream>>>(F, B, out);
    LP_CHECK("lpips_distance");
    launches_ += use_tc ? 11 : 9;
    return 0;
}

}  // namespace ddpm
