// VQ-VAE encode / quantise / decode engine. See vqvae.cuh.
#include "vqvae.cuh"

#include <stdlib.h>
#include <string.h>

#include "launch.cuh"

namespace ddpm {

int num_sms();  // api.cu

namespace {

#define VQ_CHECK_LAUNCH(what)                                                             \
    do {                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e_)); return 5; } \
    } while (0)

// fp32 [Cout][Cin][taps] -> fp16 [Cout][ktot], K = tap * Cin + ci (zero beyond taps * Cin: the arena is zero-filled).
// split: the row is three K segments of ktot each, [hi | hi | lo] with hi = fp16(w), lo = fp16(w - hi), multiplying the
// activation segments [hi | lo | hi]: a.w ~ a_hi.w_hi + a_lo.w_hi + a_hi.w_lo (the dropped a_lo.w_lo term is ~2^-22 relative).
__global__ void vq_pack_conv_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, __half* __restrict__ dst,
                                    long long ktot, int split) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    const long long row = split ? 3 * ktot : ktot;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        const long long r = i / Cin;
        const int tap = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        const float v = w[(static_cast<long long>(co) * Cin + ci) * taps + tap];
        const __half hi = __float2half_rn(v);
        const long long k = co * row + static_cast<long long>(tap) * Cin + ci;
        dst[k] = hi;
        if (split) {
            dst[k + ktot] = hi;
            dst[k + 2 * ktot] = __float2half_rn(v - __half2float(hi));
        }
    }
}

// kernel tap of a stride-2, 4-tap, pad-1 transposed conv that connects output phase p (y = 2m + p) with the low-resolution
// input m + p - 1 + a: t = 3 - p - 2a
__device__ __forceinline__ int tconv_tap(int p, int a) { return 3 - p - 2 * a; }

// ConvTranspose weight fp32 [Cin][Cout][4^d] -> fp16 sub-pixel phase matrix [2^d * Cout][2^d * Cin]
// (row = phase * Cout + co, column = tap * Cin + ci; phase / tap bits: 0 -> w, 1 -> h, 2 -> d, as conv_gemm.cu enumerates)
__global__ void vq_pack_tconv_phase_kernel(const float* __restrict__ w, int Cin, int Cout, int dims, __half* __restrict__ dst) {
    const int np = 1 << dims;
    const long long total = static_cast<long long>(np) * Cout * np * Cin;
    const int kvol = dims == 3 ? 64 : 16;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        long long r = i / Cin;
        const int tap = static_cast<int>(r % np); r /= np;
        const int co = static_cast<int>(r % Cout);
        const int phase = static_cast<int>(r / Cout);
        const int kw = tconv_tap(phase & 1, tap & 1), kh = tconv_tap((phase >> 1) & 1, (tap >> 1) & 1);
        const int kd = dims == 3 ? tconv_tap((phase >> 2) & 1, (tap >> 2) & 1) : 0;
        const int kidx = dims == 3 ? (kd * 4 + kh) * 4 + kw : kh * 4 + kw;
        dst[i] = __float2half_rn(w[(static_cast<long long>(ci) * Cout + co) * kvol + kidx]);
    }
}

// last transposed conv (few output channels): fp16 [(phase * 2^d + tap) * Cout + co][Cin] tap-product weights
__global__ void vq_pack_tconv_taps_kernel(const float* __restrict__ w, int Cin, int Cout, int dims, __half* __restrict__ dst) {
    const int np = 1 << dims;
    const long long total = static_cast<long long>(np) * np * Cout * Cin;
    const int kvol = dims == 3 ? 64 : 16;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        long long r = i / Cin;
        const int co = static_cast<int>(r % Cout); r /= Cout;
        const int tap = static_cast<int>(r % np);
        const int phase = static_cast<int>(r / np);
        const int kw = tconv_tap(phase & 1, tap & 1), kh = tconv_tap((phase >> 1) & 1, (tap >> 1) & 1);
        const int kd = dims == 3 ? tconv_tap((phase >> 2) & 1, (tap >> 2) & 1) : 0;
        const int kidx = dims == 3 ? (kd * 4 + kh) * 4 + kw : kh * 4 + kw;
        dst[i] = __float2half_rn(w[(static_cast<long long>(ci) * Cout + co) * kvol + kidx]);
    }
}

__global__ void vq_code_sq_kernel(const float* __restrict__ e, int K, int E, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float s = 0.f;
    for (int j = 0; j < E; ++j) s = fmaf(e[static_cast<size_t>(k) * E + j], e[static_cast<size_t>(k) * E + j], s);
    out[k] = s;
}

// first encoder conv (k4, s2, p1) over a few input channels: x fp32 [N][Cin][D][H][W] -> im2col rows fp16 [M][Kpad],
// M = N * Do * Ho * Wo output pixels, K = tap * Cin + ci
// (out_lo: the lo halves of the same values for the split-precision encoder, or null)
__global__ void vq_im2col_first_kernel(const float* __restrict__ x, __half* __restrict__ out, __half* __restrict__ out_lo, int N,
                                       int Cin, int D, int H, int W, int dims, int Kpad) {
    const int Do = dims == 3 ? D / 2 : 1, Ho = H / 2, Wo = W / 2;
    const long long total = static_cast<long long>(N) * Do * Ho * Wo * Kpad;
    const int taps = dims == 3 ? 64 : 16;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kpad);
        long long m = i / Kpad;
        float v = 0.f;
        if (k < taps * Cin) {
            const int ci = k % Cin, tap = k / Cin;
            const int kw = tap & 3, kh = (tap >> 2) & 3, kd = tap >> 4;
            const int wo = static_cast<int>(m % Wo); m /= Wo;
            const int ho = static_cast<int>(m % Ho); m /= Ho;
            const int d_o = static_cast<int>(m % Do);
            const int n = static_cast<int>(m / Do);
            const int wi = 2 * wo - 1 + kw, hi = 2 * ho - 1 + kh, di = dims == 3 ? 2 * d_o - 1 + kd : 0;
            if (wi >= 0 && wi < W && hi >= 0 && hi < H && di >= 0 && di < D)
                v = x[(((static_cast<long long>(n) * Cin + ci) * D + di) * H + hi) * W + wi];
        }
        const __half hv = __float2half_rn(v);
        out[i] = hv;
        if (out_lo) out_lo[i] = __float2half_rn(v - __half2float(hv));
    }
}

// last transposed conv: sum the 2^d tap products of every output voxel. taps fp32 [M_low][cols], column
// (phase * 2^d + tap) * Cout + co; image fp32 [N][Cout][D][H][W] (D, H, W = output extents)
__global__ void vq_gather_kernel(const float* __restrict__ t, const float* __restrict__ bias, float* __restrict__ img, int N,
                                 int Cout, int D, int H, int W, int dims, int cols) {
    const long long total = static_cast<long long>(N) * Cout * D * H * W;
    const int Dl = dims == 3 ? D / 2 : 1, Hl = H / 2, Wl = W / 2;
    const int np = 1 << dims;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int x = static_cast<int>(r % W); r /= W;
        const int y = static_cast<int>(r % H); r /= H;
        const int z = static_cast<int>(r % D); r /= D;
        const int co = static_cast<int>(r % Cout);
        const int n = static_cast<int>(r / Cout);
        const int pw = x & 1, ph = y & 1, pd = dims == 3 ? (z & 1) : 0;
        const int phase = pw | (ph << 1) | (pd << 2);
        float acc = bias ? bias[co] : 0.f;
        for (int tap = 0; tap < np; ++tap) {
            const int sw = (x >> 1) + pw - 1 + (tap & 1), sh = (y >> 1) + ph - 1 + ((tap >> 1) & 1);
            const int sd = dims == 3 ? (z >> 1) + pd - 1 + ((tap >> 2) & 1) : 0;
            if (sw < 0 || sw >= Wl || sh < 0 || sh >= Hl || sd < 0 || sd >= Dl) continue;
            const long long row = ((static_cast<long long>(n) * Dl + sd) * Hl + sh) * Wl + sw;
            acc += t[row * cols + (phase * np + tap) * Cout + co];
        }
        img[i] = acc;
    }
}

// Nearest codebook row per latent vector (EMAQuantizer.quantize + embed + the straight-through form x + (q - x)).
// fp32 throughout: d = (|x|^2 + |e|^2) - 2 x.e in the reference's operation order, first minimum wins (torch.max(-d)).
// 256 threads = 32 rows x 8 code lanes; codes are streamed through shared memory in chunks of 64.
// IN: 1 = z is fp32 [N][E][S] (the UNet's sample); 0 = fp16 [M][E] (the fp16 encoder's channels-last output);
// 2 = fp32 [M][E] (the split-precision encoder's output).
constexpr int kQRows = 32, kQCodes = 64;
template <int IN>
__global__ void __launch_bounds__(256) vq_quantize_kernel(const void* __restrict__ zin, const float* __restrict__ codebook,
                                                          const float* __restrict__ code_sq, int K, int E, long long M,
                                                          long long S, const int* __restrict__ idx_in, int* __restrict__ idx_out,
                                                          float* __restrict__ out_nchw, __half* __restrict__ out_half) {
    extern __shared__ float sm[];
    float* sx = sm;                          // [kQRows][E + 1]
    float* se = sm + kQRows * (E + 1);       // [kQCodes][E + 1]
    __shared__ int s_best[kQRows];
    const long long m0 = static_cast<long long>(blockIdx.x) * kQRows;
    const int tid = threadIdx.x;
    for (int i = tid; i < kQRows * E; i += 256) {
        const int r = IN == 1 ? i % kQRows : i / E, k = IN == 1 ? i / kQRows : i % E;
        const long long m = m0 + r;
        float v = 0.f;
        if (m < M && zin) {
            if (IN == 1) v = static_cast<const float*>(zin)[((m / S) * E + k) * S + (m % S)];
            else if (IN == 2) v = static_cast<const float*>(zin)[m * E + k];
            else v = __half2float(static_cast<const __half*>(zin)[m * E + k]);
        }
        sx[r * (E + 1) + k] = v;
    }
    __syncthreads();
    const int r = tid >> 3, cl = tid & 7;
    if (!idx_in) {
        float xx = 0.f;
        for (int k = 0; k < E; ++k) xx = fmaf(sx[r * (E + 1) + k], sx[r * (E + 1) + k], xx);
        float best = INFINITY;
        int best_i = 0;
        for (int c0 = 0; c0 < K; c0 += kQCodes) {
            __syncthreads();
            for (int i = tid; i < kQCodes * E; i += 256) {
                const int c = i / E, k = i % E;
                se[c * (E + 1) + k] = c0 + c < K ? codebook[static_cast<size_t>(c0 + c) * E + k] : 0.f;
            }
            __syncthreads();
#pragma unroll 1
            for (int j = 0; j < kQCodes / 8; ++j) {
                const int c = cl * (kQCodes / 8) + j;
                if (c0 + c >= K) break;
                float dot = 0.f;
                const float* xr = sx + r * (E + 1);
                const float* er = se + c * (E + 1);
                for (int k = 0; k < E; ++k) dot = fmaf(xr[k], er[k], dot);
                const float d = __fsub_rn(__fadd_rn(xx, code_sq[c0 + c]), __fmul_rn(2.f, dot));
                if (d < best) { best = d; best_i = c0 + c; }  // ascending scan: the first minimum wins
            }
        }
        // the 8 code lanes of a row: smallest distance, then smallest index
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
            if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
        }
        if (cl == 0) s_best[r] = best_i;
    } else if (cl == 0) {
        s_best[r] = m0 + r < M ? idx_in[m0 + r] : 0;
    }
    __syncthreads();
    if (idx_out && tid < kQRows && m0 + tid < M) idx_out[m0 + tid] = s_best[tid];
    for (int i = tid; i < kQRows * E; i += 256) {
        // NCHW output: consecutive threads -> consecutive rows (contiguous along S); half output: consecutive channels
        {
            const int rr = i % kQRows, k = i / kQRows;
            const long long m = m0 + rr;
            if (out_nchw && m < M) {
                const float q = codebook[static_cast<size_t>(s_best[rr]) * E + k];
                const float x = sx[rr * (E + 1) + k];
                out_nchw[((m / S) * E + k) * S + (m % S)] = idx_in ? q : __fadd_rn(x, __fsub_rn(q, x));
            }
        }
        {
            const int rr = i / E, k = i % E;
            const long long m = m0 + rr;
            if (out_half && m < M) {
                const float q = codebook[static_cast<size_t>(s_best[rr]) * E + k];
                const float x = sx[rr * (E + 1) + k];
                out_half[m * E + k] = __float2half_rn(idx_in ? q : __fadd_rn(x, __fsub_rn(q, x)));
            }
        }
    }
}

int grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// ------------------------------------------------------------------------------------------------ construction
VqVae::~VqVae() {
    if (f32_arena_) cudaFree(f32_arena_);
    if (f16_arena_) cudaFree(f16_arena_);
}

template <typename T>
T* VqVae::arena(size_t count, bool half_arena) {
    const size_t aligned = (count + 127) & ~size_t(127);
    if (half_arena) {
        T* p = sizing_ ? nullptr : reinterpret_cast<T*>(f16_arena_ + f16_used_);
        f16_used_ += aligned;
        return p;
    }
    T* p = sizing_ ? nullptr : reinterpret_cast<T*>(f32_arena_ + f32_used_);
    f32_used_ += aligned;
    return p;
}

void VqVae::slot(const std::string& name, int kind, void* dst, long long numel, int Cout, int Cin, int taps, long long ktot) {
    if (sizing_) return;
    Slot s{};
    s.kind = kind; s.dst = dst; s.numel = numel; s.Cout = Cout; s.Cin = Cin; s.taps = taps; s.ktot = ktot; s.set = false;
    slots_[name] = s;
}

VqVae::Res VqVae::make_res(const std::string& prefix, int C, int R, bool split) {
    const int taps = cfg_.spatial_dims == 3 ? 27 : 9;
    const size_t mult = split ? 3 : 1;
    const int kind = split ? 4 : 1;
    Res r{};
    r.C = C; r.R = R;
    r.w1 = arena<__half>(mult * R * taps * C, true);
    r.w2 = arena<__half>(mult * C * taps * R, true);
    r.b1 = arena<float>(R, false);
    r.b2 = arena<float>(C, false);
    slot(prefix + ".conv1.conv.weight", kind, r.w1, static_cast<long long>(R) * C * taps, R, C, taps, static_cast<long long>(taps) * C);
    slot(prefix + ".conv1.conv.bias", 0, r.b1, R);
    slot(prefix + ".conv2.conv.weight", kind, r.w2, static_cast<long long>(C) * R * taps, C, R, taps, static_cast<long long>(taps) * R);
    slot(prefix + ".conv2.conv.bias", 0, r.b2, C);
    return r;
}

int VqVae::init() {
    const VqVaeConfig& c = cfg_;
    const int sd = c.spatial_dims;
    if (sd != 2 && sd != 3) { set_error("vqvae: spatial_dims=%d unsupported", sd); return 2; }
    if (c.num_levels < 1 || c.num_levels > kVqMaxLevels) { set_error("vqvae: %d levels unsupported", c.num_levels); return 2; }
    for (int i = 0; i < c.num_levels; ++i) {
        if (c.num_channels[i] % 128 != 0 || c.num_res_channels[i] % 128 != 0) {
            set_error("vqvae: num_channels / num_res_channels must be multiples of 128 (tcgen05 N tiles)");
            return 2;
        }
    }
    if (c.embedding_dim % 128 != 0 || c.embedding_dim > 512) { set_error("vqvae: embedding_dim must be 128, 256, 384 or 512"); return 2; }
    const int np = 1 << sd, k4 = sd == 3 ? 64 : 16, k3 = sd == 3 ? 27 : 9;
    if (c.in_channels * k4 > 512) { set_error("vqvae: in_channels=%d unsupported (image-side conv)", c.in_channels); return 2; }
    if (np * np * c.out_channels > 128) { set_error("vqvae: out_channels=%d unsupported (image-side transposed conv)", c.out_channels); return 2; }
    first_kpad_ = ((c.in_channels * k4 + 63) / 64) * 64;
    last_cols_ = 128;
    const bool split = c.precise_encode != 0;
    const size_t emult = split ? 3 : 1;
    const int ekind = split ? 4 : 1;
    for (int pass = 0; pass < 2; ++pass) {
        sizing_ = pass == 0;
        f32_used_ = f16_used_ = 0;
        enc_.clear();
        dec_.clear();
        slots_.clear();
        if (!sizing_) {
            if (cudaMalloc(&f32_arena_, f32_count_ * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&f16_arena_, f16_count_ * sizeof(__half)) != cudaSuccess) {
                set_error("vqvae: cudaMalloc of the weight arenas failed");
                return 6;
            }
            cudaMemset(f32_arena_, 0, f32_count_ * sizeof(float));
            cudaMemset(f16_arena_, 0, f16_count_ * sizeof(__half));
        }
        int b = 0;
        for (int i = 0; i < c.num_levels; ++i) {
            EncLevel L{};
            L.Cin = i == 0 ? c.in_channels : c.num_channels[i - 1];
            L.Cout = c.num_channels[i];
            const long long ktot = i == 0 ? first_kpad_ : static_cast<long long>(k4) * L.Cin;
            L.w = arena<__half>(emult * L.Cout * ktot, true);
            L.b = arena<float>(L.Cout, false);
            const std::string pre = "encoder.blocks." + std::to_string(b++);
            slot(pre + ".conv.weight", ekind, L.w, static_cast<long long>(L.Cout) * L.Cin * k4, L.Cout, L.Cin, k4, ktot);
            slot(pre + ".conv.bias", 0, L.b, L.Cout);
            for (int j = 0; j < c.num_res_layers; ++j)
                L.res.push_back(make_res("encoder.blocks." + std::to_string(b++), L.Cout, c.num_res_channels[i], split));
            enc_.push_back(std::move(L));
        }
        {
            const int Cl = c.num_channels[c.num_levels - 1];
            enc_out_w_ = arena<__half>(emult * c.embedding_dim * k3 * Cl, true);
            enc_out_b_ = arena<float>(c.embedding_dim, false);
            const std::string pre = "encoder.blocks." + std::to_string(b);
            slot(pre + ".conv.weight", ekind, enc_out_w_, static_cast<long long>(c.embedding_dim) * Cl * k3, c.embedding_dim, Cl, k3,
                 static_cast<long long>(k3) * Cl);
            slot(pre + ".conv.bias", 0, enc_out_b_, c.embedding_dim);
            dec_in_w_ = arena<__half>(static_cast<size_t>(Cl) * k3 * c.embedding_dim, true);
            dec_in_b_ = arena<float>(Cl, false);
            slot("decoder.blocks.0.conv.weight", 1, dec_in_w_, static_cast<long long>(Cl) * c.embedding_dim * k3, Cl,
                 c.embedding_dim, k3, static_cast<long long>(k3) * c.embedding_dim);
            slot("decoder.blocks.0.conv.bias", 0, dec_in_b_, Cl);
        }
        b = 1;
        for (int i = 0; i < c.num_levels; ++i) {
            const int lvl = c.num_levels - 1 - i;
            DecLevel L{};
            L.Cin = c.num_channels[lvl];
            L.last = i == c.num_levels - 1;
            L.Cout = L.last ? c.out_channels : c.num_channels[lvl - 1];
            for (int j = 0; j < c.num_res_layers; ++j)
                L.res.push_back(make_res("decoder.blocks." + std::to_string(b++), L.Cin, c.num_res_channels[lvl], false));
            const std::string pre = "decoder.blocks." + std::to_string(b++);
            if (L.last) {
                L.w = arena<__half>(static_cast<size_t>(last_cols_) * L.Cin, true);
                slot(pre + ".conv.weight", 3, L.w, static_cast<long long>(L.Cin) * L.Cout * k4, L.Cout, L.Cin, k4, 0);
            } else {
                L.w = arena<__half>(static_cast<size_t>(np) * L.Cout * np * L.Cin, true);
                slot(pre + ".conv.weight", 2, L.w, static_cast<long long>(L.Cin) * L.Cout * k4, L.Cout, L.Cin, k4, 0);
            }
            L.b = arena<float>(L.Cout, false);
            slot(pre + ".conv.bias", 0, L.b, L.Cout);
            dec_.push_back(std::move(L));
        }
        codebook_ = arena<float>(static_cast<size_t>(c.num_embeddings) * c.embedding_dim, false);
        code_sq_ = arena<float>(c.num_embeddings, false);
        slot("quantizer.quantizer.embedding.weight", 0, codebook_, static_cast<long long>(c.num_embeddings) * c.embedding_dim);
        if (sizing_) { f32_count_ = f32_used_; f16_count_ = f16_used_; }
    }
    return 0;
}

int VqVae::set_param(const char* name, const float* data, long long numel, cudaStream_t stream) {
    auto it = slots_.find(name);
    if (it == slots_.end()) { set_error("vqvae: unexpected parameter '%s'", name); return 8; }
    Slot& s = it->second;
    if (numel != s.numel) { set_error("vqvae: parameter '%s' has %lld elements, expected %lld", name, numel, s.numel); return 8; }
    const int sd = cfg_.spatial_dims;
    switch (s.kind) {
        case 0:
            if (cudaMemcpyAsync(s.dst, data, numel * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess) {
                set_error("vqvae: copy of '%s' failed", name);
                return 5;
            }
            break;
        case 1:
        case 4:
            vq_pack_conv_kernel<<<grid_for(numel), 256, 0, stream>>>(data, s.Cout, s.Cin, s.taps, static_cast<__half*>(s.dst), s.ktot,
                                                                     s.kind == 4 ? 1 : 0);
            break;
        case 2:
            vq_pack_tconv_phase_kernel<<<grid_for(numel), 256, 0, stream>>>(data, s.Cin, s.Cout, sd, static_cast<__half*>(s.dst));
            break;
        default:
            vq_pack_tconv_taps_kernel<<<grid_for(numel), 256, 0, stream>>>(data, s.Cin, s.Cout, sd, static_cast<__half*>(s.dst));
            break;
    }
    VQ_CHECK_LAUNCH("vqvae: weight upload");
    s.set = true;
    finalized_ = false;
    return 0;
}

int VqVae::finalize(cudaStream_t stream) {
    for (auto& kv : slots_)
        if (!kv.second.set) { set_error("vqvae: parameter '%s' was never set", kv.first.c_str()); return 8; }
    vq_code_sq_kernel<<<(cfg_.num_embeddings + 255) / 256, 256, 0, stream>>>(codebook_, cfg_.num_embeddings, cfg_.embedding_dim, code_sq_);
    VQ_CHECK_LAUNCH("vqvae: finalize");
    finalized_ = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ plans
int VqVae::build(Plan& plan, bool decode, int N, int D, int H, int W, void* ws, size_t ws_bytes, bool dry, size_t* need,
                 const float* io_in, float* io_out, int* indices, const int* indices_in) const {
    const VqVaeConfig& c = cfg_;
    const int sd = c.spatial_dims, L = c.num_levels, E = c.embedding_dim;
    if (sd == 2 && D != 1) { set_error("vqvae: 2-D model needs D == 1"); return 2; }
    const int f = 1 << L;
    if (H % f || W % f || (sd == 3 && D % f)) { set_error("vqvae: image extents %dx%dx%d must be divisible by 2^%d", D, H, W, L); return 2; }
    auto dim = [&](int v, int lvl) { return v >> lvl; };  // extent after `lvl` halvings
    auto rows = [&](int lvl) { return static_cast<long long>(N) * (sd == 3 ? dim(D, lvl) : 1) * dim(H, lvl) * dim(W, lvl); };
    // workspace layout: three rotating activation buffers + side buffers
    size_t act = 0;
    for (int i = 0; i < L; ++i) {
        const size_t ch = static_cast<size_t>(c.num_channels[i] > c.num_res_channels[i] ? c.num_channels[i] : c.num_res_channels[i]);
        const size_t a = static_cast<size_t>(rows(i + 1)) * ch;
        if (a > act) act = a;
    }
    const bool split = !decode && c.precise_encode != 0;
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 1023) & ~size_t(1023); const size_t o = off; off += bytes; return o; };
    const size_t oA = take(act * 2), oB = take(act * 2), oC = take(act * 2);
    const size_t oCol = take(static_cast<size_t>(rows(1)) * first_kpad_ * 2);       // encode: im2col of the image
    const size_t oZ = take(static_cast<size_t>(rows(L)) * E * 2);                  // latent, channels-last fp16
    const size_t oTap = take(static_cast<size_t>(rows(1)) * last_cols_ * 4);        // decode: fp32 tap products
    // split-precision encode: lo halves of the three activation buffers and of the im2col block, fp32 latent rows
    const bool split_cfg = c.precise_encode != 0;
    const size_t oAl = split_cfg ? take(act * 2) : 0, oBl = split_cfg ? take(act * 2) : 0, oCl = split_cfg ? take(act * 2) : 0;
    const size_t oColL = split_cfg ? take(static_cast<size_t>(rows(1)) * first_kpad_ * 2) : 0;
    const size_t oZf = split_cfg ? take(static_cast<size_t>(rows(L)) * E * 4) : 0;
    if (need) *need = off + 1024;
    if (dry) return 0;
    if (off > ws_bytes) { set_error("vqvae: workspace too small (%zu < %zu)", ws_bytes, off); return 9; }
    uint8_t* base = static_cast<uint8_t*>(ws);
    __half* buf[3] = {reinterpret_cast<__half*>(base + oA), reinterpret_cast<__half*>(base + oB), reinterpret_cast<__half*>(base + oC)};
    __half* buf_lo[3] = {nullptr, nullptr, nullptr};
    if (split) {
        buf_lo[0] = reinterpret_cast<__half*>(base + oAl);
        buf_lo[1] = reinterpret_cast<__half*>(base + oBl);
        buf_lo[2] = reinterpret_cast<__half*>(base + oCl);
    }
    __half* col = reinterpret_cast<__half*>(base + oCol);
    __half* col_lo = split ? reinterpret_cast<__half*>(base + oColL) : nullptr;
    __half* zq = reinterpret_cast<__half*>(base + oZ);
    float* zf = split ? reinterpret_cast<float*>(base + oZf) : nullptr;
    float* taps = reinterpret_cast<float*>(base + oTap);
    const int sms = num_sms();
    int rc = 0;
    plan.ops.clear();
    const char* vh = getenv("DDPM_VQ_HALO");  // tests: 0 = every conv on the im2col-tile kernel (read at plan time)
    const bool halo_off = vh && atoi(vh) == 0;
    auto gemm = [&](ConvProblem q) {
        Op op{};
        // stride-1 3x3(x3) convs (residual units, the latent-side convs): the halo-tile kernel stages each input tile once
        // for all of its taps where it supports the geometry (2-D; 3-D slabs of 8 x 8 or larger); no GroupNorm here
        bool all3 = q.stride == 1 && !q.upsample2 && q.mode == EPI_STORE;
        for (int i = 0; i < q.n_seg; ++i) all3 = all3 && q.seg[i].ksize == 3;
        if (all3 && !halo_off && conv_halo_supported(q) && conv_halo_prepare(q, nullptr, 0, sms, &op.halo) == 0) {
            op.type = Op::HALO;  // (a geometry the kernel cannot stage - too many K stages per item - takes the GEMM path)
        } else {
            op.type = Op::GEMM;
            int r = conv_prepare(q, sms, &op.conv);
            if (r && !rc) rc = r;
        }
        plan.ops.push_back(op);
    };
    // in_lo != null: split-precision operands, K segments [in | in_lo | in] against weight rows [hi | hi | lo]
    auto conv = [&](const __half* in, const __half* in_lo, int lvl, int cin, int ksize, int stride, const __half* w, int w_rows,
                    const float* bias, int cout, const __half* residual, const __half* residual_lo, void* out, void* out_lo,
                    bool relu, int mode = EPI_STORE, bool up = false) {
        ConvProblem q{};
        q.spatial_dims = sd;
        q.N = N; q.D = sd == 3 ? dim(D, lvl) : 1; q.H = dim(H, lvl); q.W = dim(W, lvl);
        q.stride = stride;
        q.n_seg = in_lo ? 3 : 1;
        q.seg[0] = {in, cin, ksize};
        if (in_lo) {
            q.seg[1] = {in_lo, cin, ksize};
            q.seg[2] = {in, cin, ksize};
        }
        q.weights = w; q.w_rows = w_rows; q.Cout = cout;
        q.mode = mode;
        q.bias = bias; q.residual = residual; q.residual_lo = residual_lo; q.out = out; q.out_lo = out_lo;
        q.relu = relu ? 1 : 0;
        q.pad = ksize == 4 ? 1 : 0;
        q.upsample2 = up ? 1 : 0;
        gemm(q);
    };
    int cur = 0;  // index of the buffer holding the current activation
    auto res_unit = [&](const Res& r, int lvl) {
        const int h = (cur + 1) % 3, o = (cur + 2) % 3;
        conv(buf[cur], buf_lo[cur], lvl, r.C, 3, 1, r.w1, r.R, r.b1, r.R, nullptr, nullptr, buf[h], buf_lo[h], true);
        conv(buf[h], buf_lo[h], lvl, r.R, 3, 1, r.w2, r.C, r.b2, r.C, buf[cur], buf_lo[cur], buf[o], buf_lo[o], true);
        cur = o;
    };
    if (!decode) {
        {   // image-side conv: im2col (K = 4^d * Cin padded to 64) + 1x1 GEMM + ReLU
            Op op{};
            op.type = Op::IM2COL;
            op.src = io_in; op.dst = col; op.dst2 = col_lo; op.N = N; op.C = c.in_channels; op.D = D; op.H = H; op.W = W;
            op.K = first_kpad_;
            plan.ops.push_back(op);
            conv(col, col_lo, 1, first_kpad_, 1, 1, enc_[0].w, enc_[0].Cout, enc_[0].b, enc_[0].Cout, nullptr, nullptr, buf[cur],
                 buf_lo[cur], true);
        }
        for (int i = 0; i < L; ++i) {
            if (i > 0) {
                const int o = (cur + 1) % 3;
                conv(buf[cur], buf_lo[cur], i, enc_[i].Cin, 4, 2, enc_[i].w, enc_[i].Cout, enc_[i].b, enc_[i].Cout, nullptr, nullptr,
                     buf[o], buf_lo[o], true);
                cur = o;
            }
            for (const Res& r : enc_[i].res) res_unit(r, i + 1);
        }
        Op op{};
        if (split) {  // fp32 latent rows straight from the accumulators
            conv(buf[cur], buf_lo[cur], L, c.num_channels[L - 1], 3, 1, enc_out_w_, E, enc_out_b_, E, nullptr, nullptr, zf, nullptr,
                 false, EPI_STORE_F32);
            op.type = Op::QUANT_ROWS_F32;
            op.src = zf;
        } else {
            conv(buf[cur], nullptr, L, c.num_channels[L - 1], 3, 1, enc_out_w_, E, enc_out_b_, E, nullptr, nullptr, zq, nullptr, false);
            op.type = Op::QUANT_HALF;
            op.src = zq;
        }
        op.dst = io_out; op.dst2 = nullptr; op.idx = indices; op.rows = rows(L);
        op.N = N; op.C = E;
        plan.ops.push_back(op);
    } else {
        {
            Op op{};
            op.type = Op::QUANT_NCHW;
            op.src = io_in; op.dst = nullptr; op.dst2 = zq; op.idx = indices; op.rows = rows(L);
            op.N = N; op.C = E;
            op.K = indices_in ? 1 : 0;
            op.dst = const_cast<int*>(indices_in);  // QUANT_NCHW: dst carries the optional input indices
            plan.ops.push_back(op);
        }
        conv(zq, nullptr, L, E, 3, 1, dec_in_w_, c.num_channels[L - 1], dec_in_b_, c.num_channels[L - 1], nullptr, nullptr, buf[cur],
             nullptr, false);
        for (int i = 0; i < L; ++i) {
            const int lvl = L - i;  // halvings of the tensor this level works on
            const DecLevel& Lv = dec_[i];
            for (const Res& r : Lv.res) res_unit(r, lvl);
            if (!Lv.last) {
                const int o = (cur + 1) % 3;
                conv(buf[cur], nullptr, lvl, Lv.Cin, 2, 1, Lv.w, (1 << sd) * Lv.Cout, Lv.b, Lv.Cout, nullptr, nullptr, buf[o], nullptr,
                     true, EPI_STORE, true);
                cur = o;
            } else {
                conv(buf[cur], nullptr, lvl, Lv.Cin, 1, 1, Lv.w, last_cols_, nullptr, last_cols_, nullptr, nullptr, taps, nullptr, false,
                     EPI_STORE_F32);
                Op op{};
                op.type = Op::GATHER;
                op.src = taps; op.dst = io_out; op.dst2 = Lv.b; op.N = N; op.C = Lv.Cout; op.D = D; op.H = H; op.W = W; op.K = last_cols_;
                plan.ops.push_back(op);
            }
        }
    }
    return rc;
}

size_t VqVae::workspace_bytes(int N, int D, int H, int W) const {
    Plan p;
    size_t need = 0;
    if (build(p, false, N, D, H, W, nullptr, 0, true, &need, nullptr, nullptr, nullptr, nullptr)) return 0;
    return need;
}

int VqVae::run(const Plan& plan, cudaStream_t stream) {
    const VqVaeConfig& c = cfg_;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t qsmem = static_cast<size_t>(kQRows + kQCodes) * (c.embedding_dim + 1) * sizeof(float);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(vq_quantize_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(vq_quantize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(vq_quantize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            set_error("vqvae: cudaFuncSetAttribute failed");
            return 4;
        }
        attr_set[dev] = true;
    }
    for (const Op& op : plan.ops) {
        int rc = 0;
        switch (op.type) {
            case Op::IM2COL: {
                const long long total = static_cast<long long>(op.N) * (c.spatial_dims == 3 ? op.D / 2 : 1) * (op.H / 2) * (op.W / 2) * op.K;
                vq_im2col_first_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const float*>(op.src), static_cast<__half*>(op.dst),
                                                                             static_cast<__half*>(op.dst2), op.N, op.C, op.D, op.H, op.W,
                                                                             c.spatial_dims, op.K);
                VQ_CHECK_LAUNCH("vqvae: im2col");
                break;
            }
            case Op::GEMM:
                rc = conv_launch(op.conv, stream);
                break;
            case Op::HALO:
                rc = conv_halo_launch(op.halo, stream);
                break;
            case Op::QUANT_HALF:
            case Op::QUANT_ROWS_F32: {
                const long long S = op.rows / op.N;
                const int blocks = static_cast<int>((op.rows + kQRows - 1) / kQRows);
                if (op.type == Op::QUANT_HALF)
                    vq_quantize_kernel<0><<<blocks, 256, qsmem, stream>>>(op.src, codebook_, code_sq_, c.num_embeddings, c.embedding_dim,
                                                                          op.rows, S, nullptr, op.idx, static_cast<float*>(op.dst), nullptr);
                else
                    vq_quantize_kernel<2><<<blocks, 256, qsmem, stream>>>(op.src, codebook_, code_sq_, c.num_embeddings, c.embedding_dim,
                                                                          op.rows, S, nullptr, op.idx, static_cast<float*>(op.dst), nullptr);
                VQ_CHECK_LAUNCH("vqvae: quantize");
                break;
            }
            case Op::QUANT_NCHW: {
                const long long S = op.rows / op.N;
                vq_quantize_kernel<1><<<static_cast<int>((op.rows + kQRows - 1) / kQRows), 256, qsmem, stream>>>(
                    op.src, codebook_, code_sq_, c.num_embeddings, c.embedding_dim, op.rows, S, static_cast<const int*>(op.dst),
                    op.idx, nullptr, static_cast<__half*>(op.dst2));
                VQ_CHECK_LAUNCH("vqvae: quantize");
                break;
            }
            case Op::GATHER: {
                const long long total = static_cast<long long>(op.N) * op.C * op.D * op.H * op.W;
                vq_gather_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const float*>(op.src), static_cast<const float*>(op.dst2),
                                                                      static_cast<float*>(op.dst), op.N, op.C, op.D, op.H, op.W,
                                                                      c.spatial_dims, op.K);
                VQ_CHECK_LAUNCH("vqvae: gather");
                break;
            }
        }
        if (rc) return rc;
        ++launches_;
    }
    return 0;
}

int VqVae::encode(const float* x, float* latent, int* indices, int N, int D, int H, int W, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
    if (!finalized_) { set_error("vqvae: encode before finalize()"); return 10; }
    Plan plan;
    int rc = build(plan, false, N, D, H, W, ws, ws_bytes, false, nullptr, x, latent, indices, nullptr);
    if (rc) return rc;
    return run(plan, stream);
}

int VqVae::decode(const float* z, const int* indices_in, float* image, int* indices_out, int N, int D, int H, int W, void* ws,
                  size_t ws_bytes, cudaStream_t stream) {
    if (!finalized_) { set_error("vqvae: decode before finalize()"); return 10; }
    if (!z && !indices_in) { set_error("vqvae: decode needs a latent or indices"); return 2; }
    Plan plan;
    int rc = build(plan, true, N, D, H, W, ws, ws_bytes, false, nullptr, z, image, indices_out, indices_in);
    if (rc) return rc;
    return run(plan, stream);
}

}  // namespace ddpm
