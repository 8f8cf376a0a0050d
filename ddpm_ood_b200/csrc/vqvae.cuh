// VQ-VAE stage-1 model of the latent-diffusion path: monai-generative's `VQVAE` as the reference
// builds it from vqvae_config.json (src/trainers/base.py:44-61) and calls it once per batch (`encode_stage_2_inputs`,
// src/trainers/reconstruct.py:124) and once per t-start (`decode_stage_2_outputs`, :166).
//
//   encoder: [Conv k4 s2 p1 + ReLU, num_res_layers x ResidualUnit] per level, then Conv k3 -> embedding_dim
//   quantizer: nearest codebook row (squared L2), straight-through output
//   decoder: Conv k3, then per level [num_res_layers x ResidualUnit, ConvTranspose k4 s2 p1 (+ ReLU except the last)]
//   ResidualUnit(x) = relu(x + conv3(relu(conv3(x))))
//
// Every conv with >= 64 input channels runs on the tcgen05 implicit-GEMM kernels of conv_gemm.cu (fp16 operands, fp32
// accumulation, bias / residual / ReLU in the epilogue): the 4-tap stride-2 convs through TMA element strides, the
// transposed convs as 2^d sub-pixel phases of 2-tap convs over the low-resolution tensor (y = 2m + p reads inputs
// m + p - 1 + a with kernel tap 3 - p - 2a, a in {0, 1}: the same phase machinery as the UNet's upsample conv). The
// image-side layers have 1-3 channels: the first conv is an im2col into a 64-wide K block + a 1x1 GEMM, the last
// transposed conv a GEMM to fp32 tap products + a gather. Codebook search is fp32 on CUDA cores (indices must not
// depend on fp16 rounding of the distances).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "conv_gemm.cuh"
#include "conv_halo.cuh"

namespace ddpm {

constexpr int kVqMaxLevels = 8;

struct VqVaeConfig {
    int spatial_dims;
    int in_channels, out_channels;
    int num_levels;
    int num_res_layers;
    int num_channels[kVqMaxLevels];
    int num_res_channels[kVqMaxLevels];
    int num_embeddings, embedding_dim;
    int precise_encode;  // encoder on fp16 hi + lo operand halves (see ddpm_vqvae_config)
};

class VqVae {
   public:
    explicit VqVae(const VqVaeConfig& cfg) : cfg_(cfg) {}
    ~VqVae();
    int init();
    int set_param(const char* name, const float* data, long long numel, cudaStream_t stream);
    int finalize(cudaStream_t stream);
    int num_params_expected() const { return static_cast<int>(slots_.size()); }
    // image extents -> workspace bytes for encode / decode at batch N (0 on error)
    size_t workspace_bytes(int N, int D, int H, int W) const;
    // x: fp32 [N, Cin, D, H, W] -> latent: fp32 [N, E, d, h, w] (quantised, straight-through form x + (q - x));
    // indices: int32 [N, d, h, w] or null
    int encode(const float* x, float* latent, int* indices, int N, int D, int H, int W, void* ws, size_t ws_bytes,
               cudaStream_t stream);
    // z: fp32 [N, E, d, h, w] (d = D / 2^levels, ...) -> quantise -> image: fp32 [N, Cout, D, H, W]
    // indices_in (optional, instead of z): int32 [N, d, h, w] codebook rows to decode (decode_samples)
    int decode(const float* z, const int* indices_in, float* image, int* indices_out, int N, int D, int H, int W, void* ws,
               size_t ws_bytes, cudaStream_t stream);
    long long launches() const { return launches_; }

   private:
    struct Slot { int kind; void* dst; long long numel; int Cout, Cin, taps; long long ktot; bool set; };
    struct Res { __half *w1, *w2; float *b1, *b2; int C, R; };
    struct EncLevel { __half* w; float* b; int Cin, Cout; std::vector<Res> res; };
    struct DecLevel { std::vector<Res> res; __half* w; float* b; int Cin, Cout; bool last; };
    struct Op {
        enum Type { IM2COL, GEMM, HALO, QUANT_HALF, QUANT_ROWS_F32, QUANT_NCHW, GATHER } type;
        ConvLaunch conv;
        ConvHaloLaunch halo;  // HALO: stride-1 3x3(x3) convs on the halo-tile kernel where it supports the geometry
        // IM2COL / GATHER / QUANT geometry
        const void* src; void* dst; void* dst2; int* idx;
        int N, C, D, H, W, K;
        long long rows;
    };
    struct Plan { std::vector<Op> ops; };
    template <typename T>
    T* arena(size_t count, bool half_arena);
    Res make_res(const std::string& prefix, int C, int R, bool split);
    void slot(const std::string& name, int kind, void* dst, long long numel, int Cout = 0, int Cin = 0, int taps = 0,
              long long ktot = 0);
    int build(Plan& plan, bool decode, int N, int D, int H, int W, void* ws, size_t ws_bytes, bool dry, size_t* need,
              const float* io_in, float* io_out, int* indices, const int* indices_in) const;
    int run(const Plan& plan, cudaStream_t stream);

    VqVaeConfig cfg_;
    std::vector<EncLevel> enc_;
    __half* enc_out_w_ = nullptr; float* enc_out_b_ = nullptr;
    __half* dec_in_w_ = nullptr; float* dec_in_b_ = nullptr;
    std::vector<DecLevel> dec_;
    float* codebook_ = nullptr;   // [K][E] fp32
    float* code_sq_ = nullptr;    // [K] squared norms
    int first_kpad_ = 64;         // K of the first conv's im2col block (in_channels * 4^d rounded up to 64)
    int last_cols_ = 128;         // N of the last transposed conv's tap GEMM ((2^d phases x 2^d taps x Cout) padded to 128)
    size_t f32_count_ = 0, f16_count_ = 0, f32_used_ = 0, f16_used_ = 0;
    float* f32_arena_ = nullptr;
    __half* f16_arena_ = nullptr;
    bool sizing_ = true;
    std::map<std::string, Slot> slots_;
    bool finalized_ = false;
    long long launches_ = 0;
};

}  // namespace ddpm
