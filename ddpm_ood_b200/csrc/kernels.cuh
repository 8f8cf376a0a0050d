// Non-GEMM kernels of the reconstruction path: GroupNorm+SiLU, timestep embedding, attention core, conv_in / conv_out
// (few-channel image side of the UNet), nearest upsample, forward noising, PLMS update, clamp + MSE.
// All launchers are asynchronous on `stream` and return 0 / non-zero with ddpm::set_error().
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddpm {

// GroupNorm (+ optional SiLU) over the channel concatenation of up to two channels-last fp16 tensors.
// src0: [N, S, C0], src1: [N, S, C1] or null; gamma/beta: [C0+C1] fp32; out: [N, S, C0+C1] fp16.
int gn_silu(const __half* src0, int C0, const __half* src1, int C1, const float* gamma, const float* beta,
            __half* out, int N, int S, int groups, float eps, bool silu, cudaStream_t stream);

// Sinusoidal embedding -> Linear(E, 4E) -> SiLU -> Linear(4E, 4E) -> SiLU (the SiLU every ResnetBlock applies before its
// time_emb_proj).  timesteps: [R] int64 on the device, or null for the uniform value t_uniform with R == 1.
// With timesteps == null and R > 1, row r uses the timestep t_uniform + r (table build).
// act_out: [R, 4E] fp32.
int time_embed(const long long* timesteps, int t_uniform, int R, int E, const float* w0, const float* b0,
               const float* w1, const float* b1, float* act_out, cudaStream_t stream);
// All ResnetBlock time_emb_proj layers at once: out[R, P] = act[R, K] @ Wcat[P, K]^T + bcat[P].
int time_proj_all(const float* act, int R, int K, const float* wcat, const float* bcat, int P, float* out,
                  cudaStream_t stream);

// softmax(q k^T * scale) v per (image, head). qkv: [N*T, 3*C] fp16 (q | k | v column blocks), out: [N*T, C] fp16.
int attention_core(const __half* qkv, __half* out, int N, int T, int C, int heads, float scale, cudaStream_t stream);

// Nearest-neighbour x2 upsample, channels-last fp16. in: [N, D, H, W, C] -> out: [N, D*fd, 2H, 2W, C] (fd=2 if 3-D).
int upsample_nearest2(const __half* in, __half* out, int N, int D, int H, int W, int C, int spatial_dims,
                      cudaStream_t stream);

// conv_in for few input channels (Cin <= 8): x fp32 [N, Cin, D, H, W] -> out fp16 [N, D, H, W, Cout]; w fp32
// [Cout, Cin, taps]; 3x3(x3), pad 1.
// stats_out (optional, only when conv_in_has_stats()): GroupNorm partial sums [N][conv_in_stats_parts][Cout/4][2].
int conv_in_small(const float* x, const float* w, const float* b, __half* out, int N, int Cin, int D, int H, int W,
                  int Cout, int spatial_dims, float* stats_out, cudaStream_t stream);
int conv_in_stats_parts(int D, int H, int W);
bool conv_in_has_stats(int Cin, int Cout, int spatial_dims);

// GroupNorm (+ SiLU) from producer-side partial statistics (see ConvGemmParams::stats_out): one read + one write.
// st0/st1: [N][parts][C/4][2] fp32 partial (sum, sum of squares) per 4-channel quad.
int gn_apply(const __half* src0, int C0, const float* st0, int parts0, const __half* src1, int C1, const float* st1,
             int parts1, const float* gamma, const float* beta, __half* out, int N, int S, int groups, float eps,
             bool silu, cudaStream_t stream);

// The same statistics -> (scale, shift) table ab[N][C0+C1][2] fp32 (y = x * scale + shift is GroupNorm's affine output)
// for consumers that normalise their input on the fly (conv_halo.cuh). Bitwise the values gn_apply uses.
int gn_finalize(int C0, const float* st0, int parts0, int C1, const float* st1, int parts1, const float* gamma,
                const float* beta, float* ab, int N, int S, int groups, float eps, cudaStream_t stream);

// Layout conversions for many-channel inputs/outputs (latent models): fp32 [N, C, S] <-> fp16 [N, S, C].
int nchw_to_nhwc_half(const float* x, __half* out, int N, int C, long long S, cudaStream_t stream);

struct PlmsStep {
    // eps_bar = c[0]*eps_new + c[1]*h1 + c[2]*h2 + c[3]*h3   (h1 = newest history entry BEFORE this step's append)
    float c[4];
    // model_output' = vA * eps_bar + vB * sample   (v-prediction; (1, 0) for epsilon)
    float vA, vB;
    // prev = A * sample - Bc * model_output'
    float A, Bc;
    int use_stash;    // sample := cur_sample stash (the counter == 1 corrector step)
    int write_stash;  // cur_sample := sample (counter == 0)
    int push;         // append eps_new to the history ring (counter != 1)
    int slot_new;     // ring slot to write eps_new into when push
    int slot[3];      // ring slots of h1, h2, h3
};

// conv_out for few output channels (Cout <= 8): z fp16 [N, D, H, W, Cin] (already GN+SiLU'd) -> eps fp32
// [N, Cout, D, H, W] (eps_out may be null when fused).  If `plms` is non-null the scheduler step is fused onto the
// tail: ring ([4][numel]), stash and sample are fp32 tensors shaped like eps; sample is updated in place.
int conv_out_small(const __half* z, const float* w, const float* b, float* eps_out, int N, int Cin, int D, int H,
                   int W, int Cout, int spatial_dims, const PlmsStep* plms, float* ring, float* stash, float* sample,
                   cudaStream_t stream);
// Final GroupNorm+SiLU fused with the few-channel output conv's channel reduction (2-D): d_out[N*S][9*Cout] fp32 tap
// values, then conv_out_gather sums the 3x3 neighbourhood, adds the bias and (optionally) applies the PLMS update.
bool conv_out_taps_supported(int C, int Cout, int spatial_dims);
int gn_apply_taps(const __half* src, int C, const float* st, int parts, const float* gamma, const float* beta,
                  const float* w /*fp32 [Cout][C][9]*/, int Cout, float* d_out, int N, int S, int groups, float eps,
                  cudaStream_t stream);
int conv_out_gather(const float* d, const float* b, float* eps_out, int N, int H, int W, int Cout, const PlmsStep* plms,
                    float* ring, float* stash, float* sample, cudaStream_t stream);
// Stand-alone PLMS update (model output produced elsewhere, e.g. the drop-in scheduler.step()).
int plms_update(const float* eps_new, const PlmsStep& st, float* ring, float* stash, const float* sample_in,
                float* sample_out, long long numel, cudaStream_t stream);

// x_t = sqrt(ac[t]) * (b_scale * x0) + sqrt(1 - ac[t]) * noise   (src/trainers/reconstruct.py:143-147).
// timesteps: device int64 [N] or null for the uniform value t_uniform.
int add_noise(const float* x0, const float* noise, const float* alphas_cumprod, const long long* timesteps,
              int t_uniform, float b_scale, float* out, int N, long long per_image, cudaStream_t stream);

// recon = clamp(x / b_scale, 0, 1); mse[n] = mean((x0 - recon)^2)   (src/trainers/reconstruct.py:167-168,188-191)
int clamp_mse(const float* x, const float* x0, float b_scale, float* recon, float* mse, int N, long long per_image,
              cudaStream_t stream);

}  // namespace ddpm
