/*
 * ddpm_ood_b200 — C ABI of the B200-native reconstruction hot path of marksgraham/ddpm-ood.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t (as void*), launches
 * asynchronously on that stream, never synchronises the host, never frees caller memory, and returns 0 on success or a
 * non-zero code whose text is available from ddpm_last_error() (no exceptions cross this boundary).
 * One host thread per process, one process per GPU (the reference's torchrun model, src/trainers/base.py:22-33).
 *
 * The reference has no FFI of its own: the seam is a set of Python call sites. Each group below names
 * the call site it replaces.
 */
#ifndef DDPM_OOD_B200_H
#define DDPM_OOD_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DDPM_API __attribute__((visibility("default")))
#else
#define DDPM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ general */
DDPM_API const char* ddpm_last_error(void);
/* ABI version, bumped on any signature or struct-layout change; the Python binding refuses a library whose version or
 * struct sizes (ddpm_struct_sizes) differ from its own. */
#define DDPM_ABI_VERSION 7
DDPM_API int ddpm_abi_version(void);

/* ------------------------------------------------------------------------------------------------ building block
 * Implicit-GEMM convolution / linear layer on tcgen05 tensor cores. This is the operator behind every 3x3(x3) conv,
 * 1x1 skip conv and attention Linear inside DiffusionModelUNet.forward, which the reference calls at
 * src/trainers/reconstruct.py:150-153 (model built at src/trainers/base.py:66-75).
 * Activations are channels-last fp16 (N, D, H, W, C); weights are fp16 [w_rows][Ktot] packed by
 * ddpm_pack_conv_weight(). */
typedef struct ddpm_conv_args {
    int spatial_dims;        /* 2 or 3; for 2, D must be 1 */
    int N, D, H, W;          /* input extents */
    int stride;              /* 1 or 2 (padding is always k/2) */
    int n_seg;               /* 1..3 K-segments accumulated into one output */
    const void* seg_ptr[3];  /* fp16 NDHWC inputs */
    int seg_channels[3];     /* multiples of 64 */
    int seg_ksize[3];        /* 1 or 3 */
    const void* weights;     /* fp16 [w_rows][Ktot], Ktot = sum_s taps_s * channels_s, ordered segment, tap, channel */
    int w_rows;
    int Cout;                /* multiple of 128 */
    int b_rows_per_mtile;    /* batched GEMM: weight-row offset added per 128-row M tile; 0 for a conv */
    int mode;                /* 0 store, 1 block-diagonal softmax, 2 store with transposed V columns */
    const float* bias;       /* [Cout] or NULL */
    const float* chan_add;   /* [N][Cout] or NULL: per-image channel offset (timestep embedding projection) */
    const void* residual;    /* fp16, same layout as out, or NULL */
    void* out;               /* fp16 (N, Do, Ho, Wo, Cout) */
    float scale;             /* mode 1 */
    int group;               /* mode 1: tokens per image */
    int vt_col0;             /* mode 2 */
    void* out_vt;            /* mode 2 */
    int upsample2;           /* 1: out = conv3x3(nearest_upsample_x2(input)) computed as 2^dims sub-pixel 2x2 convs over the
                                low-resolution input (N,D,H,W = low-res extents, one segment with ksize 2, weights from
                                ddpm_pack_upconv_weight); the Upsample block of DiffusionModelUNet. 4/9 (8/27) of the MACs. */
    int impl;                /* 0: im2col-tile kernels, variant picked (CTA pairs whenever two 128-pixel tiles exist);
                                1: single-CTA im2col kernel only; 3: halo-tile kernel (stride-1 2-D 3x3 / 1x1 convs: the
                                input is staged once per 64 channels as a haloed tile - an 8 x 16 region of an image, or
                                two whole images of up to 8 x 8 pixels - and the 9 taps read shifted views of it) */
    float* stats_out;        /* mode 0, optional: GroupNorm partial statistics of the fp16-rounded output,
                                [N][ddpm_conv_stats_parts()][Cout/4][2] fp32 = (sum, sum of squares) per 4-channel quad
                                and 32-pixel part of an image; consumed by ddpm_gn_apply() / ddpm_gn_finalize();
                                impl 3 emits ddpm_conv_halo_stats_parts() parts instead */
    const float* gn_scale_shift; /* impl 3, optional: [N][gn_channels][2] fp32 (scale, shift) from ddpm_gn_finalize(); the
                                3x3 segments' inputs are normalised on the fly, z = silu(x * scale + shift) -> fp16,
                                i.e. the conv consumes GroupNorm+SiLU of its raw input without that tensor existing */
    int gn_channels;         /* total channels of the 3x3 segments */
    /* impl 3, instead of gn_scale_shift: the kernel derives the table itself, per work item, from the statistics its
       input tensors' producers left (stats_out layout) - no ddpm_gn_finalize launch. The normalised channels are the
       concatenation (segment order) of one or two tensors. */
    const float* gn_st0; int gn_parts0; int gn_c0;
    const float* gn_st1; int gn_parts1; int gn_c1;
    const float* gn_gamma; const float* gn_beta; int gn_groups; float gn_eps;
    int gn_no_act;           /* 1: gn_scale_shift without the SiLU (AttentionBlock norm); with no 3x3 segment the table
                                applies to the first (1x1) segment */
    int concat3x3;           /* impl 3: the 3x3 segments are channel slices of ONE conv weight [Cout][9][C_total] (a conv
                                over the channel concatenation of the inputs) instead of one K block per segment */
} ddpm_conv_args;
DDPM_API int ddpm_conv_forward(const ddpm_conv_args* args, void* stream);
/* Parts per image that ddpm_conv_forward emits for an output of this geometry (0: unsupported, use ddpm_gn_silu). */
DDPM_API int ddpm_conv_stats_parts(int spatial_dims, int Dout, int Hout, int Wout);
DDPM_API int ddpm_conv_halo_stats_parts(int Hout, int Wout);
/* 3-D volumes of 8 x 8 slabs on the halo kernel (impl 3, spatial_dims 3): parts for a Dout x Hout x Wout output */
DDPM_API int ddpm_conv_halo_stats_parts3(int Dout, int Hout, int Wout);

/* GroupNorm(groups, eps) (+ SiLU) over the channel concatenation of up to two channels-last fp16 tensors
 * src0 [N,S,C0], src1 [N,S,C1] (or NULL) -> out [N,S,C0+C1] fp16: the norm in front of every conv of
 * DiffusionModelUNet (ResnetBlock norm1/norm2, AttentionBlock norm, out.0).
 * ddpm_gn_silu computes the statistics itself (two passes); ddpm_gn_apply takes the producers' partial statistics
 * (st0/st1 as written through ddpm_conv_args.stats_out) and makes one pass. */
DDPM_API int ddpm_gn_silu(const void* src0, int C0, const void* src1, int C1, const float* gamma, const float* beta,
                          void* out, int N, int S, int groups, float eps, int silu, void* stream);
DDPM_API int ddpm_gn_apply(const void* src0, int C0, const float* st0, int parts0, const void* src1, int C1,
                           const float* st1, int parts1, const float* gamma, const float* beta, void* out, int N, int S,
                           int groups, float eps, int silu, void* stream);

/* Producer statistics -> per-(image, channel) GroupNorm (scale, shift) table ab [N][C0+C1][2] fp32 for
 * ddpm_conv_args.gn_scale_shift: scale = gamma * rstd, shift = beta - mean * scale. */
DDPM_API int ddpm_gn_finalize(int C0, const float* st0, int parts0, int C1, const float* st1, int parts1,
                              const float* gamma, const float* beta, float* ab, int N, int S, int groups, float eps,
                              void* stream);

/* The UNet's first conv (DiffusionModelUNet.conv_in: image channels -> num_channels[0], 3x3 / 3x3x3, pad 1) from the
 * fp32 NC(D)HW sample to the fp16 channels-last activation, with the GroupNorm statistics partials of the next norm
 * ([N, ddpm_conv_in_stats_parts, Cout/4, 2] fp32; NULL to skip, and required NULL where ddpm_conv_in_stats_parts is 0). */
DDPM_API int ddpm_conv_in_stats_parts(int Cin, int Cout, int spatial_dims, int D, int H, int W);
DDPM_API int ddpm_conv_in(const float* x, const float* w, const float* b, void* out, int N, int Cin, int D, int H, int W,
                          int Cout, int spatial_dims, float* stats_out, void* stream);

/* The UNet's tail, `out` = GroupNorm -> SiLU -> 3x3 conv to the image's few channels (DiffusionModelUNet.out,
 * generative/networks/nets/diffusion_model_unet.py, called from src/trainers/reconstruct.py:150), without materialising
 * the normalised tensor: src fp16 [N, H*W, C] with its statistics partials st [N, parts, C/4, 2] (sum, sum of squares
 * per 4-channel quad, as the producing conv's epilogue emits them); w fp32 [Cout, C, 3, 3], b fp32 [Cout];
 * taps_ws fp32 [N, H*W, 9*Cout] scratch; out fp32 [N, Cout, H, W]. 2-D, (C, Cout) in {(128,1), (128,3), (256,1)}. */
DDPM_API int ddpm_out_norm_conv(const void* src, const float* st, int parts, const float* gamma, const float* beta,
                                const float* w, const float* b, float* taps_ws, float* out, int N, int C, int H, int W,
                                int Cout, int groups, float eps, void* stream);

/* fp32 conv weight [Cout][Cin][3^dims] -> fp16 sub-pixel phase weights [2^dims * Cout][2^dims * Cin] for upsample2. */
DDPM_API int ddpm_pack_upconv_weight(const float* w, int Cout, int Cin, int spatial_dims, void* dst, void* stream);

/* Attention core of monai-generative's AttentionBlock (one head = 256 channels): out = softmax(q k^T * scale) v per
 * (image, head). qkv: fp16 [N*T, 3C] rows = tokens, columns q | k | v; out: fp16 [N*T, C].
 * impl: 0 = pick (tcgen05 kernel when 128 % T == 0 or T == 256, else the generic kernel), 1 = force generic. */
DDPM_API int ddpm_attention(const void* qkv, void* out, int N, int T, int C, int heads, float scale, int impl,
                            void* stream);

/* The WHOLE AttentionBlock of monai-generative's DiffusionModelUNet in one launch (C = 256, one head, T <= 128 tokens per
 * image): out = h + proj(softmax(q k^T * scale) v), q | k | v = Linear(GroupNorm(h)). Replaces the module the reference
 * reaches through model(x, timesteps) at src/trainers/reconstruct.py:150-153 (attention_levels of base.py:66-75).
 * h, out: fp16 [N*T, C] (tokens x channels, channels-last activations); gamma, beta: fp32 [C] (32 groups);
 * wqkv: fp16 [3C, C] (to_q | to_k | to_v weights, rows = output channels), bqkv fp32 [3C]; wproj fp16 [C, C], bproj [C];
 * stats_out: null, or fp32 [N][ddpm_attention_block_stats_parts(T)][C/4][2] GroupNorm partial sums of `out`. */
DDPM_API int ddpm_attention_block(const void* h, void* out, int N, int T, int C, int heads, int groups, float eps,
                                  float scale, const float* gamma, const float* beta, const void* wqkv, const float* bqkv,
                                  const void* wproj, const float* bproj, float* stats_out, void* stream);
DDPM_API int ddpm_attention_block_stats_parts(int T);

/* fp32 PyTorch conv weight [Cout][Cin][taps] (or Linear weight with taps == 1) -> fp16 rows of a packed matrix:
 * dst[co * ktot + koff + tap * Cin + ci]. */
DDPM_API int ddpm_pack_conv_weight(const float* w, int Cout, int Cin, int taps, void* dst, long long ktot, long long koff,
                          void* stream);

/* ------------------------------------------------------------------------------------------------ DiffusionModelUNet
 * Replaces `DiffusionModelUNet(...)` (src/trainers/base.py:66-86), `load_state_dict` (base.py:145) and
 * `model(x, timesteps=...)` (src/trainers/reconstruct.py:150-153; kw form src/trainers/ddpm_trainer.py:104).
 * The handle owns fp16-packed copies of the weights; everything else is caller memory. */
#define DDPM_MAX_LEVELS 8
typedef struct ddpm_unet_config {
    int spatial_dims;                        /* 2 or 3 */
    int in_channels, out_channels;           /* <= 8, or a multiple of 64 (in) / 128 (out) for latent models */
    int num_levels;
    int num_channels[DDPM_MAX_LEVELS];       /* multiples of 128 */
    int attention_levels[DDPM_MAX_LEVELS];   /* 0 / 1 */
    int num_res_blocks[DDPM_MAX_LEVELS];
    int num_head_channels[DDPM_MAX_LEVELS];  /* 256 */
    int norm_num_groups;                     /* 32 */
    float norm_eps;                          /* 1e-6 */
} ddpm_unet_config;

DDPM_API int ddpm_unet_create(const ddpm_unet_config* cfg, void** handle);
DDPM_API void ddpm_unet_destroy(void* handle);
/* name: a state_dict key of monai-generative's DiffusionModelUNet ("conv_in.conv.weight", ...); data: device fp32,
 * contiguous, PyTorch layout. */
DDPM_API int ddpm_unet_set_param(void* handle, const char* name, const float* data, long long numel, void* stream);
/* Fails (naming the key) unless every parameter has been set. */
DDPM_API int ddpm_unet_finalize(void* handle, void* stream);
/* Bytes of caller-provided scratch needed for a forward of this shape; 0 on error. */
DDPM_API long long ddpm_unet_workspace_bytes(void* handle, int N, int D, int H, int W);
/* x: fp32 [N, Cin, (D,) H, W]; timesteps: int64 [N] on the device; out: fp32 like x with Cout channels. */
DDPM_API int ddpm_unet_forward(void* handle, const float* x, const long long* timesteps, float* out, int N, int D, int H,
                               int W, void* workspace, long long workspace_bytes, void* stream);
/* Number of kernels launched by this handle so far (bench.py's gpu_launches). */
DDPM_API long long ddpm_unet_launch_count(void* handle);

/* Per-op-type device timing (CUDA events on the launch stream) of every `every`-th forward; 0 switches it off.
 * Types: 0 conv_in(small) 1 conv_in(gemm) 2 groupnorm 3 conv/linear gemm 4 attention core 5 upsample 6 conv_out(small,
 * + fused PLMS) 7 conv_out(gemm) 8 timestep embedding. flops/bytes are the ALGORITHMIC work of the profiled launches.
 * This is measurement support for bench.py's roofline (no reference counterpart; the reference prints wall seconds per
 * batch, src/trainers/reconstruct.py:232-236). */
#define DDPM_NUM_OP_TYPES 9
typedef struct ddpm_op_profile {
    double ms[DDPM_NUM_OP_TYPES];
    double flops[DDPM_NUM_OP_TYPES];
    double bytes[DDPM_NUM_OP_TYPES];
    long long launches[DDPM_NUM_OP_TYPES];
    long long forwards;
    double forward_ms;
} ddpm_op_profile;
DDPM_API int ddpm_unet_set_profile(void* handle, int every);
DDPM_API int ddpm_unet_read_profile(void* handle, ddpm_op_profile* out, int reset);

/* ------------------------------------------------------------------------------------------------ scheduler
 * Replaces `PNDMScheduler.add_noise` (src/trainers/reconstruct.py:143-147) and `PNDMScheduler.step`
 * (src/trainers/reconstruct.py:155-157). Scheduler STATE (counter, eps history bookkeeping, alphas_cumprod which the
 * reference overwrites at :107-117) stays with the host-side Python mirror; each call carries the coefficients. */
typedef struct ddpm_plms_step {
    float c[4];      /* eps_bar = c0*eps_new + c1*h1 + c2*h2 + c3*h3 (h1 newest history entry before this step) */
    float vA, vB;    /* model_output' = vA*eps_bar + vB*sample (v-prediction), (1, 0) for epsilon */
    float A, Bc;     /* prev_sample = A*sample - Bc*model_output' */
    int use_stash;   /* sample := stashed cur_sample (the counter == 1 corrector step) */
    int write_stash; /* stash := sample (counter == 0) */
    int push;        /* append eps_new to the history ring */
    int slot_new;    /* ring slot receiving eps_new */
    int slot[3];     /* ring slots of h1, h2, h3 */
} ddpm_plms_step;

DDPM_API int ddpm_add_noise(const float* x0, const float* noise, const float* alphas_cumprod /*device [T]*/,
                            const long long* timesteps /*device [N] or NULL*/, int t_uniform, float b_scale, float* out,
                            int N, long long per_image, void* stream);
/* ring: fp32 [4][numel]; stash/sample_in/sample_out: fp32 [numel] (sample_out may alias sample_in). */
DDPM_API int ddpm_plms_update(const float* model_output, const ddpm_plms_step* step, float* ring, float* stash,
                              const float* sample_in, float* sample_out, long long numel, void* stream);
/* The inner loop of src/trainers/reconstruct.py:149-157 in one call: for each of the n_steps timesteps (host array,
 * shared by the whole batch) run the UNet on `sample` and apply the PLMS update, fused after the output conv.
 * sample is updated in place. */
DDPM_API int ddpm_unet_run_chain(void* handle, int n_steps, const int* timesteps, const ddpm_plms_step* steps,
                                 float* sample, float* ring, float* stash, int N, int D, int H, int W, void* workspace,
                                 long long workspace_bytes, void* stream);

/* sizeof() of the structs above as this library was compiled (binding self-check: a ctypes mirror must agree). */
DDPM_API void ddpm_struct_sizes(int* conv_args, int* unet_config, int* plms_step, int* op_profile);

/* ------------------------------------------------------------------------------------------------ scoring
 * recon = clamp(x / b_scale, 0, 1) and mse[n] = mean((x0 - recon)^2): src/trainers/reconstruct.py:167-168,188-191. */
DDPM_API int ddpm_clamp_mse(const float* x, const float* x0, float b_scale, float* recon, float* mse, int N,
                            long long per_image, void* stream);

/* LPIPS (AlexNet, v0.1, linear heads, spatial mean): replaces `lpips.LPIPS(...).forward(in0, in1, normalize=...)`
 * reached through src/losses/perceptual_loss.py:100-102,125,181-183 from src/trainers/reconstruct.py:170-187.
 * Parameter names are lpips.LPIPS state_dict keys: net.slice{1..5}.{0,3,6,8,10}.{weight,bias}, lin{0..4}.model.1.weight.
 * in0/in1: fp32 [B, C, H, W] with C in {1, 3} (1 broadcasts, as the ScalingLayer does); out: fp32 [B]. */
DDPM_API int ddpm_lpips_create(void** handle);
DDPM_API void ddpm_lpips_destroy(void* handle);
DDPM_API int ddpm_lpips_set_param(void* handle, const char* name, const float* data, long long numel, void* stream);
DDPM_API int ddpm_lpips_finalize(void* handle);
DDPM_API long long ddpm_lpips_workspace_bytes(void* handle, int B, int H, int W);
DDPM_API int ddpm_lpips_forward(void* handle, const float* in0, const float* in1, float* out, int B, int C, int H, int W,
                                int normalize, void* workspace, long long workspace_bytes, void* stream);
DDPM_API long long ddpm_lpips_launch_count(void* handle);

/* ------------------------------------------------------------------------------------------------ score post-processing
 * ("next" row) The arithmetic of ood_detection.py:150-206 for score tensors already on the device,
 * [n_t, n_images] fp32 per dataset and target (mse / perceptual_difference):
 *   ddpm_val_stats   mean and sample std (ddof 1, pandas' default) per t over the validation images      (:150-158)
 *   ddpm_mean_z      per image: mean over t of (score - mean_t) / std_t                                    (:159-161,174)
 *   ddpm_auc_counts  counts[0] = #{out > in}, counts[1] = #{out == in} over all (out, in) pairs;
 *                    roc_auc_score = (counts[0] + counts[1] / 2) / (n_in * n_out)                          (:191-206) */
DDPM_API int ddpm_val_stats(const float* val, int n_t, int n_val, float* mean, float* std, void* stream);
DDPM_API int ddpm_mean_z(const float* scores, const float* mean, const float* std, int n_t, int n, float* out,
                         void* stream);
DDPM_API int ddpm_auc_counts(const float* in_scores, int n_in, const float* out_scores, int n_out,
                             unsigned long long* counts, void* stream);

/* ------------------------------------------------------------------------------------------------ data ingest
 * ("next" row) ScaleIntensityd(minv=0, maxv=1) of the reference's loader
 * (src/data/get_train_and_val_dataloader.py:76) per image on the device: src [N, per_image] uint8 (src_is_u8) or fp32
 * as stored on disk, dst fp32 in [0, 1]; a constant image maps to zeros. */
DDPM_API int ddpm_scale_intensity(const void* src, int src_is_u8, float* dst, int N, long long per_image, void* stream);

/* ------------------------------------------------------------------------------------------------ simplex noise
 * ("next" row) generate_simplex_noise of src/utils/simplex_noise.py:15-79 for --simplex_noise=1
 * (src/trainers/reconstruct.py:133-139): seeds int64 [C * B] in drawing order (channel outer, image inner; device),
 * t int64 [B] (device), out fp32 [B, C, H, W]; tables_ws: C * B * 256 bytes of scratch for the permutation tables
 * (_init, :559-577). fp64 arithmetic in the reference's operand order: bit-identical values. */
DDPM_API int ddpm_simplex_noise(const long long* seeds, const long long* t, float* out, unsigned char* tables_ws, int B,
                                int C, int H, int W, int octaves, double persistence, double frequency, void* stream);

/* ------------------------------------------------------------------------------------------------ VQ-VAE (latent models)
 * monai-generative's VQVAE as the reference builds it from vqvae_config.json (src/trainers/base.py:44-61): the stage-1
 * model of the latent-diffusion path. Replaces `self.vqvae_model.encode_stage_2_inputs(images)`
 * (src/trainers/reconstruct.py:124) and `self.vqvae_model.decode_stage_2_outputs(reconstructions)` (:166).
 * Supported: downsample_parameters (2,4,1,1) and upsample_parameters (2,4,1,1,0) per level (the reference's README
 * configuration), channel counts that are multiples of 128, ReLU, no dropout / output activation. */
typedef struct ddpm_vqvae_config {
    int spatial_dims;                 /* 2 or 3 */
    int in_channels, out_channels;    /* image channels (1..3) */
    int num_levels;
    int num_res_layers;
    int num_channels[DDPM_MAX_LEVELS];
    int num_res_channels[DDPM_MAX_LEVELS];
    int num_embeddings, embedding_dim;
    int precise_encode;               /* 1: encoder on split-precision operands (fp16 hi + fp16 lo halves of every activation
                                         and weight, three tensor-core products per MAC, fp32 accumulation) - the reference
                                         encodes in fp32 (src/trainers/reconstruct.py:124 is outside its autocast block) and a
                                         nearest-row search amplifies fp16 rounding into different rows; 0: fp16 operands */
} ddpm_vqvae_config;

DDPM_API int ddpm_vqvae_create(const ddpm_vqvae_config* cfg, void** handle);
DDPM_API void ddpm_vqvae_destroy(void* handle);
/* name: a key of VQVAE.state_dict() ("encoder.blocks.0.conv.weight", ..., "quantizer.quantizer.embedding.weight");
 * data: fp32 device pointer in PyTorch layout. */
DDPM_API int ddpm_vqvae_set_param(void* handle, const char* name, const float* data, long long numel, void* stream);
DDPM_API int ddpm_vqvae_finalize(void* handle, void* stream);
/* Workspace for encode / decode of N images of D x H x W voxels (D == 1 in 2-D). */
DDPM_API long long ddpm_vqvae_workspace_bytes(void* handle, int N, int D, int H, int W);
/* encode_stage_2_inputs: x fp32 [N, Cin, D, H, W] -> latent fp32 [N, E, D/2^L, H/2^L, W/2^L] (nearest codebook rows);
 * indices: null or int32 [N, D/2^L, H/2^L, W/2^L] (index_quantize). */
DDPM_API int ddpm_vqvae_encode(void* handle, const float* x, float* latent, int* indices, int N, int D, int H, int W,
                               void* workspace, long long workspace_bytes, void* stream);
/* decode_stage_2_outputs: z fp32 [N, E, d, h, w] -> nearest codebook rows -> image fp32 [N, Cout, D, H, W].
 * indices_in (instead of z, which may then be null): decode_samples. indices_out: null or the rows chosen. */
DDPM_API int ddpm_vqvae_decode(void* handle, const float* z, const int* indices_in, float* image, int* indices_out, int N,
                               int D, int H, int W, void* workspace, long long workspace_bytes, void* stream);
DDPM_API long long ddpm_vqvae_launch_count(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* DDPM_OOD_B200_H */
