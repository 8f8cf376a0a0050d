// DiffusionModelUNet forward engine: owns packed weights, builds a static launch plan per (batch, spatial) shape and
// replays it on a stream. Mirrors the module the reference builds at src/trainers/base.py:66-86 and calls at
// src/trainers/reconstruct.py:150-153.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "attention.cuh"
#include "attn_block.cuh"
#include "conv_gemm.cuh"
#include "conv_halo.cuh"
#include "kernels.cuh"

namespace ddpm {

constexpr int kMaxLevels = 8;

struct UNetConfig {
    int spatial_dims;
    int in_channels, out_channels;
    int num_levels;
    int num_channels[kMaxLevels];
    int attention_levels[kMaxLevels];
    int num_res_blocks[kMaxLevels];
    int num_head_channels[kMaxLevels];
    int norm_num_groups;
    float norm_eps;
};

struct ResW {
    int c0, c1;  // input channels: main tensor, concatenated skip tensor (0 if none)
    int cout;
    bool skip_conv;
    float *g1, *b1, *g2, *b2;
    __half* w1;  // [cout][9*(c0+c1)]
    float* bias1;
    __half* w2;  // [cout][9*cout (+ c0 + c1 when skip_conv)]
    float *bias2, *bias_skip, *bias2_total;
    __half* w2_id;  // identity-residual blocks only: [cout][(taps + 1) * cout] = conv2 weights | I (residual as one more K segment)
    int temb_off;
    // 3-D models: the dense form of both convs for 2 x 2 x 2 maps - under a padded 3x3x3 conv every output voxel sees every
    // input voxel through a distinct tap, so the conv is a linear layer over (voxel, channel): w1d [8 cout][8 (c0 + c1)],
    // w2d [8 cout][8 cout (+ 8 c0 + 8 c1: the 1x1 skip conv, block diagonal over voxels)], biases tiled over the voxels
    __half *w1d = nullptr, *w2d = nullptr;
    float *bias1d = nullptr, *bias2d = nullptr;
    std::string prefix;
};
struct AttnW {
    int C, heads;
    float *g, *b;
    __half* wqkv;  // [3C][C]
    float* bqkv;
    __half* wproj;  // [C][C]
    float* bproj;
    std::string prefix;
};
struct SampW {  // down / up sampler conv
    int C;
    __half* w;
    __half* w_up;  // up sampler only: sub-pixel phase weights [2^d * C][2^d * C]
    float* bias;
};

struct ParamSlot {
    enum Kind { COPY_F32, PACK_CONV, PACK_UPCONV } kind;
    void* dst2;          // PACK_UPCONV: phase-packed destination (dst keeps the plain 3x3 packing)
    void* dst;
    long long numel;     // expected element count of the source tensor
    int Cout, Cin, taps; // PACK_CONV
    long long ktot, koff;
    bool set;
};

struct Op {
    // GN_FINALIZE / CONV_HALO (appended so the profile indices of the older types stay put): GroupNorm statistics ->
    // per-(image, channel) scale/shift table, consumed by the halo-tile conv that normalises its input on the fly
    enum Type { CONV_IN_SMALL, CONV_IN_GEMM, GN, GEMM, ATTN, UPSAMPLE, CONV_OUT_SMALL, CONV_OUT_GEMM, GN_FINALIZE,
                CONV_HALO, ATTN_BLOCK } type;
    // GN
    const __half *src0, *src1;
    int C0, C1;
    const float *gamma, *beta;
    const float *st0, *st1;  // producer-side GroupNorm partial statistics (null: two-pass gn_silu)
    int parts0, parts1;
    float* dtaps;  // out norm fused with conv_out: per-pixel tap values (GN op writes, CONV_OUT_SMALL op gathers)
    __half* dst;
    int S;
    bool silu;
    float* ab;  // GN_FINALIZE: destination table [N][C0 + C1][2]
    // GEMM
    ConvLaunch conv;
    ConvHaloLaunch halo;  // CONV_HALO; CONV_IN_GEMM / CONV_OUT_GEMM with on_halo
    bool on_halo;         // CONV_IN_GEMM / CONV_OUT_GEMM: `halo` instead of `conv` (wide-channel 3-D latents at 8 x 8 slabs)
    bool uses_temb;
    int temb_off;
    // ATTN
    int T, C, heads;
    float scale;
    bool attn_tc;       // tcgen05 kernel (attention.cu) instead of the generic CUDA-core one
    AttnTcLaunch attn;
    AttnBlockLaunch attn_block;  // ATTN_BLOCK: the whole AttentionBlock in one launch (reports under ATTN in the profile)
    // UPSAMPLE
    int D, H, W;
    // profiling: algorithmic FLOPs (2 per MAC, real rows only) for GEMM-type ops, algorithmic bytes otherwise
    double flops;
    double bytes;
};

constexpr int kNumOpTypes = 9;  // Op::Type values + 8 = timestep embedding (2 launches)
struct OpProfile {
    double ms[kNumOpTypes];
    double flops[kNumOpTypes];
    double bytes[kNumOpTypes];
    long long launches[kNumOpTypes];
    long long forwards;  // profiled forwards
    double forward_ms;   // device time of the profiled forwards (first op start -> last op end)
};

struct Plan {
    ~Plan() { for (auto& e : events) cudaEventDestroy(e); }
    int N, D, H, W;
    void* ws;
    std::vector<Op> ops;
    std::vector<cudaEvent_t> events;  // ops.size() + 2 (time-embed start, each op start, end), created on the first profiled forward
    bool events_pending = false;
    float* temb_act;  // [N][4E]
    float* temb_all;  // [N][P]
    __half* z_out;    // input of conv_out (small path)
    __half* x_half;   // conv_in gemm path: input in NDHWC fp16
    __half* y_half;   // conv_out gemm path: output in NDHWC fp16
    float* eps_tmp;   // conv_out gemm path: fp32 eps of a PLMS step that does not push into the history ring
};

class UNet {
   public:
    explicit UNet(const UNetConfig& cfg);
    ~UNet();
    int init();  // allocate weight arenas; returns non-zero on failure
    int set_param(const char* name, const float* data, long long numel, cudaStream_t stream);
    int finalize(cudaStream_t stream);
    size_t workspace_bytes(int N, int D, int H, int W) const;
    // x: fp32 [N, Cin, D, H, W]; timesteps: device int64 [N] or null (then t_uniform applies to every image);
    // out: fp32 [N, Cout, D, H, W] (may be null when `plms` is given). When `plms` is non-null the scheduler update is
    // fused after the output conv and `sample` (== x allowed) is updated in place.
    int forward(const float* x, const long long* timesteps, int t_uniform, float* out, int N, int D, int H, int W,
                void* ws, size_t ws_bytes, cudaStream_t stream, const PlmsStep* plms = nullptr, float* ring = nullptr,
                float* stash = nullptr, float* sample = nullptr);
    const UNetConfig& config() const { return cfg_; }
    int num_params_expected() const { return static_cast<int>(slots_.size()); }
    long long launches() const { return launches_; }
    // The whole PLMS chain of one t-start (n_steps x [UNet forward + fused scheduler update]) as ONE CUDA graph launch:
    // captured from the ordinary launch sequence the first time a (timesteps, coefficients, buffers) combination is
    // seen, replayed afterwards (the chains of a batch recur for every batch). Keeps the programmatic-dependent-launch
    // edges; removes the per-kernel launch work of ~45 x n_steps launches per chain.
    int run_chain(int n_steps, const int* timesteps, const PlmsStep* steps, float* sample, float* ring, float* stash, int N,
                  int D, int H, int W, void* ws, size_t ws_bytes, cudaStream_t stream);
    // Profile every `every`-th forward with CUDA events around each op (0 = off). Harvesting synchronises the host
    // with the profiled forward's last event, so keep `every` large inside timed regions.
    void set_profile(int every) { profile_every_ = every; profile_tick_ = 0; }
    int read_profile(OpProfile* out, bool reset);
    double flops_per_image(int D, int H, int W) const;

   private:
    struct Layout;  // buffer planner
    int build_plan(Plan& plan, int N, int D, int H, int W, void* ws, size_t ws_bytes, bool dry, size_t* need) const;
    ResW make_res(const std::string& prefix, int c0, int c1, int cout);
    AttnW make_attn(const std::string& prefix, int C, int head_channels);
    template <typename T>
    T* arena_alloc(size_t count, bool half_arena);
    void add_copy(const std::string& name, float* dst, long long numel);
    void add_pack(const std::string& name, __half* dst, int Cout, int Cin, int taps, long long ktot, long long koff);

    UNetConfig cfg_;
    int E_;       // num_channels[0]
    int P_ = 0;   // total time_emb_proj outputs
    // weights
    float *conv_in_w_ = nullptr, *conv_in_b_ = nullptr;  // small path keeps fp32 [Cout][Cin][taps]
    __half* conv_in_wp_ = nullptr;                        // gemm path
    float *te_w0_, *te_b0_, *te_w1_, *te_b1_;
    float *tp_w_, *tp_b_;
    float *temb_table_ = nullptr, *temb_table_act_ = nullptr;  // [temb_rows_][P_], [temb_rows_][4E]
    int temb_rows_ = 1000;                                      // num_train_timesteps of every reference scheduler
    struct Level { std::vector<ResW> res; std::vector<AttnW> attn; bool has_samp; SampW samp; };
    std::vector<Level> down_, up_;
    std::vector<void*> dense2_allocs_;  // cudaMalloc'ed dense-form weights of the ResnetBlocks (3-D models)
    ResW mid1_, mid2_;
    AttnW mid_attn_;
    float *out_g_, *out_b_;
    float *conv_out_w_ = nullptr, *conv_out_b_ = nullptr;
    __half* conv_out_wp_ = nullptr;
    bool in_gemm_, out_gemm_;
    bool use_attn_tc_ = true;
    bool upconv_phases_ = true;  // nearest-x2 + conv as sub-pixel 2x2 convs (4/9 of the MACs, no upsampled tensor)
    bool fuse_gn_stats_ = true;  // GroupNorm statistics from the producers' epilogues (cpg % 4 == 0 required)
    bool use_attn_fused_ = true;  // one kernel per AttentionBlock (attn_block.cu) where its shape limits allow
    bool id_residual_mma_ = true;    // identity residuals of halo-kernel ResnetBlocks ride the MMA as an I-weighted K segment
    bool halo_gn_in_kernel_ = true;  // the halo conv derives scale/shift from producer statistics itself (no gn_finalize)
    bool use_halo_ = true;       // halo-tile conv kernel with GroupNorm+SiLU applied on the fly (2-D, images >= 16 x 8)
    // arenas
    size_t f32_count_ = 0, f16_count_ = 0, f32_used_ = 0, f16_used_ = 0;
    float* f32_arena_ = nullptr;
    __half* f16_arena_ = nullptr;
    bool sizing_ = true;
    std::map<std::string, ParamSlot> slots_;
    bool finalized_ = false;
    mutable std::map<std::tuple<int, int, int, int, void*>, std::unique_ptr<Plan>> plans_;
    long long launches_ = 0;
    struct ChainGraph { cudaGraphExec_t exec; long long launches; };
    std::map<std::string, ChainGraph> chain_graphs_;
    bool chain_warm_ = false;   // the first chain runs uncaptured (lazy cudaFuncSetAttribute calls, plan construction)
    bool use_chain_graph_ = true;
    int profile_every_ = 0;
    long long profile_tick_ = 0;
    OpProfile prof_{};
    void harvest(Plan& plan);
};

}  // namespace ddpm
