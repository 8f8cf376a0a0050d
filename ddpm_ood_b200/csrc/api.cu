// extern "C" surface declared in include/ddpm_ood_b200.h.
#include "../../include/ddpm_ood_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "attention.cuh"
#include "attn_block.cuh"
#include "conv_gemm.cuh"
#include "conv_halo.cuh"
#include "kernels.cuh"
#include "launch.cuh"

namespace ddpm {

int num_sms() {
    static int sms[ddpm::kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 148;
    int& v = sms[dev % ddpm::kMaxDevices];
    if (!v && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        v = 0;
        return 148;
    }
    return v;
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                        __half* __restrict__ dst, long long ktot, long long koff) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // destination-major so writes coalesce: i = (co * taps + tap) * Cin + ci
        const int ci = static_cast<int>(i % Cin);
        const long long r = i / Cin;
        const int tap = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        const float v = w[(static_cast<long long>(co) * Cin + ci) * taps + tap];
        dst[co * ktot + koff + static_cast<long long>(tap) * Cin + ci] = __float2half_rn(v);
    }
}

// Sub-pixel phase weights of conv3x3(nearest_upsample_x2(.)): dst[phase][co][tap2][ci] (fp16) where per dim a phase
// bit p and a tap bit a select the 3-tap subset {0} / {1,2} (p = 0) or {0,1} / {2} (p = 1); members are summed in fp32.
__global__ void pack_upconv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int dims,
                                          __half* __restrict__ dst) {
    const int phases = 1 << dims, taps2 = 1 << dims;
    int taps3 = 1;
    for (int i = 0; i < dims; ++i) taps3 *= 3;
    const long long total = static_cast<long long>(phases) * Cout * taps2 * Cin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        long long r = i / Cin;
        const int tap2 = static_cast<int>(r % taps2); r /= taps2;
        const int co = static_cast<int>(r % Cout);
        const int phase = static_cast<int>(r / Cout);
        const float* wp = w + (static_cast<long long>(co) * Cin + ci) * taps3;
        // per dim (0 = w fastest): range of 3-tap indices
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int d = 0; d < dims; ++d) {
            const int pb = (phase >> d) & 1, a = (tap2 >> d) & 1;
            if (pb == 0) { lo[d] = a ? 1 : 0; hi[d] = a ? 2 : 0; }
            else         { lo[d] = a ? 2 : 0; hi[d] = a ? 2 : 1; }
        }
        float acc = 0.f;
        for (int td = lo[2]; td <= hi[2]; ++td)
            for (int th = lo[1]; th <= hi[1]; ++th)
                for (int tw = lo[0]; tw <= hi[0]; ++tw) acc += wp[(td * 3 + th) * 3 + tw];
        dst[i] = __float2half_rn(acc);
    }
}

int pack_upconv_weight(const float* w, int Cout, int Cin, int dims, __half* dst, cudaStream_t stream) {
    const long long total = (1LL << (2 * dims)) * Cout * Cin;
    int blocks = static_cast<int>((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
    pack_upconv_weight_kernel<<<blocks, 256, 0, stream>>>(w, Cout, Cin, dims, dst);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pack_upconv_weight: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // namespace ddpm

extern "C" {

const char* ddpm_last_error(void) { return ddpm::last_error(); }
int ddpm_abi_version(void) { return DDPM_ABI_VERSION; }

void ddpm_struct_sizes(int* conv_args, int* unet_config, int* plms_step, int* op_profile) {
    if (conv_args) *conv_args = static_cast<int>(sizeof(ddpm_conv_args));
    if (unet_config) *unet_config = static_cast<int>(sizeof(ddpm_unet_config));
    if (plms_step) *plms_step = static_cast<int>(sizeof(ddpm_plms_step));
    if (op_profile) *op_profile = static_cast<int>(sizeof(ddpm_op_profile));
}

int ddpm_conv_forward(const ddpm_conv_args* a, void* stream) {
    if (!a) { ddpm::set_error("ddpm_conv_forward: null args"); return 2; }
    ddpm::ConvProblem q{};
    q.spatial_dims = a->spatial_dims;
    q.N = a->N; q.D = a->D; q.H = a->H; q.W = a->W;
    q.stride = a->stride;
    q.n_seg = a->n_seg;
    for (int s = 0; s < a->n_seg && s < ddpm::kMaxSeg; ++s) {
        q.seg[s].ptr = a->seg_ptr[s];
        q.seg[s].channels = a->seg_channels[s];
        q.seg[s].ksize = a->seg_ksize[s];
    }
    q.weights = a->weights;
    q.w_rows = a->w_rows;
    q.Cout = a->Cout;
    q.b_rows_per_mtile = a->b_rows_per_mtile;
    q.mode = a->mode;
    q.bias = a->bias;
    q.chan_add = a->chan_add;
    q.chan_add_stride = a->Cout;
    q.residual = a->residual;
    q.out = a->out;
    q.scale = a->scale;
    q.group = a->group;
    q.vt_col0 = a->vt_col0;
    q.out_vt = a->out_vt;
    q.stats_out = a->stats_out;
    q.upsample2 = a->upsample2;
    q.impl = a->impl;
    q.concat3x3 = a->concat3x3;
    q.gn_silu = a->gn_no_act ? 0 : 1;
    if (a->impl == 3) {
        ddpm::ConvHaloLaunch hl;
        ddpm::HaloGnSource src{};
        const bool from_stats = a->gn_st0 != nullptr;
        if (from_stats) {
            src.st0 = a->gn_st0; src.parts0 = a->gn_parts0; src.C0 = a->gn_c0;
            src.st1 = a->gn_st1; src.parts1 = a->gn_parts1; src.C1 = a->gn_st1 ? a->gn_c1 : 0;
            src.gamma = a->gn_gamma; src.beta = a->gn_beta;
            src.S = a->D * a->H * a->W; src.groups = a->gn_groups; src.eps = a->gn_eps;
        }
        int rc = ddpm::conv_halo_prepare(q, a->gn_scale_shift, a->gn_channels, ddpm::num_sms(), &hl,
                                         from_stats ? &src : nullptr);
        if (rc) return rc;
        rc = ddpm::conv_halo_launch(hl, static_cast<cudaStream_t>(stream));
        if (!rc && hl.p.dbg_cycles && getenv("DDPM_HALO_CYCLES_PRINT")) {  // experiment only
            long long h[16 * 128];
            if (!ddpm::conv_halo_read_cycles(hl, h, 16 * 128)) {
                double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const int nc = hl.grid / 2;
                for (int i = 0; i < nc; ++i) for (int k = 0; k < 8; ++k) s[k] += static_cast<double>(h[8 * i + k]);
                fprintf(stderr, "halo cycles (mean over %d clusters): total %.0f | MMA waits: tempty %.0f a_ready %.0f b_full %.0f | "
                        "A producer waits a_empty %.0f | transform: table %.0f wait a_full %.0f work %.0f\n",
                        nc, s[0] / nc, s[1] / nc, s[2] / nc, s[3] / nc, s[4] / nc, s[5] / nc, s[6] / nc, s[7] / nc);
                // wall-clock timeline (ns after the earliest kernel entry): per stamp the min / max over the clusters
                const long long* tl = h + 8 * 128;
                long long t0 = tl[0];
                for (int i = 0; i < nc; ++i) if (tl[8 * i] < t0) t0 = tl[8 * i];
                static const char* names[8] = {"entry", "prologue done", "pdl_wait done", "first A ready", "MMA loop end",
                                               "last acc full", "epilogue done", "exit"};
                fprintf(stderr, "halo timeline ns (min..max over clusters):");
                for (int k = 0; k < 8; ++k) {
                    long long lo = tl[k] - t0, hi = lo;
                    for (int i = 0; i < nc; ++i) { const long long v = tl[8 * i + k] - t0; if (v < lo) lo = v; if (v > hi) hi = v; }
                    fprintf(stderr, " %s %lld..%lld |", names[k], lo, hi);
                }
                fprintf(stderr, "\n");
            }
        }
        return rc;
    }
    if (a->gn_scale_shift) { ddpm::set_error("ddpm_conv_forward: gn_scale_shift needs impl 3"); return 2; }
    ddpm::ConvLaunch l;
    int rc = ddpm::conv_prepare(q, ddpm::num_sms(), &l);
    if (rc) return rc;
    return ddpm::conv_launch(l, static_cast<cudaStream_t>(stream));
}

int ddpm_conv_stats_parts(int spatial_dims, int Dout, int Hout, int Wout) {
    return ddpm::conv_stats_parts(spatial_dims, Dout, Hout, Wout);
}

int ddpm_conv_halo_stats_parts(int Hout, int Wout) { return ddpm::conv_halo_stats_parts(Hout, Wout); }
int ddpm_conv_halo_stats_parts3(int Dout, int Hout, int Wout) { return ddpm::conv_halo_stats_parts(Hout, Wout, Dout); }

int ddpm_gn_finalize(int C0, const float* st0, int parts0, int C1, const float* st1, int parts1, const float* gamma,
                     const float* beta, float* ab, int N, int S, int groups, float eps, void* stream) {
    return ddpm::gn_finalize(C0, st0, parts0, C1, st1, parts1, gamma, beta, ab, N, S, groups, eps,
                             static_cast<cudaStream_t>(stream));
}

int ddpm_gn_silu(const void* src0, int C0, const void* src1, int C1, const float* gamma, const float* beta, void* out,
                 int N, int S, int groups, float eps, int silu, void* stream) {
    return ddpm::gn_silu(static_cast<const __half*>(src0), C0, static_cast<const __half*>(src1), C1, gamma, beta,
                         static_cast<__half*>(out), N, S, groups, eps, silu != 0, static_cast<cudaStream_t>(stream));
}
int ddpm_gn_apply(const void* src0, int C0, const float* st0, int parts0, const void* src1, int C1, const float* st1,
                  int parts1, const float* gamma, const float* beta, void* out, int N, int S, int groups, float eps,
                  int silu, void* stream) {
    return ddpm::gn_apply(static_cast<const __half*>(src0), C0, st0, parts0, static_cast<const __half*>(src1), C1, st1,
                          parts1, gamma, beta, static_cast<__half*>(out), N, S, groups, eps, silu != 0,
                          static_cast<cudaStream_t>(stream));
}

int ddpm_conv_in_stats_parts(int Cin, int Cout, int spatial_dims, int D, int H, int W) {
    return ddpm::conv_in_has_stats(Cin, Cout, spatial_dims) ? ddpm::conv_in_stats_parts(D, H, W) : 0;
}

int ddpm_conv_in(const float* x, const float* w, const float* b, void* out, int N, int Cin, int D, int H, int W, int Cout,
                 int spatial_dims, float* stats_out, void* stream) {
    if (!x || !w || !b || !out) { ddpm::set_error("ddpm_conv_in: null argument"); return 2; }
    return ddpm::conv_in_small(x, w, b, static_cast<__half*>(out), N, Cin, D, H, W, Cout, spatial_dims, stats_out,
                               static_cast<cudaStream_t>(stream));
}

int ddpm_out_norm_conv(const void* src, const float* st, int parts, const float* gamma, const float* beta,
                       const float* w, const float* b, float* taps_ws, float* out, int N, int C, int H, int W, int Cout,
                       int groups, float eps, void* stream) {
    if (!src || !st || !gamma || !beta || !w || !b || !taps_ws || !out) { ddpm::set_error("ddpm_out_norm_conv: null argument"); return 2; }
    if (!ddpm::conv_out_taps_supported(C, Cout, 2)) { ddpm::set_error("ddpm_out_norm_conv: C=%d Cout=%d unsupported", C, Cout); return 2; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = ddpm::gn_apply_taps(static_cast<const __half*>(src), C, st, parts, gamma, beta, w, Cout, taps_ws, N, H * W,
                                 groups, eps, s);
    if (rc) return rc;
    return ddpm::conv_out_gather(taps_ws, b, out, N, H, W, Cout, nullptr, nullptr, nullptr, nullptr, s);
}

int ddpm_pack_upconv_weight(const float* w, int Cout, int Cin, int spatial_dims, void* dst, void* stream) {
    if (!w || !dst || (spatial_dims != 2 && spatial_dims != 3)) { ddpm::set_error("ddpm_pack_upconv_weight: bad argument"); return 2; }
    return ddpm::pack_upconv_weight(w, Cout, Cin, spatial_dims, static_cast<__half*>(dst), static_cast<cudaStream_t>(stream));
}

int ddpm_attention(const void* qkv, void* out, int N, int T, int C, int heads, float scale, int impl, void* stream) {
    if (!qkv || !out) { ddpm::set_error("ddpm_attention: null argument"); return 2; }
    if (impl == 0 && ddpm::attention_tc_supported(T, C, heads)) {
        ddpm::AttnTcLaunch l;
        int rc = ddpm::attention_tc_prepare(static_cast<const __half*>(qkv), static_cast<__half*>(out), N, T, C, heads,
                                            scale, &l);
        if (rc) return rc;
        return ddpm::attention_tc_launch(l, static_cast<cudaStream_t>(stream));
    }
    return ddpm::attention_core(static_cast<const __half*>(qkv), static_cast<__half*>(out), N, T, C, heads, scale,
                                static_cast<cudaStream_t>(stream));
}

int ddpm_attention_block(const void* h, void* out, int N, int T, int C, int heads, int groups, float eps, float scale,
                         const float* gamma, const float* beta, const void* wqkv, const float* bqkv, const void* wproj,
                         const float* bproj, float* stats_out, void* stream) {
    if (!h || !out || !gamma || !beta || !wqkv || !wproj) { ddpm::set_error("ddpm_attention_block: null argument"); return 2; }
    ddpm::AttnBlockLaunch l;
    int rc = ddpm::attn_block_prepare(static_cast<const __half*>(h), static_cast<__half*>(out), N, T, C, heads, groups, eps,
                                      scale, gamma, beta, static_cast<const __half*>(wqkv), bqkv,
                                      static_cast<const __half*>(wproj), bproj, stats_out, ddpm::num_sms(), &l);
    if (rc) return rc;
    return ddpm::attn_block_launch(l, static_cast<cudaStream_t>(stream));
}

int ddpm_attention_block_stats_parts(int T) { return ddpm::attn_block_stats_parts(T); }

int ddpm_pack_conv_weight(const float* w, int Cout, int Cin, int taps, void* dst, long long ktot, long long koff,
                          void* stream) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    if (total <= 0) return 0;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ddpm::pack_conv_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, Cout, Cin, taps, static_cast<__half*>(dst), ktot, koff);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("pack_conv_weight: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ UNet / scheduler
#include "engine.cuh"

static_assert(sizeof(ddpm_plms_step) == sizeof(ddpm::PlmsStep), "ddpm_plms_step must mirror ddpm::PlmsStep");
static_assert(DDPM_MAX_LEVELS == ddpm::kMaxLevels, "level capacity mismatch");

static ddpm::PlmsStep to_step(const ddpm_plms_step& s) {
    ddpm::PlmsStep o;
    memcpy(&o, &s, sizeof(o));
    return o;
}

extern "C" {

int ddpm_unet_create(const ddpm_unet_config* cfg, void** handle) {
    if (!cfg || !handle) { ddpm::set_error("ddpm_unet_create: null argument"); return 2; }
    ddpm::UNetConfig c{};
    c.spatial_dims = cfg->spatial_dims;
    c.in_channels = cfg->in_channels;
    c.out_channels = cfg->out_channels;
    c.num_levels = cfg->num_levels;
    for (int i = 0; i < DDPM_MAX_LEVELS; ++i) {
        c.num_channels[i] = cfg->num_channels[i];
        c.attention_levels[i] = cfg->attention_levels[i];
        c.num_res_blocks[i] = cfg->num_res_blocks[i];
        c.num_head_channels[i] = cfg->num_head_channels[i];
    }
    c.norm_num_groups = cfg->norm_num_groups;
    c.norm_eps = cfg->norm_eps;
    ddpm::UNet* u = new ddpm::UNet(c);
    int rc = u->init();
    if (rc) { delete u; *handle = nullptr; return rc; }
    *handle = u;
    return 0;
}

void ddpm_unet_destroy(void* handle) { delete static_cast<ddpm::UNet*>(handle); }

int ddpm_unet_set_param(void* handle, const char* name, const float* data, long long numel, void* stream) {
    if (!handle || !name || !data) { ddpm::set_error("ddpm_unet_set_param: null argument"); return 2; }
    return static_cast<ddpm::UNet*>(handle)->set_param(name, data, numel, static_cast<cudaStream_t>(stream));
}

int ddpm_unet_finalize(void* handle, void* stream) {
    if (!handle) { ddpm::set_error("ddpm_unet_finalize: null handle"); return 2; }
    return static_cast<ddpm::UNet*>(handle)->finalize(static_cast<cudaStream_t>(stream));
}

long long ddpm_unet_workspace_bytes(void* handle, int N, int D, int H, int W) {
    if (!handle) { ddpm::set_error("ddpm_unet_workspace_bytes: null handle"); return 0; }
    return static_cast<long long>(static_cast<ddpm::UNet*>(handle)->workspace_bytes(N, D, H, W));
}

int ddpm_unet_forward(void* handle, const float* x, const long long* timesteps, float* out, int N, int D, int H, int W,
                      void* workspace, long long workspace_bytes, void* stream) {
    if (!handle || !x || !timesteps || !out || !workspace) { ddpm::set_error("ddpm_unet_forward: null argument"); return 2; }
    return static_cast<ddpm::UNet*>(handle)->forward(x, timesteps, 0, out, N, D, H, W, workspace,
                                                     static_cast<size_t>(workspace_bytes),
                                                     static_cast<cudaStream_t>(stream));
}

long long ddpm_unet_launch_count(void* handle) {
    return handle ? static_cast<ddpm::UNet*>(handle)->launches() : 0;
}

int ddpm_unet_set_profile(void* handle, int every) {
    if (!handle) { ddpm::set_error("ddpm_unet_set_profile: null handle"); return 2; }
    static_cast<ddpm::UNet*>(handle)->set_profile(every);
    return 0;
}
int ddpm_unet_read_profile(void* handle, ddpm_op_profile* out, int reset) {
    if (!handle || !out) { ddpm::set_error("ddpm_unet_read_profile: null argument"); return 2; }
    static_assert(sizeof(ddpm_op_profile) == sizeof(ddpm::OpProfile), "ddpm_op_profile must mirror ddpm::OpProfile");
    static_assert(DDPM_NUM_OP_TYPES == ddpm::kNumOpTypes, "op type count mismatch");
    ddpm::OpProfile p;
    int rc = static_cast<ddpm::UNet*>(handle)->read_profile(&p, reset != 0);
    memcpy(out, &p, sizeof(p));
    return rc;
}

int ddpm_add_noise(const float* x0, const float* noise, const float* alphas_cumprod, const long long* timesteps,
                   int t_uniform, float b_scale, float* out, int N, long long per_image, void* stream) {
    return ddpm::add_noise(x0, noise, alphas_cumprod, timesteps, t_uniform, b_scale, out, N, per_image,
                           static_cast<cudaStream_t>(stream));
}

int ddpm_plms_update(const float* model_output, const ddpm_plms_step* step, float* ring, float* stash,
                     const float* sample_in, float* sample_out, long long numel, void* stream) {
    if (!step) { ddpm::set_error("ddpm_plms_update: null step"); return 2; }
    return ddpm::plms_update(model_output, to_step(*step), ring, stash, sample_in, sample_out, numel,
                             static_cast<cudaStream_t>(stream));
}

int ddpm_unet_run_chain(void* handle, int n_steps, const int* timesteps, const ddpm_plms_step* steps, float* sample,
                        float* ring, float* stash, int N, int D, int H, int W, void* workspace,
                        long long workspace_bytes, void* stream) {
    if (!handle || !timesteps || !steps || !sample || !ring || !stash || !workspace) {
        ddpm::set_error("ddpm_unet_run_chain: null argument");
        return 2;
    }
    ddpm::UNet* u = static_cast<ddpm::UNet*>(handle);
    std::vector<ddpm::PlmsStep> st(static_cast<size_t>(n_steps > 0 ? n_steps : 0));
    for (int i = 0; i < n_steps; ++i) st[i] = to_step(steps[i]);
    return u->run_chain(n_steps, timesteps, st.data(), sample, ring, stash, N, D, H, W, workspace,
                        static_cast<size_t>(workspace_bytes), static_cast<cudaStream_t>(stream));
}

int ddpm_clamp_mse(const float* x, const float* x0, float b_scale, float* recon, float* mse, int N,
                   long long per_image, void* stream) {
    return ddpm::clamp_mse(x, x0, b_scale, recon, mse, N, per_image, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ LPIPS
#include "lpips.cuh"

extern "C" {

int ddpm_lpips_create(void** handle) {
    if (!handle) { ddpm::set_error("ddpm_lpips_create: null argument"); return 2; }
    ddpm::Lpips* l = new ddpm::Lpips();
    int rc = l->init();
    if (rc) { delete l; *handle = nullptr; return rc; }
    *handle = l;
    return 0;
}
void ddpm_lpips_destroy(void* handle) { delete static_cast<ddpm::Lpips*>(handle); }
int ddpm_lpips_set_param(void* handle, const char* name, const float* data, long long numel, void* stream) {
    if (!handle || !name || !data) { ddpm::set_error("ddpm_lpips_set_param: null argument"); return 2; }
    return static_cast<ddpm::Lpips*>(handle)->set_param(name, data, numel, static_cast<cudaStream_t>(stream));
}
int ddpm_lpips_finalize(void* handle) {
    if (!handle) { ddpm::set_error("ddpm_lpips_finalize: null handle"); return 2; }
    return static_cast<ddpm::Lpips*>(handle)->finalize();
}
long long ddpm_lpips_workspace_bytes(void* handle, int B, int H, int W) {
    if (!handle) return 0;
    return static_cast<long long>(static_cast<ddpm::Lpips*>(handle)->workspace_bytes(B, H, W));
}
int ddpm_lpips_forward(void* handle, const float* in0, const float* in1, float* out, int B, int C, int H, int W,
                       int normalize, void* workspace, long long workspace_bytes, void* stream) {
    if (!handle || !in0 || !in1 || !out || !workspace) { ddpm::set_error("ddpm_lpips_forward: null argument"); return 2; }
    return static_cast<ddpm::Lpips*>(handle)->forward(in0, in1, out, B, C, H, W, normalize != 0, workspace,
                                                      static_cast<size_t>(workspace_bytes),
                                                      static_cast<cudaStream_t>(stream));
}
long long ddpm_lpips_launch_count(void* handle) { return handle ? static_cast<ddpm::Lpips*>(handle)->launches() : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ VQ-VAE
#include "vqvae.cuh"

static_assert(DDPM_MAX_LEVELS == ddpm::kVqMaxLevels, "level capacity mismatch");

extern "C" {

int ddpm_vqvae_create(const ddpm_vqvae_config* cfg, void** handle) {
    if (!cfg || !handle) { ddpm::set_error("ddpm_vqvae_create: null argument"); return 2; }
    ddpm::VqVaeConfig c{};
    c.spatial_dims = cfg->spatial_dims;
    c.in_channels = cfg->in_channels;
    c.out_channels = cfg->out_channels;
    c.num_levels = cfg->num_levels;
    c.num_res_layers = cfg->num_res_layers;
    for (int i = 0; i < DDPM_MAX_LEVELS; ++i) {
        c.num_channels[i] = cfg->num_channels[i];
        c.num_res_channels[i] = cfg->num_res_channels[i];
    }
    c.num_embeddings = cfg->num_embeddings;
    c.embedding_dim = cfg->embedding_dim;
    c.precise_encode = cfg->precise_encode;
    ddpm::VqVae* v = new ddpm::VqVae(c);
    int rc = v->init();
    if (rc) { delete v; *handle = nullptr; return rc; }
    *handle = v;
    return 0;
}
void ddpm_vqvae_destroy(void* handle) { delete static_cast<ddpm::VqVae*>(handle); }
int ddpm_vqvae_set_param(void* handle, const char* name, const float* data, long long numel, void* stream) {
    if (!handle || !name || !data) { ddpm::set_error("ddpm_vqvae_set_param: null argument"); return 2; }
    return static_cast<ddpm::VqVae*>(handle)->set_param(name, data, numel, static_cast<cudaStream_t>(stream));
}
int ddpm_vqvae_finalize(void* handle, void* stream) {
    if (!handle) { ddpm::set_error("ddpm_vqvae_finalize: null handle"); return 2; }
    return static_cast<ddpm::VqVae*>(handle)->finalize(static_cast<cudaStream_t>(stream));
}
long long ddpm_vqvae_workspace_bytes(void* handle, int N, int D, int H, int W) {
    if (!handle) { ddpm::set_error("ddpm_vqvae_workspace_bytes: null handle"); return 0; }
    return static_cast<long long>(static_cast<ddpm::VqVae*>(handle)->workspace_bytes(N, D, H, W));
}
int ddpm_vqvae_encode(void* handle, const float* x, float* latent, int* indices, int N, int D, int H, int W, void* workspace,
                      long long workspace_bytes, void* stream) {
    if (!handle || !x || !latent || !workspace) { ddpm::set_error("ddpm_vqvae_encode: null argument"); return 2; }
    return static_cast<ddpm::VqVae*>(handle)->encode(x, latent, indices, N, D, H, W, workspace,
                                                     static_cast<size_t>(workspace_bytes), static_cast<cudaStream_t>(stream));
}
int ddpm_vqvae_decode(void* handle, const float* z, const int* indices_in, float* image, int* indices_out, int N, int D, int H,
                      int W, void* workspace, long long workspace_bytes, void* stream) {
    if (!handle || !image || !workspace) { ddpm::set_error("ddpm_vqvae_decode: null argument"); return 2; }
    return static_cast<ddpm::VqVae*>(handle)->decode(z, indices_in, image, indices_out, N, D, H, W, workspace,
                                                     static_cast<size_t>(workspace_bytes), static_cast<cudaStream_t>(stream));
}
long long ddpm_vqvae_launch_count(void* handle) { return handle ? static_cast<ddpm::VqVae*>(handle)->launches() : 0; }

}  // extern "C"
