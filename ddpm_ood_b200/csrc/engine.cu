// DiffusionModelUNet forward engine. See engine.cuh.
#include "engine.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace ddpm {

int num_sms();  // api.cu

__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

// conv2's packed weights widened by an identity block: out = W z + I x puts the ResnetBlock's identity residual on the
// tensor pipe (fp16 x times 1.0 is exact in the fp32 accumulator) instead of latency-bound loads in the epilogue.
// Dense form of a padded 3x3x3 conv on a 2 x 2 x 2 volume (voxel v = (z * 2 + y) * 2 + x): output voxel vo reads input voxel
// vi through tap (vi - vo + 1) per axis. w: packed fp16 [cout][... | 27 taps x cin at column kcol0 | ...] (row pitch ksrc);
// dst[(vo * cout + co)][kdst0 + vi * cin + ci] = w[co][kcol0 + tap(vo, vi) * cin + ci].
// diag: a 1x1 conv (the ResnetBlock skip conv): dst[(vo, co)][kdst0 + vi * cin + ci] = vi == vo ? w[co][kcol0 + ci] : 0.
__global__ void dense2_pack_kernel(const __half* __restrict__ w, int cout, int cin, long long ksrc, long long kcol0,
                                   __half* __restrict__ dst, long long kdst, long long kdst0, int diag) {
    const long long total = 64LL * cout * cin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cin);
        long long r = i / cin;
        const int vi = static_cast<int>(r % 8); r /= 8;
        const int co = static_cast<int>(r % cout);
        const int vo = static_cast<int>(r / cout);
        __half v = __float2half(0.f);
        if (diag) {
            if (vi == vo) v = w[co * ksrc + kcol0 + ci];
        } else {
            const int dz = (vi >> 2) - (vo >> 2) + 1, dy = ((vi >> 1) & 1) - ((vo >> 1) & 1) + 1, dx = (vi & 1) - (vo & 1) + 1;
            v = w[co * ksrc + kcol0 + static_cast<long long>((dz * 3 + dy) * 3 + dx) * cin + ci];
        }
        dst[(static_cast<long long>(vo) * cout + co) * kdst + kdst0 + static_cast<long long>(vi) * cin + ci] = v;
    }
}
__global__ void tile_vec_kernel(const float* __restrict__ src, int n, int reps, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * reps) dst[i] = src[i % n];
}

__global__ void widen_with_identity_kernel(const __half* __restrict__ w, int cout, long long k, __half* __restrict__ dst) {
    const long long kw = k + cout;
    const long long total = static_cast<long long>(cout) * kw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / kw, c = i - r * kw;
        dst[i] = c < k ? w[r * k + c] : __float2half_rn(c - k == r ? 1.f : 0.f);
    }
}

__global__ void pack_conv_weight_kernel2(const float* __restrict__ w, int Cout, int Cin, int taps,
                                         __half* __restrict__ dst, long long ktot, long long koff) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        const long long r = i / Cin;
        const int tap = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        dst[co * ktot + koff + static_cast<long long>(tap) * Cin + ci] =
            __float2half_rn(w[(static_cast<long long>(co) * Cin + ci) * taps + tap]);
    }
}

__global__ void nhwc_half_to_nchw_kernel(const __half* __restrict__ y, float* __restrict__ out, int C, long long S) {
    __shared__ float tile[32][33];
    const long long n = blockIdx.z;
    const long long s0 = static_cast<long long>(blockIdx.x) * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long s = s0 + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && s < S) ? __half2float(y[(n * S + s) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long s = s0 + threadIdx.x;
        if (c < C && s < S) out[(n * C + c) * S + s] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------ construction
UNet::UNet(const UNetConfig& cfg) : cfg_(cfg), E_(cfg.num_channels[0]) {}

UNet::~UNet() {
    for (auto& kv : chain_graphs_) cudaGraphExecDestroy(kv.second.exec);
    for (void* q : dense2_allocs_) cudaFree(q);
    if (temb_table_) cudaFree(temb_table_);
    if (temb_table_act_) cudaFree(temb_table_act_);
    if (f32_arena_) cudaFree(f32_arena_);
    if (f16_arena_) cudaFree(f16_arena_);
}

template <typename T>
T* UNet::arena_alloc(size_t count, bool half_arena) {
    const size_t aligned = (count + 127) & ~size_t(127);  // 256 B (fp16) / 512 B (fp32) granules keep TMA bases aligned
    if (half_arena) {
        T* p = sizing_ ? nullptr : reinterpret_cast<T*>(f16_arena_ + f16_used_);
        f16_used_ += aligned;
        return p;
    }
    T* p = sizing_ ? nullptr : reinterpret_cast<T*>(f32_arena_ + f32_used_);
    f32_used_ += aligned;
    return p;
}

void UNet::add_copy(const std::string& name, float* dst, long long numel) {
    if (sizing_) return;
    ParamSlot s{};
    s.kind = ParamSlot::COPY_F32;
    s.dst = dst;
    s.numel = numel;
    slots_[name] = s;
}
void UNet::add_pack(const std::string& name, __half* dst, int Cout, int Cin, int taps, long long ktot,
                    long long koff) {
    if (sizing_) return;
    ParamSlot s{};
    s.kind = ParamSlot::PACK_CONV;
    s.dst = dst;
    s.numel = static_cast<long long>(Cout) * Cin * taps;
    s.Cout = Cout; s.Cin = Cin; s.taps = taps; s.ktot = ktot; s.koff = koff;
    slots_[name] = s;
}

ResW UNet::make_res(const std::string& prefix, int c0, int c1, int cout) {
    const int taps = cfg_.spatial_dims == 3 ? 27 : 9;
    ResW r{};
    r.prefix = prefix;
    r.c0 = c0; r.c1 = c1; r.cout = cout;
    const int cin = c0 + c1;
    r.skip_conv = (cin != cout);
    r.g1 = arena_alloc<float>(cin, false);
    r.b1 = arena_alloc<float>(cin, false);
    r.g2 = arena_alloc<float>(cout, false);
    r.b2 = arena_alloc<float>(cout, false);
    r.bias1 = arena_alloc<float>(cout, false);
    r.bias2 = arena_alloc<float>(cout, false);
    r.bias_skip = r.skip_conv ? arena_alloc<float>(cout, false) : nullptr;
    r.bias2_total = arena_alloc<float>(cout, false);
    const long long k1 = static_cast<long long>(taps) * cin;
    const long long k2 = static_cast<long long>(taps) * cout + (r.skip_conv ? cin : 0);
    r.w1 = arena_alloc<__half>(static_cast<size_t>(cout) * k1, true);
    r.w2 = arena_alloc<__half>(static_cast<size_t>(cout) * k2, true);
    r.w2_id = (!r.skip_conv && use_halo_ && id_residual_mma_)
                  ? arena_alloc<__half>(static_cast<size_t>(cout) * (k2 + cout), true) : nullptr;
    r.temb_off = P_;
    add_copy(prefix + ".norm1.weight", r.g1, cin);
    add_copy(prefix + ".norm1.bias", r.b1, cin);
    add_copy(prefix + ".norm2.weight", r.g2, cout);
    add_copy(prefix + ".norm2.bias", r.b2, cout);
    add_pack(prefix + ".conv1.conv.weight", r.w1, cout, cin, taps, k1, 0);
    add_copy(prefix + ".conv1.conv.bias", r.bias1, cout);
    add_pack(prefix + ".conv2.conv.weight", r.w2, cout, cout, taps, k2, 0);
    add_copy(prefix + ".conv2.conv.bias", r.bias2, cout);
    if (r.skip_conv) {
        add_pack(prefix + ".skip_connection.conv.weight", r.w2, cout, cin, 1, k2, static_cast<long long>(taps) * cout);
        add_copy(prefix + ".skip_connection.conv.bias", r.bias_skip, cout);
    }
    if (!sizing_) {
        add_copy(prefix + ".time_emb_proj.weight", tp_w_ + static_cast<size_t>(P_) * 4 * E_,
                 static_cast<long long>(cout) * 4 * E_);
        add_copy(prefix + ".time_emb_proj.bias", tp_b_ + P_, cout);
    }
    P_ += cout;
    return r;
}

AttnW UNet::make_attn(const std::string& prefix, int C, int head_channels) {
    AttnW a{};
    a.prefix = prefix;
    a.C = C;
    a.heads = head_channels > 0 ? C / head_channels : 1;
    a.g = arena_alloc<float>(C, false);
    a.b = arena_alloc<float>(C, false);
    a.bqkv = arena_alloc<float>(3 * C, false);
    a.bproj = arena_alloc<float>(C, false);
    a.wqkv = arena_alloc<__half>(static_cast<size_t>(3) * C * C, true);
    a.wproj = arena_alloc<__half>(static_cast<size_t>(C) * C, true);
    add_copy(prefix + ".norm.weight", a.g, C);
    add_copy(prefix + ".norm.bias", a.b, C);
    const char* names[3] = {"to_q", "to_k", "to_v"};
    for (int i = 0; i < 3; ++i) {
        add_pack(prefix + "." + names[i] + ".weight", sizing_ ? nullptr : a.wqkv + static_cast<size_t>(i) * C * C, C, C,
                 1, C, 0);
        add_copy(prefix + "." + names[i] + ".bias", sizing_ ? nullptr : a.bqkv + i * C, C);
    }
    add_pack(prefix + ".proj_attn.weight", a.wproj, C, C, 1, C, 0);
    add_copy(prefix + ".proj_attn.bias", a.bproj, C);
    return a;
}

int UNet::init() {
    const UNetConfig& c = cfg_;
    if (c.num_levels < 1 || c.num_levels > kMaxLevels) { set_error("unet: num_levels=%d unsupported", c.num_levels); return 2; }
    if (c.spatial_dims != 2 && c.spatial_dims != 3) { set_error("unet: spatial_dims=%d unsupported", c.spatial_dims); return 2; }
    for (int i = 0; i < c.num_levels; ++i) {
        if (c.num_channels[i] % 64 != 0) { set_error("unet: num_channels must be multiples of 64 (tcgen05 K blocks)"); return 2; }
        if (c.num_channels[i] % 128 != 0) { set_error("unet: num_channels must be multiples of 128 (tcgen05 N tiles)"); return 2; }
    }
    const int taps = c.spatial_dims == 3 ? 27 : 9;
    fuse_gn_stats_ = c.norm_num_groups > 0;
    for (int i = 0; i < c.num_levels && fuse_gn_stats_; ++i)
        fuse_gn_stats_ = (c.num_channels[i] % c.norm_num_groups == 0) && ((c.num_channels[i] / c.norm_num_groups) % 4 == 0);
    if (const char* e = getenv("DDPM_UPCONV_PHASES")) upconv_phases_ = atoi(e) != 0;  // A/B switch for tests
    if (const char* e = getenv("DDPM_ATTN_TC")) use_attn_tc_ = atoi(e) != 0;  // A/B switch for tests
    if (const char* e = getenv("DDPM_FUSE_GN")) fuse_gn_stats_ = fuse_gn_stats_ && atoi(e) != 0;  // A/B switch for tests
    if (const char* e = getenv("DDPM_CONV_HALO")) use_halo_ = atoi(e) != 0;  // A/B switch for tests
    if (const char* e = getenv("DDPM_HALO_GN_IN_KERNEL")) halo_gn_in_kernel_ = atoi(e) != 0;  // A/B switch for tests
    if (const char* e = getenv("DDPM_ID_RESIDUAL_MMA")) id_residual_mma_ = atoi(e) != 0;     // A/B switch for tests
    if (const char* e = getenv("DDPM_ATTN_FUSED")) use_attn_fused_ = atoi(e) != 0;            // A/B switch for tests
    use_halo_ = use_halo_ && fuse_gn_stats_;  // 2-D: every stride-1 3x3 conv; 3-D: the levels whose maps are 8 x 8 slabs
    in_gemm_ = (c.in_channels % 64 == 0);
    out_gemm_ = (c.out_channels % 128 == 0);
    if (!in_gemm_ && c.in_channels > 8) { set_error("unet: in_channels=%d unsupported (<=8 or multiple of 64)", c.in_channels); return 2; }
    if (!out_gemm_ && c.out_channels > 8) { set_error("unet: out_channels=%d unsupported (<=8 or multiple of 128)", c.out_channels); return 2; }

    for (int pass = 0; pass < 2; ++pass) {
        sizing_ = (pass == 0);
        f32_used_ = f16_used_ = 0;
        P_ = 0;
        down_.clear();
        up_.clear();
        slots_.clear();
        if (!sizing_) {
            if (cudaMalloc(&f32_arena_, f32_count_ * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&f16_arena_, f16_count_ * sizeof(__half)) != cudaSuccess) {
                set_error("unet: cudaMalloc of weight arenas failed (%zu + %zu bytes): %s", f32_count_ * 4, f16_count_ * 2,
                          cudaGetErrorString(cudaGetLastError()));
                return 6;
            }
            cudaMemset(f32_arena_, 0, f32_count_ * sizeof(float));
            cudaMemset(f16_arena_, 0, f16_count_ * sizeof(__half));
        }
        const int C0 = c.num_channels[0];
        const int TE = 4 * E_;
        // conv_in
        if (in_gemm_) {
            conv_in_wp_ = arena_alloc<__half>(static_cast<size_t>(C0) * taps * c.in_channels, true);
            add_pack("conv_in.conv.weight", conv_in_wp_, C0, c.in_channels, taps, static_cast<long long>(taps) * c.in_channels, 0);
        } else {
            conv_in_w_ = arena_alloc<float>(static_cast<size_t>(C0) * taps * c.in_channels, false);
            add_copy("conv_in.conv.weight", conv_in_w_, static_cast<long long>(C0) * taps * c.in_channels);
        }
        conv_in_b_ = arena_alloc<float>(C0, false);
        add_copy("conv_in.conv.bias", conv_in_b_, C0);
        // time embedding
        te_w0_ = arena_alloc<float>(static_cast<size_t>(TE) * E_, false);
        te_b0_ = arena_alloc<float>(TE, false);
        te_w1_ = arena_alloc<float>(static_cast<size_t>(TE) * TE, false);
        te_b1_ = arena_alloc<float>(TE, false);
        add_copy("time_embed.0.weight", te_w0_, static_cast<long long>(TE) * E_);
        add_copy("time_embed.0.bias", te_b0_, TE);
        add_copy("time_embed.2.weight", te_w1_, static_cast<long long>(TE) * TE);
        add_copy("time_embed.2.bias", te_b1_, TE);
        // total projection width is known only after the walk; size generously on the sizing pass
        int ptotal = 0;
        {
            for (int i = 0; i < c.num_levels; ++i) ptotal += c.num_res_blocks[i] * c.num_channels[i];
            ptotal += 2 * c.num_channels[c.num_levels - 1];
            for (int i = 0; i < c.num_levels; ++i) ptotal += (c.num_res_blocks[c.num_levels - 1 - i] + 1) * c.num_channels[c.num_levels - 1 - i];
        }
        tp_w_ = arena_alloc<float>(static_cast<size_t>(ptotal) * TE, false);
        tp_b_ = arena_alloc<float>(ptotal, false);
        // down path
        int oc = C0;
        for (int i = 0; i < c.num_levels; ++i) {
            const int ic = oc;
            oc = c.num_channels[i];
            Level L{};
            for (int j = 0; j < c.num_res_blocks[i]; ++j) {
                const std::string pre = "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
                L.res.push_back(make_res(pre, j == 0 ? ic : oc, 0, oc));
                if (c.attention_levels[i])
                    L.attn.push_back(make_attn("down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), oc,
                                               c.num_head_channels[i]));
            }
            L.has_samp = (i != c.num_levels - 1);
            if (L.has_samp) {
                L.samp.C = oc;
                L.samp.w = arena_alloc<__half>(static_cast<size_t>(oc) * taps * oc, true);
                L.samp.bias = arena_alloc<float>(oc, false);
                const std::string pre = "down_blocks." + std::to_string(i) + ".downsampler.op.conv";
                add_pack(pre + ".weight", L.samp.w, oc, oc, taps, static_cast<long long>(taps) * oc, 0);
                add_copy(pre + ".bias", L.samp.bias, oc);
            }
            down_.push_back(std::move(L));
        }
        // middle
        const int CM = c.num_channels[c.num_levels - 1];
        mid1_ = make_res("middle_block.resnet_1", CM, 0, CM);
        mid_attn_ = make_attn("middle_block.attention", CM, c.num_head_channels[c.num_levels - 1]);
        mid2_ = make_res("middle_block.resnet_2", CM, 0, CM);
        // up path
        oc = CM;
        for (int i = 0; i < c.num_levels; ++i) {
            const int lvl = c.num_levels - 1 - i;  // un-reversed level index
            const int prev = oc;
            oc = c.num_channels[lvl];
            const int ic = c.num_channels[lvl > 0 ? lvl - 1 : 0];  // reversed[min(i+1, L-1)]
            const int nres = c.num_res_blocks[lvl] + 1;
            Level L{};
            for (int j = 0; j < nres; ++j) {
                const int res_skip = (j == nres - 1) ? ic : oc;
                const int res_in = (j == 0) ? prev : oc;
                const std::string pre = "up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
                L.res.push_back(make_res(pre, res_in, res_skip, oc));
                if (c.attention_levels[lvl])
                    L.attn.push_back(make_attn("up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), oc,
                                               c.num_head_channels[lvl]));
            }
            L.has_samp = (i != c.num_levels - 1);
            if (L.has_samp) {
                L.samp.C = oc;
                L.samp.w = arena_alloc<__half>(static_cast<size_t>(oc) * taps * oc, true);
                const int ph = 1 << c.spatial_dims;
                L.samp.w_up = arena_alloc<__half>(static_cast<size_t>(ph) * oc * ph * oc, true);
                L.samp.bias = arena_alloc<float>(oc, false);
                const std::string pre = "up_blocks." + std::to_string(i) + ".upsampler.conv.conv";
                add_pack(pre + ".weight", L.samp.w, oc, oc, taps, static_cast<long long>(taps) * oc, 0);
                if (!sizing_) {
                    ParamSlot& sl = slots_[pre + ".weight"];
                    sl.kind = ParamSlot::PACK_UPCONV;
                    sl.dst2 = L.samp.w_up;
                }
                add_copy(pre + ".bias", L.samp.bias, oc);
            }
            up_.push_back(std::move(L));
        }
        // out
        out_g_ = arena_alloc<float>(C0, false);
        out_b_ = arena_alloc<float>(C0, false);
        add_copy("out.0.weight", out_g_, C0);
        add_copy("out.0.bias", out_b_, C0);
        if (out_gemm_) {
            conv_out_wp_ = arena_alloc<__half>(static_cast<size_t>(c.out_channels) * taps * C0, true);
            add_pack("out.2.conv.weight", conv_out_wp_, c.out_channels, C0, taps, static_cast<long long>(taps) * C0, 0);
        } else {
            conv_out_w_ = arena_alloc<float>(static_cast<size_t>(c.out_channels) * taps * C0, false);
            add_copy("out.2.conv.weight", conv_out_w_, static_cast<long long>(c.out_channels) * taps * C0);
        }
        conv_out_b_ = arena_alloc<float>(c.out_channels, false);
        add_copy("out.2.conv.bias", conv_out_b_, c.out_channels);
        if (sizing_) {
            f32_count_ = f32_used_;
            f16_count_ = f16_used_;
        }
        if (P_ != ptotal) { set_error("unet: internal projection width mismatch %d vs %d", P_, ptotal); return 7; }
    }
    return 0;
}

int UNet::set_param(const char* name, const float* data, long long numel, cudaStream_t stream) {
    auto it = slots_.find(name);
    if (it == slots_.end()) { set_error("unet: unexpected parameter '%s'", name); return 8; }
    ParamSlot& s = it->second;
    if (numel != s.numel) { set_error("unet: parameter '%s' has %lld elements, expected %lld", name, numel, s.numel); return 8; }
    if (s.kind == ParamSlot::COPY_F32) {
        cudaError_t e = cudaMemcpyAsync(s.dst, data, numel * sizeof(float), cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) { set_error("unet: copy of '%s' failed: %s", name, cudaGetErrorString(e)); return 5; }
    } else {
        long long blocks = (numel + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        pack_conv_weight_kernel2<<<static_cast<int>(blocks), 256, 0, stream>>>(data, s.Cout, s.Cin, s.taps,
                                                                               static_cast<__half*>(s.dst), s.ktot, s.koff);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("unet: pack of '%s' failed: %s", name, cudaGetErrorString(e)); return 5; }
        if (s.kind == ParamSlot::PACK_UPCONV) {
            int rc = pack_upconv_weight(data, s.Cout, s.Cin, cfg_.spatial_dims, static_cast<__half*>(s.dst2), stream);
            if (rc) return rc;
        }
    }
    s.set = true;
    finalized_ = false;
    return 0;
}

int UNet::finalize(cudaStream_t stream) {
    for (auto& kv : slots_) {
        if (!kv.second.set) { set_error("unet: parameter '%s' was never set", kv.first.c_str()); return 8; }
    }
    // per ResnetBlock: conv2 bias + skip-conv bias in one vector; identity-residual blocks get conv2's weights | I
    auto fold_and_widen = [&](ResW& r) {
        add_vec_kernel<<<(r.cout + 255) / 256, 256, 0, stream>>>(r.bias2, r.bias_skip, r.bias2_total, r.cout);
        if (r.w2_id) {
            const long long k2 = static_cast<long long>(cfg_.spatial_dims == 3 ? 27 : 9) * r.cout;
            widen_with_identity_kernel<<<148 * 4, 256, 0, stream>>>(r.w2, r.cout, k2, r.w2_id);
        }
    };
    // 3-D models: dense-form weights for whichever ResnetBlocks end up on 2 x 2 x 2 maps (the third level of an 8^3
    // latent); built for every block (<= 25 MB each) because the level sizes depend on the input, not on the weights
    int drc = 0;
    auto dense2 = [&](ResW& r) {
        if (cfg_.spatial_dims != 3 || drc) return;
        const int cin = r.c0 + r.c1;
        const long long k1 = 8LL * cin, k2 = 8LL * r.cout + (r.skip_conv ? 8LL * cin : 0);
        const long long ksrc1 = 27LL * cin, ksrc2 = 27LL * r.cout + (r.skip_conv ? cin : 0);
        if (!r.w1d) {
            void* q[4] = {nullptr, nullptr, nullptr, nullptr};
            if (cudaMalloc(&q[0], 8 * r.cout * k1 * sizeof(__half)) != cudaSuccess ||
                cudaMalloc(&q[1], 8 * r.cout * k2 * sizeof(__half)) != cudaSuccess ||
                cudaMalloc(&q[2], 8 * r.cout * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&q[3], 8 * r.cout * sizeof(float)) != cudaSuccess) {
                set_error("unet: cudaMalloc of the dense 2x2x2 weights failed");
                drc = 6;
                return;
            }
            for (void* x : q) dense2_allocs_.push_back(x);
            r.w1d = static_cast<__half*>(q[0]); r.w2d = static_cast<__half*>(q[1]);
            r.bias1d = static_cast<float*>(q[2]); r.bias2d = static_cast<float*>(q[3]);
        }
        dense2_pack_kernel<<<148 * 8, 256, 0, stream>>>(r.w1, r.cout, cin, ksrc1, 0, r.w1d, k1, 0, 0);
        dense2_pack_kernel<<<148 * 8, 256, 0, stream>>>(r.w2, r.cout, r.cout, ksrc2, 0, r.w2d, k2, 0, 0);
        if (r.skip_conv) {  // the 1x1 skip conv over cat(h, skip): two block-diagonal K segments after conv2's
            dense2_pack_kernel<<<148 * 8, 256, 0, stream>>>(r.w2, r.cout, r.c0, ksrc2, 27LL * r.cout, r.w2d, k2, 8LL * r.cout, 1);
            if (r.c1)
                dense2_pack_kernel<<<148 * 8, 256, 0, stream>>>(r.w2, r.cout, r.c1, ksrc2, 27LL * r.cout + r.c0, r.w2d, k2,
                                                                8LL * r.cout + 8LL * r.c0, 1);
        }
        tile_vec_kernel<<<(8 * r.cout + 255) / 256, 256, 0, stream>>>(r.bias1, r.cout, 8, r.bias1d);
        tile_vec_kernel<<<(8 * r.cout + 255) / 256, 256, 0, stream>>>(r.bias2_total, r.cout, 8, r.bias2d);
    };
    auto per_block = [&](ResW& r) { fold_and_widen(r); dense2(r); };
    for (auto& L : down_) for (auto& r : L.res) per_block(r);
    per_block(mid1_);
    per_block(mid2_);
    for (auto& L : up_) for (auto& r : L.res) per_block(r);
    if (drc) return drc;
    // timestep-embedding table: row t = all time_emb_proj outputs for timestep t (depends on weights only)
    if (!temb_table_) {
        if (cudaMalloc(&temb_table_, static_cast<size_t>(temb_rows_) * P_ * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&temb_table_act_, static_cast<size_t>(temb_rows_) * 4 * E_ * sizeof(float)) != cudaSuccess) {
            set_error("unet: cudaMalloc of the timestep-embedding table failed");
            return 6;
        }
    }
    int trc = time_embed(nullptr, 0, temb_rows_, E_, te_w0_, te_b0_, te_w1_, te_b1_, temb_table_act_, stream);
    if (!trc) trc = time_proj_all(temb_table_act_, temb_rows_, 4 * E_, tp_w_, tp_b_, P_, temb_table_, stream);
    if (trc) return trc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("unet: finalize failed: %s", cudaGetErrorString(e)); return 5; }
    finalized_ = true;
    plans_.clear();
    return 0;
}

// ------------------------------------------------------------------------------------------------ planning
static double conv_flops(const ConvProblem& q) {
    const int sd3 = q.spatial_dims == 3;
    const double Wo = (q.W + q.stride - 1) / q.stride, Ho = (q.H + q.stride - 1) / q.stride;
    const double Do = sd3 ? (q.D + q.stride - 1) / q.stride : q.D;
    double k = 0;
    for (int s = 0; s < q.n_seg; ++s) {
        const int taps = q.seg[s].ksize == 3 ? (sd3 ? 27 : 9) : (q.seg[s].ksize == 2 ? (sd3 ? 8 : 4) : 1);
        k += static_cast<double>(taps) * q.seg[s].channels;
    }
    const double phases = q.upsample2 ? (sd3 ? 8 : 4) : 1;  // executed MACs (4/9 resp. 8/27 of the reference op's)
    return 2.0 * q.N * Do * Ho * Wo * phases * q.Cout * k;
}

struct UNet::Layout {
    uint8_t* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t count) {
        off = (off + 1023) & ~size_t(1023);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

int UNet::build_plan(Plan& plan, int N, int D, int H, int W, void* ws, size_t ws_bytes, bool dry, size_t* need) const {
    const UNetConfig& c = cfg_;
    const int sd = c.spatial_dims;
    const int sms = num_sms();
    Layout lay{dry ? nullptr : static_cast<uint8_t*>(ws)};
    // stats/parts: GroupNorm partial statistics written by the tensor's producer (null/0 when unsupported)
    struct Act { __half* p; int C, D, H, W; float* stats; int parts; long long S() const { return static_cast<long long>(D) * H * W; } };
    auto take_stats = [&](int C, int parts) -> float* {
        return parts > 0 ? lay.take<float>(static_cast<size_t>(N) * parts * (C / 4) * 2) : nullptr;
    };
    auto new_act = [&](int C, int d, int h, int w, bool want_stats = true) {
        Act a{lay.take<__half>(static_cast<size_t>(N) * d * h * w * C), C, d, h, w, nullptr, 0};
        if (want_stats && fuse_gn_stats_) {
            a.parts = conv_stats_parts(sd, d, h, w);
            a.stats = take_stats(C, a.parts);  // null in the sizing (dry) pass; `parts` is what decisions key on
        }
        return a;
    };
    auto shape_act = [&](int C, int d, int h, int w) { return Act{nullptr, C, d, h, w, nullptr, 0}; };
    // scratch sized by a first dry walk: compute maxima analytically while walking (allocate lazily at the end is not
    // possible with a bump allocator, so walk twice: first to find the maxima, then to lay out).
    size_t max_z = 0, max_h = 0, max_qkv = 0, max_up = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const bool measure = (pass == 0);
        Layout saved = lay;
        __half *zA = nullptr, *zB = nullptr, *hB = nullptr, *qkv = nullptr;
        plan.ops.clear();
        if (!measure) {
            plan.temb_act = lay.take<float>(static_cast<size_t>(N) * 4 * E_);
            plan.temb_all = lay.take<float>(static_cast<size_t>(N) * P_);
            zA = lay.take<__half>(max_z);
            zB = lay.take<__half>(max_h);
            hB = lay.take<__half>(max_h);
            qkv = lay.take<__half>(max_qkv);
            plan.x_half = in_gemm_ ? lay.take<__half>(static_cast<size_t>(N) * D * H * W * c.in_channels) : nullptr;
            plan.y_half = out_gemm_ ? lay.take<__half>(static_cast<size_t>(N) * D * H * W * c.out_channels) : nullptr;
            // fp32 eps of the PLMS corrector step (counter == 1), which is blended but never enters the history ring
            plan.eps_tmp = out_gemm_ ? lay.take<float>(static_cast<size_t>(N) * D * H * W * c.out_channels) : nullptr;
        }
        int rc = 0;
        auto gn = [&](const Act& a, const Act* b, const float* g, const float* bt, __half* dst, bool silu) {
            const size_t cnt = static_cast<size_t>(N) * a.S() * (a.C + (b ? b->C : 0));
            if (measure) { if (cnt > max_z) max_z = cnt; return; }
            Op op{};
            op.type = Op::GN;
            if (a.parts > 0 && (!b || b->parts > 0)) {
                op.st0 = a.stats; op.parts0 = a.parts;
                op.st1 = b ? b->stats : nullptr; op.parts1 = b ? b->parts : 0;
            }
            op.src0 = a.p; op.C0 = a.C;
            op.src1 = b ? b->p : nullptr; op.C1 = b ? b->C : 0;
            op.gamma = g; op.beta = bt; op.dst = dst; op.S = static_cast<int>(a.S()); op.silu = silu;
            op.bytes = 4.0 * static_cast<double>(cnt);  // one fp16 read + one fp16 write per element
            plan.ops.push_back(op);
        };
        auto gemm = [&](ConvProblem q, int temb_off) {
            const bool uses_temb = temb_off >= 0;
            if (measure || dry) { if (!measure) { Op op{}; op.type = Op::GEMM; op.uses_temb = uses_temb; plan.ops.push_back(op); } return; }
            Op op{};
            op.type = Op::GEMM;
            op.uses_temb = uses_temb;
            op.temb_off = temb_off;
            op.flops = conv_flops(q);
            int r = conv_prepare(q, sms, &op.conv);
            if (r && !rc) rc = r;
            plan.ops.push_back(op);
        };
        auto conv3 = [&](const Act& in, __half* zin, int cin, const __half* w, const float* bias, int cout, int stride,
                         int temb_off, const __half* residual, __half* out, const Act* raw0, const Act* raw1,
                         float* stats_out) {
            const float* cadd = (temb_off >= 0 && plan.temb_all) ? plan.temb_all + temb_off : nullptr;
            ConvProblem q{};
            q.spatial_dims = sd;
            q.N = N; q.D = in.D; q.H = in.H; q.W = in.W;
            q.stride = stride;
            q.n_seg = 1;
            q.seg[0] = {zin, cin, 3};
            if (raw0) { q.seg[q.n_seg++] = {raw0->p, raw0->C, 1}; }
            if (raw1) { q.seg[q.n_seg++] = {raw1->p, raw1->C, 1}; }
            q.weights = w; q.w_rows = cout; q.Cout = cout;
            q.mode = EPI_STORE;
            q.bias = bias; q.chan_add = cadd; q.chan_add_stride = P_;
            q.residual = residual; q.out = out;
            q.stats_out = stats_out;
            gemm(q, temb_off);
        };
        // 1x1 conv over the real image geometry (same GEMM as `linear`, but tiles are image-structured so the epilogue
        // can emit per-image GroupNorm statistics)
        auto conv1x1 = [&](const Act& in, const __half* src, const __half* w, const float* bias, int cout,
                           const __half* residual, const Act& out) {
            ConvProblem q{};
            q.spatial_dims = sd;
            q.N = N; q.D = in.D; q.H = in.H; q.W = in.W;
            q.stride = 1;
            q.n_seg = 1;
            q.seg[0] = {src, in.C, 1};
            q.weights = w; q.w_rows = cout; q.Cout = cout;
            q.mode = EPI_STORE;
            q.bias = bias; q.residual = residual; q.out = out.p;
            q.stats_out = out.stats;
            gemm(q, -1);
        };
        auto linear = [&](const __half* in, long long rows, int cin, const __half* w, const float* bias, int cout,
                          const __half* residual, __half* out) {
            ConvProblem q{};
            q.spatial_dims = 2;
            q.N = 1; q.D = 1; q.H = 1; q.W = static_cast<int>(rows);
            q.stride = 1;
            q.n_seg = 1;
            q.seg[0] = {in, cin, 1};
            q.weights = w; q.w_rows = cout; q.Cout = cout;
            q.mode = EPI_STORE;
            q.bias = bias; q.residual = residual; q.out = out;
            gemm(q, -1);
        };
        // ResnetBlock on the halo-tile kernel: both GroupNorm+SiLU layers are applied inside the consuming conv (the
        // normalised tensors never exist), each preceded by a tiny statistics -> scale/shift kernel.
        auto halo_ok = [&](const Act& h, const Act* skip, int cout, int gn2_channels = 0) {
            if (!use_halo_ || h.parts <= 0 || (skip && skip->parts <= 0)) return false;
            // scale/shift rows the kernel stages in shared memory: conv1 normalises cat(h, skip), conv2 its own input
            if (h.C + (skip ? skip->C : 0) > 512 || gn2_channels > 512) return false;
            ConvProblem q{};
            q.spatial_dims = sd; q.N = N; q.D = h.D; q.H = h.H; q.W = h.W; q.stride = 1; q.n_seg = 1;
            q.seg[0] = {nullptr, h.C, 3};
            q.Cout = cout; q.mode = EPI_STORE;
            return conv_halo_supported(q);
        };
        auto finalize_op = [&](const Act& a, const Act* b, const float* g, const float* bt, float* ab) {
            Op op{};
            op.type = Op::GN_FINALIZE;
            op.st0 = a.stats; op.parts0 = a.parts; op.C0 = a.C;
            op.st1 = b ? b->stats : nullptr; op.parts1 = b ? b->parts : 0; op.C1 = b ? b->C : 0;
            op.gamma = g; op.beta = bt; op.ab = ab; op.S = static_cast<int>(a.S());
            op.bytes = static_cast<double>(N) * (a.parts * a.C + (b ? b->parts * b->C : 0)) * 2.0;
            plan.ops.push_back(op);
        };
        auto gn_source = [&](const Act& a, const Act* b, const float* g, const float* bt) {
            HaloGnSource src{};
            src.st0 = a.stats; src.parts0 = a.parts; src.C0 = a.C;
            src.st1 = b ? b->stats : nullptr; src.parts1 = b ? b->parts : 0; src.C1 = b ? b->C : 0;
            src.gamma = g; src.beta = bt; src.S = static_cast<int>(a.S()); src.groups = c.norm_num_groups; src.eps = c.norm_eps;
            return src;
        };
        auto halo_conv = [&](ConvProblem q, const float* ab, int ab_channels, int temb_off, const HaloGnSource* src = nullptr,
                             double flops = -1.0) {
            Op op{};
            op.type = Op::CONV_HALO;
            op.uses_temb = temb_off >= 0;
            op.temb_off = temb_off;
            op.flops = flops >= 0.0 ? flops : conv_flops(q);
            if (!dry) {
                int r = conv_halo_prepare(q, ab, ab_channels, sms, &op.halo, src);
                if (r && !rc) rc = r;
            }
            plan.ops.push_back(op);
        };
        auto resblock_halo = [&](const ResW& r, const Act& h, const Act* skip) -> Act {
            const int cin = r.c0 + r.c1;
            if (measure) {
                const size_t ch = static_cast<size_t>(N) * h.S() * r.cout;
                if (ch > max_h) max_h = ch;
                return shape_act(r.cout, h.D, h.H, h.W);
            }
            const int parts = conv_halo_stats_parts(h.H, h.W, h.D);
            // GroupNorm statistics are finalised inside the consuming kernel (halo_gn_in_kernel_) or by a tiny launch
            float* ab1 = halo_gn_in_kernel_ ? nullptr : lay.take<float>(static_cast<size_t>(N) * cin * 2);
            float* ab2 = halo_gn_in_kernel_ ? nullptr : lay.take<float>(static_cast<size_t>(N) * r.cout * 2);
            if (!halo_gn_in_kernel_) finalize_op(h, skip, r.g1, r.b1, ab1);
            Act h1{hB, r.cout, h.D, h.H, h.W, take_stats(r.cout, parts), parts};
            ConvProblem q{};
            q.spatial_dims = sd; q.N = N; q.D = h.D; q.H = h.H; q.W = h.W; q.stride = 1;
            q.n_seg = 1;
            q.seg[0] = {h.p, h.C, 3};
            if (skip) q.seg[q.n_seg++] = {skip->p, skip->C, 3};
            q.concat3x3 = 1;  // conv1's weight is one [cout][9][cin] matrix over cat(h, skip)
            q.gn_silu = 1;
            q.weights = r.w1; q.w_rows = r.cout; q.Cout = r.cout; q.mode = EPI_STORE;
            q.bias = r.bias1; q.chan_add = plan.temb_all ? plan.temb_all + r.temb_off : nullptr; q.chan_add_stride = P_;
            q.out = hB; q.stats_out = h1.stats;
            const HaloGnSource src1 = gn_source(h, skip, r.g1, r.b1);
            halo_conv(q, ab1, cin, r.temb_off, halo_gn_in_kernel_ ? &src1 : nullptr);
            if (!halo_gn_in_kernel_) finalize_op(h1, nullptr, r.g2, r.b2, ab2);
            Act out{lay.take<__half>(static_cast<size_t>(N) * h.S() * r.cout), r.cout, h.D, h.H, h.W,
                    take_stats(r.cout, parts), parts};
            ConvProblem q2{};
            q2.spatial_dims = sd; q2.N = N; q2.D = h.D; q2.H = h.H; q2.W = h.W; q2.stride = 1;
            q2.n_seg = 1;
            q2.seg[0] = {hB, r.cout, 3};
            q2.Cout = r.cout;
            double flops2 = -1.0;
            if (r.skip_conv) {
                q2.seg[q2.n_seg++] = {h.p, h.C, 1};
                if (skip) q2.seg[q2.n_seg++] = {skip->p, skip->C, 1};
            } else if (r.w2_id) {
                flops2 = conv_flops(q2);             // the identity block is not algorithmic work
                q2.seg[q2.n_seg++] = {h.p, h.C, 1};  // identity residual as a K segment against the I block of w2_id
            } else {
                q2.residual = h.p;
            }
            q2.weights = (!r.skip_conv && r.w2_id) ? r.w2_id : r.w2; q2.w_rows = r.cout; q2.Cout = r.cout; q2.mode = EPI_STORE;
            q2.bias = r.bias2_total; q2.out = out.p; q2.stats_out = out.stats;
            q2.gn_silu = 1;
            const HaloGnSource src2 = gn_source(h1, nullptr, r.g2, r.b2);
            halo_conv(q2, ab2, r.cout, -1, halo_gn_in_kernel_ ? &src2 : nullptr, flops2);
            return out;
        };
        // a ResnetBlock conv on a 2 x 2 x 2 map as ONE linear layer over (voxel, channel): rows = images, K = 8 x channels of
        // each input tensor, N = 8 cout (see ResW::w1d). 8 / 27 of the im2col form's MACs and plain 2-D GEMM tiles instead
        // of 128-row boxes gathered from 16 tiny volumes.
        auto dense2_conv = [&](const __half* z, int zc, const Act* raw0, const Act* raw1, const __half* w, const float* bias,
                               int cout, int temb_off, const __half* residual, __half* out, double flops) {
            ConvProblem q{};
            q.spatial_dims = 2; q.N = N; q.D = 1; q.H = 1; q.W = 1; q.stride = 1;
            q.n_seg = 1;
            q.seg[0] = {z, 8 * zc, 1};
            if (raw0) q.seg[q.n_seg++] = {raw0->p, 8 * raw0->C, 1};
            if (raw1) q.seg[q.n_seg++] = {raw1->p, 8 * raw1->C, 1};
            q.weights = w; q.w_rows = 8 * cout; q.Cout = 8 * cout;
            q.mode = EPI_STORE;
            q.bias = bias; q.residual = residual; q.out = out;
            if (temb_off >= 0 && plan.temb_all) { q.chan_add = plan.temb_all + temb_off; q.chan_add_stride = P_; q.chan_mod = cout; }
            if (measure || dry) { gemm(q, temb_off); return; }
            Op op{};
            op.type = Op::GEMM;
            op.uses_temb = temb_off >= 0;
            op.temb_off = temb_off;
            op.flops = flops;
            int r_ = conv_prepare(q, sms, &op.conv);
            if (r_ && !rc) rc = r_;
            plan.ops.push_back(op);
        };
        auto resblock = [&](const ResW& r, const Act& h, const Act* skip) -> Act {
            if (halo_ok(h, skip, r.cout, r.cout)) return resblock_halo(r, h, skip);
            const int cin = r.c0 + r.c1;
            if (measure) {
                const size_t ch = static_cast<size_t>(N) * h.S() * r.cout;
                if (ch > max_h) max_h = ch;
            }
            const bool dense2 = sd == 3 && h.D == 2 && h.H == 2 && h.W == 2 && r.w1d && cin % 8 == 0 && r.cout % 32 == 0;
            gn(h, skip, r.g1, r.b1, zA, true);
            Act h1{hB, r.cout, h.D, h.H, h.W, nullptr, 0};
            if (!measure && fuse_gn_stats_ && !dense2) {
                h1.parts = conv_stats_parts(sd, h.D, h.H, h.W);
                h1.stats = take_stats(r.cout, h1.parts);
            }
            if (dense2) {
                const double f1 = 2.0 * N * 8.0 * r.cout * 27.0 * cin;
                const double f2 = 2.0 * N * 8.0 * r.cout * (27.0 * r.cout + (r.skip_conv ? cin : 0));
                dense2_conv(zA, cin, nullptr, nullptr, r.w1d, r.bias1d, r.cout, r.temb_off, nullptr, hB, f1);
                gn(h1, nullptr, r.g2, r.b2, zB, true);
                Act out = measure ? shape_act(r.cout, h.D, h.H, h.W) : new_act(r.cout, h.D, h.H, h.W, false);
                dense2_conv(zB, r.cout, r.skip_conv ? &h : nullptr, r.skip_conv ? skip : nullptr, r.w2d, r.bias2d, r.cout, -1,
                            r.skip_conv ? nullptr : h.p, out.p, f2);
                return out;
            }
            conv3(h, zA, cin, r.w1, r.bias1, r.cout, 1, r.temb_off, nullptr, hB, nullptr, nullptr, h1.stats);
            gn(h1, nullptr, r.g2, r.b2, zB, true);
            Act out = measure ? shape_act(r.cout, h.D, h.H, h.W) : new_act(r.cout, h.D, h.H, h.W);
            if (r.skip_conv)
                conv3(h, zB, r.cout, r.w2, r.bias2_total, r.cout, 1, -1, nullptr, out.p, &h, skip, out.stats);
            else
                conv3(h, zB, r.cout, r.w2, r.bias2_total, r.cout, 1, -1, h.p, out.p, nullptr, nullptr, out.stats);
            return out;
        };
        auto attnblock = [&](const AttnW& a, const Act& h) -> Act {
            const long long rows = static_cast<long long>(N) * h.S();
            if (measure) {
                const size_t cq = static_cast<size_t>(rows) * 3 * a.C;
                if (cq > max_qkv) max_qkv = cq;
                const size_t ch = static_cast<size_t>(rows) * a.C;
                if (ch > max_h) max_h = ch;
            }
            // One launch per AttentionBlock (attn_block.cu): GroupNorm, q/k/v, softmax, P v, projection and the residual
            // add on 128-row tiles that never leave the SM. C = 256 / one head / T <= 128 tokens (the `small` UNet at
            // 32 x 32, 28 x 28 and 8 x 8 x 8 inputs); other shapes take the four-launch path below.
            if (use_attn_fused_ && attn_block_supported(static_cast<int>(h.S()), a.C, a.heads, c.norm_num_groups)) {
                const int parts = fuse_gn_stats_ ? attn_block_stats_parts(static_cast<int>(h.S())) : 0;
                if (measure) return shape_act(a.C, h.D, h.H, h.W);
                Act out{lay.take<__half>(static_cast<size_t>(rows) * a.C), a.C, h.D, h.H, h.W, take_stats(a.C, parts), parts};
                Op op{};
                op.type = Op::ATTN_BLOCK;
                op.T = static_cast<int>(h.S()); op.C = a.C; op.heads = a.heads;
                const double T = static_cast<double>(h.S());
                op.flops = 2.0 * N * T * a.C * 4.0 * a.C + 4.0 * N * T * T * a.C;  // q, k, v, proj Linears + the two matmuls
                if (!dry) {
                    int r = attn_block_prepare(h.p, out.p, N, op.T, a.C, a.heads, c.norm_num_groups, c.norm_eps,
                                               1.0f / sqrtf(static_cast<float>(a.C) / static_cast<float>(a.heads)), a.g, a.b,
                                               a.wqkv, a.bqkv, a.wproj, a.bproj, out.stats, sms, &op.attn_block);
                    if (r && !rc) rc = r;
                }
                plan.ops.push_back(op);
                return out;
            }
            gn(h, nullptr, a.g, a.b, zA, false);
            linear(zA, rows, a.C, a.wqkv, a.bqkv, 3 * a.C, nullptr, qkv);
            if (!measure) {
                Op op{};
                op.type = Op::ATTN;
                op.src0 = qkv; op.dst = hB; op.T = static_cast<int>(h.S()); op.C = a.C; op.heads = a.heads;
                op.scale = 1.0f / sqrtf(static_cast<float>(a.C) / static_cast<float>(a.heads));
                op.flops = 4.0 * N * static_cast<double>(h.S()) * static_cast<double>(h.S()) * a.C;
                op.attn_tc = !dry && use_attn_tc_ && attention_tc_supported(op.T, op.C, op.heads);
                if (op.attn_tc) {
                    int r = attention_tc_prepare(qkv, hB, N, op.T, op.C, op.heads, op.scale, &op.attn);
                    if (r && !rc) rc = r;
                }
                plan.ops.push_back(op);
            }
            Act out = measure ? shape_act(a.C, h.D, h.H, h.W) : new_act(a.C, h.D, h.H, h.W);
            conv1x1(h, hB, a.wproj, a.bproj, a.C, h.p, out);
            return out;
        };

        // ---- walk
        ConvProblem qin{};  // conv_in of wide-channel inputs (latents): one 3x3(x3) conv over the NDHWC fp16 copy of x
        qin.spatial_dims = sd; qin.N = N; qin.D = D; qin.H = H; qin.W = W; qin.stride = 1; qin.n_seg = 1;
        qin.seg[0] = {nullptr, c.in_channels, 3};
        qin.Cout = c.num_channels[0]; qin.mode = EPI_STORE;
        const bool in_halo = in_gemm_ && use_halo_ && c.in_channels <= 512 && conv_halo_supported(qin);
        Act h = measure ? shape_act(c.num_channels[0], D, H, W) : new_act(c.num_channels[0], D, H, W, in_gemm_ && !in_halo);
        if (!measure && in_halo) {
            h.parts = conv_halo_stats_parts(H, W, D);
            h.stats = take_stats(c.num_channels[0], h.parts);
        }
        if (!measure && !in_gemm_ && fuse_gn_stats_ && conv_in_has_stats(c.in_channels, c.num_channels[0], sd)) {
            h.parts = conv_in_stats_parts(D, H, W);
            h.stats = take_stats(c.num_channels[0], h.parts);
        }
        if (!measure) {
            Op op{};
            op.type = in_gemm_ ? Op::CONV_IN_GEMM : Op::CONV_IN_SMALL;
            op.dst = h.p; op.D = D; op.H = H; op.W = W;
            op.st0 = h.stats;
            op.flops = 2.0 * N * D * H * W * c.num_channels[0] * (sd == 3 ? 27.0 : 9.0) * c.in_channels;
            op.bytes = static_cast<double>(N) * D * H * W * (4.0 * c.in_channels + 2.0 * c.num_channels[0]);
            op.on_halo = in_halo;
            if (in_gemm_ && !dry) {
                ConvProblem q = qin;
                q.seg[0].ptr = plan.x_half;
                q.weights = conv_in_wp_; q.w_rows = c.num_channels[0];
                q.bias = conv_in_b_; q.out = h.p;
                q.stats_out = h.stats;
                int r = in_halo ? conv_halo_prepare(q, nullptr, 0, sms, &op.halo) : conv_prepare(q, sms, &op.conv);
                if (r && !rc) rc = r;
            }
            plan.ops.push_back(op);
        }
        std::vector<Act> skips;
        skips.push_back(h);
        for (size_t i = 0; i < down_.size(); ++i) {
            const Level& L = down_[i];
            for (size_t j = 0; j < L.res.size(); ++j) {
                h = resblock(L.res[j], h, nullptr);
                if (!L.attn.empty()) h = attnblock(L.attn[j], h);
                skips.push_back(h);
            }
            if (L.has_samp) {
                const int d2 = sd == 3 ? (h.D + 1) / 2 : h.D, h2 = (h.H + 1) / 2, w2 = (h.W + 1) / 2;
                Act o = measure ? shape_act(h.C, d2, h2, w2) : new_act(h.C, d2, h2, w2);
                conv3(h, h.p, h.C, L.samp.w, L.samp.bias, h.C, 2, -1, nullptr, o.p, nullptr, nullptr, o.stats);
                h = o;
                skips.push_back(h);
            }
        }
        h = resblock(mid1_, h, nullptr);
        h = attnblock(mid_attn_, h);
        h = resblock(mid2_, h, nullptr);
        for (size_t i = 0; i < up_.size(); ++i) {
            const Level& L = up_[i];
            for (size_t j = 0; j < L.res.size(); ++j) {
                Act sk = skips.back();
                skips.pop_back();
                if (sk.D != h.D || sk.H != h.H || sk.W != h.W) {
                    set_error("unet: spatial size %dx%dx%d is not divisible by 2^%d (use --latent_pad)", D, H, W, c.num_levels - 1);
                    return 2;
                }
                h = resblock(L.res[j], h, &sk);
                if (!L.attn.empty()) h = attnblock(L.attn[j], h);
            }
            if (L.has_samp) {
                const int fd = sd == 3 ? 2 : 1;
                if (upconv_phases_) {
                    // nearest x2 + 3x3 conv == 2^d sub-pixel 2x2 convs over the low-res tensor (no upsampled tensor)
                    ConvProblem q{};
                    q.spatial_dims = sd;
                    q.N = N; q.D = h.D; q.H = h.H; q.W = h.W;
                    q.stride = 1;
                    q.n_seg = 1;
                    q.seg[0] = {h.p, h.C, 2};
                    q.weights = L.samp.w_up; q.w_rows = (1 << sd) * h.C; q.Cout = h.C;
                    q.mode = EPI_STORE;
                    q.bias = L.samp.bias;
                    q.upsample2 = 1;
                    // 2-D: on the halo-tile kernel (the four taps of a phase are views of ONE staged low-resolution tile:
                    // a quarter of the im2col kernel's A traffic). Round 1 measured this slower (short 4-tap items were
                    // epilogue-bound); with 256-bit epilogue stores and prefetched addends it is +1.7 % on the same box.
                    const char* uh = getenv("DDPM_UPCONV_HALO");  // tests: 0 = the im2col-tile phases (read at plan time)
                    const bool up_halo = !(uh && atoi(uh) == 0);
                    const bool on_halo = up_halo && use_halo_ && conv_halo_supported(q);
                    Act o = shape_act(h.C, h.D * fd, h.H * 2, h.W * 2);
                    if (!measure) {
                        o = new_act(h.C, h.D * fd, h.H * 2, h.W * 2, false);
                        if (fuse_gn_stats_) {
                            o.parts = on_halo ? conv_halo_stats_parts(h.H, h.W) * 4 : conv_stats_parts(sd, h.D, h.H, h.W) * (1 << sd);
                            o.stats = take_stats(h.C, o.parts);
                        }
                    }
                    q.out = o.p;
                    q.stats_out = o.stats;
                    if (on_halo && !measure) halo_conv(q, nullptr, 0, -1);
                    else gemm(q, -1);
                    h = o;
                } else {
                    const size_t cnt = static_cast<size_t>(N) * (h.D * fd) * (h.H * 2) * (h.W * 2) * h.C;
                    if (measure && cnt > max_up) max_up = cnt;
                    Act up = measure ? shape_act(h.C, h.D * fd, h.H * 2, h.W * 2) : new_act(h.C, h.D * fd, h.H * 2, h.W * 2, false);
                    if (!measure) {
                        Op op{};
                        op.type = Op::UPSAMPLE;
                        op.src0 = h.p; op.dst = up.p; op.D = h.D; op.H = h.H; op.W = h.W; op.C = h.C;
                        op.bytes = 2.0 * cnt + 2.0 * cnt / (4.0 * fd);
                        plan.ops.push_back(op);
                    }
                    Act o = measure ? up : new_act(h.C, up.D, up.H, up.W);
                    conv3(up, up.p, h.C, L.samp.w, L.samp.bias, h.C, 1, -1, nullptr, o.p, nullptr, nullptr, o.stats);
                    h = o;
                }
            }
        }
        if (h.D != D || h.H != H || h.W != W) { set_error("unet: output spatial size mismatch"); return 7; }
        const bool out_taps = !measure && !out_gemm_ && h.parts > 0 && conv_out_taps_supported(h.C, c.out_channels, sd);
        float* dtaps = nullptr;
        if (out_taps) {
            // final GroupNorm+SiLU fused with the output conv's channel reduction; no normalised tensor in memory
            dtaps = lay.take<float>(static_cast<size_t>(N) * D * H * W * 9 * c.out_channels);
            Op op{};
            op.type = Op::GN;
            op.src0 = h.p; op.C0 = h.C; op.st0 = h.stats; op.parts0 = h.parts;
            op.gamma = out_g_; op.beta = out_b_; op.S = static_cast<int>(h.S()); op.silu = true;
            op.dtaps = dtaps;
            op.bytes = static_cast<double>(N) * h.S() * (2.0 * h.C + 36.0 * c.out_channels);
            plan.ops.push_back(op);
        }
        ConvProblem qout{};  // conv_out of wide-channel outputs: on the halo kernel the out-norm is applied in shared memory
        qout.spatial_dims = sd; qout.N = N; qout.D = D; qout.H = H; qout.W = W; qout.stride = 1; qout.n_seg = 1;
        qout.seg[0] = {nullptr, h.C, 3};
        qout.Cout = c.out_channels; qout.mode = EPI_STORE;
        const bool out_halo = out_gemm_ && use_halo_ && h.parts > 0 && h.C <= 512 && conv_halo_supported(qout);
        if (!out_taps && !out_halo) gn(h, nullptr, out_g_, out_b_, zA, true);
        if (!measure) {
            plan.z_out = zA;
            Op op{};
            op.type = out_gemm_ ? Op::CONV_OUT_GEMM : Op::CONV_OUT_SMALL;
            op.dtaps = dtaps;
            op.src0 = zA; op.D = D; op.H = H; op.W = W; op.C = h.C;
            op.flops = 2.0 * N * D * H * W * c.out_channels * (sd == 3 ? 27.0 : 9.0) * h.C;
            op.bytes = static_cast<double>(N) * D * H * W * (2.0 * h.C + 4.0 * c.out_channels);
            op.on_halo = out_halo;
            if (out_gemm_ && !dry) {
                ConvProblem q = qout;
                q.seg[0].ptr = out_halo ? h.p : zA;
                q.weights = conv_out_wp_; q.w_rows = c.out_channels;
                q.bias = conv_out_b_; q.out = plan.y_half;
                q.gn_silu = 1;
                const HaloGnSource src = gn_source(h, nullptr, out_g_, out_b_);
                int r = out_halo ? conv_halo_prepare(q, nullptr, h.C, sms, &op.halo, &src) : conv_prepare(q, sms, &op.conv);
                if (r && !rc) rc = r;
            }
            plan.ops.push_back(op);
        }
        if (rc) return rc;
        if (measure) lay = saved;
    }
    if (need) *need = lay.off + 1024;
    if (!dry && lay.off > ws_bytes) { set_error("unet: workspace too small (%zu < %zu)", ws_bytes, lay.off); return 9; }
    plan.N = N; plan.D = D; plan.H = H; plan.W = W; plan.ws = ws;
    return 0;
}

size_t UNet::workspace_bytes(int N, int D, int H, int W) const {
    Plan p{};
    size_t need = 0;
    if (build_plan(p, N, D, H, W, nullptr, 0, true, &need)) return 0;
    return need;
}

double UNet::flops_per_image(int D, int H, int W) const {
    (void)D; (void)H; (void)W;
    return 0.0;  // computed on the Python side from the layer list (oracle.unet.count_flops for tests)
}

// ------------------------------------------------------------------------------------------------ forward
int UNet::forward(const float* x, const long long* timesteps, int t_uniform, float* out, int N, int D, int H, int W,
                  void* ws, size_t ws_bytes, cudaStream_t stream, const PlmsStep* plms, float* ring, float* stash,
                  float* sample) {
    if (!finalized_) { set_error("unet: forward before finalize()"); return 10; }
    if (cfg_.spatial_dims == 2 && D != 1) { set_error("unet: 2-D model needs D == 1"); return 2; }
    if (N < 1 || D < 1 || H < 1 || W < 1) { set_error("unet: empty input (N=%d, D=%d, H=%d, W=%d)", N, D, H, W); return 2; }
    auto key = std::make_tuple(N, D, H, W, ws);
    auto it = plans_.find(key);
    if (it == plans_.end()) {
        std::unique_ptr<Plan> p(new Plan());
        int rc = build_plan(*p, N, D, H, W, ws, ws_bytes, false, nullptr);
        if (rc) return rc;
        it = plans_.emplace(key, std::move(p)).first;
    }
    Plan& plan = *it->second;
    const UNetConfig& c = cfg_;
    const int R = timesteps ? N : 1;
    const bool prof = profile_every_ > 0 && (profile_tick_++ % profile_every_ == 0);
    if (prof) {
        if (plan.events_pending) harvest(plan);
        if (plan.events.empty()) {
            plan.events.resize(plan.ops.size() + 2);
            for (auto& e : plan.events) cudaEventCreate(&e);
        }
        cudaEventRecord(plan.events[0], stream);
    }
    size_t op_idx = 0;
    // Timestep embedding -> all ResnetBlock projections. With one timestep for the whole batch (the reconstruction
    // chain) the row comes from a table computed once per weight upload; per-sample timesteps run the MLP here.
    int rc = 0;
    const float* temb_base;
    long long temb_stride;
    if (!timesteps && temb_table_ && t_uniform >= 0 && t_uniform < temb_rows_) {
        temb_base = temb_table_ + static_cast<size_t>(t_uniform) * P_;
        temb_stride = 0;
    } else {
        rc = time_embed(timesteps, t_uniform, R, E_, te_w0_, te_b0_, te_w1_, te_b1_, plan.temb_act, stream);
        if (rc) return rc;
        rc = time_proj_all(plan.temb_act, R, 4 * E_, tp_w_, tp_b_, P_, plan.temb_all, stream);
        if (rc) return rc;
        launches_ += 2;
        temb_base = plan.temb_all;
        temb_stride = timesteps ? P_ : 0;
    }
    const long long S = static_cast<long long>(D) * H * W;
    for (Op& op : plan.ops) {
        if (prof) cudaEventRecord(plan.events[1 + op_idx], stream);
        ++op_idx;
        switch (op.type) {
            case Op::CONV_IN_SMALL:
                rc = conv_in_small(x, conv_in_w_, conv_in_b_, op.dst, N, c.in_channels, D, H, W, c.num_channels[0],
                                   c.spatial_dims, const_cast<float*>(op.st0), stream);
                break;
            case Op::CONV_IN_GEMM:
                rc = nchw_to_nhwc_half(x, plan.x_half, N, c.in_channels, S, stream);
                if (!rc) rc = op.on_halo ? conv_halo_launch(op.halo, stream) : conv_launch(op.conv, stream);
                ++launches_;
                break;
            case Op::GN:
                if (op.dtaps)
                    rc = gn_apply_taps(op.src0, op.C0, op.st0, op.parts0, op.gamma, op.beta, conv_out_w_, c.out_channels,
                                       op.dtaps, N, op.S, c.norm_num_groups, c.norm_eps, stream);
                else if (op.st0)
                    rc = gn_apply(op.src0, op.C0, op.st0, op.parts0, op.src1, op.C1, op.st1, op.parts1, op.gamma, op.beta,
                                  op.dst, N, op.S, c.norm_num_groups, c.norm_eps, op.silu, stream);
                else
                    rc = gn_silu(op.src0, op.C0, op.src1, op.C1, op.gamma, op.beta, op.dst, N, op.S, c.norm_num_groups,
                                 c.norm_eps, op.silu, stream);
                break;
            case Op::GEMM:
                if (op.uses_temb) {
                    op.conv.p.chan_add = temb_base + op.temb_off;
                    op.conv.p.chan_add_stride = temb_stride;
                }
                rc = conv_launch(op.conv, stream);
                break;
            case Op::GN_FINALIZE:
                rc = gn_finalize(op.C0, op.st0, op.parts0, op.C1, op.st1, op.parts1, op.gamma, op.beta, op.ab, N, op.S,
                                 c.norm_num_groups, c.norm_eps, stream);
                break;
            case Op::CONV_HALO:
                if (op.uses_temb) {
                    op.halo.p.g.chan_add = temb_base + op.temb_off;
                    op.halo.p.g.chan_add_stride = temb_stride;
                }
                rc = conv_halo_launch(op.halo, stream);
                break;
            case Op::ATTN_BLOCK:
                rc = attn_block_launch(op.attn_block, stream);
                break;
            case Op::ATTN:
                rc = op.attn_tc ? attention_tc_launch(op.attn, stream)
                                : attention_core(op.src0, op.dst, N, op.T, op.C, op.heads, op.scale, stream);
                break;
            case Op::UPSAMPLE:
                rc = upsample_nearest2(op.src0, op.dst, N, op.D, op.H, op.W, op.C, c.spatial_dims, stream);
                break;
            case Op::CONV_OUT_SMALL:
                if (op.dtaps) {
                    rc = conv_out_gather(op.dtaps, conv_out_b_, out, N, H, W, c.out_channels, plms, ring, stash, sample,
                                         stream);
                    break;
                }
                rc = conv_out_small(op.src0, conv_out_w_, conv_out_b_, out, N, op.C, D, H, W, c.out_channels,
                                    c.spatial_dims, plms, ring, stash, sample, stream);
                break;
            case Op::CONV_OUT_GEMM: {
                rc = op.on_halo ? conv_halo_launch(op.halo, stream) : conv_launch(op.conv, stream);
                if (rc) break;
                float* eps = out ? out
                                 : (plms && !plms->push) ? plan.eps_tmp
                                                         : ring + static_cast<long long>(plms ? plms->slot_new : 0) * N * c.out_channels * S;
                dim3 grid(static_cast<unsigned>((S + 31) / 32), (c.out_channels + 31) / 32, N);
                nhwc_half_to_nchw_kernel<<<grid, dim3(32, 8), 0, stream>>>(plan.y_half, eps, c.out_channels, S);
                ++launches_;
                if (plms) {
                    rc = plms_update(eps, *plms, ring, stash, sample, sample, static_cast<long long>(N) * c.out_channels * S, stream);
                    ++launches_;
                }
                break;
            }
        }
        ++launches_;
        if (rc) return rc;
    }
    if (prof) {
        cudaEventRecord(plan.events[1 + op_idx], stream);
        plan.events_pending = true;
    }
    return 0;
}

int UNet::run_chain(int n_steps, const int* timesteps, const PlmsStep* steps, float* sample, float* ring, float* stash, int N,
                    int D, int H, int W, void* ws, size_t ws_bytes, cudaStream_t stream) {
    auto run_plain = [&]() -> int {
        for (int i = 0; i < n_steps; ++i) {
            int rc = forward(sample, nullptr, timesteps[i], nullptr, N, D, H, W, ws, ws_bytes, stream, &steps[i], ring, stash,
                             sample);
            if (rc) return rc;
        }
        return 0;
    };
    // CUDA-graph replay of the chain pays when the chain is launch-bound: at batch 256 it measured flat (2201 vs 2196
    // reconstructions/s - kernel-bound, programmatic dependent launch already hides the gaps), at batch 8 (BASELINE
    // configs[0]) a forward's ~45 kernels are shorter than their launch work. Policy: replay when one forward covers at
    // most kGraphPixels pixels; DDPM_CHAIN_GRAPH=0 / 1 forces it off / on.
    constexpr long long kGraphPixels = 64 * 1024;
    const char* ge = getenv("DDPM_CHAIN_GRAPH");  // once per chain, not per launch
    const int graph_env = ge ? (atoi(ge) != 0 ? 1 : 0) : -1;
    const bool want_graph = graph_env >= 0 ? graph_env == 1 : static_cast<long long>(N) * D * H * W <= kGraphPixels;
    if (!want_graph || !use_chain_graph_ || profile_every_ > 0 || !chain_warm_ || n_steps < 1) {
        int rc = run_plain();
        if (!rc) chain_warm_ = true;
        return rc;
    }
    // key: everything a captured launch sequence depends on
    std::string key;
    auto put = [&](const void* p, size_t n) { key.append(static_cast<const char*>(p), n); };
    put(&n_steps, sizeof(n_steps));
    put(timesteps, sizeof(int) * n_steps);
    put(steps, sizeof(PlmsStep) * n_steps);
    const void* ptrs[5] = {sample, ring, stash, ws, stream};
    put(ptrs, sizeof(ptrs));
    const int dims[4] = {N, D, H, W};
    put(dims, sizeof(dims));
    auto it = chain_graphs_.find(key);
    if (it == chain_graphs_.end()) {
        if (chain_graphs_.size() >= 256) {  // bounded cache (25-100 distinct chains per configuration in practice)
            for (auto& kv : chain_graphs_) cudaGraphExecDestroy(kv.second.exec);
            chain_graphs_.clear();
        }
        const long long before = launches_;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            use_chain_graph_ = false;
            return run_plain();
        }
        const int rc = run_plain();
        cudaGraph_t graph = nullptr;
        const cudaError_t e_end = cudaStreamEndCapture(stream, &graph);
        const long long captured = launches_ - before;
        launches_ = before;  // nothing has executed yet
        if (rc || e_end != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            use_chain_graph_ = false;  // fall back to plain launches for good (still the CUDA path, just more launch work)
            return rc ? rc : run_plain();
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e_inst != cudaSuccess || !exec) {
            cudaGetLastError();
            use_chain_graph_ = false;
            return run_plain();
        }
        it = chain_graphs_.emplace(key, ChainGraph{exec, captured}).first;
    }
    const cudaError_t e = cudaGraphLaunch(it->second.exec, stream);
    if (e != cudaSuccess) { set_error("unet: cudaGraphLaunch failed: %s", cudaGetErrorString(e)); return 5; }
    launches_ += it->second.launches;
    return 0;
}

void UNet::harvest(Plan& plan) {
    if (!plan.events_pending) return;
    const size_t n = plan.ops.size();
    cudaEventSynchronize(plan.events[n + 1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, plan.events[0], plan.events[1]);
    prof_.ms[kNumOpTypes - 1] += ms;
    prof_.launches[kNumOpTypes - 1] += 2;
    for (size_t i = 0; i < n; ++i) {
        const Op& op = plan.ops[i];
        cudaEventElapsedTime(&ms, plan.events[1 + i], plan.events[2 + i]);
        // the appended op types report under their families: GroupNorm (2) and tensor-core conv (3)
        const int t = op.type == Op::GN_FINALIZE ? static_cast<int>(Op::GN)
                      : op.type == Op::CONV_HALO ? static_cast<int>(Op::GEMM)
                      : op.type == Op::ATTN_BLOCK ? static_cast<int>(Op::ATTN) : static_cast<int>(op.type);
        prof_.ms[t] += ms;
        prof_.flops[t] += op.flops;
        prof_.bytes[t] += op.bytes;
        prof_.launches[t] += 1;
    }
    cudaEventElapsedTime(&ms, plan.events[0], plan.events[n + 1]);
    prof_.forward_ms += ms;
    prof_.forwards += 1;
    plan.events_pending = false;
}

int UNet::read_profile(OpProfile* out, bool reset) {
    for (auto& kv : plans_) harvest(*kv.second);
    if (out) *out = prof_;
    if (reset) memset(&prof_, 0, sizeof(prof_));
    return 0;
}

}  // namespace ddpm
