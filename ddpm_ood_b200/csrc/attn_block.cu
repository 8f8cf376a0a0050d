// Fused AttentionBlock kernel (tcgen05 / TMEM / TMA). See attn_block.cuh.
#include "attn_block.cuh"

#include <string.h>

#include "conv_epilogue.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace ddpm {

namespace {

constexpr int kC = 256;                     // channels == head dim
constexpr int kPanel = 128 * 128;           // one K-major SWIZZLE_128B panel: 128 rows x 64 fp16 = 16 KB
constexpr int kRingStages = 6;              // weight panels in flight (128 output rows x 64 K each)
constexpr int kWorkerWarps = 8;             // GroupNorm transform, TMEM drains, softmax, final epilogue
constexpr int kWarpTma = 8, kWarpMma = 9;
constexpr int kThreads = 320;
// shared memory map (bytes from the 1024-aligned base)
constexpr int kOffR0 = 0;                   // 64 KB: normalised h (A of q/k/v), later O as fp16 (A of the projection)
constexpr int kOffR1a = 4 * kPanel;         // 32 KB: k half (B of S), later P (A of PV); GroupNorm scratch at the start
constexpr int kOffR1b = 6 * kPanel;         // 32 KB: q half (A of S), later v half (B of PV, MN-major)
constexpr int kOffRing = 8 * kPanel;        // 6 x 16 KB weight ring
constexpr int kOffBias = kOffRing + kRingStages * kPanel;  // 256 fp32: projection bias (epilogue addend)
constexpr int kOffInv = kOffBias + kC * 4;                 // 128 fp32: 1 / softmax row sums
constexpr int kOffBars = kOffInv + 128 * 4;
constexpr int kNumBars = 2 + 2 * kRingStages + 4 + 5;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16 + 1024 /*alignment slack*/;
static_assert(kSmemBytes <= 227 * 1024, "shared memory");
// TMEM columns (fp32 accumulators, 128 lanes each)
constexpr uint32_t kColS = 0;               // S = q k^T, 128 columns
constexpr uint32_t kColO = 128;             // O = P v, 256 columns
constexpr uint32_t kColT0 = 384;            // projection scratch 0
constexpr uint32_t kColT1 = 128;            // projection scratch 1 (the first half of O, free outside the P v phase)

// MN-major SWIZZLE_128B operand as the drain leaves v: rows = keys (K index), 128 B = 64 d-values per row, the next 64
// d-values one panel (16 KB) further.
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(kPanel >> 4) << 16;  // leading byte offset
    d |= static_cast<uint64_t>(1024 >> 4) << 32;    // stride byte offset
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// 64 accumulator columns of this thread's row -> (+ bias) -> fp16 -> one K-major SWIZZLE_128B panel row (128 B).
// t_addr: TMEM address (lane quarter + first column); dst_row: shared address of the row inside the panel.
__device__ __forceinline__ void drain64(uint32_t t_addr, uint32_t dst_row, int row, const float4 (&b4)[16], float mul) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_addr + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 ba = b4[c * 8 + 2 * j], bb = b4[c * 8 + 2 * j + 1];
            uint4 o;
            o.x = pack_h2(fmaf(__uint_as_float(v[8 * j + 0]), mul, ba.x), fmaf(__uint_as_float(v[8 * j + 1]), mul, ba.y));
            o.y = pack_h2(fmaf(__uint_as_float(v[8 * j + 2]), mul, ba.z), fmaf(__uint_as_float(v[8 * j + 3]), mul, ba.w));
            o.z = pack_h2(fmaf(__uint_as_float(v[8 * j + 4]), mul, bb.x), fmaf(__uint_as_float(v[8 * j + 5]), mul, bb.y));
            o.w = pack_h2(fmaf(__uint_as_float(v[8 * j + 6]), mul, bb.z), fmaf(__uint_as_float(v[8 * j + 7]), mul, bb.w));
            const int chunk = c * 4 + j;
            sts128(dst_row + ((chunk ^ (row & 7)) << 4), o);
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) attn_block_kernel(const __grid_constant__ AttnBlockParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* r0 = smem + kOffR0;
    uint8_t* r1a = smem + kOffR1a;
    uint8_t* r1b = smem + kOffR1b;
    uint8_t* ring = smem + kOffRing;
    float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
    float* s_inv = reinterpret_cast<float*>(smem + kOffInv);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint64_t* x_full = bars;                       // TMA -> workers
    uint64_t* xhat_ready = bars + 1;               // workers -> MMA (8 warps)
    uint64_t* ring_full = bars + 2;                // TMA -> MMA
    uint64_t* ring_empty = ring_full + kRingStages;  // MMA commit -> TMA
    uint64_t* t_full = ring_empty + kRingStages;   // [2] MMA commit -> workers (projection scratch T0 / T1 complete)
    uint64_t* t_empty = t_full + 2;                // [2] workers -> MMA (scratch drained AND its fp16 copy is in smem)
    uint64_t* s_full = t_empty + 2;                // MMA commit -> softmax warps
    uint64_t* p_ready = s_full + 1;                // softmax warps (4) -> MMA
    uint64_t* o_full = p_ready + 1;                // MMA commit -> workers
    uint64_t* o_ready = o_full + 1;                // workers -> MMA (O as fp16 in R0)
    uint64_t* r0_free = o_ready + 1;               // MMA commit -> TMA (projection done reading R0)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ptx::pdl_trigger();
    if (warp == kWarpTma && lane == 0) {
        ptx::prefetch_tmap(&p.tm_x);
        ptx::prefetch_tmap(&p.tm_wqkv);
        ptx::prefetch_tmap(&p.tm_wproj);
    }
    if (warp == kWarpMma) {
        if (lane == 0) {
            ptx::mbar_init(x_full, 1);
            ptx::mbar_init(xhat_ready, kWorkerWarps);
            for (int i = 0; i < kRingStages; ++i) { ptx::mbar_init(&ring_full[i], 1); ptx::mbar_init(&ring_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], kWorkerWarps); }
            ptx::mbar_init(s_full, 1);
            ptx::mbar_init(p_ready, 4);
            ptx::mbar_init(o_full, 1);
            ptx::mbar_init(o_ready, kWorkerWarps);
            ptx::mbar_init(r0_free, 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc<512>(tmem_slot);
    }
    // projection bias -> shared memory (a weight: not produced by the preceding kernel, safe before the dependency wait)
    if (warp < kWorkerWarps && threadIdx.x < kC) s_bias[threadIdx.x] = p.bproj ? __ldg(p.bproj + threadIdx.x) : 0.f;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();

    constexpr uint32_t idesc = ptx::make_idesc_f16(128, 128);
    constexpr uint32_t idesc_bt = idesc | (1u << 16);  // B operand MN-major (v)

    if (warp == kWarpTma) {
        // ================================================================= TMA producer: h tile + the weight stream
        int st = 0;
        uint32_t ph = 0, ph_r0 = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            ptx::mbar_wait(r0_free, ph_r0 ^ 1);  // previous tile's projection has finished reading R0
            ph_r0 ^= 1;
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(x_full, 4 * kPanel);
#pragma unroll
                for (int pn = 0; pn < 4; ++pn) ptx::tma_load_3d(r0 + pn * kPanel, &p.tm_x, x_full, pn * 64, 0, tile * p.ipt);
            }
            __syncwarp();
            // jobs in the MMA warp's order: k0, q0, k1, q1, v0, v1 (rows of wqkv), proj0, proj1 (rows of wproj)
#pragma unroll 1
            for (int job = 0; job < 8; ++job) {
                const int row0 = job == 0 ? 256 : job == 1 ? 0 : job == 2 ? 384 : job == 3 ? 128 : job == 4 ? 512 : job == 5 ? 640
                                                                                                              : job == 6 ? 0 : 128;
                const CUtensorMap* m = job < 6 ? &p.tm_wqkv : &p.tm_wproj;
#pragma unroll 1
                for (int kc = 0; kc < 4; ++kc) {
                    ptx::mbar_wait(&ring_empty[st], ph ^ 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&ring_full[st], kPanel);
                        ptx::tma_load_2d(ring + st * kPanel, m, &ring_full[st], kc * 64, row0);
                    }
                    __syncwarp();
                    if (++st == kRingStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ================================================================= MMA issuer (one elected lane; the warp walks)
        int st = 0;
        uint32_t ph = 0, ph_x = 0, ph_te = 0 /*bit b: parity of scratch b*/, ph_p = 0, ph_or = 0;
        const uint32_t r0_u = ptx::smem_u32(r0), r1a_u = ptx::smem_u32(r1a), r1b_u = ptx::smem_u32(r1b),
                       ring_u = ptx::smem_u32(ring);
        // one projection half: scratch[128 x 128] = A[128 x 256] (4 K-major panels at a_base) x W panel rows^T
        auto job = [&](uint32_t a_base, uint32_t d_tmem, int tbuf) {
#pragma unroll 1
            for (int kc = 0; kc < 4; ++kc) {
                ptx::mbar_wait(&ring_full[st], ph);
                ptx::tc_fence_after();
                const uint64_t da = ptx::make_desc_k128(a_base + kc * kPanel);
                const uint64_t db = ptx::make_desc_k128(ring_u + st * kPanel);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0);
                    ptx::umma_commit(&ring_empty[st]);
                    if (kc == 3) ptx::umma_commit(&t_full[tbuf]);
                }
                __syncwarp();
                if (++st == kRingStages) { st = 0; ph ^= 1; }
            }
        };
        auto wait_te = [&](int b) {  // scratch b drained (and, for q/k/v, its fp16 copy visible in shared memory)
            ptx::mbar_wait(&t_empty[b], ((ph_te >> b) & 1u) ^ 1u);
            ph_te ^= 1u << b;
            ptx::tc_fence_after();
        };
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            wait_te(0);  // previous tile's final epilogue has drained both scratch buffers (first tile: passes)
            wait_te(1);
            ptx::mbar_wait(xhat_ready, ph_x);
            ph_x ^= 1;
            ptx::tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                job(r0_u, tmem_base + kColT0, 0);  // k half h
                job(r0_u, tmem_base + kColT1, 1);  // q half h
                wait_te(0);
                wait_te(1);
                // S (+)= q_h k_h^T over this half of the head dim (two 64-wide panels)
                if (ptx::elect_one()) {
#pragma unroll
                    for (int pn = 0; pn < 2; ++pn) {
                        const uint64_t da = ptx::make_desc_k128(r1b_u + pn * kPanel);
                        const uint64_t db = ptx::make_desc_k128(r1a_u + pn * kPanel);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_f16(tmem_base + kColS, da + 2 * k, db + 2 * k, idesc, (h | pn | k) != 0);
                    }
                    if (h == 1) ptx::umma_commit(s_full);
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                job(r0_u, tmem_base + kColT0, 0);  // v half h (overlaps the softmax)
                if (h == 0) {
                    ptx::mbar_wait(p_ready, ph_p);
                    ph_p ^= 1;
                }
                wait_te(0);
                // O[:, half h] = P v_h: K = 128 keys in steps of 16
                if (ptx::elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t da = ptx::make_desc_k128(r1a_u + (ks >> 2) * kPanel) + 2 * (ks & 3);
                        const uint64_t db = make_desc_mn128(r1b_u + ks * 16 * 128);
                        ptx::umma_f16(tmem_base + kColO + h * 128, da, db, idesc_bt, ks != 0);
                    }
                    if (h == 1) ptx::umma_commit(o_full);
                }
                __syncwarp();
            }
            ptx::mbar_wait(o_ready, ph_or);
            ph_or ^= 1;
            ptx::tc_fence_after();
            job(r0_u, tmem_base + kColT0, 0);  // out[:, 0:128)
            job(r0_u, tmem_base + kColT1, 1);  // out[:, 128:256)
            if (ptx::elect_one()) ptx::umma_commit(r0_free);
            __syncwarp();
        }
    } else {
        // ================================================================= workers (8 warps)
        const int q = warp & 3;       // TMEM lane quarter
        const int hf = warp >> 2;     // which half of a 128-column scratch this warp drains
        const int row = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t r0_u = ptx::smem_u32(r0), r1a_u = ptx::smem_u32(r1a), r1b_u = ptx::smem_u32(r1b);
        uint32_t ph_x = 0, ph_tf = 0 /*bit b: parity of scratch b*/, ph_s = 0, ph_o = 0;
        float4 b4[16];
        auto load_bias = [&](const float* g) {  // issued BEFORE the barrier wait: the L2 round trip hides behind the MMAs
#pragma unroll
            for (int i = 0; i < 16; ++i) b4[i] = g ? __ldg(reinterpret_cast<const float4*>(g) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto drain_to = [&](int tbuf, uint32_t dst_region) {  // scratch -> fp16 panel `hf` of a 2-panel region
            ptx::mbar_wait(&t_full[tbuf], (ph_tf >> tbuf) & 1u);
            ph_tf ^= 1u << tbuf;
            ptx::tc_fence_after();
            drain64(tmem_base + lane_addr + (tbuf == 0 ? kColT0 : kColT1) + hf * 64, dst_region + hf * kPanel + row * 128, row,
                    b4, 1.0f);
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&t_empty[tbuf]);
        };
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            // ---- GroupNorm of the staged tile, in place. lane = 8-channel group (one 16-byte chunk), warp = 16 rows.
            {
                const int cg = lane, pn = cg >> 3, ch = cg & 7;
                float ga[8], gb[8];
                {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * cg);
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * cg + 1);
                    const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * cg);
                    const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * cg + 1);
                    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
                    gb[0] = e0.x; gb[1] = e0.y; gb[2] = e0.z; gb[3] = e0.w; gb[4] = e1.x; gb[5] = e1.y; gb[6] = e1.z; gb[7] = e1.w;
                }
                ptx::mbar_wait(x_full, ph_x);
                ph_x ^= 1;
                float2* scratch = reinterpret_cast<float2*>(r1a);  // [16 row segments][32 groups] partial (sum, sum of squares)
                uint4 raw[16];
                const uint32_t base = r0_u + pn * kPanel;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = warp * 16 + i;
                    raw[i] = lds128(base + r * 128 + ((ch ^ (r & 7)) << 4));
                }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    float sum = 0.f, sq = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[8 * s + i]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(h2[e]);
                            sum += f.x + f.y;
                            sq = fmaf(f.x, f.x, fmaf(f.y, f.y, sq));
                        }
                    }
                    scratch[(2 * warp + s) * 32 + cg] = make_float2(sum, sq);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float inv_n = 1.0f / (8.0f * static_cast<float>(p.T));
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int row0 = (2 * warp + s) * 8;
                    const int img = row0 / p.Tpad;
                    const int seg0 = (img * p.Tpad) >> 3, nseg = p.Tpad >> 3;
                    float sum = 0.f, sq = 0.f;
                    for (int j = 0; j < nseg; ++j) { const float2 v = scratch[(seg0 + j) * 32 + cg]; sum += v.x; sq += v.y; }
                    const float mean = sum * inv_n;
                    float var = sq * inv_n - mean * mean;
                    var = var < 0.f ? 0.f : var;
                    const float rstd = rsqrtf(var + p.eps);
                    const bool img_ok = tile * p.ipt + img < p.N;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = row0 + i;
                        if (!img_ok || r - img * p.Tpad >= p.T) continue;  // zero-filled padding rows stay zero
                        uint4 v = raw[8 * s + i];
                        __half2* h2 = reinterpret_cast<__half2*>(&v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(h2[e]);
                            const float a0 = ga[2 * e] * rstd, a1 = ga[2 * e + 1] * rstd;
                            h2[e] = __floats2half2_rn(fmaf(f.x, a0, gb[2 * e] - mean * a0), fmaf(f.y, a1, gb[2 * e + 1] - mean * a1));
                        }
                        sts128(base + r * 128 + ((ch ^ (r & 7)) << 4), v);
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(xhat_ready);
            }
            // ---- q / k halves -> shared memory
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                load_bias(p.bqkv ? p.bqkv + 256 + h * 128 + hf * 64 : nullptr);
                drain_to(0, r1a_u);  // k half
                load_bias(p.bqkv ? p.bqkv + h * 128 + hf * 64 : nullptr);
                drain_to(1, r1b_u);  // q half
            }
            // ---- softmax over each row's own image block -> P (fp16, K-major) over the k region; 4 warps, thread = row
            if (hf == 0) {
                const int img = row / p.Tpad;
                const int c_lo = img * p.Tpad, c_hi = c_lo + p.T;  // key columns this row attends to
                ptx::mbar_wait(s_full, ph_s);
                ptx::tc_fence_after();
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = c_lo >> 5; c <= (c_hi - 1) >> 5; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(tmem_base + lane_addr + kColS + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        if (col >= c_lo && col < c_hi) mx = fmaxf(mx, __uint_as_float(v[j]));
                    }
                }
                const float m2 = mx * p.scale_log2e;
                float sum = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t packed[16];
                    if (c >= (c_lo >> 5) && c <= ((c_hi - 1) >> 5)) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(tmem_base + lane_addr + kColS + c * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const int col = c * 32 + j;
                            float a = 0.f, b = 0.f;
                            if (col >= c_lo && col < c_hi) a = exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -m2));
                            if (col + 1 >= c_lo && col + 1 < c_hi) b = exp2f(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -m2));
                            const __half2 hh = __floats2half2_rn(a, b);
                            const float2 back = __half22float2(hh);  // normalise by the sum of what the tensor core will see
                            sum += back.x + back.y;
                            packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) packed[j] = 0u;
                    }
                    const uint32_t prow = r1a_u + (c >> 1) * kPanel + row * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int chunk = (c & 1) * 4 + j;
                        sts128(prow + ((chunk ^ (row & 7)) << 4),
                               make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
                    }
                }
                s_inv[row] = 1.0f / sum;
                ptx::fence_proxy_async_smem();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p_ready);
            }
            ph_s ^= 1;
            // ---- v halves -> shared memory (B of P v)
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                load_bias(p.bqkv ? p.bqkv + 512 + h * 128 + hf * 64 : nullptr);
                drain_to(0, r1b_u);
            }
            // ---- O / row sum -> fp16 into R0 (A of the projection); this warp: columns [128 hf, 128 hf + 128)
            asm volatile("bar.sync 1, 256;" ::: "memory");  // s_inv of every row is written
            {
#pragma unroll
                for (int i = 0; i < 16; ++i) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const float inv = s_inv[row];
                ptx::mbar_wait(o_full, ph_o);
                ph_o ^= 1;
                ptx::tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int pn = hf * 2 + half;
                    drain64(tmem_base + lane_addr + kColO + pn * 64, r0_u + pn * kPanel + row * 128, row, b4, inv);
                }
                ptx::fence_proxy_async_smem();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(o_ready);
            }
            // ---- projection halves: + bias + residual h -> fp16 out (+ GroupNorm statistics for the next layer)
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                ptx::mbar_wait(&t_full[half], (ph_tf >> half) & 1u);
                ph_tf ^= 1u << half;
                ptx::tc_fence_after();
                conv_epilogue_tile16<128>(p.epi, tmem_base + lane_addr + (half == 0 ? kColT0 : kColT1), tile, half, 0, q, lane,
                                          s_bias + half * 128, hf * 4, hf * 4 + 4);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&t_empty[half]);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int pow2_at_least(int v) {
    int p2 = 8;
    while (p2 < v) p2 <<= 1;
    return p2;
}

bool attn_block_supported(int T, int C, int heads, int groups) {
    return C == kC && heads == 1 && groups == 32 && T >= 1 && T <= 128;
}

int attn_block_stats_parts(int T) {
    const int tp = pow2_at_least(T);
    return tp >= 32 ? tp / 32 : 0;
}

int attn_block_prepare(const __half* h, __half* out, int N, int T, int C, int heads, int groups, float eps, float scale,
                       const float* gamma, const float* beta, const __half* wqkv, const float* bqkv, const __half* wproj,
                       const float* bproj, float* stats_out, int num_sms, AttnBlockLaunch* l) {
    if (!attn_block_supported(T, C, heads, groups)) {
        set_error("attn_block: T=%d C=%d heads=%d groups=%d unsupported (C 256, 1 head, 32 groups, T <= 128)", T, C, heads, groups);
        return 2;
    }
    PFN_encodeTiled encode = get_encode();
    if (!encode) return 1;
    memset(l, 0, sizeof(*l));
    AttnBlockParams& p = l->p;
    p.N = N; p.T = T; p.Tpad = pow2_at_least(T); p.ipt = 128 / p.Tpad;
    p.num_tiles = (N + p.ipt - 1) / p.ipt;
    p.eps = eps;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.gamma = gamma; p.beta = beta; p.bqkv = bqkv; p.bproj = bproj;
    if (stats_out && p.Tpad < 32) { set_error("attn_block: fused statistics need >= 32 rows per image (T=%d)", T); return 2; }
    {
        cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(N)};
        cuuint64_t gstr[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * 2 * T};
        cuuint32_t box[3] = {64, static_cast<cuuint32_t>(p.Tpad), static_cast<cuuint32_t>(p.ipt)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&p.tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(h), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("attn_block: cuTensorMapEncodeTiled(h) failed: %d", (int)r); return 3; }
    }
    for (int i = 0; i < 2; ++i) {
        cuuint64_t gdim[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(i == 0 ? 3 * C : C)};
        cuuint64_t gstr[1] = {static_cast<cuuint64_t>(C) * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(i == 0 ? &p.tm_wqkv : &p.tm_wproj, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                            const_cast<__half*>(i == 0 ? wqkv : wproj), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("attn_block: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return 3; }
    }
    // epilogue geometry: a tile is `ipt` images of Tpad "pixels" in a row (W = T valid)
    ConvGemmParams& e = p.epi;
    e.N = N; e.D = 1; e.H = 1; e.W = T;
    e.bw = p.Tpad; e.bh = 1; e.bd = 1; e.bn = p.ipt;
    e.tiles_w = 1; e.tiles_h = 1; e.tiles_d = 1; e.tiles_n = p.num_tiles;
    e.num_m_tiles = p.num_tiles; e.num_n_tiles = 2;
    e.stride = 1;
    e.Cout = C;
    e.mode = EPI_STORE;
    e.residual = h;
    e.out = out;
    e.num_phases = 1;
    e.stats_out = stats_out;
    e.stats_parts = stats_out ? p.Tpad / 32 : 0;
    l->grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    return 0;
}

int attn_block_launch(const AttnBlockLaunch& l, cudaStream_t stream) {
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) {
            set_error("attn_block: cudaFuncSetAttribute failed");
            return 4;
        }
        attr_set[dev] = true;
    }
    cudaError_t e = launch_pdl(attn_block_kernel, dim3(l.grid), dim3(kThreads), kSmemBytes, stream, l.p);
    if (e != cudaSuccess) { set_error("attn_block: launch failed: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // namespace ddpm
