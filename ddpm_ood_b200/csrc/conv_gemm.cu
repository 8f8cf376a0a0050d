// tcgen05 / TMA implicit-GEMM convolution kernel. See conv_gemm.cuh for the contract.
#include "conv_gemm.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_epilogue.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace ddpm {

// ------------------------------------------------------------------------------------------------ error string
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ kernel
// MT = M tiles (128 pixels each) a CTA works on at once, sharing every weight tile: <256,1> and <128,2> both move
// 48 KB per 64-wide k-block for 2*128*256*64 MACs; <128,1> moves 32 KB for half the MACs and is TMA/shared-memory
// feed-bound at ~50 % of the tensor pipe (measured, profiles/r01_*), so it is kept only for problems too small to pair.
template <int BN, int MT>
struct Cfg {
    static constexpr int kStages = (BN * MT == 256) ? 4 : 6;
    static constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB per M tile
    static constexpr int kBBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = MT * kABytes + kBBytes;
    static constexpr int kAccCols = MT * BN;      // TMEM columns of one accumulator stage
    static constexpr int kTmemCols = 2 * kAccCols;  // two accumulator stages
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int MT>
__global__ void __launch_bounds__(256, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    using C = Cfg<BN, MT>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                  // [kStages][MT][16 KB]
    uint8_t* smem_b = smem + C::kStages * MT * C::kABytes;   // [kStages][BN * 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full_bar = bars;                     // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + C::kStages;       // [kStages]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * C::kStages;   // [2]        MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + 2;          // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    ptx::pdl_trigger();

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.n_seg; ++s) ptx::prefetch_tmap(&p.tmA[s]);
        ptx::prefetch_tmap(&p.tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();

    const int gpp = (p.num_m_tiles + MT - 1) / MT;  // a work item = MT consecutive M tiles (of one phase) x one N tile
    const int total_tiles = gpp * p.num_phases * p.num_n_tiles;

    if (warp == 0) {
        // ================================================================= TMA producer (one thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int g = tile / p.num_n_tiles;
                const int n_tile = tile - g * p.num_n_tiles;
                const int sub = g / gpp;  // sub-pixel phase
                const int m_group = g - sub * gpp;
                const bool phased = p.num_phases > 1;
                const int pofw = (sub & 1) - 1, pofh = ((sub >> 1) & 1) - 1, pofd = p.phase3d ? ((sub >> 2) & 1) - 1 : 0;
                int w0[MT], h0[MT], d0[MT], n0[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    int t = m_group * MT + mt;  // may run one past the last tile: its box is all out-of-bounds (zeros)
                    const int tw = t % p.tiles_w; t /= p.tiles_w;
                    const int th = t % p.tiles_h; t /= p.tiles_h;
                    const int td = t % p.tiles_d; t /= p.tiles_d;
                    w0[mt] = tw * p.bw * p.stride;
                    h0[mt] = th * p.bh * p.stride;
                    d0[mt] = td * p.bd * (p.D > 1 ? p.stride : 1);
                    n0[mt] = t * p.bn;
                }
                const int brow = n_tile * BN + m_group * p.b_rows_per_mtile + sub * p.Cout;
                int seg = 0, seg_begin = 0;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    while (kb >= p.seg_kb_end[seg]) { seg_begin = p.seg_kb_end[seg]; ++seg; }
                    const int local = kb - seg_begin;
                    const int chunks = p.seg_chunks[seg];
                    const int tap = local / chunks;
                    const int chunk = local - tap * chunks;
                    const int kw = p.seg_kw[seg], kh = p.seg_kh[seg], kd = p.seg_kd[seg], pad = p.seg_pad[seg];
                    const int iw = tap % kw;
                    const int ih = (tap / kw) % kh;
                    const int id = tap / (kw * kh);
                    const CUtensorMap* ma = seg == 0 ? &p.tmA[0] : (seg == 1 ? &p.tmA[1] : &p.tmA[2]);
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        ptx::tma_load_5d(smem_a + (stage * MT + mt) * C::kABytes, ma, &full_bar[stage], chunk * kBlockK,
                                         w0[mt] + iw + (phased ? pofw : -(kw > 1 ? pad : 0)),
                                         h0[mt] + ih + (phased ? pofh : -(kh > 1 ? pad : 0)),
                                         d0[mt] + id + (phased ? pofd : -(kd > 1 ? pad : 0)), n0[mt]);
                    ptx::tma_load_2d(smem_b + stage * C::kBBytes, &p.tmB, &full_bar[stage], kb * kBlockK, brow);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one elected lane issues; the whole
        // warp walks the loop so descriptors / barrier addresses stay in uniform registers, no R2UR waterfall per MMA)
        {
            constexpr uint32_t idesc = ptx::make_idesc_f16(kBlockM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * C::kAccCols;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    const uint64_t db = ptx::make_desc_k128(ptx::smem_u32(smem_b + stage * C::kBBytes));
                    const uint64_t da = ptx::make_desc_k128(ptx::smem_u32(smem_a + stage * MT * C::kABytes));
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                // +32 bytes (>>4 = 2) per 16-element K step inside the 128B swizzle row
                                ptx::umma_f16(d_tmem + mt * BN, da + (mt * (C::kABytes >> 4) + 2 * k), db + 2 * k, idesc,
                                              (kb | k) != 0);
                            }
                        }
                        ptx::umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
                        if (kb == p.num_kb - 1) ptx::umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue (4 warps, 1 row per thread)
        const int q = warp - 4;  // TMEM lane quarter == warp % 4
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int g = tile / p.num_n_tiles;
            const int n_tile = tile - g * p.num_n_tiles;
            const int sub = g / gpp;  // sub-pixel phase
            const int m_group = g - sub * gpp;
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
            conv_epilogue_tile<BN>(p, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * C::kAccCols + mt * BN,
                                   m_group * MT + mt, n_tile, sub, q, lane);
            }  // mt
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// Same contract, run by clusters of two CTAs driving ONE tcgen05.mma.cta_group::2 stream (M = 256): each CTA stages
// its own 128-pixel A tile(s) but only HALF of every weight tile, so the L2 -> SM traffic per MAC drops from
// (16 + 32) KB to (16 + 16) KB per 128x256x64 block (BN = 256) - the feed, not the tensor pipe, bounds the single-CTA
// kernel (profiles/r01_conv_ncu_full.md). The even CTA (leader) issues the MMAs; its full barriers collect the TMA
// bytes of both CTAs; tcgen05.commit multicasts "stage free" / "accumulator ready" to both.
template <int BN, int MT>
struct Cfg2 {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBHalfBytes = (BN / 2) * kBlockK * 2;
    static constexpr int kStageBytes = MT * kABytes + kBHalfBytes;  // per CTA
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kAccCols = MT * BN;
    static constexpr int kTmemCols = 2 * kAccCols;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + BN * 4 /*epilogue addend row*/;
};

// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 and 8-11 epilogue (two warps per TMEM lane quarter, each
// takes half of a tile's columns: with K = 256..1024 per item these kernels are epilogue-bound).
constexpr int kThreads2 = 384;
template <int BN, int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
    conv_gemm_2cta_kernel(const __grid_constant__ ConvGemmParams p) {
    using C = Cfg2<BN, MT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                      // [kStages][MT][16 KB]
    uint8_t* smem_b = smem + C::kStages * MT * C::kABytes;       // [kStages][BN/2 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full_bar = bars;                    // leader's copy is used: TMA of both CTAs -> MMA
    uint64_t* empty_bar = bars + C::kStages;      // per CTA: MMA (multicast commit) -> this CTA's producer
    uint64_t* tfull_bar = bars + 2 * C::kStages;  // per CTA: MMA (multicast commit) -> this CTA's epilogue
    uint64_t* tempty_bar = tfull_bar + 2;         // leader's copy: epilogue warps of both CTAs -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_add = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + 256);  // [BN] bias of the item's N tile

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader
    ptx::pdl_trigger();

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.n_seg; ++s) ptx::prefetch_tmap(&p.tmA[s]);
        ptx::prefetch_tmap(&p.tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 16);  // one arrive per epilogue warp (8) of both CTAs
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc_2cta<C::kTmemCols>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();  // the peer's barriers must exist before anything is signalled across the pair
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();

    // work item = 2*MT consecutive M tiles (MT per CTA) of one sub-pixel phase x one N tile
    const int gpp = (p.num_m_tiles + 2 * MT - 1) / (2 * MT);
    const int total_items = gpp * p.num_phases * p.num_n_tiles;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ================================================================= TMA producer (one thread per CTA)
        if (lane == 0) {
            int stage = 0;
            uint32_t parity = 0;
            for (int item = cluster_id; item < total_items; item += num_clusters) {
                const int g = item / p.num_n_tiles;
                const int n_tile = item - g * p.num_n_tiles;
                const int sub = g / gpp;
                const int m_group = g - sub * gpp;
                const bool phased = p.num_phases > 1;
                const int pofw = (sub & 1) - 1, pofh = ((sub >> 1) & 1) - 1, pofd = p.phase3d ? ((sub >> 2) & 1) - 1 : 0;
                int w0[MT], h0[MT], d0[MT], n0[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    int t = (m_group * 2 + static_cast<int>(rank)) * MT + mt;  // past-the-end tiles read zeros
                    const int tw = t % p.tiles_w; t /= p.tiles_w;
                    const int th = t % p.tiles_h; t /= p.tiles_h;
                    const int td = t % p.tiles_d; t /= p.tiles_d;
                    w0[mt] = tw * p.bw * p.stride;
                    h0[mt] = th * p.bh * p.stride;
                    d0[mt] = td * p.bd * (p.D > 1 ? p.stride : 1);
                    n0[mt] = t * p.bn;
                }
                const int brow = n_tile * BN + static_cast<int>(rank) * (BN / 2) + sub * p.Cout;
                int seg = 0, seg_begin = 0;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    while (kb >= p.seg_kb_end[seg]) { seg_begin = p.seg_kb_end[seg]; ++seg; }
                    const int local = kb - seg_begin;
                    const int chunks = p.seg_chunks[seg];
                    const int tap = local / chunks;
                    const int chunk = local - tap * chunks;
                    const int kw = p.seg_kw[seg], kh = p.seg_kh[seg], kd = p.seg_kd[seg], pad = p.seg_pad[seg];
                    const int iw = tap % kw;
                    const int ih = (tap / kw) % kh;
                    const int id = tap / (kw * kh);
                    const CUtensorMap* ma = seg == 0 ? &p.tmA[0] : (seg == 1 ? &p.tmA[1] : &p.tmA[2]);
                    ptx::mbar_wait(&empty_bar[stage], parity ^ 1);
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        ptx::tma_load_5d_2cta(smem_a + (stage * MT + mt) * C::kABytes, ma, &full_bar[stage],
                                              chunk * kBlockK, w0[mt] + iw + (phased ? pofw : -(kw > 1 ? pad : 0)),
                                              h0[mt] + ih + (phased ? pofh : -(kh > 1 ? pad : 0)),
                                              d0[mt] + id + (phased ? pofd : -(kd > 1 ? pad : 0)), n0[mt]);
                    ptx::tma_load_2d_2cta(smem_b + stage * C::kBHalfBytes, &p.tmB, &full_bar[stage], kb * kBlockK, brow);
                    if (++stage == C::kStages) { stage = 0; parity ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (leader CTA; one elected lane)
        if (rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(2 * kBlockM, BN);
            int stage = 0;
            uint32_t parity = 0;
            int as = 0;
            uint32_t aparity = 0;
            for (int item = cluster_id; item < total_items; item += num_clusters) {
                ptx::mbar_wait(&tempty_bar[as], aparity ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * C::kAccCols;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    const uint64_t db = ptx::make_desc_k128(ptx::smem_u32(smem_b + stage * C::kBHalfBytes));
                    const uint64_t da = ptx::make_desc_k128(ptx::smem_u32(smem_a + stage * MT * C::kABytes));
                    ptx::mbar_wait(&full_bar[stage], parity);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt)
                                ptx::umma_f16_2cta(d_tmem + mt * BN, da + (mt * (C::kABytes >> 4) + 2 * k), db + 2 * k, idesc,
                                                   (kb | k) != 0);
                        }
                        ptx::umma_commit_2cta(&empty_bar[stage]);  // frees this stage in both CTAs
                        if (kb == p.num_kb - 1) ptx::umma_commit_2cta(&tfull_bar[as]);  // accumulators ready, both CTAs
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; parity ^= 1; }
                }
                if (++as == 2) { as = 0; aparity ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue (8 warps per CTA, own 128 rows)
        const int q = warp & 3;            // TMEM lane quarter == warp % 4
        const int half = (warp - 4) >> 2;  // which half of the tile's columns
        int as = 0;
        uint32_t aparity = 0;
        for (int item = cluster_id; item < total_items; item += num_clusters) {
            const int g = item / p.num_n_tiles;
            const int n_tile = item - g * p.num_n_tiles;
            const int sub = g / gpp;
            const int m_group = g - sub * gpp;
            // fast path (plain store epilogue without a per-image addend): bias row staged in shared memory, 16-column
            // chunks with double-buffered TMEM loads (conv_epilogue_tile16), columns split between the two warp groups
            const bool fast = p.mode == EPI_STORE && p.chan_add == nullptr;
            if (fast) {
                asm volatile("bar.sync 2, 256;" ::: "memory");  // the previous item's readers are done
                for (int i = threadIdx.x - 128; i < BN; i += 256) s_add[i] = p.bias ? __ldg(p.bias + n_tile * BN + i) : 0.f;
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            ptx::mbar_wait(&tfull_bar[as], aparity);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * C::kAccCols + mt * BN;
                const int m_tile = (m_group * 2 + static_cast<int>(rank)) * MT + mt;
                if (fast)
                    conv_epilogue_tile16<BN>(p, t_addr, m_tile, n_tile, sub, q, lane, s_add, half * (BN / 32),
                                             (half + 1) * (BN / 32));
                else if (half == 0)
                    conv_epilogue_tile<BN>(p, t_addr, m_tile, n_tile, sub, q, lane);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(&tempty_bar[as]);
            if (++as == 2) { as = 0; aparity ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();  // neither CTA may retire while the other can still signal into it
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2cta<C::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ host side
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
            set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?): %s", cudaGetErrorString(e));
            return nullptr;
        }
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

static int pow2ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void tile_box(int W, int H, int D, int* bw, int* bh, int* bd, int* bn);

int conv_stats_parts(int spatial_dims, int Dout, int Hout, int Wout) {
    (void)spatial_dims;
    int bw, bh, bd, bn;
    tile_box(Wout, Hout, Dout, &bw, &bh, &bd, &bn);
    const int R = bw * bh * bd;
    if (R < 32) return 0;
    return ((Wout + bw - 1) / bw) * ((Hout + bh - 1) / bh) * ((Dout + bd - 1) / bd) * (R / 32);
}

// tile box: fill W first, then H, D, N; product is always 128 output pixels.
static void tile_box(int W, int H, int D, int* bw, int* bh, int* bd, int* bn) {
    *bw = pow2ceil(W) < kBlockM ? pow2ceil(W) : kBlockM;
    int rem = kBlockM / *bw;
    *bh = pow2ceil(H) < rem ? pow2ceil(H) : rem;
    rem /= *bh;
    *bd = pow2ceil(D) < rem ? pow2ceil(D) : rem;
    rem /= *bd;
    *bn = rem;
}

int conv_prepare(const ConvProblem& q, int num_sms, ConvLaunch* out) {
    PFN_encodeTiled encode = get_encode();
    if (!encode) return 1;
    memset(out, 0, sizeof(*out));
    ConvGemmParams& p = out->p;
    if (q.n_seg < 1 || q.n_seg > kMaxSeg) { set_error("conv: n_seg=%d out of range", q.n_seg); return 2; }
    if (q.stride != 1 && q.stride != 2) { set_error("conv: stride %d unsupported", q.stride); return 2; }
    if (q.upsample2 && (q.n_seg != 1 || q.seg[0].ksize != 2 || q.stride != 1 || q.mode != EPI_STORE || q.residual ||
                        q.b_rows_per_mtile)) {
        set_error("conv: upsample2 needs one 2x2 segment, stride 1, store epilogue, no residual");
        return 2;
    }
    const int sd = q.spatial_dims;
    if (sd != 2 && sd != 3) { set_error("conv: spatial_dims %d unsupported", sd); return 2; }
    if (sd == 2 && q.D != 1) { set_error("conv: 2-D problem needs D == 1"); return 2; }
    const int sW = q.stride, sH = q.stride, sD = (sd == 3) ? q.stride : 1;
    p.N = q.N;
    p.W = (q.W + sW - 1) / sW;
    p.H = (q.H + sH - 1) / sH;
    p.D = (q.D + sD - 1) / sD;
    p.stride = q.stride;
    tile_box(p.W, p.H, p.D, &p.bw, &p.bh, &p.bd, &p.bn);
    p.tiles_w = (p.W + p.bw - 1) / p.bw;
    p.tiles_h = (p.H + p.bh - 1) / p.bh;
    p.tiles_d = (p.D + p.bd - 1) / p.bd;
    p.tiles_n = (p.N + p.bn - 1) / p.bn;
    p.num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
    const int BN = (q.Cout % 256 == 0) ? 256 : 128;
    if (q.Cout % BN != 0) { set_error("conv: Cout=%d must be a multiple of 128", q.Cout); return 2; }
    out->block_n = BN;
    p.num_n_tiles = q.Cout / BN;
    p.Cout = q.Cout;
    p.b_rows_per_mtile = q.b_rows_per_mtile;
    p.mode = q.mode;
    p.relu = q.relu;
    if (q.relu && q.mode != EPI_STORE) { set_error("conv: relu needs the store epilogue"); return 2; }
    p.bias = q.bias;
    p.chan_add = q.chan_add;
    p.chan_add_stride = q.chan_add_stride;
    p.chan_mod = q.chan_mod;
    if (q.chan_mod && (q.chan_mod % 32 != 0 || q.Cout % q.chan_mod != 0)) { set_error("conv: chan_mod=%d unsupported", q.chan_mod); return 2; }
    p.residual = static_cast<const __half*>(q.residual);
    p.out = static_cast<__half*>(q.out);
    p.out_lo = static_cast<__half*>(q.out_lo);
    p.residual_lo = static_cast<const __half*>(q.residual_lo);
    if ((q.out_lo || q.residual_lo) && q.mode != EPI_STORE) { set_error("conv: split-precision output needs the store epilogue"); return 2; }
    p.scale = q.scale;
    p.group = q.group;
    p.vt_col0 = q.vt_col0;
    p.out_vt = static_cast<__half*>(q.out_vt);
    p.num_phases = q.upsample2 ? (sd == 3 ? 8 : 4) : 1;
    p.phase3d = (q.upsample2 && sd == 3) ? 1 : 0;
    p.stats_out = nullptr;
    p.stats_parts = 0;
    if (q.stats_out) {
        const int R = p.bw * p.bh * p.bd;
        if (q.mode != EPI_STORE || R < 32) { set_error("conv: fused GroupNorm statistics unsupported for this shape/mode"); return 2; }
        p.stats_out = q.stats_out;
        p.stats_parts = p.tiles_w * p.tiles_h * p.tiles_d * (R / 32) * p.num_phases;
    }
    if (q.mode == EPI_SOFTMAX_BD && p.num_n_tiles != 1) { set_error("conv: softmax epilogue needs one N tile"); return 2; }

    p.n_seg = q.n_seg;
    int kb = 0;
    size_t ktot = 0;
    for (int s = 0; s < q.n_seg; ++s) {
        const ConvSegment& g = q.seg[s];
        if (g.channels % kBlockK != 0) { set_error("conv: segment channels %d not a multiple of 64", g.channels); return 2; }
        if (g.ksize != 1 && g.ksize != 3 && g.ksize != 5 && !(g.ksize == 2 && q.upsample2) && !(g.ksize == 4 && q.stride == 2)) {
            set_error("conv: ksize %d unsupported", g.ksize);
            return 2;
        }
        p.seg_pad[s] = q.pad > 0 ? q.pad : g.ksize >> 1;
        if (g.ksize == 4 && p.seg_pad[s] != 1) { set_error("conv: 4-tap stride-2 convs need pad 1"); return 2; }
        p.seg_chunks[s] = g.channels / kBlockK;
        p.seg_kw[s] = g.ksize;
        p.seg_kh[s] = g.ksize;
        p.seg_kd[s] = (sd == 3) ? g.ksize : 1;
        const int taps = p.seg_kw[s] * p.seg_kh[s] * p.seg_kd[s];
        kb += taps * p.seg_chunks[s];
        p.seg_kb_end[s] = kb;
        ktot += static_cast<size_t>(taps) * g.channels;
        // A tensor map over the INPUT tensor (C, W, H, D, N)
        cuuint64_t gdim[5] = {static_cast<cuuint64_t>(g.channels), static_cast<cuuint64_t>(q.W),
                              static_cast<cuuint64_t>(q.H), static_cast<cuuint64_t>(q.D),
                              static_cast<cuuint64_t>(q.N)};
        cuuint64_t gstr[4];
        gstr[0] = static_cast<cuuint64_t>(g.channels) * 2;
        gstr[1] = gstr[0] * q.W;
        gstr[2] = gstr[1] * q.H;
        gstr[3] = gstr[2] * q.D;
        cuuint32_t box[5] = {kBlockK, static_cast<cuuint32_t>(p.bw * sW), static_cast<cuuint32_t>(p.bh * sH),
                             static_cast<cuuint32_t>(p.bd * sD), static_cast<cuuint32_t>(p.bn)};
        cuuint32_t estr[5] = {1, static_cast<cuuint32_t>(sW), static_cast<cuuint32_t>(sH),
                              static_cast<cuuint32_t>(sD), 1};
        CUresult r = encode(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(g.ptr), gdim, gstr, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv: cuTensorMapEncodeTiled(A seg %d) failed: %d", s, (int)r); return 3; }
    }
    for (int s = q.n_seg; s < kMaxSeg; ++s) p.seg_kb_end[s] = kb;
    p.num_kb = kb;
    const char* pair_env = getenv("DDPM_CONV_2CTA");  // tests: 0 = single-CTA tiles only (read at plan time)
    const int allow_pair = (pair_env && atoi(pair_env) == 0) ? 0 : 1;
    const int tiles = p.num_m_tiles * p.num_phases;  // 128-pixel tiles per N tile
    out->m_tiles_per_cta = 1;
    out->cta_pair = 0;
    if (allow_pair && q.impl == 0 && q.b_rows_per_mtile == 0 && q.mode == EPI_STORE && p.num_m_tiles >= 2) {
        // CTA pairs share every weight tile (and, for 128-wide outputs with enough work, two M tiles per CTA as well)
        out->cta_pair = 1;
        if (BN == 128 && tiles * p.num_n_tiles >= 4 * num_sms) out->m_tiles_per_cta = 2;
    }
    {
        cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(q.w_rows)};
        cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ktot) * 2};
        // a CTA of a pair stages half of the weight tile's rows
        cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(out->cta_pair ? BN / 2 : BN)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(q.weights), gdim, gstr, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv: cuTensorMapEncodeTiled(B) failed: %d", (int)r); return 3; }
    }
    if (out->cta_pair) {
        const int per_item = 2 * out->m_tiles_per_cta;
        const int items = ((p.num_m_tiles + per_item - 1) / per_item) * p.num_phases * p.num_n_tiles;
        const int clusters = items < num_sms / 2 ? items : num_sms / 2;
        out->grid = 2 * clusters;
        return 0;
    }
    // single-CTA kernel; pair M tiles (MT = 2) for 128-wide outputs when there is enough work
    if (BN == 128 && q.b_rows_per_mtile == 0 && q.mode == EPI_STORE && tiles * p.num_n_tiles >= 2 * num_sms)
        out->m_tiles_per_cta = 2;
    const int groups = (p.num_m_tiles + out->m_tiles_per_cta - 1) / out->m_tiles_per_cta;
    const int total = groups * p.num_phases * p.num_n_tiles;
    out->grid = total < num_sms ? total : num_sms;
    return 0;
}

int conv_launch(const ConvLaunch& l, cudaStream_t stream) {
    static bool attr_set_dev[kMaxDevices] = {};
    bool& g_attr_set = attr_set_dev[device_slot()];
    if (!g_attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(conv_gemm_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              Cfg<128, 1>::kSmemBytes);
        cudaError_t e2 = cudaFuncSetAttribute(conv_gemm_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              Cfg<256, 1>::kSmemBytes);
        cudaError_t e3 = cudaFuncSetAttribute(conv_gemm_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              Cfg<128, 2>::kSmemBytes);
        cudaError_t e4 = cudaFuncSetAttribute(conv_gemm_2cta_kernel<256, 1>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<256, 1>::kSmemBytes);
        cudaError_t e5 = cudaFuncSetAttribute(conv_gemm_2cta_kernel<128, 1>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 1>::kSmemBytes);
        cudaError_t e6 = cudaFuncSetAttribute(conv_gemm_2cta_kernel<128, 2>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 2>::kSmemBytes);
        const cudaError_t es[6] = {e1, e2, e3, e4, e5, e6};
        for (cudaError_t e : es) {
            if (e != cudaSuccess) {
                set_error("conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
                return 4;
            }
        }
        g_attr_set = true;
    }
    cudaError_t e;
    if (l.cta_pair) {
        if (l.block_n == 256)
            e = launch_pdl(conv_gemm_2cta_kernel<256, 1>, dim3(l.grid), dim3(kThreads2), Cfg2<256, 1>::kSmemBytes, stream, l.p);
        else if (l.m_tiles_per_cta == 2)
            e = launch_pdl(conv_gemm_2cta_kernel<128, 2>, dim3(l.grid), dim3(kThreads2), Cfg2<128, 2>::kSmemBytes, stream, l.p);
        else
            e = launch_pdl(conv_gemm_2cta_kernel<128, 1>, dim3(l.grid), dim3(kThreads2), Cfg2<128, 1>::kSmemBytes, stream, l.p);
    } else if (l.block_n == 256)
        e = launch_pdl(conv_gemm_kernel<256, 1>, dim3(l.grid), dim3(256), Cfg<256, 1>::kSmemBytes, stream, l.p);
    else if (l.m_tiles_per_cta == 2)
        e = launch_pdl(conv_gemm_kernel<128, 2>, dim3(l.grid), dim3(256), Cfg<128, 2>::kSmemBytes, stream, l.p);
    else
        e = launch_pdl(conv_gemm_kernel<128, 1>, dim3(l.grid), dim3(256), Cfg<128, 1>::kSmemBytes, stream, l.p);
    if (e != cudaSuccess) { set_error("conv: launch failed: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // namespace ddpm
