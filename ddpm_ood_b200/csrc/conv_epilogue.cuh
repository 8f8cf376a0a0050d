// Epilogue of one 128-pixel x BN-channel fp32 accumulator tile in TMEM, shared by the conv kernels (conv_gemm.cu,
// conv_halo.cu): + bias (+ per-image channel add) (+ residual) -> fp16 NDHWC, optional GroupNorm partial statistics.
#pragma once
#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace ddpm {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// Epilogue of one 128-pixel x BN-channel accumulator tile; executed by the 4 epilogue warps (q = TMEM lane quarter),
// one output pixel (row) per thread. t_addr: TMEM address of the tile for this warp's lanes. sub: sub-pixel phase.
// s_add (optional): the per-column addend (bias + this row's image's chan_add) staged in shared memory, [slot][BN] fp32
// with slot = 0 for region tiles / the row's image (0, 1) for pair tiles; replaces the global bias / chan_add loads
// (with ~230 KB of shared memory in use the L1 cache is gone and each of those loads is an L2 round trip).
template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvGemmParams& p, uint32_t t_addr, int m_tile, int n_tile,
                                                   int sub, int q, int lane, const float* s_add = nullptr) {
            const int row = q * 32 + lane;

            int t = m_tile;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; t /= p.tiles_h;
            const int td = t % p.tiles_d; t /= p.tiles_d;
            const int tn = t;
            int r = row;
            int w = tw * p.bw + r % p.bw; r /= p.bw;
            int h = th * p.bh + r % p.bh; r /= p.bh;
            const int d = td * p.bd + r % p.bd; r /= p.bd;
            int n = tn * p.bn + r;
            if (p.pair_rows) {  // conv_halo pair tiles: row = h * 16 + n' * 8 + w of images 2 * tile + n'
                w = row & 7;
                h = row >> 4;
                n = m_tile * 2 + ((row >> 3) & 1);
            }
            const bool valid = (w < p.W) && (h < p.H) && (d < p.D) && (n < p.N);
            size_t pix;
            if (p.num_phases > 1) {  // sub-pixel scatter into the doubled output grid
                const int od = p.phase3d ? 2 * d + ((sub >> 2) & 1) : d;
                const int Do = p.phase3d ? 2 * p.D : p.D;
                pix = ((static_cast<size_t>(n) * Do + od) * (2 * p.H) + (2 * h + ((sub >> 1) & 1))) * (2 * p.W) +
                      (2 * w + (sub & 1));
            } else {
                pix = ((static_cast<size_t>(n) * p.D + d) * p.H + h) * p.W + w;
            }

            if (p.mode == EPI_SOFTMAX_BD) {
                // Attention probabilities. Row = query token; its keys are the `group` columns of its own image:
                // columns [g0, g0+group) of this tile when group <= 128 (block diagonal), all BN columns otherwise.
                const int G = p.group < BN ? p.group : BN;
                const int g0 = p.group < BN ? (row / G) * G : 0;
                float mx = -INFINITY;
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_addr + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        if (col >= g0 && col < g0 + G) mx = fmaxf(mx, __uint_as_float(v[j]) * p.scale);
                    }
                }
                float sum = 0.f;
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_addr + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        if (col >= g0 && col < g0 + G) sum += __expf(__uint_as_float(v[j]) * p.scale - mx);
                    }
                }
                const float inv = 1.0f / sum;
                __half* dst = p.out + pix * static_cast<size_t>(p.Cout) + static_cast<size_t>(n_tile) * BN;
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_addr + c * 32, v);
                    ptx::tmem_ld_wait();
                    if (valid) {
                        uint32_t packed[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const int col = c * 32 + j;
                            float a = 0.f, b = 0.f;
                            if (col >= g0 && col < g0 + G) {
                                a = __expf(__uint_as_float(v[j]) * p.scale - mx) * inv;
                                b = __expf(__uint_as_float(v[j + 1]) * p.scale - mx) * inv;
                            }
                            __half2 hh = __floats2half2_rn(a, b);
                            packed[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    }
                }
            } else {
                const int col_base = n_tile * BN;
                const float* cadd = p.chan_add ? p.chan_add + static_cast<size_t>(valid ? n : 0) * p.chan_add_stride : nullptr;
                // GroupNorm partial statistics: all 32 rows of this warp lie in one image (host guarantees it)
                float* st_base = nullptr;
                if (p.stats_out && p.pair_rows) {
                    // lanes 0-7 / 16-23 hold image 0, lanes 8-15 / 24-31 image 1; part = this warp
                    const int n_l = m_tile * 2 + ((lane >> 3) & 1);
                    if (n_l < p.N) st_base = p.stats_out + (static_cast<size_t>(n_l) * p.stats_parts + q) * (p.Cout >> 1);
                } else if (p.stats_out) {
                    const int R = p.bw * p.bh * p.bd;  // pixels of one image inside the tile box (multiple of 32)
                    const int n_w = tn * p.bn + (q * 32) / R;
                    if (n_w < p.N) {
                        const int tile_sp = (td * p.tiles_h + th) * p.tiles_w + tw;
                        const int part = sub * (p.stats_parts / p.num_phases) + tile_sp * (R >> 5) + ((q * 32) % R) / 32;
                        st_base = p.stats_out + (static_cast<size_t>(n_w) * p.stats_parts + part) * (p.Cout >> 1);
                    }
                }
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_addr + c * 32, v);
                    ptx::tmem_ld_wait();
                    const int col0 = col_base + c * 32;
                    uint32_t packed[16];
                    if (valid) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (s_add) {
                            const float4* sa4 = reinterpret_cast<const float4*>(
                                s_add + (p.pair_rows ? ((row >> 3) & 1) * BN : 0) + c * 32);
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = sa4[j >> 2];
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        if (!s_add && p.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        if (!s_add && cadd) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(cadd + (p.chan_mod ? col0 % p.chan_mod : col0) + j));
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        if (p.residual) {
                            const uint4* r4 =
                                reinterpret_cast<const uint4*>(p.residual + pix * static_cast<size_t>(p.Cout) + col0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 rv = __ldg(r4 + j);
                                const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 ff = __half22float2(h2[e]);
                                    f[8 * j + 2 * e] += ff.x;
                                    f[8 * j + 2 * e + 1] += ff.y;
                                }
                            }
                        }
                        if (p.residual_lo) {
                            const uint4* r4 =
                                reinterpret_cast<const uint4*>(p.residual_lo + pix * static_cast<size_t>(p.Cout) + col0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 rv = __ldg(r4 + j);
                                const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 ff = __half22float2(h2[e]);
                                    f[8 * j + 2 * e] += ff.x;
                                    f[8 * j + 2 * e + 1] += ff.y;
                                }
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
                        if (p.mode == EPI_STORE_F32) {
                            float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * static_cast<size_t>(p.Cout) + col0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) d4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
#pragma unroll
                            for (int j = 0; j < 16; ++j) packed[j] = 0u;
                        } else if (p.mode == EPI_STORE_VT && col0 >= p.vt_col0) {
                            // transposed store for the attention V operand: out_vt[pair][c][token-in-pair], where a
                            // "pair" is the 128 consecutive tokens of one M tile.
                            const size_t tok = pix;  // token index == pixel index (N*T rows)
                            const size_t pair = tok >> 7;
                            const int tin = static_cast<int>(tok & 127);
                            const int vc = col0 - p.vt_col0;
                            const int vC = p.Cout - p.vt_col0;
                            __half* dv = p.out_vt + (pair * vC + vc) * 128 + tin;
#pragma unroll
                            for (int j = 0; j < 32; ++j) dv[static_cast<size_t>(j) * 128] = __float2half_rn(f[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                __half2 hh = __floats2half2_rn(f[j], f[j + 1]);
                                packed[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                            }
                            uint4* d4 = reinterpret_cast<uint4*>(p.out + pix * static_cast<size_t>(p.Cout) + col0);
                            if (!(p.dbg & 16)) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                            }
                            if (p.out_lo) {  // lo half: what fp16 rounding of the hi half lost
                                uint32_t plo[16];
#pragma unroll
                                for (int j = 0; j < 32; j += 2) {
                                    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&packed[j >> 1]));
                                    __half2 ll = __floats2half2_rn(f[j] - hi.x, f[j + 1] - hi.y);
                                    plo[j >> 1] = *reinterpret_cast<uint32_t*>(&ll);
                                }
                                uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + pix * static_cast<size_t>(p.Cout) + col0);
#pragma unroll
                                for (int j = 0; j < 4; ++j) l4[j] = make_uint4(plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) packed[j] = 0u;
                    }
                    if (p.stats_out && !(p.dbg & 8)) {
                        // per-lane quad sums of the ROUNDED values (what GroupNorm will read back), then a
                        // transpose-reduce over the warp's 32 pixels: 16 shuffles instead of 80.
                        float sv[16];
#pragma unroll
                        for (int qd = 0; qd < 8; ++qd) {
                            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&packed[2 * qd]));
                            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&packed[2 * qd + 1]));
                            sv[qd] = (a.x + a.y) + (b.x + b.y);
                            sv[8 + qd] = (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y);
                        }
                        if (p.pair_rows) {
                            // two images per warp: add the two rows of each image (lane ^ 16), then transpose-reduce
                            // inside the 8-lane groups: 16 + 14 shuffles; lanes 0-15 end with two values each
#pragma unroll
                            for (int i = 0; i < 16; ++i) sv[i] += __shfl_xor_sync(0xffffffffu, sv[i], 16);
#pragma unroll
                            for (int half_n = 8, off = 4; half_n >= 2; half_n >>= 1, off >>= 1) {
                                const bool hi = (lane & off) != 0;
#pragma unroll
                                for (int i = 0; i < half_n; ++i) {
                                    const float send = hi ? sv[i] : sv[i + half_n];
                                    const float keep = hi ? sv[i + half_n] : sv[i];
                                    sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                                }
                            }
                            // value index = b2 b1 b0 j: b2 -> sum / sum of squares, (b1 b0 j) -> quad
                            if (st_base && lane < 16) {
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const int quad = (col0 >> 2) + (((lane & 3) << 1) | j);
                                    st_base[quad * 2 + ((lane >> 2) & 1)] = sv[j];
                                }
                            }
                        } else {
#pragma unroll
                        for (int half_n = 8, off = 16; half_n >= 1; half_n >>= 1, off >>= 1) {
                            const bool hi = (lane & off) != 0;
#pragma unroll
                            for (int i = 0; i < half_n; ++i) {
                                const float send = hi ? sv[i] : sv[i + half_n];
                                const float keep = hi ? sv[i + half_n] : sv[i];
                                sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                        sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
                        // lane bits: b4 -> sum / sum of squares, b3 b2 b1 -> quad, b0 -> duplicate
                        if (st_base && (lane & 1) == 0) {
                            const int quad = (col0 >> 2) + ((lane >> 1) & 7);
                            st_base[quad * 2 + (lane >> 4)] = sv[0];
                        }
                        }
                    }
                }
            }
}

// Store-mode epilogue of the halo-tile kernel (conv_halo.cu) and of the CTA-pair im2col kernel: same results as conv_epilogue_tile's EPI_STORE path with
// the per-column addend staged in shared memory, restructured for latency - 16-column chunks with double-buffered
// TMEM loads (chunk c + 1 is in flight while chunk c is converted, stored and reduced), half the live registers.
// s_add: [slot][BN] fp32 (bias + chan_add), slot = 0 for region tiles / the row's image for pair tiles.
// c_begin / c_end: range of 16-column chunks this warp handles (two warps per TMEM lane quarter can split a tile).
template <int BN>
__device__ __forceinline__ void conv_epilogue_tile16(const ConvGemmParams& p, uint32_t t_addr, int m_tile, int n_tile,
                                                     int sub, int q, int lane, const float* s_add, int c_begin = 0,
                                                     int c_end = BN / 16) {
    const int row = q * 32 + lane;
    int w, h, d = 0, n;
    int part = 0, n_w = 0;  // statistics part / image of this warp's 32 rows (region and im2col tiles)
    if (p.pair_rows) {  // conv_halo pair tiles: row = h * 16 + n' * 8 + w of images 2 * tile + n'
        w = row & 7;
        h = row >> 4;
        n = m_tile * 2 + ((row >> 3) & 1);
    } else {            // tile box (bw, bh, bd, bn) of the tile grid (tiles_w, tiles_h, tiles_d, tiles_n)
        int t = m_tile;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int td = t % p.tiles_d; t /= p.tiles_d;
        const int tn = t;
        int r = row;
        w = tw * p.bw + r % p.bw; r /= p.bw;
        h = th * p.bh + r % p.bh; r /= p.bh;
        d = td * p.bd + r % p.bd; r /= p.bd;
        n = tn * p.bn + r;
        const int R = p.bw * p.bh * p.bd;  // pixels of one image inside the tile box (a multiple of 32 with statistics)
        n_w = tn * p.bn + (q * 32) / R;
        part = sub * (p.stats_parts / p.num_phases) + ((td * p.tiles_h + th) * p.tiles_w + tw) * (R >> 5) + ((q * 32) % R) / 32;
    }
    const bool valid = (w < p.W) && (h < p.H) && (d < p.D) && (n < p.N);
    size_t pix;
    if (p.num_phases > 1) {  // sub-pixel scatter into the doubled output grid
        const int od = p.phase3d ? 2 * d + ((sub >> 2) & 1) : d;
        const int Do = p.phase3d ? 2 * p.D : p.D;
        pix = ((static_cast<size_t>(n) * Do + od) * (2 * p.H) + (2 * h + ((sub >> 1) & 1))) * (2 * p.W) + (2 * w + (sub & 1));
    } else {
        pix = ((static_cast<size_t>(n) * p.D + d) * p.H + h) * p.W + w;
    }
    const int col_base = n_tile * BN;
    const float* addend = s_add + (p.pair_rows ? ((row >> 3) & 1) * BN : 0);
    float* st_base = nullptr;
    if (p.stats_out && !(p.dbg & 8)) {
        if (p.pair_rows) {
            // lanes 0-7 / 16-23 hold image 0, lanes 8-15 / 24-31 image 1; part = this warp
            const int n_l = m_tile * 2 + ((lane >> 3) & 1);
            if (n_l < p.N)
                st_base = p.stats_out + (static_cast<size_t>(n_l) * p.stats_parts + sub * (p.stats_parts / p.num_phases) + q) * (p.Cout >> 1);
        } else if (n_w < p.N) {  // all 32 rows of a warp lie in one image (host guarantees it)
            st_base = p.stats_out + (static_cast<size_t>(n_w) * p.stats_parts + part) * (p.Cout >> 1);
        }
    }
    const bool want_stats = p.stats_out && !(p.dbg & 8);
    __half* out_row = p.out + pix * static_cast<size_t>(p.Cout) + col_base;
    const __half* res_row = p.residual ? p.residual + pix * static_cast<size_t>(p.Cout) + col_base : nullptr;
    const __half* res_lo_row = p.residual_lo ? p.residual_lo + pix * static_cast<size_t>(p.Cout) + col_base : nullptr;
    __half* out_lo_row = p.out_lo ? p.out_lo + pix * static_cast<size_t>(p.Cout) + col_base : nullptr;

    // one 16-column chunk: bias / addend / residual, fp16 store; leaves the chunk's statistics candidates in sv:
    // quad sums of the ROUNDED values (what GroupNorm reads back), sv[0..3] sums, sv[4..7] sums of squares
    auto process = [&](const int c, const uint32_t (&v)[16], float* sv) {
        uint32_t packed[8];
        if (valid) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            const float4* sa4 = reinterpret_cast<const float4*>(addend + c * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b4 = sa4[j];
                f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
            }
            if (res_row) {
                const uint4* r4 = reinterpret_cast<const uint4*>(res_row + c * 16);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint4 rv = __ldg(r4 + j);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 ff = __half22float2(h2[e]);
                        f[8 * j + 2 * e] += ff.x;
                        f[8 * j + 2 * e + 1] += ff.y;
                    }
                }
            }
            if (res_lo_row) {
                const uint4* r4 = reinterpret_cast<const uint4*>(res_lo_row + c * 16);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint4 rv = __ldg(r4 + j);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 ff = __half22float2(h2[e]);
                        f[8 * j + 2 * e] += ff.x;
                        f[8 * j + 2 * e + 1] += ff.y;
                    }
                }
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                __half2 hh = __floats2half2_rn(f[j], f[j + 1]);
                packed[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            if (!(p.dbg & 16)) ptx::st_global_256(out_row + c * 16, packed);  // rows are 256-byte aligned (Cout % 128 == 0)
            if (out_lo_row) {  // lo half: what fp16 rounding of the hi half lost
                uint32_t plo[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&packed[j >> 1]));
                    __half2 ll = __floats2half2_rn(f[j] - hi.x, f[j + 1] - hi.y);
                    plo[j >> 1] = *reinterpret_cast<uint32_t*>(&ll);
                }
                uint4* l4 = reinterpret_cast<uint4*>(out_lo_row + c * 16);
                l4[0] = make_uint4(plo[0], plo[1], plo[2], plo[3]);
                l4[1] = make_uint4(plo[4], plo[5], plo[6], plo[7]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) packed[j] = 0u;
        }
        if (want_stats) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&packed[2 * qd]));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&packed[2 * qd + 1]));
                sv[qd] = (a.x + a.y) + (b.x + b.y);
                sv[4 + qd] = (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y);
            }
        }
    };
    // statistics of a chunk PAIR (c, c + 1): sv[0..7] chunk c, sv[8..15] chunk c + 1. One halving butterfly tree over the
    // warp for both chunks (half the dependent shuffle stages per chunk of the one-chunk version).
    auto reduce_pair = [&](const int c, float (&sv)[16]) {
        const int quad0 = (col_base + c * 16) >> 2;
        if (p.pair_rows) {
            // two images per warp (lanes 0-7 / 16-23 image 0, 8-15 / 24-31 image 1). xor 16 joins the two rows of an image
            // and splits the chunk pair; 4, 2, 1 halve inside the 8-lane groups. A lane ends with one value:
            // b4 -> chunk, b3 -> image, b2 -> sum / squares, b1 b0 -> quad
            {
                const bool hi = (lane & 16) != 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float send = hi ? sv[i] : sv[i + 8];
                    const float keep = hi ? sv[i + 8] : sv[i];
                    sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
#pragma unroll
            for (int half_n = 4, off = 4; half_n >= 1; half_n >>= 1, off >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < half_n; ++i) {
                    const float send = hi ? sv[i] : sv[i + half_n];
                    const float keep = hi ? sv[i + half_n] : sv[i];
                    sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            if (st_base) st_base[(quad0 + 4 * (lane >> 4) + (lane & 3)) * 2 + ((lane >> 2) & 1)] = sv[0];
        } else {
            // 32 pixels of one image: b4 -> chunk, b3 -> sum / squares, b2 b1 -> quad, then one butterfly over b0
#pragma unroll
            for (int half_n = 8, off = 16; half_n >= 1; half_n >>= 1, off >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < half_n; ++i) {
                    const float send = hi ? sv[i] : sv[i + half_n];
                    const float keep = hi ? sv[i + half_n] : sv[i];
                    sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
            if (st_base && (lane & 1) == 0)
                st_base[(quad0 + 4 * (lane >> 4) + ((lane >> 1) & 3)) * 2 + ((lane >> 3) & 1)] = sv[0];
        }
    };

    uint32_t va[16], vb[16];
    ptx::tmem_ld_32x16(t_addr + c_begin * 16, va);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 2) {  // even chunk counts
        float sv[16];
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x16(t_addr + (c + 1) * 16, vb);
        process(c, va, sv);
        ptx::tmem_ld_wait();
        if (c + 2 < c_end) ptx::tmem_ld_32x16(t_addr + (c + 2) * 16, va);
        process(c + 1, vb, sv + 8);
        if (want_stats) reduce_pair(c, sv);
    }
}

}  // namespace ddpm
