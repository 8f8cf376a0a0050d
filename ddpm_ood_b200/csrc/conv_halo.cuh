// Halo-tile 3x3 convolution on tcgen05 tensor cores (sm_100a, CTA pairs): the stride-1 2-D 3x3 convs of
// DiffusionModelUNet's ResnetBlocks (reached from src/trainers/reconstruct.py:150-153), which carry ~85 % of the
// UNet's FLOPs at the 32x32 / 16x16 levels.
//
// conv_gemm.cu fetches a fresh 128-pixel im2col tile for every one of the 9 taps; measured (profiles/
// r01_conv_ncu_full_s3.md) that operand feed, not the tensor pipe, paces the kernel. Here an M tile is an 8 x 16 pixel
// region of one image and, per 64-channel chunk, its 10 x 18 haloed input is staged ONCE in shared memory (one 5-D TMA
// box, zero padding = TMA out-of-bounds fill). Pixel (h', w') lies in row h' * 10 + w' of a K-major SWIZZLE_128B tile,
// so tap (dh, dw) is the SAME tile read through a UMMA descriptor whose start address is shifted by
// ((1 + dh) * 10 + (1 + dw)) * 128 B and whose 8-row-group stride is 10 rows (1280 B): the swizzle is a function of the
// shared-memory address bits, so shifted, non-1024-B-aligned starts address the right 16-byte chunks (validated on
// B200 by experiments/umma_halo_probe.cu). A-operand traffic drops ~6x.
//
// Because the input tile now sits in shared memory exactly once, GroupNorm + SiLU of the conv's INPUT is applied there
// in place (4 transform warps: y = silu(x * a[n,c] + b[n,c]) -> fp16, padding stays zero) before the MMAs read it: the
// normalised activation never exists in HBM and the gn_apply launches disappear. a/b come from gn_finalize()
// (kernels.cuh), computed from the producer's epilogue statistics.
//
// Up to kMaxSeg K-segments share the accumulator: 3x3 segments (the conv input, possibly a channel concatenation of two
// tensors) and 1x1 segments over raw tensors (the ResnetBlock skip conv). Weights: the same [Cout][K] fp16 matrix as
// conv_gemm (K ordered segment, tap, channel); each CTA of a pair stages half of every weight tile.
// Warp roles: 0 A producer, 1 MMA issuer (leader CTA), 2 TMEM allocator, 3 B producer, 4-7 epilogue, 8-11 transform.
#pragma once
#include "conv_gemm.cuh"

namespace ddpm {

// GroupNorm of the conv's input computed from its producers' partial statistics (ConvGemmParams::stats_out layout) by
// the kernel itself, per work item: replaces a separate statistics -> scale/shift launch. The normalised channels are
// the concatenation of (up to) two tensors, in segment order.
struct HaloGnSource {
    const float* st0; int parts0; int C0;
    const float* st1; int parts1; int C1;   // st1 null / C1 0: one tensor
    const float* gamma; const float* beta;  // [C0 + C1]
    int S;                                  // pixels per image
    int groups;
    float eps;
};

constexpr int kMaxStagesPerItem = 48;  // (segment, 64-channel chunk) stages of one work item

struct ConvHaloParams {
    ConvGemmParams g;         // geometry, K schedule, epilogue, tmA[seg] (haloed / plain boxes), tmB (half weight tile)
    int n_stages;             // K-loop schedule: stage i stages chunk sched_chunk[i] of segment sched_seg[i]
    uint8_t sched_seg[kMaxStagesPerItem], sched_chunk[kMaxStagesPerItem];
    uint8_t sched_kd[kMaxStagesPerItem];  // 3-D: depth tap (0..2) of a 3x3x3 segment's stage (9 in-plane taps each); else 0
    int slabs;                // 3-D volumes of 8 x 8 slabs: depth D of an image (a pair tile = two consecutive depth slabs
                              // of one image, g.N counts slabs); 1 for 2-D problems
    int seg_taps[kMaxSeg];    // 9 or 1
    int seg_cin[kMaxSeg];     // channels of the segment (tap stride along K)
    int seg_kcol0[kMaxSeg];   // first K column of the segment in the weight matrix
    int seg_gn[kMaxSeg];      // 1: apply silu(x * a + b) to this segment's input tile, 2: x * a + b (no activation)
    int seg_ab_off[kMaxSeg];  // channel offset of the segment in the scale/shift table
    const float2* ab;         // [N][ab_C] (scale, shift) per image and input channel, or null
    HaloGnSource gn;          // gn_from_stats: derive (scale, shift) from these statistics instead of reading `ab`
    int gn_from_stats;
    int ab_C;
    int pair_mode;            // 1: tiles are two whole images of up to 8 x 8 pixels (rows interleaved by image)
    long long* dbg_cycles;    // timing experiments only (env DDPM_HALO_CYCLES): per leader CTA [total, wait tempty,
                              // wait a_ready, wait b_full] cycles of the MMA warp
    int dbg;                  // timing experiments only (env DDPM_HALO_DBG): 1 skip transform math, 2 skip epilogue
                              // body, 4 skip the MMAs; results are wrong when non-zero
};

struct ConvHaloLaunch {
    ConvHaloParams p;
    int block_n;          // 256 (one M tile per CTA) or 128 (two M tiles per CTA)
    int m_tiles_per_cta;
    int grid;
};

// stride-1 problems of 3x3 / 1x1 segments, store epilogue. 2-D: any size (images of up to 8 x 8 pixels need
// Cout % 256 == 0). 3-D: volumes of 8 x 8 slabs (H == W == 8) with D % 4 == 0 (D % 2 == 0 when Cout % 256 == 0): a 3x3x3
// segment runs as three depth-tap stages per 64 channels, each the nine in-plane taps of one haloed pair tile.
bool conv_halo_supported(const ConvProblem& q);
// GroupNorm-statistics parts per image emitted by this kernel's epilogue for a (D x) H x W output
int conv_halo_stats_parts(int H, int W, int D = 1);
// gn_ab: null, or the (scale, shift) table [N][gn_ab_channels] applied to the 3x3 segments (channels concatenated in
// segment order)
// gn_src (optional, instead of gn_ab): compute the table inside the kernel from producer statistics.
int conv_halo_prepare(const ConvProblem& q, const float* gn_ab, int gn_ab_channels, int num_sms, ConvHaloLaunch* out,
                      const HaloGnSource* gn_src = nullptr);
int conv_halo_launch(const ConvHaloLaunch& l, cudaStream_t stream);
int conv_halo_read_cycles(const ConvHaloLaunch& l, long long* host, int n);  // experiment only (DDPM_HALO_CYCLES)

}  // namespace ddpm
