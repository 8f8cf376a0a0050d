// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{segment s} sum_{tap t in s} sum_{c} A_s[pixel shifted by t, c] * W[cout, k(s,t,c)]
//
// * A operands are NDHWC fp16 tensors read by 5-D TMA boxes (C, W, H, D, N); the 3x3(x3) halo and the zero padding
//   come from TMA out-of-bounds zero fill, a stride-2 conv from the tensor map's element strides.
// * Up to three K-segments share one accumulator: the main 3x3 conv plus e.g. the ResnetBlock 1x1 skip conv over the
//   raw block input (one or two concatenated tensors), so skip + conv2 is ONE launch.
// * B operand: fp16 weights [Cout][Ktot], K-major, K ordered segment -> tap -> channel.
// * 128-pixel x BN-channel tiles, fp32 accumulators double-buffered in TMEM, persistent CTAs, warp-specialised:
//   warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
// * Epilogue: + bias[c] (+ per-(image,channel) add, the timestep embedding) (+ residual tensor) -> fp16 NDHWC.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddpm {

constexpr int kMaxSeg = 3;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 B = one swizzle row

enum EpilogueMode : int {
    EPI_STORE = 0,       // out[pixel, c] = acc + bias (+chan_add) (+residual)
    EPI_SOFTMAX_BD = 1,  // block-diagonal softmax over groups of T columns (attention scores), writes P fp16
    EPI_STORE_VT = 2,    // like STORE but columns >= vt_col0 are written transposed ([pair][c][128 tokens])
    EPI_STORE_F32 = 3,   // out is fp32 [pixel][Cout]: acc + bias (the VQ-VAE's last transposed conv keeps fp32 tap products)
};

struct ConvGemmParams {
    CUtensorMap tmA[kMaxSeg];
    CUtensorMap tmB;
    // K schedule
    int n_seg;
    int seg_kb_end[kMaxSeg];  // exclusive prefix end in k-blocks
    int seg_chunks[kMaxSeg];  // channels/64
    int seg_kw[kMaxSeg], seg_kh[kMaxSeg], seg_kd[kMaxSeg];  // tap extents (1, 3, or 4 for the VQ-VAE's stride-2 convs)
    int seg_pad[kMaxSeg];     // zero padding per spatial dim with extent > 1: input coord = out*stride + tap - pad
    int num_kb;
    // output geometry (output pixels)
    int N, D, H, W;
    int bn, bd, bh, bw;                       // tile box, product == 128
    int tiles_n, tiles_d, tiles_h, tiles_w;   // tile counts per dim
    int num_m_tiles, num_n_tiles;
    int stride;                               // input coord = out*stride + tap - (k/2)
    int Cout;                                 // real output channels (row pitch of out)
    int b_rows_per_mtile;                     // batched GEMM: extra B row offset per m tile (0 = shared weights)
    // epilogue
    int mode;
    const float* bias;        // [Cout] or null
    const float* chan_add;    // per-image channel offsets (timestep embedding) or null: chan_add[n*chan_add_stride + c]
    long long chan_add_stride;  // row pitch in floats (0 = one shared row)
    int chan_mod;             // 0, or: chan_add holds chan_mod values per row, output column c reads entry c % chan_mod (the
                              // dense 2x2x2 form of a 3-D conv: columns are (voxel, channel))
    const __half* residual;   // same layout as out, or null
    __half* out;
    float scale;              // EPI_SOFTMAX_BD: logits scale
    int group;                // EPI_SOFTMAX_BD: tokens per image (block size)
    int vt_col0;              // EPI_STORE_VT: first transposed column
    __half* out_vt;           // EPI_STORE_VT: destination of the transposed columns
    // EPI_STORE: optional GroupNorm partial statistics of the (fp16-rounded) output, one (sum, sum of squares) pair per
    // (image, part, 4-channel quad); a part is the 32 output pixels one epilogue warp owns. Layout
    // [N][stats_parts][Cout/4][2] fp32. Summed in a fixed order: bitwise reproducible.
    float* stats_out;
    int stats_parts;
    // Nearest-x2-upsample + 3x3(x3) conv as 2^dims sub-pixel phases: phase (pd, ph, pw) is a 2x2(x2) conv over the
    // LOW-resolution input at tap offsets {p-1, p} per dim with pre-summed weights (rows [phase][Cout] of B), scattered
    // to output pixel 2*x + p. num_phases == 1: ordinary conv.
    int num_phases;
    int phase3d;
    int pair_rows;  // conv_halo pair tiles: accumulator row m = h * 16 + n' * 8 + w of images 2 * tile + n' (<= 8 x 8)
    int relu;       // EPI_STORE: max(., 0) after bias / addend / residual (VQ-VAE convs and residual units)
    // Split-precision activations (the VQ-VAE encoder, which the reference runs in fp32): a value is carried as fp16
    // hi + fp16 lo (lo = fp16(x - hi), ~22 mantissa bits together). out_lo / residual_lo: the lo halves of out / residual.
    __half* out_lo;
    const __half* residual_lo;
    int dbg;  // timing experiments only: 8 skip the statistics, 16 skip the output stores (results wrong when set)
};

// Host side: filled by conv_prepare(), launched by conv_launch().
struct ConvLaunch {
    ConvGemmParams p;
    int block_n;          // 128 or 256 (conv_halo, small batches: 64 too)
    int m_tiles_per_cta;  // 1, or 2 (block_n == 128 only): two 128-pixel tiles share each weight tile
    int cta_pair;         // 1: conv_gemm_2cta_kernel (clusters of 2, tcgen05 cta_group::2)
    int grid;
};

struct ConvSegment {
    const void* ptr;  // NDHWC fp16
    int channels;     // multiple of 64
    int ksize;        // 1, 3 or 5 (per spatial dim, "same" padding); 2 only with ConvProblem::upsample2; 4 with stride 2 and pad 1
};

struct ConvProblem {
    int spatial_dims;  // 2 or 3 (2 => D == 1)
    int N, D, H, W;    // INPUT spatial dims
    int stride;        // 1 or 2
    int n_seg;
    ConvSegment seg[kMaxSeg];
    const void* weights;  // fp16 [w_rows][Ktot]
    int w_rows;           // rows in the weight matrix (Cout, or batched)
    int Cout;
    int b_rows_per_mtile;
    int mode;
    const float* bias;
    const float* chan_add;
    long long chan_add_stride;
    int chan_mod;  // see ConvGemmParams::chan_mod (multiple of 32; needs impl-independent general epilogue: set with chan_add)
    const void* residual;
    void* out;
    float scale;
    int group;
    int vt_col0;
    void* out_vt;
    float* stats_out;  // or null; must hold N * conv_stats_parts(...) * Cout/4 * 2 floats
    // 1: the op is conv3x3(nearest_upsample_x2(input)); N,D,H,W are the LOW-resolution input extents, seg[0].ksize
    // is 2, weights are phase-packed [2^dims * Cout][2^dims * C] (pack_upconv_weight), out has the doubled extents and
    // statistics come in 2^dims * conv_stats_parts(low-res extents) parts.
    int upsample2;
    int impl;  // 0: pick; 1: single-CTA kernel only (A/B tests of the CTA-pair kernel)
    // halo kernel only: the 3x3 segments are channel slices of ONE conv weight [Cout][3x3 taps][C_total] (K ordered tap,
    // then channel over the concatenation) instead of one K block per segment - a conv over torch.cat(inputs, dim=1)
    int concat3x3;
    int gn_silu;  // halo kernel with a scale/shift table: 1 = GroupNorm + SiLU, 0 = GroupNorm only
    int relu;     // store epilogue: ReLU after bias / residual
    void* out_lo;             // split-precision output: lo half (see ConvGemmParams::out_lo)
    const void* residual_lo;  // split-precision residual: lo half
    int pad;      // zero padding of segments with ksize > 1; 0 means the default ksize / 2 ("same" for odd kernels)
};

// Number of GroupNorm-statistics parts per image the epilogue emits for an OUTPUT of this geometry (0: the tile box
// holds fewer than 32 pixels of an image, fused statistics unsupported).
int conv_stats_parts(int spatial_dims, int Dout, int Hout, int Wout);

// fp32 [Cout][Cin][3^dims] -> fp16 [2^dims phases][Cout][2^dims taps][Cin] (api.cu)
int pack_upconv_weight(const float* w, int Cout, int Cin, int dims, __half* dst, cudaStream_t stream);

// returns 0 on success; on failure sets the thread-local error string (see ddpm_last_error()).
int conv_prepare(const ConvProblem& prob, int num_sms, ConvLaunch* out);
int conv_launch(const ConvLaunch& l, cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency); null + error if absent.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode();

void set_error(const char* fmt, ...);
const char* last_error();

}  // namespace ddpm
