// One kernel per AttentionBlock of DiffusionModelUNet (monai-generative; built at src/trainers/base.py:66-86, run inside
// every UNet forward of src/trainers/reconstruct.py:150-153):
//
//   out = h + proj( softmax( q k^T / sqrt(d) ) v ),   q | k | v = Linear(GroupNorm(h)),   one head of d = C = 256.
//
// A CTA owns 128 token rows = 128 / Tpad whole images (Tpad = tokens per image rounded up to a power of two; rows beyond
// the real T tokens are zero and masked, so 7 x 7 = 49-token maps of native 28 x 28 FashionMNIST run here too) and keeps
// everything between the read of h and the write of out on chip: GroupNorm statistics are taken from the staged tile
// itself (the tile holds whole images) and applied in place in shared memory, q / k / v / P / O never exist in HBM.
// Round 1 ran four launches per block (gn_apply, q|k|v GEMM, attention core, projection GEMM) that moved h, the
// normalised h, q|k|v and O through L2 / HBM: 9.9 % of a forward for 1 % of its FLOPs (profiles/r01_launches_s7.md).
//
// What bounds it: the 512 KB of projection weights each CTA streams from L2 (6-stage ring of 128 x 64 fp16 panels),
// not the tensor pipe: per 128 rows ~10 k cycles of tcgen05.mma against ~16 k cycles of weight feed at ~32 B/clk/SM.
#pragma once
#include "conv_gemm.cuh"

namespace ddpm {

struct AttnBlockParams {
    CUtensorMap tm_x;      // h as [N][T][C] fp16: dims (C, T, N), box (64, Tpad, 128 / Tpad), SWIZZLE_128B
    CUtensorMap tm_wqkv;   // [3C][C] fp16 (q | k | v rows), box (64, 128)
    CUtensorMap tm_wproj;  // [C][C] fp16, box (64, 128)
    ConvGemmParams epi;    // output geometry / residual / statistics for conv_epilogue_tile16 (tensor maps unused)
    const float* gamma; const float* beta;  // GroupNorm affine [C]
    const float* bqkv;                       // [3C]
    const float* bproj;                      // [C]
    int N, T, Tpad, ipt;   // images, tokens per image, padded tokens, images per 128-row tile
    int num_tiles;
    float eps;
    float scale_log2e;     // softmax scale * log2(e)
};

struct AttnBlockLaunch {
    AttnBlockParams p;
    int grid;
};

// C == 256, one head, T <= 128 tokens
bool attn_block_supported(int T, int C, int heads, int groups);
// GroupNorm-statistics parts per image the kernel's epilogue emits for its OUTPUT (0: fewer than 32 rows per image)
int attn_block_stats_parts(int T);
// h, out: [N*T][C] fp16; wqkv [3C][C], wproj [C][C] fp16; stats_out: null or [N][parts][C/4][2] fp32
int attn_block_prepare(const __half* h, __half* out, int N, int T, int C, int heads, int groups, float eps, float scale,
                       const float* gamma, const float* beta, const __half* wqkv, const float* bqkv, const __half* wproj,
                       const float* bproj, float* stats_out, int num_sms, AttnBlockLaunch* l);
int attn_block_launch(const AttnBlockLaunch& l, cudaStream_t stream);

}  // namespace ddpm
