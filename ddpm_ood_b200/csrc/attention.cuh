// Fused attention core on tcgen05 tensor cores: softmax(Q K^T / sqrt(d)) V per (image, head), head dim 256, for the
// token counts the reference's UNets produce at their attention levels (AttentionBlock of monai-generative's
// DiffusionModelUNet, built at src/trainers/base.py:66-86): T tokens with 128 % T == 0 (e.g. 64 = 8x8) or T == 256.
// One CTA = 128 query rows: for T <= 128 that is 128/T whole images with a block-diagonal mask, for T == 256 half an
// image against its 256 keys. S and O accumulate in TMEM, P goes back through shared memory as the A operand of the
// second GEMM, V is consumed MN-major straight from the TMA tile (no transpose pass).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ddpm {

struct AttnTcLaunch {
    CUtensorMap tm_qkv;  // 2-D map over qkv [rows][3C] fp16, box 64 x 128, SWIZZLE_128B
    __half* out;         // [rows][C]
    int rows;            // N * T
    int T, C, heads;
    float scale_log2e;   // softmax scale * log2(e)
    int grid_x;
};

bool attention_tc_supported(int T, int C, int heads);
// qkv: [N*T, 3C] fp16 (q | k | v column blocks), out: [N*T, C] fp16.
int attention_tc_prepare(const __half* qkv, __half* out, int N, int T, int C, int heads, float scale, AttnTcLaunch* l);
int attention_tc_launch(const AttnTcLaunch& l, cudaStream_t stream);

}  // namespace ddpm
