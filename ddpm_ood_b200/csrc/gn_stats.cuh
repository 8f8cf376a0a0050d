// GroupNorm statistics -> per-channel (scale, shift), shared by gn_apply / gn_finalize (kernels.cu) and the halo-tile
// conv's transform warps (conv_halo.cu): ONE implementation, so every consumer derives bit-identical values from the
// producers' partial sums (fixed summation order, no atomics).
#pragma once
#include <cuda_runtime.h>

namespace ddpm {

// 256 cooperating threads (tid 0..255; `sync` is their barrier). st0/st1: partial (sum, sum of squares) per (image,
// part, 4-channel quad) of the (up to two, channel-concatenated) source tensors, [N][parts][C/4][2] fp32. Writes
// out(c, scale, shift) for the C0 + C1 channels of image n. Scratch: s_qs / s_qq [(C0 + C1) / 4], s_sub [256].
template <typename Sync, typename Out>
__device__ __forceinline__ void gn_scale_shift_from_parts(int tid, int n, int C0, const float* __restrict__ st0, int parts0,
                                                          int C1, const float* __restrict__ st1, int parts1,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int S, int cpg, float eps,
                                                          float* s_qs, float* s_qq, float2* s_sub, Sync sync, Out out) {
    const int C = C0 + C1;
    const int Q = C >> 2, Q0 = C0 >> 2, Q1 = C1 >> 2;
    {
        // fixed-order two-level sum of the partials: J threads per quad take parts j, j+J, ... (independent loads),
        // then one thread per quad adds the J sub-sums.
        const int J = Q <= 256 ? 256 / Q : 1;
        for (int base = 0; base < Q; base += 256) {
            const int qd = base + tid % (Q < 256 ? Q : 256);
            const int j = tid / (Q < 256 ? Q : 256);
            float a = 0.f, b = 0.f;
            if (qd < Q && j < J) {
                const float2* p;
                int parts, Qs;
                if (qd < Q0) { p = reinterpret_cast<const float2*>(st0) + static_cast<size_t>(n) * parts0 * Q0 + qd; parts = parts0; Qs = Q0; }
                else { p = reinterpret_cast<const float2*>(st1) + static_cast<size_t>(n) * parts1 * Q1 + (qd - Q0); parts = parts1; Qs = Q1; }
#pragma unroll 4
                for (int i = j; i < parts; i += J) { const float2 v = __ldg(p + static_cast<size_t>(i) * Qs); a += v.x; b += v.y; }
            }
            s_sub[tid] = make_float2(a, b);
            sync();
            if (tid < 256 && base + tid < Q && tid < (Q < 256 ? Q : 256)) {
                float sa = 0.f, sb = 0.f;
                for (int jj = 0; jj < J; ++jj) { const float2 v = s_sub[jj * (Q < 256 ? Q : 256) + tid]; sa += v.x; sb += v.y; }
                s_qs[base + tid] = sa;
                s_qq[base + tid] = sb;
            }
            sync();
        }
    }
    sync();
    const float inv_n = 1.0f / (static_cast<float>(cpg) * static_cast<float>(S));
    for (int c = tid; c < C; c += 256) {
        const int q0 = (c / cpg) * (cpg >> 2);
        float sum = 0.f, sq = 0.f;
        for (int i = 0; i < (cpg >> 2); ++i) { sum += s_qs[q0 + i]; sq += s_qq[q0 + i]; }
        const float mean = __fmul_rn(sum, inv_n);
        float var = __fsub_rn(__fmul_rn(sq, inv_n), __fmul_rn(mean, mean));  // explicit: no FMA contraction differences
        var = var < 0.f ? 0.f : var;
        const float rstd = rsqrtf(var + eps);
        const float a = __fmul_rn(gamma[c], rstd);
        out(c, a, __fsub_rn(beta[c], __fmul_rn(mean, a)));
    }
    sync();
}

}  // namespace ddpm
