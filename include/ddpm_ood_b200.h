/*
 * ddpm_ood_b200 — C ABI of the B200-native reconstruction hot path of marksgraham/ddpm-ood.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t (as void*), launches
 * asynchronously on that stream, never synchronises the host, never frees caller memory, and returns 0 on success or a
 * non-zero code whose text is available from ddpm_last_error() (no exceptions cross this boundary).
 * One host thread per process, one process per GPU (the reference's torchrun model, src/trainers/base.py:22-33).
 *
 * The reference has no FFI of its own: the seam is a set of Python call sites. Each group below names
 * the call site it replaces.
 */
#ifndef DDPM_OOD_B200_H
#define DDPM_OOD_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DDPM_API __attribute__((visibility("default")))
#else
#define DDPM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ general */
DDPM_API const char* ddpm_last_error(void);
/* ABI version, bumped on any signature change. */
DDPM_API int ddpm_abi_version(void);

/* ------------------------------------------------------------------------------------------------ building block
 * Implicit-GEMM convolution / linear layer on tcgen05 tensor cores. This is the operator behind every 3x3(x3) conv,
 * 1x1 skip conv and attention Linear inside DiffusionModelUNet.forward, which the reference calls at
 * src/trainers/reconstruct.py:150-153 (model built at src/trainers/base.py:66-75).
 * Activations are channels-last fp16 (N, D, H, W, C); weights are fp16 [w_rows][Ktot] packed by
 * ddpm_pack_conv_weight(). */
typedef struct ddpm_conv_args {
    int spatial_dims;        /* 2 or 3; for 2, D must be 1 */
    int N, D, H, W;          /* input extents */
    int stride;              /* 1 or 2 (padding is always k/2) */
    int n_seg;               /* 1..3 K-segments accumulated into one output */
    const void* seg_ptr[3];  /* fp16 NDHWC inputs */
    int seg_channels[3];     /* multiples of 64 */
    int seg_ksize[3];        /* 1 or 3 */
    const void* weights;     /* fp16 [w_rows][Ktot], Ktot = sum_s taps_s * channels_s, ordered segment, tap, channel */
    int w_rows;
    int Cout;                /* multiple of 128 */
    int b_rows_per_mtile;    /* batched GEMM: weight-row offset added per 128-row M tile; 0 for a conv */
    int mode;                /* 0 store, 1 block-diagonal softmax, 2 store with transposed V columns */
    const float* bias;       /* [Cout] or NULL */
    const float* chan_add;   /* [N][Cout] or NULL: per-image channel offset (timestep embedding projection) */
    const void* residual;    /* fp16, same layout as out, or NULL */
    void* out;               /* fp16 (N, Do, Ho, Wo, Cout) */
    float scale;             /* mode 1 */
    int group;               /* mode 1: tokens per image */
    int vt_col0;             /* mode 2 */
    void* out_vt;            /* mode 2 */
} ddpm_conv_args;
DDPM_API int ddpm_conv_forward(const ddpm_conv_args* args, void* stream);

/* fp32 PyTorch conv weight [Cout][Cin][taps] (or Linear weight with taps == 1) -> fp16 rows of a packed matrix:
 * dst[co * ktot + koff + tap * Cin + ci]. */
DDPM_API int ddpm_pack_conv_weight(const float* w, int Cout, int Cin, int taps, void* dst, long long ktot, long long koff,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDPM_OOD_B200_H */
