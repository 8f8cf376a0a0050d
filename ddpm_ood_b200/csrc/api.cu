// extern "C" surface declared in include/ddpm_ood_b200.h.
#include "../../include/ddpm_ood_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_gemm.cuh"

namespace ddpm {

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            sms = 0;
            return 148;
        }
    }
    return sms;
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                        __half* __restrict__ dst, long long ktot, long long koff) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // destination-major so writes coalesce: i = (co * taps + tap) * Cin + ci
        const int ci = static_cast<int>(i % Cin);
        const long long r = i / Cin;
        const int tap = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        const float v = w[(static_cast<long long>(co) * Cin + ci) * taps + tap];
        dst[co * ktot + koff + static_cast<long long>(tap) * Cin + ci] = __float2half_rn(v);
    }
}

}  // namespace ddpm

extern "C" {

const char* ddpm_last_error(void) { return ddpm::last_error(); }
int ddpm_abi_version(void) { return 1; }

int ddpm_conv_forward(const ddpm_conv_args* a, void* stream) {
    if (!a) { ddpm::set_error("ddpm_conv_forward: null args"); return 2; }
    ddpm::ConvProblem q{};
    q.spatial_dims = a->spatial_dims;
    q.N = a->N; q.D = a->D; q.H = a->H; q.W = a->W;
    q.stride = a->stride;
    q.n_seg = a->n_seg;
    for (int s = 0; s < a->n_seg && s < ddpm::kMaxSeg; ++s) {
        q.seg[s].ptr = a->seg_ptr[s];
        q.seg[s].channels = a->seg_channels[s];
        q.seg[s].ksize = a->seg_ksize[s];
    }
    q.weights = a->weights;
    q.w_rows = a->w_rows;
    q.Cout = a->Cout;
    q.b_rows_per_mtile = a->b_rows_per_mtile;
    q.mode = a->mode;
    q.bias = a->bias;
    q.chan_add = a->chan_add;
    q.residual = a->residual;
    q.out = a->out;
    q.scale = a->scale;
    q.group = a->group;
    q.vt_col0 = a->vt_col0;
    q.out_vt = a->out_vt;
    ddpm::ConvLaunch l;
    int rc = ddpm::conv_prepare(q, ddpm::num_sms(), &l);
    if (rc) return rc;
    return ddpm::conv_launch(l, static_cast<cudaStream_t>(stream));
}

int ddpm_pack_conv_weight(const float* w, int Cout, int Cin, int taps, void* dst, long long ktot, long long koff,
                          void* stream) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    if (total <= 0) return 0;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ddpm::pack_conv_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, Cout, Cin, taps, static_cast<__half*>(dst), ktot, koff);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("pack_conv_weight: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // extern "C"
