// Hardware probe (not part of the library): issue rate of back-to-back tcgen05.mma kind::f16 with both operands in
// shared memory, for the shapes the conv kernels use. One CTA (or CTA pair) per SM, every SM busy; a single thread
// issues `iters` x 4 MMAs (K = 16 each, walking a 64-wide k-block) and commits; cycles from clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 umma_rate_probe.cu -o umma_rate_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../ptx.cuh"

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// mode bits: 1 = shifted halo view for A (start + 11 rows, SBO 1280), 2 = two accumulators (MT = 2) sharing B
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) rank = ptx::cluster_ctarank();
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 2) ptx::tmem_alloc_2cta<512>(slot); else ptx::tmem_alloc<512>(slot);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0 && rank == 0) {
        // warp-uniform loop, one elected lane issues: descriptors stay in uniform registers (no R2UR waterfall)
        constexpr uint32_t idesc = ptx::make_idesc_f16(CG * 128, N);
        const uint32_t a0 = ptx::smem_u32(smem) + ((mode & 1) ? 11 * 128 : 0);
        const uint32_t sbo = (mode & 1) ? 1280 : 1024;
        const uint32_t b0 = ptx::smem_u32(smem + 64 * 1024);
        const int mt_n = (mode & 16) ? 3 : ((mode & 4) ? 4 : ((mode & 2) ? 2 : 1));
        const uint32_t acc_stride = (mode & 8) ? 256 : N;
        const uint64_t da_base = make_desc(a0, sbo);
        const uint64_t db_base = make_desc(b0, 1024);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint64_t da = da_base + ((it & 1) ? (23 * 1024 >> 4) : 0);
            const uint64_t db = db_base + (it & 3) * (16384 >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll 4
                    for (int mt = 0; mt < mt_n; ++mt) {
                        if (CG == 2) ptx::umma_f16_2cta(tmem + mt * acc_stride, da + 2 * k, db + 2 * k, idesc, 1);
                        else ptx::umma_f16(tmem + mt * acc_stride, da + 2 * k, db + 2 * k, idesc, 1);
                    }
                }
            }
            __syncwarp();
        }
        if (ptx::elect_one()) {
            if (CG == 2) ptx::umma_commit_2cta(bar); else ptx::umma_commit(bar);
        }
        __syncwarp();
        ptx::mbar_wait(bar, 0);
        t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    }
    if (CG == 2 && rank == 1 && warp == 0 && lane == 0) ptx::mbar_wait(bar, 0);
    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync_all();
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CG == 2) ptx::tmem_dealloc_2cta<512>(tmem); else ptx::tmem_dealloc<512>(tmem);
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int CG, int N>
void run(const char* name, int mode, long long* dout) {
    const int smem = 160 * 1024 + 1024 + 64;
    CK(cudaFuncSetAttribute(rate_kernel<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, rate_kernel<CG, N>, iters, mode, dout));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc; CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
        const int mt_n = (mode & 16) ? 3 : ((mode & 4) ? 4 : ((mode & 2) ? 2 : 1));
        const double mmas = 4.0 * iters * mt_n;
        const double macs_per_cta = 128.0 * N * 16;  // per CTA per MMA
        if (rep == 1)
            printf("%-34s mode %d: %7.1f cycles/MMA, %6.0f MACs/clk/SM, chip %.0f TFLOP/s (%.3f ms)\n", name, mode,
                   cyc / mmas, macs_per_cta * mmas / cyc, 2.0 * macs_per_cta * mmas * 148 / (ms * 1e-3) / 1e12, ms);
    }
}

int main() {
    long long* dout;
    CK(cudaMalloc(&dout, 8));
    const int modes[5] = {0, 2, 10, 16, 4};  // 1 acc; 2 accs; 2 accs 256 columns apart; 3 accs; 4 accs
    for (int mi = 0; mi < 5; ++mi) {
        const int mode = modes[mi];
        run<1, 128>("cta_group::1 M=128 N=128", mode, dout);
        run<2, 128>("cta_group::2 M=256 N=128", mode, dout);
        if (mode != 4 && mode != 16) run<2, 256>("cta_group::2 M=256 N=256", mode, dout);
    }
    return 0;
}
