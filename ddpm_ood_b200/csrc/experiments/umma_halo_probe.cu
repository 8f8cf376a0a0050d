// Hardware probe (not part of the library): can tcgen05.mma read the 9 taps of a 3x3 conv as SHIFTED views of ONE
// haloed K-major SWIZZLE_128B tile in shared memory?  A tile of 8 (w) x 16 (h) output pixels needs the 10 x 18 input
// pixels around it; pixel (h', w') sits in smem row h' * P + w' (128 B per row = 64 fp16 channels). For tap (dh, dw)
// the A operand is rows (h + 1 + dh) * P + (w + 1 + dw): 16 groups of 8 consecutive rows, group stride P rows, start
// NOT 1024-byte aligned. Modes: P = 10 (one dense TMA box, SBO = 1280 B) or P = 16 (18 one-row TMA boxes, SBO = 2048 B),
// each with descriptor base_offset = 0 or (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. umma_halo_probe.cu -o umma_halo_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../ptx.cuh"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Params {
    CUtensorMap tmA_box;  // box (64, 10, 18)
    CUtensorMap tmA_row;  // box (64, 10, 1)
    CUtensorMap tmB;      // [9*128][64], box (64, 128)
    float* d;             // [128][128]
    int h0, w0;
    int mode;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;
    d |= static_cast<uint64_t>(base_off & 7) << 49;
    d |= 2ull << 61;
    return d;
}

__global__ void __launch_bounds__(160, 1) probe_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = smem;                    // up to 18 * 2048 = 36 KB
    uint8_t* b_s = smem + 40 * 1024;        // 9 * 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 40 * 1024 + 9 * 16384);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool padded = p.mode >= 2;
    const bool use_bo = p.mode & 1;
    if (warp == 4 && lane == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc<128>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 4 && lane == 0) {
        const uint32_t a_bytes = 180 * 128;
        ptx::mbar_arrive_expect_tx(&bars[0], a_bytes + 9 * 16384);
        if (!padded) {
            ptx::tma_load_5d(a_s, &p.tmA_box, &bars[0], 0, p.w0 - 1, p.h0 - 1, 0, 0);
        } else {
            for (int r = 0; r < 18; ++r)
                ptx::tma_load_5d(a_s + r * 2048, &p.tmA_row, &bars[0], 0, p.w0 - 1, p.h0 - 1 + r, 0, 0);
        }
        for (int t = 0; t < 9; ++t) ptx::tma_load_2d(b_s + t * 16384, &p.tmB, &bars[0], 0, t * 128);
        ptx::mbar_wait(&bars[0], 0);
        ptx::tc_fence_after();
        constexpr uint32_t idesc = ptx::make_idesc_f16(128, 128);
        const uint32_t P = padded ? 16 : 10;
        for (int t = 0; t < 9; ++t) {
            const int dh = t / 3 - 1, dw = t % 3 - 1;
            const uint32_t a_addr = ptx::smem_u32(a_s) + ((1 + dh) * P + (1 + dw)) * 128;
            const uint32_t bo = use_bo ? ((a_addr >> 7) & 7) : 0;
            const uint64_t db = ptx::make_desc_k128(ptx::smem_u32(b_s + t * 16384));
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_desc(a_addr + k * 32, P * 128, bo);
                ptx::umma_f16(tmem, da, db + 2 * k, idesc, (t | k) != 0);
            }
        }
        ptx::umma_commit(&bars[1]);
    }
    if (warp < 4) {
        ptx::mbar_wait(&bars[1], 0);
        ptx::tc_fence_after();
        const int row = warp * 32 + lane;
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
            ptx::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) p.d[row * 128 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<128>(tmem);
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main() {
    const int H = 24, W = 16, C = 64;
    std::vector<__half> hx(H * W * C), hw(9 * 128 * C);
    srand(1);
    for (auto& v : hx) v = __float2half(static_cast<float>(rand() % 5 - 2));
    for (auto& v : hw) v = __float2half(static_cast<float>(rand() % 3 - 1));
    __half *dx, *dw;
    float* dd;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&dd, 128 * 128 * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(fn);
    Params p{};
    {
        cuuint64_t gdim[5] = {C, W, H, 1, 1};
        cuuint64_t gstr[4] = {C * 2, C * 2 * W, C * 2 * W * H, C * 2 * W * H};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        cuuint32_t box[5] = {64, 10, 18, 1, 1};
        CUresult r = encode(&p.tmA_box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dx, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t box2[5] = {64, 10, 1, 1, 1};
        CUresult r2 = encode(&p.tmA_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dx, gdim, gstr, box2, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t bd[2] = {C, 9 * 128};
        cuuint64_t bs[1] = {C * 2};
        cuuint32_t bb[2] = {64, 128};
        cuuint32_t be[2] = {1, 1};
        CUresult r3 = encode(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, bd, bs, bb, be, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r || r2 || r3) { printf("encode failed %d %d %d\n", (int)r, (int)r2, (int)r3); return 1; }
    }
    p.d = dd;
    const int smem = 40 * 1024 + 9 * 16384 + 1024 + 128;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int origins[2][2] = {{3, 4}, {0, 0}};
    std::vector<float> hd(128 * 128);
    for (int mode = 0; mode < 4; ++mode) {
        for (int o = 0; o < 2; ++o) {
            p.mode = mode;
            p.h0 = origins[o][0];
            p.w0 = origins[o][1];
            CK(cudaMemset(dd, 0xff, 128 * 128 * 4));
            probe_kernel<<<1, 160, smem>>>(p);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            double maxerr = 0;
            for (int m = 0; m < 128; ++m) {
                const int h = m / 8, w = m % 8;
                for (int n = 0; n < 128; ++n) {
                    float ref = 0.f;
                    for (int t = 0; t < 9; ++t) {
                        const int hh = p.h0 + h + t / 3 - 1, ww = p.w0 + w + t % 3 - 1;
                        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                        for (int c = 0; c < C; ++c)
                            ref += __half2float(hx[(hh * W + ww) * C + c]) * __half2float(hw[(t * 128 + n) * C + c]);
                    }
                    const double e = fabs(ref - hd[m * 128 + n]);
                    if (e > maxerr) maxerr = e;
                    if (e > 1e-3) ++bad;
                }
            }
            printf("mode %d (%s pitch, base_offset %s) origin (%d,%d): %d / 16384 wrong, max err %g\n", mode,
                   mode >= 2 ? "16-row" : "10-row", (mode & 1) ? "set" : "0", p.h0, p.w0, bad, maxerr);
        }
    }
    return 0;
}
