// tcgen05 attention core. See attention.cuh.
#include "attention.cuh"

#include <string.h>

#include "conv_gemm.cuh"  // set_error, get_encode
#include "launch.cuh"
#include "ptx.cuh"

namespace ddpm {

namespace {
constexpr int kHd = 256;                 // head dim
constexpr int kPanel = 128 * 64 * 2;     // one TMA box: 128 rows x 64 fp16 = 16 KB, 128B-swizzled
constexpr int kTile = 4 * kPanel;        // 128 rows x 256 fp16 = 64 KB
constexpr int kSmem = 3 * kTile + 1024 + 128;
constexpr int kThreads = 160;            // warps 0-3: softmax + epilogue (thread = query row); warp 4: TMA + MMA issue

// MN-major operand tile as TMA leaves it: rows = K index (tokens), 128 B = 64 N-values per row, 8-row swizzle atoms
// (SBO = 1024 B), next 64 N-values one panel further (LBO = 16 KB).
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(kPanel >> 4) << 16;  // leading byte offset
    d |= static_cast<uint64_t>(1024 >> 4) << 32;    // stride byte offset
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
}  // namespace

template <int NKB>  // key blocks of 128: 1 (T <= 128) or 2 (T == 256)
__global__ void __launch_bounds__(kThreads, 1) attention_tc_kernel(const __grid_constant__ AttnTcLaunch p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_s = smem;               // Q tile, later P (NKB * 32 KB)
    uint8_t* kv_s = smem + kTile;      // two key/value block buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTile);
    uint64_t* bar_q = bars;            // TMA -> MMA
    uint64_t* bar_k = bars + 1;        // [2]
    uint64_t* bar_v = bars + 3;        // [2]
    uint64_t* bar_s = bars + 5;        // S complete (tcgen05.commit)
    uint64_t* bar_p = bars + 6;        // P written (128 arrivals)
    uint64_t* bar_o = bars + 7;        // O complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ptx::pdl_trigger();
    const int head = blockIdx.y;
    const int tok0 = blockIdx.x * 128;                               // first query row of this CTA
    const int key0 = NKB == 1 ? tok0 : (tok0 / (NKB * 128)) * (NKB * 128);  // first key row
    const int colq = head * kHd, colk = p.C + head * kHd, colv = 2 * p.C + head * kHd;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&p.tm_qkv);
        ptx::mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bar_k[i], 1); ptx::mbar_init(&bar_v[i], 1); }
        ptx::mbar_init(bar_s, 1);
        ptx::mbar_init(bar_p, 128);
        ptx::mbar_init(bar_o, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc<512>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();
    const uint32_t tmem_s = tmem_base;        // S: NKB * 128 fp32 columns
    const uint32_t tmem_o = tmem_base + 256;  // O: 256 fp32 columns

    if (warp == 4) {
        if (lane == 0) {
            // ---- loads: Q and the key blocks
            ptx::mbar_arrive_expect_tx(bar_q, kTile);
            for (int pn = 0; pn < 4; ++pn) ptx::tma_load_2d(q_s + pn * kPanel, &p.tm_qkv, bar_q, colq + pn * 64, tok0);
            for (int kb = 0; kb < NKB; ++kb) {
                ptx::mbar_arrive_expect_tx(&bar_k[kb], kTile);
                for (int pn = 0; pn < 4; ++pn)
                    ptx::tma_load_2d(kv_s + kb * kTile + pn * kPanel, &p.tm_qkv, &bar_k[kb], colk + pn * 64,
                                     key0 + kb * 128);
            }
            // ---- S = Q K^T   (128 x 128 per key block, K = 256 in 16 steps)
            constexpr uint32_t idesc_s = ptx::make_idesc_f16(128, 128);
            ptx::mbar_wait(bar_q, 0);
            for (int kb = 0; kb < NKB; ++kb) {
                ptx::mbar_wait(&bar_k[kb], 0);
                ptx::tc_fence_after();
                for (int pn = 0; pn < 4; ++pn) {
                    const uint64_t da = ptx::make_desc_k128(ptx::smem_u32(q_s + pn * kPanel));
                    const uint64_t db = ptx::make_desc_k128(ptx::smem_u32(kv_s + kb * kTile + pn * kPanel));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(tmem_s + kb * 128, da + 2 * k, db + 2 * k, idesc_s, (pn | k) != 0);
                }
            }
            ptx::umma_commit(bar_s);
            // ---- the key buffers are free once S is complete: bring in the value blocks
            ptx::mbar_wait(bar_s, 0);
            for (int kb = 0; kb < NKB; ++kb) {
                ptx::mbar_arrive_expect_tx(&bar_v[kb], kTile);
                for (int pn = 0; pn < 4; ++pn)
                    ptx::tma_load_2d(kv_s + kb * kTile + pn * kPanel, &p.tm_qkv, &bar_v[kb], colv + pn * 64,
                                     key0 + kb * 128);
            }
            // ---- O = P V   (128 x 256, K = NKB * 128 keys in steps of 16; V is the MN-major operand)
            constexpr uint32_t idesc_o = ptx::make_idesc_f16(128, 256) | (1u << 16);
            ptx::mbar_wait(bar_p, 0);
            ptx::tc_fence_after();
            for (int kb = 0; kb < NKB; ++kb) {
                ptx::mbar_wait(&bar_v[kb], 0);
                ptx::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {  // 16 keys per step
                    // P: K-major, 64 keys per 16 KB panel, 32 B per step inside the swizzled row
                    const uint64_t da =
                        ptx::make_desc_k128(ptx::smem_u32(q_s + (kb * 2 + (ks >> 2)) * kPanel)) + 2 * (ks & 3);
                    // V: 16 token rows of 128 B per step
                    const uint64_t db = make_desc_mn128(ptx::smem_u32(kv_s + kb * kTile + ks * 16 * 128));
                    ptx::umma_f16(tmem_o, da, db, idesc_o, (kb | ks) != 0);
                }
            }
            ptx::umma_commit(bar_o);
        }
    } else {
        // ================================================================= softmax + epilogue, thread = query row
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        // keys this row may attend to, as columns of S
        int c_lo = 0, c_hi = NKB * 128;
        if (NKB == 1 && p.T < 128) { c_lo = (row / p.T) * p.T; c_hi = c_lo + p.T; }
        ptx::mbar_wait(bar_s, 0);
        ptx::tc_fence_after();
        float mx = -INFINITY;
        for (int c = 0; c < NKB * 4; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                if (col >= c_lo && col < c_hi) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
        }
        const float m2 = mx * p.scale_log2e;
        float sum = 0.f;
        for (int c = 0; c < NKB * 4; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
            ptx::tmem_ld_wait();
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const int col = c * 32 + j;
                float a = 0.f, b = 0.f;
                if (col >= c_lo && col < c_hi) a = exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -m2));
                if (col + 1 >= c_lo && col + 1 < c_hi) b = exp2f(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -m2));
                const __half2 hh = __floats2half2_rn(a, b);
                const float2 back = __half22float2(hh);  // normalise by the sum of what the tensor core will see
                sum += back.x + back.y;
                packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            // P[row][keys c*32 .. +31] -> K-major SWIZZLE_128B: panel = 64 keys, 16-byte chunk index XOR (row & 7)
            uint8_t* panel = q_s + (c >> 1) * kPanel + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = (c & 1) * 4 + j;
                *reinterpret_cast<uint4*>(panel + ((chunk ^ (row & 7)) << 4)) =
                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
        ptx::mbar_arrive(bar_p);
        const float inv = 1.0f / sum;
        ptx::mbar_wait(bar_o, 0);
        ptx::tc_fence_after();
        const int tok = tok0 + row;
        __half* dst = p.out + static_cast<size_t>(tok) * p.C + head * kHd;
        for (int c = 0; c < kHd / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_o + lane_addr + c * 32, v);
            ptx::tmem_ld_wait();
            if (tok < p.rows) {
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const __half2 hh = __floats2half2_rn(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
                    packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

bool attention_tc_supported(int T, int C, int heads) {
    if (heads <= 0 || C != heads * kHd) return false;
    return (T >= 1 && T <= 128 && 128 % T == 0) || T == 256;
}

int attention_tc_prepare(const __half* qkv, __half* out, int N, int T, int C, int heads, float scale, AttnTcLaunch* l) {
    if (!attention_tc_supported(T, C, heads)) { set_error("attention_tc: T=%d C=%d heads=%d unsupported", T, C, heads); return 2; }
    PFN_encodeTiled encode = get_encode();
    if (!encode) return 1;
    memset(l, 0, sizeof(*l));
    const long long rows = static_cast<long long>(N) * T;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(3 * C), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(3 * C) * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&l->tm_qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(qkv), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("attention_tc: cuTensorMapEncodeTiled failed: %d", (int)r); return 3; }
    l->out = out;
    l->rows = static_cast<int>(rows);
    l->T = T; l->C = C; l->heads = heads;
    l->scale_log2e = scale * 1.4426950408889634f;
    l->grid_x = static_cast<int>((rows + 127) / 128);
    return 0;
}

int attention_tc_launch(const AttnTcLaunch& l, cudaStream_t stream) {
    static bool attr_set_dev[kMaxDevices] = {};
    bool& attr_set = attr_set_dev[device_slot()];
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        cudaError_t e2 = cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attention_tc: cudaFuncSetAttribute failed"); return 4; }
        attr_set = true;
    }
    dim3 grid(l.grid_x, l.heads);
    cudaError_t e;
    if (l.T == 256)
        e = launch_pdl(attention_tc_kernel<2>, grid, dim3(kThreads), kSmem, stream, l);
    else
        e = launch_pdl(attention_tc_kernel<1>, grid, dim3(kThreads), kSmem, stream, l);
    if (e != cudaSuccess) { set_error("attention_tc: launch failed: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}

}  // namespace ddpm
