// Simplex noise for the reconstruction loop's --simplex_noise=1 mode on the device.
//
// Reference: src/utils/simplex_noise.py - generate_simplex_noise (:15-79) draws a seed per (channel, image), builds the
// permutation tables on the host (_init, :559-577) and calls a numba kernel per (channel, image) with a D2H / H2D round
// trip each (rand_3d_fixed_T_octaves, :141-159 -> _noise3, :704-1271: 3-D OpenSimplex on the plane z = t / frequency,
// 6 octaves). Here: one launch builds every table from its seed, one launch evaluates every pixel of every image.
//
// The arithmetic is fp64 with explicit round-to-nearest intrinsics in the reference's operand order (no FMA
// contraction), so the values are bit-identical to the reference's; the candidate-point selection follows
// oracle/simplex.py's restatement (selection rule + one uniform contribution loop instead of the spelled-out cases).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ddpm_ood_b200.h"
#include "conv_gemm.cuh"  // set_error

namespace ddpm {
namespace {

// _init: LCG-driven draw without replacement from 0..255; thread = one table
__global__ void simplex_tables_kernel(const long long* __restrict__ seeds, int n, uint8_t* __restrict__ perm) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const unsigned long long MUL = 6364136223846793005ull, INC = 1442695040888963407ull;
    unsigned long long s = static_cast<unsigned long long>(seeds[m]);  // wrapping unsigned arithmetic == c_int64 overflow
    uint8_t source[256];
    for (int i = 0; i < 256; ++i) source[i] = static_cast<uint8_t>(i);
    for (int i = 0; i < 3; ++i) s = s * MUL + INC;
    uint8_t* out = perm + static_cast<size_t>(m) * 256;
    for (int i = 255; i >= 0; --i) {
        s = s * MUL + INC;
        const long long mod = i + 1;
        long long r = static_cast<long long>(s) % mod;  // floor-mod of (seed + 31) without overflowing seed + 31
        if (r < 0) r += mod;
        r = (r + 31) % mod;
        out[i] = source[r];
        source[r] = source[i];
    }
}

struct Cand { int i, j, k; bool squish_first; };

__device__ __forceinline__ int first_set(int x, int y) { return x ? 0 : (y ? 1 : 2); }
__device__ __forceinline__ int first_clear(int x, int y) { return !x ? 0 : (!y ? 1 : 2); }

// oracle/simplex.py candidates(): up to eight lattice points of the cell, in the reference's summation order
__device__ int select_candidates(double xins, double yins, double zins, Cand (&c)[8]) {
    const double s[3] = {xins, yins, zins};
    const double in_sum = __dadd_rn(__dadd_rn(xins, yins), zins);
    auto put = [&](int idx, int i, int j, int k, bool f = false) { c[idx] = Cand{i, j, k, f}; };
    if (in_sum <= 1.0) {
        int a = 0, b = 1;
        if (s[a] >= s[b] && zins > s[b]) b = 2;
        else if (s[a] < s[b] && zins > s[a]) a = 2;
        const double w = __dsub_rn(1.0, in_sum);
        int e0[3], e1[3];
        if (w > s[a] || w > s[b]) {
            const int cc = s[b] > s[a] ? b : a;
            const int lo = cc == 0 ? 1 : 0, hi = cc == 2 ? 1 : 2;
            for (int ax = 0; ax < 3; ++ax) { e0[ax] = ax == cc; e1[ax] = ax == cc; }
            e0[lo] -= 1;
            e1[hi] -= 1;
        } else {
            const int k = 3 - a - b;
            for (int ax = 0; ax < 3; ++ax) { e0[ax] = ax != k; e1[ax] = ax != k ? 1 : -1; }
        }
        put(0, 0, 0, 0); put(1, 1, 0, 0); put(2, 0, 1, 0); put(3, 0, 0, 1);
        put(4, e0[0], e0[1], e0[2]); put(5, e1[0], e1[1], e1[2]);
        return 6;
    }
    if (in_sum >= 2.0) {
        int a = 0, b = 1;
        if (s[a] <= s[b] && zins < s[b]) b = 2;
        else if (s[a] > s[b] && zins < s[a]) a = 2;
        const double w = __dsub_rn(3.0, in_sum);
        int e0[3], e1[3];
        if (w < s[a] || w < s[b]) {
            const int cc = s[b] < s[a] ? b : a;
            const int lo = cc == 0 ? 1 : 0, hi = cc == 2 ? 1 : 2;
            for (int ax = 0; ax < 3; ++ax) { e0[ax] = ax != cc; e1[ax] = ax != cc; }
            e0[lo] += 1;
            e1[hi] += 1;
        } else {
            const int k = 3 - a - b;
            for (int ax = 0; ax < 3; ++ax) { e0[ax] = ax == k; e1[ax] = ax == k ? 2 : 0; }
        }
        put(0, 1, 1, 0); put(1, 1, 0, 1); put(2, 0, 1, 1); put(3, 1, 1, 1);
        put(4, e0[0], e0[1], e0[2]); put(5, e1[0], e1[1], e1[2]);
        return 6;
    }
    // octahedron: per pair of opposite vertices the nearer one, then the best two
    struct Pick { double score; int x, y, z; bool far; };
    auto pick = [](double p, int fx, int fy, int fz, int nx, int ny, int nz) {
        return p > 1.0 ? Pick{__dsub_rn(p, 1.0), fx, fy, fz, true} : Pick{__dsub_rn(1.0, p), nx, ny, nz, false};
    };
    Pick A = pick(__dadd_rn(xins, yins), 1, 1, 0, 0, 0, 1);
    Pick B = pick(__dadd_rn(xins, zins), 1, 0, 1, 0, 1, 0);
    const Pick C = pick(__dadd_rn(yins, zins), 0, 1, 1, 1, 0, 0);
    if (A.score <= B.score && A.score < C.score) { A.x = C.x; A.y = C.y; A.z = C.z; A.far = C.far; }
    else if (A.score > B.score && B.score < C.score) { B.x = C.x; B.y = C.y; B.z = C.z; B.far = C.far; }
    int e0[3], e1[3];
    bool flag = false;
    if (A.far == B.far) {
        if (A.far) {
            const int m = first_set(A.x & B.x, A.y & B.y);
            for (int ax = 0; ax < 3; ++ax) { e0[ax] = 1; e1[ax] = ax == m ? 2 : 0; }
        } else {
            const int k = first_clear(A.x | B.x, A.y | B.y);
            for (int ax = 0; ax < 3; ++ax) { e0[ax] = 0; e1[ax] = ax == k ? -1 : 1; }
        }
    } else {
        const Pick& f = A.far ? A : B;
        const Pick& nr = A.far ? B : A;
        const int k = first_clear(f.x, f.y);
        const int m = first_set(nr.x, nr.y);
        for (int ax = 0; ax < 3; ++ax) { e0[ax] = ax == k ? -1 : 1; e1[ax] = ax == m ? 2 : 0; }
        flag = true;
    }
    put(0, 1, 0, 0); put(1, 0, 1, 0); put(2, 0, 0, 1); put(3, 1, 1, 0); put(4, 1, 0, 1); put(5, 0, 1, 1);
    put(6, e0[0], e0[1], e0[2]); put(7, e1[0], e1[1], e1[2], flag);
    return 8;
}

__device__ double noise3(double x, double y, double z, const uint8_t* __restrict__ perm) {
    const double kStretch = -1.0 / 6, kSquish = 1.0 / 3;
    const double stretch = __dmul_rn(__dadd_rn(__dadd_rn(x, y), z), kStretch);
    const double xs = __dadd_rn(x, stretch), ys = __dadd_rn(y, stretch), zs = __dadd_rn(z, stretch);
    const double fx = floor(xs), fy = floor(ys), fz = floor(zs);
    const int xsb = static_cast<int>(fx), ysb = static_cast<int>(fy), zsb = static_cast<int>(fz);
    const double squish = __dmul_rn(static_cast<double>(xsb + ysb + zsb), kSquish);
    const double dx0 = __dsub_rn(x, __dadd_rn(fx, squish));
    const double dy0 = __dsub_rn(y, __dadd_rn(fy, squish));
    const double dz0 = __dsub_rn(z, __dadd_rn(fz, squish));
    Cand c[8];
    const int n = select_candidates(__dsub_rn(xs, fx), __dsub_rn(ys, fy), __dsub_rn(zs, fz), c);
    double value = 0.0;
    for (int q = 0; q < n; ++q) {
        const double sq = __dmul_rn(static_cast<double>(c[q].i + c[q].j + c[q].k), kSquish);
        const double di = c[q].i, dj = c[q].j, dk = c[q].k;
        double dx, dy, dz;
        if (c[q].squish_first) {
            dx = __dsub_rn(__dsub_rn(dx0, sq), di); dy = __dsub_rn(__dsub_rn(dy0, sq), dj); dz = __dsub_rn(__dsub_rn(dz0, sq), dk);
        } else {
            dx = __dsub_rn(__dsub_rn(dx0, di), sq); dy = __dsub_rn(__dsub_rn(dy0, dj), sq); dz = __dsub_rn(__dsub_rn(dz0, dk), sq);
        }
        double attn = __dsub_rn(__dsub_rn(__dsub_rn(2.0, __dmul_rn(dx, dx)), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (attn > 0.0) {
            const int g = perm[(perm[(perm[(xsb + c[q].i) & 0xFF] + ysb + c[q].j) & 0xFF] + zsb + c[q].k) & 0xFF] % 24;
            const int blk = g / 3, axis = g - 3 * blk;
            const double g0 = ((blk & 1) ? 1.0 : -1.0) * (axis == 0 ? 11.0 : 4.0);
            const double g1 = ((blk & 2) ? -1.0 : 1.0) * (axis == 1 ? 11.0 : 4.0);
            const double g2 = ((blk & 4) ? -1.0 : 1.0) * (axis == 2 ? 11.0 : 4.0);
            attn = __dmul_rn(attn, attn);
            const double ext = __dadd_rn(__dadd_rn(__dmul_rn(g0, dx), __dmul_rn(g1, dy)), __dmul_rn(g2, dz));
            value = __dadd_rn(value, __dmul_rn(__dmul_rn(attn, attn), ext));
        }
    }
    return __ddiv_rn(value, 103.0);
}

// rand_3d_fixed_T_octaves for every image: block = 256 pixels of image m (its table in shared memory)
__global__ void __launch_bounds__(256) simplex_noise_kernel(const uint8_t* __restrict__ perm_all,
                                                            const long long* __restrict__ t, float* __restrict__ out,
                                                            int B, int C, int H, int W, int octaves, double persistence,
                                                            double frequency) {
    __shared__ uint8_t perm[256];
    const int m = blockIdx.y;  // = channel * B + image: the order the seeds were drawn in
    perm[threadIdx.x] = perm_all[static_cast<size_t>(m) * 256 + threadIdx.x];
    __syncthreads();
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= H * W) return;
    const int ch = m / B, img = m - ch * B;
    const double xx = pix % W, yy = pix / W, tt = static_cast<double>(t[img]);
    double acc = 0.0, amp = 1.0, f = frequency;
    for (int o = 0; o < octaves; ++o) {
        const double v = noise3(__ddiv_rn(xx, f), __ddiv_rn(yy, f), __ddiv_rn(tt, f), perm);
        acc = __dadd_rn(acc, __dmul_rn(amp, v));
        f = __ddiv_rn(f, 2.0);
        amp = __dmul_rn(amp, persistence);
    }
    out[(static_cast<size_t>(img) * C + ch) * H * W + pix] = static_cast<float>(acc);
}

}  // namespace
}  // namespace ddpm

extern "C" int ddpm_simplex_noise(const long long* seeds, const long long* t, float* out, unsigned char* tables_ws, int B,
                                  int C, int H, int W, int octaves, double persistence, double frequency, void* stream) {
    if (!seeds || !t || !out || !tables_ws || B < 1 || C < 1 || H < 1 || W < 1 || octaves < 0) {
        ddpm::set_error("ddpm_simplex_noise: bad argument");
        return 2;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = B * C;
    ddpm::simplex_tables_kernel<<<(n + 63) / 64, 64, 0, s>>>(seeds, n, tables_ws);
    dim3 grid((H * W + 255) / 256, n);
    ddpm::simplex_noise_kernel<<<grid, 256, 0, s>>>(tables_ws, t, out, B, C, H, W, octaves, persistence, frequency);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ddpm::set_error("ddpm_simplex_noise: %s", cudaGetErrorString(e)); return 5; }
    return 0;
}
