// Stem convolution nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3) of
// _Hourglass_Wrapper (reference models/asn_stacked_hg.py:223,283): direct convolution that reads
// the NCHW image the reference API is fed with and writes the NHWC pre-BN tensor + the batch
// statistics of bn1 (:224).  Cin = 3 makes this an FFMA kernel (K = 147), 2 % of the step FLOPs.
#include <stdlib.h>
#include "common.cuh"

namespace hgk {

constexpr int ST_TH = 8, ST_TW = 16;                  // output tile (pixels)
constexpr int ST_PH = 2 * ST_TH + 5, ST_PW = 2 * ST_TW + 5;   // input patch 21 x 37
constexpr int ST_K = 147;                              // 3*7*7
constexpr int ST_CO = 64;

__device__ __forceinline__ void stem_load_patch(float* patch, const float* img, int n, int H, int W, int oy0,
                                                int ox0, int tid) {
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int i = tid; i < 3 * ST_PH * ST_PW; i += 256) {
        int c = i / (ST_PH * ST_PW);
        int r = i - c * (ST_PH * ST_PW);
        int py = r / ST_PW, px = r - py * ST_PW;
        int iy = iy0 + py, ix = ix0 + px;
        float v = 0.f;
        if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
            v = __ldg(img + ((size_t)(n * 3 + c) * H + iy) * W + ix);
        patch[i] = v;
    }
}

// Persistent CTAs (3 per SM) walk over the output tiles; 128 threads: tx = tid&7 -> 8 couts, ty = tid>>3 -> 8 pixels of one tile
// row (64 accumulators per thread).  The [k][co] weight tile is staged ONCE per CTA (a per-tile transposing store with a
// 64-float stride was a 32-way bank conflict: ~9 400 LSU cycles per tile, 40 % of the first kernel), and a thread reads the 21
// consecutive patch floats its eight stride-2 windows cover as six LDS.128 per (channel, kh) row.  An LDS.128 costs four LSU
// cycles per warp: with 8 x 4 outputs per thread (13 loads = 52 LSU cycles per 224 FFMA = 56 issue cycles) the kernel was
// LSU- and FFMA-bound at the same time (ncu: FMA pipe 51 %); 8 x 8 needs 80 LSU cycles per 448 FFMA (112 issue cycles).
// Patch rows are padded to 40 floats (alignment).  (Measured and rejected: rows of 44 floats + couts interleaved as {4 tx ..} and
// {32 + 4 tx ..}, which removes the two-way bank conflicts ncu reports, made the kernel slower: 240 us against 210.)
constexpr int ST_PWF = 40;
constexpr int ST_FWD_THREADS = 128;

__global__ void __launch_bounds__(ST_FWD_THREADS, 3) stem_conv7_fwd_kernel(const float* __restrict__ img, int N, int H, int W,
                                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                                           float* __restrict__ y, double* stat_sum, double* stat_sq,
                                                                           int tiles_x, int tiles_y) {
    __shared__ __align__(16) float Ws[ST_K * ST_CO];          // [k][co]
    __shared__ __align__(16) float patch[3 * ST_PH * ST_PWF];
    const int tid = threadIdx.x;
    const int OH = H / 2, OW = W / 2;
    for (int i = tid; i < ST_K * ST_CO; i += ST_FWD_THREADS) {   // consecutive threads -> consecutive co: conflict-free stores
        const int k = i >> 6, co = i & 63;
        Ws[i] = __ldg(w + co * ST_K + k);
    }
    const int tx = tid & 7, ty = tid >> 3;
    const int oy = ty >> 1, oxb = (ty & 1) * 8;
    float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f), bv1 = bv0;
    if (bias) { bv0 = ldg4(bias + tx * 8); bv1 = ldg4(bias + tx * 8 + 4); }
    const float bvv[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
    double d1[8], d2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { d1[j] = 0.0; d2[j] = 0.0; }
    const int total = tiles_x * tiles_y * N;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = t / (tiles_x * tiles_y);
        const int r = t - n * (tiles_x * tiles_y);
        const int tyi = r / tiles_x, txi = r - tyi * tiles_x;
        const int oy0 = tyi * ST_TH, ox0 = txi * ST_TW;
        const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
        __syncthreads();                                      // previous tile's patch fully consumed (first trip: Ws staged)
        for (int i = tid; i < 3 * ST_PH * ST_PWF; i += ST_FWD_THREADS) {
            const int cc = i / (ST_PH * ST_PWF);
            const int rr = i - cc * (ST_PH * ST_PWF);
            const int py = rr / ST_PWF, px = rr - py * ST_PWF;
            const int iy = iy0 + py, ix = ix0 + px;
            float v = 0.f;
            if (px < ST_PW && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                v = __ldg(img + ((size_t)(n * 3 + cc) * H + iy) * W + ix);
            patch[i] = v;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 1
        for (int ck = 0; ck < 21; ++ck) {                     // (channel, kh) patch rows
            const int c = ck / 7, kh = ck - c * 7;
            const float* prow = patch + (c * ST_PH + oy * 2 + kh) * ST_PWF + oxb * 2;
            float pv[24];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const float4 v = ld4(prow + 4 * q);
                pv[4 * q] = v.x; pv[4 * q + 1] = v.y; pv[4 * q + 2] = v.z; pv[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
                const float4 b0 = ld4(Ws + (ck * 7 + kw) * ST_CO + tx * 8), b1 = ld4(Ws + (ck * 7 + kw) * ST_CO + tx * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = pv[i * 2 + kw];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av, bb[j], acc[i][j]);
                }
            }
        }
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        const int gy = oy0 + oy;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gx = ox0 + oxb + i;
            if (gy < OH && gx < OW) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = acc[i][j] + bvv[j];
                    s1[j] += v[j];
                    s2[j] = fmaf(v[j], v[j], s2[j]);
                }
                float* yo = y + (((size_t)n * OH + gy) * OW + gx) * ST_CO + tx * 8;
                st4(yo, make_float4(v[0], v[1], v[2], v[3]));
                st4(yo + 4, make_float4(v[4], v[5], v[6], v[7]));
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { d1[j] += (double)s1[j]; d2[j] += (double)s2[j]; }   // fp32 over 8 pixels, fp64 from here on
    }
    if (stat_sum != nullptr) {
        __syncthreads();
        double* red = reinterpret_cast<double*>(Ws);          // [4 warps][64][2]
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double e1 = d1[j], e2 = d2[j];                    // lanes = 4 ty x 8 tx: sum over ty
            e1 += __shfl_xor_sync(0xffffffffu, e1, 8);  e2 += __shfl_xor_sync(0xffffffffu, e2, 8);
            e1 += __shfl_xor_sync(0xffffffffu, e1, 16); e2 += __shfl_xor_sync(0xffffffffu, e2, 16);
            if (lane < 8) {
                red[(warp * 64 + lane * 8 + j) * 2 + 0] = e1;
                red[(warp * 64 + lane * 8 + j) * 2 + 1] = e2;
            }
        }
        __syncthreads();
        if (tid < 64) {
            double e1 = 0.0, e2 = 0.0;
#pragma unroll
            for (int wv = 0; wv < ST_FWD_THREADS / 32; ++wv) {
                e1 += red[(wv * 64 + tid) * 2 + 0];
                e2 += red[(wv * 64 + tid) * 2 + 1];
            }
            atomicAdd(stat_sum + tid, e1);
            atomicAdd(stat_sq + tid, e2);
        }
    }
}

// dW[co][k] += sum_p dz[p][co] * patch(p)[k];  persistent CTAs loop over output tiles and flush once.
// thread -> 4 couts (tx) x 10 taps (ty*10 .. +9 of the 147, padded to 160)
__global__ void __launch_bounds__(256) stem_conv7_wgrad_kernel(const float* __restrict__ img, int N, int H, int W,
                                                               const float* __restrict__ dz, float* dw, float* dbias,
                                                               int tiles_x, int tiles_y) {
    __shared__ __align__(16) float dzs[ST_TH * ST_TW * ST_CO];     // [pixel][co]
    __shared__ float patch[3 * ST_PH * ST_PW];
    const int tid = threadIdx.x;
    const int OH = H / 2, OW = W / 2;
    const int tx = tid & 15, ty = tid >> 4;
    int koff[10];
    bool kval[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        int k = ty * 10 + j;
        kval[j] = k < ST_K;
        int kk = kval[j] ? k : 0;
        int c = kk / 49, r = kk - c * 49;
        int kh = r / 7, kw = r - kh * 7;
        koff[j] = (c * ST_PH + kh) * ST_PW + kw;
    }
    float acc[4][10];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const int total = tiles_x * tiles_y * N;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int n = t / (tiles_x * tiles_y);
        int r = t - n * (tiles_x * tiles_y);
        int tyi = r / tiles_x, txi = r - tyi * tiles_x;
        int oy0 = tyi * ST_TH, ox0 = txi * ST_TW;
        __syncthreads();
        stem_load_patch(patch, img, n, H, W, oy0, ox0, tid);
        for (int i = tid; i < ST_TH * ST_TW * (ST_CO / 4); i += 256) {
            int q = i >> 4, cv = i & 15;
            int gy = oy0 + q / ST_TW, gx = ox0 + (q % ST_TW);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy < OH && gx < OW) v = ldg4(dz + (((size_t)n * OH + gy) * OW + gx) * ST_CO + cv * 4);
            st4(dzs + q * ST_CO + cv * 4, v);
        }
        __syncthreads();
        for (int q = 0; q < ST_TH * ST_TW; ++q) {
            float4 a = ld4(dzs + q * ST_CO + tx * 4);
            const float* pb = patch + (q / ST_TW) * 2 * ST_PW + (q % ST_TW) * 2;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                float b = pb[koff[j]];
                acc[0][j] = fmaf(a.x, b, acc[0][j]);
                acc[1][j] = fmaf(a.y, b, acc[1][j]);
                acc[2][j] = fmaf(a.z, b, acc[2][j]);
                acc[3][j] = fmaf(a.w, b, acc[3][j]);
            }
        }
        if (dbias != nullptr && tid < ST_CO) {
            for (int q = 0; q < ST_TH * ST_TW; ++q) bsum += dzs[q * ST_CO + tid];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j)
            if (kval[j]) atomicAdd(dw + (size_t)(tx * 4 + i) * ST_K + ty * 10 + j, acc[i][j]);
    if (dbias != nullptr && tid < ST_CO) atomicAdd(dbias + tid, bsum);
}

// Second-generation weight gradient.  The first kernel above issues 11 shared-memory loads per 40 FFMA (one scalar
// patch load per tap): LSU-bound at ~14 TFLOP/s, and it is the LAST kernel of the backward pass, alone on the GPU for
// 0.45-0.53 ms.  Here a thread owns 4 couts x the 7 kw taps of ONE (channel, kh) patch row (21 rows -> 21 groups of 16
// threads) and walks the tile in PAIRS of horizontally adjacent output pixels: their two 7-tap windows are the 9
// consecutive floats prow[4p .. 4p+8] (stride 2), i.e. two LDS.128 + one LDS.32 feed 56 FFMA (with the two dz float4
// loads: 11 FFMA per load).  Patch rows are padded to 40 floats so that every window start is 16-byte aligned.
constexpr int ST_PWP = 40;
constexpr int ST_WG2_THREADS = 352;       // 21 x 16 workers + 16 idle lanes of the 11th warp

__global__ void __launch_bounds__(ST_WG2_THREADS, 2) stem_conv7_wgrad2_kernel(const float* __restrict__ img, int N, int H, int W,
                                                                             const float* __restrict__ dz, float* dw, float* dbias,
                                                                             int tiles_x, int tiles_y) {
    __shared__ __align__(16) float dzs[ST_TH * ST_TW * ST_CO];     // [pixel][co]
    __shared__ __align__(16) float patch[3 * ST_PH * ST_PWP];      // [c][py][px], rows padded to 40
    const int tid = threadIdx.x;
    const int OH = H / 2, OW = W / 2;
    const int tx = tid & 15, ty = tid >> 4;
    const bool worker = ty < 21;
    const int c = worker ? ty / 7 : 0, kh = worker ? ty - (ty / 7) * 7 : 0;
    float acc[4][7];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const int total = tiles_x * tiles_y * N;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = t / (tiles_x * tiles_y);
        const int r = t - n * (tiles_x * tiles_y);
        const int tyi = r / tiles_x, txi = r - tyi * tiles_x;
        const int oy0 = tyi * ST_TH, ox0 = txi * ST_TW;
        const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
        __syncthreads();                                     // previous tile fully consumed
        for (int i = tid; i < 3 * ST_PH * ST_PWP; i += ST_WG2_THREADS) {
            const int cc = i / (ST_PH * ST_PWP);
            const int rr = i - cc * (ST_PH * ST_PWP);
            const int py = rr / ST_PWP, px = rr - py * ST_PWP;
            const int iy = iy0 + py, ix = ix0 + px;
            float v = 0.f;
            if (px < ST_PW && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                v = __ldg(img + ((size_t)(n * 3 + cc) * H + iy) * W + ix);
            patch[i] = v;
        }
        for (int i = tid; i < ST_TH * ST_TW * (ST_CO / 4); i += ST_WG2_THREADS) {
            const int q = i >> 4, cv = i & 15;
            const int gy = oy0 + q / ST_TW, gx = ox0 + (q % ST_TW);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy < OH && gx < OW) v = ldg4(dz + (((size_t)n * OH + gy) * OW + gx) * ST_CO + cv * 4);
            st4(dzs + q * ST_CO + cv * 4, v);
        }
        __syncthreads();
        if (worker) {
#pragma unroll 1
            for (int oy = 0; oy < ST_TH; ++oy) {
                const float* prow = patch + (c * ST_PH + 2 * oy + kh) * ST_PWP;
                const float* arow = dzs + (oy * ST_TW) * ST_CO + tx * 4;
#pragma unroll 2
                for (int p = 0; p < ST_TW / 2; ++p) {
                    const float4 a0 = ld4(arow + (2 * p) * ST_CO);
                    const float4 a1 = ld4(arow + (2 * p + 1) * ST_CO);
                    const float4 b0 = ld4(prow + 4 * p), b1 = ld4(prow + 4 * p + 4);
                    const float b8 = prow[4 * p + 8];
                    const float b[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b8};
#pragma unroll
                    for (int kw = 0; kw < 7; ++kw) {
                        acc[0][kw] = fmaf(a0.x, b[kw], acc[0][kw]);
                        acc[1][kw] = fmaf(a0.y, b[kw], acc[1][kw]);
                        acc[2][kw] = fmaf(a0.z, b[kw], acc[2][kw]);
                        acc[3][kw] = fmaf(a0.w, b[kw], acc[3][kw]);
                        acc[0][kw] = fmaf(a1.x, b[kw + 2], acc[0][kw]);
                        acc[1][kw] = fmaf(a1.y, b[kw + 2], acc[1][kw]);
                        acc[2][kw] = fmaf(a1.z, b[kw + 2], acc[2][kw]);
                        acc[3][kw] = fmaf(a1.w, b[kw + 2], acc[3][kw]);
                    }
                }
            }
        }
        if (dbias != nullptr && tid < ST_CO) {
            for (int q = 0; q < ST_TH * ST_TW; ++q) bsum += dzs[q * ST_CO + tid];
        }
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kw = 0; kw < 7; ++kw)
                atomicAdd(dw + (size_t)(tx * 4 + i) * ST_K + (c * 7 + kh) * 7 + kw, acc[i][kw]);
    }
    if (dbias != nullptr && tid < ST_CO) atomicAdd(dbias + tid, bsum);
}

// Third-generation weight gradient: the stem is the LAST layer of the backward pass, so this kernel (and the BatchNorm-backward
// apply in front of it) ran alone on the GPU for 0.36 ms.  Two changes:
//   * the BatchNorm-backward apply is evaluated here: the kernel reads g = dL/d relu(bn1(z)) and z, forms
//     dz = cA*((g*[z*scale+shift > 0] - cC) - (z - mean)*cB) in shared memory (bn_bwd_apply_kernel's expression) -- the
//     stand-alone pass over the 100 MB tensor (read g, read z, write dz) and the re-read of dz disappear;
//   * tiles of 4 x 16 output pixels, double-buffered with cp.async (the next tile's g, z and image patch stream in while
//     the current one is in the FFMA loop; the old kernel alternated a load phase and a compute phase per CTA).
// Thread -> EIGHT couts x the 7 kw taps of one (channel, kh) patch row (56 accumulators): an LDS.128 costs four LSU cycles per
// warp, so the second generation's 4 x 7 tile (four LDS.128 + one LDS.32 = 17 LSU cycles per 56 FFMA = 14 issue cycles) was
// LSU-bound; 8 x 7 needs 25 LSU cycles per 112 FFMA (28 issue cycles).
constexpr int ST3_TH = 4, ST3_TW = 16;
constexpr int ST3_PH = 2 * ST3_TH + 5;                     // 13 input rows
constexpr int ST3_PX = ST3_TH * ST3_TW;                    // 64 output pixels per tile
constexpr int ST3_THREADS = 192;                           // 21 x 8 workers + 24 idle lanes
constexpr int ST3_BUF_FLOATS = 2 * ST3_PX * ST_CO + 3 * ST3_PH * ST_PWP;       // g | z | patch
constexpr int ST3_SMEM = 2 * ST3_BUF_FLOATS * 4 + 6 * ST_CO * 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;                         // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(ST3_THREADS, 2) stem_conv7_wgrad3_kernel(const float* __restrict__ img, int N, int H, int W,
                                                                         const float* __restrict__ g, const float* __restrict__ z,
                                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                                         int relu, const float* __restrict__ mean,
                                                                         const float* __restrict__ cA, const float* __restrict__ cB,
                                                                         const float* __restrict__ cC, float* dw, float* dbias,
                                                                         int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float st3_smem[];
    float* vecs = st3_smem + 2 * ST3_BUF_FLOATS;           // scale | shift | mean | cA | cB | cC, 64 each
    const int tid = threadIdx.x;
    const int OH = H / 2, OW = W / 2;
    const int tx = tid & 7, ty = tid >> 3;                 // tx: cout octet, ty: (channel, kh) patch row
    const bool worker = ty < 21;
    const int c = worker ? ty / 7 : 0, kh = worker ? ty - (ty / 7) * 7 : 0;
    if (tid < ST_CO) {
        vecs[tid] = __ldg(scale + tid); vecs[64 + tid] = __ldg(shift + tid); vecs[128 + tid] = __ldg(mean + tid);
        vecs[192 + tid] = __ldg(cA + tid); vecs[256 + tid] = __ldg(cB + tid); vecs[320 + tid] = __ldg(cC + tid);
    }
    float acc[8][7];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);          // bias gradient of channel quad (tid & 15), this thread's share
    const int total = tiles_x * tiles_y * N;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(st3_smem);

    auto issue = [&](int t, int b) {                        // asynchronous copies of tile t into buffer b
        const int n = t / (tiles_x * tiles_y);
        const int r = t - n * (tiles_x * tiles_y);
        const int tyi = r / tiles_x, txi = r - tyi * tiles_x;
        const int oy0 = tyi * ST3_TH, ox0 = txi * ST3_TW;
        const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
        const uint32_t gb = sbase + (uint32_t)(b * ST3_BUF_FLOATS) * 4u;
        for (int i = tid; i < ST3_PX * (ST_CO / 4); i += ST3_THREADS) {
            const int q = i >> 4, cv = i & 15;
            const int gy = oy0 + q / ST3_TW, gx = ox0 + (q % ST3_TW);
            const bool ok = gy < OH && gx < OW;
            const size_t off = ok ? ((((size_t)n * OH + gy) * OW + gx) * ST_CO + cv * 4) : 0;
            cp_async16(gb + (uint32_t)(q * ST_CO + cv * 4) * 4u, g + off, ok);
            cp_async16(gb + (uint32_t)(ST3_PX * ST_CO + q * ST_CO + cv * 4) * 4u, z + off, ok);
        }
        const uint32_t pb = gb + (uint32_t)(2 * ST3_PX * ST_CO) * 4u;
        for (int i = tid; i < 3 * ST3_PH * ST_PWP; i += ST3_THREADS) {
            const int cc = i / (ST3_PH * ST_PWP);
            const int rr = i - cc * (ST3_PH * ST_PWP);
            const int py = rr / ST_PWP, px = rr - py * ST_PWP;
            const int iy = iy0 + py, ix = ix0 + px;
            const bool ok = px < ST_PW && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
            const size_t off = ok ? (((size_t)(n * 3 + cc) * H + iy) * W + ix) : 0;
            cp_async4(pb + (uint32_t)i * 4u, img + off, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int t = blockIdx.x, b = 0;
    if (t < total) issue(t, 0);
    for (; t < total; t += gridDim.x, b ^= 1) {
        const int tn = t + gridDim.x;
        if (tn < total) {
            issue(tn, b ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                    // tile t has landed (and the vectors are staged)
        float* gs = st3_smem + b * ST3_BUF_FLOATS;
        const float* zs = gs + ST3_PX * ST_CO;
        const float* patch = gs + 2 * ST3_PX * ST_CO;
        {   // BatchNorm-backward apply in place: g -> dz (this thread always meets channel quad tid & 15: 192 % 16 == 0)
            const int cv = tid & 15;
            const float4 s4 = ld4(vecs + cv * 4), t4 = ld4(vecs + 64 + cv * 4), mu = ld4(vecs + 128 + cv * 4);
            const float4 a4 = ld4(vecs + 192 + cv * 4), b4 = ld4(vecs + 256 + cv * 4), c4 = ld4(vecs + 320 + cv * 4);
            const int n = t / (tiles_x * tiles_y);
            const int r = t - n * (tiles_x * tiles_y);
            const int oy0 = (r / tiles_x) * ST3_TH, ox0 = (r % tiles_x) * ST3_TW;
            for (int i = tid; i < ST3_PX * (ST_CO / 4); i += ST3_THREADS) {
                const int q = i >> 4;
                const bool ok = oy0 + q / ST3_TW < OH && ox0 + (q % ST3_TW) < OW;
                const float4 gv = ld4(gs + q * ST_CO + cv * 4), zv = ld4(zs + q * ST_CO + cv * 4);
                const float gx = (relu && fmaf(zv.x, s4.x, t4.x) <= 0.f) ? 0.f : gv.x;
                const float gy = (relu && fmaf(zv.y, s4.y, t4.y) <= 0.f) ? 0.f : gv.y;
                const float gz = (relu && fmaf(zv.z, s4.z, t4.z) <= 0.f) ? 0.f : gv.z;
                const float gw = (relu && fmaf(zv.w, s4.w, t4.w) <= 0.f) ? 0.f : gv.w;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {                                   // pixels outside the image contribute nothing (their dz is not defined)
                    o.x = a4.x * ((gx - c4.x) - (zv.x - mu.x) * b4.x);
                    o.y = a4.y * ((gy - c4.y) - (zv.y - mu.y) * b4.y);
                    o.z = a4.z * ((gz - c4.z) - (zv.z - mu.z) * b4.z);
                    o.w = a4.w * ((gw - c4.w) - (zv.w - mu.w) * b4.w);
                }
                st4(gs + q * ST_CO + cv * 4, o);
                bsum.x += o.x; bsum.y += o.y; bsum.z += o.z; bsum.w += o.w;
            }
        }
        __syncthreads();
        if (worker) {
#pragma unroll 1
            for (int oy = 0; oy < ST3_TH; ++oy) {
                const float* prow = patch + (c * ST3_PH + 2 * oy + kh) * ST_PWP;
                const float* arow = gs + (oy * ST3_TW) * ST_CO + tx * 4;           // couts {4 tx ..} and {32 + 4 tx ..}
#pragma unroll 2
                for (int p = 0; p < ST3_TW / 2; ++p) {
                    const float4 a00 = ld4(arow + (2 * p) * ST_CO), a01 = ld4(arow + (2 * p) * ST_CO + 32);
                    const float4 a10 = ld4(arow + (2 * p + 1) * ST_CO), a11 = ld4(arow + (2 * p + 1) * ST_CO + 32);
                    const float4 b0 = ld4(prow + 4 * p), b1 = ld4(prow + 4 * p + 4);
                    const float b8 = prow[4 * p + 8];
                    const float bb[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b8};
                    const float a0[8] = {a00.x, a00.y, a00.z, a00.w, a01.x, a01.y, a01.z, a01.w};
                    const float a1[8] = {a10.x, a10.y, a10.z, a10.w, a11.x, a11.y, a11.z, a11.w};
#pragma unroll
                    for (int kw = 0; kw < 7; ++kw) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            acc[i][kw] = fmaf(a0[i], bb[kw], acc[i][kw]);
                            acc[i][kw] = fmaf(a1[i], bb[kw + 2], acc[i][kw]);
                        }
                    }
                }
            }
        }
        __syncthreads();                                    // buffer b is free for the tile after next
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int kw = 0; kw < 7; ++kw)
                atomicAdd(dw + (size_t)((i >> 2) * 32 + tx * 4 + (i & 3)) * ST_K + (c * 7 + kh) * 7 + kw, acc[i][kw]);
    }
    if (dbias != nullptr) {
        const int cv = tid & 15;
        atomicAdd(dbias + cv * 4 + 0, bsum.x); atomicAdd(dbias + cv * 4 + 1, bsum.y);
        atomicAdd(dbias + cv * 4 + 2, bsum.z); atomicAdd(dbias + cv * 4 + 3, bsum.w);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core stem (forward): the 7x7 stride-2 convolution in SPACE-TO-DEPTH form.  With x'[sy][sx][(dy*2+dx)*3 + c] =
// X[c][2 sy + dy][2 sx + dx] (half resolution, 12 real channels padded to 16) the stem is a stride-1 convolution with 4x4
// taps at offsets -2 .. +1:  out[oy][ox] = sum_{ty,tx,c'} W'[co][c'][ty][tx] x'[oy + ty - 2][ox + tx - 2][c'],
// W'[co][(dy*2+dx)*3 + c][ty][tx] = W[co][c][2 ty + dy - 1][2 tx + dx - 1] (zero where the 7x7 index falls outside) -- exactly
// the shape the image-tile tcgen05 kernel runs (conv_tc2.cu, KS = 4): halo tile staged once, sixteen taps as shifted
// descriptors, TF32 + 2xBF16 (or 3xTF32) products, bias + BatchNorm statistics + finaliser in its epilogue.  The zero padding
// of 3 pixels becomes the zero rows / columns of x' outside [0, H/2) x [0, W/2).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_s2d_image_kernel(const float* __restrict__ img, int H, int W, unsigned total,
                                                              float* __restrict__ xs) {
    const int H2 = H >> 1, W2 = W >> 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned q = i & 3u, pix = i >> 2;                // 16-byte channel quad of one half-resolution pixel
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < 3u) {
            const unsigned sx = pix % (unsigned)W2, r = pix / (unsigned)W2;
            const unsigned sy = r % (unsigned)H2, n = r / (unsigned)H2;
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned cp = q * 4u + j;                  // c' = (dy*2 + dx)*3 + c
                const unsigned c = cp % 3u, d = cp / 3u;
                e[j] = __ldg(img + (((size_t)n * 3 + c) * H + (2u * sy + (d >> 1))) * W + (2u * sx + (d & 1u)));
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        st4(xs + (size_t)i * 4, v);
    }
}

// W [Cout][3][7][7] -> W' [Cout][32][4][4] (OIHW, the layout the weight packer of the tensor-core kernels reads: it packs K in
// blocks of 32 channels, so the 16 image channels are followed by 16 zero ones that the kernel never streams)
__global__ void stem_s2d_weight_kernel(const float* __restrict__ w, int Cout, float* __restrict__ ws) {
    const int total = Cout * 32 * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tx = i & 3, ty = (i >> 2) & 3, cp = (i >> 4) & 31, co = i >> 9;
        float v = 0.f;
        if (cp < 12) {
            const int c = cp % 3, d = cp / 3;
            const int ky = 2 * ty + (d >> 1) - 1, kx = 2 * tx + (d & 1) - 1;
            if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = __ldg(w + ((co * 3 + c) * 7 + ky) * 7 + kx);
        }
        ws[i] = v;
    }
}

}  // namespace hgk

using namespace hgk;

extern "C" int hgk_stem_conv7_fwd(const float* img, int N, int H, int W, const float* w, const float* bias, int Cout,
                                  float* y, double* stat_sum, double* stat_sq, void* stream) {
    HGK_REQUIRE(img && w && y, "hgk_stem_conv7_fwd: null pointer");
    HGK_REQUIRE(Cout == ST_CO, "hgk_stem_conv7_fwd: Cout must be 64 (got %d)", Cout);
    HGK_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "hgk_stem_conv7_fwd: H, W must be even and positive");
    HGK_REQUIRE((stat_sum == nullptr) == (stat_sq == nullptr), "hgk_stem_conv7_fwd: stat_sum/stat_sq must both be set");
    HGK_REQUIRE(N <= 65535, "hgk_stem_conv7_fwd: batch too large");
    const int tiles_x = (W / 2 + ST_TW - 1) / ST_TW, tiles_y = (H / 2 + ST_TH - 1) / ST_TH;
    const long long total = (long long)tiles_x * tiles_y * N;
    HGK_REQUIRE(total < (1LL << 31), "hgk_stem_conv7_fwd: too many tiles");
    const int grid = (int)(total < 3 * kNumSMs ? total : 3 * kNumSMs);
    stem_conv7_fwd_kernel<<<grid, ST_FWD_THREADS, 0, (cudaStream_t)stream>>>(img, N, H, W, w, bias, y, stat_sum, stat_sq, tiles_x, tiles_y);
    HGK_CHECK_LAUNCH("hgk_stem_conv7_fwd");
    return HGK_OK;
}

extern "C" int hgk_stem_conv7_wgrad(const float* img, int N, int H, int W, const float* dz, int Cout,
                                    float* dw, float* dbias, void* stream) {
    HGK_REQUIRE(img && dz && dw, "hgk_stem_conv7_wgrad: null pointer");
    HGK_REQUIRE(Cout == ST_CO, "hgk_stem_conv7_wgrad: Cout must be 64 (got %d)", Cout);
    HGK_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "hgk_stem_conv7_wgrad: H, W must be even and positive");
    int tiles_x = (W / 2 + ST_TW - 1) / ST_TW, tiles_y = (H / 2 + ST_TH - 1) / ST_TH;
    long long total = (long long)tiles_x * tiles_y * N;
    int grid = (int)(total < 2 * kNumSMs ? total : 2 * kNumSMs);
    static int gen = -1;                  // HGK_STEM_WGRAD=1 selects the first-generation kernel (A/B measurements)
    if (gen < 0) {
        const char* e = getenv("HGK_STEM_WGRAD");
        gen = (e != nullptr && e[0] == '1') ? 1 : 2;
    }
    if (gen == 1)
        stem_conv7_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, N, H, W, dz, dw, dbias, tiles_x, tiles_y);
    else
        stem_conv7_wgrad2_kernel<<<grid, ST_WG2_THREADS, 0, (cudaStream_t)stream>>>(img, N, H, W, dz, dw, dbias, tiles_x, tiles_y);
    HGK_CHECK_LAUNCH("hgk_stem_conv7_wgrad");
    return HGK_OK;
}

extern "C" int hgk_stem_conv7_wgrad_bnapply(const float* img, int N, int H, int W, const float* g, const float* z,
                                            const float* scale, const float* shift, int relu, const float* mean,
                                            const float* cA, const float* cB, const float* cC, int Cout,
                                            float* dw, float* dbias, void* stream) {
    HGK_REQUIRE(img && g && z && scale && shift && mean && cA && cB && cC && dw, "hgk_stem_conv7_wgrad_bnapply: null pointer");
    HGK_REQUIRE(Cout == ST_CO, "hgk_stem_conv7_wgrad_bnapply: Cout must be 64 (got %d)", Cout);
    HGK_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "hgk_stem_conv7_wgrad_bnapply: H, W must be even and positive");
    HGK_REQUIRE((((uintptr_t)g | (uintptr_t)z) & 15) == 0, "hgk_stem_conv7_wgrad_bnapply: g and z must be 16-byte aligned");
    const int tiles_x = (W / 2 + ST3_TW - 1) / ST3_TW, tiles_y = (H / 2 + ST3_TH - 1) / ST3_TH;
    const long long total = (long long)tiles_x * tiles_y * N;
    HGK_REQUIRE(total < (1LL << 31), "hgk_stem_conv7_wgrad_bnapply: too many tiles");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(stem_conv7_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST3_SMEM);
        if (e != cudaSuccess) {
            set_error("hgk_stem_conv7_wgrad_bnapply: cudaFuncSetAttribute(%d bytes): %s", ST3_SMEM, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    const int grid = (int)(total < 2 * kNumSMs ? total : 2 * kNumSMs);
    stem_conv7_wgrad3_kernel<<<grid, ST3_THREADS, ST3_SMEM, (cudaStream_t)stream>>>(img, N, H, W, g, z, scale, shift, relu, mean,
                                                                                  cA, cB, cC, dw, dbias, tiles_x, tiles_y);
    HGK_CHECK_LAUNCH("hgk_stem_conv7_wgrad_bnapply");
    return HGK_OK;
}


extern "C" int hgk_stem_s2d_image(const float* img, int N, int H, int W, float* xs, void* stream) {
    HGK_REQUIRE(img && xs, "hgk_stem_s2d_image: null pointer");
    HGK_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "hgk_stem_s2d_image: H and W must be even");
    const long long total = (long long)N * (H / 2) * (W / 2) * 4;
    HGK_REQUIRE(total < (1LL << 32), "hgk_stem_s2d_image: too many pixels");
    long long g = (total + 255) / 256;
    if (g > (long long)kNumSMs * 16) g = (long long)kNumSMs * 16;
    stem_s2d_image_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(img, H, W, (unsigned)total, xs);
    HGK_CHECK_LAUNCH("hgk_stem_s2d_image");
    return HGK_OK;
}

extern "C" int hgk_stem_s2d_weight(const float* w, int Cout, float* ws, void* stream) {
    HGK_REQUIRE(w && ws && Cout > 0, "hgk_stem_s2d_weight: bad arguments");
    stem_s2d_weight_kernel<<<(Cout * 512 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Cout, ws);
    HGK_CHECK_LAUNCH("hgk_stem_s2d_weight");
    return HGK_OK;
}
