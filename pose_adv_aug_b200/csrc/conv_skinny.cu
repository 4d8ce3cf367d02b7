// fp32 1x1 convolutions with a 16-channel side: the heat-map head out_conv (C -> 16,
// models/asn_stacked_hg.py:248,329), in_conv (16 -> C, :279,333) and their data gradients.  Both are pure
// HBM streams (100 MB of activations against <1 GFLOP): the generic implicit-GEMM tiles (128 x 64) spend their
// time on padding.  Same contract as hgk_conv_nhwc (BN+ReLU on load, bias, shortcut, accumulate); no
// statistics epilogue (no BatchNorm follows these layers).
//
//   conv_n16_kernel : y[p][0:16] = sum_k T(x)[p][k] w[k][0:16]   -- a warp per pixel group, lane = 4-channel
//                     quad of x with its 4x16 weights in registers, 512-byte coalesced loads, 16 warp-shuffle
//                     reduce-scatter steps per pixel, one 64-byte store.
//   conv_k16_kernel : y[p][n] = sum_{k<16} T(x)[p][k] w[k][n]    -- lane = 4 output channels with its 16x4
//                     weights in registers; x rows are warp-uniform (broadcast) loads; 512-byte stores.
#include "common.cuh"
#include "conv_args.cuh"

namespace hgk {

// named barrier of the CQ warps that share a pixel group (immediate barrier ids: a register id reserves all 16)
template <int CQ>
__device__ __forceinline__ void pair_sync(int wp) {
    if (wp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * CQ) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"n"(32 * CQ) : "memory");
}

// ---- C -> 16.  CQ = warps cooperating on one pixel (Cin = 128 * CQ); CTA = 4 warps, N16_PX pixels per warp step (with the
// prefetched next step: 2 * N16_PX * 512 bytes in flight per warp -- with two pixels the kernel ran at a quarter of the HBM rate)
constexpr int N16_PX = 4;
template <int CQ>
__global__ void __launch_bounds__(128, 4) conv_n16_kernel(const ConvArgs a) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    __shared__ float part[4][N16_PX][16];           // [warp][pixel of the step][output]: partial sums when CQ > 1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cq = warp % CQ;                       // which 128-channel slice of the pixel this warp owns
    const int wp = warp / CQ;                       // pixel-group slot of the warp inside the CTA
    constexpr int WPB = 4 / CQ;                     // pixel groups per CTA per step
    const int c0 = cq * 128 + lane * 4;
    // weights of this lane's four input channels: w[k][0:16], k = c0 .. c0+3
    float4 wr[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) wr[k][q] = ldg4(a.w + (size_t)(c0 + k) * 16 + q * 4);
    float4 sc, sh;
    load_affine4(a.x.scale, a.x.shift, c0, sc, sh);
    const bool has_aff = a.x.scale != nullptr;
    const bool has_res = a.res.z != nullptr, res_aff = a.res.scale != nullptr;
    const long long groups = (a.P + N16_PX - 1) / N16_PX;
    const long long gstride = (long long)gridDim.x * WPB;
    const float* xz = a.x.z + c0;
    auto load2 = [&](long long g, float4 (&xv)[N16_PX]) {
#pragma unroll
        for (int i = 0; i < N16_PX; ++i) {
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const long long p = g * N16_PX + i;
            if (p < a.P) xv[i] = ldg4(xz + p * a.Cin);
        }
    };
    float4 xn[N16_PX];
    long long g = (long long)blockIdx.x * WPB + wp;
    load2(g, xn);
    // the warps of a pixel group run the same number of steps (same g): named barriers pair them up
    for (; g < groups; g += gstride) {
        float4 xv[N16_PX];
#pragma unroll
        for (int i = 0; i < N16_PX; ++i) xv[i] = xn[i];
        load2(g + gstride, xn);                      // prefetch the next step while this one computes
        const long long p0 = g * N16_PX;
        float out[N16_PX];                           // after the reduce-scatter: output (lane & 15) of pixel i
#pragma unroll
        for (int i = 0; i < N16_PX; ++i) {
            if (has_aff && p0 + i < a.P) xv[i] = act4(xv[i], sc, sh, a.x.relu);
            float acc[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[q * 4 + 0] = xv[i].x * wr[0][q].x + xv[i].y * wr[1][q].x + xv[i].z * wr[2][q].x + xv[i].w * wr[3][q].x;
                acc[q * 4 + 1] = xv[i].x * wr[0][q].y + xv[i].y * wr[1][q].y + xv[i].z * wr[2][q].y + xv[i].w * wr[3][q].y;
                acc[q * 4 + 2] = xv[i].x * wr[0][q].z + xv[i].y * wr[1][q].z + xv[i].z * wr[2][q].z + xv[i].w * wr[3][q].z;
                acc[q * 4 + 3] = xv[i].x * wr[0][q].w + xv[i].y * wr[1][q].w + xv[i].z * wr[2][q].w + xv[i].w * wr[3][q].w;
            }
            // reduce-scatter over the 32 lanes: the step with offset h keeps outputs [h, 2h) on lanes with bit h set and
            // [0, h) on the others; after offsets 8, 4, 2, 1 lane l holds output (l & 15) summed over its 16-lane half;
            // the last step adds the two halves
#pragma unroll
            for (int half = 8; half >= 1; half >>= 1) {
                const bool up = (lane & half) != 0;
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    const float mine = up ? acc[j + half] : acc[j];
                    const float give = up ? acc[j] : acc[j + half];
                    acc[j] = mine + __shfl_xor_sync(0xffffffffu, give, half);
                }
            }
            out[i] = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 16);
        }
        const int o = lane & 15;
        if (CQ > 1) {
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < N16_PX; ++i) part[warp][i][o] = out[i];
            }
            pair_sync<CQ>(wp);
        }
        if (cq == 0 && lane < 16) {
#pragma unroll
            for (int i = 0; i < N16_PX; ++i) {
                if (p0 + i >= a.P) break;
                float v = out[i];
                if (CQ > 1) {
#pragma unroll
                    for (int c = 1; c < CQ; ++c) v += part[warp + c][i][o];
                }
                if (a.bias != nullptr) v += __ldg(a.bias + o);
                const long long idx = (p0 + i) * 16 + o;
                if (has_res) {
                    float r = __ldg(a.res.z + idx);
                    if (res_aff) r = act1(r, __ldg(a.res.scale + o), __ldg(a.res.shift + o), a.res.relu);
                    v += r;
                }
                if (a.accumulate) v += a.y[idx];
                a.y[idx] = v;
            }
        }
        if (CQ > 1) pair_sync<CQ>(wp);               // `part` is rewritten by the next step
    }
}

// ---- 16 -> C.  A warp covers 128 output channels of NPX consecutive pixels per step; CTA = 8 warps.  The NPX x 16 input values
// of a step are ONE coalesced 16-byte load per lane (lane = pixel (lane >> 2), channel quad (lane & 3)) and reach the FMAs by
// warp shuffles; the shortcut / previous-output rows of all NPX pixels are requested before the arithmetic starts (NPX x 512
// bytes per operand and warp in flight: with one pixel per step the launch ran at a fifth of the HBM rate).
template <int NPX, bool RES, bool ACC>
__global__ void __launch_bounds__(256, 2) conv_k16_kernel(const ConvArgs a) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nq = a.Cout >> 7;                      // 128-channel slices per pixel
    const int slice = warp % nq;                     // host guarantees 8 % nq == 0
    const int wp = warp / nq, WPB = 8 / nq;
    const int n0 = slice * 128 + lane * 4;
    float4 wr[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) wr[k] = ldg4(a.w + (size_t)k * a.Cout + n0);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias != nullptr) bv = ldg4(a.bias + n0);
    float4 rs, rt;
    load_affine4(a.res.scale, a.res.shift, n0, rs, rt);
    const bool has_aff = a.x.scale != nullptr, res_aff = a.res.scale != nullptr;
    float4 xsc, xsh;
    load_affine4(a.x.scale, a.x.shift, (lane & 3) * 4, xsc, xsh);
    const long long groups = (a.P + NPX - 1) / NPX;
    const long long gstride = (long long)gridDim.x * WPB;
    for (long long g = (long long)blockIdx.x * WPB + wp; g < groups; g += gstride) {
        const long long p0 = g * NPX;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < NPX * 4 && p0 + (lane >> 2) < a.P) {
            xv = ldg4(a.x.z + (p0 + (lane >> 2)) * 16 + (lane & 3) * 4);
            if (has_aff) xv = act4(xv, xsc, xsh, a.x.relu);
        }
        float4 rr[RES ? NPX : 1], oo[ACC ? NPX : 1];
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const bool ok = p0 + i < a.P;
            if (RES) rr[RES ? i : 0] = ok ? ldg4(a.res.z + (p0 + i) * a.Cout + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ACC) oo[ACC ? i : 0] = ok ? ld4(a.y + (p0 + i) * a.Cout + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            float4 v = bv;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int src = i * 4 + q;
                const float x0 = __shfl_sync(0xffffffffu, xv.x, src), x1 = __shfl_sync(0xffffffffu, xv.y, src);
                const float x2 = __shfl_sync(0xffffffffu, xv.z, src), x3 = __shfl_sync(0xffffffffu, xv.w, src);
                const float4 w0 = wr[q * 4 + 0], w1 = wr[q * 4 + 1], w2 = wr[q * 4 + 2], w3 = wr[q * 4 + 3];
                v.x += x0 * w0.x + x1 * w1.x + x2 * w2.x + x3 * w3.x;
                v.y += x0 * w0.y + x1 * w1.y + x2 * w2.y + x3 * w3.y;
                v.z += x0 * w0.z + x1 * w1.z + x2 * w2.z + x3 * w3.z;
                v.w += x0 * w0.w + x1 * w1.w + x2 * w2.w + x3 * w3.w;
            }
            if (RES) {
                float4 q = rr[RES ? i : 0];
                if (res_aff) q = act4(q, rs, rt, a.res.relu);
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
            }
            if (ACC) {
                const float4 o = oo[ACC ? i : 0];
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            if (p0 + i < a.P) st4(a.y + (p0 + i) * a.Cout + n0, v);
        }
    }
}

// returns 1 when one of the skinny kernels took the launch, 0 when the shape is not covered
int conv_skinny_try(const ConvArgs& a, cudaStream_t st) {
    static int off = -1;
    if (off < 0) {
        const char* e = getenv("HGK_SKINNY_OFF");
        off = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (off || a.ksize != 1 || a.stat_sum != nullptr) return 0;
    if (a.Cout == 16 && (a.Cin == 128 || a.Cin == 256)) {
        const int grid = kNumSMs * 8;
        if (a.Cin == 128) launch_pdl(conv_n16_kernel<1>, dim3(grid), dim3(128), 0, st, a);
        else launch_pdl(conv_n16_kernel<2>, dim3(grid), dim3(128), 0, st, a);
        return 1;
    }
    if (a.Cin == 16 && (a.Cout == 128 || a.Cout == 256)) {
        const bool res = a.res.z != nullptr, acc = a.accumulate != 0;
        const int grid = kNumSMs * 4;
        if (res && acc) launch_pdl(conv_k16_kernel<4, true, true>, dim3(grid), dim3(256), 0, st, a);
        else if (res) launch_pdl(conv_k16_kernel<8, true, false>, dim3(grid), dim3(256), 0, st, a);
        else if (acc) launch_pdl(conv_k16_kernel<8, false, true>, dim3(grid), dim3(256), 0, st, a);
        else launch_pdl(conv_k16_kernel<8, false, false>, dim3(grid), dim3(256), 0, st, a);
        return 1;
    }
    return 0;
}

}  // namespace hgk
