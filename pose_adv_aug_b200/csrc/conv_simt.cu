// fp32 SIMT implicit-GEMM convolution (1x1 / 3x3 pad 1, NHWC) with fused BN+ReLU-on-load
// prologue and bias / residual / accumulate / BN-statistics epilogue, plus its weight-gradient
// kernel.  This is the exact-fp32 path: it serves every channel count (C % 4 == 0), is the
// on-device reference for the tcgen05 kernels, and handles the shapes those do not cover.
//
// Replaces: nn.Conv2d + nn.BatchNorm2d(batch statistics) + nn.ReLU + `out += shortcut`
// of _Residual.forward (reference models/asn_stacked_hg.py:30-49) and their autograd.
#include "common.cuh"
#include "conv_args.cuh"

namespace hgk {

constexpr int CBM = 128, CBK = 16, CNT = 256;

__device__ __forceinline__ float f4c(const float4& v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

template <int BN>
__global__ void __launch_bounds__(CNT, BN == 64 ? 2 : 1) conv_igemm_simt(const ConvArgs a) {
    constexpr int TN = BN / 16;    // output columns per thread (4 or 8)
    constexpr int NB = BN / 64;    // float4 B loads per thread per k-tile
    __shared__ __align__(16) float smem[2 * CBM * CBK + 2 * CBK * BN];
    float* As = smem;                      // [2][BM][BK]   (k-chunk XOR-swizzled on row bit 3)
    float* Bs = smem + 2 * CBM * CBK;      // [2][BK][BN]

    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * CBM;
    const int n0 = blockIdx.y * BN;
    const int taps = a.ksize * a.ksize;
    const int KC = (a.Cin + CBK - 1) / CBK;
    const int T = taps * KC;
    const int HW = a.H * a.W;

    // ---- A loader state: two pixel rows per thread, one float4 (4 channels) each ----
    const int a_kv = tid & 3;
    int a_h[2], a_w[2];
    long long a_p[2];
    bool a_ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int row = (tid >> 2) + r * 64;
        long long p = m0 + row;
        a_ok[r] = p < a.P;
        long long pp = a_ok[r] ? p : 0;
        int rem = (int)(pp % HW);
        a_h[r] = rem / a.W;
        a_w[r] = rem - a_h[r] * a.W;
        a_p[r] = pp;
    }
    float4 a_reg[2], b_reg[NB];

    auto load_tile = [&](int it) {
        int tap = it / KC;
        int k0 = (it - tap * KC) * CBK;
        int dh = 0, dw = 0;
        if (a.ksize == 3) {
            dh = tap / 3 - 1;
            dw = tap - (tap / 3) * 3 - 1;
        }
        int c = k0 + a_kv * 4;
        float4 s, t;
        load_affine4(a.x.scale, a.x.shift, c < a.Cin ? c : 0, s, t);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            bool ok = a_ok[r] && c < a.Cin && (unsigned)(a_h[r] + dh) < (unsigned)a.H &&
                      (unsigned)(a_w[r] + dw) < (unsigned)a.W;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                v = ldg4(a.x.z + (a_p[r] + dh * a.W + dw) * a.Cin + c);
                if (a.x.scale != nullptr) v = act4(v, s, t, a.x.relu);
            }
            a_reg[r] = v;
        }
        int wt = a.flip ? (taps - 1 - tap) : tap;
        const float* wb = a.w + (size_t)wt * a.Cin * a.Cout;
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            int idx = tid + r * CNT;
            int k = idx / (BN / 4);
            int nv = idx - k * (BN / 4);
            int n = n0 + nv * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + k < a.Cin && n < a.Cout) v = ldg4(wb + (size_t)(k0 + k) * a.Cout + n);
            b_reg[r] = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int row = (tid >> 2) + r * 64;
            int kc = a_kv ^ (((row >> 3) & 1) << 1);
            st4(As + buf * CBM * CBK + row * CBK + kc * 4, a_reg[r]);
        }
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            int idx = tid + r * CNT;
            int k = idx / (BN / 4);
            int nv = idx - k * (BN / 4);
            st4(Bs + buf * CBK * BN + k * BN + nv * 4, b_reg[r]);
        }
    };

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tile(0);
    store_tile(0);
    __syncthreads();
    const int a_sw = (ty & 1) << 1;   // rows ty*8+i: bit 3 of the row == ty & 1
    for (int it = 0; it < T; ++it) {
        const int buf = it & 1;
        if (it + 1 < T) load_tile(it + 1);
        const float* Ab = As + buf * CBM * CBK + (ty * 8) * CBK;
        const float* Bb = Bs + buf * CBK * BN + tx * 4;
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = ld4(Ab + i * CBK + ((kq ^ a_sw) << 2));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float4 b0 = ld4(Bb + (kq * 4 + kk) * BN);
                float4 b1 = b0;
                if (TN == 8) b1 = ld4(Bb + (kq * 4 + kk) * BN + 64);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float av_ = f4c(av[i], kk);
                    acc[i][0] = fmaf(av_, b0.x, acc[i][0]);
                    acc[i][1] = fmaf(av_, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(av_, b0.z, acc[i][2]);
                    acc[i][3] = fmaf(av_, b0.w, acc[i][3]);
                    if (TN == 8) {
                        acc[i][4] = fmaf(av_, b1.x, acc[i][4]);
                        acc[i][5] = fmaf(av_, b1.y, acc[i][5]);
                        acc[i][6] = fmaf(av_, b1.z, acc[i][6]);
                        acc[i][7] = fmaf(av_, b1.w, acc[i][7]);
                    }
                }
            }
        }
        if (it + 1 < T) store_tile(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, residual, accumulate, store, BN statistics ----
    const bool do_stats = a.stat_sum != nullptr;
    float s1[TN], s2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) s1[j] = s2[j] = 0.f;
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
        const int n = n0 + g * 64 + tx * 4;
        if (n >= a.Cout) continue;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias != nullptr) bv = ldg4(a.bias + n);
        float4 rs, rt;
        load_affine4(a.res.scale, a.res.shift, n, rs, rt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            long long p = m0 + ty * 8 + i;
            if (p >= a.P) continue;
            float4 v = make_float4(acc[i][g * 4 + 0] + bv.x, acc[i][g * 4 + 1] + bv.y, acc[i][g * 4 + 2] + bv.z,
                                   acc[i][g * 4 + 3] + bv.w);
            if (a.res.z != nullptr) {
                float4 r = ldg4(a.res.z + p * a.Cout + n);
                if (a.res.scale != nullptr) r = act4(r, rs, rt, a.res.relu);
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            float* yp = a.y + p * a.Cout + n;
            if (a.accumulate) {
                float4 o = ld4(yp);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            st4(yp, v);
            if (do_stats) {
                s1[g * 4 + 0] += v.x; s2[g * 4 + 0] = fmaf(v.x, v.x, s2[g * 4 + 0]);
                s1[g * 4 + 1] += v.y; s2[g * 4 + 1] = fmaf(v.y, v.y, s2[g * 4 + 1]);
                s1[g * 4 + 2] += v.z; s2[g * 4 + 2] = fmaf(v.z, v.z, s2[g * 4 + 2]);
                s1[g * 4 + 3] += v.w; s2[g * 4 + 3] = fmaf(v.w, v.w, s2[g * 4 + 3]);
            }
        }
    }
    if (do_stats) {
        // per-thread fp32 partials over 8 pixels -> fp64 across the 16 row groups -> fp64 atomics
        double* red = reinterpret_cast<double*>(smem);   // [8 warps][BN][2]
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            double d1 = (double)s1[j], d2 = (double)s2[j];
            d1 += __shfl_xor_sync(0xffffffffu, d1, 16);
            d2 += __shfl_xor_sync(0xffffffffu, d2, 16);
            if (lane < 16) {
                int col = (j >> 2) * 64 + tx * 4 + (j & 3);
                red[(warp * BN + col) * 2 + 0] = d1;
                red[(warp * BN + col) * 2 + 1] = d2;
            }
        }
        __syncthreads();
        if (tid < BN && n0 + tid < a.Cout) {
            double d1 = 0.0, d2 = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                d1 += red[(w * BN + tid) * 2 + 0];
                d2 += red[(w * BN + tid) * 2 + 1];
            }
            atomicAdd(a.stat_sum + n0 + tid, d1);
            atomicAdd(a.stat_sq + n0 + tid, d2);
        }
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient: dW[tap][co][ci] += sum_p dz[p][co] * T(x)[p+off(tap)][ci]   (split over pixels)
// ------------------------------------------------------------------------------------------
struct WgradArgs {
    Act x;
    int N, H, W, Cin;
    const float* dz;
    int Cout, ksize;
    float* dw;
    long long s_co, s_ci, s_tap;
    float* dbias;
    long long P, chunk;
    int nt;   // number of ci tiles
};

template <int TM>   // rows (co) per thread: 4 -> BM 64, 8 -> BM 128
__global__ void __launch_bounds__(256) conv_wgrad_simt(const WgradArgs a) {
    constexpr int BM = TM * 16, BN = 64, BK = 16;
    constexpr int NA = BM / 64;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int mi = blockIdx.x / a.nt, ni = blockIdx.x - mi * a.nt;
    const int co0 = mi * BM, ci0 = ni * BN;
    const int tap = blockIdx.y;
    int dh = 0, dw_ = 0;
    if (a.ksize == 3) {
        dh = tap / 3 - 1;
        dw_ = tap - (tap / 3) * 3 - 1;
    }
    const long long p_begin = (long long)blockIdx.z * a.chunk;
    const long long p_end = (p_begin + a.chunk < a.P) ? (p_begin + a.chunk) : a.P;
    if (p_begin >= p_end) return;
    const int T = (int)((p_end - p_begin + BK - 1) / BK);
    const int HW = a.H * a.W;
    const bool do_bias = a.dbias != nullptr && ni == 0 && tap == 0;

    const int b_k = tid >> 4, b_nv = tid & 15;
    const int b_c = ci0 + b_nv * 4;
    float4 bs, bt;
    load_affine4(a.x.scale, a.x.shift, b_c < a.Cin ? b_c : 0, bs, bt);
    float4 a_reg[NA], b_reg;

    auto load_tile = [&](int it) {
        long long pbase = p_begin + (long long)it * BK;
#pragma unroll
        for (int r = 0; r < NA; ++r) {
            int idx = tid + r * 256;
            int k = idx / (BM / 4);
            int mv = idx - k * (BM / 4);
            long long p = pbase + k;
            int co = co0 + mv * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < p_end && co < a.Cout) v = ldg4(a.dz + p * a.Cout + co);
            a_reg[r] = v;
        }
        {
            long long p = pbase + b_k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < p_end && b_c < a.Cin) {
                int rem = (int)(p % HW);
                int h = rem / a.W, w = rem - h * a.W;
                if ((unsigned)(h + dh) < (unsigned)a.H && (unsigned)(w + dw_) < (unsigned)a.W) {
                    v = ldg4(a.x.z + (p + dh * a.W + dw_) * a.Cin + b_c);
                    if (a.x.scale != nullptr) v = act4(v, bs, bt, a.x.relu);
                }
            }
            b_reg = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int r = 0; r < NA; ++r) {
            int idx = tid + r * 256;
            int k = idx / (BM / 4);
            int mv = idx - k * (BM / 4);
            st4(&As[buf][k][mv * 4], a_reg[r]);
        }
        st4(&Bs[buf][b_k][b_nv * 4], b_reg);
    };

    const int ty = tid >> 4, tx = tid & 15;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int it = 0; it < T; ++it) {
        const int buf = it & 1;
        if (it + 1 < T) load_tile(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                float4 v = ld4(&As[buf][k][ty * TM + q * 4]);
                av[q * 4 + 0] = v.x; av[q * 4 + 1] = v.y; av[q * 4 + 2] = v.z; av[q * 4 + 3] = v.w;
            }
            float4 b = ld4(&Bs[buf][k][tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
            }
        }
        if (do_bias && tid < BM) {
#pragma unroll
            for (int k = 0; k < BK; ++k) bsum += As[buf][k][tid];
        }
        if (it + 1 < T) store_tile(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int co = co0 + ty * TM + i;
        if (co >= a.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = ci0 + tx * 4 + j;
            if (ci < a.Cin) atomicAdd(a.dw + co * a.s_co + ci * a.s_ci + tap * a.s_tap, acc[i][j]);
        }
    }
    if (do_bias && tid < BM && co0 + tid < a.Cout) atomicAdd(a.dbias + co0 + tid, bsum);
}

// multi-tensor OIHW repack (see hgk_pack_weights)
__global__ void pack_weights_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                    const long long* __restrict__ table, int n_entries) {
    const int e = blockIdx.y;
    if (e >= n_entries) return;
    const long long* t = table + (size_t)e * 6;
    const long long so = t[0], d0 = t[1];
    const int O = (int)t[2], I = (int)t[3], taps = (int)t[4], mode = (int)t[5];
    const long long total = (long long)O * I * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i enumerates the destination so that stores are coalesced
        long long tap, o, ci;
        if (mode == 0) {          // dst[tap][i][o]
            o = i % O;
            ci = (i / O) % I;
            tap = i / ((long long)O * I);
        } else {                  // dst[tap][o][i]
            ci = i % I;
            o = (i / I) % O;
            tap = i / ((long long)O * I);
        }
        dst[d0 + i] = __ldg(src + so + (o * I + ci) * taps + tap);
    }
}

}  // namespace hgk

namespace hgk {
int conv_skinny_try(const ConvArgs& a, cudaStream_t st);      // conv_skinny.cu
}
using namespace hgk;

extern "C" int hgk_conv_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                             int N, int H, int W, int Cin,
                             const float* w, int ksize, int flip, const float* bias, int Cout,
                             const float* res, const float* res_scale, const float* res_shift, int res_relu,
                             float* y, int accumulate, double* stat_sum, double* stat_sq,
                             int path, void* stream) {
    HGK_REQUIRE(x && w && y, "hgk_conv_nhwc: null pointer");
    HGK_REQUIRE(N > 0 && H > 0 && W > 0, "hgk_conv_nhwc: empty tensor N=%d H=%d W=%d", N, H, W);
    HGK_REQUIRE(ksize == 1 || ksize == 3, "hgk_conv_nhwc: ksize must be 1 or 3 (got %d)", ksize);
    HGK_REQUIRE(Cin > 0 && Cout > 0 && Cin % 4 == 0 && Cout % 4 == 0,
                "hgk_conv_nhwc: channel counts must be positive multiples of 4 (Cin=%d Cout=%d)", Cin, Cout);
    HGK_REQUIRE((stat_sum == nullptr) == (stat_sq == nullptr), "hgk_conv_nhwc: stat_sum/stat_sq must both be set");
    HGK_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), "hgk_conv_nhwc: x scale/shift must both be set");
    HGK_REQUIRE((res_scale == nullptr) == (res_shift == nullptr), "hgk_conv_nhwc: res scale/shift must both be set");
    ConvArgs a;
    a.x = Act{x, x_scale, x_shift, x_relu};
    a.N = N; a.H = H; a.W = W; a.Cin = Cin;
    a.w = w; a.ksize = ksize; a.flip = flip; a.bias = bias; a.Cout = Cout;
    a.res = Act{res, res_scale, res_shift, res_relu};
    a.y = y; a.accumulate = accumulate; a.stat_sum = stat_sum; a.stat_sq = stat_sq;
    a.P = (long long)N * H * W;
    a.bz = nullptr; a.bscale = a.bshift = a.bmean = a.binvstd = nullptr; a.brelu = 0;
    a.ffin = BnFwdFin{}; a.bfin = BnBwdFin{}; a.ap = BnApply{};
    cudaStream_t st = (cudaStream_t)stream;
    HGK_REQUIRE(path == 0 || path == 1, "hgk_conv_nhwc: this entry point is the fp32 SIMT kernel; use hgk_conv_tc_nhwc for tcgen05");
    long long mt = (a.P + CBM - 1) / CBM;
    HGK_REQUIRE(mt < 2147483647LL, "hgk_conv_nhwc: too many pixels");
    if (conv_skinny_try(a, st) == 1) {            // 16-channel heads: streaming kernels of conv_skinny.cu
        HGK_CHECK_LAUNCH("hgk_conv_nhwc");
        return HGK_OK;
    }
    if (Cout > 64) {
        dim3 grid((unsigned)mt, (unsigned)((Cout + 127) / 128));
        conv_igemm_simt<128><<<grid, CNT, 0, st>>>(a);
    } else {
        dim3 grid((unsigned)mt, 1);
        conv_igemm_simt<64><<<grid, CNT, 0, st>>>(a);
    }
    HGK_CHECK_LAUNCH("hgk_conv_nhwc");
    return HGK_OK;
}

extern "C" int hgk_conv_wgrad_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                                   int N, int H, int W, int Cin, const float* dz, int Cout, int ksize,
                                   float* dw, long long s_co, long long s_ci, long long s_tap,
                                   float* dbias, void* stream) {
    HGK_REQUIRE(x && dz && dw, "hgk_conv_wgrad_nhwc: null pointer");
    HGK_REQUIRE(N > 0 && H > 0 && W > 0, "hgk_conv_wgrad_nhwc: empty tensor");
    HGK_REQUIRE(ksize == 1 || ksize == 3, "hgk_conv_wgrad_nhwc: ksize must be 1 or 3 (got %d)", ksize);
    HGK_REQUIRE(Cin > 0 && Cout > 0 && Cin % 4 == 0 && Cout % 4 == 0,
                "hgk_conv_wgrad_nhwc: channel counts must be positive multiples of 4 (Cin=%d Cout=%d)", Cin, Cout);
    HGK_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), "hgk_conv_wgrad_nhwc: x scale/shift must both be set");
    WgradArgs a;
    a.x = Act{x, x_scale, x_shift, x_relu};
    a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.dz = dz; a.Cout = Cout; a.ksize = ksize;
    a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap; a.dbias = dbias;
    a.P = (long long)N * H * W;
    const int taps = ksize * ksize;
    const bool big = Cout > 64;
    const int BM = big ? 128 : 64;
    const int mtiles = (Cout + BM - 1) / BM;
    a.nt = (Cin + 63) / 64;
    const int tiles = mtiles * a.nt * taps;
    long long max_splits = (a.P + 127) / 128;                 // >= 8 k-iterations per CTA
    long long want = (4LL * kNumSMs + tiles - 1) / tiles;     // ~4 CTAs per SM in flight
    long long splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
    if (splits > 65535) splits = 65535;
    long long chunk = (a.P + splits - 1) / splits;
    chunk = (chunk + 15) / 16 * 16;
    splits = (a.P + chunk - 1) / chunk;
    a.chunk = chunk;
    dim3 grid((unsigned)(mtiles * a.nt), (unsigned)taps, (unsigned)splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (big) conv_wgrad_simt<8><<<grid, 256, 0, st>>>(a);
    else conv_wgrad_simt<4><<<grid, 256, 0, st>>>(a);
    HGK_CHECK_LAUNCH("hgk_conv_wgrad_nhwc");
    return HGK_OK;
}

extern "C" int hgk_pack_weights(const float* src_base, float* dst_base, const long long* table, int n_entries,
                                void* stream) {
    HGK_REQUIRE(src_base && dst_base && table, "hgk_pack_weights: null pointer");
    if (n_entries <= 0) return HGK_OK;
    HGK_REQUIRE(n_entries <= 65535, "hgk_pack_weights: too many entries");
    dim3 grid(16, (unsigned)n_entries);
    pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_base, dst_base, table, n_entries);
    HGK_CHECK_LAUNCH("hgk_pack_weights");
    return HGK_OK;
}
