// nn.BatchNorm2d (eps 1e-5, momentum 0.1) of the reference (models/asn_stacked_hg.py:19,22,25,
// 224,243) split the B200 way: the per-channel sums are produced by the convolution epilogues
// (fp64 accumulators), `bn_finalize` turns them into (scale, shift) that the *consumers* apply on
// load, and the backward is reduce -> finalize -> apply over the stored pre-BN tensor.
#include "common.cuh"
#include "bn_fin.cuh"

namespace hgk {

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sq, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* scale,
                                   float* shift, float* save_mean, float* save_invstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mean = sum[c] / count;
    double var = sq[c] / count - mean * mean;      // fp64: no cancellation issue at these magnitudes
    if (var < 0.0) var = 0.0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    float sc = (float)((double)gamma[c] * invstd);
    scale[c] = sc;
    shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)invstd;
    if (running_mean != nullptr) {
        double unb = count > 1.0 ? var * (count / (count - 1.0)) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
    }
}

__global__ void bn_eval_prepare_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                       float* scale, float* shift, float* save_mean, float* save_invstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double invstd = 1.0 / sqrt((double)rv[c] + (double)eps);
    scale[c] = (float)((double)gamma[c] * invstd);
    shift[c] = (float)((double)beta[c] - (double)rm[c] * (double)gamma[c] * invstd);
    save_mean[c] = rm[c];
    save_invstd[c] = (float)invstd;
}

// One block = 256 threads = rows_par pixel rows x (C/4) channel quads; grid-strided over pixels.
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            int relu, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, long long P, int C,
                                                            double* sum_g, double* sum_gx, int cq_per_blk, int rows_par,
                                                            long long rows_per_blk, BnBwdFin fin) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    extern __shared__ double red[];      // [rows_par][cq_per_blk*4][2]
    const int tid = threadIdx.x;
    const int cq = tid % cq_per_blk, row = tid / cq_per_blk;
    const int c = (blockIdx.y * cq_per_blk + cq) * 4;
    const bool active = row < rows_par && c < C;
    double g1[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};
    if (active) {
        float4 s = ldg4(scale + c), t = ldg4(shift + c), mu = ldg4(mean + c), is = ldg4(invstd + c);
        long long p0 = (long long)blockIdx.x * rows_per_blk;
        long long p1 = p0 + rows_per_blk < P ? p0 + rows_per_blk : P;
        float f1[4] = {0, 0, 0, 0}, f2[4] = {0, 0, 0, 0};
        int cnt = 0;
        // four rows per trip, all eight 16-byte loads issued before the arithmetic (the kernel is a pure HBM stream: with
        // one row per trip a thread had 32 bytes in flight and the launch ran at a third of the copy bandwidth)
        constexpr int UR = 4;
        const long long rstep = (long long)rows_par * UR;
        for (long long p = p0 + row; p < p1; p += rstep) {
            float4 d[UR], v[UR];
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const long long pp = p + (long long)u * rows_par;
                if (pp < p1) {
                    d[u] = ldg4(dy + pp * C + c);
                    v[u] = ldg4(z + pp * C + c);
                } else {                      // absent row: g = 0 contributes nothing to either sum
                    d[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    v[u] = mu;
                }
            }
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const float gx = (relu && fmaf(v[u].x, s.x, t.x) <= 0.f) ? 0.f : d[u].x;
                const float gy = (relu && fmaf(v[u].y, s.y, t.y) <= 0.f) ? 0.f : d[u].y;
                const float gz = (relu && fmaf(v[u].z, s.z, t.z) <= 0.f) ? 0.f : d[u].z;
                const float gw = (relu && fmaf(v[u].w, s.w, t.w) <= 0.f) ? 0.f : d[u].w;
                f1[0] += gx; f2[0] = fmaf(gx, (v[u].x - mu.x) * is.x, f2[0]);
                f1[1] += gy; f2[1] = fmaf(gy, (v[u].y - mu.y) * is.y, f2[1]);
                f1[2] += gz; f2[2] = fmaf(gz, (v[u].z - mu.z) * is.z, f2[2]);
                f1[3] += gw; f2[3] = fmaf(gw, (v[u].w - mu.w) * is.w, f2[3]);
            }
            if (++cnt == 4) {            // flush short fp32 partials (16 rows) into fp64
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g1[j] += (double)f1[j]; g2[j] += (double)f2[j];
                    f1[j] = 0.f; f2[j] = 0.f;
                }
                cnt = 0;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { g1[j] += (double)f1[j]; g2[j] += (double)f2[j]; }
    }
    if (row < rows_par) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[((row * cq_per_blk + cq) * 4 + j) * 2 + 0] = g1[j];
            red[((row * cq_per_blk + cq) * 4 + j) * 2 + 1] = g2[j];
        }
    }
    __syncthreads();
    const int nch = cq_per_blk * 4;
    if (tid < nch) {
        int cc = blockIdx.y * nch + tid;
        if (cc < C) {
            double a = 0.0, b = 0.0;
            for (int r = 0; r < rows_par; ++r) {
                a += red[(r * nch + tid) * 2 + 0];
                b += red[(r * nch + tid) * 2 + 1];
            }
            atomicAdd(sum_g + cc, a);
            atomicAdd(sum_gx + cc, b);
        }
    }
    if (fin.ticket != nullptr) {          // fused bn_bwd_finalize: run by the CTA that arrives last
        if (last_cta_arrives(fin.ticket, gridDim.x * gridDim.y)) bn_bwd_finalize_cta(fin, sum_g, sum_gx, (double)P, C);
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sum_g, const double* __restrict__ sum_gx, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int training, float* dgamma, float* dbeta,
                                       float* cA, float* cB, float* cC, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double sg = sum_g[c], sgx = sum_gx[c];
    if (dgamma != nullptr) dgamma[c] += (float)sgx;
    if (dbeta != nullptr) dbeta[c] += (float)sg;
    double A = (double)gamma[c] * (double)invstd[c];
    if (training) {
        // dz = cA * (g - cC - (z - mean) * cB): the centred form keeps the fp32 rounding error of the
        // projection term proportional to |z - mean| instead of |mean| (no cancellation)
        double c1 = sg / count, c2 = sgx / count;
        cA[c] = (float)A;
        cB[c] = (float)(c2 * (double)invstd[c]);
        cC[c] = (float)c1;
    } else {
        cA[c] = (float)A;
        cB[c] = 0.f;
        cC[c] = 0.f;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dy, const float* __restrict__ z,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           int relu, const float* __restrict__ mean,
                                                           const float* __restrict__ cA, const float* __restrict__ cB,
                                                           const float* __restrict__ cC, long long total4, int C4) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto one = [&](long long i, float4 d, float4 v, float4 s, float4 t, float4 a, float4 b, float4 cc, float4 mu) {
        float gx = (relu && fmaf(v.x, s.x, t.x) <= 0.f) ? 0.f : d.x;
        float gy = (relu && fmaf(v.y, s.y, t.y) <= 0.f) ? 0.f : d.y;
        float gz = (relu && fmaf(v.z, s.z, t.z) <= 0.f) ? 0.f : d.z;
        float gw = (relu && fmaf(v.w, s.w, t.w) <= 0.f) ? 0.f : d.w;
        float4 o;
        o.x = a.x * ((gx - cc.x) - (v.x - mu.x) * b.x);
        o.y = a.y * ((gy - cc.y) - (v.y - mu.y) * b.y);
        o.z = a.z * ((gz - cc.z) - (v.z - mu.z) * b.z);
        o.w = a.w * ((gw - cc.w) - (v.w - mu.w) * b.w);
        st4(dy + i * 4, o);
    };
    if (stride % C4 == 0) {
        // the grid stride is a multiple of the channel-quad count: a thread stays on ONE channel quad, its seven per-channel
        // vectors are loaded once and four items are in flight per trip
        const int c = (int)(i0 % C4) * 4;
        const float4 s = ldg4(scale + c), t = ldg4(shift + c), a = ldg4(cA + c), b = ldg4(cB + c), cc = ldg4(cC + c);
        const float4 mu = ldg4(mean + c);
        constexpr int UR = 4;
        for (long long i = i0; i < total4; i += stride * UR) {
            float4 d[UR], v[UR];
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const long long ii = i + u * stride;
                if (ii < total4) { d[u] = ld4(dy + ii * 4); v[u] = ldg4(z + ii * 4); }
            }
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const long long ii = i + u * stride;
                if (ii < total4) one(ii, d[u], v[u], s, t, a, b, cc, mu);
            }
        }
        return;
    }
    for (long long i = i0; i < total4; i += stride) {
        int c = (int)(i % C4) * 4;
        float4 d = ld4(dy + i * 4);
        float4 v = ldg4(z + i * 4);
        float4 s = ldg4(scale + c), t = ldg4(shift + c), a = ldg4(cA + c), b = ldg4(cB + c), cc = ldg4(cC + c);
        float4 mu = ldg4(mean + c);
        one(i, d, v, s, t, a, b, cc, mu);
    }
}

}  // namespace hgk

using namespace hgk;

extern "C" int hgk_bn_finalize(const double* sum, const double* sq, long long count, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               float* scale, float* shift, float* save_mean, float* save_invstd, int C, void* stream) {
    HGK_REQUIRE(sum && sq && gamma && beta && scale && shift && save_mean && save_invstd, "hgk_bn_finalize: null pointer");
    HGK_REQUIRE(C > 0 && count > 0, "hgk_bn_finalize: C and count must be positive");
    HGK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "hgk_bn_finalize: running stats must both be set");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sq, (double)count, gamma, beta, eps, momentum,
                                                                           running_mean, running_var, scale, shift,
                                                                           save_mean, save_invstd, C);
    HGK_CHECK_LAUNCH("hgk_bn_finalize");
    return HGK_OK;
}

extern "C" int hgk_bn_eval_prepare(const float* gamma, const float* beta, const float* running_mean,
                                   const float* running_var, float eps, float* scale, float* shift, float* save_mean,
                                   float* save_invstd, int C, void* stream) {
    HGK_REQUIRE(gamma && beta && running_mean && running_var && scale && shift && save_mean && save_invstd,
                "hgk_bn_eval_prepare: null pointer");
    HGK_REQUIRE(C > 0, "hgk_bn_eval_prepare: C must be positive");
    bn_eval_prepare_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps,
                                                                               scale, shift, save_mean, save_invstd, C);
    HGK_CHECK_LAUNCH("hgk_bn_eval_prepare");
    return HGK_OK;
}

static int bn_bwd_reduce_impl(const float* dy, const float* z, const float* scale, const float* shift, int relu,
                              const float* mean, const float* invstd, long long P, int C, double* sum_g,
                              double* sum_gx, const BnBwdFin& fin, void* stream) {
    HGK_REQUIRE(dy && z && scale && shift && mean && invstd && sum_g && sum_gx, "hgk_bn_bwd_reduce: null pointer");
    HGK_REQUIRE(P > 0 && C > 0 && C % 4 == 0, "hgk_bn_bwd_reduce: need P > 0 and C %% 4 == 0 (P=%lld C=%d)", P, C);
    int cq = C / 4;
    int cq_per_blk = cq < 64 ? cq : 64;                   // up to 256 channels per block column
    int ngrp = (cq + cq_per_blk - 1) / cq_per_blk;
    int rows_par = 256 / cq_per_blk;
    long long want_blocks = (3LL * kNumSMs + ngrp - 1) / ngrp;      // three resident CTAs per SM (85 registers)
    long long rows_per_blk = (P + want_blocks - 1) / want_blocks;
    if (rows_per_blk < 64) rows_per_blk = 64;
    long long nblk = (P + rows_per_blk - 1) / rows_per_blk;
    dim3 grid((unsigned)nblk, (unsigned)ngrp);
    size_t smem = (size_t)rows_par * cq_per_blk * 4 * 2 * sizeof(double);
    launch_pdl(bn_bwd_reduce_kernel, grid, dim3(256), smem, (cudaStream_t)stream, dy, z, scale, shift, relu, mean, invstd, P, C, sum_g,
                                                                    sum_gx, cq_per_blk, rows_par, rows_per_blk, fin);
    HGK_CHECK_LAUNCH("hgk_bn_bwd_reduce");
    return HGK_OK;
}

extern "C" int hgk_bn_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, int relu,
                                 const float* mean, const float* invstd, long long P, int C, double* sum_g,
                                 double* sum_gx, void* stream) {
    return bn_bwd_reduce_impl(dy, z, scale, shift, relu, mean, invstd, P, C, sum_g, sum_gx, BnBwdFin{}, stream);
}

extern "C" int hgk_bn_bwd_reduce_fin(const float* dy, const float* z, const float* scale, const float* shift, int relu,
                                     const float* mean, const float* invstd, long long P, int C, double* sum_g,
                                     double* sum_gx, const float* gamma, int training, float* dgamma, float* dbeta,
                                     float* cA, float* cB, float* cC, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(gamma && cA && cB && cC && ticket, "hgk_bn_bwd_reduce_fin: null pointer");
    BnBwdFin f{gamma, mean, invstd, dgamma, dbeta, cA, cB, cC, ticket, training};
    return bn_bwd_reduce_impl(dy, z, scale, shift, relu, mean, invstd, P, C, sum_g, sum_gx, f, stream);
}

extern "C" int hgk_bn_bwd_finalize(const double* sum_g, const double* sum_gx, long long count, const float* gamma,
                                   const float* mean, const float* invstd, int training, float* dgamma, float* dbeta,
                                   float* cA, float* cB, float* cC, int C, void* stream) {
    HGK_REQUIRE(sum_g && sum_gx && gamma && mean && invstd && cA && cB && cC, "hgk_bn_bwd_finalize: null pointer");
    HGK_REQUIRE(C > 0 && count > 0, "hgk_bn_bwd_finalize: C and count must be positive");
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum_g, sum_gx, (double)count, gamma, mean,
                                                                               invstd, training, dgamma, dbeta, cA, cB, cC, C);
    HGK_CHECK_LAUNCH("hgk_bn_bwd_finalize");
    return HGK_OK;
}

extern "C" int hgk_bn_bwd_apply(float* dy, const float* z, const float* scale, const float* shift, int relu,
                                const float* mean, const float* cA, const float* cB, const float* cC, long long P, int C,
                                void* stream) {
    HGK_REQUIRE(dy && z && scale && shift && mean && cA && cB && cC, "hgk_bn_bwd_apply: null pointer");
    HGK_REQUIRE(P > 0 && C > 0 && C % 4 == 0, "hgk_bn_bwd_apply: need P > 0 and C %% 4 == 0");
    long long total4 = P * (C / 4);
    long long blocks = (total4 + 255) / 256;
    if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
    launch_pdl(bn_bwd_apply_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, dy, z, scale, shift, relu, mean, cA, cB, cC, total4, C / 4);
    HGK_CHECK_LAUNCH("hgk_bn_bwd_apply");
    return HGK_OK;
}
