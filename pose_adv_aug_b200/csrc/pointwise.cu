// Bandwidth-bound ladder ops of _Hourglass / ASN (reference models/asn_stacked_hg.py:69-70,
// 140-157,192-203,407-417,431-435): 2x2 max-pool of a virtual activation (BN+ReLU applied on
// load), nearest 2x up-sample + skip add, their backward passes, NCHW<->NHWC boundary
// transposes, AvgPool2d and the tiny nn.Linear heads.  All are 128-bit coalesced NHWC streams.
#include "common.cuh"
#include "conv_args.cuh"
#include "bn_fin.cuh"

namespace hgk {

static inline unsigned stream_blocks(long long work_items) {
    long long b = (work_items + 255) / 256;
    long long cap = 16LL * kNumSMs;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

template <typename idx_t>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(Act x, int N, int H, int W, int C4, float* __restrict__ y) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    const int OH = H / 2, OW = W / 2;
    const idx_t total = (idx_t)N * OH * OW * C4;
    const int C = C4 * 4;
    for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (idx_t)gridDim.x * blockDim.x) {
        int cq = (int)(i % (idx_t)C4);
        idx_t q = i / (idx_t)C4;
        int ow = (int)(q % (idx_t)OW);
        idx_t r = q / (idx_t)OW;
        int oh = (int)(r % (idx_t)OH);
        int n = (int)(r / (idx_t)OH);
        float4 s, t;
        load_affine4(x.scale, x.shift, cq * 4, s, t);
        const float* base = x.z + (((idx_t)n * H + oh * 2) * W + ow * 2) * C + cq * 4;
        float4 v00 = ldg4(base), v01 = ldg4(base + C), v10 = ldg4(base + (idx_t)W * C), v11 = ldg4(base + (idx_t)W * C + C);
        if (x.scale != nullptr) {
            v00 = act4(v00, s, t, x.relu); v01 = act4(v01, s, t, x.relu);
            v10 = act4(v10, s, t, x.relu); v11 = act4(v11, s, t, x.relu);
        }
        float4 o;
        o.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x));
        o.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
        o.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z));
        o.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
        st4(y + i * 4, o);
    }
}

// BatchNorm-backward reduction fused into the kernel that writes the LAST contribution to dL/d relu(bn(z)) of a tensor with
// several consumers (the hourglass level inputs: skip branch + pool; the up-path sums): sum g and sum g*xhat with
// g = dy*[z*scale+shift > 0], xhat = (z - mean)*invstd -- what bn_bwd_reduce_kernel (bn.cu) computes in a pass of its own over
// both tensors (22 launches, 0.8 ms per step before this).  z == nullptr disables it.
struct BnRed {
    const float* z;          // pre-BN tensor the gradient belongs to (same shape as the gradient)
    const float* scale;
    const float* shift;
    const float* mean;
    const float* invstd;
    int relu;
    double* sum_g;
    double* sum_gx;
    BnBwdFin fin;            // last-CTA finaliser (ticket == nullptr: none)
    long long P;             // pixels of the tensor (statistics count)
};

// per-thread partial sums (channel quad fixed per thread) -> shared memory -> fp64 atomics -> last-CTA finaliser.
// Every thread of the CTA must call it.  blockDim.x == 256, 256 % C4 == 0.
__device__ __forceinline__ void bn_red_flush(const BnRed& br, int C4, const float (&s1)[4], const float (&s2)[4], double (&d1)[4],
                                             double (&d2)[4], bool last) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { d1[j] += (double)s1[j]; d2[j] += (double)s2[j]; }
    if (!last) return;
    __shared__ double red[256][8];
    const int tid = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[tid][j] = d1[j]; red[tid][4 + j] = d2[j]; }
    __syncthreads();
    const int C = C4 * 4;
    if (tid < C) {                                  // channel tid: quad tid / 4 lives in threads (tid / 4) + k * C4
        const int q = tid >> 2, e = tid & 3;
        double a = 0.0, b = 0.0;
        for (int r = q; r < 256; r += C4) { a += red[r][e]; b += red[r][4 + e]; }
        atomicAdd(br.sum_g + tid, a);
        atomicAdd(br.sum_gx + tid, b);
    }
    if (br.fin.ticket != nullptr) {
        if (last_cta_arrives(br.fin.ticket, gridDim.x)) bn_bwd_finalize_cta(br.fin, br.sum_g, br.sum_gx, (double)br.P, C);
    }
}

// first-maximum-in-scan-order wins (torch max_pool2d: `val > maxval`), so ties route like the reference
__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
    int k = 0;
    float m = a;
    if (b > m) { m = b; k = 1; }
    if (c > m) { m = c; k = 2; }
    if (d > m) { m = d; k = 3; }
    return k;
}

template <typename idx_t, bool RED>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(Act x, int N, int H, int W, int C4,
                                                           const float* __restrict__ dy, float* __restrict__ dx, int accumulate,
                                                           const BnRed br) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    const int OH = H / 2, OW = W / 2;
    const idx_t total = (idx_t)N * OH * OW * C4;
    const int C = C4 * 4;
    // RED: the grid stride is a multiple of C4 (host), so a thread stays on ONE channel quad
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu;
    if (RED) {
        const int cq0 = (int)(((idx_t)blockIdx.x * blockDim.x + threadIdx.x) % (idx_t)C4);
        mu = ldg4(br.mean + cq0 * 4);
        is = ldg4(br.invstd + cq0 * 4);
    }
    int trips = 0;
    for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (idx_t)gridDim.x * blockDim.x) {
        int cq = (int)(i % (idx_t)C4);
        idx_t q = i / (idx_t)C4;
        int ow = (int)(q % (idx_t)OW);
        idx_t r = q / (idx_t)OW;
        int oh = (int)(r % (idx_t)OH);
        int n = (int)(r / (idx_t)OH);
        float4 s, t;
        load_affine4(x.scale, x.shift, cq * 4, s, t);
        const idx_t off = (((idx_t)n * H + oh * 2) * W + ow * 2) * C + cq * 4;
        const idx_t o01 = C, o10 = (idx_t)W * C, o11 = (idx_t)W * C + C;
        const float4 z00 = ldg4(x.z + off), z01 = ldg4(x.z + off + o01), z10 = ldg4(x.z + off + o10), z11 = ldg4(x.z + off + o11);
        float4 v00 = z00, v01 = z01, v10 = z10, v11 = z11;
        if (x.scale != nullptr) {
            v00 = act4(v00, s, t, x.relu); v01 = act4(v01, s, t, x.relu);
            v10 = act4(v10, s, t, x.relu); v11 = act4(v11, s, t, x.relu);
        }
        float4 g = ldg4(dy + i * 4);
        int kx = argmax4(v00.x, v01.x, v10.x, v11.x);
        int ky = argmax4(v00.y, v01.y, v10.y, v11.y);
        int kz = argmax4(v00.z, v01.z, v10.z, v11.z);
        int kw = argmax4(v00.w, v01.w, v10.w, v11.w);
        float4 d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            d[k] = make_float4(kx == k ? g.x : 0.f, ky == k ? g.y : 0.f, kz == k ? g.z : 0.f, kw == k ? g.w : 0.f);
        const idx_t offs[4] = {0, o01, o10, o11};
        const float4 zz[4] = {z00, z01, z10, z11};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float* p = dx + off + offs[k];
            float4 o = d[k];
            if (accumulate) {
                float4 old = ld4(p);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            st4(p, o);
            if (RED) {            // o is the COMPLETE gradient w.r.t. relu(bn(z)) at this pixel
                const float4 zv = zz[k];
                const float gx = (br.relu && fmaf(zv.x, s.x, t.x) <= 0.f) ? 0.f : o.x;
                const float gy = (br.relu && fmaf(zv.y, s.y, t.y) <= 0.f) ? 0.f : o.y;
                const float gz = (br.relu && fmaf(zv.z, s.z, t.z) <= 0.f) ? 0.f : o.z;
                const float gw = (br.relu && fmaf(zv.w, s.w, t.w) <= 0.f) ? 0.f : o.w;
                s1[0] += gx; s2[0] = fmaf(gx, (zv.x - mu.x) * is.x, s2[0]);
                s1[1] += gy; s2[1] = fmaf(gy, (zv.y - mu.y) * is.y, s2[1]);
                s1[2] += gz; s2[2] = fmaf(gz, (zv.z - mu.z) * is.z, s2[2]);
                s1[3] += gw; s2[3] = fmaf(gw, (zv.w - mu.w) * is.w, s2[3]);
            }
        }
        if (RED && ++trips == 4) {            // 16 values per fp32 partial, fp64 from there on
            bn_red_flush(br, C4, s1, s2, d1, d2, false);
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
            trips = 0;
        }
    }
    if (RED) bn_red_flush(br, C4, s1, s2, d1, d2, true);
}

template <typename idx_t>
__global__ void __launch_bounds__(256) add_fwd_kernel(Act a, int a_up, Act b, int N, int H, int W, int C4, float* __restrict__ y) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    const idx_t total = (idx_t)N * H * W * C4;
    const int C = C4 * 4;
    for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (idx_t)gridDim.x * blockDim.x) {
        int cq = (int)(i % (idx_t)C4);
        idx_t q = i / (idx_t)C4;
        float4 sa, ta, sb, tb;
        load_affine4(a.scale, a.shift, cq * 4, sa, ta);
        load_affine4(b.scale, b.shift, cq * 4, sb, tb);
        idx_t ai = i * 4;
        if (a_up) {
            int w = (int)(q % (idx_t)W);
            idx_t r = q / (idx_t)W;
            int h = (int)(r % (idx_t)H);
            int n = (int)(r / (idx_t)H);
            ai = ((((idx_t)n * (H / 2) + h / 2) * (W / 2) + w / 2) * C) + cq * 4;
        }
        float4 va = ldg4(a.z + ai);
        float4 vb = ldg4(b.z + i * 4);
        if (a.scale != nullptr) va = act4(va, sa, ta, a.relu);
        if (b.scale != nullptr) vb = act4(vb, sb, tb, b.relu);
        st4(y + i * 4, make_float4(va.x + vb.x, va.y + vb.y, va.z + vb.z, va.w + vb.w));
    }
}

template <typename idx_t, bool RED>
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const float* __restrict__ dy, int N, int H, int W, int C4,
                                                            float* __restrict__ da, int accumulate, const BnRed br) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    const int OH = H / 2, OW = W / 2;
    const idx_t total = (idx_t)N * OH * OW * C4;
    const int C = C4 * 4;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu, bs = mu, bt = mu;
    if (RED) {                // the grid stride is a multiple of C4 (host): a thread stays on one channel quad
        const int cq0 = (int)(((idx_t)blockIdx.x * blockDim.x + threadIdx.x) % (idx_t)C4);
        mu = ldg4(br.mean + cq0 * 4); is = ldg4(br.invstd + cq0 * 4);
        bs = ldg4(br.scale + cq0 * 4); bt = ldg4(br.shift + cq0 * 4);
    }
    int trips = 0;
    for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (idx_t)gridDim.x * blockDim.x) {
        int cq = (int)(i % (idx_t)C4);
        idx_t q = i / (idx_t)C4;
        int ow = (int)(q % (idx_t)OW);
        idx_t r = q / (idx_t)OW;
        int oh = (int)(r % (idx_t)OH);
        int n = (int)(r / (idx_t)OH);
        const float* base = dy + (((idx_t)n * H + oh * 2) * W + ow * 2) * C + cq * 4;
        float4 a = ldg4(base), b = ldg4(base + C), c = ldg4(base + (idx_t)W * C), d = ldg4(base + (idx_t)W * C + C);
        float4 o = make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w);
        if (accumulate) {
            float4 old = ld4(da + i * 4);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        st4(da + i * 4, o);
        if (RED) {                // o is the COMPLETE gradient w.r.t. relu(bn(z)) at this pixel
            const float4 zv = ldg4(br.z + i * 4);
            const float gx = (br.relu && fmaf(zv.x, bs.x, bt.x) <= 0.f) ? 0.f : o.x;
            const float gy = (br.relu && fmaf(zv.y, bs.y, bt.y) <= 0.f) ? 0.f : o.y;
            const float gz = (br.relu && fmaf(zv.z, bs.z, bt.z) <= 0.f) ? 0.f : o.z;
            const float gw = (br.relu && fmaf(zv.w, bs.w, bt.w) <= 0.f) ? 0.f : o.w;
            s1[0] += gx; s2[0] = fmaf(gx, (zv.x - mu.x) * is.x, s2[0]);
            s1[1] += gy; s2[1] = fmaf(gy, (zv.y - mu.y) * is.y, s2[1]);
            s1[2] += gz; s2[2] = fmaf(gz, (zv.z - mu.z) * is.z, s2[2]);
            s1[3] += gw; s2[3] = fmaf(gw, (zv.w - mu.w) * is.w, s2[3]);
            if (++trips == 16) {
                bn_red_flush(br, C4, s1, s2, d1, d2, false);
#pragma unroll
                for (int j = 0; j < 4; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
                trips = 0;
            }
        }
    }
    if (RED) bn_red_flush(br, C4, s1, s2, d1, d2, true);
}

__global__ void __launch_bounds__(256) add_into_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4,
                                                       long long n, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = ldg4(src + i * 4);
        if (accumulate) {
            float4 o = ld4(dst + i * 4);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        st4(dst + i * 4, v);
    }
    if (blockIdx.x == 0) {
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) dst[i] = (accumulate ? dst[i] : 0.f) + src[i];
    }
}

// [rows][cols] -> [cols][rows] per image, 32x32 smem tiles; optional affine(+relu) indexed by the
// NHWC channel (which is `col` for nhwc->nchw)
__global__ void transpose_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                 int relu, int rows, int cols, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const long long img = (long long)blockIdx.z * rows * cols;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        int r = r0 + j, c = c0 + threadIdx.x;
        float v = 0.f;
        if (r < rows && c < cols) {
            v = __ldg(x + img + (long long)r * cols + c);
            if (scale != nullptr) v = act1(v, __ldg(scale + c), __ldg(shift + c), relu);
        }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) y[img + (long long)c * rows + r] = tile[threadIdx.x][j];
    }
}

__global__ void __launch_bounds__(256) avgpool_fwd_kernel(Act x, int N, int H, int W, int C4, int k, float* __restrict__ y) {
    const int OH = H / k, OW = W / k;
    const long long total = (long long)N * OH * OW * C4;
    const int C = C4 * 4;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int cq = (int)(i % C4);
        long long q = i / C4;
        int ow = (int)(q % OW);
        long long r = q / OW;
        int oh = (int)(r % OH);
        int n = (int)(r / OH);
        float4 s, t;
        load_affine4(x.scale, x.shift, cq * 4, s, t);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
                float4 v = ldg4(x.z + (((long long)n * H + oh * k + dy) * W + ow * k + dx) * C + cq * 4);
                if (x.scale != nullptr) v = act4(v, s, t, x.relu);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        st4(y + i * 4, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
    }
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dy, int N, int H, int W, int C4, int k,
                                                          float* __restrict__ dx, int accumulate) {
    const int OH = H / k, OW = W / k;
    const long long total = (long long)N * H * W * C4;
    const int C = C4 * 4;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int cq = (int)(i % C4);
        long long q = i / C4;
        int w = (int)(q % W);
        long long r = q / W;
        int h = (int)(r % H);
        int n = (int)(r / H);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h / k < OH && w / k < OW) {
            float4 g = ldg4(dy + (((long long)n * OH + h / k) * OW + w / k) * C + cq * 4);
            o = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
        }
        if (accumulate) {
            float4 old = ld4(dx + i * 4);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        st4(dx + i * 4, o);
    }
}

// y[m][n] = b[n] + sum_k x[m][k] w[n][k]; one warp per output element
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int M,
                                  int K, int Nout, float* __restrict__ y) {
    int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= M * Nout) return;
    int m = wid / Nout, n = wid - m * Nout;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(x + (size_t)m * K + k), __ldg(w + (size_t)n * K + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) y[wid] = acc + (b ? __ldg(b + n) : 0.f);
}

// dx[m][k] = sum_n dy[m][n] w[n][k];  dw[n][k] += sum_m dy[m][n] x[m][k];  db[n] += sum_m dy[m][n]
__global__ void linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy, int M,
                                  int K, int Nout, float* dx, float* dw, float* db) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (dx != nullptr && i < M * K) {
        int m = i / K, k = i - m * K;
        float acc = 0.f;
        for (int n = 0; n < Nout; ++n) acc = fmaf(__ldg(dy + m * Nout + n), __ldg(w + (size_t)n * K + k), acc);
        dx[i] = acc;
    }
    if (dw != nullptr && i < Nout * K) {
        int n = i / K, k = i - n * K;
        float acc = 0.f;
        for (int m = 0; m < M; ++m) acc = fmaf(__ldg(dy + m * Nout + n), __ldg(x + (size_t)m * K + k), acc);
        dw[i] += acc;
    }
    if (db != nullptr && i < Nout) {
        float acc = 0.f;
        for (int m = 0; m < M; ++m) acc += __ldg(dy + m * Nout + i);
        db[i] += acc;
    }
}

}  // namespace hgk

using namespace hgk;

// 32-bit index arithmetic whenever the largest element index fits (every layer of the hourglass): the four 64-bit divisions
// per float4 item made these streams instruction-bound (add_fwd at 64x64x256: 57 us for 225 MB, 60 % of the copy rate)
#define HGK_SMALL_IDX() ((long long)N * H * W * C < (1LL << 31))
#define HGK_NHWC_CHECK(name)                                                                      \
    HGK_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, name ": need positive dims and C %% 4 == 0 (C=%d)", C)

extern "C" int hgk_maxpool2_fwd(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W,
                                int C, float* y, void* stream) {
    HGK_REQUIRE(x && y, "hgk_maxpool2_fwd: null pointer");
    HGK_NHWC_CHECK("hgk_maxpool2_fwd");
    HGK_REQUIRE(H % 2 == 0 && W % 2 == 0, "hgk_maxpool2_fwd: H and W must be even (H=%d W=%d)", H, W);
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    if (HGK_SMALL_IDX())
        launch_pdl(maxpool2_fwd_kernel<unsigned>, dim3(stream_blocks(total)), dim3(256), 0, (cudaStream_t)stream, Act{x, x_scale, x_shift, x_relu}, N, H, W, C / 4, y);
    else
        launch_pdl(maxpool2_fwd_kernel<long long>, dim3(stream_blocks(total)), dim3(256), 0, (cudaStream_t)stream, Act{x, x_scale, x_shift, x_relu}, N, H, W, C / 4, y);
    HGK_CHECK_LAUNCH("hgk_maxpool2_fwd");
    return HGK_OK;
}

// grid of a kernel with the fused BatchNorm-backward reduction: the stride (gridDim * 256) must be a multiple of C / 4 so that a
// thread keeps its channel quad; 256 % (C / 4) == 0 is required by the caller
static unsigned red_blocks(long long total) { return stream_blocks(total); }

static int maxpool2_bwd_impl(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W,
                             int C, const float* dy, float* dx, int accumulate, const BnRed& br, void* stream) {
    HGK_REQUIRE(x && dy && dx, "hgk_maxpool2_bwd: null pointer");
    HGK_NHWC_CHECK("hgk_maxpool2_bwd");
    HGK_REQUIRE(H % 2 == 0 && W % 2 == 0, "hgk_maxpool2_bwd: H and W must be even (H=%d W=%d)", H, W);
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    const dim3 grid(red_blocks(total)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    const Act ax{x, x_scale, x_shift, x_relu};
    if (br.z != nullptr) {
        HGK_REQUIRE(C <= 1024 && 256 % (C / 4) == 0, "hgk_maxpool2_bwd_bnred: C / 4 must divide 256 (C=%d)", C);
        HGK_REQUIRE(br.z == x && x_scale != nullptr, "hgk_maxpool2_bwd_bnred: the reduction is over the pooled tensor's own BatchNorm");
        if (HGK_SMALL_IDX()) launch_pdl(maxpool2_bwd_kernel<unsigned, true>, grid, block, 0, st, ax, N, H, W, C / 4, dy, dx, accumulate, br);
        else launch_pdl(maxpool2_bwd_kernel<long long, true>, grid, block, 0, st, ax, N, H, W, C / 4, dy, dx, accumulate, br);
    } else {
        if (HGK_SMALL_IDX()) launch_pdl(maxpool2_bwd_kernel<unsigned, false>, grid, block, 0, st, ax, N, H, W, C / 4, dy, dx, accumulate, br);
        else launch_pdl(maxpool2_bwd_kernel<long long, false>, grid, block, 0, st, ax, N, H, W, C / 4, dy, dx, accumulate, br);
    }
    HGK_CHECK_LAUNCH("hgk_maxpool2_bwd");
    return HGK_OK;
}

extern "C" int hgk_maxpool2_bwd(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W,
                                int C, const float* dy, float* dx, int accumulate, void* stream) {
    return maxpool2_bwd_impl(x, x_scale, x_shift, x_relu, N, H, W, C, dy, dx, accumulate, BnRed{}, stream);
}

extern "C" int hgk_maxpool2_bwd_bnred(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W,
                                      int C, const float* dy, float* dx, int accumulate, const float* mean, const float* invstd,
                                      double* sum_g, double* sum_gx, const float* gamma, int training, float* dgamma, float* dbeta,
                                      float* cA, float* cB, float* cC, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(mean && invstd && sum_g && sum_gx && gamma && cA && cB && cC && ticket, "hgk_maxpool2_bwd_bnred: null pointer");
    BnRed br{x, x_scale, x_shift, mean, invstd, x_relu, sum_g, sum_gx,
             BnBwdFin{gamma, mean, invstd, dgamma, dbeta, cA, cB, cC, ticket, training}, (long long)N * H * W};
    return maxpool2_bwd_impl(x, x_scale, x_shift, x_relu, N, H, W, C, dy, dx, accumulate, br, stream);
}

extern "C" int hgk_add_fwd(const float* a, const float* a_scale, const float* a_shift, int a_relu, int a_up, const float* b,
                           const float* b_scale, const float* b_shift, int b_relu, int N, int H, int W, int C, float* y,
                           void* stream) {
    HGK_REQUIRE(a && b && y, "hgk_add_fwd: null pointer");
    HGK_NHWC_CHECK("hgk_add_fwd");
    HGK_REQUIRE(!a_up || (H % 2 == 0 && W % 2 == 0), "hgk_add_fwd: up-sampled output must have even H, W");
    long long total = (long long)N * H * W * (C / 4);
    if (HGK_SMALL_IDX())
        launch_pdl(add_fwd_kernel<unsigned>, dim3(stream_blocks(total)), dim3(256), 0, (cudaStream_t)stream, Act{a, a_scale, a_shift, a_relu}, a_up,
                                                                           Act{b, b_scale, b_shift, b_relu}, N, H, W, C / 4, y);
    else
        launch_pdl(add_fwd_kernel<long long>, dim3(stream_blocks(total)), dim3(256), 0, (cudaStream_t)stream, Act{a, a_scale, a_shift, a_relu}, a_up,
                                                                           Act{b, b_scale, b_shift, b_relu}, N, H, W, C / 4, y);
    HGK_CHECK_LAUNCH("hgk_add_fwd");
    return HGK_OK;
}

static int upsample2_bwd_impl(const float* dy, int N, int H, int W, int C, float* da, int accumulate, const BnRed& br, void* stream) {
    HGK_REQUIRE(dy && da, "hgk_upsample2_bwd: null pointer");
    HGK_NHWC_CHECK("hgk_upsample2_bwd");
    HGK_REQUIRE(H % 2 == 0 && W % 2 == 0, "hgk_upsample2_bwd: H and W must be even");
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    const dim3 grid(red_blocks(total)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (br.z != nullptr) {
        HGK_REQUIRE(C <= 1024 && 256 % (C / 4) == 0, "hgk_upsample2_bwd_bnred: C / 4 must divide 256 (C=%d)", C);
        if (HGK_SMALL_IDX()) launch_pdl(upsample2_bwd_kernel<unsigned, true>, grid, block, 0, st, dy, N, H, W, C / 4, da, accumulate, br);
        else launch_pdl(upsample2_bwd_kernel<long long, true>, grid, block, 0, st, dy, N, H, W, C / 4, da, accumulate, br);
    } else {
        if (HGK_SMALL_IDX()) launch_pdl(upsample2_bwd_kernel<unsigned, false>, grid, block, 0, st, dy, N, H, W, C / 4, da, accumulate, br);
        else launch_pdl(upsample2_bwd_kernel<long long, false>, grid, block, 0, st, dy, N, H, W, C / 4, da, accumulate, br);
    }
    HGK_CHECK_LAUNCH("hgk_upsample2_bwd");
    return HGK_OK;
}

extern "C" int hgk_upsample2_bwd(const float* dy, int N, int H, int W, int C, float* da, int accumulate, void* stream) {
    return upsample2_bwd_impl(dy, N, H, W, C, da, accumulate, BnRed{}, stream);
}

/* da receives the LAST contribution to dL/d relu(bn(z)) of its tensor (z: [N,H/2,W/2,C]): the BatchNorm-backward reduction and
 * its finaliser ride on this launch */
extern "C" int hgk_upsample2_bwd_bnred(const float* dy, int N, int H, int W, int C, float* da, int accumulate, const float* z,
                                       const float* scale, const float* shift, int relu, const float* mean, const float* invstd,
                                       double* sum_g, double* sum_gx, const float* gamma, int training, float* dgamma, float* dbeta,
                                       float* cA, float* cB, float* cC, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(z && scale && shift && mean && invstd && sum_g && sum_gx && gamma && cA && cB && cC && ticket,
                "hgk_upsample2_bwd_bnred: null pointer");
    BnRed br{z, scale, shift, mean, invstd, relu, sum_g, sum_gx,
             BnBwdFin{gamma, mean, invstd, dgamma, dbeta, cA, cB, cC, ticket, training}, (long long)N * (H / 2) * (W / 2)};
    return upsample2_bwd_impl(dy, N, H, W, C, da, accumulate, br, stream);
}

extern "C" int hgk_add_into(const float* src, float* dst, long long n, int accumulate, void* stream) {
    HGK_REQUIRE(src && dst && n > 0, "hgk_add_into: bad arguments");
    HGK_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "hgk_add_into: pointers must be 16-byte aligned");
    add_into_kernel<<<stream_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(src, dst, n / 4, n, accumulate);
    HGK_CHECK_LAUNCH("hgk_add_into");
    return HGK_OK;
}

extern "C" int hgk_nchw_to_nhwc(const float* x, int N, int C, int H, int W, float* y, void* stream) {
    HGK_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "hgk_nchw_to_nhwc: bad arguments");
    HGK_REQUIRE(N <= 65535, "hgk_nchw_to_nhwc: batch too large");
    int rows = C, cols = H * W;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, N), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, 0, rows, cols, y);
    HGK_CHECK_LAUNCH("hgk_nchw_to_nhwc");
    return HGK_OK;
}

extern "C" int hgk_nhwc_to_nchw(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W,
                                int C, float* y, void* stream) {
    HGK_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "hgk_nhwc_to_nchw: bad arguments");
    HGK_REQUIRE(N <= 65535, "hgk_nhwc_to_nchw: batch too large");
    HGK_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), "hgk_nhwc_to_nchw: scale/shift must both be set");
    int rows = H * W, cols = C;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, N), block(32, 8);
    HGK_REQUIRE(grid.y <= 65535, "hgk_nhwc_to_nchw: image too large");
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, x_scale, x_shift, x_relu, rows, cols, y);
    HGK_CHECK_LAUNCH("hgk_nhwc_to_nchw");
    return HGK_OK;
}

extern "C" int hgk_avgpool_fwd(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W,
                               int C, int k, float* y, void* stream) {
    HGK_REQUIRE(x && y, "hgk_avgpool_fwd: null pointer");
    HGK_NHWC_CHECK("hgk_avgpool_fwd");
    HGK_REQUIRE(k > 0 && H >= k && W >= k, "hgk_avgpool_fwd: kernel %d larger than input %dx%d", k, H, W);
    long long total = (long long)N * (H / k) * (W / k) * (C / 4);
    avgpool_fwd_kernel<<<stream_blocks(total), 256, 0, (cudaStream_t)stream>>>(Act{x, x_scale, x_shift, x_relu}, N, H, W, C / 4, k, y);
    HGK_CHECK_LAUNCH("hgk_avgpool_fwd");
    return HGK_OK;
}

extern "C" int hgk_avgpool_bwd(const float* dy, int N, int H, int W, int C, int k, float* dx, int accumulate, void* stream) {
    HGK_REQUIRE(dy && dx, "hgk_avgpool_bwd: null pointer");
    HGK_NHWC_CHECK("hgk_avgpool_bwd");
    HGK_REQUIRE(k > 0 && H >= k && W >= k, "hgk_avgpool_bwd: kernel %d larger than input %dx%d", k, H, W);
    long long total = (long long)N * H * W * (C / 4);
    avgpool_bwd_kernel<<<stream_blocks(total), 256, 0, (cudaStream_t)stream>>>(dy, N, H, W, C / 4, k, dx, accumulate);
    HGK_CHECK_LAUNCH("hgk_avgpool_bwd");
    return HGK_OK;
}

extern "C" int hgk_linear_fwd(const float* x, const float* w, const float* b, int M, int K, int Nout, float* y, void* stream) {
    HGK_REQUIRE(x && w && y && M > 0 && K > 0 && Nout > 0, "hgk_linear_fwd: bad arguments");
    long long threads = (long long)M * Nout * 32;
    linear_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, w, b, M, K, Nout, y);
    HGK_CHECK_LAUNCH("hgk_linear_fwd");
    return HGK_OK;
}

extern "C" int hgk_linear_bwd(const float* x, const float* w, const float* dy, int M, int K, int Nout, float* dx, float* dw,
                              float* db, void* stream) {
    HGK_REQUIRE(x && w && dy && M > 0 && K > 0 && Nout > 0, "hgk_linear_bwd: bad arguments");
    long long n = (long long)(M > Nout ? M : Nout) * K;
    linear_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, w, dy, M, K, Nout, dx, dw, db);
    HGK_CHECK_LAUNCH("hgk_linear_bwd");
    return HGK_OK;
}
