// Image crop / rotate / resize augmentation on the GPU (warp half): the pixel work of
// pylib/HumanAug.py:117-175 `crop` -- the function `load_batch_data` (joint-train-pose-s-r-agent.py:425-450) reaches through a
// DataLoader for every batch of agent-sampled (scale, rotation) pairs -- from images that stay resident in HBM.
//
// `crop` is byte work on top of scipy.misc / PIL, and parity here means the same BYTES:
//   * scipy.misc.toimage's `bytescale`: a float array is stretched from its own [min, max] to 0..255, + 0.5, truncated
//     (float64 arithmetic for the zero-padded crop window, float32 for a whole float32 image);
//   * PIL `Image.resize(BILINEAR)`: two passes (horizontal, then vertical), each a convolution with a triangle filter
//     widened by the down-scale factor, taps normalised in double and rounded to 22-bit fixed point, accumulated in int32
//     from 1 << 21, shifted, clipped to uint8 between the passes;
//   * PIL `Image.rotate(BILINEAR)`: inverse affine map of the pixel centre in double, the four neighbours blended as
//     a + (b - a) * d in double (edge-clamped, previous row re-used below the last one), truncated to uint8, zero outside.
// Every double / float expression below is written with explicit round-to-nearest intrinsics: the CPU library evaluates
// them without fused multiply-adds, and one contracted FMA can flip a truncation.
// All images are H x W x 3 interleaved (RGB), uint8 unless said otherwise.
#include "common.cuh"

namespace hgk {

// ---------------------------------------------------------------------------------------------------------------------
// min / max of the region [y0,y1) x [x0,x1) of an H x W x 3 image (float32 or uint8), optionally together with 0 (the crop
// window has zero padding).  out2 = {min, max} as doubles (exact for both input types).  Row-strip CTAs fold their partial
// results into two order-preserving integer keys with atomics; the last CTA to finish (ticket) decodes them, writes the
// doubles and resets the scratch {min key, max key, ticket} to its idle state {0xFFFFFFFF, 0, 0} for the next call.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int f32_key(float v) {
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_f32(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <typename T>
__global__ void __launch_bounds__(256) aug_minmax_kernel(const T* __restrict__ img, int W, int y0, int y1, int x0, int x1,
                                                           int include_zero, unsigned int* __restrict__ scratch,
                                                           double* __restrict__ out2) {
    const int rw = (x1 - x0) * 3;
    float lo = INFINITY, hi = -INFINITY;
    for (int r = y0 + blockIdx.x; r < y1; r += gridDim.x) {
        const T* row = img + ((size_t)r * W + x0) * 3;
        for (int c = threadIdx.x; c < rw; c += blockDim.x) {
            const float v = (float)row[c];
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float slo[8], shi[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
        if (lo <= hi) {                                   // this CTA saw at least one element
            atomicMin(scratch + 0, f32_key(lo));
            atomicMax(scratch + 1, f32_key(hi));
        }
        __threadfence();
        last = atomicAdd(scratch + 2, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const unsigned int klo = atomicExch(scratch + 0, 0xFFFFFFFFu), khi = atomicExch(scratch + 1, 0u);
        scratch[2] = 0u;
        const bool any = klo <= khi;
        float flo = any ? key_f32(klo) : 0.f, fhi = any ? key_f32(khi) : 0.f;
        if (include_zero) { flo = fminf(flo, 0.f); fhi = fmaxf(fhi, 0.f); }
        out2[0] = (double)flo;
        out2[1] = (double)fhi;
    }
}

// bytescale, float64 arithmetic (numpy: ((data - cmin) * scale + 0).clip(0, 255) + 0.5 -> uint8)
__device__ __forceinline__ unsigned char bytescale_f64(double v, double cmin, double scale) {
    double b = __dmul_rn(__dsub_rn(v, cmin), scale);
    b = fmin(fmax(b, 0.0), 255.0);
    return (unsigned char)(int)__dadd_rn(b, 0.5);
}

// ---------------------------------------------------------------------------------------------------------------------
// The zero-padded crop window as bytes: window pixel (y, x) = source pixel (y + oy, x + ox) inside [ny0,ny1) x [nx0,nx1),
// 0.0 elsewhere; byte-scaled in float64 over minmax (which the caller reduced over the same region, zero included when the
// window has padding).  out: Hn x Wn x 3.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) aug_window_bytes_kernel(const T* __restrict__ src, int SW, int oy, int ox, int ny0, int ny1,
                                                                 int nx0, int nx1, const double* __restrict__ minmax, int Hn, int Wn,
                                                                 unsigned char* __restrict__ out) {
    const double cmin = minmax[0];
    double cscale = __dsub_rn(minmax[1], cmin);
    if (cscale == 0.0) cscale = 1.0;
    const double scale = __ddiv_rn(255.0, cscale);
    const int total = Hn * Wn * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % 3, p = i / 3;
        const int x = p % Wn, y = p / Wn;
        double v = 0.0;
        if (y >= ny0 && y < ny1 && x >= nx0 && x < nx1) v = (double)src[((size_t)(y + oy) * SW + (x + ox)) * 3 + c];
        out[i] = bytescale_f64(v, cmin, scale);
    }
}

// whole float32 image -> bytes, float32 arithmetic (NumPy 1.x: the float64 `scale` is cast down to the array's float32)
__global__ void __launch_bounds__(256) aug_image_bytes_f32_kernel(const float* __restrict__ src, long long total,
                                                                    const double* __restrict__ minmax, unsigned char* __restrict__ out) {
    const float cmin = (float)minmax[0];
    float cscale = __fsub_rn((float)minmax[1], cmin);
    if (cscale == 0.f) cscale = 1.f;
    const float scale = (float)__ddiv_rn(255.0, (double)cscale);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float b = __fmul_rn(__fsub_rn(__ldg(src + i), cmin), scale);
        b = fminf(fmaxf(b, 0.f), 255.f);
        out[i] = (unsigned char)(int)__fadd_rn(b, 0.5f);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// PIL precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter over the full box: thread xx writes
// bounds[2 xx] = first input index, bounds[2 xx + 1] = tap count, kk[xx * ksize + j] = 22-bit fixed-point taps.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void aug_resample_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)(float)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                      // bilinear: support 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dadd_rn(0.0, __dmul_rn((double)xx + 0.5, scale));
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
        int q = 0;
        if (x < xmax) {
            double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            q = w < 0.0 ? (int)__dadd_rn(-0.5, __dmul_rn(w, 4194304.0)) : (int)__dadd_rn(0.5, __dmul_rn(w, 4194304.0));
        }
        k[x] = q;
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ unsigned char clip8_22(int acc) {
    const int v = acc >> 22;
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: tmp[y][xx][c] over the in_h x in_w sub-image of `src` (row stride SW pixels) that starts at (y_off, x_off)
__global__ void __launch_bounds__(256) aug_resize_h_kernel(const unsigned char* __restrict__ src, int SW, int y_off, int x_off,
                                                            int in_h, int out_w, const int* __restrict__ bounds,
                                                            const int* __restrict__ kk, int ksize, unsigned char* __restrict__ tmp) {
    const int total = in_h * out_w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int xx = i % out_w, y = i / out_w;
        const int xmin = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
        const int* k = kk + (size_t)xx * ksize;
        const unsigned char* row = src + ((size_t)(y + y_off) * SW + x_off + xmin) * 3;
        int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
        for (int j = 0; j < n; ++j) {
            const int q = __ldg(k + j);
            s0 += (int)row[3 * j] * q;
            s1 += (int)row[3 * j + 1] * q;
            s2 += (int)row[3 * j + 2] * q;
        }
        unsigned char* o = tmp + (size_t)i * 3;
        o[0] = clip8_22(s0);
        o[1] = clip8_22(s1);
        o[2] = clip8_22(s2);
    }
}

// vertical pass: out[yy][x][c] from tmp (in_h x out_w x 3)
__global__ void __launch_bounds__(256) aug_resize_v_kernel(const unsigned char* __restrict__ tmp, int out_w, int out_h,
                                                            const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                            unsigned char* __restrict__ out) {
    const int rw = out_w * 3;
    const int total = out_h * rw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int xc = i % rw, yy = i / rw;
        const int ymin = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
        const int* k = kk + (size_t)yy * ksize;
        int s = 1 << 21;
        for (int j = 0; j < n; ++j) s += (int)tmp[(size_t)(ymin + j) * rw + xc] * __ldg(k + j);
        out[i] = clip8_22(s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// PIL ImagingGenericTransform(affine_transform, bilinear_filter32RGB), fill 0: out (H x W x 3) from in (H x W x 3).
// m = {a0..a5}: input position of output pixel centre (x + .5, y + .5) is (a0 xin + a1 yin + a2, a3 xin + a4 yin + a5).
// ---------------------------------------------------------------------------------------------------------------------
struct Affine6 { double a[6]; };

__global__ void __launch_bounds__(256) aug_rotate_kernel(const unsigned char* __restrict__ in, int H, int W, Affine6 m,
                                                          unsigned char* __restrict__ out) {
    const int total = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int xo = i % W, yo = i / W;
        const double xc = (double)xo + 0.5, yc = (double)yo + 0.5;
        double xin = __dadd_rn(__dadd_rn(__dmul_rn(m.a[0], xc), __dmul_rn(m.a[1], yc)), m.a[2]);
        double yin = __dadd_rn(__dadd_rn(__dmul_rn(m.a[3], xc), __dmul_rn(m.a[4], yc)), m.a[5]);
        unsigned char* o = out + (size_t)i * 3;
        if (xin < 0.0 || xin >= (double)W || yin < 0.0 || yin >= (double)H) {
            o[0] = 0; o[1] = 0; o[2] = 0;
            continue;
        }
        xin = __dsub_rn(xin, 0.5);
        yin = __dsub_rn(yin, 0.5);
        const int x = xin < 0.0 ? (int)floor(xin) : (int)xin;
        const int y = yin < 0.0 ? (int)floor(yin) : (int)yin;
        const double dx = __dsub_rn(xin, (double)x), dy = __dsub_rn(yin, (double)y);
        const int x0 = min(max(x, 0), W - 1), x1 = min(max(x + 1, 0), W - 1);
        const int y0 = min(max(y, 0), H - 1);
        const bool has2 = (y + 1 >= 0) && (y + 1 < H);
        const unsigned char* r0 = in + (size_t)y0 * W * 3;
        const unsigned char* r1 = in + (size_t)(has2 ? y + 1 : y0) * W * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double p00 = (double)r0[x0 * 3 + c], p01 = (double)r0[x1 * 3 + c];
            double v1 = __dadd_rn(p00, __dmul_rn(__dsub_rn(p01, p00), dx));
            double v2 = v1;
            if (has2) {
                const double p10 = (double)r1[x0 * 3 + c], p11 = (double)r1[x1 * 3 + c];
                v2 = __dadd_rn(p10, __dmul_rn(__dsub_rn(p11, p10), dx));
            }
            v1 = __dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
            o[c] = (unsigned char)(int)v1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// utils/imutils.im_to_torch on a batch of crops: [N][res][res][3] uint8 -> [N][3][res][res] float32, divided by 255 only
// when the image's maximum exceeds 1 (one CTA per image).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) aug_to_chw_float_kernel(const unsigned char* __restrict__ img, int res, float* __restrict__ out) {
    const int n = blockIdx.x;
    const int P = res * res;
    const unsigned char* a = img + (size_t)n * P * 3;
    float* o = out + (size_t)n * P * 3;
    int mx = 0;
    for (int i = threadIdx.x; i < P * 3; i += blockDim.x) mx = max(mx, (int)a[i]);
    mx = __syncthreads_or(mx > 1);
    for (int i = threadIdx.x; i < P * 3; i += blockDim.x) {
        const int c = i / P, p = i - c * P;
        const float v = (float)a[(size_t)p * 3 + c];
        o[i] = mx ? __fdiv_rn(v, 255.f) : v;
    }
}

static int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace hgk

using namespace hgk;

extern "C" int hgk_aug_minmax(const void* img, int is_u8, int H, int W, int y0, int y1, int x0, int x1, int include_zero,
                              unsigned int* scratch3, double* out2, void* stream) {
    HGK_REQUIRE(img && out2 && scratch3, "hgk_aug_minmax: null pointer");
    HGK_REQUIRE(H > 0 && W > 0 && y0 >= 0 && x0 >= 0 && y1 <= H && x1 <= W && y1 >= y0 && x1 >= x0,
                "hgk_aug_minmax: region [%d,%d) x [%d,%d) outside the %d x %d image", y0, y1, x0, x1, H, W);
    const int rows = y1 - y0;
    const int g = rows < 1 ? 1 : (rows > kNumSMs * 4 ? kNumSMs * 4 : rows);
    if (is_u8)
        aug_minmax_kernel<unsigned char><<<g, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)img, W, y0, y1, x0, x1, include_zero, scratch3, out2);
    else
        aug_minmax_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)img, W, y0, y1, x0, x1, include_zero, scratch3, out2);
    HGK_CHECK_LAUNCH("hgk_aug_minmax");
    return HGK_OK;
}

extern "C" int hgk_aug_window_bytes(const void* src, int is_u8, int SH, int SW, int oy, int ox, int ny0, int ny1, int nx0, int nx1,
                                    const double* minmax, int Hn, int Wn, unsigned char* out, void* stream) {
    HGK_REQUIRE(src && minmax && out, "hgk_aug_window_bytes: null pointer");
    HGK_REQUIRE(Hn > 0 && Wn > 0 && (long long)Hn * Wn * 3 < (1ll << 31), "hgk_aug_window_bytes: bad window %d x %d", Hn, Wn);
    HGK_REQUIRE(ny0 >= 0 && nx0 >= 0 && ny1 <= Hn && nx1 <= Wn, "hgk_aug_window_bytes: pasted region outside the window");
    HGK_REQUIRE(ny1 <= ny0 || nx1 <= nx0 || (ny0 + oy >= 0 && nx0 + ox >= 0 && ny1 + oy <= SH && nx1 + ox <= SW),
                "hgk_aug_window_bytes: pasted region outside the %d x %d source", SH, SW);
    const int g = grid_for((long long)Hn * Wn * 3, 256);
    if (is_u8)
        aug_window_bytes_kernel<unsigned char><<<g, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)src, SW, oy, ox, ny0, ny1, nx0, nx1, minmax, Hn, Wn, out);
    else
        aug_window_bytes_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)src, SW, oy, ox, ny0, ny1, nx0, nx1, minmax, Hn, Wn, out);
    HGK_CHECK_LAUNCH("hgk_aug_window_bytes");
    return HGK_OK;
}

extern "C" int hgk_aug_image_bytes_f32(const float* src, int H, int W, const double* minmax, unsigned char* out, void* stream) {
    HGK_REQUIRE(src && minmax && out && H > 0 && W > 0, "hgk_aug_image_bytes_f32: bad arguments");
    const long long total = (long long)H * W * 3;
    aug_image_bytes_f32_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, total, minmax, out);
    HGK_CHECK_LAUNCH("hgk_aug_image_bytes_f32");
    return HGK_OK;
}

extern "C" int hgk_aug_resample_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return HGK_EINVAL;
    double scale = (double)(float)in_size / (double)out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(scale) * 2 + 1;
}

extern "C" int hgk_aug_resample_coeffs(int in_size, int out_size, int ksize, int* bounds, int* kk, void* stream) {
    HGK_REQUIRE(bounds && kk, "hgk_aug_resample_coeffs: null pointer");
    HGK_REQUIRE(in_size > 0 && out_size > 0 && in_size < (1 << 24) && ksize == hgk_aug_resample_ksize(in_size, out_size),
                "hgk_aug_resample_coeffs: in %d out %d ksize %d", in_size, out_size, ksize);
    aug_resample_coeffs_kernel<<<(out_size + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in_size, out_size, ksize, bounds, kk);
    HGK_CHECK_LAUNCH("hgk_aug_resample_coeffs");
    return HGK_OK;
}

extern "C" int hgk_aug_resize_h(const unsigned char* src, int SH, int SW, int y_off, int x_off, int in_h, int in_w, int out_w,
                                const int* bounds, const int* kk, int ksize, unsigned char* tmp, void* stream) {
    HGK_REQUIRE(src && bounds && kk && tmp, "hgk_aug_resize_h: null pointer");
    HGK_REQUIRE(in_h > 0 && in_w > 0 && out_w > 0 && y_off >= 0 && x_off >= 0 && y_off + in_h <= SH && x_off + in_w <= SW,
                "hgk_aug_resize_h: sub-image (%d,%d)+%dx%d outside the %d x %d source", y_off, x_off, in_h, in_w, SH, SW);
    HGK_REQUIRE((long long)in_h * out_w * 3 < (1ll << 31), "hgk_aug_resize_h: too large");
    aug_resize_h_kernel<<<grid_for((long long)in_h * out_w, 256), 256, 0, (cudaStream_t)stream>>>(src, SW, y_off, x_off, in_h, out_w, bounds, kk, ksize, tmp);
    HGK_CHECK_LAUNCH("hgk_aug_resize_h");
    return HGK_OK;
}

extern "C" int hgk_aug_resize_v(const unsigned char* tmp, int in_h, int out_w, int out_h, const int* bounds, const int* kk, int ksize,
                                unsigned char* out, void* stream) {
    HGK_REQUIRE(tmp && bounds && kk && out, "hgk_aug_resize_v: null pointer");
    HGK_REQUIRE(in_h > 0 && out_w > 0 && out_h > 0 && (long long)out_h * out_w * 3 < (1ll << 31), "hgk_aug_resize_v: bad sizes");
    aug_resize_v_kernel<<<grid_for((long long)out_h * out_w * 3, 256), 256, 0, (cudaStream_t)stream>>>(tmp, out_w, out_h, bounds, kk, ksize, out);
    HGK_CHECK_LAUNCH("hgk_aug_resize_v");
    return HGK_OK;
}

extern "C" int hgk_aug_rotate(const unsigned char* in, int H, int W, const double* m6, unsigned char* out, void* stream) {
    HGK_REQUIRE(in && out && m6, "hgk_aug_rotate: null pointer");
    HGK_REQUIRE(H > 0 && W > 0 && (long long)H * W * 3 < (1ll << 31), "hgk_aug_rotate: bad size %d x %d", H, W);
    Affine6 m;
    for (int i = 0; i < 6; ++i) m.a[i] = m6[i];
    aug_rotate_kernel<<<grid_for((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(in, H, W, m, out);
    HGK_CHECK_LAUNCH("hgk_aug_rotate");
    return HGK_OK;
}

extern "C" int hgk_aug_to_chw_float(const unsigned char* img, int N, int res, float* out, void* stream) {
    HGK_REQUIRE(img && out && N > 0 && res > 0 && (long long)res * res * 3 < (1ll << 31), "hgk_aug_to_chw_float: bad arguments");
    aug_to_chw_float_kernel<<<N, 1024, 0, (cudaStream_t)stream>>>(img, res, out);
    HGK_CHECK_LAUNCH("hgk_aug_to_chw_float");
    return HGK_OK;
}
