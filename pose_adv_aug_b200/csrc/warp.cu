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


// =====================================================================================================================
// Batched form: the whole crop pipeline of a BATCH of images in a dozen launches.  The per-image path above costs ~10 small
// launches per image (a batch of 24 is launch-bound: 3.5 ms for ~0.3 ms of GPU work); here blockIdx.y is the image, every
// kernel reads that image's row of a descriptor table (all fields 64-bit so that the host packs one int64 row per image),
// and images that do not need a stage (no pre-shrink, no rotation) return at once.  Same device functions, same bytes.
// =====================================================================================================================
struct AugDesc {
    long long src, H, W;                                       // float32 H x W x 3 source image
    long long pre_h, pre_w;                                    // size after the pre-shrink of ref :131 (0: none)
    long long o_bytes, o_ptmp, o_pout;                         // u8 arena offsets: whole image as bytes, horizontal pass, shrunk image
    long long c_pw_b, c_pw_k, ks_pw, c_ph_b, c_ph_k, ks_ph;    // int arena offsets / tap counts of the pre-shrink tables (width, height)
    long long Hn, Wn, ny0, ny1, nx0, nx1, oy, ox, has_zero;    // crop window: size, pasted range, source offset, zero padding present
    long long o_win, rot, o_rot, pad;                          // window bytes, rotated?, rotated bytes, padding removed after rotation
    long long in_h, in_w, o_ftmp;                              // input of the final resize (window minus padding), its horizontal pass
    long long c_fw_b, c_fw_k, ks_fw, c_fh_b, c_fh_k, ks_fh;    // tables of the final resize
    long long out_index;                                       // slot in the [N][res][res][3] output
    long long chw, flip;                                       // source is 3 x H x W (the resident image as load_image returns it),
                                                               // read W-flipped, times the colour gains of its mats row, clamped to [0,1]
};
constexpr int AUG_MATS = 9;                                    // per image: six rotation coefficients + three colour gains

// source pixel (y, x, channel c) of image d: plain H x W x 3, or -- for the agent's loader (data/joint_train_s_r_agent.py:
// 160-168: fliplr, then img[c].mul_(gain).clamp_(0, 1), then im_to_numpy) -- the flipped, colour-scaled C x H x W image
__device__ __forceinline__ float aug_src_px(const AugDesc& d, const float* __restrict__ src, const double* __restrict__ m,
                                            int y, int x, int c) {
    if (!d.chw) return src[((size_t)y * d.W + x) * 3 + c];
    const int xs = d.flip ? (int)d.W - 1 - x : x;
    const float v = __fmul_rn(src[((size_t)c * d.H + y) * d.W + xs], (float)m[6 + c]);
    return fminf(fmaxf(v, 0.f), 1.f);
}
constexpr int AUG_DESC_FIELDS = sizeof(AugDesc) / 8;

__device__ __forceinline__ void minmax_finish(float lo, float hi, int include_zero, unsigned int* scratch, double* out2) {
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
        if (lo <= hi) {
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

// stage 0: min / max of the whole float32 image (pre-shrink images only); stage 1: of the pasted region of the crop window
__global__ void __launch_bounds__(256) aug_b_minmax_kernel(const AugDesc* __restrict__ desc, int stage, const double* __restrict__ mats,
                                                            const unsigned char* __restrict__ arena, unsigned int* __restrict__ scratch,
                                                            double* __restrict__ minmax) {
    const AugDesc d = desc[blockIdx.y];
    const double* m = mats + (size_t)blockIdx.y * AUG_MATS;
    if (stage == 0 && d.pre_h == 0) return;
    unsigned int* scr = scratch + ((size_t)blockIdx.y * 2 + stage) * 3;
    double* out2 = minmax + (size_t)blockIdx.y * 4 + stage * 2;
    float lo = INFINITY, hi = -INFINITY;
    if (stage == 0) {
        const float* img = reinterpret_cast<const float*>(d.src);
        const int rw = (int)d.W * 3;
        for (int r = blockIdx.x; r < (int)d.H; r += gridDim.x)
            for (int c = threadIdx.x; c < rw; c += blockDim.x) {
                const float v = d.chw ? aug_src_px(d, img, m, r, c / 3, c % 3) : img[(size_t)r * rw + c];
                lo = fminf(lo, v); hi = fmaxf(hi, v);
            }
        minmax_finish(lo, hi, 0, scr, out2);
        return;
    }
    const int y0 = (int)(d.ny0 + d.oy), y1 = (int)(d.ny1 + d.oy), x0 = (int)(d.nx0 + d.ox), rw = (int)(d.nx1 - d.nx0) * 3;
    if (d.pre_h != 0) {
        const unsigned char* img = arena + d.o_pout;
        for (int r = y0 + blockIdx.x; r < y1; r += gridDim.x)
            for (int c = threadIdx.x; c < rw; c += blockDim.x) {
                const float v = (float)img[((size_t)r * d.pre_w + x0) * 3 + c];
                lo = fminf(lo, v); hi = fmaxf(hi, v);
            }
    } else {
        const float* img = reinterpret_cast<const float*>(d.src);
        for (int r = y0 + blockIdx.x; r < y1; r += gridDim.x)
            for (int c = threadIdx.x; c < rw; c += blockDim.x) {
                const float v = d.chw ? aug_src_px(d, img, m, r, x0 + c / 3, c % 3) : img[((size_t)r * d.W + x0) * 3 + c];
                lo = fminf(lo, v); hi = fmaxf(hi, v);
            }
    }
    minmax_finish(lo, hi, (int)d.has_zero, scr, out2);
}

__global__ void __launch_bounds__(256) aug_b_image_bytes_kernel(const AugDesc* __restrict__ desc, const double* __restrict__ mats,
                                                                 const double* __restrict__ minmax, unsigned char* __restrict__ arena) {
    const AugDesc d = desc[blockIdx.y];
    if (d.pre_h == 0) return;
    const double* m = mats + (size_t)blockIdx.y * AUG_MATS;
    const float* src = reinterpret_cast<const float*>(d.src);
    const double* mm = minmax + (size_t)blockIdx.y * 4;
    const float cmin = (float)mm[0];
    float cscale = __fsub_rn((float)mm[1], cmin);
    if (cscale == 0.f) cscale = 1.f;
    const float scale = (float)__ddiv_rn(255.0, (double)cscale);
    unsigned char* out = arena + d.o_bytes;
    const long long total = d.H * d.W * 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if (d.chw) {
            const long long p = i / 3;
            v = aug_src_px(d, src, m, (int)(p / d.W), (int)(p % d.W), (int)(i - p * 3));
        } else {
            v = __ldg(src + i);
        }
        float b = __fmul_rn(__fsub_rn(v, cmin), scale);
        b = fminf(fmaxf(b, 0.f), 255.f);
        out[i] = (unsigned char)(int)__fadd_rn(b, 0.5f);
    }
}

__device__ __forceinline__ void resample_coeffs_one(int in_size, int out_size, int ksize, int xx, int* __restrict__ bounds,
                                                    int* __restrict__ kk) {
    const double scale = __ddiv_rn((double)(float)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;
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
        ww = __dadd_rn(ww, a < 1.0 ? __dsub_rn(1.0, a) : 0.0);
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

// stage 0: tables of the pre-shrink; stage 1: of the final resize.  blockIdx.z: 0 = width table, 1 = height table
__global__ void aug_b_coeffs_kernel(const AugDesc* __restrict__ desc, int stage, int res, int* __restrict__ arena_i) {
    const AugDesc d = desc[blockIdx.y];
    if (stage == 0 && d.pre_h == 0) return;
    const bool wtab = blockIdx.z == 0;
    int in_size, out_size, ks;
    long long ob, ok;
    if (stage == 0) {
        in_size = (int)(wtab ? d.W : d.H); out_size = (int)(wtab ? d.pre_w : d.pre_h);
        ks = (int)(wtab ? d.ks_pw : d.ks_ph); ob = wtab ? d.c_pw_b : d.c_ph_b; ok = wtab ? d.c_pw_k : d.c_ph_k;
    } else {
        in_size = (int)(wtab ? d.in_w : d.in_h); out_size = res;
        ks = (int)(wtab ? d.ks_fw : d.ks_fh); ob = wtab ? d.c_fw_b : d.c_fh_b; ok = wtab ? d.c_fw_k : d.c_fh_k;
    }
    for (int xx = blockIdx.x * blockDim.x + threadIdx.x; xx < out_size; xx += gridDim.x * blockDim.x)
        resample_coeffs_one(in_size, out_size, ks, xx, arena_i + ob, arena_i + ok);
}

// horizontal pass of stage 0 (whole image bytes -> o_ptmp) or stage 1 (window / rotated window minus padding -> o_ftmp)
__global__ void __launch_bounds__(256) aug_b_resize_h_kernel(const AugDesc* __restrict__ desc, int stage, int res,
                                                              const int* __restrict__ arena_i, unsigned char* __restrict__ arena) {
    const AugDesc d = desc[blockIdx.y];
    if (stage == 0 && d.pre_h == 0) return;
    const unsigned char* src;
    int SW, y_off, x_off, in_h, out_w, ksize;
    const int *bounds, *kk;
    unsigned char* tmp;
    if (stage == 0) {
        src = arena + d.o_bytes; SW = (int)d.W; y_off = 0; x_off = 0; in_h = (int)d.H; out_w = (int)d.pre_w; ksize = (int)d.ks_pw;
        bounds = arena_i + d.c_pw_b; kk = arena_i + d.c_pw_k; tmp = arena + d.o_ptmp;
    } else {
        src = arena + (d.rot ? d.o_rot : d.o_win); SW = (int)d.Wn; y_off = x_off = (int)(d.rot ? d.pad : 0); in_h = (int)d.in_h;
        out_w = res; ksize = (int)d.ks_fw; bounds = arena_i + d.c_fw_b; kk = arena_i + d.c_fw_k; tmp = arena + d.o_ftmp;
    }
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
        o[0] = clip8_22(s0); o[1] = clip8_22(s1); o[2] = clip8_22(s2);
    }
}

__global__ void __launch_bounds__(256) aug_b_resize_v_kernel(const AugDesc* __restrict__ desc, int stage, int res,
                                                              const int* __restrict__ arena_i, unsigned char* __restrict__ arena,
                                                              unsigned char* __restrict__ out_stack) {
    const AugDesc d = desc[blockIdx.y];
    if (stage == 0 && d.pre_h == 0) return;
    const unsigned char* tmp;
    int out_w, out_h, ksize;
    const int *bounds, *kk;
    unsigned char* out;
    if (stage == 0) {
        tmp = arena + d.o_ptmp; out_w = (int)d.pre_w; out_h = (int)d.pre_h; ksize = (int)d.ks_ph;
        bounds = arena_i + d.c_ph_b; kk = arena_i + d.c_ph_k; out = arena + d.o_pout;
    } else {
        tmp = arena + d.o_ftmp; out_w = res; out_h = res; ksize = (int)d.ks_fh;
        bounds = arena_i + d.c_fh_b; kk = arena_i + d.c_fh_k; out = out_stack + (size_t)d.out_index * res * res * 3;
    }
    const int rw = out_w * 3;
    const int total = out_h * rw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int xc = i % rw, yy = i / rw;
        const int ymin = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
        const int* k = kk + (size_t)yy * ksize;
        int sacc = 1 << 21;
        for (int j = 0; j < n; ++j) sacc += (int)tmp[(size_t)(ymin + j) * rw + xc] * __ldg(k + j);
        out[i] = clip8_22(sacc);
    }
}

__global__ void __launch_bounds__(256) aug_b_window_bytes_kernel(const AugDesc* __restrict__ desc, const double* __restrict__ mats,
                                                                  const double* __restrict__ minmax, unsigned char* __restrict__ arena) {
    const AugDesc d = desc[blockIdx.y];
    const double* m = mats + (size_t)blockIdx.y * AUG_MATS;
    const double* mm = minmax + (size_t)blockIdx.y * 4 + 2;
    const double cmin = mm[0];
    double cscale = __dsub_rn(mm[1], cmin);
    if (cscale == 0.0) cscale = 1.0;
    const double scale = __ddiv_rn(255.0, cscale);
    const int Hn = (int)d.Hn, Wn = (int)d.Wn;
    const int ny0 = (int)d.ny0, ny1 = (int)d.ny1, nx0 = (int)d.nx0, nx1 = (int)d.nx1, oy = (int)d.oy, ox = (int)d.ox;
    const bool u8 = d.pre_h != 0;
    const unsigned char* s8 = arena + d.o_pout;
    const float* sf = reinterpret_cast<const float*>(d.src);
    const int SW = (int)(u8 ? d.pre_w : d.W);
    unsigned char* out = arena + d.o_win;
    const int total = Hn * Wn * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % 3, p = i / 3;
        const int x = p % Wn, y = p / Wn;
        double v = 0.0;
        if (y >= ny0 && y < ny1 && x >= nx0 && x < nx1) {
            const size_t idx = ((size_t)(y + oy) * SW + (x + ox)) * 3 + c;
            v = u8 ? (double)s8[idx] : (d.chw ? (double)aug_src_px(d, sf, m, y + oy, x + ox, c) : (double)sf[idx]);
        }
        out[i] = bytescale_f64(v, cmin, scale);
    }
}

__global__ void __launch_bounds__(256) aug_b_rotate_kernel(const AugDesc* __restrict__ desc, const double* __restrict__ mats,
                                                            unsigned char* __restrict__ arena) {
    const AugDesc d = desc[blockIdx.y];
    if (!d.rot) return;
    const double* a = mats + (size_t)blockIdx.y * AUG_MATS;
    const double a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5];
    const int H = (int)d.Hn, W = (int)d.Wn;
    const unsigned char* in = arena + d.o_win;
    unsigned char* out = arena + d.o_rot;
    const int total = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int xo = i % W, yo = i / W;
        const double xc = (double)xo + 0.5, yc = (double)yo + 0.5;
        double xin = __dadd_rn(__dadd_rn(__dmul_rn(a0, xc), __dmul_rn(a1, yc)), a2);
        double yin = __dadd_rn(__dadd_rn(__dmul_rn(a3, xc), __dmul_rn(a4, yc)), a5);
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


extern "C" int hgk_aug_desc_fields(void) { return AUG_DESC_FIELDS; }

extern "C" int hgk_aug_crop_batch(const long long* desc_host, const long long* desc_dev, const double* mats_dev, int N, int res,
                                  unsigned char* arena_u8, int* arena_i, double* minmax, unsigned int* scratch,
                                  unsigned char* out_stack, void* stream) {
    HGK_REQUIRE(desc_host && desc_dev && mats_dev && arena_u8 && arena_i && minmax && scratch && out_stack,
                "hgk_aug_crop_batch: null pointer");
    HGK_REQUIRE(N > 0 && N <= 65535 && res > 0, "hgk_aug_crop_batch: bad batch %d / resolution %d", N, res);
    const AugDesc* hd = reinterpret_cast<const AugDesc*>(desc_host);
    const AugDesc* dd = reinterpret_cast<const AugDesc*>(desc_dev);
    cudaStream_t st = (cudaStream_t)stream;
    bool any_pre = false, any_rot = false;
    long long max_img = 0, max_pre_h_out = 0, max_pre_v_out = 0, max_win = 0, max_fh = 0, max_rows = 1, max_pre_dim = 1;
    for (int i = 0; i < N; ++i) {
        const AugDesc& d = hd[i];
        HGK_REQUIRE(d.src != 0 && d.H > 0 && d.W > 0 && d.Hn > 0 && d.Wn > 0 && d.in_h > 0 && d.in_w > 0 &&
                    d.Hn * d.Wn * 3 < (1ll << 31) && d.H * d.W * 3 < (1ll << 31) && d.out_index >= 0 && d.out_index < N,
                    "hgk_aug_crop_batch: bad descriptor %d", i);
        if (d.pre_h != 0) {
            any_pre = true;
            max_img = max_img > d.H * d.W * 3 ? max_img : d.H * d.W * 3;
            max_pre_h_out = max_pre_h_out > d.H * d.pre_w ? max_pre_h_out : d.H * d.pre_w;
            max_pre_v_out = max_pre_v_out > d.pre_h * d.pre_w * 3 ? max_pre_v_out : d.pre_h * d.pre_w * 3;
            max_rows = max_rows > d.H ? max_rows : d.H;
            const long long m = d.pre_h > d.pre_w ? d.pre_h : d.pre_w;
            max_pre_dim = max_pre_dim > m ? max_pre_dim : m;
        }
        any_rot = any_rot || d.rot != 0;
        max_win = max_win > d.Hn * d.Wn ? max_win : d.Hn * d.Wn;
        max_fh = max_fh > d.in_h * res ? max_fh : d.in_h * res;
        max_rows = max_rows > d.Hn ? max_rows : d.Hn;
    }
    const int rows_grid = (int)(max_rows > 64 ? 64 : max_rows);          // x N images: enough CTAs, few atomics per image
    auto gx = [&](long long total) {                                       // grid.x for `total` items per image
        long long g = (total + 255) / 256;
        const long long cap = (long long)kNumSMs * 8 / N + 1;
        return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
    };
    if (any_pre) {
        aug_b_minmax_kernel<<<dim3(rows_grid, N), 256, 0, st>>>(dd, 0, mats_dev, arena_u8, scratch, minmax);
        aug_b_image_bytes_kernel<<<dim3(gx(max_img), N), 256, 0, st>>>(dd, mats_dev, minmax, arena_u8);
        aug_b_coeffs_kernel<<<dim3((unsigned)((max_pre_dim + 127) / 128), N, 2), 128, 0, st>>>(dd, 0, res, arena_i);
        aug_b_resize_h_kernel<<<dim3(gx(max_pre_h_out), N), 256, 0, st>>>(dd, 0, res, arena_i, arena_u8);
        aug_b_resize_v_kernel<<<dim3(gx(max_pre_v_out), N), 256, 0, st>>>(dd, 0, res, arena_i, arena_u8, out_stack);
    }
    aug_b_minmax_kernel<<<dim3(rows_grid, N), 256, 0, st>>>(dd, 1, mats_dev, arena_u8, scratch, minmax);
    aug_b_window_bytes_kernel<<<dim3(gx(max_win * 3), N), 256, 0, st>>>(dd, mats_dev, minmax, arena_u8);
    if (any_rot) aug_b_rotate_kernel<<<dim3(gx(max_win), N), 256, 0, st>>>(dd, mats_dev, arena_u8);
    aug_b_coeffs_kernel<<<dim3((unsigned)((res + 127) / 128), N, 2), 128, 0, st>>>(dd, 1, res, arena_i);
    aug_b_resize_h_kernel<<<dim3(gx(max_fh), N), 256, 0, st>>>(dd, 1, res, arena_i, arena_u8);
    aug_b_resize_v_kernel<<<dim3(gx((long long)res * res * 3), N), 256, 0, st>>>(dd, 1, res, arena_i, arena_u8, out_stack);
    HGK_CHECK_LAUNCH("hgk_aug_crop_batch");
    return HGK_OK;
}
