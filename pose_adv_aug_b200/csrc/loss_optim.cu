// Per-joint heat-map MSE (inline loss of reference stack-hg.py:156-159), the pylib/Criterion.py
// helpers, the flat multi-tensor RMSprop (stack-hg.py:51-52,165) and the library's error state.
#include <stdlib.h>
#include "common.cuh"
#include <string.h>

namespace hgk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

__device__ __forceinline__ void block_add_double(double v, double* dst) {
    __shared__ double wsum[32];
    v = warp_sum(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < (blockDim.x >> 5) ? wsum[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) atomicAdd(dst, t);
    }
}

// loss += inv_numel * sum (o-t)^2 ;  dout = [acc] + gscale * 2 (o-t) inv_numel   (one pass over o, t)
__global__ void __launch_bounds__(256) mse_fwd_bwd_kernel(const float* __restrict__ out, const float* __restrict__ target,
                                                          long long n, float inv_numel, float gscale, float* dout,
                                                          int accumulate, double* loss) {
    pdl_wait();               // programmatic dependent launch (common.cuh): no global access above
    float part = 0.f;
    double acc = 0.0;
    const float gs = 2.f * inv_numel * gscale;
    const long long n4 = n >> 2;
    int cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 o = ldg4(out + i * 4), t = ldg4(target + i * 4);
        float4 d = make_float4(o.x - t.x, o.y - t.y, o.z - t.z, o.w - t.w);
        part += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        if (++cnt == 8) { acc += (double)part; part = 0.f; cnt = 0; }
        if (dout != nullptr) {
            float4 g = make_float4(d.x * gs, d.y * gs, d.z * gs, d.w * gs);
            if (accumulate) {
                float4 old = ld4(dout + i * 4);
                g.x += old.x; g.y += old.y; g.z += old.z; g.w += old.w;
            }
            st4(dout + i * 4, g);
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            float d = out[i] - target[i];
            part += d * d;
            if (dout != nullptr) dout[i] = (accumulate ? dout[i] : 0.f) + d * gs;
        }
    }
    acc += (double)part;
    block_add_double(acc * (double)inv_numel, loss);
}

// kind 0: (p-g)^2 w ; kind 1: -(g log(p+1e-6) + (1-g) log(1-p+1e-6)) w     (both / numel)
__global__ void __launch_bounds__(256) criterion_fwd_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const float* __restrict__ weight, long long n, double* loss) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float p = pred[i], g = gt[i], w = weight[i];
        float v;
        if (kind == 0) {
            float d = p - g;
            v = d * d * w;
        } else {
            v = -(g * logf(p + 1e-6f) + (1.f - g) * logf(1.f - p + 1e-6f)) * w;
        }
        acc += (double)v;
    }
    block_add_double(acc / (double)n, loss);
}

__global__ void __launch_bounds__(256) criterion_bwd_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const float* __restrict__ weight, long long n,
                                                            const float* __restrict__ gout, float* __restrict__ dpred) {
    const float go = gout[0] / (float)n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float p = pred[i], g = gt[i], w = weight[i];
        float d;
        if (kind == 0) d = 2.f * (p - g) * w;
        else d = -(g / (p + 1e-6f) - (1.f - g) / (1.f - p + 1e-6f)) * w;
        dpred[i] = d * go;
    }
}

__device__ __forceinline__ void rmsprop_flat_body(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, long long n,
                                                  float lr, float alpha, float eps, float gscale) {
    const long long n4 = n >> 2;
    const float oma = 1.f - alpha;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 gg = ldg4(g + i * 4), vv = ld4(v + i * 4), pp = ld4(p + i * 4);
        gg.x *= gscale; gg.y *= gscale; gg.z *= gscale; gg.w *= gscale;
        vv.x = alpha * vv.x + oma * gg.x * gg.x;
        vv.y = alpha * vv.y + oma * gg.y * gg.y;
        vv.z = alpha * vv.z + oma * gg.z * gg.z;
        vv.w = alpha * vv.w + oma * gg.w * gg.w;
        pp.x -= lr * (gg.x / (sqrtf(vv.x) + eps));
        pp.y -= lr * (gg.y / (sqrtf(vv.y) + eps));
        pp.z -= lr * (gg.z / (sqrtf(vv.z) + eps));
        pp.w -= lr * (gg.w / (sqrtf(vv.w) + eps));
        st4(v + i * 4, vv);
        st4(p + i * 4, pp);
    }
    if (blockIdx.x == 0) {
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            float gg = g[i] * gscale;
            float vv = alpha * v[i] + oma * gg * gg;
            v[i] = vv;
            p[i] -= lr * (gg / (sqrtf(vv) + eps));
        }
    }
}

__global__ void __launch_bounds__(256) rmsprop_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v,
                                                           long long n, float lr, float alpha, float eps, float gscale) {
    rmsprop_flat_body(p, g, v, n, lr, alpha, eps, gscale);
}

// hyper-parameters read from device memory: a CUDA graph that captured this launch follows later changes of the learning
// rate (the reference's adjust_lr, utils/util.py) without being re-captured
__global__ void __launch_bounds__(256) rmsprop_flat_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v,
                                                               long long n, const float* __restrict__ hyper) {
    rmsprop_flat_body(p, g, v, n, hyper[0], hyper[1], hyper[2], hyper[3]);
}

__global__ void f64_to_f32_kernel(const double* __restrict__ x, float* __restrict__ y, int n, float mul) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (float)(x[i] * (double)mul);
}

static inline unsigned lo_blocks(long long items) {
    long long b = (items + 255) / 256;
    if (b > 8LL * kNumSMs) b = 8LL * kNumSMs;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace hgk

using namespace hgk;

extern "C" const char* hgk_last_error(void) { return g_err; }
namespace hgk {
// A launch may carry the programmatic-serialization attribute only when the operation in front of it IN ITS STREAM is a kernel
// launch (under stream capture a programmatic edge out of a memset / event node invalidates the capture).  The host knows
// that, the library does not: the caller arms the flag (hgk_pdl_arm) for the launches that directly follow another libhgk
// launch on the same stream; it is consumed by the next launch.  HGK_PDL=0 switches the mechanism off altogether.
static thread_local int g_pdl_armed = 0;
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("HGK_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const bool on = v == 1 && g_pdl_armed != 0;
    g_pdl_armed = 0;
    return on;
}
}  // namespace hgk

extern "C" int hgk_pdl_arm(int on) {
    hgk::g_pdl_armed = on;
    return HGK_OK;
}

extern "C" int hgk_version(void) { return 100; }

extern "C" int hgk_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return prop.major == 10 ? 1 : 0;
}

extern "C" int hgk_mse_fwd_bwd(const float* out, const float* target, long long n, float inv_numel, float gscale,
                               float* dout, int accumulate, double* loss, void* stream) {
    HGK_REQUIRE(out && target && loss && n > 0, "hgk_mse_fwd_bwd: bad arguments");
    HGK_REQUIRE(((uintptr_t)out % 16 == 0) && ((uintptr_t)target % 16 == 0) && ((uintptr_t)dout % 16 == 0),
                "hgk_mse_fwd_bwd: pointers must be 16-byte aligned");
    launch_pdl(mse_fwd_bwd_kernel, dim3(lo_blocks(n / 4)), dim3(256), 0, (cudaStream_t)stream, out, target, n, inv_numel, gscale, dout, accumulate, loss);
    HGK_CHECK_LAUNCH("hgk_mse_fwd_bwd");
    return HGK_OK;
}

extern "C" int hgk_criterion_fwd(int kind, const float* pred, const float* gt, const float* weight, long long n,
                                 double* loss, void* stream) {
    HGK_REQUIRE(pred && gt && weight && loss && n > 0, "hgk_criterion_fwd: bad arguments");
    HGK_REQUIRE(kind == 0 || kind == 1, "hgk_criterion_fwd: kind must be 0 (weighted_L2) or 1 (weighted_sigmoid_crossentropy)");
    criterion_fwd_kernel<<<lo_blocks(n), 256, 0, (cudaStream_t)stream>>>(kind, pred, gt, weight, n, loss);
    HGK_CHECK_LAUNCH("hgk_criterion_fwd");
    return HGK_OK;
}

extern "C" int hgk_criterion_bwd(int kind, const float* pred, const float* gt, const float* weight, long long n,
                                 const float* gout, float* dpred, void* stream) {
    HGK_REQUIRE(pred && gt && weight && gout && dpred && n > 0, "hgk_criterion_bwd: bad arguments");
    HGK_REQUIRE(kind == 0 || kind == 1, "hgk_criterion_bwd: kind must be 0 or 1");
    criterion_bwd_kernel<<<lo_blocks(n), 256, 0, (cudaStream_t)stream>>>(kind, pred, gt, weight, n, gout, dpred);
    HGK_CHECK_LAUNCH("hgk_criterion_bwd");
    return HGK_OK;
}

extern "C" int hgk_rmsprop_flat(float* p, const float* g, float* v, long long n, float lr, float alpha, float eps,
                                float grad_scale, void* stream) {
    HGK_REQUIRE(p && g && v && n > 0, "hgk_rmsprop_flat: bad arguments");
    HGK_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)v % 16 == 0),
                "hgk_rmsprop_flat: flat buffers must be 16-byte aligned");
    rmsprop_flat_kernel<<<lo_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, v, n, lr, alpha, eps, grad_scale);
    HGK_CHECK_LAUNCH("hgk_rmsprop_flat");
    return HGK_OK;
}

extern "C" int hgk_rmsprop_flat_dev(float* p, const float* g, float* v, long long n, const float* hyper, void* stream) {
    HGK_REQUIRE(p && g && v && hyper && n > 0, "hgk_rmsprop_flat_dev: bad arguments");
    HGK_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)v % 16 == 0),
                "hgk_rmsprop_flat_dev: flat buffers must be 16-byte aligned");
    rmsprop_flat_dev_kernel<<<lo_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, v, n, hyper);
    HGK_CHECK_LAUNCH("hgk_rmsprop_flat_dev");
    return HGK_OK;
}

extern "C" int hgk_f64_to_f32(const double* x, float* y, int n, float mul, void* stream) {
    HGK_REQUIRE(x && y && n > 0, "hgk_f64_to_f32: bad arguments");
    f64_to_f32_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, y, n, mul);
    HGK_CHECK_LAUNCH("hgk_f64_to_f32");
    return HGK_OK;
}
