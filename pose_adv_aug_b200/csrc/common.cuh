// Shared helpers for libhgk (sm_100a).  See include/hgk.h for the ABI conventions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/hgk.h"

namespace hgk {

void set_error(const char* fmt, ...);

#define HGK_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            hgk::set_error(__VA_ARGS__);               \
            return HGK_EINVAL;                         \
        }                                              \
    } while (0)

#define HGK_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            hgk::set_error("%s: CUDA error %s", name, cudaGetErrorString(e__));       \
            return HGK_ECUDA;                                                         \
        }                                                                             \
    } while (0)

// A "virtual activation": BatchNorm(+ReLU) of the stored pre-BN tensor applied on load.
struct Act {
    const float* z;
    const float* scale;   // nullptr => identity
    const float* shift;
    int relu;
};

__device__ __forceinline__ float act1(float v, float s, float t, int relu) {
    float u = fmaf(v, s, t);
    return relu ? fmaxf(u, 0.f) : u;
}

__device__ __forceinline__ float4 act4(float4 v, float4 s, float4 t, int relu) {
    float4 u;
    u.x = fmaf(v.x, s.x, t.x);
    u.y = fmaf(v.y, s.y, t.y);
    u.z = fmaf(v.z, s.z, t.z);
    u.w = fmaf(v.w, s.w, t.w);
    if (relu) {
        u.x = fmaxf(u.x, 0.f);
        u.y = fmaxf(u.y, 0.f);
        u.z = fmaxf(u.z, 0.f);
        u.w = fmaxf(u.w, 0.f);
    }
    return u;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// per-channel (scale, shift) for 4 consecutive channels; identity when scale == nullptr
__device__ __forceinline__ void load_affine4(const float* scale, const float* shift, int c, float4& s, float4& t) {
    if (scale != nullptr) {
        s = ldg4(scale + c);
        t = ldg4(shift + c);
    } else {
        s = make_float4(1.f, 1.f, 1.f, 1.f);
        t = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;   // B200

}  // namespace hgk
