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

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------------
// The train step is a chain of ~190 dependent kernels per direction, most of them short: with plain stream order kernel
// N+1's grid is launched only after kernel N has drained, so every boundary pays launch latency + the new kernel's prologue
// (barrier init, TMEM allocation, tensor-map prefetch, index set-up).  Kernels launched through launch_pdl() carry
// cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become resident as soon as every CTA of the
// predecessor has executed pdl_launch_dependents() (SM resources permitting) and run their prologue under the predecessor's
// tail; pdl_wait() (griddepcontrol.wait) then blocks until the predecessor grid has COMPLETED and its writes are visible.
// Rule for every kernel launched this way: no global-memory access of any kind before pdl_wait().
// Under stream capture the same-stream edge becomes a programmatic graph edge; cross-stream (event) edges stay full edges.
// HGK_PDL=0 launches everything with plain stream order (A/B comparisons).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace hgk
