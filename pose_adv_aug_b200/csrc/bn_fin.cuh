// BatchNorm finalisers run by the LAST CTA of the kernel that produced the per-channel sums (ticket counter +
// __threadfence, the classic last-block pattern): removes one C-length launch per BatchNorm from the critical
// path of the forward (scale/shift, running statistics) and of the backward (dgamma, dbeta, cA/cB/cC).
// Same arithmetic as bn_finalize_kernel / bn_bwd_finalize_kernel (bn.cu).
#pragma once
#include "common.cuh"
#include "conv_args.cuh"

namespace hgk {

// Every thread of the CTA must call this after its last atomic on the statistics.  Returns true in all threads of
// the one CTA that arrived last; that CTA also re-arms the ticket for the next launch (CUDA-graph replay).
__device__ __forceinline__ bool last_cta_arrives(unsigned int* ticket, unsigned int total) {
    __shared__ unsigned int s_last;
    __threadfence();                      // this thread's statistics atomics are ordered before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == total - 1u) ? 1u : 0u;
        if (t == total - 1u) atomicExch(ticket, 0u);
    }
    __syncthreads();
    const bool last = s_last != 0u;
    if (last) __threadfence();
    return last;
}

__device__ __forceinline__ void bn_fwd_finalize_cta(const BnFwdFin& f, const double* sum, const double* sq, double count, int C) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double mean = __ldcg(sum + c) / count;
        double var = __ldcg(sq + c) / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const double invstd = 1.0 / sqrt(var + (double)f.eps);
        const double g = (double)f.gamma[c];
        f.scale[c] = (float)(g * invstd);
        f.shift[c] = (float)((double)f.beta[c] - mean * g * invstd);
        f.mean[c] = (float)mean;
        f.invstd[c] = (float)invstd;
        if (f.rmean != nullptr) {
            const double unb = count > 1.0 ? var * (count / (count - 1.0)) : var;
            f.rmean[c] = (float)((1.0 - f.momentum) * (double)f.rmean[c] + f.momentum * mean);
            f.rvar[c] = (float)((1.0 - f.momentum) * (double)f.rvar[c] + f.momentum * unb);
        }
    }
}

__device__ __forceinline__ void bn_bwd_finalize_cta(const BnBwdFin& f, const double* sum_g, const double* sum_gx, double count, int C) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double sg = __ldcg(sum_g + c), sgx = __ldcg(sum_gx + c);
        if (f.dgamma != nullptr) f.dgamma[c] += (float)sgx;
        if (f.dbeta != nullptr) f.dbeta[c] += (float)sg;
        const double is = (double)f.invstd[c];
        f.cA[c] = (float)((double)f.gamma[c] * is);
        if (f.training) {
            f.cB[c] = (float)(sgx / count * is);
            f.cC[c] = (float)(sg / count);
        } else {
            f.cB[c] = 0.f;
            f.cC[c] = 0.f;
        }
    }
}

}  // namespace hgk
