// Inter-stack head of _Hourglass_Wrapper (reference models/asn_stacked_hg.py:329-334):
//     out   = out_conv(y)                       C -> J heat-maps
//     x_new = x + forth_conv(y) + in_conv(out)  C -> C and J -> C
// in_conv(out_conv(y)) is linear in y, so the J -> C convolution and the C -> C convolution are ONE C -> C convolution with
//     Wc = Wf + Wi Wo        bc = bf + bi + Wi bo
// and x_new = x + conv(y; Wc, bc) is a single tensor-core launch over y: the J -> C pass over the 100 MB activation, its
// data gradient (C -> J) and its weight gradient (three passes, two of them fp32 SIMT kernels) disappear from the step.
// The chain rule through the combination is exact:
//     dWf = dWc              dWi = dWc Wo^T          dWo += Wi^T dWc        (out_conv also has its direct gradient from the loss)
//     dbf = dbi = dbc        dbo += Wi^T dbc
// Both kernels are weight-space products (C*C*J = 1 M FMA for C = 256, J = 16) with fixed summation order.
#include "common.cuh"

namespace hgk {

// grid = C blocks (row c of Wc), any block size
__global__ void head_combine_fwd_kernel(const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ Wi,
                                        const float* __restrict__ bi, const float* __restrict__ Wo, const float* __restrict__ bo,
                                        float* __restrict__ Wc, float* __restrict__ bc, int C, int J) {
    const int c = blockIdx.x;
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < J; ++j) acc = fmaf(__ldg(Wi + (size_t)c * J + j), __ldg(Wo + (size_t)j * C + k), acc);
        Wc[(size_t)c * C + k] = __ldg(Wf + (size_t)c * C + k) + acc;
    }
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int j = 0; j < J; ++j) acc = fmaf(__ldg(Wi + (size_t)c * J + j), bo != nullptr ? __ldg(bo + j) : 0.f, acc);
        bc[c] = (bf != nullptr ? __ldg(bf + c) : 0.f) + (bi != nullptr ? __ldg(bi + c) : 0.f) + acc;
    }
}

// blocks [0, C): row c -> dWf, dWi, dbf, dbi;  blocks [C, C + J): heat-map channel j -> dWo, dbo.  256 threads.
__global__ void __launch_bounds__(256) head_combine_bwd_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc,
                                                               const float* __restrict__ Wi, const float* __restrict__ Wo,
                                                               float* dWf, float* dbf, float* dWi, float* dbi, float* dWo,
                                                               float* dbo, int C, int J) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if ((int)blockIdx.x < C) {
        const int c = blockIdx.x;
        for (int k = tid; k < C; k += 256) dWf[(size_t)c * C + k] += dWc[(size_t)c * C + k];
        for (int j = warp; j < J; j += 8) {                     // dWi[c][j] += sum_k dWc[c][k] Wo[j][k]
            float acc = 0.f;
            for (int k = lane; k < C; k += 32) acc = fmaf(dWc[(size_t)c * C + k], __ldg(Wo + (size_t)j * C + k), acc);
            acc = warp_sum(acc);
            if (lane == 0) dWi[(size_t)c * J + j] += acc;
        }
        if (tid == 0) {
            const float d = dbc[c];
            if (dbf != nullptr) dbf[c] += d;
            if (dbi != nullptr) dbi[c] += d;
        }
    } else {
        const int j = blockIdx.x - C;
        for (int k = tid; k < C; k += 256) {                    // dWo[j][k] += sum_c Wi[c][j] dWc[c][k]
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc = fmaf(__ldg(Wi + (size_t)c * J + j), dWc[(size_t)c * C + k], acc);
            dWo[(size_t)j * C + k] += acc;
        }
        if (dbo != nullptr && warp == 0) {                      // dbo[j] += sum_c Wi[c][j] dbc[c]
            float acc = 0.f;
            for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(Wi + (size_t)c * J + j), dbc[c], acc);
            acc = warp_sum(acc);
            if (lane == 0) dbo[j] += acc;
        }
    }
}

}  // namespace hgk

using namespace hgk;

extern "C" int hgk_head_combine_fwd(const float* w_forth, const float* b_forth, const float* w_in, const float* b_in,
                                    const float* w_out, const float* b_out, float* w_comb, float* b_comb, int C, int J,
                                    void* stream) {
    HGK_REQUIRE(w_forth && w_in && w_out && w_comb && b_comb, "hgk_head_combine_fwd: null pointer");
    HGK_REQUIRE(C > 0 && J > 0 && C <= 65535, "hgk_head_combine_fwd: need 0 < C <= 65535 and J > 0 (C=%d J=%d)", C, J);
    head_combine_fwd_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(w_forth, b_forth, w_in, b_in, w_out, b_out, w_comb, b_comb, C, J);
    HGK_CHECK_LAUNCH("hgk_head_combine_fwd");
    return HGK_OK;
}

extern "C" int hgk_head_combine_bwd(const float* dw_comb, const float* db_comb, const float* w_in, const float* w_out,
                                    float* dw_forth, float* db_forth, float* dw_in, float* db_in, float* dw_out, float* db_out,
                                    int C, int J, void* stream) {
    HGK_REQUIRE(dw_comb && db_comb && w_in && w_out && dw_forth && dw_in && dw_out, "hgk_head_combine_bwd: null pointer");
    HGK_REQUIRE(C > 0 && J > 0 && C + J <= 65535, "hgk_head_combine_bwd: need C > 0, J > 0, C + J <= 65535 (C=%d J=%d)", C, J);
    head_combine_bwd_kernel<<<C + J, 256, 0, (cudaStream_t)stream>>>(dw_comb, db_comb, w_in, w_out, dw_forth, db_forth, dw_in, db_in,
                                                                     dw_out, db_out, C, J);
    HGK_CHECK_LAUNCH("hgk_head_combine_bwd");
    return HGK_OK;
}
