// Argument block shared by the SIMT and tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace hgk {

struct ConvArgs {
    Act x;
    int N, H, W, Cin;
    const float* w;
    int ksize, flip;
    const float* bias;
    int Cout;
    Act res;
    float* y;
    int accumulate;
    double* stat_sum;
    double* stat_sq;
    long long P;
};

}  // namespace hgk
