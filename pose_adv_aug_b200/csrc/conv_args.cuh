// Argument block shared by the SIMT and tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace hgk {

// optional finalisers executed by the last CTA (bn_fin.cuh); ticket == nullptr disables them
struct BnFwdFin {
    const float* gamma;
    const float* beta;
    float* rmean;        // running statistics (nullptr: not updated)
    float* rvar;
    float* scale;
    float* shift;
    float* mean;
    float* invstd;
    unsigned int* ticket;
    float eps, momentum;
};
struct BnBwdFin {
    const float* gamma;
    const float* mean;
    const float* invstd;
    float* dgamma;       // += (nullptr: skipped)
    float* dbeta;
    float* cA;
    float* cB;
    float* cC;
    unsigned int* ticket;
    int training;
};

// BatchNorm-backward "apply" of the convolution whose data gradient is being computed, evaluated ON LOAD by the
// image-tile kernel: x.z holds dL/d relu(bn(z)) (= g); the operand is dz = cA*((g*[z*scale+shift > 0] - cC) -
// (z - mean)*cB), which is also written once (tile-interior pixels) to `dz` for the weight-gradient kernel.
struct BnApply {
    const float* z;      // nullptr: disabled
    const float* scale;
    const float* shift;
    const float* mean;
    const float* cA;
    const float* cB;
    const float* cC;
    float* dz;
    int relu;
};

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
    // BN-backward statistics mode (tcgen05 kernel only): the output is dL/d relu(bn(bz)); instead of sum(y), sum(y^2)
    // the epilogue accumulates stat_sum += sum g, stat_sq += sum g*xhat with g = y*[bz*bscale+bshift > 0],
    // xhat = (bz - bmean)*binvstd  (the reduction pass of nn.BatchNorm2d's backward, fused)
    const float* bz;
    const float* bscale;
    const float* bshift;
    const float* bmean;
    const float* binvstd;
    int brelu;
    BnFwdFin ffin;       // forward: (stat_sum, stat_sq) -> scale/shift/running stats
    BnApply ap;          // image-tile data-gradient kernel only
    BnBwdFin bfin;       // BN-backward statistics mode: (stat_sum, stat_sq) = (sum g, sum g*xhat) -> dgamma/dbeta/cA/cB/cC
};

}  // namespace hgk
