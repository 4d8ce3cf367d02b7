// Callers either side of the training path (SURVEY.md section 8f "next" rows N1, N2, N4 and row a8):
//   * heat-map peaks + PCK accuracy of pylib/Evaluation.py (get_preds :6-23, final_preds :169-193, calc_dists :25-39,
//     dist_acc :41-54, accuracy :56-80, accuracy_origin_res :82-104, per_person_pckh :106-167) -- the reference pulls the
//     last heat-map to the host every iteration (stack-hg.py:176-179) and runs python double loops over it;
//   * the agent's categorical sampling (joint-train-pose-s-r-agent.py:252-271: softmax -> np.random.choice per sample);
//   * the flip-test merge of stack-hg.py:225-232 (pylib/HumanAug.py:179-210: W-flip + left/right channel swap + mean);
//   * the ASN dropout mask applied to neck and skips (models/asn_stacked_hg.py:79-100).
// Everything here is index / compare / tiny-reduction work: one CTA per (sample, joint) or per sample, warp shuffles,
// no tensor cores.
#include "common.cuh"

namespace hgk {

// ---------------------------------------------------------------------------------------------------------------------
// heat-map peaks.  scores: NCHW [N,J,H,W].  mode 0 = get_preds: 1-based (x, y) of the first maximum, (0,0) when max <= 0.
// mode 1 = final_preds: + quarter-pixel shift towards the larger neighbour, + 0.5, then the inverse crop transform
// tinv[n] (row-major 2x3, fp64, computed on the host exactly as the reference does with numpy) applied to (x-1, y-1, 1),
// truncated towards zero, + 1.
// get_preds quirk kept: the row is floor(idx / size(2)) + 1, i.e. divided by H, not W (ref :19).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) heatmap_peaks_kernel(const float* __restrict__ scores, int H, int W, int mode,
                                                             int res0, int res1, const double* __restrict__ tinv,
                                                             int J, float* __restrict__ preds, float* __restrict__ maxval) {
    const int nj = blockIdx.x;
    const float* hm = scores + (size_t)nj * H * W;
    const int HW = H * W;
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float v = __ldg(hm + i);
        if (v > best || (v == best && i < bidx)) { best = v; bidx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    __shared__ float sv[4];
    __shared__ int si[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = best; si[warp] = bidx; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
        if (sv[q] > best || (sv[q] == best && si[q] < bidx)) { best = sv[q]; bidx = si[q]; }
    if (bidx == 0x7fffffff) bidx = 0;                    // all-NaN map: torch.max would return NaN; keep index 0
    float x = (float)(bidx % W) + 1.f;
    float y = floorf((float)bidx / (float)H) + 1.f;
    if (!(best > 0.f)) { x = 0.f; y = 0.f; }             // pred_mask = maxval.gt(0)
    if (maxval != nullptr) maxval[nj] = best;
    if (mode == 2) {                                     // HumanPts.heatmap2pts (pylib/HumanPts.py:118-137): 0-based x, row + 0.5
        x = (float)(bidx % W);
        y = floorf((float)bidx / (float)W) + 0.5f;
        if (!(best > 0.f)) { x = 0.f; y = 0.f; }
    }
    if (mode == 1) {
        const int px = (int)floorf(x), py = (int)floorf(y);
        if (px > 1 && px < res0 && py > 1 && py < res1) {
            const float dx = hm[(py - 1) * W + px] - hm[(py - 1) * W + px - 2];
            const float dy = hm[py * W + px - 1] - hm[(py - 2) * W + px - 1];
            x += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
            y += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
        }
        x += 0.5f;
        y += 0.5f;
        if (tinv != nullptr) {
            const double* t = tinv + (size_t)(nj / J) * 6;
            const double cx = (double)x - 1.0, cy = (double)y - 1.0;
            // np.dot(t, [cx, cy, 1]) row by row, no contraction
            const double nx = __dadd_rn(__dadd_rn(__dmul_rn(t[0], cx), __dmul_rn(t[1], cy)), t[2]);
            const double ny = __dadd_rn(__dadd_rn(__dmul_rn(t[3], cx), __dmul_rn(t[4], cy)), t[5]);
            x = (float)((long long)nx + 1);              // .astype(int) truncates towards zero
            y = (float)((long long)ny + 1);
        }
    }
    preds[(size_t)nj * 2 + 0] = x;
    preds[(size_t)nj * 2 + 1] = y;
}

// ---------------------------------------------------------------------------------------------------------------------
// calc_dists + dist_acc + the averaging loop of accuracy()/accuracy_origin_res().  One CTA.
// dists: [J,N] (joint-major, as the reference); acc: [n_idx+1], acc[0] = mean of the non-negative per-joint values.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pck_accuracy_kernel(const float* __restrict__ preds, const float* __restrict__ target,
                                                            const float* __restrict__ normalize, int N, int J, float boundary,
                                                            float thr, const int* __restrict__ idxs, int n_idx,
                                                            float* __restrict__ dists, float* __restrict__ acc) {
    for (int i = threadIdx.x; i < N * J; i += blockDim.x) {
        const int n = i / J, c = i - n * J;
        const float tx = target[i * 2], ty = target[i * 2 + 1];
        float d = -1.f;
        if (tx > boundary && ty > boundary) {
            const float dx = preds[i * 2] - tx, dy = preds[i * 2 + 1] - ty;
            d = sqrtf(dx * dx + dy * dy) / normalize[n];
        }
        dists[(size_t)c * N + n] = d;
    }
    __syncthreads();
    if (acc == nullptr) return;
    __shared__ float sacc[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < n_idx; k += (int)(blockDim.x >> 5)) {
        const float* row = dists + (size_t)idxs[k] * N;
        int valid = 0, hit = 0;
        for (int n = lane; n < N; n += 32) {
            const float d = row[n];
            if (d != -1.f) { ++valid; if (d <= thr) ++hit; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            valid += __shfl_xor_sync(0xffffffffu, valid, o);
            hit += __shfl_xor_sync(0xffffffffu, hit, o);
        }
        if (lane == 0) {
            const float a = valid > 0 ? (float)((double)hit / (double)valid) : -1.f;
            acc[k + 1] = a;
            sacc[k] = a;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        int cnt = 0;
        for (int k = 0; k < n_idx; ++k)
            if (sacc[k] >= 0.f) { sum += sacc[k]; ++cnt; }
        acc[0] = cnt ? sum / (float)cnt : 0.f;
    }
}

// dist_acc (ref :41-54) of one flat vector of distances: share of the entries != -1 that are <= thr, -1 when none is valid
__global__ void __launch_bounds__(256) dist_acc_kernel(const float* __restrict__ dists, int n, float thr, float* __restrict__ out) {
    int valid = 0, hit = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = dists[i];
        if (d != -1.f) { ++valid; if (d <= thr) ++hit; }
    }
    __shared__ int sv[8], sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        valid += __shfl_xor_sync(0xffffffffu, valid, o);
        hit += __shfl_xor_sync(0xffffffffu, hit, o);
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = valid; sh[threadIdx.x >> 5] = hit; }
    __syncthreads();
    if (threadIdx.x == 0) {
        valid = 0; hit = 0;
        for (int q = 0; q < 8; ++q) { valid += sv[q]; hit += sh[q]; }
        out[0] = valid > 0 ? (float)((double)hit / (double)valid) : -1.f;
    }
}

// per_person_pckh (ref :106-167): per sample, over the joints idxs: valid = dist != -1 and the ground-truth heat-map peak
// (gt_preds from get_preds) has both coordinates > 1; accuracy = #(dist <= thr and valid) / #valid, 0 when either set is empty.
__global__ void per_person_pckh_kernel(const float* __restrict__ dists, const float* __restrict__ gt_preds, int N, int J,
                                       const int* __restrict__ idxs, int n_idx, float thr, float* __restrict__ acc_vec) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int s1 = 0, s2 = 0, all = 0, ok = 0;
    for (int k = 0; k < n_idx; ++k) {
        const int c = idxs[k];
        const float d = dists[(size_t)c * N + n];
        const bool ind = gt_preds[((size_t)n * J + c) * 2] > 1.f && gt_preds[((size_t)n * J + c) * 2 + 1] > 1.f;
        s1 += d != -1.f;
        s2 += ind;
        if (d != -1.f && ind) { ++all; ok += d <= thr; }
    }
    acc_vec[n] = (s1 > 0 && s2 > 0) ? (float)ok / (float)all : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// flip-test merge: out[n,c,h,w] = (a[n,c,h,w] + b[n,perm[c],h,W-1-w]) / 2; a = heat-maps of the image, b = heat-maps of
// the W-flipped image, perm = the left/right joint swap.  a == nullptr: out = f(b) alone (flip_channels / shuffle_channels);
// flip_w == 0: channel permutation only.  NCHW, W % 4 == 0.
// ---------------------------------------------------------------------------------------------------------------------
struct Perm32 { int p[32]; };

__global__ void __launch_bounds__(256) flip_merge_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int H,
                                                         int W, Perm32 perm, int flip_w, long long n4, float* __restrict__ out) {
    const int W4 = W >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int w4 = (int)(i % W4);
        const long long row = i / W4;                      // (n*C + c)*H + h
        const int h = (int)(row % H);
        const long long nc = row / H;
        const int c = (int)(nc % C);
        const long long n = nc / C;
        const long long brow = ((n * C + perm.p[c]) * H + h) * (long long)W;
        float4 vb;
        if (flip_w) {
            const float4 t = ldg4(b + brow + (W - 4 - w4 * 4));
            vb = make_float4(t.w, t.z, t.y, t.x);
        } else {
            vb = ldg4(b + brow + w4 * 4);
        }
        if (a != nullptr) {
            const float4 va = ldg4(a + i * 4);
            vb = make_float4((va.x + vb.x) / 2.f, (va.y + vb.y) / 2.f, (va.z + vb.z) / 2.f, (va.w + vb.w) / 2.f);
        }
        st4(out + i * 4, vb);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// agent sampling: probs = softmax(logits) (fp32, max-subtracted), index = first k with cdf[k] > u where cdf is the fp64
// running sum of the fp32 probabilities divided by its total -- np.random.choice(K, 1, p=probs) given the uniform u
// (RandomState.choice: cdf = p.cumsum(); cdf /= cdf[-1]; cdf.searchsorted(u, side='right')).  One thread per row.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void softmax_sample_kernel(const float* __restrict__ logits, int N, int K, const double* __restrict__ u,
                                      float* __restrict__ probs, long long* __restrict__ index) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* l = logits + (size_t)n * K;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, l[k]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf(l[k] - m);
    double total = 0.0;
    for (int k = 0; k < K; ++k) {
        const float p = expf(l[k] - m) / s;
        if (probs != nullptr) probs[(size_t)n * K + k] = p;
        total += (double)p;
    }
    if (index == nullptr) return;
    const double un = u[n];
    double run = 0.0;
    int pick = K;                                         // searchsorted(..., 'right') returns K when u >= cdf[-1] = 1
    for (int k = 0; k < K; ++k) {
        run += (double)(expf(l[k] - m) / s);
        if (run / total > un) { pick = k; break; }
    }
    index[n] = pick;
}

// ---------------------------------------------------------------------------------------------------------------------
// ASN dropout (ref :79-100): y = T(x) * nearest_upsample(mask) ; mask [N,MH,MW] (the reference's [N,1,4,4]), x NHWC.
// Backward: gx = [acc ? gx : 0] + g * mask.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_mul_fwd_kernel(Act x, const float* __restrict__ mask, int H, int W, int C, int MH,
                                                           int MW, long long n4, float* __restrict__ y) {
    const int C4 = C >> 2;
    const int sh = H / MH, sw = W / MW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const long long pix = i / C4;
        const int w = (int)(pix % W);
        const long long r = pix / W;
        const int h = (int)(r % H);
        const long long n = r / H;
        float4 s, t;
        load_affine4(x.scale, x.shift, c4 * 4, s, t);
        float4 v = ldg4(x.z + i * 4);
        if (x.scale != nullptr) v = act4(v, s, t, x.relu);
        const float m = __ldg(mask + (n * MH + h / sh) * MW + w / sw);
        st4(y + i * 4, make_float4(v.x * m, v.y * m, v.z * m, v.w * m));
    }
}

__global__ void __launch_bounds__(256) mask_mul_bwd_kernel(const float* __restrict__ g, const float* __restrict__ mask, int H,
                                                           int W, int C, int MH, int MW, long long n4, float* __restrict__ gx,
                                                           int accumulate) {
    const int C4 = C >> 2;
    const int sh = H / MH, sw = W / MW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / C4;
        const int w = (int)(pix % W);
        const long long r = pix / W;
        const int h = (int)(r % H);
        const long long n = r / H;
        const float m = __ldg(mask + (n * MH + h / sh) * MW + w / sw);
        const float4 v = ldg4(g + i * 4);
        float4 o = make_float4(v.x * m, v.y * m, v.z * m, v.w * m);
        if (accumulate) {
            const float4 old = ld4(gx + i * 4);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        st4(gx + i * 4, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// HumanPts.pts2heatmap + draw_gaussian (pylib/HumanPts.py:36-48,82-116): one map per point, zeros plus the size x size
// blob g (computed on the host exactly as the reference: exp(-((x-x0)^2+(y-y0)^2)/tmp_size^2), tmp_size = ceil(3 sigma),
// size = 2 tmp_size + 1) pasted with its upper-left corner at (int(px - tmp_size), int(py - tmp_size)), clipped to the
// map.  Points with x <= 0, y <= 0, x > W or y > H leave an all-zero map and a zero row in valid_pts.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pts2heatmap_kernel(const float* __restrict__ pts, int H, int W, const float* __restrict__ g,
                                                          int size, float* __restrict__ heatmap, float* __restrict__ valid_pts) {
    const int m = blockIdx.x;
    const float px = pts[(size_t)m * 2], py = pts[(size_t)m * 2 + 1];
    const bool valid = !(px <= 0.f || py <= 0.f || px > (float)W || py > (float)H);
    const int tmp = size >> 1;
    const int ulx = (int)(px - (float)tmp), uly = (int)(py - (float)tmp);      // int() truncates towards zero
    const int brx = (int)(px + (float)tmp), bry = (int)(py + (float)tmp);
    const bool draw = valid && !(ulx >= W || uly >= H || brx < 0 || bry < 0);
    // image range [x0,x1) x [y0,y1), blob offset (gx0, gy0)
    const int x0 = max(0, ulx), x1 = min(brx + 1, W), y0 = max(0, uly), y1 = min(bry + 1, H);
    const int gx0 = max(0, -ulx), gy0 = max(0, -uly);
    float* out = heatmap + (size_t)m * H * W;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int y = i / W, x = i - y * W;
        float v = 0.f;
        if (draw && x >= x0 && x < x1 && y >= y0 && y < y1) v = __ldg(g + (gy0 + y - y0) * size + (gx0 + x - x0));
        out[i] = v;
    }
    if (threadIdx.x == 0 && valid_pts != nullptr) {
        valid_pts[(size_t)m * 2] = valid ? px : 0.f;
        valid_pts[(size_t)m * 2 + 1] = valid ? py : 0.f;
    }
}

static int stream_grid(long long n4) {
    long long b = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace hgk

using namespace hgk;

extern "C" {

int hgk_heatmap_peaks(const float* scores, int N, int J, int H, int W, int mode, int res0, int res1, const double* tinv,
                      float* preds, float* maxval, void* stream) {
    HGK_REQUIRE(scores != nullptr && preds != nullptr, "hgk_heatmap_peaks: null pointer");
    HGK_REQUIRE(N >= 0 && J >= 1 && H >= 1 && W >= 1 && (long long)H * W < (1LL << 30), "hgk_heatmap_peaks: bad shape");
    HGK_REQUIRE(mode >= 0 && mode <= 2, "hgk_heatmap_peaks: mode must be 0 (get_preds), 1 (final_preds) or 2 (heatmap2pts)");
    HGK_REQUIRE(mode != 1 || (res0 >= 1 && res0 <= W && res1 >= 1 && res1 <= H),
                "hgk_heatmap_peaks: res (%d,%d) must lie inside the %dx%d heat-map", res0, res1, W, H);
    if ((long long)N * J == 0) return HGK_OK;
    HGK_REQUIRE((long long)N * J < (1LL << 31), "hgk_heatmap_peaks: too many maps");
    heatmap_peaks_kernel<<<(unsigned)(N * J), 128, 0, (cudaStream_t)stream>>>(scores, H, W, mode, res0, res1, tinv, J, preds, maxval);
    HGK_CHECK_LAUNCH("hgk_heatmap_peaks");
    return HGK_OK;
}

int hgk_pts2heatmap(const float* pts, int M, int H, int W, const float* g, int size, float* heatmap, float* valid_pts,
                    void* stream) {
    HGK_REQUIRE(pts != nullptr && g != nullptr && heatmap != nullptr, "hgk_pts2heatmap: null pointer");
    HGK_REQUIRE(M >= 0 && H >= 1 && W >= 1 && (long long)H * W < (1LL << 30), "hgk_pts2heatmap: bad shape");
    HGK_REQUIRE(size >= 1 && (size & 1), "hgk_pts2heatmap: the blob size must be odd (2 ceil(3 sigma) + 1), got %d", size);
    if (M == 0) return HGK_OK;
    pts2heatmap_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(pts, H, W, g, size, heatmap, valid_pts);
    HGK_CHECK_LAUNCH("hgk_pts2heatmap");
    return HGK_OK;
}

int hgk_pck_accuracy(const float* preds, const float* target, const float* normalize, int N, int J, float boundary, float thr,
                     const int* idxs, int n_idx, float* dists, float* acc, void* stream) {
    HGK_REQUIRE(preds != nullptr && target != nullptr && normalize != nullptr && dists != nullptr, "hgk_pck_accuracy: null pointer");
    HGK_REQUIRE(N >= 1 && J >= 1 && (long long)N * J < (1LL << 24), "hgk_pck_accuracy: bad shape");
    HGK_REQUIRE(acc == nullptr || (idxs != nullptr && n_idx >= 1 && n_idx <= 64), "hgk_pck_accuracy: 1..64 joint indices");
    pck_accuracy_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(preds, target, normalize, N, J, boundary, thr, idxs, n_idx, dists, acc);
    HGK_CHECK_LAUNCH("hgk_pck_accuracy");
    return HGK_OK;
}

int hgk_dist_acc(const float* dists, int n, float thr, float* out, void* stream) {
    HGK_REQUIRE(dists != nullptr && out != nullptr && n >= 0, "hgk_dist_acc: bad argument");
    dist_acc_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dists, n, thr, out);
    HGK_CHECK_LAUNCH("hgk_dist_acc");
    return HGK_OK;
}

int hgk_per_person_pckh(const float* dists, const float* gt_preds, int N, int J, const int* idxs, int n_idx, float thr,
                        float* acc_vec, void* stream) {
    HGK_REQUIRE(dists != nullptr && gt_preds != nullptr && idxs != nullptr && acc_vec != nullptr, "hgk_per_person_pckh: null pointer");
    HGK_REQUIRE(N >= 1 && J >= 1 && n_idx >= 1, "hgk_per_person_pckh: bad shape");
    per_person_pckh_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dists, gt_preds, N, J, idxs, n_idx, thr, acc_vec);
    HGK_CHECK_LAUNCH("hgk_per_person_pckh");
    return HGK_OK;
}

int hgk_flip_merge_nchw(const float* a, const float* b_flipped, int N, int C, int H, int W, const int* pairs, int n_pairs,
                        int flip_w, float* out, void* stream) {
    HGK_REQUIRE(b_flipped != nullptr && out != nullptr, "hgk_flip_merge_nchw: null pointer");
    HGK_REQUIRE(out != b_flipped, "hgk_flip_merge_nchw: not an in-place operation");
    HGK_REQUIRE(N >= 0 && C >= 1 && C <= 32 && H >= 1 && W >= 4 && W % 4 == 0, "hgk_flip_merge_nchw: need C <= 32 and W %% 4 == 0");
    HGK_REQUIRE(n_pairs >= 0 && n_pairs <= 16 && (n_pairs == 0 || pairs != nullptr), "hgk_flip_merge_nchw: bad pair list");
    Perm32 perm;
    for (int c = 0; c < 32; ++c) perm.p[c] = c;
    // the reference swaps the pairs one after the other (HumanAug.py:191-195); compose the same way (host ints)
    for (int i = 0; i < n_pairs; ++i) {
        const int i1 = pairs[2 * i], i2 = pairs[2 * i + 1];
        HGK_REQUIRE(i1 >= 0 && i1 < C && i2 >= 0 && i2 < C, "hgk_flip_merge_nchw: pair index out of range");
        const int t = perm.p[i1]; perm.p[i1] = perm.p[i2]; perm.p[i2] = t;
    }
    const long long n4 = (long long)N * C * H * W / 4;
    if (n4 == 0) return HGK_OK;
    flip_merge_kernel<<<stream_grid(n4), 256, 0, (cudaStream_t)stream>>>(a, b_flipped, C, H, W, perm, flip_w != 0, n4, out);
    HGK_CHECK_LAUNCH("hgk_flip_merge_nchw");
    return HGK_OK;
}

int hgk_softmax_sample(const float* logits, int N, int K, const double* u, float* probs, long long* index, void* stream) {
    HGK_REQUIRE(logits != nullptr && (probs != nullptr || index != nullptr), "hgk_softmax_sample: null pointer");
    HGK_REQUIRE(index == nullptr || u != nullptr, "hgk_softmax_sample: sampling needs the uniforms");
    HGK_REQUIRE(N >= 1 && K >= 1 && K <= 4096, "hgk_softmax_sample: bad shape");
    softmax_sample_kernel<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(logits, N, K, u, probs, index);
    HGK_CHECK_LAUNCH("hgk_softmax_sample");
    return HGK_OK;
}

int hgk_mask_mul_fwd(const float* x, const float* x_scale, const float* x_shift, int x_relu, const float* mask, int N, int H,
                     int W, int C, int MH, int MW, float* y, void* stream) {
    HGK_REQUIRE(x != nullptr && mask != nullptr && y != nullptr, "hgk_mask_mul_fwd: null pointer");
    HGK_REQUIRE(C % 4 == 0 && C >= 4, "hgk_mask_mul_fwd: C must be a multiple of 4 (got %d)", C);
    HGK_REQUIRE(MH >= 1 && MW >= 1 && H % MH == 0 && W % MW == 0, "hgk_mask_mul_fwd: %dx%d is not a multiple of the %dx%d mask", H, W, MH, MW);
    const long long n4 = (long long)N * H * W * C / 4;
    if (n4 == 0) return HGK_OK;
    mask_mul_fwd_kernel<<<stream_grid(n4), 256, 0, (cudaStream_t)stream>>>(Act{x, x_scale, x_shift, x_relu}, mask, H, W, C, MH, MW, n4, y);
    HGK_CHECK_LAUNCH("hgk_mask_mul_fwd");
    return HGK_OK;
}

int hgk_mask_mul_bwd(const float* g, const float* mask, int N, int H, int W, int C, int MH, int MW, float* gx, int accumulate,
                     void* stream) {
    HGK_REQUIRE(g != nullptr && mask != nullptr && gx != nullptr, "hgk_mask_mul_bwd: null pointer");
    HGK_REQUIRE(C % 4 == 0 && C >= 4, "hgk_mask_mul_bwd: C must be a multiple of 4 (got %d)", C);
    HGK_REQUIRE(MH >= 1 && MW >= 1 && H % MH == 0 && W % MW == 0, "hgk_mask_mul_bwd: %dx%d is not a multiple of the %dx%d mask", H, W, MH, MW);
    const long long n4 = (long long)N * H * W * C / 4;
    if (n4 == 0) return HGK_OK;
    mask_mul_bwd_kernel<<<stream_grid(n4), 256, 0, (cudaStream_t)stream>>>(g, mask, H, W, C, MH, MW, n4, gx, accumulate);
    HGK_CHECK_LAUNCH("hgk_mask_mul_bwd");
    return HGK_OK;
}

}  // extern "C"
