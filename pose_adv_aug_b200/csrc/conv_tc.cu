// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a: 1x1 / 3x3(p1) NHWC with the
// same contract as the SIMT kernel (BN+ReLU applied on load, bias / residual / accumulate /
// BN-statistics epilogue), TF32 operands with fp32 accumulation in TMEM.
//
//   forward  : 3xTF32 error-compensated split  x*w ~= xh*wh + xh*wl + xl*wh  (fp32-class accuracy,
//              heat-map parity <= 1e-3 needs it: plain TF32 is 2.6e-2 off)
//   backward : data gradient with 1xTF32 (w_lo == NULL)
//
// CTA = 128 output pixels x BN (= Cout) channels, K streamed in 32-channel stages:
//   * all 8 warps load the activation tile from global (coalesced 128-bit), apply BN+ReLU, split
//     hi/lo and write the UMMA canonical K-major (no-swizzle) operand tile to shared memory;
//   * the weight stage is one contiguous pre-packed block fetched with cp.async.bulk (TMA bulk
//     copy) that completes on an mbarrier;
//   * one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) into a TMEM accumulator and
//     commits to the stage's mbarrier, which frees the stage for the producers;
//   * epilogue: tcgen05.ld -> shared staging tile -> coalesced bias/residual/accumulate/store and
//     fp64 per-channel statistics.
//
// SPLIT-K OVER A THREAD-BLOCK CLUSTER (the 16x16 / 8x8 / 4x4 rungs of the hourglass): a 3x3 128->128 layer at 4x4 has
// P = 384 pixels = 3 CTAs, each of which used to walk all 36 (tap, channel-chunk) stages alone -- a 48 us latency chain
// on 3 of 148 SMs, four of them per residual block on the critical path of every hourglass.  Launched with a cluster of
// S CTAs along grid.z, CTA `rank` runs stages [rank*T/S, (rank+1)*T/S) into its own TMEM accumulator, stages the partial
// 128 x BN tile in its shared memory, and after a cluster barrier every CTA sums ITS 128/S pixel rows over the S partial
// tiles through distributed shared memory (fixed order: rank 0 .. S-1) and runs the usual epilogue on them.
#include <stdlib.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"
#include "bn_fin.cuh"

namespace hgk {

long long* g_dbg_buf = nullptr;

// Per-instantiation configuration.  Everything is sized so that TWO CTAs fit on one SM (<= ~97 KB of shared
// memory, <= 256 TMEM columns, <= 112 registers): while one CTA is in its prologue (TMEM alloc, first loads)
// or epilogue, the other one keeps the tensor pipe and the memory system busy.
// BIG = one CTA per SM with 32-channel stages and the whole shared memory: used when the grid has at most one
// CTA per SM anyway (the 4x4 .. 32x32 layers), where fewer, larger stages shorten the latency-bound pipeline.
template <int BN, bool SPLIT, bool BIG = false>
struct TcCfg {
    static constexpr int BK = (!BIG && SPLIT && BN >= 128) ? 16 : 32;      // channels per pipeline stage
    static constexpr int QP = BK / 4;                                       // 16-byte channel quads per pixel per stage
    static constexpr int NJ = QP / 2;                                       // (pixel, quad) items per producer thread
    static constexpr int A_BYTES = TBM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int SMEM_BUDGET = BIG ? 196 * 1024 : 96 * 1024;
    static constexpr int NST = (SMEM_BUDGET / STAGE) > 4 ? 4 : (SMEM_BUDGET / STAGE);
    // The tensor core accumulates in fp32 with truncation: a chain of n dependent accumulations carries a
    // systematic ~n*2^-25 relative shrink (1.3e-5 measured at K = 9*256 with all three 3xTF32 terms in one
    // chain).  So the hi*hi products and the two small cross terms get separate TMEM accumulators (BN = 64:
    // two alternating hi*hi accumulators), added with round-to-nearest fp32 adds in the epilogue.  BN = 256
    // only occurs with K <= 256 (1x1 convolutions): one 96-step chain, bias ~1.5e-6.
    static constexpr int NMAIN = (SPLIT && BN <= 64) ? 2 : 1;
    static constexpr int NACC = !SPLIT ? 1 : (BN >= 256 ? 1 : NMAIN + 1);
    static constexpr int TMEM_COLS = (NACC * BN <= 64) ? 64 : (NACC * BN <= 128) ? 128 : 256;
    static constexpr int CH = BN > 128 ? 128 : BN;                          // epilogue column chunk
    static constexpr int STG_BYTES = TBM * (CH + 4) * 4 + 16384;
    static constexpr int SMEM = (NST * STAGE > STG_BYTES ? NST * STAGE : STG_BYTES) + 256;
    static_assert(NACC * BN <= 256, "two CTAs per SM share the 512 TMEM columns");
    static_assert(NST >= 2, "pipeline needs two stages");
};

__device__ __forceinline__ uint32_t cluster_nctaid_z() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%cluster_nctaid.z;" : "=r"(v));
    return v;
}
__device__ __forceinline__ uint32_t cluster_ctaid_z() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%cluster_ctaid.z;" : "=r"(v));
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in the CTA with rank `rank`
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int BN, bool SPLIT, bool BWDSTATS, bool BIG>
__global__ void __launch_bounds__(TNT + 32, BIG ? 1 : 2) conv_tc_kernel(const TcArgs args) {
    using Cfg = TcCfg<BN, SPLIT, BIG>;
    constexpr int BK = Cfg::BK, NJ = Cfg::NJ, NST = Cfg::NST, NMAIN = Cfg::NMAIN, NACC = Cfg::NACC;
    constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE = Cfg::STAGE, TMEM_COLS = Cfg::TMEM_COLS;
    constexpr uint32_t LBO_A = TBM * 16, LBO_B = BN * 16, SBO = 128;
    // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const ConvArgs& a = args.c;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * TBM;
    const int n0 = blockIdx.y * BN;
    const int taps = a.ksize * a.ksize;
    const int KC = a.Cin / BK;
    // split-K: this CTA runs stages [it0, it0 + T) of the taps * KC stages (host guarantees divisibility)
    const int SK = (int)cluster_nctaid_z(), sk_rank = (int)cluster_ctaid_z();
    const int T = taps * KC / SK;
    const int it0 = sk_rank * T;
    const int HW = a.H * a.W;
    // full_a: 256 producer arrivals; full_b: weight bulk copy (expect_tx); empty: tcgen05.commit
    const uint32_t bar_fa = smem_u32(&bars[0]), bar_fb = smem_u32(&bars[4]), bar_em = smem_u32(&bars[8]);
    if (tid == 0) HGK_STAMP(0);

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_fa + 8 * s, 256);
            mbar_init(bar_fb + 8 * s, 1);
            mbar_init(bar_em + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) HGK_STAMP(1);
    // programmatic dependent launch (common.cuh): no global access above this point
    pdl_wait();

    if (warp < 8) {
        // ===== producers (activation transform + weight bulk copies) =====
        // packed weights: [n-tile][tap][k] blocks in UMMA canonical layout; one stage = BN*BK contiguous floats
        const size_t wblk = (size_t)BN * BK;
        const float* wsrc_hi = args.w_hi + ((size_t)blockIdx.y * T * SK + it0) * wblk;
        const float* wsrc_lo = SPLIT ? args.w_lo + ((size_t)blockIdx.y * T * SK + it0) * wblk : nullptr;
        // thread -> NJ pixels x one 4-channel quad of the stage; a warp covers 8 pixels x 4 quads so that global
        // loads are full 32-byte sectors and shared stores are conflict-free 128-byte runs
        const int p_low = lane & 7, q_low = lane >> 3;
        constexpr int QH = Cfg::QP / 4;
        const int quad = (warp % QH) * 4 + q_low;
        // per-pixel state, computed once: element offset of the pixel and the 9-bit mask of taps whose
        // source pixel lies inside the image (padding = 1); the main loop only does shifts and adds
        int a_row[NJ];
        unsigned a_off[NJ], a_vm[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            a_row[j] = ((j * 8 + warp) / QH) * 8 + p_low;
            const long long p = m0 + a_row[j];
            const bool ok = p < a.P;
            const long long pp = ok ? p : 0;
            const int rem = (int)(pp % HW);
            const int h = rem / a.W, w = rem - h * a.W;
            unsigned vm = 0;
            if (ok) {
                if (a.ksize == 3) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int dh = t / 3 - 1, dw = t % 3 - 1;
                        if ((unsigned)(h + dh) < (unsigned)a.H && (unsigned)(w + dw) < (unsigned)a.W) vm |= 1u << t;
                    }
                } else {
                    vm = 1u;
                }
            }
            a_vm[j] = vm;
            a_off[j] = (unsigned)(pp * a.Cin) + quad * 4;
        }
        const float* xz = a.x.z;
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        // raw loads only (no dependent math): two register sets give two stages of latency cover
        float4 a_reg[2][NJ];
        unsigned a_msk[2];
        // running (tap, k-chunk) counters of the load stream (two stages ahead) and of the store stream
        int l_tap = it0 / KC, l_kc = it0 % KC;
        int l_toff = (a.ksize == 3) ? ((l_tap / 3 - 1) * a.W + (l_tap % 3 - 1)) * a.Cin : 0;
        int s_kc = it0 % KC;
        auto load_a = [&](int set) {
            const int coff = l_toff + l_kc * BK;
            unsigned m = 0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((a_vm[j] >> l_tap) & 1u) {
                    v = ldg4(xz + (a_off[j] + coff));
                    m |= 1u << j;
                }
                a_reg[set][j] = v;
            }
            a_msk[set] = m;
            if (++l_kc == KC) {            // next tap: offsets advance by one pixel, or by a row at the end of a tap row
                l_kc = 0;
                ++l_tap;
                l_toff += (l_tap == 3 || l_tap == 6) ? (a.W - 2) * a.Cin : a.Cin;
            }
        };
        auto store_a = [&](int s, int set) {
            float4 sc, sh;
            load_affine4(a.x.scale, a.x.shift, s_kc * BK + quad * 4, sc, sh);
            if (++s_kc == KC) s_kc = 0;
            uint8_t* base = sgen + s * STAGE + quad * LBO_A;
            const unsigned msk = a_msk[set];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float4 v = a_reg[set][j];
                // zero padding is a zero of the ACTIVATED tensor: transform only valid pixels
                if (has_aff && ((msk >> j) & 1u)) v = actc4(v, sc, sh, x_clamp);
                float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                *reinterpret_cast<float4*>(base + a_row[j] * 16) = hi;
                if (SPLIT) {
                    float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    *reinterpret_cast<float4*>(base + A_BYTES + a_row[j] * 16) = lo;
                }
            }
        };
        load_a(0);
        if (T > 1) load_a(1);
        if (tid == 0) HGK_STAMP(2);
        int s = 0;
        unsigned em_par = 1;               // parity of the previous use of the stage (toggles when s wraps)
        for (int it = 0; it < T; ++it) {
            if (it >= NST) mbar_wait(bar_em + 8 * s, em_par);          // stage s drained by the tensor core
            if (tid == 0) HGK_TRACE(0, it);
            if (tid == 0) {
                const uint32_t bb = bar_fb + 8 * s;
                const uint32_t dst = sbase + s * STAGE + (SPLIT ? 2 : 1) * A_BYTES;
                mbar_expect_tx(bb, (SPLIT ? 2 : 1) * B_BYTES);
                bulk_g2s(dst, wsrc_hi, B_BYTES, bb);
                if (SPLIT) bulk_g2s(dst + B_BYTES, wsrc_lo, B_BYTES, bb);
            }
            wsrc_hi += wblk;
            if (SPLIT) wsrc_lo += wblk;
            if (it & 1) store_a(s, 1); else store_a(s, 0);
            if (tid == 0) HGK_TRACE(6, it);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy (UMMA)
            mbar_arrive(bar_fa + 8 * s);
            if (tid == 0) HGK_TRACE(1, it);
            if (tid == 0 && it == 0) HGK_STAMP(3);
            if (it + 2 < T) { if (it & 1) load_a(1); else load_a(0); }
            if (++s == NST) { s = 0; em_par ^= 1u; }
        }
        if (tid == 0) HGK_STAMP(4);
        // all MMAs retired?
        mbar_wait(bar_em + 8 * ((T - 1) % NST), ((T - 1) / NST) & 1);
        if (tid == 0) HGK_STAMP(5);
    } else if (lane == 0) {
        // ===== MMA issuer =====
        for (int it = 0; it < T; ++it) {
            const int s = it % NST, u = it / NST;
            mbar_wait(bar_fa + 8 * s, u & 1);
            if (it == 0) HGK_STAMP(8);
            HGK_TRACE(2, it);
            mbar_wait(bar_fb + 8 * s, u & 1);
            if (it == 0) HGK_STAMP(9);
            HGK_TRACE(3, it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = sbase + s * STAGE;
            const uint32_t b_hi = a_hi + (SPLIT ? 2 : 1) * A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
                const uint64_t da = umma_desc(a_hi + k * 2 * LBO_A, LBO_A, SBO);
                const uint64_t db = umma_desc(b_hi + k * 2 * LBO_B, LBO_B, SBO);
                if (SPLIT) {
                    const uint64_t dal = umma_desc(a_hi + A_BYTES + k * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dbl = umma_desc(b_hi + B_BYTES + k * 2 * LBO_B, LBO_B, SBO);
                    if (NACC == 1) {
                        umma_tf32(tmem, dal, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem, da, dbl, IDESC, 1u);
                        umma_tf32(tmem, da, db, IDESC, 1u);
                    } else {
                        const uint32_t t_small = tmem + NMAIN * BN;
                        const uint32_t t_main = tmem + (NMAIN > 1 ? (it & 1) * BN : 0);
                        umma_tf32(t_small, dal, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(t_small, da, dbl, IDESC, 1u);
                        umma_tf32(t_main, da, db, IDESC, (it >= NMAIN || k > 0) ? 1u : 0u);
                    }
                } else {
                    umma_tf32(tmem, da, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                }
            }
            umma_commit(bar_em + 8 * s);
            HGK_TRACE(4, it);
        }
    }
    __syncwarp();      // the single-lane role reconverges before the CTA-wide barrier
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();       // every producer has observed the last commit: accumulator complete, smem reusable
    pdl_launch_dependents();   // main loop done: the next kernel of the chain may start its prologue under this epilogue

    // ---- epilogue, in column chunks of CH: TMEM -> registers -> staging tile -> coalesced
    //      bias / residual / accumulate / store + BN statistics ----
    constexpr int CH = Cfg::CH, SROW = CH + 4;
    constexpr int CG = CH / 4;           // float4 column groups of a chunk
    constexpr int RL = TNT / CG;         // row lanes (8 for CH=128, 16 for CH=64)
    constexpr int ROWS = TBM / RL;       // rows per thread (16 / 8)
    float* stg = reinterpret_cast<float*>(sgen);
    double* red = reinterpret_cast<double*>(sgen + TBM * SROW * 4);     // [RL][CH][2], behind the staging tile
    const bool epi = tid < TNT;          // the MMA warp only takes part in the barriers below
    const int cg = (epi ? tid : 0) % CG, r0 = (epi ? tid : 0) / CG;
    const bool do_stats = a.stat_sum != nullptr;
    const bool has_res = a.res.z != nullptr, res_aff = a.res.scale != nullptr;
    // split-K: this CTA finishes rows [row_base, row_base + TBM / SK) of the tile (rows_t per thread), summed over the SK
    // partial tiles of the cluster
    const int rows_t = ROWS / SK;
    const int row_base = sk_rank * (TBM / SK);
    const uint32_t stg_s = sbase;            // shared::cta address of the staging tile (same offset in every CTA of the cluster)
#pragma unroll 1
    for (int ch = 0; ch < BN / CH; ++ch) {
        if (warp < 8) {
            const int lq = warp & 3;
            const int row = lq * 32 + lane;
            const int cbeg = (warp >> 2) * (CH / 2);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + CH / 2; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * CH + c0);
                tmem_ld32(taddr, r);
                float acc[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(r[q]);
                // remaining accumulators (second hi*hi chain, cross terms): fp32 round-to-nearest adds.
                // With a single stage the second main accumulator was never written.
#pragma unroll
                for (int e = 1; e < NACC; ++e) {
                    if (NMAIN > 1 && e == 1 && T < 2) continue;
                    tmem_ld32(taddr + (uint32_t)(e * BN), r);
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[q] += __uint_as_float(r[q]);
                }
                float* dst = stg + row * SROW + c0;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    st4(dst + q * 4, make_float4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (SK > 1) cluster_sync_all();          // every partial tile of the cluster is staged
        if (ch == BN / CH - 1) {
            if (tid == 0) HGK_STAMP(6);
            if (warp == 0)
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        }
        const int n = n0 + ch * CH + cg * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias != nullptr) bv = ldg4(a.bias + n);
        float4 rs, rt;
        load_affine4(a.res.scale, a.res.shift, n, rs, rt);
        constexpr bool bwd_stats = BWDSTATS;      // separate instantiation: keeps the common kernels lean
        float4 bsc = make_float4(1.f, 1.f, 1.f, 1.f), bsh = make_float4(0.f, 0.f, 0.f, 0.f), bmu = bsh, biv = bsc;
        if (bwd_stats) { bsc = ldg4(a.bscale + n); bsh = ldg4(a.bshift + n); bmu = ldg4(a.bmean + n); biv = ldg4(a.binvstd + n); }
        double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
        float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
        // rows are processed 4 at a time with all global loads (residual / previous output / BN input) issued
        // first, so their latency is paid once per batch instead of once per row
#pragma unroll 1
        for (int g = 0; epi && g < rows_t; g += 4) {
            float4 rr[4], oo[4], zz[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long p = m0 + row_base + r0 + (g + i) * RL;
                rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                oo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                zz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g + i < rows_t && p < a.P) {
                    if (has_res) rr[i] = ldg4(a.res.z + p * a.Cout + n);
                    if (a.accumulate) oo[i] = ld4(a.y + p * a.Cout + n);
                    if (bwd_stats) zz[i] = ldg4(a.bz + p * a.Cout + n);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = row_base + r0 + (g + i) * RL;
                const long long p = m0 + r;
                if (g + i >= rows_t || p >= a.P) break;
                float4 v;
                if (SK == 1) {
                    v = ld4(stg + r * SROW + cg * 4);
                } else {                         // fixed-order sum of the cluster's partial tiles
                    const uint32_t off = stg_s + (uint32_t)(r * SROW + cg * 4) * 4u;
                    v = ld_dsmem4(dsmem_addr(off, 0));
                    for (int c = 1; c < SK; ++c) {
                        const float4 q = ld_dsmem4(dsmem_addr(off, (uint32_t)c));
                        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                    }
                }
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                if (has_res) {
                    float4 q = rr[i];
                    if (res_aff) q = act4(q, rs, rt, a.res.relu);
                    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                }
                v.x += oo[i].x; v.y += oo[i].y; v.z += oo[i].z; v.w += oo[i].w;
                st4(a.y + p * a.Cout + n, v);
                if (do_stats) {
                    if (bwd_stats) {
                        const float4 z = zz[i];
                        const float gx = (a.brelu && fmaf(z.x, bsc.x, bsh.x) <= 0.f) ? 0.f : v.x;
                        const float gy = (a.brelu && fmaf(z.y, bsc.y, bsh.y) <= 0.f) ? 0.f : v.y;
                        const float gz = (a.brelu && fmaf(z.z, bsc.z, bsh.z) <= 0.f) ? 0.f : v.z;
                        const float gw = (a.brelu && fmaf(z.w, bsc.w, bsh.w) <= 0.f) ? 0.f : v.w;
                        s1[0] += gx; s2[0] = fmaf(gx, (z.x - bmu.x) * biv.x, s2[0]);
                        s1[1] += gy; s2[1] = fmaf(gy, (z.y - bmu.y) * biv.y, s2[1]);
                        s1[2] += gz; s2[2] = fmaf(gz, (z.z - bmu.z) * biv.z, s2[2]);
                        s1[3] += gw; s2[3] = fmaf(gw, (z.w - bmu.w) * biv.w, s2[3]);
                    } else {
                        s1[0] += v.x; s2[0] = fmaf(v.x, v.x, s2[0]);
                        s1[1] += v.y; s2[1] = fmaf(v.y, v.y, s2[1]);
                        s1[2] += v.z; s2[2] = fmaf(v.z, v.z, s2[2]);
                        s1[3] += v.w; s2[3] = fmaf(v.w, v.w, s2[3]);
                    }
                }
            }
        }
        if (do_stats) {
            // fp32 partial sums over this thread's ROWS (<= 16) values, fp64 from here on (one conversion per chunk:
            // the F2F.F64 / DADD pair per 4 rows was ~40 % of the row loop, profiles/r2_tile_kernel_timeline.md)
#pragma unroll
            for (int j = 0; j < 4; ++j) { d1[j] += (double)s1[j]; d2[j] += (double)s2[j]; }
            if (epi) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    red[((r0 * CH) + cg * 4 + j) * 2 + 0] = d1[j];
                    red[((r0 * CH) + cg * 4 + j) * 2 + 1] = d2[j];
                }
            }
            __syncthreads();
            if (tid < CH) {
                double x1 = 0.0, x2 = 0.0;
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    x1 += red[((q * CH) + tid) * 2 + 0];
                    x2 += red[((q * CH) + tid) * 2 + 1];
                }
                atomicAdd(a.stat_sum + n0 + ch * CH + tid, x1);
                atomicAdd(a.stat_sq + n0 + ch * CH + tid, x2);
            }
        }
        // the staging tile is rewritten by the next chunk -- and, with split-K, read by the other CTAs of the cluster, which
        // must be done with it before it is rewritten or this CTA exits
        if (SK > 1) cluster_sync_all();
        else if (ch + 1 < BN / CH) __syncthreads();
    }
    if (tid == 0) HGK_STAMP(7);
    // fused BatchNorm finaliser: the CTA that arrives last turns the complete sums into per-channel vectors
    if (do_stats) {
        if (!BWDSTATS && a.ffin.ticket != nullptr) {
            if (last_cta_arrives(a.ffin.ticket, gridDim.x * gridDim.y * gridDim.z))
                bn_fwd_finalize_cta(a.ffin, a.stat_sum, a.stat_sq, (double)a.P, a.Cout);
        } else if (BWDSTATS && a.bfin.ticket != nullptr) {
            if (last_cta_arrives(a.bfin.ticket, gridDim.x * gridDim.y * gridDim.z))
                bn_bwd_finalize_cta(a.bfin, a.stat_sum, a.stat_sq, (double)a.P, a.Cout);
        }
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient on tcgen05:  dW[tap][co][ci] += sum_p dz[p][co] * T(x)[p+off(tap)][ci]
// GEMM with M = co (128-row tile), N = ci (= Cin), K = pixels.  tcgen05 takes 32-bit operands
// K-major only without swizzle, so every producer thread loads a 4-pixel x 4-channel block
// (four coalesced 128-bit loads), transposes it in registers and stores four 16-byte rows
// [channel][4 pixels] of the UMMA canonical K-major tile (LBO = rows*16+16 bytes between 4-pixel
// chunks -- the +16 spreads the chunks over the banks --, SBO = 128).
// Warp-specialised: warps 0-7 produce (BN+ReLU on x, TF32 rounding, smem stores, fence, mbarrier
// arrive, 2-stage register prefetch), warp 8 issues the MMAs and commits each stage back to the
// producers.  Split over pixel ranges (grid.z); partial tiles are reduced with coalesced vector
// atomics (red.global.add.v4.f32) into the [tap][Cout][Cin] destination.
// ------------------------------------------------------------------------------------------
struct WgTcArgs {
    Act x;
    int N, H, W, Cin;
    const float* dz;
    int Cout, ksize;
    float* dw;
    float* dbias;
    long long P, chunk;
};

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int WG_THREADS = 288;       // 8 producer/epilogue warps + 1 MMA warp
__host__ __device__ constexpr int wg_a_bytes() { return 8 * (TBM * 16 + 16); }
__host__ __device__ constexpr int wg_b_bytes(int BN) { return 8 * (BN * 16 + 16); }
__host__ __device__ constexpr int wg_stage_bytes(int BN) { return (wg_a_bytes() + wg_b_bytes(BN) + 127) / 128 * 128; }
// one CTA per SM, four stages (two CTAs per SM were tried: the 112-register cap spills the producers)
__host__ __device__ constexpr int wg_num_stages(int BN) { return 4; }
__host__ __device__ constexpr int wg_stg_bytes(int BN) { return TBM * ((BN > 128 ? 128 : BN) + 4) * 4; }
__host__ __device__ constexpr int wg_smem_bytes(int BN) {
    int p = wg_num_stages(BN) * wg_stage_bytes(BN);
    int s = wg_stg_bytes(BN);
    return (p > s ? p : s) + 256;
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const WgTcArgs a) {
    constexpr int NST = wg_num_stages(BN);
    constexpr int A_BYTES = wg_a_bytes(), STAGE = wg_stage_bytes(BN);
    constexpr uint32_t LBO_A = TBM * 16 + 16, LBO_B = BN * 16 + 16, SBO = 128;
    constexpr int NBLK = BN > 128 ? 2 : 1;      // 4x4 x-blocks per producer thread per stage
    // D=f32, A=B=tf32, both K-major, N>>3 at 17, M>>4 at 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NST];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int co0 = blockIdx.x * TBM, tap = blockIdx.y;
    const long long p_begin = (long long)blockIdx.z * a.chunk;
    const long long p_end = (p_begin + a.chunk < a.P) ? (p_begin + a.chunk) : a.P;
    const int T = (int)((p_end - p_begin + 31) / 32);      // host guarantees p_begin < P
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NST]);
    int dh = 0, dw_ = 0;
    if (a.ksize == 3) {
        dh = tap / 3 - 1;
        dw_ = tap - (tap / 3) * 3 - 1;
    }
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 256);
            mbar_init(bar_empty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp < 8) {
        // ===== producers: thread -> (4-pixel chunk c_low, channel quad) block of each operand =====
        const int c_low = lane & 7, q_low = lane >> 3;
        const int HW = a.H * a.W;
        const int a_quad = warp * 4 + q_low;              // 32 quads = 128 dz channels
        const int a_c = co0 + a_quad * 4;
        const bool a_cok = a_c < a.Cout;
        const bool b_on = (BN >= 128) || (warp < 4);      // BN = 64: 16 quads, warps 0-3 only
        int b_quad[NBLK];
        float4 xs[NBLK], xt[NBLK];
#pragma unroll
        for (int v = 0; v < NBLK; ++v) {
            b_quad[v] = (v * 8 + warp) * 4 + q_low;
            load_affine4(a.x.scale, a.x.shift, b_on ? b_quad[v] * 4 : 0, xs[v], xt[v]);
        }
        constexpr int PF = 2;                       // register sets = stages of load latency cover
        float4 ra[PF][4], rb[PF][NBLK][4];          // raw loads only: no dependent math until the stage is stored
        unsigned bmsk[PF];
#pragma unroll
        for (int i = 0; i < PF; ++i) bmsk[i] = 0u;
        float bsum[4] = {0.f, 0.f, 0.f, 0.f};
        const bool do_bias = a.dbias != nullptr && tap == 0;
        // running position of this thread's 4-pixel chunk in the load stream (two stages ahead of the stores)
        long long l_p = p_begin + c_low * 4;
        int l_h, l_w;
        {
            const int rem = (int)(l_p % HW);
            l_h = rem / a.W;
            l_w = rem - l_h * a.W;
        }
        const int toff = (dh * a.W + dw_) * a.Cin;
        const float* dzp = a.dz + a_c;
        const float* xzp = a.x.z;
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        auto load = [&](int set) {
            const unsigned pe = (unsigned)(p_end - l_p > 4 ? 4 : (p_end - l_p < 0 ? 0 : p_end - l_p));   // valid pixels of the chunk
            const unsigned o_dz = (unsigned)(l_p * a.Cout);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((unsigned)j < pe && a_cok) v = ldg4(dzp + (o_dz + j * a.Cout));
                ra[set][j] = v;
            }
            if (b_on) {
                int h = l_h, w = l_w;
                const unsigned o_x = (unsigned)(l_p * a.Cin + toff);
                unsigned m = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool ok = (unsigned)j < pe && (unsigned)(h + dh) < (unsigned)a.H && (unsigned)(w + dw_) < (unsigned)a.W;
                    if (ok) m |= 1u << j;
#pragma unroll
                    for (int v = 0; v < NBLK; ++v) {
                        float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) x4 = ldg4(xzp + (o_x + j * a.Cin + b_quad[v] * 4));
                        rb[set][v][j] = x4;
                    }
                    if (++w == a.W) {
                        w = 0;
                        if (++h == a.H) h = 0;
                    }
                }
                bmsk[set] = m;
            }
            l_p += 32;
            l_w += 32;
            while (l_w >= a.W) {
                l_w -= a.W;
                if (++l_h == a.H) l_h = 0;
            }
        };
        auto store = [&](int s, int set) {
            uint8_t* sa = sgen + s * STAGE + c_low * LBO_A + a_quad * 64;
            const float4 r0 = ra[set][0], r1 = ra[set][1], r2 = ra[set][2], r3 = ra[set][3];
            if (do_bias) {
                bsum[0] += (r0.x + r1.x) + (r2.x + r3.x);
                bsum[1] += (r0.y + r1.y) + (r2.y + r3.y);
                bsum[2] += (r0.z + r1.z) + (r2.z + r3.z);
                bsum[3] += (r0.w + r1.w) + (r2.w + r3.w);
            }
            // register transpose: row = channel, 4 consecutive pixels per 16-byte store
            *reinterpret_cast<float4*>(sa + 0) = tf32_rna4(make_float4(r0.x, r1.x, r2.x, r3.x));
            *reinterpret_cast<float4*>(sa + 16) = tf32_rna4(make_float4(r0.y, r1.y, r2.y, r3.y));
            *reinterpret_cast<float4*>(sa + 32) = tf32_rna4(make_float4(r0.z, r1.z, r2.z, r3.z));
            *reinterpret_cast<float4*>(sa + 48) = tf32_rna4(make_float4(r0.w, r1.w, r2.w, r3.w));
            if (b_on) {
#pragma unroll
                for (int v = 0; v < NBLK; ++v) {
                    uint8_t* sb = sgen + s * STAGE + A_BYTES + c_low * LBO_B + b_quad[v] * 64;
                    float4 x0 = rb[set][v][0], x1 = rb[set][v][1], x2 = rb[set][v][2], x3 = rb[set][v][3];
                    if (has_aff) {      // BN+ReLU on valid pixels only (padding = zero of the activated tensor)
                        const unsigned m = bmsk[set];
                        if (m & 1u) x0 = actc4(x0, xs[v], xt[v], x_clamp);
                        if (m & 2u) x1 = actc4(x1, xs[v], xt[v], x_clamp);
                        if (m & 4u) x2 = actc4(x2, xs[v], xt[v], x_clamp);
                        if (m & 8u) x3 = actc4(x3, xs[v], xt[v], x_clamp);
                    }
                    *reinterpret_cast<float4*>(sb + 0) = tf32_rna4(make_float4(x0.x, x1.x, x2.x, x3.x));
                    *reinterpret_cast<float4*>(sb + 16) = tf32_rna4(make_float4(x0.y, x1.y, x2.y, x3.y));
                    *reinterpret_cast<float4*>(sb + 32) = tf32_rna4(make_float4(x0.z, x1.z, x2.z, x3.z));
                    *reinterpret_cast<float4*>(sb + 48) = tf32_rna4(make_float4(x0.w, x1.w, x2.w, x3.w));
                }
            }
        };
        load(0);
        if (T > 1) load(1);
        int s = 0;
        unsigned em_par = 1;               // parity of the previous use of the stage (toggles when s wraps)
        for (int it = 0; it < T; ++it) {
            if (it >= NST) mbar_wait(bar_empty + 8 * s, em_par);
            if (it & 1) store(s, 1); else store(s, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_full + 8 * s);
            if (it + PF < T) { if (it & 1) load(1); else load(0); }
            if (++s == NST) { s = 0; em_par ^= 1u; }
        }
        if (do_bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = bsum[j];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (c_low == 0 && a_cok) atomicAdd(a.dbias + a_c + j, v);
            }
        }
        // all MMAs retired?
        mbar_wait(bar_empty + 8 * ((T - 1) % NST), ((T - 1) / NST) & 1);
    } else if (lane == 0) {
        // ===== MMA issuer =====
        for (int it = 0; it < T; ++it) {
            const int s = it % NST, u = it / NST;
            mbar_wait(bar_full + 8 * s, u & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = sbase + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_tf32(tmem, umma_desc(sa + k * 2 * LBO_A, LBO_A, SBO), umma_desc(sb + k * 2 * LBO_B, LBO_B, SBO), IDESC,
                          (it > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_empty + 8 * s);
        }
    }
    __syncwarp();      // the single-lane role reconverges before the CTA-wide barrier
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();       // producers have observed the last commit: accumulator complete, smem reusable

    // ---- epilogue, in column chunks of CH: TMEM -> staging tile -> coalesced vector atomics ----
    float* stg = reinterpret_cast<float*>(sgen);
    constexpr int CH = BN > 128 ? 128 : BN, SROW = CH + 4;
#pragma unroll 1
    for (int ch = 0; ch < BN / CH; ++ch) {
        if (warp < 8) {
            const int lq = warp & 3;
            const int row = lq * 32 + lane;
            const int cbeg = (warp >> 2) * (CH / 2);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + CH / 2; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * CH + c0), r);
                float* dst = stg + row * SROW + c0;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    st4(dst + q * 4, make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]),
                                                 __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3])));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (ch == BN / CH - 1 && warp == 0) {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN) : "memory");
        }
        if (tid < 256) {
            constexpr int CG = CH / 4, RL = 256 / CG;
            const int cg = tid % CG, r0 = tid / CG;
            for (int r = r0; r < TBM; r += RL) {
                const int co = co0 + r;
                if (co >= a.Cout) break;
                float4 v = ld4(stg + r * SROW + cg * 4);
                red_add_v4(a.dw + ((size_t)tap * a.Cout + co) * a.Cin + ch * CH + cg * 4, v);
            }
        }
        if (ch + 1 < BN / CH) __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// 3x3 weight gradient, three taps (one kernel row dh, dw = -1,0,+1) per CTA: the dz tile is staged once and
// shared by three MMAs, and the three column-shifted x tiles are built in registers from ONE 6-pixel load
// (pixels p-1 .. p+4 of the row shifted by dh).  Per tap this needs 3.3 instead of 8 global loads and a
// third of the barrier hand-shakes of wgrad_tc_kernel.  Requires W % 4 == 0 (4-pixel chunks never straddle
// image rows) and Cin <= 128 (3 accumulators of BN columns in TMEM).
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int wg3_stage_bytes(int BN) { return (wg_a_bytes() + 3 * wg_b_bytes(BN) + 127) / 128 * 128; }
__host__ __device__ constexpr int wg3_num_stages(int BN) { return BN > 64 ? 3 : 4; }
__host__ __device__ constexpr int wg3_smem_bytes(int BN) {
    int p = wg3_num_stages(BN) * wg3_stage_bytes(BN);
    int s = wg_stg_bytes(BN);
    return (p > s ? p : s) + 256;
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad3_tc_kernel(const WgTcArgs a) {
    constexpr int NST = wg3_num_stages(BN);
    constexpr int A_BYTES = wg_a_bytes(), B_BYTES = wg_b_bytes(BN), STAGE = wg3_stage_bytes(BN);
    constexpr uint32_t LBO_A = TBM * 16 + 16, LBO_B = BN * 16 + 16, SBO = 128;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    constexpr int TMEM_COLS = 3 * BN <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int co0 = blockIdx.x * TBM, dhg = blockIdx.y;          // dhg = tap row: taps 3*dhg .. 3*dhg+2
    const int dh = dhg - 1;
    const long long p_begin = (long long)blockIdx.z * a.chunk;
    const long long p_end = (p_begin + a.chunk < a.P) ? (p_begin + a.chunk) : a.P;
    const int T = (int)((p_end - p_begin + 31) / 32);
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[4]);
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 256);
            mbar_init(bar_empty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp < 8) {
        // ===== producers =====
        const int c_low = lane & 7, q_low = lane >> 3;
        const int HW = a.H * a.W;
        const int a_quad = warp * 4 + q_low;
        const int a_c = co0 + a_quad * 4;
        const bool a_cok = a_c < a.Cout;
        const bool b_on = (BN >= 128) || (warp < 4);
        const int b_quad = warp * 4 + q_low;
        float4 xs, xt;
        load_affine4(a.x.scale, a.x.shift, b_on ? b_quad * 4 : 0, xs, xt);
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        float4 ra[2][4], rb[2][6];
        unsigned bmsk[2] = {0u, 0u};
        float bsum[4] = {0.f, 0.f, 0.f, 0.f};
        const bool do_bias = a.dbias != nullptr && dhg == 0;
        long long l_p = p_begin + c_low * 4;
        int l_h, l_w;
        {
            const int rem = (int)(l_p % HW);
            l_h = rem / a.W;
            l_w = rem - l_h * a.W;
        }
        const float* dzp = a.dz + a_c;
        const float* xzp = a.x.z + b_quad * 4;
        auto load = [&](int set) {
            const bool pv = l_p < p_end;                              // chunks are all-valid or all-invalid (W % 4 == 0)
            const unsigned o_dz = (unsigned)(l_p * a.Cout);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pv && a_cok) v = ldg4(dzp + (o_dz + j * a.Cout));
                ra[set][j] = v;
            }
            unsigned m = 0;
            if (b_on) {
                const bool hv = pv && (unsigned)(l_h + dh) < (unsigned)a.H;
                const long long src0 = (l_p + (long long)dh * a.W - 1) * a.Cin;       // pixel p0-1 of the shifted row
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const bool ok = hv && (unsigned)(l_w - 1 + i) < (unsigned)a.W;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) {
                        v = ldg4(xzp + (src0 + (long long)i * a.Cin));
                        m |= 1u << i;
                    }
                    rb[set][i] = v;
                }
            }
            bmsk[set] = m;
            l_p += 32;
            l_w += 32;
            while (l_w >= a.W) {
                l_w -= a.W;
                if (++l_h == a.H) l_h = 0;
            }
        };
        auto store = [&](int s, int set) {
            uint8_t* sa = sgen + s * STAGE + c_low * LBO_A + a_quad * 64;
            const float4 r0 = ra[set][0], r1 = ra[set][1], r2 = ra[set][2], r3 = ra[set][3];
            if (do_bias) {
                bsum[0] += (r0.x + r1.x) + (r2.x + r3.x);
                bsum[1] += (r0.y + r1.y) + (r2.y + r3.y);
                bsum[2] += (r0.z + r1.z) + (r2.z + r3.z);
                bsum[3] += (r0.w + r1.w) + (r2.w + r3.w);
            }
            *reinterpret_cast<float4*>(sa + 0) = tf32_rna4(make_float4(r0.x, r1.x, r2.x, r3.x));
            *reinterpret_cast<float4*>(sa + 16) = tf32_rna4(make_float4(r0.y, r1.y, r2.y, r3.y));
            *reinterpret_cast<float4*>(sa + 32) = tf32_rna4(make_float4(r0.z, r1.z, r2.z, r3.z));
            *reinterpret_cast<float4*>(sa + 48) = tf32_rna4(make_float4(r0.w, r1.w, r2.w, r3.w));
            if (b_on) {
                float4 x[6];
                const unsigned m = bmsk[set];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    float4 v = rb[set][i];
                    if (has_aff && ((m >> i) & 1u)) v = actc4(v, xs, xt, x_clamp);     // padding = zero of the activated tensor
                    x[i] = tf32_rna4(v);
                }
#pragma unroll
                for (int t = 0; t < 3; ++t) {                     // tap dw = t-1: output pixel j reads source pixel j+t
                    uint8_t* sb = sgen + s * STAGE + A_BYTES + t * B_BYTES + c_low * LBO_B + b_quad * 64;
                    *reinterpret_cast<float4*>(sb + 0) = make_float4(x[t].x, x[t + 1].x, x[t + 2].x, x[t + 3].x);
                    *reinterpret_cast<float4*>(sb + 16) = make_float4(x[t].y, x[t + 1].y, x[t + 2].y, x[t + 3].y);
                    *reinterpret_cast<float4*>(sb + 32) = make_float4(x[t].z, x[t + 1].z, x[t + 2].z, x[t + 3].z);
                    *reinterpret_cast<float4*>(sb + 48) = make_float4(x[t].w, x[t + 1].w, x[t + 2].w, x[t + 3].w);
                }
            }
        };
        load(0);
        if (T > 1) load(1);
        int s = 0;
        unsigned em_par = 1;
        for (int it = 0; it < T; ++it) {
            if (it >= NST) mbar_wait(bar_empty + 8 * s, em_par);
            if (it & 1) store(s, 1); else store(s, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_full + 8 * s);
            if (it + 2 < T) { if (it & 1) load(1); else load(0); }
            if (++s == NST) { s = 0; em_par ^= 1u; }
        }
        if (do_bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = bsum[j];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (c_low == 0 && a_cok) atomicAdd(a.dbias + a_c + j, v);
            }
        }
        mbar_wait(bar_empty + 8 * ((T - 1) % NST), ((T - 1) / NST) & 1);
    } else if (lane == 0) {
        // ===== MMA issuer: 3 taps x 4 k-steps per stage =====
        for (int it = 0; it < T; ++it) {
            const int s = it % NST, u = it / NST;
            mbar_wait(bar_full + 8 * s, u & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = sbase + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc(sa + k * 2 * LBO_A, LBO_A, SBO);
#pragma unroll
                for (int t = 0; t < 3; ++t)
                    umma_tf32(tmem + t * BN, da, umma_desc(sb + t * B_BYTES + k * 2 * LBO_B, LBO_B, SBO), IDESC,
                              (it > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * s);
        }
    }
    __syncwarp();      // the single-lane role reconverges before the CTA-wide barrier
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- epilogue: per tap, TMEM -> staging tile -> coalesced vector atomics ----
    float* stg = reinterpret_cast<float*>(sgen);
    constexpr int SROW = BN + 4;
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
        if (warp < 8) {
            const int lq = warp & 3;
            const int row = lq * 32 + lane;
            const int cbeg = (warp >> 2) * (BN / 2);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(t * BN + c0), r);
                float* dst = stg + row * SROW + c0;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    st4(dst + q * 4, make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]),
                                                 __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3])));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 2 && warp == 0) {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        }
        if (tid < 256) {
            constexpr int CG = BN / 4, RL = 256 / CG;
            const int cg = tid % CG, r0 = tid / CG;
            const int tap = dhg * 3 + t;
            for (int r = r0; r < TBM; r += RL) {
                const int co = co0 + r;
                if (co >= a.Cout) break;
                float4 v = ld4(stg + r * SROW + cg * 4);
                red_add_v4(a.dw + ((size_t)tap * a.Cout + co) * a.Cin + cg * 4, v);
            }
        }
        if (t < 2) __syncthreads();
    }
}

template <int BN>
static int launch_wg3(const WgTcArgs& a, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    constexpr int smem = wg3_smem_bytes(BN);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad3_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_wgrad_tc_nhwc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    wgrad3_tc_kernel<BN><<<grid, WG_THREADS, smem, st>>>(a);
    return HGK_OK;
}

// grads of 3x3 convs are accumulated tap-major ([tap][O][I]) by the tensor-core kernel; this adds them
// into the OIHW-shaped .grad views.  table rows (5 x int64): {src_off, dst_off, O, I, taps}
__global__ void unpack_add_grads_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                        const long long* __restrict__ table, int n_entries) {
    const int e = blockIdx.y;
    if (e >= n_entries) return;
    const long long* t = table + (size_t)e * 5;
    const long long so = t[0], d0 = t[1];
    const int O = (int)t[2], I = (int)t[3], taps = (int)t[4];
    const long long total = (long long)O * I * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long tap = i % taps, oi = i / taps;           // destination order (OIHW): stores coalesced
        dst[d0 + i] += __ldg(src + so + tap * (long long)O * I + oi);
    }
}

template <int BN>
static int launch_wg(const WgTcArgs& a, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    constexpr int smem = wg_smem_bytes(BN);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_wgrad_tc_nhwc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    wgrad_tc_kernel<BN><<<grid, WG_THREADS, smem, st>>>(a);
    return HGK_OK;
}

// ------------------------------------------------------------------------------------------
// weight packing into the UMMA operand layout (+ hi/lo TF32 split)
// table rows (8 x int64): {src_off, dst_hi_off, dst_lo_off (-1: none), N, K, taps, mode, BN}
//   mode 0 (forward) : B[n][k; tap] = W[o=n][i=k][tap]              (OIHW source, O=N, I=K)
//   mode 1 (dgrad)   : B[n][k; tap] = W[o=k][i=n][taps-1-tap]       (O=K, I=N; taps pre-flipped)
//   mode 2 (forward, TF32 + 2xBF16 products): hi as mode 0; the "lo" buffer holds, per 16-channel chunk (the footprint of the
//                      fp32 lo half-block: BN x 64 bytes), the two bf16 cross-term operands in the K-major no-swizzle
//                      core-matrix layout [k/8 (2)][n (BN)][8 x bf16]: first bf16(wh), then bf16(w - wh)
// destination: [n-tile][tap][k/32] blocks, each block [quad(8)][n(BN)][4]
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_tc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                       const long long* __restrict__ table, int n_entries) {
    const int e = blockIdx.y;
    if (e >= n_entries) return;
    const long long* t = table + (size_t)e * 8;
    const long long so = t[0], dhi = t[1], dlo = t[2];
    const int N = (int)t[3], K = (int)t[4], taps = (int)t[5], mode = (int)t[6], BN = (int)t[7];
    // 32-bit index arithmetic (a weight tensor has at most a few hundred thousand elements; the five 64-bit divisions per
    // element of the first version made this launch compute-bound: 85 us alone at the head of every step)
    const unsigned total = (unsigned)N * (unsigned)K * (unsigned)taps;
    const unsigned KC = (unsigned)K / 32u, uBN = (unsigned)BN, utaps = (unsigned)taps;
    const float* sp = src + so;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned el = i & 3u;
        unsigned r = i >> 2;
        const unsigned nn = r % uBN;
        r /= uBN;
        const unsigned q = r & 7u;
        unsigned blk = r >> 3;
        const unsigned kc = blk % KC;
        blk /= KC;
        const unsigned tap = blk % utaps;
        const unsigned ntile = blk / utaps;
        const unsigned n = ntile * uBN + nn, k = kc * 32u + q * 4u + el;
        float v;
        if (mode != 1) v = __ldg(sp + (n * (unsigned)K + k) * utaps + tap);
        else v = __ldg(sp + (k * (unsigned)N + n) * utaps + (utaps - 1u - tap));
        const float hi = tf32_rna(v);
        dst[dhi + i] = hi;
        if (dlo >= 0) {
            if (mode == 2) {
                // half-block of this 16-channel chunk (in floats from the lo base), then bf16 positions inside it
                const unsigned hb = ((ntile * utaps + tap) * KC + kc) * uBN * 32u + (q >> 2) * uBN * 16u;
                const unsigned k16 = (q & 3u) * 4u + el;                 // channel inside the chunk
                const unsigned pos = (k16 >> 3) * uBN * 8u + nn * 8u + (k16 & 7u);     // [k/8][n][8]
                __nv_bfloat16* lb = reinterpret_cast<__nv_bfloat16*>(dst + dlo + hb);
                lb[pos] = __float2bfloat16_rn(hi);
                lb[uBN * 16u + pos] = __float2bfloat16_rn(v - hi);
            } else {
                dst[dlo + i] = v - hi;
            }
        }
    }
}

template <int BN, bool SPLIT, bool BWDSTATS, bool BIG>
static int launch_tc_cfg(const TcArgs& ta, cudaStream_t st, int sk = 1) {
    static bool configured = false;
    constexpr int smem = TcCfg<BN, SPLIT, BIG>::SMEM;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, SPLIT, BWDSTATS, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_tc_nhwc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    long long mt = (ta.c.P + TBM - 1) / TBM;
    dim3 grid((unsigned)mt, (unsigned)(ta.c.Cout / BN), (unsigned)sk);
    // the sk CTAs along grid.z form one thread-block cluster (split-K: partial tiles are summed through distributed shared
    // memory); programmatic dependent launch as everywhere on the chain (common.cuh)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TNT + 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 1; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = (unsigned)(sk > 1 ? sk : 1);
    cfg.attrs = at;
    cfg.numAttrs = sk > 1 ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, SPLIT, BWDSTATS, BIG>, ta);
    if (e != cudaSuccess) {
        set_error("hgk_conv_tc_nhwc: launch (split-K %d): %s", sk, cudaGetErrorString(e));
        return HGK_ECUDA;
    }
    return HGK_OK;
}

// Split-K factor for a small layer (0 = not a split-K shape): the largest cluster size in {8, 4, 2} that divides the stage
// count of the one-CTA-per-SM configuration (32-channel stages) and keeps the whole grid in one wave.  HGK_SPLITK=0 disables
// it, HGK_SPLITK_MT sets the largest layer (in 128-pixel tiles) that takes this path (default 12: the 8x8 and 4x4 rungs at batch 24;
// measured 10.66 ms/step against 10.73 with 24, 10.80 with 48 and 11.26 without split-K, gpurun_out/r4c_ab.txt).
static int splitk_factor(long long P, int Cin, int Cout, int ksize) {
    static int on = -1, max_mt = 12;
    if (on < 0) {
        const char* e = getenv("HGK_SPLITK");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
        const char* m = getenv("HGK_SPLITK_MT");
        if (m != nullptr) max_mt = atoi(m);
    }
    if (!on || (Cout != 128 && Cout != 256) || Cin % 32) return 0;
    const long long mt = (P + TBM - 1) / TBM;
    if (mt > max_mt) return 0;
    const int stages = ksize * ksize * (Cin / 32);
    for (int sk = 8; sk >= 2; sk >>= 1)
        if (stages % sk == 0 && mt * sk <= kNumSMs) return sk;
    return 0;
}

template <int BN, bool SPLIT, bool BWDSTATS = false>
static int launch_tc(const TcArgs& ta, cudaStream_t st) {
    // at most one CTA per SM anyway -> the large-stage single-CTA configuration
    const long long ctas = ((ta.c.P + TBM - 1) / TBM) * (ta.c.Cout / BN);
    static int big_off = -1;              // HGK_TC_BIG_OFF=1: always the two-CTAs-per-SM configuration (tuning knob)
    if (big_off < 0) {
        const char* e = getenv("HGK_TC_BIG_OFF");
        big_off = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (ctas <= kNumSMs && !big_off) {
        const int sk = (BN == ta.c.Cout) ? splitk_factor(ta.c.P, ta.c.Cin, ta.c.Cout, ta.c.ksize) : 0;
        return launch_tc_cfg<BN, SPLIT, BWDSTATS, true>(ta, st, sk > 1 ? sk : 1);
    }
    return launch_tc_cfg<BN, SPLIT, BWDSTATS, false>(ta, st);
}

}  // namespace hgk

namespace hgk {
bool conv_tc2_eligible(const TcArgs& ta);                             // conv_tc2.cu (16x16 image-tile kernel)
int wgrad_tc2_try(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W, int Cin,
                  const float* dz, int Cout, int ksize, float* dw, float* dbias, void* stream);   // wgrad_tc2.cu (MN-major)
int conv_tc2_launch(const TcArgs& ta, bool split, bool bwdstats, void* stream);
bool conv_tc3_eligible(const TcArgs& ta);                             // conv_tc3.cu (persistent, phase-overlapped tile kernel)
int conv_tc3_launch(const TcArgs& ta, bool split, bool bwdstats, void* stream);
}

using namespace hgk;

extern "C" int hgk_conv_tc_supported(int Cin, int Cout, int ksize) {
    return (Cin > 0 && Cin % 32 == 0 && (Cout == 64 || Cout == 128 || Cout == 256) && (ksize == 1 || ksize == 3)) ? 1 : 0;
}

// HGK_TC2_OFF=1 disables the image-tile kernel of conv_tc2.cu (every shape then takes conv_tc_kernel)
static bool use_tile_kernel() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("HGK_TC2_OFF");
        v = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// Which layer classes take the persistent kernel of conv_tc3.cu: bit 0 forward 1x1, bit 1 forward 3x3, bit 2 data gradient
// 1x1, bit 3 data gradient 3x3.  Default 13: the 3x3 forward (3xTF32, tensor-bound) is faster on the two-CTAs-per-SM kernel
// of conv_tc2.cu (profiles/r3_persistent_kernel.md).  HGK_TC3_MASK overrides (0 = conv_tc2.cu everywhere; A/B comparisons).
static int tc3_mask() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("HGK_TC3_MASK");
        v = e != nullptr ? atoi(e) : 13;
    }
    return v;
}
static bool use_tc3(int ksize, bool split) {
    const int bit = (split ? 0 : 2) + (ksize == 3 ? 1 : 0);
    return (tc3_mask() >> bit) & 1;
}

extern "C" int hgk_conv_tc_x2_supported(int N, int H, int W, int Cin, int Cout, int ksize);

static int conv_tc_impl(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                        int N, int H, int W, int Cin,
                        const float* w_hi, const float* w_lo, int ksize, const float* bias, int Cout,
                        const float* res, const float* res_scale, const float* res_shift, int res_relu,
                        float* y, int accumulate, double* stat_sum, double* stat_sq,
                        const float* bz, const float* bscale, const float* bshift, const float* bmean,
                        const float* binvstd, int brelu, const BnFwdFin* ffin, const BnBwdFin* bfin, void* stream,
                        const BnApply* ap = nullptr, int lo_bf16 = 0) {
    HGK_REQUIRE(x && w_hi && y, "hgk_conv_tc_nhwc: null pointer");
    HGK_REQUIRE(N > 0 && H > 0 && W > 0, "hgk_conv_tc_nhwc: empty tensor");
    // k = 4: the 7x7 stride-2 stem convolution in space-to-depth form (hgk_stem_s2d_image / hgk_stem_s2d_weight, stem.cu):
    // 4x4 taps at offsets -2 .. +1 over the 16-channel half-resolution image (weights packed as K = 32, upper half zero),
    // 64 outputs, forward only, image-tile kernel
    const bool stem4 = ksize == 4;
    if (stem4)
        HGK_REQUIRE(Cin == 16 && Cout == 64 && H % 16 == 0 && W % 16 == 0 && w_lo != nullptr && ap == nullptr && bz == nullptr &&
                    res == nullptr && !accumulate && (long long)N * H * W * 64 < (1LL << 32),
                    "hgk_conv_tc_nhwc: k = 4 (space-to-depth stem) needs Cin = 16, Cout = 64, H and W multiples of 16, w_lo");
    else
        HGK_REQUIRE(hgk_conv_tc_supported(Cin, Cout, ksize), "hgk_conv_tc_nhwc: unsupported shape Cin=%d Cout=%d k=%d "
                    "(need Cin %% 32 == 0, Cout in {64,128,256}, k in {1,3})", Cin, Cout, ksize);
    HGK_REQUIRE((stat_sum == nullptr) == (stat_sq == nullptr), "hgk_conv_tc_nhwc: stat_sum/stat_sq must both be set");
    HGK_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), "hgk_conv_tc_nhwc: x scale/shift must both be set");
    HGK_REQUIRE((res_scale == nullptr) == (res_shift == nullptr), "hgk_conv_tc_nhwc: res scale/shift must both be set");
    HGK_REQUIRE(((uintptr_t)w_hi % 16 == 0) && ((uintptr_t)w_lo % 16 == 0), "hgk_conv_tc_nhwc: packed weights must be 16-byte aligned");
    TcArgs ta;
    ta.c.x = Act{x, x_scale, x_shift, x_relu};
    ta.c.N = N; ta.c.H = H; ta.c.W = W; ta.c.Cin = Cin;
    ta.c.w = nullptr; ta.c.ksize = ksize; ta.c.flip = 0; ta.c.bias = bias; ta.c.Cout = Cout;
    ta.c.res = Act{res, res_scale, res_shift, res_relu};
    ta.c.y = y; ta.c.accumulate = accumulate; ta.c.stat_sum = stat_sum; ta.c.stat_sq = stat_sq;
    ta.c.P = (long long)N * H * W;
    ta.c.bz = bz; ta.c.bscale = bscale; ta.c.bshift = bshift; ta.c.bmean = bmean; ta.c.binvstd = binvstd; ta.c.brelu = brelu;
    ta.c.ffin = ffin != nullptr ? *ffin : BnFwdFin{};
    ta.c.bfin = bfin != nullptr ? *bfin : BnBwdFin{};
    ta.c.ap = ap != nullptr ? *ap : BnApply{};
    if (ap != nullptr)
        HGK_REQUIRE(w_lo == nullptr && ((use_tc3(ksize, false) && conv_tc3_eligible(ta)) || (use_tile_kernel() && conv_tc2_eligible(ta))),
                    "hgk_conv_tc_dgrad_bnapply_nhwc: shape not covered by the "
                    "image-tile kernel (see hgk_conv_tc_bnapply_supported)");
    ta.w_hi = w_hi; ta.w_lo = w_lo; ta.dbg = g_dbg_buf; ta.lo_bf16 = lo_bf16;
    if (lo_bf16 && !stem4)
        HGK_REQUIRE(w_lo != nullptr && ap == nullptr && hgk_conv_tc_x2_supported(N, H, W, Cin, Cout, ksize),
                    "hgk_conv_tc_bn_x2_nhwc: shape not covered by the TF32 + 2xBF16 tile kernel (see hgk_conv_tc_x2_supported)");
    HGK_REQUIRE((ta.c.P + TBM - 1) / TBM < 2147483647LL, "hgk_conv_tc_nhwc: too many pixels");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const bool split = w_lo != nullptr;
    if (bz != nullptr)
        HGK_REQUIRE(!split, "hgk_conv_tc_dgrad_bnstats_nhwc: only plain-TF32 data gradients carry the fused BN reduction");
    const bool small = ap == nullptr && splitk_factor(ta.c.P, Cin, Cout, ksize) > 1;      // cluster split-K (conv_tc_kernel)
    if (stem4) rc = conv_tc2_launch(ta, split, false, stream);
    else if (!small && use_tc3(ksize, split) && conv_tc3_eligible(ta)) rc = conv_tc3_launch(ta, split, bz != nullptr, stream);
    else if (!small && use_tile_kernel() && conv_tc2_eligible(ta)) rc = conv_tc2_launch(ta, split, bz != nullptr, stream);
    else if (bz != nullptr) {
        rc = Cout == 64 ? launch_tc<64, false, true>(ta, st)
                        : (Cout == 128 ? launch_tc<128, false, true>(ta, st) : launch_tc<256, false, true>(ta, st));
    } else if (Cout == 64) rc = split ? launch_tc<64, true>(ta, st) : launch_tc<64, false>(ta, st);
    else if (Cout == 128) rc = split ? launch_tc<128, true>(ta, st) : launch_tc<128, false>(ta, st);
    else rc = split ? launch_tc<256, true>(ta, st) : launch_tc<256, false>(ta, st);
    if (rc != HGK_OK) return rc;
    HGK_CHECK_LAUNCH("hgk_conv_tc_nhwc");
    return HGK_OK;
}

extern "C" int hgk_conv_tc_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                                int N, int H, int W, int Cin,
                                const float* w_hi, const float* w_lo, int ksize, const float* bias, int Cout,
                                const float* res, const float* res_scale, const float* res_shift, int res_relu,
                                float* y, int accumulate, double* stat_sum, double* stat_sq, void* stream) {
    return conv_tc_impl(x, x_scale, x_shift, x_relu, N, H, W, Cin, w_hi, w_lo, ksize, bias, Cout, res, res_scale, res_shift,
                        res_relu, y, accumulate, stat_sum, stat_sq, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
                        stream);
}

extern "C" int hgk_conv_tc_bn_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                                   int N, int H, int W, int Cin,
                                   const float* w_hi, const float* w_lo, int ksize, const float* bias, int Cout,
                                   const float* res, const float* res_scale, const float* res_shift, int res_relu,
                                   float* y, int accumulate, double* stat_sum, double* stat_sq,
                                   const float* gamma, const float* beta, float eps, float momentum,
                                   float* running_mean, float* running_var, float* scale, float* shift,
                                   float* save_mean, float* save_invstd, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(stat_sum && stat_sq && gamma && beta && scale && shift && save_mean && save_invstd && ticket,
                "hgk_conv_tc_bn_nhwc: null pointer");
    HGK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "hgk_conv_tc_bn_nhwc: running stats must both be set");
    BnFwdFin f{gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, ticket, eps, momentum};
    return conv_tc_impl(x, x_scale, x_shift, x_relu, N, H, W, Cin, w_hi, w_lo, ksize, bias, Cout, res, res_scale, res_shift,
                        res_relu, y, accumulate, stat_sum, stat_sq, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &f, nullptr,
                        stream);
}

// 1 when a forward convolution of this shape runs with TF32 + 2xBF16 products: 3x3 on the image-tile kernel (H and W multiples of
// 16, 64 / 128 output channels), 1x1 on the persistent kernel (128 / 256 output channels); never one of the small split-K
// layers.  HGK_X2=0 disables it.
extern "C" int hgk_conv_tc_x2_supported(int N, int H, int W, int Cin, int Cout, int ksize) {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("HGK_X2");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || !hgk_conv_tc_supported(Cin, Cout, ksize) || N <= 0 || H <= 0 || W <= 0) return 0;
    TcArgs ta{};
    ta.c.N = N; ta.c.H = H; ta.c.W = W; ta.c.Cin = Cin; ta.c.Cout = Cout; ta.c.ksize = ksize;
    ta.c.P = (long long)N * H * W;
    if (splitk_factor(ta.c.P, Cin, Cout, ksize) > 1) return 0;
    static int mask = -1;   // developer bisection: HGK_X2_MASK bit 0: 3x3, bit 1: 1x1 -> 128, bit 2: 1x1 -> 256, bit 3: 1x1 with Cin == 64
    if (mask < 0) {
        const char* e = getenv("HGK_X2_MASK");
        mask = e != nullptr ? atoi(e) : 15;
    }
    if (ksize == 3 && !(mask & 1)) return 0;
    if (ksize == 1 && Cout == 128 && !(mask & 2)) return 0;
    if (ksize == 1 && Cout == 256 && !(mask & 4)) return 0;
    if (ksize == 1 && Cin == 64 && !(mask & 8)) return 0;
    if (ksize == 1)         // 1x1: the persistent kernel (conv_tc3.cu) when it takes the layer
        return (use_tc3(1, true) && conv_tc3_eligible(ta)) ? 1 : 0;
    if (Cout != 64 && Cout != 128) return 0;
    if (use_tc3(ksize, true) && conv_tc3_eligible(ta)) return 0;
    return (use_tile_kernel() && conv_tc2_eligible(ta)) ? 1 : 0;
}

extern "C" int hgk_conv_tc_bn_x2_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                                      int N, int H, int W, int Cin,
                                      const float* w_hi, const float* w_x2, int ksize, const float* bias, int Cout,
                                      const float* res, const float* res_scale, const float* res_shift, int res_relu,
                                      float* y, int accumulate, double* stat_sum, double* stat_sq,
                                      const float* gamma, const float* beta, float eps, float momentum,
                                      float* running_mean, float* running_var, float* scale, float* shift,
                                      float* save_mean, float* save_invstd, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(stat_sum && stat_sq && gamma && beta && scale && shift && save_mean && save_invstd && ticket && w_x2,
                "hgk_conv_tc_bn_x2_nhwc: null pointer");
    HGK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "hgk_conv_tc_bn_x2_nhwc: running stats must both be set");
    BnFwdFin f{gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, ticket, eps, momentum};
    return conv_tc_impl(x, x_scale, x_shift, x_relu, N, H, W, Cin, w_hi, w_x2, ksize, bias, Cout, res, res_scale, res_shift,
                        res_relu, y, accumulate, stat_sum, stat_sq, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &f, nullptr,
                        stream, nullptr, 1);
}

extern "C" int hgk_conv_tc_x2_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                                   int N, int H, int W, int Cin,
                                   const float* w_hi, const float* w_x2, int ksize, const float* bias, int Cout,
                                   const float* res, const float* res_scale, const float* res_shift, int res_relu,
                                   float* y, int accumulate, double* stat_sum, double* stat_sq, void* stream) {
    HGK_REQUIRE(w_x2 != nullptr, "hgk_conv_tc_x2_nhwc: null pointer");
    return conv_tc_impl(x, x_scale, x_shift, x_relu, N, H, W, Cin, w_hi, w_x2, ksize, bias, Cout, res, res_scale, res_shift,
                        res_relu, y, accumulate, stat_sum, stat_sq, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
                        stream, nullptr, 1);
}

extern "C" int hgk_conv_tc_dgrad_bnstats_nhwc(const float* dz, int N, int H, int W, int Cin,
                                              const float* w_hi, const float* w_lo, int ksize, int Cout,
                                              const float* extra, float* dy, int accumulate,
                                              const float* bz, const float* bscale, const float* bshift, int brelu,
                                              const float* bmean, const float* binvstd,
                                              double* sum_g, double* sum_gx, void* stream) {
    HGK_REQUIRE(bz && bscale && bshift && bmean && binvstd && sum_g && sum_gx, "hgk_conv_tc_dgrad_bnstats_nhwc: null pointer");
    return conv_tc_impl(dz, nullptr, nullptr, 0, N, H, W, Cin, w_hi, w_lo, ksize, nullptr, Cout, extra, nullptr, nullptr, 0,
                        dy, accumulate, sum_g, sum_gx, bz, bscale, bshift, bmean, binvstd, brelu, nullptr, nullptr, stream);
}

extern "C" int hgk_conv_tc_dgrad_bnfin_nhwc(const float* dz, int N, int H, int W, int Cin,
                                            const float* w_hi, const float* w_lo, int ksize, int Cout,
                                            const float* extra, float* dy, int accumulate,
                                            const float* bz, const float* bscale, const float* bshift, int brelu,
                                            const float* bmean, const float* binvstd,
                                            double* sum_g, double* sum_gx,
                                            const float* gamma, int training, float* dgamma, float* dbeta,
                                            float* cA, float* cB, float* cC, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(bz && bscale && bshift && bmean && binvstd && sum_g && sum_gx && gamma && cA && cB && cC && ticket,
                "hgk_conv_tc_dgrad_bnfin_nhwc: null pointer");
    BnBwdFin f{gamma, bmean, binvstd, dgamma, dbeta, cA, cB, cC, ticket, training};
    return conv_tc_impl(dz, nullptr, nullptr, 0, N, H, W, Cin, w_hi, w_lo, ksize, nullptr, Cout, extra, nullptr, nullptr, 0,
                        dy, accumulate, sum_g, sum_gx, bz, bscale, bshift, bmean, binvstd, brelu, nullptr, &f, stream);
}

extern "C" int hgk_conv_tc_bnapply_supported(int N, int H, int W, int Cin, int Cout, int ksize) {
    if (!hgk_conv_tc_supported(Cin, Cout, ksize) || N <= 0 || H <= 0 || W <= 0) return 0;
    TcArgs ta{};
    ta.c.N = N; ta.c.H = H; ta.c.W = W; ta.c.Cin = Cin; ta.c.Cout = Cout; ta.c.ksize = ksize;
    ta.c.P = (long long)N * H * W;
    // small layers take the cluster split-K kernel, which has no apply-on-load: their BatchNorm-backward apply is a launch of
    // its own (a few microseconds at that size)
    if (splitk_factor(ta.c.P, Cin, Cout, ksize) > 1) return 0;
    if (use_tc3(ksize, false) && conv_tc3_eligible(ta)) return 1;
    return (use_tile_kernel() && conv_tc2_eligible(ta)) ? 1 : 0;
}

extern "C" int hgk_conv_tc_dgrad_bnapply_nhwc(const float* g, const float* gz, const float* gscale, const float* gshift, int grelu,
                                              const float* gmean, const float* gcA, const float* gcB, const float* gcC,
                                              float* dz_out, int N, int H, int W, int Cin,
                                              const float* w_hi, int ksize, int Cout,
                                              const float* extra, float* dy, int accumulate,
                                              const float* bz, const float* bscale, const float* bshift, int brelu,
                                              const float* bmean, const float* binvstd, double* sum_g, double* sum_gx,
                                              const float* gamma, int training, float* dgamma, float* dbeta,
                                              float* cA, float* cB, float* cC, unsigned int* ticket, void* stream) {
    HGK_REQUIRE(g && gz && gscale && gshift && gmean && gcA && gcB && gcC && dz_out, "hgk_conv_tc_dgrad_bnapply_nhwc: null pointer");
    BnApply ap{gz, gscale, gshift, gmean, gcA, gcB, gcC, dz_out, grelu};
    if (bz == nullptr)
        return conv_tc_impl(g, nullptr, nullptr, 0, N, H, W, Cin, w_hi, nullptr, ksize, nullptr, Cout, extra, nullptr, nullptr, 0,
                            dy, accumulate, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
                            stream, &ap);
    HGK_REQUIRE(bscale && bshift && bmean && binvstd && sum_g && sum_gx && gamma && cA && cB && cC && ticket,
                "hgk_conv_tc_dgrad_bnapply_nhwc: null pointer (fused reduction)");
    BnBwdFin f{gamma, bmean, binvstd, dgamma, dbeta, cA, cB, cC, ticket, training};
    return conv_tc_impl(g, nullptr, nullptr, 0, N, H, W, Cin, w_hi, nullptr, ksize, nullptr, Cout, extra, nullptr, nullptr, 0,
                        dy, accumulate, sum_g, sum_gx, bz, bscale, bshift, bmean, binvstd, brelu, nullptr, &f, stream, &ap);
}

/* developer diagnostics: per-CTA globaltimer stamps of conv_tc_kernel ([512][16] int64), NULL disables */
extern "C" int hgk_debug_set_timeline(long long* buf) {
    g_dbg_buf = buf;
    return HGK_OK;
}

extern "C" int hgk_pack_weights_tc(const float* src_base, float* dst_base, const long long* table, int n_entries,
                                   void* stream) {
    HGK_REQUIRE(src_base && dst_base && table, "hgk_pack_weights_tc: null pointer");
    if (n_entries <= 0) return HGK_OK;
    HGK_REQUIRE(n_entries <= 65535, "hgk_pack_weights_tc: too many entries");
    dim3 grid(72, (unsigned)n_entries);      // (16 blocks per entry: 85 us, latency-bound scattered 4-byte reads, alone at the head of the step)
    pack_weights_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_base, dst_base, table, n_entries);
    HGK_CHECK_LAUNCH("hgk_pack_weights_tc");
    return HGK_OK;
}

extern "C" int hgk_conv_wgrad_tc_supported(int Cin, int Cout, int ksize) {
    return ((Cin == 64 || Cin == 128 || Cin == 256) && Cout > 0 && Cout % 4 == 0 && (ksize == 1 || ksize == 3)) ? 1 : 0;
}

extern "C" int hgk_conv_wgrad_tc_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                                      int N, int H, int W, int Cin, const float* dz, int Cout, int ksize,
                                      float* dw_tap_major, float* dbias, void* stream) {
    HGK_REQUIRE(x && dz && dw_tap_major, "hgk_conv_wgrad_tc_nhwc: null pointer");
    HGK_REQUIRE(N > 0 && H > 0 && W > 0, "hgk_conv_wgrad_tc_nhwc: empty tensor");
    HGK_REQUIRE(hgk_conv_wgrad_tc_supported(Cin, Cout, ksize), "hgk_conv_wgrad_tc_nhwc: unsupported shape Cin=%d Cout=%d k=%d "
                "(need Cin in {64,128,256}, Cout %% 4 == 0, k in {1,3})", Cin, Cout, ksize);
    HGK_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), "hgk_conv_wgrad_tc_nhwc: x scale/shift must both be set");
    HGK_REQUIRE((uintptr_t)dw_tap_major % 16 == 0, "hgk_conv_wgrad_tc_nhwc: dw must be 16-byte aligned");
    {
        const int r2 = wgrad_tc2_try(x, x_scale, x_shift, x_relu, N, H, W, Cin, dz, Cout, ksize, dw_tap_major, dbias, stream);
        if (r2 < 0) return r2;
        if (r2 == 1) {
            HGK_CHECK_LAUNCH("hgk_conv_wgrad_tc_nhwc");
            return HGK_OK;
        }
    }
    WgTcArgs a;
    a.x = Act{x, x_scale, x_shift, x_relu};
    a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.dz = dz; a.Cout = Cout; a.ksize = ksize;
    a.dw = dw_tap_major; a.dbias = dbias;
    a.P = (long long)N * H * W;
    const int mtiles = (Cout + TBM - 1) / TBM;
    if (ksize == 3 && W % 4 == 0 && Cin <= 128 && getenv("HGK_WGRAD3_OFF") == nullptr) {
        // three taps per CTA (wgrad3_tc_kernel)
        long long want3 = kNumSMs / (mtiles * 3);
        if (want3 < 1) want3 = 1;
        long long max3 = (a.P + 511) / 512;
        long long sp = want3 > max3 ? max3 : want3;
        long long ck = (a.P + sp - 1) / sp;
        ck = (ck + 31) / 32 * 32;
        sp = (a.P + ck - 1) / ck;
        a.chunk = ck;
        dim3 grid3((unsigned)mtiles, 3u, (unsigned)sp);
        int rc3 = Cin == 64 ? launch_wg3<64>(a, grid3, (cudaStream_t)stream) : launch_wg3<128>(a, grid3, (cudaStream_t)stream);
        if (rc3 != HGK_OK) return rc3;
        HGK_CHECK_LAUNCH("hgk_conv_wgrad_tc_nhwc");
        return HGK_OK;
    }
    const int taps = ksize * ksize;
    // split the pixel range so that ~1 CTA per SM runs, but keep >= 16 stages (512 pixels) per CTA so the
    // atomics epilogue stays small against the main loop
    long long want = kNumSMs / (mtiles * taps);
    if (want < 1) want = 1;
    long long max_splits = (a.P + 511) / 512;
    long long splits = want > max_splits ? max_splits : want;
    if (splits > 65535) splits = 65535;
    long long chunk = (a.P + splits - 1) / splits;
    chunk = (chunk + 31) / 32 * 32;
    splits = (a.P + chunk - 1) / chunk;
    a.chunk = chunk;
    dim3 grid((unsigned)mtiles, (unsigned)taps, (unsigned)splits);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = Cin == 64 ? launch_wg<64>(a, grid, st) : (Cin == 128 ? launch_wg<128>(a, grid, st) : launch_wg<256>(a, grid, st));
    if (rc != HGK_OK) return rc;
    HGK_CHECK_LAUNCH("hgk_conv_wgrad_tc_nhwc");
    return HGK_OK;
}

extern "C" int hgk_unpack_add_grads(const float* src_base, float* dst_base, const long long* table, int n_entries,
                                    void* stream) {
    HGK_REQUIRE(src_base && dst_base && table, "hgk_unpack_add_grads: null pointer");
    if (n_entries <= 0) return HGK_OK;
    HGK_REQUIRE(n_entries <= 65535, "hgk_unpack_add_grads: too many entries");
    dim3 grid(72, (unsigned)n_entries);      // 8 elements per thread for a 128x128x9 tensor (8 blocks per entry left the launch
                                             // latency-bound: 50 us for 18 MB, alone at the end of the step)
    unpack_add_grads_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_base, dst_base, table, n_entries);
    HGK_CHECK_LAUNCH("hgk_unpack_add_grads");
    return HGK_OK;
}
