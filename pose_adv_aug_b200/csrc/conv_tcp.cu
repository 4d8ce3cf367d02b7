// Persistent, fully warp-specialised tcgen05 convolution (same contract as conv_tc_kernel).
//
// One CTA per SM walks over 128-pixel output tiles.  Three roles run concurrently and are coupled only by
// mbarriers, so the load pipeline never drains between tiles and the epilogue of tile i overlaps the main
// loop of tile i+1:
//   warps 0-7   producers : coalesced 128-bit loads (2 stages ahead, across tile boundaries), BN+ReLU on
//                           load, TF32 hi/lo split, UMMA K-major smem tile, fence.proxy.async, arrive;
//                           thread 0 also issues the weight stage as a cp.async.bulk (TMA bulk copy)
//   warp  8     MMA issuer: tcgen05.mma.kind::tf32 into one of TWO TMEM accumulator buffers,
//                           tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 9-12  epilogue  : tcgen05.ld of their 32 TMEM lanes, 32x32 transposes through a private smem patch,
//                           128-byte coalesced bias / residual / accumulate / store, fp64 BN statistics
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"

namespace hgk {

constexpr int TP_THREADS = 13 * 32;

template <int BN, bool SPLIT>
struct TcpCfg {
    static constexpr int BK = (SPLIT && BN >= 128) ? 16 : 32;
    static constexpr int QP = BK / 4;
    static constexpr int NJ = QP / 2;
    static constexpr int A_BYTES = TBM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int TR_BYTES = 4 * 32 * 33 * 4;                 // per-epilogue-warp transpose patches
    static constexpr int ST_BYTES = 2 * 4 * BN * 2 * 8;              // [tile parity][warp][BN][sum, sumsq] doubles
    static constexpr int EPI_BYTES = TR_BYTES + ST_BYTES;
    static constexpr int NST = ((200 * 1024 - EPI_BYTES) / STAGE) > 4 ? 4 : ((200 * 1024 - EPI_BYTES) / STAGE);
    // hi*hi and the two cross terms accumulate separately (fp32 accumulator truncation, see conv_tc.cu);
    // BN = 256 only occurs with K <= 256: one chain
    static constexpr int NACC = (SPLIT && BN < 256) ? 2 : 1;
    static constexpr int ACC_COLS = NACC * BN;
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int SMEM = NST * STAGE + EPI_BYTES + 256;
    static_assert(2 * ACC_COLS <= 512, "TMEM capacity");
    static_assert(NST >= 2, "pipeline needs two stages");
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(TP_THREADS, 1) conv_tcp_kernel(const TcArgs args) {
    using Cfg = TcpCfg<BN, SPLIT>;
    constexpr int BK = Cfg::BK, NJ = Cfg::NJ, NST = Cfg::NST, NACC = Cfg::NACC, ACC_COLS = Cfg::ACC_COLS;
    constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE = Cfg::STAGE, TMEM_COLS = Cfg::TMEM_COLS;
    constexpr uint32_t LBO_A = TBM * 16, LBO_B = BN * 16, SBO = 128;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const ConvArgs& a = args.c;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * 4 + 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = a.ksize * a.ksize;
    const int KC = a.Cin / BK;
    const int T = taps * KC;
    const int HW = a.H * a.W;
    const long long ntiles = (a.P + TBM - 1) / TBM;
    const uint32_t bar_fa = smem_u32(&bars[0]), bar_fb = smem_u32(&bars[4]), bar_em = smem_u32(&bars[8]);
    const uint32_t bar_accf = smem_u32(&bars[12]), bar_acce = smem_u32(&bars[14]);

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_fa + 8 * s, 256);
            mbar_init(bar_fb + 8 * s, 1);
            mbar_init(bar_em + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, 128);
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
        // ===================== producers =====================
        const size_t wblk = (size_t)BN * BK;
        const int p_low = lane & 7, q_low = lane >> 3;
        constexpr int QH = Cfg::QP / 4;
        const int quad = (warp % QH) * 4 + q_low;
        int a_row[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) a_row[j] = ((j * 8 + warp) / QH) * 8 + p_low;
        const float* xz = a.x.z;
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        // ---- load stream (runs two stages ahead of the stores, across tile boundaries) ----
        unsigned a_off[NJ], a_vm[NJ];
        long long l_tile = blockIdx.x;
        int l_it = 0, l_tap = 0, l_kc = 0, l_toff = 0;
        bool l_done = l_tile >= ntiles;
        auto setup_tile = [&]() {
            const long long m0 = l_tile * TBM;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
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
            l_it = 0; l_tap = 0; l_kc = 0;
            l_toff = (a.ksize == 3) ? -(a.W + 1) * a.Cin : 0;
        };
        if (!l_done) setup_tile();
        float4 a_reg[2][NJ];
        unsigned a_msk[2] = {0u, 0u};
        auto load_a = [&](int set) {
            if (l_done) return;
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
            if (++l_kc == KC) {
                l_kc = 0;
                ++l_tap;
                l_toff += (l_tap == 3 || l_tap == 6) ? (a.W - 2) * a.Cin : a.Cin;
            }
            if (++l_it == T) {                       // load stream enters the next tile of this CTA
                l_tile += gridDim.x;
                if (l_tile < ntiles) setup_tile(); else l_done = true;
            }
        };
        // ---- store stream ----
        int s_kc = 0, s_it = 0;
        auto store_a = [&](int s, int set) {
            float4 sc, sh;
            load_affine4(a.x.scale, a.x.shift, s_kc * BK + quad * 4, sc, sh);
            if (++s_kc == KC) s_kc = 0;
            uint8_t* base = sgen + s * STAGE + quad * LBO_A;
            const unsigned msk = a_msk[set];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float4 v = a_reg[set][j];
                if (has_aff && ((msk >> j) & 1u)) v = actc4(v, sc, sh, x_clamp);
                float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                *reinterpret_cast<float4*>(base + a_row[j] * 16) = hi;
                if (SPLIT) {
                    float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    *reinterpret_cast<float4*>(base + A_BYTES + a_row[j] * 16) = lo;
                }
            }
        };
        long long n_my = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const long long total = n_my * T;
        load_a(0);
        load_a(1);
        int s = 0;
        unsigned em_par = 1;
        for (long long g = 0; g < total; ++g) {
            if (g >= NST) mbar_wait(bar_em + 8 * s, em_par);
            if (tid == 0) {
                const uint32_t bb = bar_fb + 8 * s;
                const uint32_t dst = sbase + s * STAGE + (SPLIT ? 2 : 1) * A_BYTES;
                mbar_expect_tx(bb, (SPLIT ? 2 : 1) * B_BYTES);
                bulk_g2s(dst, args.w_hi + (size_t)s_it * wblk, B_BYTES, bb);
                if (SPLIT) bulk_g2s(dst + B_BYTES, args.w_lo + (size_t)s_it * wblk, B_BYTES, bb);
            }
            if (++s_it == T) s_it = 0;
            if (g & 1) store_a(s, 1); else store_a(s, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_fa + 8 * s);
            if (g & 1) load_a(1); else load_a(0);
            if (++s == NST) { s = 0; em_par ^= 1u; }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int s = 0;
            unsigned par = 0;
            long long i = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                const int b = (int)(i & 1);
                const unsigned u = (unsigned)(i >> 1);
                if (i >= 2) mbar_wait(bar_acce + 8 * b, (u - 1) & 1u);       // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem + b * ACC_COLS;
                for (int it = 0; it < T; ++it) {
                    mbar_wait(bar_fa + 8 * s, par);
                    mbar_wait(bar_fb + 8 * s, par);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = sbase + s * STAGE;
                    const uint32_t b_hi = a_hi + (SPLIT ? 2 : 1) * A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t da = umma_desc(a_hi + k * 2 * LBO_A, LBO_A, SBO);
                        const uint64_t db = umma_desc(b_hi + k * 2 * LBO_B, LBO_B, SBO);
                        const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
                        if (SPLIT) {
                            const uint64_t dal = umma_desc(a_hi + A_BYTES + k * 2 * LBO_A, LBO_A, SBO);
                            const uint64_t dbl = umma_desc(b_hi + B_BYTES + k * 2 * LBO_B, LBO_B, SBO);
                            const uint32_t small = acc + (NACC > 1 ? BN : 0);
                            umma_tf32(small, dal, db, IDESC, first);
                            umma_tf32(small, da, dbl, IDESC, 1u);
                            umma_tf32(acc, da, db, IDESC, NACC > 1 ? first : 1u);
                        } else {
                            umma_tf32(acc, da, db, IDESC, first);
                        }
                    }
                    umma_commit(bar_em + 8 * s);
                    if (++s == NST) { s = 0; par ^= 1u; }
                }
                umma_commit(bar_accf + 8 * b);
            }
        }
    } else {
        // ===================== epilogue warps (9..12) =====================
        const int lq = warp & 3;                     // TMEM lane quarter this warp may access
        const int wq = warp - 9;
        float* tr = reinterpret_cast<float*>(sgen + NST * STAGE) + wq * (32 * 33);
        double* stat = reinterpret_cast<double*>(sgen + NST * STAGE + Cfg::TR_BYTES);
        const bool do_stats = a.stat_sum != nullptr;
        const bool has_res = a.res.z != nullptr, res_aff = a.res.scale != nullptr;
        const int et = tid - 9 * 32;                 // 0..127
        long long i = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            const int b = (int)(i & 1);
            const unsigned u = (unsigned)(i >> 1);
            mbar_wait(bar_accf + 8 * b, u & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long m0 = tile * TBM + lq * 32;       // first pixel row of this warp
            double* st_t = stat + (size_t)b * 4 * BN * 2;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * ACC_COLS + c0);
                tmem_ld32(taddr, r);
                float acc[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(r[q]);
                if (NACC > 1) {
                    tmem_ld32(taddr + BN, r);
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[q] += __uint_as_float(r[q]);
                }
                if (c0 + 32 == BN) {                 // accumulator fully read: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(bar_acce + 8 * b);
                }
                // 32x32 transpose through the warp's private patch: lane = pixel row -> lane = channel
#pragma unroll
                for (int q = 0; q < 32; ++q) tr[lane * 33 + q] = acc[q];
                __syncwarp();
                const int n = c0 + lane;
                const float bv = a.bias != nullptr ? __ldg(a.bias + n) : 0.f;
                float rs = 1.f, rt = 0.f;
                if (res_aff) { rs = __ldg(a.res.scale + n); rt = __ldg(a.res.shift + n); }
                float s1 = 0.f, s2 = 0.f;
                double d1 = 0.0, d2 = 0.0;
#pragma unroll 1
                for (int rb = 0; rb < 32; rb += 8) {
                    float rr[8], oo[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const long long p = m0 + rb + k;
                        rr[k] = 0.f;
                        oo[k] = 0.f;
                        if (p < a.P) {
                            if (has_res) rr[k] = __ldg(a.res.z + p * a.Cout + n);
                            if (a.accumulate) oo[k] = a.y[p * a.Cout + n];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const long long p = m0 + rb + k;
                        if (p >= a.P) break;
                        float v = tr[(rb + k) * 33 + lane] + bv;
                        if (has_res) v += res_aff ? act1(rr[k], rs, rt, a.res.relu) : rr[k];
                        v += oo[k];
                        a.y[p * a.Cout + n] = v;
                        s1 += v;
                        s2 = fmaf(v, v, s2);
                    }
                    d1 += (double)s1; d2 += (double)s2;
                    s1 = 0.f; s2 = 0.f;
                }
                __syncwarp();
                if (do_stats) {
                    st_t[((size_t)wq * BN + n) * 2 + 0] = d1;
                    st_t[((size_t)wq * BN + n) * 2 + 1] = d2;
                }
            }
            if (do_stats) {
                asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps
                for (int n = et; n < BN; n += 128) {
                    double x1 = 0.0, x2 = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        x1 += st_t[((size_t)q * BN + n) * 2 + 0];
                        x2 += st_t[((size_t)q * BN + n) * 2 + 1];
                    }
                    atomicAdd(a.stat_sum + n, x1);
                    atomicAdd(a.stat_sq + n, x2);
                }
                // st_t is double-buffered by tile parity; the next use of this buffer is two tiles away and
                // separated from these reads by the bar.sync of the tile in between
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int BN, bool SPLIT>
static int launch_tcp(const TcArgs& ta, cudaStream_t st) {
    static bool configured = false;
    constexpr int smem = TcpCfg<BN, SPLIT>::SMEM;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcp_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_tc_nhwc: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    long long mt = (ta.c.P + TBM - 1) / TBM;
    unsigned grid = (unsigned)(mt < kNumSMs ? mt : kNumSMs);
    conv_tcp_kernel<BN, SPLIT><<<grid, TP_THREADS, smem, st>>>(ta);
    return HGK_OK;
}

// entry used by hgk_conv_tc_nhwc when the persistent kernel is enabled
int conv_tcp_launch(const TcArgs& ta, bool split, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int Cout = ta.c.Cout;
    if (Cout == 64) return split ? launch_tcp<64, true>(ta, st) : launch_tcp<64, false>(ta, st);
    if (Cout == 128) return split ? launch_tcp<128, true>(ta, st) : launch_tcp<128, false>(ta, st);
    return split ? launch_tcp<256, true>(ta, st) : launch_tcp<256, false>(ta, st);
}

}  // namespace hgk
