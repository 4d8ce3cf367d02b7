// Weight gradient on tcgen05 with MN-MAJOR TF32 operands (no transposes anywhere):
//     dW[tap][co][ci] += sum_p dz[p][co] * act(x)[p + off(tap)][ci]
// is a GEMM with M = co, N = ci, K = pixels.  NHWC tensors are [pixel][channel], i.e. the GEMM operands are
// MN-major (channel contiguous), which kind::tf32 accepts directly (instruction-descriptor bits 15/16) in ONE
// shared-memory layout: the 128-byte swizzle with 32-byte atoms (descriptor layout type 1; measured: the
// no-swizzle MN-major form yields zeros).  Its atom is 4 K-rows x 128 bytes = four consecutive PIXELS x 32
// channels, rows 128 bytes apart, the 32-byte chunk index XORed with the row index (address bits [5,7) ^= bits
// [7,9)); LBO = distance between 32-channel blocks, SBO = distance between 4-pixel groups.  So the operand tile is
// [32-channel block][pixel][128 bytes] and a producer thread moves one float4 (4 channels of one pixel) from a
// 512-byte-coalesced global load to one swizzled 16-byte store.
// (wgrad_tc_kernel in conv_tc.cu builds K-major tiles with 4x4 register transposes from loads that touch 8
// cache lines each: ncu shows it pinned at 80-93 % L1TEX throughput, profiles/r1b_tc_summary.md.)
//
// One CTA = one range of 8-pixel units x ALL output channels (MT accumulator row tiles of 128) x ALL input
// channels: x and dz are read exactly once per tap row.  3x3: a CTA owns one kernel row dh and stages, per
// 8-pixel unit, the 10-pixel x run [w0-1, w0+8] of image row h+dh once; the three taps dw = -1,0,+1 are three
// MMAs whose B descriptors start 0, 16 and 32 bytes into that run.
// Roles: warps 0-7 load / BN+ReLU / round / store (two stages of register prefetch), warp 8 issues
// tcgen05.mma.kind::tf32 and commits stages back; then all warps reduce the accumulators into the
// [tap][Cout][Cin] destination with red.global.add.v4.f32.
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"

namespace hgk {

struct Wg2Args {
    Act x;
    int N, H, W, Cin;
    const float* dz;
    int Cout, ksize;
    float* dw;
    float* dbias;
    long long units;          // P / 8
    long long units_per_cta;  // multiple of 4
    int lw8;                  // log2(W / 8) (3x3 only; W and H are powers of two there)
};

__device__ __forceinline__ void red_add_v4_g(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B_BASE32B (layout type 1 in bits 61-63)
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return umma_desc(saddr, lbo, sbo) | ((uint64_t)1 << 61);
}
// byte offset of the float4 (channel quad q, pixel slot p) inside an operand tile [block q/8][p][128 B]:
// 16-byte chunk q%8 of the row, its 32-byte chunk index XORed with p%4
__device__ __forceinline__ uint32_t sw_off(int q, int p, uint32_t lbo) {
    const uint32_t c16 = (uint32_t)q & 7u;
    return (uint32_t)(q >> 3) * lbo + (uint32_t)p * 128u + ((((c16 >> 1) ^ ((uint32_t)p & 3u)) << 5) | ((c16 & 1u) << 4));
}

constexpr int WG2_THREADS = 288;      // 8 producer/epilogue warps + 1 MMA warp

template <int CIN, int MT, int TAPS>
struct Wg2Cfg {
    static constexpr int HALO = TAPS == 3 ? 2 : 0;
    static constexpr int UPX = 8 + HALO;                     // x pixels staged per unit
    static constexpr int QA = MT * 32, QB = CIN / 4;         // channel quads of the dz / x operand
    static constexpr int NPA = 32, NPB = 4 * UPX;            // pixels per stage
    static constexpr int LBO_A = NPA * 128, LBO_B = NPB * 128;          // byte distance between 32-channel blocks
    static constexpr int A_BYTES = (QA / 8) * LBO_A, B_BYTES = (QB / 8) * LBO_B;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int NJA = NPA * QA / 256, NJB = (NPB * QB + 255) / 256;
    static constexpr int ACC = MT * TAPS * CIN;              // TMEM columns in use
    static constexpr int TMEM_COLS = ACC <= 64 ? 64 : ACC <= 128 ? 128 : ACC <= 256 ? 256 : 512;
    static constexpr int CH = CIN > 128 ? 128 : CIN;
    static constexpr int STG_BYTES = TBM * (CH + 4) * 4;
    // Two CTAs per SM wherever the accumulators leave room for it (<= 256 TMEM columns): the producer chain of a stage
    // (wait -> transform -> proxy fence -> arrive -> MMA -> commit) is ~1 700 cycles for 32 pixels whatever the channel
    // counts are, so a lone CTA per SM left the SM idle most of the time (ncu: 14 % warps active); the second CTA's chain
    // runs in the gaps.  Shared memory is then split in two (<= 110 KB each).
    // Only the variants whose two register-prefetch sets fit under the 96-register cap of two 288-thread CTAs
    // (NJA + NJB <= 8 float4 per set).  Measured for the larger ones (128->256, 256->128 at 64x64): two CTAs with ONE
    // prefetch set are 10 % slower than one CTA with two (45.8 -> 51.4 us, 49.9 -> 54.5 us), two CTAs with two sets spill.
    static constexpr bool OCC2 = TMEM_COLS <= 256 && NJA + NJB <= 8;
    static constexpr int NSET = 2;
    static constexpr int SMEM_BUDGET = OCC2 ? 108 * 1024 : 200 * 1024;
    static constexpr int NST = (SMEM_BUDGET / STAGE) > 4 ? 4 : (SMEM_BUDGET / STAGE);
    static constexpr int SMEM = (NST * STAGE > STG_BYTES + 4096 ? NST * STAGE : STG_BYTES + 4096) + 1024;
    static_assert(!OCC2 || SMEM <= 113 * 1024, "two CTAs per SM share 227 KB");
    static_assert(ACC <= 512, "TMEM capacity");
    static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzle atoms need 512-byte aligned tiles");
    static_assert(NST >= 2, "pipeline needs two stages");
};

template <int CIN, int MT, int TAPS>
__global__ void __launch_bounds__(WG2_THREADS, (Wg2Cfg<CIN, MT, TAPS>::OCC2 ? 2 : 1)) wgrad2_tc_kernel(const Wg2Args a) {
    using Cfg = Wg2Cfg<CIN, MT, TAPS>;
    constexpr int QA = Cfg::QA, QB = Cfg::QB, NJA = Cfg::NJA, NJB = Cfg::NJB, NST = Cfg::NST, UPX = Cfg::UPX;
    constexpr uint32_t LBO_A = Cfg::LBO_A, LBO_B = Cfg::LBO_B, A_BYTES = Cfg::A_BYTES, STAGE = Cfg::STAGE;
    // D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), A and B MN-major (bits 15, 16), N>>3 at 17, M>>4 at 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(CIN >> 3) << 17) |
                               ((uint32_t)(TBM >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NST + 1];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = TAPS == 3 ? (int)blockIdx.y - 1 : 0;
    const long long u_begin = (long long)blockIdx.x * a.units_per_cta;
    const long long u_end = (u_begin + a.units_per_cta < a.units) ? (u_begin + a.units_per_cta) : a.units;
    const int T = (int)((u_end - u_begin + 3) / 4);           // stages of 4 units (host guarantees u_begin < units)
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NST]), bar_done = smem_u32(&bars[2 * NST]);

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 256);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool do_bias = a.dbias != nullptr && blockIdx.y == 0;

    if (warp < 8) {
        // ===== producers.  item idx = tid + 256 j: pixel slot idx / Q, channel quad idx % Q (fixed per thread) =====
        const int qa = tid % QA, qb = tid % QB;
        const bool qa_ok = qa * 4 < a.Cout;                    // Cout < 128 MT: the upper accumulator rows stay unused
        float4 xs, xt;
        load_affine4(a.x.scale, a.x.shift, qb * 4, xs, xt);
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        const float* dzp = a.dz + qa * 4;
        const float* xzp = a.x.z + qb * 4;
        constexpr int NSET = Cfg::NSET;
        float4 ra[NSET][NJA], rb[NSET][NJB];
        unsigned bm[NSET] = {};
        // all index arithmetic in 32 bits (host guarantees P * max(Cin, Cout) < 2^31); H and W are powers of two for 3x3
        int l_u = (int)u_begin;                                  // first unit of the stage the next load() fetches
        const int i_end = (int)u_end;
        const int hmask = a.H - 1, umask = (1 << a.lw8) - 1;
        auto load = [&](int set) {
#pragma unroll
            for (int j = 0; j < NJA; ++j) {
                const int slot = (tid + 256 * j) / QA;           // pixel of the stage
                const int u = l_u + (slot >> 3);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < i_end && qa_ok) v = ldg4(dzp + (unsigned)(u * 8 + (slot & 7)) * (unsigned)a.Cout);
                ra[set][j] = v;
            }
            unsigned m = 0;
#pragma unroll
            for (int j = 0; j < NJB; ++j) {
                const int idx = tid + 256 * j;
                const int slot = idx / QB;                       // x pixel slot of the stage: unit slot / UPX, offset slot % UPX
                const int us = slot / UPX, i = slot - us * UPX;
                const int u = l_u + us;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                bool ok = idx < Cfg::NPB * QB && u < i_end;
                int p;
                if (TAPS == 3) {
                    const int r = u >> a.lw8;                    // global image row n*H + h
                    const int w = (u & umask) * 8 - 1 + i;
                    const int h = (r & hmask) + dh;
                    ok = ok && (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W;
                    p = (r + dh) * a.W + w;
                } else {
                    p = u * 8 + i;
                }
                if (ok) {
                    v = ldg4(xzp + (unsigned)p * (unsigned)a.Cin);
                    m |= 1u << j;
                }
                rb[set][j] = v;
            }
            bm[set] = m;
            l_u += 4;
        };
        auto store = [&](int s, int set) {
            uint8_t* sa = sgen + s * STAGE;
#pragma unroll
            for (int j = 0; j < NJA; ++j) {
                const int slot = (tid + 256 * j) / QA;
                const float4 v = ra[set][j];
                if (do_bias) { bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w; }
                *reinterpret_cast<float4*>(sa + sw_off(qa, slot, LBO_A)) = tf32_rna4(v);
            }
            uint8_t* sb = sa + A_BYTES;
#pragma unroll
            for (int j = 0; j < NJB; ++j) {
                const int idx = tid + 256 * j;
                if (idx >= Cfg::NPB * QB) continue;
                const int slot = idx / QB;
                float4 v = rb[set][j];
                // BN+ReLU on valid pixels only (padding = zero of the activated tensor)
                if (has_aff && ((bm[set] >> j) & 1u)) v = actc4(v, xs, xt, x_clamp);
                *reinterpret_cast<float4*>(sb + sw_off(qb, slot, LBO_B)) = tf32_rna4(v);
            }
        };
        load(0);
        if (NSET > 1 && T > 1) load(NSET - 1);
        int s = 0;
        unsigned em_par = 1;
        for (int it = 0; it < T; ++it) {
            if (it >= NST) mbar_wait(bar_empty + 8 * s, em_par);
            if (NSET > 1 && (it & 1)) store(s, NSET - 1); else store(s, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_full + 8 * s);
            if (it + NSET < T) { if (NSET > 1 && (it & 1)) load(NSET - 1); else load(0); }
            if (++s == NST) { s = 0; em_par ^= 1u; }
        }
        mbar_wait(bar_done, 0);
    } else if (lane == 0) {
        // ===== MMA issuer =====
        // MN-major, 128B swizzle / 32B atoms: LBO = byte distance between 32-channel blocks, SBO = between 4-pixel groups
        // (verified on hardware, tools/dbg_wg2.py; a 3x3 tap shifts the start address by whole 128-byte rows: the
        // swizzle is a function of the absolute shared-memory address, so shifted starts stay consistent)
        for (int it = 0; it < T; ++it) {
            const int s = it % NST, u = it / NST;
            mbar_wait(bar_full + 8 * s, u & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = sbase + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int t = 0; t < TAPS; ++t) {
                    const uint64_t db = umma_desc_mn32(sb + (k * UPX + t) * 128, LBO_B, 512u);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const uint64_t da = umma_desc_mn32(sa + m * 4 * LBO_A + k * 1024, LBO_A, 512u);
                        umma_tf32(tmem + (m * TAPS + t) * CIN, da, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                    }
                }
            }
            umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_done);
    }
    __syncwarp();      // the single-lane role reconverges before the CTA-wide barrier
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();       // accumulators complete, shared memory reusable

    // ---- bias gradient: per-thread partial sums of the dz quads -> shared -> one atomic per channel ----
    float* stg = reinterpret_cast<float*>(sgen);
    constexpr int CH = Cfg::CH, SROW = CH + 4;
    float* bred = stg + TBM * SROW;                 // [256][4] floats behind the staging tile
    if (do_bias) {
        if (tid < 256) st4(bred + tid * 4, bsum);
        __syncthreads();
        if (tid < QA * 4) {
            const int q = tid >> 2, e = tid & 3;
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < 256 / QA; ++r) v += bred[(r * QA + q) * 4 + e];
            if (tid < a.Cout) atomicAdd(a.dbias + tid, v);
        }
    }
    // ---- accumulators: TMEM -> staging tile -> coalesced vector atomics into [tap][Cout][Cin] ----
    constexpr int NCH = CIN / CH;
#pragma unroll 1
    for (int acc = 0; acc < MT * TAPS; ++acc) {
        const int m = acc / TAPS, t = acc - m * TAPS;
        const int tap = TAPS == 3 ? (dh + 1) * 3 + t : 0;
        if (m * TBM >= a.Cout) break;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
            if (warp < 8) {
                const int lq = warp & 3;
                const int row = lq * 32 + lane;
                const int cbeg = (warp >> 2) * (CH / 2);
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + CH / 2; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * CIN + ch * CH + c0), r);
                    float* dst = stg + row * SROW + c0;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st4(dst + q * 4, make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]),
                                                     __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3])));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid < 256) {
                constexpr int CG = CH / 4, RL = 256 / CG;
                const int cg = tid % CG, r0 = tid / CG;
                for (int r = r0; r < TBM; r += RL) {
                    const int co = m * TBM + r;
                    if (co >= a.Cout) break;
                    const float4 v = ld4(stg + r * SROW + cg * 4);
                    red_add_v4_g(a.dw + ((size_t)tap * a.Cout + co) * CIN + ch * CH + cg * 4, v);
                }
            }
            __syncthreads();                        // staging tile is rewritten by the next chunk
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
}

template <int CIN, int MT, int TAPS>
static int launch_wg2(const Wg2Args& a, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    constexpr int smem = Wg2Cfg<CIN, MT, TAPS>::SMEM;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad2_tc_kernel<CIN, MT, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_wgrad_tc_nhwc (MN-major kernel): cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    wgrad2_tc_kernel<CIN, MT, TAPS><<<grid, WG2_THREADS, smem, st>>>(a);
    return HGK_OK;
}

// returns 1 when the MN-major kernel took the launch, 0 when the shape is not covered (caller falls back), < 0 on error
int wgrad_tc2_try(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W, int Cin,
                  const float* dz, int Cout, int ksize, float* dw, float* dbias, void* stream) {
    static int mode = -1;              // HGK_WG2=0 disables this kernel (every shape then takes the K-major kernels)
    if (mode < 0) {
        const char* e = getenv("HGK_WG2");
        mode = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (mode == 0) return 0;
    const long long P = (long long)N * H * W;
    if (P % 8) return 0;
    if (!(Cout == 64 || Cout == 128 || Cout == 256)) return 0;
    if (P * (long long)(Cin > Cout ? Cin : Cout) >= (1LL << 31)) return 0;       // 32-bit index arithmetic in the producers
    const int MT = Cout > 128 ? 2 : 1;
    int lw8 = 0;
    if (ksize == 3) {
        if (W < 8 || (W & (W - 1)) || (H & (H - 1)) || Cout > 128 || Cin > 128) return 0;   // 3 x MT x Cin accumulator columns
        while ((8 << lw8) < W) ++lw8;
    } else if (MT * Cin > 512) {
        return 0;
    }
    Wg2Args a;
    a.x = Act{x, x_scale, x_shift, x_relu};
    a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.dz = dz; a.Cout = Cout; a.ksize = ksize; a.dw = dw; a.dbias = dbias;
    a.units = P / 8;
    a.lw8 = lw8;
    const int groups = ksize == 3 ? 3 : 1;
    // two CTAs per SM for the variants with Wg2Cfg::OCC2 (same predicate)
    const int acc_cols = MT * (ksize == 3 ? 3 : 1) * Cin;
    const int nja = 4 * MT, njb = (4 * (ksize == 3 ? 10 : 8) * (Cin / 4) + 255) / 256;
    const bool occ2 = acc_cols <= 256 && nja + njb <= 8;
    long long want = (long long)kNumSMs * (occ2 ? 2 : 1) / groups;
    long long max_ctas = (a.units + 15) / 16;                  // at least 4 stages per CTA
    long long ctas = want < max_ctas ? want : max_ctas;
    if (ctas < 1) ctas = 1;
    long long upc = (a.units + ctas - 1) / ctas;
    upc = (upc + 3) / 4 * 4;
    ctas = (a.units + upc - 1) / upc;
    a.units_per_cta = upc;
    dim3 grid((unsigned)ctas, (unsigned)groups);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (ksize == 3) rc = Cin == 64 ? launch_wg2<64, 1, 3>(a, grid, st) : launch_wg2<128, 1, 3>(a, grid, st);
    else if (MT == 1) rc = Cin == 64 ? launch_wg2<64, 1, 1>(a, grid, st)
                                     : (Cin == 128 ? launch_wg2<128, 1, 1>(a, grid, st) : launch_wg2<256, 1, 1>(a, grid, st));
    else rc = Cin == 64 ? launch_wg2<64, 2, 1>(a, grid, st)
                        : (Cin == 128 ? launch_wg2<128, 2, 1>(a, grid, st) : launch_wg2<256, 2, 1>(a, grid, st));
    if (rc != HGK_OK) return rc;
    return 1;
}

}  // namespace hgk
