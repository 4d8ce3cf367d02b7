// Persistent, phase-overlapped tcgen05 implicit-GEMM convolution over 16x8-pixel image tiles.
//
// Same contract as conv_tc2_kernel (conv_tc2.cu): BN+ReLU on load, 3xTF32 (forward) or plain TF32 (data gradient),
// optional BatchNorm-backward apply on load (BNAPPLY), bias / shortcut / accumulate epilogue with the BatchNorm
// statistics (or the fused BatchNorm-backward reduction, BWDSTATS) and the last-CTA finaliser.  What changed is the
// execution structure (profiles/r2_tile_kernel_timeline.md: prologue + epilogue were 40-55 % of a CTA's life and the
// two CTAs of an SM ran in lock-step):
//   * ONE CTA per SM walks over its tiles (tile = blockIdx.x + i * gridDim.x); nothing drains between tiles;
//   * the activation halo tile of a 16-channel chunk is fetched by TMA (cp.async.bulk.tensor, 4-D tensor map over
//     [N][H][W][C], out-of-bounds rows / columns zero-filled by the hardware) DIRECTLY into the K-major 64-byte-swizzle
//     UMMA layout (CU_TENSOR_MAP_SWIZZLE_64B is that layout), several chunks ahead -- no address generation, no
//     register staging, the bytes in flight are bounded by shared memory instead of registers;
//   * eight transform warps turn the landed tile IN PLACE into the TF32 "hi" operand (BN+ReLU, round) and write the
//     "lo" operand (3xTF32) next to it; for BNAPPLY the second plane receives the raw z tile (second tensor map) and
//     the transform forms dz = cA*((g*mask - cC) - (z - mean)*cB), writes it back to global once, and rounds it;
//   * one thread issues tcgen05.mma.kind::tf32 into one of TWO TMEM accumulator sets; the nine taps of a 3x3 are nine
//     descriptors into the same halo tile (start address shifted by (dh*HWD + dw)*64 bytes);
//   * eight epilogue warps drain accumulator set i&1 (tcgen05.ld -> per-warp 32x32 transpose patch -> coalesced
//     128-bit stores, shortcut / accumulate loads issued one block ahead) while the main loop of tile i+1 runs;
//     BatchNorm statistics are kept in shared memory (fp64) across tiles: one global atomic per channel per CTA.
// 1x1 convolutions see the tensor as [1][P/8][8][C] (any H, W with P % 128 == 0: also the 8x8 / 4x4 rungs).
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"
#include "bn_fin.cuh"

namespace hgk {

__device__ __forceinline__ uint64_t umma_desc_k64_3(uint32_t saddr, uint32_t sbo) {
    return umma_desc(saddr, 16u, sbo) | ((uint64_t)4 << 61);
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// developer diagnostics (tools/dbg_timeline3.py): cycles a role spends waiting on a barrier, accumulated per CTA into
// args.dbg[cta][slot]; a predicated-off branch in production (args.dbg == nullptr)
#define T3_WAIT(bar, par, acc)                          \
    do {                                                \
        if (dbg_on) {                                   \
            const long long t0__ = clock64();           \
            mbar_wait(bar, par);                        \
            acc += clock64() - t0__;                    \
        } else {                                        \
            mbar_wait(bar, par);                        \
        }                                               \
    } while (0)
#define T3_DBG(slot, val)                                                                      \
    do {                                                                                       \
        if (dbg_on && blockIdx.x < 256) args.dbg[blockIdx.x * 32 + (slot)] = (long long)(val); \
    } while (0)

constexpr int T3_NTW = 8;                                   // transform warps
constexpr int T3_NEW = 8;                                   // epilogue warps
constexpr int T3_WMMA = T3_NTW, T3_WLDA = T3_NTW + 1, T3_WLDB = T3_NTW + 2, T3_WAUX = T3_NTW + 3;
constexpr int T3_WEPI = T3_NTW + 4;                         // first epilogue warp (multiple of 4: TMEM lane quarter = warp % 4)
constexpr int T3_THREADS = (T3_WEPI + T3_NEW) * 32;         // 640
constexpr int T3_NTT = T3_NTW * 32;                         // transform threads
constexpr int T3_PATCH = 32 * 36 * 4;                       // per-epilogue-warp transpose patch (row stride 36 floats)
static_assert(T3_WEPI % 4 == 0, "epilogue warps must start at a multiple of four");

template <int BN, bool SPLIT, int KS, bool BNAPPLY>
struct T3Cfg {
    static constexpr int TH = 16, TW = 8;
    static constexpr int PAD = KS / 2;
    static constexpr int HH = TH + 2 * PAD, HWD = TW + 2 * PAD;
    static constexpr int NPIX = HH * HWD;
    static constexpr int TAPS = KS * KS;
    static constexpr int NITEM = NPIX * 4;
    static constexpr int NJ = (NITEM + T3_NTT - 1) / T3_NTT;
    static constexpr int SBO_A = HWD * 64;
    static constexpr int A_BOX = NPIX * 64;                                  // bytes one tensor-map box delivers
    static constexpr int A_PLANE = (A_BOX + 1023) / 1024 * 1024;
    static constexpr int NPLANE = (SPLIT || BNAPPLY) ? 2 : 1;                // plane 1: lo operand (SPLIT) or raw z (BNAPPLY)
    static constexpr int A_STAGE = NPLANE * A_PLANE;
    static constexpr int A_TX = (BNAPPLY ? 2 : 1) * A_BOX;
    static constexpr int B_HALF = BN * 16 * 4;
    static constexpr int B_STAGE = (SPLIT ? 2 : 1) * B_HALF;
    // accumulators per tile, see conv_tc.cu (the tensor core truncates its fp32 accumulator): long 3xTF32 chains are
    // split over hi*hi / cross-term accumulators and summed in fp32 by the epilogue
    static constexpr int NMAIN = (SPLIT && KS == 3 && BN <= 64) ? 2 : 1;
    static constexpr int NACC = (!SPLIT || KS == 1 || BN >= 256) ? 1 : NMAIN + 1;
    static constexpr int SUBCOLS = NACC * BN;
    static constexpr int TMEM_COLS = (2 * SUBCOLS <= 64) ? 64 : (2 * SUBCOLS <= 128) ? 128 : (2 * SUBCOLS <= 256) ? 256 : 512;
    static constexpr int CPW = BN / 2;                                       // columns per epilogue warp
    static constexpr int NBLK = CPW / 32;                                    // 32-column blocks per epilogue warp
    static constexpr int EPI_BYTES = T3_NEW * T3_PATCH + T3_NEW * CPW * 2 * 8;
    static constexpr int AVAIL = 227 * 1024 - 1024 - EPI_BYTES - 1024;
    static constexpr int NSB = KS == 3 ? (B_STAGE <= 8192 ? 9 : 6) : (B_STAGE >= 32768 ? 3 : 4);
    static constexpr int NSA_FIT = (AVAIL - NSB * B_STAGE) / A_STAGE;
    static constexpr int NSA = NSA_FIT > 8 ? 8 : NSA_FIT;
    static constexpr int PIPE = NSA * A_STAGE + NSB * B_STAGE;
    static constexpr int SMEM = PIPE + EPI_BYTES + 1024;
    static_assert(2 * SUBCOLS <= 512, "TMEM capacity (two accumulator sets)");
    static_assert(NSA >= 2, "activation ring needs two stages");
    static_assert(BN % 64 == 0, "two epilogue warps per lane quarter split the columns in 32-column blocks");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY>
__global__ void __launch_bounds__(T3_THREADS, 1)
conv_tc3_kernel(const TcArgs args, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_z) {
    static_assert(!(BNAPPLY && SPLIT), "the fused BatchNorm-backward apply is a data-gradient (plain TF32) mode");
    using Cfg = T3Cfg<BN, SPLIT, KS, BNAPPLY>;
    constexpr int NJ = Cfg::NJ, NSA = Cfg::NSA, NSB = Cfg::NSB, NMAIN = Cfg::NMAIN, NACC = Cfg::NACC;
    constexpr int HWD = Cfg::HWD, PAD = Cfg::PAD, TAPS = Cfg::TAPS, SUBCOLS = Cfg::SUBCOLS;
    constexpr uint32_t SBO_A = Cfg::SBO_A, LBO_B = BN * 16, SBO_B = 128;
    constexpr uint32_t A_PLANE = Cfg::A_PLANE, A_STAGE = Cfg::A_STAGE, B_HALF = Cfg::B_HALF, B_STAGE = Cfg::B_STAGE;
    constexpr uint32_t B_OFF = NSA * A_STAGE;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const ConvArgs& a = args.c;

    extern __shared__ uint8_t smem_raw[];
    // barriers: raw[NSA] (TMA landed) | afull[NSA] (transformed) | aempty[NSA] (MMAs retired) | bfull[NSB] | bempty[NSB] |
    //           accfull[2] | accempty[2]
    __shared__ __align__(8) uint64_t bars[3 * NSA + 2 * NSB + 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_w = a.W >> 3, tiles_hw = (a.H >> 4) * tiles_w;
    const int ntiles = a.N * tiles_hw;
    const int KC = a.Cin >> 4;
    const uint32_t bar_raw = smem_u32(&bars[0]), bar_fa = smem_u32(&bars[NSA]), bar_ea = smem_u32(&bars[2 * NSA]);
    const uint32_t bar_fb = smem_u32(&bars[3 * NSA]), bar_eb = smem_u32(&bars[3 * NSA + NSB]);
    const uint32_t bar_accf = smem_u32(&bars[3 * NSA + 2 * NSB]), bar_acce = smem_u32(&bars[3 * NSA + 2 * NSB + 2]);

    if (tid == 0) {
        for (int s = 0; s < NSA; ++s) {
            mbar_init(bar_raw + 8 * s, 1);
            mbar_init(bar_fa + 8 * s, T3_NTT);
            mbar_init(bar_ea + 8 * s, 1);
        }
        for (int s = 0; s < NSB; ++s) {
            mbar_init(bar_fb + 8 * s, 1);
            mbar_init(bar_eb + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, T3_NEW * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T3_WAUX) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == T3_WLDA && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
        if (BNAPPLY) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_z)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const bool dbg_on = args.dbg != nullptr;
    long long w0 = 0, w1 = 0, w2 = 0;
    const long long t_begin = dbg_on ? clock64() : 0;
    if (tid == 0) T3_DBG(14, gtimer());

    if (warp < T3_NTW) {
        // ===== transform warps: landed halo tile -> (BN+ReLU | BN-backward apply) -> TF32 hi (in place) [+ lo plane] =====
        // item idx = tid + 256*j: halo pixel idx>>2, 16-byte channel quad idx&3 (= tid&3 for every j)
        const int quad = tid & 3;
        unsigned s_off[NJ];
        int hw_j[NJ];                           // hh | ww << 8
        unsigned smask = 0, imask = 0;          // item inside the tile / pixel owned by this tile (BNAPPLY: its dz is written)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int idx = tid + T3_NTT * j;
            const int hp = idx >> 2;
            const int hh = hp / HWD, ww = hp - hh * HWD;
            hw_j[j] = hh | (ww << 8);
            // 64-byte swizzle: 16-byte chunk index ^= address bits [7,9) = (pixel >> 1) & 3 (plane bases are 1024-aligned)
            s_off[j] = (unsigned)hp * 64u + (((unsigned)quad ^ (((unsigned)hp >> 1) & 3u)) << 4);
            if (idx < Cfg::NITEM) {
                smask |= 1u << j;
                if (hh >= PAD && hh < PAD + 16 && ww >= PAD && ww < PAD + 8) imask |= 1u << j;
            }
        }
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        const BnApply& ap = a.ap;
        int sa = 0;
        unsigned par = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int n_img = tile / tiles_hw;
            const int trem = tile - n_img * tiles_hw;
            const int th0 = (trem / tiles_w) << 4, tw0 = (trem % tiles_w) << 3;
            unsigned vmask = smask;             // pixel inside the image (zero padding is a zero of the ACTIVATED tensor)
            unsigned g_off[BNAPPLY ? NJ : 1];
            if (KS == 3 || BNAPPLY) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int h = th0 - PAD + (hw_j[j] & 255), w = tw0 - PAD + (hw_j[j] >> 8);
                    if (KS == 3 && !((unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W)) vmask &= ~(1u << j);
                    if (BNAPPLY) g_off[j] = (unsigned)((n_img * a.H + h) * a.W + w) * (unsigned)a.Cin + (unsigned)quad * 4u;
                }
            }
            for (int kc = 0; kc < KC; ++kc) {
                float4 sc, sh;
                load_affine4(a.x.scale, a.x.shift, kc * 16 + quad * 4, sc, sh);
                float4 bs, bt, bmu, bA, bB, bC;
                if (BNAPPLY) {
                    const int c = kc * 16 + quad * 4;
                    bs = ldg4(ap.scale + c); bt = ldg4(ap.shift + c); bmu = ldg4(ap.mean + c);
                    bA = ldg4(ap.cA + c); bB = ldg4(ap.cB + c); bC = ldg4(ap.cC + c);
                }
                T3_WAIT(bar_raw + 8 * sa, par, w0);
                const long long tc0 = dbg_on ? clock64() : 0;
                uint8_t* base = sgen + sa * A_STAGE;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (!((smask >> j) & 1u)) continue;
                    float4 v = ld4(reinterpret_cast<const float*>(base + s_off[j]));
                    const bool in_img = (vmask >> j) & 1u;
                    if (has_aff && in_img) v = actc4(v, sc, sh, x_clamp);
                    if (BNAPPLY) {
                        // dz = cA * ((g * relu-mask - cC) - (z - mean) * cB)   (bn_bwd_apply_kernel, bn.cu)
                        const float4 z = ld4(reinterpret_cast<const float*>(base + A_PLANE + s_off[j]));
                        if (in_img) {
                            const float gx = (ap.relu && fmaf(z.x, bs.x, bt.x) <= 0.f) ? 0.f : v.x;
                            const float gy = (ap.relu && fmaf(z.y, bs.y, bt.y) <= 0.f) ? 0.f : v.y;
                            const float gz = (ap.relu && fmaf(z.z, bs.z, bt.z) <= 0.f) ? 0.f : v.z;
                            const float gw = (ap.relu && fmaf(z.w, bs.w, bt.w) <= 0.f) ? 0.f : v.w;
                            v.x = bA.x * ((gx - bC.x) - (z.x - bmu.x) * bB.x);
                            v.y = bA.y * ((gy - bC.y) - (z.y - bmu.y) * bB.y);
                            v.z = bA.z * ((gz - bC.z) - (z.z - bmu.z) * bB.z);
                            v.w = bA.w * ((gw - bC.w) - (z.w - bmu.w) * bB.w);
                            if ((imask >> j) & 1u) st4(ap.dz + (g_off[BNAPPLY ? j : 0] + (unsigned)kc * 16u), v);
                        }
                    }
                    const float4 hi = tf32_rna4(v);
                    *reinterpret_cast<float4*>(base + s_off[j]) = hi;
                    if (SPLIT) {
                        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                        *reinterpret_cast<float4*>(base + A_PLANE + s_off[j]) = lo;
                    }
                }
                const long long tc1 = dbg_on ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                mbar_arrive(bar_fa + 8 * sa);
                if (dbg_on) { w1 += tc1 - tc0; w2 += clock64() - tc1; }
                if (++sa == NSA) { sa = 0; par ^= 1u; }
            }
        }
        if (tid == 0) { T3_DBG(0, clock64() - t_begin); T3_DBG(1, w0); T3_DBG(12, w1); T3_DBG(13, w2); }
    } else if (warp == T3_WMMA) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int sa = 0, sb = 0;
            unsigned fa_par = 0, fb_par = 0;
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                const int b = i & 1;
                if (i >= 2) T3_WAIT(bar_acce + 8 * b, (unsigned)((i >> 1) - 1) & 1u, w2);   // epilogue drained this accumulator set
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_set = tmem + (uint32_t)(b * SUBCOLS);
                int it = 0;
                for (int kc = 0; kc < KC; ++kc) {
                    T3_WAIT(bar_fa + 8 * sa, fa_par, w0);
                    const uint32_t a_stage = sbase + sa * A_STAGE;
#pragma unroll 1
                    for (int tap = 0; tap < TAPS; ++tap, ++it) {
                        T3_WAIT(bar_fb + 8 * sb, fb_par, w1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_tap = a_stage + (KS == 3 ? (uint32_t)((tap / 3) * HWD + (tap % 3)) * 64u : 0u);
                        const uint32_t b_hi = sbase + B_OFF + sb * B_STAGE;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint64_t da = umma_desc_k64_3(a_tap + k * 32, SBO_A);
                            const uint64_t db = umma_desc(b_hi + k * 2 * LBO_B, LBO_B, SBO_B);
                            if (SPLIT) {
                                const uint64_t dal = umma_desc_k64_3(a_tap + A_PLANE + k * 32, SBO_A);
                                const uint64_t dbl = umma_desc(b_hi + B_HALF + k * 2 * LBO_B, LBO_B, SBO_B);
                                if (NACC == 1) {
                                    umma_tf32(t_set, dal, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                                    umma_tf32(t_set, da, dbl, IDESC, 1u);
                                    umma_tf32(t_set, da, db, IDESC, 1u);
                                } else {
                                    const uint32_t t_small = t_set + NMAIN * BN;
                                    const uint32_t t_main = t_set + (NMAIN > 1 ? (it & 1) * BN : 0);
                                    umma_tf32(t_small, dal, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                                    umma_tf32(t_small, da, dbl, IDESC, 1u);
                                    umma_tf32(t_main, da, db, IDESC, (it >= NMAIN || k > 0) ? 1u : 0u);
                                }
                            } else {
                                umma_tf32(t_set, da, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        umma_commit(bar_eb + 8 * sb);                          // weight stage free
                        if (++sb == NSB) { sb = 0; fb_par ^= 1u; }
                    }
                    umma_commit(bar_ea + 8 * sa);                              // activation stage free
                    if (++sa == NSA) { sa = 0; fa_par ^= 1u; }
                }
                umma_commit(bar_accf + 8 * b);                                 // accumulator set complete
            }
            T3_DBG(2, clock64() - t_begin); T3_DBG(3, w0); T3_DBG(4, w1); T3_DBG(5, w2);
        }
        __syncwarp();
    } else if (warp == T3_WLDA) {
        // ===== activation stream: one tensor-map box per 16-channel chunk (two for BNAPPLY: g and z) =====
        if (lane == 0) {
            int sa = 0;
            unsigned ea_par = 1;
            long long g = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int n_img = tile / tiles_hw;
                const int trem = tile - n_img * tiles_hw;
                const int th0 = (trem / tiles_w) << 4, tw0 = (trem % tiles_w) << 3;
                for (int kc = 0; kc < KC; ++kc, ++g) {
                    if (g >= NSA) T3_WAIT(bar_ea + 8 * sa, ea_par, w0);
                    const uint32_t bb = bar_raw + 8 * sa;
                    const uint32_t dst = sbase + sa * A_STAGE;
                    mbar_expect_tx(bb, Cfg::A_TX);
                    tma_load_4d(dst, &tm_x, kc * 16, tw0 - PAD, th0 - PAD, n_img, bb);
                    if (BNAPPLY) tma_load_4d(dst + A_PLANE, &tm_z, kc * 16, tw0 - PAD, th0 - PAD, n_img, bb);
                    if (++sa == NSA) { sa = 0; ea_par ^= 1u; }
                }
            }
            T3_DBG(6, clock64() - t_begin); T3_DBG(7, w0);
        }
        __syncwarp();
    } else if (warp == T3_WLDB) {
        // ===== weight stream: one cp.async.bulk per (chunk, tap) stage; packed blocks are [tap][Cin/32][8 quads][BN][4] =====
        if (lane == 0) {
            const int KC32 = a.Cin >> 5;
            int sb = 0;
            unsigned eb_par = 1;
            long long g = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kc = 0; kc < KC; ++kc) {
                    for (int tap = 0; tap < TAPS; ++tap, ++g) {
                        if (g >= NSB) T3_WAIT(bar_eb + 8 * sb, eb_par, w0);
                        const size_t off = ((size_t)tap * KC32 + (kc >> 1)) * (size_t)(BN * 32) + (size_t)(kc & 1) * (BN * 16);
                        const uint32_t bb = bar_fb + 8 * sb;
                        const uint32_t dst = sbase + B_OFF + sb * B_STAGE;
                        mbar_expect_tx(bb, B_STAGE);
                        bulk_g2s(dst, args.w_hi + off, B_HALF, bb);
                        if (SPLIT) bulk_g2s(dst + B_HALF, args.w_lo + off, B_HALF, bb);
                        if (++sb == NSB) { sb = 0; eb_par ^= 1u; }
                    }
                }
            }
            T3_DBG(8, clock64() - t_begin); T3_DBG(9, w0);
        }
        __syncwarp();
    } else if (warp >= T3_WEPI) {
        // ===== epilogue warps: TMEM -> registers -> 32x32 patch -> bias / shortcut / accumulate -> coalesced store, statistics =====
        const int we = warp - T3_WEPI;
        const int lq = warp & 3;                 // TMEM lane quarter this warp may access = 32 tile rows
        const int half = we >> 2;                // column half of the tile
        constexpr int CPW = Cfg::CPW, NBLK = Cfg::NBLK;
        float* patch = reinterpret_cast<float*>(sgen + Cfg::PIPE + we * T3_PATCH);
        double* stat = reinterpret_cast<double*>(sgen + Cfg::PIPE + T3_NEW * T3_PATCH) + (size_t)we * CPW * 2;
        const bool do_stats = a.stat_sum != nullptr;
        const bool has_res = a.res.z != nullptr, res_aff = a.res.scale != nullptr;
        const bool has_acc = a.accumulate != 0;
        for (int c = lane; c < CPW * 2; c += 32) stat[c] = 0.0;
        __syncwarp();
        const int rsub = lane >> 3, cq = lane & 7;   // row loop: lane -> (row rsub + 4 i, column quad cq)
        const unsigned uW = (unsigned)a.W, uCout = (unsigned)a.Cout;
        const unsigned col_base = (unsigned)(half * CPW + cq * 4);

        auto tile_loop = [&](auto r_tag, auto a_tag) {
            constexpr bool R = decltype(r_tag)::value, A = decltype(a_tag)::value;
            constexpr int NL = (R ? 1 : 0) + (A ? 1 : 0) + (BWDSTATS ? 1 : 0);
            constexpr int RB = NL <= 1 ? 8 : (NL == 2 ? 4 : 2);     // rows per step (prefetch registers: NL * RB float4)
            constexpr int SPB = 8 / RB, NSTEP = NBLK * SPB;
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                const int b = i & 1;
                const int n_img = tile / tiles_hw;
                const int trem = tile - n_img * tiles_hw;
                const int th0 = (trem / tiles_w) << 4, tw0 = (trem % tiles_w) << 3;
                // tile row r = lq*32 + rsub + 4*ri -> pixel (th0 + lq*4 + (ri>>1), tw0 + rsub + 4*(ri&1))
                const unsigned pix_base = (unsigned)((n_img * a.H + th0 + lq * 4) * a.W + tw0 + rsub);
                float4 rr[R ? RB : 1], oo[A ? RB : 1], zz[BWDSTATS ? RB : 1];
                auto off_of = [&](int st, int k) -> unsigned {
                    const int blk = st / SPB, ri = (st % SPB) * RB + k;
                    return (pix_base + (unsigned)(ri >> 1) * uW + (unsigned)((ri & 1) * 4)) * uCout + col_base + (unsigned)(blk * 32);
                };
                auto issue = [&](int st, int k) {
                    const unsigned off = off_of(st, k);
                    if (R) rr[R ? k : 0] = ldg4(a.res.z + off);
                    if (A) oo[A ? k : 0] = ld4(a.y + off);
                    if (BWDSTATS) zz[BWDSTATS ? k : 0] = ldg4(a.bz + off);
                };
                if (NL > 0) {
#pragma unroll
                    for (int k = 0; k < RB; ++k) issue(0, k);
                }
                T3_WAIT(bar_accf + 8 * b, (unsigned)(i >> 1) & 1u, w0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), rs, rt;
                float4 bsc = make_float4(1.f, 1.f, 1.f, 1.f), bsh = make_float4(0.f, 0.f, 0.f, 0.f), bmu = bsh, biv = bsc;
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                for (int st = 0; st < NSTEP; ++st) {
                    const int blk = st / SPB;
                    if (st % SPB == 0) {
                        // accumulator block (32 rows x 32 columns) -> transpose patch
                        const long long te0 = dbg_on ? clock64() : 0;
                        const int c0 = half * CPW + blk * 32;
#pragma unroll
                        for (int h16 = 0; h16 < 2; ++h16) {
                            uint32_t r[16];
                            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * SUBCOLS + c0 + h16 * 16);
                            tmem_ld16(taddr, r);
                            float acc[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) acc[q] = __uint_as_float(r[q]);
#pragma unroll
                            for (int e = 1; e < NACC; ++e) {
                                tmem_ld16(taddr + (uint32_t)(e * BN), r);
#pragma unroll
                                for (int q = 0; q < 16; ++q) acc[q] += __uint_as_float(r[q]);
                            }
                            float* dst = patch + lane * 36 + h16 * 16;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                st4(dst + q * 4, make_float4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]));
                        }
                        if (blk == NBLK - 1) {       // accumulator set fully read: hand it back to the MMA warp
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            mbar_arrive(bar_acce + 8 * b);
                        }
                        const int n = c0 + cq * 4;
                        if (a.bias != nullptr) bv = ldg4(a.bias + n);
                        if (R) load_affine4(a.res.scale, a.res.shift, n, rs, rt);
                        if (BWDSTATS) { bsc = ldg4(a.bscale + n); bsh = ldg4(a.bshift + n); bmu = ldg4(a.bmean + n); biv = ldg4(a.binvstd + n); }
                        __syncwarp();
                        if (dbg_on) w1 += clock64() - te0;
                    }
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        const int ri = (st % SPB) * RB + k;
                        const unsigned off = off_of(st, k);
                        float4 v = ld4(patch + (rsub + 4 * ri) * 36 + cq * 4);
                        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        if (R) {
                            float4 q = rr[R ? k : 0];
                            if (res_aff) q = act4(q, rs, rt, a.res.relu);
                            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                        }
                        if (A) { const float4 o = oo[A ? k : 0]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        st4(a.y + off, v);
                        if (do_stats) {
                            if (BWDSTATS) {
                                const float4 z = zz[BWDSTATS ? k : 0];
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
                        if (NL > 0 && st + 1 < NSTEP) issue(st + 1, k);      // next step's loads behind this row's use
                    }
                    if (st % SPB == SPB - 1) {
                        __syncwarp();                    // the patch is rewritten by the next block
                        if (do_stats) {
                            // fp32 partial sums over this lane's 8 rows; 4 lanes (rsub) share a column quad: shuffle, then fp64
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8);
                                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);
                                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
                            }
                            if (rsub == 0) {
                                double* sp = stat + (size_t)(blk * 32 + cq * 4) * 2;
#pragma unroll
                                for (int j = 0; j < 4; ++j) { sp[j * 2 + 0] += (double)s1[j]; sp[j * 2 + 1] += (double)s2[j]; }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
                        }
                    }
                }
            }
        };
        if (has_res) {
            if (has_acc) tile_loop(std::true_type{}, std::true_type{}); else tile_loop(std::true_type{}, std::false_type{});
        } else {
            if (has_acc) tile_loop(std::false_type{}, std::true_type{}); else tile_loop(std::false_type{}, std::false_type{});
        }
        if (we == 0 && lane == 0) { T3_DBG(10, clock64() - t_begin); T3_DBG(11, w0); T3_DBG(16, w1); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == T3_WAUX)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    // ---- statistics: per-warp fp64 partial sums (shared memory) -> one global atomic per channel per CTA; last CTA finalises ----
    if (a.stat_sum != nullptr) {
        const double* stat0 = reinterpret_cast<const double*>(sgen + Cfg::PIPE + T3_NEW * T3_PATCH);
        if (blockIdx.x < ntiles) {
            for (int c = tid; c < BN; c += T3_THREADS) {
                const int hf = c / Cfg::CPW, cl = c - hf * Cfg::CPW;
                double x1 = 0.0, x2 = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double* sp = stat0 + ((size_t)(hf * 4 + q) * Cfg::CPW + cl) * 2;
                    x1 += sp[0];
                    x2 += sp[1];
                }
                atomicAdd(a.stat_sum + c, x1);
                atomicAdd(a.stat_sq + c, x2);
            }
        }
        if (!BWDSTATS && a.ffin.ticket != nullptr) {
            if (last_cta_arrives(a.ffin.ticket, gridDim.x)) bn_fwd_finalize_cta(a.ffin, a.stat_sum, a.stat_sq, (double)a.P, a.Cout);
        } else if (BWDSTATS && a.bfin.ticket != nullptr) {
            if (last_cta_arrives(a.bfin.ticket, gridDim.x)) bn_bwd_finalize_cta(a.bfin, a.stat_sum, a.stat_sq, (double)a.P, a.Cout);
        }
    }
    if (tid == 0) { T3_DBG(15, gtimer()); T3_DBG(17, clock64() - t_begin); }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 4-D tensor map over an NHWC fp32 tensor, box = 16 channels x box_w x box_h x 1 image, 64-byte swizzle
static int make_map(CUtensorMap* m, const float* base, int N, int H, int W, int C, int box_w, int box_h) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) {
        set_error("hgk_conv_tc_nhwc: cuTensorMapEncodeTiled is not available from the driver");
        return HGK_ECUDA;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {16u, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("hgk_conv_tc_nhwc: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box %dx%d", (int)r, N, H, W, C, box_w, box_h);
        return HGK_ECUDA;
    }
    return HGK_OK;
}

template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY>
static int launch_tc3(TcArgs ta, cudaStream_t st) {
    using Cfg = T3Cfg<BN, SPLIT, KS, BNAPPLY>;
    static bool configured = false;
    constexpr int smem = Cfg::SMEM;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<BN, SPLIT, KS, BWDSTATS, BNAPPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_tc_nhwc (persistent tile kernel): cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    if (KS == 1) {          // a 1x1 convolution has no neighbourhood: any 128 consecutive pixels are a tile
        ta.c.H = (int)(ta.c.P / 8);
        ta.c.W = 8;
        ta.c.N = 1;
    }
    CUtensorMap mx, mz;
    int rc = make_map(&mx, ta.c.x.z, ta.c.N, ta.c.H, ta.c.W, ta.c.Cin, Cfg::HWD, Cfg::HH);
    if (rc != HGK_OK) return rc;
    if (BNAPPLY) {
        rc = make_map(&mz, ta.c.ap.z, ta.c.N, ta.c.H, ta.c.W, ta.c.Cin, Cfg::HWD, Cfg::HH);
        if (rc != HGK_OK) return rc;
    } else {
        mz = mx;
    }
    const long long tiles = (long long)ta.c.N * (ta.c.H >> 4) * (ta.c.W >> 3);
    const long long rounds = (tiles + kNumSMs - 1) / kNumSMs;
    const unsigned grid = (unsigned)((tiles + rounds - 1) / rounds);
    conv_tc3_kernel<BN, SPLIT, KS, BWDSTATS, BNAPPLY><<<grid, T3_THREADS, smem, st>>>(ta, mx, mz);
    return HGK_OK;
}

template <int BN, int KS>
static int launch_tc3_bn(const TcArgs& ta, bool split, bool bwdstats, cudaStream_t st) {
    if (ta.c.ap.z != nullptr)        // data gradient with the BatchNorm-backward apply evaluated on load
        return bwdstats ? launch_tc3<BN, false, KS, true, true>(ta, st) : launch_tc3<BN, false, KS, false, true>(ta, st);
    if (bwdstats) return launch_tc3<BN, false, KS, true, false>(ta, st);
    if (split) return launch_tc3<BN, true, KS, false, false>(ta, st);
    return launch_tc3<BN, false, KS, false, false>(ta, st);
}

// true when the persistent tile kernel covers this problem
bool conv_tc3_eligible(const TcArgs& ta) {
    const ConvArgs& c = ta.c;
    if (c.ksize == 3) {
        if (c.H % 16 || c.W % 8) return false;
        if (c.Cout > 128) return false;                                   // 3x3 with 256 outputs: no instantiation
    } else {
        if (c.P % 128) return false;
    }
    const long long cmax = c.Cin > c.Cout ? c.Cin : c.Cout;
    if (c.P * cmax >= (1LL << 32)) return false;                          // 32-bit element offsets
    if (((uintptr_t)c.x.z & 15) || (c.ap.z != nullptr && ((uintptr_t)c.ap.z & 15))) return false;   // tensor-map base alignment
    return true;
}

int conv_tc3_launch(const TcArgs& ta, bool split, bool bwdstats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int BN = ta.c.Cout;
    if (ta.c.ksize == 3) {
        if (BN == 64) return launch_tc3_bn<64, 3>(ta, split, bwdstats, st);
        return launch_tc3_bn<128, 3>(ta, split, bwdstats, st);
    }
    if (BN == 64) return launch_tc3_bn<64, 1>(ta, split, bwdstats, st);
    if (BN == 128) return launch_tc3_bn<128, 1>(ta, split, bwdstats, st);
    return launch_tc3_bn<256, 1>(ta, split, bwdstats, st);
}

}  // namespace hgk
