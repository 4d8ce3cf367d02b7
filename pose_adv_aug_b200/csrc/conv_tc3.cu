// Persistent, phase-overlapped tcgen05 implicit-GEMM convolution, CHANNELS ON THE TMEM LANES.
//
// Same contract as conv_tc2_kernel (conv_tc2.cu): BN+ReLU on load, 3xTF32 (forward) or plain TF32 (data gradient),
// optional BatchNorm-backward apply on load (BNAPPLY), bias / shortcut / accumulate epilogue with the BatchNorm
// statistics (or the fused BatchNorm-backward reduction, BWDSTATS) and the last-CTA finaliser.  What changed is the
// execution structure (profiles/r2_tile_kernel_timeline.md: prologue + epilogue were 40-55 % of a CTA's life, the two
// CTAs of an SM ran in lock-step; profiles/r3_persistent_kernel.md: the first persistent version was bound by shared-
// memory bandwidth and by its transpose-through-shared-memory epilogue):
//   * ONE CTA per SM walks over its tiles (tile = blockIdx.x + i * gridDim.x); nothing drains between tiles;
//   * the GEMM is computed TRANSPOSED: D[channel][pixel] = W[channel][k] * X[pixel][k]^T, i.e. the packed weights are the
//     M = 128 operand (two MMAs for 256 output channels) and the activation tile is the N = 128-pixel operand.  TMEM lanes
//     are output channels, so an epilogue thread owns ONE channel: per-channel bias / BatchNorm vectors are scalars, the
//     BatchNorm statistics are plain per-thread sums (no cross-lane reduction, no shared memory), and the NHWC stores
//     of a warp are 32 consecutive channels of one pixel = one full 128-byte line;
//   * the activation halo tile of a 16-channel chunk is fetched by TMA (cp.async.bulk.tensor, 4-D tensor map over
//     [N][H][W][C], out-of-bounds rows / columns zero-filled by the hardware) DIRECTLY into the K-major 64-byte-swizzle
//     UMMA layout (CU_TENSOR_MAP_SWIZZLE_64B is that layout), several chunks ahead;
//   * eight transform warps turn the landed tile IN PLACE into the TF32 "hi" operand (BN+ReLU, round) and write the
//     "lo" operand (3xTF32) next to it; for BNAPPLY the second plane receives the raw z tile (second tensor map) and
//     the transform forms dz = cA*((g*mask - cC) - (z - mean)*cB), writes it back to global once, and rounds it;
//   * one thread issues tcgen05.mma.kind::tf32 into one of TWO TMEM accumulator sets; the nine taps of a 3x3 are nine
//     descriptors into the same halo tile (start address shifted by (dh*HWD + dw)*64 bytes);
//   * eight epilogue warps drain accumulator set i&1 (tcgen05.ld -> registers -> coalesced stores) while the main loop
//     of tile i+1 runs; the shortcut / accumulate / BN-backward input tiles of the NEXT tile are pulled into L2 by
//     cp.async.bulk.prefetch one tile ahead.
// 1x1 convolutions see the tensor as [1][P/8][8][C] (any H, W with P % 8 == 0: also the 8x8 / 4x4 rungs; the ragged last
// tile is zero-filled by TMA and masked by the epilogue).  Output channels: 128 or 256 (64-channel layers: conv_tc2.cu).
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

__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

#ifndef T3_KB1
#define T3_KB1 2
#endif
constexpr int T3_NTW = 8;                                   // transform warps
constexpr int T3_NEW = 8;                                   // epilogue warps
constexpr int T3_WMMA = T3_NTW, T3_WLDA = T3_NTW + 1, T3_WLDB = T3_NTW + 2, T3_WAUX = T3_NTW + 3;
constexpr int T3_WEPI = T3_NTW + 4;                         // first epilogue warp (multiple of 4: TMEM lane quarter = warp % 4)
constexpr int T3_THREADS = (T3_WEPI + T3_NEW) * 32;         // 640
constexpr int T3_NTT = T3_NTW * 32;                         // transform threads
static_assert(T3_WEPI % 4 == 0, "epilogue warps must start at a multiple of four");

template <int BN, bool SPLIT, int KS, bool BNAPPLY>
struct T3Cfg {
    static constexpr int NM = BN / 128;                                      // M = 128-channel halves of the output
    // accumulators per (half, pixel sub-tile), see conv_tc.cu (the tensor core truncates its fp32 accumulator): the long 3x3
    // 3xTF32 chain keeps hi*hi and the cross terms apart; the epilogue sums them in fp32
    static constexpr int NACC = (SPLIT && KS == 3) ? 2 : 1;
    static constexpr int NSUBS = 512 / (256 * NM * NACC);                    // 128-pixel sub-tiles per tile (two accumulator sets fit)
    static constexpr int TH = KS == 1 ? 16 * NSUBS : 16, TW = KS == 1 ? 8 : 8 * NSUBS;
    static constexpr int NPX = TH * TW;
    static constexpr int PAD = KS / 2;
    static constexpr int HH = TH + 2 * PAD, HWD = TW + 2 * PAD;
    static constexpr int NPIX = HH * HWD;
    static constexpr int SUB_OFF = KS == 1 ? 128 * 64 : 8 * 64;              // byte offset of sub-tile 1 inside a plane
    static constexpr int TAPS = KS * KS;
    static constexpr int NITEM = NPIX * 4;
    static constexpr int NJ = (NITEM + T3_NTT - 1) / T3_NTT;
    static constexpr int SBO_X = HWD * 64;
    static constexpr int A_BOX = NPIX * 64;                                  // bytes one tensor-map box delivers
    static constexpr int A_PLANE = (A_BOX + 1023) / 1024 * 1024;
    static constexpr int NPLANE = (SPLIT || BNAPPLY) ? 2 : 1;                // plane 1: lo operand (SPLIT) or raw z (BNAPPLY)
    static constexpr int CPS = (KS == 1 && NSUBS == 1) ? 2 : 1;              // 16-channel chunks per activation stage
    static constexpr int A_STAGE = NPLANE * CPS * A_PLANE;
    static constexpr int A_TX = (BNAPPLY ? 2 : 1) * CPS * A_BOX;
    static constexpr int B_HALF = BN * 16 * 4;
    static constexpr int B_STAGE = (SPLIT ? 2 : 1) * B_HALF;
    static constexpr int SETCOLS = NM * NACC * NSUBS * 128;                  // TMEM columns of one accumulator set
    static constexpr int TMEM_COLS = (2 * SETCOLS <= 256) ? 256 : 512;
    static constexpr int NVEC = BNAPPLY ? 6 : 2;                             // per-input-channel vectors staged in shared memory
    static constexpr int VEC_BYTES = NVEC * 256 * 4;                         // Cin <= 256
    static constexpr int AVAIL = 227 * 1024 - 1024 - 1024 - VEC_BYTES;
    static constexpr int NSB = KS == 3 ? 9 : (B_STAGE >= 32768 ? 3 : 4);
    static constexpr int NSA_FIT = (AVAIL - NSB * B_STAGE) / A_STAGE;
    static constexpr int NSA = NSA_FIT > 6 ? 6 : NSA_FIT;
    static constexpr int PIPE = NSA * A_STAGE + NSB * B_STAGE;
    static constexpr int SMEM = PIPE + VEC_BYTES + 1024;
    static constexpr int NSTEP = NM * NSUBS * 8 / 2;                         // 16-pixel x 32-channel steps per epilogue warp and tile
    static_assert(BN == 128 || BN == 256, "output channels on the TMEM lanes: multiples of 128");
    static_assert(NSUBS >= 1 && 2 * SETCOLS <= 512, "TMEM capacity (two accumulator sets)");
    static_assert(NSA >= 3, "activation ring: one stage in the tensor core, one in the transform, one in flight");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

// XBF (1x1 forward only): TF32 + 2xBF16 products as in conv_tc2.cu -- the lo plane of an activation chunk holds the bf16 tiles
// xl | xh in the K-major no-swizzle layout [k/8 (2)][pixel][8 x bf16], the lo half of a weight stage the pack-mode-2 pair
// [wh_bf16 | wl_bf16]; per 16-channel chunk two kind::f16 MMAs (K = 16) replace four of the six kind::tf32 ones.
template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY, bool XBF = false>
__global__ void __launch_bounds__(T3_THREADS, 1)
conv_tc3_kernel(const TcArgs args, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_z) {
    static_assert(!(BNAPPLY && SPLIT), "the fused BatchNorm-backward apply is a data-gradient (plain TF32) mode");
    static_assert(!XBF || (SPLIT && KS == 1), "bf16 cross terms: 1x1 error-compensated forward only");
    constexpr uint32_t XPLANE = (uint32_t)T3Cfg<BN, SPLIT, KS, BNAPPLY>::NPIX * 16u;
    constexpr uint32_t IDESC_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    using Cfg = T3Cfg<BN, SPLIT, KS, BNAPPLY>;
    constexpr int NJ = Cfg::NJ, NSA = Cfg::NSA, NSB = Cfg::NSB, NACC = Cfg::NACC, NM = Cfg::NM, NSUBS = Cfg::NSUBS, CPS = Cfg::CPS;
    constexpr int HWD = Cfg::HWD, PAD = Cfg::PAD, TAPS = Cfg::TAPS, SETCOLS = Cfg::SETCOLS, TH = Cfg::TH, TW = Cfg::TW;
    constexpr uint32_t SBO_X = Cfg::SBO_X, LBO_W = BN * 16, SBO_W = 128;
    constexpr uint32_t A_PLANE = Cfg::A_PLANE, A_STAGE = Cfg::A_STAGE, B_HALF = Cfg::B_HALF, B_STAGE = Cfg::B_STAGE;
    constexpr uint32_t B_OFF = NSA * A_STAGE;
    // D = f32 (bit 4), A = B = tf32 (2 << 7, 2 << 10), both K-major, N = 128 pixels (>> 3 at bit 17), M = 128 channels (>> 4 at bit 24)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const ConvArgs& a = args.c;

    extern __shared__ uint8_t smem_raw[];
    // barriers: raw[NSA] (TMA landed) | afull[NSA] (transformed) | aempty[NSA] (MMAs retired) | bfull[NSB] | bempty[NSB] |
    //           accfull[2] | accempty[2]
    __shared__ __align__(8) uint64_t bars[3 * NSA + 2 * NSB + 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_w = a.W / TW, tiles_hw = ((a.H + TH - 1) / TH) * tiles_w;      // KS == 1: the last row tile may be ragged
    const int ntiles = a.N * tiles_hw;
    const int KR = a.Cin / (16 * CPS);                                              // activation stages (rounds) per tile
    const uint32_t bar_raw = smem_u32(&bars[0]), bar_fa = smem_u32(&bars[NSA]), bar_ea = smem_u32(&bars[2 * NSA]);
    const uint32_t bar_fb = smem_u32(&bars[3 * NSA]), bar_eb = smem_u32(&bars[3 * NSA + NSB]);
    const uint32_t bar_accf = smem_u32(&bars[3 * NSA + 2 * NSB]), bar_acce = smem_u32(&bars[3 * NSA + 2 * NSB + 2]);

    if (tid == 0) {
        for (int s = 0; s < NSA; ++s) {
            mbar_init(bar_raw + 8 * s, 1);
            mbar_init(bar_fa + 8 * s, T3_NTW);
            mbar_init(bar_ea + 8 * s, 1);
        }
        for (int s = 0; s < NSB; ++s) {
            mbar_init(bar_fb + 8 * s, 1);
            mbar_init(bar_eb + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, T3_NEW);
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
    // programmatic dependent launch: everything above ran under the previous kernel's tail; nothing below may start before
    // that kernel has completed (first global reads: the per-channel vectors)
    pdl_wait();
    // per-input-channel vectors of the on-load transform, staged once (with 227 KB of shared memory the L1 has no room
    // for them: read through __ldg they cost an L2 round trip per 16-channel chunk on the transform's critical path)
    float* svec = reinterpret_cast<float*>(sgen + Cfg::PIPE);               // [NVEC][256]
    for (int c = tid; c < a.Cin; c += T3_THREADS) {
        if (!BNAPPLY) {
            svec[c] = a.x.scale != nullptr ? __ldg(a.x.scale + c) : 1.f;
            svec[256 + c] = a.x.scale != nullptr ? __ldg(a.x.shift + c) : 0.f;
        } else {
            svec[c] = __ldg(a.ap.scale + c);
            svec[256 + c] = __ldg(a.ap.shift + c);
            svec[512 + c] = __ldg(a.ap.mean + c);
            svec[768 + c] = __ldg(a.ap.cA + c);
            svec[1024 + c] = __ldg(a.ap.cB + c);
            svec[1280 + c] = __ldg(a.ap.cC + c);
        }
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
                if (hh >= PAD && hh < PAD + TH && ww >= PAD && ww < PAD + TW) imask |= 1u << j;
            }
        }
        constexpr bool ALLV = (Cfg::NITEM % T3_NTT) == 0;       // every item slot of every thread is a real item
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (!ALLV && !((smask >> j) & 1u)) s_off[j] = s_off[0];     // idle slots re-read item 0 (never stored)
        // identity affine = (1, 0, clamp -inf): exact, so the on-load activation is applied without a branch
        const float x_clamp = (a.x.scale != nullptr && a.x.relu) ? 0.f : -INFINITY;
        const BnApply& ap = a.ap;
        const float ap_thr = ap.relu ? 0.f : -INFINITY;
        int sa = 0;
        unsigned par = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int n_img = tile / tiles_hw;
            const int trem = tile - n_img * tiles_hw;
            const int th0 = (trem / tiles_w) * TH, tw0 = (trem % tiles_w) * TW;
            unsigned vmask = 0;                 // pixel inside the image (zero padding is a zero of the ACTIVATED tensor)
            unsigned g_off[BNAPPLY ? NJ : 1];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int h = th0 - PAD + (hw_j[j] & 255), w = tw0 - PAD + (hw_j[j] >> 8);
                if (((smask >> j) & 1u) && (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W) vmask |= 1u << j;
                if (BNAPPLY) g_off[j] = (unsigned)((n_img * a.H + h) * a.W + w) * (unsigned)a.Cin + (unsigned)quad * 4u;
            }
            const unsigned wmask = vmask & imask;       // BNAPPLY: items whose dz this tile writes back
            for (int kr = 0; kr < KR; ++kr) {
                T3_WAIT(bar_raw + 8 * sa, par, w0);
                const long long tc0 = dbg_on ? clock64() : 0;
#pragma unroll
                for (int cc = 0; cc < CPS; ++cc) {
                    const int kc = kr * CPS + cc;
                    const float* vp = svec + kc * 16 + quad * 4;
                    float4 sc, sh, bs, bt, bmu, bA, bB, bC;
                    if (!BNAPPLY) {
                        sc = ld4(vp); sh = ld4(vp + 256);
                    } else {
                        bs = ld4(vp); bt = ld4(vp + 256); bmu = ld4(vp + 512);
                        bA = ld4(vp + 768); bB = ld4(vp + 1024); bC = ld4(vp + 1280);
                    }
                    uint8_t* base = sgen + sa * A_STAGE + cc * A_PLANE;
                    uint8_t* base2 = base + CPS * A_PLANE;
                    // items in batches of JB: shared-memory loads first, then the arithmetic (BNAPPLY: small batches, the six
                    // per-channel vectors already take 24 registers)
                    constexpr int JB = BNAPPLY ? 2 : NJ;
#pragma unroll
                    for (int j0 = 0; j0 < NJ; j0 += JB) {
                        float4 vv[JB], zz[BNAPPLY ? JB : 1];
#pragma unroll
                        for (int jj = 0; jj < JB; ++jj) {
                            const int j = j0 + jj;
                            if (j >= NJ) break;
                            vv[jj] = ld4(reinterpret_cast<const float*>(base + s_off[j]));
                            if (BNAPPLY) zz[BNAPPLY ? jj : 0] = ld4(reinterpret_cast<const float*>(base2 + s_off[j]));
                        }
#pragma unroll
                        for (int jj = 0; jj < JB; ++jj) {
                            const int j = j0 + jj;
                            if (j >= NJ) break;
                            const bool in_img = (vmask >> j) & 1u;
                            float4 v = vv[jj];
                            if (!BNAPPLY) {
                                const float4 t = actc4(v, sc, sh, x_clamp);
                                v.x = in_img ? t.x : 0.f; v.y = in_img ? t.y : 0.f; v.z = in_img ? t.z : 0.f; v.w = in_img ? t.w : 0.f;
                            } else {
                                // dz = cA * ((g * relu-mask - cC) - (z - mean) * cB)   (bn_bwd_apply_kernel, bn.cu)
                                const float4 z = zz[BNAPPLY ? jj : 0];
                                const float gx = fmaf(z.x, bs.x, bt.x) <= ap_thr ? 0.f : v.x;
                                const float gy = fmaf(z.y, bs.y, bt.y) <= ap_thr ? 0.f : v.y;
                                const float gz = fmaf(z.z, bs.z, bt.z) <= ap_thr ? 0.f : v.z;
                                const float gw = fmaf(z.w, bs.w, bt.w) <= ap_thr ? 0.f : v.w;
                                v.x = in_img ? bA.x * ((gx - bC.x) - (z.x - bmu.x) * bB.x) : 0.f;
                                v.y = in_img ? bA.y * ((gy - bC.y) - (z.y - bmu.y) * bB.y) : 0.f;
                                v.z = in_img ? bA.z * ((gz - bC.z) - (z.z - bmu.z) * bB.z) : 0.f;
                                v.w = in_img ? bA.w * ((gw - bC.w) - (z.w - bmu.w) * bB.w) : 0.f;
                                if ((wmask >> j) & 1u) st4(ap.dz + (g_off[BNAPPLY ? j : 0] + (unsigned)kc * 16u), v);
                            }
                            const float4 hi = tf32_rna4(v);
                            if (ALLV || ((smask >> j) & 1u)) {
                                *reinterpret_cast<float4*>(base + s_off[j]) = hi;
                                if (SPLIT && !XBF) {
                                    const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                                    *reinterpret_cast<float4*>(base2 + s_off[j]) = lo;
                                }
                                if (XBF) {      // plane = channel quad / 2, row = pixel (16 B), 8-byte half = quad & 1
                                    const uint32_t boff = (uint32_t)(quad >> 1) * XPLANE + (uint32_t)((tid + T3_NTT * j) >> 2) * 16u +
                                                          (uint32_t)(quad & 1) * 8u;
                                    uint2 xl, xh;
                                    xl.x = pack_bf16x2(v.x - hi.x, v.y - hi.y); xl.y = pack_bf16x2(v.z - hi.z, v.w - hi.w);
                                    xh.x = pack_bf16x2(hi.x, hi.y); xh.y = pack_bf16x2(hi.z, hi.w);
                                    *reinterpret_cast<uint2*>(base2 + boff) = xl;
                                    *reinterpret_cast<uint2*>(base2 + 2u * XPLANE + boff) = xh;
                                }
                            }
                        }
                    }
                }
                const long long tc1 = dbg_on ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_fa + 8 * sa);                   // one arrival per warp
                if (dbg_on) { w1 += tc1 - tc0; w2 += clock64() - tc1; }
                if (++sa == NSA) { sa = 0; par ^= 1u; }
            }
        }
        if (tid == 0) { T3_DBG(0, clock64() - t_begin); T3_DBG(1, w0); T3_DBG(12, w1); T3_DBG(13, w2); }
    } else if (warp < T3_WEPI) {
    // register budget per role (setmaxnreg acts on aligned groups of four warps and the compiler budgets the code that follows
    // it on every path, so it sits inside the role branch): 640 x 96 registers are allocated at launch; the four support warps
    // keep 48, the eight epilogue warps take 120 (two prefetch register sets): 128*48 + 256*96 + 256*120 = 61440
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
    if (warp == T3_WMMA) {
        // ===== MMA issuer: D[channel][pixel] += W[channel][k] * X[pixel][k] =====
        if (lane == 0) {
            int sa = 0, sb = 0;
            unsigned fa_par = 0, fb_par = 0;
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                const int b = i & 1;
                if (i >= 2) T3_WAIT(bar_acce + 8 * b, (unsigned)((i >> 1) - 1) & 1u, w2);   // epilogue drained this accumulator set
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_set = tmem + (uint32_t)(b * SETCOLS);
                bool first = true;                                             // first k-step of the tile overwrites the accumulators
                for (int kr = 0; kr < KR; ++kr) {
                    T3_WAIT(bar_fa + 8 * sa, fa_par, w0);
#pragma unroll 1
                    for (int cc = 0; cc < CPS; ++cc) {
                        const uint32_t x_hi0 = sbase + sa * A_STAGE + cc * A_PLANE;
#pragma unroll 1
                        for (int tap = 0; tap < TAPS; ++tap) {
                            T3_WAIT(bar_fb + 8 * sb, fb_par, w1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t x_tap = x_hi0 + (KS == 3 ? (uint32_t)((tap / 3) * HWD + (tap % 3)) * 64u : 0u);
                            const uint32_t w_hi0 = sbase + B_OFF + sb * B_STAGE;
                            if (XBF) {
                                // cross terms of this chunk: wh_bf16 * xl and wl_bf16 * xh_bf16, one K = 16 MMA each
                                const uint32_t xb = x_hi0 + CPS * A_PLANE;
#pragma unroll
                                for (int s = 0; s < NSUBS; ++s) {
                                    const uint64_t dxl = umma_desc(xb + s * 128 * 16, XPLANE, 128u);
                                    const uint64_t dxh = umma_desc(xb + 2u * XPLANE + s * 128 * 16, XPLANE, 128u);
#pragma unroll
                                    for (int mh = 0; mh < NM; ++mh) {
                                        const uint64_t dwh = umma_desc(w_hi0 + B_HALF + mh * (128 * 16), LBO_W, SBO_W);
                                        const uint64_t dwl = umma_desc(w_hi0 + B_HALF + BN * 32 + mh * (128 * 16), LBO_W, SBO_W);
                                        const uint32_t t_main = t_set + (uint32_t)(((mh * NACC) * NSUBS + s) * 128);
                                        umma_bf16(t_main, dwh, dxl, IDESC_BF, first ? 0u : 1u);
                                        umma_bf16(t_main, dwl, dxh, IDESC_BF, 1u);
                                    }
                                }
                            }
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint32_t acc_flag = (first && k == 0) ? 0u : 1u;
#pragma unroll
                                for (int s = 0; s < NSUBS; ++s) {
                                    const uint64_t dx = umma_desc_k64_3(x_tap + s * Cfg::SUB_OFF + k * 32, SBO_X);
                                    const uint64_t dxl = umma_desc_k64_3(x_tap + CPS * A_PLANE + s * Cfg::SUB_OFF + k * 32, SBO_X);
#pragma unroll
                                    for (int mh = 0; mh < NM; ++mh) {
                                        const uint64_t dw = umma_desc(w_hi0 + mh * (128 * 16) + k * 2 * LBO_W, LBO_W, SBO_W);
                                        const uint32_t t_main = t_set + (uint32_t)(((mh * NACC) * NSUBS + s) * 128);
                                        if (XBF) {
                                            umma_tf32(t_main, dw, dx, IDESC, 1u);      // (the bf16 MMAs above initialised the tile)
                                        } else if (SPLIT) {
                                            const uint64_t dwl = umma_desc(w_hi0 + B_HALF + mh * (128 * 16) + k * 2 * LBO_W, LBO_W, SBO_W);
                                            const uint32_t t_small = t_set + (uint32_t)(((mh * NACC + (NACC - 1)) * NSUBS + s) * 128);
                                            if (NACC == 1) {
                                                umma_tf32(t_main, dw, dxl, IDESC, acc_flag);
                                                umma_tf32(t_main, dwl, dx, IDESC, 1u);
                                                umma_tf32(t_main, dw, dx, IDESC, 1u);
                                            } else {
                                                umma_tf32(t_small, dw, dxl, IDESC, acc_flag);
                                                umma_tf32(t_small, dwl, dx, IDESC, 1u);
                                                umma_tf32(t_main, dw, dx, IDESC, acc_flag);
                                            }
                                        } else {
                                            umma_tf32(t_main, dw, dx, IDESC, acc_flag);
                                        }
                                    }
                                }
                            }
                            first = false;
                            umma_commit(bar_eb + 8 * sb);                      // weight stage free
                            if (++sb == NSB) { sb = 0; fb_par ^= 1u; }
                        }
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
                const int th0 = (trem / tiles_w) * TH, tw0 = (trem % tiles_w) * TW;
                for (int kr = 0; kr < KR; ++kr, ++g) {
                    if (g >= NSA) T3_WAIT(bar_ea + 8 * sa, ea_par, w0);
                    const uint32_t bb = bar_raw + 8 * sa;
                    const uint32_t dst = sbase + sa * A_STAGE;
                    mbar_expect_tx(bb, Cfg::A_TX);
#pragma unroll
                    for (int cc = 0; cc < CPS; ++cc) {
                        const int c0 = (kr * CPS + cc) * 16;
                        tma_load_4d(dst + cc * A_PLANE, &tm_x, c0, tw0 - PAD, th0 - PAD, n_img, bb);
                        if (BNAPPLY) tma_load_4d(dst + (CPS + cc) * A_PLANE, &tm_z, c0, tw0 - PAD, th0 - PAD, n_img, bb);
                    }
                    if (++sa == NSA) { sa = 0; ea_par ^= 1u; }
                }
            }
            T3_DBG(6, clock64() - t_begin); T3_DBG(7, w0);
        }
        __syncwarp();
    } else if (warp == T3_WLDB) {
        // ===== weight stream: one cp.async.bulk per (chunk, tap) stage; packed blocks are [tap][Cin/32][8 quads][BN][4] =====
        if (lane == 0) {
            const int KC = a.Cin >> 4, KC32 = a.Cin >> 5;
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
    } else if (warp == T3_WAUX) {
        // ===== L2 prefetch of the epilogue's input tiles (shortcut / previous output / BN-backward z), one tile ahead =====
        const float* src = lane == 0 ? a.res.z : (lane == 1 ? (a.accumulate ? a.y : nullptr) : (lane == 2 && BWDSTATS ? a.bz : nullptr));
        if (lane < 3 && src != nullptr) {
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                // pace: the rows of tile i are requested when the main loop of tile i-1 has finished
                if (i >= 1) mbar_wait(bar_accf + 8 * ((i - 1) & 1), (unsigned)((i - 1) >> 1) & 1u);
                const int n_img = tile / tiles_hw;
                const int trem = tile - n_img * tiles_hw;
                const int th0 = (trem / tiles_w) * TH, tw0 = (trem % tiles_w) * TW;
                if (KS == 1) {
                    const long long p0 = (long long)tile * Cfg::NPX;
                    long long np = a.P - p0;
                    if (np > Cfg::NPX) np = Cfg::NPX;
                    l2_prefetch(src + p0 * a.Cout, (uint32_t)(np * a.Cout * 4));
                } else {
                    for (int r = 0; r < TH; ++r)
                        l2_prefetch(src + ((size_t)(n_img * a.H + th0 + r) * a.W + tw0) * a.Cout, (uint32_t)(TW * a.Cout * 4));
                }
            }
        }
        __syncwarp();
    }
    } else {
        // ===== epilogue warps: thread = output channel; 16 pixels (TMEM columns) per step =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 120;" ::: "memory");
        const int we = warp - T3_WEPI;
        const int lq = warp & 3;                 // TMEM lane quarter this warp may access = 32 channels of a 128-channel half
        const int hsel = we >> 2;                // the two warps of a lane quarter alternate over the 16-pixel steps
        constexpr int NSTEP = Cfg::NSTEP;
        const bool do_stats = a.stat_sum != nullptr;
        const bool has_res = a.res.z != nullptr, res_aff = a.res.scale != nullptr;
        const bool has_acc = a.accumulate != 0;
        const float* __restrict__ resp = a.res.z;
        const float* __restrict__ bzp = a.bz;
        float* __restrict__ yp = a.y;
        const unsigned row_stride = (unsigned)a.W * (unsigned)BN;      // Cout == BN: per-pixel offsets are immediates
        // shortcut activation as a branch-free (scale, shift, clamp): identity = (1, 0, -inf)
        const float r_clamp = (res_aff && a.res.relu) ? 0.f : -INFINITY;
        const float b_thr = a.brelu ? 0.f : -INFINITY;       // BN-backward ReLU mask: g = (bz*scale+shift <= thr) ? 0 : dy
        double d1[NM], d2[NM];
        float bv[NM], rs[NM], rt[NM], bsc[NM], bsh[NM], bmu[NM], biv[NM];
#pragma unroll
        for (int mh = 0; mh < NM; ++mh) {
            const int ch = mh * 128 + lq * 32 + lane;
            d1[mh] = 0.0; d2[mh] = 0.0;
            bv[mh] = a.bias != nullptr ? __ldg(a.bias + ch) : 0.f;
            rs[mh] = res_aff ? __ldg(a.res.scale + ch) : 1.f;
            rt[mh] = res_aff ? __ldg(a.res.shift + ch) : 0.f;
            bsc[mh] = 1.f; bsh[mh] = 0.f; bmu[mh] = 0.f; biv[mh] = 1.f;
            if (BWDSTATS) { bsc[mh] = __ldg(a.bscale + ch); bsh[mh] = __ldg(a.bshift + ch); bmu[mh] = __ldg(a.bmean + ch); biv[mh] = __ldg(a.binvstd + ch); }
        }
        // element offset of pixel q (0..15) of a step relative to the step's first pixel: compile-time for a 1x1 (consecutive
        // pixels), one runtime row stride for a 3x3 (two image rows of 8 pixels)
#define T3_QOFF(q) (KS == 1 ? (unsigned)((q) * BN) : ((q) >> 3) * row_stride + (unsigned)(((q) & 7) * BN))

        auto tile_loop = [&](auto r_tag, auto a_tag) {
            constexpr bool R = decltype(r_tag)::value, A = decltype(a_tag)::value;
            constexpr int NL = (R ? 1 : 0) + (A ? 1 : 0) + (BWDSTATS ? 1 : 0);
            // The epilogue's global inputs (shortcut / previous output / BN-backward z) are fetched in BATCHES of KB steps: all
            // loads of a batch are issued together and nothing is issued until the whole batch has been consumed.  (Loads that
            // are issued between the issue and the use of an older set share the warp's six scoreboard slots with it, so the
            // wait for the old set also waits for the new one: register double-buffering gained nothing, ncu showed the warps
            // in long-scoreboard stalls half of the time.)  The first batch of a tile is issued before the tile's accumulator
            // is awaited -- at the end of the previous tile -- so per tile only the later batch boundaries expose a latency.
            constexpr int KB = NL == 1 ? T3_KB1 : 1;
            constexpr int NBATCH = NSTEP / KB;
            static_assert(NSTEP % KB == 0, "steps per tile are a multiple of the prefetch batch");
            float rr[R ? KB * 16 : 1], oo[A ? KB * 16 : 1], zz[BWDSTATS ? KB * 16 : 1];
            // tile -> first pixel and number of pixels that exist (KS == 1: the last tile of the linear view may be ragged;
            // P % 16 == 0, so a 16-pixel step is either complete or absent)
            auto tile_setup = [&](int tile, unsigned& pix0, int& npx) {
                const int n_img = tile / tiles_hw;
                const int trem = tile - n_img * tiles_hw;
                const int th0 = (trem / tiles_w) * TH, tw0 = (trem % tiles_w) * TW;
                pix0 = (unsigned)((n_img * a.H + th0) * a.W + tw0);
                npx = Cfg::NPX;
                if (KS == 1) {
                    const long long left = a.P - (long long)tile * Cfg::NPX;
                    if (left < Cfg::NPX) npx = (int)left;
                }
            };
            // step -> (channel half mh, pixel sub-tile s, 16-pixel group j): this warp takes every other group
            auto step_base = [&](int st, unsigned pix0, int& mh, unsigned& col, unsigned& off, int& lp0) {
                const int per_mh = NSUBS * 4;                  // steps of this warp per channel half
                mh = st / per_mh;
                const int rem = st - mh * per_mh;
                const int s = rem >> 2, j = ((rem & 3) << 1) + hsel;      // 16-pixel group j of sub-tile s: MMA rows 16j..16j+15
                col = (unsigned)(((mh * NACC) * NSUBS + s) * 128 + j * 16);
                // MMA row r of sub-tile s -> tile pixel (KS==1: row s*16 + r/8, column r%8; KS==3: row r/8, column s*8 + r%8)
                const int trow = (KS == 1 ? s * 16 : 0) + j * 2, tcol = (KS == 1 ? 0 : s * 8);
                lp0 = trow * 8;                                // KS == 1: linear index of the group's first pixel inside the tile
                off = (pix0 + (unsigned)trow * (unsigned)a.W + (unsigned)tcol) * (unsigned)BN + (unsigned)(mh * 128 + lq * 32 + lane);
            };
            auto issue_batch = [&](int bi, unsigned pix0, int npx) {
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    int mh, lp0; unsigned col, off;
                    step_base(bi * KB + k, pix0, mh, col, off, lp0);
                    if (KS == 1 && lp0 >= npx) continue;
                    const float* pr = resp + off;
                    const float* po = yp + off;
                    const float* pz = bzp + off;
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        if (R) rr[R ? k * 16 + q : 0] = __ldg(pr + T3_QOFF(q));
                        if (A) oo[A ? k * 16 + q : 0] = po[T3_QOFF(q)];
                        if (BWDSTATS) zz[BWDSTATS ? k * 16 + q : 0] = __ldg(pz + T3_QOFF(q));
                    }
                }
            };
            unsigned pix0 = 0;
            int npx = 0;
            if (blockIdx.x < ntiles) {
                tile_setup(blockIdx.x, pix0, npx);
                if (NL > 0) issue_batch(0, pix0, npx);
            }
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                const int b = i & 1;
                T3_WAIT(bar_accf + 8 * b, (unsigned)(i >> 1) & 1u, w0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // TMEM loads run ONE STEP AHEAD of their use: next to the tensor core's accumulator traffic of the following
                // tile's main loop a tcgen05.ld takes ~1000 cycles (ncu: half of the epilogue's samples sat on its scoreboard)
                uint32_t r[16], r2[16];              // r2: second accumulator (NACC == 2 only)
                auto tmem_issue = [&](int st) {
                    int mh, lp0; unsigned col, off;
                    step_base(st, pix0, mh, col, off, lp0);
                    if (KS == 1 && lp0 >= npx) return;
                    const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * SETCOLS) + col;
                    tmem_ld16_nowait(taddr, r);
                    if (NACC > 1) tmem_ld16_nowait(taddr + (uint32_t)((NACC - 1) * NSUBS * 128), r2);
                };
                tmem_issue(0);
#pragma unroll 1
                for (int bi = 0; bi < NBATCH; ++bi) {
#pragma unroll
                    for (int k = 0; k < KB; ++k) {
                        const int st = bi * KB + k;
                        int mh, lp0; unsigned col, off;
                        step_base(st, pix0, mh, col, off, lp0);
                        const bool present = KS != 1 || lp0 < npx;      // warp-uniform
                        const long long te0 = dbg_on ? clock64() : 0;
                        tmem_ld_wait();
                        float acc[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q] = __uint_as_float(r[q]);
                        if (NACC > 1) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) acc[q] += __uint_as_float(r2[q]);
                        }
                        if (st + 1 < NSTEP) {
                            tmem_issue(st + 1);
                        } else {                     // accumulator set fully read: hand it back to the MMA warp
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_acce + 8 * b);
                        }
                        if (dbg_on) w1 += clock64() - te0;
                        if (present) {
                            const float bias = NM > 1 ? (mh ? bv[NM - 1] : bv[0]) : bv[0];
                            const float rsc = NM > 1 ? (mh ? rs[NM - 1] : rs[0]) : rs[0], rsh = NM > 1 ? (mh ? rt[NM - 1] : rt[0]) : rt[0];
                            const float s_c = NM > 1 ? (mh ? bsc[NM - 1] : bsc[0]) : bsc[0], s_h = NM > 1 ? (mh ? bsh[NM - 1] : bsh[0]) : bsh[0];
                            const float s_m = NM > 1 ? (mh ? bmu[NM - 1] : bmu[0]) : bmu[0], s_i = NM > 1 ? (mh ? biv[NM - 1] : biv[0]) : biv[0];
                            float* py = yp + off;
                            float s1 = 0.f, s2 = 0.f;
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                float v = acc[q] + bias;
                                if (R) v += fmaxf(fmaf(rr[R ? k * 16 + q : 0], rsc, rsh), r_clamp);
                                if (A) v += oo[A ? k * 16 + q : 0];
                                py[T3_QOFF(q)] = v;
                                if (BWDSTATS) {          // sum g, sum g * xhat
                                    const float z = zz[BWDSTATS ? k * 16 + q : 0];
                                    const float gv = fmaf(z, s_c, s_h) <= b_thr ? 0.f : v;
                                    s1 += gv;
                                    s2 = fmaf(gv, (z - s_m) * s_i, s2);
                                } else {                 // sum y, sum y^2
                                    s1 += v;
                                    s2 = fmaf(v, v, s2);
                                }
                            }
                            if (do_stats) {              // fp32 partial sums over 16 pixels, fp64 from here on
                                if (NM > 1 && mh) { d1[NM - 1] += (double)s1; d2[NM - 1] += (double)s2; }
                                else { d1[0] += (double)s1; d2[0] += (double)s2; }
                            }
                        }
                    }
                    // the whole batch is consumed: fetch the next one (of this tile, or the first of this CTA's next tile)
                    if (NL > 0) {
                        if (bi + 1 < NBATCH) {
                            issue_batch(bi + 1, pix0, npx);
                        } else if (tile + (int)gridDim.x < ntiles) {
                            tile_setup(tile + gridDim.x, pix0, npx);
                            issue_batch(0, pix0, npx);
                        }
                    } else if (bi + 1 == NBATCH && tile + (int)gridDim.x < ntiles) {
                        tile_setup(tile + gridDim.x, pix0, npx);
                    }
                }
            }
        };
#undef T3_QOFF
        if (has_res) {
            if (has_acc) tile_loop(std::true_type{}, std::true_type{}); else tile_loop(std::true_type{}, std::false_type{});
        } else {
            if (has_acc) tile_loop(std::false_type{}, std::true_type{}); else tile_loop(std::false_type{}, std::false_type{});
        }
        if (do_stats) {
#pragma unroll
            for (int mh = 0; mh < NM; ++mh) {
                const int ch = mh * 128 + lq * 32 + lane;
                atomicAdd(a.stat_sum + ch, d1[mh]);
                atomicAdd(a.stat_sq + ch, d2[mh]);
            }
        }
        if (we == 0 && lane == 0) { T3_DBG(10, clock64() - t_begin); T3_DBG(11, w0); T3_DBG(16, w1); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    pdl_launch_dependents();      // every tile of this CTA is stored: the next kernel of the chain may start its prologue
    if (warp == T3_WAUX)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    // ---- the CTA that arrives last turns the complete sums into the per-channel BatchNorm vectors ----
    if (a.stat_sum != nullptr) {
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

template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY, bool XBF = false>
static int launch_tc3(TcArgs ta, cudaStream_t st) {
    using Cfg = T3Cfg<BN, SPLIT, KS, BNAPPLY>;
    static bool configured = false;
    constexpr int smem = Cfg::SMEM;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<BN, SPLIT, KS, BWDSTATS, BNAPPLY, XBF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_tc_nhwc (persistent tile kernel): cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    if (KS == 1) {          // a 1x1 convolution has no neighbourhood: any run of consecutive pixels is a tile
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
    const long long tiles = (long long)ta.c.N * ((ta.c.H + Cfg::TH - 1) / Cfg::TH) * (ta.c.W / Cfg::TW);
    const long long rounds = (tiles + kNumSMs - 1) / kNumSMs;
    const unsigned grid = (unsigned)((tiles + rounds - 1) / rounds);
    cudaError_t le = launch_pdl(conv_tc3_kernel<BN, SPLIT, KS, BWDSTATS, BNAPPLY, XBF>, dim3(grid), dim3(T3_THREADS), (size_t)smem, st, ta, mx, mz);
    if (le != cudaSuccess) {
        set_error("hgk_conv_tc_nhwc (persistent tile kernel): launch: %s", cudaGetErrorString(le));
        return HGK_ECUDA;
    }
    return HGK_OK;
}

template <int BN, int KS>
static int launch_tc3_bn(const TcArgs& ta, bool split, bool bwdstats, cudaStream_t st) {
    if (ta.c.ap.z != nullptr)        // data gradient with the BatchNorm-backward apply evaluated on load
        return bwdstats ? launch_tc3<BN, false, KS, true, true>(ta, st) : launch_tc3<BN, false, KS, false, true>(ta, st);
    if (bwdstats) return launch_tc3<BN, false, KS, true, false>(ta, st);
    if (split) {
        if (ta.lo_bf16) {       // TF32 + 2xBF16 products (w_lo in pack mode 2): 1x1 forward
            if (KS == 1) return launch_tc3<BN, true, 1, false, false, KS == 1>(ta, st);
            set_error("hgk_conv_tc_x2_nhwc: no TF32 + 2xBF16 instantiation of the persistent kernel for k=%d", KS);
            return HGK_EINVAL;
        }
        return launch_tc3<BN, true, KS, false, false>(ta, st);
    }
    return launch_tc3<BN, false, KS, false, false>(ta, st);
}

// true when the persistent tile kernel covers this problem
bool conv_tc3_eligible(const TcArgs& ta) {
    const ConvArgs& c = ta.c;
    if (c.Cout != 128 && c.Cout != 256) return false;
    if (c.ksize == 3) {
        if (c.H % 16 || c.W % 16) return false;
        if (c.Cout > 128) return false;                                   // 3x3 with 256 outputs: no instantiation
    } else {
        if (c.P % 16) return false;                                       // a 16-pixel epilogue step is complete or absent
    }
    if (c.Cin > 256) return false;                                        // per-channel vectors staged in shared memory
    const long long cmax = c.Cin > c.Cout ? c.Cin : c.Cout;
    if ((c.P + 256) * cmax >= (1LL << 32)) return false;                  // 32-bit element offsets
    if (((uintptr_t)c.x.z & 15) || (c.ap.z != nullptr && ((uintptr_t)c.ap.z & 15))) return false;   // tensor-map base alignment
    return true;
}

int conv_tc3_launch(const TcArgs& ta, bool split, bool bwdstats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (ta.c.ksize == 3) return launch_tc3_bn<128, 3>(ta, split, bwdstats, st);
    if (ta.c.Cout == 128) return launch_tc3_bn<128, 1>(ta, split, bwdstats, st);
    return launch_tc3_bn<256, 1>(ta, split, bwdstats, st);
}

}  // namespace hgk
