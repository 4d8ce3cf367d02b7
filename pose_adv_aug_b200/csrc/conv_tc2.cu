// tcgen05 implicit-GEMM convolution over 16x16-pixel IMAGE TILES (M = 256 output pixels per CTA) -- the
// kernel for every layer whose H and W are multiples of 16 (16x16 .. 128x128).  Same contract as
// conv_tc_kernel (conv_tc.cu): BN+ReLU on load, 3xTF32 (forward) or plain TF32 (data gradient), bias /
// shortcut / accumulate / BN-statistics (or fused BN-backward reduction) epilogue.
//
// Why tiles: the 128-linear-pixel kernel re-reads its activation tile once per 3x3 tap and re-streams the
// whole packed weight matrix per 128 pixels; at 64x64 both flows go through L2 at ~10x the compulsory HBM
// bytes and L2 -> SM bandwidth (about the HBM bandwidth on B200) becomes the bound (profiles/r1b_tc_summary.md).
// Here
//   * the activation HALO tile (18x18 pixels x 16 channels for 3x3; 16x16 for 1x1) is loaded, transformed
//     (BN+ReLU, TF32 hi/lo split) and stored to shared memory ONCE per 16-channel chunk in the K-major
//     64-byte-swizzle UMMA layout: halo pixel (hh, ww) is one 64-byte row at (hh*HW + ww)*64, its four 16-byte
//     channel quads XOR-swizzled with address bits [7,9).  An 8-pixel run of one image row is one 8-row group,
//     the next image row is SBO = HW*64 bytes further.  The nine taps are nine MMAs on the SAME tile with the
//     descriptor start address shifted by (dh*HW + dw)*64 bytes -- no im2col, no re-load.  (The swizzle is a
//     function of the absolute shared-memory address, so row-shifted starts stay consistent; the first version
//     used the no-swizzle layout with 16-byte pixel rows, whose shifted core matrices straddle 128-byte lines:
//     ncu showed every 3x3 MMA occupying the tensor pipe for 128 instead of 64 cycles.)
//   * each weight stage (one tap x 16 channels x Cout, hi and lo) arrives by cp.async.bulk (TMA bulk copy) and
//     is used by both 128-pixel halves of the tile (columns 0-7 and 8-15), halving the weight stream per pixel;
//   * roles: warps 0-7 load/transform/store the halo tiles (two chunks of register prefetch), warp 8 issues
//     tcgen05.mma.kind::tf32 into TMEM and commits stages back, warp 9 streams the weights; all 8 producer
//     warps then run the epilogue (tcgen05.ld -> smem tile -> coalesced stores + fp64 channel statistics).
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"
#include "bn_fin.cuh"

// TF32 + 2xBF16 forward products (template parameter XBF; 3x3 forward, the tensor-bound launches of the step).  The fp32-class
// product x*w ~= xh*wh + xl*wh + xh*wl keeps its main term in TF32 (kind::tf32, K = 8) and evaluates the two cross terms --
// each 2^-11 of the main term -- with bf16 operands (kind::f16, K = 16): bf16(xl)*bf16(wh) + bf16(xh)*bf16(wl), error
// ~3 * 2^-20 per product.  Four MMAs per 16-channel chunk and tap instead of six, same shared-memory bytes: the fp32 lo plane
// of the activation stage is replaced by two bf16 tiles (xl, xh) in the K-major NO-swizzle core-matrix layout
// [k/8 (2 planes)][halo pixel][8 x bf16 = 16 B] (pixel rows 16 bytes apart, so a 3x3 tap is again a shifted start address:
// (dh*HWD + dw) * 16 bytes; 8-pixel row group -> next image row = SBO = HWD * 16), the fp32 lo half of the weight stage by
// [wh_bf16 | wl_bf16] in [k/8][n][8 x bf16] (pack mode 2, conv_tc.cu).
namespace hgk {

// UMMA shared-memory descriptor, K-major, SWIZZLE_64B (layout type 4 in bits 61-63): rows of 64 bytes, 8-row groups
// SBO bytes apart; LBO is not used by swizzled K-major layouts (canonical value 1)
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t saddr, uint32_t sbo) {
    return umma_desc(saddr, 16u, sbo) | ((uint64_t)4 << 61);
}

constexpr int T2_THREADS = 320;      // 8 producer/epilogue warps + MMA warp + weight-copy warp

// NSUB = 128-pixel halves per tile: 2 (16x16 tiles, each weight stage feeds two MMAs per k-step) for the large
// layers; 1 (16 rows x 8 columns) when the 16x16 grid would leave SMs idle (16x16 and 32x32 layers): twice the
// CTAs, half the serial work per CTA, half the TMEM (so two CTAs fit on an SM)
template <int BN, bool SPLIT, int KS, int NSUB = 2>
struct T2Cfg {
    static constexpr int TH = 16, TW = 8 * NSUB;
    // KS = 4: the space-to-depth form of the 7x7 stride-2 stem convolution (stem.cu) -- 4x4 taps at offsets -2 .. +1
    static constexpr int PAD = KS / 2;                                     // halo rows / columns in FRONT of the tile
    static constexpr int HH = TH + KS - 1, HWD = TW + KS - 1;             // halo tile
    static constexpr int NPIX = HH * HWD;
    static constexpr int TAPS = KS * KS;
    static constexpr int BK = 16, QP = 4;                                  // channels / 16-byte quads per chunk
    static constexpr int NITEM = NPIX * QP;                                // (pixel, quad) items per chunk
    static constexpr int NJ = (NITEM + 255) / 256;                         // items per producer thread
    static constexpr int SBO_A = HWD * 64;                                 // 8-pixel row group -> next image row
    static constexpr int A_HALF = (NPIX * 64 + 1023) / 1024 * 1024;        // hi (or lo) part of a stage
    static constexpr int A_STAGE = (SPLIT ? 2 : 1) * A_HALF;
    static constexpr int B_HALF = BN * BK * 4;
    static constexpr int B_STAGE = (SPLIT ? 2 : 1) * B_HALF;
    // accumulators per 128-pixel half, see conv_tc.cu (fp32 accumulator truncation of the tensor core): 3x3 chains
    // (K = 9 Cin) split into hi*hi / cross-term accumulators; a 1x1 chain is at most 3*256/8 = 96 MMAs long -- the
    // same length conv_tc.cu accepts for its single-accumulator BN = 256 case
    // (the stem, KS = 4, keeps them too: with ONE accumulator for its 96-MMA chain the error was 1.6e-6 of max instead of 4e-7,
    //  and the launch was no faster: 115 vs 117 us)
    static constexpr int NMAIN = (SPLIT && KS > 1 && BN <= 64) ? 2 : 1;
    static constexpr int NACC = (!SPLIT || KS == 1 || BN >= 256) ? 1 : NMAIN + 1;
    static constexpr int SUBCOLS = NACC * BN;
    static constexpr int TMEM_COLS = (NSUB * SUBCOLS <= 64) ? 64 : (NSUB * SUBCOLS <= 128) ? 128 : (NSUB * SUBCOLS <= 256) ? 256 : 512;
    // two CTAs per SM wherever TMEM (<= 256 columns each) allows it: the prologue / epilogue of one CTA then overlaps
    // the main loop of the other (every phase of this kernel is a serial latency chain inside one CTA)
    static constexpr bool OCC2 = TMEM_COLS <= 256;
    // PAIR: the producers fill TWO 16-channel stages per round (one proxy fence, twice the independent work per
    // thread): for the plain-TF32 1x1 kernels (data gradients) the producers' serial chain per chunk (~1300 cycles:
    // transform ~600 + fence / arrive ~300-1000) is what bounds the main loop, not the tensor pipe (184-366 cycles per
    // chunk) -- profiles/r2_tile_kernel_timeline.md.  Needs an even number (>= 4) of activation stages.  The kernel
    // switches it off for the BNAPPLY instantiations (measured slower there: 113 -> 119 us on the 64x64 256->128 data
    // gradient; the doubled register sets spill inside the producer loop).
    static constexpr bool PAIR = KS == 1 && NSUB == 1 && !SPLIT;
    // KS = 4 (stem): a single 16-channel chunk (12 real channels), so one activation stage; its sixteen 8 KB weight stages
    // are latency-bound bulk copies -- eight in flight (three left the tile waiting ~0.25 us per tap: 199 us per launch)
    static constexpr int NSA = KS == 4 ? 1 : KS > 1 ? 2 : (PAIR ? 4 : (OCC2 && SPLIT ? 2 : 3));      // activation stages (one per chunk)
    static constexpr int NSB = KS == 4 ? 8 : KS > 1 ? (OCC2 ? (SPLIT ? 3 : 4) : 8) : NSA;   // weight stages (one per chunk x tap)
    static constexpr int PIPE = NSA * A_STAGE + NSB * B_STAGE;
    static constexpr int CH = BN > 128 ? 128 : BN;                         // epilogue column chunk
    static constexpr int STG_BYTES = TBM * (CH + 4) * 4 + 16384;
    static constexpr int SMEM = (PIPE > STG_BYTES ? PIPE : STG_BYTES) + 1024;
    static_assert(NSUB * SUBCOLS <= 512, "TMEM capacity");
    static_assert(A_HALF % 1024 == 0 && B_HALF % 1024 == 0, "stage alignment (swizzle atoms)");
    static_assert(SMEM <= (OCC2 ? 113 : 227) * 1024, "shared memory");
};

template <int BN, bool SPLIT, int KS, int NSUB>
__host__ __device__ constexpr uint32_t Cfg_NPIX16() { return (uint32_t)T2Cfg<BN, SPLIT, KS, NSUB>::NPIX * 16u; }

template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY, int NSUB, bool XBF = false>
__global__ void __launch_bounds__(T2_THREADS, (T2Cfg<BN, SPLIT, KS, NSUB>::OCC2 ? 2 : 1)) conv_tc2_kernel(const TcArgs args) {
    static_assert(!(BNAPPLY && SPLIT), "the fused BatchNorm-backward apply is a data-gradient (plain TF32) mode");
    static_assert(!XBF || SPLIT, "bf16 cross terms belong to the error-compensated forward");
    // XBF: second half of an activation stage = xl tile (2 planes of NPIX x 16 B) followed by the xh tile
    constexpr uint32_t XPLANE = Cfg_NPIX16<BN, SPLIT, KS, NSUB>();
    constexpr uint32_t IDESC_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    using Cfg = T2Cfg<BN, SPLIT, KS, NSUB>;
    constexpr int NJ = Cfg::NJ, NSA = Cfg::NSA, NSB = Cfg::NSB, NMAIN = Cfg::NMAIN, NACC = Cfg::NACC;
    constexpr int HWD = Cfg::HWD, PAD = Cfg::PAD, TAPS = Cfg::TAPS, SUBCOLS = Cfg::SUBCOLS;
    constexpr uint32_t SBO_A = Cfg::SBO_A, LBO_B = BN * 16, SBO_B = 128;
    constexpr uint32_t A_HALF = Cfg::A_HALF, A_STAGE = Cfg::A_STAGE, B_HALF = Cfg::B_HALF, B_STAGE = Cfg::B_STAGE;
    constexpr uint32_t B_OFF = NSA * A_STAGE;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const ConvArgs& a = args.c;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NSA + 2 * NSB + 1];
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) HGK_STAMP(0);          // developer timeline (tools/dbg_timeline2.py); a predicated-off store otherwise
    const int tiles_w = a.W / Cfg::TW, tiles_hw = (a.H >> 4) * tiles_w;
    const int n_img = blockIdx.x / tiles_hw;
    const int trem = blockIdx.x - n_img * tiles_hw;
    const int th0 = (trem / tiles_w) << 4, tw0 = (trem % tiles_w) * Cfg::TW;
    const int KC = KS == 4 ? 1 : a.Cin >> 4;     // stem: 16-channel pixels; the weights are packed as K = 32 (upper half zero)
    const uint32_t bar_fa = smem_u32(&bars[0]), bar_ea = smem_u32(&bars[NSA]);
    const uint32_t bar_fb = smem_u32(&bars[2 * NSA]), bar_eb = smem_u32(&bars[2 * NSA + NSB]);
    const uint32_t bar_done = smem_u32(&bars[2 * NSA + 2 * NSB]);

    if (tid == 0) {
        for (int s = 0; s < NSA; ++s) {
            mbar_init(bar_fa + 8 * s, 256);
            mbar_init(bar_ea + 8 * s, 1);
        }
        for (int s = 0; s < NSB; ++s) {
            mbar_init(bar_fb + 8 * s, 1);
            mbar_init(bar_eb + 8 * s, 1);
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
    // The CTA-wide rendezvous that publishes the barriers and the TMEM base is a NAMED barrier executed inside each role
    // branch (bar.sync 1, 320): the producers issue their first global loads BEFORE it, so the ~0.6 us of barrier
    // initialisation / TMEM allocation overlaps the first memory round trip (profiles/r2_tile_kernel_timeline.md).
#define T2_RENDEZVOUS()                                                           \
    do {                                                                          \
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");          \
        asm volatile("barrier.sync 1, 320;" ::: "memory");   /* non-aligned form: reached from different code */  \
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");           \
    } while (0)
    static_assert(T2_THREADS == 320, "T2_RENDEZVOUS counts 320 threads");
    uint32_t tmem = 0;
    // programmatic dependent launch (common.cuh): barrier initialisation and the TMEM allocation above overlap the previous
    // kernel's tail; no global access before this point
    pdl_wait();

    if (warp < 8) {
        // ===== producers: halo tile of one 16-channel chunk -> registers -> BN+ReLU, hi/lo -> shared memory =====
        // item idx = tid + 256*j: halo pixel idx>>2, channel quad idx&3 (= tid&3 for every j)
        const int quad = tid & 3;
        unsigned a_off[NJ], s_off[NJ];
        unsigned vmask = 0, smask = 0;          // pixel inside the image / item inside the tile
        unsigned imask = 0;                     // BNAPPLY: pixel owned by this tile (its dz is written back)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int idx = tid + 256 * j;
            const int hp = idx >> 2;
            const int hh = hp / HWD, ww = hp - hh * HWD;
            const int h = th0 - PAD + hh, w = tw0 - PAD + ww;
            const bool in_tile = idx < Cfg::NITEM;
            const bool in_img = in_tile && (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W;
            a_off[j] = in_img ? (unsigned)((((long long)n_img * a.H + h) * a.W + w) * a.Cin) + quad * 4 : 0u;
            // 64-byte swizzle: 16-byte chunk index ^= address bits [7,9) = (pixel >> 1) & 3 (stage bases are 1024-aligned)
            s_off[j] = (unsigned)hp * 64u + (((unsigned)quad ^ (((unsigned)hp >> 1) & 3u)) << 4);
            if (in_img) vmask |= 1u << j;
            if (in_tile) smask |= 1u << j;
            if (BNAPPLY && in_tile && hh >= PAD && hh < PAD + 16 && ww >= PAD && ww < PAD + Cfg::TW) imask |= 1u << j;
        }
        const float* xz = a.x.z;
        const bool has_aff = a.x.scale != nullptr;
        const float x_clamp = a.x.relu ? 0.f : -INFINITY;
        // register prefetch depth: two chunks when the CTA has the SM to itself, one when a second CTA covers the latency
        constexpr bool PAIR = Cfg::PAIR && !BNAPPLY;
        constexpr int NSET = (PAIR || !Cfg::OCC2) ? 2 : 1;
        float4 a_reg[NSET][NJ];
        float4 z_reg[BNAPPLY ? NSET : 1][BNAPPLY ? NJ : 1];   // BNAPPLY: the pre-BN output z next to its gradient
        const BnApply& ap = a.ap;
        auto load_a = [&](int set, int kc) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((vmask >> j) & 1u) v = ldg4(xz + (a_off[j] + (unsigned)kc * 16u));
                a_reg[set][j] = v;
                if (BNAPPLY) {
                    float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((vmask >> j) & 1u) zz = ldg4(ap.z + (a_off[j] + (unsigned)kc * 16u));
                    z_reg[BNAPPLY ? set : 0][BNAPPLY ? j : 0] = zz;
                }
            }
        };
        auto store_a = [&](int s, int set, int kc) {
            float4 sc, sh;
            load_affine4(a.x.scale, a.x.shift, kc * 16 + quad * 4, sc, sh);
            float4 bs, bt, bmu, bA, bB, bC;
            if (BNAPPLY) {
                const int c = kc * 16 + quad * 4;
                bs = ldg4(ap.scale + c); bt = ldg4(ap.shift + c); bmu = ldg4(ap.mean + c);
                bA = ldg4(ap.cA + c); bB = ldg4(ap.cB + c); bC = ldg4(ap.cC + c);
            }
            uint8_t* base = sgen + s * A_STAGE;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (!((smask >> j) & 1u)) continue;
                float4 v = a_reg[set][j];
                // zero padding is a zero of the ACTIVATED tensor: transform only pixels inside the image
                if (has_aff && ((vmask >> j) & 1u)) v = actc4(v, sc, sh, x_clamp);
                if (BNAPPLY && ((vmask >> j) & 1u)) {
                    // dz = cA * ((g * relu-mask - cC) - (z - mean) * cB)   (bn_bwd_apply_kernel, bn.cu)
                    const float4 z = z_reg[BNAPPLY ? set : 0][BNAPPLY ? j : 0];
                    const float gx = (ap.relu && fmaf(z.x, bs.x, bt.x) <= 0.f) ? 0.f : v.x;
                    const float gy = (ap.relu && fmaf(z.y, bs.y, bt.y) <= 0.f) ? 0.f : v.y;
                    const float gz = (ap.relu && fmaf(z.z, bs.z, bt.z) <= 0.f) ? 0.f : v.z;
                    const float gw = (ap.relu && fmaf(z.w, bs.w, bt.w) <= 0.f) ? 0.f : v.w;
                    v.x = bA.x * ((gx - bC.x) - (z.x - bmu.x) * bB.x);
                    v.y = bA.y * ((gy - bC.y) - (z.y - bmu.y) * bB.y);
                    v.z = bA.z * ((gz - bC.z) - (z.z - bmu.z) * bB.z);
                    v.w = bA.w * ((gw - bC.w) - (z.w - bmu.w) * bB.w);
                    if ((imask >> j) & 1u) st4(ap.dz + (a_off[j] + (unsigned)kc * 16u), v);
                }
                const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                *reinterpret_cast<float4*>(base + s_off[j]) = hi;
                if (SPLIT && !XBF) {
                    const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    *reinterpret_cast<float4*>(base + A_HALF + s_off[j]) = lo;
                }
                if (XBF) {
                    // bf16 tiles: plane = channel quad / 2 (8-channel K group), row = halo pixel (16 B), 8-byte half = quad & 1
                    const uint32_t boff = (uint32_t)(quad >> 1) * XPLANE + (uint32_t)((tid + 256 * j) >> 2) * 16u + (uint32_t)(quad & 1) * 8u;
                    uint2 xl, xh;
                    xl.x = pack_bf16x2(v.x - hi.x, v.y - hi.y); xl.y = pack_bf16x2(v.z - hi.z, v.w - hi.w);
                    xh.x = pack_bf16x2(hi.x, hi.y); xh.y = pack_bf16x2(hi.z, hi.w);
                    *reinterpret_cast<uint2*>(base + A_HALF + boff) = xl;
                    *reinterpret_cast<uint2*>(base + A_HALF + 2u * XPLANE + boff) = xh;
                }
            }
        };
        load_a(0, 0);
        if (NSET > 1 && KC > 1) load_a(NSET - 1, 1);
        if (tid == 0) HGK_STAMP(2);
        T2_RENDEZVOUS();
        if (tid == 0) HGK_STAMP(1);
        int sa = 0;
        unsigned ea_par = 1;                 // parity of the previous use of the stage (toggles when sa wraps)
        if (PAIR) {
            // two chunks per round (KC is even: Cin % 32 == 0); stages sa, sa+1 wrap together (NSA even)
            for (int kc = 0; kc < KC; kc += 2) {
                if (kc >= NSA) {
                    mbar_wait(bar_ea + 8 * sa, ea_par);
                    mbar_wait(bar_ea + 8 * (sa + 1), ea_par);
                }
                if (tid == 0) HGK_TRACE(0, kc);
                store_a(sa, 0, kc);
                store_a(sa + 1, NSET - 1, kc + 1);
                if (tid == 0) HGK_TRACE(6, kc);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                mbar_arrive(bar_fa + 8 * sa);
                mbar_arrive(bar_fa + 8 * (sa + 1));
                if (tid == 0) { HGK_TRACE(1, kc); if (kc == 0) HGK_STAMP(3); }
                if (kc + 2 < KC) { load_a(0, kc + 2); load_a(NSET - 1, kc + 3); }
                sa += 2;
                if (sa == NSA) { sa = 0; ea_par ^= 1u; }
            }
        } else {
            for (int kc = 0; kc < KC; ++kc) {
                if (kc >= NSA) mbar_wait(bar_ea + 8 * sa, ea_par);            // stage drained by the tensor core
                if (tid == 0) HGK_TRACE(0, kc);
                if (NSET > 1 && (kc & 1)) store_a(sa, NSET - 1, kc); else store_a(sa, 0, kc);
                if (tid == 0) HGK_TRACE(6, kc);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                mbar_arrive(bar_fa + 8 * sa);
                if (tid == 0) { HGK_TRACE(1, kc); if (kc == 0) HGK_STAMP(3); }
                if (kc + NSET < KC) { if (NSET > 1 && (kc & 1)) load_a(NSET - 1, kc + NSET); else load_a(0, kc + NSET); }
                if (++sa == NSA) { sa = 0; ea_par ^= 1u; }
            }
        }
        if (tid == 0) HGK_STAMP(4);
        mbar_wait(bar_done, 0);              // every MMA retired: accumulators complete, shared memory reusable
        pdl_launch_dependents();             // main loop done: the next kernel of the chain may start its prologue
        if (tid == 0) HGK_STAMP(5);
    } else if (warp == 8) {
        // ===== MMA issuer =====
        T2_RENDEZVOUS();
        tmem = tmem_slot;
        if (lane == 0) {
            int sa = 0, sb = 0, it = 0;
            unsigned fa_par = 0, fb_par = 0;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(bar_fa + 8 * sa, fa_par);
                if (kc == 0) HGK_STAMP(8);
                HGK_TRACE(2, kc);
                const uint32_t a_stage = sbase + sa * A_STAGE;
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap, ++it) {
                    mbar_wait(bar_fb + 8 * sb, fb_par);
                    if (it == 0) HGK_STAMP(9);
                    if (tap == 0) HGK_TRACE(3, kc);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_tap = a_stage + (KS > 1 ? (uint32_t)((tap / KS) * HWD + (tap % KS)) * 64u : 0u);
                    const uint32_t b_hi = sbase + B_OFF + sb * B_STAGE;
#pragma unroll
                    for (int sub = 0; sub < NSUB; ++sub) {
                        const uint32_t t_sub = tmem + sub * SUBCOLS;
                        if (XBF) {
                            // cross terms, one K = 16 bf16 MMA each: xl * wh_bf16 and xh_bf16 * wl_bf16
                            const uint32_t x_tap = a_stage + A_HALF + (KS > 1 ? (uint32_t)((tap / KS) * HWD + (tap % KS)) * 16u : 0u) + sub * 128;
                            const uint64_t dxl = umma_desc(x_tap, XPLANE, (uint32_t)HWD * 16u);
                            const uint64_t dxh = umma_desc(x_tap + 2u * XPLANE, XPLANE, (uint32_t)HWD * 16u);
                            const uint64_t dwh = umma_desc(b_hi + B_HALF, LBO_B, SBO_B);
                            const uint64_t dwl = umma_desc(b_hi + B_HALF + BN * 32, LBO_B, SBO_B);
                            const uint32_t t_x = NACC == 1 ? t_sub : t_sub + NMAIN * BN;
                            umma_bf16(t_x, dxl, dwh, IDESC_BF, (NACC == 1 || it > 0) ? (it > 0 ? 1u : 0u) : 0u);
                            umma_bf16(t_x, dxh, dwl, IDESC_BF, 1u);
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint64_t da = umma_desc_k64(a_tap + sub * 512 + k * 32, SBO_A);
                            const uint64_t db = umma_desc(b_hi + k * 2 * LBO_B, LBO_B, SBO_B);
                            if (XBF) {
                                const uint32_t t_main = NACC == 1 ? t_sub : t_sub + (NMAIN > 1 ? (it & 1) * BN : 0);
                                // (NACC == 1: the bf16 MMA above already initialised the accumulator of this stage)
                                umma_tf32(t_main, da, db, IDESC, NACC == 1 ? 1u : ((it >= NMAIN || k > 0) ? 1u : 0u));
                            } else if (SPLIT) {
                                const uint64_t dal = umma_desc_k64(a_tap + A_HALF + sub * 512 + k * 32, SBO_A);
                                const uint64_t dbl = umma_desc(b_hi + B_HALF + k * 2 * LBO_B, LBO_B, SBO_B);
                                if (NACC == 1) {
                                    umma_tf32(t_sub, dal, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                                    umma_tf32(t_sub, da, dbl, IDESC, 1u);
                                    umma_tf32(t_sub, da, db, IDESC, 1u);
                                } else {
                                    const uint32_t t_small = t_sub + NMAIN * BN;
                                    const uint32_t t_main = t_sub + (NMAIN > 1 ? (it & 1) * BN : 0);
                                    umma_tf32(t_small, dal, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                                    umma_tf32(t_small, da, dbl, IDESC, 1u);
                                    umma_tf32(t_main, da, db, IDESC, (it >= NMAIN || k > 0) ? 1u : 0u);
                                }
                            } else {
                                umma_tf32(t_sub, da, db, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(bar_eb + 8 * sb);                          // weight stage free
                    if (++sb == NSB) { sb = 0; fb_par ^= 1u; }
                }
                umma_commit(bar_ea + 8 * sa);                              // activation stage free
                HGK_TRACE(4, kc);
                if (++sa == NSA) { sa = 0; fa_par ^= 1u; }
            }
            umma_commit(bar_done);
        }
        __syncwarp();
    } else {
        // ===== weight stream: one cp.async.bulk per (chunk, tap) stage; packed blocks are [tap][Cin/32][8 quads][BN][4] =====
        T2_RENDEZVOUS();
        if (lane == 0) {
            const int KC32 = KS == 4 ? 1 : a.Cin >> 5;
            int sb = 0, it = 0;
            unsigned eb_par = 1;
            for (int kc = 0; kc < KC; ++kc) {
                for (int tap = 0; tap < TAPS; ++tap, ++it) {
                    if (it >= NSB) mbar_wait(bar_eb + 8 * sb, eb_par);
                    if (tap == 0) HGK_TRACE(5, kc);
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
        __syncwarp();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    tmem = tmem_slot;

    // ---- epilogue: per column chunk of CH and per 128-pixel half: TMEM -> registers -> staging tile -> coalesced
    //      bias / shortcut / accumulate / store + BN statistics (row r of half `sub` = pixel (th0 + r/8, tw0 + 8 sub + r%8)) ----
    constexpr int CH = Cfg::CH, SROW = CH + 4;
    constexpr int CG = CH / 4;           // float4 column groups of a chunk
    constexpr int RL = TNT / CG;         // row lanes (8 for CH=128, 16 for CH=64)
    constexpr int ROWS = TBM / RL;       // rows per thread (16 / 8)
    float* stg = reinterpret_cast<float*>(sgen);
    double* red = reinterpret_cast<double*>(sgen + TBM * SROW * 4);     // [RL][CH][2], behind the staging tile
    const bool epi = tid < TNT;
    const int cg = (epi ? tid : 0) % CG, r0 = (epi ? tid : 0) / CG;
    const bool do_stats = a.stat_sum != nullptr;
    const bool has_res = a.res.z != nullptr, res_aff = a.res.scale != nullptr;
    // element offsets fit 32 bits (conv_tc2_eligible: P * max(Cin, Cout) < 2^32)
    const unsigned pix0 = (unsigned)((n_img * a.H + th0) * a.W + tw0);
    const unsigned uW = (unsigned)a.W, uCout = (unsigned)a.Cout;
#pragma unroll 1
    for (int ch = 0; ch < BN / CH; ++ch) {
        const int n = ch * CH + cg * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias != nullptr) bv = ldg4(a.bias + n);
        float4 rs, rt;
        load_affine4(a.res.scale, a.res.shift, n, rs, rt);
        float4 bsc = make_float4(1.f, 1.f, 1.f, 1.f), bsh = make_float4(0.f, 0.f, 0.f, 0.f), bmu = bsh, biv = bsc;
        if (BWDSTATS) { bsc = ldg4(a.bscale + n); bsh = ldg4(a.bshift + n); bmu = ldg4(a.bmean + n); biv = ldg4(a.binvstd + n); }
        double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
#pragma unroll 1
        for (int sub = 0; sub < NSUB; ++sub) {
            if (tid == 0) HGK_TRACE(8, ch * NSUB + sub);
            if (warp < 8) {
                const int lq = warp & 3;
                const int row = lq * 32 + lane;
                const int cbeg = (warp >> 2) * (CH / 2);
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + CH / 2; c0 += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(sub * SUBCOLS + ch * CH + c0);
                    tmem_ld32(taddr, r);
                    float acc[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(r[q]);
#pragma unroll
                    for (int e = 1; e < NACC; ++e) {
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
            if (ch == BN / CH - 1 && sub == NSUB - 1 && warp == 0)
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
            if (tid == 0) HGK_TRACE(9, ch * NSUB + sub);
            float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
            // rows are processed RB at a time with all global loads (shortcut / previous output / BN input) issued first
            constexpr int RB = Cfg::OCC2 ? 2 : 4;
            // PLAIN (no shortcut, no accumulation -- conv1 / conv2 of every residual block): the row loop is issue-bound
            // (~40 instructions per row), so the unused shortcut / accumulate registers and adds are compiled out
            auto row_loop = [&](auto plain_tag) {
                constexpr bool PLAIN = decltype(plain_tag)::value;
#pragma unroll 1
                for (int g = 0; epi && g < ROWS; g += RB) {
                    float4 rr[RB], oo[RB], zz[RB];
                    unsigned pp[RB];
#pragma unroll
                    for (int i = 0; i < RB; ++i) {
                        const int r = r0 + (g + i) * RL;
                        pp[i] = (pix0 + (unsigned)(r >> 3) * uW + (unsigned)(sub * 8 + (r & 7))) * uCout + (unsigned)n;
                        if (!PLAIN) {
                            rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            oo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (has_res) rr[i] = ldg4(a.res.z + pp[i]);
                            if (a.accumulate) oo[i] = ld4(a.y + pp[i]);
                        }
                        if (BWDSTATS) zz[i] = ldg4(a.bz + pp[i]);
                    }
#pragma unroll
                    for (int i = 0; i < RB; ++i) {
                        const int r = r0 + (g + i) * RL;
                        float4 v = ld4(stg + r * SROW + cg * 4);
                        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        if (!PLAIN) {
                            if (has_res) {
                                float4 q = rr[i];
                                if (res_aff) q = act4(q, rs, rt, a.res.relu);
                                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                            }
                            v.x += oo[i].x; v.y += oo[i].y; v.z += oo[i].z; v.w += oo[i].w;
                        }
                        st4(a.y + pp[i], v);
                        if (do_stats) {
                            if (BWDSTATS) {
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
            };
            if (!has_res && !a.accumulate) row_loop(std::true_type{}); else row_loop(std::false_type{});
            if (do_stats) {      // fp32 partial sums over this thread's ROWS (16 / 8) values, fp64 from here on
#pragma unroll
                for (int j = 0; j < 4; ++j) { d1[j] += (double)s1[j]; d2[j] += (double)s2[j]; }
            }
            if (tid == 0) HGK_TRACE(10, ch * NSUB + sub);
            if (!(ch == BN / CH - 1 && sub == NSUB - 1)) __syncthreads();   // staging tile is rewritten by the next half / chunk
        }
        if (do_stats) {
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
                atomicAdd(a.stat_sum + ch * CH + tid, x1);
                atomicAdd(a.stat_sq + ch * CH + tid, x2);
            }
            if (ch + 1 < BN / CH) __syncthreads();                       // `red` is rewritten by the next chunk
            if (tid == 0) HGK_TRACE(11, ch);
        }
    }
    if (tid == 0) HGK_STAMP(6);
    // fused BatchNorm finaliser: the CTA that arrives last turns the complete sums into per-channel vectors
    if (do_stats) {
        if (!BWDSTATS && a.ffin.ticket != nullptr) {
            if (last_cta_arrives(a.ffin.ticket, gridDim.x)) bn_fwd_finalize_cta(a.ffin, a.stat_sum, a.stat_sq, (double)a.P, a.Cout);
        } else if (BWDSTATS && a.bfin.ticket != nullptr) {
            if (last_cta_arrives(a.bfin.ticket, gridDim.x)) bn_bwd_finalize_cta(a.bfin, a.stat_sum, a.stat_sq, (double)a.P, a.Cout);
        }
    }
}

template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY, int NSUB, bool XBF = false>
static int launch_tc2_sub(const TcArgs& ta, cudaStream_t st) {
    static bool configured = false;
    constexpr int smem = T2Cfg<BN, SPLIT, KS, NSUB>::SMEM;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, SPLIT, KS, BWDSTATS, BNAPPLY, NSUB, XBF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("hgk_conv_tc_nhwc (tile kernel): cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
            return HGK_ECUDA;
        }
        configured = true;
    }
    const unsigned grid = (unsigned)(ta.c.N * (ta.c.H >> 4) * (ta.c.W / (8 * NSUB)));
    cudaError_t le = launch_pdl(conv_tc2_kernel<BN, SPLIT, KS, BWDSTATS, BNAPPLY, NSUB, XBF>, dim3(grid), dim3(T2_THREADS), (size_t)smem, st, ta);
    if (le != cudaSuccess) {
        set_error("hgk_conv_tc_nhwc (image-tile kernel): launch: %s", cudaGetErrorString(le));
        return HGK_ECUDA;
    }
    return HGK_OK;
}

template <int BN, bool SPLIT, int KS, bool BWDSTATS, bool BNAPPLY = false, bool XBF = false>
static int launch_tc2_cfg(const TcArgs& ta, cudaStream_t st) {
    // 16x8 tiles (twice the CTAs, two per SM: prologue / epilogue of one overlaps the main loop of the other) up to the
    // 64x64 layers; 16x16 tiles (half the weight stream per pixel) for the 128x128 layers.  Measured on the config-2
    // step: switch-over at 148 tiles 12.44 ms, at 400 tiles 12.28 ms, always 16x8 12.31 ms.
    const long long tiles16 = (long long)ta.c.N * (ta.c.H >> 4) * (ta.c.W >> 4);
    static long long below = -1;          // HGK_TC2_SUB1_BELOW overrides the switch-over point (tuning knob)
    if (below < 0) {
        const char* e = getenv("HGK_TC2_SUB1_BELOW");
        below = e != nullptr ? atoll(e) : 400;
    }
    if (tiles16 < below) return launch_tc2_sub<BN, SPLIT, KS, BWDSTATS, BNAPPLY, 1, XBF>(ta, st);
    return launch_tc2_sub<BN, SPLIT, KS, BWDSTATS, BNAPPLY, 2, XBF>(ta, st);
}

template <int BN, int KS>
static int launch_tc2_bn(const TcArgs& ta, bool split, bool bwdstats, cudaStream_t st) {
    if (ta.c.ap.z != nullptr)        // data gradient with the BatchNorm-backward apply evaluated on load
        return bwdstats ? launch_tc2_cfg<BN, false, KS, true, true>(ta, st) : launch_tc2_cfg<BN, false, KS, false, true>(ta, st);
    if (bwdstats) return launch_tc2_cfg<BN, false, KS, true>(ta, st);
    if (split) {
        // TF32 + 2xBF16 products (w_lo in pack mode 2): 3x3 forward with 64 / 128 output channels only
        if (ta.lo_bf16) {
            // 64 output channels (the 128x128 layer): 16x8 tiles -- 192 TMEM columns, so two CTAs share an SM and the
            // producer chain / epilogue of one runs under the MMAs of the other (283 -> 227 us; the data gradient of the
            // same layer is faster on 16x16 tiles and keeps them)
            if (KS == 3 && BN == 64) return launch_tc2_sub<BN, true, KS == 3 ? 3 : 1, false, false, 1, (KS == 3 && BN == 64)>(ta, st);
            if (KS == 3 && BN <= 128) return launch_tc2_cfg<BN, true, KS == 3 ? 3 : 1, false, false, (KS == 3 && BN <= 128)>(ta, st);
            set_error("hgk_conv_tc_bn_x2_nhwc: no TF32 + 2xBF16 instantiation for k=%d Cout=%d", KS, BN);
            return HGK_EINVAL;
        }
        return launch_tc2_cfg<BN, true, KS, false>(ta, st);
    }
    return launch_tc2_cfg<BN, false, KS, false>(ta, st);
}

// true when the tile kernel covers this problem (the caller falls back to conv_tc_kernel otherwise)
bool conv_tc2_eligible(const TcArgs& ta) {
    const ConvArgs& c = ta.c;
    if (c.H % 16 || c.W % 16) return false;
    if (c.ksize == 3 && c.Cout > 128) return false;                       // 3x3 with 256 outputs: no instantiation
    const long long cmax = c.Cin > c.Cout ? c.Cin : c.Cout;
    if (c.P * cmax >= (1LL << 32)) return false;                          // 32-bit element offsets in the producers
    if ((long long)c.N * (c.H >> 4) * (c.W >> 4) >= (1LL << 31)) return false;
    return true;
}

int conv_tc2_launch(const TcArgs& ta, bool split, bool bwdstats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int BN = ta.c.Cout;
    if (ta.c.ksize == 4) {
        // the stem in space-to-depth form (32 -> 64 channels, forward only): 16x8 tiles, two CTAs per SM
        if (BN != 64 || !split || bwdstats || ta.c.ap.z != nullptr) {
            set_error("hgk_conv_tc_nhwc: k = 4 is the forward stem convolution only (32 -> 64 channels, w_lo given)");
            return HGK_EINVAL;
        }
        return ta.lo_bf16 ? launch_tc2_sub<64, true, 4, false, false, 1, true>(ta, st)
                          : launch_tc2_sub<64, true, 4, false, false, 1, false>(ta, st);
    }
    if (ta.c.ksize == 3) {
        if (BN == 64) return launch_tc2_bn<64, 3>(ta, split, bwdstats, st);
        return launch_tc2_bn<128, 3>(ta, split, bwdstats, st);
    }
    if (BN == 64) return launch_tc2_bn<64, 1>(ta, split, bwdstats, st);
    if (BN == 128) return launch_tc2_bn<128, 1>(ta, split, bwdstats, st);
    return launch_tc2_bn<256, 1>(ta, split, bwdstats, st);
}

}  // namespace hgk
