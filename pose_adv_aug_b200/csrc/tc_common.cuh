// Shared device helpers of the tcgen05 kernels (mbarrier, bulk copy, UMMA descriptors/issue, TMEM loads).
#pragma once
#include "common.cuh"
#include "conv_args.cuh"

namespace hgk {

struct TcArgs {
    ConvArgs c;
    const float* w_hi;
    const float* w_lo;
    long long* dbg;      // optional timeline buffer (developer diagnostics): [cta][16] globaltimer stamps
    int lo_bf16;         // w_lo holds the bf16 cross-term operands [wh_bf16 | wl_bf16] (pack mode 2) instead of fp32 w - wh
};

extern long long* g_dbg_buf;

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define HGK_STAMP(slot)                                                                   \
    do {                                                                                  \
        if (args.dbg != nullptr && blockIdx.x < 256) args.dbg[blockIdx.x * 16 + (slot)] = gtimer(); \
    } while (0)

// per-chunk trace of CTA 0 (SM cycle counter): row 256 + kind, column = chunk index (< 16)
#define HGK_TRACE(kind, col)                                                                      \
    do {                                                                                          \
        if (args.dbg != nullptr && blockIdx.x == 0 && (col) < 16) args.dbg[(256 + (kind)) * 16 + (col)] = clock64(); \
    } while (0)

constexpr int TBM = 128, TNT = 256;
__host__ __device__ constexpr int tc_staging_bytes(int BN) { return TBM * (BN + 4) * 4 + 16384; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Blocking wait.  The suspend-time hint matters: without it try_wait returns after a very short hardware nap and the
// waiting warps (a whole role of a warp-specialised kernel) spin through the loop, taking issue slots from the warps of the
// same scheduler that have work (profiles/r3_persistent_kernel.md: the epilogue warps ran at a third of their speed).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "HGK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra HGK_DONE_%=;\n\t"
        "bra HGK_WAIT_%=;\n\t"
        "HGK_DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// kind::f16 with bf16 operands (K = 16 per instruction), fp32 accumulation in TMEM
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// two fp32 -> one register of two bf16 (round to nearest even), low half = first argument
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B (rows 16 B apart);
// LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version 1 (Blackwell)
    return d;
}
// round-to-nearest (ties away) to TF32 = add half an ulp of the 10-bit mantissa and clear the low 13 bits.
// (cvt.rna.tf32.f32 has no single-instruction lowering on sm_100a: ptxas expands it to ~8 integer ops,
// which made the producers issue-bound; inf stays inf, NaN stays NaN.)
__device__ __forceinline__ float tf32_rna(float v) {
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 tf32_rna4(float4 v) {
    return make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
}
// BN+ReLU on load with the ReLU expressed as a clamp value (0 or -inf): one FFMA + one FMNMX per element
__device__ __forceinline__ float4 actc4(float4 v, float4 s, float4 t, float clampv) {
    return make_float4(fmaxf(fmaf(v.x, s.x, t.x), clampv), fmaxf(fmaf(v.y, s.y, t.y), clampv),
                       fmaxf(fmaf(v.z, s.z, t.z), clampv), fmaxf(fmaf(v.w, s.w, t.w), clampv));
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


}  // namespace hgk
