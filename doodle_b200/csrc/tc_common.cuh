// Thin inline-PTX layer for the Blackwell (sm_100a) features the splat kernels use:
// mbarrier, tcgen05 (alloc / mma kind::tf32 / commit / ld / fences), UMMA shared-memory descriptors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace helio {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (An arrive WITHOUT release semantics, mbarrier.arrive.relaxed, was tried for the producers' "stage written" signal: the
// MEMBAR.ALL.CTA in front of it stays -- it belongs to fence.proxy.async, not to the release -- so it bought nothing.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint, in nanoseconds,
// runs out) instead of returning after its short default limit.  Without it the waiting warps of the pipeline spin --
// ncu counted ~1e9 of the 3.5e9 warp instructions of the backward splat in TRYWAIT / BRA loops -- and steal issue slots
// from the warps that share their scheduler (21 warps per SM: the kernel is bound by instruction issue).
#ifndef HELIO_MBAR_SUSPEND_NS
#define HELIO_MBAR_SUSPEND_NS 20000
#endif
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)HELIO_MBAR_SUSPEND_NS)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if HELIO_MBAR_SUSPEND_NS > 0
    while (!mbar_try_wait_hint(bar, parity)) {
    }
#else
    while (!mbar_try_wait(bar, parity)) {
    }
#endif
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the same for every state space (a CTA pair's leader orders its peer's shared memory too)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }

// ---- explicit shared-state-space accesses (32-bit shared addresses) --------------------------
// The operand tiles are addressed through pointers the compiler cannot prove to be shared memory; generic
// LD/ST would go through the slower generic path (and the long scoreboard), so the hot loops use these.
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {  // read-only tables: may be hoisted / CSE'd
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// one lane of the (converged) warp: the lane tcgen05.mma / commit are issued from
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05 -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with fp16 operands (kind::f16, K = 16 per instruction)
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread -> one arrival on `bar` when they complete
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
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
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- thread-block clusters / CTA pairs (cta_group::2) ------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.  Default semantics
// (release at CTA scope): the data this guards is shared memory consumed by the async proxy, made visible by
// fence.proxy.async before the arrive; a cluster-scope release would add MEMBAR.GPU + an L1 invalidate per call.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
// long waits (epilogue waiting for a whole tile): back off so the spin does not eat issue slots
template <int NS = 64>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(NS);
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {  // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 (128 rows per CTA) and B split by N halves across the pair;
// issued by ONE thread of the leader CTA, descriptors are CTA-relative and apply to both CTAs.
__device__ __forceinline__ void mma_tf32_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all previously issued MMAs -> one arrival on `bar` (same offset) in every CTA of `mask`
__device__ __forceinline__ void mma_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------
// K-major operand tile, SWIZZLE_128B: rows of 128 bytes (32 tf32), 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
//  layout_type [61,64) with SWIZZLE_128B = 2.)  The tile base must be 1024-byte aligned; a K step of
// 8 tf32 inside the 128-byte row is +32 bytes on the start address.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;              // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;    // SBO: 8 rows * 128 B
    d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}
// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) {
    return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4);
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 instruction descriptor: D=f32, A=B=fp16 (format 0), both K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void sts_v4_b32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// v (already scaled into fp16 range) -> two fp16 pieces, p1 = rn(v), p2 = rn(v - p1): 11 + 11 significant bits
__device__ __forceinline__ void split_f16x2(float v0, float v1, uint32_t& p1, uint32_t& p2);
__device__ __forceinline__ void sts_v2_b32(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// two fp32 -> packed fp16x2 (round to nearest) and back
__device__ __forceinline__ uint32_t f2h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void h22f(uint32_t h, float& lo, float& hi) {
    asm("{\n\t.reg .f16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(lo), "=f"(hi) : "r"(h));
}

__device__ __forceinline__ void split_f16x2(float v0, float v1, uint32_t& p1, uint32_t& p2) {
    p1 = f2h2(v0, v1);
    float f0, f1;
    h22f(p1, f0, f1);
    p2 = f2h2(v0 - f0, v1 - f1);
}
struct f32x2;
__device__ __forceinline__ void split_f16x2_packed(unsigned long long v2, uint32_t& p1, uint32_t& p2);

// split an fp32 into a tf32-exact high part and the fp32 remainder (the tensor core reads the top
// 19 bits of each operand; hi + lo == v exactly)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    lo = v - hi;
}

// shared-memory read that must not be hoisted or merged (staging buffers rewritten every iteration)
__device__ __forceinline__ float4 lds_v4_volatile(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, two IEEE results per instruction) ------------
struct f32x2 {
    unsigned long long r;
};
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p.r) : "f"(lo), "f"(hi));
    return p;
}
__device__ __forceinline__ void unpack2(f32x2 p, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p.r));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 c;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c.r) : "l"(a.r), "l"(b.r));
    return c;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 c;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c.r) : "l"(a.r), "l"(b.r));
    return c;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.r) : "l"(a.r), "l"(b.r), "l"(c.r));
    return d;
}

// Truncating split of a pair {v0, v1} (already scaled into fp16 range) into fp16 pieces: hi = v with the 13 low mantissa
// bits cleared (exactly an fp16 value once |v| >= 2^-14), lo = v - hi (exact), p1 = cvt(hi), p2 = rn(lo).  The same
// 11 + 11 bits as the rounding split to within one unit of the last place (|v - p1 - p2| <= 2^-21 |v|, what the tf32
// hi / lo split keeps), with two logic operations and one packed subtraction instead of two fp16 -> fp32 conversions.
// Below 2^-14 the first conversion rounds to a subnormal: absolute error <= 2^-25, i.e. 2^-39 of the 2^14 scale.
__device__ __forceinline__ void split_f16x2_packed(unsigned long long v2, uint32_t& p1, uint32_t& p2) {
    float v0, v1, r0, r1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(v2));
    const float h0 = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u);
    p1 = f2h2(h0, h1);
    f32x2 d = add2(f32x2{v2}, pack2(-h0, -h1));
    unpack2(d, r0, r1);
    p2 = f2h2(r0, r1);
}

}  // namespace tc
}  // namespace helio
