// K2 / K3 on the 5th-generation tensor cores (tcgen05, accumulators in TMEM), 3xTF32.
//
//   forward : img[b]      = Gx^T diag(amp) Gy                       (K = heliostats)
//   backward: T[n,i]      = sum_j g[i,j] Gy[n,j]                    (K = image columns)
//             U[n,j]      = sum_i g[i,j] amp Gx[n,i]                (K = image rows)
//             S0 = sum_i amp Gx T,  Sx = sum_i amp Gx (x_i-a) T,  Sxx likewise with (x_i-a)^2,
//             Sy = sum_j Gy (y_j-b) U,  Syy likewise              -> moments {S0,Sx,Sy,Sxx+Syy}
//
// Both kernels are persistent (one CTA per SM), warp-specialised, and generate their Gaussian
// operands on the fly: no [B,N,R,R] tensor, no [N,R] operand matrices in HBM.
//
//   producer warps : evaluate exp2(-k2 (x - c)^2), split each value into a tf32-exact high part
//                    and the fp32 remainder, and write both as K-major SWIZZLE_128B tiles;
//                    (backward) a second producer group stages the image-gradient tile of g,
//                    split the same way (transposed on the fly for the U product);
//   MMA warp       : one elected thread issues, per 32-deep K stage, 4 K-steps x {hi*hi, hi*lo,
//                    lo*hi} tcgen05.mma kind::tf32 into a fp32 accumulator in TMEM; two
//                    accumulators ping-pong so the epilogue overlaps the next tile;
//   epilogue warps : tcgen05.ld the accumulator; forward stores image rows, backward folds the
//                    recomputed Gaussian weights into the per-heliostat moments.
// Pipelines: smem full/empty per stage (producers <-> MMA), TMEM full/empty per accumulator
// (MMA <-> epilogue), static round-robin tile schedule.
//
// CG = 2 runs the same kernels on CTA pairs (cta_group::2, cluster of two SMs): the pair shares one
// 256-row accumulator tile (128 TMEM lanes per CTA), each CTA generates its 128 A rows and only its
// half of the B rows, and the leader's MMA reads both shared memories.  That cuts the operand bytes
// each SM writes and the tensor core reads per MMA by a third and leaves room for a third stage.
//
// Accuracy: hi*hi + hi*lo + lo*hi with fp32 accumulation drops only lo*lo (~2^-22 relative) and
// the truncation of lo (~2^-21): fp32-class results from the tensor pipe.
#pragma once
#include <type_traits>
#include "helio_common.cuh"
#include "tc_common.cuh"

#ifndef HELIO_FWD_DEFER
#define HELIO_FWD_DEFER 1   // measured on B200: -5.7 % forward time (3.73 -> 3.53 ms at N=2000, R=256, B=4096)
#endif
#ifndef HELIO_PACKED2
#define HELIO_PACKED2 0
#endif
#ifndef HELIO_FWD_F16_TRUNC
// f16x3 forward producers: 1 = packed fp32x2 exponent arithmetic + truncating two-piece split (piece 1 = the top 11
// significand bits, piece 2 = fp16(v - piece 1): |error| <= 2^-22 v, the K3 split), 0 = round-to-nearest piece 1 with the
// residual taken through an fp16 -> fp32 round trip (2^-23 v, 1.5 more instructions per Gaussian).
#define HELIO_FWD_F16_TRUNC 1
#endif
#ifndef HELIO_FWD_DUO
#define HELIO_FWD_DUO 1     // forward, images of at most 64 pixels a side: two images per 128 x 128 tile (0: one per 128 x 64 tile)
#endif
#ifndef HELIO_FWD_SWP
#define HELIO_FWD_SWP 1     // f16x3 forward producers software-pipelined across stages (see the kernel)
#endif
#ifndef HELIO_FWD_SWP_MIN_N
// the pipeline restarts at every tile (its first stage is evaluated on its own), so it needs several stages per sun to pay:
// measured on B200 -3.5 % at N = 2000 (R = 256), -9 % at N = 5000 (R = 64), +10 % at N = 50 (two stages per tile)
#define HELIO_FWD_SWP_MIN_N 256
#endif
#ifndef HELIO_FWD_FULL_STAGE
// 1: stages whose 32 heliostats all exist (every stage but a sun's last) skip the per-heliostat bounds work: no index clamps
// on the parameter loads (one pointer, four immediate offsets) and no validity selects on the exponent offsets.  Measured on
// B200: -7 % with the round-to-nearest split, nothing on top of the truncating split (the producers are then no longer bound by
// instruction issue) at +40 % code size, so it stays an A/B switch.
#define HELIO_FWD_FULL_STAGE 0
#endif
#ifndef HELIO_BWD_DEFER
#define HELIO_BWD_DEFER 0
#endif

#ifndef HELIO_CONSUMER_FENCE
// Where the generic -> async proxy fence between the producers' st.shared and the tensor core's operand reads sits.
// 0: in every producer thread before its warp's arrive (the CUTLASS pattern).  fence.proxy.async compiles to MEMBAR.ALL.CTA +
//    FENCE.VIEW.ASYNC.S, and the MEMBAR also waits for the thread's outstanding global loads (the parameter prefetch).
// 1: in the MMA warp after its wait on the full barrier: st.shared -> __syncwarp -> arrive.release -> wait.acquire ->
//    fence.proxy.async -> tcgen05.mma is a causality chain with the proxy fence in it (PTX memory model, proxy-preserved base
//    causality order), and the producers no longer stall on a membar.
#define HELIO_CONSUMER_FENCE 0
#endif
#ifndef HELIO_BWD_STAGER_ROWMAP
// 1: the K3 stagers of product 0 visit the rows of a swizzle atom in the bank-conflict-free order of the forward producers
// (measured on B200: K3 -3.4 % at R = 64, -1.2 % at R = 128, unchanged at R = 256); 0: consecutive rows
#define HELIO_BWD_STAGER_ROWMAP 1
#endif
#ifndef HELIO_EPI_SLEEP_FWD_NS
// nanosleep between the epilogue warps' polls of the accumulator-full barrier.  The polls are 12-16 % of the warp instructions
// the two kernels execute (ncu source page); 256 / 1000 / 4000 ns instead of 64 changed neither kernel's time on B200.
#define HELIO_EPI_SLEEP_FWD_NS 64
#endif
#ifndef HELIO_EPI_SLEEP_BWD_NS
#define HELIO_EPI_SLEEP_BWD_NS 64
#endif
#ifndef HELIO_TC_STATS
#define HELIO_TC_STATS 0    // 1: per-warp cycle accounting of the pipeline roles (debug builds only; scripts/tc_stats.py)
#endif

namespace helio {

#if HELIO_TC_STATS
// [CTA][warp][slot]: 0 = total cycles in the role loop, 1..3 = cycles blocked in the role's waits (see the kernels)
__device__ unsigned long long g_tc_stats[160][24][4];
__device__ __forceinline__ unsigned long long tc_clock() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}
struct TcStat {
    unsigned long long t0, acc[4];
    __device__ TcStat() : t0(tc_clock()), acc{0, 0, 0, 0} {}
    __device__ void flush() {
        acc[0] = tc_clock() - t0;
        if ((threadIdx.x & 31) == 0 && blockIdx.x < 160)
            for (int k = 0; k < 4; ++k) g_tc_stats[blockIdx.x][threadIdx.x >> 5][k] = acc[k];
    }
};
#define TC_STAT_DECL TcStat tcst
#define TC_STAT_BEGIN unsigned long long tcs_t = tc_clock()
#define TC_STAT_END(slot) tcst.acc[slot] += tc_clock() - tcs_t
#define TC_STAT_FLUSH tcst.flush()
#else
#define TC_STAT_DECL
#define TC_STAT_BEGIN
#define TC_STAT_END(slot)
#define TC_STAT_FLUSH
#endif

constexpr int kTcMaxR = 1024;  // coordinate tables live in shared memory

// Always-on clock probe: CTA 0 of a tcgen05 kernel leaves {SM cycles, nanoseconds} of its own lifetime here ([0] forward,
// [1] backward), i.e. the SM clock the kernel actually HELD (B200 power-throttles inside these kernels: ~1.6-1.8 GHz
// under full tensor load while NVML still reports the 1965 MHz application clock).  helio_tc_clock_mhz reads it; bench.py
// quotes the tensor-pipe utilisation at this clock.
__device__ unsigned long long g_tc_clock[2][2];
__device__ __forceinline__ unsigned long long probe_cycles() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long probe_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------------------------------------
// pieces shared by both kernels
// ------------------------------------------------------------------------------------------------
template <int NT, int CG, int BROWS, int ASPLIT, int BSPLIT = 1, int TILES = 2, int EPI_BYTES = 0, int KC = 32, int F16PIECES = 0,
          int BALT = 1, int ESPLIT = 1>
struct SplatTcLayout {
    static constexpr int kNT = NT;                       // UMMA N (accumulator columns)
    static constexpr int kM = 128;                       // A rows per CTA = TMEM lanes
    static constexpr int kBRows = BROWS;                 // B operand rows produced by each CTA (NT / CG)
    static constexpr int kKC = KC;                       // K per stage: 32 (one 128-byte swizzle row of tf32) or 64 (of fp16)
    // F16PIECES = 1 (backward "f16x3", K = 64): the two tiles of an operand hold fp16 pieces p1 / p2 of 64 K values per
    // 128-byte row instead of tf32 hi / lo of 32; same bytes, same descriptors, kind::f16 MMAs (K = 16 = 32 bytes per step)
    static constexpr int kF16Pieces = F16PIECES;
    static_assert(KC == 32 || (KC == 64 && F16PIECES == 1 && TILES == 2), "K per stage");
    static constexpr int kASplit = ASPLIT;               // warps sharing one 32-row slab of A (each takes kKC / ASPLIT of K)
    static constexpr int kAWarps = kM / 32 * ASPLIT;
    static constexpr int kBSplit = BSPLIT;               // same for the B operand rows
    // BALT groups of B-operand warps take the stages in turn (backward gradient stagers: a group then has BALT stage
    // periods for the load latency of its tile instead of one); only one group arrives on a stage's full barrier
    static constexpr int kBAlt = BALT;
    static constexpr int kBWarps = kBRows / 32 * BSPLIT * BALT;
    static constexpr int kMmaWarp = kAWarps + kBWarps;
    static constexpr int kEpiWarp0 = kMmaWarp + 1;
    // ESPLIT epilogue warps per TMEM lane quarter (each takes 1 / ESPLIT of an accumulator's columns; backward only)
    static constexpr int kEpiSplit = ESPLIT;
    static constexpr int kEpiWarps = 4 * ESPLIT;
    static constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
    static constexpr int kABytes = kM * 128;             // one of {hi, lo}
    static constexpr int kBBytes = kBRows * 128;
    // TILES = 2: {hi, lo} tf32 tiles per operand (3xTF32).  TILES = 1: one tile per operand whose 128-byte rows hold
    // two fp16 pieces of the 32 K values, [piece 1 | piece 2] (forward "f16x3" variant, see splat_fwd_tc_kernel).
    static constexpr int kTiles = TILES;
    static constexpr int kBOff = TILES * kABytes;        // byte offset of the B tiles inside a stage
    static constexpr int kStageBytes = TILES * (kABytes + kBBytes);
    static constexpr int kTmemCols = 2 * NT;             // two accumulators
    static constexpr int kTableBytes = 2 * kTcMaxR * 4;
    static constexpr int kEpiBytes = EPI_BYTES;          // epilogue staging (forward: 4 warps x 32 rows x 128 B)
    static constexpr int kFixedBytes = kTableBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI_BYTES;
    static constexpr int kFit = (227 * 1024 - kFixedBytes) / kStageBytes;
    static constexpr int kStages = kFit > (TILES == 1 ? 6 : 4) ? (TILES == 1 ? 6 : 4) : kFit;
    static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
    static constexpr int kFullCount = (kAWarps + kBWarps / BALT) * CG;   // one arrival per producer warp (of the group whose turn it is) of the pair
    static constexpr int kTEmptyCount = kEpiWarps * CG;           // one arrival per epilogue warp of the pair
    static_assert(CG == 1 || CG == 2, "CTA group size");
    static_assert(BROWS * CG == NT, "each CTA of the group stages NT / CG rows of B");
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32, "TMEM columns");
    static_assert(kStages >= 2 && kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert((2 * kStages + 4) * 8 + 8 <= 224, "barrier block (the last 32 bytes hold the clock probe)");
};

// shared-memory carve-up + one-time setup common to the forward and backward kernels
template <class C, int CG>
struct SplatTcCtx {
    uint8_t* smem;        // operand stages (1024-byte aligned)
    uint32_t smem_u;      // same, as a 32-bit shared-state-space address
    float *sX, *sY;       // pixel-centre tables
    uint32_t sX_u, sY_u;
    uint64_t *full, *empty, *tfull, *tempty;
    uint32_t epi_u;       // epilogue staging buffer (32-bit shared address)
    uint32_t tmem_base;
    uint32_t rank;        // CTA rank inside the pair (0 when CG == 1)
    uint64_t* probe;      // clock probe start values (CTA 0, thread 0)

    __device__ __forceinline__ void setup(uint8_t* smem_raw, int R, const Axis& ax, const Axis& ay) {
        smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
        sX = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
        sY = sX + kTcMaxR;
        smem_u = tc::smem_u32(smem), sX_u = tc::smem_u32(sX), sY_u = tc::smem_u32(sY);
        uint64_t* bars = reinterpret_cast<uint64_t*>(sY + kTcMaxR);
        full = bars;                       // [kStages]  producers -> MMA   (leader's copy is the live one)
        empty = bars + C::kStages;         // [kStages]  MMA -> producers
        tfull = bars + 2 * C::kStages;     // [2]        MMA -> epilogue
        tempty = tfull + 2;                // [2]        epilogue -> MMA   (leader's copy is the live one)
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
        epi_u = tc::smem_u32(reinterpret_cast<uint8_t*>(bars) + 256);
        probe = bars + 28;                 // bytes 224..239 of the barrier block
        if (blockIdx.x == 0 && threadIdx.x == 0) probe[0] = probe_cycles(), probe[1] = probe_ns();
        rank = CG == 2 ? tc::cluster_ctarank() : 0u;
        const int warp = threadIdx.x >> 5;

        // pixel-centre tables, padded with 0 (rows/columns >= R are computed but never stored)
        for (int i = threadIdx.x; i < kTcMaxR; i += C::kThreads) {
            sX[i] = i < R ? axis_at(ax, i) : 0.f;
            sY[i] = i < R ? axis_at(ay, i) : 0.f;
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < C::kStages; ++s) {
                tc::mbar_init(&full[s], C::kFullCount);
                tc::mbar_init(&empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) {
                tc::mbar_init(&tfull[a], 1);
                tc::mbar_init(&tempty[a], C::kTEmptyCount);
            }
            tc::mbar_fence_init();
        }
        if (warp == C::kMmaWarp) {
            if constexpr (CG == 2) {
                tc::tmem_alloc_2cta(tmem_slot, C::kTmemCols);
                tc::tmem_relinquish_2cta();
            } else {
                tc::tmem_alloc(tmem_slot, C::kTmemCols);
                tc::tmem_relinquish();
            }
        }
        tc::tc_fence_before();
        __syncthreads();
        if constexpr (CG == 2) tc::cluster_sync();   // peer barriers initialised before any remote arrive
        tc::tc_fence_after();
        tmem_base = *tmem_slot;
    }

    __device__ __forceinline__ void teardown(int which) {
        tc::tc_fence_before();
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            g_tc_clock[which][0] = probe_cycles() - probe[0];
            g_tc_clock[which][1] = probe_ns() - probe[1];
        }
        if constexpr (CG == 2) tc::cluster_sync();   // no CTA leaves while its peer may still signal it
        if ((threadIdx.x >> 5) == C::kMmaWarp) {
            tc::tc_fence_after();
            if constexpr (CG == 2) tc::tmem_dealloc_2cta(tmem_base, C::kTmemCols);
            else tc::tmem_dealloc(tmem_base, C::kTmemCols);
        }
    }

    // producer warp: stage s written (all lanes) -> one arrival on the group's full barrier
    __device__ __forceinline__ void producer_commit(int s) const {
#if !HELIO_CONSUMER_FENCE
        tc::fence_proxy_async_smem();
#endif
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            if constexpr (CG == 2) tc::mbar_arrive_remote(&full[s], 0);
            else tc::mbar_arrive(&full[s]);
        }
    }
    // producer warp: wait until the MMAs that read stage s have retired
    __device__ __forceinline__ void producer_acquire(int s, uint32_t ph) const {
        if ((threadIdx.x & 31) == 0) tc::mbar_wait(&empty[s], ph ^ 1);
        __syncwarp();
    }
    // epilogue warp: accumulator drained -> one arrival on the group's tempty barrier
    __device__ __forceinline__ void epilogue_release(int acc) const {
        tc::tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            if constexpr (CG == 2) tc::mbar_arrive_remote(&tempty[acc], 0);
            else tc::mbar_arrive(&tempty[acc]);
        }
    }
    // MMA thread: one stage = 4 K-steps x {hi*hi, hi*lo, lo*hi}; `first` clears the accumulator
    __device__ __forceinline__ void issue_stage(int s, uint32_t d_tmem, bool first, bool last, int acc) const {
        const uint32_t sa = smem_u + (uint32_t)(s * C::kStageBytes);
        if constexpr (C::kTiles == 1) {
            // fp16 pieces: rows are [a1 (K 0..31, 64 B) | a2 (64 B)]; products a1 b1 + a1 b2 + a2 b1, K = 16 (32 bytes) per MMA
            constexpr uint32_t idesc16 = tc::make_idesc_f16(C::kM * CG, C::kNT);
            const uint64_t a = tc::make_desc_k_sw128(sa), b = tc::make_desc_k_sw128(sa + C::kBOff);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2);           // +32 bytes per K step (descriptor units of 16 bytes)
                if constexpr (CG == 2) {
                    tc::mma_f16_ss_2cta(d_tmem, a + o, b + o, idesc16, !(first && ks == 0));
                    tc::mma_f16_ss_2cta(d_tmem, a + o, b + 4 + o, idesc16, 1);
                    tc::mma_f16_ss_2cta(d_tmem, a + 4 + o, b + o, idesc16, 1);
                } else {
                    tc::mma_f16_ss(d_tmem, a + o, b + o, idesc16, !(first && ks == 0));
                    tc::mma_f16_ss(d_tmem, a + o, b + 4 + o, idesc16, 1);
                    tc::mma_f16_ss(d_tmem, a + 4 + o, b + o, idesc16, 1);
                }
            }
            if constexpr (CG == 2) {
                tc::mma_commit_2cta(&empty[s], 3);
                if (last) tc::mma_commit_2cta(&tfull[acc], 3);
            } else {
                tc::mma_commit(&empty[s]);
                if (last) tc::mma_commit(&tfull[acc]);
            }
            return;
        }
        constexpr uint32_t idesc = C::kF16Pieces ? tc::make_idesc_f16(C::kM * CG, C::kNT) : tc::make_idesc_tf32(C::kM * CG, C::kNT);
        const uint64_t a_hi = tc::make_desc_k_sw128(sa);
        const uint64_t a_lo = tc::make_desc_k_sw128(sa + C::kABytes);
        const uint64_t b_hi = tc::make_desc_k_sw128(sa + 2 * C::kABytes);
        const uint64_t b_lo = tc::make_desc_k_sw128(sa + 2 * C::kABytes + C::kBBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);   // +32 bytes per K step of 8 tf32
            if constexpr (C::kF16Pieces) {           // p1 q1 + p1 q2 + p2 q1, 16 fp16 (32 bytes) of K per instruction
                if constexpr (CG == 2) {
                    tc::mma_f16_ss_2cta(d_tmem, a_hi + ko, b_hi + ko, idesc, !(first && k == 0));
                    tc::mma_f16_ss_2cta(d_tmem, a_hi + ko, b_lo + ko, idesc, 1);
                    tc::mma_f16_ss_2cta(d_tmem, a_lo + ko, b_hi + ko, idesc, 1);
                } else {
                    tc::mma_f16_ss(d_tmem, a_hi + ko, b_hi + ko, idesc, !(first && k == 0));
                    tc::mma_f16_ss(d_tmem, a_hi + ko, b_lo + ko, idesc, 1);
                    tc::mma_f16_ss(d_tmem, a_lo + ko, b_hi + ko, idesc, 1);
                }
            } else if constexpr (CG == 2) {
                tc::mma_tf32_ss_2cta(d_tmem, a_hi + ko, b_hi + ko, idesc, !(first && k == 0));
#ifndef HELIO_DEBUG_1XTF32   // timing experiment only: results are wrong without the cross terms
                tc::mma_tf32_ss_2cta(d_tmem, a_hi + ko, b_lo + ko, idesc, 1);
                tc::mma_tf32_ss_2cta(d_tmem, a_lo + ko, b_hi + ko, idesc, 1);
#endif
            } else {
                tc::mma_tf32_ss(d_tmem, a_hi + ko, b_hi + ko, idesc, !(first && k == 0));
                tc::mma_tf32_ss(d_tmem, a_hi + ko, b_lo + ko, idesc, 1);
                tc::mma_tf32_ss(d_tmem, a_lo + ko, b_hi + ko, idesc, 1);
            }
        }
        if constexpr (CG == 2) {
            tc::mma_commit_2cta(&empty[s], 3);               // stage reusable in both CTAs when these MMAs retire
            if (last) tc::mma_commit_2cta(&tfull[acc], 3);   // accumulator complete in both CTAs
        } else {
            tc::mma_commit(&empty[s]);
            if (last) tc::mma_commit(&tfull[acc]);
        }
    }
    __device__ __forceinline__ void mma_wait_full(int s, uint32_t ph) const {
        tc::mbar_wait(&full[s], ph);
#if HELIO_CONSUMER_FENCE
        tc::fence_proxy_async_all();
#endif
        tc::tc_fence_after();
    }
    __device__ __forceinline__ void mma_wait_tempty(int acc, uint32_t aph) const {
        tc::mbar_wait(&tempty[acc], aph ^ 1);
        tc::tc_fence_after();
    }
};

// ================================================================================================
// forward
// ================================================================================================
// PREC = 0: 3xTF32 (tf32 hi/lo tiles).  PREC = 1 ("f16x3", opt-in): both operands are Gaussians in [0, 1], so they are
// scaled by 2^14 (exponent offset, exact) and split into two fp16 pieces v = p1 + p2 (11 + 11 significant bits, the same
// 2^-22 the tf32 hi/lo split keeps, for every value above 2^-17; absolute error < 2^-39 below); the three products
// p1 q1 + p1 q2 + p2 q1 run as kind::f16 MMAs at twice the tf32 rate on half the operand bytes (32 KB stages: six fit),
// and the epilogue unscales by 2^-28.  The image gradient in the backward has no such bound, so K3 stays 3xTF32.
template <int NT, int CG, int PS, int PREC = 0, int STG = 0>
using SplatFwdTc = SplatTcLayout<NT, CG, NT / CG, PS, PS, PREC == 1 ? 1 : 2, STG ? 16384 : 0>;

// Work fused into the forward epilogue while the accumulator row sits in registers (the kernel is tensor-bound and
// leaves HBM almost idle, so the HBM-bound passes of the loss block ride along for free):
//   kFuseMax : tile_max[b] = max over the image (atomicMax on the int pattern: values are >= 0, order-free), which
//              replaces image_max_kernel for the target render;
//   kFuseLoss: per-warp partials {sum diff^2, sum |diff| dmaps, sum |diff|}, diff = img/t - target/t, written to
//              partials[b][tile][cta][warp][3] and combined in index order by loss_pack_partials_kernel, which
//              replaces loss_fwd_kernel's pass over the image.
//   kFuseFeed: the encoder feed of the COM trainer (layers/center_of_mass.py:21-60, train_with_env.py:182-209): per-warp
//              partials {sum w, sum w j, sum w i}, w = max(img, 0), combined in index order by com_pack_partials_kernel
//              (replaces com_fwd_kernel's pass over the image), and an optional second copy of the image written
//              straight into a caller-provided slot (e.g. hist[:, -1] of the rollout's history buffer).
enum : int { kFuseNone = 0, kFuseMax = 1, kFuseLoss = 2, kFuseFeed = 3 };
struct FwdFuse {
    float* tile_max;         // [B]           kFuseMax (initialised to the floor by the caller)
    const float* target;     // [B][R][R]     kFuseLoss
    const float* dmaps;      // [B][R][R]
    const float* tx;         // [B]           (clamped to 1e-6 here)
    float* partials;         // [B][tiles][CG][4][3]   kFuseLoss / kFuseFeed
    float* img2;             // kFuseFeed: second destination of the image (may be NULL), image b at img2 + b * img2_bstride
    long long img2_bstride;  //            in floats
};

// Producer mapping: a warp owns 32 operand rows (image rows for A, image columns for B); a lane owns
// 4 consecutive heliostats of the stage (one 16-byte chunk of the K-major row) and walks 8 of the
// rows, so the footprint parameters sit in registers (loaded once per stage, prefetched one stage
// ahead) and every warp store writes four full 128-byte rows of the swizzled tile, conflict-free.
// STG = 1: the epilogue transposes each warp's 32 x 32 block through shared memory so that global stores are full 128-byte
// row segments.  Measured on B200: -3...-12 % when the epilogue is a large share of the tile (few heliostats, and with the
// f16x3 operands), +3...+5 % for large N under 3xTF32 (the staging competes with the operand traffic), so the launcher
// picks it by N and operand format.
template <int NT, int CG, int PS, int FUSE, int PREC, int STG>
__global__ void __launch_bounds__(SplatFwdTc<NT, CG, PS, PREC, STG>::kThreads, 1)
splat_fwd_tc_kernel(const float4* __restrict__ params, const int* __restrict__ counts, float* __restrict__ img, int N, int R,
                    Axis ax, Axis ay, int tiles_i, int tiles_j, int num_tiles, FwdFuse fz, int duo_B) {
    // duo_B > 0 ("two images per tile", images of at most 64 pixels a side on the 128 x 128 single-CTA tile): tile t holds
    // images 2 t and 2 t + 1 of duo_B, A = [Gx_2t ; Gx_2t+1], B = [Gy_2t ; Gy_2t+1] (64 rows each), and only the two diagonal
    // 64 x 64 blocks of the accumulator are images.  The off-diagonal half of the MMA work is wasted, but at this size the
    // kernel is bound by its producer warps, and a 128 x 64 tile would leave the two that own A rows 64..127 idle.
    using C = SplatFwdTc<NT, CG, PS, PREC, STG>;
    extern __shared__ uint8_t smem_raw[];
    SplatTcCtx<C, CG> cx;
    cx.setup(smem_raw, R, ax, ay);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heliostats contracted for sun b: all N, or the first counts[b] of a culled (compacted) parameter row; at least one
    // (all-padding) stage so that the accumulator is always written
    auto sun_count = [&](int b) { return counts ? __ldg(counts + b) : N; };
    auto sun_chunks = [&](int cnt) { return max(1, (cnt + C::kKC - 1) / C::kKC); };
    const int tiles_per_img = tiles_i * tiles_j;
    // stages of a tile: every role counts them alike (two-image tiles: the longer of the two culled lists)
    auto tile_chunks = [&](int tile) {
        if (duo_B > 0) return sun_chunks(max(sun_count(2 * tile), sun_count(min(2 * tile + 1, duo_B - 1))));
        return sun_chunks(sun_count(tile / tiles_per_img));
    };
    const int group = blockIdx.x / CG, ngroups = gridDim.x / CG;
    constexpr int kTileM = C::kM * CG;

    if (warp < C::kMmaWarp) {
        // ================= producers =================
        // A warp owns 32 / PS operand rows and all 32 heliostats of a stage (PS = 2: twice the producer warps, each
        // with half the rows; a K split instead would halve the bytes per row a store instruction covers and double
        // the shared-memory wavefronts -- measured slower).  lane = (row subgroup rs, K chunk ch): per step a warp
        // covers 4 rows x 32 heliostats, each lane evaluating its 4 heliostats for one row and storing them as one
        // 16-byte chunk, so every store instruction writes four full 128-byte rows of the swizzled tile.
        constexpr int PSX = C::kASplit;                  // = kBSplit
        constexpr int kRows = 32 / PSX;                  // operand rows per warp
        constexpr int kRS = 4;                           // rows per step
        constexpr int kSteps = kRows / kRS;
        static_assert(PSX == 1 || PSX == 2, "producer split");
        static_assert(kSteps % 2 == 0, "a pair of steps covers one 8-row swizzle atom");
        auto produce = [&](auto is_a_tag) {
        constexpr bool isA = decltype(is_a_tag)::value;                   // the two operands get their own instantiation: no per-stage selects
        const int pw = isA ? warp : warp - C::kAWarps;                   // producer index inside its operand
        const int wrow = pw * kRows;                                     // first operand row of this warp (CTA-local)
        const int rs = lane >> 3, ch = lane & 7;
        const uint32_t region = (isA ? 0u : (uint32_t)C::kBOff) + (uint32_t)(wrow >> 3) * 1024u;
        [[maybe_unused]] const uint32_t lo_delta = isA ? C::kABytes : C::kBBytes;
        // rows visited by this lane: wrow + lane_row(step).  A step covers rows {0, 4, 1, 5} or {2, 6, 3, 7} of an 8-row swizzle
        // atom, in that order over the lane groups rs = 0..3.  The f16x3 stores are 8 bytes per lane (one 64-byte half row per
        // piece and row) and the LSU handles a 64-bit warp store as two half-warps: each half (rs = 0, 1 / rs = 2, 3) then holds
        // one row whose half lands in banks 0-15 (swizzle phase < 4) and one in banks 16-31, i.e. 128 bytes over 32 banks, one
        // wavefront.  Consecutive rows {4 st .. 4 st + 3} put both half rows of a half-warp on the same 16 banks (ncu source
        // page: 3.6 wavefronts per STS.64 instead of 2, half of the kernel's shared-memory store wavefronts).
        auto lane_row = [&](int st) -> uint32_t {
            return 8u * (uint32_t)(st >> 1) + 2u * (uint32_t)(st & 1) + 4u * (uint32_t)(rs & 1) + (uint32_t)(rs >> 1);
        };
        auto row_off = [&](int st) -> uint32_t {
            const uint32_t row = lane_row(st);
            return (row >> 3) * 1024u + (row & 7u) * 128u + ((((uint32_t)ch) ^ (row & 7u)) << 4);
        };
        uint32_t it = 0;                             // global stage counter
        TC_STAT_DECL;
        if (PREC == 1 && HELIO_FWD_SWP && N >= HELIO_FWD_SWP_MIN_N) {
            // ---- f16x3, software-pipelined across stages --------------------------------------------------------------------
            // A warp issues in order, and a stage has two phases that use different pipes: the Gaussians (32 MUFU.EX2 per lane, 8
            // issue cycles each on the XU, little else) and the split + store (LOP3 / F2FP / FADD2 / STS, no MUFU).  Run back to
            // back they serialise (ncu: no pipe above 40 %, the scheduler idle 59 % of the cycles); here step st of stage c is
            // split and stored while step st of stage c + 1 is evaluated into the registers it frees, in one basic block, so
            // the MUFUs of one stage fill under the ALU work of the previous one.  Same arithmetic, bit-identical results.
            for (int tile = group; tile < num_tiles; tile += ngroups) {
                int b = tile / tiles_per_img;
                const int t = tile % tiles_per_img;
                int g0 = (isA ? (t / tiles_j) * kTileM + (int)cx.rank * C::kM : (t % tiles_j) * NT + (int)cx.rank * C::kBRows) + wrow;
                bool off_img = false;                // two-image tile, odd batch: the second image of the last tile does not exist
                if (duo_B > 0) {
                    b = 2 * tile + (wrow >> 6);
                    off_img = b >= duo_B;
                    b = min(b, duo_B - 1);
                    g0 = wrow & 63;
                }
                const float4* pb = params + (size_t)b * N;
                const int cnt = sun_count(b), nchunks = tile_chunks(tile), last = max(cnt, 1) - 1;
                if (g0 >= R || off_img) {
                    // rows beyond the image feed accumulator rows / columns that are never stored: nothing to write
                    for (int c = 0; c < nchunks; ++c, ++it) {
                        const int s = it % C::kStages;
                        cx.producer_acquire(s, (it / C::kStages) & 1);
                        cx.producer_commit(s);
                    }
                    continue;
                }
                const uint32_t tab = (isA ? cx.sX_u : cx.sY_u) + (uint32_t)g0 * 4u;
                float xr[kSteps];
#pragma unroll
                for (int st = 0; st < kSteps; ++st) xr[st] = tc::lds_f32(tab + 4u * lane_row(st));
                float4 pr[4];
                auto load_params = [&](int c) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) pr[e] = __ldg(pb + min(c * C::kKC + 4 * ch + e, last));
                };
                tc::f32x2 nc2[2], k22[2], la2[2], vp[kSteps][2];
                auto decode = [&](int c) {                   // pr (raw footprints of stage c) -> packed exponent coefficients
                    float ctr[4], nk2[4], la[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool have = c * C::kKC + 4 * ch + e < cnt;
                        ctr[e] = isA ? pr[e].x : pr[e].y;
                        nk2[e] = -pr[e].z;
                        la[e] = have ? (isA ? lg2_ftz(pr[e].w) : 0.f) + 14.f : -INFINITY;     // amplitude and the 2^14 scale; K padding: exact zeros
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        nc2[h] = tc::pack2(-ctr[2 * h], -ctr[2 * h + 1]);
                        k22[h] = tc::pack2(nk2[2 * h], nk2[2 * h + 1]);
                        la2[h] = tc::pack2(la[2 * h], la[2 * h + 1]);
                    }
                };
                auto eval = [&](int st) {
                    const tc::f32x2 x2 = tc::pack2(xr[st], xr[st]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const tc::f32x2 d2 = tc::add2(x2, nc2[h]);
                        float a0, a1;
                        tc::unpack2(tc::fma2(tc::mul2(d2, k22[h]), d2, la2[h]), a0, a1);
                        vp[st][h] = tc::pack2(ex2(a0), ex2(a1));
                    }
                };
                auto store = [&](uint32_t base, int st) {
                    uint32_t p1a, p1b, p2a, p2b;
                    tc::split_f16x2_packed(vp[st][0].r, p1a, p2a);
                    tc::split_f16x2_packed(vp[st][1].r, p1b, p2b);
                    const uint32_t row = lane_row(st), sw = row & 7u;
                    const uint32_t rb = base + (row >> 3) * 1024u + sw * 128u + 8u * ((uint32_t)ch & 1u);
                    tc::sts_v2_b32(rb + ((((uint32_t)ch >> 1) ^ sw) << 4), p1a, p1b);
                    tc::sts_v2_b32(rb + (((4u + ((uint32_t)ch >> 1)) ^ sw) << 4), p2a, p2b);
                };
                load_params(0);
                decode(0);
                if (nchunks > 1) load_params(1);
#pragma unroll
                for (int st = 0; st < kSteps; ++st) eval(st);
#pragma unroll 1
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % C::kStages;
                    const bool more = c + 1 < nchunks;
                    if (more) {
                        decode(c + 1);
                        if (c + 2 < nchunks) load_params(c + 2);      // lands under this stage's work
                    }
                    {
                        TC_STAT_BEGIN;
                        cx.producer_acquire(s, (it / C::kStages) & 1);
                        TC_STAT_END(2);
                    }
                    const uint32_t base = cx.smem_u + (uint32_t)(s * C::kStageBytes) + region;
                    if (more) {
#pragma unroll
                        for (int st = 0; st < kSteps; ++st) {
                            store(base, st);
                            eval(st);
                        }
                    } else {
#pragma unroll
                        for (int st = 0; st < kSteps; ++st) store(base, st);
                    }
                    {
                        TC_STAT_BEGIN;
                        cx.producer_commit(s);
                        TC_STAT_END(1);
                    }
                }
            }
        } else {
        int pending = -1;                            // stage whose stores are issued but not yet handed to the MMA warp
        static_assert(HELIO_FWD_DEFER == 1, "the forward producers are written for the deferred hand-off (measured: -5.7 %)");
        for (int tile = group; tile < num_tiles; tile += ngroups) {
            int b = tile / tiles_per_img;
            const int t = tile % tiles_per_img;
            int g0 = (isA ? (t / tiles_j) * kTileM + (int)cx.rank * C::kM : (t % tiles_j) * NT + (int)cx.rank * C::kBRows) + wrow;
            bool off_img = false;
            if (duo_B > 0) {
                b = 2 * tile + (wrow >> 6);
                off_img = b >= duo_B;
                b = min(b, duo_B - 1);
                g0 = wrow & 63;
            }
            // operand rows beyond the image only feed accumulator rows / columns that are never stored: skip the
            // Gaussians and leave the stage bytes as they are (accumulator rows and columns are independent)
            const bool dead = g0 >= R || off_img;
            const uint32_t tab = (isA ? cx.sX_u : cx.sY_u) + (uint32_t)g0 * 4u;
            float xr[kSteps];
#pragma unroll
            for (int st = 0; st < kSteps; ++st) xr[st] = tc::lds_f32(tab + 4u * lane_row(st));
            const float4* pb = params + (size_t)b * N;
            const int cnt = sun_count(b), nchunks = tile_chunks(tile), last = max(cnt, 1) - 1;
            // This lane's 4 heliostats of a stage arrive as raw float4 (index clamped so the load never needs a select),
            // prefetched one stage ahead into the buffer the current stage has just decoded.  The hand-over's
            // fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR waits for every outstanding
            // memory operation of the thread, i.e. for what is left of the L2 latency of that prefetch: 14 % of a producer
            // warp's time (scripts/tc_stats.py).  Moving the loads behind the fence, two stages ahead in two register buffers, was
            // measured: the wait drops to 8 %, but 16 more registers hit the 128-register ceiling of a 13-warp CTA (four warps
            // share one scheduler's 16 K registers), ptxas spills, and the forward gets slower (3.51 -> 4.5 ms).
            float4 prA[4];
            auto load_params = [&](float4 (&pr)[4], int c) {
                if (HELIO_FWD_FULL_STAGE && (c + 1) * C::kKC <= cnt) {
                    const float4* q = pb + (c * C::kKC + 4 * ch);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pr[e] = __ldg(q + e);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) pr[e] = __ldg(pb + min(c * C::kKC + 4 * ch + e, last));
                }
            };
            auto stage = [&](float4 (&pr)[4], const int c, auto full_tag) {
                float ctr[4], nk2[4], la[4];
                constexpr bool full = decltype(full_tag)::value;       // all 32 heliostats of the stage exist
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool have = full || c * C::kKC + 4 * ch + e < cnt;
                    ctr[e] = isA ? pr[e].x : pr[e].y;
                    nk2[e] = -pr[e].z;
                    // amplitude folded into the exponent (amp ~ 1: lg2.approx is exact to 2^-22 absolute there);
                    // K padding: 2^-inf = exact zeros
                    la[e] = have ? (isA ? lg2_ftz(pr[e].w) : 0.f) + (PREC == 1 ? 14.f : 0.f) : -INFINITY;   // f16x3: operands x 2^14
                }
                load_params(pr, c + 1 < nchunks ? c + 1 : c);      // next stage's footprints, into the buffer just decoded
                const int s = it % C::kStages;
                // Hand the PREVIOUS stage over only after this stage's Gaussians have been evaluated: its shared-memory
                // stores drain under the arithmetic.
                float v[kSteps][4];
                [[maybe_unused]] tc::f32x2 vp[kSteps][2];
                if (!dead) {
                    if constexpr (HELIO_PACKED2 || (HELIO_FWD_F16_TRUNC && PREC == 1)) {
                    // FADD2 / FMUL2 / FFMA2: the same IEEE operations, two heliostats per instruction
                    tc::f32x2 nc2[2], k22[2], la2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        nc2[h] = tc::pack2(-ctr[2 * h], -ctr[2 * h + 1]);
                        k22[h] = tc::pack2(nk2[2 * h], nk2[2 * h + 1]);
                        la2[h] = tc::pack2(la[2 * h], la[2 * h + 1]);
                    }
#pragma unroll
                    for (int st = 0; st < kSteps; ++st) {
                        const tc::f32x2 x2 = tc::pack2(xr[st], xr[st]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const tc::f32x2 d2 = tc::add2(x2, nc2[h]);
                            const tc::f32x2 a2 = tc::fma2(tc::mul2(d2, k22[h]), d2, la2[h]);
                            float a0, a1;
                            tc::unpack2(a2, a0, a1);
                            v[st][2 * h] = ex2(a0);
                            v[st][2 * h + 1] = ex2(a1);
                            if constexpr (PREC == 1) vp[st][h] = tc::pack2(v[st][2 * h], v[st][2 * h + 1]);
                        }
                    }
                    } else {
#pragma unroll
                    for (int st = 0; st < kSteps; ++st)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float d = xr[st] - ctr[e];
                            v[st][e] = ex2(fmaf(d * nk2[e], d, la[e]));
                        }
                    }
                }
                {
                    TC_STAT_BEGIN;
                    if (pending >= 0) cx.producer_commit(pending);
                    TC_STAT_END(1);
                }
                {
                    TC_STAT_BEGIN;
                    cx.producer_acquire(s, (it / C::kStages) & 1);
                    TC_STAT_END(2);
                }
                const uint32_t base = cx.smem_u + (uint32_t)(s * C::kStageBytes) + region;
                if constexpr (PREC == 1) {
                    // two fp16 pieces per value, packed [p1 (8 bytes) ... | p2 ...]: this lane's 4 heliostats are bytes
                    // 8 ch .. 8 ch + 7 of the first 64 bytes of the row (p1) and of the second 64 bytes (p2)
                    if (!dead) {
#pragma unroll
                        for (int st = 0; st < kSteps; ++st) {
#if HELIO_FWD_F16_TRUNC
                            uint32_t p1a, p1b, p2a, p2b;
                            tc::split_f16x2_packed(vp[st][0].r, p1a, p2a);
                            tc::split_f16x2_packed(vp[st][1].r, p1b, p2b);
#else
                            const uint32_t p1a = tc::f2h2(v[st][0], v[st][1]), p1b = tc::f2h2(v[st][2], v[st][3]);
                            float f0, f1, f2, f3;
                            tc::h22f(p1a, f0, f1);
                            tc::h22f(p1b, f2, f3);
                            const uint32_t p2a = tc::f2h2(v[st][0] - f0, v[st][1] - f1), p2b = tc::f2h2(v[st][2] - f2, v[st][3] - f3);
#endif
                            const uint32_t row = lane_row(st), sw = row & 7u;
                            const uint32_t rb = base + (row >> 3) * 1024u + sw * 128u + 8u * ((uint32_t)ch & 1u);
                            tc::sts_v2_b32(rb + ((((uint32_t)ch >> 1) ^ sw) << 4), p1a, p1b);
                            tc::sts_v2_b32(rb + (((4u + ((uint32_t)ch >> 1)) ^ sw) << 4), p2a, p2b);
                        }
                    }
                } else
                if (!dead) {
#pragma unroll
                    for (int st = 0; st < kSteps; ++st) {
                        float hi[4], lo[4];
#if HELIO_PACKED2
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            hi[2 * h] = __uint_as_float(__float_as_uint(v[st][2 * h]) & 0xFFFFE000u);
                            hi[2 * h + 1] = __uint_as_float(__float_as_uint(v[st][2 * h + 1]) & 0xFFFFE000u);
                            const tc::f32x2 l2 = tc::add2(tc::pack2(v[st][2 * h], v[st][2 * h + 1]), tc::pack2(-hi[2 * h], -hi[2 * h + 1]));
                            tc::unpack2(l2, lo[2 * h], lo[2 * h + 1]);
                        }
#else
#pragma unroll
                        for (int e = 0; e < 4; ++e) tc::split_tf32(v[st][e], hi[e], lo[e]);
#endif
                        const uint32_t dst = base + row_off(st);
                        tc::sts_v4(dst, hi[0], hi[1], hi[2], hi[3]);
                        tc::sts_v4(dst + lo_delta, lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                pending = s;
                ++it;
            };
            load_params(prA, 0);
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                if (HELIO_FWD_FULL_STAGE && (c + 1) * C::kKC <= cnt) stage(prA, c, std::true_type{});
                else stage(prA, c, std::false_type{});
            }
        }
#if HELIO_FWD_DEFER
        if (pending >= 0) cx.producer_commit(pending);
#endif
        }
        TC_STAT_FLUSH;
        };
        if (warp < C::kAWarps) produce(std::true_type{}); else produce(std::false_type{});
    } else if (warp == C::kMmaWarp) {
        // ================= MMA issuer (leader CTA of the group) =================
        if (cx.rank == 0) {
            TC_STAT_DECL;
            uint32_t it = 0, tcount = 0;
            for (int tile = group; tile < num_tiles; tile += ngroups, ++tcount) {
                const int acc = tcount & 1;
                const int nchunks = tile_chunks(tile);
                {
                    TC_STAT_BEGIN;
                    cx.mma_wait_tempty(acc, (tcount >> 1) & 1);
                    TC_STAT_END(1);
                }
                const uint32_t d_tmem = cx.tmem_base + (uint32_t)(acc * NT);
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % C::kStages;
                    {
                        TC_STAT_BEGIN;
                        cx.mma_wait_full(s, (it / C::kStages) & 1);
                        TC_STAT_END(2);
                    }
                    if (tc::elect_one()) cx.issue_stage(s, d_tmem, c == 0, c == nchunks - 1, acc);
                    __syncwarp();
                }
            }
            TC_STAT_FLUSH;
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        uint32_t tcount = 0;
        TC_STAT_DECL;
        for (int tile = group; tile < num_tiles; tile += ngroups, ++tcount) {
            int b = tile / tiles_per_img;
            const int t = tile % tiles_per_img;
            int i0 = (t / tiles_j) * kTileM + (int)cx.rank * C::kM, j0 = (t % tiles_j) * NT;
            // two-image tile: lane quarters 0, 1 hold image 2 t (accumulator columns 0..63), quarters 2, 3 image 2 t + 1 (64..127)
            int qrow = q, coff = 0, ncols = NT;
            bool img_ok = true;
            if (duo_B > 0) {
                b = 2 * tile + (q >> 1);
                img_ok = b < duo_B;
                b = min(b, duo_B - 1);
                i0 = 0, j0 = 0, qrow = q & 1, coff = (q >> 1) * 64, ncols = 64;
            }
            const int acc = tcount & 1;
            float tinv_t = 1.f;                      // kFuseLoss: the image's normaliser, fetched before the wait
            if constexpr (FUSE == kFuseLoss) tinv_t = fmaxf(__ldg(fz.tx + b), 1e-6f);
            {
                TC_STAT_BEGIN;
                tc::mbar_wait_sleep<HELIO_EPI_SLEEP_FWD_NS>(&cx.tfull[acc], (tcount >> 1) & 1);
                TC_STAT_END(1);
            }
            tc::tc_fence_after();
            const int i = img_ok ? i0 + qrow * 32 + lane : R;      // a missing second image: every row is out of range
            const size_t row_off = ((size_t)b * R + min(i, R - 1)) * R + j0;
            float* dst = img + row_off;
            // kFuseFeed: optional second copy of the image (e.g. the newest slot of a rollout's history buffer, batch stride given)
            float* dst2 = nullptr;
            if constexpr (FUSE == kFuseFeed) dst2 = fz.img2 ? fz.img2 + (size_t)b * fz.img2_bstride : nullptr;
            const uint32_t taddr = cx.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
            const bool vec = (R & 3) == 0;
            // kFuseMax: f0 = running max; kFuseLoss: {sum diff^2, sum |diff| dmaps, sum |diff|}; kFuseFeed: {sum w, sum w j, sum w i}
            float f0 = 0.f, f1 = 0.f, f2 = 0.f;
            // one image element (value p at row gi, column gj) for the sums that ride in the epilogue; tq / dq = target / dmaps there
            auto fuse_one = [&](float p, float tq, float dq, int gi, int gj) {
                if constexpr (FUSE == kFuseLoss) {
                    const float diff = p / tinv_t - tq / tinv_t;          // same arithmetic as loss_fwd_kernel
                    const float ae = fabsf(diff);
                    f0 = fmaf(diff, diff, f0);
                    f1 = fmaf(ae, dq, f1);
                    f2 += ae;
                }
                if constexpr (FUSE == kFuseFeed) {                        // CenterOfMass2D: w = max(x, 0), x = column, y = row
                    const float w = fmaxf(p, 0.f);
                    f0 += w;
                    f1 = fmaf(w, (float)gj, f1);
                    f2 = fmaf(w, (float)gi, f2);
                }
            };
#pragma unroll 1
            for (int cc = 0; cc < ncols; cc += 32) {
                const int cb = cc;                                   // column offset inside the image tile
                if (j0 + cb >= R) break;
                float v[32];
                tc::tmem_ld_32x32(taddr + coff + cc, v);
                if constexpr (PREC == 1) {               // f16x3: both operands carried a factor 2^14
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] *= 3.7252902984619140625e-09f;   // 2^-28, exact
                }
                const bool full = vec && j0 + cb + 32 <= R;      // warp-uniform
                if constexpr (STG == 1)
                if (full) {
                    // The accumulator arrives one image row per thread; storing it that way touches 32 cache lines per
                    // instruction.  Transpose the warp's 32 x 32 block through shared memory (XOR-swizzled, conflict-free
                    // both ways) so that every store instruction writes four full 128-byte row segments -- and the loads of
                    // target / dmaps for the fused loss sums are full row segments too.
                    const uint32_t stg = cx.epi_u + (uint32_t)q * 4096u;
                    __syncwarp();                                // the previous block has been read out
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        tc::sts_v4(stg + (uint32_t)lane * 128u + ((uint32_t)(c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    __syncwarp();
                    const int rrow = lane >> 3, ch = lane & 7;
                    const int gj = j0 + cb + 4 * ch;
                    float4 tq[8], dq[8];
                    if constexpr (FUSE == kFuseLoss) {           // all 16 loads in flight before the first use
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int gi = i0 + qrow * 32 + 4 * k + rrow;
                            const size_t o = ((size_t)b * R + min(gi, R - 1)) * R + gj;
                            tq[k] = __ldg(reinterpret_cast<const float4*>(fz.target + o));
                            dq[k] = __ldg(reinterpret_cast<const float4*>(fz.dmaps + o));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int row = 4 * k + rrow, gi = i0 + qrow * 32 + row;
                        const float4 x = tc::lds_v4_volatile(stg + (uint32_t)row * 128u + ((uint32_t)(ch ^ (row & 7)) << 4));
                        if (gi < R && img_ok) {
                            *reinterpret_cast<float4*>(img + ((size_t)b * R + gi) * R + gj) = x;
                            if constexpr (FUSE == kFuseFeed)
                                if (dst2) *reinterpret_cast<float4*>(dst2 + (size_t)gi * R + gj) = x;
                            if constexpr (FUSE == kFuseLoss) {
                                fuse_one(x.x, tq[k].x, dq[k].x, gi, gj), fuse_one(x.y, tq[k].y, dq[k].y, gi, gj + 1);
                                fuse_one(x.z, tq[k].z, dq[k].z, gi, gj + 2), fuse_one(x.w, tq[k].w, dq[k].w, gi, gj + 3);
                            }
                            if constexpr (FUSE == kFuseFeed) {
                                fuse_one(x.x, 0.f, 0.f, gi, gj), fuse_one(x.y, 0.f, 0.f, gi, gj + 1);
                                fuse_one(x.z, 0.f, 0.f, gi, gj + 2), fuse_one(x.w, 0.f, 0.f, gi, gj + 3);
                            }
                        }
                    }
                }
                if (i < R) {
                    if (full) {
                        if constexpr (STG == 0) {
#pragma unroll
                            for (int e = 0; e < 32; e += 4)
                                *reinterpret_cast<float4*>(dst + cb + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                            static_assert(FUSE == kFuseNone || FUSE == kFuseMax, "the loss / feed epilogues use the staged (coalesced) path");
                        }
                    } else {
                        // ragged edge (R % 32 != 0 or R % 4 != 0): one image row per thread, scalar
                        float* d2 = dst2 ? dst2 + (size_t)i * R + j0 : nullptr;
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (j0 + cb + e < R) {
                                dst[cb + e] = v[e];
                                if constexpr (FUSE == kFuseFeed) {
                                    if (d2) d2[cb + e] = v[e];
                                    fuse_one(v[e], 0.f, 0.f, i, j0 + cb + e);
                                }
                                if constexpr (FUSE == kFuseLoss)
                                    fuse_one(v[e], __ldg(fz.target + row_off + cb + e), __ldg(fz.dmaps + row_off + cb + e), i, j0 + cb + e);
                            }
                    }
                    if constexpr (FUSE == kFuseMax) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (full || j0 + cb + e < R) f0 = fmaxf(f0, v[e]);
                    }
                }
            }
            cx.epilogue_release(acc);
            if constexpr (FUSE == kFuseMax) {
                f0 = warp_max(f0);
                if (lane == 0 && img_ok) atomicMax(reinterpret_cast<int*>(fz.tile_max + b), __float_as_int(f0));
            }
            if constexpr (FUSE == kFuseLoss || FUSE == kFuseFeed) {
                f0 = warp_sum(f0), f1 = warp_sum(f1), f2 = warp_sum(f2);
                if (lane == 0) {
                    float* pp = fz.partials + ((((size_t)b * tiles_per_img + t) * CG + cx.rank) * 4 + q) * 3;
                    pp[0] = f0, pp[1] = f1, pp[2] = f2;
                }
            }
        }
        TC_STAT_FLUSH;
    }
    cx.teardown(0);
}

inline bool splat_tc_fwd_supported(int B, int N, int R) { return B > 0 && N > 0 && R >= 8 && R <= kTcMaxR; }
inline bool splat_tc_fwd_preferred(int B, int N, int R) {
    // measured on B200 over N in {50..5000}, R in {64..512}, B in {25..16384}: the tensor path wins (or ties within a
    // few microseconds) everywhere, including N = 50; only very small images leave most of a 128-row tile dead
    return R >= 48;
}

// launch `kernel` as a persistent grid of CTA groups (clusters of CG CTAs)
template <int CG, class Kernel, class... Args>
inline cudaError_t launch_tc_groups(Kernel kernel, long long num_tiles, int num_sms, int threads, int smem_bytes,
                                    cudaStream_t st, Args... args) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    long long groups = num_sms / CG;
    if (num_tiles < groups) groups = num_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * CG));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CG > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// tiles per image and partial records per image of the shape the forward picks for R (for the fused-loss buffers)
inline int splat_tc_fwd_cg(int R, int num_sms, int pair) { return (R > 128 && pair != 1 && num_sms >= 2) ? 2 : 1; }
inline int splat_tc_fwd_partials_per_image(int R, int num_sms, int pair) {
    const int cg = splat_tc_fwd_cg(R, num_sms, pair);
    const int nt = R > 128 ? 256 : (R > 64 ? 128 : 64);
    const int tiles = ((R + 128 * cg - 1) / (128 * cg)) * ((R + nt - 1) / nt);
    return tiles * cg * 4;
}

template <int NT, int CG, int PS, int PREC = 0, int STG = 0>
inline cudaError_t launch_splat_fwd_tc(const float* params, const int* counts, float* img, int B, int N, int R, float width,
                                       float height, int num_sms, cudaStream_t st, int fuse, const FwdFuse& fz, bool duo = false) {
    using C = SplatFwdTc<NT, CG, PS, PREC, STG>;
    // duo: two images of at most 64 pixels a side per 128 x 128 tile (see the kernel); plain max / no epilogue fusion only
    if (duo && !(NT == 128 && CG == 1 && PS == 1 && R <= 64 && (fuse == kFuseNone || fuse == kFuseMax))) return cudaErrorInvalidValue;
    const int tiles_i = duo ? 1 : (R + C::kM * CG - 1) / (C::kM * CG), tiles_j = duo ? 1 : (R + NT - 1) / NT;
    const long long num_tiles = duo ? ((long long)B + 1) / 2 : (long long)B * tiles_i * tiles_j;
    if (num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    auto go = [&](auto kernel) {
        return launch_tc_groups<CG>(kernel, num_tiles, num_sms, C::kThreads, C::kSmemBytes, st,
                                    reinterpret_cast<const float4*>(params), counts, img, N, R, make_axis(width, R),
                                    make_axis(height, R), tiles_i, tiles_j, (int)num_tiles, fz, duo ? B : 0);
    };
    if (fuse == kFuseMax) return go(splat_fwd_tc_kernel<NT, CG, PS, kFuseMax, PREC, STG>);
    if constexpr (STG == 1) {                        // the loss / feed epilogues work on the staged (coalesced) layout
        if (fuse == kFuseLoss) return go(splat_fwd_tc_kernel<NT, CG, PS, kFuseLoss, PREC, 1>);
        if (fuse == kFuseFeed) return go(splat_fwd_tc_kernel<NT, CG, PS, kFuseFeed, PREC, 1>);
    }
    if (fuse == kFuseLoss || fuse == kFuseFeed) return cudaErrorInvalidValue;
    return go(splat_fwd_tc_kernel<NT, CG, PS, kFuseNone, PREC, STG>);
}

inline int sun_avg_k(int N) { return N; }   // contraction length a tile sees (culling only shortens it)

// pair = 0: auto (CTA pairs for images taller than 128 rows), 1: single CTA, 2: CTA pairs
// split = producer warps per 32-row operand slab (1 or 2; 0 = auto)
// counts (may be NULL): per-sun number of valid entries of a culled parameter row (cull.cuh)
inline cudaError_t splat_tc_fwd(const float* params, float* img, int B, int N, int R, float width, float height, int num_sms,
                                cudaStream_t st, int pair = 0, int split = 0, int fuse = kFuseNone, const FwdFuse& fz = FwdFuse{},
                                const int* counts = nullptr, int prec = 0) {
#define HELIO_FWD(NT_, CG_, PS_) launch_splat_fwd_tc<NT_, CG_, PS_>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz)
    // staged epilogue (see splat_fwd_tc_kernel): when the epilogue is a large share of a tile
    if (fuse == kFuseLoss || fuse == kFuseFeed) split = 1;      // those epilogues exist for the staged stores only
    const bool stg = fuse == kFuseLoss || fuse == kFuseFeed || (split != 2 && (prec == 1 || sun_avg_k(N) < 1024));
    // images of at most 64 pixels a side: two per 128 x 128 tile (all eight producer warps live), see the kernel
    const bool duo = HELIO_FWD_DUO && R <= 64 && B > 1 && split != 2 && (fuse == kFuseNone || fuse == kFuseMax);
    if (prec == 1) {                                 // opt-in f16x3 operands (see SplatFwdTc)
#define HELIO_FWD16(NT_, CG_, PS_, STG_) launch_splat_fwd_tc<NT_, CG_, PS_, 1, STG_>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz)
        if (R > 128) {
            if (splat_tc_fwd_cg(R, num_sms, pair) == 2) return split == 2 ? HELIO_FWD16(256, 2, 2, 0) : (stg ? HELIO_FWD16(256, 2, 1, 1) : HELIO_FWD16(256, 2, 1, 0));
            return stg ? HELIO_FWD16(256, 1, 1, 1) : HELIO_FWD16(256, 1, 1, 0);
        }
        if (R > 64) return split == 2 ? HELIO_FWD16(128, 1, 2, 0) : (stg ? HELIO_FWD16(128, 1, 1, 1) : HELIO_FWD16(128, 1, 1, 0));
        if (duo) return stg ? launch_splat_fwd_tc<128, 1, 1, 1, 1>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz, true)
                            : launch_splat_fwd_tc<128, 1, 1, 1, 0>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz, true);
        return stg ? HELIO_FWD16(64, 1, 1, 1) : HELIO_FWD16(64, 1, 1, 0);
#undef HELIO_FWD16
    }
#define HELIO_FWDS(NT_, CG_) launch_splat_fwd_tc<NT_, CG_, 1, 0, 1>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz)
    if (stg) {
        if (R > 128) return splat_tc_fwd_cg(R, num_sms, pair) == 2 ? HELIO_FWDS(256, 2) : HELIO_FWDS(256, 1);
        if (duo) return launch_splat_fwd_tc<128, 1, 1, 0, 1>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz, true);
        return R > 64 ? HELIO_FWDS(128, 1) : HELIO_FWDS(64, 1);
    }
#undef HELIO_FWDS
    if (R > 128) {
        if (splat_tc_fwd_cg(R, num_sms, pair) == 2) return split == 2 ? HELIO_FWD(256, 2, 2) : HELIO_FWD(256, 2, 1);
        return HELIO_FWD(256, 1, 1);
    }
    if (R > 64) return split == 2 ? HELIO_FWD(128, 1, 2) : HELIO_FWD(128, 1, 1);
    // measured on B200: twice the producer warps (split = 2) is 6-25 % slower at every shape, whether the extra warps
    // split the rows or the K range of a stage; kept as an A/B switch (HELIO_TC_FWD_SPLIT=2) only
    if (duo) return launch_splat_fwd_tc<128, 1, 1, 0, 0>(params, counts, img, B, N, R, width, height, num_sms, st, fuse, fz, true);
    return split == 2 ? HELIO_FWD(64, 1, 2) : HELIO_FWD(64, 1, 1);
#undef HELIO_FWD
}

// ================================================================================================
// backward
// ================================================================================================
// Tile = (sun b, block of 128*CG heliostats, 128 per CTA).  Per tile two products run back to back
// through the same pipeline, each split into ceil(R/NT) accumulators of 128 x NT per CTA:
//   product 0 (T): A rows = Gy[n, j-chunk]       B rows = g[i, j-chunk]   (i = accumulator column)
//   product 1 (U): A rows = amp Gx[n, i-chunk]   B rows = g[i-chunk, j]^T (j = accumulator column)
// The epilogue thread that owns TMEM lane n keeps {S0,Sx,Sxx} from product 0 in registers, adds
// {Sy,Syy} from product 1 and writes one float4 per heliostat.
//
// PREC = 1 ("f16x3", K = 64 per stage): the Gaussian operand lies in [0, 1] and is scaled by 2^14 as in the forward; the
// image gradient is scaled PER IMAGE by the power of two that brings max |g[b]| just under 2^14 (gmax[b], produced by the
// loss backward); both split into two fp16 pieces v = p1 + p2 (11 + 11 significant bits for every value within 2^-17 of the
// image's maximum, absolute error below 2^-38 of that maximum for smaller ones).  An operand row of 64 K values is 128
// bytes per piece -- the bytes a row of 32 tf32 values takes -- so a stage keeps its 64 KB, the ring its three stages and
// the MMA issue its descriptors, but a stage now covers TWICE the contraction depth with the same 12 instructions
// (kind::f16, K = 16): half the tensor work per eval, half the stage hand-overs, and the stagers' tile-load latency is
// amortised over twice the data.  The epilogue unscales the moments by the exact power of two.
#ifndef HELIO_BWD_PACKED2
// 1: packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2: the same IEEE operations, two results per instruction) in the backward
// epilogue and in the f16x3 producers.  With 21 warps per SM the f16x3 backward is bound by instruction issue across its
// roles (scripts/tc_stats.py), so halving the FP32 instruction count pays there (it did not in the latency-bound forward).
#define HELIO_BWD_PACKED2 1
#endif
#ifndef HELIO_BWD_STAGER_GROUPS_F16
#define HELIO_BWD_STAGER_GROUPS_F16 2    // f16x3 backward: two groups of gradient stagers take the stages in turn
#endif
#ifndef HELIO_BWD_STAGER_GROUPS
#define HELIO_BWD_STAGER_GROUPS 1        // 3xTF32 backward
#endif
#ifndef HELIO_BWD_EPI_SPLIT_F16
#define HELIO_BWD_EPI_SPLIT_F16 2        // f16x3 backward: two epilogue warps per TMEM lane quarter (half of the columns each)
#endif
template <int NT, int CG, int PREC = 0>
using SplatBwdTc = SplatTcLayout<NT, CG, NT / CG, 2, 1, 2, (PREC && HELIO_BWD_EPI_SPLIT_F16 > 1) ? 4096 : 0, PREC ? 64 : 32, PREC,
                                 // (the single-CTA 256-column variant already has eight stager warps: no second group, 25 warps)
                                 PREC ? ((NT == 256 && CG == 1) ? 1 : HELIO_BWD_STAGER_GROUPS_F16) : HELIO_BWD_STAGER_GROUPS,
                                 PREC ? HELIO_BWD_EPI_SPLIT_F16 : 1>;

// exponent e with |v| < 2^(e+1) (v finite, > 0), else 0; and the power of two 2^k as a float
__device__ __forceinline__ int float_exponent(float v) { return v > 0.f ? (int)((__float_as_uint(v) >> 23) & 0xFFu) - 127 : 0; }
__device__ __forceinline__ float pow2i(int k) { return __uint_as_float((uint32_t)(min(max(k, -126), 127) + 127) << 23); }
// per-image scale exponent of the gradient operand: |g| * 2^sexp < 2^14 (fp16 tops out at 65504)
__device__ __forceinline__ int grad_scale_exponent(float gmax_b) {
    return gmax_b > 0.f && gmax_b <= 3.0e38f ? min(max(13 - float_exponent(gmax_b), -100), 100) : 0;
}

template <int NT, int CG, int PREC>
__global__ void __launch_bounds__(SplatBwdTc<NT, CG, PREC>::kThreads, 1)
splat_bwd_tc_kernel(const float4* __restrict__ params, const int* __restrict__ counts, const int* __restrict__ index,
                    const float* __restrict__ g_img, const float* __restrict__ gmax, float4* __restrict__ moments, int N, int R,
                    Axis ax, Axis ay, int nblocks, int num_tiles) {
    using C = SplatBwdTc<NT, CG, PREC>;
    static_assert(PREC == 0 || HELIO_BWD_DEFER == 0, "the f16x3 backward is written for the plain hand-off");
    extern __shared__ uint8_t smem_raw[];
    SplatTcCtx<C, CG> cx;
    cx.setup(smem_raw, R, ax, ay);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = (R + C::kKC - 1) / C::kKC;   // K stages per accumulator
    const int pblocks = (R + NT - 1) / NT;           // accumulators per product
    const bool vec = (R & 3) == 0;
    const int group = blockIdx.x / CG, ngroups = gridDim.x / CG;
    constexpr int kTileH = C::kM * CG;               // heliostats per tile
    // culled input (cull.cuh): params rows are compacted, counts[b] entries are valid and index maps them back; a tile
    // past the end of its sun's list is skipped by every role alike (the caller zero-fills the moments)
    auto sun_count = [&](int b) { return counts ? __ldg(counts + b) : N; };
    auto tile_empty = [&](int tile) { return (tile % nblocks) * kTileH >= sun_count(tile / nblocks); };

    if (warp < C::kAWarps) {
        // ================= Gaussian operand: thread = heliostat row, warp = (32-row slab, K slice) =================
        // kASplit warps share a slab and each generates kKC / kASplit of the stage's K range, which halves the
        // per-stage latency of a producer warp (the critical path of this kernel).  K padding (k >= R) needs no
        // zeroing here: the matching rows of the gradient operand are exact zeros and these values are finite.
        constexpr int kQ = 8 / C::kASplit;               // 16-byte chunks of the 128-byte row per warp
        const int r = (warp % (C::kM / 32)) * 32 + lane;
        const int q0 = (warp / (C::kM / 32)) * kQ;
        uint32_t it = 0;
        // this row's footprint for a tile; the next tile's is fetched while the current one is generated
        auto tile_params = [&](int tile, bool& live) {
            const int b = tile / nblocks, nb = tile % nblocks;
            const int n = nb * kTileH + (int)cx.rank * C::kM + r;
            live = tile < num_tiles && n < sun_count(b);
            return live ? __ldg(params + (size_t)b * N + n) : make_float4(0.f, 0.f, 0.f, 1.f);
        };
        bool live_next;
        float4 p_next = tile_params(group, live_next);
        TC_STAT_DECL;
#if HELIO_BWD_DEFER
        int pending = -1;
#endif
        for (int tile = group; tile < num_tiles; tile += ngroups) {
            const bool live = live_next;
            const float4 p = p_next;
            p_next = tile_params(tile + ngroups, live_next);
            if (tile_empty(tile)) continue;
            const float nk2 = -p.z;
            const float dead = live ? 0.f : -INFINITY;   // rows beyond N: 2^-inf = 0
#pragma unroll 1
            for (int prod = 0; prod < 2; ++prod) {
                const float ctr = prod == 0 ? p.y : p.x;
                const float la = dead + (prod == 0 ? 0.f : lg2_ftz(p.w)) + (PREC == 1 ? 14.f : 0.f);   // amplitude (and the f16x3 scale 2^14) folded into the exponent
                const uint32_t tab = (prod == 0 ? cx.sY_u : cx.sX_u) + (uint32_t)q0 * (PREC == 1 ? 32u : 16u);
#pragma unroll 1
                for (int pbk = 0; pbk < pblocks; ++pbk) {
#pragma unroll 1
                    for (int c = 0; c < kchunks; ++c, ++it) {
                        const int s = it % C::kStages;
                        const int k0 = c * C::kKC;
#if HELIO_BWD_DEFER
                        // evaluate first, hand the PREVIOUS stage over (its stores have drained meanwhile), then store
                        float v[kQ][4];
#pragma unroll
                        for (int q = 0; q < kQ; ++q) {
                            const float4 xs = tc::lds_v4(tab + (uint32_t)(k0 + 4 * q) * 4u);
                            const float x[4] = {xs.x, xs.y, xs.z, xs.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float d = x[e] - ctr;
                                v[q][e] = ex2(fmaf(d * nk2, d, la));
                            }
                        }
                        if (pending >= 0) cx.producer_commit(pending);
                        cx.producer_acquire(s, (it / C::kStages) & 1);
                        const uint32_t hi_base = cx.smem_u + (uint32_t)(s * C::kStageBytes);
                        const uint32_t lo_base = hi_base + C::kABytes;
#pragma unroll
                        for (int q = 0; q < kQ; ++q) {
                            float hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) tc::split_tf32(v[q][e], hi[e], lo[e]);
                            const uint32_t off = tc::sw128_offset((uint32_t)r, (uint32_t)(q0 + q));
                            tc::sts_v4(hi_base + off, hi[0], hi[1], hi[2], hi[3]);
                            tc::sts_v4(lo_base + off, lo[0], lo[1], lo[2], lo[3]);
                        }
                        pending = s;
#else
                        {
                            TC_STAT_BEGIN;
                            cx.producer_acquire(s, (it / C::kStages) & 1);
                            TC_STAT_END(2);
                        }
                        const uint32_t hi_base = cx.smem_u + (uint32_t)(s * C::kStageBytes);
                        const uint32_t lo_base = hi_base + C::kABytes;
                        if constexpr (PREC == 1) {
                            // 16-byte chunk (q0 + q) of the row = 8 fp16 = K values k0 + 8 (q0 + q) .. + 7
#pragma unroll
                            for (int q = 0; q < kQ; ++q) {
                                const float4 xa = tc::lds_v4(tab + (uint32_t)(k0 + 8 * q) * 4u), xb = tc::lds_v4(tab + (uint32_t)(k0 + 8 * q + 4) * 4u);
                                const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                                uint32_t p1[4], p2[4];
#if HELIO_BWD_PACKED2
                                const tc::f32x2 nc2 = tc::pack2(-ctr, -ctr), k22 = tc::pack2(nk2, nk2), la2 = tc::pack2(la, la);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const tc::f32x2 d2 = tc::add2(tc::pack2(x[2 * e], x[2 * e + 1]), nc2);
                                    float a0, a1;
                                    tc::unpack2(tc::fma2(tc::mul2(d2, k22), d2, la2), a0, a1);
                                    tc::split_f16x2_packed(tc::pack2(ex2(a0), ex2(a1)).r, p1[e], p2[e]);
                                }
#else
                                float v[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float d = x[e] - ctr;
                                    v[e] = ex2(fmaf(d * nk2, d, la));
                                }
#pragma unroll
                                for (int e = 0; e < 4; ++e) tc::split_f16x2(v[2 * e], v[2 * e + 1], p1[e], p2[e]);
#endif
                                const uint32_t off = tc::sw128_offset((uint32_t)r, (uint32_t)(q0 + q));
                                tc::sts_v4_b32(hi_base + off, p1[0], p1[1], p1[2], p1[3]);
                                tc::sts_v4_b32(lo_base + off, p2[0], p2[1], p2[2], p2[3]);
                            }
                        } else {
#pragma unroll
                        for (int q = 0; q < kQ; ++q) {
                            const float4 xs = tc::lds_v4(tab + (uint32_t)(k0 + 4 * q) * 4u);
                            const float x[4] = {xs.x, xs.y, xs.z, xs.w};
                            float hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float d = x[e] - ctr;
                                const float v = ex2(fmaf(d * nk2, d, la));
                                tc::split_tf32(v, hi[e], lo[e]);
                            }
                            const uint32_t off = tc::sw128_offset((uint32_t)r, (uint32_t)(q0 + q));
                            tc::sts_v4(hi_base + off, hi[0], hi[1], hi[2], hi[3]);
                            tc::sts_v4(lo_base + off, lo[0], lo[1], lo[2], lo[3]);
                        }
                        }
                        {
                            TC_STAT_BEGIN;
                            cx.producer_commit(s);
                            TC_STAT_END(1);
                        }
#endif
                    }
                }
            }
        }
#if HELIO_BWD_DEFER
        if (pending >= 0) cx.producer_commit(pending);
#endif
        TC_STAT_FLUSH;
    } else if (warp < C::kMmaWarp) {
        // ================= gradient tile stagers =================
        // kBAlt groups of kBRows / 32 warps; group sg fills the stages with it % kBAlt == sg
        const int sw_ = (int)(threadIdx.x >> 5) - C::kAWarps;              // stager warp index
        const int sg = sw_ / (C::kBRows / 32);                             // group
        const int gw = sw_ % (C::kBRows / 32);                             // 32-row slab inside the group
        const int t = gw * 32 + lane;                    // 0..kBRows-1: operand row inside this CTA's share
        const int row_base = (int)cx.rank * C::kBRows;   // first accumulator column this CTA stages
        uint32_t it = 0;
        TC_STAT_DECL;
#if HELIO_BWD_DEFER
        int spending = -1;
#endif
        for (int tile = group; tile < num_tiles; tile += ngroups) {
            if (tile_empty(tile)) continue;
            const int b = tile / nblocks;
            const float* gb = g_img + (size_t)b * R * R;
            float gs = 1.f;                                  // f16x3: per-image power-of-two scale of the gradient operand
            if constexpr (PREC == 1) gs = pow2i(grad_scale_exponent(__ldg(gmax + b)));
#pragma unroll 1
            for (int prod = 0; prod < 2; ++prod) {
#pragma unroll 1
                for (int pbk = 0; pbk < pblocks; ++pbk) {
#pragma unroll 1
                    for (int c = 0; c < kchunks; ++c, ++it) {
                        if (C::kBAlt > 1 && (int)(it % C::kBAlt) != sg) continue;  // another group's turn (the loop still counts the stage)
                        const int s = it % C::kStages;
                        const int k0 = c * C::kKC;
                        float4 vals[8];
                        // product 0: 8 lanes per 128-byte row segment, a warp instruction covers 4 rows: rows {0, 4, 1, 5} / {2, 6, 3, 7}
                        // of an 8-row swizzle atom, which keeps the 8-byte f16x3 stores free of bank conflicts (see the forward
                        // producers), or consecutive rows with HELIO_BWD_STAGER_ROWMAP = 0
                        const int ch = lane & 7, rl = lane >> 3;
#if HELIO_BWD_STAGER_ROWMAP
                        const int row_l = gw * 32 + 4 * (rl & 1) + (rl >> 1);
                        auto qrow = [](int q) { return 8 * (q >> 1) + 2 * (q & 1); };
#else
                        const int row_l = gw * 32 + rl;
                        auto qrow = [](int q) { return 4 * q; };
#endif
                        auto srow = [&](int q) { return row_l + qrow(q); };
                        const int j = pbk * NT + row_base + t;                     // product 1: this thread's image column
                        // 32 K values (kb .. kb + 31) of this thread's share of the operand tile -> vals[8]
                        //   product 0: operand row = image row i (accumulator column), K = image column j: 8 lanes cover one
                        //              128-byte row segment, a warp instruction covers 4 rows; vals[q] = row srow(q), K kb + 4 ch .. + 3
                        //   product 1: operand row = image column j, K = image row i: lanes read consecutive columns of one
                        //              image row (coalesced), transposing in registers; vals[q] = K kb + 4 q .. + 3 of row t
                        auto load32 = [&](const int kb) {
                            // whole operand tile inside the image: no per-element bounds checks (the common case)
                            const bool inside = vec && kb + 32 <= R && pbk * NT + row_base + C::kBRows <= R;
                            if (prod == 0) {
                                if (inside && R == NT) {
                                    // R = 64 / 128 / 256 exactly (the usual resolutions): the row stride is a compile-time constant and
                                    // the loads of a stage share one address register with immediate offsets
                                    const float4* src = reinterpret_cast<const float4*>(gb + (size_t)(row_base + row_l) * NT + kb) + ch;
#pragma unroll
                                    for (int q = 0; q < 8; ++q) vals[q] = __ldg(src + qrow(q) * (NT / 4));
                                } else if (inside) {
                                    const float4* src = reinterpret_cast<const float4*>(gb + (size_t)(pbk * NT + row_base + row_l) * R + kb) + ch;
                                    const size_t R4 = (size_t)(R >> 2);                                 // one image row in float4
#pragma unroll
                                    for (int q = 0; q < 8; ++q) vals[q] = __ldg(src + (size_t)qrow(q) * R4);
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        const int i = pbk * NT + row_base + srow(q), jj = kb + ch * 4;
                                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                                        if (i < R) {
                                            const float* src = gb + (size_t)i * R + jj;
                                            if (vec && jj + 3 < R) {
                                                v = __ldg(reinterpret_cast<const float4*>(src));
                                            } else {
                                                if (jj < R) v.x = __ldg(src);
                                                if (jj + 1 < R) v.y = __ldg(src + 1);
                                                if (jj + 2 < R) v.z = __ldg(src + 2);
                                                if (jj + 3 < R) v.w = __ldg(src + 3);
                                            }
                                        }
                                        vals[q] = v;
                                    }
                                }
                            } else {
                                if (inside && R == NT) {
                                    const float* src = gb + (size_t)kb * NT + j;
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        vals[q].x = __ldg(src + (4 * q) * NT);
                                        vals[q].y = __ldg(src + (4 * q + 1) * NT);
                                        vals[q].z = __ldg(src + (4 * q + 2) * NT);
                                        vals[q].w = __ldg(src + (4 * q + 3) * NT);
                                    }
                                } else if (inside) {
                                    const float* src = gb + (size_t)kb * R + j;
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        vals[q].x = __ldg(src + (size_t)(4 * q) * R);
                                        vals[q].y = __ldg(src + (size_t)(4 * q + 1) * R);
                                        vals[q].z = __ldg(src + (size_t)(4 * q + 2) * R);
                                        vals[q].w = __ldg(src + (size_t)(4 * q + 3) * R);
                                    }
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        float x[4];
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const int i = kb + 4 * q + e;
                                            x[e] = (i < R && j < R) ? __ldg(gb + (size_t)i * R + j) : 0.f;
                                        }
                                        vals[q] = make_float4(x[0], x[1], x[2], x[3]);
                                    }
                                }
                            }
                        };
                        const uint32_t hi_base = cx.smem_u + (uint32_t)(s * C::kStageBytes + 2 * C::kABytes);
                        const uint32_t lo_base = hi_base + C::kBBytes;
                        if constexpr (PREC == 1) {
                            // two halves of 32 K values; rows of the fp16 tiles hold 64 K values (128 bytes) per piece
#pragma unroll 1
                            for (int h = 0; h < 2; ++h) {
                                load32(k0 + 32 * h);
                                if (h == 0) {
                                    TC_STAT_BEGIN;
                                    cx.producer_acquire(s, (it / C::kStages) & 1);
                                    TC_STAT_END(2);
                                }
                                if (prod == 0) {
                                    // this lane's 4 K values = bytes 64 h + 8 ch .. + 7 of the row: half of 16-byte chunk 4 h + ch / 2
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        uint32_t p1a, p2a, p1b, p2b;
#if HELIO_BWD_PACKED2
                                        const tc::f32x2 gs2 = tc::pack2(gs, gs);
                                        tc::split_f16x2_packed(tc::mul2(tc::pack2(vals[q].x, vals[q].y), gs2).r, p1a, p2a);
                                        tc::split_f16x2_packed(tc::mul2(tc::pack2(vals[q].z, vals[q].w), gs2).r, p1b, p2b);
#else
                                        tc::split_f16x2(vals[q].x * gs, vals[q].y * gs, p1a, p2a);
                                        tc::split_f16x2(vals[q].z * gs, vals[q].w * gs, p1b, p2b);
#endif
                                        const uint32_t row = (uint32_t)srow(q), sw = row & 7u;
                                        const uint32_t off = (row >> 3) * 1024u + sw * 128u + ((((uint32_t)(4 * h) + ((uint32_t)ch >> 1)) ^ sw) << 4) + 8u * ((uint32_t)ch & 1u);
                                        tc::sts_v2_b32(hi_base + off, p1a, p1b);
                                        tc::sts_v2_b32(lo_base + off, p2a, p2b);
                                    }
                                } else {
                                    // vals[2 m], vals[2 m + 1] = K values 32 h + 8 m .. + 7 of row t: 16-byte chunk 4 h + m
#pragma unroll
                                    for (int m = 0; m < 4; ++m) {
                                        uint32_t p1[4], p2[4];
#if HELIO_BWD_PACKED2
                                        const tc::f32x2 gs2 = tc::pack2(gs, gs);
                                        tc::split_f16x2_packed(tc::mul2(tc::pack2(vals[2 * m].x, vals[2 * m].y), gs2).r, p1[0], p2[0]);
                                        tc::split_f16x2_packed(tc::mul2(tc::pack2(vals[2 * m].z, vals[2 * m].w), gs2).r, p1[1], p2[1]);
                                        tc::split_f16x2_packed(tc::mul2(tc::pack2(vals[2 * m + 1].x, vals[2 * m + 1].y), gs2).r, p1[2], p2[2]);
                                        tc::split_f16x2_packed(tc::mul2(tc::pack2(vals[2 * m + 1].z, vals[2 * m + 1].w), gs2).r, p1[3], p2[3]);
#else
                                        tc::split_f16x2(vals[2 * m].x * gs, vals[2 * m].y * gs, p1[0], p2[0]);
                                        tc::split_f16x2(vals[2 * m].z * gs, vals[2 * m].w * gs, p1[1], p2[1]);
                                        tc::split_f16x2(vals[2 * m + 1].x * gs, vals[2 * m + 1].y * gs, p1[2], p2[2]);
                                        tc::split_f16x2(vals[2 * m + 1].z * gs, vals[2 * m + 1].w * gs, p1[3], p2[3]);
#endif
                                        const uint32_t off = tc::sw128_offset((uint32_t)t, (uint32_t)(4 * h + m));
                                        tc::sts_v4_b32(hi_base + off, p1[0], p1[1], p1[2], p1[3]);
                                        tc::sts_v4_b32(lo_base + off, p2[0], p2[1], p2[2], p2[3]);
                                    }
                                }
                            }
                            {
                                TC_STAT_BEGIN;
                                cx.producer_commit(s);
                                TC_STAT_END(1);
                            }
                        } else {
                        load32(k0);
                        uint32_t offs[8];
                        if (prod == 0) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) offs[q] = tc::sw128_offset((uint32_t)srow(q), (uint32_t)ch);
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) offs[q] = tc::sw128_offset((uint32_t)t, (uint32_t)q);
                        }
#if HELIO_BWD_DEFER
                        if (spending >= 0) cx.producer_commit(spending);   // previous stage: drained under the loads above
#endif
                        {
                            TC_STAT_BEGIN;
                            cx.producer_acquire(s, (it / C::kStages) & 1);
                            TC_STAT_END(2);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 h, l;
                            tc::split_tf32(vals[q].x, h.x, l.x);
                            tc::split_tf32(vals[q].y, h.y, l.y);
                            tc::split_tf32(vals[q].z, h.z, l.z);
                            tc::split_tf32(vals[q].w, h.w, l.w);
                            tc::sts_v4(hi_base + offs[q], h.x, h.y, h.z, h.w);
                            tc::sts_v4(lo_base + offs[q], l.x, l.y, l.z, l.w);
                        }
#if HELIO_BWD_DEFER
                        spending = s;
#else
                        {
                            TC_STAT_BEGIN;
                            cx.producer_commit(s);
                            TC_STAT_END(1);
                        }
#endif
                        }
                    }
                }
            }
        }
#if HELIO_BWD_DEFER
        if (spending >= 0) cx.producer_commit(spending);
#endif
        TC_STAT_FLUSH;
    } else if (warp == C::kMmaWarp) {
        // ================= MMA issuer (leader CTA of the group) =================
        if (cx.rank == 0) {
            TC_STAT_DECL;
            uint32_t it = 0, sub = 0;
            const int subs_per_tile = 2 * pblocks;
            for (int tile = group; tile < num_tiles; tile += ngroups) {
                if (tile_empty(tile)) continue;
                for (int sb = 0; sb < subs_per_tile; ++sb, ++sub) {
                    const int acc = sub & 1;
                    {
                        TC_STAT_BEGIN;
                        cx.mma_wait_tempty(acc, (sub >> 1) & 1);
                        TC_STAT_END(1);
                    }
                    const uint32_t d_tmem = cx.tmem_base + (uint32_t)(acc * NT);
                    for (int c = 0; c < kchunks; ++c, ++it) {
                        const int s = it % C::kStages;
                        {
                            TC_STAT_BEGIN;
                            cx.mma_wait_full(s, (it / C::kStages) & 1);
                            TC_STAT_END(2);
                        }
                        if (tc::elect_one()) cx.issue_stage(s, d_tmem, c == 0, c == kchunks - 1, acc);
                        __syncwarp();
                    }
                }
            }
            TC_STAT_FLUSH;
        }
    } else {
        // ================= epilogue: accumulator -> moments =================
        // kEpiSplit warps share a TMEM lane quarter, each folding 1 / kEpiSplit of an accumulator's columns; their partial
        // moments are combined through shared memory at the end of the tile (slot by tile parity, one named barrier)
        const int q = warp & 3;
        const int eh = (warp - C::kEpiWarp0) >> 2;            // which share of the columns
        constexpr int kColsPer = NT / C::kEpiSplit;
        uint32_t sub = 0, tdone = 0;
        auto tile_params = [&](int tile, bool& live) {
            const int b = tile / nblocks, nb = tile % nblocks;
            const int n = nb * kTileH + (int)cx.rank * C::kM + q * 32 + lane;
            live = tile < num_tiles && n < sun_count(b);
            return live ? __ldg(params + (size_t)b * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        bool live_next;
        float4 p_next = tile_params(group, live_next);
        TC_STAT_DECL;
        for (int tile = group; tile < num_tiles; tile += ngroups) {
            const int b = tile / nblocks, nb = tile % nblocks;
            const int n = nb * kTileH + (int)cx.rank * C::kM + q * 32 + lane;
            const bool live = live_next;
            const float4 p = p_next;
            p_next = tile_params(tile + ngroups, live_next);
            if (tile_empty(tile)) continue;
            const float nk2 = -p.z;
            float S0 = 0.f, Sx = 0.f, Sy = 0.f, S2 = 0.f;
#pragma unroll 1
            for (int prod = 0; prod < 2; ++prod) {
                const float ctr = prod == 0 ? p.x : p.y;
                const uint32_t tab = prod == 0 ? cx.sX_u : cx.sY_u;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#if HELIO_BWD_PACKED2
                tc::f32x2 s0p = tc::pack2(0.f, 0.f), s1p = s0p, s2p = s0p;      // even / odd columns
                const tc::f32x2 nc2 = tc::pack2(-ctr, -ctr), k22 = tc::pack2(nk2, nk2);
#endif
#pragma unroll 1
                for (int pbk = 0; pbk < pblocks; ++pbk, ++sub) {
                    const int acc = sub & 1;
                    {
                        TC_STAT_BEGIN;
                        tc::mbar_wait_sleep<HELIO_EPI_SLEEP_BWD_NS>(&cx.tfull[acc], (sub >> 1) & 1);
                        TC_STAT_END(1);
                    }
                    tc::tc_fence_after();
                    const uint32_t taddr = cx.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
                    const int col0 = pbk * NT;
#pragma unroll 1
                    for (int cb = eh * kColsPer; cb < (eh + 1) * kColsPer; cb += 32) {
                        if (col0 + cb >= R) break;
                        float v[32];
                        tc::tmem_ld_32x32(taddr + cb, v);
#pragma unroll
                        for (int e4 = 0; e4 < 32; e4 += 4) {
                            const float4 xs = tc::lds_v4(tab + (uint32_t)(col0 + cb + e4) * 4u);
                            const float x[4] = {xs.x, xs.y, xs.z, xs.w};
#if HELIO_BWD_PACKED2
#pragma unroll
                            for (int e = 0; e < 4; e += 2) {
                                // columns >= R hold exact zeros (their operand rows are zero)
                                const tc::f32x2 d2 = tc::add2(tc::pack2(x[e], x[e + 1]), nc2);
                                const tc::f32x2 dd2 = tc::mul2(d2, d2);
                                float a0, a1;
                                tc::unpack2(tc::mul2(dd2, k22), a0, a1);
                                const tc::f32x2 w2 = tc::mul2(tc::pack2(ex2(a0), ex2(a1)), tc::pack2(v[e4 + e], v[e4 + e + 1]));
                                s0p = tc::add2(s0p, w2);
                                s1p = tc::fma2(w2, d2, s1p);
                                s2p = tc::fma2(w2, dd2, s2p);
                            }
#else
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // columns >= R hold exact zeros (their operand rows are zero)
                                const float d = x[e] - ctr;
                                const float dd = d * d;
                                const float w = ex2(dd * nk2) * v[e4 + e];
                                s0 += w;
                                s1 = fmaf(w, d, s1);
                                s2 = fmaf(w, dd, s2);
                            }
#endif
                        }
                    }
                    cx.epilogue_release(acc);
                }
#if HELIO_BWD_PACKED2
                {
                    float lo, hi;
                    tc::unpack2(s0p, lo, hi), s0 = lo + hi;
                    tc::unpack2(s1p, lo, hi), s1 = lo + hi;
                    tc::unpack2(s2p, lo, hi), s2 = lo + hi;
                }
#endif
                if (prod == 0) {
                    S0 = s0 * p.w, Sx = s1 * p.w, S2 = s2 * p.w;
                } else {
                    Sy = s1, S2 += s2;
                }
            }
            if constexpr (C::kEpiSplit > 1) {
                static_assert(C::kEpiSplit == 2, "two-way split");
                const uint32_t slot = cx.epi_u + (uint32_t)((tdone & 1u) * 2048u + (uint32_t)(q * 32 + lane) * 16u);
                if (eh == 1) tc::sts_v4(slot, S0, Sx, Sy, S2);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");      // the two warps of this lane quarter
                ++tdone;
                if (eh == 1) continue;                        // the first warp of the quarter owns the result
                const float4 o = tc::lds_v4_volatile(slot);
                S0 += o.x, Sx += o.y, Sy += o.z, S2 += o.w;
            }
            if constexpr (PREC == 1) {                       // undo 2^14 (Gaussian operand) x 2^sexp (gradient operand): exact
                const float us = pow2i(-14 - grad_scale_exponent(__ldg(gmax + b)));
                S0 *= us, Sx *= us, Sy *= us, S2 *= us;
            }
            if (live) moments[(size_t)b * N + (index ? __ldg(index + (size_t)b * N + n) : n)] = make_float4(S0, Sx, Sy, S2);
        }
        TC_STAT_FLUSH;
    }
    cx.teardown(1);
}

inline bool splat_tc_bwd_supported(int B, int N, int R) { return B > 0 && N > 0 && R >= 8 && R <= kTcMaxR; }
inline bool splat_tc_bwd_preferred(int B, int N, int R) { return R >= 48; }

template <int NT, int CG, int PREC = 0>
inline cudaError_t launch_splat_bwd_tc(const float* params, const int* counts, const int* index, const float* g_img, const float* gmax,
                                       float* moments, int B, int N, int R, float width, float height, int num_sms, cudaStream_t st) {
    using C = SplatBwdTc<NT, CG, PREC>;
    const int nblocks = (N + C::kM * CG - 1) / (C::kM * CG);
    const long long num_tiles = (long long)B * nblocks;
    if (num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    return launch_tc_groups<CG>(splat_bwd_tc_kernel<NT, CG, PREC>, num_tiles, num_sms, C::kThreads, C::kSmemBytes, st,
                                reinterpret_cast<const float4*>(params), counts, index, g_img, gmax, reinterpret_cast<float4*>(moments),
                                N, R, make_axis(width, R), make_axis(height, R), nblocks, (int)num_tiles);
}

// counts / index (may be NULL): culled (compacted) parameter rows and their original heliostat indices (cull.cuh); the
// caller zero-fills `moments` first, only the kept heliostats are written
// gmax (may be NULL = 3xTF32): per-image max |g_img[b]|; when given, the f16x3 operand format (K = 64 per stage) is used
inline cudaError_t splat_tc_bwd(const float* params, const float* g_img, float* moments, int B, int N, int R, float width,
                                float height, int num_sms, cudaStream_t st, int pair = 0, const int* counts = nullptr,
                                const int* index = nullptr, const float* gmax = nullptr) {
    if (gmax != nullptr) {
#define HELIO_BWD16(NT_, CG_) launch_splat_bwd_tc<NT_, CG_, 1>(params, counts, index, g_img, gmax, moments, B, N, R, width, height, num_sms, st)
        if (R > 128) {
            if (pair != 1 && num_sms >= 2 && (pair == 2 || N > 128)) return HELIO_BWD16(256, 2);
            return HELIO_BWD16(256, 1);
        }
        if (R > 64) return HELIO_BWD16(128, 1);
        return HELIO_BWD16(64, 1);
#undef HELIO_BWD16
    }
#define HELIO_BWD(NT_, CG_) launch_splat_bwd_tc<NT_, CG_>(params, counts, index, g_img, nullptr, moments, B, N, R, width, height, num_sms, st)
    if (R > 128) {
        // CTA pairs need a second block of 128 heliostats to be worth it
        if (pair != 1 && num_sms >= 2 && (pair == 2 || N > 128)) return HELIO_BWD(256, 2);
        return HELIO_BWD(256, 1);
    }
    if (R > 64) return HELIO_BWD(128, 1);
    return HELIO_BWD(64, 1);
#undef HELIO_BWD
}

}  // namespace helio
