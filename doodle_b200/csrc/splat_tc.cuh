// K2 / K3 on the 5th-generation tensor cores: img[b] = Gx^T diag(amp) Gy as a 3xTF32 contraction.
//
// Forward (splat_fwd_tc_kernel), one persistent CTA per SM, warp-specialised:
//   producers  (128 + NT threads) : thread r owns ONE operand row (image row i for A, image column j
//                                   for B); per stage of 32 heliostats it evaluates its 32 Gaussians
//                                   exp2(-k2 (x - a)^2), splits each into tf32 hi + fp32 remainder and
//                                   writes them as K-major SWIZZLE_128B rows (conflict-free STS.128);
//   MMA warp   (1 elected thread) : per stage 4 K-steps x {hi*hi, hi*lo, lo*hi} tcgen05.mma kind::tf32,
//                                   M=128 x N=NT fp32 accumulator in TMEM, double-buffered (2*NT cols);
//   epilogue   (4 warps)          : tcgen05.ld 32x32b -> 128-byte-per-thread row segments -> st.global.
// Pipelines: smem full/empty per stage (producers <-> MMA), TMEM full/empty per accumulator
// (MMA <-> epilogue), static round-robin tile schedule (tile = sun b, 128 image rows, NT columns).
//
// Accuracy: hi*hi + hi*lo + lo*hi with fp32 accumulation drops only lo*lo (~2^-22 relative) and
// the truncation of lo (~2^-20), i.e. fp32-class results from the tensor pipe.
#pragma once
#include "helio_common.cuh"
#include "tc_common.cuh"

namespace helio {

template <int NT>
struct SplatFwdTc {
    static constexpr int kNT = NT;                       // UMMA N (image columns per tile)
    static constexpr int kM = 128;                       // UMMA M (image rows per tile)
    static constexpr int kKC = 32;                       // heliostats per stage (one 128-byte swizzle row)
    static constexpr int kStages = (NT == 256) ? 2 : 3;
    static constexpr int kProducerThreads = kM + NT;
    static constexpr int kProducerWarps = kProducerThreads / 32;
    static constexpr int kMmaWarp = kProducerWarps;
    static constexpr int kEpiWarp0 = kProducerWarps + 1;
    static constexpr int kThreads = (kProducerWarps + 1 + 4) * 32;
    static constexpr int kABytes = kM * 128;             // one of {hi, lo}
    static constexpr int kBBytes = NT * 128;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kTmemCols = 2 * NT;             // two accumulators
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns: power of two <= 512");
    static_assert(kEpiWarp0 % 4 == 1 || true, "epilogue warps cover the four TMEM lane quarters via warp_idx % 4");
};

template <int NT>
__global__ void __launch_bounds__(SplatFwdTc<NT>::kThreads, 1)
splat_fwd_tc_kernel(const float4* __restrict__ params, float* __restrict__ img, int N, int R, Axis ax, Axis ay,
                    int tiles_i, int tiles_j, int num_tiles) {
    using C = SplatFwdTc<NT>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned operand stages, then barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full = bars;                       // [kStages]  producers -> MMA
    uint64_t* empty = bars + C::kStages;         // [kStages]  MMA -> producers
    uint64_t* tfull = bars + 2 * C::kStages;     // [2]        MMA -> epilogue
    uint64_t* tempty = tfull + 2;                // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            tc::mbar_init(&full[s], C::kProducerThreads);
            tc::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tfull[a], 1);
            tc::mbar_init(&tempty[a], 128);
        }
        tc::mbar_fence_init();
    }
    if (warp == C::kMmaWarp) {
        tc::tmem_alloc(tmem_slot, C::kTmemCols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nchunks = (N + C::kKC - 1) / C::kKC;
    const int tiles_per_img = tiles_i * tiles_j;

    if (warp < C::kProducerWarps) {
        // ================= producers =================
        const int row = threadIdx.x;                 // 0..127: A rows, 128..: B rows
        const bool isA = row < C::kM;
        const int r = isA ? row : row - C::kM;
        uint32_t it = 0;                             // global stage counter
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int i0 = (t / tiles_j) * C::kM, j0 = (t % tiles_j) * NT;
            const int g = isA ? i0 + r : j0 + r;     // image row (A) or column (B) of this thread
            const bool live = g < R;
            const float x = isA ? axis_at(ax, g) : axis_at(ay, g);
            const float4* pb = params + (size_t)b * N;
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = (it / C::kStages) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * C::kStageBytes;
                uint8_t* hi_base = st + (isA ? 0 : 2 * C::kABytes);
                uint8_t* lo_base = hi_base + (isA ? C::kABytes : C::kBBytes);
                const int n0 = c * C::kKC;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int n = n0 + 4 * q + e;
                        float v = 0.f;
                        if (live && n < N) {
                            const float4 p = __ldg(pb + n);
                            const float d = x - (isA ? p.x : p.y);
                            v = ex2(-p.z * d * d);
                            if (isA) v *= p.w;
                        }
                        tc::split_tf32(v, hi[e], lo[e]);
                    }
                    const uint32_t off = tc::sw128_offset((uint32_t)r, (uint32_t)q);
                    *reinterpret_cast<float4*>(hi_base + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(lo_base + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&full[s]);
            }
        }
    } else if (warp == C::kMmaWarp) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = tc::make_idesc_tf32(C::kM, NT);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int acc = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            tc::mbar_wait(&tempty[acc], aph ^ 1);
            tc::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NT);
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = (it / C::kStages) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = tc::smem_u32(smem + s * C::kStageBytes);
                    const uint64_t a_hi = tc::make_desc_k_sw128(sa);
                    const uint64_t a_lo = tc::make_desc_k_sw128(sa + C::kABytes);
                    const uint64_t b_hi = tc::make_desc_k_sw128(sa + 2 * C::kABytes);
                    const uint64_t b_lo = tc::make_desc_k_sw128(sa + 2 * C::kABytes + C::kBBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ko = (uint64_t)(k * 32 >> 4);   // +32 bytes per K step of 8 tf32
                        tc::mma_tf32_ss(d_tmem, a_hi + ko, b_hi + ko, idesc, (c | k) != 0);
                        tc::mma_tf32_ss(d_tmem, a_hi + ko, b_lo + ko, idesc, 1);
                        tc::mma_tf32_ss(d_tmem, a_lo + ko, b_hi + ko, idesc, 1);
                    }
                    tc::mma_commit(&empty[s]);                          // stage reusable when these MMAs retire
                    if (c == nchunks - 1) tc::mma_commit(&tfull[acc]);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int i0 = (t / tiles_j) * C::kM, j0 = (t % tiles_j) * NT;
            const int acc = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            tc::mbar_wait(&tfull[acc], aph);
            tc::tc_fence_after();
            const int i = i0 + q * 32 + lane;
            float* dst = img + ((size_t)b * R + i) * R + j0;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
            const bool vec = (R & 3) == 0;
#pragma unroll 1
            for (int cb = 0; cb < NT; cb += 32) {
                float v[32];
                tc::tmem_ld_32x32(taddr + cb, v);
                if (i < R) {
                    if (vec && j0 + cb + 32 <= R) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(dst + cb + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (j0 + cb + e < R) dst[cb + e] = v[e];
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty[acc]);
        }
    }
    // ---- teardown ----
    tc::tc_fence_before();
    __syncthreads();
    if (warp == C::kMmaWarp) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

inline bool splat_tc_fwd_supported(int B, int N, int R) { return B > 0 && N > 0 && R >= 8; }
inline bool splat_tc_fwd_preferred(int B, int N, int R) {
    // tensor path pays off once the contraction dimension and the image are large enough
    return R >= 128 && N >= 64;
}

template <int NT>
inline cudaError_t launch_splat_fwd_tc(const float* params, float* img, int B, int N, int R, float width, float height,
                                       int num_sms, cudaStream_t st) {
    using C = SplatFwdTc<NT>;
    const int tiles_i = (R + C::kM - 1) / C::kM, tiles_j = (R + NT - 1) / NT;
    const long long num_tiles = (long long)B * tiles_i * tiles_j;
    if (num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(splat_fwd_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    const int grid = (int)(num_tiles < num_sms ? num_tiles : num_sms);
    splat_fwd_tc_kernel<NT><<<grid, C::kThreads, C::kSmemBytes, st>>>(reinterpret_cast<const float4*>(params), img, N, R,
                                                                      make_axis(width, R), make_axis(height, R), tiles_i,
                                                                      tiles_j, (int)num_tiles);
    return cudaGetLastError();
}

inline cudaError_t splat_tc_fwd(const float* params, float* img, int B, int N, int R, float width, float height, int num_sms,
                                cudaStream_t st) {
    if (R > 128) return launch_splat_fwd_tc<256>(params, img, B, N, R, width, height, num_sms, st);
    return launch_splat_fwd_tc<128>(params, img, B, N, R, width, height, num_sms, st);
}

// ---- backward: not on the tensor path yet -------------------------------------------------------
inline bool splat_tc_bwd_supported(int, int, int) { return false; }
inline bool splat_tc_bwd_preferred(int, int, int) { return false; }
inline cudaError_t splat_tc_bwd(const float*, const float*, float*, int, int, int, float, float, int, cudaStream_t) {
    return cudaErrorNotSupported;
}

}  // namespace helio
