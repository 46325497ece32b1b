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
//                    lo*hi} tcgen05.mma kind::tf32 into a 128 x NT fp32 accumulator in TMEM;
//                    two accumulators ping-pong so the epilogue overlaps the next tile;
//   epilogue warps : tcgen05.ld the accumulator; forward stores image rows, backward folds the
//                    recomputed Gaussian weights into the per-heliostat moments.
// Pipelines: smem full/empty per stage (producers <-> MMA), TMEM full/empty per accumulator
// (MMA <-> epilogue), static round-robin tile schedule.
//
// Accuracy: hi*hi + hi*lo + lo*hi with fp32 accumulation drops only lo*lo (~2^-22 relative) and
// the truncation of lo (~2^-21): fp32-class results from the tensor pipe.
#pragma once
#include "helio_common.cuh"
#include "tc_common.cuh"

namespace helio {

constexpr int kTcMaxR = 1024;  // coordinate tables live in shared memory

// ================================================================================================
// forward
// ================================================================================================
template <int NT>
struct SplatFwdTc {
    static constexpr int kNT = NT;                       // UMMA N (image columns per tile)
    static constexpr int kM = 128;                       // UMMA M (image rows per tile)
    static constexpr int kKC = 32;                       // heliostats per stage (one 128-byte swizzle row)
    static constexpr int kStages = (NT == 256) ? 2 : 3;
    static constexpr int kProducerWarps = (kM + NT) / 32;
    static constexpr int kMmaWarp = kProducerWarps;
    static constexpr int kEpiWarp0 = kProducerWarps + 1;
    static constexpr int kThreads = (kProducerWarps + 1 + 4) * 32;
    static constexpr int kABytes = kM * 128;             // one of {hi, lo}
    static constexpr int kBBytes = NT * 128;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kTmemCols = 2 * NT;             // two accumulators
    static constexpr int kTableBytes = 2 * kTcMaxR * 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + kTableBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns: power of two <= 512");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// Producer mapping: a warp owns 32 operand rows (image rows for A, image columns for B) and its
// lanes own the 32 heliostats of the stage, so the footprint parameters sit in registers (one
// coalesced 512-byte load per warp and stage, prefetched one stage ahead) and every store is one
// conflict-free 128-byte row of the swizzled tile.
template <int NT>
__global__ void __launch_bounds__(SplatFwdTc<NT>::kThreads, 1)
splat_fwd_tc_kernel(const float4* __restrict__ params, float* __restrict__ img, int N, int R, Axis ax, Axis ay,
                    int tiles_i, int tiles_j, int num_tiles) {
    using C = SplatFwdTc<NT>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned operand stages, then coordinate tables, then barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* sX = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
    float* sY = sX + kTcMaxR;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sY + kTcMaxR);
    uint64_t* full = bars;                       // [kStages]  producers -> MMA
    uint64_t* empty = bars + C::kStages;         // [kStages]  MMA -> producers
    uint64_t* tfull = bars + 2 * C::kStages;     // [2]        MMA -> epilogue
    uint64_t* tempty = tfull + 2;                // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // pixel-centre tables, padded with 0 (rows/columns >= R are computed but never stored)
    for (int i = threadIdx.x; i < kTcMaxR; i += C::kThreads) {
        sX[i] = i < R ? axis_at(ax, i) : 0.f;
        sY[i] = i < R ? axis_at(ay, i) : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            tc::mbar_init(&full[s], C::kProducerWarps);
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
        const bool isA = warp < C::kM / 32;
        const int wrow = (isA ? warp : warp - C::kM / 32) * 32;          // first operand row of this warp
        const uint32_t region = (isA ? 0u : 2u * C::kABytes) + (uint32_t)(wrow >> 3) * 1024u;
        const uint32_t lo_delta = isA ? C::kABytes : C::kBBytes;
        // byte offset of this lane's element inside a 128-byte row whose (row & 7) == c
        uint32_t xoff[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) xoff[c] = ((((uint32_t)lane >> 2) ^ (uint32_t)c) << 4) + ((uint32_t)lane & 3u) * 4u;
        uint32_t it = 0;                             // global stage counter
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int g0 = (isA ? (t / tiles_j) * C::kM : (t % tiles_j) * NT) + wrow;
            const float* tab = (isA ? sX : sY) + g0;
            float xr[32];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                const float4 v = *reinterpret_cast<const float4*>(tab + e);
                xr[e] = v.x, xr[e + 1] = v.y, xr[e + 2] = v.z, xr[e + 3] = v.w;
            }
            const float4* pb = params + (size_t)b * N;
            float4 pn = lane < N ? __ldg(pb + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < nchunks; ++c, ++it) {
                const float4 p = pn;
                const bool have = c * C::kKC + lane < N;
                const int nn = (c + 1) * C::kKC + lane;
                pn = nn < N ? __ldg(pb + nn) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float ctr = isA ? p.x : p.y;
                const float nk2 = -p.z;
                const float scale = have ? (isA ? p.w : 1.f) : 0.f;     // K padding: exact zeros
                const int s = it % C::kStages;
                const uint32_t ph = (it / C::kStages) & 1;
                if (lane == 0) tc::mbar_wait(&empty[s], ph ^ 1);
                __syncwarp();
                uint8_t* base = smem + s * C::kStageBytes + region;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float d = xr[e] - ctr;
                    const float v = ex2((d * nk2) * d) * scale;
                    float hi, lo;
                    tc::split_tf32(v, hi, lo);
                    uint8_t* dst = base + (e >> 3) * 1024 + (e & 7) * 128 + xoff[e & 7];
                    *reinterpret_cast<float*>(dst) = hi;
                    *reinterpret_cast<float*>(dst + lo_delta) = lo;
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&full[s]);
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
                if (j0 + cb >= R) break;
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

inline bool splat_tc_fwd_supported(int B, int N, int R) { return B > 0 && N > 0 && R >= 8 && R <= kTcMaxR; }
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

// ================================================================================================
// backward
// ================================================================================================
// Tile = (sun b, block of 128 heliostats).  Per tile two products run back to back through the
// same pipeline, each split into ceil(R/NT) accumulators of 128 x NT:
//   product 0 (T): A rows = Gy[n, j-chunk]       B rows = g[i, j-chunk]   (i = accumulator column)
//   product 1 (U): A rows = amp Gx[n, i-chunk]   B rows = g[i-chunk, j]^T (j = accumulator column)
// The epilogue thread that owns TMEM lane n keeps {S0,Sx,Sxx} from product 0 in registers, adds
// {Sy,Syy} from product 1 and writes one float4 per heliostat.
template <int NT>
struct SplatBwdTc {
    static constexpr int kNT = NT;                       // UMMA N (pixels per accumulator)
    static constexpr int kM = 128;                       // UMMA M (heliostats per tile)
    static constexpr int kKC = 32;                       // pixels per stage
    static constexpr int kStages = (NT == 256) ? 2 : (NT == 128 ? 3 : 4);
    static constexpr int kAWarps = kM / 32;              // Gaussian-operand producers
    static constexpr int kGWarps = NT / 32;              // gradient-tile stagers
    static constexpr int kMmaWarp = kAWarps + kGWarps;
    static constexpr int kEpiWarp0 = kMmaWarp + 1;
    static constexpr int kThreads = (kEpiWarp0 + 4) * 32;
    static constexpr int kABytes = kM * 128;
    static constexpr int kGBytes = NT * 128;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kGBytes;
    static constexpr int kTmemCols = 2 * NT;
    static constexpr int kTableBytes = 2 * kTcMaxR * 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + kTableBytes + 1024 + 256;
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32, "TMEM columns");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int NT>
__global__ void __launch_bounds__(SplatBwdTc<NT>::kThreads, 1)
splat_bwd_tc_kernel(const float4* __restrict__ params, const float* __restrict__ g_img, float4* __restrict__ moments,
                    int N, int R, Axis ax, Axis ay, int nblocks, int num_tiles) {
    using C = SplatBwdTc<NT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* sX = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
    float* sY = sX + kTcMaxR;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sY + kTcMaxR);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::kStages;
    uint64_t* tfull = bars + 2 * C::kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kTcMaxR; i += C::kThreads) {
        sX[i] = i < R ? axis_at(ax, i) : 0.f;
        sY[i] = i < R ? axis_at(ay, i) : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            tc::mbar_init(&full[s], C::kAWarps + C::kGWarps);
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

    const int kchunks = (R + C::kKC - 1) / C::kKC;   // K stages per accumulator
    const int pblocks = (R + NT - 1) / NT;           // accumulators per product
    const bool vec = (R & 3) == 0;

    if (warp < C::kAWarps) {
        // ================= Gaussian operand: thread = heliostat row =================
        const int r = threadIdx.x;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / nblocks, nb = tile % nblocks;
            const int n = nb * C::kM + r;
            const bool live = n < N;
            const float4 p = live ? __ldg(params + (size_t)b * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float nk2 = -p.z;
#pragma unroll 1
            for (int prod = 0; prod < 2; ++prod) {
                const float ctr = prod == 0 ? p.y : p.x;
                const float scale = live ? (prod == 0 ? 1.f : p.w) : 0.f;
                const float* tab = prod == 0 ? sY : sX;
#pragma unroll 1
                for (int pbk = 0; pbk < pblocks; ++pbk) {
#pragma unroll 1
                    for (int c = 0; c < kchunks; ++c, ++it) {
                        const int s = it % C::kStages;
                        const uint32_t ph = (it / C::kStages) & 1;
                        const int k0 = c * C::kKC;
                        if (lane == 0) tc::mbar_wait(&empty[s], ph ^ 1);
                        __syncwarp();
                        uint8_t* hi_base = smem + s * C::kStageBytes;
                        uint8_t* lo_base = hi_base + C::kABytes;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 xs = *reinterpret_cast<const float4*>(tab + k0 + 4 * q);
                            const float x[4] = {xs.x, xs.y, xs.z, xs.w};
                            float hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float d = x[e] - ctr;
                                float v = ex2((d * nk2) * d) * scale;
                                if (k0 + 4 * q + e >= R) v = 0.f;               // K padding: exact zeros
                                tc::split_tf32(v, hi[e], lo[e]);
                            }
                            const uint32_t off = tc::sw128_offset((uint32_t)r, (uint32_t)q);
                            *reinterpret_cast<float4*>(hi_base + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<float4*>(lo_base + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        }
                        tc::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&full[s]);
                    }
                }
            }
        }
    } else if (warp < C::kMmaWarp) {
        // ================= gradient tile stagers =================
        const int t = threadIdx.x - C::kAWarps * 32;     // 0..NT-1
        const int gw = t >> 5;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / nblocks;
            const float* gb = g_img + (size_t)b * R * R;
#pragma unroll 1
            for (int prod = 0; prod < 2; ++prod) {
#pragma unroll 1
                for (int pbk = 0; pbk < pblocks; ++pbk) {
#pragma unroll 1
                    for (int c = 0; c < kchunks; ++c, ++it) {
                        const int s = it % C::kStages;
                        const uint32_t ph = (it / C::kStages) & 1;
                        const int k0 = c * C::kKC;
                        float4 vals[8];
                        uint32_t offs[8];
                        if (prod == 0) {
                            // operand row = image row i (accumulator column), K = image column j:
                            // 8 lanes cover one 128-byte row segment, a warp instruction covers 4 rows
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int row = gw * 32 + q * 4 + (lane >> 3), ch = lane & 7;
                                const int i = pbk * NT + row, j = k0 + ch * 4;
                                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (i < R) {
                                    const float* src = gb + (size_t)i * R + j;
                                    if (vec && j + 3 < R) {
                                        v = __ldg(reinterpret_cast<const float4*>(src));
                                    } else {
                                        if (j < R) v.x = __ldg(src);
                                        if (j + 1 < R) v.y = __ldg(src + 1);
                                        if (j + 2 < R) v.z = __ldg(src + 2);
                                        if (j + 3 < R) v.w = __ldg(src + 3);
                                    }
                                }
                                vals[q] = v;
                                offs[q] = tc::sw128_offset((uint32_t)row, (uint32_t)ch);
                            }
                        } else {
                            // operand row = image column j (accumulator column), K = image row i:
                            // lanes read consecutive columns of one image row (coalesced), transposing in registers
                            const int j = pbk * NT + t;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                float x[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int i = k0 + 4 * q + e;
                                    x[e] = (i < R && j < R) ? __ldg(gb + (size_t)i * R + j) : 0.f;
                                }
                                vals[q] = make_float4(x[0], x[1], x[2], x[3]);
                                offs[q] = tc::sw128_offset((uint32_t)t, (uint32_t)q);
                            }
                        }
                        if (lane == 0) tc::mbar_wait(&empty[s], ph ^ 1);
                        __syncwarp();
                        uint8_t* hi_base = smem + s * C::kStageBytes + 2 * C::kABytes;
                        uint8_t* lo_base = hi_base + C::kGBytes;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 h, l;
                            tc::split_tf32(vals[q].x, h.x, l.x);
                            tc::split_tf32(vals[q].y, h.y, l.y);
                            tc::split_tf32(vals[q].z, h.z, l.z);
                            tc::split_tf32(vals[q].w, h.w, l.w);
                            *reinterpret_cast<float4*>(hi_base + offs[q]) = h;
                            *reinterpret_cast<float4*>(lo_base + offs[q]) = l;
                        }
                        tc::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&full[s]);
                    }
                }
            }
        }
    } else if (warp == C::kMmaWarp) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = tc::make_idesc_tf32(C::kM, NT);
        uint32_t it = 0, sub = 0;
        const int subs_per_tile = 2 * pblocks;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int sb = 0; sb < subs_per_tile; ++sb, ++sub) {
                const int acc = sub & 1;
                const uint32_t aph = (sub >> 1) & 1;
                tc::mbar_wait(&tempty[acc], aph ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NT);
                for (int c = 0; c < kchunks; ++c, ++it) {
                    const int s = it % C::kStages;
                    const uint32_t ph = (it / C::kStages) & 1;
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = tc::smem_u32(smem + s * C::kStageBytes);
                        const uint64_t a_hi = tc::make_desc_k_sw128(sa);
                        const uint64_t a_lo = tc::make_desc_k_sw128(sa + C::kABytes);
                        const uint64_t b_hi = tc::make_desc_k_sw128(sa + 2 * C::kABytes);
                        const uint64_t b_lo = tc::make_desc_k_sw128(sa + 2 * C::kABytes + C::kGBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ko = (uint64_t)(k * 32 >> 4);
                            tc::mma_tf32_ss(d_tmem, a_hi + ko, b_hi + ko, idesc, (c | k) != 0);
                            tc::mma_tf32_ss(d_tmem, a_hi + ko, b_lo + ko, idesc, 1);
                            tc::mma_tf32_ss(d_tmem, a_lo + ko, b_hi + ko, idesc, 1);
                        }
                        tc::mma_commit(&empty[s]);
                        if (c == kchunks - 1) tc::mma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================= epilogue: accumulator -> moments =================
        const int q = warp & 3;
        uint32_t sub = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / nblocks, nb = tile % nblocks;
            const int n = nb * C::kM + q * 32 + lane;
            const bool live = n < N;
            const float4 p = live ? __ldg(params + (size_t)b * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float nk2 = -p.z;
            float S0 = 0.f, Sx = 0.f, Sy = 0.f, S2 = 0.f;
#pragma unroll 1
            for (int prod = 0; prod < 2; ++prod) {
                const float ctr = prod == 0 ? p.x : p.y;
                const float* tab = prod == 0 ? sX : sY;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int pbk = 0; pbk < pblocks; ++pbk, ++sub) {
                    const int acc = sub & 1;
                    const uint32_t aph = (sub >> 1) & 1;
                    tc::mbar_wait(&tfull[acc], aph);
                    tc::tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
                    const int col0 = pbk * NT;
#pragma unroll 1
                    for (int cb = 0; cb < NT; cb += 32) {
                        if (col0 + cb >= R) break;
                        float v[32];
                        tc::tmem_ld_32x32(taddr + cb, v);
#pragma unroll
                        for (int e4 = 0; e4 < 32; e4 += 4) {
                            const float4 xs = *reinterpret_cast<const float4*>(tab + col0 + cb + e4);
                            const float x[4] = {xs.x, xs.y, xs.z, xs.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // columns >= R hold exact zeros (their operand rows are zero)
                                const float d = x[e] - ctr;
                                const float w = ex2((d * nk2) * d) * v[e4 + e];
                                s0 += w;
                                s1 = fmaf(w, d, s1);
                                s2 = fmaf(w * d, d, s2);
                            }
                        }
                    }
                    tc::tc_fence_before();
                    tc::mbar_arrive(&tempty[acc]);
                }
                if (prod == 0) {
                    S0 = s0 * p.w, Sx = s1 * p.w, S2 = s2 * p.w;
                } else {
                    Sy = s1, S2 += s2;
                }
            }
            if (live) moments[(size_t)b * N + n] = make_float4(S0, Sx, Sy, S2);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == C::kMmaWarp) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

inline bool splat_tc_bwd_supported(int B, int N, int R) { return B > 0 && N > 0 && R >= 8 && R <= kTcMaxR; }
inline bool splat_tc_bwd_preferred(int B, int N, int R) { return R >= 128 && N >= 64; }

template <int NT>
inline cudaError_t launch_splat_bwd_tc(const float* params, const float* g_img, float* moments, int B, int N, int R,
                                       float width, float height, int num_sms, cudaStream_t st) {
    using C = SplatBwdTc<NT>;
    const int nblocks = (N + C::kM - 1) / C::kM;
    const long long num_tiles = (long long)B * nblocks;
    if (num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(splat_bwd_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    const int grid = (int)(num_tiles < num_sms ? num_tiles : num_sms);
    splat_bwd_tc_kernel<NT><<<grid, C::kThreads, C::kSmemBytes, st>>>(
        reinterpret_cast<const float4*>(params), g_img, reinterpret_cast<float4*>(moments), N, R, make_axis(width, R),
        make_axis(height, R), nblocks, (int)num_tiles);
    return cudaGetLastError();
}

inline cudaError_t splat_tc_bwd(const float* params, const float* g_img, float* moments, int B, int N, int R, float width,
                                float height, int num_sms, cudaStream_t st) {
    if (R > 128) return launch_splat_bwd_tc<256>(params, g_img, moments, B, N, R, width, height, num_sms, st);
    if (R > 64) return launch_splat_bwd_tc<128>(params, g_img, moments, B, N, R, width, height, num_sms, st);
    return launch_splat_bwd_tc<64>(params, g_img, moments, B, N, R, width, height, num_sms, st);
}

}  // namespace helio
