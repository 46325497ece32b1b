// K2 / K3, CUDA-core path: shared-memory-staged heliostat tiles.
//
// Used for small fields / odd resolutions and as the independent cross-check of the tcgen05 path.
// Forward : img[b] = sum_n amp_n Gx_n (x) Gy_n  as register-tiled rank-1 updates (1 FMA per
//           heliostat-pixel), Gaussians staged per chunk of heliostats in shared memory,
//           coalesced float4 image stores.
// Backward: per (b, n) the moments {S0,Sx,Sy,S2} of g*G, recomputing the Gaussians; one warp owns
//           NH heliostats, lanes own image columns, g rows staged in shared memory (2 FMA / eval).
#pragma once
#include "helio_common.cuh"

namespace helio {

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int TI, int TJ, int MI, int MJ, int CH>
struct SplatFwdCfg {
    static constexpr int kTI = TI, kTJ = TJ, kMI = MI, kMJ = MJ, kCH = CH;
    static constexpr int kThreadsI = TI / MI, kThreadsJ = TJ / MJ;
    static constexpr int kThreads = kThreadsI * kThreadsJ;
    static constexpr int kVI = MI / 4, kVJ = MJ / 4;  // float4 groups per thread along i / j
    static_assert(MI % 4 == 0 && MJ % 4 == 0, "micro-tile is built from float4 groups");
    static constexpr int kSmemBytes = CH * (TI + TJ) * 4;
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, 2)
splat_fwd_simt_kernel(const float4* __restrict__ params, float* __restrict__ img, int N, int R, Axis ax, Axis ay,
                      int tiles_i, int tiles_j) {
    constexpr int TI = Cfg::kTI, TJ = Cfg::kTJ, CH = Cfg::kCH, VI = Cfg::kVI, VJ = Cfg::kVJ;
    constexpr int NT = Cfg::kThreads;
    __shared__ __align__(16) float sGx[CH][TI];
    __shared__ __align__(16) float sGy[CH][TJ];
    __shared__ float4 sPar[CH];

    const int tile = blockIdx.x % (tiles_i * tiles_j);
    const int b = blockIdx.x / (tiles_i * tiles_j);
    const int i0 = (tile / tiles_j) * TI, j0 = (tile % tiles_j) * TJ;
    const int tid = threadIdx.x;
    const int ti = tid / Cfg::kThreadsJ, tj = tid % Cfg::kThreadsJ;

    float acc[Cfg::kMI][Cfg::kMJ];
#pragma unroll
    for (int i = 0; i < Cfg::kMI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::kMJ; ++j) acc[i][j] = 0.f;

    const float4* pb = params + (size_t)b * N;
    for (int n0 = 0; n0 < N; n0 += CH) {
        __syncthreads();  // previous chunk fully consumed
        if (tid < CH) sPar[tid] = (n0 + tid < N) ? __ldg(pb + n0 + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        // stage amp*Gx[c][i] and Gy[c][j] for the chunk (padding heliostats have amp = 0)
        for (int e = tid; e < CH * TI; e += NT) {
            const int c = e / TI, i = e % TI;
            const float4 p = sPar[c];
            const float d = axis_at(ax, i0 + i) - p.x;
            sGx[c][i] = (i0 + i < R) ? p.w * ex2(-p.z * d * d) : 0.f;
        }
        for (int e = tid; e < CH * TJ; e += NT) {
            const int c = e / TJ, j = e % TJ;
            const float4 p = sPar[c];
            const float d = axis_at(ay, j0 + j) - p.y;
            sGy[c][j] = ex2(-p.z * d * d);
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < CH; ++c) {
            float gx[Cfg::kMI], gy[Cfg::kMJ];
#pragma unroll
            for (int v = 0; v < VI; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(&sGx[c][v * (TI / VI) + ti * 4]);
                gx[4 * v] = t.x, gx[4 * v + 1] = t.y, gx[4 * v + 2] = t.z, gx[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int v = 0; v < VJ; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(&sGy[c][v * (TJ / VJ) + tj * 4]);
                gy[4 * v] = t.x, gy[4 * v + 1] = t.y, gy[4 * v + 2] = t.z, gy[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < Cfg::kMI; ++i)
#pragma unroll
                for (int j = 0; j < Cfg::kMJ; ++j) acc[i][j] = fmaf(gx[i], gy[j], acc[i][j]);
        }
    }
    // store: each float4 group is 16 B aligned when R % 4 == 0
    float* ib = img + (size_t)b * R * R;
    const bool vec = (R & 3) == 0;
#pragma unroll
    for (int vi = 0; vi < VI; ++vi)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + vi * (TI / VI) + ti * 4 + e;
            if (i >= R) continue;
#pragma unroll
            for (int vj = 0; vj < VJ; ++vj) {
                const int j = j0 + vj * (TJ / VJ) + tj * 4;
                const float* a = &acc[vi * 4 + e][vj * 4];
                if (vec && j + 3 < R) {
                    *reinterpret_cast<float4*>(ib + (size_t)i * R + j) = make_float4(a[0], a[1], a[2], a[3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (j + q < R) ib[(size_t)i * R + j + q] = a[q];
                }
            }
        }
}

using SplatFwdBig = SplatFwdCfg<128, 128, 8, 8, 32>;   // 256 threads, 8x8 per thread
using SplatFwdSmall = SplatFwdCfg<64, 64, 4, 4, 32>;   // 256 threads, 4x4 per thread
using SplatFwdTiny = SplatFwdCfg<32, 32, 4, 4, 32>;    // 64 threads

template <class Cfg>
inline cudaError_t launch_splat_fwd_simt(const float* params, float* img, int B, int N, int R, float width,
                                         float height, cudaStream_t st) {
    const int tiles_i = (R + Cfg::kTI - 1) / Cfg::kTI, tiles_j = (R + Cfg::kTJ - 1) / Cfg::kTJ;
    const long long grid = (long long)B * tiles_i * tiles_j;
    splat_fwd_simt_kernel<Cfg><<<(unsigned)grid, Cfg::kThreads, 0, st>>>(
        reinterpret_cast<const float4*>(params), img, N, R, make_axis(width, R), make_axis(height, R), tiles_i, tiles_j);
    return cudaGetLastError();
}

inline cudaError_t splat_fwd_simt(const float* params, float* img, int B, int N, int R, float width, float height,
                                  int num_sms, cudaStream_t st) {
    // pick the largest tile that still gives every SM a CTA
    const long long big = (long long)B * ((R + 127) / 128) * ((R + 127) / 128);
    const long long small = (long long)B * ((R + 63) / 64) * ((R + 63) / 64);
    if (R >= 96 && big >= 2LL * num_sms) return launch_splat_fwd_simt<SplatFwdBig>(params, img, B, N, R, width, height, st);
    if (R >= 48 && small >= num_sms) return launch_splat_fwd_simt<SplatFwdSmall>(params, img, B, N, R, width, height, st);
    return launch_splat_fwd_simt<SplatFwdTiny>(params, img, B, N, R, width, height, st);
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// CTA = (sun b, chunk of WARPS*NH heliostats).  Lanes own JPL image columns of a 32*JPL-wide
// column block; rows of g are staged TR at a time in shared memory.
template <int WARPS, int NH, int JPL, int TR>
struct SplatBwdCfg {
    static constexpr int kWarps = WARPS, kNH = NH, kJPL = JPL, kTR = TR;
    static constexpr int kThreads = WARPS * 32;
    static constexpr int kHel = WARPS * NH;  // heliostats per CTA
    static constexpr int kCols = 32 * JPL;   // columns per block
    static_assert(JPL % 4 == 0, "lanes read float4 groups");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, 1)
splat_bwd_simt_kernel(const float4* __restrict__ params, const float* __restrict__ g_img, float4* __restrict__ moments,
                      int N, int R, Axis ax, Axis ay, int chunks) {
    constexpr int NH = Cfg::kNH, JPL = Cfg::kJPL, TR = Cfg::kTR, COLS = Cfg::kCols, NT = Cfg::kThreads;
    constexpr int VJ = JPL / 4;
    extern __shared__ __align__(16) float smem[];
    float* sG = smem;                 // [TR][COLS]   staged rows of g
    float* sGx = smem + TR * COLS;    // [kHel][R]    amp*Gx table of this CTA's heliostats
    float* sXs = sGx + Cfg::kHel * R;  // [R]          x_i

    const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nbase = chunk * Cfg::kHel;
    const float4* pb = params + (size_t)b * N;
    const float* gb = g_img + (size_t)b * R * R;

    for (int i = tid; i < R; i += NT) sXs[i] = axis_at(ax, i);
    for (int e = tid; e < Cfg::kHel * R; e += NT) {
        const int c = e / R, i = e % R;
        const int n = nbase + c;
        float val = 0.f;
        if (n < N) {
            const float4 p = __ldg(pb + n);
            const float d = axis_at(ax, i) - p.x;
            val = p.w * ex2(-p.z * d * d);
        }
        sGx[e] = val;
    }
    // per-warp heliostat parameters
    float4 par[NH];
#pragma unroll
    for (int q = 0; q < NH; ++q) {
        const int n = nbase + wid * NH + q;
        par[q] = (n < N) ? __ldg(pb + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float S0[NH], Sx[NH], Sxx[NH], Sy[NH], Syy[NH];
#pragma unroll
    for (int q = 0; q < NH; ++q) S0[q] = Sx[q] = Sxx[q] = Sy[q] = Syy[q] = 0.f;

    const bool vec = (R & 3) == 0;
    for (int j0 = 0; j0 < R; j0 += COLS) {
        // this lane's columns: j0 + v*(COLS/VJ) + lane*4 + e
        float gy[NH][JPL], U0[NH][JPL];
#pragma unroll
        for (int q = 0; q < NH; ++q)
#pragma unroll
            for (int v = 0; v < VJ; ++v)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = j0 + v * (COLS / VJ) + lane * 4 + e;
                    const float d = axis_at(ay, j < R ? j : R - 1) - par[q].y;
                    gy[q][v * 4 + e] = (j < R) ? ex2(-par[q].z * d * d) : 0.f;
                    U0[q][v * 4 + e] = 0.f;
                }
        for (int r0 = 0; r0 < R; r0 += TR) {
            __syncthreads();  // sG free (and, first time, sGx/sXs written)
            for (int e = tid * 4; e < TR * COLS; e += NT * 4) {
                const int r = e / COLS, c = e % COLS;
                const int i = r0 + r, j = j0 + c;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < R) {
                    const float* src = gb + (size_t)i * R + j;
                    if (vec && j + 3 < R) {
                        t = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (j < R) t.x = __ldg(src);
                        if (j + 1 < R) t.y = __ldg(src + 1);
                        if (j + 2 < R) t.z = __ldg(src + 2);
                        if (j + 3 < R) t.w = __ldg(src + 3);
                    }
                }
                *reinterpret_cast<float4*>(sG + e) = t;
            }
            __syncthreads();
            const int rows = min(TR, R - r0);
            for (int r = 0; r < rows; ++r) {
                float g[JPL];
#pragma unroll
                for (int v = 0; v < VJ; ++v) {
                    const float4 t = *reinterpret_cast<const float4*>(sG + r * COLS + v * (COLS / VJ) + lane * 4);
                    g[4 * v] = t.x, g[4 * v + 1] = t.y, g[4 * v + 2] = t.z, g[4 * v + 3] = t.w;
                }
                const float xi = sXs[r0 + r];
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    const float gx = sGx[(wid * NH + q) * R + r0 + r];
                    float row = 0.f;
#pragma unroll
                    for (int k = 0; k < JPL; ++k) {
                        row = fmaf(g[k], gy[q][k], row);
                        U0[q][k] = fmaf(g[k], gx, U0[q][k]);
                    }
                    const float dx = xi - par[q].x;
                    const float w0 = gx * row;
                    S0[q] += w0;
                    Sx[q] = fmaf(w0, dx, Sx[q]);
                    Sxx[q] = fmaf(w0 * dx, dx, Sxx[q]);
                }
            }
        }
        // fold this column block's U0 into the y moments
#pragma unroll
        for (int q = 0; q < NH; ++q)
#pragma unroll
            for (int v = 0; v < VJ; ++v)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = j0 + v * (COLS / VJ) + lane * 4 + e;
                    const float dy = axis_at(ay, j < R ? j : R - 1) - par[q].y;
                    const float w = gy[q][v * 4 + e] * U0[q][v * 4 + e];
                    Sy[q] = fmaf(w, dy, Sy[q]);
                    Syy[q] = fmaf(w * dy, dy, Syy[q]);
                }
    }
#pragma unroll
    for (int q = 0; q < NH; ++q) {
        const float s0 = warp_sum(S0[q]), sx = warp_sum(Sx[q]), sy = warp_sum(Sy[q]);
        const float s2 = warp_sum(Sxx[q] + Syy[q]);
        const int n = nbase + wid * NH + q;
        if (lane == 0 && n < N) moments[(size_t)b * N + n] = make_float4(s0, sx, sy, s2);
    }
}

using SplatBwdBig = SplatBwdCfg<16, 4, 8, 16>;   // 512 threads, 64 heliostats / CTA, 256-column blocks
using SplatBwdSmall = SplatBwdCfg<8, 2, 4, 16>;  // 256 threads, 16 heliostats / CTA, 128-column blocks

template <class Cfg>
inline cudaError_t launch_splat_bwd_simt(const float* params, const float* g_img, float* moments, int B, int N, int R,
                                         float width, float height, cudaStream_t st) {
    const int chunks = (N + Cfg::kHel - 1) / Cfg::kHel;
    const size_t smem = (size_t)(Cfg::kTR * Cfg::kCols + Cfg::kHel * R + R) * sizeof(float);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(splat_bwd_simt_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    splat_bwd_simt_kernel<Cfg><<<(unsigned)((long long)B * chunks), Cfg::kThreads, smem, st>>>(
        reinterpret_cast<const float4*>(params), g_img, reinterpret_cast<float4*>(moments), N, R, make_axis(width, R),
        make_axis(height, R), chunks);
    return cudaGetLastError();
}

inline cudaError_t splat_bwd_simt(const float* params, const float* g_img, float* moments, int B, int N, int R,
                                  float width, float height, int num_sms, cudaStream_t st) {
    const long long big = (long long)B * ((N + SplatBwdBig::kHel - 1) / SplatBwdBig::kHel);
    if (R > 128 && R <= 768 && big >= num_sms)
        return launch_splat_bwd_simt<SplatBwdBig>(params, g_img, moments, B, N, R, width, height, st);
    return launch_splat_bwd_simt<SplatBwdSmall>(params, g_img, moments, B, N, R, width, height, st);
}

}  // namespace helio
