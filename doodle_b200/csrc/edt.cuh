// Exact Euclidean distance transform of the thresholded target images: the GPU replacement of
// make_distance_maps (test_environment.py:92-97), which the reference runs per image on the host through
// scipy.ndimage.distance_transform_edt (scipy is unpinned in the reference's requirements.txt:3).
//
//   mask[b] = img[b] > thr * max(img[b])           (float32 compare, as numpy evaluates it)
//   dmap[b][i][j] = float32( sqrt( min over mask pixels (i',j') of (i-i')^2 + (j-j')^2 ) )      (float64 sqrt)
//
// Two exact integer passes (the separable squared-distance decomposition of Meijster et al. / Felzenszwalb):
//   columns: g[i][j]  = min_i' |i - i'| over mask pixels of column j (forward + backward scan, one thread per
//            (b, j), coalesced over j), saturating at INF = 2R when the column has no mask pixel;
//   rows   : d2[i][j] = min_j' (j - j')^2 + g[i][j']^2, one warp per image row with g^2 of the row in shared
//            memory; each lane scans outwards from its pixel in segments of 8 columns, skipping a segment whose
//            lower bound (gap^2 + segment minimum) cannot improve the result and stopping once gap^2 >= best.
// Everything is integer arithmetic, so the result is bit-identical to scipy's exact transform.  An image without
// any mask pixel reproduces scipy's behaviour for an input without background (distance to a virtual pixel at
// (-1, 0)); it only occurs for an all-zero image.
#pragma once
#include "helio_common.cuh"

namespace helio {

constexpr int kEdtThreads = 256;
constexpr int kEdtMaxR = 4096;   // INF = 2R must fit int16, d2 must fit int32

// columns pass: one thread per (b, j)
__global__ void __launch_bounds__(kEdtThreads)
edt_cols_kernel(const float* __restrict__ img, const float* __restrict__ mx, float thr, int B, int R, short* __restrict__ g) {
    const long long t = (long long)blockIdx.x * kEdtThreads + threadIdx.x;
    if (t >= (long long)B * R) return;
    const int b = (int)(t / R), j = (int)(t % R);
    const float cut = __fmul_rn(thr, __ldg(mx + b));
    const float* src = img + (size_t)b * R * R + j;
    short* dst = g + (size_t)b * R * R + j;
    const int inf = 2 * R;
    int d = inf;
    for (int i = 0; i < R; ++i) {
        d = (__ldg(src + (size_t)i * R) > cut) ? 0 : min(d + 1, inf);
        dst[(size_t)i * R] = (short)d;
    }
    d = inf;
    for (int i = R - 1; i >= 0; --i) {
        const int f = dst[(size_t)i * R];
        d = f == 0 ? 0 : min(d + 1, inf);
        if (d < f) dst[(size_t)i * R] = (short)d;
    }
}

// rows pass: one warp per (b, i).  Exact minimisation of (j - j')^2 + g^2[j'] with two prunings that never drop the
// minimiser: a direction stops once gap^2 >= best, and a segment of 8 columns is skipped when gap^2 + min(g^2 in the
// segment) >= best.  Rows far from the mask (g^2 large and flat) finish after a few segment tests; rows through the mask
// walk one segment minimum per 8 columns up to the distance itself.
constexpr int kEdtSeg = 8;

__global__ void __launch_bounds__(kEdtThreads)
edt_rows_kernel(const short* __restrict__ g, int B, int R, float* __restrict__ dmaps) {
    extern __shared__ int sEdt[];                          // [warps][R + nseg]  g^2 of each warp's row, then segment minima
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (kEdtThreads / 32) + warp;
    if (row >= (long long)B * R) return;                   // whole warp leaves together
    const int i = (int)(row % R);
    const int nseg = (R + kEdtSeg - 1) / kEdtSeg;
    int* g2 = sEdt + warp * (R + nseg);
    int* seg = g2 + R;
    const short* src = g + (size_t)row * R;
    for (int j = lane; j < R; j += 32) {
        const int v = src[j];
        g2[j] = v * v;
    }
    __syncwarp();
    for (int s = lane; s < nseg; s += 32) {
        int m = 0x7fffffff;
        for (int e = 0; e < kEdtSeg; ++e) {
            const int jj = s * kEdtSeg + e;
            if (jj < R) m = min(m, g2[jj]);
        }
        seg[s] = m;
    }
    __syncwarp();
    const int inf2 = 4 * R * R;
    float* dst = dmaps + (size_t)row * R;
    for (int j = lane; j < R; j += 32) {
        int best = g2[j];
        const int js = j / kEdtSeg;
        auto scan = [&](int s) {
            for (int e = 0; e < kEdtSeg; ++e) {
                const int jj = s * kEdtSeg + e;
                if (jj < R) {
                    const int d = jj - j;
                    best = min(best, g2[jj] + d * d);
                }
            }
        };
        scan(js);
        for (int s = js - 1; s >= 0; --s) {
            const int gap = j - (s * kEdtSeg + kEdtSeg - 1);
            const int gg = gap * gap;
            if (gg >= best) break;
            if (seg[s] + gg < best) scan(s);
        }
        for (int s = js + 1; s < nseg; ++s) {
            const int gap = s * kEdtSeg - j;
            const int gg = gap * gap;
            if (gg >= best) break;
            if (seg[s] + gg < best) scan(s);
        }
        if (best >= inf2) best = (i + 1) * (i + 1) + j * j;      // no mask pixel in the image: scipy's virtual pixel
        dst[j] = (float)sqrt((double)best);
    }
}

}  // namespace helio
