// Footprint culling (opt-in): per sun, compact the heliostats whose Gaussian can reach the receiver at all.
//
// With large orientation errors most reflected rays miss the receiver by many sigma (BASELINE shape: 90 mrad errors at
// ~95 m put the hit point ~17 m off per axis against a 7.5 m half-width), and their footprint on every pixel is below
// 2^-40 of its peak: an exact zero at fp32 image precision.  The dense contraction spends the same tensor work on them
// as on the mirrors that hit.  Culling keeps, per sun and in the original order, the heliostats with
//     k2 * (dx^2 + dy^2) <= kCullExponent,   dx = max(|a| - w/2, 0), dy = max(|b| - h/2, 0)
// ({a, b, k2, amp} = K1's footprint; invalid rays have k2 = 0 and are always kept), so that K2 contracts over
// counts[b] <= N heliostats and K3 visits only the kept ones; culled heliostats receive exactly zero image gradient.
// Every dropped term is < amp * 2^-40 on every pixel: N = 5000 of them move a pixel by < 5e-9 (tolerance: 1e-6 + 1e-4 |v|).
#pragma once
#include "helio_common.cuh"

namespace helio {

constexpr float kCullExponent = 40.f;
constexpr int kCullThreads = 256;

struct CullBuffers {          // carved out of one caller-provided workspace
    float4* cparams;          // [B][N] compacted footprints (first counts[b] entries of each row are valid)
    int* index;               // [B][N] original heliostat index of each compacted entry
    int* counts;              // [B]
};

inline int64_t cull_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (int64_t)B * N * 16 + (int64_t)B * N * 4 + (((int64_t)B * 4 + 255) / 256) * 256;
}
inline CullBuffers cull_carve(void* ws, int B, int N) {
    CullBuffers c;
    char* p = reinterpret_cast<char*>(ws);
    c.cparams = reinterpret_cast<float4*>(p);
    c.index = reinterpret_cast<int*>(p + (size_t)B * N * 16);
    c.counts = reinterpret_cast<int*>(p + (size_t)B * N * 20);
    return c;
}

// one CTA per sun; order-preserving compaction (ballot + warp prefix + block prefix)
__global__ void __launch_bounds__(kCullThreads)
cull_kernel(const float4* __restrict__ params, int N, float half_w, float half_h, CullBuffers out) {
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float4* pb = params + (size_t)b * N;
    float4* cb = out.cparams + (size_t)b * N;
    int* ib = out.index + (size_t)b * N;
    __shared__ int warp_tot[kCullThreads / 32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += kCullThreads) {
        const int n = n0 + threadIdx.x;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (n < N) {
            p = __ldg(pb + n);
            const float dx = fmaxf(fabsf(p.x) - half_w, 0.f), dy = fmaxf(fabsf(p.y) - half_h, 0.f);
            keep = !(p.z * (dx * dx + dy * dy) > kCullExponent);      // NaN compares false: kept
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < wid; ++w) off += warp_tot[w];
        if (keep) {
            const int dst = off + __popc(m & ((1u << lane) - 1u));
            cb[dst] = p;
            ib[dst] = n;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < kCullThreads / 32; ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out.counts[b] = base;
        if (base == 0) cb[0] = make_float4(0.f, 0.f, 0.f, 1.f);   // benign entry for the all-padding stage of an empty sun
    }
}

}  // namespace helio
