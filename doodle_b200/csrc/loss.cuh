// K4: HelioEnv.step image losses (test_environment.py:436-457, 492), forward partials and adjoint.
// HBM-bound elementwise + per-image reductions: one CTA per image slice, float4 loads, warp-shuffle
// reductions, ordered combination (deterministic).
#pragma once
#include "helio_common.cuh"

namespace helio {

constexpr int kLossThreads = 256;

template <int NV>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* sh /* [NV][32] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sh[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float x = lane < nw ? sh[k * 32 + lane] : 0.f;
            v[k] = warp_sum(x);
        }
    }
}

// tx[b] = max(max target[b], 1e-6).  One CTA per image.
__global__ void __launch_bounds__(kLossThreads) image_max_kernel(const float* __restrict__ target, int R, float* __restrict__ tx) {
    const size_t npix = (size_t)R * R;
    const float* t = target + (size_t)blockIdx.x * npix;
    float m = -INFINITY;
    if ((npix & 3) == 0) {
        const float4* t4 = reinterpret_cast<const float4*>(t);
        for (size_t i = threadIdx.x; i < npix / 4; i += kLossThreads) {
            const float4 q = __ldg(t4 + i);
            m = fmaxf(fmaxf(fmaxf(m, q.x), fmaxf(q.y, q.z)), q.w);
        }
    } else {
        for (size_t i = threadIdx.x; i < npix; i += kLossThreads) m = fmaxf(m, __ldg(t + i));
    }
    __shared__ float sh[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float x = threadIdx.x < kLossThreads / 32 ? sh[threadIdx.x] : -INFINITY;
        x = warp_max(x);
        if (threadIdx.x == 0) tx[blockIdx.x] = fmaxf(x, 1e-6f);
    }
}

// per_img[b] = { sum diff^2, sum |diff| dmaps, sum |diff| },  diff = (img - target)/tx[b]
__global__ void __launch_bounds__(kLossThreads)
loss_fwd_kernel(const float* __restrict__ img, const float* __restrict__ target, const float* __restrict__ dmaps,
                const float* __restrict__ tx, int R, float* __restrict__ per_img) {
    const size_t npix = (size_t)R * R, off = (size_t)blockIdx.x * npix;
    const float t = __ldg(tx + blockIdx.x);   // reference divides both images by tx
    float acc[3] = {0.f, 0.f, 0.f};
    auto one = [&](float p, float q, float d) {
        const float diff = p / t - q / t;
        const float e = fabsf(diff);
        acc[0] = fmaf(diff, diff, acc[0]);
        acc[1] = fmaf(e, d, acc[1]);
        acc[2] += e;
    };
    if ((npix & 3) == 0) {
        const float4* a = reinterpret_cast<const float4*>(img + off);
        const float4* c = reinterpret_cast<const float4*>(target + off);
        const float4* d = reinterpret_cast<const float4*>(dmaps + off);
        for (size_t i = threadIdx.x; i < npix / 4; i += kLossThreads) {
            const float4 p = __ldg(a + i), q = __ldg(c + i), w = __ldg(d + i);
            one(p.x, q.x, w.x), one(p.y, q.y, w.y), one(p.z, q.z, w.z), one(p.w, q.w, w.w);
        }
    } else {
        for (size_t i = threadIdx.x; i < npix; i += kLossThreads) one(__ldg(img + off + i), __ldg(target + off + i), __ldg(dmaps + off + i));
    }
    __shared__ float sh[3 * 32];
    block_reduce<3>(acc, sh);
    if (threadIdx.x == 0) {
        per_img[3 * blockIdx.x] = acc[0];
        per_img[3 * blockIdx.x + 1] = acc[1];
        per_img[3 * blockIdx.x + 2] = acc[2];
    }
}

// g_img = (2 g0 diff + (g1 dmaps + g2) sign(diff)) / tx (+ g_img_in)
__global__ void __launch_bounds__(kLossThreads)
loss_bwd_kernel(const float* __restrict__ img, const float* __restrict__ target, const float* __restrict__ dmaps,
                const float* __restrict__ tx, const float* __restrict__ g_per_img, const float* __restrict__ g_in, int R,
                int slices, float* __restrict__ g_img) {
    const int b = blockIdx.x / slices, s = blockIdx.x % slices;
    const size_t npix = (size_t)R * R, off = (size_t)b * npix;
    const float t = __ldg(tx + b);
    const float g0 = __ldg(g_per_img + 3 * b), g1 = __ldg(g_per_img + 3 * b + 1), g2 = __ldg(g_per_img + 3 * b + 2);
    auto one = [&](float p, float q, float d, float gi) {
        const float diff = p / t - q / t;
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        return (2.f * g0 * diff + (g1 * d + g2) * sg) / t + gi;
    };
    if ((npix & 3) == 0) {
        const float4* a = reinterpret_cast<const float4*>(img + off);
        const float4* c = reinterpret_cast<const float4*>(target + off);
        const float4* d = reinterpret_cast<const float4*>(dmaps + off);
        const float4* gi = g_in ? reinterpret_cast<const float4*>(g_in + off) : nullptr;
        float4* o = reinterpret_cast<float4*>(g_img + off);
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix / 4; i += (size_t)slices * kLossThreads) {
            const float4 p = __ldg(a + i), q = __ldg(c + i), w = __ldg(d + i);
            const float4 z = gi ? __ldg(gi + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            o[i] = make_float4(one(p.x, q.x, w.x, z.x), one(p.y, q.y, w.y, z.y), one(p.z, q.z, w.z, z.z), one(p.w, q.w, w.w, z.w));
        }
    } else {
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix; i += (size_t)slices * kLossThreads)
            g_img[off + i] = one(__ldg(img + off + i), __ldg(target + off + i), __ldg(dmaps + off + i), g_in ? __ldg(g_in + off + i) : 0.f);
    }
}

}  // namespace helio
