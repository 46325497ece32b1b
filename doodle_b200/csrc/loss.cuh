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

// Both forward kernels run one thread-block CLUSTER per image (`slices` CTAs, 1..8): each CTA reduces an
// interleaved slice of the pixels, the partials meet in the leader's view of the cluster's shared memory (DSMEM)
// and are combined in rank order, so the result is deterministic and small batches still fill the machine.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// tx[b] = max(max target[b], floor)   (floor = 1e-6 for the loss normaliser, -inf for the raw maximum).
__global__ void __launch_bounds__(kLossThreads)
image_max_kernel(const float* __restrict__ target, int R, int slices, float floor, float* __restrict__ tx) {
    const int b = blockIdx.x / slices, s = blockIdx.x % slices;
    const size_t npix = (size_t)R * R, stride = (size_t)slices * kLossThreads;
    const float* t = target + (size_t)b * npix;
    float m = -INFINITY;
    if ((npix & 3) == 0) {
        const float4* t4 = reinterpret_cast<const float4*>(t);
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix / 4; i += stride) {
            const float4 q = __ldg(t4 + i);
            m = fmaxf(fmaxf(fmaxf(m, q.x), fmaxf(q.y, q.z)), q.w);
        }
    } else {
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix; i += stride) m = fmaxf(m, __ldg(t + i));
    }
    __shared__ float sh[32];
    __shared__ float part;
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float x = threadIdx.x < kLossThreads / 32 ? sh[threadIdx.x] : -INFINITY;
        x = warp_max(x);
        if (threadIdx.x == 0) part = x;
    }
    if (slices > 1) {
        cluster_barrier();
        if (s == 0 && threadIdx.x == 0) {
            float x = part;
            for (int r = 1; r < slices; ++r) x = fmaxf(x, ld_dsmem(&part, (uint32_t)r));
            tx[b] = fmaxf(x, floor);
        }
        cluster_barrier();   // peers keep their shared memory alive until the leader has read it
    } else if (threadIdx.x == 0) {
        tx[b] = fmaxf(part, floor);
    }
}

// per_img[b] = { sum diff^2, sum |diff| dmaps, sum |diff| },  diff = (img - target)/tx[b]
__global__ void __launch_bounds__(kLossThreads)
loss_fwd_kernel(const float* __restrict__ img, const float* __restrict__ target, const float* __restrict__ dmaps,
                const float* __restrict__ tx, int R, int slices, float* __restrict__ per_img) {
    const int b = blockIdx.x / slices, s = blockIdx.x % slices;
    const size_t npix = (size_t)R * R, off = (size_t)b * npix, stride = (size_t)slices * kLossThreads;
    const float t = fmaxf(__ldg(tx + b), 1e-6f);   // reference divides both images by tx (clamped, test_environment.py:436)
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
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix / 4; i += stride) {
            const float4 p = __ldg(a + i), q = __ldg(c + i), w = __ldg(d + i);
            one(p.x, q.x, w.x), one(p.y, q.y, w.y), one(p.z, q.z, w.z), one(p.w, q.w, w.w);
        }
    } else {
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix; i += stride)
            one(__ldg(img + off + i), __ldg(target + off + i), __ldg(dmaps + off + i));
    }
    __shared__ float sh[3 * 32];
    __shared__ float part[3];
    block_reduce<3>(acc, sh);
    if (slices > 1) {
        if (threadIdx.x == 0) part[0] = acc[0], part[1] = acc[1], part[2] = acc[2];
        cluster_barrier();
        if (s == 0 && threadIdx.x == 0) {
            for (int r = 1; r < slices; ++r) {
                acc[0] += ld_dsmem(&part[0], (uint32_t)r);
                acc[1] += ld_dsmem(&part[1], (uint32_t)r);
                acc[2] += ld_dsmem(&part[2], (uint32_t)r);
            }
        }
        cluster_barrier();
    }
    if (s == 0 && threadIdx.x == 0) {
        per_img[3 * b] = acc[0];
        per_img[3 * b + 1] = acc[1];
        per_img[3 * b + 2] = acc[2];
    }
}

// cluster size for the per-image forward kernels: enough CTAs to cover the machine twice, at most 8 per image and
// never fewer than one CTA-sized slice of float4s each
inline int loss_slices(int B, int R, int num_sms) {
    const size_t vecs = ((size_t)R * R + 3) / 4;
    int s = 1;
    while (s < 8 && (long long)B * s < 8LL * num_sms && vecs / (size_t)(2 * s) >= (size_t)kLossThreads) s *= 2;
    return s;
}

template <class Kernel, class... Args>
inline cudaError_t launch_image_clusters(Kernel kernel, int B, int slices, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((long long)B * slices));
    cfg.blockDim = dim3(kLossThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)slices;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = slices > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// packed[0..1] = sum_b per_img[b][0..1] (the numerators of mse and dist, test_environment.py:456-457).  One CTA,
// fixed summation order, fp64 accumulation: deterministic.
__global__ void __launch_bounds__(kLossThreads) loss_pack_kernel(const float* __restrict__ per_img, int B, float* __restrict__ packed) {
    double a0 = 0.0, a1 = 0.0;
    for (int b = threadIdx.x; b < B; b += kLossThreads) {
        a0 += (double)__ldg(per_img + 3 * b);
        a1 += (double)__ldg(per_img + 3 * b + 1);
    }
    __shared__ double sh[2][kLossThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = a0, sh[1][threadIdx.x >> 5] = a1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) s0 += sh[0][w], s1 += sh[1][w];
        packed[0] = (float)s0;
        packed[1] = (float)s1;
    }
}

// Fused-epilogue variant: the forward splat left P partial records per image ({sum diff^2, sum |diff| dmaps,
// sum |diff|} per epilogue warp); combine them in index order into per_img[b][3], then the batch sums as above.
__global__ void __launch_bounds__(kLossThreads)
loss_pack_partials_kernel(const float* __restrict__ partials, int P, int B, float* __restrict__ per_img, float* __restrict__ packed) {
    double a0 = 0.0, a1 = 0.0;
    for (int b = threadIdx.x; b < B; b += kLossThreads) {
        const float* pp = partials + (size_t)b * P * 3;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int k = 0; k < P; ++k) s0 += __ldg(pp + 3 * k), s1 += __ldg(pp + 3 * k + 1), s2 += __ldg(pp + 3 * k + 2);
        per_img[3 * b] = s0, per_img[3 * b + 1] = s1, per_img[3 * b + 2] = s2;
        a0 += (double)s0;
        a1 += (double)s1;
    }
    __shared__ double sh[2][kLossThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = a0, sh[1][threadIdx.x >> 5] = a1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) s0 += sh[0][w], s1 += sh[1][w];
        packed[0] = (float)s0;
        packed[1] = (float)s1;
    }
}

// Feed epilogue (splat_tc.cuh, kFuseFeed): P partial records {sum w, sum w j, sum w i} per image -> com_sums[b][3] and
// coords[b][2] exactly as com_fwd_kernel forms them (layers/center_of_mass.py:44-58).  Index order: deterministic.
__global__ void __launch_bounds__(kLossThreads)
com_pack_partials_kernel(const float* __restrict__ partials, int P, int B, float eps, float* __restrict__ coords,
                         float* __restrict__ sums) {
    const int b = blockIdx.x * kLossThreads + threadIdx.x;
    if (b >= B) return;
    const float* pp = partials + (size_t)b * P * 3;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < P; ++k) s0 += __ldg(pp + 3 * k), s1 += __ldg(pp + 3 * k + 1), s2 += __ldg(pp + 3 * k + 2);
    const bool mass = s0 > 0.f;
    coords[2 * b] = mass ? s1 / (s0 + eps) : -1.f;
    coords[2 * b + 1] = mass ? s2 / (s0 + eps) : -1.f;
    if (sums) sums[3 * b] = s0, sums[3 * b + 1] = s1, sums[3 * b + 2] = s2;
}

// g_img = (2 g0 diff + (g1 dmaps + g2) sign(diff)) / tx (+ g_img_in); {g0,g1,g2} = g_per_img[b] (may be NULL) plus the
// batch-wide {g_packed[0], g_packed[1], 0} (may be NULL): the adjoint of loss_pack_kernel folded in.
__global__ void __launch_bounds__(kLossThreads)
loss_bwd_kernel(const float* __restrict__ img, const float* __restrict__ target, const float* __restrict__ dmaps,
                const float* __restrict__ tx, const float* __restrict__ g_per_img, const float* __restrict__ g_packed,
                const float* __restrict__ g_in, int R, int slices, float* __restrict__ g_img, float* __restrict__ gmax) {
    // gmax (may be NULL): [B], zero-initialised by the caller; receives max |g_img[b]| (atomicMax on the int pattern of the
    // non-negative values, order-free) -- the per-image power-of-two scale of K3's fp16-piece operands comes from it
    const int b = blockIdx.x / slices, s = blockIdx.x % slices;
    const size_t npix = (size_t)R * R, off = (size_t)b * npix;
    const float t = fmaxf(__ldg(tx + b), 1e-6f);
    float mx = 0.f;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (g_per_img) g0 = __ldg(g_per_img + 3 * b), g1 = __ldg(g_per_img + 3 * b + 1), g2 = __ldg(g_per_img + 3 * b + 2);
    if (g_packed) g0 += __ldg(g_packed), g1 += __ldg(g_packed + 1);
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
            const float4 r = make_float4(one(p.x, q.x, w.x, z.x), one(p.y, q.y, w.y, z.y), one(p.z, q.z, w.z, z.z), one(p.w, q.w, w.w, z.w));
            o[i] = r;
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(r.x), fabsf(r.y))), fmaxf(fabsf(r.z), fabsf(r.w)));
        }
    } else {
        for (size_t i = (size_t)s * kLossThreads + threadIdx.x; i < npix; i += (size_t)slices * kLossThreads) {
            const float r = one(__ldg(img + off + i), __ldg(target + off + i), __ldg(dmaps + off + i), g_in ? __ldg(g_in + off + i) : 0.f);
            g_img[off + i] = r;
            mx = fmaxf(mx, fabsf(r));
        }
    }
    if (gmax) {
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(reinterpret_cast<int*>(gmax + b), __float_as_int(mx));
    }
}

}  // namespace helio
