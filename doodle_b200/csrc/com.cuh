// Centre of mass of each image: the encoder feed of the reference's COM trainer
// (layers/center_of_mass.py:21-60, CenterOfMass2D.forward), forward and adjoint.
//
//   w = max(x, 0);  S = sum w;  X = sum w * col;  Y = sum w * row
//   coords[b] = (X, Y) / (S + eps), or (-1, -1) when S <= 0
//   dL/dx[i][j] = [x >= 0] * (g_x (j - x_com) + g_y (i - y_com)) / (S + eps)      (zero for S <= 0)
//
// HBM-bound: one pass over the image each way.  Forward runs one thread-block cluster per image (same scheme as
// the loss kernels: slice partials combined through DSMEM in rank order, deterministic).
#pragma once
#include "helio_common.cuh"
#include "loss.cuh"

namespace helio {

__global__ void __launch_bounds__(kLossThreads)
com_fwd_kernel(const float* __restrict__ img, int H, int W, int slices, float eps, float* __restrict__ coords,
               float* __restrict__ sums) {
    const int b = blockIdx.x / slices, s = blockIdx.x % slices;
    const size_t npix = (size_t)H * W;
    const float* x = img + (size_t)b * npix;
    float acc[3] = {0.f, 0.f, 0.f};
    if ((W & 3) == 0) {
        // flat float4 sweep (many independent loads in flight, as in the loss kernels); a float4 never straddles a row
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int nvec = (int)(npix / 4), wv = W / 4;
#pragma unroll 4
        for (int v = s * kLossThreads + threadIdx.x; v < nvec; v += slices * kLossThreads) {
            const float4 q = __ldg(x4 + v);
            const int i = v / wv, j = 4 * (v - i * wv);
            const float w0 = fmaxf(q.x, 0.f), w1 = fmaxf(q.y, 0.f), w2 = fmaxf(q.z, 0.f), w3 = fmaxf(q.w, 0.f);
            const float rs = (w0 + w1) + (w2 + w3);
            acc[0] += rs;
            acc[1] = fmaf(w0, (float)j, fmaf(w1, (float)(j + 1), fmaf(w2, (float)(j + 2), fmaf(w3, (float)(j + 3), acc[1]))));
            acc[2] = fmaf(rs, (float)i, acc[2]);
        }
    } else {
        // rows are dealt to (slice, warp) pairs, lanes stride over the columns
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = kLossThreads / 32;
        for (int i = s * nw + wid; i < H; i += slices * nw) {
            const float* row = x + (size_t)i * W;
            float rs = 0.f, rx = 0.f;
            for (int j = lane; j < W; j += 32) {
                const float w = fmaxf(__ldg(row + j), 0.f);
                rs += w;
                rx = fmaf(w, (float)j, rx);
            }
            acc[0] += rs;
            acc[1] += rx;
            acc[2] = fmaf(rs, (float)i, acc[2]);
        }
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
        const bool mass = acc[0] > 0.f;
        coords[2 * b] = mass ? acc[1] / (acc[0] + eps) : -1.f;
        coords[2 * b + 1] = mass ? acc[2] / (acc[0] + eps) : -1.f;
        if (sums) sums[3 * b] = acc[0], sums[3 * b + 1] = acc[1], sums[3 * b + 2] = acc[2];
    }
}

__global__ void __launch_bounds__(kLossThreads)
com_bwd_kernel(const float* __restrict__ img, const float* __restrict__ sums, const float* __restrict__ g_coords, int H, int W,
               int slices, float eps, float* __restrict__ g_img) {
    const int b = blockIdx.x / slices, s = blockIdx.x % slices;
    const size_t npix = (size_t)H * W;
    const float* x = img + (size_t)b * npix;
    float* g = g_img + (size_t)b * npix;
    const float S = __ldg(sums + 3 * b);
    const bool mass = S > 0.f;
    const float inv = mass ? 1.f / (S + eps) : 0.f;
    const float xc = __ldg(sums + 3 * b + 1) * inv, yc = __ldg(sums + 3 * b + 2) * inv;
    const float gx = __ldg(g_coords + 2 * b) * inv, gy = __ldg(g_coords + 2 * b + 1) * inv;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = kLossThreads / 32;
    for (int i = s * nw + wid; i < H; i += slices * nw) {
        const float gyi = gy * ((float)i - yc);
        for (int j = lane; j < W; j += 32) {
            const float v = __ldg(x + (size_t)i * W + j);
            g[(size_t)i * W + j] = v >= 0.f ? fmaf(gx, (float)j - xc, gyi) : 0.f;
        }
    }
}

}  // namespace helio
