// Shared device/host helpers for libhelio_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/helio_b200.h"

namespace helio {

// ---- error plumbing -----------------------------------------------------------------------
inline char* last_error_buf() {
    static thread_local char buf[512] = "";
    return buf;
}
inline int set_error(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(last_error_buf(), 512, fmt, a, b);
    return code;
}
#define HELIO_CUDA_OK(expr)                                                                    \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return helio::set_error((int)e__, "%s: %s", #expr, cudaGetErrorString(e__));       \
    } while (0)
#define HELIO_REQUIRE(cond, msg)                                                               \
    do {                                                                                       \
        if (!(cond)) return helio::set_error(HELIO_E_BADARG, "bad argument: %s (%s)", msg, #cond); \
    } while (0)

// ---- small vector helpers -----------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__host__ __device__ __forceinline__ V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ void st3(float* p, V3 a) {
    p[0] = a.x;
    p[1] = a.y;
    p[2] = a.z;
}

// Scene in kernel-argument form (passed by value; 27 floats).
struct Scene {
    V3 p, n, u, v, w;  // target_pos, unit normal, plane_u, plane_v, u x v
    float width, height, sigma_scale;
    V3 bp, bn, bu, bv;  // boundary(): targ_pos, targ_norm (raw), east, up
    float bw, bh;
};
inline Scene make_scene(const helio_scene_t* s) {
    Scene c;
    c.p = v3(s->target_pos);
    c.n = v3(s->target_normal);
    c.u = v3(s->plane_u);
    c.v = v3(s->plane_v);
    c.w = cross(c.u, c.v);
    c.width = s->width;
    c.height = s->height;
    c.sigma_scale = s->sigma_scale;
    c.bp = v3(s->bnd_targ_pos);
    c.bn = v3(s->bnd_targ_norm);
    c.bu = v3(s->bnd_u);
    c.bv = v3(s->bnd_v);
    c.bw = s->bnd_width;
    c.bh = s->bnd_height;
    return c;
}

// torch.linspace(-w/2, w/2, R)[i] as the reference evaluates it (fp32; first half fma(step,i,start),
// second half fma(-step, R-1-i, end)); newenv_rl_test_multi_error.py:129-130.
struct Axis {
    float start, end, step;
    int R, half;
};
inline Axis make_axis(float extent, int R) {
    Axis a;
    a.start = -extent / 2;
    a.end = extent / 2;
    a.step = R > 1 ? (a.end - a.start) / (float)(R - 1) : 0.f;
    a.R = R;
    a.half = R / 2;
    return a;
}
__device__ __forceinline__ float axis_at(const Axis& a, int i) {
    return i < a.half ? __fmaf_rn(a.step, (float)i, a.start) : __fmaf_rn(-a.step, (float)(a.R - 1 - i), a.end);
}

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log2 for the amplitude fold (amp ~ 1): one MUFU; a denormal amplitude counts as zero (lg2 -> -inf -> weight 0), which
// spares the producers the scale-and-correct sequence of the non-ftz form (5 instructions per heliostat per stage)
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace helio
