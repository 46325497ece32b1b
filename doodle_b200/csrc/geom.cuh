// K1: fused per-(sun, heliostat) geometry, forward and adjoint.
//
// One thread per (b, n).  Follows the reference's fp32 operation order so that hit points agree
// to rounding (newenv_rl_test_multi_error.py:356-389, 126-127, 146); see include/helio_b200.h
// for the list of reference functions this replaces.
#pragma once
#include "helio_common.cuh"

namespace helio {

constexpr float kLeakySlope = 0.01f;  // F.leaky_relu default (newenv_rl_test_multi_error.py:369)
constexpr int kGeomThreads = 256;

struct GeomState {
    float ce, se, cu, su;  // cos/sin of the East / Up error angles
    V3 rot;                // rotated action normal (before the Up-axis guard)
    V3 v;                  // after leaky-ReLU on z
    float vn;              // max(|v|, 1e-9)
    V3 actual;             // v / vn  (returned actual_normals)
    V3 h, ihat;            // heliostat position, unit incident direction
    float an;              // max(|actual|, 1e-9)   (reflect_vectors renormalises)
    V3 aunit;
    float dots;            // -(ihat . aunit)
    V3 r;
    float rn;
    V3 rhat;  // returned reflected_rays
    float denom, num, t;
    bool valid;
    V3 P, dvec;
    float dist, sigma, two_s2;
};

__device__ __forceinline__ GeomState geom_forward(const Scene& sc, V3 h, V3 sun, V3 act, float e_east, float e_up,
                                                  bool has_err) {
    GeomState g;
    g.h = h;
    if (has_err) {
        sincosf(e_east * 1e-3f, &g.se, &g.ce);
        sincosf(e_up * 1e-3f, &g.su, &g.cu);
    } else {
        g.se = g.su = 0.f;
        g.ce = g.cu = 1.f;
    }
    // rotate_normals_batch :93-104  (Rz(up) then Rx(east))
    float x_u = g.cu * act.x - g.su * act.y;
    float y_u = g.su * act.x + g.cu * act.y;
    float y_e = g.ce * y_u - g.se * act.z;
    float z_e = g.se * y_u + g.ce * act.z;
    g.rot = v3(x_u, y_e, z_e);
    // :369-372
    g.v = v3(x_u, y_e, z_e > 0.f ? z_e : z_e * kLeakySlope);
    g.vn = fmaxf(norm(g.v), 1e-9f);
    g.actual = v3(g.v.x / g.vn, g.v.y / g.vn, g.v.z / g.vn);
    // :377-380
    V3 inc = sun - h;
    float inn = fmaxf(norm(inc), 1e-9f);
    g.ihat = v3(inc.x / inn, inc.y / inn, inc.z / inn);
    // reflect_vectors :46-50
    g.an = fmaxf(norm(g.actual), 1e-9f);
    g.aunit = v3(g.actual.x / g.an, g.actual.y / g.an, g.actual.z / g.an);
    g.dots = -dot(g.ihat, g.aunit);
    g.r = -g.ihat - (2.f * g.dots) * g.aunit;
    g.rn = fmaxf(norm(g.r), 1e-9f);
    g.rhat = v3(g.r.x / g.rn, g.r.y / g.rn, g.r.z / g.rn);
    // ray_plane_intersection_batch :60-75  (sc.n is already unit)
    g.denom = dot(g.rhat, sc.n);
    g.valid = fabsf(g.denom) > 1e-9f;
    float safe_denom = g.valid ? g.denom : 1e-9f;
    g.num = dot(sc.p - h, sc.n);
    g.t = g.valid ? g.num / safe_denom : 0.f;
    V3 inter = v3(h.x + g.t * g.rhat.x, h.y + g.t * g.rhat.y, h.z + g.t * g.rhat.z);
    g.P = g.valid ? inter : v3(0.f, 0.f, 0.f);
    // gaussian_blur_batch :126-127, :146
    g.dvec = g.P - h;
    g.dist = norm(g.dvec);
    g.sigma = fmaxf(sc.sigma_scale * g.dist, 1e-9f);
    g.two_s2 = fmaxf(2.f * g.sigma * g.sigma, 1e-12f);
    return g;
}

// boundary(return_all=True), test_environment.py:101-130.  Returns out; if grad != nullptr also
// d out / d vect.
__device__ __forceinline__ float boundary_one(const Scene& sc, V3 h, V3 vct, V3* grad) {
    const float tol = 0.75f, eps = 1e-6f;
    float dots = -dot(vct, sc.bn);
    bool valid = fabsf(dots) > eps;
    float den = dots + (valid ? 0.f : eps);
    float pv = dot(sc.bp, vct);
    float t = pv / den;
    V3 inter = v3(h.x + vct.x * t, h.y + vct.y * t, h.z + vct.z * t);
    V3 local = inter - sc.bp;
    float xl = dot(local, sc.bu), yl = dot(local, sc.bv);
    float hw = sc.bw * tol / 2.f, hh = sc.bh * tol / 2.f;
    float ax = fabsf(xl) - hw * tol, ay = fabsf(yl) - hh * tol;
    // F.relu propagates NaN (fmaxf would drop it): a NaN action must reach the NaN asserts of step (:495-501)
    float dx = ax > 0.f ? ax : (ax != ax ? ax : 0.f), dy = ay > 0.f ? ay : (ay != ay ? ay : 0.f);
    float dist = sqrtf(dx * dx + dy * dy + 1e-8f);
    bool inside = (fabsf(xl) <= hw) && (fabsf(yl) <= hh) && valid;
    float outm = inside ? 0.f : 1.f;
    if (grad) {
        float g_dx = (ax > 0.f) ? outm * dx / dist : 0.f;
        float g_dy = (ay > 0.f) ? outm * dy / dist : 0.f;
        float g_xl = g_dx * (xl > 0.f ? 1.f : (xl < 0.f ? -1.f : 0.f));
        float g_yl = g_dy * (yl > 0.f ? 1.f : (yl < 0.f ? -1.f : 0.f));
        V3 g_inter = g_xl * sc.bu + g_yl * sc.bv;
        float g_t = dot(g_inter, vct);
        float g_pv = g_t / den;
        float g_den = -g_t * pv / (den * den);
        *grad = t * g_inter + g_pv * sc.bp - g_den * sc.bn;
    }
    return dist * outm;
}

// calculate_angles_mrad, test_environment.py:132-155 (fp32: clamp to +-nextafter(1,0)).
__device__ __forceinline__ float angle_mrad(V3 ideal, V3 actual, float* g_dot) {
    const float hi = 0.99999994f, lo = -0.99999994f;
    float d = dot(ideal, actual);
    float c = d != d ? d : fminf(fmaxf(d, lo), hi);    // torch.clamp propagates NaN
    if (g_dot) *g_dot = (d >= lo && d <= hi) ? -1000.f / sqrtf(1.f - c * c) : 0.f;
    return acosf(c) * 1000.f;
}

// calculate_ideal_normals, newenv_rl_test_multi_error.py:270-278
__device__ __forceinline__ V3 ideal_normal(const Scene& sc, V3 h, V3 ihat) {
    V3 refl = sc.p - h;
    float rn = fmaxf(norm(refl), 1e-9f);
    V3 s = v3(ihat.x + refl.x / rn, ihat.y + refl.y / rn, ihat.z + refl.z / rn);
    float sn = fmaxf(norm(s), 1e-9f);
    return v3(s.x / sn, s.y / sn, s.z / sn);
}

struct GeomWorkspace {  // layout of the caller-provided workspace
    unsigned int ticket;
    unsigned int pad[3];
    // followed by float partials[nblocks][2]
};

__global__ void __launch_bounds__(kGeomThreads)
geom_fwd_kernel(Scene sc, const float* __restrict__ helio, const float* __restrict__ sun,
                const float* __restrict__ action, const float* __restrict__ errs, int B, int N,
                float4* __restrict__ params, float* __restrict__ actual, float* __restrict__ refl,
                float* __restrict__ ideal, float* __restrict__ bounds, float* __restrict__ angles,
                float* __restrict__ sums, GeomWorkspace* ws) {
    const long long M = (long long)B * N;
    const long long idx = (long long)blockIdx.x * kGeomThreads + threadIdx.x;
    float bnd = 0.f, ang = 0.f;
    if (idx < M) {
        const int b = (int)(idx / N), n = (int)(idx - (long long)b * N);
        V3 h = ld3(helio + 3 * n), s = ld3(sun + 3 * b), a;
        if (action) {
            a = ld3(action + 3 * idx);
        } else {
            // no action: aim with the ideal normal (the target render of HelioEnv.step, test_environment.py:429-433);
            // same arithmetic as the ideal-normal output below, so it equals feeding that output back in
            V3 inc = s - h;
            float inn = fmaxf(norm(inc), 1e-9f);
            a = ideal_normal(sc, h, v3(inc.x / inn, inc.y / inn, inc.z / inn));
        }
        float e0 = 0.f, e1 = 0.f;
        if (errs) {
            float2 e = __ldg(reinterpret_cast<const float2*>(errs) + idx);
            e0 = e.x;
            e1 = e.y;
        }
        GeomState g = geom_forward(sc, h, s, a, e0, e1, errs != nullptr);
        // footprint parameters in receiver coordinates
        V3 d = g.P - sc.p;
        float fa = dot(d, sc.u), fb = dot(d, sc.v), fc = dot(d, sc.w);
        float k2 = g.valid ? kLog2e / g.two_s2 : 0.f;
        float amp = g.valid ? exp2f(-(fc * fc) * k2) : 1.f;
        params[idx] = make_float4(g.valid ? fa : 0.f, g.valid ? fb : 0.f, k2, amp);
        st3(actual + 3 * idx, g.actual);
        st3(refl + 3 * idx, g.rhat);
        if (ideal || angles || sums) {
            V3 id = ideal_normal(sc, h, g.ihat);
            if (ideal) st3(ideal + 3 * idx, id);
            ang = angle_mrad(id, g.actual, nullptr);
            if (angles) angles[idx] = ang;
        }
        if (bounds || sums) {
            bnd = boundary_one(sc, h, a, nullptr);
            if (bounds) bounds[idx] = bnd;
        }
    }
    if (sums == nullptr) return;
    // warp-shuffle reduction, then an ordered reduction over blocks by the last block to finish
    __shared__ float sb[kGeomThreads / 32], sa[kGeomThreads / 32];
    __shared__ bool is_last;
    float wb = warp_sum(bnd), wa = warp_sum(ang);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sb[wid] = wb;
        sa[wid] = wa;
    }
    __syncthreads();
    float* partials = reinterpret_cast<float*>(ws + 1);
    if (threadIdx.x == 0) {
        float tb = 0.f, ta = 0.f;
#pragma unroll
        for (int w = 0; w < kGeomThreads / 32; ++w) {
            tb += sb[w];
            ta += sa[w];
        }
        partials[2 * blockIdx.x] = tb;
        partials[2 * blockIdx.x + 1] = ta;
        __threadfence();
        unsigned int t = atomicAdd(&ws->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double accb = 0.0, acca = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kGeomThreads) {
        accb += (double)__ldcg(partials + 2 * i);
        acca += (double)__ldcg(partials + 2 * i + 1);
    }
    __shared__ double db[kGeomThreads], da[kGeomThreads];
    db[threadIdx.x] = accb;
    da[threadIdx.x] = acca;
    __syncthreads();
    for (int s = kGeomThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            db[threadIdx.x] += db[threadIdx.x + s];
            da[threadIdx.x] += da[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[0] = (float)db[0];
        sums[1] = (float)da[0];
        ws->ticket = 0;  // leave the workspace zeroed for the next launch
    }
}

__global__ void __launch_bounds__(kGeomThreads)
geom_bwd_kernel(Scene sc, const float* __restrict__ helio, const float* __restrict__ sun,
                const float* __restrict__ action, const float* __restrict__ errs, int B, int N,
                const float4* __restrict__ g_moments, const float* __restrict__ g_actual,
                const float* __restrict__ g_refl, const float* __restrict__ g_bounds,
                const float* __restrict__ g_angles, const float* __restrict__ g_sums,
                float* __restrict__ g_action) {
    const long long M = (long long)B * N;
    const long long idx = (long long)blockIdx.x * kGeomThreads + threadIdx.x;
    if (idx >= M) return;
    const int b = (int)(idx / N), n = (int)(idx - (long long)b * N);
    V3 h = ld3(helio + 3 * n), s = ld3(sun + 3 * b), a = ld3(action + 3 * idx);
    float e0 = 0.f, e1 = 0.f;
    if (errs) {
        float2 e = __ldg(reinterpret_cast<const float2*>(errs) + idx);
        e0 = e.x;
        e1 = e.y;
    }
    GeomState g = geom_forward(sc, h, s, a, e0, e1, errs != nullptr);

    // ---- image term: moments -> dL/dP, dL/d(two_s2)   (see include/helio_b200.h, K3) ----
    V3 gP = v3(0.f, 0.f, 0.f);
    if (g_moments && g.valid) {
        float4 m = __ldg(g_moments + idx);
        V3 d = g.P - sc.p;
        float fc = dot(d, sc.w);
        float inv = 1.f / g.two_s2;
        float ga = 2.f * m.y * inv, gb = 2.f * m.z * inv, gc = -2.f * fc * m.x * inv;
        gP = ga * sc.u + gb * sc.v + gc * sc.w;
        float g_ts2 = (m.w + fc * fc * m.x) * inv * inv;
        // two_s2 = max(2 sigma^2, 1e-12); sigma = max(scale*dist, 1e-9); dist = |P - h|
        float g_sigma = (2.f * g.sigma * g.sigma >= 1e-12f) ? g_ts2 * 4.f * g.sigma : 0.f;
        float g_dist = (sc.sigma_scale * g.dist >= 1e-9f) ? g_sigma * sc.sigma_scale : 0.f;
        if (g.dist > 0.f) gP = gP + (g_dist / g.dist) * g.dvec;
    }
    // P = where(valid, h + t rhat, 0)
    V3 g_rhat = g.t * gP;  // t == 0 when invalid
    if (g_refl) g_rhat = g_rhat + ld3(g_refl + 3 * idx);
    if (g.valid) {
        float g_t = dot(gP, g.rhat);
        float g_denom = -g_t * g.num / (g.denom * g.denom);
        g_rhat = g_rhat + g_denom * sc.n;
    }
    // rhat = r / max(|r|, 1e-9)
    V3 g_r = (1.f / g.rn) * (g_rhat - dot(g_rhat, g.rhat) * g.rhat);
    // r = -i - 2 dots aunit ; dots = -(i . aunit)
    float g_dots = -2.f * dot(g_r, g.aunit);
    V3 g_aunit = (-2.f * g.dots) * g_r - g_dots * g.ihat;
    // aunit = actual / max(|actual|, 1e-9)
    V3 g_act = (1.f / g.an) * (g_aunit - dot(g_aunit, g.aunit) * g.aunit);
    if (g_actual) g_act = g_act + ld3(g_actual + 3 * idx);
    // alignment: angle = 1000 acos(clamp(ideal . actual))
    float w_ang = (g_angles ? __ldg(g_angles + idx) : 0.f) + (g_sums ? __ldg(g_sums + 1) : 0.f);
    if (w_ang != 0.f) {
        V3 id = ideal_normal(sc, h, g.ihat);
        float gd;
        angle_mrad(id, g.actual, &gd);
        g_act = g_act + (w_ang * gd) * id;
    }
    // actual = v / max(|v|, 1e-9)
    V3 g_v = (1.f / g.vn) * (g_act - dot(g_act, g.actual) * g.actual);
    // v.z = leaky_relu(rot.z)
    V3 g_rot = v3(g_v.x, g_v.y, g_v.z * (g.rot.z > 0.f ? 1.f : kLeakySlope));
    // rot = Rx(e_east) Rz(e_up) action  -> transpose
    float gy_u = g.ce * g_rot.y + g.se * g_rot.z;
    float gz = -g.se * g_rot.y + g.ce * g_rot.z;
    float gx = g.cu * g_rot.x + g.su * gy_u;
    float gy = -g.su * g_rot.x + g.cu * gy_u;
    V3 out = v3(gx, gy, gz);
    // boundary(action)
    float w_bnd = (g_bounds ? __ldg(g_bounds + idx) : 0.f) + (g_sums ? __ldg(g_sums) : 0.f);
    if (w_bnd != 0.f) {
        V3 gb;
        boundary_one(sc, h, a, &gb);
        out = out + w_bnd * gb;
    }
    st3(g_action + 3 * idx, out);
}

}  // namespace helio
