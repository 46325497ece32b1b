// extern "C" entry points of libhelio_sm100.so (see include/helio_b200.h for the contract).
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "com.cuh"
#include "cull.cuh"
#include "edt.cuh"
#include "geom.cuh"
#include "loss.cuh"
#include "splat_simt.cuh"
#include "splat_tc.cuh"

using namespace helio;

#ifndef HELIO_BWD_PREC_DEFAULT
#define HELIO_BWD_PREC_DEFAULT 1
#endif

namespace {

struct DeviceInfo {
    int ok = 0;
    int sms = 0;
    int cc_major = 0, cc_minor = 0;
};

// per-device capability cache (one process per GPU is the deployment model, but stay correct
// if a process touches several devices)
const DeviceInfo* device_info() {
    static DeviceInfo info[64];
    static std::once_flag once[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::call_once(once[dev], [dev]() {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return;
        info[dev].sms = p.multiProcessorCount;
        info[dev].cc_major = p.major;
        info[dev].cc_minor = p.minor;
        info[dev].ok = (p.major == 10);
    });
    return &info[dev];
}

int require_device(const DeviceInfo** out) {
    const DeviceInfo* d = device_info();
    if (d == nullptr) return set_error(HELIO_E_NOSM100, "no usable CUDA device (%s)%s", cudaGetErrorString(cudaGetLastError()));
    if (!d->ok) {
        char cc[32];
        snprintf(cc, sizeof cc, "%d.%d", d->cc_major, d->cc_minor);
        return set_error(HELIO_E_NOSM100, "libhelio_sm100 is built for sm_100a only; device is cc %s%s", cc);
    }
    *out = d;
    return 0;
}

// tcgen05 kernels: 0 = auto (CTA pairs, cta_group::2, where the shape allows), 1 = single-CTA only, 2 = force pairs.
// Initial value from HELIO_TC_PAIR; helio_set_tc_pair_mode() changes it at run time (A/B tests, parity tests).
std::atomic<int>& tc_pair_state() {
    static std::atomic<int> mode{[]() {
        const char* e = std::getenv("HELIO_TC_PAIR");
        const int v = e ? std::atoi(e) : 0;
        return (v == 1 || v == 2) ? v : 0;
    }()};
    return mode;
}
int tc_pair_mode() { return tc_pair_state().load(std::memory_order_relaxed); }

// forward splat operand format: 0 = 3xTF32 everywhere, 1 = f16x3 everywhere (two fp16 pieces of the 2^14-scaled Gaussians),
// 2 = auto (default) = f16x3.  Both carry 22 significant bits per operand; measured error against fp64 of f16x3 is not
// larger than 3xTF32's at any tested shape (test_forward_f16x3_...), it halves the tensor work (the B200 is power-limited
// inside these kernels) and its half-size stages deepen the ring: forward -3 ... -12 % at R <= 128, -10 ... -14 % at R = 256.
// 3xTF32 -- the format BASELINE.json names -- stays selectable (mode 0) and is what bench.py reports as `both_3xtf32`.
std::atomic<int>& fwd_prec_state() {
    static std::atomic<int> mode{[]() {
        const char* e = std::getenv("HELIO_FWD_PREC");
        const int v = e ? std::atoi(e) : 2;
        return (v >= 0 && v <= 2) ? v : 2;
    }()};
    return mode;
}
// backward splat operand format inside helio_step_bwd: 0 = 3xTF32, 1 = f16x3 with K = 64 per stage (see SplatBwdTc).
// The f16x3 backward needs the per-image maximum of |dL/dimg|, which the loss backward produces on the way; the standalone
// helio_splat_bwd (arbitrary g_img, no scratch) always uses 3xTF32.
std::atomic<int>& bwd_prec_state() {
    static std::atomic<int> mode{[]() {
        const char* e = std::getenv("HELIO_BWD_PREC");
        const int v = e ? std::atoi(e) : HELIO_BWD_PREC_DEFAULT;
        return (v == 0 || v == 1) ? v : HELIO_BWD_PREC_DEFAULT;
    }()};
    return mode;
}
int fwd_prec_for(int R) {
    const int m = fwd_prec_state().load(std::memory_order_relaxed);
    (void)R;
    return m == 2 ? 1 : m;
}

// ---- opt-in per-kernel timing (helio_profile_*): CUDA events recorded around every kernel this library
// enqueues, on the stream it is enqueued on.  Off by default; not for use under stream capture.
struct ProfRecord {
    const char* name;
    cudaEvent_t e0, e1;
};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRecord> g_prof;

struct KernelTimer {
    cudaStream_t st;
    ProfRecord rec{nullptr, nullptr, nullptr};
    KernelTimer(const char* name, void* stream) : st((cudaStream_t)stream) {
        if (!g_prof_on.load(std::memory_order_relaxed)) return;
        if (cudaEventCreate(&rec.e0) != cudaSuccess || cudaEventCreate(&rec.e1) != cudaSuccess) return;
        rec.name = name;
        cudaEventRecord(rec.e0, st);
    }
    ~KernelTimer() {
        if (!rec.name) return;
        cudaEventRecord(rec.e1, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(rec);
    }
};

inline unsigned geom_blocks(int B, int N) { return (unsigned)(((long long)B * N + kGeomThreads - 1) / kGeomThreads); }

// tx[b] = floor: the fused per-image maximum (atomicMax on the int pattern of non-negative floats) then yields
// max(max_ij target, floor) directly, the same clamped value helio_image_max returns (test_environment.py:436)
__global__ void fill_kernel(float* __restrict__ p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

extern "C" {

HELIO_API int helio_abi_version(void) { return HELIO_ABI_VERSION; }

HELIO_API const char* helio_last_error(void) { return last_error_buf(); }

HELIO_API int helio_device_ok(void) {
    const DeviceInfo* d = nullptr;
    return require_device(&d) == 0 ? 1 : 0;
}

HELIO_API int helio_set_tc_pair_mode(int mode) {
    if (mode < 0 || mode > 2) return set_error(HELIO_E_BADARG, "bad argument: %s%s", "tc pair mode must be 0, 1 or 2");
    tc_pair_state().store(mode, std::memory_order_relaxed);
    return 0;
}

HELIO_API int helio_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    g_prof_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return 0;
}

HELIO_API int helio_profile_count(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    return (int)g_prof.size();
}

HELIO_API int helio_profile_get(int index, const char** name, float* ms) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    HELIO_REQUIRE(index >= 0 && index < (int)g_prof.size() && name && ms, "profile record index");
    const ProfRecord& r = g_prof[index];
    HELIO_CUDA_OK(cudaEventSynchronize(r.e1));
    HELIO_CUDA_OK(cudaEventElapsedTime(ms, r.e0, r.e1));
    *name = r.name;
    return 0;
}

#if HELIO_TC_STATS
namespace {
int splat_bwd_impl(const float* params, const float* g_img, int B, int N, int R, float width, float height, float* moments,
                   int impl, void* stream, const int* counts, const int* index, const float* gmax);
}
// debug builds only: the f16x3 backward splat on its own (gmax[B] = per-image max |g_img|, computed by the caller)
HELIO_API int helio_debug_splat_bwd_f16(const float* params, const float* g_img, const float* gmax, int B, int N, int R, float width,
                                        float height, float* moments, void* stream) {
    return splat_bwd_impl(params, g_img, B, N, R, width, height, moments, HELIO_SPLAT_TC, stream, nullptr, nullptr, gmax);
}
// debug builds only (-DHELIO_TC_STATS=1, scripts/tc_stats.py): copy out / clear the per-warp cycle counters of the last
// tcgen05 kernel.  out: [160][24][4] uint64.
HELIO_API int helio_debug_tc_stats(unsigned long long* out_host, int clear) {
    HELIO_CUDA_OK(cudaDeviceSynchronize());
    if (out_host) HELIO_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_tc_stats, sizeof(g_tc_stats)));
    if (clear) {
        void* p = nullptr;
        HELIO_CUDA_OK(cudaGetSymbolAddress(&p, g_tc_stats));
        HELIO_CUDA_OK(cudaMemset(p, 0, sizeof(g_tc_stats)));
    }
    return 0;
}
#endif

HELIO_API int helio_tc_clock_mhz(int which, float* mhz_host) {
    HELIO_REQUIRE((which == 0 || which == 1) && mhz_host, "which must be 0 (forward) or 1 (backward)");
    unsigned long long v[2][2];
    HELIO_CUDA_OK(cudaMemcpyFromSymbol(v, g_tc_clock, sizeof(v)));     // synchronises with the device
    *mhz_host = v[which][1] ? (float)((double)v[which][0] * 1e3 / (double)v[which][1]) : 0.f;
    return 0;
}

HELIO_API int helio_set_bwd_precision(int mode) {
    if (mode != 0 && mode != 1) return set_error(HELIO_E_BADARG, "bad argument: %s%s", "backward precision mode must be 0 or 1");
    bwd_prec_state().store(mode, std::memory_order_relaxed);
    return 0;
}

HELIO_API int helio_set_fwd_precision(int mode) {
    if (mode < 0 || mode > 2) return set_error(HELIO_E_BADARG, "bad argument: %s%s", "forward precision mode must be 0, 1 or 2");
    fwd_prec_state().store(mode, std::memory_order_relaxed);
    return 0;
}

HELIO_API int64_t helio_geom_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (int64_t)sizeof(GeomWorkspace) + (int64_t)geom_blocks(B, N) * 2 * sizeof(float);
}

HELIO_API int helio_geom_fwd(const helio_scene_t* scene, const float* helio_pos, const float* sun, const float* action,
                   const float* errs, int B, int N, float* params, float* actual, float* refl, float* ideal,
                   float* bounds, float* angles, float* sums, void* workspace, int64_t workspace_bytes, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(scene && helio_pos && sun && params && actual && refl, "null pointer");   // action may be NULL (ideal aim)
    HELIO_REQUIRE(B > 0 && N > 0, "B, N must be positive");
    if (sums) {
        HELIO_REQUIRE(workspace != nullptr, "sums requested without workspace");
        if (workspace_bytes < helio_geom_workspace_bytes(B, N))
            return set_error(HELIO_E_WORKSPACE, "geom workspace too small%s%s");
    }
    KernelTimer timer("geom_fwd", stream);
    geom_fwd_kernel<<<geom_blocks(B, N), kGeomThreads, 0, (cudaStream_t)stream>>>(
        make_scene(scene), helio_pos, sun, action, errs, B, N, reinterpret_cast<float4*>(params), actual, refl, ideal,
        bounds, angles, sums, reinterpret_cast<GeomWorkspace*>(workspace));
    HELIO_CUDA_OK(cudaGetLastError());
    return 0;
}

HELIO_API int helio_geom_bwd(const helio_scene_t* scene, const float* helio_pos, const float* sun, const float* action,
                   const float* errs, int B, int N, const float* g_moments, const float* g_actual, const float* g_refl,
                   const float* g_bounds, const float* g_angles, const float* g_sums, float* g_action, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(scene && helio_pos && sun && action && g_action, "null pointer");
    HELIO_REQUIRE(B > 0 && N > 0, "B, N must be positive");
    KernelTimer timer("geom_bwd", stream);
    geom_bwd_kernel<<<geom_blocks(B, N), kGeomThreads, 0, (cudaStream_t)stream>>>(
        make_scene(scene), helio_pos, sun, action, errs, B, N, reinterpret_cast<const float4*>(g_moments), g_actual,
        g_refl, g_bounds, g_angles, g_sums, g_action);
    HELIO_CUDA_OK(cudaGetLastError());
    return 0;
}

namespace {
bool splat_fwd_uses_tc(int impl, int B, int N, int R) {
    const bool tc_ok = splat_tc_fwd_supported(B, N, R);
    return (impl == HELIO_SPLAT_TC && tc_ok) || (impl == HELIO_SPLAT_AUTO && tc_ok && splat_tc_fwd_preferred(B, N, R));
}

int splat_fwd_impl(const float* params, int B, int N, int R, float width, float height, float* img, int impl, void* stream,
                   int fuse, const FwdFuse& fz, const int* counts = nullptr) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(params && img, "null pointer");
    HELIO_REQUIRE(B > 0 && N > 0 && R > 0, "B, N, R must be positive");
    HELIO_REQUIRE(impl >= HELIO_SPLAT_AUTO && impl <= HELIO_SPLAT_TC, "unknown impl");
    KernelTimer timer("splat_fwd", stream);
    if (impl == HELIO_SPLAT_TC && !splat_tc_fwd_supported(B, N, R))
        return set_error(HELIO_E_BADARG, "tcgen05 splat forward does not support this shape%s%s");
    if (splat_fwd_uses_tc(impl, B, N, R)) {
        static const int split = []() {   // tuning / A-B switch: producer warps per operand slab (0 = auto)
            const char* e = std::getenv("HELIO_TC_FWD_SPLIT");
            return e ? std::atoi(e) : 0;
        }();
        HELIO_CUDA_OK(splat_tc_fwd(params, img, B, N, R, width, height, d->sms, (cudaStream_t)stream, tc_pair_mode(), split, fuse, fz, counts,
                                   fwd_prec_for(R)));
    } else {
        HELIO_REQUIRE(fuse == kFuseNone && counts == nullptr, "epilogue fusion / culled input need the tcgen05 path");
        HELIO_CUDA_OK(splat_fwd_simt(params, img, B, N, R, width, height, d->sms, (cudaStream_t)stream));
    }
    return 0;
}
}  // namespace

HELIO_API int helio_splat_fwd(const float* params, int B, int N, int R, float width, float height, float* img, int impl,
                    void* stream) {
    return splat_fwd_impl(params, B, N, R, width, height, img, impl, stream, kFuseNone, FwdFuse{});
}

namespace {
bool splat_bwd_uses_tc(int impl, int B, int N, int R) {
    const bool tc_ok = splat_tc_bwd_supported(B, N, R);
    return (impl == HELIO_SPLAT_TC && tc_ok) || (impl == HELIO_SPLAT_AUTO && tc_ok && splat_tc_bwd_preferred(B, N, R));
}

int splat_bwd_impl(const float* params, const float* g_img, int B, int N, int R, float width, float height, float* moments,
                   int impl, void* stream, const int* counts = nullptr, const int* index = nullptr, const float* gmax = nullptr) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(params && g_img && moments, "null pointer");
    HELIO_REQUIRE(B > 0 && N > 0 && R > 0, "B, N, R must be positive");
    HELIO_REQUIRE(impl >= HELIO_SPLAT_AUTO && impl <= HELIO_SPLAT_TC, "unknown impl");
    KernelTimer timer("splat_bwd", stream);
    if (impl == HELIO_SPLAT_TC && !splat_tc_bwd_supported(B, N, R))
        return set_error(HELIO_E_BADARG, "tcgen05 splat backward does not support this shape%s%s");
    if (splat_bwd_uses_tc(impl, B, N, R)) {
        if (counts) HELIO_CUDA_OK(cudaMemsetAsync(moments, 0, (size_t)B * N * 16, (cudaStream_t)stream));   // culled heliostats: zero
        HELIO_CUDA_OK(splat_tc_bwd(params, g_img, moments, B, N, R, width, height, d->sms, (cudaStream_t)stream, tc_pair_mode(),
                                   counts, index, gmax));
    } else {
        HELIO_REQUIRE(counts == nullptr, "culled input needs the tcgen05 path");
        HELIO_CUDA_OK(splat_bwd_simt(params, g_img, moments, B, N, R, width, height, d->sms, (cudaStream_t)stream));
    }
    return 0;
}
}  // namespace

HELIO_API int helio_splat_bwd(const float* params, const float* g_img, int B, int N, int R, float width, float height,
                    float* moments, int impl, void* stream) {
    return splat_bwd_impl(params, g_img, B, N, R, width, height, moments, impl, stream);
}

namespace {
// noisy splat with the encoder feed: fused into the tcgen05 epilogue, or (CUDA-core shapes) splat + helio_com_fwd + a strided copy
int splat_fwd_feed_impl(const float* params, int B, int N, int R, float width, float height, float* img, int impl,
                        const helio_feed_t* feed, void* stream, const int* counts = nullptr) {
    HELIO_REQUIRE(feed != nullptr, "null pointer");
    HELIO_REQUIRE((feed->com_coords == nullptr) == (feed->com_sums == nullptr), "com_coords and com_sums go together");
    HELIO_REQUIRE(feed->img2 == nullptr || feed->img2_batch_stride >= (int64_t)R * R, "img2 batch stride smaller than an image");
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    if (splat_fwd_uses_tc(impl, B, N, R) && feed->partials != nullptr) {
        if (feed->partials_floats < (int64_t)B * splat_tc_fwd_partials_per_image(R, d->sms, tc_pair_mode()) * 3)
            return set_error(HELIO_E_WORKSPACE, "feed partials buffer too small%s%s");
        FwdFuse fz{};
        fz.partials = feed->partials, fz.img2 = feed->img2, fz.img2_bstride = (long long)feed->img2_batch_stride;
        if (int rc = splat_fwd_impl(params, B, N, R, width, height, img, impl, stream, kFuseFeed, fz, counts)) return rc;
        if (feed->com_coords) {
            KernelTimer timer("com_pack", stream);
            com_pack_partials_kernel<<<(B + kLossThreads - 1) / kLossThreads, kLossThreads, 0, (cudaStream_t)stream>>>(
                feed->partials, splat_tc_fwd_partials_per_image(R, d->sms, tc_pair_mode()), B, feed->eps, feed->com_coords, feed->com_sums);
            HELIO_CUDA_OK(cudaGetLastError());
        }
        return 0;
    }
    if (int rc = splat_fwd_impl(params, B, N, R, width, height, img, impl, stream, kFuseNone, FwdFuse{}, counts)) return rc;
    if (feed->com_coords)
        if (int rc = helio_com_fwd(img, B, R, R, feed->eps, feed->com_coords, feed->com_sums, stream)) return rc;
    if (feed->img2)
        HELIO_CUDA_OK(cudaMemcpy2DAsync(feed->img2, (size_t)feed->img2_batch_stride * 4, img, (size_t)R * R * 4, (size_t)R * R * 4, (size_t)B,
                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
}  // namespace

HELIO_API int helio_splat_fwd_feed(const float* params, int B, int N, int R, float width, float height, float* img, int impl,
                         const helio_feed_t* feed, void* stream) {
    return splat_fwd_feed_impl(params, B, N, R, width, height, img, impl, feed, stream);
}

HELIO_API int64_t helio_cull_workspace_bytes(int B, int N) { return cull_workspace_bytes(B, N); }

HELIO_API int helio_cull(const float* params, int B, int N, float width, float height, void* workspace, int64_t workspace_bytes,
               void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(params && workspace, "null pointer");
    HELIO_REQUIRE(B > 0 && N > 0, "B, N must be positive");
    if (workspace_bytes < cull_workspace_bytes(B, N)) return set_error(HELIO_E_WORKSPACE, "cull workspace too small%s%s");
    KernelTimer timer("cull", stream);
    cull_kernel<<<B, kCullThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(params), N, 0.5f * width, 0.5f * height,
                                                              cull_carve(workspace, B, N));
    HELIO_CUDA_OK(cudaGetLastError());
    return 0;
}

HELIO_API int helio_splat_fwd_culled(const void* cull_workspace, int B, int N, int R, float width, float height, float* img,
                           void* stream) {
    HELIO_REQUIRE(cull_workspace, "null pointer");
    const CullBuffers c = cull_carve(const_cast<void*>(cull_workspace), B, N);
    return splat_fwd_impl(reinterpret_cast<const float*>(c.cparams), B, N, R, width, height, img, HELIO_SPLAT_TC, stream, kFuseNone,
                          FwdFuse{}, c.counts);
}

HELIO_API int helio_splat_bwd_culled(const void* cull_workspace, const float* g_img, int B, int N, int R, float width, float height,
                           float* moments, void* stream) {
    HELIO_REQUIRE(cull_workspace, "null pointer");
    const CullBuffers c = cull_carve(const_cast<void*>(cull_workspace), B, N);
    return splat_bwd_impl(reinterpret_cast<const float*>(c.cparams), g_img, B, N, R, width, height, moments, HELIO_SPLAT_TC, stream,
                          c.counts, c.index);
}

HELIO_API int helio_image_max(const float* target, int B, int R, float* tx, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(target && tx, "null pointer");
    HELIO_REQUIRE(B > 0 && R > 0, "B, R must be positive");
    const int slices = loss_slices(B, R, d->sms);
    KernelTimer timer("image_max", stream);
    HELIO_CUDA_OK(launch_image_clusters(image_max_kernel, B, slices, (cudaStream_t)stream, target, R, slices, 1e-6f, tx));
    return 0;
}

HELIO_API int64_t helio_distance_maps_workspace_bytes(int B, int R) {
    if (B <= 0 || R <= 0) return 0;
    return (((int64_t)B * 4 + 255) / 256) * 256 + (int64_t)B * R * R * 2;
}

HELIO_API int helio_distance_maps(const float* img, int B, int R, float thr, float* dmaps, void* workspace,
                        int64_t workspace_bytes, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(img && dmaps && workspace, "null pointer");
    HELIO_REQUIRE(B > 0 && R > 0 && R <= kEdtMaxR, "B, R must be positive (R <= 4096)");
    if (workspace_bytes < helio_distance_maps_workspace_bytes(B, R)) return set_error(HELIO_E_WORKSPACE, "edt workspace too small%s%s");
    float* mx = reinterpret_cast<float*>(workspace);
    short* g = reinterpret_cast<short*>(reinterpret_cast<char*>(workspace) + (((int64_t)B * 4 + 255) / 256) * 256);
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer timer("edt_max", stream);
        const int slices = loss_slices(B, R, d->sms);
        HELIO_CUDA_OK(launch_image_clusters(image_max_kernel, B, slices, st, img, R, slices, -INFINITY, mx));
    }
    {
        KernelTimer timer("edt_cols", stream);
        edt_cols_kernel<<<(unsigned)(((long long)B * R + kEdtThreads - 1) / kEdtThreads), kEdtThreads, 0, st>>>(img, mx, thr, B, R, g);
        HELIO_CUDA_OK(cudaGetLastError());
    }
    {
        KernelTimer timer("edt_rows", stream);
        constexpr int warps = kEdtThreads / 32;
        const size_t smem = (size_t)warps * (R + (R + kEdtSeg - 1) / kEdtSeg) * sizeof(int);
        HELIO_CUDA_OK(cudaFuncSetAttribute(edt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edt_rows_kernel<<<(unsigned)(((long long)B * R + warps - 1) / warps), kEdtThreads, smem, st>>>(g, B, R, dmaps);
        HELIO_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

HELIO_API int helio_loss_fwd(const float* img, const float* target, const float* dmaps, const float* tx, int B, int R,
                   float* per_img, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(img && target && dmaps && tx && per_img, "null pointer");
    HELIO_REQUIRE(B > 0 && R > 0, "B, R must be positive");
    const int slices = loss_slices(B, R, d->sms);
    KernelTimer timer("loss_fwd", stream);
    HELIO_CUDA_OK(launch_image_clusters(loss_fwd_kernel, B, slices, (cudaStream_t)stream, img, target, dmaps, tx, R, slices, per_img));
    return 0;
}

HELIO_API int helio_loss_bwd(const float* img, const float* target, const float* dmaps, const float* tx, const float* g_per_img,
                   const float* g_img_in, int B, int R, float* g_img, void* stream) {
    HELIO_REQUIRE(g_per_img, "null pointer");
    return helio_loss_bwd_packed(img, target, dmaps, tx, g_per_img, nullptr, g_img_in, B, R, g_img, stream);
}

namespace {
int loss_bwd_impl(const float* img, const float* target, const float* dmaps, const float* tx, const float* g_per_img,
                  const float* g_packed, const float* g_img_in, int B, int R, float* g_img, float* gmax, void* stream);
}

HELIO_API int helio_loss_bwd_packed(const float* img, const float* target, const float* dmaps, const float* tx,
                          const float* g_per_img, const float* g_packed, const float* g_img_in, int B, int R, float* g_img,
                          void* stream) {
    return loss_bwd_impl(img, target, dmaps, tx, g_per_img, g_packed, g_img_in, B, R, g_img, nullptr, stream);
}

namespace {
int loss_bwd_impl(const float* img, const float* target, const float* dmaps, const float* tx, const float* g_per_img,
                  const float* g_packed, const float* g_img_in, int B, int R, float* g_img, float* gmax, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(img && target && dmaps && tx && (g_per_img || g_packed) && g_img, "null pointer");
    HELIO_REQUIRE(B > 0 && R > 0, "B, R must be positive");
    // enough CTAs to fill the machine even for small B
    const size_t vecs = ((size_t)R * R + 3) / 4;
    int slices = (int)((vecs + 4 * kLossThreads - 1) / (4 * kLossThreads));
    const int want = (16 * d->sms + B - 1) / B;
    if (slices > want) slices = want;
    if (slices < 1) slices = 1;
    KernelTimer timer("loss_bwd", stream);
    if (gmax) {
        fill_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gmax, B, 0.f);
        HELIO_CUDA_OK(cudaGetLastError());
    }
    loss_bwd_kernel<<<(unsigned)((long long)B * slices), kLossThreads, 0, (cudaStream_t)stream>>>(
        img, target, dmaps, tx, g_per_img, g_packed, g_img_in, R, slices, g_img, gmax);
    HELIO_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // namespace

HELIO_API int helio_loss_pack(const float* per_img, int B, float* packed, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(per_img && packed, "null pointer");
    HELIO_REQUIRE(B > 0, "B must be positive");
    KernelTimer timer("loss_pack", stream);
    loss_pack_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(per_img, B, packed);
    HELIO_CUDA_OK(cudaGetLastError());
    return 0;
}

HELIO_API int helio_com_fwd(const float* img, int B, int H, int W, float eps, float* coords, float* sums, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(img && coords, "null pointer");
    HELIO_REQUIRE(B > 0 && H > 0 && W > 0, "B, H, W must be positive");
    int slices = 1;
    while (slices < 8 && (long long)B * slices < 8LL * d->sms && H / (2 * slices) >= kLossThreads / 32) slices *= 2;
    KernelTimer timer("com_fwd", stream);
    HELIO_CUDA_OK(launch_image_clusters(com_fwd_kernel, B, slices, (cudaStream_t)stream, img, H, W, slices, eps, coords, sums));
    return 0;
}

HELIO_API int helio_com_bwd(const float* img, const float* sums, const float* g_coords, int B, int H, int W, float eps,
                  float* g_img, void* stream) {
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    HELIO_REQUIRE(img && sums && g_coords && g_img, "null pointer");
    HELIO_REQUIRE(B > 0 && H > 0 && W > 0, "B, H, W must be positive");
    int slices = (H + kLossThreads / 32 - 1) / (kLossThreads / 32);
    const int want = (16 * d->sms + B - 1) / B;
    if (slices > want) slices = want;
    if (slices < 1) slices = 1;
    KernelTimer timer("com_bwd", stream);
    com_bwd_kernel<<<(unsigned)((long long)B * slices), kLossThreads, 0, (cudaStream_t)stream>>>(img, sums, g_coords, H, W,
                                                                                              slices, eps, g_img);
    HELIO_CUDA_OK(cudaGetLastError());
    return 0;
}

HELIO_API int64_t helio_step_partials_floats(int B, int N, int R, int impl) {
    const DeviceInfo* d = device_info();
    if (!d || B <= 0 || N <= 0 || R <= 0 || !splat_fwd_uses_tc(impl, B, N, R)) return 0;
    return (int64_t)B * splat_tc_fwd_partials_per_image(R, d->sms, tc_pair_mode()) * 3;
}

HELIO_API int helio_step_fwd(const helio_scene_t* scene, const float* helio_pos, const float* sun, const float* action,
                   const float* errs, const float* dmaps, int B, int N, int R, int impl, int render_target, float* params,
                   float* actual, float* refl, float* ideal, float* bounds, float* angles, float* img, float* target,
                   float* tx, float* per_img, float* packed, float* tgt_params, float* tgt_actual, float* tgt_refl,
                   float* loss_partials, void* cull_workspace, void* workspace, int64_t workspace_bytes, void* stream) {
    return helio_step_fwd_feed(scene, helio_pos, sun, action, errs, dmaps, B, N, R, impl, render_target, params, actual, refl, ideal,
                               bounds, angles, img, target, tx, per_img, packed, tgt_params, tgt_actual, tgt_refl, loss_partials,
                               cull_workspace, workspace, workspace_bytes, nullptr, stream);
}

HELIO_API int helio_step_fwd_feed(const helio_scene_t* scene, const float* helio_pos, const float* sun, const float* action,
                        const float* errs, const float* dmaps, int B, int N, int R, int impl, int render_target, float* params,
                        float* actual, float* refl, float* ideal, float* bounds, float* angles, float* img, float* target,
                        float* tx, float* per_img, float* packed, float* tgt_params, float* tgt_actual, float* tgt_refl,
                        float* loss_partials, void* cull_workspace, void* workspace, int64_t workspace_bytes,
                        const helio_feed_t* feed, void* stream) {
    HELIO_REQUIRE(scene && target && tx, "null pointer");
    HELIO_REQUIRE(action == nullptr || (ideal && bounds && angles && img && per_img && packed && dmaps), "null pointer");
    HELIO_REQUIRE(action != nullptr || render_target, "nothing to do");
    HELIO_REQUIRE(R > 0, "R must be positive");
    const DeviceInfo* d = nullptr;
    if (int rc = require_device(&d)) return rc;
    // With the tensor-core splat the HBM-bound passes of the loss block ride in its epilogue (see FwdFuse): the
    // target's per-image maximum, and the three per-image loss sums of the noisy image.
    const bool tc = splat_fwd_uses_tc(impl, B, N, R);
    const bool fused = loss_partials != nullptr && tc && feed == nullptr;   // loss sums in the noisy splat's epilogue
    static const bool fuse_max_on = []() {                // A/B switch (default on), read once
        const char* fm = std::getenv("HELIO_FUSE_MAX");
        return !(fm && fm[0] == '0');
    }();
    const bool fused_max = tc && fuse_max_on;            // target maximum in the target splat's epilogue
    if (render_target) {
        // error-free field aimed with the ideal normals (test_environment.py:429-436).  It depends on the suns only,
        // so it goes first: a caller streaming the action in from the host overlaps that copy with these kernels.
        HELIO_REQUIRE(tgt_params && tgt_actual && tgt_refl, "target render needs scratch buffers");
        if (int rc = helio_geom_fwd(scene, helio_pos, sun, nullptr, nullptr, B, N, tgt_params, tgt_actual, tgt_refl, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, 0, stream)) return rc;
        if (fused_max) {
            fill_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tx, B, 1e-6f);   // tx leaves this call clamped on every path
            HELIO_CUDA_OK(cudaGetLastError());
            FwdFuse fz{};
            fz.tile_max = tx;
            if (int rc = splat_fwd_impl(tgt_params, B, N, R, scene->width, scene->height, target, impl, stream, kFuseMax, fz)) return rc;
        } else {
            if (int rc = helio_splat_fwd(tgt_params, B, N, R, scene->width, scene->height, target, impl, stream)) return rc;
            if (int rc = helio_image_max(target, B, R, tx, stream)) return rc;
        }
    }
    if (action == nullptr) return 0;     // target-only call (phase 1 of a host-action step)
    // noisy field: K1 (with ideal normals, boundary, alignment and their sums -> packed[2..3])
    if (int rc = helio_geom_fwd(scene, helio_pos, sun, action, errs, B, N, params, actual, refl, ideal, bounds, angles,
                                packed + 2, workspace, workspace_bytes, stream)) return rc;
    // opt-in culling: contract only over the heliostats whose footprint can reach the receiver (cull.cuh)
    const float* sp = params;
    const int* counts = nullptr;
    if (cull_workspace && tc) {
        if (int rc = helio_cull(params, B, N, scene->width, scene->height, cull_workspace, cull_workspace_bytes(B, N), stream)) return rc;
        const CullBuffers cb = cull_carve(cull_workspace, B, N);
        sp = reinterpret_cast<const float*>(cb.cparams);
        counts = cb.counts;
    }
    if (feed) {
        if (int rc = splat_fwd_feed_impl(sp, B, N, R, scene->width, scene->height, img, counts ? HELIO_SPLAT_TC : impl, feed, stream, counts)) return rc;
    } else if (!fused) {
        if (int rc = splat_fwd_impl(sp, B, N, R, scene->width, scene->height, img, impl, stream, kFuseNone, FwdFuse{}, counts)) return rc;
    }
    if (fused) {
        FwdFuse fz{};
        fz.target = target, fz.dmaps = dmaps, fz.tx = tx, fz.partials = loss_partials;
        if (int rc = splat_fwd_impl(sp, B, N, R, scene->width, scene->height, img, impl, stream, kFuseLoss, fz, counts)) return rc;
        KernelTimer timer("loss_pack", stream);
        loss_pack_partials_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(
            loss_partials, splat_tc_fwd_partials_per_image(R, d->sms, tc_pair_mode()), B, per_img, packed);
        HELIO_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (int rc = helio_loss_fwd(img, target, dmaps, tx, B, R, per_img, stream)) return rc;
    return helio_loss_pack(per_img, B, packed, stream);
}

HELIO_API int helio_step_bwd(const helio_scene_t* scene, const float* helio_pos, const float* sun, const float* action,
                   const float* errs, const float* params, const float* img, const float* target, const float* dmaps,
                   const float* tx, int B, int N, int R, int impl, const float* g_packed, const float* g_per_img,
                   const float* g_img_in, const float* g_actual, const float* g_refl, const float* g_bounds,
                   const float* g_angles, const void* cull_workspace, float* g_img, float* moments, float* g_action,
                   void* stream) {
    HELIO_REQUIRE(scene && params, "null pointer");
    // culled forward: the backward visits the same compacted heliostat lists (the workspace helio_step_fwd filled)
    const float* sp = params;
    const int *counts = nullptr, *index = nullptr;
    if (cull_workspace && splat_bwd_uses_tc(impl, B, N, R)) {
        const CullBuffers cb = cull_carve(const_cast<void*>(cull_workspace), B, N);
        sp = reinterpret_cast<const float*>(cb.cparams), counts = cb.counts, index = cb.index;
    }
    HELIO_REQUIRE(moments || !(g_packed || g_per_img || g_img_in), "moments scratch missing");
    const float* g_mom = nullptr;
    if (g_packed || g_per_img) {
        HELIO_REQUIRE(g_img && g_action, "g_img scratch missing");
        // f16x3 backward: the loss backward also leaves max |g_img[b]| per image, in the first B floats of g_action -- free
        // until the geometry adjoint, the last kernel of this call, overwrites the whole buffer
        float* gmax = (bwd_prec_state().load(std::memory_order_relaxed) == 1 && splat_bwd_uses_tc(impl, B, N, R)) ? g_action : nullptr;
        if (int rc = loss_bwd_impl(img, target, dmaps, tx, g_per_img, g_packed, g_img_in, B, R, g_img, gmax, stream)) return rc;
        if (int rc = splat_bwd_impl(sp, g_img, B, N, R, scene->width, scene->height, moments, impl, stream, counts, index, gmax)) return rc;
        g_mom = moments;
    } else if (g_img_in) {
        if (int rc = splat_bwd_impl(sp, g_img_in, B, N, R, scene->width, scene->height, moments, impl, stream, counts, index)) return rc;
        g_mom = moments;
    }
    return helio_geom_bwd(scene, helio_pos, sun, action, errs, B, N, g_mom, g_actual, g_refl, g_bounds, g_angles,
                          g_packed ? g_packed + 2 : nullptr, g_action, stream);
}

}  // extern "C"
