/* helio_b200.h -- C ABI of libhelio_sm100.so: the DOODLE flux-renderer hot path on B200 (sm_100a).
 *
 * Drop-in boundary for ONE path of l3th4l/DOODLE: HelioField.render forward+backward and the
 * HelioEnv.step loss block.  The reference is pure Python/PyTorch and has no FFI of its own; each
 * entry point below names the reference Python it replaces (file:line relative to the reference
 * root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous fp32 unless the name ends in _host;
 *  - the library never allocates, frees or keeps pointers: the caller owns all buffers
 *    (outputs and workspaces included), so every call is CUDA-graph capturable;
 *  - calls only enqueue work on `stream` (a cudaStream_t passed as void*); no implicit sync;
 *  - return 0 on success, a positive cudaError_t or a negative HELIO_E_* code otherwise; the
 *    message is available from helio_last_error() (thread-local);
 *  - B = suns / episodes, N = heliostats, R = receiver resolution; images are [B][R][R] with
 *    axis 1 <-> plane_u (x, width) and axis 2 <-> plane_v (y, height)
 *    (meshgrid indexing="ij", newenv_rl_test_multi_error.py:129-131).
 */
#ifndef HELIO_B200_H
#define HELIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HELIO_ABI_VERSION 5

#if defined(__GNUC__)
#define HELIO_API __attribute__((visibility("default")))
#else
#define HELIO_API
#endif

#define HELIO_E_BADARG   (-1) /* null pointer / non-positive size / unsupported shape */
#define HELIO_E_NOSM100  (-2) /* device is not compute capability 10.x                 */
#define HELIO_E_WORKSPACE (-3) /* workspace too small                                   */

/* splat implementation selector */
#define HELIO_SPLAT_AUTO 0 /* tcgen05 when the shape supports it, else SIMT */
#define HELIO_SPLAT_SIMT 1 /* CUDA-core shared-memory-tiled path             */
#define HELIO_SPLAT_TC   2 /* tcgen05 3xTF32 path (error if unsupported)     */

/* Scene constants of one HelioField (newenv_rl_test_multi_error.py:162-216) plus the constants
 * HelioEnv.step hands to boundary() (test_environment.py:460-470).  Plain floats, passed by
 * pointer from HOST memory. */
typedef struct helio_scene {
    float target_pos[3];    /* HelioField.target_position                                  */
    float target_normal[3]; /* HelioField.target_normal, unit length (:189-192)            */
    float plane_u[3];       /* :206                                                        */
    float plane_v[3];       /* :207-213                                                    */
    float width, height;    /* target_area (:187)                                          */
    float sigma_scale;      /* :197                                                        */
    /* boundary(): HelioEnv passes ITS OWN targ_pos/targ_norm (not normalised) and the fixed
     * east/up axes (1,0,0)/(0,0,1) (test_environment.py:461-470). */
    float bnd_targ_pos[3];
    float bnd_targ_norm[3];
    float bnd_u[3];
    float bnd_v[3];
    float bnd_width, bnd_height;
} helio_scene_t;

HELIO_API int helio_abi_version(void);
HELIO_API const char* helio_last_error(void);
/* 1 if the current device can run the kernels (cc 10.x), else 0 (and last_error is set). */
HELIO_API int helio_device_ok(void);

/* tcgen05 splat kernels: 0 = auto (CTA pairs / cta_group::2 where the shape allows; default), 1 = single-CTA
 * kernels only, 2 = force CTA pairs.  Process-wide; initial value from the environment variable HELIO_TC_PAIR.
 * No reference counterpart (tuning / A-B switch of this library). */
HELIO_API int helio_set_tc_pair_mode(int mode);

/* Forward splat operand format on the tcgen05 path.  0: 3xTF32 everywhere.  1: "f16x3" everywhere -- both operands of
 * the forward are Gaussians in [0,1]; they are scaled by 2^14 and split into two fp16 pieces (11 + 11 significant bits,
 * the accuracy the tf32 hi/lo split keeps), three kind::f16 MMAs per K-step, fp32 accumulation, exact 2^-28 unscale in
 * the epilogue.  2 (default): auto = f16x3 (same measured accuracy as 3xTF32, half the tensor work: faster at every
 * shape); mode 0 keeps the 3xTF32 contraction BASELINE.json names selectable.
 * Process-wide; initial value from HELIO_FWD_PREC.  The backward has its own switch (helio_set_bwd_precision). */
HELIO_API int helio_set_fwd_precision(int mode);

/* Backward splat operand format inside helio_step_bwd.  0: 3xTF32.  1: "f16x3, K = 64" -- the Gaussian operand scaled by
 * 2^14 as in the forward, the image gradient scaled PER IMAGE by the power of two that brings max |dL/dimg[b]| under 2^14
 * (the loss backward produces that maximum on the way), both split into two fp16 pieces; a 64 KB pipeline stage then
 * covers 64 instead of 32 steps of the contraction with the same twelve MMAs (kind::f16): half the tensor work, half the
 * stage hand-overs.  Accuracy as 3xTF32 (22 significant bits per operand; entries more than 2^17 below their image's
 * maximum keep an absolute error below 2^-38 of that maximum).  helio_splat_bwd (arbitrary g_img, no scratch for the
 * maxima) always uses 3xTF32.  Process-wide; initial value from HELIO_BWD_PREC. */
HELIO_API int helio_set_bwd_precision(int mode);

/* Opt-in per-kernel timing (no reference counterpart; SURVEY.md section 5 "tracing / profiling").
 * helio_profile_enable(1) clears earlier records and makes every entry point record a CUDA event pair
 * around each kernel it enqueues, on the caller's stream; helio_profile_enable(0) stops and clears.
 * helio_profile_get waits for record `index` and returns the kernel's name (static string: geom_fwd,
 * splat_fwd, image_max, loss_fwd, loss_pack, loss_bwd, splat_bwd, geom_bwd) and its duration in ms.
 * Creating events allocates driver resources: keep it off in production and under stream capture. */
HELIO_API int helio_profile_enable(int on);
HELIO_API int helio_profile_count(void);
HELIO_API int helio_profile_get(int index, const char** name, float* ms);

/* Diagnostic (no reference counterpart): SM clock in MHz that CTA 0 of the most recent tcgen05 splat kernel held over its
 * lifetime (which = 0 forward, 1 backward; cycles / wall nanoseconds measured inside the kernel).  B200 power-throttles
 * under sustained tensor load, so this -- not the application clock NVML reports -- is the clock the roofline of those
 * kernels should be quoted at.  Synchronises with the device.  0 if that kernel has not run. */
HELIO_API int helio_tc_clock_mhz(int which, float* mhz_host);

/* Bytes of workspace helio_geom_fwd needs for (B, N) (block partials + ticket counter).  The
 * workspace must be zero-initialised ONCE by the caller; the kernel leaves it zeroed.  A workspace holds
 * the ticket counter of the ordered reduction: use it from ONE stream at a time (one workspace per
 * concurrently running call). */
HELIO_API int64_t helio_geom_workspace_bytes(int B, int N);

/* K1 forward.  Replaces, fused per (b, n):
 *   rotate_normals_batch            newenv_rl_test_multi_error.py:78-104
 *   Up-axis leaky-ReLU + normalise  :369-373
 *   incident dir, reflect_vectors   :376-383, :46-50
 *   ray_plane_intersection_batch    :52-75
 *   sigma from distance             :126-127, :146
 *   calculate_ideal_normals         :256-278          (ideal, optional)
 *   boundary(return_all=True)       test_environment.py:101-130   (bounds, optional)
 *   calculate_angles_mrad           test_environment.py:132-155   (angles, optional)
 * in : helio[N][3], sun[B][3], action[B][N][3] (NULL = aim every mirror with its ideal normal, the
 *      target render of test_environment.py:429-433), errs[B][N][2] in mrad (NULL = zero errors)
 * out: params[B][N][4] = {a, b, k2, amp}: the heliostat's footprint on the receiver is
 *        amp * exp2(-k2 (x_i - a)^2) * exp2(-k2 (y_j - b)^2)   (k2 = log2(e)/max(2 sigma^2,1e-12);
 *        an invalid ray has k2 = 0, amp = 1: +1 on every pixel, :141-143)
 *      actual[B][N][3], refl[B][N][3] (= reflected_rays [B*N][3]);
 *      ideal[B][N][3], bounds[B][N], angles[B][N] may each be NULL;
 *      sums[2] = {sum bounds, sum angles} (warp-shuffle + ordered block reduction; NULL to skip,
 *      needs `workspace`). */
HELIO_API int helio_geom_fwd(const helio_scene_t* scene_host, const float* helio, const float* sun,
                   const float* action, const float* errs, int B, int N,
                   float* params, float* actual, float* refl, float* ideal,
                   float* bounds, float* angles, float* sums,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* K1 backward (recomputes the forward chain; what autograd does for the ops listed above).
 * g_moments[B][N][4] = {S0,Sx,Sy,S2} from helio_splat_bwd; g_actual, g_refl [B][N][3];
 * g_bounds, g_angles [B][N]; g_sums[2] (device) = upstream grads of sums.  Any may be NULL.
 * out: g_action[B][N][3] (overwritten). */
HELIO_API int helio_geom_bwd(const helio_scene_t* scene_host, const float* helio, const float* sun,
                   const float* action, const float* errs, int B, int N,
                   const float* g_moments, const float* g_actual, const float* g_refl,
                   const float* g_bounds, const float* g_angles, const float* g_sums,
                   float* g_action, void* stream);

/* K2: Gaussian flux splat + sum over heliostats.  Replaces gaussian_blur_batch
 * (newenv_rl_test_multi_error.py:107-149) and images = gauss.sum(1) (:404-406) without
 * materialising [B][N][R][R]:  img[b] = Gx^T diag(amp) Gy.
 * impl: HELIO_SPLAT_*.  img[B][R][R] is overwritten. */
HELIO_API int helio_splat_fwd(const float* params, int B, int N, int R, float width, float height,
                    float* img, int impl, void* stream);

/* K3: adjoint of K2 w.r.t. the footprint parameters, recomputing the Gaussians.
 * in : params[B][N][4], g_img[B][R][R]
 * out: moments[B][N][4] = { S0 = sum g G, Sx = sum g G (x_i-a), Sy = sum g G (y_j-b),
 *                           S2 = sum g G ((x_i-a)^2 + (y_j-b)^2) },  G = amp Gx_i Gy_j. */
HELIO_API int helio_splat_bwd(const float* params, const float* g_img, int B, int N, int R,
                    float width, float height, float* moments, int impl, void* stream);

/* Footprint culling (opt-in; no reference counterpart -- the reference evaluates every heliostat on every
 * pixel, newenv_rl_test_multi_error.py:140-148).  helio_cull compacts, per sun and in the original order,
 * the heliostats whose Gaussian can exceed 2^-40 of its peak anywhere on the receiver:
 *   k2 (dx^2 + dy^2) <= 40,  dx = max(|a| - width/2, 0),  dy = max(|b| - height/2, 0)   ({a,b,k2,amp} = params)
 * into `workspace` (compacted params [B][N][4], original indices [B][N], counts [B]).  The *_culled splats
 * (tcgen05 path only) contract over counts[b] heliostats / visit only the kept ones; culled heliostats get
 * exactly zero moments.  Every dropped term is below amp * 2^-40 on every pixel. */
HELIO_API int64_t helio_cull_workspace_bytes(int B, int N);
HELIO_API int helio_cull(const float* params, int B, int N, float width, float height,
               void* workspace, int64_t workspace_bytes, void* stream);
HELIO_API int helio_splat_fwd_culled(const void* cull_workspace, int B, int N, int R, float width, float height,
                           float* img, void* stream);
HELIO_API int helio_splat_bwd_culled(const void* cull_workspace, const float* g_img, int B, int N, int R,
                           float width, float height, float* moments, void* stream);

/* tx[b] = max(max_ij target[b], 1e-6)   (test_environment.py:436). */
HELIO_API int helio_image_max(const float* target, int B, int R, float* tx, void* stream);

/* Distance maps of set_sun_pos: replaces make_distance_maps (test_environment.py:92-97), i.e. per image
 *   mask = img > thr * img.max();  dmap = float32(scipy.ndimage.distance_transform_edt(1 - mask))
 * with an exact integer separable transform on the GPU (bit-identical to scipy's exact EDT; an image
 * without mask pixels reproduces scipy's result for an input without background).
 * img, dmaps [B][R][R]; workspace: helio_distance_maps_workspace_bytes(B, R) bytes, no initialisation. */
HELIO_API int64_t helio_distance_maps_workspace_bytes(int B, int R);
HELIO_API int helio_distance_maps(const float* img, int B, int R, float thr, float* dmaps,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* K4 forward: per-image loss partials of HelioEnv.step (test_environment.py:438-457,492):
 *   diff = (img - target)/tx ;  per_img[b] = { sum diff^2, sum |diff| dmaps, sum |diff| }.
 * The caller forms mse = sum_b per_img[b][0]/(B R^2), dist = mean_b per_img[b][1],
 * mae_image[b] = per_img[b][2]/R^2 (and the error-mask variants) from per_img[B][3]. */
HELIO_API int helio_loss_fwd(const float* img, const float* target, const float* dmaps, const float* tx,
                   int B, int R, float* per_img, void* stream);

/* K4 backward: g_img = (2 g0 diff + (g1 dmaps + g2) sign(diff)) / tx with g_per_img[B][3] =
 * {g0,g1,g2}; g_img_in (NULL or [B][R][R]) is added (gradient arriving through obs['img']). */
HELIO_API int helio_loss_bwd(const float* img, const float* target, const float* dmaps, const float* tx,
                   const float* g_per_img, const float* g_img_in, int B, int R, float* g_img,
                   void* stream);

/* Centre of mass of each image: CenterOfMass2D.forward (layers/center_of_mass.py:21-60), the encoder
 * feed of train_with_env_com_trunc_advantage_ttt.py:42-53.  img[B][H][W] (row i = y, column j = x);
 * coords[B][2] = (sum w j, sum w i) / (sum w + eps) with w = max(img, 0), or (-1, -1) when sum w <= 0;
 * sums[B][3] = {sum w, sum w j, sum w i} (may be NULL; needed by helio_com_bwd). */
HELIO_API int helio_com_fwd(const float* img, int B, int H, int W, float eps, float* coords, float* sums, void* stream);
/* its adjoint: g_img[i][j] = [img >= 0] (g_x (j - x_com) + g_y (i - y_com)) / (sum w + eps), zero for
 * images without mass.  g_coords[B][2]. */
HELIO_API int helio_com_bwd(const float* img, const float* sums, const float* g_coords, int B, int H, int W,
                  float eps, float* g_img, void* stream);

/* helio_loss_bwd with the batch-wide gradients folded in: {g0,g1,g2} = g_per_img[b] (may be NULL) +
 * {g_packed[0], g_packed[1], 0} (may be NULL), where g_packed[4] (device) are the upstream grads of
 * helio_step_fwd's packed sums. */
HELIO_API int helio_loss_bwd_packed(const float* img, const float* target, const float* dmaps, const float* tx,
                          const float* g_per_img, const float* g_packed, const float* g_img_in,
                          int B, int R, float* g_img, void* stream);

/* packed[0..1] = sum_b per_img[b][0..1]: the numerators of F.mse_loss(pred_n, targ_n) and of
 * (err * distance_maps).sum((1,2)).mean() (test_environment.py:456-457).  Ordered, deterministic. */
HELIO_API int helio_loss_pack(const float* per_img, int B, float* packed, void* stream);

/* One call for the whole forward of HelioEnv.step (test_environment.py:402-457): enqueues
 *   K1 noisy field (+ ideal normals, boundary, alignment)   :414-421, :455, :460-470
 *   K2 img
 *   [render_target != 0]  K1 + K2 of the error-free field aimed with the ideal normals -> target,
 *                         tx = max(target).clamp_min(1e-6)                                :429-436
 *   K4 per_img, packed[4] = { sum diff^2, sum |diff| dmaps, sum bounds, sum angles }      :438-457
 * The caller divides packed by {B R^2, B, B N, B N} (after an all-reduce when B is sharded).
 * render_target == 0 reuses the caller's target / tx (they depend on sun only).
 * tgt_params[B][N][4], tgt_actual/tgt_refl[B][N][3] are scratch for the target render (may be NULL
 * when render_target == 0).  `workspace` as for helio_geom_fwd.  Same results as calling the individual
 * entry points (to summation order); it exists to cut host overhead for small fields and to fuse the
 * loss passes into the splat epilogues for large ones.
 *
 * The target render is enqueued first (it depends on the suns only); action == NULL stops after it
 * (phase 1 of a step whose action is still being copied in from the host).
 *
 * loss_partials (may be NULL): helio_step_partials_floats(B, N, R, impl) floats of scratch.  When given and
 * the shape takes the tcgen05 splat, image_max and loss_fwd do not run as separate passes: the target's
 * per-image maximum and the three per-image loss sums are accumulated in the splat epilogues while the
 * image rows are in registers.  tx holds max(max_ij target, 1e-6) on every path (test_environment.py:436).
 * A NaN pixel does not propagate into tx (fmaxf / integer max drop it; torch.amax would return NaN): the
 * NaN then reaches the metrics through img / target themselves.
 *
 * cull_workspace (may be NULL = dense): helio_cull_workspace_bytes(B, N) bytes.  When given and the shape
 * takes the tcgen05 splat, the noisy render contracts only over the heliostats kept by helio_cull; pass
 * the same (untouched) workspace to helio_step_bwd. */
HELIO_API int64_t helio_step_partials_floats(int B, int N, int R, int impl);
HELIO_API int helio_step_fwd(const helio_scene_t* scene_host, const float* helio, const float* sun,
                   const float* action, const float* errs, const float* dmaps,
                   int B, int N, int R, int impl, int render_target,
                   float* params, float* actual, float* refl, float* ideal, float* bounds, float* angles,
                   float* img, float* target, float* tx, float* per_img, float* packed,
                   float* tgt_params, float* tgt_actual, float* tgt_refl, float* loss_partials,
                   void* cull_workspace, void* workspace, int64_t workspace_bytes, void* stream);

/* Encoder feed of the reference's rollout (train_with_env.py:182-209: hist[:, -1] = img, then the policy's encoder;
 * train_with_env_com_trunc_advantage_ttt.py:42-53: CenterOfMass2D of every history frame, layers/center_of_mass.py:21-60)
 * produced WHILE the image tile is still in the splat epilogue's registers, instead of by extra passes over the image:
 *   com_coords[B][2], com_sums[B][3]  centre of mass of the noisy image exactly as helio_com_fwd returns it
 *                                     (coords = (sum w j, sum w i) / (sum w + eps), w = max(img, 0); (-1,-1) without mass);
 *   img2 (may be NULL)                a second copy of the image, image b at img2 + b * img2_batch_stride floats
 *                                     (e.g. the newest slot of a [B][k][R][R] history buffer: stride k R R).
 * partials: helio_step_partials_floats(B, N, R, impl) floats of scratch (0 floats = shape not on the tcgen05 path: the
 * same results then come from helio_com_fwd and a strided copy, nothing is fused).  Plain host struct, passed by pointer. */
typedef struct helio_feed {
    float* img2;
    int64_t img2_batch_stride;
    float eps;              /* CenterOfMass2D.eps (1e-12) */
    float* com_coords;      /* [B][2], may be NULL together with com_sums */
    float* com_sums;        /* [B][3] */
    float* partials;
    int64_t partials_floats; /* floats available at `partials` (HELIO_E_WORKSPACE if fewer than helio_step_partials_floats) */
} helio_feed_t;

/* helio_splat_fwd with the feed outputs (K2 + centre of mass + optional second image copy in one kernel). */
HELIO_API int helio_splat_fwd_feed(const float* params, int B, int N, int R, float width, float height, float* img, int impl,
                         const helio_feed_t* feed_host, void* stream);

/* helio_step_fwd with the feed outputs taken from the noisy render's epilogue (all other arguments as helio_step_fwd;
 * with feed_host != NULL the loss sums run as the separate loss_fwd pass: loss_partials is ignored). */
HELIO_API int helio_step_fwd_feed(const helio_scene_t* scene_host, const float* helio, const float* sun,
                        const float* action, const float* errs, const float* dmaps,
                        int B, int N, int R, int impl, int render_target,
                        float* params, float* actual, float* refl, float* ideal, float* bounds, float* angles,
                        float* img, float* target, float* tx, float* per_img, float* packed,
                        float* tgt_params, float* tgt_actual, float* tgt_refl, float* loss_partials,
                        void* cull_workspace, void* workspace, int64_t workspace_bytes,
                        const helio_feed_t* feed_host, void* stream);

/* Backward of helio_step_fwd down to g_action[B][N][3]: K4' (g_img) -> K3 (moments) -> K1'.
 * g_packed[4] (device) = upstream grads of packed; g_per_img[B][3], g_img_in[B][R][R] (gradient
 * arriving through obs['img']), g_actual, g_refl [B][N][3], g_bounds, g_angles [B][N] may be NULL.
 * g_img[B][R][R] and moments[B][N][4] are scratch. */
HELIO_API int helio_step_bwd(const helio_scene_t* scene_host, const float* helio, const float* sun,
                   const float* action, const float* errs, const float* params,
                   const float* img, const float* target, const float* dmaps, const float* tx,
                   int B, int N, int R, int impl,
                   const float* g_packed, const float* g_per_img, const float* g_img_in,
                   const float* g_actual, const float* g_refl, const float* g_bounds, const float* g_angles,
                   const void* cull_workspace, float* g_img, float* moments, float* g_action, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HELIO_B200_H */
