"""CUDA-graph replay of ``HelioEnv.step``.

For a small field (BASELINE.json configs[1]: N=50, 128x128, B=25) the kernels of a step take a few microseconds
each and ``env.step`` + ``loss.backward()`` is all launch + Python overhead.  Every entry point of
libhelio_sm100.so only enqueues kernels on the caller's stream and never allocates, so the step can be captured
once and replayed.  Two users:

``StepGraph`` (internal; what ``HelioEnv(graph="auto")`` uses -- TRANSPARENT to the caller)
    ``env.step(action)`` keeps its contract: metrics carry autograd history down to ``action``, the caller's own
    ``loss.backward()`` works, any number of steps may be alive at once (the reference's rollout accumulates the
    losses of T steps before one backward, train_with_env.py:190-209; its test-time-compute loop does
    step / backward / optimiser step, train_with_env_com_trunc_advantage_ttt.py:291-312).
      forward : one tiny copy of the action into the static arena, ONE graph launch (the same kernels
                helio_step_fwd enqueues + the glue that forms means / mae / aux), ONE clone of the arena so that this
                step's outputs and saved tensors outlive the next replay;
      backward: upstream gradients copied into static buffers, ONE graph launch (helio_step_bwd), ONE clone of the
                action gradient.  A step that is no longer the most recent one first copies its saved arena back.
    Results are bit-identical to the eager fused step (same kernels, same order).

``GraphedStep`` (public, explicit)
    step + objective + backward in ONE graph for loops that want ``(loss, grad)`` per call with static outputs.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import torch

from . import _lib
from .functional import _cf, _ptr, _stream


def _default_objective(m: Dict[str, torch.Tensor]) -> torch.Tensor:
    return m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]


# ======================================================================================================================
# transparent replay inside HelioEnv.step
# ======================================================================================================================
class StepGraph:
    """Static buffers + captured graphs of one HelioEnv's fused step.  Owned by the env; see HelioEnv._graph_step."""

    # slots of the arena, in floats per (B, N, R); everything a step returns or saves for its backward
    _SLOTS = ("action", "errs", "params", "actual", "refl", "ideal", "bounds", "angles", "img", "per_img", "packed", "means",
              "mae", "aux")

    def __init__(self, env):
        self.env = env
        nf = env.noisy_field
        self.dev = nf.device
        B, N, R = env.batch_size, env.num_heliostats, env.resolution
        self.B, self.N, self.R = B, N, R
        self.impl = nf.splat_impl
        self.impl_bwd = nf.splat_impl if nf.splat_impl_bwd is None else nf.splat_impl_bwd
        sizes = dict(action=3 * B * N, errs=2 * B * N, params=4 * B * N, actual=3 * B * N, refl=3 * B * N, ideal=3 * B * N,
                     bounds=B * N, angles=B * N, img=B * R * R, per_img=3 * B, packed=4, means=4, mae=B, aux=B * (3 + 3 * N))
        self.offsets, off = {}, 0
        for k in self._SLOTS:
            self.offsets[k] = (off, sizes[k])
            off += (sizes[k] + 3) // 4 * 4                      # 16-byte aligned slots (float4 kernels)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.arena = torch.zeros(off, **f32)
        self.arena_gen = 0                                      # which step's data the static arena holds
        self.s = self.shaped(self.arena)                        # views of the STATIC arena (what the graphs read / write)
        self.s["action2d"] = self.s["action"].view(B, 3 * N)
        # (shape, stride, offset) of the per-step outputs inside a cloned arena: one as_strided each, no slice + view pairs
        o = {k: v[0] for k, v in self.offsets.items()}
        self.out_specs = (((B, R, R), (R * R, R, 1), o["img"]), ((4,), (1,), o["means"]), ((B, 3 + 3 * N), (3 + 3 * N, 1), o["aux"]),
                          ((B * N, 3), (3, 1), o["refl"]), ((B, N), (N, 1), o["bounds"]), ((B * N,), (1,), o["angles"]),
                          ((B, 1), (1, 1), o["mae"]), ((B * N, 3), (3, 1), o["ideal"]))
        # per-environment inputs, copied here so that replaced tensors (reset -> new error tensor, set_sun_pos) only
        # cost a copy, never a recapture
        self.sun = torch.empty(B, 3, **f32)
        self.dmaps = torch.empty(B, R, R, **f32)
        self.target = torch.empty(B, R, R, **f32)
        self.tx = torch.empty(B, **f32)
        self.helio = nf.heliostat_positions
        # sharded env: the packed sums are all-reduced inside the forward graph and divided by the GLOBAL counts
        # (dist.global_means); the backward of that all-reduce is the identity, so d means / d packed_local is the same factor
        self.reduce = getattr(env, "_graph_reduce", None)          # (process group, world size): only dist.make_sharded_env sets it
        self.inv_counts = env._inv_counts if self.reduce is None else env._inv_counts / self.reduce[1]
        self.scene = nf.scene()
        self.workspace = torch.zeros_like(nf._geom_workspace(B))
        self.src_keys = {}                                      # (data_ptr, version) of the env tensors last copied in
        self.sun_gen = 0
        # backward statics
        self.reduced = torch.zeros(4, **f32)
        self.g_means = torch.zeros(4, **f32)
        self.g_packed = torch.zeros(4, **f32)
        self.g_per_img = torch.zeros(B, 3, **f32)
        self.g_img_in = None                                    # allocated on first use
        self.g_refl = None
        self.g_bounds = None
        self.g_img = torch.empty(B, R, R, **f32)
        self.moments = torch.empty(B, N, 4, **f32)
        self.g_action = torch.empty(B, N, 3, **f32)
        self.g_action2d = self.g_action.view(B, 3 * N)
        self._errs_src = None                                   # (tensor, version) of the error tensor last copied in
        # deferred finite check (test_environment.py:495-501): the three means land in pinned host memory through a
        # copy node of the forward graph and are examined at the start of the NEXT step -- no sync in the step itself
        self.host_means = torch.zeros(4, dtype=torch.float32, device="cpu").pin_memory()   # explicit: callers may set a CUDA default device
        self.host_means_np = self.host_means.numpy()            # same memory; reading three floats costs no torch dispatch
        self.fwd_done = torch.cuda.Event()
        self._stream_obj, self._stream_raw = None, None         # torch Stream object of the raw handle last seen (building one costs ~20 us)
        self.pending_check = False
        self.fwd_graph = None
        self.bwd_graphs = {}
        self.launches_fwd = 0

    # ---- views -------------------------------------------------------------------------------------------------------
    def view(self, arena: torch.Tensor, k: str) -> torch.Tensor:
        o, n = self.offsets[k]
        return arena[o:o + n]

    def shaped(self, arena: torch.Tensor):
        B, N, R = self.B, self.N, self.R
        v = lambda k: self.view(arena, k)
        return dict(action=v("action").view(B, N, 3), errs=v("errs").view(B, N, 2), params=v("params").view(B, N, 4),
                    actual=v("actual").view(B, N, 3), refl=v("refl").view(B * N, 3), ideal=v("ideal").view(B, N, 3),
                    bounds=v("bounds").view(B, N), angles=v("angles").view(B, N), img=v("img").view(B, R, R),
                    per_img=v("per_img").view(B, 3), packed=v("packed"), means=v("means"), mae=v("mae").view(B, 1),
                    aux=v("aux").view(B, 3 + 3 * N))

    # ---- inputs that belong to the environment -----------------------------------------------------------------------
    def refresh_inputs(self):
        """Copy sun / errors / distance maps / target into the static buffers when their source tensors changed
        (replaced -- reset() draws a new error tensor, set_sun_pos() new suns -- or edited in place).  The common case
        (nothing changed) costs three identity + version comparisons.  A changed sun batch bumps ``sun_gen``."""
        env, nf = self.env, self.env.noisy_field
        B = self.B
        # the tensor render() would slice its errors from (newenv_rl_test_multi_error.py:340-353); None = fresh draw per call
        src = nf.error_angles_mrad if B == 1 else (nf.batch_error_angles_mrad if nf.batch_error_angles_mrad is not None
                                                   and B <= nf.batch_error_angles_mrad.shape[0] else None)
        k = self._errs_src
        if src is None or k is None or k[0] is not src or k[1] != src._version:
            self.s["errs"].copy_(nf._select_errors(B).reshape(B, self.N, 2))
            self._errs_src = None if src is None else (src, src._version)
        t = env.distance_maps
        k = self.src_keys.get("dmaps")
        if k is None or k[0] is not t or k[1] != t._version:
            self.dmaps.copy_(t)
            self.src_keys["dmaps"] = (t, t._version)
        t = env.sun_pos
        k = self.src_keys.get("sun")
        if k is None or k[0] is not t or k[1] != t._version:
            self.sun.copy_(t)
            self.src_keys["sun"] = (t, t._version)
            self.sun_gen += 1
            self._render_target()

    def _render_target(self):
        """target / tx of the current suns (test_environment.py:429-436): the env's exact cache when it holds them, else
        the target phase of helio_step_fwd (eager, once per set_sun_pos)."""
        env = self.env
        target, tx = env._cached_target()
        if target is None:
            lib = _lib.load()
            B, N, R = self.B, self.N, self.R
            target = torch.empty(B, R, R, dtype=torch.float32, device=self.dev)
            tx = torch.empty(B, dtype=torch.float32, device=self.dev)
            scratch = torch.empty(B * N * 10, dtype=torch.float32, device=self.dev)
            rc = lib.helio_step_fwd(C.byref(self.scene), _ptr(self.helio), _ptr(self.sun), None, None, None, B, N, R, self.impl, 1,
                                    None, None, None, None, None, None, None, _ptr(target), _ptr(tx), None, None, _ptr(scratch),
                                    _ptr(scratch[4 * B * N:]), _ptr(scratch[7 * B * N:]), None, None, _ptr(self.workspace),
                                    self.workspace.numel() * 4, _stream())
            _lib.check(rc, "helio_step_fwd (target)")
            env._store_target(target, tx)
        self.target.copy_(target)
        self.tx.copy_(tx)

    # ---- capture -----------------------------------------------------------------------------------------------------
    def _enqueue_forward(self):
        lib = _lib.load()
        B, N, R = self.B, self.N, self.R
        s = self.s
        rc = lib.helio_step_fwd(
            C.byref(self.scene), _ptr(self.helio), _ptr(self.sun), _ptr(s["action"]), _ptr(s["errs"]), _ptr(self.dmaps), B, N, R,
            self.impl, 0, _ptr(s["params"]), _ptr(s["actual"]), _ptr(s["refl"]), _ptr(s["ideal"]), _ptr(s["bounds"]), _ptr(s["angles"]),
            _ptr(s["img"]), _ptr(self.target), _ptr(self.tx), _ptr(s["per_img"]), _ptr(s["packed"]), None, None, None, None, None,
            _ptr(self.workspace), self.workspace.numel() * 4, _stream())
        _lib.check(rc, "helio_step_fwd")
        # the glue HelioEnv.step performs in torch, same ops (bit-identical): means, mae_image, aux
        if self.reduce is not None:
            import torch.distributed as dist
            self.reduced.copy_(s["packed"])             # keep the local sums in the arena (they are this rank's outputs)
            dist.all_reduce(self.reduced, op=dist.ReduceOp.SUM, group=self.reduce[0])
            torch.mul(self.reduced, self.inv_counts, out=s["means"])
        else:
            torch.mul(s["packed"], self.inv_counts, out=s["means"])
        torch.div(s["per_img"][:, 2:3], float(R * R), out=s["mae"])
        s["aux"][:, :3].copy_(self.sun)
        s["aux"][:, 3:].copy_(s["action"].view(B, 3 * N))
        self.host_means.copy_(s["means"], non_blocking=True)

    def _capture(self, fn):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side), torch.no_grad():          # warm-up off the capture: one-time setup happens here
            fn()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            fn()
        return g

    def capture_forward(self):
        from . import functional as Fn
        self.refresh_inputs()
        self.fwd_graph = self._capture(self._enqueue_forward)
        self.launches_fwd = Fn._step_fwd_kernels(False, False, int(_lib.load().helio_step_partials_floats(self.B, self.N, self.R, self.impl)) > 0)

    def _enqueue_backward(self, sig):
        has_means, has_mae, has_img, has_refl, has_bounds = sig
        lib = _lib.load()
        B, N, R = self.B, self.N, self.R
        s = self.s
        if has_means:
            torch.mul(self.g_means, self.inv_counts, out=self.g_packed)      # d means / d packed
        rc = lib.helio_step_bwd(
            C.byref(self.scene), _ptr(self.helio), _ptr(self.sun), _ptr(s["action"]), _ptr(s["errs"]), _ptr(s["params"]), _ptr(s["img"]),
            _ptr(self.target), _ptr(self.dmaps), _ptr(self.tx), B, N, R, self.impl_bwd,
            _ptr(self.g_packed) if has_means else None, _ptr(self.g_per_img) if has_mae else None,
            _ptr(self.g_img_in) if has_img else None, None, _ptr(self.g_refl) if has_refl else None,
            _ptr(self.g_bounds) if has_bounds else None, None, None, _ptr(self.g_img), _ptr(self.moments), _ptr(self.g_action), _stream())
        _lib.check(rc, "helio_step_bwd")

    def backward_graph(self, sig):
        g = self.bwd_graphs.get(sig)
        if g is None:
            f32 = dict(dtype=torch.float32, device=self.dev)
            if sig[2] and self.g_img_in is None:
                self.g_img_in = torch.zeros(self.B, self.R, self.R, **f32)
            if sig[3] and self.g_refl is None:
                self.g_refl = torch.zeros(self.B * self.N, 3, **f32)
            if sig[4] and self.g_bounds is None:
                self.g_bounds = torch.zeros(self.B, self.N, **f32)
            g = self.bwd_graphs[sig] = self._capture(lambda: self._enqueue_backward(sig))
        return g

    # ---- deferred finite check ---------------------------------------------------------------------------------------
    def check_pending(self):
        if not self.pending_check:
            return
        self.pending_check = False
        self.fwd_done.synchronize()                             # long done by the time the caller comes back
        mse, dist_l, bound = self.host_means_np[:3].tolist()
        if mse - mse == 0.0 and dist_l - dist_l == 0.0 and bound - bound == 0.0:      # all finite: the common case
            return
        import math
        assert not math.isnan(mse), "MSE is NaN"
        assert not math.isnan(dist_l), "Distance loss is NaN"
        assert not math.isnan(bound), "Boundary loss is NaN"
        assert not math.isinf(mse), "MSE is Inf"
        assert not math.isinf(dist_l), "Distance loss is Inf"
        assert not math.isinf(bound), "Boundary loss is Inf"


class GraphStepFn(torch.autograd.Function):
    """env.step through StepGraph: forward / backward are graph replays; outputs are views of a per-step clone of the
    arena, so they (and the tensors the backward needs) survive later steps."""

    @staticmethod
    def forward(ctx, action, sg: StepGraph):
        from . import functional as Fn
        sg.refresh_inputs()
        two_d = action.dim() == 2
        (sg.s["action2d"] if two_d else sg.s["action"]).copy_(action)       # [B,3N] or [B,N,3], any dtype / strides
        sg.fwd_graph.replay()
        sg.arena_gen += 1
        if sg.env.check_finite:
            raw = _stream()
            if raw != sg._stream_raw:
                sg._stream_obj, sg._stream_raw = torch.cuda.current_stream(sg.dev), raw
            sg.fwd_done.record(sg._stream_obj)
            sg.pending_check = True
        Fn._LAUNCHES += sg.launches_fwd
        mine = sg.arena.clone()                                 # this step's outputs + saved tensors
        img, means, aux, refl, bounds, angles, mae, ideal = [mine.as_strided(*spec) for spec in sg.out_specs]
        ctx.sg, ctx.gen, ctx.sun_gen, ctx.two_d = sg, sg.arena_gen, sg.sun_gen, two_d
        ctx.save_for_backward(mine)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(ideal, angles)
        return img, means, aux, refl, bounds, angles, mae, ideal

    @staticmethod
    def backward(ctx, g_img, g_means, g_aux, g_refl, g_bounds, g_angles, g_mae, g_ideal):
        from . import functional as Fn
        sg: StepGraph = ctx.sg
        (mine,) = ctx.saved_tensors
        if ctx.sun_gen != sg.sun_gen:
            raise RuntimeError("HelioEnv(graph=...): the sun positions changed between this step and its backward; "
                               "call backward before set_sun_pos / a resampling reset, or construct the env with graph=False")
        if ctx.gen != sg.arena_gen:                             # an older step: put its tensors back under the graph
            sg.arena.copy_(mine)
            sg.arena_gen = ctx.gen
        B, N, R = sg.B, sg.N, sg.R
        sig = (g_means is not None, g_mae is not None, g_img is not None, g_refl is not None, g_bounds is not None)
        if any(sig):
            graph = sg.backward_graph(sig)
            if sig[0]:
                sg.g_means.copy_(g_means)
            if sig[1]:
                sg.g_per_img[:, 2:3].copy_(g_mae.reshape(B, 1) / float(R * R))
            if sig[2]:
                sg.g_img_in.copy_(g_img)
            if sig[3]:
                sg.g_refl.copy_(g_refl.reshape(B * N, 3))
            if sig[4]:
                sg.g_bounds.copy_(g_bounds)
            graph.replay()
            Fn._LAUNCHES += 1 + (1 if (sig[0] or sig[1]) else 0) + (1 if (sig[0] or sig[1] or sig[2]) else 0)
            g_action = (sg.g_action2d if ctx.two_d else sg.g_action).clone()
            if g_aux is not None:
                g_action += g_aux[:, 3:].reshape(g_action.shape)
        elif g_aux is not None:
            g_action = g_aux[:, 3:].reshape((B, 3 * N) if ctx.two_d else (B, N, 3)).clone()
        else:
            return None, None
        return g_action, None


# ======================================================================================================================
# explicit: step + objective + backward in one graph
# ======================================================================================================================
class GraphedStep:
    """One CUDA graph holding ``env.step(action)`` and the backward of ``objective(metrics)`` to ``action``.

        gstep = GraphedStep(env, objective=lambda m: m["dist"])
        for _ in range(fine_steps):
            loss, grad = gstep(candidate)        # grad = d objective / d candidate, static buffers

    Replays read ``env.sun_pos``, ``env.distance_maps`` and the noisy field's error tensors through the addresses
    captured, so in-place updates of those tensors are seen; *replacing* them (``set_sun_pos``, ``reset`` with
    ``new_errors_every_reset``) needs ``recapture()``.  (``HelioEnv(graph="auto")`` needs none of this: see StepGraph.)
    """

    def __init__(self, env, objective: Optional[Callable[[Dict[str, torch.Tensor]], torch.Tensor]] = None,
                 warmup: int = 3):
        self.env = env
        self.objective = objective or _default_objective
        B, N = env.batch_size, env.num_heliostats
        self.action = torch.zeros(B, N, 3, device=env.device, requires_grad=True)
        self._warmup = warmup
        self.graph = None
        self.recapture()

    def _run(self):
        obs, metrics, monitor = self.env.step(self.action)
        loss = self.objective(metrics)
        grad, = torch.autograd.grad(loss, self.action)
        return obs, metrics, monitor, loss, grad

    def recapture(self, action: Optional[torch.Tensor] = None):
        env = self.env
        if action is None:
            action = env.ideal_normals if getattr(env, "ideal_normals", None) is not None else None
        with torch.no_grad():
            if action is not None:
                self.action.copy_(action.detach().reshape(self.action.shape))
        check, graph_mode = env.check_finite, env.graph
        env.check_finite = False                       # the finite check is the caller's business here (read gstep.metrics)
        env.graph = False                              # the eager fused step is what gets captured here
        try:
            side = torch.cuda.Stream(device=self.action.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):              # warm-up off the default stream: one-time setup
                for _ in range(self._warmup):          # (workspaces, func attributes, cached scene, cached target) happens here
                    self._run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import functional as Fn
            l0 = Fn.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.obs, self.metrics, self.monitor, self.loss, self.grad = self._run()
            self.helio_kernels_per_replay = Fn.launch_count() - l0
        finally:
            env.check_finite, env.graph = check, graph_mode

    def __call__(self, action: torch.Tensor):
        """Replays the step on ``action`` ([B,N,3] or [B,3N]); returns (loss, grad) living in static buffers that
        the next call overwrites (``self.obs`` / ``self.metrics`` / ``self.monitor`` likewise)."""
        with torch.no_grad():
            self.action.copy_(action.detach().reshape(self.action.shape))
        self.graph.replay()
        return self.loss, self.grad
