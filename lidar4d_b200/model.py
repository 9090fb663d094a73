"""Drop-in `LiDAR4D` / `LiDAR_Renderer` nn.Modules backed by liblidar4d_b200.so.

API mirrored from the reference (paths relative to the reference checkout):
  LiDAR4D.__init__ kwargs           model/lidar4d.py:23-45, main_lidar4d.py:155-179
  render(rays_o, rays_d, time, staged, max_ray_batch, **kw) -> dict   model/renderer.py:142-186
  run(rays_o, rays_d, time, num_steps, perturb, **kw) -> dict         model/renderer.py:44-140
  flow(x, t) -> {"forward","backward"}                                model/lidar4d.py:124-137
  density(x, t) -> {"sigma","geo_feat"}                               model/lidar4d.py:139-188
  get_params(lr)                                                      model/lidar4d.py:226-237
  state_dict keys / shapes                                            SURVEY.md 8(b)

There is no PyTorch/CPU implementation of the math in this file: every call
lands in the CUDA library through ctypes and raises if that is impossible.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _capi
from .geometry import FieldConfig, make_frame


# =============================================================================
# parameter containers with the reference's module / parameter names
# =============================================================================
class _TcnnParams(nn.Module):
    """Stands where a tcnn.Encoding / tcnn.Network stood: a flat fp32 `params`
    vector and `n_output_dims` (hash_field.py:137-139, lidar4d.py:84,96,108)."""

    def __init__(self, n_params: int, n_output_dims: int, init: str, **kw):
        super().__init__()
        self.n_output_dims = n_output_dims
        p = torch.empty(n_params)
        if init == "hash":
            p.uniform_(-1e-4, 1e-4)                      # tcnn grid init [tcnn-ext]
        elif init == "mlp":
            o = 0
            dims = kw["dims"]
            for i in range(len(dims) - 1):               # xavier-uniform per [out,in] matrix [tcnn-ext]
                n = dims[i] * dims[i + 1]
                s = math.sqrt(6.0 / (dims[i] + dims[i + 1]))
                p[o:o + n].uniform_(-s, s)
                o += n
        elif init == "empty":
            pass
        self.params = nn.Parameter(p)

    def forward(self, *a, **k):
        raise RuntimeError("encoders/MLPs are fused into the render kernel; call LiDAR4D.render/run/flow/density")


class _HashGridT(nn.Module):
    def __init__(self, n_params, time_resolution, n_output_dims):
        super().__init__()
        self.hash_t = nn.ModuleList([_TcnnParams(n_params, n_output_dims * 4, "hash") for _ in range(time_resolution)])
        self.n_output_dims = n_output_dims


class _HashGrid4D(nn.Module):
    def __init__(self, cfg: FieldConfig):
        super().__init__()
        gs = cfg.static_grid()
        self.hash_static = _TcnnParams(gs.n_params, gs.n_output_dims, "hash")
        self.hash_dynamic = nn.ModuleList([
            _HashGridT(cfg.dynamic_grid(p).n_params, cfg.time_resolution, cfg.hash_dynamic_dim // 3) for p in range(3)])
        self.n_output_dims = cfg.hash_static_dim + cfg.hash_dynamic_dim


class _Planes4D(nn.Module):
    """planes_field.py:144-190: ModuleList over scales of ParameterList over the 6 planes."""

    def __init__(self, cfg: FieldConfig):
        super().__init__()
        self.planes = nn.ModuleList()
        combs = list(itertools.combinations(range(4), 2))
        for mult in cfg.plane_scales:
            reso = [cfg.min_resolution * mult] * 3 + [cfg.time_resolution]
            pl = nn.ParameterList()
            for comb in combs:
                t = torch.empty([1, cfg.n_features_per_level_plane] + [reso[cc] for cc in comb[::-1]])
                if 3 in comb:
                    nn.init.ones_(t)
                else:
                    nn.init.uniform_(t, a=0.1, b=0.5)
                pl.append(nn.Parameter(t))
            self.planes.append(pl)
        self.n_output_dims = 2 * cfg.plane_dim


class _FlowField(nn.Module):
    def __init__(self, cfg: FieldConfig):
        super().__init__()
        gf = cfg.flow_grid()
        self.grid_enc = _TcnnParams(gf.n_params, gf.n_output_dims, "hash")
        h = cfg.hidden_dim_flow
        self.mlp = nn.Sequential(nn.Linear(gf.n_output_dims // 4, h, bias=False), nn.ReLU(),
                                 nn.Linear(h, h, bias=False), nn.ReLU(), nn.Linear(h, 6, bias=False))
        torch.nn.init.normal_(self.mlp[-1].weight.data, 0, 0.001)     # flow_field.py:100


# =============================================================================
# engine: flat arenas, staging, pointer tables, launches
# =============================================================================
ARENA_ALIGN = 1024          # floats; = L4D_ADAM_CHUNK of include/lidar4d_b200.h


class _Engine:
    """Owns what the kernels need besides the caller's tensors:

    * ONE flat fp32 parameter arena - every hot-path parameter's ``.data`` is a view into it (state_dict keys, shapes
      and ``load_state_dict`` are unchanged; after ``model.to(device)`` the views are rebuilt lazily) - and ONE flat
      fp32 gradient arena that every ``.grad`` views: the backward kernels accumulate straight into it (no per-launch
      allocation, no autograd accumulation pass), ``RayShardedDP`` all-reduces it in place, ``lidar4d_b200.optim.Adam``
      updates both arenas in one launch;
    * the staged working set (fp16 tables, channels-last planes, MLP operand copies), refreshed when a master changed;
    * the persistent gradient work buffer (zero between launches: the fold kernels clear what they consume).
    """

    def __init__(self, owner: "LiDAR4D", cfg=None):
        self.owner = owner
        self.cfg = cfg if cfg is not None else owner.cfg
        # default = what the reference does: tiny-cuda-nn's FullyFusedMLP consumes fp16 copies of its fp32 parameters;
        # this is also the mode that runs the dense kernels on the tcgen05 tensor cores
        self.mlp_fp16 = True
        self.ccfg = _capi.make_config(self.cfg, True)
        self.names = _capi.param_names(self.cfg)
        self.lib = None
        self.staged = None
        self._stamp = None
        self._staged_grad_mode = None
        self._chk = None
        self.stage_gen = 0         # bumped by every re-staging; forward records it, backward checks it
        self.static_params = False # True: skip the content check of no-grad calls (tight inference loops)
        self.flat_p = None
        self.flat_g = None
        self._works = {}
        self._tensors = None
        self._mtab = None
        self.offsets = None        # name -> (offset, numel) in floats
        self.total = 0
        self._gtab = None
        self.hash_grads_hook = None  # (torch.cuda.Event, callable): see parallel.RayShardedDP.final_backward
        self.timing = None         # {'fwd': [(ev0, ev1)], 'bwd': [...]}: CUDA events around the fused kernels

    def __deepcopy__(self, memo):
        """copy.deepcopy(model) (some EMA / checkpoint helpers do this): the copy gets a FRESH engine - ctypes pointer
        tables cannot be copied and its arenas are rebuilt lazily around the copied parameters on first use."""
        owner = memo.get(id(self.owner), self.owner)
        new = _Engine.__new__(_Engine)
        memo[id(self)] = new
        _Engine.__init__(new, owner, cfg=self.cfg)       # (the owner's own attributes are still being copied at this point)
        new.set_mlp_fp16(self.mlp_fp16)
        new.static_params = self.static_params
        return new

    def set_mlp_fp16(self, on: bool):
        if bool(on) != self.mlp_fp16:
            self.mlp_fp16 = bool(on)
            self.ccfg = _capi.make_config(self.cfg, self.mlp_fp16)
            self._stamp = None            # force re-staging

    def _lib(self):
        if self.lib is None:
            self.lib = _capi.load_library()
        return self.lib

    def tensors(self) -> Dict[str, torch.Tensor]:
        """name -> Parameter of every tensor the hot path reads (cached: walking named_parameters() costs ~0.1 ms and
        this is called several times per step; LiDAR4D._apply / load_state_dict drop the cache)."""
        ts = self._tensors
        if ts is None:
            sd = dict(self.owner.named_parameters())
            ts = self._tensors = {n: sd[n] for n in self.names}
        return ts

    def drop_caches(self):
        self._tensors, self._mtab = None, None

    def device(self) -> torch.device:
        return self.owner.aabb.device

    def _require_cuda(self, *ts):
        dev = self.device()
        if dev.type != "cuda":
            raise RuntimeError("lidar4d_b200 runs on CUDA (sm_100a) only; move the model with .cuda() "
                               "(there is no CPU fallback)")
        for t in ts:
            if t is not None and t.device != dev:
                raise RuntimeError(f"tensor on {t.device}, model on {dev}")

    def stream(self) -> int:
        return torch.cuda.current_stream(self.device()).cuda_stream

    def _events(self, kind):
        if self.timing is None:
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self.timing[kind].append(ev)
        return ev

    @property
    def n_launches(self) -> int:
        """Kernels this library launched (counted inside the library, l4d_launch_count)."""
        return int(self._lib().l4d_launch_count())

    # ---- flat arenas ------------------------------------------------------------
    def _layout(self, ts):
        if self.offsets is None:
            o, offs = 0, {}
            for n in self.names:
                offs[n] = (o, ts[n].numel())
                o += (ts[n].numel() + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
            self.offsets, self.total = offs, o
        return self.offsets

    def ensure_flat(self, ts=None) -> bool:
        """Make every hot-path parameter a view into the flat arena.  Returns True if the arena was (re)built."""
        ts = ts or self.tensors()
        offs = self._layout(ts)
        dev = self.device()
        if self.flat_p is not None and self.flat_p.device == dev:
            base = self.flat_p.data_ptr()
            if all(t.data_ptr() == base + 4 * offs[n][0] for n, t in ts.items()):
                return False
        for n, t in ts.items():
            if t.dtype != torch.float32 or t.device != dev:
                raise RuntimeError(f"parameter {n} must be fp32 on {dev} (got {t.dtype} on {t.device})")
        flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for n, t in ts.items():
                o, k = offs[n]
                v = flat[o:o + k].view(t.shape)
                v.copy_(t)
                t.data = v
        self.flat_p, self.flat_g, self._gtab, self._stamp, self._mtab = flat, None, None, None, None
        return True

    def attach_grads(self):
        """Point every trainable parameter's .grad at its slice of the flat gradient arena (zeroing slices whose
        .grad was None / copying foreign gradients in) and return the pointer table the kernels accumulate into."""
        ts = self.tensors()
        self.ensure_flat(ts)
        fresh = False
        if self.flat_g is None or self.flat_g.device != self.flat_p.device:
            self.flat_g = torch.zeros(self.total, dtype=torch.float32, device=self.flat_p.device)
            self._gtab, fresh = None, True
        base = self.flat_g.data_ptr()
        stale, n_train = [], 0
        for n, t in ts.items():
            if not t.requires_grad:
                continue
            n_train += 1
            g = t.grad
            if g is None or g.data_ptr() != base + 4 * self.offsets[n][0] or g.dtype != torch.float32 or g.shape != t.shape:
                stale.append((n, t, g))
        if stale:
            all_none = len(stale) == n_train and all(g is None for _, _, g in stale)
            with torch.no_grad():
                if all_none and not fresh:
                    self.flat_g.zero_()                     # the usual case: optimizer.zero_grad(set_to_none=True)
                for n, t, g in stale:
                    o, k = self.offsets[n]
                    v = self.flat_g[o:o + k].view(t.shape)
                    if not all_none:
                        if g is None:
                            v.zero_()
                        else:
                            v.copy_(g)
                    t.grad = v
        if self._gtab is None:
            self._gtab = _capi.L4DMasterGrads()
            _capi.fill_pointer_table(self._gtab, self.cfg, lambda n: base + 4 * self.offsets[n][0])
        return self._gtab

    def work(self):
        """Persistent zero-initialised gradient work buffer (plane / MLP grads in working layout, slice-independent
        hash accumulators) of the CURRENT stream: every kernel that consumes a region clears it again, so launches
        that overlap on different streams need their own (a fold kernel of one must not clear what the scatter of the
        other is still adding)."""
        key = (self.device(), self.stream())
        w = self._works.get(key)
        if w is None:
            n = self._lib().l4d_grad_work_bytes(C.byref(self.ccfg))
            w = self._works[key] = torch.zeros(n, dtype=torch.uint8, device=self.device())
        return w

    def reset_work(self):
        """Call after an exception inside a backward left a work buffer half-consumed."""
        for w in self._works.values():
            w.zero_()

    def infer_workspace(self, n_rays: int, S: int):
        """Persistent exchange workspace of no-grad renders through the split pipeline (grown on demand)."""
        n = self._lib().l4d_saved_bytes(C.byref(self.ccfg), n_rays, S)
        ws = getattr(self, "_infer_ws", None)
        if ws is None or ws.device != self.device() or ws.numel() < n:
            self._infer_ws = ws = torch.empty(n, dtype=torch.uint8, device=self.device())
        return ws, n

    # ---- staging ----------------------------------------------------------------
    def invalidate_staged(self):
        """Force the next call to refresh the kernels' working set.  Needed after in-place writes that bypass
        autograd's version counter in GRAD mode (``p.data.copy_()``); no-grad calls check the content themselves."""
        self._stamp = None

    def _checksum(self) -> int:
        return int(self.flat_p.view(torch.int32).sum(dtype=torch.int64))

    def master_table(self):
        base = self.flat_p.data_ptr()
        if self._mtab is None or self._mtab[0] != base:
            tab = _capi.L4DMasterParams()
            _capi.fill_pointer_table(tab, self.cfg, lambda n: base + 4 * self.offsets[n][0])
            self._mtab = (base, tab)
        return self._mtab[1]

    def ensure_staged(self):
        """Refresh the fp16 / channels-last / transposed working set when a master parameter changed.
        Change detection: autograd version counters (optimizer.step, load_state_dict) in grad mode; in no-grad mode
        additionally the CONTENT (one reduction + one 8-byte read), because torch_ema's copy_to()/restore()
        (runner.py:565-567,680,820) write through ``.data`` and leave the counters alone; and always when the grad mode
        flipped since the last staging (eval -> train after ema.restore())."""
        lib = self._lib()
        ts = self.tensors()
        reflat = self.ensure_flat(ts)
        grad_mode = torch.is_grad_enabled()
        stamp = tuple(t._version for t in ts.values())
        need = (self.staged is None or reflat or stamp != self._stamp or self.staged.device != self.device()
                or grad_mode != self._staged_grad_mode)
        chk = None
        if not need and not grad_mode and not self.static_params:
            chk = self._checksum()
            need = chk != self._chk
        if not need:
            return
        nbytes = lib.l4d_staged_bytes(C.byref(self.ccfg))
        if nbytes == 0:
            raise RuntimeError("unsupported configuration: " + lib.l4d_last_error().decode())
        if self.staged is None or self.staged.device != self.device() or self.staged.numel() != nbytes:
            self.staged = torch.zeros(nbytes, dtype=torch.uint8, device=self.device())
        self.stage(_capi.STAGE_TABLES | _capi.STAGE_SMALL)
        self._stamp, self._staged_grad_mode = stamp, grad_mode
        if not grad_mode and not self.static_params:
            self._chk = chk if chk is not None else self._checksum()

    def stage(self, what: int):
        lib = self._lib()
        tab = self.master_table()
        with torch.cuda.device(self.device()):
            rc = lib.l4d_stage_params_ex(C.byref(self.ccfg), C.byref(tab), self.staged.data_ptr(), self.staged.numel(),
                                         what, self.stream())
        _capi.check(lib, rc, "l4d_stage_params_ex")
        self.stage_gen += 1

    def params_updated_by_optimizer(self, tables_staged: bool):
        """lidar4d_b200.optim.Adam wrote the parameter arena (and, fused, the fp16 tables): finish the refresh."""
        if self.staged is None:
            self._stamp = None
            return
        self.stage(_capi.STAGE_SMALL if tables_staged else (_capi.STAGE_TABLES | _capi.STAGE_SMALL))
        self._chk = None
        self._staged_grad_mode = True

    def check_generation(self, ctx_gen: int, what: str):
        if ctx_gen != self.stage_gen:
            raise RuntimeError(
                f"{what}: the model's parameters (or its MLP precision mode) changed between this forward and its "
                "backward; the kernels' working set was re-staged in between.  Run backward before optimizer.step() / "
                "set_mlp_fp16() / load_state_dict().")

    def frame(self, time) -> _capi.L4DFrame:
        if torch.is_tensor(time):
            time = float(time.reshape(-1)[0])          # same host sync as lidar4d.py:143
        return _capi.make_frame_struct(make_frame(time, self.cfg.num_frames, self.cfg.time_resolution))


class _RenderFn(torch.autograd.Function):
    """Fused render forward / backward.  Parameters are passed as inputs so that autograd schedules the backward;
    their gradients are accumulated by the kernels straight into the flat gradient arena that every ``.grad`` views
    (model.grad_mode == "arena", the default), or returned to autograd as fresh tensors ("autograd": works with
    torch.autograd.grad, costs one arena allocation + one accumulation pass per backward)."""

    @staticmethod
    def forward(ctx, eng: _Engine, rays_o, rays_d, frame, S, perturb, seed, ray_offset, want_weights, train, *params):
        lib = eng._lib()
        N = rays_o.shape[0]
        dev = rays_o.device
        depth = torch.empty(N, device=dev)
        image = torch.empty(N, 2, device=dev)
        wsum = torch.empty(N, device=dev)
        weights = torch.empty(N, S, device=dev) if want_weights else None
        zvals = torch.empty(N, S, device=dev) if want_weights else None
        rays = _capi.L4DRays()
        rays.rays_o, rays.rays_d = rays_o.data_ptr(), rays_d.data_ptr()
        rays.n_rays, rays.n_steps, rays.perturb = N, S, int(bool(perturb))
        rays.seed, rays.ray_offset = int(seed), int(ray_offset)
        fused = eng.owner.pipeline == "fused"
        rays.reserved = 1 if fused else 0
        saved, nsaved = None, 0
        chunk = N
        if train:
            nsaved = lib.l4d_saved_bytes(C.byref(eng.ccfg), N, S)
            saved = torch.empty(nsaved, dtype=torch.uint8, device=dev)
        elif not fused:
            # inference through the split pipeline: the kernels exchange feature tiles through a workspace; walk the rays
            # in chunks over ONE persistent workspace (config 4: 131,072 rays per frame would otherwise need 160 GB)
            chunk = min(N, int(eng.owner.infer_ray_chunk))
            saved, nsaved = eng.infer_workspace(chunk, S)
        ev = eng._events("fwd")
        with torch.cuda.device(dev):
            if ev: ev[0].record()
            for h in range(0, N, chunk):
                n = min(chunk, N - h)
                rays.rays_o, rays.rays_d, rays.n_rays = rays_o.data_ptr() + 12 * h, rays_d.data_ptr() + 12 * h, n
                rays.ray_offset = int(ray_offset) + h
                rc = lib.l4d_render_forward(
                    C.byref(eng.ccfg), eng.staged.data_ptr(), C.byref(frame), C.byref(rays),
                    depth.data_ptr() + 4 * h, image.data_ptr() + 8 * h, wsum.data_ptr() + 4 * h,
                    weights.data_ptr() + 4 * h * S if want_weights else None, zvals.data_ptr() + 4 * h * S if want_weights else None,
                    saved.data_ptr() if saved is not None else None, nsaved, eng.stream())
                _capi.check(lib, rc, "l4d_render_forward")
            if ev: ev[1].record()
        ctx.eng, ctx.frame, ctx.rays_args = eng, frame, (N, S, int(bool(perturb)), int(seed), int(ray_offset))
        ctx.saved, ctx.nsaved = (saved, nsaved) if train else (None, 0)
        ctx.fused = fused
        ctx.rays_o, ctx.rays_d = rays_o, rays_d
        ctx.stage_gen = eng.stage_gen
        ctx.want_weights = want_weights
        if want_weights:
            ctx.mark_non_differentiable(zvals)
            return depth, image, wsum, weights, zvals
        return depth, image, wsum

    @staticmethod
    def backward(ctx, g_depth, g_image, g_wsum, *rest):
        eng = ctx.eng
        lib = eng._lib()
        if ctx.saved is None:
            raise RuntimeError("render was run without saving activations (no_grad / eval)")
        eng.check_generation(ctx.stage_gen, "render backward")
        N, S, perturb, seed, ray_offset = ctx.rays_args
        dev = ctx.rays_o.device
        g_weights = rest[0] if (ctx.want_weights and len(rest) > 0) else None
        z = lambda t, shape: (torch.zeros(shape, device=dev) if t is None else t.contiguous().float())
        g_depth, g_image = z(g_depth, (N,)), z(g_image, (N, 2))
        g_wsum = None if g_wsum is None else g_wsum.contiguous().float()
        g_weights = None if g_weights is None else g_weights.contiguous().float()
        arena = eng.owner.grad_mode == "arena"
        tab, views = (eng.attach_grads(), None) if arena else _scratch_grads(eng)
        work = eng.work()
        rays = _capi.L4DRays()
        rays.rays_o, rays.rays_d = ctx.rays_o.data_ptr(), ctx.rays_d.data_ptr()
        rays.n_rays, rays.n_steps, rays.perturb, rays.seed, rays.ray_offset = N, S, perturb, seed, ray_offset
        rays.reserved = 1 if ctx.fused else 0
        ev = eng._events("bwd")
        with torch.cuda.device(dev):
            if ev: ev[0].record()
            hook = eng.hash_grads_hook if arena else None      # (event, callback) armed by RayShardedDP for the step's last backward
            rc = lib.l4d_render_backward_ex(
                C.byref(eng.ccfg), eng.staged.data_ptr(), C.byref(ctx.frame), C.byref(rays),
                ctx.saved.data_ptr(), ctx.nsaved, g_depth.data_ptr(), g_image.data_ptr(),
                g_wsum.data_ptr() if g_wsum is not None else None,
                g_weights.data_ptr() if g_weights is not None else None,
                C.byref(tab), work.data_ptr(), work.numel(), hook[0].cuda_event if hook else None, eng.stream())
            if ev: ev[1].record()
            _capi.check(lib, rc, "l4d_render_backward")
            if hook:
                eng.hash_grads_hook = None
                hook[1]()
            rc = lib.l4d_unstage_grads(C.byref(eng.ccfg), work.data_ptr(), work.numel(), C.byref(tab), eng.stream())
            _capi.check(lib, rc, "l4d_unstage_grads")
        ctx.saved = None
        if arena:
            return (None,) * (10 + len(eng.names))
        return (None,) * 10 + tuple(views[n] for n in eng.names)


def _scratch_grads(eng: _Engine):
    """grad_mode == "autograd": a fresh zeroed arena whose views are handed to autograd."""
    ts = eng.tensors()
    eng.ensure_flat(ts)
    flat = torch.zeros(eng.total, dtype=torch.float32, device=eng.device())
    views = {n: flat[o:o + k].view(ts[n].shape) for n, (o, k) in eng.offsets.items()}
    tab = _capi.L4DMasterGrads()
    _capi.fill_pointer_table(tab, eng.cfg, lambda n: views[n].data_ptr())
    return tab, views


class _FlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: _Engine, x, frame, train, *params):
        lib = eng._lib()
        n = x.shape[0]
        flow = torch.empty(n, 6, device=x.device)
        saved = torch.empty(16, n, device=x.device) if train else None
        with torch.cuda.device(x.device):
            rc = lib.l4d_flow_forward(C.byref(eng.ccfg), eng.staged.data_ptr(), C.byref(frame), x.data_ptr(), n,
                                      flow.data_ptr(), saved.data_ptr() if train else None, eng.stream())
        _capi.check(lib, rc, "l4d_flow_forward")
        ctx.eng, ctx.frame, ctx.x, ctx.saved, ctx.stage_gen = eng, frame, x, saved, eng.stage_gen
        return flow

    @staticmethod
    def backward(ctx, g_flow):
        eng = ctx.eng
        lib = eng._lib()
        if ctx.saved is None:
            raise RuntimeError("flow was run without saving activations")
        eng.check_generation(ctx.stage_gen, "flow backward")
        x = ctx.x
        n = x.shape[0]
        g_flow = g_flow.contiguous().float()
        arena = eng.owner.grad_mode == "arena"
        tab, views = (eng.attach_grads(), None) if arena else _scratch_grads(eng)
        work = eng.work()
        with torch.cuda.device(x.device):
            rc = lib.l4d_flow_backward(C.byref(eng.ccfg), eng.staged.data_ptr(), C.byref(ctx.frame), x.data_ptr(), n,
                                       ctx.saved.data_ptr(), g_flow.data_ptr(), C.byref(tab), work.data_ptr(), work.numel(),
                                       eng.stream())
            _capi.check(lib, rc, "l4d_flow_backward")
            rc = lib.l4d_unstage_grads(C.byref(eng.ccfg), work.data_ptr(), work.numel(), C.byref(tab), eng.stream())
            _capi.check(lib, rc, "l4d_unstage_grads")
        if arena:
            return (None,) * (4 + len(eng.names))
        flow_names = {"flow_net.grid_enc.params", "flow_net.mlp.0.weight", "flow_net.mlp.2.weight", "flow_net.mlp.4.weight"}
        return (None,) * 4 + tuple(views[n] if n in flow_names else None for n in eng.names)


# =============================================================================
# public modules
# =============================================================================
class LiDAR_Renderer(nn.Module):
    """model/renderer.py:13-186 API."""

    def __init__(self, bound=1, near_lidar=0.01, far_lidar=0.81, density_scale=1, active_sensor=False):
        super().__init__()
        self.bound = bound
        self.near_lidar = near_lidar
        self.far_lidar = far_lidar
        self.density_scale = density_scale
        self.active_sensor = active_sensor
        aabb = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer("aabb", aabb)

    def forward(self, x, d):
        raise NotImplementedError()

    def run(self, rays_o, rays_d, time, num_steps=768, perturb=False, **kwargs):
        raise NotImplementedError()

    def render(self, rays_o, rays_d, time, staged=False, max_ray_batch=4096, **kwargs):
        """renderer.py:142-186.  `staged` keeps the reference's chunked contract
        (only depth/image returned); chunks may be larger here because nothing of
        size [N*S, C] is materialised."""
        B, N = rays_o.shape[:2]
        device = rays_o.device
        if staged:
            depth = torch.empty((B, N), device=device)
            image = torch.empty((B, N, self.out_lidar_dim), device=device)
            for b in range(B):
                head = 0
                while head < N:
                    tail = min(head + max_ray_batch, N)
                    r = self.run(rays_o[b:b + 1, head:tail], rays_d[b:b + 1, head:tail], time[b:b + 1], **kwargs)
                    depth[b:b + 1, head:tail] = r["depth_lidar"]
                    image[b:b + 1, head:tail] = r["image_lidar"]
                    head += max_ray_batch
            return {"depth_lidar": depth, "image_lidar": image}
        return self.run(rays_o, rays_d, time, **kwargs)


class LiDAR4D(LiDAR_Renderer):
    def __init__(self, min_resolution=32, base_resolution=512, max_resolution=32768, time_resolution=8,
                 n_levels_plane=4, n_features_per_level_plane=8, n_levels_hash=8, n_features_per_level_hash=4,
                 log2_hashmap_size=19, num_layers_flow=3, hidden_dim_flow=64, num_layers_sigma=2,
                 hidden_dim_sigma=64, geo_feat_dim=15, num_layers_lidar=3, hidden_dim_lidar=64, out_lidar_dim=2,
                 num_frames=51, bound=1, unet: Optional[nn.Module] = None, **kwargs):
        renderer_kw = {k: kwargs[k] for k in ("near_lidar", "far_lidar", "density_scale", "active_sensor") if k in kwargs}
        super().__init__(bound, **renderer_kw)
        self.cfg = FieldConfig(
            min_resolution=min_resolution, base_resolution=base_resolution, max_resolution=max_resolution,
            time_resolution=time_resolution, n_levels_plane=n_levels_plane,
            n_features_per_level_plane=n_features_per_level_plane, n_levels_hash=n_levels_hash,
            n_features_per_level_hash=n_features_per_level_hash, log2_hashmap_size=log2_hashmap_size,
            hash_size_dynamic=tuple(kwargs.get("hash_size_dynamic", (15, 13, 13))),
            num_layers_flow=num_layers_flow, hidden_dim_flow=hidden_dim_flow, num_layers_sigma=num_layers_sigma,
            hidden_dim_sigma=hidden_dim_sigma, geo_feat_dim=geo_feat_dim, num_layers_lidar=num_layers_lidar,
            hidden_dim_lidar=hidden_dim_lidar, out_lidar_dim=out_lidar_dim, num_frames=num_frames, bound=float(bound),
            near_lidar=float(self.near_lidar), far_lidar=float(self.far_lidar), density_scale=float(self.density_scale),
            active_sensor=bool(self.active_sensor),
            **{k: kwargs[k] for k in ("flow_base_resolution", "flow_max_resolution", "flow_log2_hashmap_size") if k in kwargs})
        c = self.cfg
        c.validate()
        self.out_lidar_dim = out_lidar_dim
        self.num_frames = num_frames

        self.planes_encoder = _Planes4D(c)
        self.hash_encoder = _HashGrid4D(c)
        self.view_encoder = _TcnnParams(0, c.view_dim, "empty")
        self.flow_net = _FlowField(c)
        self.sigma_net = _TcnnParams(c.mlp_param_count(c.sigma_in_pad, 64, 1), 1 + geo_feat_dim, "mlp",
                                     dims=[c.sigma_in_pad, 64, 16])
        self.intensity_net = _TcnnParams(c.mlp_param_count(c.attr_in_pad, 64, 2), 1, "mlp", dims=[c.attr_in_pad, 64, 64, 16])
        self.raydrop_net = _TcnnParams(c.mlp_param_count(c.attr_in_pad, 64, 2), 1, "mlp", dims=[c.attr_in_pad, 64, 64, 16])
        # U-Net ray-drop refinement is a per-image post-process outside the hot path;
        # callers that need it pass the reference's module in (model/unet.py), its state_dict keys stay `unet.*`.
        self.unet = unet if unet is not None else nn.Identity()

        self.materialize_weights = True    # return `weights` / `z_vals` like renderer.py:134-140
        # "split": gather / dense / scatter kernels exchanging SoA planes (default, faster);
        # "fused": the single-kernel forward and backward (csrc/l4d_kernels.cu)
        self.pipeline = "split"
        # "arena": the backward kernels accumulate into the flat gradient arena that every .grad views (default);
        # "autograd": gradients are returned to autograd as tensors (torch.autograd.grad, hooks), slower
        self.grad_mode = "arena"
        self.infer_ray_chunk = 2048        # rays per launch of a no-grad render (bounds the exchange workspace: 2.5 GB at L=16)
        self.jitter_seed = 0
        self._jitter_calls = 0
        self._engine = _Engine(self)

    # ---- precision / pipeline switches ------------------------------------------------------
    def set_mlp_fp16(self, on: bool = True):
        """True: MLP weights are consumed as fp16-rounded working copies of the fp32 masters (exactly how
        tiny-cuda-nn's FullyFusedMLP holds them) and the dense kernels of the split pipeline run on the
        tcgen05 tensor cores with hi/lo-split fp16 activations (fp32-class accuracy) -- the default.
        False: fp32 weights on the fp32-FMA kernels (bit-for-bit the fp32 oracle's arithmetic, ~2x slower)."""
        self._engine.set_mlp_fp16(on)
        return self

    def invalidate_staged(self):
        """Refresh the kernels' working set at the next call.  Only needed after in-place parameter writes that bypass
        autograd's version counter WHILE GRADIENTS ARE ENABLED (``p.data.copy_(...)``); optimizer steps,
        ``load_state_dict`` and anything done before a ``torch.no_grad()`` render (torch_ema ``copy_to``/``restore``
        around evaluation, runner.py:565-567,680) are detected automatically."""
        self._engine.invalidate_staged()
        return self

    def prepare_grads(self):
        """Attach every parameter's .grad to the flat gradient arena (clearing it if the grads were None) on the current
        stream.  Call once after ``zero_grad`` when the backward passes of one optimiser step are spread over several
        CUDA streams, before those streams fork: otherwise the first backward to run would clear the arena on ITS stream."""
        self._engine._require_cuda()
        self._engine.ensure_staged()          # a pending re-staging must not happen on one of the forked streams either
        self._engine.attach_grads()
        return self

    def static_params(self, on: bool = True):
        """Promise that parameters do not change between no-grad calls (skips the per-call content check, one
        reduction + an 8-byte device->host read): for tight inference loops such as a trajectory render."""
        self._engine.static_params = bool(on)
        return self

    # ---- nn.Module plumbing -----------------------------------------------------------------
    def forward(self, x, d, t):
        pass

    def _apply(self, fn, *a, **k):          # .to() / .cuda() / .float(): parameters may be re-created
        r = super()._apply(fn, *a, **k)
        if hasattr(self, "_engine"):
            self._engine.drop_caches()
        return r

    def _params_list(self) -> List[torch.Tensor]:
        ts = self._engine.tensors()
        return [ts[n] for n in self._engine.names]

    # ---- LiDAR_Renderer.run (renderer.py:44-140) -----------------------------------------------
    def run(self, rays_o, rays_d, time, num_steps=768, perturb=False, ray_offset=0, **kwargs):
        eng = self._engine
        prefix = rays_o.shape[:-1]
        ro = rays_o.detach().contiguous().view(-1, 3).float()
        rd = rays_d.detach().contiguous().view(-1, 3).float()
        eng._require_cuda(ro, rd)
        eng.ensure_staged()
        frame = eng.frame(time)
        params = self._params_list()
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        seed = 0
        if perturb:
            seed = self.jitter_seed + self._jitter_calls
            self._jitter_calls += 1
        want_w = bool(self.materialize_weights)
        out = _RenderFn.apply(eng, ro, rd, frame, int(num_steps), bool(perturb), seed, int(ray_offset), want_w, train, *params)
        res = {"depth_lidar": out[0].view(*prefix), "image_lidar": out[1].view(*prefix, self.out_lidar_dim),
               "weights_sum_lidar": out[2]}
        if want_w:
            res["weights"], res["z_vals"] = out[3], out[4]
        return res

    # ---- LiDAR4D.flow (lidar4d.py:124-137) -------------------------------------------------------
    def flow(self, x, t):
        eng = self._engine
        x = x.detach().contiguous().view(-1, 3).float()
        eng._require_cuda(x)
        eng.ensure_staged()
        frame = eng.frame(t)
        params = self._params_list()
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        fl = _FlowFn.apply(eng, x, frame, train, *params)
        return {"forward": fl[:, :3], "backward": fl[:, 3:]}

    # ---- LiDAR4D.density (lidar4d.py:139-188), inference / parity only ---------------------------
    @torch.no_grad()
    def density(self, x, t=None, return_features: bool = False):
        eng = self._engine
        lib = eng._lib()
        x = x.detach().contiguous().view(-1, 3).float()
        eng._require_cuda(x)
        eng.ensure_staged()
        frame = eng.frame(t)
        n = x.shape[0]
        sigma = torch.empty(n, device=x.device)
        geo = torch.empty(n, 15, device=x.device)
        feats = torch.empty(n, self.cfg.sigma_in_dim, device=x.device) if return_features else None
        flow = torch.empty(n, 6, device=x.device) if return_features else None
        with torch.cuda.device(x.device):
            rc = lib.l4d_density_forward(C.byref(eng.ccfg), eng.staged.data_ptr(), C.byref(frame), x.data_ptr(), n,
                                         sigma.data_ptr(), geo.data_ptr(),
                                         feats.data_ptr() if return_features else None,
                                         flow.data_ptr() if return_features else None, eng.stream())
        _capi.check(lib, rc, "l4d_density_forward")
        out = {"sigma": sigma, "geo_feat": geo}
        if return_features:
            out.update(features=feats, flow=flow)
        return out

    # ---- LiDAR4D.attribute (lidar4d.py:191-223), inference / parity only; render() has the heads fused in ----
    @torch.no_grad()
    def attribute(self, x, d, mask=None, geo_feat=None, **kwargs):
        """x is unused (as in the reference); d [N,3] unit directions, geo_feat [N,15], mask [N] bool or None.
        Returns [N,2] = (raydrop, intensity); rows outside the mask are zero."""
        eng = self._engine
        lib = eng._lib()
        d = d.detach().contiguous().view(-1, 3).float()
        eng._require_cuda(d)
        eng.ensure_staged()
        n = d.shape[0]
        geo = geo_feat.detach().contiguous().view(n, 15).float()
        m8 = None if mask is None else mask.detach().contiguous().view(n).to(torch.uint8)
        out = torch.empty(n, self.out_lidar_dim, device=d.device)
        with torch.cuda.device(d.device):
            rc = lib.l4d_attribute_forward(C.byref(eng.ccfg), eng.staged.data_ptr(), d.data_ptr(), geo.data_ptr(),
                                           m8.data_ptr() if m8 is not None else None, n, out.data_ptr(), eng.stream())
        _capi.check(lib, rc, "l4d_attribute_forward")
        return out

    @torch.no_grad()
    def hash_indices(self, grid_id: int, level: int, x: torch.Tensor):
        """Parity hook: uint32 corner indices + weights of one level (grid_id 0 static, 1-3 dynamic, 4 flow)."""
        eng = self._engine
        lib = eng._lib()
        x = x.detach().contiguous().float()
        eng._require_cuda(x)
        n, D = x.shape
        idx = torch.empty(n, 1 << D, dtype=torch.int32, device=x.device)
        w = torch.empty(n, 1 << D, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.l4d_hash_indices(C.byref(eng.ccfg), grid_id, level, x.data_ptr(), n, idx.data_ptr(), w.data_ptr(),
                                      eng.stream())
        _capi.check(lib, rc, "l4d_hash_indices")
        return idx, w

    # ---- optimizer utils (lidar4d.py:226-237) ----------------------------------------------------
    def get_params(self, lr):
        return [
            {"params": self.planes_encoder.parameters(), "lr": lr},
            {"params": self.hash_encoder.parameters(), "lr": lr},
            {"params": self.view_encoder.parameters(), "lr": lr},
            {"params": self.flow_net.parameters(), "lr": 0.1 * lr},
            {"params": self.sigma_net.parameters(), "lr": 0.1 * lr},
            {"params": self.intensity_net.parameters(), "lr": 0.1 * lr},
            {"params": self.raydrop_net.parameters(), "lr": 0.1 * lr},
        ]

    @property
    def gpu_launches(self) -> int:
        return self._engine.n_launches
