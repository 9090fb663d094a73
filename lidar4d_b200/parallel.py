"""Ray-sharded data parallelism for the LiDAR4D hot path.

The reference has no distributed code at all.  Rays are
independent (compositing is intra-ray, model/renderer.py:98-129), parameters
are replicated, so the only exchange is ONE all-reduce (sum) of the flat fp32
gradient arena per optimiser step; there is no data-path collective.

One process per GPU (torchrun); NCCL over NVLink/NVSwitch on the GPU box, gloo
in the CPU tests (tests/test_parallel.py).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rays: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous 1/R chunk of a ray batch; the start doubles as `ray_offset`,
    so jitter streams (seed, global ray index, sample) do not depend on R."""
    base, rem = divmod(n_rays, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def flat_grads(params: List[torch.Tensor]):
    """Return (flat, needs_copy_back).  When every .grad is a view into ONE
    storage (the flat arena the backward kernel accumulated into, see
    model._Engine.attach_grads) that storage is reduced in place - a single
    collective with no packing copies; otherwise the gradients are packed."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None, False
    st = grads[0].untyped_storage()
    same = all(g.untyped_storage().data_ptr() == st.data_ptr() and g.dtype == torch.float32 and g.is_contiguous()
               for g in grads)
    if same and st.nbytes() % 4 == 0:
        flat = torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, 0, (st.nbytes() // 4,))
        return flat, False
    return torch.cat([g.reshape(-1).float() for g in grads]), True


class RayShardedDP:
    """Thin wrapper: shard rays, all-reduce gradients once per step."""

    def __init__(self, model: torch.nn.Module, world_size: Optional[int] = None, rank: Optional[int] = None):
        self.model = model
        self.world_size = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        self._comm, self._event, self._early = None, None, None

    def params(self) -> List[torch.Tensor]:
        return [p for p in self.model.parameters() if p.requires_grad and p.numel() > 0]

    def shard(self, rays_o: torch.Tensor, rays_d: torch.Tensor):
        """[1,N,3] rays -> this rank's contiguous slice and its global ray offset."""
        n = rays_o.shape[-2]
        a, b = shard_range(n, self.world_size, self.rank)
        return rays_o[..., a:b, :], rays_d[..., a:b, :], a

    def render(self, rays_o, rays_d, time, **kw):
        """Render this rank's shard; losses must be normalised by the GLOBAL ray
        count (runner.py:213,219 use .sum()/.mean() over rays) so that the summed
        gradients equal the single-GPU gradient."""
        ro, rd, off = self.shard(rays_o, rays_d)
        return self.model.render(ro, rd, time, ray_offset=off, **kw)

    # ---- overlapped reduction -----------------------------------------------------------------------------------
    def final_backward(self, loss: torch.Tensor) -> None:
        """`loss.backward()` for the LAST backward of an optimiser step.  The backward kernels record an event once the
        hash_static / hash_dynamic gradients are final (60 % of the gradient bytes; the flow-net backward still has to
        run); that bucket's all-reduce starts behind the event on a side stream and overlaps the rest of the backward.
        `allreduce_grads()` afterwards reduces the remaining buckets and joins the side stream."""
        eng = getattr(self.model, "_engine", None)
        if self.world_size == 1 or eng is None or not loss.is_cuda or getattr(self.model, "grad_mode", "") != "arena":
            loss.backward()
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=loss.device)
            self._event = torch.cuda.Event()
            self._event.record()                      # materialise the cudaEvent_t handle
        self._early = None

        def start_bucket():
            lo = eng.offsets["hash_encoder.hash_static.params"][0]
            hi = eng.offsets["flow_net.grid_enc.params"][0]
            self._comm.wait_event(self._event)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(eng.flat_g[lo:hi], op=dist.ReduceOp.SUM)
            self._early = (lo, hi)

        eng.hash_grads_hook = (self._event, start_bucket)
        loss.backward()
        eng.hash_grads_hook = None                    # not consumed (e.g. the loss did not reach the renderer)

    @torch.no_grad()
    def allreduce_grads(self) -> None:
        if self.world_size == 1:
            return
        params = self.params()
        eng = getattr(self.model, "_engine", None)
        early, self._early = getattr(self, "_early", None), None
        if early is not None and eng is not None and eng.flat_g is not None:
            lo, hi = early
            dist.all_reduce(eng.flat_g[:lo], op=dist.ReduceOp.SUM)          # planes
            dist.all_reduce(eng.flat_g[hi:], op=dist.ReduceOp.SUM)          # flow grid + MLPs
            torch.cuda.current_stream().wait_stream(self._comm)
            return
        # a parameter that got no gradient on this rank still takes part in the reduction
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        flat, copy_back = flat_grads(params)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if copy_back:
            o = 0
            for p in params:
                n = p.grad.numel()
                p.grad.copy_(flat[o:o + n].view_as(p.grad))
                o += n
