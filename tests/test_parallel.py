"""CPU, world_size 2, gloo: host-side logic of the ray-sharded data parallelism
(SURVEY.md 8(e)): shard ranges cover the batch exactly once, ray offsets make
the jitter stream independent of the rank count, and the single all-reduce of
the flat gradient arena reproduces the 1-rank gradient."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar4d_b200.parallel import RayShardedDP, flat_grads, shard_range
from oracle.lidar4d_oracle import jitter_uniform


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 1024, 4096, 65536, 67980):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_jitter_stream_is_rank_count_independent():
    full = jitter_uniform(5, np.arange(64), 16)
    for w in (2, 4):
        parts = [jitter_uniform(5, np.arange(*shard_range(64, w, r)), 16) for r in range(w)]
        assert np.array_equal(np.concatenate(parts, 0), full)


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Parameter(torch.randn(40))
        self.b = torch.nn.Parameter(torch.randn(8, 5))

    def loss(self, x, n_global):
        return ((x @ self.b.t()).sum(-1) * self.a[:x.shape[0]].mean()).sum() / n_global


def _worker(rank, world, port, q, use_arena):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = Toy()
    x = torch.arange(30 * 5, dtype=torch.float32).view(30, 5) / 50
    dp = RayShardedDP(m, world_size=world, rank=rank)
    a, b = shard_range(30, world, rank)
    loss = m.loss(x[a:b], 30)
    if use_arena:
        loss.backward()
    else:
        dp.final_backward(loss)       # no CUDA engine behind this model: must degrade to a plain backward
    if use_arena:      # gradients as views of one flat arena, like the CUDA backward returns them
        arena = torch.zeros(128)
        va, vb = arena[:40], arena[64:104].view(8, 5)
        va.copy_(m.a.grad); vb.copy_(m.b.grad)
        m.a.grad, m.b.grad = va, vb
        flat, copy = flat_grads(dp.params())
        assert not copy and flat.data_ptr() == arena.data_ptr()
    dp.allreduce_grads()
    q.put((rank, m.a.grad.clone(), m.b.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def _run(use_arena, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, use_arena)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = Toy()
    x = torch.arange(30 * 5, dtype=torch.float32).view(30, 5) / 50
    # single-rank reference: same global normalisation, chunked the same way
    (m.loss(x[:15], 30) + m.loss(x[15:], 30)).backward()
    for _, ga, gb in res:
        assert torch.allclose(ga, m.a.grad, atol=1e-6) and torch.allclose(gb, m.b.grad, atol=1e-6)


def test_allreduce_matches_single_rank_packed():
    _run(False, 29611)


def test_allreduce_matches_single_rank_arena():
    _run(True, 29612)
