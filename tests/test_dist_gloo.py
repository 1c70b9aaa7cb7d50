"""world_size-2 gloo test of the ray-sharding + flat-bucket gradient all-reduce logic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from joint_tensorf_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # a channel-last "plane", a line and a dense weight, identical on both ranks
    plane = torch.nn.Parameter(torch.randn(1, 8, 6, 5).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    line = torch.nn.Parameter(torch.randn(1, 8, 7, 1))
    w = torch.nn.Parameter(torch.randn(3, 4))
    bucket = parallel.GradBucket([plane, line, w])
    bucket.attach()
    # every rank "renders" its shard of 10 rays: loss = sum over its rays
    rays = torch.arange(10, dtype=torch.float32)
    (mine,) = parallel.shard_rays(rays, rays, rank, world)[:1]
    loss = (plane.sum() + 2 * line.sum() + 3 * w.sum()) * mine.sum()
    loss.backward()
    assert plane.grad.data_ptr() == bucket.views[0].data_ptr(), "autograd must accumulate into the bucket"
    assert plane.grad.stride() == plane.stride()
    bucket.all_reduce(average=False)
    total = float(rays.sum())
    ok = (torch.allclose(plane.grad, torch.full_like(plane, total)) and
          torch.allclose(line.grad, torch.full_like(line, 2 * total)) and
          torch.allclose(w.grad, torch.full_like(w, 3 * total)))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _worker_overlap(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sync = parallel.OverlappedGradSync()
    flat = torch.arange(20, dtype=torch.float32) * (rank + 1)          # the render node's flat gradient bucket
    sync.on_app_grads(flat[:12])                                        # appearance part first (async)
    sync.on_rest(flat[12:])                                             # the rest, then wait for both
    se3 = torch.nn.Parameter(torch.zeros(4, 6))
    se3.grad = torch.full((4, 6), float(rank + 1))
    frozen = torch.nn.Parameter(torch.zeros(3))                         # no gradient: skipped
    sync.finish([se3, frozen])
    tot = sum(r + 1 for r in range(world))
    ok = torch.equal(flat, torch.arange(20, dtype=torch.float32) * tot) and \
        torch.equal(se3.grad, torch.full((4, 6), float(tot))) and not sync.works
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_grad_sync_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_overlap, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_overlapped_grad_sync_is_a_noop_without_a_process_group():
    sync = parallel.OverlappedGradSync()
    flat = torch.ones(8)
    sync.on_app_grads(flat[:4])
    sync.on_rest(flat[4:])
    sync.finish([])
    assert torch.equal(flat, torch.ones(8)) and sync.bytes == 0


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.frames_of_rank(10, 1, 4) == [1, 5, 9]


def test_gradient_bucket_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


class _TinyModel(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        # a channel-last factor (as B200_VMSplit stores them) and a dense weight, DIFFERENT on every rank
        p = torch.randn(1, 4, 3, 5, generator=g)
        self.plane = torch.nn.Parameter(p.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
        self.w = torch.nn.Parameter(torch.randn(3, 2, generator=g))
        self.grad_sync = None


def _worker_attach(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _TinyModel(seed=10 + rank)
    se3 = torch.nn.Parameter(torch.full((2, 6), float(rank)))
    ref = _TinyModel(seed=10)                                  # what rank 0 holds
    sync = parallel.OverlappedGradSync().attach(m, [se3])
    ok = m.grad_sync is sync and torch.equal(m.plane, ref.plane) and torch.equal(m.w, ref.w) and \
        torch.equal(se3, torch.zeros(2, 6)) and m.plane.stride() == ref.plane.stride()
    # paused(): no collective is issued (a rank-local backward must not wait for the other ranks)
    flat = torch.full((6,), float(rank + 1))
    with sync.paused():
        sync.on_app_grads(flat[:3])
        sync.on_rest(flat[3:])
        sync.finish([se3])
    ok = ok and torch.equal(flat, torch.full((6,), float(rank + 1))) and sync._active()
    # the split-backward policy: 2 ranks -> single launch; forced on / off
    ok = ok and not sync.use_split21(90, 100)
    ok = ok and parallel.OverlappedGradSync(split21=True).use_split21(10, 100)
    ok = ok and not parallel.OverlappedGradSync(split21=False).use_split21(90, 100)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_attach_broadcasts_and_paused_skips_collectives_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_attach, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
