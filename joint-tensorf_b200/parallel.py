"""Ray-sharded data parallelism.

One process per GPU; every rank holds a replica of the VM factors, basis_mat, the
shading head and the per-image pose table, renders its own contiguous slice of
the step's ray list, and the gradients are summed with ONE collective over a
flat fp32 bucket (NCCL over NVLink/NVSwitch on the GPU box; gloo in the CPU
tests). The reference has no multi-GPU path (options.py:126); this is the
B200-native addition named by the north star.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_o, rays_d, rank, world, *extra):
    lo, hi = shard_bounds(rays_o.shape[0], rank, world)
    return (rays_o[lo:hi], rays_d[lo:hi], *[e[lo:hi] for e in extra])


def frames_of_rank(n_frames, rank, world):
    """Image-sharded full-frame rendering: rank r renders frames r, r+W, ..."""
    return list(range(rank, n_frames, world))


class GradBucket:
    """Flat fp32 bucket over the gradients of `params`.

    `attach()` points every p.grad at a view of one contiguous buffer (keeping each
    parameter's strides, so channel-last factors stay channel-last) -- autograd
    then accumulates straight into the bucket and `all_reduce()` is a single
    collective with no flatten/unflatten copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros((sum(self.sizes),), device=dev, dtype=torch.float32)
        self.views = []
        o = 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(torch.as_strided(self.flat, p.shape, p.stride(), o))
            o += n

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None, average=True, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average and not async_op:
            self.flat.div_(dist.get_world_size(group))
        return work


class OverlappedGradSync:
    """Sum the gradients of a step across ranks while its backward is still running.

    `B200_VMSplit.grad_sync = OverlappedGradSync()` makes the render node's backward hand over its flat gradient
    bucket in two parts: the appearance-factor gradients (3/4 of the bytes, complete after the appearance scatter)
    are all-reduced on NCCL's stream while the density scatter runs; the rest (density factors, basis_mat, head)
    follows and the backward's stream waits for both before the node returns, so whatever consumes the gradients
    next (p.grad, the adjoint blur -- linear, so reducing before it is the same sum) sees cross-rank sums.
    `finish(extra)` reduces the few gradients produced by later autograd nodes (se3_refine). Sums only: scale the
    loss by 1/world_size for a mean (no extra pass over the bucket). World size 1: every call is a no-op."""

    def __init__(self, group=None):
        self.group = group
        self.works = []
        self.bytes = 0

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, t):
        if self._active() and t.numel() > 0:
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.bytes += t.numel() * t.element_size()

    def on_app_grads(self, region):
        self._reduce(region)

    def on_rest(self, region):
        self._reduce(region)
        self.wait()

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []

    def finish(self, extra=()):
        """extra: parameters whose .grad was produced outside the render node (e.g. se3_refine)."""
        gs = [p.grad for p in extra if p.grad is not None]
        if gs and self._active():
            cat = torch.cat([g.reshape(-1) for g in gs])
            self._reduce(cat)
            self.wait()
            o = 0
            for g in gs:
                g.copy_(cat[o:o + g.numel()].view_as(g))
                o += g.numel()
        self.wait()


def seed_all_ranks(seed):
    """Host RNG draws the reference makes per step (np.random.choice blur scale
    tensorf.py:198, torch.rand bg flip batBase.py:154, randperm ray_idx) must be
    identical on all ranks: seed them identically."""
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
