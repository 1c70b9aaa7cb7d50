"""Data-parallel parity on real GPUs (needs >= 2 devices, skipped otherwise): ray-sharded forward + backward with
parallel.OverlappedGradSync over NCCL must give every rank the gradients of the whole batch (the sum of the
shards' gradients), with and without the blur node downstream. The single-GPU run of the full batch is the
reference; fp32 atomics reorder sums -> <= 1e-3 of the largest entry."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, blur, ret):
    import torch.distributed as dist

    from common import load_golden, rel_err
    from gpu_common import default_opt, forward_kwargs, module_from_golden
    from joint_tensorf_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    g = load_golden("cubic_blur" if blur else "cubic_mlp")
    opt = default_opt(g["case"]["shading"], False)

    def grads(module, o, d, kw, w_rgb, w_acc, scale):
        for p in module.parameters():
            p.grad = None
        o = o.clone().requires_grad_(True)
        d = d.clone().requires_grad_(True)
        rgb, _, acc = module.forward(opt, o, d, **kw)
        (((rgb * w_rgb).sum() + (acc * w_acc).sum()) * scale).backward()
        return {k: p.grad.detach().clone() for k, p in module.named_parameters() if p.grad is not None}

    m = module_from_golden(g, dev)
    m.head_precision = "fp32"
    kw = forward_kwargs(g, dev)
    o, d = g["rays_o"].to(dev), g["rays_d"].to(dev)
    wr, wa = g["w_rgb"].to(dev), g["w_acc"].to(dev)
    full = grads(m, o, d, kw, wr, wa, 1.0)                      # whole batch on one GPU, no sync
    lo, hi = parallel.shard_bounds(o.shape[0], rank, world)
    kws = dict(kw)
    if kws.get("jitter") is not None:
        kws["jitter"] = kws["jitter"][lo:hi]
    m.grad_sync = parallel.OverlappedGradSync()
    part = grads(m, o[lo:hi], d[lo:hi], kws, wr[lo:hi], wa[lo:hi], 1.0)
    m.grad_sync.finish([])
    torch.cuda.synchronize()
    worst = max(rel_err(part[k], full[k]) for k in full)
    ret[rank] = (worst, sorted(part) == sorted(full), m.grad_sync.bytes > 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("blur", [False, True])
def test_ray_sharded_gradients_equal_full_batch_gradients_nccl(blur):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(2, _free_port(), blur, ret), nprocs=2, join=True)
    for r in (0, 1):
        worst, same_keys, reduced = ret[r]
        assert same_keys and reduced
        assert worst <= 1e-3, worst
