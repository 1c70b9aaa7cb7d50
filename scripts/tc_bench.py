"""Time the tensor-core head kernels in isolation (CUDA events, L2 flushed between runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from joint_tensorf_b200 import ops
from oracle import vm_oracle as vo

A = int(os.environ.get("A", 2265497))
dev = "cuda"
p = vo.init_params([8, 8, 8], [16] * 3, [48] * 3, 27, "MLP_Fea", 64, 2, 2, 0.1, 0.0, seed=0)
names = ["basis_mat.weight", "renderModule.mlp.0.weight", "renderModule.mlp.0.bias", "renderModule.mlp.2.weight",
         "renderModule.mlp.2.bias", "renderModule.mlp.4.weight", "renderModule.mlp.4.bias"]
d = [p[k].to(dev).contiguous() for k in names]
comps = torch.rand(A, 144, device=dev) * 0.05
S, n_rays = 1000, 4096
rays_d = torch.randn(n_rays, 3, device=dev)
sidx = (torch.arange(A, device=dev) % (n_rays * S)).int()
aidx = torch.arange(A, device=dev).int()
cnt = torch.tensor([A], device=dev, dtype=torch.int32)
rgb = torch.zeros(A, 4, device=dev); feat = torch.zeros(A, 28, device=dev)
stage = ops.head_tc_stage(A, dev)
dout = torch.randn(A, 4, device=dev) * 0.1
dcomps = torch.zeros(A, 144, device=dev)
grads = [torch.zeros_like(t) for t in d]
flush = torch.empty(64 * 1024 * 1024, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps

for split in (1, 2):
    for save in (False, True):
        t = timeit(lambda: ops.head_fwd_tc(split, comps, aidx, sidx, rays_d, S, False, *d, cnt, A, 1.0, 1.0, rgb,
                                           feat if save else None, stage if save else None))
        print(f"head_fwd_tc split={split} save={save}: {t:.3f} ms")
ops.head_fwd_tc(2, comps, aidx, sidx, rays_d, S, False, *d, cnt, A, 1.0, 1.0, rgb, feat, stage)
t = timeit(lambda: ops.head_bwd_tc(dout, feat, d[0], d[1], d[3], d[5], cnt, A, 1.0, dcomps, stage, grads))
print(f"head_bwd_tc (data + wgrad): {t:.3f} ms")
