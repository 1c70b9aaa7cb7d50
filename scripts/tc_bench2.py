"""Time the forward tensor-core kernels in isolation (CUDA events, L2 flushed between runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import joint_tensorf_b200 as jt
from joint_tensorf_b200 import ops
from joint_tensorf_b200.ops import FactorSet

dev = "cuda"
kw, run = jt.synth.config("cfg2")
kw = dict(kw)
torch.manual_seed(0)
m = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), dev, **kw)
o, d, _ = jt.synth.blender_rays(4096, 32, seed=1)
o, d = o.to(dev), d.to(dev)
S = run["n_samples"]
jit = torch.rand(4096, device=dev)
comp = ops.march_compact(o, d, jit, False, S, m._h_geom(), None)
V = int(comp.count.item())
A = V
aidx = torch.arange(V, device=dev, dtype=torch.int32)
cnt = torch.tensor([A], device=dev, dtype=torch.int32)
afs = FactorSet(list(m.app_plane), list(m.app_line))
head = [t.detach().contiguous() for t in m.renderModule.head_params()]
wb = m.basis_mat.weight.detach().contiguous()
featdir = torch.zeros(A + 128, 32, device=dev)
rgb = torch.zeros(A + 128, 4, device=dev)
stage = ops.head_tc_stage(A + 128, dev)
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

print("A =", A)
for split in (1, 2):
    for save in (False, True):
        st = stage if save else None
        t = timeit(lambda: ops.app_basis_fwd_tc(split, afs, comp.samp, aidx, comp.sidx, d, S, False, wb, cnt, A, featdir, st))
        t2 = timeit(lambda: ops.head_mlp_fwd_tc(split, featdir, *head, cnt, A, 1.0, 1.0, rgb, st))
        print(f"split={split} save={save}: app_basis {t:.3f} ms   head_mlp {t2:.3f} ms")
t3 = timeit(lambda: ops.head_mlp_fwd_tc(3, featdir, *head, cnt, A, 1.0, 1.0, rgb, None))
print(f"split=3 (fp16 operand tiles, inference): head_mlp {t3:.3f} ms")
comps = torch.zeros(A, 144, device=dev)
print("vm_app_fwd (SIMT gather only):", timeit(lambda: ops.vm_gather_fwd(1, afs, comp.samp, aidx, cnt, A, comps)))
