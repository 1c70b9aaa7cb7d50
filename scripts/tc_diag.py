"""Print (do not assert) the error of every tensor-core check -- one GPU round trip, full picture."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_tc as t
from joint_tensorf_b200 import ops

for k, n in [(16, 16), (32, 16), (144, 32), (160, 64), (80, 64)]:
    try:
        g = torch.Generator().manual_seed(0)
        a = torch.randn(128, k, generator=g).cuda(); b = torch.randn(n, k, generator=g).cuda()
        d = ops.tc_selftest(0, a, b, k, n); torch.cuda.synchronize()
        ref = t._bf(a) @ t._bf(b).T
        print(f"kmajor K={k} N={n}: max err {float((d-ref).abs().max()):.3e} ref max {float(ref.abs().max()):.3e} "
              f"frac rows ok {float(((d-ref).abs().max(1).values < 1e-2).float().mean()):.3f} "
              f"frac cols ok {float(((d-ref).abs().max(0).values < 1e-2).float().mean()):.3f}")
    except Exception:
        traceback.print_exc()
for ma, n in [(64, 80), (32, 144), (8, 80)]:
    try:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(128, ma, generator=g).cuda(); y = torch.randn(128, n, generator=g).cuda()
        d = ops.tc_selftest(1, x, y, 128, n, ma); torch.cuda.synchronize()
        ref = t._bf(x).T @ t._bf(y)
        print(f"mnmajor Ma={ma} N={n}: max err {float((d[:ma]-ref).abs().max()):.3e} ref max {float(ref.abs().max()):.3e} "
              f"pad rows max {float(d[ma:].abs().max()):.3e}")
    except Exception:
        traceback.print_exc()
for split in (1, 2):
    for a_count in (128, 1000):
        try:
            p, comps, rays_d, sidx, aidx, S = t._head_inputs(a_count)
            ref, feat_ref = t._head_reference(p, comps, rays_d, sidx, aidx, S, 0.8, 0.6)
            d = {k: v.cuda().contiguous() for k, v in p.items()}
            rgb = torch.zeros((a_count, 4), device="cuda"); feat = torch.zeros((a_count, 28), device="cuda")
            cnt = torch.tensor([a_count], device="cuda", dtype=torch.int32)
            ops.head_fwd_tc(split, comps.cuda(), aidx.cuda(), sidx.cuda(), rays_d.cuda(), S, False,
                            d["basis_mat.weight"], d["renderModule.mlp.0.weight"], d["renderModule.mlp.0.bias"],
                            d["renderModule.mlp.2.weight"], d["renderModule.mlp.2.bias"], d["renderModule.mlp.4.weight"],
                            d["renderModule.mlp.4.bias"], cnt, a_count, 0.8, 0.6, rgb, feat)
            torch.cuda.synchronize()
            print(f"head split={split} A={a_count}: feat err {float((feat[:, :27].cpu()-feat_ref).abs().max()):.3e} "
                  f"rgb err {float((rgb[:, :3].cpu()-ref).abs().max()):.3e}")
        except Exception:
            traceback.print_exc()

print("---- backward")
names = ["basis_mat.weight", "renderModule.mlp.0.weight", "renderModule.mlp.0.bias", "renderModule.mlp.2.weight",
         "renderModule.mlp.2.bias", "renderModule.mlp.4.weight", "renderModule.mlp.4.bias"]
from oracle import vm_oracle as vo
for a_count, bias in ((128, 0.0), (1000, 0.0), (40000, 0.0), (1000, 3.0), (40000, 3.0)):
    try:
        p, comps, rays_d, sidx, aidx, S = t._head_inputs(a_count, seed=3)
        p["renderModule.mlp.0.bias"] += bias      # bias 3.0: every unit active -> no relu-mask flips between precisions
        p["renderModule.mlp.2.bias"] += bias
        g = torch.Generator().manual_seed(7)
        dout = torch.randn(a_count, 3, generator=g) * 0.1
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        cr = comps.clone().requires_grad_(True)
        feat = cr @ pr["basis_mat.weight"].T
        dirs = rays_d[(sidx[aidx.long()] // S).long()]
        x = torch.cat([feat, dirs, vo.positional_encoding(feat, 2, 0.8), vo.positional_encoding(dirs, 2, 0.6)], -1)
        h = torch.relu(x @ pr[names[1]].T + pr[names[2]])
        h = torch.relu(h @ pr[names[3]].T + pr[names[4]])
        out = h @ pr[names[5]].T + pr[names[6]]
        (out * dout).sum().backward()
        d = {k: v.cuda().contiguous() for k, v in p.items()}
        dout4 = torch.zeros((a_count, 4), device="cuda"); dout4[:, :3] = dout.cuda()
        dcomps = torch.zeros((a_count, 144), device="cuda")
        grads = [torch.zeros_like(d[k]) for k in names]
        cnt = torch.tensor([a_count], device="cuda", dtype=torch.int32)
        rgb = torch.zeros((a_count, 4), device="cuda"); feat_d = torch.zeros((a_count, 28), device="cuda")
        stage = ops.head_tc_stage(a_count, "cuda")
        ops.head_fwd_tc(2, comps.cuda(), aidx.cuda(), sidx.cuda(), rays_d.cuda(), S, False, *[d[k] for k in names],
                        cnt, a_count, 0.8, 0.6, rgb, feat_d, stage)
        ref_rgb = torch.sigmoid(out.detach())
        print(f"   fwd(save) rgb err {float((rgb[:, :3].cpu() - ref_rgb).abs().max()):.3e}")
        ops.head_bwd_tc(dout4, feat_d, d[names[0]], d[names[1]], d[names[3]], d[names[5]], cnt, a_count, 0.8,
                        dcomps, stage, grads)
        torch.cuda.synchronize()
        rel = lambda a, b: float((a.cpu().double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        print(f"bwd A={a_count} bias+{bias}: dcomps rel {rel(dcomps, cr.grad):.3e} " + " ".join(f"{k.split('.')[-2]}.{k.split('.')[-1][0]} {rel(gk, pr[k].grad):.2e}" for k, gk in zip(names, grads)))
    except Exception:
        traceback.print_exc()
