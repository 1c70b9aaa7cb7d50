"""Where does the full-size cfg4 density-gradient discrepancy come from? (diagnostic, not a test)
Compares, on the 64-ray slice of tests/test_gpu_parity.py::test_full_size_cfg4_properties[0.4-None]:
  (a) per-sample dL/d sigma_feat: jt_render_bwd (captured from VMRender.backward) vs oracle fp32 vs oracle f64
  (b) the density scatter alone: jt_vm_scatter_rays fed with the ORACLE's fp32 dL/d sigma_feat vs the oracle's plane
      gradients restricted to the density path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import joint_tensorf_b200 as jt
from joint_tensorf_b200 import ops, render
from joint_tensorf_b200.options import default_opt
from oracle import vm_oracle as vo

DEV = "cuda:0"
near = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
kw, run = jt.synth.config("cfg4"); kw = dict(kw); kw["near_far"] = [near, 1.0]
grid = list(kw["gridSize"])
torch.manual_seed(0)
m = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), DEV, **kw)
m.head_precision = "fp32"
S = run["n_samples"]
o, d, _ = jt.synth.llff_ndc_rays(4096, 8)
sl = torch.arange(0, 4096, 64)
o, d = o[sl].contiguous(), d[sl].contiguous()
jit = torch.rand(1, S, generator=torch.Generator().manual_seed(3))
n = o.shape[0]
g = torch.Generator().manual_seed(4)
w_rgb, w_acc = torch.rand(n, 3, generator=g), torch.rand(n, generator=g)
sd = {k: v.detach().cpu().contiguous().clone() for k, v in m.state_dict().items()}
field_kw = dict(aabb=m.aabb.cpu(), grid=grid, near_far=[near, 1.0], step_ratio=0.3, density_shift=0.0,
                distance_scale=25.0, weight_thres=1e-7, act="relu", shading="MLP_Fea_WeakView")

def oracle(dtype):
    params = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    f = vo.Field(params=params, **field_kw)
    rgb, depth, acc, det = vo.render(f, o.clone(), d.clone(), n_samples=S, white_bg=False, jitter=jit, ndc=True,
                                     exact=(dtype == torch.float64), detail=True)
    det["sigma_feat"].retain_grad()
    ((rgb * w_rgb).sum() + (acc * w_acc).sum()).backward()
    return det["sigma_feat"].grad.detach(), {k: p.grad for k, p in params.items() if p.grad is not None}, det

gs32, g32, det32 = oracle(torch.float32)
gs64, g64, det64 = oracle(torch.float64)
print("app samples: f32", int(det32["app_mask"].sum()), "f64", int(det64["app_mask"].sum()),
      "differ", int((det32["app_mask"] != det64["app_mask"]).sum()))

import joint_tensorf_b200._lib as L
cap = {}
render.VMRender.debug_capture = cap
og, dg = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
rgb, depth, acc = m.forward(default_opt("MLP_Fea_WeakView", True), og, dg, white_bg=False, is_train=True, ndc_ray=True,
                            N_samples=S, jitter=jit.to(DEV), bg_coin=False)
V = int(jt.VMRender.last_counts[0].item())
((rgb * w_rgb.to(DEV)).sum() + (acc * w_acc.to(DEV)).sum()).backward()
torch.cuda.synchronize()
render.VMRender.debug_capture = None
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
print("V", V, "oracle valid", gs32.numel())
mine_p = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
for k in ("density_plane.0", "density_line.1"):
    print(k, "mine vs ref32", rel(mine_p[k], g32[k]), "mine vs f64", rel(mine_p[k], g64[k]), "ref32 vs f64", rel(g32[k], g64[k]))

# (b) scatter alone with the oracle's fp32 dL/dsigma_feat
fs = ops.FactorSet([p.detach() for p in m.density_plane], [p.detach() for p in m.density_line])
aux = (m._ndc_table(S, False) + jit.reshape(-1).to(DEV) * ((1.0 - near) / S)).contiguous()
comp = ops.march_compact(og.detach().contiguous(), dg.detach().contiguous(), aux, True, S, m._h_geom(), None)
assert int(comp.count.item()) == gs32.numel()
for name, gs in (("ref32 dsig", gs32), ("f64 dsig (rounded)", gs64.float())):
    gp, gl = fs.zero_grads()
    d_o = torch.zeros((n, 3), device=DEV); d_d = torch.zeros((n, 3), device=DEV)
    gin = torch.zeros((comp.cap,), device=DEV); gin[:V] = gs.to(DEV)
    ops.vm_scatter_rays(0, fs, gp, gl, comp.samp, None, comp.sidx, comp.count, comp.cap, gin, S,
                        L.floats(m._h_inv.tolist()), d_o, d_d)
    torch.cuda.synchronize()
    gpn, gln = ops.FactorSet.grads_as_nchw(gp, gl)
    print(f"scatter alone fed with {name}: plane0 vs ref32 {rel(gpn[0].cpu(), g32['density_plane.0']):.2e} vs f64 "
          f"{rel(gpn[0].cpu(), g64['density_plane.0']):.2e}; line1 vs ref32 {rel(gln[1].cpu(), g32['density_line.1']):.2e} "
          f"vs f64 {rel(gln[1].cpu(), g64['density_line.1']):.2e}")

# (a) per-sample dL/d sigma_feat of the composite backward (jt_render_bwd) against the oracle's
mine_s = cap["dsig"][:V].cpu()
print("per-sample dL/dsigma_feat: mine vs ref32", rel(mine_s, gs32), " mine vs f64", rel(mine_s, gs64), " ref32 vs f64", rel(gs32, gs64))
# where do they differ most, relative to the ray-local scale?
dif = (mine_s.double() - gs64).abs()
j = int(dif.argmax())
print("worst sample", j, "mine", float(mine_s[j]), "ref32", float(gs32[j]), "f64", float(gs64[j]), "max|g|", float(gs64.abs().max()))

# ---- structure of the per-sample differences: which rays / which depth along the ray?
sid = comp.sidx[:V].cpu().long()
ray, kk = sid // S, sid % S
d_m, d_r = (mine_s.double() - gs64), (gs32.double() - gs64)
scale = float(gs64.abs().max())
for name, dd_ in (("mine-f64", d_m), ("ref32-f64", d_r), ("mine-ref32", mine_s.double() - gs32.double())):
    per_ray_abs = torch.zeros(n, dtype=torch.float64).index_add_(0, ray, dd_.abs())
    per_ray_sgn = torch.zeros(n, dtype=torch.float64).index_add_(0, ray, dd_)
    coh = float((per_ray_sgn.abs() / per_ray_abs.clamp_min(1e-300)).mean())
    print(f"{name}: max {float(dd_.abs().max())/scale:.2e} rms {float(dd_.pow(2).mean().sqrt())/scale:.2e} "
          f"mean sign-coherence per ray {coh:.3f}; worst ray {int(per_ray_abs.argmax())}")
top = torch.topk(d_m.abs(), 8).indices
for t_ in top.tolist():
    print(f"  sample {t_}: ray {int(ray[t_])} k {int(kk[t_])} mine {float(mine_s[t_]):.6e} ref32 {float(gs32[t_]):.6e} f64 {float(gs64[t_]):.6e}")
# the same for the values feeding the backward: weights and transmittance
w32 = det32["weight"][det32["valid"]]; w64 = det64["weight"][det64["valid"]]
print("oracle weights ref32 vs f64 (rel to max):", rel(w32, w64))

mw = cap["weight"][:V].cpu()
print("my weights vs ref32:", rel(mw, w32), " vs f64:", rel(mw, w64))
msf = cap["sigfeat"][:V].cpu()
print("my sigma_feat vs ref32:", rel(msf, det32["sigma_feat"].detach()), " vs f64:", rel(msf, det64["sigma_feat"].detach()),
      " ref32 vs f64:", rel(det32["sigma_feat"].detach(), det64["sigma_feat"].detach()))
A_ = int(cap["a_count"].item())
mrgb = cap["rgb"][:A_, :3].cpu()
r32 = det32["rgb"][det32["app_mask"]]; r64 = det64["rgb"][det64["app_mask"]]
print("my rgb vs ref32:", float((mrgb - r32).abs().max()), " vs f64:", float((mrgb.double() - r64).abs().max()),
      " ref32 vs f64:", float((r32.double() - r64).abs().max()))
