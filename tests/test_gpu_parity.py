"""GPU parity against golden vectors of the live reference and against the CPU oracle.

Tolerances (BASELINE.json north_star): valid masks / sample indices bit-exact;
rgb, depth, opacity <= 1e-4 max-abs (fp32); gradients <= 1e-4 of the largest gradient
entry with the strict-fp32 head (fp32 atomics reorder sums, the
measured errors are ~1e-6 and are logged to gpurun_out/parity_errors.jsonl by every
test), <= 2e-2 with the tensor-core head whose backward GEMMs take bf16 operands."""
import pytest
import torch

import joint_tensorf_b200 as jt
from common import (field_from_golden, golden_names, golden_valid, load_golden, rel_err,
                    render_kwargs_from_golden, vo)
from gpu_common import (default_opt, forward_kwargs, module_from_golden, record_err, run_module_on_golden,
                        slice_parity)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ABS_TOL = 1e-4
GRAD_REL_TOL = 1e-4
# tensor-core head ("tc"): forward with split bf16 operands is fp32-class (<= 1e-4); its backward
# GEMMs are plain bf16 -> the north star's bf16 tolerance, <= 2e-2 relative
TC_GRAD_REL_TOL = 2e-2


@pytest.mark.parametrize("name", golden_names())
def test_sample_mask_and_indices_bit_exact(name):
    g = load_golden(name)
    m = module_from_golden(g, DEV)
    case = g["case"]
    ndc = case.get("ndc", False)
    o, d = g["rays_o"].to(DEV), g["rays_d"].to(DEV)
    jit = g["jitter"].to(DEV) if g["jitter"] is not None else None
    if ndc:
        pts, z, valid = m.sample_ray_ndc(o, d, is_train=case["train"], N_samples=g["n_samples"], jitter=jit)
    else:
        pts, z, valid = m.sample_ray(o, d, is_train=case["train"], N_samples=g["n_samples"], jitter=jit)
    ref_valid = golden_valid(g)
    if g["mask_volume"] is None or case["blur"] is not None:
        assert torch.equal(valid.cpu(), ref_valid), f"{int((valid.cpu() != ref_valid).sum())} mask bits differ"
    assert torch.equal(z.cpu().expand_as(g["z"]), g["z"]), "sample depths differ bitwise"
    # compacted list == nonzero(mask), including the alpha-mask cull
    geom = m._h_geom()
    aux = None
    if ndc:
        aux = (m._ndc_table(g["n_samples"], False) + (jit.reshape(-1) * (2.0 / g["n_samples"]) if jit is not None else 0)).contiguous()
    elif jit is not None:
        aux = jit.reshape(-1).contiguous()
    mask = m.alphaMask if (m.alphaMask is not None and case["blur"] is None) else None
    comp = jt.ops.march_compact(o.contiguous(), d.contiguous(), aux, ndc, g["n_samples"], geom, mask)
    v = int(comp.count.item())
    assert v == g["valid_count"]
    want = torch.nonzero(ref_valid.reshape(-1)).reshape(-1).to(torch.int32)
    assert torch.equal(comp.sidx[:v].cpu(), want)
    off = comp.ray_off.cpu()
    assert torch.equal(off[1:] - off[:-1], ref_valid.sum(-1).to(torch.int32))


@pytest.mark.parametrize("name", golden_names())
def test_forward_backward_matches_reference_golden(name):
    g = load_golden(name)
    out = run_module_on_golden(g, DEV)
    record_err("golden:" + name, head="fp32", rgb=(out["rgb"].cpu() - g["rgb"]).abs().max(),
               depth=(out["depth"].cpu() - g["depth"]).abs().max(),
               d_rays_o=rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]), d_rays_d=rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]),
               **{"g:" + k: rel_err(out["grads"][k].cpu(), ref) for k, ref in g["grads"].items()})
    assert (out["rgb"].cpu() - g["rgb"]).abs().max() <= ABS_TOL
    assert (out["acc"].cpu() - g["acc"]).abs().max() <= ABS_TOL
    assert (out["depth"].cpu() - g["depth"]).abs().max() <= ABS_TOL
    assert rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]) <= GRAD_REL_TOL
    assert rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]) <= GRAD_REL_TOL
    for k, ref in g["grads"].items():
        got = out["grads"][k].cpu()
        assert got.shape == ref.shape, k
        assert rel_err(got, ref) <= GRAD_REL_TOL, (k, rel_err(got, ref))
    for k, (s, sabs, mx) in g["grad_sums"].items():
        got = out["grads"][k].double().abs().sum().item()
        assert abs(got - sabs) <= 1e-3 * max(sabs, 1e-12), (k, got, sabs)


def test_tensor_core_head_on_reference_golden():
    """cubic_mlp golden (16/48 comps, MLP_Fea 64) through the tcgen05 head."""
    g = load_golden("cubic_mlp")
    out = run_module_on_golden(g, DEV, head="tc")
    record_err("golden:cubic_mlp", head="tc", rgb=(out["rgb"].cpu() - g["rgb"]).abs().max(),
               d_rays_o=rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]), d_rays_d=rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]),
               **{"g:" + k: rel_err(out["grads"][k].cpu(), ref) for k, ref in g["grads"].items()})
    assert (out["rgb"].cpu() - g["rgb"]).abs().max() <= ABS_TOL
    assert (out["acc"].cpu() - g["acc"]).abs().max() <= ABS_TOL
    assert rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]) <= TC_GRAD_REL_TOL
    assert rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]) <= TC_GRAD_REL_TOL
    for k, ref in g["grads"].items():
        assert rel_err(out["grads"][k].cpu(), ref) <= TC_GRAD_REL_TOL, (k, rel_err(out["grads"][k].cpu(), ref))


def test_tensor_core_sh_on_reference_golden():
    """sh_vm48 golden (16/48 comps, app_dim 27, SH shading = BASELINE configs[1]) through the fused
    gather + basis_mat + SHRender kernel and the tcgen05 SH backward."""
    g = load_golden("sh_vm48")
    out = run_module_on_golden(g, DEV, head="tc")
    record_err("golden:sh_vm48", head="tc", rgb=(out["rgb"].cpu() - g["rgb"]).abs().max(),
               d_rays_o=rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]), d_rays_d=rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]),
               **{"g:" + k: rel_err(out["grads"][k].cpu(), ref) for k, ref in g["grads"].items()})
    assert (out["rgb"].cpu() - g["rgb"]).abs().max() <= ABS_TOL
    assert (out["acc"].cpu() - g["acc"]).abs().max() <= ABS_TOL
    assert (out["depth"].cpu() - g["depth"]).abs().max() <= ABS_TOL
    assert rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]) <= TC_GRAD_REL_TOL
    assert rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]) <= TC_GRAD_REL_TOL
    for k, ref in g["grads"].items():
        assert rel_err(out["grads"][k].cpu(), ref) <= TC_GRAD_REL_TOL, (k, rel_err(out["grads"][k].cpu(), ref))


@pytest.mark.parametrize("name", ["ndc_weakview", "ndc_weakview_blur"])
def test_tensor_core_weakview_on_reference_golden(name):
    """LLFF-shaped goldens (3x16 / 3x20 comps, app_dim 20, MLP_Fea_WeakView 32, NDC rays, non-cubic grid) through
    the tcgen05 WeakView head (csrc/weakview_tc.cu): hi+lo bf16 forward (fp32 class), bf16 backward GEMMs."""
    g = load_golden(name)
    out = run_module_on_golden(g, DEV, head="tc")
    errs = dict(rgb=(out["rgb"].cpu() - g["rgb"]).abs().max(), acc=(out["acc"].cpu() - g["acc"]).abs().max(),
                depth=(out["depth"].cpu() - g["depth"]).abs().max(),
                d_rays_o=rel_err(out["d_rays_o"].cpu(), g["d_rays_o"]), d_rays_d=rel_err(out["d_rays_d"].cpu(), g["d_rays_d"]),
                **{"g:" + k: rel_err(out["grads"][k].cpu(), ref) for k, ref in g["grads"].items()})
    record_err("golden:" + name, head="tc", **errs)
    assert errs["rgb"] <= ABS_TOL and errs["acc"] <= ABS_TOL and errs["depth"] <= ABS_TOL, errs
    bad = {k: float(v) for k, v in errs.items() if k not in ("rgb", "acc", "depth") and v > TC_GRAD_REL_TOL}
    assert not bad, bad
    for k, (s, sabs, mx) in g["grad_sums"].items():
        got = out["grads"][k].double().abs().sum().item()
        assert abs(got - sabs) <= TC_GRAD_REL_TOL * max(sabs, 1e-12), (k, got, sabs)


@pytest.mark.parametrize("head", ["fp32", "tc"])
def test_midsize_sh_against_oracle(head):
    """128^3, 16/48 comps, SH shading, 300 rays (a ragged last tile): CUDA path vs the CPU oracle;
    the inference call (no staging, no autograd) must reproduce the training colours."""
    kw, run = jt.synth.config("cfg1")
    kw["shadingMode"] = "SH"
    torch.manual_seed(0)
    m = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), DEV, **kw)
    m.head_precision = head
    gtol = GRAD_REL_TOL if head == "fp32" else TC_GRAD_REL_TOL
    with torch.no_grad():
        for i in range(3):
            m.density_plane[i].mul_(4.0)
            m.density_line[i].mul_(4.0)
    n = 300
    o, d, _ = jt.synth.blender_rays(n, 10, seed=5)
    S = run["n_samples"]
    jit = torch.rand(n, 1, generator=torch.Generator().manual_seed(11))
    params = {k: v.detach().cpu().contiguous().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    field = vo.Field(aabb=m.aabb.cpu(), grid=[128] * 3, params=params, near_far=[2.0, 6.0], step_ratio=0.5,
                     density_shift=-10.0, distance_scale=25.0, weight_thres=1e-6, act="softplus", shading="SH")
    oc, dc = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    rgb_ref, depth_ref, acc_ref = vo.render(field, oc, dc, n_samples=S, white_bg=True, jitter=jit)
    w = torch.rand(n, 3, generator=torch.Generator().manual_seed(4))
    (rgb_ref * w).sum().backward()
    og, dg = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    fkw = dict(white_bg=True, is_train=True, N_samples=S, jitter=jit.to(DEV), bg_coin=False)
    rgb, depth, acc = m.forward(default_opt("SH"), og, dg, **fkw)
    (rgb * w.to(DEV)).sum().backward()
    assert (rgb.cpu() - rgb_ref).abs().max() <= ABS_TOL
    assert (acc.cpu() - acc_ref).abs().max() <= ABS_TOL
    assert (depth.cpu() - depth_ref).abs().max() <= ABS_TOL
    assert rel_err(og.grad.cpu(), oc.grad) <= gtol
    assert rel_err(dg.grad.cpu(), dc.grad) <= gtol
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), params[k].grad) <= gtol, (k, rel_err(p.grad.cpu(), params[k].grad))
    with torch.no_grad():
        rgb_inf, _, _ = m.forward(default_opt("SH"), o.to(DEV), d.to(DEV), **fkw)
    assert torch.equal(rgb_inf, rgb.detach())


@pytest.mark.parametrize("blur", [None, (0.1, 0.15)])
@pytest.mark.parametrize("head", ["fp32", "tc"])
def test_midsize_against_oracle(blur, head):
    """128^3 (cfg1 shape), 256 rays: CUDA path vs the CPU oracle on identical inputs."""
    kw, run = jt.synth.config("cfg1")
    torch.manual_seed(0)
    m = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), DEV, **kw)
    m.head_precision = head
    gtol = GRAD_REL_TOL if head == "fp32" else TC_GRAD_REL_TOL
    with torch.no_grad():
        for i in range(3):
            m.density_plane[i].mul_(4.0)
            m.density_line[i].mul_(4.0)
    o, d, _ = jt.synth.blender_rays(256, 8, seed=3)
    S = run["n_samples"]
    jit = torch.rand(256, 1, generator=torch.Generator().manual_seed(9))
    params = {k: v.detach().cpu().contiguous().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    field = vo.Field(aabb=m.aabb.cpu(), grid=[128] * 3, params=params, near_far=[2.0, 6.0], step_ratio=0.5,
                     density_shift=-10.0, distance_scale=25.0, weight_thres=1e-6, act="softplus", shading="MLP_Fea")
    okw = dict(n_samples=S, white_bg=True, jitter=jit)
    fkw = dict(white_bg=True, is_train=True, N_samples=S, jitter=jit.to(DEV), bg_coin=False)
    if blur:
        okw.update(blur_mode="uniform-gaussian", blur_density=blur[0], blur_color=blur[1], kernel_size=64)
        fkw.update(c2f_mode="uniform-gaussian", c2f_parameter_density=blur[0], c2f_parameter_color=blur[1],
                   c2f_kernel_size=64)
    oc, dc = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    rgb_ref, depth_ref, acc_ref = vo.render(field, oc, dc, **okw)
    w = torch.rand(256, 3, generator=torch.Generator().manual_seed(4))
    (rgb_ref * w).sum().backward()
    og, dg = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    rgb, depth, acc = m.forward(default_opt(), og, dg, **fkw)
    (rgb * w.to(DEV)).sum().backward()
    assert (rgb.cpu() - rgb_ref).abs().max() <= ABS_TOL
    assert (acc.cpu() - acc_ref).abs().max() <= ABS_TOL
    assert (depth.cpu() - depth_ref).abs().max() <= ABS_TOL
    assert rel_err(og.grad.cpu(), oc.grad) <= gtol
    assert rel_err(dg.grad.cpu(), dc.grad) <= gtol
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), params[k].grad) <= gtol, (k, rel_err(p.grad.cpu(), params[k].grad))
    # the no-grad call (fp16 operand tiles in the tensor-core MLP head) meets the same forward bound
    with torch.no_grad():
        rgb_inf, depth_inf, acc_inf = m.forward(default_opt(), o.to(DEV), d.to(DEV), **fkw)
    assert (rgb_inf.cpu() - rgb_ref).abs().max() <= ABS_TOL
    assert torch.equal(acc_inf, acc.detach())


@pytest.mark.parametrize("head,wl,blur", [("fp32", "cfg2", None), ("tc", "cfg2", None), ("tc", "cfg2_sh", None),
                                          ("fp32", "cfg2_sh", None), ("fp32", "cfg2", (0.09, 0.15)),
                                          ("tc", "cfg2", (0.09, 0.15)), ("tc", "cfg2_sh", (0.09, 0.15))])
def test_full_size_cfg2_properties(head, wl, blur):
    """300^3 / 4096 rays / S=1000 (the benchmark workloads: SH shading = BASELINE configs[1], the MLP_Fea head on
    the same field = configs[2], and both with the per-step blur of configs[2] at the benchmark's sigma
    parameters 0.15 colour / 0.15 x 0.6 density): size-independent invariants + a 48-ray slice against the CPU
    oracle for rgb / depth / opacity and ALL gradients (factors through the adjoint blur, basis_mat, head, rays)."""
    kw, run = jt.synth.config(wl)
    shading = kw["shadingMode"]
    torch.manual_seed(0)
    m = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), DEV, **kw)
    m.head_precision = head
    gtol = GRAD_REL_TOL if head == "fp32" else TC_GRAD_REL_TOL
    with torch.no_grad():
        for i in range(3):
            m.density_plane[i].mul_(3.0)
            m.density_line[i].mul_(3.0)
    o, d, _ = jt.synth.blender_rays(4096, 32)
    o, d = o.to(DEV), d.to(DEV)
    S = run["n_samples"]
    jit = torch.rand(4096, device=DEV)
    bkw, okw_b = {}, {}
    if blur:
        bkw = dict(c2f_mode="uniform-gaussian", c2f_parameter_density=blur[0], c2f_parameter_color=blur[1],
                   c2f_kernel_size=64)
        okw_b = dict(blur_mode="uniform-gaussian", blur_density=blur[0], blur_color=blur[1], kernel_size=64)
    # (1) compaction: sorted, unique, per-ray counts == dense mask popcount
    pts, z, valid = m.sample_ray(o, d, is_train=True, N_samples=S, jitter=jit)
    comp = jt.ops.march_compact(o, d, jit.contiguous(), False, S, m._h_geom(), None)
    v = int(comp.count.item())
    assert v == int(valid.sum())
    sidx = comp.sidx[:v]
    assert bool((sidx[1:] > sidx[:-1]).all())
    assert torch.equal(sidx.long(), torch.nonzero(valid.reshape(-1)).reshape(-1))
    assert 0.5 < v / (4096 * S) < 0.8
    del pts, z
    # (2) render: weights + background transmittance partition unity; rgb in [0,1]
    og, dg = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    rgb, depth, acc = m.forward(default_opt(shading), og, dg, white_bg=True, is_train=True, N_samples=S, jitter=jit,
                                **bkw)
    assert bool(((rgb >= 0) & (rgb <= 1)).all()) and bool(((acc >= -1e-5) & (acc <= 1 + 1e-4)).all())
    assert torch.isfinite(depth).all()
    # (3) linearity of the backward pass in the upstream gradient (a fresh forward per gradient: the strict-fp32 MLP
    # head's backward is single-use, and the tensor-core / SH paths are exercised re-entrantly)
    w1, w2 = torch.rand_like(rgb), torch.rand_like(rgb)
    params = [m.density_plane[0], m.app_line[1], m.basis_mat.weight]
    reentrant = head == "tc" or shading == "SH"

    def grads_for(w, out=None, keep=False):
        if out is None:
            out = m.forward(default_opt(shading), og, dg, white_bg=True, is_train=True, N_samples=S, jitter=jit, **bkw)[0]
        return torch.autograd.grad((out * w).sum(), params + [og], retain_graph=keep)

    if reentrant:
        ga, gb, gc = grads_for(w1, rgb, True), grads_for(w2, rgb, True), grads_for(w1 + 2 * w2, rgb)
    else:
        ga, gb, gc = grads_for(w1), grads_for(w2), grads_for(w1 + 2 * w2)
        with pytest.raises(jt._lib.JtError):             # the single-use guard
            out = m.forward(default_opt(shading), og, dg, white_bg=True, is_train=True, N_samples=S, jitter=jit, **bkw)[0]
            torch.autograd.grad(out.sum(), [og], retain_graph=True)
            torch.autograd.grad(out.sum(), [og])
    for a, b, c in zip(ga, gb, gc):
        assert rel_err(a + 2 * b, c) <= max(gtol, 1e-3)          # self-consistency (atomics reorder the sums)
    del rgb, depth, acc, ga, gb, gc
    # (4) a 48-ray slice of the same full-size field against the CPU oracle, every gradient
    sl = slice(100, 148)
    fkw = dict(opt=default_opt(shading), white_bg=True, is_train=True, N_samples=S, jitter=jit[sl], bg_coin=False, **bkw)
    okw = dict(n_samples=S, white_bg=True, jitter=jit[sl].cpu().reshape(-1, 1), **okw_b)
    field_kw = dict(aabb=m.aabb.cpu(), grid=[300] * 3, near_far=[2.0, 6.0], step_ratio=0.5, density_shift=-10.0,
                    distance_scale=25.0, weight_thres=1e-6, act="softplus", shading=shading)
    slice_parity(m, field_kw, o, d, jit, sl, fkw, okw, head, f"full:{wl}:blur={blur}", ABS_TOL, gtol, vo, rel_err)


@pytest.mark.parametrize("near,blur,head", [(-1.0, None, "fp32"), (0.4, None, "fp32"), (-1.0, (0.09, 0.15), "fp32"),
                                            (0.4, (0.09, 0.15), "fp32"), (-1.0, None, "tc"), (0.4, (0.09, 0.15), "tc")])
def test_full_size_cfg4_properties(near, blur, head):
    """BASELINE configs[3] at size: LLFF NDC rays (1008x756 views), 617x687x617 grid, 3x16 / 3x20 components,
    MLP_Fea_WeakView 32 head, relu density, S=1000, both ends of tensorf_near_plane_schedule (near 0.4 and -1);
    blur on exercises the non-cubic H/W re-interpretation of bateRF.py:21-39,68,76 (SURVEY B-3) at 617x687.
    Compaction bit-exact against the dense mask; a 64-ray slice against the CPU oracle with every gradient."""
    kw, run = jt.synth.config("cfg4")
    kw = dict(kw)
    kw["near_far"] = [near, 1.0]
    grid = list(kw["gridSize"])
    torch.manual_seed(0)
    m = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), DEV, **kw)
    m.head_precision = head
    gtol = GRAD_REL_TOL if head == "fp32" else TC_GRAD_REL_TOL
    S = run["n_samples"]
    o, d, _ = jt.synth.llff_ndc_rays(4096, 8)
    o, d = o.to(DEV), d.to(DEV)
    jit = torch.rand(1, S, generator=torch.Generator().manual_seed(3)).to(DEV)
    bkw, okw_b = {}, {}
    if blur:
        bkw = dict(c2f_mode="uniform-gaussian", c2f_parameter_density=blur[0], c2f_parameter_color=blur[1],
                   c2f_kernel_size=64)
        okw_b = dict(blur_mode="uniform-gaussian", blur_density=blur[0], blur_color=blur[1], kernel_size=64)
    # (1) compaction == nonzero(dense mask), bit-exact, sorted
    pts, z, valid = m.sample_ray_ndc(o, d, is_train=True, N_samples=S, jitter=jit)
    aux = (m._ndc_table(S, False) + jit.reshape(-1) * ((1.0 - near) / S)).contiguous()
    comp = jt.ops.march_compact(o, d, aux, True, S, m._h_geom(), None)
    v = int(comp.count.item())
    assert v == int(valid.sum()) and v > 0.5 * 4096 * S
    assert torch.equal(comp.sidx[:v].long(), torch.nonzero(valid.reshape(-1)).reshape(-1))
    off = comp.ray_off
    assert torch.equal((off[1:] - off[:-1]).long(), valid.sum(-1))
    del pts, z, valid, comp
    # (2) the full batch renders: finite, in range, appearance census in the expected class
    og, dg = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    opt = default_opt("MLP_Fea_WeakView", True)
    rgb, depth, acc = m.forward(opt, og, dg, white_bg=False, is_train=True, ndc_ray=True, N_samples=S, jitter=jit,
                                bg_coin=False, **bkw)
    a_cnt = int(jt.VMRender.last_counts[1].item())
    record_err(f"full:cfg4:near={near}:blur={blur}", head=head, V=v, A=a_cnt)
    assert bool(((rgb >= 0) & (rgb <= 1)).all()) and torch.isfinite(depth).all() and 0 < a_cnt < v
    rgb.sum().backward()
    assert torch.isfinite(og.grad).all() and torch.isfinite(m.app_plane[1].grad).all()
    del rgb, depth, acc
    # (3) 64-ray slice (8 rays of every view) against the CPU oracle, every gradient
    sl = torch.arange(0, 4096, 64, device=DEV)
    fkw = dict(opt=opt, white_bg=False, is_train=True, ndc_ray=True, N_samples=S, jitter=jit, bg_coin=False, **bkw)
    okw = dict(n_samples=S, white_bg=False, jitter=jit.cpu(), ndc=True, **okw_b)
    field_kw = dict(aabb=m.aabb.cpu(), grid=grid, near_far=[near, 1.0], step_ratio=0.3, density_shift=0.0,
                    distance_scale=25.0, weight_thres=1e-7, act="relu", shading="MLP_Fea_WeakView")
    slice_parity(m, field_kw, o, d, jit, sl, fkw, okw, head, f"full:cfg4:near={near}:blur={blur}", ABS_TOL, gtol,
                 vo, rel_err)


def test_empty_and_degenerate_batches():
    g = load_golden("cubic_blur")
    m = module_from_golden(g, DEV)
    opt = default_opt()
    # rays that miss the box entirely: no valid sample, white background
    o = torch.tensor([[10.0, 10.0, 10.0], [0.0, 0.0, 8.0]], device=DEV)
    d = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], device=DEV).requires_grad_(True)
    rgb, depth, acc = m.forward(opt, o, d, white_bg=True, is_train=False, N_samples=64)
    assert torch.equal(acc, torch.zeros_like(acc)) and torch.equal(rgb, torch.ones_like(rgb))
    rgb.sum().backward()
    assert torch.equal(d.grad, torch.zeros_like(d.grad))
    # a single ray, a ragged count (not a multiple of the warp size)
    o1, d1 = g["rays_o"][:1].to(DEV), g["rays_d"][:1].to(DEV)
    r1 = m.forward(opt, o1, d1, white_bg=True, is_train=False, N_samples=g["n_samples"])[0]
    o37, d37 = g["rays_o"][:37].to(DEV), g["rays_d"][:37].to(DEV)
    r37 = m.forward(opt, o37, d37, white_bg=True, is_train=False, N_samples=g["n_samples"])[0]
    assert torch.allclose(r1, r37[:1], atol=1e-6)


def test_empty_batches_on_the_sh_tensor_core_path():
    """No valid sample / no appearance sample at all: the tcgen05 SH kernels must not be launched on garbage
    (persistent grids read the counts from device memory) and the gradients are exact zeros."""
    g = load_golden("sh_vm48")
    m = module_from_golden(g, DEV)
    m.head_precision = "tc"
    opt = default_opt("SH")
    o = torch.tensor([[10.0, 10.0, 10.0], [0.0, 0.0, 8.0], [9.0, 0.0, 0.0]], device=DEV)
    d = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]], device=DEV).requires_grad_(True)
    rgb, depth, acc = m.forward(opt, o, d, white_bg=True, is_train=True, N_samples=g["n_samples"], bg_coin=False)
    assert torch.equal(acc, torch.zeros_like(acc)) and torch.equal(rgb, torch.ones_like(rgb))
    rgb.sum().backward()
    assert torch.equal(d.grad, torch.zeros_like(d.grad))
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert torch.equal(p.grad, torch.zeros_like(p.grad)), k
    # strongly negative density feature everywhere (softplus -> 0): valid samples exist, but no weight passes
    # the threshold -> A = 0
    with torch.no_grad():
        for i in range(3):
            m.density_plane[i].abs_().add_(0.05)
            m.density_line[i].fill_(-50.0)
    o2, d2 = g["rays_o"][:40].to(DEV), g["rays_d"][:40].to(DEV).requires_grad_(True)
    rgb, depth, acc = m.forward(opt, o2, d2, white_bg=True, is_train=True, N_samples=g["n_samples"], bg_coin=False)
    assert int(jt.VMRender.last_counts[1].item()) == 0 and int(jt.VMRender.last_counts[0].item()) > 0
    rgb.sum().backward()
    assert torch.isfinite(d2.grad).all()
    assert torch.equal(m.basis_mat.weight.grad, torch.zeros_like(m.basis_mat.weight.grad))
