"""Build a B200_VMSplit from a golden fixture and run forward+backward on the GPU."""
import torch

import joint_tensorf_b200 as jt
from common import render_kwargs_from_golden  # noqa: F401


from joint_tensorf_b200.options import default_opt  # noqa: F401,E402


def module_from_golden(g, device="cuda:0"):
    case, kw = g["case"], g["field_kw"]
    m = jt.B200_VMSplit(g["aabb"].clone(), list(case["grid"]), device, density_n_comp=list(case["dens"]),
                        appearance_n_comp=list(case["app"]), app_dim=case["app_dim"], shadingMode=case["shading"],
                        alphaMask_thres=1e-4, distance_scale=25.0, pos_pe=2, view_pe=2, fea_pe=2,
                        featureC=max(case["hidden"], 1), dtype=torch.float32, **kw)
    missing, unexpected = m.load_state_dict(g["state_dict"], strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if not k.startswith("alphaMask")], missing
    if g.get("mask_volume") is not None:
        m.alphaMask = jt.AlphaGridMask(device, g["aabb"].clone().to(device), g["mask_volume"].to(device))
    return m


def forward_kwargs(g, device):
    case = g["case"]
    ndc = case.get("ndc", False)
    kw = dict(white_bg=not ndc, is_train=case["train"], ndc_ray=ndc, N_samples=g["n_samples"], bg_coin=False)
    if g["jitter"] is not None:
        kw["jitter"] = g["jitter"].to(device)
    if case["blur"] is not None:
        kw.update(c2f_parameter_density=case["blur"][0], c2f_parameter_color=case["blur"][1],
                  c2f_mode="uniform-gaussian", c2f_kernel_size=64)
    return kw


def run_module_on_golden(g, device="cuda:0", head="fp32"):
    m = module_from_golden(g, device)
    m.head_precision = head
    o = g["rays_o"].to(device).requires_grad_(True)
    d = g["rays_d"].to(device).requires_grad_(True)
    opt = default_opt(g["case"]["shading"], g["case"].get("ndc", False))
    rgb, depth, acc = m.forward(opt, o, d, **forward_kwargs(g, device))
    loss = (rgb * g["w_rgb"].to(device)).sum() + (acc * g["w_acc"].to(device)).sum()
    loss.backward()
    grads = {k: p.grad.detach() for k, p in m.named_parameters() if p.grad is not None}
    return dict(module=m, rgb=rgb.detach(), depth=depth.detach(), acc=acc.detach(), d_rays_o=o.grad, d_rays_d=d.grad,
                grads=grads)


# ------------------------------------------------------------------ measured-error log + full-size slice parity
import json
import os

_ERRLOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_errors.jsonl")


def record_err(test, **vals):
    """Append the MEASURED errors of a parity check to gpurun_out/parity_errors.jsonl (copied to profiles/ per
    round), so that the achieved numbers are on record next to the tolerance the assert uses."""
    try:
        os.makedirs(os.path.dirname(_ERRLOG), exist_ok=True)
        with open(_ERRLOG, "a") as f:
            f.write(json.dumps(dict(test=test, **{k: (float(v) if not isinstance(v, (str, int)) else v)
                                                  for k, v in vals.items()})) + "\n")
    except OSError:
        pass


def slice_parity(m, field_kw, o, d, jit, sl, fkw, okw, head, tag, abs_tol, grad_tol, vo, rel_err, seed=4):
    """Render the ray slice `sl` of a full-size field on the GPU module `m` and with the CPU oracle on a copy of its
    parameters; compare rgb / depth / opacity and EVERY gradient (12 factors, basis_mat, head, rays_o, rays_d)."""
    dev = o.device
    params = {k: v.detach().cpu().contiguous().clone().requires_grad_(True) for k, v in m.state_dict().items()
              if not k.startswith("alphaMask")}
    field = vo.Field(params=params, **field_kw)
    os_, ds_ = o[sl].cpu().clone().requires_grad_(True), d[sl].cpu().clone().requires_grad_(True)
    n = os_.shape[0]
    g = torch.Generator().manual_seed(seed)
    w_rgb, w_acc = torch.rand(n, 3, generator=g), torch.rand(n, generator=g)
    rgb_ref, depth_ref, acc_ref = vo.render(field, os_, ds_, **okw)
    ((rgb_ref * w_rgb).sum() + (acc_ref * w_acc).sum()).backward()
    for p in m.parameters():
        p.grad = None
    og, dg = o[sl].clone().requires_grad_(True), d[sl].clone().requires_grad_(True)
    m.head_precision = head
    rgb, depth, acc = m.forward(fkw.pop("opt"), og, dg, **fkw)
    ((rgb * w_rgb.to(dev)).sum() + (acc * w_acc.to(dev)).sum()).backward()
    errs = dict(rgb=(rgb.detach().cpu() - rgb_ref).abs().max(), acc=(acc.detach().cpu() - acc_ref).abs().max(),
                depth=(depth.cpu() - depth_ref).abs().max(), d_rays_o=rel_err(og.grad.cpu(), os_.grad),
                d_rays_d=rel_err(dg.grad.cpu(), ds_.grad))
    for k, p in m.named_parameters():
        ref = params[k].grad
        if ref is None:
            continue
        assert p.grad is not None, k
        errs["g:" + k] = rel_err(p.grad.cpu(), ref)
    record_err(tag, head=head, **errs)
    assert errs["rgb"] <= abs_tol and errs["acc"] <= abs_tol, errs
    assert errs["depth"] <= 2 * abs_tol, errs
    bad = {k: v for k, v in errs.items() if (k.startswith("g:") or k.startswith("d_rays")) and not v <= grad_tol}
    assert not bad, bad
    return errs
