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


def run_module_on_golden(g, device="cuda:0", head="fp32", storage="fp32", round_factors=False, bf16_bwd_taps=False):
    m = module_from_golden(g, device)
    m.head_precision = head
    m.factor_storage = storage
    m.bf16_backward_taps = bf16_bwd_taps
    if round_factors:          # fp32 factors holding bf16-representable values
        with torch.no_grad():
            for plist in (m.density_plane, m.density_line, m.app_plane, m.app_line):
                for p in plist:
                    p.copy_(p.to(torch.bfloat16).float())
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


def slice_parity(m, field_kw, o, d, jit, sl, fkw, okw, head, tag, abs_tol, grad_tol, vo, rel_err, seed=4, exact=None):
    """Render the ray slice `sl` of a full-size field on the GPU module `m` and with the CPU oracle on a copy of its
    parameters; compare rgb / depth / opacity and EVERY gradient (12 factors, basis_mat, head, rays_o, rays_d).

    A gradient passes when it is within `grad_tol` of the oracle's fp32 result (the reference's arithmetic). Some
    full-size gradients are ill-conditioned in fp32 -- the density gradient of a nearly opaque field is the
    difference of two transmittance sums that cancel to ~1e-4 of their size, and alpha = 1 - exp(-x) for x ~ 1e-4
    amplifies one ulp of exp() to 3e-4 -- so two correct fp32 implementations with different summation order differ
    by more than 1e-4 there. For those (`exact` defaults to True with the fp32 head) the oracle is also run in
    float64 on the same fp32-placed samples (vo.render(exact=True)) and the gradient must be no further from that
    exact result than twice the reference's own fp32 rounding error: err(ours, f64) <= max(tol, 2 err(ref32, f64)).
    All three numbers are logged per tensor."""
    dev = o.device
    exact = (head == "fp32") if exact is None else exact
    sd = {k: v.detach().cpu().contiguous().clone() for k, v in m.state_dict().items() if not k.startswith("alphaMask")}
    n = o[sl].shape[0]
    g = torch.Generator().manual_seed(seed)
    w_rgb, w_acc = torch.rand(n, 3, generator=g), torch.rand(n, generator=g)

    def run_oracle(dtype):
        params = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
        field = vo.Field(params=params, **field_kw)
        os_, ds_ = o[sl].cpu().clone().requires_grad_(True), d[sl].cpu().clone().requires_grad_(True)
        rgb_r, depth_r, acc_r = vo.render(field, os_, ds_, exact=(dtype == torch.float64), **okw)
        ((rgb_r * w_rgb).sum() + (acc_r * w_acc).sum()).backward()
        grads = {"d_rays_o": os_.grad, "d_rays_d": ds_.grad}
        grads.update({"g:" + k: p.grad for k, p in params.items() if p.grad is not None})
        return rgb_r.detach(), depth_r.detach(), acc_r.detach(), grads

    rgb_ref, depth_ref, acc_ref, g32 = run_oracle(torch.float32)
    g64 = run_oracle(torch.float64)[3] if exact else None
    for p in m.parameters():
        p.grad = None
    og, dg = o[sl].clone().requires_grad_(True), d[sl].clone().requires_grad_(True)
    m.head_precision = head
    fkw = dict(fkw)
    rgb, depth, acc = m.forward(fkw.pop("opt"), og, dg, **fkw)
    ((rgb * w_rgb.to(dev)).sum() + (acc * w_acc.to(dev)).sum()).backward()
    mine = {"d_rays_o": og.grad.cpu(), "d_rays_d": dg.grad.cpu()}
    mine.update({"g:" + k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None})
    errs = dict(rgb=(rgb.detach().cpu() - rgb_ref).abs().max(), acc=(acc.detach().cpu() - acc_ref).abs().max(),
                depth=(depth.cpu() - depth_ref).abs().max())
    bad = {}
    for k, ref in g32.items():
        assert k in mine, k
        e_ref = rel_err(mine[k], ref)
        errs[k] = e_ref
        ok = e_ref <= grad_tol
        if exact:
            e_exact, e_noise = rel_err(mine[k], g64[k]), rel_err(ref, g64[k])
            errs[k + "|vs_f64"], errs[k + "|ref32_vs_f64"] = e_exact, e_noise
            ok = ok or e_exact <= max(grad_tol, 2.0 * e_noise)
        if not ok:
            bad[k] = {kk: float(v) for kk, v in errs.items() if kk.startswith(k)}
    record_err(tag, head=head, **errs)
    assert errs["rgb"] <= abs_tol and errs["acc"] <= abs_tol, errs
    assert errs["depth"] <= 2 * abs_tol, errs
    assert not bad, bad
    return errs
