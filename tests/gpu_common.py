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
