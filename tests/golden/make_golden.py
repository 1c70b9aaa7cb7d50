"""Generate golden input/output vectors from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case below it constructs the unmodified reference `BAT_VMSplit`
(model/tensorf_repr/bateRF.py:7) on CPU, runs `forward` (batBase.py:44) and the
backward of a sum-reduced loss on seeded inputs, and stores inputs + outputs as
`tests/golden/<case>.pt`. Random draws the reference makes internally
(`torch.rand_like` for the stratified jitter, `torch.rand` for the background
coin flip) are replaced by recorded tensors so that the same numbers can be fed
to the oracle and to the CUDA path.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402
import joint_tensorf_b200.synth as synth  # noqa: E402

torch.set_num_threads(8)

CASES = {
    # name: (field kwargs overrides, run kwargs)
    "cubic_mlp": dict(grid=[32, 32, 32], dens=[16] * 3, app=[48] * 3, app_dim=27, shading="MLP_Fea", hidden=64,
                      n_rays=96, train=True, blur=None, dens_scale=3.5),
    "cubic_blur": dict(grid=[24, 24, 24], dens=[8] * 3, app=[12] * 3, app_dim=27, shading="MLP_Fea", hidden=64,
                       n_rays=64, train=True, blur=(0.09, 0.15)),
    "noncubic_blur": dict(grid=[24, 28, 20], dens=[8] * 3, app=[12] * 3, app_dim=27, shading="MLP_Fea", hidden=64,
                          n_rays=64, train=True, blur=(0.2, 0.1)),
    "alpha_mask": dict(grid=[28, 28, 28], dens=[8] * 3, app=[12] * 3, app_dim=27, shading="MLP_Fea", hidden=64,
                       n_rays=64, train=False, blur=None, mask=True),
    "sh": dict(grid=[24, 24, 24], dens=[8] * 3, app=[12] * 3, app_dim=27, shading="SH", hidden=0,
               n_rays=64, train=True, blur=None),
    # BASELINE configs[1] component counts (16 / 48 per plane, app_dim 27, SH shading): the shape the
    # tensor-core SH path (jt_app_basis_sh_fwd_tc / jt_sh_bwd_tc) is specialised for
    "sh_vm48": dict(grid=[32, 32, 32], dens=[16] * 3, app=[48] * 3, app_dim=27, shading="SH", hidden=0,
                    n_rays=96, train=True, blur=None, dens_scale=3.5),
    "ndc_weakview": dict(grid=[24, 28, 24], dens=[16] * 3, app=[20] * 3, app_dim=20, shading="MLP_Fea_WeakView",
                         hidden=32, n_rays=64, train=True, blur=None, ndc=True, dens_scale=0.12),
    "ndc_weakview_blur": dict(grid=[24, 28, 24], dens=[16] * 3, app=[20] * 3, app_dim=20, shading="MLP_Fea_WeakView",
                              hidden=32, n_rays=64, train=True, blur=(0.12, 0.08), ndc=True, dens_scale=0.12),
}

FULL_GRADS = ("density_plane.0", "density_line.2", "app_plane.1", "app_line.0", "basis_mat.weight")


@contextlib.contextmanager
def patched_rng(jitter):
    real_rl, real_r = torch.rand_like, torch.rand

    def rand_like(x, *a, **k):
        assert tuple(x.shape) == tuple(jitter.shape), (x.shape, jitter.shape)
        return jitter.clone().to(x.dtype)

    def rand(*a, **k):       # background coin flip batBase.py:154 -> "not white"
        return torch.full((1,), 0.9)

    torch.rand_like, torch.rand = rand_like, rand
    try:
        yield
    finally:
        torch.rand_like, torch.rand = real_rl, real_r


def ref_mat_mode():
    return [[0, 1], [0, 2], [1, 2]]     # tensorBase.py:405


def ref_vec_mode():
    return [2, 1, 0]                    # tensorBase.py:406


def build_reference(case, seed=0):
    tr, _ = ref_loader.load()
    ndc = case.get("ndc", False)
    if ndc:
        aabb = torch.tensor([[-1.5, -1.67, -2.0], [1.5, 1.67, 1.0]])
        kw = dict(near_far=[-1.0, 1.0], density_shift=0.0, step_ratio=0.3, fea2denseAct="relu",
                  volume_init_scale=0.05, volume_init_bias=0.2, rayMarch_weight_thres=1e-7)
    else:
        aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
        kw = dict(near_far=[2.0, 6.0], density_shift=-10, step_ratio=0.5, fea2denseAct="softplus",
                  volume_init_scale=0.1, volume_init_bias=0.0, rayMarch_weight_thres=1e-6)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = tr.BAT_VMSplit(aabb, list(case["grid"]), "cpu", density_n_comp=list(case["dens"]),
                           appearance_n_comp=list(case["app"]), app_dim=case["app_dim"],
                           shadingMode=case["shading"], alphaMask_thres=1e-4, distance_scale=25.0,
                           pos_pe=2, view_pe=2, fea_pe=2, featureC=max(case["hidden"], 1), dtype=torch.float32, **kw)
    if case["shading"] == "SH":      # 5-arg call of a 3-arg function
        from model.tensorf_repr import tensorBase as _tb
        m.renderModule = lambda p, v, f, *_: _tb.SHRender(p, v, f)
    # make the field non-trivial: random init keeps softplus(f-10) ~ 1e-4, so all
    # weights are tiny; add a density blob so that transmittance actually decays.
    with torch.no_grad():
        if True:
            for i in range(3):
                m.density_plane[i].mul_(case.get("dens_scale", 6.0))
                m.density_line[i].mul_(case.get("dens_scale", 6.0))
    return m, aabb, kw


def run_case(name, case):
    tr, _ = ref_loader.load()
    ndc = case.get("ndc", False)
    m, aabb, kw = build_reference(case)
    g = case["grid"]
    step_ratio = kw["step_ratio"]
    n_samples = min(1000, int(np.linalg.norm(g) / step_ratio))
    n = case["n_rays"]
    if ndc:
        o, d, _ = synth.llff_ndc_rays(n_rays=n, n_views=8, seed=11)
    else:
        o, d, _ = synth.blender_rays(n_rays=n, n_views=8, seed=11)
    gen = torch.Generator().manual_seed(5)
    if case["train"]:
        jitter = torch.rand((1, n_samples), generator=gen) if ndc else torch.rand((n, 1), generator=gen)
    else:
        jitter = None
    w_rgb = torch.rand((n, 3), generator=gen)
    w_acc = torch.rand((n,), generator=gen)

    if case.get("mask"):
        # binary occupancy volume: a ball + a slab, on a grid different from gridSize
        md, mh, mw = 20, 22, 24
        zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, md), torch.linspace(-1, 1, mh),
                                    torch.linspace(-1, 1, mw), indexing="ij")
        vol = (((xx - 0.1) ** 2 + yy ** 2 + (zz + 0.2) ** 2) < 0.45).float()
        vol[:, :, :3] = 1.0
        m.alphaMask = tr.AlphaGridMask("cpu", aabb, vol)

    opt = ref_loader.default_opt(case["shading"], ndc=ndc)
    o = o.clone().requires_grad_(True)
    d = d.clone().requires_grad_(True)
    blur = case["blur"]
    fw = dict(white_bg=not ndc, is_train=case["train"], ndc_ray=ndc, N_samples=n_samples)
    if blur is not None:
        fw.update(c2f_parameter_density=blur[0], c2f_parameter_color=blur[1], c2f_mode="uniform-gaussian",
                  c2f_kernel_size=64)
    with patched_rng(jitter if jitter is not None else torch.zeros(1)):
        rgb, depth, acc = m.forward(opt, o, d, **fw)
        # intermediate: the valid mask (re-run the sampler with the same jitter)
        if ndc:
            _, z, valid = m.sample_ray_ndc(o, d, is_train=case["train"], N_samples=n_samples)
        else:
            _, z, valid = m.sample_ray(o, d, is_train=case["train"], N_samples=n_samples)
        if m.alphaMask is not None and blur is None:
            pts = o[:, None, :] + d[:, None, :] * z[..., None]
            keep = m.alphaMask.sample_alpha(pts[valid]) > 0
            bad = ~valid
            bad[valid] |= ~keep
            valid = ~bad
    loss = (rgb * w_rgb).sum() + (acc * w_acc).sum()
    loss.backward()

    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    grads = {}
    sums = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        sums[k] = (float(p.grad.double().sum()), float(p.grad.double().abs().sum()), float(p.grad.abs().max()))
        if k in FULL_GRADS or k.startswith("renderModule"):
            grads[k] = p.grad.detach().clone()
    out = dict(
        name=name, case=dict(case), aabb=aabb, field_kw=kw, n_samples=n_samples,
        rays_o=o.detach().clone(), rays_d=d.detach().clone(), jitter=jitter, w_rgb=w_rgb, w_acc=w_acc,
        state_dict=sd,
        mask_volume=(m.alphaMask.alpha_volume[0, 0].clone() if m.alphaMask is not None else None),
        valid_packed=torch.from_numpy(np.packbits(valid.numpy().reshape(-1))), valid_count=int(valid.sum()),
        z=z.detach().clone(), rgb=rgb.detach().clone(), depth=depth.detach().clone(), acc=acc.detach().clone(),
        d_rays_o=o.grad.clone(), d_rays_d=d.grad.clone(), grads=grads, grad_sums=sums,
    )
    torch.save(out, os.path.join(HERE, f"{name}.pt"))
    print(f"{name}: S={n_samples} valid={int(valid.sum())}/{valid.numel()} acc_mean={float(acc.mean()):.4f} "
          f"rgb_mean={float(rgb.mean()):.4f} |d_o|max={float(o.grad.abs().max()):.3e}")
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        run_case(nm, CASES[nm])
