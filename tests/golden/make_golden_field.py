"""Golden vectors for the per-step sweeps and the field maintenance ops, from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_field.py

field_reg_adam.pt    : BAT_VMSplit.density_L1 / TV_loss_density / TV_loss_app (tensoRF.py:212-228, TVLoss
                       tensorBase.py:16-41), their autograd, and three steps of the optimiser the reference
                       builds (torch.optim.Adam(get_optparam_groups(...), betas=(0.9, 0.99)), model/tensorf.py:473-475)
                       with the per-step lr decay of tensorf.py:431-436.
field_maintenance.pt : getDenseAlpha -> updateAlphaMask (tensorBase.py:618-661) -> shrink (tensoRF.py:297-334) ->
                       upsample_volume_grid (tensoRF.py:274-295) on a field with a density blob.
"""
import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden as mg  # noqa: E402
import ref_loader  # noqa: E402

torch.set_num_threads(8)

FACTORS = [f"{pre}_{kind}.{i}" for pre in ("density", "app") for kind in ("plane", "line") for i in range(3)]


def reg_adam():
    tr, _ = ref_loader.load()
    case = dict(grid=[24, 28, 20], dens=[8] * 3, app=[12] * 3, app_dim=27, shading="MLP_Fea", hidden=64, dens_scale=1.0)
    m, aabb, kw = mg.build_reference(case, seed=3)
    gen = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k in FACTORS:                    # mixed signs and a few exact zeros (abs backward: sign(0) = 0)
                p.sub_(0.06)
                p.view(-1)[torch.randperm(p.numel(), generator=gen)[:7]] = 0.0
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    reg = tr.TVLoss()
    weights = (0.3, 10.0, 7.0)
    with contextlib.redirect_stdout(io.StringIO()):
        l1, tvd, tva = m.density_L1(), m.TV_loss_density(reg), m.TV_loss_app(reg)
    (weights[0] * l1 + weights[1] * tvd + weights[2] * tva).backward()
    reg_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    # the optimiser exactly as the reference builds it (model/tensorf.py:473-475)
    lr_index, lr_basis = 0.02, 1e-3
    optim = torch.optim.Adam(m.get_optparam_groups(lr_index, lr_basis), betas=(0.9, 0.99))
    decay = 0.1 ** (1 / 300)
    step_grads = []
    names = [k for k, _ in m.named_parameters()]
    for it in range(3):
        grads = {}
        for k, p in m.named_parameters():
            g = torch.randn(p.shape, generator=gen) * (10.0 ** -(it + 1))
            if it == 2 and k.startswith("app_line"):
                g = torch.zeros_like(g)          # a tensor with an all-zero gradient still moves (momentum)
            p.grad = g.clone()
            grads[k] = g
        step_grads.append(grads)
        optim.step()
        for group in optim.param_groups:         # tensorf.py:433-434
            group["lr"] = group["lr"] * decay
    final = {k: p.detach().clone() for k, p in m.named_parameters()}
    state = {k: {"exp_avg": optim.state[p]["exp_avg"].clone(), "exp_avg_sq": optim.state[p]["exp_avg_sq"].clone(),
                 "step": float(optim.state[p]["step"])} for k, p in m.named_parameters()}
    out = dict(case=case, aabb=aabb, field_kw=kw, state_dict=sd0, weights=weights,
               values=(float(l1), float(tvd), float(tva)), reg_grads=reg_grads, lr_index=lr_index, lr_basis=lr_basis,
               decay=decay, step_grads=step_grads, final=final, adam_state=state, param_names=names,
               final_lrs=[g["lr"] for g in optim.param_groups])
    torch.save(out, os.path.join(HERE, "field_reg_adam.pt"))
    print(f"field_reg_adam: L1={float(l1):.6f} TVd={float(tvd):.6e} TVa={float(tva):.6e} "
          f"|g|max={max(float(v.abs().max()) for v in reg_grads.values()):.3e}")


def maintenance():
    tr, _ = ref_loader.load()
    case = dict(grid=[32, 28, 36], dens=[8] * 3, app=[12] * 3, app_dim=27, shading="MLP_Fea", hidden=64, dens_scale=5.0)
    m, aabb, kw = mg.build_reference(case, seed=4)
    with torch.no_grad():                        # an off-centre density blob: planes and lines fade out
        ctr, wid = (0.15, -0.1, 0.2), 0.4
        g = case["grid"]

        def bump(axis):
            t = torch.linspace(-1, 1, g[axis])
            return torch.exp(-((t - ctr[axis]) / wid) ** 2)
        for i in range(3):
            m0, m1 = mg.ref_mat_mode()[i]
            v = mg.ref_vec_mode()[i]
            m.density_line[i].mul_(1.8 * bump(v).view(1, 1, -1, 1))
            m.density_plane[i].mul_(1.8 * bump(m1).view(1, 1, -1, 1) * bump(m0).view(1, 1, 1, -1))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    # compute_alpha reads state cached by the last forward (batBase.py:37,46-59; SURVEY B-7): run a tiny un-blurred one
    opt = ref_loader.default_opt(case["shading"], ndc=False)
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
        m.forward(opt, torch.tensor([[0.0, 0.0, 4.0]] * 4), torch.tensor([[0.0, 0.0, -1.0]] * 4), white_bg=True,
                  is_train=False, N_samples=8)
    mask_grid = (20, 24, 28)
    with contextlib.redirect_stdout(io.StringIO()):
        alpha, _ = m.getDenseAlpha(mask_grid)
        new_aabb = m.updateAlphaMask(mask_grid)
    vol = m.alphaMask.alpha_volume[0, 0].clone()
    pts = (torch.rand((512, 3), generator=torch.Generator().manual_seed(9)) * 2 - 1) * 1.6
    with contextlib.redirect_stdout(io.StringIO()):
        alpha_pts = m.compute_alpha(pts, m.stepSize)      # with the new alpha mask in place
        m.shrink(new_aabb)
    sd1 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    aabb1, grid1 = m.aabb.clone(), m.gridSize.tolist()
    step1, nsamp1 = float(m.stepSize), int(m.nSamples)
    up = [40, 36, 44]
    with contextlib.redirect_stdout(io.StringIO()):
        m.upsample_volume_grid(up)
    sd2 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out = dict(case=case, aabb=aabb, field_kw=kw, state_dict=sd0, mask_grid=mask_grid, dense_alpha=alpha.clone(),
               mask_volume=vol, new_aabb=new_aabb.clone(), alpha_thres=1e-4, pts=pts, alpha_pts=alpha_pts.clone(),
               shrunk=sd1, shrunk_aabb=aabb1, shrunk_grid=grid1, shrunk_step=step1, shrunk_nsamples=nsamp1,
               up_target=up, upsampled=sd2, up_step=float(m.stepSize))
    torch.save(out, os.path.join(HERE, "field_maintenance.pt"))
    print(f"field_maintenance: mask on {int(vol.sum())}/{vol.numel()} voxels, new_aabb={new_aabb.tolist()}, "
          f"shrunk grid {grid1}, alpha max {float(alpha.max()):.3e}")


if __name__ == "__main__":
    reg_adam()
    maintenance()
