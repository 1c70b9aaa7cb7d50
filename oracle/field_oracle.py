"""CPU oracle for the per-step factor sweeps and the field maintenance ops (TEST INFRASTRUCTURE).

Checker only (tests/, smoke(), bench.py CPU legs); the product never imports it. Restates, as plain
functions over torch CPU tensors, what the reference does around the render path every step
 and between steps (8f-3). Each function cites the reference file:line.

Parity pin: `tests/golden/make_golden_field.py` runs the LIVE reference (`BAT_VMSplit.density_L1`,
`TV_loss_*` with `TVLoss`, `torch.optim.Adam` exactly as model/tensorf.py:474-475 builds it,
`updateAlphaMask`, `shrink`, `upsample_volume_grid`) on seeded fields and stores the results in
`tests/golden/field_*.pt`; `tests/test_oracle_golden.py` checks this file against them.

Third-party arithmetic: PyTorch (`torch.optim.Adam` single-tensor algorithm, torch/optim/adam.py;
`F.interpolate(mode="bilinear", align_corners=True)`, `F.max_pool3d`).
"""
import math

import torch
import torch.nn.functional as F

MAT_MODE = ((0, 1), (0, 2), (1, 2))  # tensorBase.py:405
VEC_MODE = (2, 1, 0)                 # tensorBase.py:406


# --------------------------------------------------------------------------- regularisers (8f-2)
def tv_loss(x, weight=1.0):
    """TVLoss.forward (tensorBase.py:21-38) on a [B,C,H,W] tensor."""
    b, c, h, w = x.shape
    count_h = c * (h - 1) * w
    count_w = c * h * (w - 1)
    total = 0
    if count_h > 0:
        total = total + torch.pow(x[:, :, 1:, :] - x[:, :, :h - 1, :], 2).sum() / count_h
    if count_w > 0:
        total = total + torch.pow(x[:, :, :, 1:] - x[:, :, :, :w - 1], 2).sum() / count_w
    return weight * 2 * total / b


def density_l1(params):
    """TensorVMSplit.density_L1 (tensoRF.py:212-216)."""
    total = 0
    for i in range(3):
        total = total + torch.mean(torch.abs(params[f"density_plane.{i}"])) + torch.mean(torch.abs(params[f"density_line.{i}"]))
    return total


def tv_loss_density(params):
    """TensorVMSplit.TV_loss_density (tensoRF.py:218-222)."""
    return sum(tv_loss(params[f"density_plane.{i}"]) * 1e-2 for i in range(3))


def tv_loss_app(params):
    """TensorVMSplit.TV_loss_app (tensoRF.py:224-228)."""
    return sum(tv_loss(params[f"app_plane.{i}"]) * 1e-2 for i in range(3))


# --------------------------------------------------------------------------- optimiser (8f-2)
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.99, eps=1e-8):
    """One torch.optim.Adam update (torch/optim/adam.py `_single_tensor_adam`, weight_decay 0, amsgrad off),
    the optimiser model/tensorf.py:474-475 constructs with betas (0.9, 0.99). In place on p, m, v;
    `step` is the count after the increment."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


def lr_decay_factor(target_ratio, duration):
    """model/tensorf.py:285-287; applied to every param_group each step (tensorf.py:431-436)."""
    return target_ratio ** (1 / duration)


# --------------------------------------------------------------------------- maintenance (8f-3)
def upsample_factors(params, res_target):
    """TensorVMSplit.up_sampling_VM / upsample_volume_grid (tensoRF.py:274-295)."""
    out = dict(params)
    for pre in ("app", "density"):
        for i in range(3):
            m0, m1 = MAT_MODE[i]
            out[f"{pre}_plane.{i}"] = F.interpolate(params[f"{pre}_plane.{i}"], size=(res_target[m1], res_target[m0]),
                                                    mode="bilinear", align_corners=True)
            out[f"{pre}_line.{i}"] = F.interpolate(params[f"{pre}_line.{i}"], size=(res_target[VEC_MODE[i]], 1),
                                                   mode="bilinear", align_corners=True)
    return out


def dense_grid_points(aabb, grid):
    """getDenseAlpha's sample positions (tensorBase.py:621-626): [gx,gy,gz,3]."""
    samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, grid[0]), torch.linspace(0, 1, grid[1]),
                                         torch.linspace(0, 1, grid[2]), indexing="ij"), -1)
    return aabb[0] * (1 - samples) + aabb[1] * samples


def alpha_mask_from_dense(alpha, dense_xyz, thres):
    """updateAlphaMask after getDenseAlpha (tensorBase.py:639-657): alpha [gx,gy,gz] -> (volume [gz,gy,gx] of
    {0,1}, new_aabb [2,3])."""
    grid = list(alpha.shape)
    dense_xyz = dense_xyz.transpose(0, 2).contiguous()
    a = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
    a = F.max_pool3d(a, kernel_size=5, padding=2, stride=1).view(grid[::-1])
    a = (a >= thres).to(a.dtype)
    valid = dense_xyz[a > 0.5]
    return a, torch.stack((valid.amin(0), valid.amax(0)))


def shrink_bounds(aabb, units, grid, new_aabb):
    """Voxel range kept by TensorVMSplit.shrink (tensoRF.py:300-305): (t_l, b_r) long tensors."""
    xyz_min, xyz_max = new_aabb
    t_l, b_r = (xyz_min - aabb[0]) / units, (xyz_max - aabb[0]) / units
    t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
    b_r = torch.stack([b_r, torch.as_tensor(grid, dtype=torch.long)]).amin(0)
    return t_l, b_r


def shrink_factors(params, t_l, b_r):
    """The crops of tensoRF.py:307-321."""
    out = dict(params)
    tl, br = t_l.tolist(), b_r.tolist()
    for pre in ("density", "app"):
        for i in range(3):
            v = VEC_MODE[i]
            m0, m1 = MAT_MODE[i]
            out[f"{pre}_line.{i}"] = params[f"{pre}_line.{i}"][..., tl[v]:br[v], :]
            out[f"{pre}_plane.{i}"] = params[f"{pre}_plane.{i}"][..., tl[m1]:br[m1], tl[m0]:br[m0]]
    return out


def shrink_corrected_aabb(aabb, grid, t_l, b_r):
    """tensoRF.py:324-330 (taken when the mask grid differs from gridSize)."""
    g = torch.as_tensor(grid, dtype=torch.long)
    lo, hi = t_l / (g - 1), (b_r - 1) / (g - 1)
    return torch.stack(((1 - lo) * aabb[0] + lo * aabb[1], (1 - hi) * aabb[0] + hi * aabb[1]))
