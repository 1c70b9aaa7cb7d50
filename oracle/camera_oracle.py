"""CPU oracle for the pose -> ray generation that feeds the render hot path (TEST INFRASTRUCTURE).

Checker only: imported by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU
legs, never by the product (`joint-tensorf_b200/`).

Restates, over torch CPU tensors (differentiable through autograd, so it also yields
the reference gradients w.r.t. the se(3) refinement), what one training step of the
reference does between the pose parameters and the `center` / `ray_dir` arguments of
`BAT_VMSplit.forward`:

    model/bat.py:341-353     get_pose: se3_refine.weight[idx] -> lie.se3_to_SE3 -> pose.compose
    model/tensorf.py:144-166 render: camera.get_center_and_ray for ALL H*W pixels -> [:, ray_idx] -> convert_NDC
    camera.py:81-99          Lie.se3_to_SE3 (Taylor series, nth = 8)
    camera.py:43-58          Pose.compose / compose_pair
    camera.py:231-261        get_center_and_ray
    camera.py:303-340        convert_NDC

Parity pin: `tests/golden/make_golden_camera.py` runs the LIVE reference `camera.py`
(imported from /root/reference in the build container) on seeded inputs and commits
inputs + outputs + autograd gradients as `tests/golden/camera_*.pt`;
`tests/test_oracle_golden.py` checks this file against them.
"""
import torch


def skew_symmetric(w):
    """camera.py:113-119."""
    w0, w1, w2 = w.unbind(dim=-1)
    O = torch.zeros_like(w0)
    return torch.stack([torch.stack([O, -w2, w1], dim=-1),
                        torch.stack([w2, O, -w0], dim=-1),
                        torch.stack([-w1, w0, O], dim=-1)], dim=-2)


def taylor_A(x, nth=10):
    """sin(x)/x, camera.py:121-129."""
    ans = torch.zeros_like(x)
    denom = 1.
    for i in range(nth + 1):
        if i > 0:
            denom *= (2 * i) * (2 * i + 1)
        ans = ans + (-1) ** i * (x ** (2 * i) / denom)
    return ans


def taylor_B(x, nth=10):
    """(1-cos(x))/x^2, camera.py:130-137."""
    ans = torch.zeros_like(x)
    denom = 1.
    for i in range(nth + 1):
        denom *= (2 * i + 1) * (2 * i + 2)
        ans = ans + (-1) ** i * (x ** (2 * i) / denom)
    return ans


def taylor_C(x, nth=10):
    """(x-sin(x))/x^3, camera.py:138-145."""
    ans = torch.zeros_like(x)
    denom = 1.
    for i in range(nth + 1):
        denom *= (2 * i + 2) * (2 * i + 3)
        ans = ans + (-1) ** i * (x ** (2 * i) / denom)
    return ans


def se3_to_SE3(wu):
    """camera.py:81-99: [...,6] (w, u) -> [...,3,4] (R | V u), series truncated at nth = 8."""
    w, u = wu.split([3, 3], dim=-1)
    wx = skew_symmetric(w)
    theta = w.norm(dim=-1)[..., None, None]
    I = torch.eye(3, dtype=torch.float32)
    A = taylor_A(theta, nth=8)
    B = taylor_B(theta, nth=8)
    C = taylor_C(theta, nth=8)
    R = I + A * wx + B * wx @ wx
    V = I + B * wx + C * wx @ wx
    return torch.cat([R, (V @ u[..., None])], dim=-1)


def compose_pair(pose_a, pose_b):
    """camera.py:50-58: pose_new(x) = pose_b o pose_a(x)."""
    R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
    R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
    R_new = R_b @ R_a
    t_new = (R_b @ t_a + t_b)[..., 0]
    return torch.cat([R_new, t_new[..., None]], dim=-1)


def refined_pose(se3_refine, pose):
    """model/bat.py:350-353: pose = compose([se3_to_SE3(se3_refine), pose])."""
    return compose_pair(se3_to_SE3(se3_refine), pose)


def get_center_and_ray(H, W, pose, intr_inv):
    """camera.py:231-261 for all H*W pixels: centers, ray_dirs [B,HW,3]."""
    y_range = torch.arange(H, dtype=torch.float32).add_(0.5)
    x_range = torch.arange(W, dtype=torch.float32).add_(0.5)
    Y, X = torch.meshgrid(y_range, x_range, indexing="ij")
    xy_grid = torch.stack([X, Y], dim=-1).view(-1, 2)
    xy_grid = xy_grid.repeat(len(pose), 1, 1)
    xy_hom = torch.cat([xy_grid, torch.ones_like(xy_grid[..., :1])], dim=-1)       # to_hom, camera.py:202-205
    grid_3D = xy_hom @ intr_inv.transpose(-1, -2)                                   # img2cam, camera.py:213-214
    t = pose[..., 3:]
    R_inv = pose[..., :3]
    ray_dirs = grid_3D @ R_inv
    centers = -(t.transpose(-2, -1) @ R_inv).repeat(1, H * W, 1)
    return centers, ray_dirs


def convert_NDC(center, ray, intr, near=1.0, center_shift=True, detach_shift=False):
    """camera.py:303-340."""
    if center_shift:
        shift = (near - center[..., 2:]) / ray[..., 2:] * ray
        center = center + (shift.detach() if detach_shift else shift)
    cx, cy, cz = center.unbind(dim=-1)
    rx, ry, rz = ray.unbind(dim=-1)
    scale_x = intr[:, 0, 0] / intr[:, 0, 2]
    scale_y = intr[:, 1, 1] / intr[:, 1, 2]
    cxoz, cyoz = cx / cz, cy / cz
    rxoz, ryoz = rx / rz, ry / rz
    cnx = scale_x[:, None] * cxoz
    cny = scale_y[:, None] * cyoz
    cnz = 1 - 2 * near / cz
    rnx = scale_x[:, None] * (rxoz - cxoz)
    rny = scale_y[:, None] * (ryoz - cyoz)
    rnz = 2 * near / cz
    return torch.stack([cnx, cny, cnz], dim=-1), torch.stack([rnx, rny, rnz], dim=-1)


def rays_of_step(se3_refine, pose, intr_inv, H, W, ray_idx, intr=None, ndc=False, near=1.0, center_shift=True,
                 detach_shift=False):
    """One training step's ray set (model/tensorf.py:144-166 after bat.py:341-353):
    center, ray_dir [B, R, 3] for the pixel subset `ray_idx` ([R], shared by all views)."""
    p = refined_pose(se3_refine, pose) if se3_refine is not None else pose
    center, ray = get_center_and_ray(H, W, p, intr_inv)
    center, ray = center[:, ray_idx], ray[:, ray_idx]
    if ndc:
        center, ray = convert_NDC(center, ray, intr, near, center_shift, detach_shift)
    return center, ray
