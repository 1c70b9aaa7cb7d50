"""GPU parity of the pose -> ray kernels (csrc/pose_rays.cu, jt_pose_rays_fwd / _bwd) against golden
vectors of the live reference's camera.py and against oracle/camera_oracle.py on larger seeded inputs.

Tolerance class: fp32 floating point. center / ray <= 1e-5 relative to the largest entry (Taylor series
evaluated in powers of theta^2 instead of pow(theta, 2i)); d/d se3_refine <= 1e-4 relative (a sum over
B*R rays in a different order)."""
import math

import pytest
import torch

import joint_tensorf_b200 as jt
from common import camera_golden_names, load_golden, rel_err
from joint_tensorf_b200.options import Namespace
from oracle import camera_oracle as co

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _opt(H, W, ndc=False, shift=True, detach=False):
    return Namespace(H=H, W=W, camera=dict(model="perspective", ndc=ndc),
                     arch=dict(ndc_near_plane=1.0, ndc_center_shift=shift, detach_ndc_center_shift=detach))


@pytest.mark.parametrize("name", camera_golden_names())
def test_pose_rays_match_reference_golden(name):
    g = load_golden(name)
    c = g["case"]
    opt = _opt(c["H"], c["W"], c["ndc"], c.get("shift", True), c.get("detach", False))
    se3 = g["se3"].to(DEV).requires_grad_(True)
    center, ray = jt.camera.get_center_and_ray(opt, g["pose"].to(DEV), g["intr_inv"].to(DEV), ray_idx=g["ray_idx"].to(DEV),
                                               intr=g["intr"].to(DEV), se3_refine=se3)
    assert rel_err(center.cpu(), g["center"]) <= 1e-5
    assert rel_err(ray.cpu(), g["ray"]) <= 1e-5
    loss = (center * g["g_center"].to(DEV)).sum() + (ray * g["g_ray"].to(DEV)).sum()
    loss.backward()
    assert rel_err(se3.grad.cpu(), g["d_se3"]) <= 1e-4


def _blender_views(B, seed):
    gen = torch.Generator().manual_seed(seed)
    u, phi = torch.rand((B,), generator=gen), torch.rand((B,), generator=gen) * 2 * math.pi
    cz = u * 0.9 + 0.05
    cr = torch.sqrt(1 - cz * cz)
    centers = 4.0 * torch.stack([cr * torch.cos(phi), cr * torch.sin(phi), cz], dim=-1)
    rot, t = jt.synth._look_at_w2c(centers)
    return torch.cat([rot, t[..., None]], dim=-1), gen


def test_training_shape_against_oracle():
    """bench shape: 32 views x 128 shared pixels of an 800x800 image, se3 noise 0.15 (bat_blender_VM.yaml:108)."""
    B, H, W, R = 32, 800, 800, 128
    pose, gen = _blender_views(B, 3)
    intr = torch.tensor([[1111.1, 0, W / 2], [0, 1111.1, H / 2], [0, 0, 1]]).float()[None].repeat(B, 1, 1)
    se3 = 0.15 * torch.randn((B, 6), generator=gen)
    ray_idx = torch.randperm(H * W, generator=gen)[:R]
    gc, gr = torch.randn((B, R, 3), generator=gen), torch.randn((B, R, 3), generator=gen)
    se3_ref = se3.clone().requires_grad_(True)
    # the oracle generates all H*W rays like the reference; index a sub-image worth of them to stay fast
    c_ref, r_ref = co.rays_of_step(se3_ref, pose, intr.inverse(), H, W, ray_idx)
    ((c_ref * gc).sum() + (r_ref * gr).sum()).backward()
    se3_d = se3.to(DEV).requires_grad_(True)
    c, r = jt.camera.get_center_and_ray(_opt(H, W), pose.to(DEV), intr.inverse().to(DEV), ray_idx=ray_idx.to(DEV),
                                        se3_refine=se3_d)
    assert rel_err(c.cpu(), c_ref.detach()) <= 1e-5 and rel_err(r.cpu(), r_ref.detach()) <= 1e-5
    ((c * gc.to(DEV)).sum() + (r * gr.to(DEV)).sum()).backward()
    assert rel_err(se3_d.grad.cpu(), se3_ref.grad) <= 1e-4


def test_view_subset_rows_and_per_view_pixels():
    """se3_refine.weight[var.idx] (bat.py:350): rows picked by view_idx, gradient lands in those rows only;
    per-view pixel lists [B,R]."""
    n_rows, B, H, W, R = 10, 4, 60, 80, 33
    pose, gen = _blender_views(B, 5)
    intr = torch.tensor([[70.0, 0, W / 2], [0, 70.0, H / 2], [0, 0, 1]]).float()[None].repeat(B, 1, 1)
    table = 0.1 * torch.randn((n_rows, 6), generator=gen)
    idx = torch.tensor([7, 2, 9, 4])
    pix = torch.stack([torch.randperm(H * W, generator=gen)[:R] for _ in range(B)])
    gc, gr = torch.randn((B, R, 3), generator=gen), torch.randn((B, R, 3), generator=gen)
    t_ref = table.clone().requires_grad_(True)
    c_all, r_all = co.get_center_and_ray(H, W, co.refined_pose(t_ref[idx], pose), intr.inverse())
    c_ref = torch.gather(c_all, 1, pix[..., None].expand(-1, -1, 3))
    r_ref = torch.gather(r_all, 1, pix[..., None].expand(-1, -1, 3))
    ((c_ref * gc).sum() + (r_ref * gr).sum()).backward()
    t_d = table.to(DEV).requires_grad_(True)
    c, r = jt.camera.get_center_and_ray(_opt(H, W), pose.to(DEV), intr.inverse().to(DEV), ray_idx=pix.to(DEV),
                                        se3_refine=t_d, view_idx=idx.to(DEV))
    assert rel_err(c.cpu(), c_ref.detach()) <= 1e-5 and rel_err(r.cpu(), r_ref.detach()) <= 1e-5
    ((c * gc.to(DEV)).sum() + (r * gr.to(DEV)).sum()).backward()
    assert rel_err(t_d.grad.cpu(), t_ref.grad) <= 1e-4
    untouched = [i for i in range(n_rows) if i not in idx.tolist()]
    assert float(t_d.grad[untouched].abs().max()) == 0.0


def test_render_slice_and_pose_gradient():
    """pix = None: a contiguous render_by_slices slice (nerf.py:728-740); no se3 -> gradient w.r.t. the pose."""
    B, H, W = 2, 50, 70
    pose, gen = _blender_views(B, 9)
    intr = torch.tensor([[61.0, 0, W / 2], [0, 61.0, H / 2], [0, 0, 1]]).float()[None].repeat(B, 1, 1)
    base, n = 1234, 999
    p_ref = pose.clone().requires_grad_(True)
    c_all, r_all = co.get_center_and_ray(H, W, p_ref, intr.inverse())
    c_ref, r_ref = c_all[:, base:base + n], r_all[:, base:base + n]
    gc, gr = torch.randn((B, n, 3), generator=gen), torch.randn((B, n, 3), generator=gen)
    ((c_ref * gc).sum() + (r_ref * gr).sum()).backward()
    p_d = pose.to(DEV).requires_grad_(True)
    c, r = jt.camera.get_center_and_ray(_opt(H, W), p_d, intr.inverse().to(DEV), pix_base=base, n_rays=n)
    assert rel_err(c.cpu(), c_ref.detach()) <= 1e-5 and rel_err(r.cpu(), r_ref.detach()) <= 1e-5
    ((c * gc.to(DEV)).sum() + (r * gr.to(DEV)).sum()).backward()
    assert rel_err(p_d.grad.cpu(), p_ref.grad) <= 1e-4
    # whole image with one call: every pixel exactly once, row-major like the reference's meshgrid
    c_full, r_full = jt.camera.get_center_and_ray(_opt(H, W), pose.to(DEV), intr.inverse().to(DEV))
    assert c_full.shape == (B, H * W, 3)
    assert rel_err(r_full.cpu(), r_all.detach()) <= 1e-5


def test_errors_are_loud():
    opt = _opt(8, 8, ndc=True)
    pose = torch.eye(3, 4)[None].to(DEV)
    with pytest.raises(RuntimeError):
        jt.camera.get_center_and_ray(opt, pose, torch.eye(3)[None].to(DEV))           # ndc without intr
    with pytest.raises(RuntimeError):
        jt.camera.get_center_and_ray(_opt(8, 8), torch.eye(3, 4)[None], torch.eye(3)[None])   # CPU tensors
