"""Golden vectors for the pose -> ray path from the LIVE reference `camera.py`.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_camera.py

For each case: se3_refine -> camera.lie.se3_to_SE3 (camera.py:81-99) -> camera.pose.compose
(camera.py:43-58, as model/bat.py:350-353 calls it) -> camera.get_center_and_ray (camera.py:231-261)
-> [:, ray_idx] (model/tensorf.py:157-159) -> camera.convert_NDC (camera.py:303-340) when ndc,
then the autograd gradient of sum(center*g_center + ray*g_ray) w.r.t. se3_refine.
Stored as tests/golden/camera_<case>.pt (inputs + outputs).
"""
import math
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402
import joint_tensorf_b200.synth as synth  # noqa: E402

CASES = {
    "blender": dict(B=6, H=40, W=48, focal=66.0, R=64, ndc=False, se3_scale=0.15, seed=11),
    "blender_zero": dict(B=3, H=24, W=24, focal=33.0, R=32, ndc=False, se3_scale=0.0, seed=12),
    "llff_ndc": dict(B=4, H=36, W=48, focal=40.8, R=48, ndc=True, se3_scale=0.05, seed=13, shift=True, detach=False),
    "llff_ndc_detach": dict(B=4, H=36, W=48, focal=40.8, R=48, ndc=True, se3_scale=0.05, seed=14, shift=True, detach=True),
    "llff_ndc_noshift": dict(B=2, H=20, W=28, focal=24.0, R=40, ndc=True, se3_scale=0.05, seed=15, shift=False, detach=False),
}


def make_inputs(c):
    g = torch.Generator().manual_seed(c["seed"])
    B, H, W = c["B"], c["H"], c["W"]
    if c["ndc"]:
        # forward-facing cameras near the identity (LLFF-shaped, pose_eye o small motion)
        rv = 0.05 * torch.randn((B, 6), generator=g)
        _, cam = ref_loader.load()
        pose = cam.lie.se3_to_SE3(rv).detach()
    else:
        u = torch.rand((B,), generator=g)
        phi = torch.rand((B,), generator=g) * 2 * math.pi
        cz = u * 0.9 + 0.05
        cr = torch.sqrt(1 - cz * cz)
        centers = 4.0 * torch.stack([cr * torch.cos(phi), cr * torch.sin(phi), cz], dim=-1)
        rot, t = synth._look_at_w2c(centers)
        pose = torch.cat([rot, t[..., None]], dim=-1)
    intr = torch.tensor([[c["focal"], 0, W / 2], [0, c["focal"], H / 2], [0, 0, 1]]).float()[None].repeat(B, 1, 1)
    se3 = c["se3_scale"] * torch.randn((B, 6), generator=g)
    if c["se3_scale"] > 0:
        se3[0] = 0.0                           # one view still at its initial value
    ray_idx = torch.randperm(H * W, generator=g)[: c["R"]]
    g_center = torch.randn((B, c["R"], 3), generator=g)
    g_ray = torch.randn((B, c["R"], 3), generator=g)
    return pose, intr, se3, ray_idx, g_center, g_ray


def main():
    _, cam = ref_loader.load()
    warnings.filterwarnings("ignore")
    for name, c in CASES.items():
        pose, intr, se3, ray_idx, g_center, g_ray = make_inputs(c)
        opt = ref_loader.AttrDict(H=c["H"], W=c["W"], device="cpu", camera=dict(model="perspective", ndc=c["ndc"]),
                                  arch=dict(ndc_near_plane=1.0, ndc_center_shift=c.get("shift", True),
                                            detach_ndc_center_shift=c.get("detach", False)))
        se3_p = se3.clone().requires_grad_(True)
        pose_refine = cam.lie.se3_to_SE3(se3_p)                             # bat.py:351
        p = cam.pose.compose([pose_refine, pose])                           # bat.py:352
        center, ray = cam.get_center_and_ray(opt, p, intr_inv=intr.inverse())  # tensorf.py:152
        center, ray = center[:, ray_idx], ray[:, ray_idx]                   # tensorf.py:157-159
        if c["ndc"]:
            center, ray = cam.convert_NDC(opt, center, ray, intr=intr)      # tensorf.py:160-163
        loss = (center * g_center).sum() + (ray * g_ray).sum()
        (d_se3,) = torch.autograd.grad(loss, se3_p)
        out = dict(case=c, pose=pose, intr=intr, intr_inv=intr.inverse(), se3=se3, ray_idx=ray_idx, g_center=g_center,
                   g_ray=g_ray, refined_pose=p.detach(), center=center.detach(), ray=ray.detach(), d_se3=d_se3)
        path = os.path.join(HERE, f"camera_{name}.pt")
        torch.save(out, path)
        print(name, "center", tuple(center.shape), "|d_se3|max", float(d_se3.abs().max()), "->", path)


if __name__ == "__main__":
    main()
