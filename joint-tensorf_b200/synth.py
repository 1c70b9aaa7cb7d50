"""Deterministic synthetic scenes and ray batches.

No dataset is available offline, so benchmarks and parity tests run on
Blender-shaped and LLFF-shaped rays generated here:

* Blender-shaped: pinhole cameras on the upper hemisphere of radius 4 looking
  at the origin (+z forward, y down -- the convention the reference ends up with
  after data/blender.py:86-91), 800x800 pixels, focal 1111.1; ray directions
  are un-normalised with camera-space z == 1 exactly as
  camera.get_center_and_ray builds them (camera.py:231-261).
* LLFF-shaped: forward-facing cameras near the identity pose, 1008x756,
  focal 0.85 W, rays converted to NDC with the near plane at 1
  (camera.convert_NDC, camera.py:303-340).

Everything is pure torch on CPU so that the oracle and the CUDA path see
bit-identical inputs.
"""
import math

import torch


def _look_at_w2c(centers):
    """World-to-camera rotations for cameras at `centers` looking at the origin.
    Camera axes: +z forward, +y down, +x right. Returns R [B,3,3] (rows = camera
    axes in world coordinates) and t = -R c."""
    fwd = -centers / centers.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    down = torch.cross(fwd, right, dim=-1)
    rot = torch.stack([right, down, fwd], dim=-2)
    t = -(rot @ centers[..., None])[..., 0]
    return rot, t


def blender_rays(n_rays=4096, n_views=32, hw=(800, 800), focal=1111.1, radius=4.0, seed=1):
    """n_rays rays spread evenly over n_views hemisphere cameras; the same pixel
    subset is used for all views (reference nerf.py:657-658). Returns
    (rays_o [N,3], rays_d [N,3], view_index [N]) fp32."""
    g = torch.Generator().manual_seed(seed)
    h, w = hw
    per_view = max(1, n_rays // n_views)
    u = torch.rand((n_views,), generator=g)
    phi = torch.rand((n_views,), generator=g) * 2 * math.pi
    cz = u * 0.9 + 0.05
    cr = torch.sqrt(1 - cz * cz)
    centers = radius * torch.stack([cr * torch.cos(phi), cr * torch.sin(phi), cz], dim=-1)
    rot, _ = _look_at_w2c(centers)
    pix = torch.randperm(h * w, generator=g)[:per_view]
    py = (pix // w).float() + 0.5
    px = (pix % w).float() + 0.5
    cam = torch.stack([(px - w / 2) / focal, (py - h / 2) / focal, torch.ones_like(px)], dim=-1)  # [P,3]
    dirs = cam[None] @ rot                                   # [B,P,3]  (grid_3D @ R_inv, camera.py:251)
    orig = centers[:, None, :].expand_as(dirs)
    view = torch.arange(n_views)[:, None].expand(n_views, per_view)
    o, d, v = orig.reshape(-1, 3), dirs.reshape(-1, 3), view.reshape(-1)
    return o[:n_rays].contiguous(), d[:n_rays].contiguous(), v[:n_rays].contiguous()


def blender_views(n_views=32, hw=(800, 800), focal=1111.1, radius=4.0, seed=1):
    """The cameras of `blender_rays` as matrices: world-to-camera poses [B,3,4] and intrinsics [B,3,3]
    (what the reference keeps per view: var.pose, var.intr; data/blender.py:86-104)."""
    g = torch.Generator().manual_seed(seed)
    h, w = hw
    u = torch.rand((n_views,), generator=g)
    phi = torch.rand((n_views,), generator=g) * 2 * math.pi
    cz = u * 0.9 + 0.05
    cr = torch.sqrt(1 - cz * cz)
    centers = radius * torch.stack([cr * torch.cos(phi), cr * torch.sin(phi), cz], dim=-1)
    rot, t = _look_at_w2c(centers)
    pose = torch.cat([rot, t[..., None]], dim=-1)
    intr = torch.tensor([[focal, 0.0, w / 2], [0.0, focal, h / 2], [0.0, 0.0, 1.0]])[None].repeat(n_views, 1, 1)
    return pose.contiguous(), intr.contiguous()


def frame_rays(view=0, hw=(800, 800), focal=1111.1, radius=4.0, seed=2):
    """All H*W rays of one hemisphere camera (full-frame render, config 5)."""
    g = torch.Generator().manual_seed(seed + 7919 * view)
    h, w = hw
    u = torch.rand((1,), generator=g)
    phi = torch.rand((1,), generator=g) * 2 * math.pi
    cz = u * 0.9 + 0.05
    cr = torch.sqrt(1 - cz * cz)
    center = radius * torch.stack([cr * torch.cos(phi), cr * torch.sin(phi), cz], dim=-1)
    rot, _ = _look_at_w2c(center)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32) + 0.5,
                            torch.arange(w, dtype=torch.float32) + 0.5, indexing="ij")
    cam = torch.stack([(xs - w / 2) / focal, (ys - h / 2) / focal, torch.ones_like(xs)], dim=-1).view(-1, 3)
    d = cam @ rot[0]
    o = center.expand_as(d)
    return o.contiguous(), d.contiguous()


def llff_views(n_views=8, hw=(756, 1008), seed=1):
    """The cameras of `llff_ndc_rays` as matrices: world-to-camera poses [B,3,4] (small random rigid
    motions of the identity) and intrinsics [B,3,3] with focal 0.85 W."""
    g = torch.Generator().manual_seed(seed)
    h, w = hw
    focal = 0.85 * w
    rv = 0.05 * torch.randn((n_views, 3), generator=g)
    tv = 0.05 * torch.randn((n_views, 3), generator=g)
    ang = rv.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    k = rv / ang
    kx = torch.zeros(n_views, 3, 3)
    kx[:, 0, 1], kx[:, 0, 2], kx[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    kx[:, 1, 2], kx[:, 2, 0], kx[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    a = ang[..., None]
    rot = torch.eye(3)[None] + torch.sin(a) * kx + (1 - torch.cos(a)) * (kx @ kx)
    pose = torch.cat([rot, tv[..., None]], dim=-1)
    intr = torch.tensor([[focal, 0.0, w / 2], [0.0, focal, h / 2], [0.0, 0.0, 1.0]])[None].repeat(n_views, 1, 1)
    return pose.contiguous(), intr.contiguous()


def llff_ndc_rays(n_rays=4096, n_views=8, hw=(756, 1008), seed=1, near=1.0):
    """Forward-facing NDC rays (cfg4). Poses are small random rigid motions of
    the identity; rays are shifted to the near plane and projected as in
    camera.convert_NDC (camera.py:303-340). Returns (o [N,3], d [N,3], view [N])."""
    g = torch.Generator().manual_seed(seed)
    h, w = hw
    focal = 0.85 * w
    per_view = max(1, n_rays // n_views)
    rv = 0.05 * torch.randn((n_views, 3), generator=g)
    tv = 0.05 * torch.randn((n_views, 3), generator=g)
    ang = rv.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    k = rv / ang
    kx = torch.zeros(n_views, 3, 3)
    kx[:, 0, 1], kx[:, 0, 2], kx[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    kx[:, 1, 2], kx[:, 2, 0], kx[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    a = ang[..., None]
    rot = torch.eye(3)[None] + torch.sin(a) * kx + (1 - torch.cos(a)) * (kx @ kx)   # w2c rotation
    pix = torch.randperm(h * w, generator=g)[:per_view]
    py = (pix // w).float() + 0.5
    px = (pix % w).float() + 0.5
    cam = torch.stack([(px - w / 2) / focal, (py - h / 2) / focal, torch.ones_like(px)], dim=-1)
    ray = cam[None] @ rot                                       # [B,P,3]
    center = -(tv[:, None, :] @ rot).expand_as(ray)             # camera.py:252
    center = center + (near - center[..., 2:]) / ray[..., 2:] * ray
    cx, cy, cz = center.unbind(-1)
    rx, ry, rz = ray.unbind(-1)
    sx, sy = focal / (w / 2), focal / (h / 2)
    o = torch.stack([sx * (cx / cz), sy * (cy / cz), 1 - 2 * near / cz], dim=-1)
    d = torch.stack([sx * (rx / rz - cx / cz), sy * (ry / rz - cy / cz), 2 * near / cz], dim=-1)
    view = torch.arange(n_views)[:, None].expand(n_views, per_view)
    return (o.reshape(-1, 3)[:n_rays].contiguous(), d.reshape(-1, 3)[:n_rays].contiguous(),
            view.reshape(-1)[:n_rays].contiguous())


# ---------------------------------------------------------------- workload configurations
def describe(name, n_rays=None):
    """One-line description of a workload for bench.py's `config.workload`."""
    kw, run = config(name)
    g = kw["gridSize"]
    grid = f"{g[0]}^3" if g[0] == g[1] == g[2] else "x".join(str(x) for x in g)
    head = {"SH": "SH(deg 2) shading", "MLP_Fea": f"MLP_Fea {kw['featureC']} head",
            "MLP_Fea_WeakView": f"MLP_Fea_WeakView {kw['featureC']} head"}[kw["shadingMode"]]
    rays = "LLFF forward-facing NDC rays" if run["ndc"] else "Blender-shaped rays"
    s = (f"{name}: TensoRF-VM {grid}, density 3x{kw['density_n_comp'][0]} / appearance 3x{kw['appearance_n_comp'][0]} "
         f"comps, app_dim {kw['app_dim']}, {head}, S={run['n_samples']}, {rays}")
    return s + (f", {n_rays} rays/GPU" if n_rays else "")


def config(name):
    """Constructor keyword sets of the benchmark configurations.
    `n_samples` follows model/tensorf.py:449-461:
    min(sample_intvs, int(||grid||_2 / step_ratio))."""
    blender = dict(aabb=[[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], density_n_comp=[16] * 3,
                   appearance_n_comp=[48] * 3, app_dim=27, near_far=[2.0, 6.0], shadingMode="MLP_Fea",
                   alphaMask_thres=1e-4, density_shift=-10, distance_scale=25.0, pos_pe=2, view_pe=2,
                   fea_pe=2, featureC=64, step_ratio=0.5, fea2denseAct="softplus",
                   volume_init_scale=0.1, rayMarch_weight_thres=1e-6, volume_init_bias=0.0)
    if name == "cfg1":
        g = [128] * 3
        return dict(blender, gridSize=g), dict(n_samples=min(1000, int(math.sqrt(3 * 128 ** 2) / 0.5)), ndc=False, white_bg=True)
    if name == "cfg2":
        g = [300] * 3
        return dict(blender, gridSize=g), dict(n_samples=1000, ndc=False, white_bg=True)
    if name == "cfg2_sh":
        g = [300] * 3
        return dict(blender, gridSize=g, shadingMode="SH"), dict(n_samples=1000, ndc=False, white_bg=True)
    if name == "cfg4":
        llff = dict(aabb=[[-1.5, -1.67, -2.0], [1.5, 1.67, 1.0]], gridSize=[617, 687, 617],
                    density_n_comp=[16] * 3, appearance_n_comp=[20] * 3, app_dim=20, near_far=[-1.0, 1.0],
                    shadingMode="MLP_Fea_WeakView", alphaMask_thres=1e-4, density_shift=0.0, distance_scale=25.0,
                    pos_pe=2, view_pe=2, fea_pe=2, featureC=32, step_ratio=0.3, fea2denseAct="relu",
                    volume_init_scale=0.05, rayMarch_weight_thres=1e-7, volume_init_bias=0.2)
        return llff, dict(n_samples=1000, ndc=True, white_bg=False)
    raise KeyError(name)
