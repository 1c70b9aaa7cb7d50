"""Pose -> ray generation on the B200 path.

Host-side mirror of the slice of the reference's `camera.py` / `model/bat.py` that sits
between the pose parameters and `B200_VMSplit.forward`:

    reference                                             here
    ---------------------------------------------------   ------------------------------------------
    camera.lie.se3_to_SE3 (camera.py:81-99)               folded into `get_center_and_ray(..., se3_refine=)`
    camera.pose.compose   (camera.py:43-58)               (one kernel computes the refined pose per view)
    camera.get_center_and_ray(opt, pose, intr_inv)        `get_center_and_ray(opt, pose, intr_inv, ray_idx=...)`
      -> [:, ray_idx]  (model/tensorf.py:157-159)         only the requested rays are generated
    camera.convert_NDC(opt, center, ray, intr)            folded in when `opt.camera.ndc` (intr= required)

All arithmetic is in csrc/pose_rays.cu (C ABI: jt_pose_rays_fwd / jt_pose_rays_bwd); gradients
flow to `se3_refine` (joint pose optimisation, test-time pose optimisation) and, when the caller
composed the pose itself in torch, to `pose`. There is no CPU implementation.
"""
import torch

from . import _lib, ops
from ._lib import check
from .ops import TIMER, _p, _stream


class _RayCfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class PoseRays(torch.autograd.Function):
    """(se3_refine | None, pose [B,3,4]) -> center, ray [B,R,3]."""

    @staticmethod
    def forward(ctx, cfg, se3, pose, intr_inv, intr, pix, view_idx):
        ops._need_cuda(pose, "pose")
        lib = _lib.lib()
        dev = pose.device
        pose_c = pose.detach().contiguous().float()
        B, R = cfg.n_views, cfg.n_rays
        if pose_c.dim() == 2:
            pose_c = pose_c[None]
        if pose_c.shape[0] not in (1, B):
            raise _lib.JtError(f"pose has {pose_c.shape[0]} entries for {B} views")
        base_per_view = 0 if (pose_c.shape[0] == 1 and B > 1) else 1
        se3_c = se3.detach().contiguous().float() if se3 is not None else None
        kinv = intr_inv.detach().contiguous().float()
        kinv = kinv[None] if kinv.dim() == 2 else kinv
        kinv_per_view = 0 if (kinv.shape[0] == 1 and B > 1) else 1
        k = None
        k_per_view = 0
        if cfg.ndc:
            if intr is None:
                raise _lib.JtError("opt.camera.ndc needs intr= (camera.convert_NDC reads intr[:,0,0]/intr[:,0,2])")
            k = intr.detach().contiguous().float()
            k = k[None] if k.dim() == 2 else k
            k_per_view = 0 if (k.shape[0] == 1 and B > 1) else 1
        pix_c, pix_per_view = None, 0
        if pix is not None:
            pix_c = pix.to(device=dev, dtype=torch.int32).contiguous()
            pix_per_view = 1 if pix_c.dim() == 2 else 0
        vi = view_idx.to(device=dev, dtype=torch.int32).contiguous() if view_idx is not None else None
        center = torch.empty((B, R, 3), device=dev)
        ray = torch.empty((B, R, 3), device=dev)
        args = (_p(se3_c), _p(vi), _p(pose_c), base_per_view, _p(kinv), kinv_per_view, _p(k), k_per_view, _p(pix_c),
                pix_per_view, int(cfg.pix_base), B, R, int(cfg.width), int(cfg.ndc), int(cfg.center_shift),
                int(cfg.detach_shift), float(cfg.near))
        with TIMER.span("pose_rays_fwd"):
            check(lib.jt_pose_rays_fwd(*args, _p(center), _p(ray), 0, _stream()), "jt_pose_rays_fwd")
        ctx.args = args
        ctx.keep = (se3_c, vi, pose_c, kinv, k, pix_c)         # keep the buffers behind the raw pointers alive
        ctx.se3_shape = None if se3 is None else tuple(se3.shape)
        ctx.pose_shape = tuple(pose.shape)
        ctx.base_per_view = base_per_view
        return center, ray

    @staticmethod
    def backward(ctx, g_center, g_ray):
        lib = _lib.lib()
        dev = g_center.device
        B = ctx.args[11]
        want_se3 = ctx.se3_shape is not None and ctx.needs_input_grad[1]
        want_pose = ctx.needs_input_grad[2]
        if not (want_se3 or want_pose):
            return (None,) * 7
        if want_pose and not ctx.base_per_view:
            raise _lib.JtError("gradient w.r.t. a pose shared by several views is not implemented")
        if want_pose and ctx.se3_shape is not None:
            raise _lib.JtError("pose gradient is only provided when no se3_refine is folded in")
        g_center = g_center.contiguous().float()
        g_ray = g_ray.contiguous().float()
        scratch = torch.empty((B, 12), device=dev)
        d_se3 = torch.zeros(ctx.se3_shape, device=dev) if want_se3 else None
        d_pose = torch.empty((B, 3, 4), device=dev) if want_pose else None
        with TIMER.span("pose_rays_bwd"):
            check(lib.jt_pose_rays_bwd(*ctx.args, _p(g_center), _p(g_ray), _p(scratch), _p(d_se3), _p(d_pose),
                                       _stream()), "jt_pose_rays_bwd")
        if d_pose is not None:
            d_pose = d_pose.view(ctx.pose_shape)
        return None, d_se3, d_pose, None, None, None, None


def get_center_and_ray(opt, pose, intr_inv=None, ray_idx=None, intr=None, se3_refine=None, view_idx=None,
                       pix_base=0, n_rays=None):
    """Drop-in for `camera.get_center_and_ray(opt, pose, intr_inv)` (camera.py:231-261) followed by the
    `[:, ray_idx]` selection and `camera.convert_NDC` of `Graph.render` (model/tensorf.py:157-164).

    pose      [B,3,4] world-to-camera. With `se3_refine` ([n_rows,6], e.g. `se3_refine.weight`) the pose
              handed in is the *un-refined* one (`pose_noise o GT`, bat.py:346-348) and
              `compose([se3_to_SE3(se3_refine[view_idx]), pose])` (bat.py:350-353) happens inside the kernel;
              gradients then flow to `se3_refine`. Without it, gradients flow to `pose`.
    ray_idx   [R] pixel indices shared by all views (nerf.py:657-658), [B,R] per-view indices, or None:
              the contiguous pixels pix_base .. pix_base + n_rays - 1 (a `render_by_slices` slice;
              the whole image when n_rays is None).
    Reads opt.H, opt.W, opt.camera.ndc, opt.arch.ndc_near_plane / ndc_center_shift / detach_ndc_center_shift
    exactly like the reference. Returns center, ray [B,R,3]."""
    if getattr(opt.camera, "model", "perspective") != "perspective":
        raise _lib.JtError("only the perspective camera model is implemented (camera.py:233)")
    if intr_inv is None:
        raise _lib.JtError("intr_inv is required")
    B = pose.shape[0] if pose.dim() == 3 else 1
    if view_idx is not None:
        B = int(view_idx.shape[0])
    elif se3_refine is not None and pose.dim() == 3 and pose.shape[0] == 1:
        B = int(se3_refine.shape[0])
    if ray_idx is not None:
        R = int(ray_idx.shape[-1])
    else:
        R = int(n_rays) if n_rays is not None else int(opt.H) * int(opt.W) - int(pix_base)
    arch = getattr(opt, "arch", None)
    ndc = bool(getattr(opt.camera, "ndc", False))
    near = float(getattr(arch, "ndc_near_plane", 0.1)) if arch is not None and hasattr(arch, "ndc_near_plane") else 0.1
    shift = not (arch is not None and hasattr(arch, "ndc_center_shift") and arch.ndc_center_shift is False)
    detach = bool(arch is not None and getattr(arch, "detach_ndc_center_shift", False))
    cfg = _RayCfg(n_views=B, n_rays=R, width=int(opt.W), ndc=ndc, near=near, center_shift=shift, detach_shift=detach,
                  pix_base=int(pix_base))
    return PoseRays.apply(cfg, se3_refine, pose, intr_inv, intr, ray_idx, view_idx)


@torch.no_grad()
def refined_pose(se3, pose, view_idx=None):
    """`camera.pose.compose([camera.lie.se3_to_SE3(se3), pose])` (camera.py:81-99,43-58) -> [B,3,4], no gradient.
    Used for the fixed part of the training pose, `compose([pose_noise, GT])` (bat.py:346-348), and for
    reading back the current refined poses (evaluation / logging)."""
    ops._need_cuda(pose, "pose")
    lib = _lib.lib()
    dev = pose.device
    pose_c = pose.detach().contiguous().float()
    pose_c = pose_c[None] if pose_c.dim() == 2 else pose_c
    se3_c = se3.detach().contiguous().float()
    vi = view_idx.to(device=dev, dtype=torch.int32).contiguous() if view_idx is not None else None
    B = int(vi.shape[0]) if vi is not None else int(se3_c.shape[0])
    if pose_c.shape[0] not in (1, B):
        raise _lib.JtError(f"pose has {pose_c.shape[0]} entries for {B} views")
    eye = torch.eye(3, device=dev)
    scratch = torch.empty((2, B, 1, 3), device=dev)
    out = torch.empty((B, 3, 4), device=dev)
    check(lib.jt_pose_rays_fwd(_p(se3_c), _p(vi), _p(pose_c), 0 if (pose_c.shape[0] == 1 and B > 1) else 1, _p(eye), 0,
                               0, 0, 0, 0, 0, B, 1, 1, 0, 0, 0, 0.0, _p(scratch[0]), _p(scratch[1]), _p(out),
                               _stream()), "jt_pose_rays_fwd")
    return out
