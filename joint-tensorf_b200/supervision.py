"""2-D supervision pre-processing and the render loss on csrc/image_prep.cu.

Mirrors, with the reference's names and `opt` fields:
    process_GT_images  <- Model.process_GT_images   (model/nerf.py:57-113; wandb / matplotlib logging left out)
    get_edge_mask      <- Model.get_edge_mask       (model/nerf.py:116-149)
    render_loss        <- the render term of Graph.compute_loss (model/tensorf.py:99-124), pixel gathers fused
The reference re-runs the first two every 500 iterations (nerf.py:172-175) for every scale of
`c2f_alternate_2D_scale_pool`, and the third every step. No CPU / ATen fallback: CUDA tensors only.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, floats
from .ops import TIMER, _need_cuda, _p, _stream
from .vmsplit import average_taps, gaussian_taps


def interp_schedule(x, schedule, left=0, right=1):
    """reference util.py:217-225."""
    assert left <= x <= right
    xs = np.linspace(left, right, len(schedule))
    return np.interp(x, xs, np.asarray(schedule, dtype=np.float64))


def _scales(opt):
    if getattr(opt, "c2f_alternate_2D_mode", None) == "sample":          # nerf.py:63-66
        return list(opt.c2f_alternate_2D_scale_pool)
    return [0.0, 1.0]


def blur_taps_2d(opt, it, scale):
    """(taps [ksize] fp32 host tensor, kernel_width) for one blur scale (nerf.py:70-83)."""
    blur_param = float(interp_schedule(float(it / opt.max_iter), opt.blur_2d_c2f_schedule)) * scale
    width = blur_param * (opt.W + opt.H) / 2
    if opt.blur_2d_mode == "uniform-gaussian":
        taps = gaussian_taps(width, opt.blur_2d_c2f_kernel_size)
    elif opt.blur_2d_mode == "uniform-average":
        taps = average_taps(width, opt.blur_2d_c2f_kernel_size)
    else:
        raise NotImplementedError("illegal blur_2d_mode")
    return taps.float(), width


def image_blur(images, taps):
    """Separable replicate-padded correlation of [..., H, W] fp32 CUDA images with `taps` (host list / tensor)
    along W then H (nerf.py:98-110)."""
    _need_cuda(images, "images")
    assert images.dtype == torch.float32 and images.dim() >= 2
    x = images.contiguous()
    h, w = x.shape[-2:]
    n_img = x.numel() // (h * w)
    vals = [float(v) for v in (taps.tolist() if torch.is_tensor(taps) else taps)]
    out = torch.empty_like(x)
    tmp = torch.empty_like(x)
    with TIMER.span("image_blur"):
        check(_lib.lib().jt_image_blur(_p(x), _p(out), _p(tmp), n_img, h, w, floats(vals), len(vals), _stream()),
              "jt_image_blur")
    return out


@torch.no_grad()
def process_GT_images(opt, images, it):
    """images [B,3,H,W] -> {scale: blurred images [B,3,H,W]} (scales whose kernel width is < 0.01 share the input,
    nerf.py:95-97)."""
    assert images.shape[-2:] == (opt.H, opt.W)
    out = {}
    for sc in _scales(opt):
        taps, width = blur_taps_2d(opt, it, sc)
        out[sc] = images if width < 0.01 else image_blur(images, taps)
    return out


def edge_mask(images, soft=False, thresh=1.25, return_stats=False):
    """images [B,3,H,W] -> [B, H*W] soft (float, GG / max) or hard (uint8, GG > mean * thresh) edge mask."""
    _need_cuda(images, "images")
    assert images.dim() == 4 and images.shape[1] == 3 and images.dtype == torch.float32
    x = images.contiguous()
    b, _, h, w = x.shape
    dev = x.device
    gg = torch.empty((b, h * w), device=dev, dtype=torch.float32)
    ws = torch.empty((int(_lib.lib().jt_edge_mask_ws_floats(b, h, w)),), device=dev, dtype=torch.float32)
    stats = torch.empty((b, 2), device=dev, dtype=torch.float32)
    mask = torch.empty((b, h * w), device=dev, dtype=torch.float32 if soft else torch.uint8)
    with TIMER.span("edge_mask"):
        check(_lib.lib().jt_edge_mask(_p(x), b, h, w, int(bool(soft)), float(thresh), _p(gg), _p(ws), _p(stats),
                                      _p(mask) if soft else None, None if soft else _p(mask), _stream()),
              "jt_edge_mask")
    return (mask, gg, stats) if return_stats else mask


@torch.no_grad()
def get_edge_mask(opt, blurred_gt_cached_images):
    """{scale: images} -> {scale: mask [B, H*W]} (nerf.py:116-149)."""
    soft = bool(getattr(opt, "soft_edge_mask", False))
    thresh = getattr(opt, "hard_edge_mask_mean_thresh", 1.25)
    return {sc: edge_mask(blurred_gt_cached_images[sc], soft, thresh) for sc in _scales(opt)}


def _i32(t):
    if t is None:
        return None
    return t if t.dtype == torch.int32 else t.to(torch.int32)


class RenderLoss(torch.autograd.Function):
    """loss = MSE / edge-weighted MSE between rgb [B,n,3] and the pixels `ray_idx` of views `view_idx` of the
    image cache; one launch forward (gather + both sums), one backward."""

    @staticmethod
    def forward(ctx, rgb, images, mask, ray_idx, view_idx, mode, fe, fn):
        _need_cuda(rgb, "rgb")
        _need_cuda(images, "images")
        assert rgb.dim() == 3 and rgb.shape[-1] == 3 and rgb.dtype == torch.float32
        assert images.dim() == 4 and images.shape[1] == 3 and images.dtype == torch.float32
        rgb_c = rgb.detach().contiguous()
        images = images.contiguous()
        b, n, _ = rgb_c.shape
        hw = images.shape[2] * images.shape[3]
        kind = 0
        if mode != 0:
            assert mask is not None and mask.dim() == 2 and mask.shape[1] == hw and mask.is_contiguous()
            kind = {torch.float32: 1, torch.uint8: 2}[mask.dtype]
        assert view_idx is not None or images.shape[0] == b
        ws = torch.empty((4,), device=rgb.device, dtype=torch.float64)
        loss = torch.empty((), device=rgb.device, dtype=torch.float32)
        args = (_p(rgb_c), _p(images), _p(mask) if mode != 0 else None, kind, _p(ray_idx), _p(view_idx), b, n, hw,
                int(mode), float(fe), float(fn))
        with TIMER.span("render_loss_fwd"):
            check(_lib.lib().jt_render_loss_fwd(*args, _p(ws), _p(loss), _stream()), "jt_render_loss_fwd")
        ctx.keep = (rgb_c, images, mask, ray_idx, view_idx, ws)
        ctx.args = args
        return loss

    @staticmethod
    def backward(ctx, g):
        rgb_c = ctx.keep[0]
        d = torch.empty_like(rgb_c)
        g = g.contiguous().float()
        with TIMER.span("render_loss_bwd"):
            check(_lib.lib().jt_render_loss_bwd(*ctx.args, _p(ctx.keep[5]), _p(g), _p(d), _stream()),
                  "jt_render_loss_bwd")
        return d, None, None, None, None, None, None, None


def render_loss(opt, rgb, images, ray_idx=None, train_edge_masks=None, it=0, mode="train", view_idx=None):
    """`loss.render` of Graph.compute_loss (model/tensorf.py:99-124).

    rgb [B,n,3]; images [Bc,3,H,W] (var.image: the blurred GT cache of the sampled scale, or the raw images);
    ray_idx [n] pixel indices shared by the views (None in validation: all pixels); train_edge_masks [Bc,H*W] or
    None; view_idx [B] rows of the cache (None: the cache holds exactly the B views, as in the reference)."""
    if getattr(opt, "edge_mask_on_render_loss", False):                        # tensorf.py:104-110
        edge_loss_on = it % 2 == 0 if getattr(opt, "alternate_edge_loss", False) else True
    else:
        edge_loss_on = False
    kind = 0
    if edge_loss_on and mode in ["train"] and it < opt.edge_mask_before_iter:  # tensorf.py:112
        kind = 1 if getattr(opt, "soft_edge_loss", False) else 2
        assert train_edge_masks is not None
    fe = float(getattr(opt, "edge_loss_factor", 1.0))
    fn = float(getattr(opt, "non_edge_loss_factor", 1.0))
    if mode not in ["train", "test-optim"]:
        ray_idx = None                                                         # tensorf.py:101-102
    return RenderLoss.apply(rgb, images, train_edge_masks if kind else None, _i32(ray_idx), _i32(view_idx), kind,
                            fe, fn)
