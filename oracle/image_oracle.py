"""CPU oracle for the 2-D supervision pre-processing and the render loss (TEST INFRASTRUCTURE).

Checker only (tests/, smoke(), bench.py CPU legs); the product never imports it. Restates, as plain functions over
torch CPU tensors with the same ATen operators the reference calls, SURVEY.md section 8f-4:
    Model.process_GT_images (model/nerf.py:57-113), Model.get_edge_mask (model/nerf.py:116-149) and the render term
    of Graph.compute_loss (model/tensorf.py:99-124, MSE_loss base.py:259-261).

Parity pin: `tests/golden/make_golden_image.py` imports the LIVE reference (`model.nerf.Model`, `model.tensorf.Graph`,
logging packages mocked), calls those three methods on seeded images / renders and stores inputs + outputs in
`tests/golden/image_*.pt`; `tests/test_oracle_golden.py` checks this file against them.

Third-party arithmetic: PyTorch ATen (`F.pad(mode="replicate")`, `F.conv1d`, `F.conv2d`, `nanmean`).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def interp_schedule(x, schedule, left=0, right=1):
    """util.py:217-225."""
    xs = np.linspace(left, right, len(schedule))
    return np.interp(x, xs, schedule)


def gaussian_kernel(t, kernel_size):
    """kernels.get_gaussian_kernel (kernels.py:16-22); t: 0-d tensor as in the reference."""
    ns = torch.arange(-(kernel_size // 2), kernel_size // 2 + 1, dtype=torch.float32)
    exponent = -0.5 * (ns / max(t, 0.0001)) * (ns / max(t, 0.0001))
    kernel = 1 / (max(t, 0.0001) * math.sqrt(2 * math.pi)) * torch.exp(exponent)
    return torch.clamp(kernel, max=1.0)


def average_kernel(t, kernel_size):
    """kernels.get_average_kernel (kernels.py:24-41)."""
    if kernel_size % 2 == 0:
        kernel_size += 1
    if isinstance(t, torch.Tensor):
        t = t.item()
    t0 = min(math.floor(t), kernel_size // 2)
    k0 = torch.zeros(kernel_size)
    k0[kernel_size // 2 - t0:kernel_size // 2 + t0 + 1] = 1 / (t0 * 2 + 1)
    t1 = min(math.ceil(t), kernel_size // 2)
    k1 = torch.zeros(kernel_size)
    k1[kernel_size // 2 - t1:kernel_size // 2 + t1 + 1] = 1 / (t1 * 2 + 1)
    return (t % 1.0) * k1 + (1 - t % 1.0) * k0


def scales(opt):
    """nerf.py:63-66."""
    if opt["c2f_alternate_2D_mode"] == "sample":
        return list(opt["c2f_alternate_2D_scale_pool"])
    return [0.0, 1.0]


def blur_images(images, kernel_1d):
    """The separable convolution of nerf.py:98-110. images [B,3,H,W], kernel_1d [K]."""
    b, _, h, w = images.shape
    k = kernel_1d.float().expand(1, 1, -1)
    pad = (k.shape[-1] // 2, k.shape[-1] // 2)
    x = images.reshape(b * 3, h, w)
    x = F.pad(x, pad, mode="replicate")
    x = F.conv1d(x, k.expand(h, 1, -1), bias=None, stride=1, padding=0, dilation=1, groups=h)
    x = x.permute(0, 2, 1)
    x = F.pad(x, pad, mode="replicate")
    x = F.conv1d(x, k.expand(w, 1, -1), bias=None, stride=1, padding=0, dilation=1, groups=w)
    return x.permute(0, 2, 1).reshape(b, 3, h, w).contiguous()


def process_gt_images(opt, images, it):
    """nerf.py:57-113 -> {scale: images}; opt is a plain dict of the YAML fields."""
    h, w = images.shape[-2:]
    out = {}
    for sc in scales(opt):
        blur_param = torch.tensor(interp_schedule(float(it / opt["max_iter"]), opt["blur_2d_c2f_schedule"]))
        blur_param *= sc
        width = blur_param * (w + h) / 2
        if opt["blur_2d_mode"] == "uniform-gaussian":
            k = gaussian_kernel(width, opt["blur_2d_c2f_kernel_size"])
        elif opt["blur_2d_mode"] == "uniform-average":
            k = average_kernel(width, opt["blur_2d_c2f_kernel_size"])
        else:
            raise NotImplementedError
        out[sc] = images if width < 0.01 else blur_images(images, k)
    return out


def sobel_magnitude(images):
    """nerf.py:124-139: [B,3,H,W] -> GG [B, H*W]."""
    b, _, h, w = images.shape
    kx = torch.tensor([[1, 0, -1], [2, 0, -2], [1, 0, -1]], dtype=torch.float32)[None, None].expand(1, 3, -1, -1)
    ky = torch.tensor([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], dtype=torch.float32)[None, None].expand(1, 3, -1, -1)
    x = F.pad(images, (1, 1, 1, 1), mode="replicate")
    gx = F.conv2d(x, kx, padding=0)
    gy = F.conv2d(x, ky, padding=0)
    return torch.sqrt(gx ** 2 + gy ** 2).view(b, h * w)


def edge_mask(images, soft=False, thresh=1.25):
    """nerf.py:140-148."""
    gg = sobel_magnitude(images)
    if soft:
        return gg / gg.max(dim=1, keepdim=True)[0]
    return (gg > gg.mean(dim=1, keepdim=True) * thresh).to(torch.uint8)


def mse(pred, label):
    """base.py:259-261."""
    return ((pred.contiguous() - label) ** 2).nanmean()


def render_loss(rgb, images, ray_idx, edge_masks, kind, fe, fn):
    """tensorf.py:99-124. rgb [B,n,3]; images [B,3,H,W]; kind 0 plain / 1 soft-edge / 2 hard-edge loss."""
    b = rgb.shape[0]
    image = images.reshape(b, 3, -1).permute(0, 2, 1)
    if ray_idx is not None:
        image = image[:, ray_idx]
    if kind == 0:
        return mse(rgb, image)
    m = edge_masks[:, ray_idx].view(b, len(ray_idx), 1)
    if kind == 1:
        m = m.expand(-1, -1, 3) * fe + fn
        return mse(rgb * m, image * m)
    m = m.expand(-1, -1, 3)
    return fe * mse(rgb * m, image * m) + fn * mse(rgb * (1 - m), image * (1 - m))
