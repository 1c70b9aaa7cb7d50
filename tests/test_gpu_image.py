"""GPU parity of the 2-D supervision pre-processing and the render loss (csrc/image_prep.cu through the C ABI,
SURVEY.md section 8f-4) against the live-reference golden vectors (tests/golden/image_prep.pt) and against
oracle/image_oracle.py on larger seeded images.

Tolerances: blurred images <= 1e-5 max-abs (values in [0, ~2]; 201-tap fp32 sums in a different order), soft masks
<= 1e-5, hard masks bit-exact except pixels whose gradient magnitude lies within 1e-5 (relative) of the threshold
(the per-image mean is summed in a different order), loss <= 1e-6 relative, d loss / d rgb <= 1e-5 of the largest
entry."""
import pytest
import torch

from common import load_golden, rel_err
from oracle import image_oracle as io
from test_oracle_golden import _image_loss_kind

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _opt(o, h, w):
    from joint_tensorf_b200.options import Namespace
    d = {k: v for k, v in o.items() if k != "it"}
    d.update(H=h, W=w)
    return Namespace(d)


def _check_hard_mask(got, images_cpu, thresh):
    gg = io.sobel_magnitude(images_cpu)
    thr = gg.mean(dim=1, keepdim=True) * thresh
    ref = (gg > thr).to(torch.uint8)
    bad = got.cpu() != ref
    near = (gg - thr).abs() <= 1e-5 * thr
    assert not bool((bad & ~near).any()), int((bad & ~near).sum())
    assert int(bad.sum()) <= max(2, ref.numel() // 10000)


@pytest.mark.parametrize("name", ["blender", "box_soft"])
def test_process_gt_images_and_edge_masks_on_reference_golden(name):
    from joint_tensorf_b200 import supervision as sv
    g = load_golden("image_prep")
    o, ref = g["opts"][name], g["prep"][name]
    images = g["images"].to(DEV)
    opt = _opt(o, *images.shape[-2:])
    blurred = sv.process_GT_images(opt, images, o["it"])
    assert sorted(blurred) == sorted(ref["blurred"])
    assert blurred[0.0] is images
    for sc, img in blurred.items():
        assert (img.cpu() - ref["blurred"][sc]).abs().max() <= 1e-5, (name, sc)
    # masks are computed from OUR blurred images (error propagates through the Sobel filter: 16 * 1e-5 at most)
    masks = sv.get_edge_mask(opt, blurred)
    for sc, m in masks.items():
        r = ref["edge"][sc]
        assert m.dtype == r.dtype and m.shape == r.shape
        if m.dtype == torch.uint8:
            _check_hard_mask(m, blurred[sc].cpu(), o["hard_edge_mask_mean_thresh"])
            assert float((m.cpu() != r).float().mean()) <= 2e-3
        else:
            assert (m.cpu() - r).abs().max() <= 2e-5, (name, sc)


def test_render_loss_on_reference_golden():
    from joint_tensorf_b200 import supervision as sv
    from joint_tensorf_b200.options import Namespace
    g = load_golden("image_prep")
    images = g["images"].to(DEV)
    h, w = images.shape[-2:]
    masks = {k: g["prep"][k]["edge"][1.0].to(DEV) for k in g["opts"]}
    ray_idx = g["ray_idx"].to(DEV)
    for l in g["losses"]:
        opt = Namespace(dict(l["flags"], edge_loss_factor=1.5, non_edge_loss_factor=0.5, H=h, W=w))
        rgb = g["rgb"].clone()
        if l["nan"]:
            rgb[1, 7, 2] = float("nan")
        rgb = rgb.to(DEV).requires_grad_(True)
        val = sv.render_loss(opt, rgb, images, ray_idx, masks.get(l["mask"]), it=l["it"], mode="train")
        assert abs(float(val) - float(l["loss"])) <= 1e-6 * max(1.0, abs(float(l["loss"]))), l["tag"]
        (val * l["upstream"]).backward()
        ref = l["d_rgb"]
        got = rgb.grad.cpu()
        assert torch.equal(torch.isnan(got), torch.isnan(ref)), l["tag"]
        ok = ~torch.isnan(ref)
        assert (got[ok] - ref[ok]).abs().max() <= 1e-5 * ref[ok].abs().max(), l["tag"]
    v = g["loss_val"]
    opt = Namespace(edge_loss_factor=1.5, non_edge_loss_factor=0.5, H=h, W=w)
    val = sv.render_loss(opt, v["rgb"].to(DEV), images, ray_idx, None, it=0, mode="val")
    assert abs(float(val) - float(v["loss"])) <= 1e-6


@pytest.mark.parametrize("shape,ntaps", [((2, 3, 200, 304), 201), ((1, 3, 131, 77), 65), ((2, 3, 37, 530), 257),
                                         ((1, 3, 9, 11), 201), ((1, 3, 64, 64), 1)])
def test_image_blur_against_oracle(shape, ntaps):
    """Ragged sizes (not multiples of the 128 x 32 tile), images smaller than the halo, 1 and 257 taps."""
    from joint_tensorf_b200 import supervision as sv
    gen = torch.Generator().manual_seed(5)
    images = torch.rand(shape, generator=gen)
    taps = io.gaussian_kernel(torch.tensor(3.7, dtype=torch.float64), ntaps) if ntaps > 1 else torch.tensor([0.6])
    taps = taps * (1 + 0.1 * torch.rand(taps.shape, generator=gen))      # asymmetric: catches a flipped stencil
    ref = io.blur_images(images, taps)
    got = sv.image_blur(images.to(DEV), taps).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("soft", [False, True])
def test_edge_mask_against_oracle_800(soft):
    """Blender-sized images (800 x 800): Sobel magnitude, per-image max / mean and the mask."""
    from joint_tensorf_b200 import supervision as sv
    gen = torch.Generator().manual_seed(6)
    images = torch.rand((2, 3, 800, 800), generator=gen)
    images = io.blur_images(images, io.gaussian_kernel(torch.tensor(2.0, dtype=torch.float64), 15))
    mask, gg, stats = sv.edge_mask(images.to(DEV), soft=soft, thresh=1.25, return_stats=True)
    ref_gg = io.sobel_magnitude(images)
    assert (gg.cpu() - ref_gg).abs().max() <= 1e-5
    assert rel_err(stats[:, 0].cpu(), ref_gg.max(dim=1)[0]) <= 1e-6
    assert rel_err(stats[:, 1].cpu(), ref_gg.mean(dim=1)) <= 1e-5
    if soft:
        assert (mask.cpu() - io.edge_mask(images, True)).abs().max() <= 1e-5
    else:
        _check_hard_mask(mask, images, 1.25)


def test_render_loss_full_size_properties():
    """cfg2-sized batch (32 views x 128 rays of 800 x 800 images): the fused gather + loss equals the oracle, the
    backward is linear in the upstream gradient, and the hard-edge loss with an all-ones mask equals
    edge_factor * plain MSE (the (1-m) branch is the nanmean of zeros)."""
    from joint_tensorf_b200 import supervision as sv
    from joint_tensorf_b200.options import Namespace
    gen = torch.Generator().manual_seed(7)
    b, n, h, w = 32, 128, 800, 800
    cache = torch.rand((40, 3, h, w), generator=gen)
    view_idx = torch.randperm(40, generator=gen)[:b]
    ray_idx = torch.randperm(h * w, generator=gen)[:n]
    mask = (torch.rand((40, h * w), generator=gen) > 0.8).to(torch.uint8)
    rgb = torch.rand((b, n, 3), generator=gen)
    opt = Namespace(edge_mask_on_render_loss=True, edge_mask_before_iter=10, edge_loss_factor=1.5, non_edge_loss_factor=0.5,
                    H=h, W=w)
    ref = io.render_loss(rgb, cache[view_idx], ray_idx, mask[view_idx], 2, 1.5, 0.5)
    r = rgb.to(DEV).requires_grad_(True)
    args = (cache.to(DEV), ray_idx.to(DEV), mask.to(DEV))
    val = sv.render_loss(opt, r, *args, it=0, mode="train", view_idx=view_idx.to(DEV))
    assert abs(float(val) - float(ref)) <= 1e-6
    (g1,) = torch.autograd.grad(val, r, retain_graph=True)
    (g3,) = torch.autograd.grad(val * 3.0, r)
    assert (g3 - 3.0 * g1).abs().max() <= 1e-6 * g1.abs().max()
    rr = rgb.clone().requires_grad_(True)
    (gref,) = torch.autograd.grad(io.render_loss(rr, cache[view_idx], ray_idx, mask[view_idx], 2, 1.5, 0.5), rr)
    assert (g1.cpu() - gref).abs().max() <= 1e-5 * gref.abs().max()
    ones = torch.ones_like(args[2])
    plain = sv.render_loss(Namespace(H=h, W=w), r, args[0], args[1], None, it=0, mode="train", view_idx=view_idx.to(DEV))
    hard = sv.render_loss(opt, r, args[0], args[1], ones, it=0, mode="train", view_idx=view_idx.to(DEV))
    assert abs(float(hard) - 1.5 * float(plain)) <= 1e-6


def test_supervision_rejects_cpu_tensors():
    from joint_tensorf_b200 import _lib, supervision as sv
    with pytest.raises(_lib.JtError):
        sv.image_blur(torch.rand(1, 3, 8, 8), [1.0])
    with pytest.raises(_lib.JtError):
        sv.edge_mask(torch.rand(1, 3, 8, 8))
