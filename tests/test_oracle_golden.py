"""Pin the CPU oracle (oracle/vm_oracle.py) against golden vectors produced by the
LIVE reference (tests/golden/make_golden.py). Valid masks must be bit-exact;
floating outputs and gradients agree to <= 2e-6 (same ATen kernels, same order)."""
import pytest
import torch

from common import (camera_golden_names, field_from_golden, golden_names, golden_valid, load_golden, rel_err,
                    render_kwargs_from_golden, vo)
from oracle import camera_oracle as co

torch.set_num_threads(max(1, torch.get_num_threads()))


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    field = field_from_golden(g)
    o = g["rays_o"].clone().requires_grad_(True)
    d = g["rays_d"].clone().requires_grad_(True)
    rgb, depth, acc, aux = vo.render(field, o, d, detail=True, **render_kwargs_from_golden(g))
    # bit-exact class
    assert torch.equal(aux["valid"], golden_valid(g))
    assert int(aux["valid"].sum()) == g["valid_count"]
    assert torch.equal(aux["z"].expand_as(g["z"]), g["z"])
    # floating class
    assert (rgb - g["rgb"]).abs().max() <= 2e-6
    assert (acc - g["acc"]).abs().max() <= 2e-6
    assert (depth - g["depth"]).abs().max() <= 1e-5
    loss = (rgb * g["w_rgb"]).sum() + (acc * g["w_acc"]).sum()
    loss.backward()
    assert rel_err(o.grad, g["d_rays_o"]) <= 1e-5
    assert rel_err(d.grad, g["d_rays_d"]) <= 1e-5
    for k, ref in g["grads"].items():
        assert rel_err(field.params[k].grad, ref) <= 1e-5, k
    for k, (s, sabs, mx) in g["grad_sums"].items():
        got = field.params[k].grad
        assert abs(float(got.double().abs().sum()) - sabs) <= 1e-5 * max(sabs, 1e-12), k


@pytest.mark.parametrize("name", camera_golden_names())
def test_camera_oracle_matches_reference_golden(name):
    """oracle/camera_oracle.py against the live reference's camera.py (make_golden_camera.py)."""
    g = load_golden(name)
    c = g["case"]
    se3 = g["se3"].clone().requires_grad_(True)
    assert (co.refined_pose(se3, g["pose"]) - g["refined_pose"]).abs().max() <= 1e-6
    center, ray = co.rays_of_step(se3, g["pose"], g["intr_inv"], c["H"], c["W"], g["ray_idx"], intr=g["intr"],
                                  ndc=c["ndc"], near=1.0, center_shift=c.get("shift", True),
                                  detach_shift=c.get("detach", False))
    assert rel_err(center, g["center"]) <= 1e-6
    assert rel_err(ray, g["ray"]) <= 1e-6
    loss = (center * g["g_center"]).sum() + (ray * g["g_ray"]).sum()
    loss.backward()
    assert rel_err(se3.grad, g["d_se3"]) <= 1e-5


def test_field_oracle_regularisers_and_adam_match_reference_golden():
    """oracle/field_oracle.py against the live reference's density_L1 / TV_loss_* / torch.optim.Adam run
    (tests/golden/make_golden_field.py)."""
    from oracle import field_oracle as fo
    g = load_golden("field_reg_adam")
    params = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    l1, tvd, tva = fo.density_l1(params), fo.tv_loss_density(params), fo.tv_loss_app(params)
    for got, ref in zip((l1, tvd, tva), g["values"]):
        assert abs(float(got) - ref) <= 1e-6 * abs(ref)
    w = g["weights"]
    (w[0] * l1 + w[1] * tvd + w[2] * tva).backward()
    for k, ref in g["reg_grads"].items():
        assert rel_err(params[k].grad, ref) <= 1e-6, k
    # three Adam steps with the reference's groups (factors lr_index, basis + head lr_basis) and lr decay
    p = {k: v.clone() for k, v in g["state_dict"].items()}
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items()}
    lr = {k: (g["lr_index"] if k.split(".")[0] in ("density_plane", "density_line", "app_plane", "app_line")
              else g["lr_basis"]) for k in g["param_names"]}
    for it, grads in enumerate(g["step_grads"]):
        for k in g["param_names"]:
            fo.adam_step(p[k], grads[k], m[k], v2[k], it + 1, lr[k] * g["decay"] ** it)
    for k in g["param_names"]:
        assert rel_err(p[k], g["final"][k]) <= 1e-6, k
        assert rel_err(m[k], g["adam_state"][k]["exp_avg"]) <= 1e-6, k
        assert rel_err(v2[k], g["adam_state"][k]["exp_avg_sq"]) <= 1e-6, k


def test_field_oracle_maintenance_matches_reference_golden():
    """updateAlphaMask / shrink / upsample restatements against the live reference run."""
    from oracle import field_oracle as fo
    g = load_golden("field_maintenance")
    grid = list(g["case"]["grid"])
    pts = fo.dense_grid_points(g["aabb"], g["mask_grid"])
    field = field_from_golden(dict(g, mask_volume=None), requires_grad=False)
    fc = vo.grid_constants(field.aabb, field.grid, field.step_ratio)
    with torch.no_grad():
        alpha = vo.compute_alpha(field, pts.view(-1, 3), fc["step"]).view(g["mask_grid"])
    assert (alpha - g["dense_alpha"]).abs().max() <= 1e-7
    vol, new_aabb = fo.alpha_mask_from_dense(g["dense_alpha"], pts, g["alpha_thres"])
    assert torch.equal(vol, g["mask_volume"])
    assert torch.equal(new_aabb, g["new_aabb"])
    t_l, b_r = fo.shrink_bounds(g["aabb"], fc["units"], grid, g["new_aabb"])
    assert (b_r - t_l).tolist() == g["shrunk_grid"]
    shr = fo.shrink_factors(g["state_dict"], t_l, b_r)
    for k, ref in g["shrunk"].items():
        assert torch.equal(shr[k], ref), k
    assert torch.allclose(fo.shrink_corrected_aabb(g["aabb"], grid, t_l, b_r), g["shrunk_aabb"], atol=0, rtol=0)
    up = fo.upsample_factors(g["shrunk"], g["up_target"])
    for k, ref in g["upsampled"].items():
        assert torch.equal(up[k], ref), k


def _image_loss_kind(l):
    """The branch of model/tensorf.py:104-124 a golden loss case takes."""
    f = l["flags"]
    on = f.get("edge_mask_on_render_loss", False) and (l["it"] % 2 == 0 if f.get("alternate_edge_loss", False) else True)
    if not (on and l["it"] < f["edge_mask_before_iter"]):
        return 0
    return 1 if f.get("soft_edge_loss", False) else 2


def test_image_oracle_matches_reference_golden():
    """oracle/image_oracle.py against the live reference's process_GT_images / get_edge_mask / compute_loss
    (tests/golden/make_golden_image.py). Same ATen operators in the same order -> exact or 1e-7."""
    from oracle import image_oracle as io
    g = load_golden("image_prep")
    images = g["images"]
    masks = {}
    for name, o in g["opts"].items():
        ref = g["prep"][name]
        blurred = io.process_gt_images(o, images, o["it"])
        assert sorted(blurred) == sorted(ref["blurred"])
        for sc, img in blurred.items():
            assert (img - ref["blurred"][sc]).abs().max() <= 1e-6, (name, sc)
            m = io.edge_mask(img, o.get("soft_edge_mask", False), o.get("hard_edge_mask_mean_thresh", 1.25))
            assert m.dtype == ref["edge"][sc].dtype
            if m.dtype == torch.uint8:
                assert torch.equal(m, ref["edge"][sc]), (name, sc)
            else:
                assert (m - ref["edge"][sc]).abs().max() <= 1e-6, (name, sc)
        assert blurred[0.0] is images                       # width < 0.01: the raw images (nerf.py:95-97)
        masks[name] = ref["edge"][1.0]
    for l in g["losses"]:
        rgb = g["rgb"].clone()
        if l["nan"]:
            rgb[1, 7, 2] = float("nan")
        rgb.requires_grad_(True)
        kind = _image_loss_kind(l)
        val = io.render_loss(rgb, images, g["ray_idx"], masks.get(l["mask"]), kind, 1.5, 0.5)
        assert abs(float(val) - float(l["loss"])) <= 1e-7, l["tag"]
        (d,) = torch.autograd.grad(val * l["upstream"], rgb)
        assert torch.allclose(d, l["d_rgb"], rtol=1e-6, atol=1e-9, equal_nan=True), l["tag"]
    v = g["loss_val"]
    assert abs(float(io.render_loss(v["rgb"], images, None, None, 0, 1.5, 0.5)) - float(v["loss"])) <= 1e-7


@pytest.mark.parametrize("name", ["cubic_mlp", "sh_vm48", "ndc_weakview_blur"])
def test_float64_yardstick_mode_tracks_the_reference(name):
    """vo.render(exact=True): samples placed in fp32 exactly as the reference does (same valid mask, same appearance
    set), everything downstream in float64. The GPU tests use it to tell the reference's own fp32 rounding from a real
    discrepancy on ill-conditioned full-size gradients; here it must agree with the reference's fp32 outputs to fp32
    rounding on the well-conditioned golden cases."""
    g = load_golden(name)
    f64 = field_from_golden(g)
    f64.params = {k: v.detach().double().requires_grad_(True) for k, v in f64.params.items()}
    o, d = g["rays_o"].clone().requires_grad_(True), g["rays_d"].clone().requires_grad_(True)
    rgb, depth, acc, det = vo.render(f64, o, d, exact=True, detail=True, **render_kwargs_from_golden(g))
    assert rgb.dtype == torch.float64
    assert torch.equal(det["valid"], golden_valid(g))
    assert (rgb.float() - g["rgb"]).abs().max() <= 2e-6 and (acc.float() - g["acc"]).abs().max() <= 2e-6
    ((rgb * g["w_rgb"]).sum() + (acc * g["w_acc"]).sum()).backward()
    assert rel_err(o.grad, g["d_rays_o"]) <= 1e-4
    for k, ref in g["grads"].items():
        assert rel_err(f64.params[k].grad.float(), ref) <= 1e-4, k
