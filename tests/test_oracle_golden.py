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
