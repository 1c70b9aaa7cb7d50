"""GPU parity of the per-step sweeps (regularisers, fused Adam) and the field
maintenance ops (alpha-mask update, shrink, upsample; 8f-3) against golden vectors from the live reference
(tests/golden/make_golden_field.py) and the CPU oracle (oracle/field_oracle.py). Tolerances: fp32 sums in a
different order -> <= 1e-5 relative; data movement (shrink) and the mask bits are bit-exact."""
import pytest
import torch

import joint_tensorf_b200 as jt
from common import load_golden, rel_err, vo
from gpu_common import module_from_golden
from joint_tensorf_b200.sweeps import FusedAdam
from oracle import field_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FACTOR_PREFIXES = ("density_plane", "density_line", "app_plane", "app_line")


def _module(name):
    g = load_golden(name)
    g = dict(g, case=dict(g["case"]))
    return g, module_from_golden(dict(g, mask_volume=None), DEV)


def test_regularisers_match_reference_golden_autograd_and_fused():
    g, m = _module("field_reg_adam")
    reg = type("TV", (), {"TVLoss_weight": 1})()
    l1, tvd, tva = m.density_L1(), m.TV_loss_density(reg), m.TV_loss_app(reg)
    for got, ref in zip((l1, tvd, tva), g["values"]):
        assert abs(float(got) - ref) <= 1e-5 * abs(ref), (float(got), ref)
    w = g["weights"]
    (w[0] * l1 + w[1] * tvd + w[2] * tva).backward()
    for k, ref in g["reg_grads"].items():
        p = dict(m.named_parameters())[k]
        assert rel_err(p.grad.cpu(), ref) <= 1e-5, k
    assert m.basis_mat.weight.grad is None and m.app_line[0].grad is None
    # fused path: the same gradients, ACCUMULATED in place (two calls -> twice the gradient), values from the same call
    auto = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    for p in m.parameters():
        p.grad = None
    vals = m.regularize_(*w)
    m.regularize_(*w, values=False)
    assert torch.allclose(vals.cpu(), torch.tensor(g["values"]), rtol=1e-5, atol=0)
    for k, ref in auto.items():
        p = dict(m.named_parameters())[k]
        assert rel_err(p.grad, 2.0 * ref) <= 1e-5, k
    assert float(m.app_line[1].grad.abs().max()) == 0.0       # no term touches the appearance lines


def test_regulariser_weights_and_cache():
    g, m = _module("field_reg_adam")
    a = float(m.density_L1())
    with torch.no_grad():
        m.density_line[0].mul_(2.0)               # version bump -> the cached node is stale
    b = float(m.density_L1())
    assert b > a
    ref = fo.density_l1({k: v.detach().cpu() for k, v in m.state_dict().items()})
    assert abs(b - float(ref)) <= 1e-5 * float(ref)
    # a TVLoss weight other than 1 scales the TV terms (tensorBase.py:17-19,38)
    reg = type("TV", (), {"TVLoss_weight": 3.0})()
    assert abs(float(m.TV_loss_app(reg)) - 3.0 * g["values"][2]) <= 1e-5 * 3.0 * g["values"][2]
    # zero host weights: nothing is swept, gradients untouched
    for p in m.parameters():
        p.grad = None
    m.regularize_(0.0, 0.0, 0.0, values=False)
    assert all(float(p.grad.abs().max()) == 0.0 for k, p in m.named_parameters()
               if k.split(".")[0] in FACTOR_PREFIXES)


def test_regularisers_at_full_plane_size_against_oracle():
    """300 x 300 x 16/48 planes (cfg2 factor sizes): values and gradients against the oracle."""
    gen = torch.Generator().manual_seed(2)
    params = {}
    for pre, c in (("density", 16), ("app", 48)):
        for i in range(3):
            params[f"{pre}_plane.{i}"] = torch.randn((1, c, 300, 300), generator=gen) * 0.1
            params[f"{pre}_line.{i}"] = torch.randn((1, c, 300, 1), generator=gen) * 0.1
    m = jt.B200_VMSplit(torch.tensor([[-1.5] * 3, [1.5] * 3]), [300] * 3, DEV, density_n_comp=[16] * 3,
                        appearance_n_comp=[48] * 3, app_dim=27, shadingMode="SH")
    m.load_state_dict(params, strict=False)
    cpu = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref = (fo.density_l1(cpu), fo.tv_loss_density(cpu), fo.tv_loss_app(cpu))
    (0.5 * ref[0] + 2.0 * ref[1] + 3.0 * ref[2]).backward()
    vals = m.regularize_(0.5, 2.0, 3.0)
    for got, r in zip(vals.tolist(), ref):
        assert abs(got - float(r)) <= 2e-5 * float(r)
    for k, p in m.named_parameters():
        if k in cpu and cpu[k].grad is not None:
            assert rel_err(p.grad.cpu(), cpu[k].grad) <= 1e-5, k


def test_fused_adam_matches_reference_golden():
    g, m = _module("field_reg_adam")
    opt = FusedAdam(m.get_optparam_groups(g["lr_index"], g["lr_basis"]), betas=(0.9, 0.99))
    named = dict(m.named_parameters())
    assert sorted(named) == sorted(g["param_names"])
    for grads in g["step_grads"]:
        for k, p in named.items():
            p.grad = torch.empty_like(p, memory_format=torch.preserve_format).copy_(grads[k].to(DEV))
        opt.step()
        for group in opt.param_groups:            # model/tensorf.py:433-434
            group["lr"] = group["lr"] * g["decay"]
    assert [grp["lr"] for grp in opt.param_groups] == g["final_lrs"]
    for k, p in named.items():
        assert rel_err(p.detach().cpu(), g["final"][k]) <= 2e-6, k
        st = opt.state[p]
        assert float(st["step"]) == g["adam_state"][k]["step"]
        assert rel_err(st["exp_avg"].cpu(), g["adam_state"][k]["exp_avg"]) <= 2e-6, k
        assert rel_err(st["exp_avg_sq"].cpu(), g["adam_state"][k]["exp_avg_sq"]) <= 2e-6, k
    # state_dict has torch.optim.Adam's layout (the reference saves / restores it)
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and len(sd["param_groups"]) == 6
    ref_opt = torch.optim.Adam(m.get_optparam_groups(g["lr_index"], g["lr_basis"]), betas=(0.9, 0.99))
    ref_opt.load_state_dict(sd)


@pytest.mark.parametrize("zero_grad", [False, True])
def test_fused_adam_ragged_sizes_unaligned_and_options(zero_grad):
    gen = torch.Generator().manual_seed(4)
    sizes = [1, 3, 5, 1027, 4096, 70001]
    store = torch.zeros(sum(sizes) + 16, device=DEV)
    ps, off = [], 1                                # odd offsets: 4-byte aligned only
    for n in sizes:
        ps.append(torch.nn.Parameter(store[off:off + n]))
        off += n
    ps.append(torch.nn.Parameter(torch.zeros((0,), device=DEV)))          # empty tensor
    cpu_p = [torch.randn((p.numel(),), generator=gen) for p in ps]
    with torch.no_grad():
        for p, c in zip(ps, cpu_p):
            p.copy_(c)
    opt = FusedAdam([{"params": ps[:3], "lr": 0.05}, {"params": ps[3:], "lr": 0.002}], betas=(0.9, 0.99),
                    zero_grad_in_step=zero_grad)
    opt.grad_scale = 0.25
    ms = [torch.zeros_like(c) for c in cpu_p]
    vs = [torch.zeros_like(c) for c in cpu_p]
    for it in range(4):
        gs = [torch.randn((p.numel(),), generator=gen) * 0.3 for p in ps]
        for p, gg in zip(ps, gs):
            if zero_grad and it > 0:
                assert float(p.grad.abs().sum()) == 0.0
                p.grad.copy_(gg)
            else:
                p.grad = gg.to(DEV)
        opt.step()
        for i, (c, gg) in enumerate(zip(cpu_p, gs)):
            fo.adam_step(c, gg * 0.25, ms[i], vs[i], it + 1, 0.05 if i < 3 else 0.002)
    for p, c in zip(ps, cpu_p):
        if c.numel():
            assert rel_err(p.detach().cpu(), c) <= 2e-6
    assert float(store[0]) == 0.0 and float(store[off:].abs().sum()) == 0.0          # no out-of-bounds writes


def test_alpha_mask_update_shrink_upsample_match_reference_golden():
    g, m = _module("field_maintenance")
    m.kernel_density, m.c2f_mode = None, None
    mg = list(g["mask_grid"])
    alpha, dense_xyz = m.getDenseAlpha(mg)
    assert tuple(alpha.shape) == tuple(mg) and tuple(dense_xyz.shape) == (*mg, 3)
    assert (alpha.cpu() - g["dense_alpha"]).abs().max() <= 1e-6
    assert torch.equal(dense_xyz.cpu(), fo.dense_grid_points(g["aabb"], mg))
    new_aabb = m.updateAlphaMask(mg)
    vol = m.alphaMask.alpha_volume[0, 0].cpu()
    # the mask is a threshold of a float: a voxel may differ only if its pooled alpha is within rounding of the threshold
    pooled = torch.nn.functional.max_pool3d(g["dense_alpha"].clamp(0, 1).transpose(0, 2).contiguous()[None, None], 5, 1, 2)[0, 0]
    diff = vol != g["mask_volume"]
    assert int(diff.sum()) == 0 or bool(((pooled[diff] - g["alpha_thres"]).abs() <= 1e-6 * g["alpha_thres"]).all())
    if int(diff.sum()) == 0:
        assert torch.equal(new_aabb.cpu(), g["new_aabb"])
    # the bit-packed copy the ray marcher reads agrees with the float volume
    ref_mask = jt.AlphaGridMask(DEV, g["aabb"].to(DEV), m.alphaMask.alpha_volume[0, 0].clone())
    assert torch.equal(ref_mask.bits, m.alphaMask.bits)
    got = m.compute_alpha(g["pts"].to(DEV), m.stepSize)
    assert (got.cpu() - g["alpha_pts"]).abs().max() <= 1e-6
    assert bool((g["alpha_pts"] == 0).any()) and bool((g["alpha_pts"] > 0).any())      # both mask outcomes exercised
    m.shrink(g["new_aabb"].to(DEV))
    assert m.gridSize.tolist() == g["shrunk_grid"] and m.nSamples == g["shrunk_nsamples"]
    assert torch.equal(m.aabb.cpu(), g["shrunk_aabb"])
    assert abs(float(m.stepSize) - g["shrunk_step"]) == 0.0
    for k, ref in g["shrunk"].items():
        assert torch.equal(m.state_dict()[k].cpu(), ref), k
    m.upsample_volume_grid(g["up_target"])
    for k, ref in g["upsampled"].items():
        got = m.state_dict()[k]
        assert got.shape == ref.shape, k
        assert rel_err(got.cpu(), ref) <= 1e-6, k
        if got.dim() == 4:
            assert got.permute(0, 2, 3, 1).is_contiguous(), k
    assert abs(float(m.stepSize) - g["up_step"]) == 0.0


def test_update_alpha_mask_raises_when_empty_and_respects_previous_mask():
    g, m = _module("field_maintenance")
    m.kernel_density, m.c2f_mode = None, None
    m.alphaMask_thres = 2.0                       # alpha <= 1: nothing passes
    with pytest.raises(RuntimeError):
        m.updateAlphaMask((12, 12, 12))
    m.alphaMask_thres = 1e-4
    m.updateAlphaMask(list(g["mask_grid"]))
    first = m.alphaMask.alpha_volume.clone()
    # second update sees the first mask (batBase.py:29-31): masked-out points have alpha 0
    a2 = m._dense_alpha_zyx(list(g["mask_grid"]))
    outside = torch.nn.functional.max_pool3d(first, 3, 1, 1)[0, 0] == 0      # not even a neighbour kept
    assert float(a2[outside].abs().max()) == 0.0


@pytest.mark.parametrize("shape,target", [((1, 8, 5, 7), (11, 3)), ((1, 4, 9, 1), (17, 1)), ((1, 48, 64, 64), (64, 64)),
                                          ((1, 16, 33, 20), (1, 1)), ((1, 12, 1, 6), (4, 9))])
def test_resize_bilinear_against_aten(shape, target):
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    ref = torch.nn.functional.interpolate(x, size=target, mode="bilinear", align_corners=True)
    got = jt.ops.resize_bilinear_cl(x.to(DEV).contiguous(memory_format=torch.channels_last), *target)
    assert got.shape == ref.shape
    assert rel_err(got.cpu(), ref) <= 1e-6
