"""Host-side logic of the drop-in module, checked on CPU (no kernels run)."""
import math

import pytest
import torch

import joint_tensorf_b200 as jt
from common import load_golden, vo
from gpu_common import module_from_golden
from joint_tensorf_b200 import vmsplit


def small(grid=(20, 24, 28), shading="MLP_Fea", **kw):
    return jt.B200_VMSplit(torch.tensor([[-1.5, -1.67, -2.0], [1.5, 1.67, 1.0]]), list(grid), "cpu",
                           density_n_comp=[8] * 3, appearance_n_comp=[12] * 3, app_dim=27, shadingMode=shading,
                           featureC=64, pos_pe=2, view_pe=2, fea_pe=2, step_ratio=0.5, near_far=[2.0, 6.0], **kw)


def test_state_dict_matches_reference_names_and_shapes():
    for name in ("cubic_mlp", "ndc_weakview", "sh"):
        g = load_golden(name)
        m = module_from_golden(g, device="cpu")
        sd = m.state_dict()
        assert set(sd) == set(g["state_dict"]), (set(sd) ^ set(g["state_dict"]))
        for k, v in g["state_dict"].items():
            assert tuple(sd[k].shape) == tuple(v.shape), k
            assert torch.equal(sd[k], v), k


def test_factors_are_channel_last_in_memory():
    m = small()
    for plist in (m.density_plane, m.app_plane, m.density_line, m.app_line):
        for p in plist:
            assert p.permute(0, 2, 3, 1).is_contiguous()
    # survives load_state_dict and .to()
    sd = {k: v.contiguous() for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    m = m.to("cpu")
    assert m.app_plane[1].permute(0, 2, 3, 1).is_contiguous()
    assert m.app_plane[0].shape == (1, 12, 24, 20) and m.app_plane[1].shape == (1, 12, 28, 20)
    assert m.app_line[0].shape == (1, 12, 28, 1)


def test_step_size_constants_match_oracle():
    m = small()
    fc = vo.grid_constants(m.aabb, [20, 24, 28], 0.5)
    assert torch.equal(m.stepSize.cpu(), fc["step"])
    assert torch.equal(m.invaabbSize.cpu(), fc["inv"])
    assert torch.equal(m.units.cpu(), fc["units"])
    assert m.nSamples == fc["n_samples"]
    assert m.gridSize.tolist() == [20, 24, 28]


@pytest.mark.parametrize("mode", ["uniform-gaussian", "uniform-average"])
@pytest.mark.parametrize("param", [0.3, 0.15, 0.07, 0.0, 1e-5])
def test_blur_taps_match_oracle(mode, param):
    m = small()
    fc = vo.grid_constants(m.aabb, [20, 24, 28], 0.5)
    mine = m.get_kernel(None, mode, param, 64)
    ref = vo.blur_taps(fc, mode, param, 64)
    assert mine.shape == (65,)
    assert torch.allclose(mine, ref, rtol=0, atol=1e-7)
    with pytest.raises(RuntimeError):
        m.get_kernel(None, "diff", 0.1, 64)


def test_ndc_depth_table_matches_oracle_bitwise():
    m = small()
    m.near_far = [-1.0, 1.0]
    j = torch.rand(1, 77)
    z = m._ndc_table(77, False) + j.reshape(-1) * ((1.0 - -1.0) / 77)
    assert torch.equal(z, vo.ndc_depth_table([-1.0, 1.0], 77, j)[0])


def test_optimizer_groups_and_freeze():
    m = small()
    groups = m.get_optparam_groups(0.02, 0.001)
    assert len(groups) == 6
    assert [g["lr"] for g in groups] == [0.02] * 4 + [0.001] * 2
    n = sum(p.numel() for g in groups for p in g["params"])
    assert n == sum(p.numel() for p in m.parameters())
    m.freeze_scene(None)
    assert not any(p.requires_grad for p in m.parameters())
    m.unfreeze_scene(None)
    assert all(p.requires_grad for p in m.parameters())
    assert small(shading="SH").renderModule is None


def test_param_state_roundtrip():
    m = small(grid=(24, 20, 28))
    assert m.density_plane[0].shape == (1, 8, 20, 24) and m.density_line[0].shape == (1, 8, 28, 1)
    assert m.app_plane[2].permute(0, 2, 3, 1).is_contiguous()
    assert m.gridSize.tolist() == [24, 20, 28]
    ck = m.save_param_state()
    assert ck["tensorf_reset_kwargs"]["gridSize"] == [24, 20, 28]
    m2 = small(grid=(16, 16, 16))
    m2.load_param_state(ck)
    assert m2.density_plane[0].shape == (1, 8, 20, 24)
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m2.app_plane[1], m.app_plane[1])


def test_sweeps_and_maintenance_have_no_cpu_fallback():
    """Regularisers, optimiser step, upsampling and the alpha-mask update run as CUDA kernels only."""
    from joint_tensorf_b200.sweeps import FusedAdam
    m = small()
    for call in (m.density_L1, lambda: m.TV_loss_density(None), lambda: m.regularize_(1e-4),
                 lambda: m.upsample_volume_grid([20, 20, 20]), lambda: m.updateAlphaMask((8, 8, 8)),
                 lambda: m.compute_alpha(torch.zeros(4, 3), 0.1)):
        with pytest.raises(jt._lib.JtError):
            call()
    opt = FusedAdam(m.get_optparam_groups(0.02, 1e-3), betas=(0.9, 0.99))
    assert [g["lr"] for g in opt.param_groups] == [0.02] * 4 + [1e-3] * 2
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(jt._lib.JtError):
        opt.step()
    with pytest.raises(jt._lib.JtError):
        FusedAdam(m.parameters(), weight_decay=0.1)


def test_unsupported_options_fail_loudly():
    from gpu_common import default_opt
    m = small()
    opt = default_opt()
    opt["arch"]["abs_components"] = True
    with pytest.raises(jt._lib.JtError):
        m._check_opt(opt)
    with pytest.raises(Exception):
        small(shading="MLP_PE")
    with pytest.raises(jt._lib.JtError):
        m.compute_densityfeature(torch.zeros(4, 3), interp_mode="bicubic")


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    from gpu_common import default_opt
    m = small()
    with pytest.raises(jt._lib.JtError):
        m.forward(default_opt(), torch.zeros(4, 3), torch.ones(4, 3), N_samples=8)
    with pytest.raises(jt._lib.JtError):
        m.compute_densityfeature(torch.zeros(4, 3))


def test_mask_bit_packing():
    vol = (torch.rand(5, 6, 7) > 0.5).float()
    am = jt.AlphaGridMask("cpu", torch.tensor([[-1.0] * 3, [1.0] * 3]), vol)
    flat = vol.reshape(-1) > 0
    for n in (0, 1, 31, 32, 33, 100, flat.numel() - 1):
        word = int(am.bits[n >> 5]) & 0xFFFFFFFF
        assert bool((word >> (n & 31)) & 1) == bool(flat[n])
    assert am.gridSize.tolist() == [7, 6, 5]


def test_synthetic_rays_are_deterministic_and_hit_the_box():
    o1, d1, v1 = jt.synth.blender_rays(256, 8)
    o2, d2, _ = jt.synth.blender_rays(256, 8)
    assert torch.equal(o1, o2) and torch.equal(d1, d2)
    assert abs(float(o1.norm(dim=-1).mean()) - 4.0) < 1e-4
    fc = vo.grid_constants(torch.tensor([[-1.5] * 3, [1.5] * 3]), [128] * 3, 0.5)
    _, _, valid = vo.sample_ray(fc, [2.0, 6.0], o1, d1, 443, None)
    assert 0.3 < float(valid.float().mean()) < 0.9
    o, d, _ = jt.synth.llff_ndc_rays(128, 4)
    assert torch.allclose(o[:, 2], torch.full((128,), -1.0), atol=1e-5)


def test_supervision_blur_taps_and_scales_match_oracle():
    """supervision.blur_taps_2d / _scales (host side of process_GT_images, nerf.py:63-83) against the oracle's
    restatement for both golden option sets, every scale, several iterations."""
    import torch
    from common import load_golden
    from joint_tensorf_b200 import supervision as sv
    from joint_tensorf_b200.options import Namespace
    from oracle import image_oracle as io
    g = load_golden("image_prep")
    h, w = g["images"].shape[-2:]
    for name, o in g["opts"].items():
        opt = Namespace(dict({k: v for k, v in o.items() if k != "it"}, H=h, W=w))
        assert sv._scales(opt) == io.scales(o)
        for it in (0, o["it"], 333, 999):
            for sc in io.scales(o):
                taps, width = sv.blur_taps_2d(opt, it, sc)
                bp = torch.tensor(io.interp_schedule(float(it / o["max_iter"]), o["blur_2d_c2f_schedule"])) * sc
                wref = bp * (w + h) / 2
                assert abs(width - float(wref)) <= 1e-12
                kref = (io.gaussian_kernel if o["blur_2d_mode"] == "uniform-gaussian" else io.average_kernel)(
                    wref, o["blur_2d_c2f_kernel_size"])
                assert taps.shape == kref.shape and (taps - kref).abs().max() <= 1e-7, (name, it, sc)


def test_llff_views_reproduce_the_ndc_ray_generator():
    """synth.llff_views (pose / intrinsics matrices handed to the pose->ray kernel in bench.py's cfg4 leg) describes the
    same cameras as synth.llff_ndc_rays (the closed-form generator used for the oracle-side LLFF rays)."""
    import joint_tensorf_b200 as jt
    pose, intr = jt.synth.llff_views(8, (756, 1008), seed=1)
    assert pose.shape == (8, 3, 4) and intr.shape == (8, 3, 3)
    rot = pose[:, :, :3]
    eye = torch.eye(3).expand(8, 3, 3)
    assert (rot @ rot.transpose(1, 2) - eye).abs().max() < 1e-5          # proper rotations
    assert float(intr[0, 0, 0]) == pytest.approx(0.85 * 1008)
    o, d, view = jt.synth.llff_ndc_rays(64, 8, (756, 1008), seed=1)
    assert o.shape == (64, 3) and torch.isfinite(o).all() and torch.isfinite(d).all()
    assert (o[:, 2] + 1.0).abs().max() < 1e-4                              # origins sit on the NDC near plane z = -1


def test_workload_descriptions_name_the_baseline_config():
    import joint_tensorf_b200 as jt
    s = jt.synth.describe("cfg2_sh", 4096)
    assert "300^3" in s and "3x16" in s and "3x48" in s and "app_dim 27" in s and "SH" in s and "4096 rays/GPU" in s
    assert "MLP_Fea" in jt.synth.describe("cfg2") and "617x687x617" in jt.synth.describe("cfg4")


def test_opt_fixtures_match_the_reference_options_py():
    """tests/golden/opt_*.json (used by tests/test_gpu_reference_callsite.py on the GPU box) are exactly what the
    reference's options.py builds from its shipped YAMLs; re-derived here whenever /root/reference is mounted."""
    import json
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_opt_fixture as mof
    import ref_loader
    for y, mdl in mof.YAMLS.items():
        path = os.path.join(gold, f"opt_{y}.json")
        assert os.path.exists(path), path
        have = json.load(open(path))
        assert have["arch"]["tensorf"]["model"] == "BAT_VMSplit" and "c2f_kernel_size" in have
        if ref_loader.available():
            want = json.loads(json.dumps(mof.build(y, mdl), sort_keys=True, default=str))
            assert want == have, y
