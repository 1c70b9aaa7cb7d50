"""Drive B200_VMSplit exactly the way the reference engine does.

`opt` is the reference's own options.py output on its shipped YAMLs (tests/golden/opt_*.json, made by
tests/golden/make_opt_fixture.py from /root/reference; tests/test_host_logic.py re-derives it when the reference is
mounted). The constructor keywords are those of model/tensorf.py:375-397, resolution / n_samples follow
tensorf.py:449-461, the forward keywords those of tensorf.py:246-262 (with the schedule values the engine computes at
the start of training), then the per-step regularisers of tensorf.py:127-130, the optimiser groups of tensorf.py:473
and the maintenance calls of tensorf.py:415-425,480-489."""
import json
import os

import numpy as np
import pytest
import torch

import joint_tensorf_b200 as jt
from common import GOLDEN_DIR, vo
from joint_tensorf_b200.options import Namespace
from oracle import field_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def load_opt(name):
    return Namespace(json.load(open(os.path.join(GOLDEN_DIR, f"opt_{name}.json"))))


def find_resolution(opt, bbox, n_voxels):          # tensorf.py:449-456
    xyz_min, xyz_max = bbox[0], bbox[1]
    voxel_size = ((xyz_max - xyz_min).prod() / n_voxels).pow(1 / 3)
    scale = torch.tensor(opt.train_schedule.resolution_scale_init)
    return ((xyz_max - xyz_min) / voxel_size * scale).long().tolist()


def find_n_samples(opt, resolution):               # tensorf.py:458-461
    return min(int(opt.nerf.sample_intvs), int(np.linalg.norm(resolution) / opt.nerf.step_ratio))


def construct(opt, bbox, resolution, cls):         # tensorf.py:375-397, keyword for keyword
    density_n_comp = list(map(int, opt.arch.tensorf.density_components))
    appearance_n_comp = list(map(int, opt.arch.tensorf.color_components))
    return cls(
        bbox, resolution, opt.device,
        density_n_comp=density_n_comp, appearance_n_comp=appearance_n_comp,
        app_dim=3 if opt.arch.shading == "RGB" else opt.arch.shading.app_dim,
        near_far=opt.nerf.depth.range, shadingMode=opt.arch.shading.model,
        alphaMask_thres=opt.train_schedule.alpha_mask_threshold, density_shift=opt.arch.density_shift,
        distance_scale=opt.arch.distance_scale, pos_pe=opt.arch.shading.pose_pe, view_pe=opt.arch.shading.view_pe,
        fea_pe=opt.arch.shading.fea_pe, featureC=opt.arch.shading.mlp_hidden_dim, step_ratio=opt.nerf.step_ratio,
        fea2denseAct=opt.arch.feature_to_density_activation, dtype=torch.float32,
        volume_init_scale=opt.arch.tensorf.volume_init_scale,
        rayMarch_weight_thres=opt.arch.tensorf.rayMarch_weight_thres,
        volume_init_bias=opt.arch.tensorf.volume_init_bias)


class TVLossWeight:          # the only thing TV_loss_* read from the reference's TVLoss module (tensorBase.py:16-20)
    TVLoss_weight = 1.0


@pytest.mark.parametrize("yaml", ["bat_blender_VM_MLP", "bat_llff_VM_MLP"])
def test_module_driven_like_the_reference_engine(yaml):
    opt = load_opt(yaml)
    ndc = bool(opt.camera.ndc)
    bbox = torch.tensor(opt.data.scene_bbox).float().view(2, 3)
    resolution = find_resolution(opt, bbox, opt.train_schedule.n_voxel_init)
    n_samples = find_n_samples(opt, resolution)
    torch.manual_seed(0)
    m = construct(opt, bbox.to(DEV), resolution, jt.B200_VMSplit)
    assert [tuple(p.shape) for p in m.density_plane] == [(1, 16, resolution[1], resolution[0]), (1, 16, resolution[2], resolution[0]),
                                                         (1, 16, resolution[2], resolution[1])]
    n = 192
    if ndc:
        m.near_far[0] = opt.tensorf_near_plane_schedule[0]          # tensorf.py:231 writes the schedule value in place
        o, d, _ = jt.synth.llff_ndc_rays(n, 8, seed=21)
    else:
        o, d, _ = jt.synth.blender_rays(n, 8, seed=21)
        with torch.no_grad():                                      # make the random-init field non-trivial
            for i in range(3):
                m.density_plane[i].mul_(5.0)
                m.density_line[i].mul_(5.0)
    center, ray = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    scale = opt.c2f_random_density_scale_pool[3]                   # np.random.choice(...) at tensorf.py:198
    fkw = dict(white_bg=opt.nerf.setbg_opaque, is_test_optim=False, ndc_ray=opt.camera.ndc, N_samples=n_samples,
               c2f_parameter_density=opt.c2f_schedule_density[1] * scale, c2f_parameter_color=opt.c2f_schedule_color[1],
               c2f_mode=opt.c2f_mode, c2f_kernel_size=opt.c2f_kernel_size, fea_pe_progress=opt.c2f_fea_pe_schedule[0],
               view_pe_progress=opt.c2f_view_pe_schedule[0])
    # ---- a training call exactly as tensorf.py:246-262 issues it (stratified sampling: is_train True)
    torch.manual_seed(5)
    rgb, depth, opacity = m.forward(opt, center=center.view(-1, 3), ray_dir=ray.view(-1, 3),
                                    is_train=opt.nerf.sample_stratified, **fkw)
    assert rgb.shape == (n, 3) and depth.shape == (n,) and opacity.shape == (n,)
    loss_render = ((rgb - 0.5) ** 2).mean()
    # regularisers as tensorf.py:127-130 calls them, weights of the YAML
    l1, tvd, tva = m.density_L1(), m.TV_loss_density(TVLossWeight()), m.TV_loss_app(TVLossWeight())
    w_l1 = float(opt.loss_weight.L1.init)
    total = loss_render + w_l1 * l1 + float(opt.loss_weight.TV_density) * tvd + float(opt.loss_weight.TV_color) * tva
    total.backward()
    assert torch.isfinite(center.grad).all() and torch.isfinite(ray.grad).all()
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    # ---- the same call in evaluation mode against the CPU oracle (no stratified jitter to inject)
    sd = {k: v.detach().cpu().contiguous().clone() for k, v in m.state_dict().items()}
    field = vo.Field(aabb=bbox, grid=resolution, params=sd, near_far=[float(v) for v in m.near_far],
                     step_ratio=opt.nerf.step_ratio, density_shift=float(opt.arch.density_shift),
                     distance_scale=opt.arch.distance_scale, weight_thres=opt.arch.tensorf.rayMarch_weight_thres,
                     act=opt.arch.feature_to_density_activation, shading=opt.arch.shading.model,
                     view_pe=opt.arch.shading.view_pe, fea_pe=opt.arch.shading.fea_pe)
    with torch.no_grad():
        rgb_e, depth_e, acc_e = m.forward(opt, center=center.view(-1, 3), ray_dir=ray.view(-1, 3), is_train=False, **fkw)
        rgb_r, depth_r, acc_r = vo.render(field, o, d, n_samples=n_samples, white_bg=opt.nerf.setbg_opaque, ndc=ndc,
                                          blur_mode=opt.c2f_mode, blur_density=fkw["c2f_parameter_density"],
                                          blur_color=fkw["c2f_parameter_color"], kernel_size=opt.c2f_kernel_size)
    assert (rgb_e.cpu() - rgb_r).abs().max() <= 1e-4 and (acc_e.cpu() - acc_r).abs().max() <= 1e-4
    assert (depth_e.cpu() - depth_r).abs().max() <= 2e-4
    assert abs(float(l1) - float(fo.density_l1(sd))) <= 1e-5 * max(1.0, float(l1))
    assert abs(float(tvd) - float(fo.tv_loss_density(sd))) <= 1e-4 * max(1e-6, float(tvd))
    assert abs(float(tva) - float(fo.tv_loss_app(sd))) <= 1e-4 * max(1e-6, float(tva))
    # ---- optimiser groups (tensorf.py:473-475) and one Adam step through them
    groups = m.get_optparam_groups(opt.optim.lr_index, opt.optim.lr_basis)
    assert sum(len(list(g["params"])) for g in m.get_optparam_groups()) == len(list(m.parameters()))
    torch.optim.Adam(groups, betas=(0.9, 0.99)).step()
    # ---- maintenance as update_schedule / _update_alphamask drive it (tensorf.py:415-425, 480-489)
    n_voxel_list = torch.round(torch.exp(torch.linspace(np.log(opt.train_schedule.n_voxel_init),
                                                        np.log(opt.train_schedule.n_voxel_final),
                                                        len(opt.train_schedule.upsample_iters) + 1))).long().tolist()[1:]
    res2 = find_resolution(opt, bbox, n_voxel_list[0])
    m.upsample_volume_grid(res2)
    assert m.gridSize.tolist() == res2
    with torch.no_grad():
        out = m.forward(opt, center=center.view(-1, 3), ray_dir=ray.view(-1, 3), is_train=False,
                        **dict(fkw, N_samples=find_n_samples(opt, res2)))
    assert all(torch.isfinite(t).all() for t in out)
    if res2[0] * res2[1] * res2[2] < 256 ** 3 and not ndc:           # tensorf.py:483 (Blender: a shrinkable object)
        new_aabb = m.updateAlphaMask(tuple(res2))
        m.shrink(new_aabb)
        ck = m.save_param_state()
        assert "alphaMask.aabb" in ck and ck["tensorf_reset_kwargs"]["gridSize"] == m._grid
        with torch.no_grad():
            out = m.forward(opt, center=center.view(-1, 3), ray_dir=ray.view(-1, 3), is_train=False,
                            **dict(fkw, c2f_parameter_density=None, c2f_parameter_color=None, c2f_mode=None,
                                   c2f_kernel_size=None, N_samples=find_n_samples(opt, m._grid)))
        assert all(torch.isfinite(t).all() for t in out)
