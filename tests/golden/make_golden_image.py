"""Golden vectors for the 2-D supervision pre-processing and the render loss, from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_image.py

image_prep.pt : Model.process_GT_images + Model.get_edge_mask (model/nerf.py:57-149) called unbound on a stand-in
                `self` that carries only the attributes they read (train_data.all.image, it, tb), for two option
                sets (Blender YAML style: gaussian, 201 taps, sampled scale pool, hard masks; and a box-filter /
                soft-mask set), plus the render term of Graph.compute_loss (model/tensorf.py:99-124) and its
                autograd w.r.t. var.rgb for every loss / mask combination the YAML flags allow.
The engine modules import logging / plotting packages that are not installed here (lpips, matplotlib, visdom,
imageio, roma) and wandb; they are replaced by inert mocks -- none of them touches the arithmetic.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402

torch.set_num_threads(8)


class _MockPkg(MagicMock):
    __path__ = []


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = ("lpips", "matplotlib", "visdom", "imageio", "roma", "wandb", "mcubes", "trimesh", "mpl_toolkits")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _MockPkg()
        m.__name__, m.__path__, m.__spec__ = spec.name, [], spec
        return m

    def exec_module(self, m):
        pass


def load_engine():
    ref_loader.load()
    sys.meta_path.insert(0, _MockFinder())
    import importlib
    nerf = importlib.import_module("model.nerf")
    tensorf = importlib.import_module("model.tensorf")
    base = importlib.import_module("model.base")
    nerf.util_vis.tb_wandb_image = lambda *a, **k: None
    return nerf, tensorf, base


def smooth_images(gen, b, h, w):
    """Images with structure at several scales (so Sobel magnitudes spread around their mean), in [0,1]."""
    ys = torch.linspace(0, 1, h)[None, None, :, None]
    xs = torch.linspace(0, 1, w)[None, None, None, :]
    img = torch.zeros(b, 3, h, w)
    for f in (1.0, 3.0, 7.0):
        ph = torch.rand(b, 3, 1, 1, generator=gen) * 6.28
        a = torch.rand(b, 3, 1, 1, generator=gen)
        img = img + a * torch.sin(f * 6.28 * xs + ph) * torch.cos(f * 4.0 * ys + 0.5 * ph) / f
    img = img + 0.15 * torch.rand(b, 3, h, w, generator=gen)
    img[:, :, h // 3:h // 2, w // 4:w // 2] += 0.8            # a hard-edged box
    img = (img - img.amin()) / (img.amax() - img.amin())
    return img.contiguous()


OPTS = {
    "blender": dict(c2f_alternate_2D_mode="sample", c2f_alternate_2D_scale_pool=[0.0, 0.25, 0.5, 0.75, 1.0],
                    max_iter=1000, blur_2d_c2f_schedule=[0.025, 0.0125, 0.00625, 0.0, 0.0], blur_2d_mode="uniform-gaussian",
                    blur_2d_c2f_kernel_size=201, hard_edge_mask_mean_thresh=1.25, soft_edge_mask=False, it=100),
    "box_soft": dict(c2f_alternate_2D_mode="none", c2f_alternate_2D_scale_pool=[0.0, 1.0],
                     max_iter=1000, blur_2d_c2f_schedule=[0.07, 0.035, 0.015, 0.0], blur_2d_mode="uniform-average",
                     blur_2d_c2f_kernel_size=31, soft_edge_mask=True, it=200),
}


def main():
    nerf, tensorf, base = load_engine()
    gen = torch.Generator().manual_seed(23)
    b, h, w = 3, 40, 56
    images = smooth_images(gen, b, h, w)
    out = dict(images=images, opts=OPTS, prep={})
    masks = {}
    for name, o in OPTS.items():
        opt = ref_loader.AttrDict({k: v for k, v in o.items() if k != "it"})
        opt.H, opt.W, opt.device = h, w, "cpu"
        opt.tb = dict(num_images=[1, 1])
        me = types.SimpleNamespace(train_data=types.SimpleNamespace(all=ref_loader.AttrDict(image=images)),
                                   it=o["it"], tb=None)
        blurred = nerf.Model.process_GT_images(me, opt)
        edge = nerf.Model.get_edge_mask(me, opt, blurred)
        out["prep"][name] = dict(blurred={float(k): v.clone() for k, v in blurred.items()},
                                 edge={float(k): v.clone() for k, v in edge.items()})
        masks[name] = edge[1.0]

    # ---- render loss (tensorf.py:99-124)
    n = 50
    ray_idx = torch.randperm(h * w, generator=gen)[:n]
    rgb0 = torch.rand(b, n, 3, generator=gen)
    losses = []
    combos = [
        dict(tag="plain", flags=dict(), mask=None, it=0),
        dict(tag="alternate_off_iteration", flags=dict(edge_mask_on_render_loss=True, alternate_edge_loss=True,
                                                       edge_mask_before_iter=8000), mask="blender", it=1),
        dict(tag="hard_loss_hard_mask", flags=dict(edge_mask_on_render_loss=True, alternate_edge_loss=True,
                                                   edge_mask_before_iter=8000), mask="blender", it=2),
        dict(tag="soft_loss_hard_mask", flags=dict(edge_mask_on_render_loss=True, soft_edge_loss=True,
                                                   edge_mask_before_iter=8000), mask="blender", it=3),
        dict(tag="hard_loss_soft_mask", flags=dict(edge_mask_on_render_loss=True, edge_mask_before_iter=8000),
             mask="box_soft", it=4),
        dict(tag="soft_loss_soft_mask", flags=dict(edge_mask_on_render_loss=True, soft_edge_loss=True,
                                                   edge_mask_before_iter=8000), mask="box_soft", it=5),
        dict(tag="past_edge_iters", flags=dict(edge_mask_on_render_loss=True, edge_mask_before_iter=4), mask="blender",
             it=6),
        dict(tag="nan_pixel", flags=dict(edge_mask_on_render_loss=True, edge_mask_before_iter=8000), mask="blender",
             it=8, nan=True),
    ]
    for c in combos:
        opt = ref_loader.AttrDict(dict(c["flags"], edge_loss_factor=1.5, non_edge_loss_factor=0.5))
        opt.H, opt.W = h, w
        opt.loss_weight = dict(render=0)
        opt.nerf = dict(ray_sampling_strategy="all_view_rand_rays")
        rgb = rgb0.clone()
        if c.get("nan"):
            rgb[1, 7, 2] = float("nan")
        rgb.requires_grad_(True)
        var = ref_loader.AttrDict(idx=torch.arange(b), image=images, ray_idx=ray_idx, rgb=rgb,
                                  train_edge_masks=None if c["mask"] is None else masks[c["mask"]])
        me = types.SimpleNamespace(it=c["it"], MSE_loss=lambda p, l=0: base.Graph.MSE_loss(None, p, l), nerf=MagicMock(),
                                   tvloss=None)
        loss = tensorf.Graph.compute_loss(me, opt, var, mode="train")
        val = loss.render
        (g,) = torch.autograd.grad(val * 0.7, rgb)
        losses.append(dict(tag=c["tag"], flags=c["flags"], mask=c["mask"], it=c["it"], nan=bool(c.get("nan")),
                           loss=val.detach().clone(), d_rgb=g.clone(), upstream=0.7))
    # validation: every pixel, no ray_idx (tensorf.py:101-102)
    opt = ref_loader.AttrDict(edge_loss_factor=1.5, non_edge_loss_factor=0.5, H=h, W=w, loss_weight=dict(render=0),
                              nerf=dict(ray_sampling_strategy="all_view_rand_rays"))
    rgb_full = torch.rand(b, h * w, 3, generator=gen).requires_grad_(True)
    var = ref_loader.AttrDict(idx=torch.arange(b), image=images, ray_idx=ray_idx, rgb=rgb_full, train_edge_masks=None)
    me = types.SimpleNamespace(it=0, MSE_loss=lambda p, l=0: base.Graph.MSE_loss(None, p, l), nerf=MagicMock(), tvloss=None)
    val = tensorf.Graph.compute_loss(me, opt, var, mode="val").render
    out["loss_val"] = dict(rgb=rgb_full.detach().clone(), loss=val.detach().clone())
    out.update(ray_idx=ray_idx, rgb=rgb0, losses=losses, edge_loss_factor=1.5, non_edge_loss_factor=0.5)
    path = os.path.join(HERE, "image_prep.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for name in OPTS:
        e = out["prep"][name]["edge"][1.0]
        print(name, "scales", sorted(out["prep"][name]["blurred"]), "mask mean", float(e.float().mean()))
    for l in losses:
        print(l["tag"], float(l["loss"]))


if __name__ == "__main__":
    main()
