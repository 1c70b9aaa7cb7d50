"""Shared helpers for the test-suite: golden fixtures -> oracle Field / run."""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import vm_oracle as vo  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names():
    """Render-path fixtures (tests/golden/make_golden.py)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))
    return [n for n in names if not n.startswith(("camera_", "field_", "image_"))]


def camera_golden_names():
    """Pose -> ray fixtures (tests/golden/make_golden_camera.py)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "camera_*.pt")))
    return names


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def golden_valid(g):
    n, s = g["rays_o"].shape[0], g["n_samples"]
    bits = np.unpackbits(g["valid_packed"].numpy())[: n * s].reshape(n, s)
    return torch.from_numpy(bits.astype(bool))


def field_from_golden(g, requires_grad=True):
    case, kw = g["case"], g["field_kw"]
    params = {k: v.clone().requires_grad_(requires_grad) for k, v in g["state_dict"].items()}
    return vo.Field(aabb=g["aabb"].clone(), grid=list(case["grid"]), params=params,
                    near_far=list(kw["near_far"]), step_ratio=kw["step_ratio"],
                    density_shift=float(kw["density_shift"]), distance_scale=25.0,
                    weight_thres=kw["rayMarch_weight_thres"], act=kw["fea2denseAct"],
                    shading=case["shading"], view_pe=2, fea_pe=2,
                    mask_volume=g["mask_volume"], mask_aabb=(g["aabb"].clone() if g["mask_volume"] is not None else None))


def render_kwargs_from_golden(g):
    case = g["case"]
    ndc = case.get("ndc", False)
    kw = dict(n_samples=g["n_samples"], white_bg=not ndc, jitter=g["jitter"], ndc=ndc)
    if case["blur"] is not None:
        kw.update(blur_mode="uniform-gaussian", blur_density=case["blur"][0], blur_color=case["blur"][1], kernel_size=64)
    return kw


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
