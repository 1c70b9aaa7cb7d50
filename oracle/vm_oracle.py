"""CPU oracle for the TensoRF-VM volume-rendering hot path (TEST INFRASTRUCTURE).

This file is the checker, never the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it. The product (`joint-tensorf_b200/`) never does.

It restates, as plain functions over torch CPU tensors, the algorithm of the
reference field layer `model/tensorf_repr` of Nemo1999/Joint-TensoRF. Each
function cites the reference file:line it follows. It calls the same ATen
operators the reference calls (`grid_sample`, `conv1d`, `cumprod`,
`softplus`, `Linear`) so that (a) results are bit-close to the reference and
(b) its CPU timing is representative of the reference's own CPU path.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against the LIVE reference, imported in the build
container: `tests/golden/make_golden.py` runs reference `BAT_VMSplit.forward`
+ `backward` on seeded inputs, stores inputs and outputs under
`tests/golden/*.pt`, and `tests/test_oracle_golden.py` checks this file against
those fixtures (valid masks bit-exact, everything else <= 2e-6).

Third-party arithmetic: PyTorch ATen (reference pins torch 1.13.1,
`env_setup/install.sh:12`; this image has 2.11.0). Semantics used:
`grid_sample(mode="bilinear", padding_mode="zeros", align_corners=True)`,
`pad(mode="replicate")`, `conv1d` (cross-correlation), `cumprod`, `softplus`
(beta=1, threshold=20).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

MAT_MODE = ((0, 1), (0, 2), (1, 2))  # tensorBase.py:405
VEC_MODE = (2, 1, 0)                 # tensorBase.py:406


# --------------------------------------------------------------------------- state
@dataclass
class Field:
    """Everything `forward` reads besides the rays (tensorBase.py:374-452)."""
    aabb: torch.Tensor                   # [2,3] fp32
    grid: List[int]                      # gridSize (x,y,z)
    params: Dict[str, torch.Tensor]      # reference state_dict names
    near_far: List[float] = field(default_factory=lambda: [2.0, 6.0])
    step_ratio: float = 0.5
    density_shift: float = -10.0
    distance_scale: float = 25.0
    weight_thres: float = 1e-6           # rayMarch_weight_thres
    act: str = "softplus"                # fea2denseAct
    shading: str = "MLP_Fea"             # MLP_Fea | MLP_Fea_WeakView | SH
    view_pe: int = 2
    fea_pe: int = 2
    mask_volume: Optional[torch.Tensor] = None   # [D,H,W] float {0,1}
    mask_aabb: Optional[torch.Tensor] = None


def grid_constants(aabb, grid, step_ratio):
    """tensorBase.py:477-488 (update_stepSize). All fp32 torch arithmetic."""
    aabb = aabb.to(torch.float32)
    size = aabb[1] - aabb[0]
    g = torch.tensor(list(grid), dtype=torch.long)
    units = size / (g - 1)
    step = torch.mean(units) * step_ratio
    diag = torch.sqrt(torch.sum(torch.square(size)))
    return dict(aabb=aabb, size=size, inv=2.0 / size, units=units, step=step,
                diag=diag, n_samples=int((diag / step).item()) + 1, g=g)


# --------------------------------------------------------------------------- sampling
def sample_ray(fc, near_far, rays_o, rays_d, n_samples, jitter=None):
    """tensorBase.py:572-612. `jitter` [N,1] replaces `rand_like(rng[:, [0]])`
    (None == is_train False). Returns pts [N,S,3], z [N,S], valid [N,S]."""
    near, far = near_far
    od, dd = rays_o.detach(), rays_d.detach()
    vec = torch.where(dd == 0, torch.full_like(dd, 1e-6), dd)
    ra = (fc["aabb"][1] - od) / vec
    rb = (fc["aabb"][0] - od) / vec
    t_min = torch.minimum(ra, rb).amax(-1).clamp(min=near, max=far)
    rng = torch.arange(n_samples, dtype=torch.float32)[None]
    if jitter is not None:
        rng = rng.repeat(rays_d.shape[-2], 1)
        rng = rng + jitter
    z = t_min[..., None] + fc["step"] * rng
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
    outside = ((fc["aabb"][0] > pts) | (pts > fc["aabb"][1])).any(dim=-1)
    return pts, z, ~outside


def ndc_depth_table(near_far, n_samples, jitter=None):
    """tensorBase.py:556-559: linspace(near, far, S)[None] (+ rand * (far-near)/S)."""
    near, far = near_far
    z = torch.linspace(near, far, n_samples, dtype=torch.float32).unsqueeze(0)
    if jitter is not None:
        z = z + jitter * ((far - near) / n_samples)
    return z


def sample_ray_ndc(fc, near_far, rays_o, rays_d, n_samples, jitter=None):
    """tensorBase.py:554-571 with simulate_euclid_* False (all YAMLs)."""
    z = ndc_depth_table(near_far, n_samples, jitter)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
    outside = ((fc["aabb"][0] > pts) | (pts > fc["aabb"][1])).any(dim=-1)
    return pts, z, ~outside


def alpha_mask_lookup(volume, mask_aabb, xyz):
    """AlphaGridMask.sample_alpha, tensorBase.py:80-98. volume [D,H,W]."""
    size = mask_aabb[1] - mask_aabb[0]
    inv = 1.0 / size * 2
    u = (xyz - mask_aabb[0]) * inv - 1
    vol = volume.view(1, 1, *volume.shape[-3:])
    return F.grid_sample(vol, u.view(1, -1, 1, 1, 3), align_corners=True).view(-1)


def normalize_coord(fc, xyz):
    """tensorBase.py:502-503."""
    return (xyz - fc["aabb"][0]) * fc["inv"] - 1


# --------------------------------------------------------------------------- blur
def gaussian_taps(t, kernel_size):
    """kernels.py:16-22. `t` is a 0-d fp32 tensor; taps clamped to <= 1, not normalised."""
    ns = torch.arange(-(kernel_size // 2), kernel_size // 2 + 1, dtype=torch.float32)
    tt = max(t, 0.0001)
    expo = -0.5 * (ns / tt) * (ns / tt)
    k = 1 / (tt * math.sqrt(2 * math.pi)) * torch.exp(expo)
    return torch.clamp(k, max=1.0)


def average_taps(t, kernel_size):
    """kernels.py:24-41."""
    if kernel_size % 2 == 0:
        kernel_size += 1
    if isinstance(t, torch.Tensor):
        t = t.item()
    half = kernel_size // 2
    lo = min(math.floor(t), half)
    k0 = torch.zeros(kernel_size)
    k0[half - lo:half + lo + 1] = 1 / (lo * 2 + 1)
    hi = min(math.ceil(t), half)
    k1 = torch.zeros(kernel_size)
    k1[half - hi:half + hi + 1] = 1 / (hi * 2 + 1)
    return (t % 1.0) * k1 + (1 - t % 1.0) * k0


def blur_taps(fc, mode, param, kernel_size):
    """BatBase.get_kernel, batBase.py:13-25."""
    scale = torch.mean(fc["g"] / fc["size"])
    if mode == "uniform-gaussian":
        return gaussian_taps(scale * param, kernel_size).to(torch.float32)
    if mode == "uniform-average":
        return average_taps(scale * param, kernel_size).to(torch.float32)
    raise RuntimeError(f"invalid c2f_mode {mode}")


def blur_line(taps, line):
    """BAT_VMSplit.convolute_line, bateRF.py:8-19. line [1,C,L,1]."""
    c = line.shape[1]
    half = taps.shape[-1] // 2
    taps = taps.to(line.dtype)
    x = line.squeeze(-1).view(c, 1, -1)
    x = F.pad(x, (half, half), mode="replicate")
    x = F.conv1d(x, taps.view(1, 1, -1))
    return x.view(1, c, -1).unsqueeze(-1)


def blur_plane(taps, plane, hh, ww):
    """BAT_VMSplit.convolute_plane, bateRF.py:21-39. NOTE the reference calls it
    with (hh, ww) = (g[m0], g[m1]) although storage is [1,C,g[m1],g[m0]]
    (bateRF.py:68,76) -- a memory re-interpretation on non-cubic grids
    (SURVEY.md Appendix B-3). Reproduced, not fixed."""
    c = plane.shape[1]
    half = taps.shape[-1] // 2
    k = taps.to(plane.dtype).view(1, 1, -1)
    x = plane.reshape(c, hh, ww)
    x = F.pad(x, (half, half), mode="replicate")
    x = F.conv1d(x, k.expand(hh, 1, -1), groups=hh)
    x = x.view(c, hh, ww).permute(0, 2, 1)
    x = F.pad(x, (half, half), mode="replicate")
    x = F.conv1d(x, k.expand(ww, 1, -1), groups=ww)
    return x.view(1, c, ww, hh).permute(0, 1, 3, 2).contiguous()


# --------------------------------------------------------------------------- VM features
def _plane_line_grids(u):
    cp = torch.stack([u[..., list(MAT_MODE[i])] for i in range(3)]).view(3, -1, 1, 2)
    cl = torch.stack([u[..., VEC_MODE[i]] for i in range(3)])
    cl = torch.stack((torch.zeros_like(cl), cl), dim=-1).view(3, -1, 1, 2)
    return cp, cl


def _factors(field, fc, prefix, taps):
    out = []
    for i in range(3):
        p = field.params[f"{prefix}_plane.{i}"]
        l = field.params[f"{prefix}_line.{i}"]
        if taps is not None:
            m0, m1 = MAT_MODE[i]
            p = blur_plane(taps, p, int(fc["g"][m0]), int(fc["g"][m1]))
            l = blur_line(taps, l)
        out.append((p, l))
    return out


def density_feature(field, fc, u, taps=None):
    """BAT_VMSplit.compute_densityfeature, bateRF.py:41-94 (all arch flags False)."""
    cp, cl = _plane_line_grids(u)
    sig = torch.zeros((u.shape[0],), dtype=u.dtype)
    for i, (p, l) in enumerate(_factors(field, fc, "density", taps)):
        pc = F.grid_sample(p, cp[[i]], mode="bilinear", align_corners=True).view(-1, u.shape[0])
        lc = F.grid_sample(l, cl[[i]], mode="bilinear", align_corners=True).view(-1, u.shape[0])
        sig = sig + torch.sum(pc * lc, dim=0)
    return sig


def app_components(field, fc, u, taps=None):
    """The [A, sum C_app] plane*line products before basis_mat (bateRF.py:97-128)."""
    cp, cl = _plane_line_grids(u)
    pcs, lcs = [], []
    for i, (p, l) in enumerate(_factors(field, fc, "app", taps)):
        pcs.append(F.grid_sample(p, cp[[i]], mode="bilinear", align_corners=True).view(-1, u.shape[0]))
        lcs.append(F.grid_sample(l, cl[[i]], mode="bilinear", align_corners=True).view(-1, u.shape[0]))
    return (torch.cat(pcs) * torch.cat(lcs)).T


def app_feature(field, fc, u, taps=None):
    """BAT_VMSplit.compute_appfeature, bateRF.py:97-130."""
    return F.linear(app_components(field, fc, u, taps), field.params["basis_mat.weight"])


def feature2density(field, x):
    """tensorBase.py:696-700."""
    if field.act == "softplus":
        return F.softplus(x + field.density_shift)
    return F.relu(x + field.density_shift)


def raw2alpha(sigma, dist):
    """tensorBase.py:57-65."""
    alpha = 1.0 - torch.exp(-sigma * dist)
    t = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1, dtype=alpha.dtype), 1.0 - alpha + 1e-10], -1), -1)
    return alpha, alpha * t[:, :-1], t[:, -1:]


# --------------------------------------------------------------------------- shading
def positional_encoding(x, freqs, progress=1.0):
    """tensorBase.py:43-55."""
    lv = torch.arange(freqs)
    bands = 2 ** lv
    mask = (progress * freqs - lv).clamp_(min=0.0, max=1)
    p = x[..., None] * bands
    p = torch.cat([torch.sin(p) * mask, torch.cos(p) * mask], dim=-1)
    return p.reshape(x.shape[:-1] + (freqs * 2 * x.shape[-1],))


def shade_mlp_fea(field, dirs, feat, view_prog=1.0, fea_prog=1.0):
    """MLPRender_Fea.forward, tensorBase.py:116-126."""
    q = field.params
    x = [feat, dirs]
    if field.fea_pe > 0:
        x.append(positional_encoding(feat, field.fea_pe, fea_prog))
    if field.view_pe > 0:
        x.append(positional_encoding(dirs, field.view_pe, view_prog))
    h = torch.cat(x, dim=-1)
    h = F.relu(F.linear(h, q["renderModule.mlp.0.weight"], q["renderModule.mlp.0.bias"]))
    h = F.relu(F.linear(h, q["renderModule.mlp.2.weight"], q["renderModule.mlp.2.bias"]))
    return torch.sigmoid(F.linear(h, q["renderModule.mlp.4.weight"], q["renderModule.mlp.4.bias"]))


def shade_weakview(field, dirs, feat, view_prog=1.0, fea_prog=1.0):
    """MLPRender_Fea_WeakView.forward, tensorBase.py:198-214."""
    q = field.params
    x = [feat]
    if field.fea_pe > 0:
        x.append(positional_encoding(feat, field.fea_pe, fea_prog))
    h = torch.cat(x, dim=-1)
    h = F.relu(F.linear(h, q["renderModule.layer1.weight"], q["renderModule.layer1.bias"]))
    h = F.relu(F.linear(h, q["renderModule.layer2.weight"], q["renderModule.layer2.bias"]))
    mid = []
    if field.view_pe > 0:
        mid.append(positional_encoding(dirs, field.view_pe, view_prog))
    mid.append(h)
    return torch.sigmoid(F.linear(torch.cat(mid, dim=-1), q["renderModule.layer3.weight"],
                                  q["renderModule.layer3.bias"]))


_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
       -1.0925484305920792, 0.5462742152960396)


def sh_basis_deg2(d):
    """sh.py:88-113 for deg == 2 (9 bases)."""
    x, y, z = d.unbind(-1)
    out = torch.empty((*d.shape[:-1], 9), dtype=d.dtype)
    out[..., 0] = _C0
    out[..., 1] = -_C1 * y
    out[..., 2] = _C1 * z
    out[..., 3] = -_C1 * x
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    out[..., 4] = _C2[0] * xy
    out[..., 5] = _C2[1] * yz
    out[..., 6] = _C2[2] * (2.0 * zz - xx - yy)
    out[..., 7] = _C2[3] * xz
    out[..., 8] = _C2[4] * (xx - yy)
    return out


def shade_sh(field, dirs, feat, *_):
    """SHRender, tensorBase.py:68-72 (the reference crashes calling it through
    forward -- SURVEY.md Appendix B-1 -- the maths is what is restated)."""
    sh = sh_basis_deg2(dirs)[:, None]
    f = feat.view(-1, 3, sh.shape[-1])
    return torch.relu(torch.sum(sh * f, dim=-1) + 0.5)


_SHADERS = {"MLP_Fea": shade_mlp_fea, "MLP_Fea_WeakView": shade_weakview, "SH": shade_sh}


# --------------------------------------------------------------------------- forward
def render(field, rays_o, rays_d, *, n_samples, white_bg=True, jitter=None, ndc=False,
           blur_mode=None, blur_density=None, blur_color=None, kernel_size=None,
           view_prog=1.0, fea_prog=1.0, detail=False, exact=False):
    """BatBase.forward, batBase.py:44-165 (detach_viewdirs/detach_xyz True,
    two-stage / predict_density heads off, `bg coin flip` folded into white_bg).

    jitter: [N,1] (metric rays) or [1,S] (NDC) uniform numbers, None == is_train False.
    Differentiable w.r.t. field.params, rays_o, rays_d. Returns rgb[N,3],
    depth[N], acc[N] (+ dict of intermediates when detail=True).

    exact=True (not the reference's arithmetic; a yardstick for it): the samples are placed in fp32
    exactly as above -- the valid mask is the bit-exact class -- and everything after that runs in
    float64 (field.params must be float64). The tests use it to tell a real discrepancy from the
    reference's own fp32 rounding where a gradient is ill-conditioned (transmittance cancellation in
    nearly opaque fields, 1 - exp(-x) for x ~ 1e-4)."""
    fc = grid_constants(field.aabb, field.grid, field.step_ratio)
    dirs = rays_d
    if ndc:
        pts, z, valid = sample_ray_ndc(fc, field.near_far, rays_o, dirs, n_samples, jitter)
        dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
        norm = torch.norm(dirs, dim=-1, keepdim=True)
        dists = dists * norm
        dirs = dirs / norm
    else:
        pts, z, valid = sample_ray(fc, field.near_far, rays_o, dirs, n_samples, jitter)
        dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
    dirs = dirs.view(-1, 1, 3).expand(pts.shape).detach()     # detach_viewdirs
    if exact:
        pts, z, dists, dirs = pts.double(), z.double(), dists.double(), dirs.double()

    blur_on = blur_density is not None or blur_color is not None
    if field.mask_volume is not None and not blur_on:          # batBase.py:76-82
        keep = alpha_mask_lookup(field.mask_volume, field.mask_aabb, pts[valid]) > 0
        bad = ~valid
        bad[valid] |= ~keep
        valid = ~bad

    taps_d = taps_c = None
    if blur_mode is not None:                                  # batBase.py:91-101
        taps_d = blur_taps(fc, blur_mode, blur_density, kernel_size)
        taps_c = blur_taps(fc, blur_mode, blur_color, kernel_size)

    sigma = torch.zeros(pts.shape[:-1], dtype=pts.dtype)
    rgb = torch.zeros((*pts.shape[:2], 3), dtype=pts.dtype)
    u = normalize_coord(fc, pts)
    sig_feat = None
    if valid.any():
        sig_feat = density_feature(field, fc, u[valid], taps_d)
        sigma[valid] = feature2density(field, sig_feat)
    alpha, weight, bg = raw2alpha(sigma, dists * field.distance_scale)
    app_mask = weight > field.weight_thres
    if app_mask.any():
        feat = app_feature(field, fc, u[app_mask], taps_c)
        rgb[app_mask] = _SHADERS[field.shading](field, dirs[app_mask], feat, view_prog, fea_prog)

    acc = torch.sum(weight, -1)
    rgb_map = torch.sum(weight[..., None] * rgb, -2)
    with torch.no_grad():
        depth = torch.sum(weight * z, -1) + (1.0 - acc) * rays_d[..., -1]
        depth = depth - field.near_far[0] + 0.05
    if white_bg:
        rgb_map = rgb_map + (1.0 - acc[..., None])
    rgb_map = rgb_map.clamp(0, 1)
    if detail:
        return rgb_map, depth, acc, dict(valid=valid, z=z, pts=pts, sigma_feat=sig_feat, sigma=sigma,
                                         weight=weight, app_mask=app_mask, rgb=rgb, taps_d=taps_d, taps_c=taps_c)
    return rgb_map, depth, acc


def compute_alpha(field, xyz, length, taps=None):
    """BatBase.compute_alpha, batBase.py:27-42 (kernel cached from last forward -> `taps`)."""
    fc = grid_constants(field.aabb, field.grid, field.step_ratio)
    if field.mask_volume is not None:
        keep = alpha_mask_lookup(field.mask_volume, field.mask_aabb, xyz) > 0
    else:
        keep = torch.ones_like(xyz[:, 0], dtype=torch.bool)
    sigma = torch.zeros(xyz.shape[:-1])
    if keep.any():
        sigma[keep] = feature2density(field, density_feature(field, fc, normalize_coord(fc, xyz[keep]), taps))
    return 1 - torch.exp(-sigma * length)


# --------------------------------------------------------------------------- construction helpers
def init_params(grid, dens_comp, app_comp, app_dim, shading, hidden, view_pe, fea_pe,
                scale, bias, seed=0):
    """Random-init factors |bias + scale*N(0,1)| (tensoRF.py:159-169) and default
    torch Linear init for basis_mat / shading head (tensoRF.py:156,
    tensorBase.py:104-114,184-196). Returns a reference-named state dict."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for pre, comps in (("density", dens_comp), ("app", app_comp)):
        for i in range(3):
            m0, m1 = MAT_MODE[i]
            p[f"{pre}_plane.{i}"] = torch.abs(bias + scale * torch.randn((1, comps[i], grid[m1], grid[m0]), generator=g))
            p[f"{pre}_line.{i}"] = torch.abs(bias + scale * torch.randn((1, comps[i], grid[VEC_MODE[i]], 1), generator=g))

    def lin(n_out, n_in, zero_bias=False, has_bias=True):
        bound = 1 / math.sqrt(n_in)
        w = (torch.rand((n_out, n_in), generator=g) * 2 - 1) * bound
        b = torch.zeros(n_out) if zero_bias else (torch.rand((n_out,), generator=g) * 2 - 1) * bound
        return (w, b) if has_bias else (w, None)

    p["basis_mat.weight"] = lin(app_dim, sum(app_comp), has_bias=False)[0]
    if shading == "MLP_Fea":
        n_in = 2 * view_pe * 3 + 2 * fea_pe * app_dim + 3 + app_dim
        for name, (o, i, zb) in {"mlp.0": (hidden, n_in, False), "mlp.2": (hidden, hidden, False),
                                 "mlp.4": (3, hidden, True)}.items():
            w, b = lin(o, i, zb)
            p[f"renderModule.{name}.weight"], p[f"renderModule.{name}.bias"] = w, b
    elif shading == "MLP_Fea_WeakView":
        n_in = (2 * fea_pe + 1) * app_dim
        for name, (o, i, zb) in {"layer1": (hidden, n_in, False), "layer2": (hidden, hidden, False),
                                 "layer3": (3, hidden + 2 * view_pe * 3, True)}.items():
            w, b = lin(o, i, zb)
            p[f"renderModule.{name}.weight"], p[f"renderModule.{name}.bias"] = w, b
    return p
