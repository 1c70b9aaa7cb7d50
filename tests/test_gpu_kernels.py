"""Per-kernel GPU checks against the oracle's pieces / plain fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

import joint_tensorf_b200 as jt
from common import load_golden, rel_err, vo
from gpu_common import module_from_golden
from joint_tensorf_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cl(x):
    return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


@pytest.mark.parametrize("shape", [(16, 40, 40, 40, 40), (12, 28, 24, 24, 28), (8, 300, 7, 7, 300), (4, 5, 9, 9, 5)])
@pytest.mark.parametrize("ntaps", [65, 9])
def test_blur_plane_forward_and_adjoint(shape, ntaps):
    c, h_store, w_store, hq, wq = shape          # storage [1,C,h_store,w_store], blurred as (hq, wq)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, c, h_store, w_store), generator=g)
    taps = torch.rand((ntaps,), generator=g) * 0.2
    ref = vo.blur_plane(taps, x, hq, wq)
    xd = _cl(x.to(DEV)).requires_grad_(True)
    y = ops.BlurFactor.apply(xd, taps.to(DEV), hq, wq, 3)
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max() <= 2e-5
    # adjoint: <K x, r> == <x, K^T r>, and against autograd of the oracle
    r = torch.randn(ref.shape, generator=g)
    xr = x.clone().requires_grad_(True)
    (vo.blur_plane(taps, xr, hq, wq) * r).sum().backward()
    (y * r.to(DEV)).sum().backward()
    assert rel_err(xd.grad.cpu(), xr.grad) <= 2e-5


@pytest.mark.parametrize("L", [2, 33, 300, 687])
def test_blur_line(L):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((1, 16, L, 1), generator=g)
    taps = vo.gaussian_taps(torch.tensor(3.7), 64)
    xr = x.clone().requires_grad_(True)
    ref = vo.blur_line(taps, xr)
    xd = _cl(x.to(DEV)).requires_grad_(True)
    y = ops.BlurFactor.apply(xd, taps.to(DEV), L, 1, 2)
    assert (y.cpu() - ref).abs().max() <= 2e-5
    r = torch.randn(ref.shape, generator=g)
    (ref * r).sum().backward()
    (y * r.to(DEV)).sum().backward()
    assert rel_err(xd.grad.cpu(), xr.grad) <= 2e-5


@pytest.mark.parametrize("m,n,k,act", [(1000, 64, 150, 1), (777, 27, 144, 0), (130, 3, 64, 2), (5, 144, 27, 0), (0, 8, 8, 0)])
def test_gemm_nt_and_tn(m, n, k, act):
    g = torch.Generator().manual_seed(2)
    ldx, ldy = (k + 3) & ~3, (n + 3) & ~3
    x = torch.zeros(max(m, 1), ldx)
    x[:, :k] = torch.randn(max(m, 1), k, generator=g)
    w = torch.randn(n, k, generator=g) * 0.2
    b = torch.randn(n, generator=g)
    ref = x[:m, :k] @ w.T + b
    ref = torch.relu(ref) if act == 1 else torch.sigmoid(ref) if act == 2 else ref
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    y = torch.full((max(m, 1), ldy), 7.0, device=DEV)
    cnt = torch.tensor([m], device=DEV, dtype=torch.int32)
    ops.gemm_nt(xd, ldx, wd, k, 0, bd, y, ldy, None, 0, cnt, max(m, 1) + 300, n, k, act)
    if m:
        assert (y[:m, :n].cpu() - ref).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))
        assert (y[:m, n:] == 0).all()
    # transposed-weight form: X @ W  with W given as [K,N]
    y2 = torch.zeros((max(m, 1), ldy), device=DEV)
    ops.gemm_nt(xd, ldx, wd.T.contiguous(), n, 1, None, y2, ldy, None, 0, cnt, max(m, 1), n, k, 0)
    if m:
        assert (y2[:m, :n].cpu() - x[:m, :k] @ w.T).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))
    # weight gradient
    dy = torch.zeros(max(m, 1), ldy)
    dy[:, :n] = torch.randn(max(m, 1), n, generator=g)
    dw = torch.zeros(n, k, device=DEV)
    db = torch.zeros(n, device=DEV)
    ops.gemm_tn(dy.to(DEV), ldy, xd, ldx, cnt, max(m, 1) + 77, n, k, dw, k, db)
    assert rel_err(dw.cpu(), dy[:m, :n].T @ x[:m, :k]) <= 1e-4 if m else float(dw.abs().max()) == 0
    assert rel_err(db.cpu(), dy[:m, :n].sum(0)) <= 1e-4 if m else float(db.abs().max()) == 0


@pytest.mark.parametrize("name", ["cubic_mlp", "noncubic_blur", "ndc_weakview"])
def test_feature_ops_match_oracle_including_out_of_range(name):
    g = load_golden(name)
    m = module_from_golden(g, DEV)
    field = vo.Field(aabb=g["aabb"], grid=list(g["case"]["grid"]),
                     params={k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()})
    fc = vo.grid_constants(field.aabb, field.grid, 0.5)
    gen = torch.Generator().manual_seed(3)
    u = torch.rand((501, 3), generator=gen) * 2.4 - 1.2          # some coordinates fall outside [-1,1]
    u[0] = torch.tensor([1.0, -1.0, 1.0])
    u[1] = torch.tensor([-1.0, -1.0, -1.0])
    taps = None
    if g["case"]["blur"] is not None:
        taps = vo.blur_taps(fc, "uniform-gaussian", 0.1, 64)
    ur = u.clone().requires_grad_(True)
    sref = vo.density_feature(field, fc, ur, taps)
    aref = vo.app_feature(field, fc, ur, taps)
    wa = torch.randn(aref.shape, generator=gen)
    ws = torch.randn(sref.shape, generator=gen)
    ((sref * ws).sum() + (aref * wa).sum()).backward()
    ud = u.to(DEV).requires_grad_(True)
    kd = taps.to(DEV) if taps is not None else None
    mode = "uniform-gaussian" if taps is not None else None
    s = m.compute_densityfeature(ud, kd, mode)
    a = m.compute_appfeature(ud, kd, mode)
    assert (s.cpu() - sref).abs().max() <= 1e-5 * max(1.0, float(sref.abs().max()))
    assert (a.cpu() - aref).abs().max() <= 1e-5 * max(1.0, float(aref.abs().max()))
    ((s * ws.to(DEV)).sum() + (a * wa.to(DEV)).sum()).backward()
    assert rel_err(ud.grad.cpu(), ur.grad) <= 1e-4
    for k in ("density_plane.0", "density_line.1", "app_plane.2", "app_line.0", "basis_mat.weight"):
        got = dict(m.named_parameters())[k].grad.cpu()
        assert rel_err(got, field.params[k].grad) <= 1e-4, k


def test_alpha_mask_lookup_matches_grid_sample():
    g = load_golden("alpha_mask")
    m = module_from_golden(g, DEV)
    gen = torch.Generator().manual_seed(5)
    xyz = torch.rand((20000, 3), generator=gen) * 3.4 - 1.7
    # exact lattice points too (integer indices, where the +1 corner has zero weight)
    lat = torch.stack(torch.meshgrid(torch.linspace(-1.5, 1.5, 24), torch.linspace(-1.5, 1.5, 22),
                                     torch.linspace(-1.5, 1.5, 20), indexing="ij"), -1).reshape(-1, 3)
    xyz = torch.cat([xyz, lat])
    ref = vo.alpha_mask_lookup(g["mask_volume"], g["aabb"], xyz) > 0
    got = m.alphaMask.sample_alpha(xyz.to(DEV)) > 0
    assert torch.equal(got.cpu(), ref), int((got.cpu() != ref).sum())


def test_compute_alpha_and_update_alpha_mask():
    g = load_golden("cubic_mlp")
    m = module_from_golden(g, DEV)
    field = vo.Field(aabb=g["aabb"], grid=list(g["case"]["grid"]), params=g["state_dict"])
    xyz = torch.rand((4000, 3), generator=torch.Generator().manual_seed(6)) * 3 - 1.5
    ref = vo.compute_alpha(field, xyz, float(m.stepSize))
    got = m.compute_alpha(xyz.to(DEV), m.stepSize)
    assert (got.cpu() - ref).abs().max() <= 1e-5
    with torch.no_grad():
        for i in range(3):
            m.density_plane[i].mul_(3.0)
            m.density_line[i].mul_(3.0)
    new_aabb = m.updateAlphaMask((24, 24, 24))
    assert m.alphaMask is not None and new_aabb.shape == (2, 3)
    m.shrink(new_aabb)
    rgb, _, _ = m.forward(__import__("gpu_common").default_opt(), g["rays_o"].to(DEV), g["rays_d"].to(DEV),
                          white_bg=True, is_train=False, N_samples=64)
    assert torch.isfinite(rgb).all()


def test_vector_red_accumulates_collisions():
    """All samples hit the same texel: the 16-byte RED scatter must sum, not overwrite."""
    m = jt.B200_VMSplit(torch.tensor([[-1.0] * 3, [1.0] * 3]), [8, 8, 8], DEV, density_n_comp=[4] * 3,
                        appearance_n_comp=[4] * 3, app_dim=27, shadingMode="SH", step_ratio=0.5)
    u = torch.zeros((4096, 3), device=DEV) + 0.1
    s = m.compute_densityfeature(u)
    s.sum().backward()
    one = m.compute_densityfeature(u[:1])
    g_all = m.density_plane[0].grad.clone()
    m.density_plane[0].grad = None
    one.sum().backward()
    assert rel_err(g_all, 4096 * m.density_plane[0].grad) <= 1e-4


def test_appearance_capacity_bound_and_tracker():
    """Memory model of the appearance stage: a bounded capacity truncates the appearance list safely and
    raises the device flag; the automatic tracker engages after a few training calls, sizes the stage from the
    observed counts, and a later overflow makes the next forward raise."""
    import joint_tensorf_b200 as jt
    from common import load_golden
    from gpu_common import default_opt, forward_kwargs, module_from_golden
    g = load_golden("cubic_mlp")
    m = module_from_golden(g, DEV)
    opt = default_opt("MLP_Fea")
    o, d = g["rays_o"].to(DEV), g["rays_d"].to(DEV)
    fkw = forward_kwargs(g, DEV)
    m.app_capacity = None
    og = o.clone().requires_grad_(True)
    ref = m.forward(opt, og, d, **fkw)
    a_full = int(jt.VMRender.last_counts[1].item())
    assert a_full > 1000 and int(jt.VMRender.last_app_used[1].item()) == 0
    # explicit bound below A: flagged, finite, backward runs, nothing is overrun
    m.app_capacity = a_full // 2
    og2 = o.clone().requires_grad_(True)
    out = m.forward(opt, og2, d, **fkw)
    used = jt.VMRender.last_app_used.tolist()
    assert used == [a_full // 2, 1] and int(jt.VMRender.last_counts[1].item()) == a_full
    out[0].sum().backward()
    assert torch.isfinite(og2.grad).all() and all(torch.isfinite(t).all() for t in out)
    assert not torch.equal(out[0], ref[0])
    # automatic tracking: exact results while it learns, then a bounded stage with identical results
    m.app_capacity = "auto"
    m._app_tracker.reset()
    for _ in range(m._app_tracker.warm + 3):
        og3 = o.clone().requires_grad_(True)
        out = m.forward(opt, og3, d, **fkw)
        torch.cuda.synchronize()
        assert torch.equal(out[0], ref[0])
    cap_now = m._app_tracker.capacity(o.shape[0], o.shape[0] * g["n_samples"])
    assert cap_now is not None and a_full <= cap_now <= o.shape[0] * g["n_samples"]
    # a sudden 50x denser request than the history allows: flagged, and the NEXT forward raises
    m._app_tracker.max_per_ray, m._app_tracker.floor = 1e-3, 16
    out = m.forward(opt, o.clone().requires_grad_(True), d, **fkw)
    torch.cuda.synchronize()
    assert int(jt.VMRender.last_app_used[1].item()) == 1
    with pytest.raises(jt._lib.JtError):
        m.forward(opt, o.clone().requires_grad_(True), d, **fkw)
    # ... after which the capacity is unbounded again
    out = m.forward(opt, o.clone().requires_grad_(True), d, **fkw)
    assert torch.equal(out[0], ref[0])
    # no-grad calls allocate exactly (host read of A) and reproduce the training colours
    with torch.no_grad():
        out_ng = m.forward(opt, o, d, **fkw)
    assert (out_ng[0] - ref[0]).abs().max() <= 1e-4      # the no-grad MLP_Fea head takes fp16 operand tiles


def test_cuda_graph_replay_of_a_training_step():
    """joint_tensorf_b200.graphs.GraphedStep: forward + loss + backward captured once and replayed reproduces the
    eagerly executed step (deterministic inputs: injected jitter, white background), and follows in-place updates of
    its static inputs."""
    import joint_tensorf_b200 as jt
    from common import load_golden
    from gpu_common import default_opt, forward_kwargs, module_from_golden
    g = load_golden("sh_vm48")
    m = module_from_golden(g, DEV)
    m.head_precision = "tc"
    m.app_capacity = None
    opt = default_opt("SH")
    fkw = forward_kwargs(g, DEV)
    o_s = g["rays_o"].to(DEV).clone().requires_grad_(True)
    d_s = g["rays_d"].to(DEV).clone()
    w = g["w_rgb"].to(DEV)
    params = list(m.parameters())

    def fn():
        for p in params + [o_s]:
            p.grad = None
        rgb, depth, acc = m.forward(opt, o_s, d_s, **fkw)
        loss = (rgb * w).sum()
        loss.backward()
        return loss.detach(), rgb.detach()

    loss_e, rgb_e = (t.clone() for t in fn())
    g_app_e, g_o_e = m.app_plane[1].grad.clone(), o_s.grad.clone()
    graphed = jt.graphs.GraphedStep(fn, warmup=2)
    loss_g, rgb_g = graphed()
    torch.cuda.synchronize()
    assert torch.equal(rgb_g, rgb_e) and torch.equal(loss_g, loss_e)
    assert rel_err(m.app_plane[1].grad, g_app_e) <= 1e-5 and rel_err(o_s.grad, g_o_e) <= 1e-5     # atomics reorder sums
    # new ray directions through the same static buffer
    with torch.no_grad():
        d_s.copy_(d_s.flip(0))
    loss_g2, rgb_g2 = (t.clone() for t in graphed())
    loss_e2, rgb_e2 = fn()
    torch.cuda.synchronize()
    assert torch.equal(rgb_g2, rgb_e2) and not torch.equal(rgb_g2, rgb_e)
