"""The UNMODIFIED reference on the host cores (TEST / BASELINE INFRASTRUCTURE, never the product).

`bench.py --impl reference` and the `cpu_baseline` leg time the reference's own
`BAT_VMSplit.forward` + backward (model/tensorf_repr/batBase.py:44-165) on CPU. The
reference is pure Python: it is not pip-installable (no setup.py / pyproject) and
`/root/reference` does not exist on the GPU box, so `stage()` -- called by
`__graft_entry__.build()` in the build container, where the reference is mounted --
copies the handful of files the field layer imports (model/tensorf_repr/*.py,
model/kernels.py, util.py, camera.py; verbatim, unmodified) into `oracle/_ref/`.
That directory is git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels to the GPU box like a built `.so`. When neither
`/root/reference` nor `oracle/_ref` is present, the callers fall back to the
restatement `oracle/vm_oracle.py` (kind "port").

Only tests/, smoke() and bench.py's baseline legs may import this module.
"""
import contextlib
import io
import os
import shutil
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(HERE, "_ref")
LIVE = os.environ.get("JT_REFERENCE_ROOT", "/root/reference")
FILES = ("util.py", "camera.py", "model/kernels.py", "model/tensorf_repr/__init__.py", "model/tensorf_repr/batBase.py",
         "model/tensorf_repr/bateRF.py", "model/tensorf_repr/sh.py", "model/tensorf_repr/tensoRF.py",
         "model/tensorf_repr/tensorBase.py")


def stage():
    """Copy the reference's field-layer files (verbatim) into oracle/_ref. No-op without /root/reference."""
    if not os.path.isdir(os.path.join(LIVE, "model", "tensorf_repr")):
        return False
    for f in FILES:
        dst = os.path.join(STAGED, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(LIVE, f), dst)
    with open(os.path.join(STAGED, "STAGED_FROM"), "w") as fh:
        fh.write(f"verbatim copies of {', '.join(FILES)} from {LIVE} (git-ignored; made by oracle/reference_arm.stage)\n")
    return True


def root():
    for r in (LIVE, STAGED):
        if os.path.isfile(os.path.join(r, "model", "tensorf_repr", "bateRF.py")):
            return r
    return None


def load():
    """(tensorf_repr module, default_opt factory) of the unmodified reference, or None when it is absent."""
    r = root()
    if r is None:
        return None
    gold = os.path.join(ROOT, "tests", "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    import ref_loader
    ref_loader.REFERENCE_ROOT = r
    tr, _ = ref_loader.load()
    return tr, ref_loader.default_opt


def build_field(workload):
    """The reference `BAT_VMSplit` for a bench workload (same constructor keywords as the B200 module)."""
    from joint_tensorf_b200 import synth
    tr, default_opt = load()
    kw, run = synth.config(workload)
    kw = dict(kw)
    aabb, grid = torch.tensor(kw.pop("aabb")), kw.pop("gridSize")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = tr.BAT_VMSplit(aabb, list(grid), "cpu", dtype=torch.float32, **kw)
    if kw["shadingMode"] == "SH":
        # the reference's own SHRender call site passes 5 arguments to a 3-argument function (tensorBase.py:68 vs
        # batBase.py:137): wrap the unmodified function so that forward runs at all
        from model.tensorf_repr import tensorBase as _tb
        m.renderModule = lambda p, v, f, *_: _tb.SHRender(p, v, f)
    return m, default_opt(kw["shadingMode"], run["ndc"]), run


def time_reference(workload, n_rays, steps, warmup, blur=0.0, budget_s=240.0):
    """rays/s of reference forward + MSE + backward on all host cores. Stops early when `budget_s` is spent;
    returns (rays_per_s, sec_per_step, cores, steps_done, warmup_done)."""
    from joint_tensorf_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m, opt, run = build_field(workload)
    if run["ndc"]:
        o, d, _ = synth.llff_ndc_rays(n_rays, 8, seed=1)
    else:
        o, d, _ = synth.blender_rays(n_rays, max(1, min(32, n_rays // 16)), seed=1)
    target = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(2))
    fkw = dict(white_bg=run["white_bg"], is_train=True, ndc_ray=run["ndc"], N_samples=run["n_samples"])
    if blur > 0:
        fkw.update(c2f_mode="uniform-gaussian", c2f_parameter_density=blur * 0.6, c2f_parameter_color=blur,
                   c2f_kernel_size=64)
    t_start = time.perf_counter()
    times, warm_done = [], 0
    for it in range(warmup + steps):
        timed = it >= warmup
        spent = time.perf_counter() - t_start
        if not timed and warm_done >= 1 and spent > 0.25 * budget_s:
            continue                                   # out of warm-up budget: go straight to the timed steps
        if timed and times and spent > budget_s:
            break
        oc, dc = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
        for p in m.parameters():
            p.grad = None
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            rgb, _, _ = m.forward(opt, oc, dc, **fkw)
        loss = ((rgb - target) ** 2).mean()
        loss.backward()
        dt = time.perf_counter() - t0
        if timed:
            times.append(dt)
        else:
            warm_done += 1
    sec = sum(times) / len(times)
    return n_rays / sec, sec, cores, len(times), warm_done
