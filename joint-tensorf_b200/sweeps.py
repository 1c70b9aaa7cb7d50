"""Per-step full-factor sweeps: regularisers + optimiser update.

The reference evaluates `density_L1`, `TV_loss_density`, `TV_loss_app` (tensoRF.py:212-228 with
`TVLoss`, tensorBase.py:16-41) on every training step (model/tensorf.py:126-130) and then steps
`torch.optim.Adam(grad_vars, betas=(0.9, 0.99))` (tensorf.py:474-475) with a per-step learning-rate
decay (tensorf.py:431-436). Here each is one streaming kernel of csrc/field_sweep.cu over the
channel-last factor storage. No CPU / ATen fallback: CUDA tensors only.
"""
import torch

from . import _lib
from ._lib import check, doubles, floats, ints, longlongs, ptrs
from .ops import TIMER, _need_cuda, _p, _stream, phys_cl


def _meta(arrays, terms, l1s):
    dims = []
    for a in arrays:
        assert a.dim() == 3 and a.is_contiguous() and a.dtype == torch.float32, (a.shape, a.stride(), a.dtype)
        dims += list(a.shape)
    return ints(dims), ints(terms), ints([1 if v else 0 for v in l1s])


def reg_values(arrays, terms, l1s):
    """arrays: channel-last [H,W,C] fp32 CUDA buffers -> tensor [3] = (L1, TV_density, TV_app)."""
    for a in arrays:
        _need_cuda(a, "factor")
    dev = arrays[0].device
    out = torch.empty((3,), device=dev, dtype=torch.float32)
    ws = torch.empty((36,), device=dev, dtype=torch.float64)
    dims, tm, l1 = _meta(arrays, terms, l1s)
    with TIMER.span("reg_values"):
        check(_lib.lib().jt_reg_values(len(arrays), ptrs([a.data_ptr() for a in arrays]), dims, tm, l1, _p(ws),
                                       _p(out), _stream()), "jt_reg_values")
    return out


def reg_grads_(arrays, grads, terms, l1s, coef3, up3=None):
    """grads[i] += d/d arrays[i] (coef3 . (L1, TV_density, TV_app)); in place, one launch."""
    for a, g in zip(arrays, grads):
        _need_cuda(a, "factor")
        _need_cuda(g, "gradient")
        assert g.shape == a.shape and g.is_contiguous() and g.dtype == torch.float32
    dims, tm, l1 = _meta(arrays, terms, l1s)
    with TIMER.span("reg_grads"):
        check(_lib.lib().jt_reg_grads(len(arrays), ptrs([a.data_ptr() for a in arrays]),
                                      ptrs([g.data_ptr() for g in grads]), dims, tm, l1, floats(coef3), _p(up3),
                                      _stream()), "jt_reg_grads")


class FieldRegularizers(torch.autograd.Function):
    """(L1, TV_density, TV_app) of a field as ONE autograd node over its factors.

    forward: one sweep (every factor element read once); backward: one sweep that writes
    d/dx (g . values) for every factor. `terms[i]` / `l1s[i]` as in jt_reg_values."""

    @staticmethod
    def forward(ctx, terms, l1s, *factors):
        xs = [phys_cl(x.detach()) for x in factors]
        ctx.xs, ctx.terms, ctx.l1s = xs, list(terms), list(l1s)
        return reg_values(xs, terms, l1s)

    @staticmethod
    def backward(ctx, g):
        xs = ctx.xs
        need = ctx.needs_input_grad[2:]
        idx = [i for i in range(len(xs)) if need[i] and (ctx.terms[i] or ctx.l1s[i])]
        out = [None] * len(xs)
        if idx:
            gs = [torch.zeros_like(xs[i]) for i in idx]
            reg_grads_([xs[i] for i in idx], gs, [ctx.terms[i] for i in idx], [ctx.l1s[i] for i in idx],
                       (1.0, 1.0, 1.0), g.contiguous().float())
            for i, gx in zip(idx, gs):
                out[i] = gx.unsqueeze(0).permute(0, 3, 1, 2)
        return (None, None, *out)


def _dense_flat(t):
    """1-D view over the dense storage of `t` (contiguous or channel-last); None if not dense."""
    if t.is_contiguous():
        return t.view(-1)
    if t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous():
        return t.permute(0, 2, 3, 1).reshape(-1)
    return None


class FusedAdam(torch.optim.Optimizer):
    """`torch.optim.Adam` (no weight decay, no amsgrad) as one multi-tensor kernel launch per step.

    Same constructor arguments, `param_groups` (the reference rescales `param_group['lr']` every step,
    tensorf.py:433-434) and `state_dict` layout (`step`, `exp_avg`, `exp_avg_sq`) as torch's, so the
    reference's optimiser save/restore code works unchanged. Extras: `grad_scale` (multiplied into the
    gradients, e.g. 1/world_size after a summing all-reduce) and `zero_grad_in_step` (the gradient
    buffers are cleared by the same pass, replacing the separate `optimizer.zero_grad()` sweep)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 zero_grad_in_step=False):
        if weight_decay != 0 or amsgrad:
            raise _lib.JtError("FusedAdam implements the reference's configuration: weight_decay=0, amsgrad=False")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self.zero_grad_in_step = zero_grad_in_step
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # one launch per distinct (betas, eps); a field has one
        buckets = {}
        for group in self.param_groups:
            key = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            for p in group["params"]:
                if p.grad is None:
                    continue
                _need_cuda(p, "parameter")
                if p.dtype != torch.float32 or p.grad.is_sparse:
                    raise _lib.JtError("FusedAdam: dense fp32 parameters only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                pf = _dense_flat(p)
                if pf is None:
                    raise _lib.JtError("FusedAdam: parameter storage must be dense (contiguous or channel-last)")
                g = p.grad
                if g.stride() != p.stride():
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                    if self.zero_grad_in_step:
                        p.grad.zero_()
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if m.stride() != p.stride() or v.stride() != p.stride():    # state loaded from a foreign layout
                    m = st["exp_avg"] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(m)
                    v = st["exp_avg_sq"] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(v)
                buckets.setdefault(key, []).append((pf, _dense_flat(g), _dense_flat(m), _dense_flat(v),
                                                    float(group["lr"]), int(st["step"]), p))
        for (b1, b2, eps), items in buckets.items():
            with TIMER.span("adam"):
                check(_lib.lib().jt_adam_multi(
                    len(items), ptrs([t[0].data_ptr() for t in items]), ptrs([t[1].data_ptr() for t in items]),
                    ptrs([t[2].data_ptr() for t in items]), ptrs([t[3].data_ptr() for t in items]),
                    longlongs([t[0].numel() for t in items]), doubles([t[4] for t in items]),
                    ints([t[5] for t in items]), b1, b2, eps, float(self.grad_scale),
                    int(self.zero_grad_in_step), _stream()), "jt_adam_multi")
            for t in items:       # the kernel wrote through raw pointers: tell autograd the parameters changed
                torch.autograd.graph.increment_version(t[6])
        return loss
