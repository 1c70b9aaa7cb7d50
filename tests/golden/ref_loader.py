"""Import the live reference (Nemo1999/Joint-TensoRF) field module on CPU.

Used ONLY by `tests/golden/make_golden.py` (run in the build container, where
`/root/reference` is mounted) and by the optional live-reference test. Nothing
that runs on the GPU box touches this file: `/root/reference` does not exist
there.

The reference's field layer (`model/tensorf_repr`) imports four small packages
that are not installed in this image (`icecream`, `ipdb`, `termcolor`,
`easydict`); they are only used for debug printing / attribute dicts, so we
install inert stand-ins into `sys.modules` before importing it.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JT_REFERENCE_ROOT", "/root/reference")


class AttrDict(dict):
    """dict with attribute access, recursive on nested dicts (easydict stand-in)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _install_shims():
    if "icecream" not in sys.modules:
        m = types.ModuleType("icecream")

        class _IC:
            def __call__(self, *a, **k):
                return a[0] if len(a) == 1 else a

            def enable(self):
                pass

            def disable(self):
                pass

            def configureOutput(self, *a, **k):
                pass

        m.ic = _IC()
        sys.modules["icecream"] = m
    if "ipdb" not in sys.modules:
        m = types.ModuleType("ipdb")
        m.set_trace = lambda *a, **k: None
        sys.modules["ipdb"] = m
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.colored = lambda s, *a, **k: str(s)
        sys.modules["termcolor"] = m
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = AttrDict
        sys.modules["easydict"] = m


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model", "tensorf_repr"))


def load():
    """Returns (tensorf_repr module, camera module). Raises if reference absent."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from model import tensorf_repr  # noqa
        import camera  # noqa
    return tensorf_repr, camera


def default_opt(shading="MLP_Fea", ndc=False):
    """The `opt` fields the field layer reads on every forward
    (reference batBase.py:46-62, tensorBase.py:581)."""
    return AttrDict(
        arch=dict(
            abs_components=False,
            component_wise_feature2density=False,
            plane_feature2density=False,
            convolve_plane_only=False,
            convolve_positive_only=False,
            ignore_negative_split=False,
            ndc_near_plane=1.0,
            shading=dict(model=shading, detach_viewdirs=True, detach_xyz=True),
            tensorf=dict(grid_sample_interp_mode="bilinear"),
        ),
        camera=dict(ndc=ndc, ndc_simulate_euclid_sample=False, ndc_simulate_euclid_depth=False),
        nerf=dict(),
        device="cpu",
    )
