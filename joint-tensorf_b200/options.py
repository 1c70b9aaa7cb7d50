"""The slice of the reference's `opt` namespace that the field layer reads on every
forward (batBase.py:46-62, tensorBase.py:581): a plain attribute dict, so that
callers without the reference's options.py/EasyDict can drive the module."""


class Namespace(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return Namespace(v) if isinstance(v, dict) and not isinstance(v, Namespace) else v

    def __setattr__(self, k, v):
        self[k] = v


def default_opt(shading="MLP_Fea", ndc=False):
    return Namespace(
        arch=dict(abs_components=False, component_wise_feature2density=False, plane_feature2density=False,
                  convolve_plane_only=False, convolve_positive_only=False, ignore_negative_split=False,
                  ndc_near_plane=1.0, shading=dict(model=shading, detach_viewdirs=True, detach_xyz=True),
                  tensorf=dict(grid_sample_interp_mode="bilinear")),
        camera=dict(ndc=ndc, ndc_simulate_euclid_sample=False, ndc_simulate_euclid_depth=False), nerf=dict())
