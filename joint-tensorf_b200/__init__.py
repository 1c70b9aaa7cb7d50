"""joint-tensorf_b200: B200-native (sm_100a) TensoRF-VM volume-rendering hot path.

Public surface (mirrors reference `model.tensorf_repr`):
    B200_VMSplit   -- drop-in for BAT_VMSplit / TensorVMSplit
    AlphaGridMask  -- occupancy mask container
    camera.get_center_and_ray -- sparse pose -> ray generation with backward to se(3) (section 8f-1)
Lower level: `ops` (C-ABI wrappers), `render.VMRender` (fused autograd node),
`synth` (deterministic synthetic scenes/rays), `parallel` (ray-sharded data parallel).
"""
from . import _lib, camera, graphs, ops, options, synth  # noqa: F401
from .render import RenderCfg, VMRender  # noqa: F401
from .vmsplit import AlphaGridMask, B200_VMSplit  # noqa: F401

__all__ = ["B200_VMSplit", "AlphaGridMask", "VMRender", "RenderCfg", "ops", "synth"]
