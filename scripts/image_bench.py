"""Timing of the 2-D supervision kernels on Blender-sized data (100 views, 800 x 800, 201 taps)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from joint_tensorf_b200 import supervision as sv  # noqa: E402
from joint_tensorf_b200.vmsplit import gaussian_taps  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = "cuda:0"
    images = torch.rand((100, 3, 800, 800), device=dev)
    taps = gaussian_taps(20.0, 201)
    out = {}
    ms = timeit(lambda: sv.image_blur(images, taps))
    out["image_blur_100x3x800x800_201taps_ms"] = ms
    out["image_blur_tflops"] = 2 * 2 * 201 * images.numel() / ms / 1e9
    out["image_blur_hbm_gbs"] = 4 * 4 * images.numel() / ms / 1e6
    out["edge_mask_hard_ms"] = timeit(lambda: sv.edge_mask(images, False, 1.25))
    # the reference's own op sequence with stock ATen kernels on this GPU (nerf.py:98-110)
    import torch.nn.functional as F

    def aten():
        k = taps.to(dev).expand(1, 1, -1)
        x = images.reshape(300, 800, 800)
        x = F.pad(x, (100, 100), mode="replicate")
        x = F.conv1d(x, k.expand(800, 1, -1), groups=800).permute(0, 2, 1)
        x = F.pad(x, (100, 100), mode="replicate")
        x = F.conv1d(x, k.expand(800, 1, -1), groups=800)
        return x.permute(0, 2, 1).reshape(100, 3, 800, 800).contiguous()
    out["aten_gpu_image_blur_ms"] = timeit(aten, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
