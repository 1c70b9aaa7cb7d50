"""CUDA-graph replay of a training step.

One step of the hot path is ~45 kernel launches issued through ctypes from Python: 1.4 ms of host time against 2.8 ms
of device time at 4096 rays x 1000 samples (fine), but the same 1.4 ms against 0.6 ms when the batch is small or
strong-scaled over 8 GPUs (512 rays per GPU) -- the step becomes launch-bound. The kernels never synchronise, read
their sample counts from device memory and take every size bound from the host-side ray count, so the whole step
(pose -> rays -> forward -> loss -> backward, including the NCCL all-reduces of a data-parallel step) can be captured
once and replayed: `GraphedStep(fn)` warms `fn` up on a side stream, captures it into a `torch.cuda.CUDAGraph` and
replays it on every call.

Rules for `fn` (the usual CUDA-graph contract):
  * its inputs are STATIC tensors the caller refreshes in place (`pix.copy_(...)`) before each call; its return
    value(s) are static tensors overwritten by every replay (read them after the call);
  * host-side decisions are baked in at capture time: blur taps / schedule values, the background coin flip of
    batBase.py:154 (pass `bg_coin=` or use `white_bg=True`), `near_far`;
  * the module must not synchronise: `app_capacity` must be None or a fixed number (the automatic tracker polls
    events), and no-grad calls (which read the appearance count on the host) cannot be captured;
  * device-side random numbers (the stratified jitter, `torch.rand(..., device=cuda)`) advance on every replay.
"""
import torch


class GraphedStep:
    def __init__(self, fn, warmup=3, pool=None):
        self.fn = fn
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
