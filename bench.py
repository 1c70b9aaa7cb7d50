#!/usr/bin/env python
"""bench.py -- rays/s forward+backward of the TensoRF-VM render hot path.

    python bench.py --gpus N --steps K --warmup W            (own arm, B200 kernels)
    python bench.py --impl reference --gpus N ...            (reference arm: CPU port)

Workload (BASELINE.json configs[1], "cfg2_sh"): TensoRF-VM 300^3 grid, density 3x16 /
appearance 3x48 components, app_dim 27, SH shading, 4096-ray batch, S=1000 samples per
ray, forward + backward to all factor / basis / ray gradients. Synthetic Blender-shaped
rays and random-init factors (joint_tensorf_b200.synth). The same field with the
MLP_Fea shading head (`--workload cfg2`, the head of the BAT VM_MLP configs) is timed in
the same run and reported under "also".

A step = one training iteration without the optimizer (SURVEY.md section 8d timing
protocol): se3_refine + fixed poses -> rays of the sampled pixels (jt_pose_rays_fwd) ->
stratified jitter -> forward -> MSE -> backward to all factor / head / se3_refine
gradients (+ one NCCL all-reduce of the flat gradient bucket when N > 1).
N > 1 is weak scaling: every rank renders its own 4096 rays.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rays/sec fwd+bwd (300^3 VM, 4096-ray batch)"
UNIT = "rays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_sh",
                    help="cfg2_sh = BASELINE.json configs[1] (300^3, 3x16 / 3x48 comps, app_dim 27, SH shading); "
                         "cfg2 = same field with the MLP_Fea head; cfg1 = 128^3; cfg4 = LLFF NDC 617x687x617")
    ap.add_argument("--no-also", action="store_true",
                    help="skip the second head variant of the 300^3 field (cfg2 <-> cfg2_sh) timed after the headline")
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--blur", type=float, default=0.0, help="c2f blur parameter (0 = off, cfg3 uses 0.15)")
    ap.add_argument("--cpu-rays", type=int, default=2048,
                    help="rays in the bounded CPU-baseline sample (2 timed fwd+bwd iterations: ~10-20 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the 800x800 full-frame inference timing")
    ap.add_argument("--render-chunk", type=int, default=16384, help="rays per forward call of the full-frame render")
    ap.add_argument("--head", default="auto", choices=["auto", "tc", "fp32"],
                    help="shading head: tcgen05 tensor-core kernels (auto/tc) or strict-fp32 SIMT GEMMs")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled (NVML, nvidia-smi as fallback) while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES remapping via the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self.h = h
                        break
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        flags = []
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                          ("sw_power_cap", 0x4)):
            flags.append("Active" if (r & bit) else "Not Active")
        return [str(sm), str(mx), "0"] + flags

    def run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.005 if self.nvml is not None else 0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------- CPU port (oracle) timing
def oracle_field(workload, seed=0):
    from joint_tensorf_b200 import synth
    from oracle import vm_oracle as vo
    kw, run = synth.config(workload)
    params = vo.init_params(kw["gridSize"], kw["density_n_comp"], kw["appearance_n_comp"], kw["app_dim"],
                            kw["shadingMode"], kw["featureC"], kw["view_pe"], kw["fea_pe"],
                            kw["volume_init_scale"], kw["volume_init_bias"], seed=seed)
    for v in params.values():
        v.requires_grad_(True)
    field = vo.Field(aabb=torch.tensor(kw["aabb"]), grid=kw["gridSize"], params=params, near_far=kw["near_far"],
                     step_ratio=kw["step_ratio"], density_shift=float(kw["density_shift"]),
                     distance_scale=kw["distance_scale"], weight_thres=kw["rayMarch_weight_thres"],
                     act=kw["fea2denseAct"], shading=kw["shadingMode"], view_pe=kw["view_pe"], fea_pe=kw["fea_pe"])
    return field, run


def time_cpu_port(workload, n_rays, steps, warmup, blur):
    """Reference algorithm (oracle port, same ATen CPU operators as the reference) on
    the host cores: forward + MSE + backward on a bounded sample of the workload."""
    from joint_tensorf_b200 import synth
    from oracle import vm_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    field, run = oracle_field(workload)
    o, d, _ = synth.blender_rays(n_rays, max(1, min(32, n_rays // 16)), seed=1)
    target = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(2))
    kw = dict(n_samples=run["n_samples"], white_bg=run["white_bg"], ndc=run["ndc"])
    if blur > 0:
        kw.update(blur_mode="uniform-gaussian", blur_density=blur * 0.6, blur_color=blur, kernel_size=64)
    times = []
    for it in range(warmup + steps):
        oc, dc = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
        jit = torch.rand(n_rays, 1)
        for p in field.params.values():
            p.grad = None
        t0 = time.perf_counter()
        rgb, _, _ = vo.render(field, oc, dc, jitter=jit, **kw)
        loss = ((rgb - target) ** 2).mean()
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return n_rays / sec, sec, cores


def time_aten_gpu(workload, n_rays, steps, warmup, dev):
    """Same-box bar: the reference ALGORITHM with stock ATen CUDA operators on this B200 -- the oracle
    port run with its tensors on the device (grid_sample / cumprod / addmm kernels, autograd backward). This is
    what the unmodified reference would execute on a GPU; reported next to the CPU baseline, never shipped."""
    from joint_tensorf_b200 import synth
    from oracle import vm_oracle as vo
    field, run = oracle_field(workload)
    field.params = {k: v.detach().to(dev).requires_grad_(True) for k, v in field.params.items()}
    field.aabb = field.aabb.to(dev)
    o, d, _ = synth.blender_rays(n_rays, 32, seed=1)
    o, d = o.to(dev), d.to(dev)
    target = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(2)).to(dev)
    kw = dict(n_samples=run["n_samples"], white_bg=run["white_bg"], ndc=run["ndc"])
    times = []
    with torch.device(dev):
        for it in range(warmup + steps):
            oc, dc = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
            jit = torch.rand(n_rays, 1)
            for p in field.params.values():
                p.grad = None
            torch.cuda.synchronize()
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ev.record()
            rgb, _, _ = vo.render(field, oc, dc, jitter=jit, **kw)
            loss = ((rgb - target) ** 2).mean()
            loss.backward()
            e_ev.record()
            torch.cuda.synchronize()
            if it >= warmup:
                times.append(s_ev.elapsed_time(e_ev) * 1e-3)
    sec = sum(times) / len(times)
    return n_rays / sec, sec


def synth_describe(workload, n_rays=None):
    from joint_tensorf_b200 import synth
    return synth.describe(workload, n_rays)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_rays
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    rps, sec, cores = time_cpu_port(args.workload, n, steps, warm, args.blur)
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": synth_describe(args.workload) + f", {n}-ray sample of the {args.rays}-ray batch, fwd+bwd",
                   "blur": args.blur},
        "cpu_baseline": {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} of 4096 rays per step, {steps} steps, oracle/vm_oracle.py (torch CPU, "
                                   f"{cores} threads)"},
        "e2e": {"value": rps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- own arm
def algorithmic_bytes(name, V, A, cd, ca, ctot_a):
    """bytes a kernel must move per launch (fp32 factors)."""
    s = 4
    dens_f = V * (18 * cd * s + 16)                 # 3 x (4+2) taps x C_d x 4 B + coords in + feature out
    app_f = A * (18 * ca * s + 12 + 4 * ctot_a)     # taps + coords + component row out (un-fused)
    app_fused = A * (18 * ca * s + 12)              # taps + coords; the component row stays on chip
    table = {
        "vm_density_fwd": dens_f,
        "vm_app_fwd": app_f,
        "app_basis_fwd_tc": app_fused + A * 128,    # + feat/dir row out
        "app_basis_sh_fwd_tc": app_fused + A * 32,  # + rgb and view direction out
        "vm_density_bwd": V * (3 * 18 * cd * s + 4 + 16),
        "vm_app_bwd": A * (3 * 18 * ca * s + 4 * ctot_a + 16),
    }
    return table.get(name)


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of these
# workloads (profiles/r01_ncu_full_v14.txt for the SH path and the gather / scatter kernels, r01_ncu_full_v4.txt for
# the MLP_Fea head kernels; cold caches, so an upper bound on a warm step)
NCU_DRAM_BYTES = {
    "vm_app_bwd": 1.923781e9 + 0.330893e9, "vm_density_bwd": 0.088398e9 + 0.003505e9,
    "vm_density_fwd": 0.052985e9 + 0.001150e9, "app_basis_sh_fwd_tc": 0.580590e9 + 0.725284e9,
    "sh_bwd_tc": 0.322786e9 + 0.724659e9 + 0.784006e9 + 0.003986e9,        # data kernel + basis weight-gradient kernel
    "app_basis_fwd_tc": 0.453102e9 + 0.899889e9, "head_mlp_fwd_tc": 0.299442e9 + 1.428690e9,
    "head_bwd_tc": 0.906751e9 + 1.360633e9 + 2.863843e9 + 0.003652e9,      # data kernel + weight-gradient kernel
}


def gemm_flops(name, A, F, ctot, in_dim, H):
    table = {
        "basis_fwd": 2 * A * F * ctot, "basis_bwd_x": 2 * A * F * ctot, "basis_bwd_w": 2 * A * F * ctot,
        "mlp_l1_fwd": 2 * A * in_dim * H, "mlp_l1_bwd_x": 2 * A * in_dim * H, "mlp_l1_bwd_w": 2 * A * in_dim * H,
        "mlp_l2_fwd": 2 * A * H * H, "mlp_l2_bwd_x": 2 * A * H * H, "mlp_l2_bwd_w": 2 * A * H * H,
        "mlp_l3_fwd": 2 * A * H * 3, "mlp_l3_bwd_x": 2 * A * H * 3, "mlp_l3_bwd_w": 2 * A * H * 3,
        # fused tensor-core head: basis + 3 layers forward; data + weight gradients backward
        "head_fwd_tc": 2 * A * (F * ctot + in_dim * H + H * H + H * 3),
        "head_bwd_tc": 4 * A * (F * ctot + in_dim * H + H * H + H * 3),
        "head_mlp_fwd_tc": 2 * A * (in_dim * H + H * H + H * 3),
        "sh_bwd_tc": 4 * A * F * ctot,
    }
    return table.get(name)


def own_arm(args):
    import torch.distributed as dist

    import joint_tensorf_b200 as jt
    from joint_tensorf_b200 import ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured" if peaks else "fallback"

    kw, run = jt.synth.config(args.workload)
    kw = dict(kw)
    torch.manual_seed(0)                       # identical replicas on every rank
    model = jt.B200_VMSplit(torch.tensor(kw.pop("aabb")), kw.pop("gridSize"), dev, **kw)
    from joint_tensorf_b200.options import default_opt
    model.head_precision = args.head
    opt = default_opt(model.shadingMode, run["ndc"])
    N, S = args.rays, run["n_samples"]
    # Pose side of the step (SURVEY 8d protocol): 32 hemisphere views, pose noise N(0, 0.15^2) composed into the
    # fixed pose (bat.py:34,346-348), se3_refine = 0 and trainable (bat.py:350), 128 pixels shared by all views
    # (nerf.py:657-658) -> N = 4096 rays generated by jt_pose_rays_fwd, gradients back to se3_refine.
    # LLFF (cfg4): 8 forward-facing views = small rigid motions of the identity, 1008x756, focal 0.85 W, rays
    # converted to NDC (camera.py:303-340) inside the same kernel.
    ndc = bool(run["ndc"])
    n_views = 8 if ndc else 32
    H_img, W_img = (756, 1008) if ndc else (800, 800)
    R_pix = max(1, N // n_views)
    N = n_views * R_pix
    if ndc:
        gt_pose, intr = jt.synth.llff_views(n_views, (H_img, W_img), seed=1 + rank)
        base_pose = gt_pose.to(dev)
    else:
        gt_pose, intr = jt.synth.blender_views(n_views, (H_img, W_img), seed=1 + rank)
        g_noise = torch.Generator().manual_seed(50 + rank)
        pose_noise = 0.15 * torch.randn((n_views, 6), generator=g_noise)
        base_pose = jt.camera.refined_pose(pose_noise.to(dev), gt_pose.to(dev))
    intr_inv_d = intr.inverse().to(dev)
    intr_d = intr.to(dev) if ndc else None       # convert_NDC reads the focal length / principal point
    se3_refine = torch.nn.Parameter(torch.zeros((n_views, 6), device=dev))
    cam_opt = jt.options.Namespace(H=H_img, W=W_img, camera=dict(model="perspective", ndc=ndc),
                                   arch=dict(ndc_near_plane=1.0) if ndc else dict())
    pix_h = torch.randperm(H_img * W_img, generator=torch.Generator().manual_seed(7 + rank))[:R_pix].to(torch.int32)
    tgt_h = torch.rand(N, 3, generator=torch.Generator().manual_seed(100 + rank))
    pix_h, tgt_h = pix_h.pin_memory(), tgt_h.pin_memory()
    pix_d, tgt_d = pix_h.to(dev), tgt_h.to(dev)
    loss_h = torch.empty((), pin_memory=True)
    fkw = dict(white_bg=run["white_bg"], is_train=True, ndc_ray=run["ndc"], N_samples=S)
    if args.blur > 0:
        fkw.update(c2f_mode="uniform-gaussian", c2f_parameter_density=args.blur * 0.6,
                   c2f_parameter_color=args.blur, c2f_kernel_size=64)
    params = [p for p in model.parameters()] + [se3_refine]
    # data parallel: the render node's backward reduces its flat gradient bucket across ranks itself (appearance
    # part overlapped with the density scatter); the loss carries the 1/world factor, so the sums are means
    sync = parallel.OverlappedGradSync() if world > 1 else None
    model.grad_sync = sync
    inv_world = 1.0 / world

    def make_step(model, opt, params):
        def step(pix, tgt, reduce=True):
            for p in params:
                p.grad = None
            model.grad_sync = sync if reduce else None
            center, ray = jt.camera.get_center_and_ray(cam_opt, base_pose, intr_inv_d, ray_idx=pix, intr=intr_d,
                                                       se3_refine=se3_refine)
            rgb, depth, acc = model(opt, center.view(-1, 3), ray.view(-1, 3), **fkw)
            loss = ((rgb - tgt) ** 2).mean()
            (loss * inv_world if world > 1 else loss).backward()
            if sync is not None and reduce:
                sync.finish([se3_refine])
            return loss
        return step

    step = make_step(model, opt, params)

    def step_e2e():
        pix = pix_h.to(dev, non_blocking=True)
        t = tgt_h.to(dev, non_blocking=True)
        loss = step(pix, t)
        loss_h.copy_(loss.detach(), non_blocking=True)

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)          # > 126 MB L2

    def timed(fn, k):
        """Sum of the K per-step device times (CUDA events on the launching stream). The L2
        flush between steps is on the stream but outside every event pair; the host does not
        synchronise inside the loop, so it queues launches ahead as a training loop does."""
        evs = []
        t_host = time.perf_counter()
        # The timed region starts right after a device synchronize: without queued work the GPU would execute the
        # first step's ~45 launches as fast as the host can issue them (3.2 ms instead of 2.7 ms, `step_ms_rank0` of
        # profiles/r01_bench_v15.json). A few extra untimed L2-flush writes give the host the head start it has in
        # every later step; nothing inside an event pair changes.
        for _ in range(6):
            flush.fill_(1.0)
        for _ in range(k):
            flush.fill_(1.0)                                           # evict L2 between steps (untimed)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        timed.host_ms = (time.perf_counter() - t_host) * 1e3 / k      # host time to ENQUEUE a step (launch-bound check)
        torch.cuda.synchronize()
        timed.per_step = [round(s.elapsed_time(e), 3) for s, e in evs]
        return sum(s.elapsed_time(e) for s, e in evs)

    tok = torch.zeros(1, device=dev)

    def barrier():
        """Host barrier + device synchronize (the contract's bracket), then -- multi-GPU only -- a one-element
        all-reduce left on the stream: a DEVICE-side rendezvous, so that the ranks' first timed step starts together
        even though their hosts leave dist.barrier() a few ms apart (measured at N=8: the first timed step took
        7.8 ms instead of 3.65 ms because rank 0 waited for the last host inside its first gradient all-reduce)."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(tok)

    for _ in range(max(args.warmup, 3)):
        step(pix_d, tgt_d)
    barrier()
    for _ in range(2):
        step_e2e()
    barrier()
    l0 = jt._lib.launch_count()
    with ClockSampler(local) as clk:             # clocks are sampled across both timed regions
        ms = timed(lambda: step(pix_d, tgt_d), args.steps)
        host_ms, per_step = timed.host_ms, timed.per_step
        barrier()
        launches = jt._lib.launch_count() - l0
        ms_e2e = timed(step_e2e, args.steps)
        barrier()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    roof, breakdown = None, None
    if rank == 0 and not args.no_breakdown:
        ops.TIMER.enabled = True                 # rank-0-only section: no collectives in here
        for _ in range(3):
            step(pix_d, tgt_d, reduce=False)
        ops.TIMER.summary()
        reps = 5
        for _ in range(reps):
            flush.fill_(1.0)
            step(pix_d, tgt_d, reduce=False)
        summ = ops.TIMER.summary()
        ops.TIMER.enabled = False
        V, A = (int(t.item()) for t in jt.VMRender.last_counts)      # measured on the last step's batch
        breakdown = {k: round(v[0] / reps, 4) for k, v in sorted(summ.items(), key=lambda kv: -kv[1][0])}
        top = next(iter(breakdown))
        cd, ca = model.density_n_comp[0], model.app_n_comp[0]
        F_, H_ = model.app_dim, model.featureC
        in_dim = F_ + 3 + 2 * model.fea_pe * F_ + 6 * model.view_pe
        per_launch_ms = summ[top][0] / summ[top][1]
        by = algorithmic_bytes(top, V, A, cd, ca, sum(model.app_n_comp))
        fl = gemm_flops(top, A, F_, sum(model.app_n_comp), in_dim, H_)
        if by is not None:
            ach = by / (per_launch_ms * 1e-3) / 1e9
            roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": NCU_DRAM_BYTES.get(top), "peak_source": peak_src,
                    "ms_per_launch": per_launch_ms, "algorithmic_bytes": by,
                    "note": "algorithmic bytes count every tap as an HBM access (SURVEY 8d); the 69 MB of factors "
                            "stay in the 126 MB L2 and consecutive samples of a ray share cells, so frac > 1 and "
                            "DRAM traffic << algorithmic bytes (profiles/r01_ncu_full_v14.txt)"}
        elif fl is not None:
            ach = fl / (per_launch_ms * 1e-3) / 1e12
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": ach / tf_peak, "traffic": NCU_DRAM_BYTES.get(top), "peak_source": peak_src,
                    "ms_per_launch": per_launch_ms}
        gk = [k for k in ("vm_density_fwd", "vm_app_fwd", "app_basis_fwd_tc", "app_basis_sh_fwd_tc", "vm_density_bwd",
                          "vm_app_bwd") if k in breakdown]
        gather = sum(breakdown[k] for k in gk)
        gbytes = sum(algorithmic_bytes(k, V, A, cd, ca, sum(model.app_n_comp)) for k in gk)
        if roof is not None and gather > 0:
            roof["vm_gather_fwd_bwd"] = {"ms": gather, "achieved": gbytes / (gather * 1e-3) / 1e9,
                                         "frac": gbytes / (gather * 1e-3) / 1e9 / hbm_peak, "V": V, "A": A,
                                         "kernels": gk}

    # second half of BASELINE.json's metric: one 800x800 frame (640 000 rays, cfg2 field, is_train=False, no
    # blur) rendered like the reference's render_by_slices (model/nerf.py:728-740), rays resident on the device
    def time_render(mdl, mopt):
        n_pix = H_img * W_img
        f_pose, f_kinv = gt_pose[:1].to(dev), intr_inv_d[:1]
        rkw = dict(white_bg=run["white_bg"], is_train=False, ndc_ray=run["ndc"], N_samples=S)
        chunk = args.render_chunk

        def render_frame():
            """render_by_slices (nerf.py:728-740): per slice, rays of that slice only (jt_pose_rays_fwd) -> forward."""
            outs = []
            with torch.no_grad():
                for c in range(0, n_pix, chunk):
                    ce, ra = jt.camera.get_center_and_ray(cam_opt, f_pose, f_kinv, pix_base=c, n_rays=min(chunk, n_pix - c),
                                                          intr=intr_d[:1] if ndc else None)
                    outs.append(mdl(mopt, ce.view(-1, 3), ra.view(-1, 3), **rkw)[0])
            return torch.cat(outs)

        render_frame()
        torch.cuda.synchronize()
        frames = 3
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1.0)
        s_ev.record()
        for _ in range(frames):
            img = render_frame()
        e_ev.record()
        torch.cuda.synchronize()
        fms = s_ev.elapsed_time(e_ev) / frames
        return {"ms_per_frame": fms, "rays_per_s": n_pix / (fms * 1e-3), "frame": f"{W_img}x{H_img}",
                "frames_timed": frames, "rays_per_call": chunk, "finite": bool(torch.isfinite(img).all())}

    render = time_render(model, opt) if rank == 0 and not args.no_render else None

    # the other shading head on the same 300^3 field (cfg2_sh <-> cfg2): same rays, same protocol, timed after the
    # headline so that it cannot disturb it. Single-GPU runs only.
    also = None
    other = {"cfg2_sh": "cfg2", "cfg2": "cfg2_sh"}.get(args.workload)
    if rank == 0 and world == 1 and other and not args.no_also and args.blur == 0:
        kw_o, _ = jt.synth.config(other)
        kw_o = dict(kw_o)
        torch.manual_seed(0)
        model_o = jt.B200_VMSplit(torch.tensor(kw_o.pop("aabb")), kw_o.pop("gridSize"), dev, **kw_o)
        model_o.head_precision = args.head
        opt_o = default_opt(model_o.shadingMode, run["ndc"])
        step_o = make_step(model_o, opt_o, [p for p in model_o.parameters()] + [se3_refine])
        for _ in range(max(args.warmup, 3)):
            step_o(pix_d, tgt_d, reduce=False)
        torch.cuda.synchronize()
        ms_o = timed(lambda: step_o(pix_d, tgt_d, reduce=False), args.steps)
        also = {"workload": jt.synth.describe(other, N), "value": N * args.steps / (ms_o * 1e-3), "unit": UNIT,
                "ms_per_step": ms_o / args.steps, "steps": args.steps}
        if not args.no_render:
            also["render_800x800"] = time_render(model_o, opt_o)
        del model_o, step_o
        torch.cuda.empty_cache()

    # next rows of SURVEY 8f, timed after everything above because the optimiser changes the parameters:
    # 8f-2 a full training iteration = the step above + density_L1 regulariser (Blender weight 8e-5, TV weights 0,
    # bat_blender_VM.yaml:134-139) + Adam update of every parameter; 8f-3 the between-step maintenance ops.
    extras = None
    if rank == 0 and not args.no_breakdown:
        from joint_tensorf_b200.sweeps import FusedAdam
        fused = FusedAdam(model.get_optparam_groups(0.02, 1e-3) + [{"params": [se3_refine], "lr": 1e-3}],
                          betas=(0.9, 0.99))
        decay = 0.1 ** (1 / 30000)

        def train_step():
            step(pix_d, tgt_d, reduce=False)
            model.regularize_(8e-5, 0.0, 0.0)
            fused.step()
            for grp in fused.param_groups:
                grp["lr"] *= decay

        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
        ms_train = timed(train_step, 5) / 5
        ops.TIMER.enabled = True
        model.regularize_(8e-5, 1.0, 1.0)
        fused.step()
        ops.TIMER.summary()
        for _ in range(5):
            flush.fill_(1.0)
            model.regularize_(8e-5, 1.0, 1.0)            # all three terms: the LLFF configuration (weights 10.0)
            fused.step()
        sw = {k: round(v[0] / 5, 4) for k, v in ops.TIMER.summary().items()}
        n_par = sum(p.numel() for grp in fused.param_groups for p in grp["params"])
        # arrays the three-term sweep touches: density planes + lines (L1, TV) and appearance planes (TV)
        n_sw = sum(p.numel() for p in [*model.density_plane, *model.density_line, *model.app_plane])
        # algorithmic bytes: Adam reads p, g, m, v and writes p, m, v; the value sweep reads x once; the gradient
        # sweep reads x and read-modify-writes g
        sweep_bytes = {"adam": 28 * n_par, "reg_values": 4 * n_sw, "reg_grads": 12 * n_sw}
        extras = {"train_step_ms": round(ms_train, 4),
                  "train_step_what": "step + density_L1 (value + gradient sweep) + fused Adam over all parameters",
                  "sweep_ms": sw,
                  "sweep_gbs": {k: round(sweep_bytes[k] / (sw[k] * 1e-3) / 1e9, 1) for k in sweep_bytes if sw.get(k)},
                  "sweep_bytes": sweep_bytes, "hbm_peak_gbs": hbm_peak}
        # maintenance: dense alpha on 200^3 (tensorf.py update_alphamask), mask build, 150^3 -> 300^3 upsample
        model.kernel_density, model.c2f_mode = None, None
        kw2, _ = jt.synth.config(args.workload)
        kw2 = dict(kw2)
        kw2.pop("gridSize")
        aabb2 = torch.tensor(kw2.pop("aabb"))
        mt = None
        for rep in range(2):                          # rep 0 warms up, rep 1 is measured
            ops.TIMER.summary()
            a_zyx = model._dense_alpha_zyx([200, 200, 200])
            ops.alpha_mask_build(a_zyx, model.alphaMask_thres)
            small = jt.B200_VMSplit(aabb2, [150, 150, 150], dev, **kw2)
            small.upsample_volume_grid([300, 300, 300])
            del small
            mt = ops.TIMER.summary()
        ops.TIMER.enabled = False
        extras["maintenance_ms"] = {"dense_alpha_200^3": round(mt["field_alpha"][0], 4),
                                    "maxpool5_threshold_pack_200^3": round(mt["alpha_mask_build"][0], 4),
                                    "upsample_150^3_to_300^3_all_factors": round(mt["resize_bilinear"][0], 4)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rps, sec, cores = time_cpu_port(args.workload, args.cpu_rays, 2, 1, args.blur)
        cpu = {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_rays} of {N} rays, 2 timed fwd+bwd iterations of oracle/vm_oracle.py "
                         f"(torch CPU, {cores} threads)"}

    aten = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rps, sec = time_aten_gpu(args.workload, N, 3, 1, dev)
            aten = {"value": rps, "unit": UNIT, "ms_per_step": sec * 1e3,
                    "what": f"reference algorithm (oracle port) with stock ATen CUDA kernels on this GPU, {N} rays, "
                            "fwd+bwd, 3 timed iterations; rays resident, no pose generation"}
        except Exception as ex:      # a baseline leg must never take the bench line down
            aten = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        torch.cuda.empty_cache()

    if rank == 0:
        rays_total = N * world
        h2d = pix_h.numel() * 4 + tgt_h.numel() * 4
        line = {
            "metric": METRIC, "value": rays_total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.head == "fp32" else "f32 (basis/shading GEMMs on tcgen05: bf16 operands, f32 accumulate)",
            "data": "synthetic",
            "config": {"workload": jt.synth.describe(args.workload, N) +
                                   f" ({n_views} views x {R_pix} pixels, rays generated from se3_refine + pose inside the step), "
                                   "fwd+bwd to factor/basis/head/se3 gradients, optimizer step excluded",
                       "head": args.head,
                       "head_arith": "forward: hi+lo split bf16 operands (3 MMAs per product, fp32-class, rgb within 1e-4); "
                                     "backward: bf16 operands, f32 accumulate (gradients within 2e-2 rel)",
                       "blur": args.blur, "l2": "256 MB write between timed steps (L2 flushed)",
                       "parallelism": f"ray-sharded x{world}, NCCL all-reduce of the flat fp32 gradient bucket inside the backward "
                                      "(appearance part overlapped with the density scatter)"},
            "e2e": {"value": rays_total * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 3),
            "step_ms_rank0": per_step,
            "clocks": clk.summary(),
        }
        if roof is not None:
            line["roofline"] = roof
        if breakdown is not None:
            line["kernel_ms_per_step"] = breakdown
        if render is not None:
            line["render_800x800"] = render
        if also is not None:
            line["also"] = also
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if aten is not None:
            line["aten_gpu_baseline"] = aten
        if extras is not None:
            line["next_rows"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        own_arm(a)
