#!/usr/bin/env python
"""bench.py -- images/sec of the FPD hourglass training step (BASELINE.json configs[1]/[2]):
student hourglass(stacks=4, features=128) + frozen teacher hourglass(stacks=8, features=256), FPD loss,
256x256 synthetic inputs, batch 32 per GPU, pure data parallel.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's algorithm on the host CPU cores (oracle port)

One JSON line on stdout (rank 0). A "step" = student fwd + teacher fwd + fused FPD loss + student bwd +
gradient all-reduce (N>1) + Adam on one batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NS = types.SimpleNamespace
FLOP_PER_IMAGE = 79.633e9  # student fwd 7.8145 + bwd 15.629 + teacher fwd 56.189 GFLOP
CONV_H_3X3_DRAM_BYTES = 87.09e6   # dram__bytes_read + write of that kernel and shape, profiles/r1c_prof_conv_h_3x3.md
WORKLOAD = "hourglass FPD train: student s4 f128 + frozen teacher s8 f256, 256x256, batch 32/GPU"


def cfg(f, s, j=16):
    return NS(MODEL=NS(EXTRA=NS(NUM_FEATURES=f, NUM_STACKS=s, NUM_BLOCKS=1), NUM_JOINTS=j))


def synthetic_batch(B, seed, H=256, W=256, J=16):
    """N(0,1) images, reference-style Gaussian targets (sigma 2, 13x13, peak 1), 0/1 weights."""
    import numpy as np
    import torch
    rng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    h, w = H // 4, W // 4
    t = np.zeros((B, J, h, w), np.float32)
    xs = np.arange(13, dtype=np.float32)
    gauss = np.exp(-((xs[None] - 6) ** 2 + (xs[:, None] - 6) ** 2) / 8.0)
    for b in range(B):
        for j in range(J):
            mx, my = rng.randint(0, w), rng.randint(0, h)
            x0, y0 = mx - 6, my - 6
            gx0, gx1 = max(0, -x0), min(x0 + 13, w) - x0
            gy0, gy1 = max(0, -y0), min(y0 + 13, h) - y0
            t[b, j, max(0, y0):min(y0 + 13, h), max(0, x0):min(x0 + 13, w)] = gauss[gy0:gy1, gx0:gx1]
    tw = (rng.rand(B, J, 1) > 0.2).astype(np.float32)
    return x, torch.from_numpy(t), torch.from_numpy(tw)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx = float(p[2])
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm) -- cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------------
def cpu_fpd_steps(B, max_steps, max_seconds, warmup=1, as_written=False):
    """Times the reference algorithm (oracle restatement of lib/core/function.py:119-147 on
    lib/models/hourglass.py) on the host CPU. Returns (images_per_s, steps_timed, threads)."""
    import torch
    from oracle import hourglass_oracle as O
    import fpd_b200  # noqa: F401
    from fpd_b200.lib.models import hourglass as H  # parameter containers only (no compute on CPU)
    # measured on the 128-thread B200 host (gpurun_out/cpu_threads.log): 16 threads 5.2 img/s, 32: 3.3, 64: 1.4,
    # 128: 0.09 -- the batch-4 sample cannot use more threads productively, so cap at 16 unless overridden
    threads = int(os.environ.get("FPD_CPU_THREADS", "0")) or min(16, len(os.sched_getaffinity(0)))
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    s_sd = {k: v.clone() for k, v in H.get_pose_net(cfg(128, 4), True).state_dict().items()}
    t_sd = {k: v.clone() for k, v in H.get_pose_net(cfg(256, 8), False).state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in s_sd.items() if v.is_floating_point() and "running" not in k}
    s_sd.update(params)
    if as_written:  # the reference never freezes / detaches the teacher (function.py:120-121,146)
        tparams = {k: v.requires_grad_(True) for k, v in t_sd.items() if v.is_floating_point() and "running" not in k}
        t_sd.update(tparams)
    opt = torch.optim.Adam(list(params.values()), lr=2.5e-4)
    x, target, tw = synthetic_batch(B, 0)

    def one():
        outs = O.hourglass_net(s_sd, x, 4, training=True)
        if as_written:
            tout = O.hourglass_net(t_sd, x, 8, training=False)[-1]
        else:
            with torch.no_grad():
                tout = O.hourglass_net(t_sd, x, 8, training=False)[-1]
        loss, _, _ = O.fpd_loss(outs, target, tw, tout, 0.5)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < max_seconds):
        one()
        n += 1
    dt = time.perf_counter() - t0
    return B * n / dt, n, threads


def run_reference(args, rank):
    if rank != 0:
        return
    B = 4
    ips, n, threads = cpu_fpd_steps(B, max_steps=max(1, args.steps), max_seconds=150.0, warmup=min(args.warmup, 1))
    sample = "oracle port (teacher under no_grad), batch %d per step, %d timed steps, fp32, %d threads" % (B, n, threads)
    line = {"impl": "reference", "metric": "images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * B / ips, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def time_dominant_kernel(B):
    """Live CUDA-event timing of the dominant kernel (conv_tc_h_kernel<f16>: tcgen05 kind::f16 implicit-GEMM conv, halo
    tile fetched + BN/ReLU-transformed + split once per channel block, taps as shifted copies into TMEM) on its heaviest
    shape: the teacher's 3x3 128->128 @64x64 (18 launches/step = 38.7 % of the teacher's MACs), 3xFP16.
    Eight launches per timed region over four rotating input/output sets (4 x 67 MB > the 126 MB L2: every launch reads
    cold data), replayed from a CUDA graph so the host-side launch cost (tensor-map encodes, ~40 us) is not in the
    region -- the same way the kernel runs inside the training step."""
    import torch
    from fpd_b200 import ops
    H = W = 64
    Cin = Cout = 128
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(B, H, W, Cin, device="cuda", generator=g) for _ in range(4)]
    ys = [torch.empty(B, H, W, Cout, device="cuda") for _ in range(4)]
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * 0.03
    w_hi, w_lo = ops.weight_prep_f16(w)
    mean = torch.zeros(Cin, device="cuda")
    scale = torch.ones(Cin, device="cuda")
    shift = torch.zeros(Cin, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def launch(i):
        ops.conv2d_tc_h(xs[i % 4], w_hi, w_lo, 3, mean=mean, scale=scale, shift=shift, relu=True, out=ys[i % 4])
    for i in range(4):
        launch(i)
    torch.cuda.synchronize()
    reps, per, tot = 5, 8, 0.0
    # the eight launches are replayed from a CUDA graph (as in the training step), so the region holds device time only
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(per):
                launch(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for _ in range(reps):
        flush.zero_()  # L2 flush before each timed group
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / (reps * per)
    flops = 2.0 * B * H * W * Cin * Cout * 9
    return ms, flops


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import fpd_b200  # noqa: F401
    from fpd_b200 import _native as N
    from fpd_b200.lib.models import hourglass as H
    from fpd_b200.train_step import FPDTrainStep

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(0)
    student = H.get_pose_net(cfg(128, 4), True).to(dev)
    teacher = H.get_pose_net(cfg(256, 8), False).to(dev)
    if world > 1:  # identical replicas: broadcast rank 0's weights once (no per-step broadcast, unlike DataParallel)
        for t in list(student.parameters()) + list(student.buffers()):
            dist.broadcast(t.data, 0)
    step = FPDTrainStep(student, teacher, alpha=0.5, lr=2.5e-4, use_graph=not args.no_graph)
    x, target, tw = synthetic_batch(B, 1000 + rank)
    xh, th, wh = x.pin_memory(), target.pin_memory(), tw.pin_memory()
    xd, td, wd = xh.to(dev), th.to(dev), wh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step.step(xd, td, wd, next_x=xd)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: step.step(xd, td, wd, next_x=xd), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1000.0)

    # ---- end to end: pinned-host inputs copied H2D every step, loss read back D2H every step
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step():
        # the next batch's images are what crosses PCIe each step (this step's were staged by the previous call)
        losses = step.step(xh, th, wh, next_x=xh)
        loss_host.copy_(losses, non_blocking=False)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * B / (ms_e2e / 1000.0)
    h2d = xh.numel() * 4 + th.numel() * 4 + wh.numel() * 4
    final_loss = float(loss_host[2])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    line = {"metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            # forward + data-gradient convs: 3xFP16 (fp16 hi/lo operand pairs, fp32 accumulate); weight gradients: 3xTF32;
            # both fp32-grade (parity <= 1e-3 against the fp32 reference). FPD_PRECISION=tf32 = single-pass TF32.
            "dtype": "f16x3+tf32x3" if os.environ.get("FPD_PRECISION", "tf32x3") != "tf32" else "tf32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "per_gpu_batch": B,
                       "parallelism": "dp%d" % world, "cuda_graph": not args.no_graph,
                       "teacher": "second stream, software-pipelined one batch ahead" if step.pipeline else
                                  ("second stream" if step.overlap_teacher else "same stream"),
                       "l2": "per-step working set (activations ~ several GB) exceeds the 126 MB L2; no flush needed",
                       "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(step.launches_per_step or 0) * args.steps,
            "clocks": clocks}
    try:
        k_ms, k_flops = time_dominant_kernel(B)
        f16_peak = peaks["bf16_tflops"]  # kind::f16 runs at the bf16 rate; burst figure: the kernel is timed alone
        ach = k_flops / (k_ms * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": f16_peak, "unit": "TFLOP/s",
                            "frac": ach / f16_peak,
                            # dram__bytes_read+write of this kernel/shape: profiles/r1c_prof_conv_h_3x3.md
                            "traffic": CONV_H_3X3_DRAM_BYTES if B == 32 else None,
                            "kernel": "conv_tc_h_kernel<f16> 3x3 128->128 @64x64 B=%d (3xFP16: 3 MMA passes per "
                                      "algorithmic FLOP; executed-MMA fraction of peak = 3 x frac)" % B,
                            "kernel_ms": k_ms, "peak_source": peak_src + ", kind::f16 = bf16 rate",
                            "step_frac_of_f16_peak": value / world * FLOP_PER_IMAGE / 1e12 / peaks["bf16_tflops_sustained"]}
    except Exception as exc:  # never lose the headline line to a side measurement
        line["roofline"] = {"error": repr(exc)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            ips, n, threads = cpu_fpd_steps(4, max_steps=3, max_seconds=25.0)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": "oracle port of the FPD step, batch 4, %d timed steps, teacher under "
                                              "no_grad" % n}
        except Exception as exc:
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
