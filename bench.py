#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the JPerceiver training hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" is one full training iteration of ``Baseline`` on one synthetic batch: forward of the three
ResNet-18 stacks + decoders + CCT, ``compute_losses``, backward, gradient all-reduce, clip + Adam.
Workload at every N: BASELINE.json configs[1] — cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20
(type "static", frames [0,-1,1]) at the harness shape 320x1024, batch 4 per GPU (weak scaling), layout branch
under the non-square rule of SURVEY.md §8 a-8.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIG_NAME = "cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20"
H, W, B_PER_GPU = 320, 1024, 4


def model_options(batch):
    """``cfg.model`` of config/cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20.py:22-54 with the BASELINE.json
    harness overrides (320x1024, batch 4, pretrained paths nulled: no checkpoints in this environment)."""
    return dict(name="Baseline", depth_num_layers=18, pose_num_layers=18, frame_ids=[0, -1, 1], imgs_per_gpu=batch,
                height=H, width=W, scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, depth_pretrained_path=None,
                pose_pretrained_path=None, automask=True, disp_norm=True, smoothness_weight=1e-3, scale_weight=0.1,
                dynamic_weight=15.0, static_weight=5.0, occ_map_size=256, num_class=2, loss_type="iou", loss_weight=20,
                loss2_type="boundary", loss2_weight=20, type="static", loss_sum=3, split="odometry", debug_outputs=False)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, batch=1):
    """Forward + compute_losses + backward + clip + Adam of the oracle port (CPU, fp32, all host threads) on a
    bounded sample of the workload: the same config and shape at batch ``batch``."""
    from oracle import port as O
    from jperceiver_b200.model import MONO
    torch.set_num_threads(os.cpu_count() or 1)
    opt = model_options(batch)
    shapes = {k: v for k, v in MONO.module_dict["Baseline"](opt).state_dict().items()}   # shapes only (host-side holders)
    P = O.synth_params(shapes, seed=0)
    params = [v.requires_grad_(True) for k, v in P.items() if v.is_floating_point() and "running" not in k]
    optim = torch.optim.Adam(params, lr=1e-4)
    inp = O.synth_inputs(opt, batch, seed=1024)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        optim.zero_grad()
        _, losses = O.forward(P, opt, inp, training=True)
        O.total_loss(losses).backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 35)
        optim.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": batch / sec, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d step(s) of the same config/shape at batch %d after %d warm-up (oracle/port.py: fwd+losses+bwd+clip+Adam, "
                      "fp32, torch CPU)" % (steps, batch, warmup), "s_per_step": sec}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = cpu_reference_run(steps, warm, batch=1)
    line = {"impl": "reference", "metric": "training images/sec", "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s @%dx%d, type=static, frames [0,-1,1]; CPU arm runs batch 1 per step (bounded sample of batch %d)"
                                   % (CONFIG_NAME, H, W, B_PER_GPU)},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from jperceiver_b200 import _lib, functional as JF, netops, synthetic
    from jperceiver_b200.apis import TrainEngine, change_input_variable
    from jperceiver_b200.model import MONO

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    B = args.batch
    opt = model_options(B)
    torch.manual_seed(1024)
    model = MONO.module_dict["Baseline"](opt).to(dev).train()
    engine = TrainEngine(model, dict(type="Adam", lr=1e-4, weight_decay=0), dict(max_norm=35, norm_type=2))
    host = synthetic.make_batch(opt, B, seed=1024 + rank, pin=True)
    resident = change_input_variable(host, dev)
    torch.cuda.synchronize()

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
        return float(ms.item())

    d2h = torch.empty(32, dtype=torch.float32).pin_memory()
    work_stream = torch.cuda.Stream()          # all steps (eager and captured) run on one non-default stream
    work_stream.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(work_stream)
    # ---- eager profiling pass: per-kernel CUDA-event timings (roofline inputs) and the launch count of one step
    for _ in range(args.warmup):
        engine.step(resident, need_log=False)
    torch.cuda.synchronize()
    JF.PROFILE.clear()
    JF.PROFILE_DETAIL.clear()
    JF.PROFILE_ON = True
    launches0 = _lib.launches
    prof_steps = 3
    for _ in range(prof_steps):
        engine.step(resident, need_log=False)
    torch.cuda.synchronize()
    JF.PROFILE_ON = False
    launches_per_step = (_lib.launches - launches0) // prof_steps
    kern = JF.profile_summary()
    # tensor-core convolution work of one step: algorithmic FLOPs (2*M*N*K per launch, from the launch tags) and kernel time
    conv_flops, conv_ms = 0.0, 0.0
    for (name, tag), evs in JF.PROFILE_DETAIL.items():
        if name in ("conv_fwd", "conv_dgrad", "conv_wgrad"):
            conv_flops += 2.0 * tag[0] * tag[1] * tag[2] * len(evs) / prof_steps
            conv_ms += sum(a.elapsed_time(b) for a, b in evs) / prof_steps
    # the same launches against the tighter of their two rooflines: per launch max(FLOPs / tensor peak,
    # minimum bytes / HBM peak) with minimum bytes = output + weights + input read once (input taken as M*K/kh^2 elements: a lower
    # bound for strided / concatenated layers, so the fraction below is not flattered)
    conv_bound_ms = None
    try:
        pk0, _ = peaks()
        tot = 0.0
        for (name, tag), evs in JF.PROFILE_DETAIL.items():
            if name in ("conv_fwd", "conv_dgrad", "conv_wgrad"):
                M_, N_, K_, kh_ = float(tag[0]), float(tag[1]), float(tag[2]), float(max(int(tag[3]), 1))
                t_tensor = 2.0 * M_ * N_ * K_ / (pk0.get("bf16_tflops_sustained", pk0["bf16_tflops"]) / 2.0 * 1e12)
                t_hbm = 4.0 * (M_ * N_ + N_ * K_ + M_ * K_ / (kh_ * kh_)) / (pk0["hbm_gbs"] * 1e9)
                tot += max(t_tensor, t_hbm) * 1e3 * len(evs) / prof_steps
        conv_bound_ms = tot
    except Exception as e:   # noqa: BLE001  (an extra figure must never cost the bench line)
        sys.stderr.write("bench: per-kernel conv bound not computed (%r)\n" % (e,))
    for v in kern.values():
        v["ms_per_step"] = v["ms_total"] / prof_steps
        v["launches_per_step"] = v["launches"] // prof_steps

    pipelined = [False]
    if args.graph:
        engine.capture(resident, warmup=2)
        step_resident = lambda: engine.replay()

        pipelined[0] = not args.no_prefetch

        def step_e2e():
            # every step: one H2D of a whole batch from pinned memory + the step + a blocking D2H of the loss scalars.  With
            # prefetch (default) the H2D of the NEXT batch overlaps the step on a side stream (double-buffered staging), as a
            # prefetching loader would; --no-prefetch copies this step's batch first, then steps.
            if pipelined[0]:
                try:
                    out = engine.replay_pipelined(host)
                except Exception as e:                         # never lose the e2e number to the optional overlap
                    sys.stderr.write("bench: input prefetch disabled (%r)\n" % (e,))
                    pipelined[0] = False
                    out = engine.replay(host)
            else:
                out = engine.replay(host)
            d2h[:out.numel()].copy_(out, non_blocking=False)   # D2H of the step's result (blocks: the loss is read)
    else:
        step_resident = lambda: engine.step(resident, need_log=False)

        def step_e2e():
            data = change_input_variable(host, dev)
            out = engine.step(data, need_log=True)
            d2h[:out.numel()].copy_(out, non_blocking=False)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop()
    launches = launches_per_step * args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    if rank != 0:
        _finish(world)
        return
    pk, pk_src = peaks()
    # the convolutions are timed inside a long step: their denominator is the SUSTAINED tensor figure when the driver measured one
    tf32_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2.0
    tf32_src = pk_src + (" sustained" if "bf16_tflops_sustained" in pk else "") + " bf16 dense peak / 2 (kind::tf32 issues at half the bf16 rate)"
    ms_step = ms / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_value = B * world / (ms_e2e / args.steps / 1e3)
    # roofline of the dominant hand-written kernel: fused photometric forward (one launch per scale)
    F_src = len(opt["frame_ids"]) - 1
    alg = [B * (4 * H * W * (3 + 3 * F_src) + 4 * (H >> (s + 1)) * (W >> (s + 1))) for s in range(4)]
    pf = kern.get("photometric_fwd", {"ms_per_launch": None, "launches": 0})
    achieved = (sum(alg) / 4) / (pf["ms_per_launch"] * 1e-3) / 1e9 if pf["ms_per_launch"] else None
    line = {
        "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 storage, tf32 tensor-core convolutions, f32 loss chain", "data": "synthetic",
        "config": {"workload": "%s @%dx%d (harness shape), type=static, frames [0,-1,1], batch %d/GPU, full training step "
                               "(fwd + compute_losses + bwd + allreduce + clip + Adam)" % (CONFIG_NAME, H, W, B),
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (activations, several GB) exceeds the 126 MB L2; no explicit flush",
                   "frames_per_s": value * (1 + F_src), "operator_kernels": dict(netops.KERNELS),
                   "execution": "one CUDA graph per step (captured from the eager step)" if args.graph else "eager"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": synthetic.batch_bytes(host),
                "d2h_bytes_per_step": 4 * (len(engine.last_names)),
                "input_pipeline": ("H2D of the next batch overlapped with the step (double-buffered staging, side stream)"
                                   if pipelined[0] else "H2D of the step's batch, then the step")},
        "gpu_launches": launches, "clocks": clocks,
        # dominant kernels of the step: the tcgen05 implicit-GEMM convolutions (forward, data gradient, weight gradient)
        "roofline": {"kernel": "conv_tc_fwd / conv_tc_fwd2 / conv_tc_wgrad (tcgen05 kind::tf32, all %d launches of a step)"
                               % sum(kern.get(k, {}).get("launches_per_step", 0) for k in ("conv_fwd", "conv_dgrad", "conv_wgrad")),
                     "bound": "tensor", "achieved": conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                     "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": (conv_flops / (conv_ms * 1e-3) / 1e12) / tf32_peak if conv_ms else None,
                     "traffic": None, "peak_source": tf32_src,
                     "algorithmic_flops_per_step": conv_flops, "ms_per_step": conv_ms,
                     "per_kernel_bound_ms_per_step": conv_bound_ms,
                     "frac_of_per_kernel_bound": (conv_bound_ms / conv_ms) if (conv_bound_ms and conv_ms) else None,
                     "share_of_step": conv_ms / ms_step if conv_ms else None},
        # the kernel BASELINE.json names: fused photometric loss, forward, one launch per scale
        "roofline_photometric": {"kernel": "photometric_fwd_kernel (mean over the 4 scale launches)", "bound": "hbm", "achieved": achieved,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (achieved / pk["hbm_gbs"]) if achieved else None,
                                 "traffic": 52171520, "traffic_source": "profiles/r1_ncu_photo_v2.csv (scale 0: dram read 48.56 MB + write 3.61 MB)",
                                 "peak_source": pk_src, "algorithmic_bytes_per_launch": sum(alg) / 4,
                                 "ms_per_launch": pf["ms_per_launch"],
                                 "note": "FP32-issue bound, not HBM bound: ~1800 thread instructions per pixel (ncu), DRAM traffic == algorithmic bytes"},
        "kernels": {k: {"ms_per_step": round(v["ms_per_step"], 4), "launches_per_step": v["launches_per_step"],
                        "ms_per_launch": round(v["ms_per_launch"], 5)} for k, v in kern.items()},
        "kernel_timing": "CUDA events around each C-ABI call during %d eager steps before the timed region" % prof_steps,
    }
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(1, 1, batch=1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    _finish(world)


def _finish(world):
    """Multi-rank exit.  Tearing the NCCL communicator down while the captured step graph (which holds NCCL kernels) is still
    alive blocked the processes after the result line had been printed; every rank has finished its collectives here, so
    flush and leave without the teardown."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e: copy each step's batch before the step instead of overlapping the next batch's H2D")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the jperceiver_b200 arm has no CPU path (use --impl reference)")
        run_ours(args)


if __name__ == "__main__":
    main()
