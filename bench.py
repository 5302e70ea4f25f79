#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the JPerceiver training hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores
    python bench.py --config C3 --batch 8 ...                # the other BASELINE.json configurations (see CONFIGS)

A "step" is one full training iteration of ``Baseline`` on one synthetic batch: forward of the three
ResNet-18 stacks + decoders + CCT, ``compute_losses``, backward, gradient all-reduce, clip + Adam.
Default workload at every N: BASELINE.json configs[1] — cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20
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

# BASELINE.json `configs` (index -> reference config file, harness shape, per-GPU batch, scaling mode).  C2 is the bench line
# (configs[1], the configuration the metric is quoted on at N = 1); the others are run on request (--config) and their lines are
# committed under profiles/.
CONFIGS = {
    "C1": dict(name="cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20_B1", type="static", split="odometry", H=192, W=640,
               frame_ids=[0, -1, 1], batch=1, scaling="weak"),
    "C2": dict(name="cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20", type="static", split="odometry", H=320, W=1024,
               frame_ids=[0, -1, 1], batch=4, scaling="weak"),
    "C3": dict(name="cfg_kitti_baseline_raw_boundary_ce_iou_1024_20", type="static_raw", split="raw", H=320, W=1024,
               frame_ids=[0, -1, 1], batch=8, scaling="weak"),
    "C4": dict(name="cfg_kitti_baseline_argo_both_boundary_ce_iou_1024_20_B1", type="Argo_both", split="argo", H=1024, W=1024,
               frame_ids=[0, -1], batch=1, scaling="weak", extra=dict(loss_weightS=20, loss2_weightS=20)),
    "C5": dict(name="cfg_kitti_baseline_kitti_odom_8pugsB24_lr1e-4_ce_eigen", type="static_eigen", split="eigen", H=320, W=1024,
               frame_ids=[0, -1, 1], batch=24, scaling="strong", extra=dict(loss_sum=0)),   # global batch 24 split over the GPUs
}


def model_options(cfg, batch, debug_outputs=True):
    """``cfg.model`` of the reference config file (config/<name>.py:22-54) with the BASELINE.json harness overrides (shape,
    batch, pretrained paths nulled: no checkpoints in this environment)."""
    opt = dict(name="Baseline", depth_num_layers=18, pose_num_layers=18, frame_ids=list(cfg["frame_ids"]), imgs_per_gpu=batch,
               height=cfg["H"], width=cfg["W"], scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, depth_pretrained_path=None,
               pose_pretrained_path=None, automask=True, disp_norm=True, smoothness_weight=1e-3, scale_weight=0.1,
               dynamic_weight=15.0, static_weight=5.0, occ_map_size=256, num_class=2, loss_type="iou", loss_weight=20,
               loss2_type="boundary", loss2_weight=20, type=cfg["type"], loss_sum=3, split=cfg["split"], debug_outputs=debug_outputs)
    opt.update(cfg.get("extra", {}))
    return opt


def per_gpu_batch(cfg, args, world):
    if args.batch:
        return args.batch
    if cfg["scaling"] == "strong":
        if cfg["batch"] % world:
            raise SystemExit("global batch %d does not divide over %d GPUs" % (cfg["batch"], world))
        return cfg["batch"] // world
    return cfg["batch"]


def workload_text(cfg, B, world):
    return ("%s @%dx%d (harness shape), type=%s, frames %s, batch %d/GPU, full training step (fwd + compute_losses + bwd + "
            "allreduce + clip + Adam)" % (cfg["name"], cfg["H"], cfg["W"], cfg["type"], cfg["frame_ids"], B))


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


def ncu_traffic():
    """DRAM traffic per launch of the named kernels from the committed ``ncu`` captures (profiles/r2_ncu_traffic.json, written
    by tools/ncu_traffic.py from ``dram__bytes_read.sum + dram__bytes_write.sum``); {} when the file is absent."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:   # noqa: BLE001
        return {}


# ------------------------------------------------------------------------------------------------
# baselines: the oracle port on the host cores (reference arm / cpu_baseline) and on the GPU in eager torch
# ------------------------------------------------------------------------------------------------
def oracle_train_run(cfg, batch, steps, warmup, device="cpu"):
    """Forward + compute_losses + backward + clip + Adam of the oracle port (oracle/port.py: the reference's arithmetic
    restated, pinned to the reference's own outputs).  ``device='cpu'``: fp32 on all host threads — the reference's CPU path.
    ``device='cuda'``: the same code on the GPU in eager torch with torch's defaults (cuDNN, ``allow_tf32=True``), including the
    reference's host round-trips (scipy EDT and cv2 inside the loss) — what the reference itself does on a GPU."""
    from oracle import port as O
    from jperceiver_b200.model import MONO
    torch.set_num_threads(os.cpu_count() or 1)
    opt = model_options(cfg, batch)
    shapes = {k: v for k, v in MONO.module_dict["Baseline"](opt).state_dict().items()}   # shapes only (host-side holders)
    dev = torch.device(device)
    P = {k: v.to(dev) for k, v in O.synth_params(shapes, seed=0).items()}
    params = [v.requires_grad_(True) for k, v in P.items() if v.is_floating_point() and "running" not in k]
    optim = torch.optim.Adam(params, lr=1e-4)
    inp = {k: v.to(dev) for k, v in O.synth_inputs(opt, batch, seed=1024, hw_full=(2056, 2464) if cfg["split"] == "argo" else (375, 1242)).items()}
    times = []
    with torch.device(dev):
        for it in range(warmup + steps):
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            optim.zero_grad()
            _, losses = O.forward(P, opt, inp, training=True)
            O.total_loss(losses).backward()
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 35)
            optim.step()
            if dev.type == "cuda":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return {"value": batch / sec, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d step(s) of %s @%dx%d at batch %d after %d warm-up (oracle/port.py: fwd+losses+bwd+clip+Adam, %s)"
                      % (steps, cfg["name"], cfg["H"], cfg["W"], batch, warmup,
                         "fp32, torch CPU, all host threads" if dev.type == "cpu" else "eager torch on cuda:0, cuDNN TF32 (torch defaults)"),
            "s_per_step": sec}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = per_gpu_batch(cfg, args, world)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = oracle_train_run(cfg, B, steps, warm)
    line = {"impl": "reference", "metric": "training images/sec", "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(cfg, B, 1) + "; CPU arm: one process on the host cores, %d step(s) — a bounded sample" % steps},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from jperceiver_b200 import _lib, conv as JC, functional as JF, netops, synthetic
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
    cfg = CONFIGS[args.config]
    B = per_gpu_batch(cfg, args, world)
    H, W = cfg["H"], cfg["W"]
    opt = model_options(cfg, B, debug_outputs=not args.no_debug_outputs)
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
    # ---- eager profiling pass ON ONE STREAM (branch streams off, so that per-kernel CUDA-event times do not overlap):
    # per-kernel timings (roofline inputs), the launch count and the serial eager step time
    model.branch_streams = False
    for _ in range(args.warmup):
        engine.step(resident, need_log=False)
    torch.cuda.synchronize()
    JF.PROFILE.clear()
    JF.PROFILE_DETAIL.clear()
    JF.PROFILE_ON = True
    # keep the GPU behind the host during the profiled steps (8 ms of device spin per 48 launches, ~2 ms of host work), so the
    # event pairs bracket kernel time and not the host's launch latency
    JF.PROFILE_BACKPRESSURE = 0 if args.no_backpressure else 16_000_000
    launches0 = _lib.launches
    prof_steps = 3
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(prof_steps):
        engine.step(resident, need_log=False)
    p1.record()
    torch.cuda.synchronize()
    eager_serial_ms = p0.elapsed_time(p1) / prof_steps
    JF.PROFILE_ON = False
    JF.PROFILE_BACKPRESSURE = 0
    model.branch_streams = True
    launches_per_step = (_lib.launches - launches0) // prof_steps
    kern = JF.profile_summary()
    pk, pk_src = peaks()
    tf32_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2.0
    # tensor-core convolution work of one step: algorithmic FLOPs (2*M*N*K per launch, from the launch tags), kernel time, and the
    # same launches against the tighter of their two rooflines: per launch max(FLOPs / tensor peak, minimum bytes /
    # HBM peak), minimum bytes = output + weights + input read once (input as M*K/kh^2 elements: a lower bound)
    conv_flops = conv_ms = conv_bound_ms = 0.0
    top = None
    for (name, tag), evs in JF.PROFILE_DETAIL.items():
        if name in ("conv_fwd", "conv_dgrad", "conv_wgrad"):
            M_, N_, K_, kh_ = float(tag[0]), float(tag[1]), float(tag[2]), float(max(int(tag[3]), 1))
            n = len(evs) / prof_steps
            ms_k = sum(a.elapsed_time(b) for a, b in evs) / prof_steps
            conv_flops += 2.0 * M_ * N_ * K_ * n
            conv_ms += ms_k
            t_tensor = 2.0 * M_ * N_ * K_ / (tf32_peak * 1e12)
            t_hbm = 4.0 * (M_ * N_ + N_ * K_ + M_ * K_ / (kh_ * kh_)) / (pk["hbm_gbs"] * 1e9)
            conv_bound_ms += max(t_tensor, t_hbm) * 1e3 * n
            if top is None or ms_k > top[0]:
                top = (ms_k, name, tag, n)
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
        _finish(world, engine)
        return
    tf32_src = pk_src + (" sustained" if "bf16_tflops_sustained" in pk else "") + " bf16 dense peak / 2 (kind::tf32 issues at half the bf16 rate)"
    ms_step = ms / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_value = B * world / (ms_e2e / args.steps / 1e3)
    traffic = ncu_traffic()
    # roofline of the kernel BASELINE.json names: fused photometric forward (one launch per scale).  Algorithmic bytes per launch:
    # target + F source frames + disp_s read once, and — with the drop-in's default outputs — the warped frames
    # ("color", f, s) and min_index written once.
    F_src = len(opt["frame_ids"]) - 1
    dbg_bytes = (12 * H * W * F_src + 8 * H * W) if opt["debug_outputs"] else 0
    alg = [B * (4 * H * W * (3 + 3 * F_src) + 4 * (H >> (s + 1)) * (W >> (s + 1)) + dbg_bytes) for s in range(4)]
    pf = kern.get("photometric_fwd", {"ms_per_launch": None, "launches": 0})
    pb = kern.get("photometric_bwd", {"ms_per_launch": None, "launches": 0})
    achieved = (sum(alg) / 4) / (pf["ms_per_launch"] * 1e-3) / 1e9 if pf["ms_per_launch"] else None
    alg_b = [B * (4 * H * W * (3 + 3 * F_src) + 8 * (H >> (s + 1)) * (W >> (s + 1)) + (12 * H * W * F_src if opt["debug_outputs"] else 0) + H * W)
             for s in range(4)]
    achieved_b = (sum(alg_b) / 4) / (pb["ms_per_launch"] * 1e-3) / 1e9 if pb["ms_per_launch"] else None
    conv_ach = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None
    line = {
        "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "f32 storage, tf32 tensor-core convolutions (truncation-compensated), f32 loss chain", "data": "synthetic",
        "config": {"workload": workload_text(cfg, B, world), "config_id": args.config,
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (activations, several GB) exceeds the 126 MB L2; no explicit flush",
                   "frames_per_s": value * (1 + F_src), "operator_kernels": dict(netops.KERNELS),
                   "outputs": ("drop-in default: (\"color\",f,s), (\"min_index\",s), (\"depth\",0,s) materialised every step"
                               if opt["debug_outputs"] else "loss-only (debug_outputs=False)"),
                   "execution": ("one CUDA graph per step (captured from the eager step); depth / pose / layout trunks and the four BEV "
                                 "decoders on concurrent streams inside it") if args.graph else "eager"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": synthetic.batch_bytes(host),
                "d2h_bytes_per_step": 4 * (len(engine.last_names)),
                "input_pipeline": ("H2D of the next batch overlapped with the step (double-buffered staging, side stream)"
                                   if pipelined[0] else "H2D of the step's batch, then the step")},
        "gpu_launches": launches, "clocks": clocks,
        # dominant kernels of the step: the tcgen05 implicit-GEMM convolutions (forward, data gradient, weight gradient)
        "roofline": {"kernel": "conv_tc_patch / conv_tc_fwd / conv_tc_fwd2 / conv_tc_wgrad (tcgen05 kind::tf32, all %d launches of a step)"
                               % sum(kern.get(k, {}).get("launches_per_step", 0) for k in ("conv_fwd", "conv_dgrad", "conv_wgrad")),
                     "bound": "tensor", "achieved": conv_ach, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": (conv_ach / tf32_peak) if conv_ach else None,
                     "traffic": traffic.get("conv_step_dram_bytes"), "traffic_source": traffic.get("conv_source"),
                     "peak_source": tf32_src, "algorithmic_flops_per_step": conv_flops, "ms_per_step": conv_ms,
                     "per_kernel_bound_ms_per_step": conv_bound_ms,
                     "frac_of_per_kernel_bound": (conv_bound_ms / conv_ms) if (conv_bound_ms and conv_ms) else None,
                     "share_of_step_kernel_time": (conv_ms / sum(v["ms_per_step"] for v in kern.values())) if conv_ms else None,
                     "largest_launch": ({"kind": top[1], "M": top[2][0], "N": top[2][1], "K": top[2][2], "ms_per_step": top[0],
                                         "tflops": 2.0 * top[2][0] * top[2][1] * top[2][2] * top[3] / (top[0] * 1e-3) / 1e12} if top else None),
                     "note": "cta_group::1 kind::tf32 tops out at ~355 TFLOP/s on this part with every load and the epilogue removed "
                             "(profiles/README.md); timing: CUDA events on one stream, eager"},
        # the kernel BASELINE.json names: fused photometric loss, forward, one launch per scale
        "roofline_photometric": {"kernel": "photometric_fwd_kernel (mean over the 4 scale launches)", "bound": "hbm", "achieved": achieved,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (achieved / pk["hbm_gbs"]) if achieved else None,
                                 "traffic": traffic.get("photometric_fwd_dram_bytes"), "traffic_source": traffic.get("photometric_source"),
                                 "peak_source": pk_src, "algorithmic_bytes_per_launch": sum(alg) / 4,
                                 "ms_per_launch": pf["ms_per_launch"],
                                 "backward": {"ms_per_launch": pb["ms_per_launch"], "algorithmic_bytes_per_launch": sum(alg_b) / 4,
                                              "achieved": achieved_b, "frac": (achieved_b / pk["hbm_gbs"]) if achieved_b else None},
                                 "fp32_issue_bound_note": "the kernel is FP32-issue bound, not HBM bound (ncu: DRAM traffic == algorithmic "
                                                          "bytes, issue slots 41-57 % busy); see profiles/README.md"},
        "kernels": {k: {"ms_per_step": round(v["ms_per_step"], 4), "launches_per_step": v["launches_per_step"],
                        "ms_per_launch": round(v["ms_per_launch"], 5)} for k, v in kern.items()},
        "kernel_timing": "CUDA events around each C-ABI call during %d eager single-stream steps before the timed region, the GPU kept "
                         "behind the host by a device spin every %d launches so that the pairs bracket device time, not launch latency "
                         "(that pass: %.2f ms per step incl. the spins; the timed region replays the multi-stream graph)"
                         % (prof_steps, JF.PROFILE_EVERY, eager_serial_ms),
    }
    if world == 1 and not args.no_gpu_eager_baseline:
        try:
            engine.release_graph()
            torch.cuda.empty_cache()
            g = oracle_train_run(cfg, B, 3, 2, device="cuda")
            line["gpu_eager_baseline"] = {"value": g["value"], "unit": "images/s", "ms_per_step": g["s_per_step"] * 1e3, "sample": g["sample"],
                                          "note": "the honest GPU bar on the same box: the reference's arithmetic in eager torch + cuDNN TF32"}
        except Exception as e:   # noqa: BLE001  (a baseline must never cost the bench line)
            line["gpu_eager_baseline"] = {"unavailable": repr(e)[:300]}
    if not args.no_cpu_baseline and world == 1:
        r = oracle_train_run(cfg, B if cfg["H"] * cfg["W"] * B <= 4 * 320 * 1024 else 1, 1, 1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    _finish(world, engine)


def _finish(world, engine=None):
    """Multi-rank exit: drop the captured step graph (it holds NCCL kernels: tearing the communicator down while it is alive
    blocked the processes in round 1), then a normal ``destroy_process_group``.  A watchdog ends the process if the teardown
    still does not return, so the driver is never left waiting on a finished run."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        t = threading.Timer(20.0, lambda: (sys.stderr.write("bench: NCCL teardown did not return in 20 s, exiting\n"), os._exit(0)))
        t.daemon = True
        t.start()
        try:
            if engine is not None:
                engine.release_graph()
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        finally:
            t.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default C2 = configs[1])")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the configuration's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--no-debug-outputs", action="store_true", help="loss-only step: do not materialise (\"color\",f,s) / min_index / depth outputs")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-backpressure", action="store_true", help="per-kernel event timing without the device spin that keeps the GPU behind the host")
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
