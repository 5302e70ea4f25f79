"""Per-layer table of the tensor-core convolutions inside one eager training step (bench workload): CUDA-event time per
distinct (kind, shape), FLOPs, achieved TFLOP/s, the per-layer roofline time max(FLOPs/peak_tf32, min bytes/HBM) and the
excess over it — the list the conv kernel work is prioritised from.  Usage: python tools/conv_layers.py [--batch 4]"""
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from jperceiver_b200 import synthetic, functional as JF
from jperceiver_b200.apis import TrainEngine, change_input_variable
from jperceiver_b200.model import MONO

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--tf32-peak", type=float, default=0.0, help="TFLOP/s; 0 = measure a cuBLAS TF32 GEMM here")
args = ap.parse_args()
dev = torch.device("cuda:0")
pk, _ = bench.peaks()
if args.tf32_peak <= 0:
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
    for _ in range(3): a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): a @ b
    e1.record(); torch.cuda.synchronize()
    args.tf32_peak = 10 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b
print("tf32 cuBLAS peak %.0f TFLOP/s, hbm %.0f GB/s" % (args.tf32_peak, pk["hbm_gbs"]))
opt = bench.model_options(bench.CONFIGS["C2"], args.batch)
torch.manual_seed(1024)
model = MONO.module_dict["Baseline"](opt).to(dev).train()
model.branch_streams = False          # one stream: the per-call event pairs must not overlap other trunks' kernels
engine = TrainEngine(model)
data = change_input_variable(synthetic.make_batch(opt, args.batch, seed=1024, pin=True), dev)
for _ in range(3):
    engine.step(data, need_log=False)
torch.cuda.synchronize()
JF.PROFILE.clear(); JF.PROFILE_DETAIL.clear(); JF.PROFILE_ON = True
JF.PROFILE_BACKPRESSURE = 16_000_000   # keep the GPU behind the host: event pairs bracket device time (see functional.py)
steps = 2
for _ in range(steps):
    engine.step(data, need_log=False)
torch.cuda.synchronize()
JF.PROFILE_ON = False
rows = []
for (name, tag), evs in JF.PROFILE_DETAIL.items():
    M, N, K, kh, stride, srcC, ups, refl, split = tag
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    n = len(evs) // steps
    fl = 2.0 * M * N * K
    by = 4.0 * (M * N + N * K + (M * sum(srcC) / (1 if name != "conv_fwd" or stride == 1 else 1)))   # output + weight + input once (approx.)
    t_roof = max(fl / (args.tf32_peak * 1e12), by / (pk["hbm_gbs"] * 1e9)) * 1e3
    rows.append(dict(kind=name, M=M, N=N, K=K, k=kh, s=stride, src=list(srcC), up=list(ups), refl=refl, split=split, n=n, ms=ms,
                     ms_each=ms / n, tflops=fl * n / (ms * 1e-3) / 1e12, roof_ms=t_roof * n, excess_ms=ms - t_roof * n))
rows.sort(key=lambda r: -r["excess_ms"])
tot = sum(r["ms"] for r in rows); roof = sum(r["roof_ms"] for r in rows)
print("conv total %.2f ms/step, roofline %.2f ms (%.1f%%)" % (tot, roof, 100 * roof / tot))
for r in rows:
    print("%-10s M=%7d N=%3d K=%5d k%d s%d src=%-16s up=%-10s r%d split=%3d n=%2d  %7.3f ms (%6.1f us each) %6.1f TF/s roof %6.3f ms excess %6.3f" %
          (r["kind"], r["M"], r["N"], r["K"], r["k"], r["s"], r["src"], r["up"], r["refl"], r["split"], r["n"], r["ms"], r["ms_each"] * 1e3,
           r["tflops"], r["roof_ms"], r["excess_ms"]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(tf32_peak=args.tf32_peak, rows=rows), open(os.path.join(ROOT, "gpurun_out", "conv_layers.json"), "w"))
