"""Per-kernel GPU time of one eager training step via torch.profiler (CUPTI): true device durations, no replay."""
import os, sys, json, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from jperceiver_b200 import synthetic
from jperceiver_b200.apis import TrainEngine, change_input_variable
from jperceiver_b200.model import MONO

dev = torch.device("cuda:0")
opt = bench.model_options(bench.CONFIGS["C2"], 4)
torch.manual_seed(1024)
model = MONO.module_dict["Baseline"](opt).to(dev).train()
engine = TrainEngine(model)
data = change_input_variable(synthetic.make_batch(opt, 4, seed=1024, pin=True), dev)
for _ in range(3):
    engine.step(data, need_log=False)
torch.cuda.synchronize()
steps = 2
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        engine.step(data, need_log=False)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))[:80]
        agg[n][0] += 1
        agg[n][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print("total GPU kernel time per step: %.2f ms over %d kernels" % (tot / steps / 1e3, sum(v[0] for v in agg.values()) // steps))
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for n, (c, t) in rows[:45]:
    print("%6.2f%% %9.1f us/step n=%4d  %s" % (100 * t / tot, t / steps, c // steps, n))
json.dump({n: {"count_per_step": c // steps, "us_per_step": t / steps} for n, (c, t) in rows}, open(os.path.join(ROOT, "gpurun_out", "step_kernel_times.json"), "w"), indent=0)
