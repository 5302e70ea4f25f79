"""Per-stream GPU kernel time of one eager multi-stream training step via torch.profiler (CUPTI): which stream bounds the step
(the sum of a stream's kernel durations is a lower bound of the graph-replay step time), and what runs on it."""
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
engine = TrainEngine(model, dict(type="Adam", lr=1e-4, weight_decay=0), dict(max_norm=35, norm_type=2))
data = change_input_variable(synthetic.make_batch(opt, 4, seed=1024, pin=True), dev)
for _ in range(3):
    engine.step(data, need_log=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    engine.step(data, need_log=False)
    torch.cuda.synchronize()
trace = os.path.join(ROOT, "gpurun_out", "step_trace.json")
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
per = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0]))
tot = collections.defaultdict(float)
for e in ev:
    st = e["args"].get("stream", -1)
    n = re.sub(r"\(.*", "", e["name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))[:60]
    per[st][n][0] += 1
    per[st][n][1] += e["dur"]
    tot[st] += e["dur"]
print("kernels %d, total %.2f ms" % (len(ev), sum(tot.values()) / 1e3))
for st, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("stream %s: %.2f ms over %d launches" % (st, t / 1e3, sum(v[0] for v in per[st].values())))
    for n, (c, d) in sorted(per[st].items(), key=lambda kv: -kv[1][1])[:14]:
        print("    %8.1f us n=%4d  %s" % (d, c, n))
os.remove(trace)
