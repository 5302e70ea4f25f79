"""Timeline of ONE replay of the captured multi-stream step graph (CUPTI via torch.profiler): per stream the first / last kernel
time and busy time, and how much of the step has 0 / 1 / 2+ kernels in flight — where the step waits on a single stream."""
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
ws = torch.cuda.Stream(); ws.wait_stream(torch.cuda.current_stream()); torch.cuda.set_stream(ws)
for _ in range(3):
    engine.step(data, need_log=False)
engine.capture(data, warmup=2)
for _ in range(3):
    engine.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    engine.replay()
    torch.cuda.synchronize()
trace = os.path.join(ROOT, "gpurun_out", "replay_trace.json")
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
os.remove(trace)
t0 = min(e["ts"] for e in ev); t1 = max(e["ts"] + e["dur"] for e in ev)
print("replay: %d kernels, span %.2f ms, sum of kernel time %.2f ms" % (len(ev), (t1 - t0) / 1e3, sum(e["dur"] for e in ev) / 1e3))
per = collections.defaultdict(list)
for e in ev:
    per[e["args"].get("stream", -1)].append(e)
for st, es in sorted(per.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print("stream %3s: %4d kernels, busy %6.2f ms, first at %6.2f ms, last ends at %6.2f ms" %
          (st, len(es), sum(e["dur"] for e in es) / 1e3, (min(e["ts"] for e in es) - t0) / 1e3, (max(e["ts"] + e["dur"] for e in es) - t0) / 1e3))
# concurrency histogram
pts = []
for e in ev:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
hist = collections.defaultdict(float); cur = 0; last = t0
for t, d in pts:
    hist[min(cur, 4)] += t - last; last = t; cur += d
print("time with k kernels in flight (ms):", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
# the 12 longest single-kernel-in-flight stretches: which kernel runs alone
alone = []
cur = []; active = {}
evs = sorted(ev, key=lambda e: e["ts"])
import heapq
ends = []
i = 0
timeline = sorted([(e["ts"], 0, idx) for idx, e in enumerate(evs)] + [(e["ts"] + e["dur"], 1, idx) for idx, e in enumerate(evs)])
act = set(); last = t0
solo = collections.defaultdict(float)
for t, kind, idx in timeline:
    if len(act) == 1:
        n = re.sub(r"\(.*", "", evs[next(iter(act))]["name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))[:50]
        solo[n] += t - last
    last = t
    if kind == 0: act.add(idx)
    else: act.discard(idx)
print("kernels running ALONE (ms):")
for n, v in sorted(solo.items(), key=lambda kv: -kv[1])[:16]:
    print("   %7.3f  %s" % (v / 1e3, n))
