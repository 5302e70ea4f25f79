"""Diagnostic: per-parameter difference between plain-autograd gradients and TrainEngine.forward_backward (direct accumulation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import port as O
from jperceiver_b200 import _lib, netops, functional as JF, conv as JC
from jperceiver_b200.apis import TrainEngine
from jperceiver_b200.model import MONO
_lib.lib()
dev = torch.device("cuda:0")
opt = dict(name="Baseline", depth_num_layers=18, pose_num_layers=18, frame_ids=[0, -1, 1], imgs_per_gpu=2, height=128, width=384,
           scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, depth_pretrained_path=None, pose_pretrained_path=None,
           automask=True, disp_norm=True, smoothness_weight=1e-3, scale_weight=0.1, dynamic_weight=15.0, static_weight=5.0,
           occ_map_size=64, num_class=2, loss_type="iou", loss_weight=20, loss2_type="boundary", loss2_weight=20,
           type="static", loss_sum=3, split="odometry", automask_noise=0.0)
model = MONO.module_dict["Baseline"](opt)
model.load_state_dict(O.synth_params(model.state_dict(), seed=5))
model.to(dev).train()
model.DepthDecoder.drop_p = 0.0
inp = O.synth_inputs(opt, 2, seed=2, hw_full=(120, 400))
oK = inp[("odometry_K", 0, 0)]
oK[:, 0, 0] *= 0.3; oK[:, 1, 1] *= 0.3; oK[:, 0, 2] = 200.0; oK[:, 1, 2] = 40.0
data = {k: v.to(dev) for k, v in inp.items()}
engine = TrainEngine(model)
names = [n for n, p in model.named_parameters() if p.requires_grad]

def plain():
    engine.flat.zero_grad()
    _, losses = model(data)
    sum(losses.values()).backward()
    return engine.flat.grad.clone(), {k: float(v) for k, v in losses.items()}

def report(tag, a, b):
    rows = []
    for n, (off, num) in zip(names, engine.flat.views):
        x, y = a[off:off + num], b[off:off + num]
        d = (x - y).abs().max().item()
        s = max(x.abs().max().item(), 1e-12)
        rows.append((d / s, d, s, n, num))
    rows.sort(reverse=True)
    print("==", tag, "max abs diff %.4g, scale %.4g" % ((a - b).abs().max().item(), a.abs().max().item()))
    for r in rows[:12]:
        print("   rel %.3e abs %.3e scale %.3e  %s (%d)" % r)

g0, l0 = plain()
g1, l1 = plain()
report("plain vs plain (run-to-run)", g0, g1)
print("loss diffs", max(abs(l0[k] - l1[k]) for k in l0))
for mode in ("direct+wt", "direct only", "wt only"):
    for call in range(2):
        engine.flat.zero_grad()
        JC.WT.enabled = mode != "direct only"
        if JC.WT.enabled:
            JC.WT.refresh()
        _, losses = model(data)
        JF.DIRECT_GRAD = mode != "wt only"
        sum(losses.values()).backward()
        JF.DIRECT_GRAD = False
        JC.WT.enabled = JC.WT.fresh = False
        report("%s call %d vs plain" % (mode, call), engine.flat.grad.clone(), g0)
