"""Diagnostic: run-to-run noise of the depth trunk (forward outputs and flat gradient) under different switches."""
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
           type="static_eigen", loss_sum=3, split="odometry", automask_noise=0.0)
model = MONO.module_dict["Baseline"](opt)
model.load_state_dict(O.synth_params(model.state_dict(), seed=5))
model.to(dev).train()
model.DepthDecoder.drop_p = 0.0
img = O.synth_inputs(opt, 2, seed=2, hw_full=(120, 400))[("color_aug", 0, 0)].to(dev)
engine = TrainEngine(model)
names = [n for n, p in model.named_parameters() if p.requires_grad]
tgt = [torch.rand(2, 1, 128 >> (s + 1), 384 >> (s + 1), generator=torch.Generator().manual_seed(s)).to(dev) for s in range(4)]

def run():
    engine.flat.zero_grad()
    feats = model.DepthEncoder(img)
    out = model.DepthDecoder(feats)
    loss = sum(((out[("disp", 0, s)] - tgt[s]) ** 2).mean() for s in range(4))
    loss.backward()
    return engine.flat.grad.clone(), [f.detach().clone() for f in feats], [out[("disp", 0, s)].detach().clone() for s in range(4)]

def compare(tag):
    g0, f0, d0 = run()
    g1, f1, d1 = run()
    print("==", tag)
    print("  encoder feature max|diff|/max:", ["%.2e" % ((a - b).abs().max().item() / a.abs().max().item()) for a, b in zip(f0, f1)])
    print("  disp max|diff|:", ["%.2e" % (a - b).abs().max().item() for a, b in zip(d0, d1)])
    n = g0.norm().item()
    print("  grad rel L2 noise %.3e" % ((g0 - g1).norm().item() / n))
    rows = []
    for nm, (off, num) in zip(names, engine.flat.views):
        d = (g0[off:off + num] - g1[off:off + num]).norm().item()
        rows.append((d / n, g0[off:off + num].norm().item() / n, nm))
    rows.sort(reverse=True)
    for r in rows[:6]:
        print("     contrib %.3e (param norm share %.3e) %s" % r)

compare("default")
netops.FUSE_BN_STATS = False
compare("no fused BN statistics")
netops.FUSE_BN_STATS = True
JC.KSPLIT_SLOTS = 1
compare("no split-K (forward/dgrad)")
JC.KSPLIT_SLOTS = 296
JC.SMALLN = False
compare("disparity heads through the tensor-core kernel")
