"""Micro-benchmark of the memory-bound network kernels (BatchNorm fwd/bwd, max-pool fwd/bwd) at the largest shapes of the
320x1024 / B=4 step; prints ms and achieved GB/s against the bytes each call must move."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jperceiver_b200 import functional as JF  # noqa: E402
CL = torch.channels_last
dev = torch.device("cuda:0")


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, (B, C, H, W), relu, has_res in [("layout stem bn 64@512x512", (4, 64, 512, 512), True, False), ("layout layer1 bn2 64@256x256", (4, 64, 256, 256), True, True),
                                         ("depth stem bn 64@160x512", (4, 64, 160, 512), True, False), ("layer3 bn 256@20x64", (4, 256, 20, 64), True, True)]:
    x = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=CL).requires_grad_(True)
    res = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=CL) if has_res else None
    g, b = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbytes = x.numel() * 4
    tf = bench(lambda: JF.batchnorm_train(x, res, g, b, rm, rv, 0.1, 1e-5, relu))
    y = JF.batchnorm_train(x, res, g, b, rm, rv, 0.1, 1e-5, relu)
    gy = torch.randn_like(y)
    tb = bench(lambda: torch.autograd.grad(y, [x], gy, retain_graph=True))
    fb = nbytes * (3 + (1 if has_res else 0))
    bb = nbytes * (3 + 4 + (1 if has_res else 0))
    print(json.dumps({"op": name, "fwd_ms": tf, "fwd_gbs": fb / tf / 1e6, "bwd_ms": tb, "bwd_gbs": bb / tb / 1e6}))
for name, (B, C, H, W), k, s, p in [("crp pool 5x5 256@80x256", (4, 256, 80, 256), 5, 1, 2), ("stem pool 3x3s2 64@512x512", (4, 64, 512, 512), 3, 2, 1)]:
    x = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=CL).requires_grad_(True)
    tf = bench(lambda: JF.maxpool(x, k, s, p))
    y = JF.maxpool(x, k, s, p)
    gy = torch.randn_like(y)
    tb = bench(lambda: torch.autograd.grad(y, [x], gy, retain_graph=True))
    fb = x.numel() * 4 + y.numel() * 5
    bb = y.numel() * 5 + x.numel() * 4
    print(json.dumps({"op": name, "fwd_ms": tf, "fwd_gbs": fb / tf / 1e6, "bwd_ms": tb, "bwd_gbs": bb / tb / 1e6}))
