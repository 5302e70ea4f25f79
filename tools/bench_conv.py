"""Micro-benchmark of the tcgen05 convolution kernels on representative layers (B=4, 320x1024 network shapes).
Prints per-layer ms and TFLOP/s (2*M*N*K, unpadded K) for forward, data gradient and weight gradient."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jperceiver_b200 import conv as JC  # noqa: E402

CL = torch.channels_last
dev = torch.device("cuda:0")
LAYERS = [
    # name, sources (C,H,W,up), Cout, k, stride, pad, reflect, act
    ("iconv1 cat(256,up256,1)->256 @80x256", [(256, 80, 256, 0), (256, 40, 128, 1), (1, 80, 256, 0)], 256, 3, 1, 1, 1, "leaky"),
    ("merge1 256->256 @80x256", [(256, 80, 256, 0)], 256, 3, 1, 1, 1, "leaky"),
    ("crp1 1x1 256->256 @80x256", [(256, 80, 256, 0)], 256, 1, 1, 0, 0, "none"),
    ("layer1 64->64 @80x256", [(64, 80, 256, 0)], 64, 3, 1, 1, 0, "none"),
    ("layer2 128->128 @40x128", [(128, 40, 128, 0)], 128, 3, 1, 1, 0, "none"),
    ("layer3 256->256 @20x64", [(256, 20, 64, 0)], 256, 3, 1, 1, 0, "none"),
    ("layer4 512->512 @10x32", [(512, 10, 32, 0)], 512, 3, 1, 1, 0, "none"),
    ("layout layer1 64->64 @256x256", [(64, 256, 256, 0)], 64, 3, 1, 1, 0, "none"),
    ("stem 7x7 s2 (3->4)->64 @320x1024", [(4, 320, 1024, 0)], 64, 7, 2, 3, 0, "none"),
    ("iconv2 cat(256,up256,1)->256 @40x128", [(256, 40, 128, 0), (256, 20, 64, 1), (1, 40, 128, 0)], 256, 3, 1, 1, 1, "leaky"),
    ("layout layer2 128->128 @128x128", [(128, 128, 128, 0)], 128, 3, 1, 1, 0, "none"),
    ("layout layer3 256->256 @64x64", [(256, 64, 64, 0)], 256, 3, 1, 1, 0, "none"),
    ("layout layer4 512->512 @32x32", [(512, 32, 32, 0)], 512, 3, 1, 1, 0, "none"),
    ("pose layer1 64->64 @48x160", [(64, 48, 160, 0)], 64, 3, 1, 1, 0, "none"),
    ("layoutdec up16->16 @256x256", [(16, 128, 128, 1)], 16, 3, 1, 1, 0, "none"),
    ("crp2 1x1 256->256 @40x128", [(256, 40, 128, 0)], 256, 1, 1, 0, 0, "none"),
]


def bench(fn, iters=10):
    """Device time per call: `iters` calls captured into ONE CUDA graph and replayed (host launch overhead — ctypes, tensor-map
    encoding, allocation: ~50 us per call — would otherwise floor every layer below ~55 us)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main(only=None, iters=10):
    B = 4
    out = []
    for name, srcs, cout, k, stride, pad, reflect, act in LAYERS:
        if only and only not in name:
            continue
        g = torch.Generator().manual_seed(0)
        xs = [torch.randn(B, c, h, w, generator=g).to(dev).contiguous(memory_format=CL) for c, h, w, up in srcs]
        ups = [bool(u) for *_, u in srcs]
        cin_t = sum(c for c, *_ in srcs)
        cin_w = 3 if k == 7 else cin_t
        w_ = (torch.randn(cout, cin_w, k, k, generator=g) / (cin_w * k * k) ** 0.5).to(dev).contiguous(memory_format=CL)
        b_ = torch.randn(cout, generator=g).to(dev)
        y = JC.conv2d_tc(xs, ups, w_, b_, stride, pad, reflect, act, None)
        M = y.shape[0] * y.shape[2] * y.shape[3]
        flops = 2.0 * M * cout * cin_w * k * k
        t_f = bench(lambda: JC.conv2d_tc(xs, ups, w_, b_, stride, pad, reflect, act, None), iters)
        dz = torch.randn_like(y)
        rec = {"layer": name, "M": M, "N": cout, "K": cin_w * k * k, "gflop": flops / 1e9, "fwd_ms": t_f, "fwd_tflops": flops / t_f / 1e9}
        if k != 7:
            t_d = bench(lambda: JC.conv_dgrad(dz, w_, xs, ups, stride, pad, reflect, [True] * len(xs)), iters)
            rec.update(dgrad_ms=t_d, dgrad_tflops=flops / t_d / 1e9)
        t_w = bench(lambda: JC.conv_wgrad(dz, w_, xs, ups, stride, pad, reflect), iters)
        rec.update(wgrad_ms=t_w, wgrad_tflops=flops / t_w / 1e9)
        out.append(rec)
        print(json.dumps(rec))
    return out


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None, int(sys.argv[2]) if len(sys.argv) > 2 else 10)
