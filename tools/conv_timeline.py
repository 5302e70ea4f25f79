"""Per-CTA timeline of the tcgen05 convolution forward kernel (globaltimer stamps written by the kernel's debug hook):
where one tile's time goes — setup, offset table, first operands, main loop, epilogue.  python tools/conv_timeline.py <layer substring>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from jperceiver_b200 import conv as JC
import tools.bench_conv as BC
CL = torch.channels_last
dev = torch.device("cuda:0")
only = sys.argv[1] if len(sys.argv) > 1 else "layer1 64"
for name, srcs, cout, k, stride, pad, reflect, act in BC.LAYERS:
    if only not in name:
        continue
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(4, c, h, w, generator=g).to(dev).contiguous(memory_format=CL) for c, h, w, up in srcs]
    ups = [bool(u) for *_, u in srcs]
    cin = sum(c for c, *_ in srcs); cin_w = 3 if k == 7 else cin
    w_ = (torch.randn(cout, cin_w, k, k, generator=g) / (cin_w * k * k) ** 0.5).to(dev).contiguous(memory_format=CL)
    b_ = torch.randn(cout, generator=g).to(dev)
    for _ in range(3):
        JC.conv2d_tc(xs, ups, w_, b_, stride, pad, reflect, act, None)
    st = torch.zeros(512, 6, 8, dtype=torch.int64, device=dev)
    JC.DBG_STAMPS = st
    JC.conv2d_tc(xs, ups, w_, b_, stride, pad, reflect, act, None)
    JC.DBG_STAMPS = None
    torch.cuda.synchronize()
    s = st.cpu()
    t0 = s[:, :, 0][s[:, :, 0] > 0].min()
    print(name)
    print("cta: start | setup(sync1) offsets(sync2) | producer: loop end, accum ready, epilogue end | mma: first full, last issue | exit   (us, relative to kernel start)")
    for cta in list(range(0, 6)) + list(range(296, 300)) + list(range(506, 512)):
        if s[cta, 0, 0] == 0:
            continue
        r = lambda w, i: (int(s[cta, w, i]) - int(t0)) / 1e3
        print("%4d: %7.2f | %7.2f %7.2f | %7.2f %7.2f %7.2f | %7.2f %7.2f | %7.2f" % (cta, r(0, 0), r(0, 1), r(0, 2), r(0, 3), r(0, 4), r(0, 5), r(5, 3), r(5, 4), r(0, 6)))
