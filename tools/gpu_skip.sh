#!/bin/bash
for v in 0 2; do for sk in 0 1 2 3; do
echo "== variant $v skip $sk"
JPB_CONV_VARIANT=$v JPB_CONV_SKIP=$sk timeout 120 python - <<PY
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import bench_conv as BC, torch, json
from jperceiver_b200 import conv as JC
CL=torch.channels_last; dev=torch.device('cuda:0')
for name, srcs, cout, k, stride, pad, reflect, act in BC.LAYERS:
    if not any(t in name for t in ("merge1","layout layer2 128","layout layer1 64")): continue
    g=torch.Generator().manual_seed(0)
    xs=[torch.randn(4,c,h,w,generator=g).to(dev).contiguous(memory_format=CL) for c,h,w,up in srcs]
    ups=[bool(u) for *_,u in srcs]; cin=sum(c for c,*_ in srcs)
    w_=(torch.randn(cout,cin,k,k,generator=g)/(cin*k*k)**0.5).to(dev).contiguous(memory_format=CL)
    b_=torch.randn(cout,generator=g).to(dev)
    t=BC.bench(lambda: JC.conv2d_tc(xs,ups,w_,b_,stride,pad,reflect,act,None),10)
    print("%-40s %7.1f us"%(name[:40],t*1e3))
PY
done; done
