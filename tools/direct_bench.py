"""Micro-benchmark of the image-facing conv layers (direct kernels vs the generic paths).  usage: direct_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops
dev = "cuda"
SHAPES = [("E stem 3->16 @256 x128", 128, 3, 16, 256, 256, 3), ("E stem dgrad 16->3 @256 x8", 8, 16, 3, 256, 256, 3),
          ("vgg1_1 3->64 @256 x30", 30, 3, 64, 256, 256, 3), ("vgg1_1 dgrad 64->3 @256 x30", 30, 64, 3, 256, 256, 3),
          ("head 128->3 @64 x8", 8, 128, 3, 64, 64, 3), ("head 64->3 @128 x8", 8, 64, 3, 128, 128, 3),
          ("head 32->3 k7 @256 x8", 8, 32, 3, 256, 256, 7), ("head dgrad 3->128 @64 x8", 8, 3, 128, 64, 64, 3),
          ("head dgrad 3->64 @128 x8", 8, 3, 64, 128, 128, 3), ("head dgrad 3->32 k7 @256 x8", 8, 3, 32, 256, 256, 7)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for name, n, cin, cout, h, w, k in SHAPES:
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
    wt = torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5
    ops.conv2d(x, wt); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv2d(x, wt); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    byts = (n * h * w * (cin + cout)) * 4
    print(f"{name:32s} {ts[2]:8.3f} ms   {byts / ts[2] / 1e6:8.1f} GB/s of algorithmic bytes   direct={os.environ.get('PVG_NO_DIRECT','0')!='1'}", flush=True)
