"""VGG conv1_1 (3 -> 64, 3x3, N frames of 256x256) forward with the operand planes of its output: stem kernel vs the
previous path (per-pixel kernel + split pass).  usage: python tools/stem_bench.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
x = ops.empty_nhwc((n, 3, 256, 256), "cuda").normal_()
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.2
b = torch.randn(64, device="cuda") * 0.1
w2 = torch.randn(64, 64, 3, 3, device="cuda") * 0.05
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for stem in (True, False):
    ops.stem_kernel = stem
    ts = []
    for it in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = ops.conv2d(x, w, b, act=ops.ACT_RELU, out_planes=True)
        pl = ops.planes_of(y) or {2: ops._split(y, 2, 2)[1]}
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del y, pl
    t = sorted(ts)[1]
    gb = n * 65536 * 64 * 8 / 1e9
    print(f"stem_kernel={stem}: conv1_1 + planes {t:.3f} ms  ({gb / t * 1e3:.0f} GB/s of y + planes written)", flush=True)
