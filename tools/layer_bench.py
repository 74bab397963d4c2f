"""Per-layer fwd + bwd micro-benchmark through the C ABI: CUDA events around every C-ABI call of one conv layer
(forward, activation backward / splits, data gradient, weight gradient), median of 5 with an L2 flush in between.
usage: layer_bench.py [precision] [set]     set = heads | vgg | model | all"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import _lib, ops

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
which = sys.argv[2] if len(sys.argv) > 2 else "all"
ops.set_precision(prec)
dev = "cuda"
SETS = {
    "heads": [("head 32->3 k7 @256 x8", 8, 32, 3, 256, 256, 7, True), ("head 64->3 @128 x8", 8, 64, 3, 128, 128, 3, True),
              ("head 128->3 @64 x8", 8, 128, 3, 64, 64, 3, True), ("vgg1_1 3->64 @256 x120", 120, 3, 64, 256, 256, 3, True),
              ("E stem 3->16 @256 x128", 128, 3, 16, 256, 256, 3, True)],
    "enc": [("E res 16->16 @128 x128", 128, 16, 16, 128, 128, 3, True), ("E res 16->32 @128 x128", 128, 16, 32, 128, 128, 3, True),
            ("E res 16->16 @128 x8", 8, 16, 16, 128, 128, 3, True), ("E res 32->32 @64 x128", 128, 32, 32, 64, 64, 3, True)],
    "vgg": [("vgg1_2 64->64 @256 x120", 120, 64, 64, 256, 256, 3, True), ("vgg2_2 128->128 @128 x120", 120, 128, 128, 128, 128, 3, True),
            ("vgg3_2 256->256 @64 x120", 120, 256, 256, 64, 64, 3, True), ("vgg4_2 512->512 @32 x120", 120, 512, 512, 32, 32, 3, True)],
    "model": [("D up2 64->32 @256 x8", 8, 64, 32, 256, 256, 3, True), ("D res 64->64 @128 x8", 8, 64, 64, 128, 128, 3, True),
              ("D up0 128->128 @64 x8", 8, 128, 128, 64, 64, 3, True), ("LSTM1 544->1024 @16 x8", 8, 544, 1024, 16, 16, 3, True)],
}
shapes = sum(SETS.values(), []) if which == "all" else SETS[which]
flush = torch.empty(64 * 1024 * 1024, device=dev)
records = []
orig = _lib.call


def timed(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = orig(name, *args)
    e1.record()
    key = name
    if args and isinstance(args[0], _lib.ConvDesc):
        d = args[0]
        key = f"{name}[{d.Cin}->{d.Cout} algo{d.algo}]"
    records.append((key, e0, e1))
    return rc


ops.call = timed
_lib.call = timed
for name, n, cin, cout, h, w, k, need_dx in shapes:
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_().requires_grad_(need_dx)
    wt = (torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5).requires_grad_(True)
    y = ops.conv2d(x, wt)
    g = torch.randn_like(y)
    y.backward(g)
    torch.cuda.synchronize()
    per = collections.defaultdict(list)
    for _ in range(5):
        flush.zero_()
        records.clear()
        y = ops.conv2d(x, wt)
        y.backward(g)
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for key, a, b in records:
            agg[key] = agg.get(key, 0.0) + a.elapsed_time(b)
        for key, v in agg.items():
            per[key].append(v)
    flops = 2.0 * n * h * w * cout * k * k * cin
    tot = 0.0
    parts = []
    for key, v in per.items():
        v.sort()
        m = v[len(v) // 2]
        tot += m
        parts.append(f"{key.replace('pvg_', '')} {m:.3f}")
    print(f"{name:28s} [{prec}] total {tot:7.3f} ms ({3 * flops / tot / 1e9:6.1f} TF/s fwd+bwd) | " + " | ".join(parts), flush=True)
