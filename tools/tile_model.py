"""Fits time-per-tile = a + b * k_iters of the tensor-core conv kernel: N=120 @64x64, Cout=128 (one N tile), Cin sweeps.
usage: tile_model.py [precision]   (PVG_2CTA=0/1 selects the kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import _lib, ops
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ops.set_precision(prec)
dev = "cuda"
n, h, w = 120, 64, 64
flush = torch.empty(64 * 1024 * 1024, device=dev)
rows = []
for cin in (32, 64, 128, 256, 512):
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
    wt = torch.randn(cout, cin, 3, 3, device=dev) * (cin * 9) ** -0.5
    algo, nprod, fmt = ops._conv_algo(cin, cout, 3, "fwd")
    packs = ops._get_packs(wt, cin, True)
    split = ops._split(x, nprod, fmt) if nprod >= 2 else None
    y = ops._conv_forward(x, packs, 0, cout, 3, None, 0, 0.0, algo, nprod, fmt, split=split)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops._conv_forward(x, packs, 0, cout, 3, None, 0, 0.0, algo, nprod, fmt, split=split)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[2]
    k_iters = 9 * cin // 32
    tiles_per_sm = (n * h * w / 128) * ((cout + 127) // 128) / 148
    rows.append((k_iters, ms * 1e3 / tiles_per_sm))
    print(f"Cin {cin:4d} k_iters {k_iters:4d}: {ms:7.3f} ms  {2.0 * n * h * w * cout * 9 * cin / ms / 1e9:7.1f} TF/s  per-tile {ms * 1e3 / tiles_per_sm:7.2f} us", flush=True)
(k0, t0), (k1, t1) = rows[1], rows[-1]
b = (t1 - t0) / (k1 - k0)
print(f"[{prec} 2CTA={os.environ.get('PVG_2CTA','auto')} cout={cout}] per k-iter {b * 1e3:.0f} ns = {b * 1965:.0f} clk ; fixed per tile {t0 - b * k0:.2f} us")
