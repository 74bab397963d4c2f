"""bf16- vs tf32-evaluated correction terms of the split product on VGG-like shapes and data distributions:
relative L2 error of the conv output against a float64 CPU convolution."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from playablevideogeneration_b200 import ops

dev = "cuda"
SHAPES = [(8, 64, 64, 16, 16), (8, 128, 128, 8, 8), (8, 256, 256, 4, 4), (8, 512, 512, 2, 2), (8, 512, 512, 1, 1),
          (6, 512, 512, 2, 2), (8, 128, 64, 8, 8), (8, 512, 256, 2, 2), (8, 64, 64, 64, 64)]
DATA = {"randn": lambda t: t, "relu": lambda t: t.clamp_min(0), "tiny": lambda t: t * 1e-7,
        "sparse_sign": lambda t: torch.sign(t) * (t.abs() > 1.5) * 1e-6}
g = torch.Generator().manual_seed(0)
for (n, cin, cout, h, w) in SHAPES:
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (cin * 9) ** -0.5
    for dname, f in DATA.items():
        x = f(torch.randn(n, cin, h, w, generator=g))
        ref = F.conv2d(x.double(), wt.double(), padding=1)
        out = {}
        for corr in ("tf32", "bf16", "fp16"):
            ops.set_correction(corr, corr, corr)
            y = ops.conv2d(x.to(dev), wt.to(dev)).cpu().double()
            out[corr] = float((y - ref).norm() / ref.norm().clamp_min(1e-300))
        ref32 = F.conv2d(x, wt, padding=1).double()
        e32 = float((ref32 - ref).norm() / ref.norm().clamp_min(1e-300))
        print(f"N{n} {cin}->{cout} @{h}x{w} {dname:12s} relL2: cpu-fp32 {e32:.2e}  tf32-corr {out['tf32']:.2e}  bf16-corr {out['bf16']:.2e}  fp16-corr {out['fp16']:.2e}", flush=True)

print("--- signed bias: all-positive operands, mean((y - ref) / ref) ---")
for (n, cin, cout, h, w) in [(8, 64, 64, 16, 16), (8, 512, 512, 2, 2)]:
    for wname, wf in (("w>0", lambda t: t.abs()), ("w random", lambda t: t)):
        wt = wf(torch.randn(cout, cin, 3, 3, generator=g)) * (cin * 9) ** -0.5
        x = torch.randn(n, cin, h, w, generator=g).abs()
        ref = F.conv2d(x.double(), wt.double(), padding=1)
        for corr in ("tf32", "bf16", "fp16"):
            ops.set_correction(corr, corr, corr)
            y = ops.conv2d(x.to(dev), wt.to(dev)).cpu().double()
            scale = ref.abs().mean()
            print(f"N{n} {cin}->{cout} @{h}x{w} {wname:9s} {corr}: mean signed err / mean|ref| = {float((y - ref).mean() / scale):+.3e}   "
                  f"rms err / mean|ref| = {float((y - ref).pow(2).mean().sqrt() / scale):.3e}", flush=True)
        y32 = F.conv2d(x, wt, padding=1).double()
        print(f"   cpu fp32: mean signed {float((y32 - ref).mean() / scale):+.3e}")

print("--- error that is coherent over pixels: rms over output channels of mean_pixels(y - ref), / rms(ref); relu inputs ---")
for (n, cin, cout, h, w) in [(8, 64, 64, 32, 32), (8, 256, 256, 8, 8)]:
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (cin * 9) ** -0.5
    x = torch.randn(n, cin, h, w, generator=g).clamp_min(0)
    ref = F.conv2d(x.double(), wt.double(), padding=1)
    rms = ref.pow(2).mean().sqrt()
    for corr in ("tf32", "bf16", "fp16"):
        ops.set_correction(corr, corr, corr)
        y = ops.conv2d(x.to(dev), wt.to(dev)).cpu().double()
        coh = (y - ref).mean(dim=(0, 2, 3)).pow(2).mean().sqrt() / rms
        print(f"N{n} {cin}->{cout} @{h}x{w} {corr}: coherent {float(coh):.3e}", flush=True)
    y32 = F.conv2d(x, wt, padding=1).double()
    print(f"   cpu fp32: coherent {float((y32 - ref).mean(dim=(0, 2, 3)).pow(2).mean().sqrt() / rms):.3e}")
