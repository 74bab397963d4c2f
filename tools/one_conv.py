"""One tensor-core conv launch (for ncu): python tools/one_conv.py N Cin Cout H W [k] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops
n, cin, cout, h, w = (int(v) for v in sys.argv[1:6])
k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
ops.set_precision(os.environ.get("PVG_PRECISION", "tf32x3"))
dev = "cuda"
x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
wt = torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5
for _ in range(reps):
    y = ops.conv2d(x, wt)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
