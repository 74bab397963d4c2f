"""One weight-gradient launch (for ncu): python tools/one_wgrad.py N Cin Cout H W [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops
n, cin, cout, h, w = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = "cuda"
x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
wt = (torch.randn(cout, cin, 3, 3, device=dev) * (cin * 9) ** -0.5).requires_grad_(True)
for _ in range(reps):
    y = ops.conv2d(x, wt)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok", float(wt.grad.abs().mean()))
