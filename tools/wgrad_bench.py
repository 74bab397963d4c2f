"""Per-layer micro-benchmark of the tensor-core weight-gradient kernel on the BAIR-256 model shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
ops.set_precision(prec)
dev = "cuda"
SHAPES = [("D up0 128->128 @64 x8", 8, 128, 128, 64, 64, 3), ("D res 128->128 @64 x8", 8, 128, 128, 64, 64, 3),
          ("D up1 128->64 @128 x8", 8, 128, 64, 128, 128, 3), ("D res 64->64 @128 x8", 8, 64, 64, 128, 128, 3),
          ("D up2 64->32 @256 x8", 8, 64, 32, 256, 256, 3), ("D final 32->3 k7 @256 x8", 8, 32, 3, 256, 256, 7),
          ("LSTM0 224->512 @32 x8", 8, 224, 512, 32, 32, 3), ("LSTM1 544->1024 @16 x8", 8, 544, 1024, 16, 16, 3),
          ("LSTM2 288->512 @32 x8", 8, 288, 512, 32, 32, 3), ("R same 160->256 @32 x8", 8, 160, 256, 32, 32, 3),
          ("E res 64->64 @32 x128", 128, 64, 64, 32, 32, 3), ("E res 32->32 @64 x128", 128, 32, 32, 64, 64, 3),
          ("E res 64->65 @32 x128", 128, 64, 65, 32, 32, 3)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for name, n, cin, cout, h, w, k in SHAPES:
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
    wt = (torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5).requires_grad_(True)
    y = ops.conv2d(x, wt)
    g = torch.randn_like(y)
    y.backward(g); torch.cuda.synchronize()
    ops.wgrad_profile = []
    for _ in range(3):
        flush.zero_()
        y = ops.conv2d(x, wt)
        y.backward(g)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b, _ in ops.wgrad_profile) or [float("nan")]
    ops.wgrad_profile = None
    flops = 2.0 * n * h * w * cout * k * k * cin
    print(f"{name:28s} wgrad {ms[len(ms)//2]:8.3f} ms  {flops / ms[len(ms)//2] / 1e9:7.1f} TFLOP/s (algorithmic) [{prec}]", flush=True)
