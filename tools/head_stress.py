"""Stress / sanitizer target for the shared-memory tiled head kernels (conv_direct.cu): repeated fwd + bwd of the 7x7 and
3x3 three-channel heads against cuDNN fp32, interleaved with tensor-core convs.  usage: head_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from playablevideogeneration_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
worst = {}
for it in range(reps):
    for (n, cin, cout, h, w, k) in [(1, 32, 3, 64, 64, 7), (2, 32, 3, 40, 70, 7), (2, 64, 3, 20, 68, 3), (1, 16, 3, 21, 36, 7)]:
        x = torch.randn(n, cin, h, w, device=dev, generator=g).requires_grad_(True)
        wt = (torch.randn(cout, cin, k, k, device=dev, generator=g) * (cin * k * k) ** -0.5).requires_grad_(True)
        b = torch.randn(cout, device=dev, generator=g).requires_grad_(True)
        xo = torch.randn(2, 64, 16, 16, device=dev, generator=g)              # a tensor-core conv in between (allocator churn)
        ops.conv2d(xo, torch.randn(64, 64, 3, 3, device=dev, generator=g) * 0.04)
        y = ops.conv2d(x, wt, b, act=3)
        gy = torch.randn(y.shape, device=dev, generator=g)
        y.backward(gy)
        xr, wr, br = x.detach().clone().requires_grad_(True), wt.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        yr = torch.tanh(F.conv2d(xr, wr, br, padding=k // 2))
        yr.backward(gy)
        for name, a, r in (("y", y, yr), ("dx", x.grad, xr.grad), ("dw", wt.grad, wr.grad), ("db", b.grad, br.grad)):
            e = float((a - r).abs().max() / r.abs().max())
            key = (name, cin, cout, k)
            worst[key] = max(worst.get(key, 0.0), e)
            if e > 2e-5:
                print(f"iter {it}: {key} rel err {e:.3e}  bad {(int(((a - r).abs() > 2e-5 * r.abs().max()).sum()))} / {a.numel()}", flush=True)
torch.cuda.synchronize()
print("worst:", {str(k): f"{v:.1e}" for k, v in worst.items()})
