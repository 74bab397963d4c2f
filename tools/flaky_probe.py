"""Repeat the 7x7 / small-N tensor-core conv several times and report the error each time (race hunting)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from playablevideogeneration_b200 import ops
dev = "cuda"
def rnd(*s, seed): return torch.randn(s, generator=torch.Generator().manual_seed(seed))
for (n, cin, cout, h, w, k) in [(1, 32, 3, 64, 64, 7), (1, 128, 3, 16, 16, 3), (2, 64, 16, 32, 32, 3), (2, 32, 32, 32, 32, 3), (1, 64, 64, 64, 64, 3)]:
    x = rnd(n, cin, h, w, seed=1); wt = rnd(cout, cin, k, k, seed=2) * (cin * k * k) ** -0.5; b = rnd(cout, seed=3)
    ref = F.conv2d(x, wt, b, padding=k // 2)
    xg, wg, bg = x.to(dev), wt.to(dev), b.to(dev)
    errs = []
    outs = []
    for it in range(12):
        y = ops.conv2d(xg, wg, bg)
        outs.append(y.cpu())
        errs.append(float((outs[-1] - ref).abs().max()))
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print((n, cin, cout, h, w, k), "errs", ["%.1e" % e for e in errs], "bitwise-identical" if same else "NON-DETERMINISTIC")
    if not same:
        d = (outs[0] - outs[[i for i, o in enumerate(outs) if not torch.equal(outs[0], o)][0]]).abs()
        idx = (d > 0).nonzero()
        print("   differing elements:", idx.shape[0], "first", idx[:5].tolist(), "max diff", float(d.max()))
