"""Reproduces the test order in which conv_umma3[bf16](1,32,32,3,64,64,7,True,3) fails and localises the corruption."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import tests.test_kernels_gpu as T
from playablevideogeneration_b200 import ops

DEV = "cuda"
for shape in T.CONV_SIMT_SHAPES:
    T.test_conv_simt_forward(shape)
T.test_tf32_probe_reports_rounding_mode()
ops.set_correction("bf16", "bf16", "bf16")
for i, shape in enumerate(T.CONV_UMMA_SHAPES[:6]):
    x, wt, b, ref = T._conv_case(shape)
    ops.set_precision("tf32x3")
    xd, wd, bd = x.to(DEV), wt.to(DEV), (b.to(DEV) if b is not None else None)
    got = ops.conv2d(xd, wd, bd, act=shape[8])
    e = float((got.cpu() - ref).abs().max())
    print(i, shape, "err", f"{e:.3e}", flush=True)
    if e > 1e-5:
        got2 = ops.conv2d(xd, wd, bd, act=shape[8])
        print("  second call err", f"{float((got2.cpu() - ref).abs().max()):.3e}", " first vs second", f"{float((got - got2).abs().max()):.3e}")
        print("  x on device intact:", bool((xd.cpu() == x).all()), " w intact:", bool((wd.cpu() == wt).all()))
        pre_ref = F.conv2d(x[:, :shape[1]], wt, b, padding=shape[6] // 2)
        got_na = ops.conv2d(xd, wd, bd, act=0)
        d = (got_na.cpu() - pre_ref).abs()
        print("  no-activation conv err", f"{float(d.max()):.3e}", "bad", int((d > 1e-4).sum()), "of", d.numel())
        idx = (d > 1e-4).nonzero()
        print("  bad (n, c, h, w) sample:", idx[:12].tolist(), " h range", int(idx[:, 2].min()) if len(idx) else None, int(idx[:, 2].max()) if len(idx) else None,
              " w range", int(idx[:, 3].min()) if len(idx) else None, int(idx[:, 3].max()) if len(idx) else None)
        packs = ops._get_packs(wd, shape[2], False)
        exp = wt.permute(0, 2, 3, 1).reshape(-1)
        print("  fwd pack matches OIHW->OHWI:", bool((packs.hi(0).cpu() == exp).all()))
        g64 = torch.tanh(F.conv2d(x.double()[:, :shape[1]], wt.double(), b.double(), padding=shape[6] // 2)) if shape[8] == 3 else None
        if g64 is not None:
            print("  vs fp64 reference: ours", f"{float((got.cpu().double() - g64).abs().max()):.3e}", " cpu fp32 ref", f"{float((ref.double() - g64).abs().max()):.3e}")
