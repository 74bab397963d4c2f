"""Micro-benchmark of single tensor-core conv launches through the C ABI (no autograd, pre-split operands): kernel time only.
usage: python tools/conv_micro.py            (a fixed list of model shapes x epilogue options)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops, _lib
from playablevideogeneration_b200._lib import ConvDesc
dev = "cuda"
FMT = _lib.CORR_FP16_ALL
SHAPES = [  # n, cin, cout, h, w, groups
    (8, 128, 128, 64, 64, 1), (48, 128, 128, 64, 64, 6), (8, 64, 64, 32, 32, 1), (8, 288, 128, 16, 16, 1),
    (8, 544, 1024, 16, 16, 0), (120, 64, 128, 128, 128, 0), (120, 64, 64, 256, 256, 0), (120, 256, 256, 64, 64, 0),
    (120, 128, 128, 128, 128, 0), (8, 64, 32, 256, 256, 1), (128, 16, 16, 128, 128, 1)]
flush = torch.empty(48 * 1024 * 1024, device=dev)


def run(n, cin, cout, h, w, bias, act, planes, groups):
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
    wt = torch.randn(cout, cin, 3, 3, device=dev) * (cin * 9) ** -0.5
    b = torch.randn(cout, device=dev) if bias else None
    xp = ops._split(x, 2, FMT)[1]
    packs = ops._get_packs(wt, cin, True, cout)
    wp = packs.lo(0, 2, FMT)
    y = ops.empty_nhwc((n, cout, h, w), dev)
    yp = torch.empty((2 * y.numel(),), dtype=torch.float16, device=dev) if planes else None
    sums = torch.zeros((groups, 2, cout), dtype=torch.float64, device=dev) if groups else None
    d = ConvDesc(n, h, w, cin, cout, 3, 3, 1, act, 0.0, _lib.ALGO_UMMA, 2, FMT)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.call("pvg_conv2d_fwd_planes", d, xp.data_ptr(), wp.data_ptr(), b.data_ptr() if bias else None, y.data_ptr(),
                 yp.data_ptr() if planes else None, None, sums.data_ptr() if groups else None, groups, None, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[1] * 1e3


def run_lstm(n, cin, hidden, h, w):
    cout = 4 * hidden
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
    wt = torch.randn(cout, cin, 3, 3, device=dev) * (cin * 9) ** -0.5
    b = torch.randn(cout, device=dev)
    xp = ops._split(x, 2, FMT)[1]
    wp = ops._get_packs(wt, cin, True, cout).lo(0, 2, FMT)
    c_prev = ops.empty_nhwc((n, hidden, h, w), dev).normal_()
    c_new, h_new = torch.empty_like(c_prev), torch.empty_like(c_prev)
    gates = ops.empty_nhwc((n, cout, h, w), dev)
    d = ConvDesc(n, h, w, cin, cout, 3, 3, 1, 0, 0.0, _lib.ALGO_UMMA, 2, FMT)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.call("pvg_convlstm_step", d, xp.data_ptr(), wp.data_ptr(), b.data_ptr(), c_prev.data_ptr(), c_new.data_ptr(),
                 h_new.data_ptr(), gates.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[1] * 1e3


if len(sys.argv) > 1 and sys.argv[1] == "lstm":
    for n, cin, hidden, h, w in ((8, 544, 256, 16, 16), (8, 288, 128, 32, 32), (8, 224, 128, 32, 32), (64, 544, 256, 16, 16)):
        us = run_lstm(n, cin, hidden, h, w)
        fl = 2.0 * n * h * w * 4 * hidden * 9 * cin
        print(f"lstm N={n} {h}x{w} {cin}->{4 * hidden}: {us:8.1f} us ({fl / us / 1e6:6.1f} TF/s)", flush=True)
    sys.exit(0)
if len(sys.argv) > 6:          # one shape, one epilogue mode (for ncu): n cin cout h w mode[plain|bias+relu|planes|sums]
    n, cin, cout, h, w = (int(v) for v in sys.argv[1:6])
    mode = sys.argv[6]
    us = run(n, cin, cout, h, w, mode != "plain", _lib.ACT_RELU if mode in ("bias+relu", "planes") else 0, mode == "planes",
             1 if mode == "sums" else 0)
    print(mode, us, "us")
    sys.exit(0)
for n, cin, cout, h, w, groups in SHAPES:
    fl = 2.0 * n * h * w * cout * 9 * cin
    res = []
    for name, bias, act, planes, g in (("plain", False, 0, False, 0), ("bias+relu", True, _lib.ACT_RELU, False, 0),
                                       ("bias+relu+planes", True, _lib.ACT_RELU, True, 0), ("bn_sums", False, 0, False, groups)):
        if name == "bn_sums" and not groups:
            continue
        us = run(n, cin, cout, h, w, bias, act, planes, g)
        res.append(f"{name} {us:8.1f} us ({fl / us / 1e6:6.1f} TF/s)")
    print(f"N={n:3d} {h}x{w} {cin}->{cout}: " + " | ".join(res), flush=True)
