"""BatchNorm (+ pool / activation) forward and backward through ops.pool_bn_act on model shapes: kernel time per call.
usage: python tools/bn_micro.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops
dev = "cuda"
SHAPES = [(8, 32, 256, 256, 1, False), (8, 64, 128, 128, 1, False), (8, 128, 64, 64, 1, False), (128, 16, 128, 128, 1, False),
          (128, 16, 128, 128, 1, True), (48, 128, 64, 64, 6, False), (8, 256, 16, 16, 1, False), (128, 72, 32, 32, 1, False)]
flush = torch.empty(48 * 1024 * 1024, device=dev)
for n, c, h, w, groups, pool in SHAPES:
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    x = ops.empty_nhwc((n, c, h, w), dev).normal_().requires_grad_(True)
    oh, ow = (h // 2, w // 2) if pool else (h, w)
    gy = ops.empty_nhwc((n, c, oh, ow), dev).normal_()
    tf, tb = [], []
    for it in range(5):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda._sleep(3_000_000)           # the host queues the whole call while the GPU spins: events bracket GPU time only
        e0.record()
        y = ops.pool_bn_act(x, bn, pool=pool, act=ops.ACT_LRELU, groups=groups, planes=ops.conv_input_planes())
        e1.record()
        flush.zero_()
        torch.cuda._sleep(3_000_000)
        e1b = torch.cuda.Event(enable_timing=True); e1b.record()
        y.backward(gy)
        e2.record(); torch.cuda.synchronize()
        tf.append(e0.elapsed_time(e1)); tb.append(e1b.elapsed_time(e2))
        x.grad = None
    f, b = sorted(tf)[1] * 1e3, sorted(tb)[1] * 1e3
    mb = n * c * h * w * 4 / 1e6
    print(f"N={n:3d} C={c:3d} {h}x{w} groups={groups} pool={int(pool)}: fwd {f:7.1f} us  bwd {b:7.1f} us   (tensor {mb:6.1f} MB)", flush=True)
