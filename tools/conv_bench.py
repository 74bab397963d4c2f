"""Per-layer micro-benchmark of the tensor-core conv kernels on the BAIR-256 training shapes (CUDA events, L2 flushed
between repetitions by the 120+ MB activation tensors themselves).  usage: python tools/conv_bench.py [precision] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import ops

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ops.set_precision(prec)
dev = "cuda"
# name, N, Cin(phys), Cout, H, W, k
SHAPES = [
    ("vgg1_2 64->64 @256 x30", 30, 64, 64, 256, 256, 3),
    ("vgg2_2 128->128 @128 x30", 30, 128, 128, 128, 128, 3),
    ("vgg3_2 256->256 @64 x30", 30, 256, 256, 64, 64, 3),
    ("vgg4_2 512->512 @32 x30", 30, 512, 512, 32, 32, 3),
    ("vgg5_1 512->512 @16 x120", 120, 512, 512, 16, 16, 3),
    ("D up0 128->128 @64 x8", 8, 128, 128, 64, 64, 3),
    ("D up1 128->64 @128 x8", 8, 128, 64, 128, 128, 3),
    ("D up2 64->32 @256 x8", 8, 64, 32, 256, 256, 3),
    ("D final 32->3 k7 @256 x8", 8, 32, 3, 256, 256, 7),
    ("LSTM0 224->512 @32 x8", 8, 224, 512, 32, 32, 3),
    ("LSTM1 544->1024 @16 x8", 8, 544, 1024, 16, 16, 3),
    ("R same 160->256 @32 x8", 8, 160, 256, 32, 32, 3),
    ("E res 64->64 @32 x128", 128, 64, 64, 32, 32, 3),
    ("E res 32->32 @64 x128", 128, 32, 32, 64, 64, 3),
]
out = []
flush = torch.empty(64 * 1024 * 1024, device=dev)          # 256 MB > L2
for name, n, cin, cout, h, w, k in SHAPES:
    x = ops.empty_nhwc((n, cin, h, w), dev).normal_()
    wt = torch.randn(cout, cin, k, k, device=dev) * (cin * k * k) ** -0.5
    y = ops.conv2d(x, wt)                                    # warm-up (packs weights, sets attributes)
    torch.cuda.synchronize()
    ops.conv_profile = []
    for _ in range(reps):
        flush.zero_()
        ops.conv2d(x, wt)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b, _ in ops.conv_profile)
    ops.conv_profile = None
    med = ms[len(ms) // 2]
    flops = 2.0 * n * h * w * cout * k * k * cin
    out.append(dict(layer=name, ms=med, tflops=flops / med / 1e9, gflop=flops / 1e9))
    print(f"{name:28s} {med:8.3f} ms  {flops / med / 1e9:7.1f} TFLOP/s (algorithmic)  [{prec}]", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/conv_bench_{prec}.json", "w"), indent=1)
