"""Per-entry-point breakdown of one eager training step of the bench workload: CUDA events around every C-ABI call,
aggregated by (entry point, shape).  usage: python tools/step_profile.py [precision] [out.json]"""
import os, sys, json, random, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from playablevideogeneration_b200.configs import build_config
from playablevideogeneration_b200 import _lib, ops
from playablevideogeneration_b200.caddy import Model
from playablevideogeneration_b200.training.step import TrainStep
from playablevideogeneration_b200.vgg import Vgg19

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
ops.set_precision(prec)
w = bench.WORKLOADS["bair256_b8_t16"]
dev = torch.device("cuda")
cfg = build_config(dict(config=w["config"], H=w["H"], W=w["W"], S=w["S"]))
torch.manual_seed(0); random.seed(0)
model = Model(cfg).to(dev)
vgg = Vgg19(allow_random_init=True)
step = TrainStep(cfg, model, vgg)
batch = tuple(t.to(dev) for t in bench.synthetic_batch(w))
for _ in range(2):
    step.step(batch, w["gt_init"], 1.0)
torch.cuda.synchronize()

records = []
orig = _lib.call
def timed(name, *args):
    key = name
    if args and isinstance(args[0], _lib.ConvDesc):
        d = args[0]
        key = f"{name} N{d.N} {d.H}x{d.W} {d.Cin}->{d.Cout} k{d.R} algo{d.algo}"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = orig(name, *args)
    e1.record()
    records.append((key, e0, e1))
    return rc
ops.call = timed
_lib.call = timed
import playablevideogeneration_b200.training.step as S
if hasattr(S, "call"):
    S.call = timed
torch.cuda._sleep(int(4e9))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
step.step(batch, w["gt_init"], 1.0)
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, a, b in records:
    ms = a.elapsed_time(b)
    c = agg.setdefault(key, [0, 0.0])
    c[0] += 1; c[1] += ms
tot = sum(v[1] for v in agg.values())
byname = collections.Counter()
for k, v in agg.items():
    byname[k.split()[0]] += v[1]
print(f"step (events, incl. glue) {t0.elapsed_time(t1):.1f} ms ; sum of C-ABI calls {tot:.1f} ms ; {len(records)} calls")
for k, v in byname.most_common():
    print(f"  {k:28s} {v:8.2f} ms")
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"  {v[1]:8.2f} ms  x{v[0]:<4d} {k}")
if len(sys.argv) > 2:
    json.dump(dict(step_ms=t0.elapsed_time(t1), calls={k: dict(n=v[0], ms=v[1]) for k, v in agg.items()}), open(sys.argv[2], "w"), indent=1)
