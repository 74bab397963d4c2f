"""Which tensor-core convs with a channel count that is not a multiple of 32 lose accuracy with fp16 correction planes on
the pretrain_bair test case?  Re-runs every such conv of one forward pass with tf32 / fp16 corrections against float64."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests.golden_util import batch_tuple, case_inputs, load_case
from tests.test_model_gpu import _build, _to_dev
from playablevideogeneration_b200 import ops

case, g = load_case("pretrain_bair")
cfg, sd, vgg_sd, obs = case_inputs(case)
model, step = _build(case, cfg, sd, vgg_sd)
model.train()
seen = []
orig = ops.conv2d


def spy(x, weight, bias=None, act=0, slope=0.0):
    y = orig(x, weight, bias, act, slope)
    if x.shape[1] % 32 != 0 and x.shape[1] % 8 == 0 and len(seen) < 12:
        xd, wd = x.detach().double().cpu(), weight.detach().double().cpu()
        ref = F.conv2d(xd[:, :weight.shape[1]], wd, None if bias is None else bias.detach().double().cpu(), padding=weight.shape[2] // 2)
        res = {}
        for corr in ("tf32", "fp16", "bf16"):
            ops.set_correction(corr, corr, corr)
            yy = orig(x.detach(), weight.detach(), None if bias is None else bias.detach(), 0, 0.0).double().cpu()
            res[corr] = float((yy - ref).norm() / ref.norm())
        ops.set_correction()
        seen.append((tuple(x.shape), tuple(weight.shape), float(x.abs().max()), float(x.abs().mean()), float(weight.abs().max()),
                     float(weight.abs().mean()), res))
    return y


ops.conv2d = spy
import playablevideogeneration_b200.caddy as C
for mod in (C,):
    if hasattr(mod, "conv2d"):
        mod.conv2d = spy
torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
total, info, res = step.compute_losses(_to_dev(batch_tuple(obs)), case["gt_init"], case["gumbel_temperature"], pretraining=True)
for s in seen:
    print("x", s[0], "w", s[1], f"|x|max {s[2]:.3e} mean {s[3]:.3e}  |w|max {s[4]:.3e} mean {s[5]:.3e}  relL2 err:", {k: f"{v:.2e}" for k, v in s[6].items()})
print("convs spied:", len(seen))
