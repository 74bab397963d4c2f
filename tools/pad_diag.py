"""Diagnostic: gradients of one training step with / without the physically padded 65-channel block."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import batch_tuple, case_inputs, load_case
from playablevideogeneration_b200 import ops
from playablevideogeneration_b200.caddy import Model
from playablevideogeneration_b200.training.step import TrainStep
from playablevideogeneration_b200.vgg import Vgg19
case, g = load_case("full_bair_feedback")
cfg, sd, vgg_sd, obs = case_inputs(case)
res = {}
for mode in ("pad", "nopad"):
    os.environ["PVG_NO_COUT_PAD"] = "1" if mode == "nopad" else "0"
    model = Model(cfg); model.load_state_dict({k: v.clone() for k, v in sd.items()}); model = model.cuda()
    step = TrainStep(cfg, model, Vgg19(vgg_sd))
    bt = tuple(t.cuda() for t in batch_tuple(obs))
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    model.train()
    total, info, r = step.compute_losses(bt, case["gt_init"], case["gumbel_temperature"])
    step.arena.zero_grad(); total.backward()
    res[mode] = ({k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}, float(total.detach().cpu()[0]),
                 set(step.arena.touched))
print("loss", res["pad"][1], res["nopad"][1], "touched equal", res["pad"][2] == res["nopad"][2], len(res["pad"][2]), len(res["nopad"][2]))
rows = []
for k in res["pad"][0]:
    a, b = res["pad"][0][k].double(), res["nopad"][0][k].double()
    rows.append((float((a - b).norm() / (b.norm() + 1e-300)), k, float(b.norm()), int((a == 0).sum()), int((b == 0).sum()), a.numel()))
rows.sort(reverse=True)
for r in rows[:15]:
    print("%.3e %-70s |g| %.3e zeros pad/nopad %d/%d of %d" % r)
