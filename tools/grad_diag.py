"""Per-parameter gradient comparison of the CUDA path against the CPU oracle on the GPU box (diagnostic, not a test).
usage: python tools/grad_diag.py <case> <precision>"""
import random
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import caddy_oracle as O
from oracle.cases import CASES
from tests.golden_util import batch_tuple, case_inputs

name, precision = sys.argv[1], sys.argv[2]
case = CASES[name]
cfg, sd, vgg_sd, obs = case_inputs(case)
params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")) and "centroid" not in k)
          for k, v in sd.items()}
mi = O.MutualInformation(cfg["data"]["actions_count"], cfg["training"]["mutual_information_estimation_alpha"] if case["smooth_mi"] else None)
torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
total, comp, res = O.compute_losses(params, vgg_sd, cfg, mi, batch_tuple(obs), case["gt_init"], case["gumbel_temperature"],
                                    pretraining=case["mode"] == "pretraining")
for r in res:
    if isinstance(r, torch.Tensor) and r.requires_grad:
        r.retain_grad()
for r in res[1]:
    r.retain_grad()
total.backward()

from playablevideogeneration_b200 import ops
from playablevideogeneration_b200.caddy import Model
from playablevideogeneration_b200.training.step import TrainStep
from playablevideogeneration_b200.vgg import Vgg19
ops.set_precision(precision)
model = Model(cfg, reduced=case.get("reduced", False))
model.load_state_dict({k: v.clone() for k, v in sd.items()})
model = model.cuda().train()
step = TrainStep(cfg, model, Vgg19(vgg_sd))
torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
tot2, info, res2 = step.compute_losses(tuple(t.cuda() for t in batch_tuple(obs)), case["gt_init"], case["gumbel_temperature"],
                                       pretraining=case["mode"] == "pretraining")
for r in res2:
    if isinstance(r, torch.Tensor) and r.requires_grad:
        r.retain_grad()
for r in res2[1]:
    r.retain_grad()
step.arena.zero_grad()
tot2.backward()
print(f"== {name} {precision}: loss {float(tot2):.8f} vs oracle {float(total):.8f}")
print("-- gradients w.r.t. the returned tensors (index: rel err)")
flat1 = list(res[1]) + [r for r in res if isinstance(r, torch.Tensor)]
flat2 = list(res2[1]) + [r for r in res2 if isinstance(r, torch.Tensor)]
for i, (a, b) in enumerate(zip(flat1, flat2)):
    if a.grad is not None and b.grad is not None:
        d = (b.grad.cpu() - a.grad).norm() / (a.grad.norm() + 1e-30)
        print(f"  out[{i}] shape {tuple(a.shape)}: rel {float(d):.2e}  |g| {float(a.grad.norm()):.3e}")
print("-- parameter gradients (registration order)")
for k, p in model.named_parameters():
    ref = params[k].grad
    if ref is None and p.grad is None:
        continue
    got = p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(params[k])
    if ref is None:
        ref = torch.zeros_like(got)
    rel = float((got - ref).norm() / (ref.norm() + 1e-30))
    flag = " <<<" if rel > 1e-3 else ""
    print(f"  {k:75s} rel {rel:.2e} |g| {float(ref.norm()):.3e}{flag}")
