"""Sub-network gradient checks against the CPU oracle (diagnostic)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import caddy_oracle as O
from oracle.cases import CASES, build_config
from playablevideogeneration_b200 import ops
from playablevideogeneration_b200.caddy import Model

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ops.set_precision(prec)
case = CASES["full_bair_feedback"]
cfg = build_config(case)
sd = O.make_weights(cfg, 1)
model = Model(cfg); model.load_state_dict({k: v.clone() for k, v in sd.items()}); model = model.cuda().train()
params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var")) and "centroid" not in k) for k, v in sd.items()}
st = O._State(params, True)
g = torch.Generator().manual_seed(5)
def rnd(*s): return torch.randn(*s, generator=g)

def rel(a, b): return float((a.cpu() - b).norm() / (b.norm() + 1e-30))

def report(tag, pairs):
    print(f"[{tag}] " + "  ".join(f"{n}={rel(a, b):.2e}" for n, a, b in pairs))

def param_report(tag, prefix):
    worst = (0, "")
    for k, p in model.named_parameters():
        if k.startswith(prefix) and p.grad is not None and params[k].grad is not None:
            r = rel(p.grad, params[k].grad)
            if r > worst[0]: worst = (r, k)
    print(f"[{tag}] worst param grad rel err {worst[0]:.2e} ({worst[1]})")
    for k, p in model.named_parameters(): p.grad = None
    for v in params.values(): v.grad = None

# (a) encoder, input requires grad
x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1)
xc = x.clone().requires_grad_(True); xg = x.cuda().requires_grad_(True)
w1, w2 = rnd(2, 64, 8, 8), rnd(2, 1, 8, 8)
s_ref, a_ref = O.encoder(st, xc); ((s_ref * w1).sum() + (a_ref * w2).sum()).backward()
s, a = model.representation_network(xg); ((s * w1.cuda()).sum() + (a * w2.cuda()).sum()).backward()
report("E", [("state", s, s_ref.detach()), ("att", a, a_ref.detach()), ("dx", xg.grad, xc.grad)])
param_report("E", "representation_network")

# (b) decoder
h = rnd(2, 128, 8, 8)
hc = h.clone().requires_grad_(True); hg = h.cuda().requires_grad_(True)
ws = [rnd(2, 3, 64, 64), rnd(2, 3, 32, 32), rnd(2, 3, 16, 16)]
outs_ref = O.decoder(st, hc); sum((o * w).sum() for o, w in zip(outs_ref, ws)).backward()
_, outs = model.rendering_network(hg); sum((o * w.cuda()).sum() for o, w in zip(outs, ws)).backward()
report("D", [(f"o{i}", o, r.detach()) for i, (o, r) in enumerate(zip(outs, outs_ref))] + [("dh", hg.grad, hc.grad)])
param_report("D", "rendering_network")

# (c) dynamics, 3 steps
sts = [rnd(2, 64, 8, 8) for _ in range(3)]
acts = [torch.softmax(rnd(2, 7), 1) for _ in range(3)]
vars_ = [rnd(2, 2) for _ in range(3)]
sc = [t.clone().requires_grad_(True) for t in sts]; sg = [t.cuda().requires_grad_(True) for t in sts]
ac = [t.clone().requires_grad_(True) for t in acts]; ag = [t.cuda().requires_grad_(True) for t in acts]
wo = [rnd(2, 128, 8, 8) for _ in range(3)]
dyn = O.Dynamics(st); lref = 0
for t in range(3): lref = lref + (dyn.step(sc[t], ac[t], vars_[t]) * wo[t]).sum()
lref.backward()
model.dynamics_network.reinit_memory(2); l = 0
for t in range(3): l = l + (model.dynamics_network(sg[t], ag[t], vars_[t].cuda()) * wo[t].cuda()).sum()
l.backward()
report("R", [("loss", l.detach().reshape(1), lref.detach().reshape(1))] + [(f"ds{t}", sg[t].grad, sc[t].grad) for t in range(3)] + [(f"da{t}", ag[t].grad, ac[t].grad) for t in range(3)])
param_report("R", "dynamics_network")

# (d) chain: E -> R -> D -> E -> R -> D
x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1)
wf = rnd(2, 3, 64, 64)
def chain_ref():
    dyn = O.Dynamics(st); s, _ = O.encoder(st, x); tot = 0
    for t in range(3):
        hh = dyn.step(s, acts[t], vars_[t]); o = O.decoder(st, hh)[0]; tot = tot + (o * wf).sum(); s, _ = O.encoder(st, o)
    return tot
def chain_gpu():
    model.dynamics_network.reinit_memory(2); s, _ = model.representation_network(x.cuda()); tot = 0
    for t in range(3):
        hh = model.dynamics_network(s, acts[t].cuda(), vars_[t].cuda()); o, _ = model.rendering_network(hh); tot = tot + (o * wf.cuda()).sum(); s, _ = model.representation_network(o)
    return tot
lr_ = chain_ref(); lr_.backward(); lg = chain_gpu(); lg.backward()
report("chain", [("loss", lg.detach().reshape(1), lr_.detach().reshape(1))])
param_report("chain", "")
