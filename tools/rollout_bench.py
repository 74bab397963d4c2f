"""BASELINE.json configs[4]: play.py-style autoregressive rollout, BAIR 256x256, batch 64, 100 steps, eval mode (secondary
metric: generated frames/s).  usage: python tools/rollout_bench.py [batch] [steps] [precision] [graph]
``graph`` replays one CUDA graph per generated frame (Model.enable_graphed_inference); batch 1 is play.py's own case."""
import os, sys, json, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200.configs import build_config
from playablevideogeneration_b200 import ops
from playablevideogeneration_b200.caddy import Model
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ops.set_precision(sys.argv[3] if len(sys.argv) > 3 else "tf32x3")
dev = torch.device("cuda")
cfg = build_config(dict(config="bair", H=256, W=256, S=1))
torch.manual_seed(0); random.seed(0)
model = Model(cfg).to(dev).eval()
graphed = len(sys.argv) > 4 and sys.argv[4] == "graph"
if graphed:
    model.enable_graphed_inference()
g = torch.Generator().manual_seed(0)
obs = (torch.rand((batch, 3, 256, 256), generator=g) * 2 - 1).to(dev)
actions = torch.randint(0, 7, (steps, batch), generator=g).to(dev)
with torch.no_grad():
    model.start_inference()
    model.dynamics_network.reinit_memory(batch)
    for t in range(3):                                        # warm-up (captures the graph in graphed mode)
        frames, obs = model.generate_next_batch(obs, actions[t])
    model.start_inference()
    model.dynamics_network.reinit_memory(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        frames, obs = model.generate_next_batch(obs, actions[t])
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps(dict(metric="rollout frames/sec (BAIR 256x256, eval, batch %d, %d steps)" % (batch, steps), value=batch * steps / (ms * 1e-3),
                      ms_per_step=ms / steps, precision=ops.get_precision(), cuda_graph=graphed, finite=bool(torch.isfinite(frames).all()))))
