"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests: load a golden case, rebuild its seeded inputs."""
import json
import os

import numpy as np
import torch

from oracle import caddy_oracle as O
from oracle.cases import CASES, build_config, sample_tensor, RESULT_NAMES_FULL, RESULT_NAMES_PRE

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = json.loads(str(g["case_json"]))
    assert case == json.loads(json.dumps(CASES[name])), "golden fixture is stale: re-run oracle/make_golden.py"
    return case, g


def case_inputs(case):
    cfg = build_config(case)
    reduced = case.get("reduced", False)
    sd = O.make_weights(cfg, case["weight_seed"], reduced)
    vgg_sd = O.make_vgg_weights(case.get("vgg_seed", 1234))
    if case["mode"] == "rollout":
        obs = O.make_observations(1, 1, 3 * case["S"], case["H"], case["W"], case["input_seed"])[0, 0]
    else:
        obs = O.make_observations(case["B"], case["T"], 3 * case["S"], case["H"], case["W"], case["input_seed"])
    return cfg, sd, vgg_sd, obs


def batch_tuple(obs):
    B, T = obs.shape[:2]
    return (obs, torch.zeros((B, T), dtype=torch.int32), torch.zeros((B, T)), torch.zeros((B, T), dtype=torch.bool))


def compare_results(g, names, results, rtol, atol, prefix="res."):
    """Checks a 20-tuple against the stored (strided) reference tensors; returns the worst abs error seen."""
    worst = 0.0
    for name, val in zip(names, results):
        vals = list(val) if isinstance(val, (list, tuple)) else [val]
        keys = [f"{prefix}{name}.{i}" for i in range(len(vals))] if isinstance(val, (list, tuple)) else [prefix + name]
        for k, v in zip(keys, vals):
            ref = g[k]
            got = sample_tensor(v.detach().cpu())
            assert got.shape == ref.shape, (k, got.shape, ref.shape)
            if ref.dtype == np.int64:
                assert (got == ref).all(), k
                continue
            err = np.abs(got - ref)
            tol = atol + rtol * np.abs(ref)
            assert (err <= tol).all(), f"{k}: max err {err.max():.3e} (tol {tol.min():.1e}), ref absmax {np.abs(ref).max():.3e}"
            worst = max(worst, float(err.max()))
    return worst


def oracle_run(case, dtype=torch.float32, threads=16):
    """One forward + losses + backward of the CPU oracle in ``dtype`` on the case's seeded inputs/weights/noise.
    float64 gives the ground truth that both fp32 implementations (the reference's CPU kernels and the CUDA path) are
    measured against: the feedback recurrence with train-mode BatchNorm amplifies rounding differences by ~1e3 per few
    steps, so "how far is fp32 from exact" is the only meaningful yardstick for intermediate tensors and gradients."""
    import random
    torch.set_num_threads(min(threads, torch.get_num_threads()))
    cfg, sd, vgg_sd, obs = case_inputs(case)
    conv = lambda v: v.to(dtype) if v.is_floating_point() else v
    params = {k: conv(v).clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
                                                and "centroid" not in k) for k, v in sd.items()}
    vgg = {k: conv(v) for k, v in vgg_sd.items()}
    mi = O.MutualInformation(cfg["data"]["actions_count"],
                             cfg["training"]["mutual_information_estimation_alpha"] if case["smooth_mi"] else None)
    mi.matrix = mi.matrix.to(dtype)
    bt = batch_tuple(obs)
    bt = (bt[0].to(dtype),) + bt[1:]
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    total, comp, res = O.compute_losses(params, vgg, cfg, mi, bt, case["gt_init"], case["gumbel_temperature"],
                                        pretraining=case["mode"] == "pretraining")
    total.backward()
    grads = {k: v.grad.detach() for k, v in params.items() if v.grad is not None}
    return total.detach(), res, grads


def flat_results(res):
    out = []
    for r in res:
        out.extend(list(r) if isinstance(r, (list, tuple)) else [r])
    return out


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))
