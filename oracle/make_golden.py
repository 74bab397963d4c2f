"""Generates tests/golden/*.npz by running the UNMODIFIED reference (this container only).

    python oracle/make_golden.py            # writes every case
    python oracle/make_golden.py full_bair  # one case

Each case = one seeded run of the reference ``Model`` + ``Trainer.compute_losses[_pretraining]`` + backward (or the
``generate_next`` rollout) on the deterministic weights/inputs of ``oracle/caddy_oracle.py``.  Large tensors are
stored as a strided sample (``flatten()[::STRIDE]``) plus their mean/abs-mean so fixtures stay small.
The case definitions (``CASES``) are shared with the tests, which rebuild inputs and weights from the same seeds.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import caddy_oracle as O          # noqa: E402
from oracle import ref_harness as R           # noqa: E402
from oracle.cases import CASES, STRIDE, build_config, sample_tensor, RESULT_NAMES_FULL, RESULT_NAMES_PRE  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _record_results(out: dict, names, results):
    for name, val in zip(names, results):
        if isinstance(val, (list, tuple)):
            for i, v in enumerate(val):
                out[f"res.{name}.{i}"] = sample_tensor(v)
        else:
            out[f"res.{name}"] = sample_tensor(val)


def run_train_case(case: dict) -> dict:
    cfg = build_config(case)
    reduced = case.get("reduced", False)
    vgg_sd = O.make_vgg_weights(case.get("vgg_seed", 1234))
    R.install_shims(vgg_sd)
    import torchvision.models as tvm
    tvm.vgg19.weights = vgg_sd
    sd = O.make_weights(cfg, case["weight_seed"], reduced)
    model = R.build_model(cfg, sd, reduced)
    model.train()
    dp = torch.nn.DataParallel(model)
    trainer = R.build_trainer(cfg, dp, smooth=case.get("smooth_mi", True))
    obs = O.make_observations(case["B"], case["T"], 3 * case["S"], case["H"], case["W"], case["input_seed"])
    batch = R.make_batch(obs)
    captured = {}
    model.register_forward_hook(lambda m, i, o: captured.__setitem__("res", o))
    # the trainer reads its schedule; pin the quantities the case names
    trainer.get_ground_truth_observations_count = lambda: case["gt_init"]
    trainer.get_gumbel_temperature = lambda: case["gumbel_temperature"]
    out = {}
    n_steps = case.get("steps", 1)
    for step in range(n_steps):
        torch.manual_seed(case["noise_seed"] + step)
        random.seed(case["noise_seed"] + step)
        if case["mode"] == "pretraining":
            total, info, _ = trainer.compute_losses_pretraining(dp, batch, case["T"])
        else:
            total, info, _ = trainer.compute_losses(dp, batch, case["T"])
        trainer.optimizer.zero_grad()
        total.backward()
        tag = "" if step == 0 else f"step{step}."
        out[tag + "total_loss"] = total.detach().double().numpy()
        for k, v in info.items():
            if isinstance(v, (int, float)):
                out[tag + "info." + k] = np.float64(v)
        if step == 0:
            names = RESULT_NAMES_PRE if case["mode"] == "pretraining" else RESULT_NAMES_FULL
            _record_results(out, names, captured["res"])
            for k, p in model.named_parameters():
                if p.grad is not None:
                    out["gradnorm." + k] = p.grad.double().norm().numpy()
                    out["gradsample." + k] = sample_tensor(p.grad, stride=max(1, p.numel() // 64))
            msd = model.state_dict()
            for k in msd:
                if k.endswith("running_mean") or k.endswith("running_var") or "centroid" in k:
                    out["buf." + k] = msd[k].detach().float().numpy().copy()
            if hasattr(trainer.mutual_information_loss, "matrix_estimator"):
                out["mi_matrix"] = trainer.mutual_information_loss.matrix_estimator.estimated_matrix.detach().numpy().copy()
        if n_steps > 1:
            trainer.optimizer.step()
            if step == n_steps - 1:
                for k, p in model.named_parameters():
                    out["param_after." + k] = sample_tensor(p.detach(), stride=max(1, p.numel() // 64))
    return out


def run_eval_case(case: dict) -> dict:
    """evaluation/evaluation_dataset_builder.py:47-54: model(batch, ground_truth_observations_init, OneHotActionSampler,
    ZeroActionVariationSampler, gumbel_temperature) in eval mode under no_grad."""
    cfg = build_config(case)
    R.install_shims()
    sys.path.insert(0, R.REF_ROOT)
    from evaluation.action_sampler import OneHotActionSampler
    from evaluation.action_variation_sampler import ZeroActionVariationSampler
    sd = O.make_weights(cfg, case["weight_seed"], False)
    model = R.build_model(cfg, sd, False)
    model.eval()
    obs = O.make_observations(case["B"], case["T"], 3 * case["S"], case["H"], case["W"], case["input_seed"])
    batch = R.make_batch(obs)
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    with torch.no_grad():
        res = model(batch.to_tuple(), ground_truth_observations_init=case["gt_init"], action_sampler=OneHotActionSampler(),
                    action_variation_sampler=ZeroActionVariationSampler(), gumbel_temperature=case["gumbel_temperature"])
    out = {}
    _record_results(out, RESULT_NAMES_FULL, res)
    return out


def run_rollout_case(case: dict) -> dict:
    cfg = build_config(case)
    reduced = case.get("reduced", False)
    R.install_shims()
    sd = O.make_weights(cfg, case["weight_seed"], reduced)
    model = R.build_model(cfg, sd, reduced)
    model.eval()
    obs = O.make_observations(1, 1, 3 * case["S"], case["H"], case["W"], case["input_seed"])[0, 0]
    out = {}
    torch.manual_seed(case["noise_seed"])
    with torch.no_grad():
        model.start_inference()
        for i, a in enumerate(case["actions"]):
            frame, obs = model.generate_next(obs, a, noise=case.get("noise", False))
            out[f"frame.{i}"] = frame.numpy()
        for i, (a1, a2, f) in enumerate(case.get("interp", [])):
            frame, obs = model.generate_next_interpolation(obs, a1, a2, f)
            out[f"iframe.{i}"] = frame.numpy()
    return out


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    want = set(argv[1:])
    for name, case in CASES.items():
        if want and name not in want:
            continue
        torch.set_num_threads(os.cpu_count())
        out = (run_rollout_case(case) if case["mode"] == "rollout" else run_eval_case(case) if case["mode"] == "eval"
               else run_train_case(case))
        out["case_json"] = np.array(json.dumps(case))
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB, "
              f"total_loss={out.get('total_loss', 'n/a')}")


if __name__ == "__main__":
    main(sys.argv)
