"""Pins the CPU oracle (oracle/caddy_oracle.py) against outputs of the UNMODIFIED reference stored in tests/golden/.

The oracle and the reference both run PyTorch fp32 CPU kernels, so agreement is expected to ~1e-6 (the only
differences are op ordering, e.g. expand-vs-repeat).  Tolerances: tensors 2e-5 abs + 2e-5 rel; total loss 1e-6 rel.
"""
import random

import numpy as np
import pytest
import torch

from oracle import caddy_oracle as O
from oracle.cases import CASES, RESULT_NAMES_FULL, RESULT_NAMES_PRE
from tests.golden_util import load_case, case_inputs, batch_tuple, compare_results

TRAIN_CASES = [n for n, c in CASES.items() if c["mode"] in ("full", "pretraining")]
ROLLOUT_CASES = [n for n, c in CASES.items() if c["mode"] == "rollout"]


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_oracle_train_step_matches_reference(name):
    case, g = load_case(name)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
                                          and "centroid" not in k) for k, v in sd.items()}
    mi = O.MutualInformation(cfg["data"]["actions_count"],
                             cfg["training"]["mutual_information_estimation_alpha"] if case["smooth_mi"] else None)
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    total, comp, res = O.compute_losses(params, vgg_sd, cfg, mi, batch_tuple(obs), case["gt_init"],
                                        case["gumbel_temperature"], pretraining=case["mode"] == "pretraining")
    names = RESULT_NAMES_PRE if case["mode"] == "pretraining" else RESULT_NAMES_FULL
    compare_results(g, names, res, rtol=2e-5, atol=2e-5)
    ref_total = float(g["total_loss"][0])
    assert abs(float(total) - ref_total) <= 1e-6 * abs(ref_total), (float(total), ref_total)
    for r in range(3):
        for k in (f"perceptual_loss_r{r}", f"observations_rec_loss_r{r}"):
            assert abs(float(comp[k]) - float(g["info." + k])) <= 2e-6 * abs(float(g["info." + k])) + 1e-9, k
    for k in ("states_rec_loss", "entropy_loss", "action_directions_kl_loss", "action_mutual_information_loss",
              "action_state_distribution_kl_loss"):
        ref = float(g["info." + k])
        assert abs(float(comp[k]) - ref) <= 1e-5 * abs(ref) + 1e-8, (k, float(comp[k]), ref)
    total.backward()
    for k, p in params.items():
        key = "gradnorm." + k
        if key in g.files:
            ref = float(g[key])
            got = float(p.grad.double().norm()) if p.grad is not None else 0.0
            assert abs(got - ref) <= 1e-3 * ref + 1e-7, (k, got, ref)
    for k in g.files:
        if k.startswith("buf."):
            np.testing.assert_allclose(params[k[4:]].detach().numpy(), g[k], rtol=1e-4, atol=1e-5, err_msg=k)
    if case["smooth_mi"]:
        np.testing.assert_allclose(mi.matrix.numpy(), g["mi_matrix"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_oracle_rollout_matches_reference(name):
    case, g = load_case(name)
    cfg, sd, _, obs = case_inputs(case)
    roll = O.Rollout({k: v.clone() for k, v in sd.items()}, cfg)
    torch.manual_seed(case["noise_seed"])
    with torch.no_grad():
        roll.start_inference()
        for i, a in enumerate(case["actions"]):
            frame, obs = roll.generate_next(obs, a, noise=case.get("noise", False))
            np.testing.assert_allclose(frame.numpy(), g[f"frame.{i}"], rtol=2e-5, atol=2e-5)
        for i, (a1, a2, f) in enumerate(case.get("interp", [])):
            frame, obs = roll.generate_next_interpolation(obs, a1, a2, f)
            np.testing.assert_allclose(frame.numpy(), g[f"iframe.{i}"], rtol=2e-5, atol=2e-5)


def test_weight_recipe_covers_reference_state_dict():
    """model_param_spec must list exactly the reference's checkpoint keys - 9 856 367 params for BAIR."""
    from oracle.cases import build_config
    cfg = build_config(dict(config="bair", H=256, W=256, S=1))
    spec = O.model_param_spec(cfg)
    n = sum(int(np.prod(s)) for name, s, kind in spec if kind not in ("rmean", "rvar", "nbt"))
    assert n == 9856367       # SURVEY.md 8b (includes the (7,2) requires_grad=False centroid Parameter)


EVAL_CASES = [n for n, c in CASES.items() if c["mode"] == "eval"]


@pytest.mark.parametrize("name", EVAL_CASES)
def test_oracle_eval_forward_with_samplers_matches_reference(name):
    """build_evaluation_dataset.py path: eval-mode forward_full_model with OneHotActionSampler / ZeroActionVariationSampler."""
    from playablevideogeneration_b200.evaluation.samplers import OneHotActionSampler, ZeroActionVariationSampler
    case, g = load_case(name)
    cfg, sd, _, obs = case_inputs(case)
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    with torch.no_grad():
        res = O.forward_full_model({k: v.clone() for k, v in sd.items()}, cfg, batch_tuple(obs), case["gt_init"],
                                   case["gumbel_temperature"], OneHotActionSampler(), ZeroActionVariationSampler(), train=False)
    compare_results(g, RESULT_NAMES_FULL, res, rtol=2e-5, atol=2e-5)
