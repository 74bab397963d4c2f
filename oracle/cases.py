"""Parity-case definitions shared by the golden generator, the oracle tests and the GPU parity tests.

TEST INFRASTRUCTURE.  A case names a config (BASELINE.json ``configs``, shrunk so the CPU reference finishes in
seconds), the seeds of weights / inputs / noise, and the schedule values the trainer would have produced.
``build_config`` returns the plain-dict equivalent of the reference YAML (configs/01_bair.yaml, 02_breakout.yaml,
03_tennis.yaml) after ``Configuration.check_config`` defaults (utils/configuration.py:37-92), with the spatial
sizes overridden by the case.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

STRIDE = 7   # large tensors are stored as flatten()[::STRIDE]

RESULT_NAMES_FULL = ["reconstructed_observations", "multiresolution_reconstructed_observations", "reconstructed_states",
                     "states", "hidden_states", "selected_actions", "action_logits", "action_samples", "attention",
                     "reconstructed_attention", "action_directions_distribution", "sampled_action_directions",
                     "action_states_distribution", "sampled_action_states", "action_variations",
                     "reconstructed_action_logits", "reconstructed_action_directions_distribution",
                     "reconstructed_sampled_action_directions", "reconstructed_action_states_distribution",
                     "reconstructed_sampled_action_states"]
RESULT_NAMES_PRE = ["reconstructed_observations", "multiresolution_reconstructed_observations", "reconstructed_states",
                    "states", "reconstructed_hidden_states", "hidden_states", "selected_actions", "action_logits",
                    "action_samples", "attention", "action_directions_distribution", "sampled_action_directions",
                    "action_states_distribution", "sampled_action_states", "action_variations",
                    "reconstructed_action_logits", "reconstructed_action_directions_distribution",
                    "reconstructed_sampled_action_directions", "reconstructed_action_states_distribution",
                    "reconstructed_sampled_action_states"]

_LOSS_WEIGHTS_BAIR = {
    "reconstruction_loss_lambda": 1.0, "reconstruction_loss_lambda_pretraining": 1.0,
    "perceptual_loss_lambda": 1.0, "perceptual_loss_lambda_pretraining": 1.0,
    "action_divergence_lambda": 0.0, "action_divergence_lambda_pretraining": 0.0,
    "states_rec_lambda": 0.2, "states_rec_lambda_pretraining": 0.2,
    "hidden_states_rec_lambda_pretraining": 1.0,
    "entropy_lambda": 0.0, "entropy_lambda_pretraining": 0.0,
    "action_directions_kl_lambda": 0.0001, "action_directions_kl_lambda_pretraining": 0.0001,
    "action_mutual_information_lambda": 0.15, "action_mutual_information_lambda_pretraining": 0.15,
    "action_state_distribution_kl_lambda": 0.0, "action_state_distribution_kl_lambda_pretraining": 0.0,
}

_BASE = {
    "logging": {"run_name": "case", "output_root": "/tmp/pvg_results", "save_root": "/tmp/pvg_checkpoints",
                "output_images_directory": "/tmp/pvg_results/images", "save_root_directory": "/tmp/pvg_checkpoints/case"},
    "data": {"data_root": "/tmp", "crop": None, "actions_count": 7, "ground_truth_available": False},
    "model": {
        "architecture": "model.main_model.model",
        "representation_network": {"target_input_size": [256, 256], "state_features": 64, "state_resolution": [32, 32]},
        "dynamics_network": {"hidden_state_size": 128, "embedding_mlp_size": 128, "random_noise_size": 32},
        "rendering_network": {"input_shape": [64, 32, 32]},
        "action_network": {"use_gumbel": True, "hard_gumbel": False, "ensamble_size": 1, "gumbel_temperature": 1.0,
                           "action_space_dimension": 2, "use_variations": True},
        "centroid_estimator": {"alpha": 0.1},
    },
    "training": {
        "trainer": "training.smooth_mi_trainer", "use_ground_truth_actions": False, "learning_rate": 0.0004,
        "weight_decay": 0.000001, "pretraining_steps": 1000, "pretraining_detach": False,
        "lr_schedule": [300000, 10000000000], "lr_gamma": 0.3333, "max_steps": 300000, "save_freq": 3000,
        "ground_truth_observations_start": 6, "ground_truth_observations_end": 6, "ground_truth_observations_steps": 16000,
        "gumbel_temperature_start": 1.0, "gumbel_temperature_end": 0.4, "gumbel_temperature_steps": 20000,
        "mutual_information_estimation_alpha": 0.2,
        "batching": {"batch_size": 8, "observations_count": 12, "observations_count_start": 7,
                     "observations_count_steps": 25000, "skip_frames": 0, "observation_stacking": 1, "num_workers": 0},
        "loss_weights": _LOSS_WEIGHTS_BAIR,
        "action_direction_plotting_freq": 1000, "use_motion_weights": False, "motion_weights_bias": 0.0,
        "action_mutual_information_entropy_lambda": 1.0, "max_steps_per_epoch": 10000,
    },
    "evaluation": {"evaluator": "evaluation.evaluator", "max_evaluation_batches": 20, "eval_freq": 8000,
                   "batching": {"batch_size": 8, "observations_count": 30, "skip_frames": 0, "observation_stacking": 1,
                                "num_workers": 0}},
}


def build_config(case: dict) -> dict:
    cfg = copy.deepcopy(_BASE)
    kind = case.get("config", "bair")
    H, W, S = case["H"], case["W"], case["S"]
    cfg["model"]["representation_network"]["target_input_size"] = [W, H]
    cfg["model"]["representation_network"]["state_resolution"] = [H // 8, W // 8]
    cfg["training"]["batching"]["observation_stacking"] = S
    cfg["evaluation"]["batching"]["observation_stacking"] = S
    cfg["training"]["batching"]["batch_size"] = case.get("B", 8)
    if kind == "bair":
        pass
    elif kind == "breakout":        # configs/02_breakout.yaml: reduced model, A=3, D=1, hidden 64
        cfg["model"]["architecture"] = "model.reduced_model.model"
        cfg["data"]["actions_count"] = 3
        cfg["model"]["dynamics_network"]["hidden_state_size"] = 64
        cfg["model"]["action_network"]["action_space_dimension"] = 1
    elif kind == "tennis":          # configs/03_tennis.yaml: S=4, D=5, plain trainer, KL-state lambda 1e-5
        cfg["model"]["action_network"]["action_space_dimension"] = 5
        cfg["training"]["trainer"] = "training.trainer"
        lw = cfg["training"]["loss_weights"]
        lw["action_state_distribution_kl_lambda"] = 0.00001
        lw["action_state_distribution_kl_lambda_pretraining"] = 0.00001
    else:
        raise ValueError(kind)
    return cfg


def sample_tensor(t, stride: int | None = None) -> np.ndarray:
    """Golden storage form: small tensors whole, large ones as a strided sample (float32 / int64)."""
    t = t.detach()
    if t.dtype in (torch.int64, torch.int32, torch.bool):
        return t.to(torch.int64).numpy()
    flat = t.float().contiguous().reshape(-1)
    if stride is None:
        stride = STRIDE if flat.numel() > 4096 else 1
    return flat[::stride].numpy().copy()


# BASELINE.json configs, shrunk for the CPU reference.  gt_init is what the trainer would pass (capped to T-1).
CASES = {
    # configs[0]: configs/01_bair.yaml, batch=2, seq_len=4, 64x64, full model + all losses (one train step)
    "full_bair": dict(mode="full", config="bair", B=2, T=4, S=1, H=64, W=64, gt_init=3, gumbel_temperature=1.0,
                      weight_seed=0, input_seed=0, noise_seed=123, smooth_mi=True),
    # same, free-running from frame 1 (exercises the D -> E feedback path for T-2 steps) and two optimiser steps
    "full_bair_feedback": dict(mode="full", config="bair", B=2, T=5, S=1, H=64, W=64, gt_init=1, gumbel_temperature=0.7,
                               weight_seed=1, input_seed=1, noise_seed=321, smooth_mi=True, steps=2),
    "pretrain_bair": dict(mode="pretraining", config="bair", B=2, T=4, S=1, H=64, W=64, gt_init=3,
                          gumbel_temperature=1.0, weight_seed=2, input_seed=2, noise_seed=77, smooth_mi=True),
    # configs[2] shape family: reduced model, non-square frames (Breakout is 208x160 -> here 96x64), A=3, D=1
    "full_breakout": dict(mode="full", config="breakout", reduced=True, B=2, T=4, S=1, H=96, W=64, gt_init=2,
                          gumbel_temperature=0.9, weight_seed=3, input_seed=3, noise_seed=11, smooth_mi=True),
    # configs[3] shape family: Tennis, observation stacking 4 (12 input channels), D=5, plain MI
    "full_tennis": dict(mode="full", config="tennis", B=2, T=5, S=4, H=64, W=128, gt_init=2, gumbel_temperature=0.6,
                        weight_seed=4, input_seed=4, noise_seed=5, smooth_mi=False),
    # configs[4] family: play.py rollout, eval mode, batch 1
    "rollout_bair": dict(mode="rollout", config="bair", S=1, H=64, W=64, weight_seed=5, input_seed=5, noise_seed=9,
                         actions=[0, 3, 6, 1], noise=False),
    "rollout_tennis_noise": dict(mode="rollout", config="tennis", S=4, H=32, W=64, weight_seed=6, input_seed=6,
                                 noise_seed=10, actions=[2, 2, 5], noise=True),
    # build_evaluation_dataset.py: eval-mode forward with the evaluation sampler plug-ins
    "eval_bair_builder": dict(mode="eval", config="bair", B=2, T=5, S=1, H=64, W=64, gt_init=2, gumbel_temperature=0.7,
                              weight_seed=9, input_seed=9, noise_seed=13),
    # interpolate.py: generate_next_interpolation(observation, first_action, second_action, factor) after one plain step
    "rollout_bair_interp": dict(mode="rollout", config="bair", S=1, H=64, W=64, weight_seed=7, input_seed=8, noise_seed=12,
                                actions=[4], noise=False, interp=[[0, 3, 0.25], [1, 5, 0.8], [6, 2, 0.5]]),
}
