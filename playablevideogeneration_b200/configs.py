"""Plain-dict equivalents of the reference YAML configs (configs/01_bair.yaml, 02_breakout.yaml, 03_tennis.yaml) after
``Configuration.check_config`` has injected its defaults (utils/configuration.py:37-92), with the spatial sizes taken from the
caller.  What ``train.py`` / ``play.py`` would hand to ``model(config)`` (train.py:38-39) - used by bench.py, the tools and the
tests to build the model and the training step without a YAML file on disk.

    build_config(dict(config="bair" | "breakout" | "tennis", H=..., W=..., S=observation_stacking[, B=batch_size]))
"""
from __future__ import annotations

import copy

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
