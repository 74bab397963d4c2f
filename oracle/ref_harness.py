"""Imports the UNMODIFIED reference from /root/reference (this container only) with the shims of SURVEY.md 8c.

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` to produce ``tests/golden/*`` and by
``tests/test_oracle_golden.py`` (when /root/reference exists) to re-check the oracle restatement live.  Nothing on
the GPU box needs it: /root/reference does not travel.

Shims (monkey patches, reference files untouched):
  1. stub matplotlib (imported by utils/tensor_displayer.py:11-12, which model.py:15 imports);
  2. ``Tensor.cuda`` / ``Module.cuda`` -> identity on a CPU-only host (hard-coded .cuda() calls);
  3. ``collections.Sequence`` (training/trainer.py:179; removed in py3.10);
  4. ``torchvision.models.vgg19(pretrained=True)`` -> architecture only, weights from ``make_vgg_weights``;
  5. WANDB disabled, fake dataset/logger for the Trainer constructor.
"""
from __future__ import annotations

import collections
import collections.abc
import copy
import os
import sys
import types

import torch
import torch.nn as nn
import yaml

REF_ROOT = os.environ.get("PVG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "main_model"))


_installed = False


def install_shims(vgg_sd=None):
    global _installed
    if _installed:
        return
    os.environ.setdefault("WANDB_MODE", "disabled")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches", "matplotlib.colors"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    import torchvision.models as tvm
    _orig_vgg19 = tvm.vgg19

    def _vgg19(pretrained=False, **kw):
        net = _orig_vgg19(weights=None)
        if _vgg19.weights is not None:
            own = net.state_dict()
            own.update({k: v.clone() for k, v in _vgg19.weights.items()})
            net.load_state_dict(own)
        return net
    _vgg19.weights = vgg_sd
    tvm.vgg19 = _vgg19
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def load_config(yaml_name: str, overrides: dict | None = None) -> dict:
    """Reads configs/<yaml_name> from the reference and applies the defaults of utils/configuration.py:37-92."""
    with open(os.path.join(REF_ROOT, "configs", yaml_name)) as f:
        cfg = yaml.safe_load(f)
    tr = cfg["training"]
    tr.setdefault("use_motion_weights", False)
    tr.setdefault("motion_weights_bias", 0.0)
    tr.setdefault("action_direction_plotting_freq", 1000)
    tr.setdefault("action_mutual_information_entropy_lambda", 1.0)
    tr.setdefault("max_steps_per_epoch", 10000)
    cfg["model"]["action_network"].setdefault("use_variations", True)
    cfg["data"].setdefault("ground_truth_available", True)
    cfg["logging"]["output_images_directory"] = "/tmp/pvg_ref_out"

    def merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(dst.get(k), dict):
                merge(dst[k], v)
            else:
                dst[k] = v
    if overrides:
        merge(cfg, overrides)
    return cfg


def build_model(cfg: dict, sd: dict, reduced: bool = False):
    """Reference ``model(config)`` factory (train.py:38-39) with the golden weights loaded."""
    install_shims()
    import importlib
    mod = importlib.import_module("model.reduced_model.model" if reduced else "model.main_model.model")
    m = getattr(mod, "model")(copy.deepcopy(cfg))
    missing = m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m


class _FakeBatch:
    def __init__(self, observations, actions):
        self.observations = observations
        self.actions = actions
        self.rewards = torch.zeros(actions.shape, dtype=torch.float32)
        self.dones = torch.zeros(actions.shape, dtype=torch.bool)

    def to_tuple(self, cuda=True):
        return self.observations, self.actions, self.rewards, self.dones


class _FakeLogger:
    def print(self, *a, **k):
        pass

    def get_wandb(self):
        return None


def build_trainer(cfg: dict, model, smooth: bool):
    """Reference Trainer / SmoothMITrainer over a fake dataset (only the loss code is exercised)."""
    install_shims()
    import importlib
    mod = importlib.import_module("training.smooth_mi_trainer" if smooth else "training.trainer")
    c = copy.deepcopy(cfg)
    c["training"]["batching"]["num_workers"] = 0
    t = getattr(mod, "trainer")(c, model, [0] * 64, _FakeLogger())
    t.global_step = 1        # avoid the plotting branch at step 0 (trainer.py:544)
    return t


def make_batch(observations, actions=None):
    if actions is None:
        actions = torch.zeros(observations.shape[:2], dtype=torch.int32)
    return _FakeBatch(observations, actions)
