"""The drop-in boundary exercised by the REAL caller: the unmodified reference ``Trainer`` / ``SmoothMITrainer``
(training/trainer.py:266,426) drives ``nn.DataParallel(<this package's Model>)`` and must reproduce the golden losses of the
all-reference run - first with the reference's own loss classes, then with ``training.losses`` replaced by this package's
stand-in (playablevideogeneration_b200.integration) without editing a reference file.

CPU test: the kernels are replaced by the test-only stand-ins of tests/fake_ops.py, so what is checked is the protocol - the
factory, ``.module.centroid_estimator``, the 20-tuple, the loss classes' names / signatures / return conventions, the quirks
the trainer relies on.  Needs /root/reference (this container only; skipped on the GPU box)."""
import importlib
import random
import sys

import numpy as np
import pytest
import torch

from oracle import ref_harness as R
from tests import fake_ops
from tests.golden_util import case_inputs, load_case

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference is not present (reference files do not travel)")
CASES = ["full_bair", "pretrain_bair", "full_tennis", "full_breakout"]


def _run(case, g, cfg, sd, obs, trainer_of):
    from playablevideogeneration_b200.caddy import Model
    model = Model(cfg, reduced=case.get("reduced", False))
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    model.train()
    dp = torch.nn.DataParallel(model)               # what train.py:67 hands to the trainer (no devices here: calls the module)
    trainer = trainer_of(cfg, dp, case.get("smooth_mi", True))
    trainer.get_ground_truth_observations_count = lambda: case["gt_init"]
    trainer.get_gumbel_temperature = lambda: case["gumbel_temperature"]
    batch = R.make_batch(obs)
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    if case["mode"] == "pretraining":
        total, info, _ = trainer.compute_losses_pretraining(dp, batch, case["T"])
    else:
        total, info, _ = trainer.compute_losses(dp, batch, case["T"])
    ref = float(g["total_loss"][0])
    assert abs(float(total) - ref) <= 2e-6 * abs(ref), (float(total), ref)
    for k, v in info.items():
        if isinstance(v, (int, float)) and ("info." + k) in g.files:
            r = float(g["info." + k])
            assert abs(v - r) <= 2e-5 * abs(r) + 1e-7, (k, v, r)
    trainer.optimizer.zero_grad()
    total.backward()
    for k, p in model.named_parameters():
        key = "gradnorm." + k
        if key in g.files and p.grad is not None:
            r = float(g[key])
            assert abs(float(p.grad.double().norm()) - r) <= 2e-3 * r + 1e-7, (k, float(p.grad.double().norm()), r)
    trainer.optimizer.step()                        # torch.optim.Adam over this model's parameters (trainer.py:36,586)
    return trainer


@pytest.mark.parametrize("name", CASES)
def test_reference_trainer_drives_this_model(name, monkeypatch):
    fake_ops.install(monkeypatch)
    case, g = load_case(name)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    R.install_shims(vgg_sd)
    import torchvision.models as tvm
    tvm.vgg19.weights = vgg_sd
    t = _run(case, g, cfg, sd, obs, R.build_trainer)
    assert type(t.perceptual_loss).__module__ == "training.losses"


@pytest.mark.parametrize("name", ["full_bair", "full_tennis"])
def test_reference_trainer_with_this_packages_losses(name, monkeypatch):
    """``integration.install`` swaps ``training.losses`` before training/trainer.py is imported: the reference file is used
    as it is, yet its ``Trainer.__init__`` builds this package's loss objects."""
    fake_ops.install(monkeypatch)
    case, g = load_case(name)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    R.install_shims(vgg_sd)
    import playablevideogeneration_b200.integration as integ
    for m in ("training.losses", "training.trainer", "training.smooth_mi_trainer"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    monkeypatch.setitem(sys.modules, "training.losses", integ.losses_module(vgg_sd))
    t = _run(case, g, cfg, sd, obs, R.build_trainer)
    assert type(t.perceptual_loss).__module__.startswith("playablevideogeneration_b200")
    assert type(t.observations_loss).__module__.startswith("playablevideogeneration_b200")
    for m in ("training.trainer", "training.smooth_mi_trainer"):       # leave no trainer bound to the stand-in behind
        monkeypatch.delitem(sys.modules, m, raising=False)


def test_install_registers_the_factories(monkeypatch):
    import playablevideogeneration_b200.integration as integ
    for m in ("model.main_model.model", "model.reduced_model.model", "training.losses"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    try:
        integ.install(allow_random_vgg=True)
        assert importlib.import_module("model.main_model.model").model.__module__.startswith("playablevideogeneration_b200")
        assert importlib.import_module("model.reduced_model.model").model.__module__.startswith("playablevideogeneration_b200")
        assert hasattr(importlib.import_module("training.losses"), "ParallelPerceptualLoss")
    finally:
        for m in ("model.main_model.model", "model.reduced_model.model", "training.losses"):
            sys.modules.pop(m, None)
