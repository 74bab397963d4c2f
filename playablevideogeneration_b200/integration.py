"""One call that makes the UNMODIFIED reference scripts (train.py, play.py, interpolate.py, build_evaluation_dataset.py) run
on this package - no reference file is edited.

    import playablevideogeneration_b200.integration as pvg
    pvg.install(vgg_features_state_dict=torchvision.models.vgg19(weights=...).state_dict())      # before `import train`

``install`` registers modules in ``sys.modules`` under the dotted names the reference looks up:

* ``model.main_model.model`` / ``model.reduced_model.model``  (``config["model"]["architecture"]``, train.py:38-39,
  play.py:45-46) -> the factories of this package;
* ``training.losses`` (imported by training/trainer.py:17-19 and training/smooth_mi_trainer.py:7) -> a module with the
  reference's class names backed by the pvg_b200 kernels, so ``Trainer.__init__`` (trainer.py:41-54) builds the tensor-core
  VGG19 perceptual loss instead of the cuDNN one (74 % of the training FLOPs).

Everything else the trainer does (Adam, schedules, logging, checkpoints) stays the reference's own code.
"""
from __future__ import annotations

import sys
import types
from typing import Dict, Optional

import torch


def losses_module(vgg_features_state_dict: Optional[Dict[str, torch.Tensor]] = None, allow_random_vgg: bool = False) -> types.ModuleType:
    """A stand-in for the reference's ``training/losses.py``: same names, same constructor and call signatures."""
    from .training import losses as L
    from .vgg import Vgg19

    shared = {}

    def vgg():
        if "vgg" not in shared:        # one frozen VGG19 for every loss object, like the reference's module-level usage
            shared["vgg"] = Vgg19(vgg_features_state_dict, allow_random_init=allow_random_vgg)
        return shared["vgg"]

    class ParallelPerceptualLoss(L.ParallelPerceptualLoss):
        """losses.py:379-390: constructed without arguments by the trainer (trainer.py:51)."""

        def __init__(self):
            super().__init__(vgg())

    class UnmeanedPerceptualLoss(L.UnmeanedPerceptualLoss):
        def __init__(self):
            super().__init__(vgg())

    class PerceptualLoss:
        """losses.py:494-588.  The trainer constructs one (trainer.py:43) and overwrites it eight lines later (:51) without
        ever calling it; constructing it here is free (the reference loads a second VGG19 for nothing)."""

        def __call__(self, *args, **kwargs):
            raise NotImplementedError("PerceptualLoss is never called by the reference trainer; use ParallelPerceptualLoss")

    mod = types.ModuleType("training.losses")
    mod.__doc__ = "pvg_b200 stand-in for the reference's training/losses.py (playablevideogeneration_b200.integration)"
    for name in ("StatesLoss", "HiddenStatesLoss", "ObservationsLoss", "KLDivergence", "KLGaussianDivergenceLoss",
                 "KLGeneralGaussianDivergenceLoss", "FixedMatrixEstimator", "MutualInformationLoss", "SmoothMutualInformationLoss",
                 "EntropyLogitLoss", "EntropyProbabilityLoss", "MotionLossWeightMaskCalculator", "SequenceLossEvaluator"):
        setattr(mod, name, getattr(L, name))
    mod.ParallelPerceptualLoss = ParallelPerceptualLoss
    mod.UnmeanedPerceptualLoss = UnmeanedPerceptualLoss
    mod.PerceptualLoss = PerceptualLoss
    return mod


def install(vgg_features_state_dict: Optional[Dict[str, torch.Tensor]] = None, models: bool = True, losses: bool = True,
            allow_random_vgg: bool = False) -> None:
    """Registers the stand-ins (see the module docstring).  Call before the reference modules are imported."""
    if models:
        from .model.main_model import model as main_model
        from .model.reduced_model import model as reduced_model
        sys.modules["model.main_model.model"] = main_model
        sys.modules["model.reduced_model.model"] = reduced_model
    if losses:
        sys.modules["training.losses"] = losses_module(vgg_features_state_dict, allow_random_vgg)
