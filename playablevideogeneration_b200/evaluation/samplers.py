"""Sampler plug-ins of the evaluation path (SURVEY.md 8b "Sampler plug-ins", 8f rank 2) and the frame conversion of the
evaluation-dataset builder.  Same class names and call signatures as ``evaluation/action_sampler.py`` and
``evaluation/action_variation_sampler.py``; tensors stay on the device of their inputs (the reference hard-codes .cuda())."""
from __future__ import annotations

from typing import Dict

import torch


class OneHotActionSampler:
    """evaluation/action_sampler.py:6-34: one-hot of the most probable action."""

    def __call__(self, log_probabilities: torch.Tensor, ground_truth: torch.Tensor) -> torch.Tensor:
        onehot = torch.zeros_like(log_probabilities, dtype=torch.float)
        return onehot.scatter_(1, log_probabilities.argmax(dim=1, keepdim=True), 1)


class GroundTruthActionSampler:
    """evaluation/action_sampler.py:37-84: one-hot of the ground-truth action translated into the model's action space."""

    def __init__(self, ground_truth_to_actions_mapping: Dict):
        self.mapping_dict = ground_truth_to_actions_mapping

    def translate_ground_truth_indexes(self, ground_truth: torch.Tensor) -> torch.Tensor:
        out = ground_truth.clone()
        for gt_idx, idx in self.mapping_dict.items():
            out[ground_truth == gt_idx] = idx
        return out

    def __call__(self, log_probabilities: torch.Tensor, ground_truth: torch.Tensor) -> torch.Tensor:
        onehot = torch.zeros_like(log_probabilities, dtype=torch.float)
        idx = self.translate_ground_truth_indexes(ground_truth).reshape((-1, 1)).long().to(log_probabilities.device)
        return onehot.scatter_(1, idx, 1)


class ZeroActionVariationSampler:
    """evaluation/action_variation_sampler.py:6-25."""

    def __call__(self, sampled_action_directions: torch.Tensor, action_samples: torch.Tensor) -> torch.Tensor:
        return sampled_action_directions * 0


def normalize_range(observations: torch.Tensor) -> torch.Tensor:
    """evaluation_dataset_builder.py:142-154 (check_and_normalize_range): [-1, 1] -> [0, 1] when any value is negative - decided
    on the device (the reference reads the minimum back with .item())."""
    return torch.where(observations.min() < 0, (observations + 1) / 2, observations)


def frames_to_uint8_hwc(observations: torch.Tensor) -> torch.Tensor:
    """(..., C, H, W) float frames -> (..., H, W, C) uint8 on the device, the conversion the reference does on the host
    (np.moveaxis + (x * 255).astype(np.uint8), evaluation_dataset_builder.py:66-68 / utils/tensor_displayer.py:25-27):
    a quarter of the bytes cross PCIe."""
    from .. import ops
    x = observations
    if x.is_cuda and x.dtype == torch.float32 and x.dim() >= 4:
        lead = x.shape[:-3]
        flat = x.reshape((-1,) + tuple(x.shape[-3:]))
        if flat.stride()[1:] != (1, flat.shape[3] * flat.shape[1], flat.shape[1]) and flat.shape[1] != 1:
            flat = flat.contiguous(memory_format=torch.channels_last)          # frames from the model already are channels-last
        u8 = ops.frames_to_uint8(flat)                                          # one min pass + one conversion pass on the device
        return u8.permute(0, 2, 3, 1).reshape(tuple(lead) + (flat.shape[2], flat.shape[3], flat.shape[1]))
    x = normalize_range(observations)
    return (x * 255).clamp(0, 255).to(torch.uint8).movedim(-3, -1).contiguous()
