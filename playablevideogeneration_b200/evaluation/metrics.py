"""Cheap per-frame evaluation metrics of the reference evaluator, on device.

Same class names, call signatures and return shapes ``(bs, observations_count)`` as ``evaluation/metrics/{mse,psnr,
motion_masked_mse,vgg_cosine_similarity}.py``; inputs are ``(bs, observations_count, channels, height, width)``.
The pixel metrics are single fused reductions over (C, H, W); ``VGGCosineSimilarity`` runs the VGG19 pyramid of the
perceptual loss (``playablevideogeneration_b200.vgg.Vgg19``: the tcgen05 conv kernels) once per branch."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..vgg import Vgg19


class MSE(nn.Module):
    """evaluation/metrics/mse.py:12-21."""

    def forward(self, reference_observations: torch.Tensor, generated_observations: torch.Tensor) -> torch.Tensor:
        return ops.sqdiff_mean(reference_observations, generated_observations)          # one fused reduction kernel


class PSNR(nn.Module):
    """evaluation/metrics/psnr.py:10-28: -10 log10(mse + 1e-8) of the observations divided by ``range``."""

    def forward(self, reference_observations: torch.Tensor, generated_observations: torch.Tensor, range=1.0) -> torch.Tensor:
        if range == 1.0:
            mse = ops.sqdiff_mean(reference_observations, generated_observations)
        else:
            mse = ops.sqdiff_mean(reference_observations / range, generated_observations / range)
        return -10 * torch.log10(mse + 1e-8)


def frame_difference_motion_mask(observations: torch.Tensor) -> torch.Tensor:
    """evaluation/metrics/motion_mask.py:13-34: |frame_t - frame_{t-1}| averaged over the 3 channels, zero for the first frame.
    (bs, T, 3, h, w) -> (bs, T, 1, h, w)."""
    assert observations.size(2) == 3
    mask = torch.abs(observations[:, 1:] - observations[:, :-1]).sum(dim=2, keepdim=True) / 3
    return torch.cat([torch.zeros_like(mask[:, 0:1]), mask], dim=1)


class MotionMaskedMSE(nn.Module):
    """evaluation/metrics/motion_masked_mse.py:14-27: squared error weighted by the reference's frame-difference mask."""

    def forward(self, reference_observations: torch.Tensor, generated_observations: torch.Tensor) -> torch.Tensor:
        if reference_observations.size(2) != 3:
            raise AssertionError("the motion mask is defined on RGB frames (evaluation/metrics/motion_mask.py:22)")
        return ops.sqdiff_mean(reference_observations, generated_observations, motion_mask=True)     # mask computed in the same pass


class VGGCosineSimilarity(nn.Module):
    """evaluation/metrics/vgg_cosine_similarity.py:10-61: mean over the five VGG19 feature levels of the cosine similarity
    (dim = flattened features, eps 1e-6) between the two branches; inputs are first mapped by (x / range - 0.5) / (0.5 + 1e-6)."""

    def __init__(self, vgg: Optional[Vgg19] = None):
        super().__init__()
        self.vgg = vgg if vgg is not None else Vgg19()

    def forward(self, reference_observations: torch.Tensor, generated_observations: torch.Tensor, range=1.0) -> torch.Tensor:
        bs, count = reference_observations.shape[:2]

        def prep(x):
            x = (x / range - 0.5) / (0.5 + 1e-6)
            return x.reshape((bs * count,) + tuple(x.shape[2:]))

        with torch.no_grad():
            ref_feats = self.vgg(prep(reference_observations))
            gen_feats = self.vgg(prep(generated_observations))
        cos = nn.CosineSimilarity(dim=1, eps=1e-6)
        sims = torch.zeros((bs, count), dtype=torch.float32, device=reference_observations.device)
        for a, b in zip(ref_feats, gen_feats):
            sims += cos(a.reshape(bs * count, -1), b.reshape(bs * count, -1)).reshape(bs, count)
        return sims / len(ref_feats)
