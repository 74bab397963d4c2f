"""VGG19 feature pyramid for the perceptual loss (model/layers/vgg.py:8-56) on the pvg_b200 conv kernels:
13 x (conv3x3 + bias + ReLU fused in the conv epilogue) and 4 max-pools; weights frozen."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_RELU

CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]     # indices inside torchvision vgg19().features
TAPS = (0, 5, 10, 19, 28)                                        # relu1_1, relu2_1, relu3_1, relu4_1, relu5_1


class Vgg19(nn.Module):
    """``features_state_dict``: torchvision-style ``features.<idx>.{weight,bias}`` tensors (ImageNet weights when the
    caller has them; the reference downloads them, vgg.py:16 - there is no network here)."""

    def __init__(self, features_state_dict: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        self.convs = nn.ModuleDict()
        cin = 3
        it = iter(CONV_IDX)
        for v in CFG:
            if v == "M":
                continue
            idx = next(it)
            self.convs[str(idx)] = nn.Conv2d(cin, v, 3, padding=1)
            cin = v
        if features_state_dict is not None:
            self.load_features(features_state_dict)
        for p in self.parameters():
            p.requires_grad = False

    def load_features(self, sd: Dict[str, torch.Tensor]) -> None:
        with torch.no_grad():
            for idx in CONV_IDX:
                self.convs[str(idx)].weight.copy_(sd[f"features.{idx}.weight"])
                self.convs[str(idx)].bias.copy_(sd[f"features.{idx}.bias"])

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        feats = []
        it = iter(CONV_IDX)
        for v in CFG:
            if v == "M":
                x = ops.maxpool2(x)
                continue
            idx = next(it)
            conv = self.convs[str(idx)]
            x = ops.conv2d(x, conv.weight, conv.bias, act=ACT_RELU)
            if idx in TAPS:
                feats.append(x)
        return feats
