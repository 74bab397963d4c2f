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
    """``features_state_dict``: torchvision-style ``features.<idx>.{weight,bias}`` tensors.  The reference always loads the
    ImageNet weights (``models.vgg19(pretrained=True)``, vgg.py:16); without a state dict this class looks for the same
    checkpoint in the local torch hub cache and RAISES when it is not there - a randomly initialised feature extractor is
    only built on request (``allow_random_init=True``: benchmarks and parity tests that inject seeded weights afterwards)."""

    HUB_FILES = ("vgg19-dcbb9e9d.pth",)

    def __init__(self, features_state_dict: Optional[Dict[str, torch.Tensor]] = None, allow_random_init: bool = False):
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
        if features_state_dict is None and not allow_random_init:
            features_state_dict = self._cached_imagenet_weights()
            if features_state_dict is None:
                raise RuntimeError("Vgg19() needs the torchvision ImageNet weights (features.<idx>.weight/bias): pass "
                                   "features_state_dict, place vgg19-dcbb9e9d.pth in the torch hub cache, or pass "
                                   "allow_random_init=True for a benchmark / test with injected weights")
        if features_state_dict is not None:
            self.load_features(features_state_dict)
        for p in self.parameters():
            p.requires_grad = False

    @classmethod
    def _cached_imagenet_weights(cls) -> Optional[Dict[str, torch.Tensor]]:
        import os
        try:
            hub = torch.hub.get_dir()
        except Exception:
            return None
        for name in cls.HUB_FILES:
            path = os.path.join(hub, "checkpoints", name)
            if os.path.isfile(path):
                return torch.load(path, map_location="cpu")
        return None

    def load_features(self, sd: Dict[str, torch.Tensor]) -> None:
        with torch.no_grad():
            for idx in CONV_IDX:
                self.convs[str(idx)].weight.copy_(sd[f"features.{idx}.weight"])
                self.convs[str(idx)].bias.copy_(sd[f"features.{idx}.bias"])

    def forward(self, x: torch.Tensor, l1_targets: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """The five tapped features - or, with ``l1_targets`` (the features of the ground truth), the five per-sample L1
        distances mean|target - feature| (losses.py:450-465), each computed where the feature is produced so that its
        gradient is folded into the producing convolution's backward pass (ops.tap_l1)."""
        feats = []
        convs = [v for v in CFG if v != "M"]
        fwd_planes = ops.conv_input_planes(weight_grad=False)      # frozen weights: only the forward operand format
        it = iter(CONV_IDX)
        for pos, v in enumerate(CFG):
            if v == "M":
                x = ops.maxpool2(x, planes=fwd_planes)              # every pool of VGG19 is followed by a convolution
                continue
            idx = next(it)
            conv = self.convs[str(idx)]
            next_is_conv = pos + 1 < len(CFG) and CFG[pos + 1] != "M"
            # conv -> ReLU -> conv chains: the epilogue writes the next convolution's fp16 operand planes itself
            x = ops.conv2d(x, conv.weight, conv.bias, act=ACT_RELU, out_planes=next_is_conv)
            if idx in TAPS:
                if l1_targets is not None:
                    x, dist = ops.tap_l1(x, l1_targets[len(feats)])
                    feats.append(dist)
                else:
                    feats.append(x)
        return feats
