"""Kink floor of the training-step gradients (test infrastructure, CPU only).

The CADDY loss is not smooth: L1 / perceptual terms differentiate |a - b| (a sign), the networks contain ReLU / LeakyReLU
kinks and max-pool ties.  Its gradient is therefore a discontinuous function of the forward values: two correct fp32
evaluations whose forward results differ in the last bits can land on different sides of a kink, and the gradient of the
few parameters that see that element jumps by a fixed quantum (e.g. 5.7e-4 relative for the lowest-resolution tanh head
of the pretrain_bair case) - for the reference's own CPU arithmetic as much as for the CUDA path.

This script measures that floor with the oracle itself: it re-runs the fp32 oracle with relative Gaussian noise of 1e-6
(the measured per-convolution error of the fp32-equivalent tensor-core product, tests/test_kernels_gpu.py) added to every
convolution output, for several seeds, and stores per parameter the largest relative L2 distance to the float64 gradient.
tests/test_model_gpu.py accepts a gradient when it is within 10x the unperturbed fp32 oracle's error OR within 2x this
floor.  usage: python oracle/make_kink_floor.py [case ...]   ->  tests/golden/kink_floor_<case>.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle.cases import CASES
from tests.golden_util import GOLDEN_DIR, oracle_run, rel_l2

NOISE, SEEDS = 1e-6, 8


def main():
    names = sys.argv[1:] or [n for n, c in CASES.items() if c["mode"] in ("full", "pretraining")]
    orig = F.conv2d
    for name in names:
        case = CASES[name]
        _, _, g64 = oracle_run(case, torch.float64)
        floor = {}
        for seed in range(SEEDS):
            gen = torch.Generator().manual_seed(1000 + seed)

            def noisy(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
                y = orig(x, w, b, stride, padding, dilation, groups)
                return y + y.detach().abs().mean() * NOISE * torch.randn(y.shape, generator=gen, dtype=y.dtype)

            F.conv2d = noisy
            try:
                _, _, g = oracle_run(case, torch.float32)
            finally:
                F.conv2d = orig
            for k, v in g.items():
                if k in g64 and float(g64[k].norm()) > 1e-7:
                    floor[k] = max(floor.get(k, 0.0), rel_l2(v, g64[k]))
            print(name, "seed", seed, "worst so far", f"{max(floor.values()):.3e}", flush=True)
        with open(os.path.join(GOLDEN_DIR, f"kink_floor_{name}.json"), "w") as f:
            json.dump(dict(noise=NOISE, seeds=SEEDS, floor=floor), f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
