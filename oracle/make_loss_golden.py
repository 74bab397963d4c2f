"""Golden values of the reference's small distribution losses (TEST INFRASTRUCTURE; needs /root/reference).

Runs the UNMODIFIED ``training/losses.py`` classes KLDivergence, EntropyLogitLoss, EntropyProbabilityLoss,
KLGaussianDivergenceLoss, KLGeneralGaussianDivergenceLoss on seeded inputs, values and input gradients
-> tests/golden/small_losses.npz.     usage: python oracle/make_loss_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as R           # noqa: E402

SEED = 321


def inputs():
    """(bs, observations, actions) logits x2, probabilities, (bs, observations, 2, space) Gaussian parameters x2"""
    g = torch.Generator().manual_seed(SEED)
    logits_a = torch.randn((3, 5, 7), generator=g)
    logits_b = torch.randn((3, 5, 7), generator=g)
    probs = torch.softmax(torch.randn((3, 5, 7), generator=g) * 2.0, dim=-1)
    gauss_a = torch.randn((3, 5, 2, 4), generator=g)
    gauss_b = torch.randn((3, 5, 2, 4), generator=g)
    gauss_a[:, :, 1] = gauss_a[:, :, 1].abs() * 0.5 + 0.01          # variances; some fall under the 0.05 clamp of the general KL
    gauss_b[:, :, 1] = gauss_b[:, :, 1].abs() * 0.5 + 0.01
    return logits_a, logits_b, probs, gauss_a, gauss_b


def evaluate(L):
    la, lb, pr, ga, gb = [t.clone().requires_grad_(True) for t in inputs()]
    out = {}

    def rec(name, value, *wrt):
        out[name] = value.detach().numpy()
        grads = torch.autograd.grad(value, wrt, allow_unused=True)
        for i, (t, gr) in enumerate(zip(wrt, grads)):
            out[f"{name}.grad{i}"] = (gr if gr is not None else torch.zeros_like(t)).numpy()

    rec("kl_divergence", L.KLDivergence()(la, lb), la, lb)
    rec("entropy_logit", L.EntropyLogitLoss()(la), la)
    rec("entropy_probability", L.EntropyProbabilityLoss()(pr), pr)
    rec("kl_gaussian", L.KLGaussianDivergenceLoss()(ga), ga)
    rec("kl_general_gaussian", L.KLGeneralGaussianDivergenceLoss()(ga, gb), ga, gb)
    return out


def main():
    R.install_shims()
    sys.path.insert(0, R.REF_ROOT)
    import training.losses as L
    out = evaluate(L)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "small_losses.npz")
    np.savez_compressed(path, **out)
    print({k: np.asarray(v).reshape(-1)[:2] for k, v in out.items()})


if __name__ == "__main__":
    main()
