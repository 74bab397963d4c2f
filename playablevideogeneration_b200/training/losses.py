"""Losses of the CADDY hot path (reference: training/losses.py) on the pvg_b200 kernels.

Same class names and call signatures as the reference so a trainer can swap them in.  Pixel/feature-space work
(ground-truth resize, L1, VGG19 features, per-level |a-b| means and their backward) runs in hand-written kernels; the
(B x T x <=7)-element action statistics (KL, entropy, mutual information) are plain torch expressions.
"""
from __future__ import annotations

import sys
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..vgg import Vgg19


def _flatten(t: torch.Tensor) -> torch.Tensor:
    return t.reshape((-1,) + tuple(t.shape[2:]))


def _aligned_ground_truth(observations: torch.Tensor, reconstructed: torch.Tensor) -> torch.Tensor:
    """losses.py:71-92 / :414-450: current frame only; drop the first frame when the reconstruction is T-1 long;
    bilinear-resize to the reconstruction's resolution."""
    gt = observations[:, :, :3]
    seq, rec_seq = gt.size(1), reconstructed.size(1)
    if rec_seq != seq:
        if rec_seq != seq - 1:
            raise Exception(f"Received an input batch with sequence length {seq}, but got a reconstructed batch of {rec_seq}")
        gt = gt[:, 1:]
    gt = _flatten(gt)
    h, w = reconstructed.shape[3:]
    if gt.shape[2] != h or gt.shape[3] != w:
        gt = ops.resize_bilinear(gt, (h, w))
    return gt


class StatesLoss:
    """losses.py:14-27."""

    def __call__(self, states, reconstructed_states):
        return F.mse_loss(states, reconstructed_states)


class HiddenStatesLoss:
    """losses.py:30-53."""

    def __call__(self, hidden_states, reconstructed_hidden_states):
        seq, rec_seq = hidden_states.size(1), reconstructed_hidden_states.size(1)
        if rec_seq != seq:
            if rec_seq - 1 != seq:
                raise Exception(f"Received an input batch with sequence length {seq}, but got a reconstructed batch of {rec_seq}")
            reconstructed_hidden_states = reconstructed_hidden_states[:, 1:]
        return F.mse_loss(hidden_states, reconstructed_hidden_states)


class ObservationsLoss:
    """losses.py:56-118 (unweighted branch; the weight-mask branch is dead in the reference)."""

    def __call__(self, observations, reconstructed_observations, weight_mask=None):
        if weight_mask is not None:
            raise NotImplementedError("motion weight masks are not on the hot path (use_motion_weights defaults to False)")
        gt = _aligned_ground_truth(observations, reconstructed_observations)
        rec = _flatten(reconstructed_observations)
        return _global_l1(gt, rec)


def _global_l1(gt: torch.Tensor, rec: torch.Tensor) -> torch.Tensor:
    """nn.L1Loss(mean) over all elements = one 'sample' holding everything."""
    gt, rec = ops.nhwc(gt), ops.nhwc(rec)
    return _L1All.apply(gt, rec)


class _L1All(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gt, rec):
        out = torch.zeros((1,), dtype=torch.float64, device=rec.device)
        ops.call("pvg_absdiff_mean_fwd", gt.data_ptr(), rec.data_ptr(), 1, rec.numel(), out.data_ptr(), ops._stream())
        ctx.save_for_backward(gt, rec)
        return out.float()[0]

    @staticmethod
    def backward(ctx, g):
        gt, rec = ctx.saved_tensors
        d = torch.empty_like(rec)
        gg = g.reshape(1).float().contiguous()
        ops.call("pvg_absdiff_mean_bwd", gt.data_ptr(), rec.data_ptr(), gg.data_ptr(), 1, rec.numel(), d.data_ptr(), ops._stream())
        return None, d


class UnmeanedPerceptualLoss(nn.Module):
    """losses.py:393-491 (unweighted branch)."""

    def __init__(self, vgg: Optional[Vgg19] = None):
        super().__init__()
        self.vgg = vgg if vgg is not None else Vgg19()

    def forward(self, observations, reconstructed_observations, weight_mask=None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        if weight_mask is not None:
            raise NotImplementedError("motion weight masks are not on the hot path")
        gt = _aligned_ground_truth(observations, reconstructed_observations)
        rec = _flatten(reconstructed_observations)
        with torch.no_grad():
            gt_feats = self.vgg(gt.detach())
        dists = self.vgg(rec, l1_targets=gt_feats)          # |gt - rec|.mean(dim=[1,2,3]) per image and tap
        total = None
        singles = []
        for cur in dists:
            total = cur if total is None else total + cur
            singles.append(cur)
        # Reference quirk (losses.py:484-488): ``total_loss = current_loss`` followed by the in-place ``total_loss +=``
        # makes single_losses[0] alias the running total; callers therefore see the level TOTAL in slot 0.
        singles[0] = total
        return total, singles


class ParallelPerceptualLoss:
    """losses.py:379-390."""

    def __init__(self, vgg: Optional[Vgg19] = None):
        self.perceptual_loss = UnmeanedPerceptualLoss(vgg)

    def cuda(self):
        self.perceptual_loss.cuda()
        return self

    def __call__(self, observations, reconstructed_observations, weight_mask=None):
        total, singles = self.perceptual_loss(observations, reconstructed_observations, weight_mask)
        return total.mean(), [s.mean() for s in singles]


class KLDivergence:
    """training/losses.py:121-143 (debug-only in the reference): KL(softmax(target) || softmax(input)), batch mean."""

    def __call__(self, input_logits: torch.Tensor, target_logits: torch.Tensor) -> torch.Tensor:
        a = input_logits.size(-1)
        logp = F.log_softmax(input_logits.reshape(-1, a), dim=1)
        q = F.softmax(target_logits.reshape(-1, a), dim=1)
        return F.kl_div(logp, q, reduction="batchmean")


class KLGaussianDivergenceLoss:
    """losses.py:146-169."""

    def __call__(self, distribution_parameters):
        d = distribution_parameters.reshape(-1, 2, distribution_parameters.size(-1))
        mean, var = d[:, 0], d[:, 1]
        return -0.5 * (1 + torch.log(var) - mean.pow(2) - var).sum(dim=-1).mean()


class KLGeneralGaussianDivergenceLoss:
    """losses.py:172-209 (variances detached; log taken before the clamp)."""

    def __call__(self, distribution_parameters, reference_distribution_parameters, eps=0.05):
        dim = distribution_parameters.size(-1)
        d = distribution_parameters.reshape(-1, 2, dim)
        r = reference_distribution_parameters.reshape(-1, 2, dim)
        mean, var = d[:, 0], d[:, 1].detach()
        rmean, rvar = r[:, 0], r[:, 1].detach()
        lv, rlv = torch.log(var), torch.log(rvar)
        var, rvar = torch.clamp(var, min=eps), torch.clamp(rvar, min=eps)
        kl = rlv - lv - 1 + var / rvar + (rmean - mean).pow(2) / rvar
        return 0.5 * kl.sum(dim=-1).mean()


class EntropyLogitLoss:
    """losses.py:339-356."""

    def __call__(self, logits):
        fl = logits.reshape((-1, logits.size(-1)))
        return -1 * torch.sum(F.softmax(fl, dim=1) * F.log_softmax(fl, dim=1)) / fl.size(0)


class EntropyProbabilityLoss:
    """losses.py:359-376."""

    def __call__(self, probabilities):
        fp = probabilities.reshape((-1, probabilities.size(-1)))
        return -1 * torch.sum(fp * torch.log(fp)) / fp.size(0)


class FixedMatrixEstimator(nn.Module):
    """losses.py:212-235."""

    def __init__(self, rows, columns, initial_alpha=0.2, initial_value=None):
        super().__init__()
        self.alpha = initial_alpha
        if initial_value is None:
            initial_value = torch.full((rows, columns), 1.0 / (rows * columns), dtype=torch.float32)
        self.estimated_matrix = nn.Parameter(initial_value, requires_grad=False)

    def forward(self, latest):
        out = self.estimated_matrix * (1 - self.alpha) + latest * self.alpha
        with torch.no_grad():       # reference: ``.data = out.detach()``; in-place keeps the address stable for CUDA graphs
            self.estimated_matrix.data.copy_(out)
        return out


class _AllReduceSum(torch.autograd.Function):
    """Sum over the ranks of a data-parallel job.  Every rank then evaluates the SAME loss on the reduced value, so the
    gradient w.r.t. the local addend is the incoming gradient unchanged."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        torch.distributed.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class MutualInformationLoss(nn.Module):
    """losses.py:238-302.  ``process_group``: under data parallelism the reference evaluates this loss on the GATHERED
    batch (trainer.py:471-473 runs on GPU0 after DataParallel's gather); the joint matrix is a plain sum over samples,
    so summing the A x A partial matrices over the ranks (49 floats) reproduces it exactly."""

    process_group = None

    def compute_joint_probability_matrix(self, distribution_1, distribution_2):
        dim = distribution_1.size(-1)
        assert distribution_2.size(-1) == dim
        d1, d2 = distribution_1.reshape(-1, dim), distribution_2.reshape(-1, dim)
        assert d1.size(0) == d2.size(0)
        p = (d1.unsqueeze(2) * d2.unsqueeze(1)).sum(dim=0)
        if self.process_group is not None:
            p = _AllReduceSum.apply(p, self.process_group)
        p = (p + p.t()) / 2.0
        return p / p.sum()

    def __call__(self, distribution_1, distribution_2, lamb=1.0, eps=sys.float_info.epsilon):
        p = self.compute_joint_probability_matrix(distribution_1, distribution_2)
        n = p.size(0)
        mr = p.sum(dim=1).view(n, 1).expand(n, n)
        mc = p.sum(dim=0).view(1, n).expand(n, n)
        p = torch.where(p < eps, torch.full_like(p, eps), p)
        mr = torch.where(mr < eps, torch.full_like(mr, eps), mr)
        mc = torch.where(mc < eps, torch.full_like(mc, eps), mc)
        return -1 * (p * (torch.log(p) - lamb * torch.log(mr) - lamb * torch.log(mc))).sum()


class SmoothMutualInformationLoss(MutualInformationLoss):
    """losses.py:305-336 (EMA-smoothed joint matrix; its state is the checkpoint's "mi_estimator" entry)."""

    def __init__(self, config):
        super().__init__()
        self.actions_count = config["data"]["actions_count"]
        self.mi_estimation_alpha = config["training"]["mutual_information_estimation_alpha"]
        self.matrix_estimator = FixedMatrixEstimator(self.actions_count, self.actions_count, self.mi_estimation_alpha)

    def compute_joint_probability_matrix(self, distribution_1, distribution_2):
        return self.matrix_estimator(super().compute_joint_probability_matrix(distribution_1, distribution_2))


class MotionLossWeightMaskCalculator:
    """training/losses.py:591-647: per-pixel loss weights |frame_t - frame_{t-1}| of the ground truth plus the same of the
    reconstruction, summed over the 3 channels, + ``weight_bias``; ones for the first frame.  No gradient flows."""

    def __init__(self, weight_bias: float = 0.0):
        self.weight_bias = weight_bias

    def compute_weight_mask(self, observations: torch.Tensor, reconstructed_observations: torch.Tensor) -> torch.Tensor:
        observations = observations.detach()[:, :, :3]
        reconstructed_observations = reconstructed_observations.detach()
        t, tr = observations.size(1), reconstructed_observations.size(1)
        if tr != t:
            if tr != t - 1:
                raise Exception(f"Received an input batch with sequence length {t}, but got a reconstructed batch of {tr}")
            reconstructed_observations = torch.cat([observations[:, 0:1], reconstructed_observations], dim=1)
        mask = torch.abs(observations[:, 1:] - observations[:, :-1]) + \
            torch.abs(reconstructed_observations[:, 1:] - reconstructed_observations[:, :-1])
        assert mask.size(2) == 3
        mask = mask.sum(dim=2, keepdim=True) + self.weight_bias
        return torch.cat([torch.ones_like(mask[:, 0:1]), mask], dim=1)


class SequenceLossEvaluator:
    """training/losses.py:650-713 (used by evaluation/evaluator.py:191-203): evaluates ``loss`` at every sequence position;
    a reconstruction that is one element shorter is aligned to the right and position 0 scores 0 (and is left out of the
    average).  Returns (average, per-position tensor)."""

    def __init__(self, loss):
        self.loss = loss

    def __call__(self, ground_truth_sequence: torch.Tensor, reconstructed_sequence: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        t, tr = ground_truth_sequence.size(1), reconstructed_sequence.size(1)
        if tr != t and tr != t - 1:
            raise Exception(f"Received an input batch with sequence length {t}, but got a reconstructed batch of {tr}")
        shift = t - tr
        terms = [torch.zeros((), dtype=torch.float32, device=ground_truth_sequence.device)] * shift
        for i in range(tr):
            cur = self.loss(ground_truth_sequence[:, i + shift:i + shift + 1], reconstructed_sequence[:, i:i + 1])
            if type(cur) == tuple:                  # losses that return several tensors: the first is the term of interest
                cur = cur[0]
            terms.append(cur.reshape(()).float())
        loss_terms = torch.stack(terms)             # one stack instead of T scalar writes (each a host-visible op)
        return (loss_terms.mean() if shift == 0 else loss_terms[1:].mean()), loss_terms
