"""CPU oracle for the CADDY hot path (TEST INFRASTRUCTURE - never imported by the product).

A functional, plain-PyTorch fp32 *restatement* of the reference algorithm
(willi-menapace/PlayableVideoGeneration).  It runs on the CPU only, works directly on a flat
``state_dict`` whose keys/shapes are the reference's checkpoint layout, and is what the CUDA path is
compared against in ``tests/``, in ``__graft_entry__.smoke()`` and in ``bench.py``'s ``cpu_baseline``
leg.  Nothing under ``playablevideogeneration_b200/`` may import this file.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the *unmodified* reference (imported from
/root/reference with the shims of SURVEY.md 8c) on seeded inputs/weights/noise and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against them.

Every function cites the reference file:line it restates.  Layout is the reference's: NCHW activations,
OIHW weights, all arithmetic fp32; RNG draws come from the global CPU generator in the reference's order.
"""
from __future__ import annotations

import math
import random
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LRELU = 0.2
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
# torchvision vgg19.features indices of the 13 convs used up to relu5_1 (model/layers/vgg.py:25-34)
VGG19_CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]
# feature taps: after the ReLU of conv idx 0, 5, 10, 19, 28 (slices end at 2, 7, 12, 21, 30)
VGG19_TAPS = [0, 5, 10, 19, 28]


# --------------------------------------------------------------------------------------------------------------
# deterministic weights recipe (shared by the golden generator, the oracle tests and the GPU parity tests)
# --------------------------------------------------------------------------------------------------------------

def decoder_channels(reduced: bool) -> Tuple[int, int, int, int]:
    """main: 128->128->64->32 (model/main_model/rendering_network.py:30-42); reduced: 64->64->32->16
    (model/reduced_model/rendering_network.py:31-41)."""
    return (64, 64, 32, 16) if reduced else (128, 128, 64, 32)


def model_param_spec(cfg: dict, reduced: bool = False) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every entry of the reference ``Model.state_dict()`` in registration order.
    kinds: conv, convbias, bias, bnw, bnb, rmean, rvar, nbt, fc, state0, centroid.  Layout per SURVEY.md 8b."""
    S = cfg["training"]["batching"]["observation_stacking"]
    A = cfg["data"]["actions_count"]
    D = cfg["model"]["action_network"]["action_space_dimension"]
    sf = cfg["model"]["representation_network"]["state_features"]
    sh, sw = cfg["model"]["representation_network"]["state_resolution"]
    hid = cfg["model"]["dynamics_network"]["hidden_state_size"]
    ens = cfg["model"]["action_network"]["ensamble_size"]
    aux = A + D
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,), "bnw"), (prefix + ".bias", (c,), "bnb"),
                     (prefix + ".running_mean", (c,), "rmean"), (prefix + ".running_var", (c,), "rvar"),
                     (prefix + ".num_batches_tracked", (), "nbt")])

    def res(prefix, cin, cout, ds):
        spec.append((prefix + ".conv1.weight", (cout, cin, 3, 3), "conv"))
        bn(prefix + ".bn1", cout)
        spec.append((prefix + ".conv2.weight", (cout, cout, 3, 3), "conv"))
        bn(prefix + ".bn2", cout)
        if ds != 1 or cin != cout:
            spec.append((prefix + ".downsample.0.weight", (cout, cin, 1, 1), "conv"))
            bn(prefix + ".downsample.2", cout)

    spec.append(("state_to_hidden_state_layer.0.weight", (hid, sf, 3, 3), "conv"))
    spec.append(("state_to_hidden_state_layer.0.bias", (hid,), "bias"))
    for e in range(ens):
        p = f"action_network.{e}"
        res(p + ".residuals.0", sf, 2 * sf, 2)
        res(p + ".residuals.1", 2 * sf, 2 * sf, 1)
        spec.extend([(p + ".mean_fc.weight", (D, 2 * sf), "fc"), (p + ".mean_fc.bias", (D,), "bias"),
                     (p + ".variance_fc.weight", (D, 2 * sf), "fc"), (p + ".variance_fc.bias", (D,), "bias"),
                     (p + ".final_fc.weight", (A, D), "fc"), (p + ".final_fc.bias", (A,), "bias")])
    lstm = [(sf + aux, hid, sh, sw), (2 * hid + aux, 2 * hid, sh // 2, sw // 2), (hid + aux, hid, sh, sw)]
    for i, (cin, cout, h, w) in enumerate(lstm):
        p = f"dynamics_network.recurrent_layers_blocks.{i}"
        spec.append((p + ".0.initial_hidden_state", (cout, h, w), "state0"))
        spec.append((p + ".0.initial_hidden_cell_state", (cout, h, w), "state0"))
        for g in ("input_gate", "forget_gate", "output_gate", "cell_gate"):
            spec.append((p + f".0.cell.{g}.weight", (cout, cin + cout, 3, 3), "conv"))
            spec.append((p + f".0.cell.{g}.bias", (cout,), "bias"))
        bn(p + ".1", cout)
    p = "dynamics_network.non_recurrent_blocks"
    spec.append((p + ".0.conv1.weight", (2 * hid, hid + aux, 3, 3), "conv")); bn(p + ".0.bn1", 2 * hid)
    spec.append((p + ".1.conv.weight", (hid, 2 * hid + aux, 3, 3), "conv")); bn(p + ".1.norm", hid)
    spec.append((p + ".2.conv1.weight", (hid, hid + aux, 3, 3), "conv")); bn(p + ".2.bn1", hid)
    p = "representation_network"
    spec.append((p + ".conv1.weight", (16, 3 * S, 3, 3), "conv")); bn(p + ".bn1", 16)
    for i, (cin, cout, ds) in enumerate([(16, 16, 1), (16, 32, 2), (32, 32, 1), (32, 64, 2), (64, 64, 1),
                                         (64, sf + 1, 1)]):
        res(p + f".residuals.{i}", cin, cout, ds)
    c0, c1, c2, c3 = decoder_channels(reduced)
    p = "rendering_network"
    spec.append((p + ".upsample_blocks.0.0.conv.weight", (c1, c0, 3, 3), "conv")); bn(p + ".upsample_blocks.0.0.norm", c1)
    res(p + ".upsample_blocks.0.1", c1, c1, 1)
    spec.append((p + ".upsample_blocks.1.0.conv.weight", (c2, c1, 3, 3), "conv")); bn(p + ".upsample_blocks.1.0.norm", c2)
    res(p + ".upsample_blocks.1.1", c2, c2, 1)
    spec.append((p + ".upsample_blocks.2.conv.weight", (c3, c2, 3, 3), "conv")); bn(p + ".upsample_blocks.2.norm", c3)
    for i, (c, k) in enumerate([(c1, 3), (c2, 3), (c3, 7)]):
        spec.append((p + f".final_blocks.{i}.conv.weight", (3, c, k, k), "conv"))
        spec.append((p + f".final_blocks.{i}.conv.bias", (3,), "bias"))
    spec.append(("centroid_estimator.estimated_centroids", (A, D), "centroid"))
    return spec


def make_weights(cfg: dict, seed: int = 0, reduced: bool = False) -> Dict[str, Tensor]:
    """Seeded synthetic weights for every state_dict entry (non-trivial BN affine/running stats, non-zero initial
    LSTM state) so that every term of the path is exercised.  NOT the reference's init - both sides of every parity
    test load these same tensors."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape, kind in model_param_spec(cfg, reduced):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "fc":
            t = torch.randn(shape, generator=g) * math.sqrt(1.0 / shape[1])
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "bnw":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "bnb":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "rmean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "rvar":
            t = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif kind == "nbt":
            t = torch.zeros((), dtype=torch.long)
        elif kind == "state0":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "centroid":
            t = torch.randn(shape, generator=g)
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd


def make_vgg_weights(seed: int = 1234) -> Dict[str, Tensor]:
    """Seeded stand-in for torchvision VGG19 ImageNet weights (no network here).  Keys are
    torchvision's ``features.<idx>.{weight,bias}``.  He-init so activations keep unit scale through 13 layers."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = 3
    it = iter(VGG19_CONV_IDX)
    for v in VGG19_CFG:
        if v == "M":
            continue
        idx = next(it)
        sd[f"features.{idx}.weight"] = torch.randn((v, cin, 3, 3), generator=g) * math.sqrt(2.0 / (cin * 9))
        sd[f"features.{idx}.bias"] = torch.randn((v,), generator=g) * 0.05
        cin = v
    return sd


def make_observations(batch: int, seq: int, channels: int, height: int, width: int, seed: int = 0) -> Tensor:
    """U[-1,1] synthetic frames (SURVEY.md 8d 'Synthetic inputs')."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((batch, seq, channels, height, width), generator=g) * 2.0 - 1.0


# --------------------------------------------------------------------------------------------------------------
# layer blocks
# --------------------------------------------------------------------------------------------------------------

class _State:
    """Weights + mutable buffers.  ``train`` selects BatchNorm batch statistics (and running-stat updates)."""

    def __init__(self, sd: Dict[str, Tensor], train: bool):
        self.sd = sd
        self.train = train

    def __getitem__(self, k):
        return self.sd[k]


def _bn(st: _State, x: Tensor, p: str) -> Tensor:
    """nn.BatchNorm2d (training: biased batch var for normalisation, unbiased for the running update)."""
    if st.train:
        with torch.no_grad():
            st.sd[p + ".num_batches_tracked"] = st.sd[p + ".num_batches_tracked"] + 1
    return F.batch_norm(x, st.sd[p + ".running_mean"], st.sd[p + ".running_var"], st[p + ".weight"], st[p + ".bias"],
                        st.train, BN_MOMENTUM, BN_EPS)


def _residual_block(st: _State, x: Tensor, p: str, ds: int) -> Tensor:
    """model/layers/residual_block.py:49-68 - conv3x3 -> avgpool(ds) -> BN -> lrelu -> conv3x3 -> BN -> (+id) -> lrelu;
    identity = conv1x1 -> avgpool(ds) -> BN when present (:41-47)."""
    out = F.conv2d(x, st[p + ".conv1.weight"], padding=1)
    if ds != 1:
        out = F.avg_pool2d(out, ds)
    out = F.leaky_relu(_bn(st, out, p + ".bn1"), LRELU)
    out = F.conv2d(out, st[p + ".conv2.weight"], padding=1)
    out = _bn(st, out, p + ".bn2")
    if (p + ".downsample.0.weight") in st.sd:
        idn = F.conv2d(x, st[p + ".downsample.0.weight"])
        if ds != 1:
            idn = F.avg_pool2d(idn, ds)
        idn = _bn(st, idn, p + ".downsample.2")
    else:
        idn = x
    return F.leaky_relu(out + idn, LRELU)


def _same_block(st: _State, x: Tensor, p: str, ds: int) -> Tensor:
    """model/layers/same_block.py:36-47."""
    out = F.conv2d(x, st[p + ".conv1.weight"], padding=1)
    if ds != 1:
        out = F.avg_pool2d(out, ds)
    return F.leaky_relu(_bn(st, out, p + ".bn1"), LRELU)


def _up_block(st: _State, x: Tensor, p: str, late: bool) -> Tensor:
    """model/layers/up_block.py:30-44 (bilinear x2, align_corners=False)."""
    if not late:
        x = F.interpolate(x, scale_factor=2, mode="bilinear")
    x = F.leaky_relu(_bn(st, F.conv2d(x, st[p + ".conv.weight"], padding=1), p + ".norm"), LRELU)
    if late:
        x = F.interpolate(x, scale_factor=2, mode="bilinear")
    return x


def _final_block(st: _State, x: Tensor, p: str, pad: int) -> Tensor:
    """model/layers/final_block.py:24-29."""
    return torch.tanh(F.conv2d(x, st[p + ".conv.weight"], st[p + ".conv.bias"], padding=pad))


# --------------------------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------------------------

def encoder(st: _State, obs: Tensor) -> Tuple[Tensor, Tensor]:
    """RepresentationNetwork.forward, model/main_model/representation_network.py:32-58."""
    p = "representation_network"
    x = F.conv2d(obs, st[p + ".conv1.weight"], padding=1)
    x = F.leaky_relu(_bn(st, F.avg_pool2d(x, 2), p + ".bn1"), LRELU)
    for i, ds in enumerate([1, 2, 1, 2, 1, 1]):
        x = _residual_block(st, x, f"{p}.residuals.{i}", ds)
    return x[:, :-1], torch.sigmoid(x[:, -1:])


def _sample(mean: Tensor, var: Tensor) -> Tensor:
    """ActionNetwork.sample, model/main_model/action_network.py:36-48 (CPU generator draw)."""
    return torch.randn(mean.size(), dtype=torch.float32) * torch.sqrt(var) + mean


def action_network(st: _State, states: Tensor, attention: Tensor, ens: int = 0):
    """ActionNetwork.forward, model/main_model/action_network.py:62-118."""
    p = f"action_network.{ens}"
    B, T = states.shape[:2]
    x = (states * attention).reshape((B * T,) + tuple(states.shape[2:]))
    x = _residual_block(st, x, p + ".residuals.0", 2)
    x = _residual_block(st, x, p + ".residuals.1", 1)
    x = x.mean(dim=(2, 3))
    mean = F.linear(x, st[p + ".mean_fc.weight"], st[p + ".mean_fc.bias"])
    var = torch.abs(F.linear(x, st[p + ".variance_fc.weight"], st[p + ".variance_fc.bias"]))
    state_dist = torch.stack([mean, var], dim=1).reshape(B, T, 2, -1)
    sampled_states = _sample(mean, var).reshape(B, T, -1)
    mean, var = mean.reshape(B, T, -1), var.reshape(B, T, -1)
    dmean = mean[:, 1:] - mean[:, :-1]
    dvar = var[:, 1:] + var[:, :-1]
    dir_dist = torch.stack([dmean, dvar], dim=2)
    sampled_dirs = _sample(dmean, dvar)
    logits = F.linear(sampled_dirs.reshape(B * (T - 1), -1), st[p + ".final_fc.weight"], st[p + ".final_fc.bias"])
    return logits.reshape(B, T - 1, -1), dir_dist, sampled_dirs, state_dist, sampled_states


def gumbel_softmax(logp: Tensor, temperature: float, hard: bool = False, eps: float = 1e-20) -> Tensor:
    """GumbelSoftmax.forward, model/layers/gumbel_softmax.py:26-72 (U drawn on the CPU generator)."""
    u = torch.rand(logp.size())
    y = F.softmax((logp - torch.log(-torch.log(u + eps) + eps)) / temperature, dim=-1)
    if hard:
        hard_y = torch.zeros_like(y).scatter_(1, y.argmax(dim=-1, keepdim=True), 1.0)
        y = (hard_y - y).detach() + y
    return y


def update_centroids(st: _State, dir_dist: Tensor, probs: Tensor, alpha: float) -> None:
    """CentroidEstimator.update_centroids, model/layers/centroid_estimator.py:38-68 (training only)."""
    if not st.train:
        return
    k = "centroid_estimator.estimated_centroids"
    with torch.no_grad():
        means = dir_dist.reshape(-1, 2, dir_dist.shape[-1])[:, 0]
        est = (means.unsqueeze(1) * probs.unsqueeze(-1)).sum(0) / probs.sum(0).unsqueeze(-1)
        st.sd[k] = (st.sd[k] * (1 - alpha) + est * alpha).detach()


def compute_variations(st: _State, points: Tensor, assign: Tensor) -> Tensor:
    """CentroidEstimator.compute_variations, model/layers/centroid_estimator.py:70-94."""
    c = st.sd["centroid_estimator.estimated_centroids"]
    return (assign.unsqueeze(-1) * (points.unsqueeze(1) - c)).sum(1)


class Dynamics:
    """ConvDynamicsNetwork + ConvLSTM state, model/main_model/conv_dynamics_network.py:111-133,
    model/layers/convolutional_lstm.py:36-74, convolutional_lstm_cell.py:77-103."""

    def __init__(self, st: _State):
        self.st = st
        self.h: List[Optional[Tensor]] = [None] * 3
        self.c: List[Optional[Tensor]] = [None] * 3

    def _lstm(self, i: int, x: Tensor, aux: Tensor) -> Tensor:
        st = self.st
        p = f"dynamics_network.recurrent_layers_blocks.{i}.0"
        B = x.shape[0]
        if self.h[i] is None:
            self.h[i] = st[p + ".initial_hidden_state"].repeat((B, 1, 1, 1))
            self.c[i] = st[p + ".initial_hidden_cell_state"].repeat((B, 1, 1, 1))
        H, W = x.shape[2:]
        z = torch.cat([x, aux[:, :, None, None].expand(-1, -1, H, W), self.h[i]], dim=1)
        gate = lambda n: F.conv2d(z, st[f"{p}.cell.{n}.weight"], st[f"{p}.cell.{n}.bias"], padding=1)
        ig, fg, og = torch.sigmoid(gate("input_gate")), torch.sigmoid(gate("forget_gate")), torch.sigmoid(gate("output_gate"))
        cg = torch.tanh(gate("cell_gate"))
        self.c[i] = fg * self.c[i] + ig * cg
        self.h[i] = og * torch.tanh(self.c[i])
        return _bn(st, self.h[i], f"dynamics_network.recurrent_layers_blocks.{i}.1")

    def step(self, states: Tensor, actions: Tensor, variations: Tensor) -> Tensor:
        st = self.st
        aux = torch.cat([actions, variations], dim=1)
        nb = "dynamics_network.non_recurrent_blocks"
        cat = lambda x: torch.cat([x, aux[:, :, None, None].expand(-1, -1, x.shape[2], x.shape[3])], dim=1)
        x = self._lstm(0, states, aux)
        x = _same_block(st, cat(x), nb + ".0", 2)
        x = self._lstm(1, x, aux)
        x = _up_block(st, cat(x), nb + ".1", late=True)
        x = self._lstm(2, x, aux)
        return _same_block(st, cat(x), nb + ".2", 1)


def decoder(st: _State, hidden: Tensor) -> List[Tensor]:
    """RenderingNetwork.forward, model/main_model/rendering_network.py:52-71 -> [high ... low] resolution."""
    p = "rendering_network"
    x = _up_block(st, hidden, p + ".upsample_blocks.0.0", late=False)
    x = _residual_block(st, x, p + ".upsample_blocks.0.1", 1)
    o0 = _final_block(st, x, p + ".final_blocks.0", 1)
    x = _up_block(st, x, p + ".upsample_blocks.1.0", late=False)
    x = _residual_block(st, x, p + ".upsample_blocks.1.1", 1)
    o1 = _final_block(st, x, p + ".final_blocks.1", 1)
    x = _up_block(st, x, p + ".upsample_blocks.2", late=False)
    o2 = _final_block(st, x, p + ".final_blocks.2", 3)
    return [o2, o1, o0]


# --------------------------------------------------------------------------------------------------------------
# model forward passes
# --------------------------------------------------------------------------------------------------------------

def _action_head(st, cfg, folded_states, folded_attention, actions, gumbel_temperature, action_sampler,
                 action_variation_sampler):
    """Shared front part of forward_full_model / forward_pretraining (model.py:151-205 == :357-410)."""
    an = cfg["model"]["action_network"]
    ens = random.choice(list(range(an["ensamble_size"])))          # model.py:152 consumes python RNG
    logits, dir_dist, sampled_dirs, state_dist, sampled_states = action_network(st, folded_states, folded_attention, ens)
    B, Tm1, A = logits.shape
    flat_logits = logits.reshape(-1, A)
    logp, prob = torch.log_softmax(flat_logits, 1), torch.softmax(flat_logits, 1)
    update_centroids(st, dir_dist.reshape(-1, 2, dir_dist.shape[-1]), prob, cfg["model"]["centroid_estimator"]["alpha"])
    if action_sampler is not None:
        samples = action_sampler(logp, actions[:, :-1].reshape((-1,)))
    elif an["use_gumbel"]:
        temp = gumbel_temperature if gumbel_temperature is not None else an["gumbel_temperature"]
        samples = gumbel_softmax(logp, temp, an["hard_gumbel"])
    else:
        samples = prob
    if cfg["training"]["use_ground_truth_actions"]:
        raise Exception("The use of ground truth actions during training is not supported by the selected model")
    flat_dirs = sampled_dirs.reshape(B * Tm1, -1)
    variations = compute_variations(st, flat_dirs, samples)
    if not an.get("use_variations", True):
        variations = variations * 0
    if action_variation_sampler is not None:
        variations = action_variation_sampler(flat_dirs, samples)
    samples = samples.reshape(B, Tm1, A)
    variations = variations.reshape(B, Tm1, -1)
    return ens, logits, dir_dist, sampled_dirs, state_dist, sampled_states, samples, variations, samples.argmax(dim=2)


def _current_observation(idx, gt_init, stacking, gt_obs, recs):
    """Model.compute_current_observation, model.py:499-543 (frames most-recent-first on the channel axis)."""
    assert gt_init > 0 and len(recs) >= idx
    if idx < gt_init:
        return gt_obs[:, idx]
    frames = []
    start = idx - stacking + 1
    if start < gt_init:
        frames.append(gt_obs[:, gt_init - 1, :(gt_init - start) * 3])
    for f in range(max(start, gt_init), idx + 1):
        frames.insert(0, recs[f - 1])
    return torch.cat(frames, dim=1)


def forward_full_model(sd, cfg, batch_tuple, gt_init: int, gumbel_temperature=None, action_sampler=None,
                       action_variation_sampler=None, train: bool = True):
    """Model.forward_full_model, model/main_model/model.py:84-286.  Returns the reference's 20-tuple."""
    if gt_init <= 0:
        raise Exception("To forward the full model specify a number of ground truth observations > 0")
    st = _State(sd, train)
    obs, actions = batch_tuple[0], batch_tuple[1]
    B, T = obs.shape[:2]
    S = cfg["training"]["batching"]["observation_stacking"]
    states_flat, att_flat = encoder(st, obs.reshape((-1,) + tuple(obs.shape[2:])))
    states = states_flat.reshape((B, T) + tuple(states_flat.shape[1:]))
    attention = att_flat.reshape((B, T) + tuple(att_flat.shape[1:]))
    ens, logits, dir_dist, sampled_dirs, state_dist, sampled_states, samples, variations, selected = _action_head(
        st, cfg, states, attention, actions, gumbel_temperature, action_sampler, action_variation_sampler)
    dyn = Dynamics(st)
    rec_states, rec_att, hidden_all, recs = [states[:, 0]], [attention[:, 0]], [], []
    pyr: Optional[List[List[Tensor]]] = None
    noise_size = cfg["model"]["dynamics_network"]["random_noise_size"]
    for t in range(T - 1):
        torch.randn((B, noise_size))                      # model.py:220,496: drawn, never used by R
        h = dyn.step(rec_states[-1], samples[:, t], variations[:, t])
        outs = decoder(st, h)
        hidden_all.append(h)
        recs.append(outs[0])
        if pyr is None:
            pyr = [[] for _ in outs]
        for i, o in enumerate(outs):
            pyr[i].append(o)
        if t + 1 < gt_init:
            s, a = states[:, t + 1], attention[:, t + 1]
        else:
            s, a = encoder(st, _current_observation(t + 1, gt_init, S, obs, recs))
        rec_states.append(s)
        rec_att.append(a)
    f_rec_states = torch.stack(rec_states, 1)
    f_rec_att = torch.stack(rec_att[1:], 1)
    f_hidden = torch.stack(hidden_all, 1)
    f_pyr = [torch.stack(p, 1) for p in pyr]
    r_logits, r_dir_dist, r_sampled_dirs, r_state_dist, r_sampled_states = action_network(
        st, f_rec_states, torch.stack(rec_att, 1), ens)
    return (f_pyr[0], f_pyr, f_rec_states, states, f_hidden, selected, logits, samples, attention, f_rec_att,
            dir_dist, sampled_dirs, state_dist, sampled_states, variations,
            r_logits, r_dir_dist, r_sampled_dirs, r_state_dist, r_sampled_states)


def _stacked_observations(obs: Tensor, stacking: int) -> Tensor:
    """Model.compute_stacked_observations, model.py:470-486."""
    seqs = [obs]
    for s in range(1, stacking):
        seqs.append(torch.cat([obs[:, 0:1].repeat([1, s, 1, 1, 1]), obs[:, :-s]], dim=1))
    return torch.cat(seqs, dim=2)


def forward_pretraining(sd, cfg, batch_tuple, gumbel_temperature=None, action_sampler=None,
                        action_variation_sampler=None, train: bool = True):
    """Model.forward_pretraining, model/main_model/model.py:290-468."""
    if cfg["training"]["pretraining_detach"]:
        raise Exception("Pretraining detach is not supported by the current model")
    st = _State(sd, train)
    obs, actions = batch_tuple[0], batch_tuple[1]
    B, T = obs.shape[:2]
    S = cfg["training"]["batching"]["observation_stacking"]
    states_flat, att_flat = encoder(st, obs.reshape((-1,) + tuple(obs.shape[2:])))
    states = states_flat.reshape((B, T) + tuple(states_flat.shape[1:]))
    attention = att_flat.reshape((B, T) + tuple(att_flat.shape[1:]))
    ens, logits, dir_dist, sampled_dirs, state_dist, sampled_states, samples, variations, selected = _action_head(
        st, cfg, states, attention, actions, gumbel_temperature, action_sampler, action_variation_sampler)
    rec_hidden_flat = F.conv2d(states_flat, st["state_to_hidden_state_layer.0.weight"],
                               st["state_to_hidden_state_layer.0.bias"], padding=1)
    rec_hidden = rec_hidden_flat.reshape((B, T) + tuple(rec_hidden_flat.shape[1:]))
    outs = decoder(st, rec_hidden_flat)
    f_pyr = [o.reshape((B, T) + tuple(o.shape[1:])) for o in outs]
    dyn = Dynamics(st)
    noise_size = cfg["model"]["dynamics_network"]["random_noise_size"]
    hidden_all = []
    for t in range(T - 1):
        torch.randn((B, noise_size))
        hidden_all.append(dyn.step(states[:, t], samples[:, t], variations[:, t]))
    f_hidden = torch.stack(hidden_all, 1)
    stacked = _stacked_observations(f_pyr[0], S)
    rs_flat, ra_flat = encoder(st, stacked.reshape((-1,) + tuple(stacked.shape[2:])))
    f_rec_states = rs_flat.reshape((B, T) + tuple(rs_flat.shape[1:]))
    f_rec_att = ra_flat.reshape((B, T) + tuple(ra_flat.shape[1:]))
    r_logits, r_dir_dist, r_sampled_dirs, r_state_dist, r_sampled_states = action_network(st, f_rec_states, f_rec_att, ens)
    return (f_pyr[0], f_pyr, f_rec_states, states, rec_hidden, f_hidden, selected, logits, samples, attention,
            dir_dist, sampled_dirs, state_dist, sampled_states, variations,
            r_logits, r_dir_dist, r_sampled_dirs, r_state_dist, r_sampled_states)


class Rollout:
    """start_inference / generate_next, model/main_model/model.py:561-607 (batch 1, eval mode)."""

    def __init__(self, sd, cfg):
        self.st = _State(sd, train=False)
        self.cfg = cfg
        self.dyn = Dynamics(self.st)

    def start_inference(self):
        self.dyn = Dynamics(self.st)

    def generate_next(self, observation: Tensor, action: int, noise: bool = False):
        cfg = self.cfg
        A = cfg["data"]["actions_count"]
        D = cfg["model"]["action_network"]["action_space_dimension"]
        onehot = torch.zeros((1, A)); onehot[0, action] = 1.0
        var = torch.randn((1, D)) if noise else torch.zeros((1, D))
        state, _ = encoder(self.st, observation.unsqueeze(0))
        torch.randn((1, cfg["model"]["dynamics_network"]["random_noise_size"]))
        h = self.dyn.step(state, onehot, var)
        frame = decoder(self.st, h)[0].squeeze(0)
        return frame, torch.cat([frame, observation[:-3]], dim=0)

    def generate_next_interpolation(self, observation: Tensor, first_action: int, second_action: int, factor: float):
        """Model.generate_next_interpolation, model/main_model/model.py:609-655: the action whose centroid is nearer on the
        interpolation line is selected (factor > 0.5 -> second), the variation is the offset of the interpolated point
        from that centroid."""
        cfg = self.cfg
        A = cfg["data"]["actions_count"]
        c = self.st.sd["centroid_estimator.estimated_centroids"]
        selected = second_action if factor > 0.5 else first_action
        point = (c[second_action] - c[first_action]) * factor + c[first_action]
        var = (point - c[selected]).unsqueeze(0)
        onehot = torch.zeros((1, A)); onehot[0, selected] = 1.0
        state, _ = encoder(self.st, observation.unsqueeze(0))
        torch.randn((1, cfg["model"]["dynamics_network"]["random_noise_size"]))
        h = self.dyn.step(state, onehot, var)
        frame = decoder(self.st, h)[0].squeeze(0)
        return frame, torch.cat([frame, observation[:-3]], dim=0)


# --------------------------------------------------------------------------------------------------------------
# losses (training/losses.py) and the trainer's weighted sum (training/trainer.py)
# --------------------------------------------------------------------------------------------------------------

def vgg19_features(vgg_sd: Dict[str, Tensor], x: Tensor) -> List[Tensor]:
    """Vgg19.forward, model/layers/vgg.py:41-56: relu1_1, relu2_1, relu3_1, relu4_1, relu5_1."""
    feats = []
    it = iter(VGG19_CONV_IDX)
    for v in VGG19_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2)
            continue
        idx = next(it)
        x = F.relu(F.conv2d(x, vgg_sd[f"features.{idx}.weight"], vgg_sd[f"features.{idx}.bias"], padding=1))
        if idx in VGG19_TAPS:
            feats.append(x)
    return feats


def _align_gt(obs: Tensor, rec: Tensor) -> Tensor:
    """losses.py:71-92 / :414-450: current frame only, drop first GT frame when rec is T-1 long, flatten (B*T)."""
    gt = obs[:, :, :3]
    if rec.shape[1] != gt.shape[1]:
        if rec.shape[1] != gt.shape[1] - 1:
            raise Exception(f"Received an input batch with sequence length {gt.shape[1]}, but got a reconstructed batch of {rec.shape[1]}")
        gt = gt[:, 1:]
    return gt.reshape((-1,) + tuple(gt.shape[2:]))


def observations_loss(obs: Tensor, rec: Tensor) -> Tensor:
    """ObservationsLoss.__call__, training/losses.py:56-118 (unweighted branch)."""
    gt = F.interpolate(_align_gt(obs, rec), tuple(rec.shape[3:]), mode="bilinear")
    return F.l1_loss(gt, rec.reshape((-1,) + tuple(rec.shape[2:])))


def perceptual_loss(vgg_sd, obs: Tensor, rec: Tensor) -> Tuple[Tensor, List[Tensor]]:
    """ParallelPerceptualLoss -> UnmeanedPerceptualLoss, training/losses.py:379-491 (unweighted branch)."""
    gt = _align_gt(obs, rec)
    h, w = rec.shape[3:]
    if gt.shape[2] != h or gt.shape[3] != w:
        gt = F.interpolate(gt, (h, w), mode="bilinear")
    with torch.no_grad():
        fg = vgg19_features(vgg_sd, gt.detach())
    fr = vgg19_features(vgg_sd, rec.reshape((-1,) + tuple(rec.shape[2:])))
    per_level = [torch.abs(a.detach() - b).mean(dim=[1, 2, 3]) for a, b in zip(fg, fr)]
    total = per_level[0]
    for l in per_level[1:]:
        total = total + l
    # Reference quirk (losses.py:484-488): ``total_loss = current_loss`` then ``total_loss += ...`` is an IN-PLACE add on
    # the tensor that is also ``single_losses[0]``, so the "level 0" entry the trainer receives is the level TOTAL.
    # The trainer's weighted sum (trainer.py:452) therefore counts level 0 once and levels 1-4 twice.
    return total.mean(), [total.mean()] + [l.mean() for l in per_level[1:]]


def kl_gaussian(dist: Tensor) -> Tensor:
    """KLGaussianDivergenceLoss, training/losses.py:146-169."""
    d = dist.reshape(-1, 2, dist.shape[-1])
    mean, var = d[:, 0], d[:, 1]
    return -0.5 * (1 + torch.log(var) - mean.pow(2) - var).sum(-1).mean()


def kl_general_gaussian(dist: Tensor, ref: Tensor, eps: float = 0.05) -> Tensor:
    """KLGeneralGaussianDivergenceLoss, training/losses.py:172-209 (log taken before the clamp)."""
    d, r = dist.reshape(-1, 2, dist.shape[-1]), ref.reshape(-1, 2, ref.shape[-1])
    mean, var = d[:, 0], d[:, 1].detach()
    rmean, rvar = r[:, 0], r[:, 1].detach()
    lv, rlv = torch.log(var), torch.log(rvar)
    var, rvar = torch.clamp(var, min=eps), torch.clamp(rvar, min=eps)
    kl = rlv - lv - 1 + var / rvar + (rmean - mean).pow(2) / rvar
    return 0.5 * kl.sum(-1).mean()


def entropy_logit(logits: Tensor) -> Tensor:
    """EntropyLogitLoss, training/losses.py:339-356."""
    fl = logits.reshape(-1, logits.shape[-1])
    return -torch.sum(F.softmax(fl, 1) * F.log_softmax(fl, 1)) / fl.shape[0]


class MutualInformation:
    """MutualInformationLoss / SmoothMutualInformationLoss + FixedMatrixEstimator, training/losses.py:212-336."""

    def __init__(self, actions_count: int, smooth_alpha: Optional[float]):
        self.alpha = smooth_alpha
        self.matrix = torch.full((actions_count, actions_count), 1.0 / (actions_count * actions_count))

    def joint(self, p1: Tensor, p2: Tensor) -> Tensor:
        d = p1.shape[-1]
        p = (p1.reshape(-1, d).unsqueeze(2) * p2.reshape(-1, d).unsqueeze(1)).sum(0)
        p = (p + p.t()) / 2.0
        p = p / p.sum()
        if self.alpha is not None:
            p = self.matrix * (1 - self.alpha) + p * self.alpha
            self.matrix = p.detach()
        return p

    def __call__(self, p1: Tensor, p2: Tensor, lamb: float = 1.0, eps: float = sys.float_info.epsilon) -> Tensor:
        p = self.joint(p1, p2)
        n = p.shape[0]
        mr = p.sum(1).view(n, 1).expand(n, n)
        mc = p.sum(0).view(1, n).expand(n, n)
        p = torch.where(p < eps, torch.full_like(p, eps), p)          # losses.py:290 (in-place masked write)
        mr = torch.where(mr < eps, torch.full_like(mr, eps), mr)
        mc = torch.where(mc < eps, torch.full_like(mc, eps), mc)
        return -(p * (torch.log(p) - lamb * torch.log(mr) - lamb * torch.log(mc))).sum()


def compute_losses(sd, vgg_sd, cfg, mi: MutualInformation, batch_tuple, gt_init: int, gumbel_temperature: float,
                   pretraining: bool = False, train: bool = True):
    """Trainer.compute_losses (training/trainer.py:400-550) / compute_losses_pretraining (:241-398): forward + every
    loss + the weighted sum (float64 accumulators, :447-449).  Returns (total (1,) f64, components dict, results)."""
    lw = cfg["training"]["loss_weights"]
    sfx = "_pretraining" if pretraining else ""
    T = batch_tuple[0].shape[1]
    if gt_init >= T:
        gt_init = T - 1
    if pretraining:
        res = forward_pretraining(sd, cfg, batch_tuple, gumbel_temperature=gumbel_temperature, train=train)
        (rec, pyr, rec_states, states, rec_hidden, hidden, selected, logits, samples, attention, dir_dist,
         sampled_dirs, state_dist, sampled_states, variations, r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates) = res
    else:
        res = forward_full_model(sd, cfg, batch_tuple, gt_init, gumbel_temperature=gumbel_temperature, train=train)
        (rec, pyr, rec_states, states, hidden, selected, logits, samples, attention, rec_att, dir_dist,
         sampled_dirs, state_dist, sampled_states, variations, r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates) = res
    obs = batch_tuple[0]
    perc = torch.zeros((1,), dtype=torch.float64)
    perc_term = torch.zeros((1,), dtype=torch.float64)
    l1 = torch.zeros((1,), dtype=torch.float64)
    comp: Dict[str, Tensor] = {}
    lam_p = lw["perceptual_loss_lambda" + sfx]
    for r, cur in enumerate(pyr):
        p_tot, p_lv = perceptual_loss(vgg_sd, obs, cur)
        term = p_lv[0] * 0.0
        for l in p_lv:
            term = term + l * lam_p
        o = observations_loss(obs, cur)
        perc = perc + p_tot
        perc_term = perc_term + term
        l1 = l1 + o
        comp[f"perceptual_loss_r{r}"] = p_tot.detach()
        comp[f"observations_rec_loss_r{r}"] = o.detach()
        for li, l in enumerate(p_lv):
            comp[f"perceptual_loss_r{r}_l{li}"] = l.detach()
    n = len(pyr)
    perc, perc_term, l1 = perc / n, perc_term / n, l1 / n
    states_rec = F.mse_loss(states.detach(), rec_states)
    ent = entropy_logit(logits)
    kl_dir = kl_gaussian(dir_dist)
    mi_loss = mi(torch.softmax(logits, -1), torch.softmax(r_logits, -1),
                 lamb=cfg["training"].get("action_mutual_information_entropy_lambda", 1.0))
    kl_state = kl_general_gaussian(r_state_dist, state_dist.detach())
    total = (lw["reconstruction_loss_lambda" + sfx] * l1 + perc_term
             + lw["states_rec_lambda" + sfx] * states_rec + lw["entropy_lambda" + sfx] * ent
             + lw["action_directions_kl_lambda" + sfx] * kl_dir
             + lw["action_mutual_information_lambda" + sfx] * mi_loss
             + lw["action_state_distribution_kl_lambda" + sfx] * kl_state)
    if pretraining:
        hid_rec = F.mse_loss(hidden, rec_hidden.detach()[:, 1:] if rec_hidden.shape[1] != hidden.shape[1] else rec_hidden.detach())
        total = total + lw["hidden_states_rec_lambda_pretraining"] * hid_rec
        comp["hidden_states_rec_loss"] = hid_rec.detach()
    comp.update({"avg_observations_rec_loss": l1.detach(), "avg_perceptual_loss": perc.detach(),
                 "perceptual_loss_term": perc_term.detach(), "states_rec_loss": states_rec.detach(),
                 "entropy_loss": ent.detach(), "action_directions_kl_loss": kl_dir.detach(),
                 "action_mutual_information_loss": mi_loss.detach(),
                 "action_state_distribution_kl_loss": kl_state.detach()})
    return total, comp, res


def adam_step(params: Sequence[Tensor], grads: Sequence[Tensor], exp_avg, exp_avg_sq, step: int, lr: float,
              weight_decay: float, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """torch.optim.Adam as the trainer configures it (training/trainer.py:36): L2 decay folded into the gradient."""
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g + weight_decay * p
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
