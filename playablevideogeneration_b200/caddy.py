"""CADDY (encoder E, action network A, ConvLSTM dynamics R, decoder D) on the pvg_b200 kernels.

Host-side mirror of the reference's module protocol: the module tree registers exactly the reference's
parameter / buffer names and shapes (OIHW weights), so ``load_state_dict`` of a reference checkpoint works and
``state_dict()`` round-trips, and ``Model`` exposes ``forward`` (20-tuple), ``start_inference``, ``generate_next`` and
``generate_next_interpolation`` with the reference's signatures, error behaviour and CPU-generator RNG draw order.
``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear`` objects are used as parameter containers only - their ``forward`` is
never called; all convolution / normalisation / resampling / recurrent arithmetic runs in ``ops`` (hand-written
sm_100a kernels).  Tiny (B x T x <=7) vector algebra of the action head stays in torch.

Reference files restated here: model/main_model/{model,representation_network,action_network,conv_dynamics_network,
rendering_network}.py, model/reduced_model/*, model/layers/*.py.
"""
from __future__ import annotations

import random
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_TANH

SLOPE = 0.2


class NoiseSource:
    """Where the model's random numbers come from.

    ``cpu`` (default): drawn on the global CPU generator at the point of use and copied to the device - exactly the
    reference's behaviour and draw order (action_network.py:45, gumbel_softmax.py:33, model.py:496).
    ``record``: like ``cpu`` but remembers the sequence of draws of one step.
    ``static``: every draw returns a persistent device buffer; ``refill()`` (host side, before a CUDA-graph replay)
    redraws the whole sequence on the CPU generator in the recorded order and uploads it.  This keeps the reference's
    RNG stream while letting the step be captured in a CUDA graph (no host work between kernels)."""

    def __init__(self):
        self.mode = "cpu"
        self.plan = []          # (kind, shape, device or None)
        self.bufs = []
        self.cursor = 0

    def _draw(self, kind, shape):
        return torch.randn(shape, dtype=torch.float32) if kind == "randn" else torch.rand(shape)

    def sample(self, kind, shape, device):
        shape = tuple(shape)
        if self.mode == "static":
            buf = self.bufs[self.cursor]
            self.cursor += 1
            if device is None:
                return None
            assert buf is not None and tuple(buf.shape) == shape
            return buf
        t = self._draw(kind, shape)
        if self.mode == "record":
            self.plan.append((kind, shape, device))
        return t.to(device) if device is not None else t

    def begin_step(self):
        self.cursor = 0
        if self.mode == "record":
            self.plan = []

    def freeze(self):
        """All recorded draws become views of ONE device buffer, mirrored by ONE pinned host buffer: a refill is the CPU draws
        (in the recorded order - the reference's RNG stream) plus a single asynchronous upload.  (One pageable copy per draw -
        about 60 per training step - blocked the host behind the running replay and left the GPU idle for ~1 ms per step.)"""
        sizes = [int(torch.Size(shape).numel()) if dev is not None else 0 for _, shape, dev in self.plan]
        devs = [dev for _, _, dev in self.plan if dev is not None]
        total = sum(sizes)
        self._flat = torch.empty((max(total, 1),), dtype=torch.float32, device=devs[0]) if devs else None
        self._host = torch.empty((max(total, 1),), dtype=torch.float32)
        if self._flat is not None and self._flat.is_cuda:
            self._host = self._host.pin_memory()
        self._host_event = self._commit_event = self._stage_buf = None
        self._staged = False
        self.bufs, self._spans, off = [], [], 0
        for (kind, shape, dev), n in zip(self.plan, sizes):
            if dev is None:
                self.bufs.append(None)
                self._spans.append(None)
            else:
                self.bufs.append(self._flat[off:off + n].view(shape))
                self._spans.append((off, n))
                off += n
        self.mode = "static"
        self.refill()

    def refill(self):
        """Draw the next step's numbers and put them where the (captured) step reads them."""
        self.stage()
        self.commit()

    def stage(self, stream=None):
        """First half of a refill: the CPU draws (recorded order = the reference's RNG stream) and their upload into a STAGING
        device buffer, optionally on a side stream - safe while a replay that reads the live buffer is still running."""
        if self._host_event is not None:
            self._host_event.synchronize()              # the previous upload has read the pinned buffer
        for (kind, shape, dev), span in zip(self.plan, self._spans):
            t = self._draw(kind, shape)                 # drawn even when unused: keeps the CPU RNG stream of the reference
            if span is not None:
                self._host[span[0]:span[0] + span[1]].copy_(t.reshape(-1))
        if self._flat is None:
            return
        if not self._flat.is_cuda:
            self._stage_buf = self._host
            self._staged = True
            return
        if self._stage_buf is None:
            self._stage_buf = torch.empty_like(self._flat)
        cur = torch.cuda.current_stream()
        st = stream if stream is not None else cur
        if self._commit_event is not None:
            st.wait_event(self._commit_event)           # the last commit has read the staging buffer
        with torch.cuda.stream(st):
            self._stage_buf.copy_(self._host, non_blocking=True)
            self._host_event = torch.cuda.Event()
            self._host_event.record(st)
        self._staged = True

    def commit(self):
        """Second half: staging -> live buffer, device to device on the current stream (in front of the replay)."""
        if self._flat is not None:
            if not self._staged:
                self.stage()
            if self._flat.is_cuda:
                torch.cuda.current_stream().wait_event(self._host_event)
                self._flat.copy_(self._stage_buf, non_blocking=True)
                self._commit_event = torch.cuda.Event()
                self._commit_event.record()
            else:
                self._flat.copy_(self._stage_buf)
        self._staged = False
        self.cursor = 0


def _conv(cin, cout, k, bias):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=(k - 1) // 2, bias=bias)


_DEFAULT_NOISE = NoiseSource()


def _fold_eval_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """Inference: BatchNorm (running statistics) folded into the preceding bias-free convolution - ``w' = w * s``, ``b' = beta -
    mean * s`` with ``s = gamma / sqrt(var + eps)`` - so conv -> BN -> LeakyReLU is ONE launch (bias + activation in the conv
    epilogue) instead of three.  Cached on the conv module; rebuilt when the weight, the BatchNorm parameters / statistics or
    the optimiser epoch move."""
    parts = (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple(-1 if t is None else t._version for t in parts) + (ops.weights_epoch, conv.weight.data_ptr())
    hit = getattr(conv, "_pvg_folded", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    with torch.no_grad():
        s = torch.rsqrt(bn.running_var + bn.eps)
        if bn.weight is not None:
            s = s * bn.weight
        w = (conv.weight * s[:, None, None, None]).contiguous()
        b = -bn.running_mean * s
        if bn.bias is not None:
            b = b + bn.bias
        b = b.contiguous()
    conv._pvg_folded = (key, w, b)
    return w, b


def _can_fold(bn: nn.BatchNorm2d) -> bool:
    return (not bn.training) and (not torch.is_grad_enabled()) and bn.running_mean is not None and ops.fold_eval_batchnorm


class ResidualBlock(nn.Module):
    """model/layers/residual_block.py:14-68."""

    def __init__(self, in_planes, out_planes, downsample_factor=1):
        super().__init__()
        if downsample_factor not in (1, 2):
            raise ValueError("downsample_factor must be 1 or 2")
        self.conv1 = _conv(in_planes, out_planes, 3, False)
        self.bn1 = nn.BatchNorm2d(out_planes)
        self.conv2 = _conv(out_planes, out_planes, 3, False)
        self.bn2 = nn.BatchNorm2d(out_planes)
        self.downsample_factor = downsample_factor
        self.downsample = None
        if downsample_factor != 1 or in_planes != out_planes:
            self.downsample = nn.Sequential(_conv(in_planes, out_planes, 1, False), nn.AvgPool2d(downsample_factor),
                                            nn.BatchNorm2d(out_planes))

    def forward(self, x, out_planes: bool = False, groups: int = 1):
        """``groups``: the batch holds that many independent reference calls (time steps) - BatchNorm statistics per chunk.
        ``out_planes``: the block's output feeds another tensor-core convolution - the last BatchNorm pass writes that
        convolution's 16-bit operand planes next to the fp32 result (no separate split pass).
        A block whose channel count is not a multiple of 8 (the encoder's 65-channel tail, representation_network.py:28)
        computes on tensors physically padded with zero channels to the next multiple of 8: its three convolutions and their
        gradients then run on the tensor cores instead of the fp32 CUDA-core kernels.  The result keeps the padding."""
        pool = self.downsample_factor == 2
        need = ops.conv_input_planes()
        cout = self.conv1.out_channels
        cphys = (cout + 7) // 8 * 8 if (cout % 8 and x.shape[1] % 8 == 0 and ops.supports_padded_cout()) else None
        fold = cphys is None and not pool and _can_fold(self.bn1)
        if fold:                # inference: conv1 -> bn1 -> LeakyReLU in one launch
            w1, b1 = _fold_eval_bn(self.conv1, self.bn1)
            out = ops.conv2d(x, w1, b1, act=ACT_LRELU, slope=SLOPE, out_planes=True)
        else:
            # training: the statistics of an un-pooled BatchNorm input are accumulated by the convolution's own epilogue
            st = groups if self.bn1.training else 0
            out = ops.conv2d(x, self.conv1.weight, cout_phys=cphys, bn_stats_groups=0 if pool else st)
            out = ops.pool_bn_act(out, self.bn1, pool=pool, act=ACT_LRELU, slope=SLOPE, planes=need, groups=groups)
        out = ops.conv2d(out, self.conv2.weight, cout_phys=cphys, bn_stats_groups=groups if self.bn2.training else 0)
        if self.downsample is not None:
            if fold:
                wd, bd = _fold_eval_bn(self.downsample[0], self.downsample[2])
                idn = ops.conv2d(x, wd, bd)
            else:
                idn = ops.conv2d(x, self.downsample[0].weight, cout_phys=cphys,
                                 bn_stats_groups=groups if (self.downsample[2].training and not pool) else 0)
                idn = ops.pool_bn_act(idn, self.downsample[2], pool=pool, act=ACT_NONE, groups=groups)
        else:
            idn = x
        return ops.pool_bn_act(out, self.bn2, residual=idn, act=ACT_LRELU, slope=SLOPE, planes=need if out_planes else (),
                               groups=groups)


class SameBlock(nn.Module):
    """model/layers/same_block.py:10-47."""

    def __init__(self, in_planes, out_planes, downsample_factor=1):
        super().__init__()
        self.downsample_factor = downsample_factor
        self.conv1 = _conv(in_planes, out_planes, 3, False)
        self.bn1 = nn.BatchNorm2d(out_planes)

    def forward(self, x):
        if self.downsample_factor != 2 and _can_fold(self.bn1):
            w, b = _fold_eval_bn(self.conv1, self.bn1)
            return ops.conv2d(x, w, b, act=ACT_LRELU, slope=SLOPE)
        pool = self.downsample_factor == 2
        out = ops.conv2d(x, self.conv1.weight, bn_stats_groups=1 if (self.bn1.training and not pool) else 0)
        return ops.pool_bn_act(out, self.bn1, pool=pool, act=ACT_LRELU, slope=SLOPE)


class UpBlock(nn.Module):
    """model/layers/up_block.py:5-44 (bilinear x2, before or after the conv block)."""

    def __init__(self, in_features, out_features, late_upscaling=False):
        super().__init__()
        self.late_upscaling = late_upscaling
        self.conv = _conv(in_features, out_features, 3, False)
        self.norm = nn.BatchNorm2d(out_features, affine=True)

    def forward(self, x, out_planes: bool = False, groups: int = 1):
        if not self.late_upscaling:
            x = ops.upsample2x(x, planes=ops.conv_input_planes())
        if _can_fold(self.norm):
            w, b = _fold_eval_bn(self.conv, self.norm)
            x = ops.conv2d(x, w, b, act=ACT_LRELU, slope=SLOPE, out_planes=out_planes and not self.late_upscaling)
        else:
            x = ops.pool_bn_act(ops.conv2d(x, self.conv.weight, bn_stats_groups=groups if self.norm.training else 0), self.norm,
                                act=ACT_LRELU, slope=SLOPE, groups=groups,
                                planes=ops.conv_input_planes() if (out_planes and not self.late_upscaling) else ())
        if self.late_upscaling:
            x = ops.upsample2x(x)
        return x


class FinalBlock(nn.Module):
    """model/layers/final_block.py:9-29: conv (+bias) -> tanh, fused in the conv epilogue."""

    def __init__(self, in_planes, out_planes, kernel_size=3):
        super().__init__()
        self.conv = _conv(in_planes, out_planes, kernel_size, True)

    def forward(self, x):
        return ops.conv2d(x, self.conv.weight, self.conv.bias, act=ACT_TANH)


class ConvLSTMCell(nn.Module):
    """Parameter container of model/layers/convolutional_lstm_cell.py:6-25 (four gate convolutions)."""

    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.input_gate = _conv(in_planes + out_planes, out_planes, 3, True)
        self.forget_gate = _conv(in_planes + out_planes, out_planes, 3, True)
        self.output_gate = _conv(in_planes + out_planes, out_planes, 3, True)
        self.cell_gate = _conv(in_planes + out_planes, out_planes, 3, True)

    def fused(self, interleaved: bool = False):
        """The four gate convolutions as one weight / bias: stacked [i; f; o; g] (N = 4C), or with the output channels
        INTERLEAVED (row 4c + gate) - the layout whose conv epilogue holds all four gates of a channel in one thread."""
        gates = (self.input_gate, self.forget_gate, self.output_gate, self.cell_gate)
        if interleaved:
            w = torch.stack([g.weight for g in gates], dim=1)
            return w.reshape((-1,) + tuple(w.shape[2:])), torch.stack([g.bias for g in gates], dim=1).reshape(-1)
        return torch.cat([g.weight for g in gates], 0), torch.cat([g.bias for g in gates], 0)


class ConvLSTM(nn.Module):
    """model/layers/convolutional_lstm.py:9-74.  The four gate convolutions run as ONE implicit GEMM (N = 4C) over the
    zero-padded channel concat [inputs..., h]; sigmoid/tanh/c-update/h-output are one fused point-wise kernel."""

    def __init__(self, in_planes, out_planes, size):
        super().__init__()
        self.cell = ConvLSTMCell(in_planes, out_planes)
        self.initial_hidden_state = nn.Parameter(torch.zeros(out_planes, size[0], size[1]))
        self.initial_hidden_cell_state = nn.Parameter(torch.zeros(out_planes, size[0], size[1]))
        self._h = self._c = self._w = self._b = None

    def reinit_memory(self, batch_size: int):
        self._h = self._c = self._w = self._b = None

    def forward(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        batch = inputs[0].size(0)
        if self._h is None:
            self._h = ops.nhwc(self.initial_hidden_state.unsqueeze(0).expand(batch, -1, -1, -1))
            self._c = ops.nhwc(self.initial_hidden_cell_state.unsqueeze(0).expand(batch, -1, -1, -1))
        fused_cell = ops.supports_fused_lstm() and self.cell.input_gate.out_channels % 4 == 0
        if self._w is None:
            self._w, self._b = self.cell.fused(interleaved=fused_cell)
        z = ops.concat_pad(list(inputs) + [self._h], planes=ops.conv_input_planes())
        if fused_cell:          # gate convolution with sigmoid / tanh / c-update / h-output fused into its epilogue: one launch
            self._h, self._c = ops.convlstm_step(z, self._w, self._b, self._c)
        else:
            gates = ops.conv2d(z, self._w, self._b)
            self._h, self._c = ops.lstm_cell(gates, self._c)
        return self._h


class ConvDynamicsNetwork(nn.Module):
    """model/main_model/conv_dynamics_network.py:14-133."""

    def __init__(self, config):
        super().__init__()
        hid = config["model"]["dynamics_network"]["hidden_state_size"]
        res = config["model"]["representation_network"]["state_resolution"]
        sf = config["model"]["representation_network"]["state_features"]
        aux = config["data"]["actions_count"] + config["model"]["action_network"]["action_space_dimension"]
        self.recurrent_layers = [ConvLSTM(sf + aux, hid, res), ConvLSTM(2 * hid + aux, 2 * hid, (res[0] // 2, res[1] // 2)),
                                 ConvLSTM(hid + aux, hid, res)]
        self.recurrent_layers_blocks = nn.ModuleList([
            nn.Sequential(self.recurrent_layers[0], nn.BatchNorm2d(hid)),
            nn.Sequential(self.recurrent_layers[1], nn.BatchNorm2d(2 * hid)),
            nn.Sequential(self.recurrent_layers[2], nn.BatchNorm2d(hid))])
        self.non_recurrent_blocks = nn.ModuleList([SameBlock(hid + aux, 2 * hid, downsample_factor=2),
                                                   UpBlock(2 * hid + aux, hid, late_upscaling=True),
                                                   SameBlock(hid + aux, hid, downsample_factor=1)])

    def reinit_memory(self, batch_size: int):
        for layer in self.recurrent_layers:
            layer.reinit_memory(batch_size)

    def _recurrent(self, i, x, actions, variations):
        lstm, bn = self.recurrent_layers_blocks[i][0], self.recurrent_layers_blocks[i][1]
        return ops.pool_bn_act(lstm([x, actions, variations]), bn, act=ACT_NONE)

    def forward(self, states, actions, variations, random_noise=None):
        need = ops.conv_input_planes()
        x = self._recurrent(0, states, actions, variations)
        x = self.non_recurrent_blocks[0](ops.concat_pad([x, actions, variations], planes=need))
        x = self._recurrent(1, x, actions, variations)
        x = self.non_recurrent_blocks[1](ops.concat_pad([x, actions, variations], planes=need))
        x = self._recurrent(2, x, actions, variations)
        return self.non_recurrent_blocks[2](ops.concat_pad([x, actions, variations], planes=need))


class RepresentationNetwork(nn.Module):
    """model/main_model/representation_network.py:8-58."""

    def __init__(self, config):
        super().__init__()
        cin = config["training"]["batching"]["observation_stacking"] * 3
        sf = config["model"]["representation_network"]["state_features"]
        self.conv1 = _conv(cin, 16, 3, False)
        self.bn1 = nn.BatchNorm2d(16)
        self.residuals = nn.Sequential(ResidualBlock(16, 16, 1), ResidualBlock(16, 32, 2), ResidualBlock(32, 32, 1),
                                       ResidualBlock(32, 64, 2), ResidualBlock(64, 64, 1), ResidualBlock(64, sf + 1, 1))

    def forward(self, observations):
        x = ops.conv2d(observations, self.conv1.weight)
        x = ops.pool_bn_act(x, self.bn1, pool=True, act=ACT_LRELU, slope=SLOPE, planes=ops.conv_input_planes())
        last = len(self.residuals) - 1
        for i, block in enumerate(self.residuals):
            x = block(x, out_planes=i < last)
        sf = self.residuals[last].conv1.out_channels - 1          # x may carry zero padding channels beyond sf + 1
        return x[:, :sf], torch.sigmoid(x[:, sf:sf + 1])


class ActionNetwork(nn.Module):
    """model/main_model/action_network.py:9-118."""

    def __init__(self, config):
        super().__init__()
        sf = config["model"]["representation_network"]["state_features"]
        dim = config["model"]["action_network"]["action_space_dimension"]
        self.residuals = nn.Sequential(ResidualBlock(sf, 2 * sf, 2), ResidualBlock(2 * sf, 2 * sf, 1))
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.mean_fc = nn.Linear(2 * sf, dim)
        self.variance_fc = nn.Linear(2 * sf, dim)
        self.final_fc = nn.Linear(dim, config["data"]["actions_count"])

    noise = None        # NoiseSource shared with the owning Model (set by Model.__init__)

    def sample(self, mean, variance):
        src = self.noise if self.noise is not None else _DEFAULT_NOISE
        noise = src.sample("randn", mean.size(), mean.device)                     # CPU generator, as the reference (:45)
        return noise * torch.sqrt(variance) + mean

    def forward(self, states, attention):
        b, t = states.shape[:2]
        x = (states * attention).reshape((b * t,) + tuple(states.shape[2:]))
        x = self.residuals[0](x, out_planes=True)
        x = self.residuals[1](x)
        x = x.mean(dim=(2, 3))
        mean = F.linear(x, self.mean_fc.weight, self.mean_fc.bias)
        var = torch.abs(F.linear(x, self.variance_fc.weight, self.variance_fc.bias))
        state_dist = torch.stack([mean, var], dim=1).reshape(b, t, 2, -1)
        sampled_states = self.sample(mean, var).reshape(b, t, -1)
        mean, var = mean.reshape(b, t, -1), var.reshape(b, t, -1)
        dmean, dvar = mean[:, 1:] - mean[:, :-1], var[:, 1:] + var[:, :-1]
        dir_dist = torch.stack([dmean, dvar], dim=2)
        sampled_dirs = self.sample(dmean, dvar)
        logits = F.linear(sampled_dirs.reshape(b * (t - 1), -1), self.final_fc.weight, self.final_fc.bias)
        return logits.reshape(b, t - 1, -1), dir_dist, sampled_dirs, state_dist, sampled_states


class RenderingNetwork(nn.Module):
    """model/main_model/rendering_network.py:14-71; ``reduced`` selects model/reduced_model/rendering_network.py:31-41."""

    def __init__(self, config, reduced=False):
        super().__init__()
        c0, c1, c2, c3 = (64, 64, 32, 16) if reduced else (128, 128, 64, 32)
        if config["model"]["dynamics_network"]["hidden_state_size"] != c0:
            raise ValueError(f"the rendering network takes {c0} hidden-state channels")
        self.bottleneck_blocks = nn.Sequential()
        self.upsample_blocks = nn.ModuleList([nn.Sequential(UpBlock(c0, c1), ResidualBlock(c1, c1)),
                                              nn.Sequential(UpBlock(c1, c2), ResidualBlock(c2, c2)), UpBlock(c2, c3)])
        self.final_blocks = nn.ModuleList([FinalBlock(c1, 3, 3), FinalBlock(c2, 3, 3), FinalBlock(c3, 3, 7)])

    def forward(self, hidden_states, groups: int = 1):
        """``groups`` > 1: ``hidden_states`` stacks that many decoder calls of the reference along the batch axis (the
        teacher-forced steps of forward_full_model, whose frames nothing downstream waits for); every BatchNorm keeps the
        per-call batch statistics and replays the running-statistics updates in call order."""
        x = hidden_states
        outs = []
        for up, final in zip(self.upsample_blocks, self.final_blocks):
            if isinstance(up, nn.Sequential):          # UpBlock -> ResidualBlock: the UpBlock's output feeds the block's convs
                x = up[1](up[0](x, out_planes=True, groups=groups), groups=groups)
            else:
                x = up(x, groups=groups)
            outs.append(final(x))
        outs = list(reversed(outs))
        return outs[0], outs


class GumbelSoftmax(nn.Module):
    """model/layers/gumbel_softmax.py:7-72 (uniform noise drawn on the CPU generator, :33)."""

    noise = None

    def __init__(self, initial_temperature, hard=True):
        super().__init__()
        self.current_temperature = initial_temperature
        self.hard = hard

    def forward(self, logp, temperature=None):
        if temperature is not None:
            self.current_temperature = temperature
        src = self.noise if self.noise is not None else _DEFAULT_NOISE
        u = src.sample("rand", logp.size(), logp.device)
        g = -torch.log(-torch.log(u + 1e-20) + 1e-20)
        soft = F.softmax((logp + g) / self.current_temperature, dim=-1)
        if self.hard:
            hard = torch.zeros_like(soft).scatter_(1, soft.argmax(dim=-1, keepdim=True), 1.0)
            return (hard - soft).detach() + soft
        return soft


class CentroidEstimator(nn.Module):
    """model/layers/centroid_estimator.py:5-94."""

    process_group = None        # data-parallel job: sum numerator / denominator over ranks = the gathered-batch update

    def __init__(self, centroids_count, space_dimensions, alpha):
        super().__init__()
        self.centroids_count, self.space_dimensions, self.alpha = centroids_count, space_dimensions, alpha
        self.estimated_centroids = nn.Parameter(torch.randn((centroids_count, space_dimensions)), requires_grad=False)

    def get_estimated_centroids(self):
        return self.estimated_centroids

    def update_centroids(self, points_priors, centroid_assignments):
        if not self.training:
            return
        with torch.no_grad():
            means = points_priors.reshape(-1, 2, self.space_dimensions)[:, 0]
            assign = centroid_assignments.reshape(-1, self.centroids_count)
            num, den = (means.unsqueeze(1) * assign.unsqueeze(-1)).sum(0), assign.sum(0).unsqueeze(-1)
            if self.process_group is not None:
                packed = torch.cat([num, den], dim=1)
                torch.distributed.all_reduce(packed, group=self.process_group)
                num, den = packed[:, :-1], packed[:, -1:]
            est = num / den
            # the reference rebinds ``.data``; an in-place copy is numerically identical and keeps the buffer address
            # stable (CUDA-graph replays read and write the same storage)
            self.estimated_centroids.data.copy_(self.estimated_centroids * (1 - self.alpha) + est * self.alpha)

    def compute_variations(self, points, centroid_assignments):
        lead = list(points.size())[:-1]
        pts = points.reshape(-1, self.space_dimensions)
        assign = centroid_assignments.reshape(-1, self.centroids_count)
        var = (assign.unsqueeze(-1) * (pts.unsqueeze(1) - self.estimated_centroids)).sum(1)
        return var.reshape(tuple(lead + [-1]))


class GraphedRollout:
    """One CUDA graph per rollout step (E -> R -> D, eval mode) for a fixed batch and frame shape.

    play.py / interpolate.py generate one frame per call at batch 1: ~150 kernel launches of a few microseconds each, i.e.
    the step is pure launch latency.  The graph replays them with one cudaGraphLaunch.  What makes the
    step capturable: the ConvLSTM memory lives in static buffers that the captured step reads and overwrites, the inputs
    (observation, one-hot action, variation) are copied into static buffers before each replay, and the reference's unused
    per-step noise draw (model.py:496) stays on the host, outside the graph, so the CPU RNG stream is unchanged.
    Weight packs are captured by address: rebuild the graph (``Model.enable_graphed_inference``) after loading new weights."""

    def __init__(self, model: "Model", obs_shape, device):
        self.model = model
        self.key = (tuple(obs_shape), torch.device(device))
        b = obs_shape[0]
        dim = model.config["model"]["action_network"]["action_space_dimension"]
        self.obs = torch.zeros(tuple(obs_shape), dtype=torch.float32, device=device)
        self.onehot = torch.zeros((b, model.actions_count), dtype=torch.float32, device=device)
        self.var = torch.zeros((b, dim), dtype=torch.float32, device=device)
        self.lstms = list(model.dynamics_network.recurrent_layers)
        with torch.no_grad():
            self.h = [ops.nhwc(l.initial_hidden_state.detach().unsqueeze(0).expand(b, -1, -1, -1)).clone() for l in self.lstms]
            self.c = [ops.nhwc(l.initial_hidden_cell_state.detach().unsqueeze(0).expand(b, -1, -1, -1)).clone() for l in self.lstms]
            rng = torch.get_rng_state()          # warm-up and capture run the host-side noise draw: not part of the rollout
            for l in self.lstms:
                l._w, l._b = l.cell.fused(interleaved=ops.supports_fused_lstm() and l.cell.input_gate.out_channels % 4 == 0)
            for _ in range(2):                    # eager warm-up: weight packs, kernel attributes, the tf32 probe
                self._bind()
                model._rollout_step(self.obs, self.onehot, self.var)
            torch.cuda.synchronize(device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._bind()
                frames = model._rollout_step(self.obs, self.onehot, self.var)
                for l, h, c in zip(self.lstms, self.h, self.c):
                    h.copy_(l._h)
                    c.copy_(l._c)
                self.frames = frames
                self.next_obs = torch.cat([frames, self.obs[:, :-3]], dim=1)
            for l in self.lstms:                  # the Python-side memory is not used while the graph drives the rollout
                l._h = l._c = None
            torch.set_rng_state(rng)
        self.reset()

    def _bind(self):
        for l, h, c in zip(self.lstms, self.h, self.c):
            l._h, l._c = h, c

    def reset(self):
        """start_inference: the recurrent memory goes back to the learned initial state."""
        with torch.no_grad():
            for l, h, c in zip(self.lstms, self.h, self.c):
                h.copy_(l.initial_hidden_state.detach().unsqueeze(0).expand_as(h))
                c.copy_(l.initial_hidden_cell_state.detach().unsqueeze(0).expand_as(c))

    def step(self, observations, onehot, variations):
        self.obs.copy_(observations)
        self.onehot.copy_(onehot)
        self.var.copy_(variations)
        self.model.generate_noise(observations.shape[0])      # model.py:496, host-side draw, once per generated frame
        self.graph.replay()
        return self.frames.clone(), self.next_obs.clone()


class Model(nn.Module):
    """model/main_model/model.py:19-655 (``reduced=True``: model/reduced_model/model.py)."""

    def __init__(self, config, reduced: bool = False):
        super().__init__()
        self.config = config
        m = config["model"]
        self.action_network_ensable_size = m["action_network"]["ensamble_size"]
        self.random_noise_size = m["dynamics_network"]["random_noise_size"]
        self.training_observation_stacking = config["training"]["batching"]["observation_stacking"]
        self.use_ground_truth_actions = config["training"]["use_ground_truth_actions"]
        self.pretraining_detach = config["training"]["pretraining_detach"]
        self.actions_count = config["data"]["actions_count"]
        self.state_features = m["representation_network"]["state_features"]
        self.state_resolution = m["representation_network"]["state_resolution"]
        self.hidden_state_size = m["dynamics_network"]["hidden_state_size"]
        self.state_to_hidden_state_layer = nn.Sequential(_conv(self.state_features, self.hidden_state_size, 3, True))
        self.gumbel_softmax = GumbelSoftmax(m["action_network"]["gumbel_temperature"], m["action_network"]["hard_gumbel"])
        self.action_network = nn.ModuleList([ActionNetwork(config) for _ in range(self.action_network_ensable_size)])
        self.dynamics_network = ConvDynamicsNetwork(config)
        self.representation_network = RepresentationNetwork(config)
        self.rendering_network = RenderingNetwork(config, reduced)
        self.centroid_estimator = CentroidEstimator(self.actions_count, m["action_network"]["action_space_dimension"],
                                                    m["centroid_estimator"]["alpha"])
        self.train_forward_counts = 0
        self.batch_teacher_forced_decoder = True     # decode the teacher-forced steps of forward_full_model in one batch
        self.noise = NoiseSource()
        self._graph_inference = False
        self._graphed_rollout: Optional[GraphedRollout] = None
        for net in self.action_network:
            net.noise = self.noise
        self.gumbel_softmax.noise = self.noise

    # ----------------------------------------------------------------------------------------------------------
    def forward(self, batch_tuple, ground_truth_observations_init=0, pretraining=False, gumbel_temperature=None,
                action_sampler=None, action_variation_sampler=None):
        try:
            if pretraining:
                return self.forward_pretraining(batch_tuple, gumbel_temperature=gumbel_temperature,
                                                action_sampler=action_sampler, action_variation_sampler=action_variation_sampler)
            if ground_truth_observations_init <= 0:
                raise Exception("To forward the full model specify a number of ground truth observations > 0")
            return self.forward_full_model(batch_tuple, ground_truth_observations_init, gumbel_temperature=gumbel_temperature,
                                           action_sampler=action_sampler, action_variation_sampler=action_variation_sampler)
        finally:
            ops.flush_deferred()          # BatchNorm num_batches_tracked increments of this forward, one launch

    def _encode_sequence(self, observations):
        b, t = observations.shape[:2]
        flat = observations.reshape((-1,) + tuple(observations.shape[2:]))
        states_flat, att_flat = self.representation_network(flat)
        states = states_flat.reshape((b, t) + tuple(states_flat.shape[1:]))
        attention = att_flat.reshape((b, t) + tuple(att_flat.shape[1:]))
        return states_flat, states, attention

    def _action_head(self, states, attention, actions, gumbel_temperature, action_sampler, action_variation_sampler):
        """model.py:151-205 (== :357-410)."""
        net = random.choice(self.action_network)
        logits, dir_dist, sampled_dirs, state_dist, sampled_states = net(states, attention)
        b, tm1, a = logits.shape
        flat_logits = logits.reshape(-1, a)
        logp, prob = torch.log_softmax(flat_logits, dim=1), torch.softmax(flat_logits, dim=1)
        self.centroid_estimator.update_centroids(dir_dist.reshape((-1,) + tuple(dir_dist.shape[2:])), prob)
        if action_sampler is not None:
            samples = action_sampler(logp, actions[:, :-1].reshape((-1,)))
        elif self.config["model"]["action_network"]["use_gumbel"]:
            samples = self.gumbel_softmax(logp, temperature=gumbel_temperature)
        else:
            samples = torch.softmax(flat_logits, dim=1)
        if self.use_ground_truth_actions:
            raise Exception("The use of ground truth actions during training is not supported by the selected model")
        flat_dirs = sampled_dirs.reshape(b * tm1, -1)
        variations = self.centroid_estimator.compute_variations(flat_dirs, samples)
        if not self.config["model"]["action_network"]["use_variations"]:
            variations = variations * 0
        if action_variation_sampler is not None:
            variations = action_variation_sampler(flat_dirs, samples)
        samples = samples.reshape(b, tm1, -1)
        variations = variations.reshape(b, tm1, -1)
        selected = torch.argmax(samples, dim=2)
        assert logits.size(1) == states.size(1) - 1
        return net, logits, dir_dist, sampled_dirs, state_dist, sampled_states, samples, variations, selected

    def forward_full_model(self, batch_tuple, ground_truth_observations_init, gumbel_temperature=None, action_sampler=None,
                           action_variation_sampler=None):
        observations, actions, rewards, dones = batch_tuple
        b, t = observations.shape[:2]
        _, states, attention = self._encode_sequence(observations)
        (net, logits, dir_dist, sampled_dirs, state_dist, sampled_states, samples, variations,
         selected) = self._action_head(states, attention, actions, gumbel_temperature, action_sampler, action_variation_sampler)
        self.dynamics_network.reinit_memory(b)
        rec_states, rec_att, hidden_all, recs = [states[:, 0]], [attention[:, 0]], [], []
        pyramid: Optional[List[List[torch.Tensor]]] = None
        pending: List[torch.Tensor] = []            # hidden states of teacher-forced steps whose frames nobody waits for

        def render(hiddens):
            """model.py:241-243 for one step, or for several teacher-forced steps in ONE decoder launch sequence."""
            nonlocal pyramid
            k = len(hiddens)
            rec, multi = self.rendering_network(hiddens[0] if k == 1 else torch.cat(hiddens, dim=0), groups=k)
            if pyramid is None:
                pyramid = [[] for _ in multi]
            for j in range(k):
                recs.append(rec if k == 1 else rec[j * b:(j + 1) * b])
                for i, m in enumerate(multi):
                    pyramid[i].append(m if k == 1 else m[j * b:(j + 1) * b])

        for idx in range(t - 1):
            self.generate_noise(b)          # reference draws (and never uses) the noise: keeps the RNG stream aligned
            hidden = self.dynamics_network(rec_states[-1], samples[:, idx], variations[:, idx])
            hidden_all.append(hidden)
            pending.append(hidden)
            if idx + 1 < ground_truth_observations_init:
                # teacher forcing: the next state comes from the ground truth, so this step's frame is not needed yet - the
                # decoder runs later, batched with the other teacher-forced steps (per-step BatchNorm statistics kept)
                s, a = states[:, idx + 1], attention[:, idx + 1]
                if self.batch_teacher_forced_decoder and idx + 2 <= t - 1:
                    rec_states.append(s)
                    rec_att.append(a)
                    continue
                render(pending); pending = []
            else:
                render(pending); pending = []
                obs = self.compute_current_observation(idx + 1, ground_truth_observations_init, observations, recs)
                s, a = self.representation_network(obs)
            rec_states.append(s)
            rec_att.append(a)
        if pending:
            render(pending)
        f_rec_states = torch.stack(rec_states, dim=1)
        f_rec_att = torch.stack(rec_att[1:], dim=1)
        f_hidden = torch.stack(hidden_all, dim=1)
        f_pyr = [torch.stack(p, dim=1) for p in pyramid]
        r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates = net(f_rec_states, torch.stack(rec_att, dim=1))
        return (f_pyr[0], f_pyr, f_rec_states, states, f_hidden, selected, logits, samples, attention, f_rec_att,
                dir_dist, sampled_dirs, state_dist, sampled_states, variations,
                r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates)

    def forward_pretraining(self, batch_tuple, gumbel_temperature=None, action_sampler=None, action_variation_sampler=None):
        observations, actions, rewards, dones = batch_tuple
        b, t = observations.shape[:2]
        states_flat, states, attention = self._encode_sequence(observations)
        if self.pretraining_detach:
            raise Exception("Pretraining detach is not supported by the current model")
        (net, logits, dir_dist, sampled_dirs, state_dist, sampled_states, samples, variations,
         selected) = self._action_head(states, attention, actions, gumbel_temperature, action_sampler, action_variation_sampler)
        layer = self.state_to_hidden_state_layer[0]
        rec_hidden_flat = ops.conv2d(states_flat, layer.weight, layer.bias)
        rec_hidden = rec_hidden_flat.reshape((b, -1, self.hidden_state_size, self.state_resolution[0], self.state_resolution[1]))
        _, multi = self.rendering_network(rec_hidden_flat)
        f_pyr = [m.reshape((b, t) + tuple(m.shape[1:])) for m in multi]
        self.dynamics_network.reinit_memory(b)
        hidden_all = []
        for idx in range(t - 1):
            self.generate_noise(b)
            hidden_all.append(self.dynamics_network(states[:, idx], samples[:, idx], variations[:, idx]))
        f_hidden = torch.stack(hidden_all, dim=1)
        stacked = self.compute_stacked_observations(f_pyr[0])
        rs_flat, ra_flat = self.representation_network(stacked.reshape((-1,) + tuple(stacked.shape[2:])))
        f_rec_states = rs_flat.reshape((b, t) + tuple(rs_flat.shape[1:]))
        f_rec_att = ra_flat.reshape((b, t) + tuple(ra_flat.shape[1:]))
        r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates = net(f_rec_states, f_rec_att)
        return (f_pyr[0], f_pyr, f_rec_states, states, rec_hidden, f_hidden, selected, logits, samples, attention,
                dir_dist, sampled_dirs, state_dist, sampled_states, variations,
                r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates)

    # ----------------------------------------------------------------------------------------------------------
    def compute_stacked_observations(self, observations):
        """model.py:470-486."""
        seqs = [observations]
        for s in range(1, self.training_observation_stacking):
            seqs.append(torch.cat([observations[:, 0:1].repeat([1, s, 1, 1, 1]), observations[:, :-s]], dim=1))
        return torch.cat(seqs, dim=2)

    def generate_noise(self, batch_size: int):
        """model.py:488-497.  Drawn on the CPU generator like the reference; the dynamics network ignores it
        (conv_dynamics_network.py:111-133), so it is not copied to the device."""
        return self.noise.sample("randn", (batch_size, self.random_noise_size), None)

    def compute_current_observation(self, idx, ground_truth_observations_init, ground_truth_observations,
                                    all_reconstructed_observations):
        """model.py:499-543: channels go from the most recent frame to the oldest."""
        assert ground_truth_observations_init > 0
        assert len(all_reconstructed_observations) >= idx
        if idx < ground_truth_observations_init:
            return ground_truth_observations[:, idx]
        frames = []
        start = idx - self.training_observation_stacking + 1
        if start < ground_truth_observations_init:
            frames.append(ground_truth_observations[:, ground_truth_observations_init - 1,
                                                    :(ground_truth_observations_init - start) * 3])
        for f in range(max(start, ground_truth_observations_init), idx + 1):
            frames.insert(0, all_reconstructed_observations[f - 1])
        return torch.cat(frames, dim=1)

    def actions_one_hot(self, actions):
        """model.py:545-559."""
        onehot = torch.zeros((actions.size(0), self.actions_count), dtype=torch.float, device=actions.device)
        onehot.scatter_(1, actions.reshape((-1, 1)).long(), 1)
        return onehot

    # ----------------------------------------------------------------------------------------------------------
    def enable_graphed_inference(self, enabled: bool = True):
        """Opt-in: ``generate_next`` / ``generate_next_interpolation`` / ``generate_next_batch`` replay one CUDA graph per step
        (eval mode, CUDA tensors, no autograd).  Call again after loading new weights (the graph captures weight packs)."""
        self._graph_inference = enabled
        self._graphed_rollout = None
        return self

    def start_inference(self):
        """model.py:561-568."""
        self.dynamics_network.reinit_memory(batch_size=1)
        if self._graphed_rollout is not None:
            self._graphed_rollout.reset()

    def _rollout(self, observation_batch, actions_batch, variation_batch):
        """One rollout step, through the CUDA graph when graphed inference is enabled."""
        if (self._graph_inference and observation_batch.is_cuda and not self.training and not torch.is_grad_enabled()):
            key = (tuple(observation_batch.shape), observation_batch.device)
            if self._graphed_rollout is None or self._graphed_rollout.key != key:
                self._graphed_rollout = GraphedRollout(self, observation_batch.shape, observation_batch.device)
            frames, _ = self._graphed_rollout.step(observation_batch, actions_batch, variation_batch)
            return frames
        return self._rollout_step(observation_batch, actions_batch, variation_batch)

    def _device(self):
        return self.estimated_device if hasattr(self, "estimated_device") else next(self.parameters()).device

    def _rollout_step(self, observation_batch, actions_batch, variation_batch):
        state, _ = self.representation_network(observation_batch)
        self.generate_noise(observation_batch.shape[0])
        hidden = self.dynamics_network(state, actions_batch, variation_batch)
        frame, _ = self.rendering_network(hidden)
        return frame

    def generate_next(self, observation, action, noise=False):
        """model.py:570-607."""
        dev = observation.device
        dim = self.config["model"]["action_network"]["action_space_dimension"]
        actions_batch = torch.zeros((1, self.actions_count), dtype=torch.float32, device=dev)
        actions_batch[0, action] = 1.0
        if noise:
            variation = self.noise.sample("randn", (1, dim), dev)
        else:
            variation = torch.zeros((1, dim), dtype=torch.float32, device=dev)
        frame = self._rollout(observation.unsqueeze(0), actions_batch, variation).squeeze(0)
        return frame, torch.cat([frame, observation[:-3]], dim=0)

    def generate_next_interpolation(self, observation, first_action, second_action, interpolation_factor):
        """model.py:609-655."""
        dev = observation.device
        selected = second_action if interpolation_factor > 0.5 else first_action
        c = self.centroid_estimator.estimated_centroids
        point = (c[second_action] - c[first_action]) * interpolation_factor + c[first_action]
        variation = (point - c[selected]).unsqueeze(0)
        actions_batch = torch.zeros((1, self.actions_count), dtype=torch.float32, device=dev)
        actions_batch[0, selected] = 1.0
        frame = self._rollout(observation.unsqueeze(0), actions_batch, variation).squeeze(0)
        return frame, torch.cat([frame, observation[:-3]], dim=0)

    def generate_next_batch(self, observations, actions, variations=None):
        """Batched rollout step (BASELINE.json configs[4]; the reference's generate_next is hard-wired to batch 1,
        model.py:568,580,586).  observations (B, 3S, H, W), actions (B,) int -> (frames (B,3,H,W), next observations).
        Call ``dynamics_network.reinit_memory(B)`` first."""
        onehot = self.actions_one_hot(actions)
        if variations is None:
            dim = self.config["model"]["action_network"]["action_space_dimension"]
            variations = torch.zeros((observations.shape[0], dim), dtype=torch.float32, device=observations.device)
        frames = self._rollout(observations, onehot, variations)
        return frames, torch.cat([frames, observations[:, :-3]], dim=1)


def model(config):
    """Factory looked up by config string (train.py:38-39, play.py:45-46)."""
    return Model(config)


def reduced_model(config):
    return Model(config, reduced=True)
