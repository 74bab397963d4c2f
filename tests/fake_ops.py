"""CPU stand-ins for playablevideogeneration_b200.ops, for HOST-LOGIC tests only (control flow, RNG draw order, tuple
layout, state_dict handling, loss weighting).  Installed by the ``fake_ops`` fixture via monkeypatch; the product
never imports this file and has no CPU path of its own."""
import torch
import torch.nn.functional as F

from playablevideogeneration_b200 import ops
from playablevideogeneration_b200._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH


def _act(x, act, slope):
    if act == ACT_LRELU:
        return F.leaky_relu(x, slope)
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_TANH:
        return torch.tanh(x)
    if act == ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def conv2d(x, weight, bias=None, act=ACT_NONE, slope=0.0, out_planes=False, cout_phys=None, bn_stats_groups=0):
    cin = weight.shape[1]
    return _act(F.conv2d(x[:, :cin], weight, bias, padding=weight.shape[2] // 2), act, slope)


def pool_bn_act(x, bn, residual=None, pool=False, act=ACT_NONE, slope=0.2, groups=1, planes=()):
    if pool:
        x = F.avg_pool2d(x, 2)
    n = x.shape[0] // groups                       # reference semantics: one BatchNorm call per group, in order
    y = bn(x) if groups == 1 else torch.cat([bn(x[g * n:(g + 1) * n]) for g in range(groups)], dim=0)
    if residual is not None:
        y = y + residual
    return _act(y, act, slope)


def upsample2x(x, planes=()):
    return F.interpolate(x, scale_factor=2, mode="bilinear")


def resize_bilinear(x, size):
    return F.interpolate(x.detach(), size, mode="bilinear")


def maxpool2(x, planes=()):
    return F.max_pool2d(x, 2)


def lstm_cell(gates, c_prev):
    i, f, o, g = gates.chunk(4, dim=1)
    c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c), c


def concat_pad(parts, multiple=32, planes=()):
    ref = next(p for p in parts if p.dim() == 4)
    h, w = ref.shape[2:]
    exp = [p if p.dim() == 4 else p[:, :, None, None].expand(-1, -1, h, w) for p in parts]
    out = torch.cat(exp, dim=1)
    pad = (-out.shape[1]) % multiple
    if pad:
        out = torch.cat([out, out.new_zeros(out.shape[0], pad, h, w)], dim=1)
    return out


def sqdiff_mean(reference, generated, motion_mask=False):
    d = (reference - generated).pow(2)
    if motion_mask:
        m = torch.abs(reference[:, 1:] - reference[:, :-1]).sum(dim=2, keepdim=True) / reference.shape[2]
        d = d * torch.cat([torch.zeros_like(m[:, 0:1]), m], dim=1)
    return d.mean(dim=[2, 3, 4])


def absdiff_mean(a, b):
    return (a.detach() - b).abs().reshape(a.shape[0], -1).mean(dim=1)


def tap_l1(x, target):
    return x, absdiff_mean(target, x)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    with torch.no_grad():
        g = g * grad_scale + weight_decay * p
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / (1 - beta2 ** step) ** 0.5).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - beta1 ** step))


def install(monkeypatch):
    from playablevideogeneration_b200.training import losses
    for name in ("conv2d", "pool_bn_act", "upsample2x", "resize_bilinear", "maxpool2", "lstm_cell", "concat_pad",
                 "absdiff_mean", "tap_l1", "sqdiff_mean", "adam_step"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, "nhwc", lambda x: x)
    monkeypatch.setattr(ops, "conv_input_planes", lambda weight_grad=True: ())
    monkeypatch.setattr(ops, "supports_padded_cout", lambda: False)
    monkeypatch.setattr(ops, "supports_fused_lstm", lambda: False)
    monkeypatch.setattr(ops, "fold_eval_batchnorm", False)
    monkeypatch.setattr(losses, "_global_l1", lambda gt, rec: F.l1_loss(rec, gt))
