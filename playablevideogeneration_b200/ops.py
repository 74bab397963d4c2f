"""torch.autograd.Function wrappers over the C ABI (include/pvg_b200.h).

Tensors are logical NCHW with *channels_last* strides, i.e. physically NHWC fp32 - the layout every kernel uses - so
the reference's NCHW-indexed glue (slicing, stacking, the 20-tuple the trainer unpacks) keeps working unchanged.
PyTorch is used for storage, streams and the autograd tape only; every op below launches hand-written sm_100a
kernels and raises if the extension is missing or the tensor is not on a CUDA device (no CPU / eager fallback).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ALGO_SIMT, ALGO_UMMA, ConvDesc, call

Tensor = torch.Tensor

# ---------------------------------------------------------------------------------------------------------------
# precision mode of the tensor-core convolutions
#   "tf32x3": error-compensated 3xTF32, fp32-equivalent (default: meets the loss <= 1e-5 rel parity tolerance)
#   "tf32"  : one TF32 product (what cuDNN does by default for the reference on Ampere+ GPUs)
#   "fp32"  : force the CUDA-core fp32 path everywhere (debug / cross-check)
# ---------------------------------------------------------------------------------------------------------------
_precision = os.environ.get("PVG_PRECISION", "tf32x3")
# how the two correction products of "tf32x3" are evaluated, per kernel role (forward conv / data gradient / weight gradient):
#   "tf32": all three products in TF32 (C-ABI nprod = 3; the original scheme)
#   "bf16" / "fp16": kind::f16 MMAs on 16-bit copies of the operands (nprod = 2): the corrections are 2^-11 of the result,
#           so few mantissa bits suffice.  fp16 (11 bits) holds the tf32 mantissa of a weight exactly - with bf16 weights
#           the rounding error of a weight, identical for every pixel, showed up 30x above the fp32 CPU oracle in
#           cancellation-heavy gradients of the ill-conditioned training graph - and suits the O(1) forward activations;
#           bf16 has fp32's exponent range, which the gradients (1e-6 .. 1e-12) need.  The weight gradient multiplies two
#           activation tensors (no weight operand): bf16.
#   "h3": ALL three products as kind::f16 MMAs on fp16 plane pairs {f16((x - f16(x)) * 2^12), f16(x)} (PVG_CORR_FP16_ALL): the
#           pair carries 22 bits of x, products are exact, accumulation is fp32 - the same arithmetic as TF32 + fp16
#           corrections at 3/4 of its tensor time and half of its shared-memory traffic (the conv never reads the fp32 tensor).
#           Needs operands inside fp16's range: forward activations and weights as they are; gradients after a per-tensor
#           power-of-two scaling chosen from their largest magnitude (pvg_amax + pvg_split_16_scaled), undone by the kernel.
#           "h3" for the backward roles takes effect when BOTH dgrad and wgrad select it (they share the scaled planes of dY).
CORR_MODES = ("tf32", "bf16", "fp16", "h3")
DEFAULT_CORR = {"fwd": os.environ.get("PVG_CORR", "h3"), "dgrad": os.environ.get("PVG_DGRAD_CORR", "h3"),
                "wgrad": os.environ.get("PVG_WGRAD_CORR", "h3")}
_corr = dict(DEFAULT_CORR)
_tf32_truncates: Optional[bool] = None     # does tcgen05 kind::tf32 truncate raw fp32 operands? (probed lazily)


def set_precision(mode: str) -> None:
    global _precision
    if mode not in ("tf32x3", "tf32", "fp32"):
        raise ValueError(f"unknown precision mode {mode!r}")
    _precision = mode


def get_precision() -> str:
    return _precision


def set_correction(fwd: Optional[str] = None, dgrad: Optional[str] = None, wgrad: Optional[str] = None) -> None:
    """Selects how the correction products are evaluated per kernel role; None restores that role's default."""
    for role, mode in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
        mode = DEFAULT_CORR[role] if mode is None else mode
        if mode not in CORR_MODES:
            raise ValueError(f"correction mode must be one of {CORR_MODES}, got {mode!r}")
        _corr[role] = mode


def _mode(role: str) -> Tuple[int, int]:
    """(nprod, corr_fmt) of the C ABI for the tensor-core kernel playing ``role`` under the current precision mode."""
    if _precision != "tf32x3":
        return 1, 0
    c = _corr[role]
    if c == "h3" and role != "fwd" and not (_corr["dgrad"] == "h3" and _corr["wgrad"] == "h3"):
        c = "bf16"                          # the scaled-gradient planes are shared by both backward kernels
    if c == "tf32" or (c != "h3" and not tf32_truncates()):
        return 3, 0
    return 2, {"fp16": _lib.CORR_FP16, "bf16": _lib.CORR_BF16, "h3": _lib.CORR_FP16_ALL}[c]


def _plane_dtype(fmt: int):
    return torch.bfloat16 if fmt == _lib.CORR_BF16 else torch.float16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _check_cuda(t: Tensor) -> None:
    if not t.is_cuda:
        raise _lib.PvgError("pvg_b200 ops run on CUDA tensors only (no CPU fallback exists for the CADDY hot path)")
    if t.dtype != torch.float32:
        raise _lib.PvgError(f"pvg_b200 ops take float32 tensors, got {t.dtype}")


def nhwc_strides(shape: Sequence[int]) -> Tuple[int, int, int, int]:
    n, c, h, w = shape
    return (h * w * c, 1, w * c, c)


def empty_nhwc(shape: Sequence[int], device, zero: bool = False) -> Tensor:
    t = torch.empty_strided(tuple(shape), nhwc_strides(shape), dtype=torch.float32, device=device)
    return t.zero_() if zero else t


def nhwc(x: Tensor) -> Tensor:
    """Returns x itself when it already is physically NHWC-dense, else a packed copy."""
    _check_cuda(x)
    want = nhwc_strides(x.shape)
    ok = all(s == w or d == 1 for s, w, d in zip(x.stride(), want, x.shape))
    if ok:
        return x
    out = empty_nhwc(x.shape, x.device)
    out.copy_(x)
    return out


# ---------------------------------------------------------------------------------------------------------------
# launch-count hygiene: a training step issues ~16 k kernel launches; ~2 000 of them were one-element memsets / counters
# ---------------------------------------------------------------------------------------------------------------
class _ZeroPool:
    """Zero-filled scratch for accumulators that are consumed inside the op that requests them (BatchNorm sums, split-K
    weight-gradient scratch, loss partial sums): ONE memset per optimiser step (``begin``) instead of one fill kernel per
    request.  Outside a step (``active`` false) or when the arena is exhausted, requests fall back to ``torch.zeros``.
    Arenas are never freed or resized in place - a captured CUDA graph may still point into an old one."""

    def __init__(self):
        self.buf: Optional[Tensor] = None
        self.off = 0
        self.demand = 0
        self.active = False
        self._keep: List[Tensor] = []

    def begin(self, device) -> None:
        need = int(self.demand * 1.25) + 4096
        capturing = torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing()
        if self.demand and (self.buf is None or self.buf.numel() < need or self.buf.device != device) and not capturing:
            self.buf = torch.empty((need,), dtype=torch.uint8, device=device)
            self._keep.append(self.buf)
        self.demand = 0
        self.off = 0
        self.active = self.buf is not None and self.buf.device == device
        if self.active:
            self.buf.zero_()

    def end(self) -> None:
        self.active = False

    def zeros(self, shape, dtype, device) -> Tensor:
        n = int(torch.empty((), dtype=dtype).element_size())
        for d in shape:
            n *= int(d)
        n_al = (n + 255) // 256 * 256
        self.demand += n_al
        if self.active and self.off + n_al <= self.buf.numel() and self.buf.device == device:
            v = self.buf[self.off:self.off + n].view(dtype).view(tuple(shape))
            self.off += n_al
            return v
        return torch.zeros(tuple(shape), dtype=dtype, device=device)


zero_pool = _ZeroPool()
_deferred_counts = {}        # id(tensor) -> [tensor, pending increment]


def defer_count(t: Tensor, inc: int) -> None:
    """``t += inc`` for an integer counter buffer (BatchNorm's num_batches_tracked), applied by ``flush_deferred`` in one
    multi-tensor launch instead of one tiny kernel per BatchNorm call (762 per BAIR-256 step)."""
    e = _deferred_counts.get(id(t))
    if e is None:
        _deferred_counts[id(t)] = [t, inc]
    else:
        e[1] += inc


def flush_deferred() -> None:
    if _deferred_counts:
        items = list(_deferred_counts.values())
        _deferred_counts.clear()
        torch._foreach_add_([t for t, _ in items], [c for _, c in items])


# ---------------------------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------------------------
fold_eval_batchnorm = os.environ.get("PVG_NO_BN_FOLD") != "1"      # inference: BatchNorm folded into the preceding conv (caddy.py)
weights_epoch = 0      # bumped by optimisers that update parameters through raw pointers (no autograd version bump)


def invalidate_weight_cache() -> None:
    global weights_epoch
    weights_epoch += 1


class _Packs:
    """Packed copies of one conv weight.  ``f32[0]``: forward hi/lo planes [Cout][R][S][CinK]; ``f32[2]``: data-gradient
    hi/lo planes [CinRows][R][S][CoutK] (flipped taps).  ``lo(which, 2, fmt)``: the 16-bit plane pair
    {f16(lo * 2^12), f16(hi)} of the forward (which = 0) or data-gradient (which = 2) pack, built on first use."""

    def __init__(self, fwd: Tensor, bwd: Tensor):
        self.f32 = {0: fwd, 2: bwd}
        self._bf = {}

    def hi(self, which: int) -> Tensor:
        return self.f32[which][0]

    def lo(self, which: int, nprod: int, fmt: int = 0) -> Optional[Tensor]:
        planes = self.f32[which]
        if nprod == 3:
            return planes[1]
        if nprod == 2:
            t = self._bf.get((which, fmt))
            if t is None:
                n = planes.shape[1]
                t = torch.empty((2 * n,), dtype=_plane_dtype(fmt), device=planes.device)
                call("pvg_pack_16x2", planes[0].data_ptr(), planes[1].data_ptr(), t.data_ptr(), n, fmt, _stream())
                self._bf[(which, fmt)] = t
            return t
        return None


def _get_packs(weight: Tensor, cin_rows: int, tensor_core: bool, cout_rows: Optional[int] = None) -> _Packs:
    """Packed (+tf32-split) copies of a conv weight for an activation with ``cin_rows`` physical channels, cached ON the
    tensor object (so a recycled allocation can never alias a stale pack) and rebuilt when the tensor's autograd version or
    the global weights epoch moves.  Tensor-core consumers get tf32-rounded hi planes and K-side channel counts rounded up
    to 32 (zero filled); CUDA-core consumers the exact fp32 values and the unpadded counts."""
    cache = getattr(weight, "_pvg_packs", None)
    if cache is None:
        cache = {}
        try:
            weight._pvg_packs = cache
        except Exception:
            pass
    cout, cin, r, s = weight.shape
    cout_rows = cout if cout_rows is None else cout_rows        # physical output channels (zero rows beyond the real ones)
    key = (cin_rows, tensor_core, cout_rows)
    hit = cache.get(key)
    ver = (weight._version, weights_epoch)
    if hit is not None and hit[0] == ver:
        return hit[1]
    cin_k, cout_k = (_pad32(cin_rows), _pad32(cout_rows)) if tensor_core else (cin_rows, cout_rows)
    w = weight.detach().contiguous()
    fwd = torch.empty((2, cout_rows * r * s * cin_k), dtype=torch.float32, device=weight.device)
    bwd = torch.empty((2, cin_rows * r * s * cout_k), dtype=torch.float32, device=weight.device)
    call("pvg_pack_conv_weight_ex", w.data_ptr(), cout, cin, r, s, cin_rows, cin_k, cout_k, cout_rows, 1 if tensor_core else 0,
         fwd[0].data_ptr(), fwd[1].data_ptr(), bwd[0].data_ptr(), bwd[1].data_ptr(), _stream())
    packs = _Packs(fwd, bwd)
    cache[key] = (ver, packs)
    return packs


# channel counts the tensor-core kernels take: any multiple of 8 (16-byte TMA strides of the 16-bit planes); counts that are
# not a multiple of 32 are completed with zeros by TMA out-of-bounds fill (weights are packed with the padded count)
_TC_CIN_MULTIPLE = 32 if os.environ.get("PVG_NO_OOB_PAD") == "1" else 8


def _pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def _conv_algo(cin_phys: int, cout: int = 1 << 30, ksize: int = 3, role: str = "fwd") -> Tuple[int, int, int]:
    """(algo, nprod, corr_fmt) for a conv whose A operand has cin_phys physical channels and which produces cout channels.
    The CUDA-core path (fp32, incl. the direct kernels of conv_direct.cu) takes the image-facing layers: A operands that
    are not a multiple of 32 channels wide, and outputs of <= 4 channels (tanh heads, gradients w.r.t. images)."""
    if _precision == "fp32" or cin_phys % _TC_CIN_MULTIPLE != 0 or (cout <= 4 and ksize <= (7 if cout <= 3 else 3)
                                                                      and os.environ.get("PVG_NO_DIRECT") != "1"):
        return ALGO_SIMT, 1, 0
    return (ALGO_UMMA,) + _mode(role)


def tf32_truncates() -> bool:
    """Probes (once) whether tcgen05 kind::tf32 truncates or rounds raw fp32 operands: one conv on crafted inputs."""
    global _tf32_truncates
    if _tf32_truncates is None:
        dev = torch.device("cuda")
        x = empty_nhwc((1, 32, 8, 16), dev, zero=True)
        x[0, 0, 0, 0] = 1.0 + 2.0 ** -11 + 2.0 ** -13           # trunc -> 1.0 ; round-to-nearest -> 1 + 2^-10
        w = torch.zeros((16, 32, 1, 1), device=dev)
        w[0, 0, 0, 0] = 1.0
        packs = _get_packs(w, 32, True)
        y = empty_nhwc((1, 16, 8, 16), dev)
        _run_conv(x, None, packs.hi(0), None, None, y, 1, 0, ACT_NONE, 0.0, ALGO_UMMA, 1, 0)
        v = float(y[0, 0, 0, 0])
        if v == 1.0:
            _tf32_truncates = True
        elif v == 1.0 + 2.0 ** -10:
            _tf32_truncates = False
        else:
            raise _lib.PvgError(f"tf32 probe returned {v!r}: the tensor-core conv path is broken")
    return _tf32_truncates


# ---------------------------------------------------------------------------------------------------------------
# 16-bit operand planes that travel WITH a tensor: producers (BatchNorm apply, conv epilogue, max-pool, upsample, concat) can
# write the plane pairs their consumer convolution needs in the same pass (``planes=`` arguments below); they are attached to
# the returned tensor object as ``_pvg_planes = {format: planes}`` and picked up by ``conv2d``.  Any torch op in between
# (reshape, slice, stack) simply drops the attribute and the consumer falls back to a pvg_split_16 pass.
# ---------------------------------------------------------------------------------------------------------------
def conv_input_planes(weight_grad: bool = True) -> Tuple[int, ...]:
    """Plane formats a tensor-core convolution consuming a tensor will ask for under the current precision mode: the forward
    operand format and, when its weight gradient will be taken, the weight-gradient operand format."""
    if _precision != "tf32x3" or os.environ.get("PVG_NO_PRODUCER_PLANES") == "1":
        return ()
    out = []
    n, f = _mode("fwd")
    if n == 2:
        out.append(f)
    if weight_grad and torch.is_grad_enabled():
        n, f = _mode("wgrad")
        if n == 2 and f not in out:
            out.append(f)
    return tuple(out)


def planes_of(t: Tensor) -> dict:
    return getattr(t, "_pvg_planes", None) or {}


def _attach_planes(t: Tensor, fmts: Sequence[int], planes: Sequence[Optional[Tensor]]) -> Tensor:
    d = {f: p for f, p in zip(fmts, planes) if p is not None}
    if d:
        t._pvg_planes = d
    return t


def _alloc_planes(fmts: Sequence[int], numel: int, channels: int, device) -> List[Optional[Tensor]]:
    """(up to two) plane-pair buffers for a producer kernel; none when the channel count breaks the 16-byte TMA stride rule."""
    if channels % 8 != 0:
        return [None, None]
    out = [torch.empty((2 * numel,), dtype=_plane_dtype(f), device=device) for f in list(fmts)[:2]]
    return out + [None] * (2 - len(out))


conv_profile = None     # bench.py sets this to a list: (start_event, end_event, algorithmic_flops, kernel family) per tensor-core conv launch


def _run_conv(x, x_lo, w, w_lo, bias, y, ksize, pad, act, slope, algo, nprod, fmt, alg_flops=0.0):
    n, cin, h, wd = x.shape
    d = ConvDesc(n, h, wd, cin, y.shape[1], ksize, ksize, pad, act, float(slope), algo, nprod, fmt)
    prof = conv_profile is not None and algo == ALGO_UMMA
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        profile_spin()
        e0.record()
    call("pvg_conv2d_fwd", d, x.data_ptr(), _p(x_lo), w.data_ptr(), _p(w_lo), _p(bias), y.data_ptr(), _stream())
    if prof:
        e1.record()
        kind = "single" if nprod == 1 else ("h3" if (nprod == 2 and fmt == _lib.CORR_FP16_ALL) else "tf32")
        conv_profile.append((e0, e1, alg_flops, kind))


def _split(x: Tensor, nprod: int = 3, fmt: int = 0) -> Tuple[Tensor, Tensor]:
    """(hi, lo) operands of the split product for an activation tensor: nprod == 3 -> fp32 residual plane,
    nprod == 2 -> the 16-bit plane pair {f16((x - trunc_tf32(x)) * 2^12), f16(x)} (x itself is the hi operand)."""
    if nprod == 2:
        planes = torch.empty((2 * x.numel(),), dtype=_plane_dtype(fmt), device=x.device)
        call("pvg_split_16", x.data_ptr(), planes.data_ptr(), x.numel(), fmt, _stream())
        return x, planes
    lo = torch.empty_like(x)
    if tf32_truncates():
        call("pvg_split_tf32", x.data_ptr(), None, lo.data_ptr(), x.numel(), _stream())
        return x, lo
    hi = torch.empty_like(x)
    call("pvg_split_tf32", x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    return hi, lo


def _conv_forward(x: Tensor, packs: _Packs, which: int, cout: int, ksize: int, bias, act: int, slope: float,
                  algo: int, nprod: int, fmt: int, alg_flops: float = 0.0, split=None) -> Tensor:
    """``split``: optional pre-computed (hi, lo) operands of x for this (nprod, fmt)."""
    n, cin, h, w = x.shape
    y = empty_nhwc((n, cout, h, w), x.device)
    if algo == ALGO_UMMA and nprod >= 2:
        hi, lo = split if split is not None else _split(x, nprod, fmt)
        _run_conv(hi, lo, packs.hi(which), packs.lo(which, nprod, fmt), bias, y, ksize, (ksize - 1) // 2, act, slope, algo,
                  nprod, fmt, alg_flops)
    else:
        _run_conv(x, None, packs.hi(which), None, bias, y, ksize, (ksize - 1) // 2, act, slope, algo, nprod, fmt, alg_flops)
    return y


wgrad_profile = None    # like conv_profile, for the tensor-core weight-gradient kernel
profile_spin_cycles = 0  # > 0: a spin kernel of that many clocks is queued in front of every profiled launch


def profile_spin() -> None:
    """Keeps the GPU busy while the host queues [start event, kernel, end event] of a profiled launch.  Without it an eager,
    host-bound step measures the host's launch latency: the start event is stamped by an idle GPU and the kernel arrives tens
    of microseconds later (a 43 us conv read 96 us, a 178 us one 1 ms in the first round-2 layer table)."""
    if profile_spin_cycles > 0:
        torch.cuda._sleep(int(profile_spin_cycles))
profile_shapes = {}     # id(start event) -> (role, N, H, W, Cin, Cout, R, writes planes, writes BN sums) of a profiled launch
small_wgrad_kernel = os.environ.get("PVG_NO_SMALL_WGRAD") != "1"    # pvg_conv2d_wgrad_small for 16 / 32-channel 3x3 layers
stem_kernel = os.environ.get("PVG_NO_STEM_KERNEL") != "1"      # pvg_conv2d_stem_planes for 3 -> 64 channel 3x3 layers


class Conv2dFn(torch.autograd.Function):
    """nn.Conv2d(stride 1, 'same' padding) (+ bias) (+ activation) - reference call sites listed in pvg_b200.h.
    ``weight`` is the reference's OIHW parameter; its input-channel count may be smaller than x's physical channel
    count (zero-padded concat buffers)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, slope, x_planes, out_planes, cout_phys=None, bn_groups=0):
        """``x_planes``: {format: planes} that came with x (may be empty); ``out_planes``: also return the forward-operand
        plane pair of y (all-fp16 forward kernel only)."""
        # the auxiliary outputs (operand planes, BatchNorm sums) never receive a gradient: do not let autograd allocate and
        # zero-fill stand-ins for them in the backward pass (298 fp16 fills = 6 GB of writes per BAIR-256 step, ncu launch list)
        ctx.set_materialize_grads(False)
        x_in = x
        x = nhwc(x)
        if x is not x_in:
            x_planes = {}
        _check_cuda(weight)
        cout, cin_log, r, s = weight.shape
        cin_p = x.shape[1]
        if cin_log > cin_p or r != s:
            raise _lib.PvgError(f"conv weight {tuple(weight.shape)} does not fit input with {cin_p} channels")
        algo, nprod, fmt = _conv_algo(cin_p, cout, r, "fwd")
        cphys = cout if cout_phys is None else int(cout_phys)
        if cphys != cout and not (algo == ALGO_UMMA and nprod == 2 and fmt == _lib.CORR_FP16_ALL and bias is None and cphys % 8 == 0):
            raise _lib.PvgError("physically padded outputs exist on the all-fp16 tensor-core path only (see supports_padded_cout)")
        packs = _get_packs(weight, cin_p, algo == ALGO_UMMA, cphys)
        b = bias.detach().contiguous() if bias is not None else None
        flops = 2.0 * x.shape[0] * x.shape[2] * x.shape[3] * cout * r * s * cin_log
        y_planes = None
        xp_used = None
        bn_sums = None
        if algo == ALGO_UMMA and nprod == 2 and fmt == _lib.CORR_FP16_ALL:
            # all-fp16 forward conv straight from plane pairs (and, on request, to the plane pair of y)
            xp = xp_used = x_planes[fmt] if fmt in x_planes else _split(x, nprod, fmt)[1]
            n, _, h, w = x.shape
            y = empty_nhwc((n, cphys, h, w), x.device)
            if out_planes and cphys % 8 == 0:
                y_planes = torch.empty((2 * y.numel(),), dtype=torch.float16, device=x.device)
            d = ConvDesc(n, h, w, cin_p, cphys, r, r, (r - 1) // 2, act, float(slope), ALGO_UMMA, 2, fmt)
            prof = conv_profile is not None
            if prof:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                profile_spin()
                e0.record()
            if bn_groups and act == ACT_NONE and b is None:
                bn_sums = zero_pool.zeros((bn_groups, 2, cphys), torch.float64, x.device)
            call("pvg_conv2d_fwd_planes", d, xp.data_ptr(), packs.lo(0, 2, fmt).data_ptr(), _p(b), y.data_ptr(), _p(y_planes), None,
                 _p(bn_sums), int(bn_groups) if bn_sums is not None else 0, None, _stream())
            if prof:
                e1.record()
                conv_profile.append((e0, e1, flops, "h3"))
                profile_shapes[id(e0)] = ("fwd", n, h, w, cin_p, cphys, r, bool(y_planes is not None), bool(bn_sums is not None))
        elif (algo == ALGO_SIMT and cin_p == 3 and cout == 64 and r == 3 and _precision != "fp32" and stem_kernel
              and _is_nhwc_dense(x)):
            # VGG19 conv1_1: output-bound CUDA-core kernel with coalesced stores that also writes the planes conv1_2 reads
            n, _, h, w = x.shape
            y = empty_nhwc((n, cout, h, w), x.device)
            if out_planes:
                y_planes = torch.empty((2 * y.numel(),), dtype=torch.float16, device=x.device)
            d = ConvDesc(n, h, w, 3, cout, 3, 3, 1, act, float(slope), ALGO_SIMT, 1, 0)
            call("pvg_conv2d_stem_planes", d, x.data_ptr(), packs.hi(0).data_ptr(), _p(b), y.data_ptr(), _p(y_planes), _stream())
        else:
            split = (x, x_planes[fmt]) if (algo == ALGO_UMMA and nprod == 2 and fmt in x_planes) else None
            y = _conv_forward(x, packs, 0, cout, r, b, act, slope, algo, nprod, fmt, flops, split=split)
        ctx.save_for_backward(x, weight, y if act != ACT_NONE else None)
        ctx.meta = (act, slope, bias is not None, cin_log)
        ctx.cphys = cphys
        wm = _mode("wgrad")
        ctx.x_wplanes = x_planes.get(wm[1]) if wm[0] == 2 else None      # weight-gradient operand planes that came with x
        if ctx.x_wplanes is None and wm == (2, _lib.CORR_FP16_ALL) and xp_used is not None and weight.requires_grad:
            ctx.x_wplanes = xp_used          # the forward operand planes double as the weight-gradient operand
        ctx.x_wfmt = wm[1]
        for t in (y_planes, bn_sums):
            if t is not None:
                ctx.mark_non_differentiable(t)
        return y, y_planes, bn_sums

    @staticmethod
    def backward(ctx, dy, _dy_planes=None, _dsums=None):
        if dy is None:
            return (None,) * 9
        x, weight, y = ctx.saved_tensors
        act, slope, has_bias, cin_log = ctx.meta
        cout, _, r, s = weight.shape
        n, cin_p, h, w = x.shape
        tap = pending_tap(dy)             # a feature-matching L1 gradient still to be added to dy (TapL1Fn)
        if tap is not None and (act == ACT_NONE or dy.numel() // dy.shape[0] % 8 != 0 or not _is_nhwc_dense(dy)):
            dy, tap = _materialize_tap(dy, y, tap), None
        dy = nhwc(dy)
        dmode = _mode("dgrad")            # (nprod, fmt) of the data-gradient kernel (conv_umma.cu)
        wmode = _mode("wgrad")            # ... of the weight-gradient kernel (conv_wgrad_umma.cu)
        H3 = (2, _lib.CORR_FP16_ALL)
        if ctx.cphys != cout and not (dmode == H3 and wmode == H3 and _precision == "tf32x3"):
            raise _lib.PvgError("physically padded outputs need the all-fp16 backward kernels")
        if dmode == H3 or wmode == H3:
            if (dmode == H3 and wmode == H3 and cin_p % _TC_CIN_MULTIPLE == 0 and ctx.cphys % 8 == 0 and _precision == "tf32x3"
                    and not (cin_p <= 4 and r <= 7) and not (r == 7 and cout <= 3)):
                return Conv2dFn._backward_h3(ctx, dy, x, weight, y, tap)
            # shapes the all-fp16 kernels do not take (image-facing layers, channel counts that are not a multiple of 8):
            # TF32 main product + bf16 corrections
            BF = (2, _lib.CORR_BF16)
            dmode = BF if dmode == H3 else dmode
            wmode = BF if wmode == H3 else wmode
        want_dx = (ctx.needs_input_grad[0] and cout % _TC_CIN_MULTIPLE == 0 and dmode[0] >= 2
                   and _conv_algo(cout, cin_p, r, "dgrad")[0] == ALGO_UMMA)      # else: the planes of g would have no reader
        want_dw = ctx.needs_input_grad[1] and cin_p % _TC_CIN_MULTIPLE == 0 and cout % 4 == 0 and wmode[0] >= 2
        g_splits = {}                     # (nprod, fmt) -> (hi, lo) of g
        if act != ACT_NONE:
            g = torch.empty_like(dy)
            fused = dmode if want_dx else (wmode if want_dw else (0, 0))
            if tap is not None and fused[0] in (2, 3):
                dy, tap = _materialize_tap(dy, y, tap), None
            if tap is not None:          # image-facing VGG conv1_1: activation backward with the tap gradient folded in
                target, gl, _ = tap
                call("pvg_act_bwd_tap", dy.data_ptr(), y.data_ptr(), act, float(slope), g.data_ptr(), dy.shape[0],
                     dy.numel() // dy.shape[0], target.data_ptr(), gl.data_ptr(), _stream())
            elif fused[0] == 2:          # activation backward and the 16-bit planes of g in one pass
                planes = torch.empty((2 * dy.numel(),), dtype=_plane_dtype(fused[1]), device=dy.device)
                call("pvg_act_bwd_split_16", dy.data_ptr(), y.data_ptr(), act, float(slope), g.data_ptr(), planes.data_ptr(),
                     dy.numel(), fused[1], _stream())
                g_splits[fused] = (g, planes)
            elif fused[0] == 3:          # activation backward and the fp32 residual plane of g in one pass
                lo = torch.empty_like(dy)
                hi = None if tf32_truncates() else torch.empty_like(dy)
                call("pvg_act_bwd_split", dy.data_ptr(), y.data_ptr(), act, float(slope), g.data_ptr(), _p(hi), lo.data_ptr(),
                     dy.numel(), _stream())
                g_splits[(3, 0)] = (g if hi is None else hi, lo)
            else:
                call("pvg_act_bwd", dy.data_ptr(), y.data_ptr(), act, float(slope), g.data_ptr(), dy.numel(), _stream())
        else:
            g = dy
        dx = dw = db = None

        def split_g(mode):
            if mode not in g_splits:
                g_splits[mode] = _split(g, *mode)
            return g_splits[mode]

        if ctx.needs_input_grad[0]:
            algo, np_, fmt_ = _conv_algo(cout, cin_p, r, "dgrad")
            if algo == ALGO_UMMA and (np_, fmt_) == H3:
                np_, fmt_ = dmode
            packs = _get_packs(weight, cin_p, algo == ALGO_UMMA)
            # data gradient = "same" convolution of g with the tap-flipped, transposed pack [CinP][R][S][Cout]
            dx = _conv_forward(g, packs, 2, cin_p, r, None, ACT_NONE, 0.0, algo, np_, fmt_,
                               2.0 * n * h * w * cout * r * s * cin_log,
                               split=split_g((np_, fmt_)) if (algo == ALGO_UMMA and np_ >= 2) else None)
        if ctx.needs_input_grad[1]:
            nprod, wfmt = wmode
            d = ConvDesc(n, h, w, cin_p, cout, r, s, (r - 1) // 2, ACT_NONE, 0.0, ALGO_SIMT, nprod, wfmt)
            head7 = r == 7 and cout <= 3 and cin_p <= 32 and os.environ.get("PVG_NO_DIRECT") != "1"
            tc_wgrad = _precision != "fp32" and cin_p % _TC_CIN_MULTIPLE == 0 and not head7
            # the tensor-core path overwrites dw (split-K partials meet in its zero-pool scratch); the CUDA-core kernels add into it
            dw = (torch.empty_like if tc_wgrad else torch.zeros_like)(weight, memory_format=torch.contiguous_format)
            if tc_wgrad:
                # tensor-core weight gradient; dY needs a channel count that is a multiple of 4 (16-byte TMA strides):
                # the 3-channel image heads and the 65-channel encoder tail are zero-padded (a few MB)
                cout4 = (cout + 3) // 4 * 4
                if nprod == 2:
                    cout4 = (cout + 7) // 8 * 8          # bf16 planes: 16-byte strides need 8 channels
                if cout4 != cout:
                    g4 = empty_nhwc((n, cout4, h, w), dy.device)
                    g4[:, :cout].copy_(g)
                    g4[:, cout:].zero_()
                    dw4 = torch.empty((cout4, cin_log, r, s), dtype=torch.float32, device=dy.device)
                    d = ConvDesc(n, h, w, cin_p, cout4, r, s, (r - 1) // 2, ACT_NONE, 0.0, ALGO_SIMT, nprod, wfmt)
                    g_pair = _split(g4, nprod, wfmt) if nprod >= 2 else (g4, None)
                else:
                    g4, dw4 = g, dw
                    g_pair = split_g(wmode) if nprod >= 2 else (g, None)
                scratch = zero_pool.zeros((cout4 * r * s * _pad32(cin_p),), torch.float32, dy.device)
                if nprod == 2 and ctx.x_wplanes is not None and ctx.x_wfmt == wfmt:
                    x_hi, x_lo = x, ctx.x_wplanes                  # written by x's producer in the forward pass
                else:
                    x_hi, x_lo = _split(x, nprod, wfmt) if nprod >= 2 else (x, None)
                g_hi, g_lo = g_pair
                prof = wgrad_profile is not None
                if prof:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    profile_spin()
                    e0.record()
                call("pvg_conv2d_wgrad_umma", d, cin_log, x_hi.data_ptr(), _p(x_lo), g_hi.data_ptr(), _p(g_lo),
                     scratch.data_ptr(), dw4.data_ptr(), 0, _stream())
                if prof:
                    e1.record()
                    wgrad_profile.append((e0, e1, 2.0 * n * h * w * cout * r * s * cin_log))
                if cout4 != cout:
                    dw = dw4[:cout].contiguous()
            else:
                call("pvg_conv2d_wgrad", d, cin_log, x.data_ptr(), g.data_ptr(), dw.data_ptr(), _stream())
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.empty((cout,), dtype=torch.float32, device=dy.device)
            scratch = torch.empty((cout,), dtype=torch.float64, device=dy.device)
            call("pvg_channel_sum", g.data_ptr(), n * h * w, cout, scratch.data_ptr(), db.data_ptr(), _stream())
        return dx, dw, db, None, None, None, None, None, None


# A gradient tensor can carry the largest magnitude of its elements, written by the kernel that produced it (conv data gradient,
# BatchNorm backward) into a one-element device buffer: the consumer's power-of-two scale then needs no pass over the tensor.
# The tag is only trusted while the tensor has not been modified (autograd may accumulate a second gradient INTO the buffer in
# place: that bumps its version counter).
track_amax = os.environ.get("PVG_NO_AMAX_TAGS") != "1"


def tag_amax(t: Tensor, amax_bits: Tensor) -> Tensor:
    t._pvg_amax = (amax_bits, t._version)
    return t


def known_amax(t: Tensor) -> Optional[Tensor]:
    tag = getattr(t, "_pvg_amax", None)
    if tag is None or not track_amax or tag[1] != t._version:
        return None
    return tag[0]


def _backward_h3(ctx, dy, x, weight, y, tap=None):
    """Data and weight gradient as all-fp16 split products (conv_h3.cu with the flipped pack; conv_wgrad_umma.cu NPROD = 4):
    dY is scaled by a power of two chosen from max|dY| so that its fp16 plane pair is exact to 22 bits, both kernels undo the
    scale; x is consumed through the very planes the forward convolution read."""
    act, slope, has_bias, cin_log = ctx.meta
    cout_log, _, r, s = weight.shape
    cout = ctx.cphys                    # physical channels of dY (>= the weight's: zero rows in the packs)
    n, cin_p, h, w = x.shape
    dev = dy.device
    st = _stream()
    inv = torch.empty((1,), dtype=torch.float32, device=dev)
    planes = torch.empty((2 * dy.numel(),), dtype=torch.float16, device=dev)
    amax = known_amax(dy)                 # written by the kernel that produced dy (an upper bound of max|dy| is what is needed)
    if amax is None:
        amax = zero_pool.zeros((1,), torch.int32, dev)
        call("pvg_amax", dy.data_ptr(), dy.numel(), amax.data_ptr(), st)
    need_g = has_bias and ctx.needs_input_grad[2]
    if act != ACT_NONE and tap is not None:
        target, gl, _ = tap                # dy + the feature-matching L1 gradient, activation backward and planes in one pass
        g = torch.empty_like(dy) if need_g else None
        call("pvg_act_bwd_tap_split_16_scaled", dy.data_ptr(), y.data_ptr(), act, float(slope), _p(g), planes.data_ptr(),
             dy.shape[0], dy.numel() // dy.shape[0], amax.data_ptr(), inv.data_ptr(), target.data_ptr(), gl.data_ptr(), st)
    elif act != ACT_NONE:
        g = torch.empty_like(dy) if need_g else None
        call("pvg_act_bwd_split_16_scaled", dy.data_ptr(), y.data_ptr(), act, float(slope), _p(g), planes.data_ptr(), dy.numel(),
             amax.data_ptr(), inv.data_ptr(), st)
    else:
        g = dy
        call("pvg_split_16_scaled", dy.data_ptr(), planes.data_ptr(), dy.numel(), amax.data_ptr(), inv.data_ptr(), st)
    dx = dw = db = None
    packs = _get_packs(weight, cin_p, True, cout)
    if ctx.needs_input_grad[0]:
        dx = empty_nhwc((n, cin_p, h, w), dev)
        d = ConvDesc(n, h, w, cout, cin_p, r, r, (r - 1) // 2, ACT_NONE, 0.0, ALGO_UMMA, 2, _lib.CORR_FP16_ALL)
        prof = conv_profile is not None
        if prof:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            profile_spin()
            e0.record()
        dx_amax = zero_pool.zeros((1,), torch.int32, dev) if track_amax else None
        call("pvg_conv2d_fwd_planes", d, planes.data_ptr(), packs.lo(2, 2, _lib.CORR_FP16_ALL).data_ptr(), None, dx.data_ptr(), None,
             inv.data_ptr(), None, 0, _p(dx_amax), st)
        if dx_amax is not None:
            tag_amax(dx, dx_amax)
        if prof:
            e1.record()
            conv_profile.append((e0, e1, 2.0 * n * h * w * cout_log * r * s * cin_log, "h3"))
            profile_shapes[id(e0)] = ("dgrad", n, h, w, cout, cin_p, r, False, False)
    if ctx.needs_input_grad[1]:
        # 16 / 32 channels on both sides: tiny GEMM, K = every pixel - fp32 CUDA-core kernel on x and dY themselves
        # (measured, r02 layer table: 16 -> 16 over 128 frames 929 -> 445 us, 16 -> 32 951 -> 631 us; with 32 input channels the
        #  tensor-core kernel is the faster one, so those stay there)
        use_small = small_wgrad_kernel and act == ACT_NONE and r == 3 and cin_p == 16 and cout in (16, 32)
        xp = None
        if not use_small:
            xp = ctx.x_wplanes if ctx.x_wplanes is not None else _split(x, 2, _lib.CORR_FP16_ALL)[1]
        deferred = wgrad_defer.active
        if deferred:
            dw, scratch = None, wgrad_defer.scratch_for(weight, cin_p, cout, r, s, dev)
        else:
            dw = torch.empty((cout, cin_log, r, s), dtype=torch.float32, device=dev)
            scratch = zero_pool.zeros((cout * r * s * _pad32(cin_p),), torch.float32, dev)
        d = ConvDesc(n, h, w, cin_p, cout, r, s, (r - 1) // 2, ACT_NONE, 0.0, ALGO_UMMA, 2, _lib.CORR_FP16_ALL)
        prof = wgrad_profile is not None
        if prof:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            profile_spin()
            e0.record()
        if use_small:
            call("pvg_conv2d_wgrad_small", d, cin_log, x.data_ptr(), dy.data_ptr(), scratch.data_ptr(), _p(dw), 0, st)
        else:
            call("pvg_conv2d_wgrad_planes", d, cin_log, xp.data_ptr(), planes.data_ptr(), inv.data_ptr(), scratch.data_ptr(),
                 _p(dw), 0, st)
        if prof:
            e1.record()
            wgrad_profile.append((e0, e1, 2.0 * n * h * w * cout_log * r * s * cin_log))
            profile_shapes[id(e0)] = ("wgrad", n, h, w, cin_p, cout, r, False, False)
        if dw is not None and cout != cout_log:
            dw = dw[:cout_log]
    if need_g:
        db = torch.empty((cout,), dtype=torch.float32, device=dev)
        scr = torch.empty((cout,), dtype=torch.float64, device=dev)
        call("pvg_channel_sum", g.data_ptr(), n * h * w, cout, scr.data_ptr(), db.data_ptr(), st)
    return dx, dw, db, None, None, None, None, None, None


Conv2dFn._backward_h3 = staticmethod(_backward_h3)


def supports_padded_cout() -> bool:
    """True when conv2d(..., cout_phys=) is available: every conv role on the all-fp16 tensor-core kernels."""
    H3 = (2, _lib.CORR_FP16_ALL)
    return (_precision == "tf32x3" and os.environ.get("PVG_NO_COUT_PAD") != "1" and _mode("fwd") == H3 and _mode("dgrad") == H3
            and _mode("wgrad") == H3)


def conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, act: int = ACT_NONE, slope: float = 0.0,
           out_planes: bool = False, cout_phys: Optional[int] = None, bn_stats_groups: int = 0) -> Tensor:
    """``out_planes``: the consumer of the result is another tensor-core convolution - have the epilogue write its operand planes.
    ``cout_phys``: write the result physically padded (zero channels) to this many channels - a multiple of 8 - so that a layer
    with an odd channel count (the 65-channel encoder tail, representation_network.py:28) and its consumers stay on the
    tensor cores.
    ``bn_stats_groups`` > 0: the result goes straight into a training-mode BatchNorm with that many batch groups - the conv
    epilogue accumulates its per-channel statistics (no separate statistics pass over the result)."""
    if os.environ.get("PVG_NO_EPILOGUE_STATS") == "1":
        bn_stats_groups = 0
    want = bool(out_planes) and _lib.CORR_FP16_ALL in conv_input_planes(False)
    y, yp, sums = Conv2dFn.apply(x, weight, bias, act, slope, planes_of(x), want, cout_phys, int(bn_stats_groups))
    if yp is not None:
        y._pvg_planes = {_lib.CORR_FP16_ALL: yp}
    if sums is not None:
        y._pvg_bn_sums = (sums, int(bn_stats_groups))          # picked up by pool_bn_act (training mode, no pooling)
    return y


# ---------------------------------------------------------------------------------------------------------------
# [avg_pool2d(2)] -> BatchNorm2d -> (+ residual) -> activation
# ---------------------------------------------------------------------------------------------------------------
class PoolBNActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, training, pool, act, slope, eps, momentum,
                groups, plane_fmts=(), pre_sums=None):
        ctx.set_materialize_grads(False)
        x_in = x
        x = nhwc(x)
        if x is not x_in:
            pre_sums = None
        n, c, h, w = x.shape
        dev = x.device
        st = _stream()
        if residual is not None:
            residual = nhwc(residual)
        oh, ow = (h // 2, w // 2) if pool else (h, w)
        cp = c if running_mean is None else int(running_mean.shape[0])     # channels that have parameters (<= c: zero padding)
        if weight is not None:
            cp = int(weight.shape[0])
        padded = cp != c
        mean = (torch.zeros if padded else torch.empty)((groups, c), dtype=torch.float32, device=dev)
        invstd = (torch.zeros if padded else torch.empty)((groups, c), dtype=torch.float32, device=dev)
        have_stats = pre_sums is not None and training and not pool and tuple(pre_sums.shape) == (groups, 2, c)
        sums = pre_sums if have_stats else (zero_pool.zeros((groups, 2, c), torch.float64, dev) if (training or pool) else None)
        if pool:
            xp = empty_nhwc((n, c, oh, ow), dev)
            call("pvg_pool2_stats", x.data_ptr(), n, h, w, c, xp.data_ptr(), groups, sums.data_ptr(), st)
        else:
            xp = x
            if training and not have_stats:           # else: accumulated by the producing convolution's epilogue
                call("pvg_bn_stats", x.data_ptr(), n, h * w, c, groups, sums.data_ptr(), st)
        y = empty_nhwc((n, c, oh, ow), dev)
        pa, pb = _alloc_planes(plane_fmts, y.numel(), c, dev)
        fa = plane_fmts[0] if len(plane_fmts) > 0 else 0
        fb = plane_fmts[1] if len(plane_fmts) > 1 else 0
        wd = weight.detach() if weight is not None else None
        bd = bias.detach() if bias is not None else None
        if training and groups * c * 8 <= 48 * 1024:      # statistics finalisation fused into the apply pass
            call("pvg_bn_finalize_apply_ex", xp.data_ptr(), n, oh * ow, c, groups, sums.data_ptr(), (n // groups) * oh * ow,
                 float(eps), float(momentum), _p(running_mean), _p(running_var), mean.data_ptr(), invstd.data_ptr(), _p(wd),
                 _p(bd), _p(residual), act, float(slope), y.data_ptr(), _p(pa), fa, _p(pb), fb, cp, st)
        else:
            if padded and training:
                raise _lib.PvgError("channel-padded BatchNorm needs the fused finalize + apply kernel")
            if training:
                call("pvg_bn_finalize", sums.data_ptr(), (n // groups) * oh * ow, groups, c, float(eps), float(momentum),
                     _p(running_mean), _p(running_var), mean.data_ptr(), invstd.data_ptr(), st)
            else:
                if groups != 1:
                    raise _lib.PvgError("grouped statistics only exist in training mode")
                call("pvg_bn_eval_prepare", running_mean.data_ptr(), running_var.data_ptr(), cp, float(eps), mean.data_ptr(),
                     invstd.data_ptr(), st)
            call("pvg_bn_apply_ex", xp.data_ptr(), n, oh * ow, c, groups, mean.data_ptr(), invstd.data_ptr(), _p(wd), _p(bd),
                 _p(residual), act, float(slope), y.data_ptr(), _p(pa), fa, _p(pb), fb, cp, st)
        ctx.save_for_backward(xp, weight, mean, invstd, y if act != ACT_NONE else None)
        ctx.meta = (training, pool, act, slope, groups, residual is not None, (n, c, h, w))
        ctx.cp = cp
        for t in (pa, pb):
            if t is not None:
                ctx.mark_non_differentiable(t)
        return y, pa, pb

    @staticmethod
    def backward(ctx, dy, _dpa=None, _dpb=None):
        if dy is None:
            return (None,) * 15
        xp, weight, mean, invstd, y = ctx.saved_tensors
        training, pool, act, slope, groups, has_res, (n, c, h, w) = ctx.meta
        dy = nhwc(dy)
        dev = dy.device
        st = _stream()
        oh, ow = (h // 2, w // 2) if pool else (h, w)
        sums2 = zero_pool.zeros((groups, 2, c), torch.float64, dev)
        need_params = weight is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        if training or need_params:
            call("pvg_bn_bwd_reduce", dy.data_ptr(), _p(y), xp.data_ptr(), n, oh * ow, c, groups, mean.data_ptr(),
                 invstd.data_ptr(), act, float(slope), sums2.data_ptr(), st)
        dx = empty_nhwc((n, c, h, w), dev) if ctx.needs_input_grad[0] else None
        g_out = empty_nhwc((n, c, oh, ow), dev) if (has_res and ctx.needs_input_grad[3]) else None
        dweight = dbias = None
        if need_params:
            dweight = torch.empty((ctx.cp,), dtype=torch.float32, device=dev)
            dbias = torch.empty((ctx.cp,), dtype=torch.float32, device=dev)
        if dx is not None or g_out is not None:
            if dx is None:        # only the residual gradient is wanted
                dx_buf = empty_nhwc((n, c, h, w), dev)
            else:
                dx_buf = dx
            dx_amax = zero_pool.zeros((1,), torch.int32, dev) if (track_amax and dx is not None) else None
            call("pvg_bn_bwd_apply_ex", dy.data_ptr(), _p(y), xp.data_ptr(), n, h, w, c, groups, mean.data_ptr(),
                 invstd.data_ptr(), _p(weight.detach() if weight is not None else None), act, float(slope),
                 sums2.data_ptr(), 0 if training else 1, 1 if pool else 0, dx_buf.data_ptr(), _p(g_out), _p(dweight), _p(dbias),
                 ctx.cp, _p(dx_amax), st)          # dweight / dbias: the parameter gradients ride along in the same launch
            if dx_amax is not None:
                tag_amax(dx, dx_amax)
        elif need_params:
            if ctx.cp != c:
                raise _lib.PvgError("channel-padded BatchNorm backward needs the input gradient path")
            call("pvg_bn_bwd_params", sums2.data_ptr(), groups, c, dweight.data_ptr(), dbias.data_ptr(), st)
        return (dx, dweight, dbias, g_out) + (None,) * 11


def pool_bn_act(x, bn, residual=None, pool=False, act=ACT_NONE, slope=0.2, groups=1, planes: Sequence[int] = ()):
    """``bn`` is an nn.BatchNorm2d used purely as the parameter/buffer container (reference state_dict names).
    ``planes``: plane formats (``conv_input_planes()``) the apply pass writes next to y for the convolution that consumes it."""
    training = bn.training
    if not training:
        groups = 1                                          # running statistics: the chunks of a stacked batch are not told apart
    if training and bn.num_batches_tracked is not None:
        defer_count(bn.num_batches_tracked, groups)         # applied by flush_deferred() at the end of Model.forward
    planes = tuple(planes)[:2]
    pre = getattr(x, "_pvg_bn_sums", None)
    pre_sums = pre[0] if (pre is not None and pre[1] == groups and training and not pool) else None
    y, pa, pb = PoolBNActFn.apply(x, bn.weight, bias_or_none(bn), residual, bn.running_mean, bn.running_var, training, pool, act,
                                  slope, bn.eps, bn.momentum if bn.momentum is not None else 0.1, groups, planes, pre_sums)
    return _attach_planes(y, planes, (pa, pb))


def bias_or_none(m):
    return getattr(m, "bias", None)


# ---------------------------------------------------------------------------------------------------------------
# resampling
# ---------------------------------------------------------------------------------------------------------------
class Upsample2xFn(torch.autograd.Function):
    """F.interpolate(scale_factor=2, mode='bilinear', align_corners=False), model/layers/up_block.py:35,43."""

    @staticmethod
    def forward(ctx, x, plane_fmts=()):
        ctx.set_materialize_grads(False)
        x = nhwc(x)
        n, c, h, w = x.shape
        y = empty_nhwc((n, c, 2 * h, 2 * w), x.device)
        pa, pb = _alloc_planes(plane_fmts, y.numel(), c, x.device)
        call("pvg_resize_bilinear_ex", x.data_ptr(), n, h, w, c, y.data_ptr(), 2 * h, 2 * w, _p(pa),
             plane_fmts[0] if len(plane_fmts) > 0 else 0, _p(pb), plane_fmts[1] if len(plane_fmts) > 1 else 0, _stream())
        ctx.shape = (n, c, h, w)
        for t in (pa, pb):
            if t is not None:
                ctx.mark_non_differentiable(t)
        return y, pa, pb

    @staticmethod
    def backward(ctx, dy, _dpa=None, _dpb=None):
        if dy is None:
            return None, None
        n, c, h, w = ctx.shape
        dy = nhwc(dy)
        dx = empty_nhwc((n, c, h, w), dy.device)
        call("pvg_upsample2x_bwd", dy.data_ptr(), n, h, w, c, dx.data_ptr(), _stream())
        return dx, None


def upsample2x(x: Tensor, planes: Sequence[int] = ()) -> Tensor:
    planes = tuple(planes)[:2]
    y, pa, pb = Upsample2xFn.apply(x, planes)
    return _attach_planes(y, planes, (pa, pb))


def resize_bilinear(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """F.interpolate(x, size, mode='bilinear') of a tensor that needs no gradient (ground truth, losses.py:92,450)."""
    x = nhwc(x.detach())
    n, c, h, w = x.shape
    y = empty_nhwc((n, c, size[0], size[1]), x.device)
    call("pvg_resize_bilinear", x.data_ptr(), n, h, w, c, y.data_ptr(), size[0], size[1], _stream())
    return y


class MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) of torchvision VGG19 (model/layers/vgg.py:16)."""

    @staticmethod
    def forward(ctx, x, plane_fmts=()):
        ctx.set_materialize_grads(False)
        x = nhwc(x)
        n, c, h, w = x.shape
        y = empty_nhwc((n, c, h // 2, w // 2), x.device)
        pa, _ = _alloc_planes(plane_fmts[:1], y.numel(), c, x.device)
        call("pvg_maxpool2_fwd_ex", x.data_ptr(), n, h, w, c, y.data_ptr(), _p(pa), plane_fmts[0] if len(plane_fmts) > 0 else 0,
             _stream())
        ctx.save_for_backward(x, y)
        if pa is not None:
            ctx.mark_non_differentiable(pa)
        return y, pa

    @staticmethod
    def backward(ctx, dy, _dpa=None):
        if dy is None:
            return None, None
        x, y = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = nhwc(dy)
        dx = empty_nhwc((n, c, h, w), dy.device)
        if (h % 2) or (w % 2):
            dx.zero_()
        call("pvg_maxpool2_bwd", dy.data_ptr(), x.data_ptr(), y.data_ptr(), n, h, w, c, 0, dx.data_ptr(), _stream())
        known = known_amax(dy)
        if known is not None:              # every element of dx is an element of dy or zero
            tag_amax(dx, known)
        return dx, None


def maxpool2(x: Tensor, planes: Sequence[int] = ()) -> Tensor:
    planes = tuple(planes)[:1]
    y, pa = MaxPool2Fn.apply(x, planes)
    return _attach_planes(y, planes, (pa,))


# ---------------------------------------------------------------------------------------------------------------
# ConvLSTM cell (point-wise part) and the channel concat that feeds the gate convolution
# ---------------------------------------------------------------------------------------------------------------
class LSTMCellFn(torch.autograd.Function):
    """convolutional_lstm_cell.py:92-101 given the 4C gate pre-activations (order input, forget, output, cell)."""

    @staticmethod
    def forward(ctx, gates, c_prev):
        gates, c_prev = nhwc(gates), nhwc(c_prev)
        n, c, h, w = c_prev.shape
        c_new, h_new = empty_nhwc((n, c, h, w), gates.device), empty_nhwc((n, c, h, w), gates.device)
        call("pvg_lstm_fwd", gates.data_ptr(), c_prev.data_ptr(), n * h * w, c, c_new.data_ptr(), h_new.data_ptr(), _stream())
        ctx.save_for_backward(gates, c_prev, c_new)
        return h_new, c_new

    @staticmethod
    def backward(ctx, dh, dc):
        gates, c_prev, c_new = ctx.saved_tensors
        n, c, h, w = c_prev.shape
        dh = nhwc(dh) if dh is not None else None
        dc = nhwc(dc) if dc is not None else None
        dgates, dc_prev = torch.empty_like(gates), torch.empty_like(c_prev)
        call("pvg_lstm_bwd", gates.data_ptr(), c_prev.data_ptr(), c_new.data_ptr(), _p(dh), _p(dc), n * h * w, c,
             dgates.data_ptr(), dc_prev.data_ptr(), _stream())
        return dgates, dc_prev


def lstm_cell(gates: Tensor, c_prev: Tensor) -> Tuple[Tensor, Tensor]:
    return LSTMCellFn.apply(gates, c_prev)


class _WgradDefer:
    """One weight-gradient scratch per weight and optimiser step.  A weight of the time loop is used T - 1 times per step;
    autograd would run, per use, a split-K weight-gradient launch, an unpack to OIHW and an accumulation kernel (1 765
    accumulations per BAIR-256 step, round 1).  Between ``begin()`` and ``flush()`` the all-fp16 weight-gradient kernel of every
    use adds its partial sums (already multiplied by that use's 1 / S) into ONE packed scratch; ``flush()`` unpacks each scratch
    once, straight into ``weight.grad`` (leaf parameters: the flat gradient arena) or through ``torch.autograd.backward`` for
    derived weights (the interleaved ConvLSTM gate weight).  Same sum as autograd's (trainer.py:584-587), fewer launches."""

    def __init__(self):
        self.active = False
        self.entries = {}

    def begin(self):
        self.entries = {}
        self.active = os.environ.get("PVG_NO_WGRAD_DEFER") != "1"

    def scratch_for(self, weight, cin_p, cout_phys, r, s, device):
        key = (weight.data_ptr(), tuple(weight.shape), cin_p, cout_phys)
        e = self.entries.get(key)
        if e is None:
            e = dict(weight=weight, cin_p=cin_p, cout=cout_phys, r=r, s=s,
                     scratch=zero_pool.zeros((cout_phys * r * s * _pad32(cin_p),), torch.float32, device))
            self.entries[key] = e
        return e["scratch"]

    def flush(self, on_leaf_grad=None, resolve=None):
        """Unpacks every scratch into its weight's gradient; ``on_leaf_grad(param)`` is called for each leaf parameter that
        received one (what a post-accumulate hook would have seen); ``resolve(data_ptr)`` maps a weight's storage address to the
        owning parameter object (the flat arena knows it), so the gradient lands in THE parameter's ``.grad`` whatever tensor
        object autograd handed to the backward function."""
        self.active = False
        entries, self.entries = self.entries, {}
        derived = []
        for e in entries.values():
            w = e["weight"]
            if resolve is not None and w.is_leaf:
                owner = resolve(w.data_ptr())
                if owner is not None and tuple(owner.shape) == tuple(w.shape):
                    w = owner
            cout_log, cin_log, r, s = w.shape
            st = _stream()
            if w.is_leaf and e["cout"] == cout_log and w.grad is not None and w.grad.is_contiguous():
                call("pvg_unpack_dw", e["scratch"].data_ptr(), cout_log, cin_log, r, s, e["cin_p"], w.grad.data_ptr(), 1, st)
                if on_leaf_grad is not None:
                    on_leaf_grad(w)
                continue
            dw = torch.empty((e["cout"], cin_log, r, s), dtype=torch.float32, device=w.device)
            call("pvg_unpack_dw", e["scratch"].data_ptr(), e["cout"], cin_log, r, s, e["cin_p"], dw.data_ptr(), 0, st)
            dw = dw[:cout_log]
            if w.is_leaf:
                if w.grad is None:
                    w.grad = dw.contiguous()
                else:
                    w.grad.add_(dw)
                if on_leaf_grad is not None:
                    on_leaf_grad(w)
            else:
                derived.append((w, dw))
        if derived:         # e.g. the interleaved gate weight of a ConvLSTM: its stack / reshape graph has not been walked yet
            torch.autograd.backward([w for w, _ in derived], [g for _, g in derived])


wgrad_defer = _WgradDefer()


class _CtxShim:
    """The fields ``_backward_h3`` reads from a Conv2dFn context."""

    def __init__(self, meta, cphys, x_wplanes, needs_input_grad):
        self.meta, self.cphys, self.x_wplanes, self.needs_input_grad = meta, cphys, x_wplanes, needs_input_grad


class ConvLSTMStepFn(torch.autograd.Function):
    """convolutional_lstm_cell.py:88-101 as ONE launch: the gate convolution over the padded concat z = [inputs..., h] with the
    cell update fused into its epilogue (pvg_convlstm_step).  ``w_il`` / ``b_il``: the four gate weights / biases INTERLEAVED
    along the output channels (row 4c + gate).  Backward: pvg_lstm_bwd_act, then the all-fp16 data / weight gradient kernels."""

    @staticmethod
    def forward(ctx, z, w_il, b_il, c_prev, z_planes):
        z_in = z
        z, c_prev = nhwc(z), nhwc(c_prev)
        if z is not z_in:
            z_planes = {}
        fmt = _lib.CORR_FP16_ALL
        n, cin_p, h, w = z.shape
        cout, cin_log, r, _ = w_il.shape
        c = cout // 4
        packs = _get_packs(w_il, cin_p, True)
        zp = z_planes[fmt] if fmt in z_planes else _split(z, 2, fmt)[1]
        need_grad = any(ctx.needs_input_grad[:4])
        gates = empty_nhwc((n, cout, h, w), z.device) if need_grad else None
        c_new, h_new = empty_nhwc((n, c, h, w), z.device), empty_nhwc((n, c, h, w), z.device)
        d = ConvDesc(n, h, w, cin_p, cout, r, r, (r - 1) // 2, ACT_NONE, 0.0, ALGO_UMMA, 2, fmt)
        prof = conv_profile is not None
        if prof:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            profile_spin()
            e0.record()
        call("pvg_convlstm_step", d, zp.data_ptr(), packs.lo(0, 2, fmt).data_ptr(), b_il.detach().contiguous().data_ptr(),
             c_prev.data_ptr(), c_new.data_ptr(), h_new.data_ptr(), _p(gates), _stream())
        if prof:
            e1.record()
            conv_profile.append((e0, e1, 2.0 * n * h * w * cout * r * r * cin_log, "h3"))
            profile_shapes[id(e0)] = ("lstm", n, h, w, z.shape[1], cout, r, False, False)
        ctx.save_for_backward(z, w_il, gates, c_prev, c_new)
        ctx.zp = zp if w_il.requires_grad else None
        return h_new, c_new

    @staticmethod
    def backward(ctx, dh, dc):
        z, w_il, gates, c_prev, c_new = ctx.saved_tensors
        n, c, h, w = c_prev.shape
        dh = nhwc(dh) if dh is not None else None
        dc = nhwc(dc) if dc is not None else None
        dgates, dc_prev = torch.empty_like(gates), torch.empty_like(c_prev)
        call("pvg_lstm_bwd_act", gates.data_ptr(), c_prev.data_ptr(), c_new.data_ptr(), _p(dh), _p(dc), n * h * w, c,
             dgates.data_ptr(), dc_prev.data_ptr(), _stream())
        shim = _CtxShim((ACT_NONE, 0.0, True, w_il.shape[1]), w_il.shape[0], ctx.zp, tuple(ctx.needs_input_grad[:3]))
        dz, dw, db = _backward_h3(shim, dgates, z, w_il, None)[:3]
        return dz, dw, db, dc_prev, None


def supports_fused_lstm() -> bool:
    return supports_padded_cout() and os.environ.get("PVG_NO_FUSED_LSTM") != "1"


def convlstm_step(z: Tensor, w_il: Tensor, b_il: Tensor, c_prev: Tensor) -> Tuple[Tensor, Tensor]:
    """(h_new, c_new) of one ConvLSTM cell step; z = padded concat [inputs..., h] (its planes are picked up when attached)."""
    return ConvLSTMStepFn.apply(z, w_il, b_il, c_prev, planes_of(z))


class ConcatPadFn(torch.autograd.Function):
    """Channel concat of 4-D maps and 2-D (N, C) vectors broadcast over H x W (conv_dynamics_network.py:64-109,
    convolutional_lstm_cell.py:35-75), zero-padded to ``c_pad`` physical channels so the result is a legal
    tensor-core A operand (K chunks of 32 channels)."""

    @staticmethod
    def forward(ctx, c_pad, plane_fmts, *parts):
        ctx.set_materialize_grads(False)
        ref = next(p for p in parts if p.dim() == 4)
        n, _, h, w = ref.shape
        out = empty_nhwc((n, c_pad, h, w), ref.device)
        if len(parts) > _lib.ConcatDesc.MAX_PARTS:
            raise _lib.PvgError("too many parts for one concat")
        d = _lib.ConcatDesc()
        d.N, d.H, d.W, d.Cpad, d.nparts = n, h, w, c_pad, len(parts)
        keep, spans, off = [], [], 0
        for k, p in enumerate(parts):
            _check_cuda(p)
            c = p.shape[1]
            q = p.detach()
            if p.dim() == 4:
                # a map is consumed in place when every sample is NHWC-dense (the batch stride is free: time slices of (B, T, ...))
                if not (c == 1 or q.stride()[1] == 1) or q.stride()[2] != w * c or q.stride()[3] != c:
                    q = nhwc(q)
            elif q.stride(1) != 1:
                q = q.contiguous()
            keep.append(q)
            d.c[k], d.is_vec[k], d.bstride[k], d.src[k] = c, 0 if p.dim() == 4 else 1, q.stride(0), q.data_ptr()
            spans.append((off, c, p.dim()))
            off += c
        if off > c_pad:
            raise _lib.PvgError("concat wider than its padded size")
        pa, pb = _alloc_planes(plane_fmts, out.numel(), c_pad, ref.device)
        call("pvg_concat_pad", d, out.data_ptr(), _p(pa), plane_fmts[0] if len(plane_fmts) > 0 else 0, _p(pb),
             plane_fmts[1] if len(plane_fmts) > 1 else 0, _stream())
        ctx.spans = spans
        for t in (pa, pb):
            if t is not None:
                ctx.mark_non_differentiable(t)
        return out, pa, pb

    @staticmethod
    def backward(ctx, dout, _dpa=None, _dpb=None):
        if dout is None:
            return (None,) * (2 + len(ctx.spans))
        grads = []
        for i, (off, c, dim) in enumerate(ctx.spans):
            if not ctx.needs_input_grad[i + 2]:
                grads.append(None)
            elif dim == 4:
                grads.append(dout[:, off:off + c])
            else:
                grads.append(dout[:, off:off + c].sum(dim=(2, 3)))
        return (None, None) + tuple(grads)


def concat_pad(parts: Sequence[Tensor], multiple: int = 32, planes: Sequence[int] = ()) -> Tensor:
    total = sum(p.shape[1] for p in parts)
    c_pad = (total + multiple - 1) // multiple * multiple
    planes = tuple(planes)[:2]
    y, pa, pb = ConcatPadFn.apply(c_pad, planes, *parts)
    return _attach_planes(y, planes, (pa, pb))


# ---------------------------------------------------------------------------------------------------------------
# input pipeline
# ---------------------------------------------------------------------------------------------------------------
def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's ``precompute_coeffs`` (box = the whole axis, BILINEAR: triangle filter, support 1, widened by the scale when
    shrinking) followed by ``normalize_coeffs_8bpc`` (22-bit fixed point), in the same double-precision operation order
    (src/libImaging/Resample.c): (bounds int32 [out_size, 2] = (first, count), coefficients int32 [out_size, ksize], ksize)."""
    import math
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = [[0, 0] for _ in range(out_size)]
    kk = [[0] * ksize for _ in range(out_size)]
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)                # C's (int): truncation toward zero, like Python's int()
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            wv = 1.0 - a if a < 1.0 else 0.0
            ws.append(wv)
            ww += wv
        for x in range(xmax):
            v = ws[x] / ww if ww != 0.0 else ws[x]
            kk[xx][x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = [xmin, xmax]
    return bounds, kk, ksize


_resample_tables = {}


def _resample_table(in_size: int, out_size: int, device):
    key = (in_size, out_size, str(device))
    hit = _resample_tables.get(key)
    if hit is None:
        b, k, ks = pil_bilinear_coeffs(in_size, out_size)
        hit = (torch.tensor(b, dtype=torch.int32, device=device), torch.tensor(k, dtype=torch.int32, device=device), ks)
        _resample_tables[key] = hit
    return hit


def resize_frames_uint8(frames: Tensor, crop: Optional[Sequence[int]], size: Sequence[int]) -> Tensor:
    """PIL ``crop`` + ``resize(size, Image.BILINEAR)`` of dataset/transforms.py:15-32 on uint8 frames on the device, bit-identical
    to Pillow (antialiased two-pass resampling in 22-bit fixed point: horizontal pass, then vertical).  frames: (N, Hs, Ws, 3)
    uint8 CUDA tensor; crop = [left, upper, right, lower] or None; size = (width, height) as in ``target_input_size``."""
    if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise _lib.PvgError("resize_frames_uint8 takes a (N, H, W, 3) uint8 CUDA tensor")
    frames = frames.contiguous()
    n, hs, ws, _ = frames.shape
    left, top, right, bottom = (0, 0, ws, hs) if crop is None else (int(v) for v in crop)
    hin, win = bottom - top, right - left
    ow, oh = int(size[0]), int(size[1])
    cur, cur_h, cur_w, box = frames, hs, ws, (left, top, hin, win)
    if ow != win:                                          # horizontal pass first (ImagingResample)
        b, k, ks = _resample_table(win, ow, frames.device)
        tmp = torch.empty((n, hin, ow, 3), dtype=torch.uint8, device=frames.device)
        call("pvg_resample_u8", cur.data_ptr(), n, cur_h, cur_w, box[0], box[1], hin, win, 0, ow, b.data_ptr(), k.data_ptr(), ks,
             tmp.data_ptr(), _stream())
        cur, cur_h, cur_w, box = tmp, hin, ow, (0, 0, hin, ow)
    if oh != hin:
        b, k, ks = _resample_table(hin, oh, frames.device)
        out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=frames.device)
        call("pvg_resample_u8", cur.data_ptr(), n, cur_h, cur_w, box[0], box[1], hin, ow, 1, oh, b.data_ptr(), k.data_ptr(), ks,
             out.data_ptr(), _stream())
        cur, cur_h, cur_w, box = out, oh, ow, (0, 0, oh, ow)
    if cur is frames:                                      # nothing to resample: the crop itself
        cur = frames[:, top:bottom, left:right].contiguous()
    return cur


def frames_from_uint8(frames: Tensor, crop: Optional[Sequence[int]] = None, mean: float = 0.5, std: float = 0.5,
                      size: Optional[Sequence[int]] = None) -> Tensor:
    """PIL ``crop`` (+ ``resize(size, BILINEAR)``) + ``ToTensor`` + ``Normalize(mean, std)`` of dataset/transforms.py:15-32,90-108
    on the device.  frames: (N, Hs, Ws, 3) uint8 CUDA tensor (decoded RGB frames); crop = [left, upper, right, lower] as in the
    YAML (``data.crop``); size = (width, height) = ``target_input_size`` (None: the crop already has it, as in every shipped
    config).  Returns the (N, 3, H, W) fp32 observation tensor (channels_last storage, what the kernels consume), bit-identical
    to the reference's CPU transform."""
    if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise _lib.PvgError("frames_from_uint8 takes a (N, H, W, 3) uint8 CUDA tensor")
    frames = frames.contiguous()
    n, hs, ws, _ = frames.shape
    left, top, right, bottom = (0, 0, ws, hs) if crop is None else (int(v) for v in crop)
    h, w = bottom - top, right - left
    if size is not None and (int(size[0]), int(size[1])) != (w, h):
        frames = resize_frames_uint8(frames, crop, size)
        n, hs, ws, _ = frames.shape
        left, top, h, w = 0, 0, hs, ws
    out = empty_nhwc((n, 3, h, w), frames.device)
    call("pvg_frames_u8_to_nhwc", frames.data_ptr(), n, hs, ws, left, top, h, w, float(mean), float(std), out.data_ptr(), _stream())
    return out


# ---------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------
class AbsDiffMeanFn(torch.autograd.Function):
    """out[n] = mean |a[n] - b[n]| over everything but the first dim; gradient flows to ``b`` only
    (a = detached ground-truth branch, losses.py:465 / nn.L1Loss at :118)."""

    @staticmethod
    def forward(ctx, a, b):
        _check_cuda(a); _check_cuda(b)
        if a.shape != b.shape:
            raise _lib.PvgError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
        if a.dim() == 4:
            a, b = nhwc(a), nhwc(b)           # same physical order on both sides; the mean is order independent
        else:
            a, b = a.contiguous(), b.contiguous()
        n = a.shape[0]
        count = a.numel() // n
        out = zero_pool.zeros((n,), torch.float64, a.device)
        call("pvg_absdiff_mean_fwd", a.data_ptr(), b.data_ptr(), n, count, out.data_ptr(), _stream())
        ctx.save_for_backward(a, b)
        return out.float()

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        n = a.shape[0]
        db = torch.empty_like(b)
        g = gout.contiguous().float()
        call("pvg_absdiff_mean_bwd", a.data_ptr(), b.data_ptr(), g.data_ptr(), n, a.numel() // n, db.data_ptr(), _stream())
        return None, db


def absdiff_mean(a: Tensor, b: Tensor) -> Tensor:
    return AbsDiffMeanFn.apply(a.detach(), b)


# A feature-matching L1 term on the output of a convolution that ALSO feeds deeper layers (the VGG taps of the perceptual loss,
# model/layers/vgg.py:41-56 + training/losses.py:450-465) gives that output two gradients: autograd would materialise the L1
# one (pvg_absdiff_mean_bwd) and add it to the deeper one - six passes over the largest feature maps of the step.  TapL1Fn
# sits IN the chain instead (identity for the feature, the loss as a second output); in backward it hands the deeper gradient on
# unchanged with the pending term attached, and the producing convolution's backward folds it into its activation-backward
# pass (pvg_act_bwd_tap*).  Anything else that might receive the tagged tensor never does: tap_l1 only defers for a feature that
# comes straight out of Conv2dFn, and Conv2dFn.backward materialises the sum itself whenever it cannot fuse it.
fuse_taps = os.environ.get("PVG_NO_TAP_FUSION") != "1"


def _is_nhwc_dense(t: Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def pending_tap(dy: Optional[Tensor]):
    tag = getattr(dy, "_pvg_tap", None) if dy is not None else None
    if tag is None:
        return None
    target, gl, version, feature = tag
    if version != dy._version:
        raise _lib.PvgError("a gradient carrying a deferred feature-matching term was modified before its convolution saw it")
    del dy._pvg_tap
    return target, gl, feature


def _materialize_tap(dy: Tensor, y: Optional[Tensor], tap) -> Tensor:
    target, gl, feature = tap
    y = feature if y is None else y           # a convolution without activation does not save its output: the tap node did
    n = y.shape[0]
    db = torch.empty_like(y)
    call("pvg_absdiff_mean_bwd", target.data_ptr(), y.data_ptr(), gl.data_ptr(), n, y.numel() // n, db.data_ptr(), _stream())
    return db.add_(dy)


class TapL1Fn(torch.autograd.Function):
    """(feature, mean|target - feature| per sample): AbsDiffMeanFn for a feature that is also consumed by deeper layers."""

    @staticmethod
    def forward(ctx, x, target, defer):
        _check_cuda(x); _check_cuda(target)
        if x.shape != target.shape or not (_is_nhwc_dense(x) and _is_nhwc_dense(target)):
            raise _lib.PvgError("tap_l1 needs two channels-last feature maps of one shape")
        n = x.shape[0]
        out = zero_pool.zeros((n,), torch.float64, x.device)
        call("pvg_absdiff_mean_fwd", target.data_ptr(), x.data_ptr(), n, x.numel() // n, out.data_ptr(), _stream())
        ctx.save_for_backward(target, x)
        ctx.defer = defer
        ctx.set_materialize_grads(False)
        return x.view_as(x), out.float()

    @staticmethod
    def backward(ctx, g_x, g_loss):
        if g_loss is None:
            return g_x, None, None
        target, x = ctx.saved_tensors
        gl = g_loss.contiguous().float()
        if g_x is not None and ctx.defer and _is_nhwc_dense(g_x):
            g_x._pvg_tap = (target, gl, g_x._version, x)     # consumed by Conv2dFn.backward of the producing convolution
            return g_x, None, None
        n = x.shape[0]
        db = torch.empty_like(x)
        call("pvg_absdiff_mean_bwd", target.data_ptr(), x.data_ptr(), gl.data_ptr(), n, x.numel() // n, db.data_ptr(), _stream())
        return (db if g_x is None else db.add_(g_x)), None, None


def tap_l1(x: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    """Returns (x, loss[n] = mean|target[n] - x[n]|): ``x`` (an alias carrying x's operand planes) is what deeper layers consume."""
    fn = x.grad_fn
    defer = bool(fuse_taps and fn is not None and type(fn).__name__ == "Conv2dFnBackward")
    out, loss = TapL1Fn.apply(x, target.detach(), defer)
    pl = planes_of(x)
    if pl:
        out._pvg_planes = pl
    return out, loss


# ---------------------------------------------------------------------------------------------------------------
# evaluator-side reductions / frame conversion (no gradients)
# ---------------------------------------------------------------------------------------------------------------
def _frame_layout(t: Tensor) -> Optional[bool]:
    """True: every (b, t) frame of a (bs, T, C, H, W) tensor is channels-last dense; False: planar dense; None: neither."""
    bs, T, c, h, w = t.shape
    st = t.stride()
    if st[0] == T * c * h * w and st[1] == c * h * w:
        if st[2:] == (h * w, w, 1):
            return False
        if (c == 1 or st[2] == 1) and st[3] == w * c and st[4] == c:
            return True
    return None


def sqdiff_mean(reference: Tensor, generated: Tensor, motion_mask: bool = False) -> Tensor:
    """(bs, T) mean over (C, H, W) of (reference - generated)^2, optionally weighted per pixel by the reference's frame-difference
    motion mask (evaluation/metrics/mse.py:21, motion_masked_mse.py:23-26) - one fused reduction."""
    _check_cuda(reference); _check_cuda(generated)
    if reference.shape != generated.shape or reference.dim() != 5:
        raise _lib.PvgError("sqdiff_mean takes two (bs, T, C, H, W) tensors of the same shape")
    la, lb = _frame_layout(reference), _frame_layout(generated)
    if la is None or la != lb:
        reference, generated = reference.contiguous(), generated.contiguous()
        la = False
    bs, T, c, h, w = reference.shape
    out = torch.zeros((bs * T,), dtype=torch.float64, device=reference.device)
    call("pvg_sqdiff_mean", reference.data_ptr(), generated.data_ptr(), bs * T, T, c, h * w, 1 if la else 0, 1 if motion_mask else 0,
         out.data_ptr(), _stream())
    return out.float().reshape(bs, T)


def frames_to_uint8(x: Tensor) -> Tensor:
    """uint8 copy of x in its own element order: (x * 255) truncated, after (x + 1) / 2 when any element is negative."""
    _check_cuda(x)
    if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
        x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device).as_strided(x.shape, x.stride())
    scratch = torch.empty((1,), dtype=torch.int32, device=x.device)
    call("pvg_frames_to_u8", x.data_ptr(), x.numel(), scratch.data_ptr(), out.data_ptr(), _stream())
    return out


# ---------------------------------------------------------------------------------------------------------------
# optimiser
# ---------------------------------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999,
              eps: float = 1e-8, weight_decay: float = 0.0, grad_scale: float = 1.0) -> None:
    """In-place torch.optim.Adam update (L2-style weight decay) over flat fp32 buffers."""
    call("pvg_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), _stream())


def adam_step_dev(p: Tensor, g: Tensor, m: Tensor, v: Tensor, hyper: Tensor) -> None:
    """Adam update whose step-dependent scalars live in device memory (``hyper``, 7 floats) - graph-capturable."""
    call("pvg_adam_step_dev", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), hyper.data_ptr(), _stream())
