"""Kernel-level parity (GPU): every C-ABI op against the same op in plain PyTorch fp32 on the CPU (what the oracle is
built from).  Tolerances are written next to each check.  Runs on the B200 box: ``pytest -m gpu``."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from playablevideogeneration_b200 import ops
    return ops


def _log(name, **kw):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "kernel_errors.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **kw)) + "\n")
    except Exception:
        pass


def _err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float()
    return float((got - ref).abs().max()), float(ref.abs().max())


def _close(name, got, ref, rel, abs_=0.0):
    e, scale = _err(got, ref)
    extra = {}
    if e > rel * scale + abs_:          # where and how many: tells a systematic error from a localised one
        d = (got.detach().float().cpu() - ref.detach().float()).abs()
        idx = int(d.reshape(-1).argmax())
        extra = dict(bad=int((d > rel * scale + abs_).sum()), numel=d.numel(), argmax=idx,
                     got=float(got.detach().float().cpu().reshape(-1)[idx]), ref=float(ref.detach().float().reshape(-1)[idx]))
    _log(name, max_abs_err=e, ref_absmax=scale, tol=rel * scale + abs_, **extra)
    assert e <= rel * scale + abs_, f"{name}: max abs err {e:.3e} > {rel:g} * {scale:.3e} + {abs_:g}"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


# -----------------------------------------------------------------------------------------------------------------
CONV_SIMT_SHAPES = [  # N, Cin, Cout, H, W, k, bias
    (2, 3, 16, 20, 20, 3, False), (1, 12, 16, 12, 36, 3, False), (2, 16, 32, 16, 16, 1, False),
    (2, 16, 16, 10, 14, 3, True), (1, 3, 64, 32, 32, 3, True), (2, 20, 3, 16, 16, 7, True),
    # shared-memory tiled head kernels (conv_direct.cu): ragged tiles, several channel chunks
    (2, 32, 3, 40, 70, 7, True), (1, 16, 3, 21, 36, 7, True), (2, 64, 3, 20, 68, 3, True), (1, 128, 2, 17, 64, 3, False),
    (2, 3, 32, 9, 70, 7, False), (1, 3, 16, 12, 33, 7, True), (1, 3, 64, 6, 64, 7, False)]


@pytest.mark.parametrize("shape", CONV_SIMT_SHAPES)
def test_conv_simt_forward(shape):
    ops = _ops()
    n, cin, cout, h, w, k, has_bias = shape
    x, wt = _rand(n, cin, h, w, seed=1), _rand(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5)
    b = _rand(cout, seed=3) if has_bias else None
    ref = _conv_ref64(x, wt, b, k)
    ops.set_precision("fp32")
    try:
        got = ops.conv2d(x.to(DEV), wt.to(DEV), b.to(DEV) if has_bias else None)
    finally:
        ops.set_precision("tf32x3")
    _close(f"conv_simt{shape}", got, ref, 2e-6, 1e-6)      # fp32 FMA chains of <= 1k terms, different summation order


CONV_UMMA_SHAPES = [  # N, Cin_logical, Cin_physical, Cout, H, W, k, bias, act
    (2, 32, 32, 64, 16, 16, 3, False, 0), (1, 64, 64, 65, 32, 32, 3, False, 0), (2, 128, 128, 128, 32, 32, 3, False, 0),
    (2, 201, 224, 512, 8, 8, 3, True, 0), (3, 64, 64, 128, 12, 20, 1, False, 0), (1, 32, 32, 3, 64, 64, 7, True, 3),
    (2, 64, 64, 64, 26, 20, 3, True, 2), (1, 128, 128, 3, 16, 16, 3, True, 3), (8, 256, 256, 256, 4, 4, 3, False, 0),
    (1, 64, 64, 64, 128, 128, 3, True, 2), (2, 521, 544, 1024, 16, 16, 3, True, 0),
    # channel counts that are not a multiple of 32: the K loop is completed by TMA out-of-bounds zero fill
    (2, 16, 16, 16, 20, 24, 3, False, 0), (1, 16, 16, 32, 16, 16, 1, True, 0), (2, 40, 40, 48, 12, 12, 3, False, 2),
    # the shapes bench.py times: 256x256 maps (61 440-tile VGG conv1_2 family, 64 -> 32 decoder layer), the 13x10 / 6x16 state
    # maps of full-size Breakout / Tennis, batches of >= 64 frames, the N = 4C ConvLSTM gate conv
    (2, 64, 64, 64, 256, 256, 3, True, 2), (2, 64, 64, 32, 256, 256, 3, False, 0), (64, 64, 64, 128, 13, 10, 3, False, 0),
    (64, 128, 128, 128, 6, 16, 3, True, 0), (96, 32, 32, 32, 26, 20, 3, False, 0), (128, 16, 16, 16, 48, 128, 3, False, 0),
    (8, 265, 288, 512, 32, 32, 3, True, 0), (120, 256, 256, 256, 16, 16, 3, True, 2)]


def _conv_ref64(x, wt, b, k):
    """CPU reference in float64, rounded to fp32.  (The fp32 oneDNN convolution of the GPU box's host was seen returning
    bf16-grade results - 1e-2 off on 8 % of the outputs of a 7x7 conv - depending on which convolutions the process had run
    before, while the CUDA result was bit-identical to the true fp32 value; float64 takes the plain reference path.)"""
    return F.conv2d(x.double(), wt.double(), None if b is None else b.double(), padding=k // 2).float()


def _conv_case(shape, seed=0):
    n, cin, cinp, cout, h, w, k, has_bias, act = shape
    x = _rand(n, cinp, h, w, seed=seed + 1)
    x[:, cin:] = 0
    wt = _rand(cout, cin, k, k, seed=seed + 2, scale=(cin * k * k) ** -0.5)
    b = _rand(cout, seed=seed + 3) if has_bias else None
    ref = _conv_ref64(x[:, :cin], wt, b, k)
    if act == 2:
        ref = F.relu(ref)
    elif act == 3:
        ref = torch.tanh(ref.double()).float()
    return x, wt, b, ref


def test_tf32_probe_reports_rounding_mode():
    ops = _ops()
    trunc = ops.tf32_truncates()
    _log("tf32_probe", truncates=bool(trunc))
    assert trunc in (True, False)


@pytest.fixture(params=["bf16", "fp16", "tf32", "h3"])
def corr(request):
    """Every evaluation of the two correction products of the fp32-equivalent split (C-ABI nprod = 2 with bf16 / fp16
    planes, nprod = 3), applied to the forward, data-gradient and weight-gradient kernels alike."""
    ops = _ops()
    ops.set_correction(request.param, request.param, request.param)
    yield request.param
    ops.set_correction()


@pytest.mark.parametrize("shape", CONV_UMMA_SHAPES)
def test_conv_umma_forward_tf32x3(shape, corr):
    """error-compensated split tensor-core conv (TF32 main product + bf16 or TF32 corrections) vs fp32 CPU:
    fp32-equivalent (tolerance 1e-5 of the output scale)."""
    ops = _ops()
    x, wt, b, ref = _conv_case(shape)
    ops.set_precision("tf32x3")
    xd, wd, bd = x.to(DEV), wt.to(DEV), (b.to(DEV) if b is not None else None)
    got = ops.conv2d(xd, wd, bd, act=shape[8])
    e, scale = _err(got, ref)
    if e > 1e-5 * scale + 1e-6:          # localise before failing: input corruption, non-determinism or arithmetic?
        got2 = ops.conv2d(xd, wd, bd, act=shape[8])
        pre = ops.conv2d(xd, wd, bd, act=0)
        pre_ref = F.conv2d(x[:, :shape[1]], wt, b, padding=shape[6] // 2)
        d = (pre.cpu() - pre_ref).abs()
        idx = (d > 1e-4 * float(pre_ref.abs().max())).nonzero()
        ref64 = F.conv2d(x[:, :shape[1]].double(), wt.double(), None if b is None else b.double(), padding=shape[6] // 2)
        _log(f"conv_umma3_diag[{corr}]{shape}", second_call_err=_err(got2, ref)[0], first_vs_second=float((got - got2).abs().max()),
             x_intact=bool((xd.cpu() == x).all()), w_intact=bool((wd.cpu() == wt).all()), pre_act_err=float(d.max()),
             pre_act_bad=int(len(idx)), bad_sample=idx[:8].tolist(), algo=str(ops._conv_algo(shape[2], shape[3], shape[6], "fwd")),
             ours_vs_fp64=float((pre.cpu().double() - ref64).abs().max()), cpu32_vs_fp64=float((pre_ref.double() - ref64).abs().max()),
             threads=torch.get_num_threads())
    _close(f"conv_umma3[{corr}]{shape}", got, ref, 1e-5, 1e-6)


@pytest.mark.parametrize("shape", CONV_UMMA_SHAPES[:5])
def test_conv_umma_forward_tf32(shape):
    """single TF32 product: 10-bit mantissa operands -> 2e-3 of the output scale."""
    ops = _ops()
    x, wt, b, ref = _conv_case(shape)
    ops.set_precision("tf32")
    try:
        got = ops.conv2d(x.to(DEV), wt.to(DEV), b.to(DEV) if b is not None else None, act=shape[8])
    finally:
        ops.set_precision("tf32x3")
    _close(f"conv_umma1{shape}", got, ref, 2e-3, 1e-5)


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 16, 16, 3, True, 0), (2, 201, 224, 512, 8, 8, 3, True, 0),
                                   (2, 3, 3, 16, 16, 16, 3, False, 0), (1, 64, 64, 65, 16, 16, 1, False, 0),
                                   (2, 32, 32, 3, 16, 16, 7, True, 3), (2, 137, 160, 256, 16, 16, 3, False, 0),
                                   (2, 32, 32, 3, 40, 70, 7, True, 3), (1, 16, 16, 3, 21, 36, 7, True, 0),
                                   (2, 16, 16, 16, 16, 16, 3, False, 0), (2, 16, 16, 32, 12, 20, 3, True, 0),
                                   (1, 16, 16, 32, 16, 16, 1, False, 0), (2, 40, 40, 24, 12, 12, 3, False, 0),
                                   (2, 3, 3, 16, 37, 45, 3, True, 0), (1, 12, 12, 16, 12, 36, 3, False, 0),
                                   (1, 3, 3, 64, 20, 20, 3, False, 0),
                                   (8, 64, 64, 32, 64, 64, 3, False, 0), (3, 64, 64, 128, 26, 20, 3, True, 0),
                                   (2, 96, 96, 16, 16, 16, 1, False, 0), (2, 256, 256, 132, 8, 8, 3, False, 0),
                                   # timed shapes: 256x256 decoder layer, full-size Breakout / Tennis state maps, N >= 64
                                   (2, 64, 64, 32, 256, 256, 3, False, 0), (64, 64, 64, 128, 13, 10, 3, False, 0),
                                   (64, 128, 128, 128, 6, 16, 3, True, 0), (96, 32, 32, 32, 26, 20, 3, False, 0)])
def test_conv_backward(shape, corr):
    """dx (tensor-core dgrad with flipped packed weights or SIMT), dw (split-K wgrad), db vs torch autograd on CPU."""
    _conv_backward_case(shape, corr)


def _conv_backward_case(shape, corr):
    ops = _ops()
    n, cin, cinp, cout, h, w, k, has_bias, act = shape
    x, wt, b, _ = _conv_case(shape, seed=10)
    xr = x[:, :cin].double().requires_grad_(True)          # float64 autograd reference (see _conv_ref64)
    wr = wt.double().requires_grad_(True)
    br = b.double().requires_grad_(True) if has_bias else None
    ref = F.conv2d(xr, wr, br, padding=k // 2)
    if act == 3:
        ref = torch.tanh(ref)
    gy = _rand(*ref.shape, seed=20)
    ref.backward(gy.double())
    xg = x.to(DEV).requires_grad_(True)
    wg = wt.to(DEV).requires_grad_(True)
    bg = b.to(DEV).requires_grad_(True) if has_bias else None
    out = ops.conv2d(xg, wg, bg, act=act)
    out.backward(gy.to(DEV))
    _close(f"conv_bwd_dx[{corr}]{shape}", xg.grad[:, :cin], xr.grad, 1e-5, 1e-6)
    _close(f"conv_bwd_dw[{corr}]{shape}", wg.grad, wr.grad, 2e-5, 1e-6)
    if has_bias:
        _close(f"conv_bwd_db[{corr}]{shape}", bg.grad, br.grad, 1e-5, 1e-6)


@pytest.mark.parametrize("shape", [(2, 201, 224, 512, 8, 8, 3, True, 0), (2, 128, 128, 128, 16, 16, 3, False, 0),
                                   (2, 64, 64, 3, 16, 16, 3, True, 3)])
def test_conv_backward_fp32_simt(shape):
    """same as above with every conv forced onto the CUDA-core fp32 path (the cross-check implementation)."""
    ops = _ops()
    ops.set_precision("fp32")
    try:
        _conv_backward_case(shape, "fp32")
    finally:
        ops.set_precision("tf32x3")


# -----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [dict(c=16, pool=True, res=False, train=True, groups=1),
                                 dict(c=65, pool=False, res=True, train=True, groups=1),
                                 dict(c=128, pool=False, res=False, train=True, groups=3),
                                 dict(c=64, pool=True, res=False, train=False, groups=1),
                                 dict(c=256, pool=False, res=True, train=False, groups=1)])
def test_pool_bn_act(cfg):
    ops = _ops()
    c, pool, use_res, train, groups = cfg["c"], cfg["pool"], cfg["res"], cfg["train"], cfg["groups"]
    n, h, w = 6, 8, 12
    x = _rand(n, c, h, w, seed=1).requires_grad_(True)
    oh, ow = (h // 2, w // 2) if pool else (h, w)
    res = _rand(n, c, oh, ow, seed=2).requires_grad_(True) if use_res else None
    bn_ref = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn_ref.weight.copy_(1 + 0.1 * _rand(c, seed=3)); bn_ref.bias.copy_(0.1 * _rand(c, seed=4))
        bn_ref.running_mean.copy_(0.1 * _rand(c, seed=5)); bn_ref.running_var.copy_(1 + 0.2 * torch.rand(c))
    import copy
    bn_gpu = copy.deepcopy(bn_ref).to(DEV)
    bn_ref.train(train); bn_gpu.train(train)
    xp = F.avg_pool2d(x, 2) if pool else x
    outs = []
    for g in range(groups):                       # reference semantics: one BatchNorm call per group, in order
        sl = slice(g * n // groups, (g + 1) * n // groups)
        outs.append(bn_ref(xp[sl]))
    y = torch.cat(outs, 0)
    if use_res:
        y = y + res
    ref = F.leaky_relu(y, 0.2)
    gy = _rand(*ref.shape, seed=6)
    ref.backward(gy)
    xg = x.detach().to(DEV).requires_grad_(True)
    rg = res.detach().to(DEV).requires_grad_(True) if use_res else None
    got = ops.pool_bn_act(xg, bn_gpu, residual=rg, pool=pool, act=ops.ACT_LRELU, slope=0.2, groups=groups)
    got.backward(gy.to(DEV))
    tag = str(cfg)
    _close("bn_fwd" + tag, got, ref, 2e-6, 2e-6)
    _close("bn_dx" + tag, xg.grad, x.grad, 1e-5, 1e-6)
    if use_res:
        _close("bn_dres" + tag, rg.grad, res.grad, 1e-6, 1e-7)
    _close("bn_dw" + tag, bn_gpu.weight.grad, bn_ref.weight.grad, 1e-5, 1e-5)
    _close("bn_db" + tag, bn_gpu.bias.grad, bn_ref.bias.grad, 1e-5, 1e-5)
    _close("bn_rm" + tag, bn_gpu.running_mean, bn_ref.running_mean, 1e-5, 1e-6)
    _close("bn_rv" + tag, bn_gpu.running_var, bn_ref.running_var, 1e-5, 1e-6)
    ops.flush_deferred()          # num_batches_tracked increments are batched (Model.forward flushes them)
    assert int(bn_gpu.num_batches_tracked) == int(bn_ref.num_batches_tracked)


@pytest.mark.parametrize("shape", [(2, 16, 5, 7), (1, 65, 8, 8), (3, 128, 16, 4)])
def test_upsample2x(shape):
    ops = _ops()
    x = _rand(*shape, seed=1).requires_grad_(True)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear")
    gy = _rand(*ref.shape, seed=2)
    ref.backward(gy)
    xg = x.detach().to(DEV).requires_grad_(True)
    got = ops.upsample2x(xg)
    got.backward(gy.to(DEV))
    _close(f"up2x_fwd{shape}", got, ref, 1e-6, 1e-7)
    _close(f"up2x_bwd{shape}", xg.grad, x.grad, 2e-6, 1e-7)


@pytest.mark.parametrize("case", [((2, 3, 64, 64), (32, 32)), ((2, 3, 64, 64), (16, 16)), ((1, 3, 96, 64), (24, 16)),
                                  ((1, 3, 30, 50), (17, 23))])
def test_resize_bilinear(case):
    ops = _ops()
    shape, size = case
    x = _rand(*shape, seed=1)
    _close(f"resize{case}", ops.resize_bilinear(x.to(DEV), size), F.interpolate(x, size, mode="bilinear"), 1e-6, 1e-7)


def test_maxpool2():
    ops = _ops()
    x = F.relu(_rand(2, 64, 12, 20, seed=1)).requires_grad_(True)      # many exact zeros -> ties, like VGG
    ref = F.max_pool2d(x, 2)
    gy = _rand(*ref.shape, seed=2)
    ref.backward(gy)
    xg = x.detach().to(DEV).requires_grad_(True)
    got = ops.maxpool2(xg)
    got.backward(gy.to(DEV))
    _close("maxpool_fwd", got, ref, 0.0, 0.0)
    nz = (x.detach() > 0)
    _close("maxpool_bwd_nonzero", xg.grad.cpu() * nz, x.grad * nz, 0.0, 0.0)   # ties at 0 are killed by relu' anyway


def test_lstm_cell():
    ops = _ops()
    n, c, h, w = 2, 32, 6, 5
    gates = _rand(n, 4 * c, h, w, seed=1).requires_grad_(True)
    cprev = _rand(n, c, h, w, seed=2).requires_grad_(True)
    i, f, o, g = gates.chunk(4, dim=1)
    cn = torch.sigmoid(f) * cprev + torch.sigmoid(i) * torch.tanh(g)
    hn = torch.sigmoid(o) * torch.tanh(cn)
    gh, gc = _rand(n, c, h, w, seed=3), _rand(n, c, h, w, seed=4)
    (hn * gh + cn * gc).sum().backward()
    gg = gates.detach().to(DEV).requires_grad_(True)
    cg = cprev.detach().to(DEV).requires_grad_(True)
    h2, c2 = ops.lstm_cell(gg, cg)
    (h2 * gh.to(DEV) + c2 * gc.to(DEV)).sum().backward()
    _close("lstm_h", h2, hn, 2e-6, 1e-6)
    _close("lstm_c", c2, cn, 2e-6, 1e-6)
    _close("lstm_dgates", gg.grad, gates.grad, 1e-5, 1e-6)
    _close("lstm_dc", cg.grad, cprev.grad, 1e-5, 1e-6)


@pytest.mark.parametrize("shape", [(2, 201, 224, 128, 32, 32), (3, 137, 160, 64, 13, 10), (8, 521, 544, 256, 16, 16)])
def test_fused_convlstm_step_matches_fp64(shape):
    """pvg_convlstm_step (gate convolution + cell update in one launch, interleaved gate columns) and its backward against the
    reference cell (convolutional_lstm_cell.py:88-101) in float64: four separate gate convolutions, sigmoid / tanh, c', h'."""
    ops = _ops()
    n, cin, cin_p, c, h, w = shape
    z = _rand(n, cin_p, h, w, seed=1)
    z[:, cin:] = 0
    ws = [_rand(c, cin, 3, 3, seed=10 + k, scale=(cin * 9) ** -0.5) for k in range(4)]
    bs = [_rand(c, seed=20 + k, scale=0.1) for k in range(4)]
    cp = _rand(n, c, h, w, seed=3)
    zr = z[:, :cin].double().requires_grad_(True)
    wr = [t.double().requires_grad_(True) for t in ws]
    br = [t.double().requires_grad_(True) for t in bs]
    cr = cp.double().requires_grad_(True)
    gi, gf, go = (torch.sigmoid(F.conv2d(zr, wr[k], br[k], padding=1)) for k in range(3))
    gc = torch.tanh(F.conv2d(zr, wr[3], br[3], padding=1))
    cn = gf * cr + gi * gc
    hn = go * torch.tanh(cn)
    gh, gcn = _rand(n, c, h, w, seed=4), _rand(n, c, h, w, seed=5)
    (hn * gh.double() + cn * gcn.double()).sum().backward()
    zg = z.to(DEV).requires_grad_(True)
    wg = [t.to(DEV).requires_grad_(True) for t in ws]
    bg = [t.to(DEV).requires_grad_(True) for t in bs]
    cg = cp.to(DEV).requires_grad_(True)
    w_il = torch.stack(wg, dim=1).reshape(4 * c, cin, 3, 3)
    b_il = torch.stack(bg, dim=1).reshape(-1)
    assert ops.supports_fused_lstm()
    h2, c2 = ops.convlstm_step(zg, w_il, b_il, cg)
    (h2 * gh.to(DEV) + c2 * gcn.to(DEV)).sum().backward()
    tag = str(shape)
    _close("fused_lstm_h" + tag, h2, hn, 1e-5, 2e-6)
    _close("fused_lstm_c" + tag, c2, cn, 1e-5, 2e-6)
    _close("fused_lstm_dz" + tag, zg.grad[:, :cin], zr.grad, 2e-5, 1e-6)
    _close("fused_lstm_dc" + tag, cg.grad, cr.grad, 1e-5, 1e-6)
    for k in range(4):
        _close(f"fused_lstm_dw{k}" + tag, wg[k].grad, wr[k].grad, 3e-5, 1e-6)
        _close(f"fused_lstm_db{k}" + tag, bg[k].grad, br[k].grad, 2e-5, 1e-5)


def test_concat_pad():
    ops = _ops()
    a, v, hdd = _rand(2, 64, 4, 6, seed=1).requires_grad_(True), _rand(2, 9, seed=2).requires_grad_(True), _rand(2, 128, 4, 6, seed=3).requires_grad_(True)
    ref = torch.cat([a, v[:, :, None, None].expand(-1, -1, 4, 6), hdd], dim=1)
    gy = _rand(2, 224, 4, 6, seed=4)
    (ref * gy[:, :201]).sum().backward()
    ag, vg, hg = (t.detach().to(DEV).requires_grad_(True) for t in (a, v, hdd))
    got = ops.concat_pad([ag, vg, hg])
    assert got.shape == (2, 224, 4, 6) and float(got[:, 201:].abs().max()) == 0.0
    (got * gy.to(DEV)).sum().backward()
    _close("concat_fwd", got[:, :201], ref, 0.0, 0.0)
    _close("concat_dv", vg.grad, v.grad, 1e-6, 1e-6)
    _close("concat_da", ag.grad, a.grad, 0.0, 0.0)


def test_producer_planes_equal_a_split_pass():
    """Producers that write the 16-bit operand planes of their result next to it (BatchNorm apply, conv epilogue, max-pool,
    bilinear upsample, channel concat) must write exactly what a separate pvg_split_16 pass over the result would."""
    ops = _ops()
    from playablevideogeneration_b200 import _lib
    fmts = (_lib.CORR_FP16_ALL, _lib.CORR_BF16)

    def check(name, y):
        got = ops.planes_of(y)
        assert got, name
        for f, pl in got.items():
            ref = ops._split(y.detach(), 2, f)[1]
            assert torch.equal(pl.view(torch.int16), ref.view(torch.int16)), (name, f)

    x = _rand(4, 64, 12, 20, seed=1).to(DEV)
    bn = torch.nn.BatchNorm2d(64).to(DEV).train()
    check("bn_train", ops.pool_bn_act(x, bn, pool=True, act=ops.ACT_LRELU, planes=fmts))
    res = _rand(4, 64, 12, 20, seed=2).to(DEV)
    check("bn_eval_res", ops.pool_bn_act(x, bn.eval(), residual=res, act=ops.ACT_LRELU, planes=fmts))
    check("upsample", ops.upsample2x(x, planes=fmts))
    check("maxpool", ops.maxpool2(x, planes=fmts[:1]))
    a, v, hdd = _rand(2, 64, 4, 6, seed=1).to(DEV), _rand(2, 9, seed=2).to(DEV), _rand(2, 128, 4, 6, seed=3).to(DEV)
    check("concat", ops.concat_pad([a, v, hdd], planes=fmts))
    wt = _rand(72, 64, 3, 3, seed=4, scale=0.05).to(DEV)
    y = ops.conv2d(ops.nhwc(x), wt, act=ops.ACT_RELU, out_planes=True)
    assert list(ops.planes_of(y)) == [_lib.CORR_FP16_ALL]
    check("conv_epilogue", y)
    # and a consumer fed by producer planes gives the result of the split path, bit for bit
    wt2 = _rand(32, 72, 3, 3, seed=5, scale=0.05).to(DEV)
    z_planes = ops.conv2d(y, wt2)
    z_split = ops.conv2d(y.detach().clone(), wt2)
    assert torch.equal(z_planes, z_split)


@pytest.mark.parametrize("shape", [(4, 64, 128, 16, 24, 2), (6, 32, 64, 13, 10, 3), (8, 128, 72, 32, 32, 1), (12, 64, 32, 6, 16, 4)])
def test_conv_epilogue_batchnorm_statistics(shape):
    """BatchNorm statistics accumulated by the conv epilogue (per batch group) against sums over the conv's own output."""
    ops = _ops()
    n, cin, cout, h, w, groups = shape
    x = _rand(n, cin, h, w, seed=1).to(DEV)
    wt = _rand(cout if cout % 8 == 0 else 65, cin, 3, 3, seed=2, scale=(cin * 9) ** -0.5).to(DEV)
    y = ops.conv2d(x, wt, bn_stats_groups=groups, cout_phys=cout if wt.shape[0] != cout else None)
    sums, g = y._pvg_bn_sums
    assert g == groups and tuple(sums.shape) == (groups, 2, y.shape[1])
    yy = y.detach().double().reshape(groups, n // groups, y.shape[1], h * w)
    ref = torch.stack([yy.sum(dim=(1, 3)), (yy * yy).sum(dim=(1, 3))], dim=1)
    _close(f"epilogue_stats{shape}", sums.float(), ref.float().cpu(), 2e-6, 1e-5)
    # and the BatchNorm that consumes them equals the one that computes its own statistics
    bn = torch.nn.BatchNorm2d(wt.shape[0]).to(DEV).train()
    import copy
    bn2 = copy.deepcopy(bn)
    a = ops.pool_bn_act(y, bn, act=ops.ACT_LRELU, groups=groups)
    y_plain = y.detach().clone()
    b = ops.pool_bn_act(y_plain, bn2, act=ops.ACT_LRELU, groups=groups)
    _close(f"epilogue_stats_bn{shape}", a, b.cpu(), 2e-6, 2e-6)


def test_gradient_amax_tags_replace_the_amax_pass_bit_for_bit(monkeypatch):
    """conv data-gradient epilogue and BatchNorm backward write max|dx| next to dx; the conv backward above them takes its
    power-of-two scale from that tag instead of a pvg_amax pass.  Same data gradient, bit for bit, and fewer amax launches;
    a tag on a tensor that was modified afterwards is not trusted."""
    ops = _ops()

    def run(track):
        monkeypatch.setattr(ops, "track_amax", track)
        calls = []
        real = ops.call
        monkeypatch.setattr(ops, "call", lambda name, *a: (calls.append(name), real(name, *a))[1])
        torch.manual_seed(0)
        x = ops.nhwc(_rand(4, 32, 16, 24, seed=1).to(DEV)).requires_grad_(True)
        w1 = _rand(64, 32, 3, 3, seed=2, scale=0.06).to(DEV).requires_grad_(True)
        w2 = _rand(64, 64, 3, 3, seed=3, scale=0.04).to(DEV).requires_grad_(True)
        w3 = _rand(32, 64, 3, 3, seed=4, scale=0.04).to(DEV).requires_grad_(True)
        bn1 = torch.nn.BatchNorm2d(64).to(DEV).train()
        bn2 = torch.nn.BatchNorm2d(64).to(DEV).train()
        h = ops.pool_bn_act(ops.conv2d(x, w1), bn1, pool=True, act=ops.ACT_LRELU)
        h = ops.maxpool2(ops.pool_bn_act(ops.conv2d(h, w2), bn2, act=ops.ACT_RELU))
        y = ops.conv2d(h, w3, act=ops.ACT_TANH)
        y.backward(_rand(*y.shape, seed=5).to(DEV))
        monkeypatch.setattr(ops, "call", real)
        return [t.grad.clone() for t in (x, w1, w2, w3, bn1.weight, bn2.bias)], calls.count("pvg_amax")

    tagged, n_tagged = run(True)
    plain, n_plain = run(False)
    assert n_plain == 3 and n_tagged == 1, (n_plain, n_tagged)      # only the loss-side gradient still needs the pass
    assert torch.equal(tagged[0], plain[0])            # same scale -> same operand planes -> the same data gradient
    for i, (a, b) in enumerate(zip(tagged[1:], plain[1:])):     # split-K / statistics accumulate with atomics: order varies
        _close(f"amax_tags_grad{i}", a, b.cpu(), 1e-6, 1e-6)
    monkeypatch.setattr(ops, "track_amax", True)
    t = torch.ones(8, device=DEV)
    ops.tag_amax(t, torch.zeros(1, dtype=torch.int32, device=DEV))
    assert ops.known_amax(t) is not None
    t.add_(1.0)
    assert ops.known_amax(t) is None


def test_feature_tap_l1_folded_into_the_conv_backward(monkeypatch):
    """ops.tap_l1 (the perceptual loss's L1 on a VGG feature that deeper layers also consume): same loss values as
    absdiff_mean, and the gradient that reaches the image equals the one autograd assembles from the separate L1 backward and
    the deeper gradient - on the image-facing CUDA-core conv (3 -> 64), on tensor-core convs, through a max-pool, and for the
    last tap (no deeper consumer)."""
    ops = _ops()
    img = _rand(6, 3, 16, 24, seed=1).to(DEV)
    ws = [_rand(64, 3, 3, 3, seed=2, scale=0.2), _rand(64, 64, 3, 3, seed=3, scale=0.05), _rand(128, 64, 3, 3, seed=4, scale=0.05),
          _rand(128, 128, 3, 3, seed=5, scale=0.04)]
    ws = [w.to(DEV) for w in ws]
    bs = [_rand(w.shape[0], seed=10 + i, scale=0.1).to(DEV) for i, w in enumerate(ws)]
    coef = _rand(6, seed=20).to(DEV)           # a different upstream gradient per sample and tap

    def chain(x, targets, fused):
        losses, feats = [], []
        for i, (w, b) in enumerate(zip(ws, bs)):
            if i == 2:
                x = ops.maxpool2(x, planes=ops.conv_input_planes(weight_grad=False))
            x = ops.conv2d(x, w, b, act=ops.ACT_RELU, out_planes=i in (0, 2))
            if i in (0, 2, 3):
                feats.append(x)
                if targets is None:
                    continue
                if fused:
                    x, l = ops.tap_l1(x, targets[len(losses)])
                else:
                    l = ops.absdiff_mean(targets[len(losses)], x)
                losses.append(l)
        return feats, losses

    with torch.no_grad():
        targets, _ = chain(ops.nhwc(_rand(6, 3, 16, 24, seed=7).to(DEV)), None, False)
    results = []
    for fused in (True, False):
        calls = []
        real = ops.call
        monkeypatch.setattr(ops, "call", lambda name, *a: (calls.append(name), real(name, *a))[1])
        x = ops.nhwc(img.clone()).requires_grad_(True)
        _, losses = chain(x, targets, fused)
        total = sum((l * coef * (k + 1)).sum() for k, l in enumerate(losses))
        total.backward()
        monkeypatch.setattr(ops, "call", real)
        results.append(([l.detach().clone() for l in losses], x.grad.clone(), calls))
    (lf, gf, cf), (lu, gu, cu) = results
    for a, b in zip(lf, lu):
        assert torch.equal(a, b)
    _close("tap_l1_image_grad", gf, gu.cpu(), 2e-6, 1e-9)
    assert cf.count("pvg_act_bwd_tap") == 1 and cf.count("pvg_act_bwd_tap_split_16_scaled") == 1, cf
    assert cf.count("pvg_absdiff_mean_bwd") == 1 and cu.count("pvg_absdiff_mean_bwd") == 3      # fused: only the last tap
    # a tapped convolution WITHOUT activation cannot fold the term into an activation-backward pass: its backward adds it itself
    grads = []
    for fused in (True, False):
        x = ops.nhwc(_rand(2, 64, 8, 16, seed=30).to(DEV)).requires_grad_(True)
        f = ops.conv2d(x, ws[1])
        tgt = _rand(2, 64, 8, 16, seed=31).to(DEV)
        tgt = ops.nhwc(tgt)
        if fused:
            f, l = ops.tap_l1(f, tgt)
        else:
            l = ops.absdiff_mean(tgt, f)
        (ops.conv2d(f, ws[1], bs[1], act=ops.ACT_RELU).sum() * 1e-3 + l.sum()).backward()
        grads.append(x.grad.clone())
    _close("tap_l1_no_activation", grads[0], grads[1].cpu(), 2e-6, 1e-9)


@pytest.mark.parametrize("shape", [(2, 16, 32), (3, 13, 40), (1, 256, 256), (5, 9, 7)])
def test_stem_conv_3_to_64_with_planes(shape):
    """pvg_conv2d_stem_planes (VGG conv1_1): fp32 CUDA-core result against float64, and the plane pair it writes for
    conv1_2 against a split pass over its own output; ragged tiles included."""
    ops = _ops()
    from playablevideogeneration_b200 import _lib
    n, h, w = shape
    x = _rand(n, 3, h, w, seed=1).to(DEV)
    wt = _rand(64, 3, 3, 3, seed=2, scale=0.3).to(DEV)
    b = _rand(64, seed=3, scale=0.2).to(DEV)
    calls = []
    real = ops.call
    ops.call = lambda name, *a: (calls.append(name), real(name, *a))[1]
    try:
        y = ops.conv2d(ops.nhwc(x), wt, b, act=ops.ACT_RELU, out_planes=True)
    finally:
        ops.call = real
    assert "pvg_conv2d_stem_planes" in calls
    ref = F.relu(F.conv2d(x.double().cpu(), wt.double().cpu(), b.double().cpu(), padding=1))
    _close(f"stem{shape}", y, ref.float(), 2e-6, 1e-6)
    pl = ops.planes_of(y)[_lib.CORR_FP16_ALL]
    want = ops._split(y.detach(), 2, _lib.CORR_FP16_ALL)[1]
    assert torch.equal(pl.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("shape", [(3, 16, 16, 20, 40), (2, 16, 32, 13, 33), (4, 16, 32, 8, 8), (2, 16, 16, 37, 64), (128, 16, 16, 128, 128)])
def test_small_channel_weight_gradient_kernel(shape):
    """pvg_conv2d_wgrad_small (16 input channels, the first encoder stage) through conv2d's backward, against float64; the
    data gradient of the same call still comes from the tensor-core kernel.  The 32-input-channel instantiations of the
    kernel are exercised through the C ABI directly."""
    ops = _ops()
    n, cin, cout, h, w = shape
    x = ops.nhwc(_rand(n, cin, h, w, seed=1).to(DEV)).requires_grad_(True)
    wt = _rand(cout, cin, 3, 3, seed=2, scale=(cin * 9) ** -0.5).to(DEV).requires_grad_(True)
    gy = _rand(n, cout, h, w, seed=3).to(DEV)
    calls = []
    real = ops.call
    ops.call = lambda name, *a: (calls.append(name), real(name, *a))[1]
    try:
        y = ops.conv2d(x, wt)
        y.backward(gy)
    finally:
        ops.call = real
    assert "pvg_conv2d_wgrad_small" in calls and "pvg_conv2d_wgrad_planes" not in calls
    if n * h * w <= 200000:
        xd, wd = x.detach().double().cpu().requires_grad_(True), wt.detach().double().cpu().requires_grad_(True)
        F.conv2d(xd, wd, padding=1).backward(gy.double().cpu())
        _close(f"wgrad_small{shape}", wt.grad, wd.grad.float(), 2e-6, 1e-6)
        _close(f"wgrad_small_dx{shape}", x.grad, xd.grad.float(), 2e-6, 1e-6)
    else:                       # full encoder size: against the tensor-core kernel
        small = wt.grad.clone()
        wt.grad = None; x.grad = None
        ops.small_wgrad_kernel = False
        try:
            ops.conv2d(x, wt).backward(gy)
        finally:
            ops.small_wgrad_kernel = True
        _close(f"wgrad_small_vs_tc{shape}", small, wt.grad.cpu(), 1e-5, 1e-5)


def test_small_channel_weight_gradient_kernel_32_inputs():
    ops = _ops()
    from playablevideogeneration_b200._lib import ConvDesc
    from playablevideogeneration_b200 import _lib
    for cout in (16, 32):
        n, cin, h, w = 3, 32, 11, 40
        x = ops.nhwc(_rand(n, cin, h, w, seed=1).to(DEV))
        g = ops.nhwc(_rand(n, cout, h, w, seed=2).to(DEV))
        scratch = torch.zeros(cout * 9 * 32, device=DEV)
        dw = torch.empty(cout, cin, 3, 3, device=DEV)
        d = ConvDesc(n, h, w, cin, cout, 3, 3, 1, 0, 0.0, _lib.ALGO_SIMT, 1, 0)
        ops.call("pvg_conv2d_wgrad_small", d, cin, x.data_ptr(), g.data_ptr(), scratch.data_ptr(), dw.data_ptr(), 0,
                 torch.cuda.current_stream().cuda_stream)
        wd = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
        F.conv2d(x.double().cpu(), wd, padding=1).backward(g.double().cpu())
        _close(f"wgrad_small_c32_{cout}", dw, wd.grad.float(), 2e-6, 1e-6)


def test_concat_pad_strided_time_slices():
    """Maps handed to the concat as time slices of a (B, T, C, H, W) tensor (batch-strided, NHWC-dense per sample) are read in
    place."""
    ops = _ops()
    seq = ops.empty_nhwc((3 * 5, 64, 4, 6), DEV).normal_().reshape(3, 5, 64, 4, 6)
    v = _rand(3, 9, seed=2).to(DEV)
    got = ops.concat_pad([seq[:, 2], v])
    ref = torch.cat([seq[:, 2], v[:, :, None, None].expand(-1, -1, 4, 6)], dim=1)
    assert got.shape == (3, 96, 4, 6) and torch.equal(got[:, :73], ref) and float(got[:, 73:].abs().max()) == 0.0


def test_padded_65_channel_block_runs_on_tensor_cores_and_matches_fp64():
    """The encoder tail ResidualBlock(64 -> 65) (representation_network.py:28) on tensors physically padded to 72 channels:
    forward, input gradient and every parameter gradient against a float64 torch reference; padding channels stay zero; the
    BatchNorm running statistics (65 entries) are updated as by nn.BatchNorm2d."""
    import copy
    ops = _ops()
    from playablevideogeneration_b200 import _lib
    from playablevideogeneration_b200.caddy import ResidualBlock
    torch.manual_seed(3)
    blk = ResidualBlock(64, 65, 1)
    ref = copy.deepcopy(blk).double()
    blk = blk.to(DEV).train()
    x = _rand(4, 64, 16, 16, seed=5)
    xr = x.double().requires_grad_(True)
    out = ref.bn1(F.conv2d(xr, ref.conv1.weight, padding=1))
    out = F.leaky_relu(out, 0.2)
    out = ref.bn2(F.conv2d(out, ref.conv2.weight, padding=1))
    idn = ref.downsample[2](F.conv2d(xr, ref.downsample[0].weight))
    yr = F.leaky_relu(out + idn, 0.2)
    gy = _rand(4, 65, 16, 16, seed=6)
    yr.backward(gy.double())
    assert ops.supports_padded_cout()
    calls = []
    orig = _lib.call
    def spy(name, *a):
        calls.append((name, a[0].algo if a and isinstance(a[0], _lib.ConvDesc) else None))
        return orig(name, *a)
    ops.call = spy
    try:
        xg = x.to(DEV).requires_grad_(True)
        y = blk(xg)
        assert y.shape[1] == 72 and float(y[:, 65:].abs().max()) == 0.0
        gpad = torch.zeros_like(y)
        gpad[:, :65] = gy.to(DEV)
        y.backward(gpad)
    finally:
        ops.call = orig
    assert not any(n in ("pvg_conv2d_fwd", "pvg_conv2d_wgrad") and algo == _lib.ALGO_SIMT for n, algo in calls), \
        "a 65-channel convolution fell back to the CUDA-core kernels"
    _close("pad65_fwd", y[:, :65], yr, 2e-5, 2e-6)
    _close("pad65_dx", xg.grad, xr.grad, 2e-5, 1e-6)
    for (k, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        _close("pad65_d" + k, p.grad, q.grad, 5e-5, 1e-6)
    _close("pad65_running_mean", blk.bn2.running_mean, ref.bn2.running_mean, 1e-5, 1e-6)
    _close("pad65_running_var", blk.bn1.running_var, ref.bn1.running_var, 1e-5, 1e-6)


def test_absdiff_mean():
    ops = _ops()
    a, b = _rand(5, 64, 9, 7, seed=1), _rand(5, 64, 9, 7, seed=2).requires_grad_(True)
    ref = (a - b).abs().mean(dim=[1, 2, 3])
    gw = _rand(5, seed=3)
    (ref * gw).sum().backward()
    bg = b.detach().to(DEV).requires_grad_(True)
    got = ops.absdiff_mean(a.to(DEV), bg)
    (got * gw.to(DEV)).sum().backward()
    _close("absdiff_fwd", got, ref, 1e-6, 0.0)
    _close("absdiff_bwd", bg.grad, b.grad, 1e-6, 1e-9)


def test_evaluator_reductions_and_uint8_frames():
    """pvg_sqdiff_mean (MSE / PSNR / motion-masked MSE of evaluation/metrics/*.py in one pass, both frame layouts) and
    pvg_frames_to_u8 (the builder's uint8 conversion) against their plain torch expressions."""
    ops = _ops()
    from playablevideogeneration_b200.evaluation.samplers import frames_to_uint8_hwc
    ref = torch.rand((2, 5, 3, 20, 28), generator=torch.Generator().manual_seed(1))
    gen = (ref + 0.1 * _rand(2, 5, 3, 20, 28, seed=2)).clamp(0, 1)
    want = (ref - gen).pow(2).mean(dim=[2, 3, 4])
    mask = torch.abs(ref[:, 1:] - ref[:, :-1]).sum(dim=2, keepdim=True) / 3
    mask = torch.cat([torch.zeros_like(mask[:, 0:1]), mask], dim=1)
    want_m = ((ref - gen).pow(2) * mask).mean(dim=[2, 3, 4])
    for layout in ("planar", "channels_last"):
        a, b = ref.to(DEV), gen.to(DEV)
        if layout == "channels_last":
            a = a.reshape(10, 3, 20, 28).contiguous(memory_format=torch.channels_last).reshape(2, 5, 3, 20, 28)
            b = b.reshape(10, 3, 20, 28).contiguous(memory_format=torch.channels_last).reshape(2, 5, 3, 20, 28)
        _close("sqdiff_" + layout, ops.sqdiff_mean(a, b), want, 2e-6, 1e-9)
        _close("sqdiff_mask_" + layout, ops.sqdiff_mean(a, b, motion_mask=True), want_m, 2e-6, 1e-10)
    for rng in ("signed", "unit"):
        x = _rand(3, 3, 16, 24, seed=3).clamp(-1, 1) if rng == "signed" else torch.rand((3, 3, 16, 24), generator=torch.Generator().manual_seed(4))
        xn = (x + 1) / 2 if float(x.min()) < 0 else x
        want_u8 = (xn * 255).clamp(0, 255).to(torch.uint8).movedim(-3, -1).contiguous()
        for fmt in (torch.contiguous_format, torch.channels_last):
            got = frames_to_uint8_hwc(x.to(DEV).contiguous(memory_format=fmt))
            assert got.dtype == torch.uint8 and tuple(got.shape) == (3, 16, 24, 3)
            assert torch.equal(got.cpu(), want_u8), (rng, fmt, int((got.cpu() != want_u8).sum()))


def test_adam_matches_torch():
    ops = _ops()
    p0, steps = _rand(1000, seed=1), 3
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=4e-4, weight_decay=1e-6)
    p = p0.to(DEV); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for s in range(1, steps + 1):
        g = _rand(1000, seed=10 + s)
        p_ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(DEV), m, v, s, 4e-4, weight_decay=1e-6)
    _close("adam", p, p_ref, 1e-6, 1e-7)


def test_ops_refuse_cpu_tensors():
    ops = _ops()
    from playablevideogeneration_b200._lib import PvgError
    with pytest.raises(PvgError):
        ops.conv2d(torch.zeros(1, 32, 8, 8), torch.zeros(16, 32, 3, 3))
