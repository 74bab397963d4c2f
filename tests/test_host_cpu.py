"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/pvg_b200.h declares, the module
tree reproduces the reference checkpoint layout, error behaviour matches, and - with the kernels replaced by CPU
stand-ins (tests/fake_ops.py, test-only) - the host logic (control flow, RNG draw order, 20-tuple layout, loss weighting,
flat-arena optimiser step) reproduces the golden outputs of the unmodified reference."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import caddy_oracle as O
from oracle.cases import CASES, RESULT_NAMES_FULL, RESULT_NAMES_PRE, build_config
from tests.golden_util import batch_tuple, case_inputs, compare_results, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from playablevideogeneration_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pvg_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pvg_[a-z0-9_]+)\s*\(", header)))
    assert os.path.isfile(_lib.LIB_PATH), "libpvg_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == _lib.EXPORTED_SYMBOLS, set(declared) ^ set(_lib.EXPORTED_SYMBOLS)
    lib.pvg_version.restype = ctypes.c_int
    assert lib.pvg_version() >= 100
    _lib.load()


def test_header_is_plain_c_and_matches_the_ctypes_struct():
    """include/pvg_b200.h is the boundary: it must compile as C99 on its own (no C++ or torch types), and the ctypes mirror
    of pvg_conv_desc must have the layout the C compiler gives the struct."""
    import shutil
    import subprocess
    import tempfile
    from playablevideogeneration_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "probe.c")
        exe = os.path.join(td, "probe")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include <stddef.h>\n#include "pvg_b200.h"\n'
                    'int main(void) { printf("%zu %zu %zu %zu\\n", sizeof(pvg_conv_desc), offsetof(pvg_conv_desc, slope), '
                    'offsetof(pvg_conv_desc, nprod), offsetof(pvg_conv_desc, corr_fmt)); return 0; }\n')
        subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src],
                       check=True)
        size, o_slope, o_nprod, o_fmt = (int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split())
    D = _lib.ConvDesc
    assert ctypes.sizeof(D) == size
    assert (D.slope.offset, D.nprod.offset, D.corr_fmt.offset) == (o_slope, o_nprod, o_fmt)


def test_ops_have_no_cpu_fallback():
    from playablevideogeneration_b200 import ops
    from playablevideogeneration_b200._lib import PvgError
    with pytest.raises(PvgError):
        ops.conv2d(torch.zeros(1, 32, 8, 8), torch.zeros(16, 32, 3, 3))
    with pytest.raises(PvgError):
        ops.upsample2x(torch.zeros(1, 4, 8, 8))


@pytest.mark.parametrize("kind,reduced,hw", [("bair", False, (256, 256)), ("breakout", True, (208, 160)),
                                             ("tennis", False, (96, 256))])
def test_state_dict_layout_matches_reference(kind, reduced, hw):
    from playablevideogeneration_b200.caddy import Model
    S = 4 if kind == "tennis" else 1
    cfg = build_config(dict(config=kind, H=hw[0], W=hw[1], S=S))
    m = Model(cfg, reduced=reduced)
    sd = m.state_dict()
    spec = O.model_param_spec(cfg, reduced)
    assert [k for k, _, _ in spec] == list(sd.keys())
    assert all(tuple(sd[k].shape) == s for k, s, _ in spec)
    m.load_state_dict(O.make_weights(cfg, 0, reduced), strict=True)
    if kind == "bair":
        assert sum(p.numel() for p in m.parameters()) == 9856367


def test_factories_and_error_behaviour():
    import importlib
    cfg = build_config(dict(config="bair", H=64, W=64, S=1))
    for mod in ("playablevideogeneration_b200.model.main_model.model",):
        m = getattr(importlib.import_module(mod), "model")(cfg)
        assert hasattr(m, "generate_next") and hasattr(m, "start_inference") and hasattr(m, "centroid_estimator")
    cfg_r = build_config(dict(config="breakout", H=96, W=64, S=1))
    mr = getattr(importlib.import_module("playablevideogeneration_b200.model.reduced_model.model"), "model")(cfg_r)
    assert mr.rendering_network.final_blocks[2].conv.weight.shape == (3, 16, 7, 7)
    obs = torch.zeros(1, 3, 3, 64, 64)
    with pytest.raises(Exception, match="ground truth observations > 0"):
        m(batch_tuple(obs), ground_truth_observations_init=0)
    cfg2 = build_config(dict(config="bair", H=64, W=64, S=1)); cfg2["training"]["pretraining_detach"] = True


@pytest.fixture
def fake_ops(monkeypatch):
    from tests import fake_ops as fo
    fo.install(monkeypatch)


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["mode"] in ("full", "pretraining")])
def test_host_logic_reproduces_reference_with_cpu_standins(name, fake_ops):
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.training.step import TrainStep
    from playablevideogeneration_b200.vgg import Vgg19
    case, g = load_case(name)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    model = Model(cfg, reduced=case.get("reduced", False))
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    step = TrainStep(cfg, model, Vgg19(vgg_sd))
    model.train()
    n_steps = case.get("steps", 1)
    for s in range(n_steps):
        torch.manual_seed(case["noise_seed"] + s); random.seed(case["noise_seed"] + s)
        total, info, res = step.compute_losses(batch_tuple(obs), case["gt_init"], case["gumbel_temperature"],
                                               pretraining=case["mode"] == "pretraining")
        tag = "" if s == 0 else f"step{s}."
        ref_total = float(g[tag + "total_loss"][0])
        assert abs(float(total) - ref_total) <= 2e-6 * abs(ref_total), (s, float(total), ref_total)
        if s == 0:
            names = RESULT_NAMES_PRE if case["mode"] == "pretraining" else RESULT_NAMES_FULL
            compare_results(g, names, res, rtol=2e-5, atol=2e-5)
        step.arena.zero_grad()
        total.backward()
        if s == 0:
            for k, p in model.named_parameters():
                key = "gradnorm." + k
                if key in g.files:
                    ref = float(g[key])
                    assert abs(float(p.grad.double().norm()) - ref) <= 1e-3 * ref + 1e-7, k
        if n_steps > 1:
            step.optimizer_step()
    if n_steps > 1:
        for k, p in model.named_parameters():
            key = "param_after." + k
            if key in g.files:
                from oracle.cases import sample_tensor
                got = sample_tensor(p.detach(), stride=max(1, p.numel() // 64))
                # Adam's lr*g/(|g|+eps) is sign-like: where a gradient is rounding noise the update flips between
                # implementations, so bound the difference by the maximum possible drift (steps * lr) and require the
                # bulk of the entries to agree closely.
                diff = np.abs(got - g[key])
                assert diff.max() <= 2 * 4e-4 * 1.05 and (diff.size < 32 or (diff > 2e-5).mean() <= 0.4), (k, diff.max())


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["mode"] == "rollout"])
def test_rollout_host_logic_with_cpu_standins(name, fake_ops):
    from playablevideogeneration_b200.caddy import Model
    case, g = load_case(name)
    cfg, sd, _, obs = case_inputs(case)
    model = Model(cfg, reduced=case.get("reduced", False))
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    model.eval()
    torch.manual_seed(case["noise_seed"])
    with torch.no_grad():
        model.start_inference()
        for i, a in enumerate(case["actions"]):
            frame, obs = model.generate_next(obs, a, noise=case.get("noise", False))
            np.testing.assert_allclose(frame.numpy(), g[f"frame.{i}"], rtol=2e-5, atol=2e-5)
        for i, (a1, a2, f) in enumerate(case.get("interp", [])):
            frame, obs = model.generate_next_interpolation(obs, a1, a2, f)
            np.testing.assert_allclose(frame.numpy(), g[f"iframe.{i}"], rtol=2e-5, atol=2e-5)


def test_zero_pool_and_deferred_counters():
    """Launch hygiene helpers of ops.py (host logic only): the per-step zero pool hands out disjoint, zero-filled, correctly
    typed views after one memset, falls back to torch.zeros outside a step or when exhausted, never resizes an arena in
    place; deferred integer counters are applied once by flush_deferred()."""
    from playablevideogeneration_b200 import ops
    pool = ops._ZeroPool()
    dev = torch.device("cpu")
    a = pool.zeros((3, 2, 5), torch.float64, dev)                 # outside a step: plain zeros, but demand is recorded
    assert a.dtype == torch.float64 and a.shape == (3, 2, 5) and float(a.abs().sum()) == 0.0 and pool.demand > 0
    pool.begin(dev)                                               # sizes the arena from the recorded demand
    first = pool.buf
    assert pool.active and first is not None
    x = pool.zeros((7,), torch.float32, dev)
    y = pool.zeros((2, 2, 4), torch.float64, dev)
    x.fill_(3.0); y.fill_(5.0)
    assert float(x.sum()) == 21.0 and float(y.sum()) == 80.0      # disjoint regions
    assert x.data_ptr() % 256 == first.data_ptr() % 256 and (y.data_ptr() - x.data_ptr()) % 256 == 0
    big = pool.zeros((1 << 20,), torch.float32, dev)              # exhausted -> fallback, still zero
    assert float(big.abs().sum()) == 0.0 and big.data_ptr() != first.data_ptr()
    pool.end()
    assert not pool.active
    pool.begin(dev)                                               # demand grew: a NEW arena, the old one is kept alive
    assert pool.buf is not first and any(b is first for b in pool._keep)
    z = pool.zeros((7,), torch.float32, dev)
    assert float(z.abs().sum()) == 0.0                            # the memset of begin() cleared whatever lived here
    pool.end()

    c1, c2 = torch.zeros((), dtype=torch.long), torch.tensor(4, dtype=torch.long)
    ops.defer_count(c1, 1); ops.defer_count(c2, 2); ops.defer_count(c1, 3)
    assert int(c1) == 0 and int(c2) == 4                          # nothing applied yet
    ops.flush_deferred()
    assert int(c1) == 4 and int(c2) == 6
    ops.flush_deferred()                                          # idempotent when nothing is pending
    assert int(c1) == 4 and int(c2) == 6


def test_trainstep_checkpoint_is_the_reference_trainers(tmp_path):
    """TrainStep.state_dict() is the dictionary Trainer.save_checkpoint writes (training/trainer.py:100,
    smooth_mi_trainer.py:43-45): the flat-arena Adam state round-trips through torch.optim.Adam's own state_dict layout
    (same parameter order, parameters without a gradient have no entry, per-parameter step counts), the learning-rate
    state loads into a real MultiStepLR, and a checkpoint file written by one TrainStep resumes another."""
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.training.step import TrainStep
    from playablevideogeneration_b200.vgg import Vgg19
    cfg = build_config(dict(config="bair", H=64, W=64, S=1))
    torch.manual_seed(0)
    model = Model(cfg)
    step = TrainStep(cfg, model, Vgg19(O.make_vgg_weights()))
    params = list(model.parameters())
    tr = cfg["training"]
    # a real torch.optim.Adam on the same parameters, two steps, some parameters never receive a gradient
    opt = torch.optim.Adam(params, lr=tr["learning_rate"], weight_decay=tr["weight_decay"])
    g = torch.Generator().manual_seed(1)
    skipped = {0, 5} | {i for i, p in enumerate(params) if not p.requires_grad}
    for it in range(2):
        for i, p in enumerate(params):
            p.grad = None if (i in skipped or (it == 0 and i == 7)) else torch.randn(p.shape, generator=g) * 1e-2
        opt.step()
    ref = opt.state_dict()
    step.load_optimizer_state_dict(ref)
    a = step.arena
    index = step._adam_index()
    assert len(index) == len(params) - 1                      # the centroid parameter has requires_grad = False (no arena slot)
    for k, (p, o, i) in enumerate(zip(a.params, a.offsets, index)):
        n = p.numel()
        assert p is params[i]
        if i in ref["state"]:
            st = ref["state"][i]
            assert a.steps[k] == int(float(st["step"])) == (1 if i == 7 else 2)
            assert torch.equal(step.exp_avg[o:o + n].view(p.shape), st["exp_avg"])
            assert torch.equal(step.exp_avg_sq[o:o + n].view(p.shape), st["exp_avg_sq"])
        else:
            assert i in skipped and a.steps[k] == 0 and float(step.exp_avg[o:o + n].abs().sum()) == 0.0
    mine = step.optimizer_state_dict()
    assert sorted(mine["state"]) == sorted(ref["state"]) and mine["param_groups"][0]["params"] == ref["param_groups"][0]["params"]
    opt2 = torch.optim.Adam(params, lr=tr["learning_rate"], weight_decay=tr["weight_decay"])
    opt2.load_state_dict(mine)                                        # torch accepts it as its own
    for i, st in ref["state"].items():
        assert torch.equal(opt2.state[params[i]]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(opt2.state[params[i]]["step"]) == float(st["step"])
    # learning-rate schedule
    step.global_step = 7
    sched = torch.optim.lr_scheduler.MultiStepLR(opt2, milestones=[int(m) for m in tr["lr_schedule"]], gamma=tr["lr_gamma"])
    sched.load_state_dict(step.lr_scheduler_state_dict())
    assert sched.last_epoch == 7 and abs(sched.get_last_lr()[0] - step.current_lr()) < 1e-12
    # file round trip with the reference's keys
    path = str(tmp_path / "latest.pth.tar")
    step.save_checkpoint(path)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"model", "optimizer", "lr_scheduler", "step", "mi_estimator"} and ckpt["step"] == 7
    assert list(ckpt["model"].keys()) == [k for k, _, _ in O.model_param_spec(cfg, False)]
    torch.manual_seed(5)
    model2 = Model(cfg)
    step2 = TrainStep(cfg, model2, Vgg19(O.make_vgg_weights()))
    step2.load_checkpoint(path)
    assert step2.global_step == 7 and step2.arena.steps == a.steps
    assert torch.equal(step2.exp_avg, step.exp_avg) and torch.equal(step2.arena.flat, a.flat)
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def _metric_inputs():
    g = torch.Generator().manual_seed(123)
    shape = (2, 4, 3, 32, 48)
    ref = torch.rand(shape, generator=g)
    gen = (ref + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    return ref, gen


def test_evaluation_metrics_match_the_reference(fake_ops):
    """evaluation/metrics/{mse,psnr,motion_masked_mse,vgg_cosine_similarity}.py: golden values produced by the unmodified
    reference classes (oracle/make_metric_golden.py) on the same seeded inputs; VGG through the CPU stand-ins here, through
    the CUDA kernels in tests/test_model_gpu.py."""
    from playablevideogeneration_b200.evaluation.metrics import MSE, PSNR, MotionMaskedMSE, VGGCosineSimilarity
    from playablevideogeneration_b200.vgg import Vgg19
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    ref, gen = _metric_inputs()
    np.testing.assert_allclose(MSE()(ref, gen).numpy(), g["mse"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(PSNR()(ref, gen).numpy(), g["psnr"], rtol=1e-6)
    np.testing.assert_allclose(PSNR()(ref * 255, gen * 255, range=255.0).numpy(), g["psnr_range255"], rtol=1e-5)
    np.testing.assert_allclose(MotionMaskedMSE()(ref, gen).numpy(), g["motion_masked_mse"], rtol=1e-6, atol=1e-9)
    vcs = VGGCosineSimilarity(Vgg19(O.make_vgg_weights()))
    got = vcs(ref, gen)
    assert got.shape == (2, 4)
    np.testing.assert_allclose(got.numpy(), g["vgg_cosine"], rtol=2e-5)
    # sequence helpers of training/losses.py:591-713 (evaluator.py:191-203)
    from playablevideogeneration_b200.training.losses import MotionLossWeightMaskCalculator, SequenceLossEvaluator, StatesLoss
    np.testing.assert_allclose(MotionLossWeightMaskCalculator(0.3).compute_weight_mask(ref, gen).numpy(), g["weight_mask_same"], rtol=1e-6)
    np.testing.assert_allclose(MotionLossWeightMaskCalculator(0.0).compute_weight_mask(ref, gen[:, 1:]).numpy(), g["weight_mask_short"],
                               rtol=1e-6, atol=1e-7)
    ev = SequenceLossEvaluator(StatesLoss())
    for tag, rec in (("same", gen), ("short", gen[:, 1:])):
        avg, terms = ev(ref, rec)
        np.testing.assert_allclose(terms.numpy(), g[f"seq_{tag}_terms"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(float(avg), float(g[f"seq_{tag}_avg"]), rtol=1e-6)
    with pytest.raises(Exception):
        ev(ref, gen[:, 2:])


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["mode"] == "eval"])
def test_eval_forward_with_samplers_host_logic(name, fake_ops):
    """build_evaluation_dataset.py path (evaluation_dataset_builder.py:47-54) through the module mirror with CPU stand-ins:
    eval mode, OneHotActionSampler + ZeroActionVariationSampler plug-ins, against the unmodified reference's outputs; and
    the builder's frame conversion against numpy."""
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.evaluation.samplers import (GroundTruthActionSampler, OneHotActionSampler,
                                                                  ZeroActionVariationSampler, frames_to_uint8_hwc)
    case, g = load_case(name)
    cfg, sd, _, obs = case_inputs(case)
    model = Model(cfg)
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    model.eval()
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    with torch.no_grad():
        res = model(batch_tuple(obs), ground_truth_observations_init=case["gt_init"], action_sampler=OneHotActionSampler(),
                    action_variation_sampler=ZeroActionVariationSampler(), gumbel_temperature=case["gumbel_temperature"])
    compare_results(g, RESULT_NAMES_FULL, res, rtol=2e-5, atol=2e-5)
    # frame conversion of the builder: pad with the first ground-truth frame, [-1, 1] -> [0, 1], channels last, uint8
    rec = torch.cat([obs[:, 0:1, 0:3], res[0]], dim=1)
    want = (np.moveaxis(((rec + 1) / 2).numpy(), 2, -1) * 255).clip(0, 255).astype(np.uint8)
    got = frames_to_uint8_hwc(rec)
    assert got.dtype == torch.uint8 and tuple(got.shape) == want.shape and np.array_equal(got.numpy(), want)
    # GroundTruthActionSampler: one-hot of the translated ground-truth action
    logp = torch.log_softmax(torch.randn(5, 4), dim=1)
    oh = GroundTruthActionSampler({0: 2, 1: 0, 2: 1})(logp, torch.tensor([0, 1, 2, 3, 1]))
    assert oh.argmax(dim=1).tolist() == [2, 0, 1, 3, 0] and float(oh.sum()) == 5.0
    assert OneHotActionSampler()(logp, None).argmax(dim=1).tolist() == logp.argmax(dim=1).tolist()


def test_input_pipeline_formula_is_the_reference_transform():
    """tests/golden/input_pipeline.npz holds the output of the unmodified TransformsGenerator.get_final_transforms (PIL crop +
    ToTensor + Normalize, dataset/transforms.py:90-108) on seeded uint8 frames.  The formula the CUDA kernel implements -
    ((u8 / 255) - 0.5) / 0.5 in fp32 on the crop box - must reproduce it bit for bit (the GPU test then holds the kernel
    to the same formula)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "input_pipeline.npz"))
    left, top, right, bottom = (int(v) for v in g["crop"])
    u8 = torch.from_numpy(g["frames"])[:, top:bottom, left:right]                   # (N, H, W, 3)
    got = ((u8.float() / 255.0) - 0.5) / 0.5
    assert np.array_equal(got.permute(0, 3, 1, 2).numpy(), g["out"])


def test_small_distribution_losses_match_the_reference():
    """KLDivergence, EntropyLogitLoss, EntropyProbabilityLoss, KLGaussianDivergenceLoss, KLGeneralGaussianDivergenceLoss
    (training/losses.py:121-209, 339-376) against values and input gradients of the unmodified reference classes
    (tests/golden/small_losses.npz, generated by oracle/make_loss_golden.py)."""
    import numpy as np
    from oracle.make_loss_golden import evaluate
    from playablevideogeneration_b200.training import losses as L
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_losses.npz"))
    got = evaluate(L)
    assert sorted(got) == sorted(golden.files)
    for k in golden.files:
        np.testing.assert_allclose(got[k], golden[k], rtol=2e-6, atol=1e-7, err_msg=k)


def _numpy_resample(img, bounds, kk, axis):
    """Integer restatement of Pillow's 8-bit resampling pass (test infrastructure): img (H, W, 3) uint8."""
    import numpy as np
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((len(bounds),) + src.shape[1:], dtype=np.uint8)
    for a, ((first, count), k) in enumerate(zip(bounds, kk)):
        ss = np.full(src.shape[1:], 1 << 21, dtype=np.int64)
        for x in range(count):
            ss = ss + src[first + x] * k[x]
        out[a] = np.clip(ss >> 22, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def test_pil_resize_coefficients_reproduce_pillow():
    """ops.pil_bilinear_coeffs (the host half of pvg_resample_u8) + an integer restatement of the two passes against what the
    reference's own transform (PIL crop + resize(BILINEAR), dataset/transforms.py:15-32) produced:
    tests/golden/input_resize.npz from oracle/make_metric_golden.py --input-pipeline.  Bit-exact."""
    import numpy as np
    from playablevideogeneration_b200 import ops
    g = np.load(os.path.join(ROOT, "tests", "golden", "input_resize.npz"))
    for name in ("down", "up", "mixed", "bair"):
        frames, crop, size, want = g[f"{name}.frames"], g[f"{name}.crop"], g[f"{name}.size"], g[f"{name}.u8"]
        if crop[0] >= 0:
            frames = frames[:, crop[1]:crop[3], crop[0]:crop[2]]
        ow, oh = int(size[0]), int(size[1])
        for f, w in zip(frames, want):
            cur = f
            if ow != cur.shape[1]:
                b, k, _ = ops.pil_bilinear_coeffs(cur.shape[1], ow)
                cur = _numpy_resample(cur, b, k, 1)
            if oh != cur.shape[0]:
                b, k, _ = ops.pil_bilinear_coeffs(cur.shape[0], oh)
                cur = _numpy_resample(cur, b, k, 0)
            assert np.array_equal(cur, w), name


def test_pil_resize_restatement_against_pillow_on_random_sizes():
    """Wider pin of ops.pil_bilinear_coeffs than the four committed goldens: when Pillow is importable, 60 random (input size,
    output size) pairs - shrinking by up to 7x, growing by up to 5x, prime sizes - must reproduce Image.resize(BILINEAR)."""
    import numpy as np
    PIL = pytest.importorskip("PIL")
    from PIL import Image
    from playablevideogeneration_b200 import ops
    rng = np.random.RandomState(5)
    for _ in range(60):
        h, w = int(rng.randint(3, 97)), int(rng.randint(3, 97))
        oh, ow = int(rng.randint(1, 120)), int(rng.randint(1, 120))
        img = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        cur = img
        if ow != w:
            b, k, _ = ops.pil_bilinear_coeffs(w, ow)
            cur = _numpy_resample(cur, b, k, 1)
        if oh != h:
            b, k, _ = ops.pil_bilinear_coeffs(h, oh)
            cur = _numpy_resample(cur, b, k, 0)
        assert np.array_equal(cur, want), (h, w, oh, ow)

