"""End-to-end parity (GPU): the CUDA path (playablevideogeneration_b200) against (a) the golden outputs of the UNMODIFIED
reference stored in tests/golden/ and (b) the CPU oracle on a fresh seed.

Stated tolerances (BASELINE.json north_star): reconstruction pixel MSE <= 1e-4, total loss <= 1e-5 relative - checked
in the default fp32-equivalent 3xTF32 mode.  Intermediate tensors and gradients are judged against a float64 run of
the oracle, relative to how far the fp32 CPU oracle itself is from it (see the test docstring)."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import caddy_oracle as O
from oracle.cases import CASES, RESULT_NAMES_FULL, RESULT_NAMES_PRE, sample_tensor
from tests.golden_util import batch_tuple, case_inputs, compare_results, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda"
TRAIN_CASES = [n for n, c in CASES.items() if c["mode"] in ("full", "pretraining")]
ROLLOUT_CASES = [n for n, c in CASES.items() if c["mode"] == "rollout"]
ROLLOUT_TOL = 2e-5      # max abs error of a generated frame (tanh output in [-1, 1]) against the reference golden: what the CPU
                        # oracle itself is held to (tests/test_oracle_golden.py); the CUDA path measures ~3e-6


def _log(name, **kw):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "model_errors.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **kw)) + "\n")
    except Exception:
        pass


def _build(case, cfg, sd, vgg_sd):
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.training.step import TrainStep
    from playablevideogeneration_b200.vgg import Vgg19
    model = Model(cfg, reduced=case.get("reduced", False))
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    model = model.to(DEV)
    step = TrainStep(cfg, model, Vgg19(vgg_sd)) if vgg_sd is not None else None
    return model, step


def _to_dev(bt):
    return tuple(t.to(DEV) for t in bt)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_step_matches_reference_golden(name):
    """(1) the contract against the UNMODIFIED reference's stored outputs: total loss <= 1e-5 rel, reconstruction pixel
    MSE <= 1e-4, each per-resolution loss term <= 2e-5 rel.
    (2) every returned tensor and every parameter gradient, judged by conditioning: with e(x) = relative L2 distance of x
    to the float64 oracle, require e(CUDA path) <= 10 * e(fp32 CPU oracle) + 1e-5.  (A fixed tolerance is meaningless
    here: E, R, D individually agree with the oracle to 2e-6 incl. gradients - tools/grad_diag2.py - while three
    free-running R->D->E steps with batch-2 train-mode BatchNorm amplify that to 1e-2 for BOTH fp32 implementations.)"""
    from tests.golden_util import flat_results, oracle_run, rel_l2
    case, g = load_case(name)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    model, step = _build(case, cfg, sd, vgg_sd)
    model.train()
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    total, info, res = step.compute_losses(_to_dev(batch_tuple(obs)), case["gt_init"], case["gumbel_temperature"],
                                           pretraining=case["mode"] == "pretraining")
    rec = sample_tensor(res[0].detach().cpu())
    mse = float(np.mean((rec - g["res.reconstructed_observations"]) ** 2))
    ref_total = float(g["total_loss"][0])
    got_total = float(total.detach().cpu()[0])
    rel = abs(got_total - ref_total) / abs(ref_total)
    _log(name, recon_mse=mse, loss=got_total, loss_ref=ref_total, loss_rel_err=rel)
    assert mse <= 1e-4, f"reconstruction MSE {mse:.3e} > 1e-4"
    assert rel <= 1e-5, f"total loss {got_total!r} vs reference {ref_total!r}: rel err {rel:.3e} > 1e-5"
    # index work is exact: the argmax-selected actions must equal the UNMODIFIED reference's, element for element
    sel_idx = (RESULT_NAMES_PRE if case["mode"] == "pretraining" else RESULT_NAMES_FULL).index("selected_actions")
    sel = res[sel_idx].detach().cpu().to(torch.int64).numpy()
    assert sel.shape == g["res.selected_actions"].shape and (sel == g["res.selected_actions"]).all(), \
        (sel.tolist(), g["res.selected_actions"].tolist())
    host = step.fetch_info(info)
    for r in range(3):
        for k in (f"perceptual_loss_r{r}", f"observations_rec_loss_r{r}"):
            ref = float(g["info." + k])
            assert abs(host[k] - ref) <= 2e-5 * abs(ref), (k, host[k], ref)
    step.arena.zero_grad()
    total.backward()
    # ---- (2) conditioning-aware comparison ---------------------------------------------------------------------
    t64, res64, g64 = oracle_run(case, torch.float64)
    t32, res32, g32 = oracle_run(case, torch.float32)
    names = RESULT_NAMES_PRE if case["mode"] == "pretraining" else RESULT_NAMES_FULL
    flat_names = []
    for nme, r in zip(names, res64):
        flat_names.extend([f"{nme}.{i}" for i in range(len(r))] if isinstance(r, (list, tuple)) else [nme])
    bad, worst_ratio = [], 0.0
    for nme, a, b32, b64 in zip(flat_names, flat_results(res), flat_results(res32), flat_results(res64)):
        if not b64.is_floating_point():
            agree = float((a.cpu() == b64).float().mean())
            if agree < 0.9:
                bad.append((nme, agree))
            continue
        e_ours, e_ref = rel_l2(a, b64), rel_l2(b32, b64)
        worst_ratio = max(worst_ratio, e_ours / (e_ref + 1e-6))
        if e_ours > 10 * e_ref + 1e-5:
            bad.append((nme, e_ours, e_ref))
    # The loss is not smooth (|a - b| signs, ReLU kinks, max-pool ties): a forward value that differs in the last bits can put
    # an element on the other side of a kink and move the gradient of the few parameters that see it by a fixed quantum.
    # tests/golden/kink_floor_<case>.json (oracle/make_kink_floor.py) holds, per parameter, how far the fp32 CPU ORACLE's own
    # gradient moves from the fp64 one when rounding-level noise (1e-6) is added to its convolutions; a gradient passes
    # when it is within 10x the unperturbed oracle's error or within 2x that floor.
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"kink_floor_{name}.json")) as f:
        kink = json.load(f)["floor"]
    gbad, gworst, at_kink = [], 0.0, []
    for k, p in model.named_parameters():
        if k in g64 and p.grad is not None and float(g64[k].norm()) > 1e-7:     # skip gradients that are pure rounding noise
            e_ours, e_ref = rel_l2(p.grad, g64[k]), rel_l2(g32[k], g64[k])
            if e_ours > 10 * e_ref + 1e-5:
                if e_ours <= 2 * kink.get(k, 0.0):
                    at_kink.append((k, e_ours, e_ref, kink[k]))
                    continue
                gbad.append((k, e_ours, e_ref, kink.get(k, 0.0)))
            gworst = max(gworst, e_ours / (e_ref + 1e-6))
    _log(name + ":conditioning", tensors_bad=len(bad), grads_bad=len(gbad), worst_tensor_ratio=worst_ratio,
         worst_grad_ratio=gworst, grads_within_kink_floor=len(at_kink), first=str((bad + gbad + at_kink)[:3]))
    assert not bad, f"outputs further from the fp64 truth than 10x the fp32 CPU oracle: {bad[:4]}"
    assert not gbad, f"{len(gbad)} gradients further from the fp64 truth than 10x the fp32 CPU oracle and 2x its kink floor: {gbad[:4]}"
    msd = model.state_dict()
    for k in g.files:
        if k.startswith("buf."):
            np.testing.assert_allclose(msd[k[4:]].detach().cpu().numpy(), g[k], rtol=2e-3, atol=2e-4, err_msg=k)


def test_two_optimizer_steps_match_reference():
    """forward + losses + backward + Adam, twice (golden 'full_bair_feedback' holds step-2 loss and parameters)."""
    name = "full_bair_feedback"
    case, g = load_case(name)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    model, step = _build(case, cfg, sd, vgg_sd)
    bt = _to_dev(batch_tuple(obs))
    losses = []
    for s in range(case["steps"]):
        torch.manual_seed(case["noise_seed"] + s); random.seed(case["noise_seed"] + s)
        total, _ = step.step(bt, case["gt_init"], case["gumbel_temperature"])
        losses.append(float(total.cpu()[0]))
    ref2 = float(g["step1.total_loss"][0])
    rel = abs(losses[1] - ref2) / abs(ref2)
    _log("two_steps", loss_step2=losses[1], ref=ref2, rel=rel)
    assert rel <= 5e-3, (losses, ref2)           # Adam's first update is sign-like (lr * g/|g|): rounding-noise gradients flip
    worst, worst_key = 0.0, None
    for k, p in model.named_parameters():
        key = "param_after." + k
        if key in g.files:
            got = sample_tensor(p.detach().cpu(), stride=max(1, p.numel() // 64))
            e = float(np.abs(got - g[key]).max())
            if e > worst:
                worst, worst_key = e, k
    _log("two_steps_params", worst_abs=worst, worst_param=worst_key)
    # Adam's first steps move each weight by ~lr = 4e-4 per step regardless of gradient scale (sign-like): a gradient element
    # that is rounding noise can differ by 2 * lr per step between any two fp32 implementations, and the EMA centroids follow.
    # This case is the ill-conditioned one (free-running feedback with batch-2 BatchNorm: the fp32 CPU oracle's own gradients
    # are 6.6 % from the float64 ones, and moving ONE block of this code between the CUDA-core and tensor-core kernels moves
    # them by 4 %, tools/pad_diag.py); over the arithmetic variants of this repo (TF32 / bf16 / fp16 / all-fp16 products, padded
    # or unpadded 65-channel tail) the worst entry - always the EMA centroids - measured 2.2e-3 ... 3.3e-3.
    assert worst <= 4e-3, (worst, worst_key)


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_rollout_matches_reference_golden(name):
    case, g = load_case(name)
    cfg, sd, _, obs = case_inputs(case)
    model, _ = _build(case, cfg, sd, None)
    model.eval()
    obs = obs.to(DEV)
    torch.manual_seed(case["noise_seed"])
    with torch.no_grad():
        model.start_inference()
        for i, a in enumerate(case["actions"]):
            frame, obs = model.generate_next(obs, a, noise=case.get("noise", False))
            err = float(np.abs(frame.cpu().numpy() - g[f"frame.{i}"]).max())
            _log(name, step=i, max_abs_err=err)
            assert err <= ROLLOUT_TOL, (i, err)
        for i, (a1, a2, f) in enumerate(case.get("interp", [])):          # interpolate.py:152
            frame, obs = model.generate_next_interpolation(obs, a1, a2, f)
            err = float(np.abs(frame.cpu().numpy() - g[f"iframe.{i}"]).max())
            _log(name, interp_step=i, max_abs_err=err)
            assert err <= ROLLOUT_TOL, (i, err)


def test_cuda_path_matches_cpu_oracle_on_fresh_seed():
    """Not in the goldens: new weights/inputs/noise; oracle (CPU) and CUDA path run side by side on the box.  The oracle runs
    in float64: that is the exact value both fp32 implementations approximate, and it keeps the check independent of which
    fp32 convolution primitive the box's oneDNN picks (see tests/test_kernels_gpu.py:_conv_ref64)."""
    from tests.golden_util import oracle_run
    case = dict(CASES["full_bair"], weight_seed=41, input_seed=42, noise_seed=43, gt_init=2, T=4, gumbel_temperature=0.8)
    cfg, sd, vgg_sd, obs = case_inputs(case)
    ref_total, ref_res, _ = oracle_run(case, torch.float64)
    model, step = _build(case, cfg, sd, vgg_sd)
    model.train()
    torch.manual_seed(43); random.seed(43)
    total, info, res = step.compute_losses(_to_dev(batch_tuple(obs)), 2, 0.8)
    mse = float(((res[0].detach().cpu().double() - ref_res[0].detach()) ** 2).mean())
    rel = abs(float(total.detach().cpu()[0]) - float(ref_total)) / abs(float(ref_total))
    _log("fresh_seed", recon_mse=mse, loss_rel_err=rel)
    assert mse <= 1e-4 and rel <= 1e-5, (mse, rel)


# BASELINE.json configs[1..3] at their FULL frame sizes (the shapes bench.py times: 256x256 BAIR, 208x160 Breakout with the
# reduced model and 26x20 -> 13x10 state maps, 96x256 Tennis with 4-frame stacking and 12x32 -> 6x16 maps), batch 2 and six
# frames so that the float64 oracle finishes in about a minute on the box's host cores.  No golden exists at these sizes (the
# reference would need minutes per case on the CPU): the yardstick is the oracle, which tests/test_oracle_golden.py pins to
# the unmodified reference on every small case.
FULL_SIZE_CASES = {
    "bair256": dict(CASES["full_bair"], B=2, T=6, H=256, W=256, gt_init=3, weight_seed=51, input_seed=52, noise_seed=53),
    "breakout208x160": dict(CASES["full_breakout"], B=2, T=6, H=208, W=160, gt_init=3, weight_seed=54, input_seed=55, noise_seed=56),
    "tennis96x256": dict(CASES["full_tennis"], B=2, T=6, H=96, W=256, gt_init=3, weight_seed=57, input_seed=58, noise_seed=59),
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE_CASES))
def test_full_size_train_step_matches_fp64_oracle(name):
    """Same contract as the golden cases, at the frame sizes of BASELINE.json configs[1..3]: total loss <= 1e-5 relative and
    reconstruction MSE <= 1e-4 against the float64 oracle; every returned tensor no further from it than 10x the fp32 CPU
    oracle; selected actions identical.  Gradients: >= 90 % of the parameter tensors within 10x the fp32 oracle's own error,
    none beyond 30x + 2e-3 (the per-parameter kink floors of tests/golden/kink_floor_*.json - how far an |a - b| / ReLU /
    max-pool kink moves a gradient under rounding-level noise - are not precomputed at this size; the small cases measured
    quanta of up to 6e-4)."""
    from tests.golden_util import flat_results, oracle_run, rel_l2
    case = FULL_SIZE_CASES[name]
    cfg, sd, vgg_sd, obs = case_inputs(case)
    model, step = _build(case, cfg, sd, vgg_sd)
    model.train()
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    total, info, res = step.compute_losses(_to_dev(batch_tuple(obs)), case["gt_init"], case["gumbel_temperature"])
    step.arena.zero_grad()
    total.backward()
    torch.cuda.synchronize()
    t64, res64, g64 = oracle_run(case, torch.float64)
    t32, res32, g32 = oracle_run(case, torch.float32)
    got, ref = float(total.detach().cpu()[0]), float(t64)
    rel = abs(got - ref) / abs(ref)
    mse = float(((res[0].detach().cpu().double() - res64[0].detach()) ** 2).mean())
    rel32 = abs(float(t32) - ref) / abs(ref)
    bad, worst = [], 0.0
    for i, (a, b32, b64) in enumerate(zip(flat_results(res), flat_results(res32), flat_results(res64))):
        if not b64.is_floating_point():
            if not bool((a.cpu() == b64).all()):
                bad.append((i, "integer tensor differs"))
            continue
        e_ours, e_ref = rel_l2(a, b64), rel_l2(b32, b64)
        worst = max(worst, e_ours / (e_ref + 1e-6))
        if e_ours > 10 * e_ref + 1e-5:
            bad.append((i, e_ours, e_ref))
    over10, over30, n_grads, gworst = [], [], 0, 0.0
    for k, p in model.named_parameters():
        if k in g64 and p.grad is not None and float(g64[k].norm()) > 1e-7:
            n_grads += 1
            e_ours, e_ref = rel_l2(p.grad, g64[k]), rel_l2(g32[k], g64[k])
            gworst = max(gworst, e_ours / (e_ref + 1e-6))
            if e_ours > 10 * e_ref + 1e-5:
                over10.append((k, e_ours, e_ref))
            if e_ours > 30 * e_ref + 2e-3:
                over30.append((k, e_ours, e_ref))
    _log("full_size:" + name, loss=got, loss_fp64=ref, loss_rel_err=rel, fp32_oracle_loss_rel_err=rel32, recon_mse=mse,
         tensors_bad=len(bad), worst_tensor_ratio=worst, grads=n_grads, grads_over_10x=len(over10), grads_over_30x=len(over30),
         worst_grad_ratio=gworst, first=str((bad + over30 + over10)[:3]))
    assert rel <= 1e-5, f"total loss {got!r} vs float64 oracle {ref!r}: rel err {rel:.3e} > 1e-5"
    assert mse <= 1e-4, f"reconstruction MSE {mse:.3e} > 1e-4"
    assert not bad, f"outputs further from the fp64 truth than 10x the fp32 CPU oracle: {bad[:4]}"
    assert not over30, f"gradients beyond 30x the fp32 oracle's error + 2e-3: {over30[:4]}"
    assert len(over10) <= 0.1 * n_grads, f"{len(over10)} of {n_grads} gradients beyond 10x the fp32 oracle's error: {over10[:4]}"


def test_graphed_replays_then_eager_eval_uses_current_weights():
    """After N CUDA-graph replays the weight packs cached for EAGER use must follow the arena (the replay re-packs
    the weights it used and then Adam moves them): an eval-mode forward right after the replays must equal a fresh model
    loaded from ``step.state_dict()``."""
    from playablevideogeneration_b200.training.step import GraphedTrainStep
    case, _ = load_case("full_bair")
    cfg, sd, vgg_sd, obs = case_inputs(case)
    bt = _to_dev(batch_tuple(obs))
    model, step = _build(case, cfg, sd, vgg_sd)
    torch.manual_seed(700); random.seed(700)
    with torch.no_grad():
        model.eval()
        model(bt, ground_truth_observations_init=case["gt_init"], gumbel_temperature=1.0)     # populate the eager pack cache
    graphed = GraphedTrainStep(step, bt, case["gt_init"], 1.0, warmup=1)
    for _ in range(3):
        graphed(bt)
    torch.cuda.synchronize()
    ckpt = step.state_dict()
    fresh, _ = _build(case, cfg, {k: v.detach().cpu() for k, v in ckpt["model"].items()}, None)
    outs = []
    for m in (model, fresh):
        m.eval()
        torch.manual_seed(701); random.seed(701)
        with torch.no_grad():
            outs.append(m(bt, ground_truth_observations_init=case["gt_init"], gumbel_temperature=1.0)[0])
    err = float((outs[0] - outs[1]).abs().max())
    _log("graph_then_eager_eval", max_abs_diff=err)
    assert err == 0.0, err


def test_batched_rollout_step_equals_single():
    """generate_next_batch (configs[4] entry point) must reproduce generate_next sample by sample (eval mode)."""
    case, _ = load_case("rollout_bair")
    cfg, sd, _, obs = case_inputs(case)
    model, _ = _build(case, cfg, sd, None)
    model.eval()
    obs = obs.to(DEV)
    with torch.no_grad():
        model.start_inference()
        f1, _ = model.generate_next(obs, 3)
        model.dynamics_network.reinit_memory(4)
        fb, nb = model.generate_next_batch(obs.unsqueeze(0).repeat(4, 1, 1, 1), torch.tensor([3, 1, 3, 0], device=DEV))
    assert float((fb[0] - f1).abs().max()) <= 1e-5 and float((fb[2] - f1).abs().max()) <= 1e-5
    assert float((fb[1] - f1).abs().max()) > 1e-4


def test_graphed_step_matches_eager_steps():
    """The CUDA-graph replay of the whole optimiser step must reproduce the kernel-by-kernel (eager) step: same seeds ->
    same losses over three steps (2e-5 rel: the only difference is the order of fp32 atomics, amplified by the sign-like
    first Adam updates)."""
    from playablevideogeneration_b200.training.step import GraphedTrainStep
    case, _ = load_case("full_bair")
    cfg, sd, vgg_sd, obs = case_inputs(case)
    bt = _to_dev(batch_tuple(obs))
    _, eager = _build(case, cfg, sd, vgg_sd)
    ref_losses = []
    for s in range(3):
        torch.manual_seed(500 + s); random.seed(500 + s)
        total, _ = eager.step(bt, case["gt_init"], 0.9)
        ref_losses.append(float(total.cpu()[0]))
    _, gs = _build(case, cfg, sd, vgg_sd)
    torch.manual_seed(500); random.seed(500)
    graphed = GraphedTrainStep(gs, bt, case["gt_init"], 0.9, warmup=1)       # warm-up = eager step 0
    got = []
    for s in (1, 2):
        torch.manual_seed(500 + s); random.seed(500 + s)
        total, info = graphed(bt)
        got.append(float(total.cpu()[0]))
    _log("graph_vs_eager", eager=ref_losses, graphed=got)
    for a, b in zip(got, ref_losses[1:]):
        assert abs(a - b) <= 2e-5 * abs(b), (got, ref_losses)


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_graphed_rollout_equals_eager_and_golden(name):
    """Opt-in CUDA-graph inference (Model.enable_graphed_inference): one graph replay per generated frame must give the
    frames of the kernel-by-kernel path bit for bit (same kernels, same order) - hence the reference's - over two rollouts
    (start_inference resets the static recurrent memory), without moving the CPU RNG stream."""
    case, g = load_case(name)
    cfg, sd, _, obs0 = case_inputs(case)

    def rollout(model):
        frames = []
        obs = obs0.to(DEV)
        torch.manual_seed(case["noise_seed"])
        with torch.no_grad():
            model.start_inference()
            for a in case["actions"]:
                f, obs = model.generate_next(obs, a, noise=case.get("noise", False))
                frames.append(f.clone())
            for a1, a2, fac in case.get("interp", []):
                f, obs = model.generate_next_interpolation(obs, a1, a2, fac)
                frames.append(f.clone())
        return frames, torch.get_rng_state()

    model, _ = _build(case, cfg, sd, None)
    model.eval()
    eager, rng_eager = rollout(model)
    model.enable_graphed_inference()
    for attempt in range(2):                                   # the second rollout reuses the captured graph
        graphed, rng_graphed = rollout(model)
        assert torch.equal(rng_eager, rng_graphed), "graphed inference moved the CPU RNG stream"
        for i, (a, b) in enumerate(zip(eager, graphed)):
            assert torch.equal(a, b), (attempt, i, float((a - b).abs().max()))
    keys = [f"frame.{i}" for i in range(len(case["actions"]))] + [f"iframe.{i}" for i in range(len(case.get("interp", [])))]
    for k, f in zip(keys, graphed):
        assert float(np.abs(f.cpu().numpy() - g[k]).max()) <= ROLLOUT_TOL, k


def test_evaluation_metrics_match_reference_golden():
    """The evaluator's cheap metrics on the device (VGG cosine similarity through the tensor-core conv kernels) against the
    values of the unmodified reference classes (tests/golden/metrics.npz, oracle/make_metric_golden.py)."""
    from playablevideogeneration_b200.evaluation.metrics import MSE, PSNR, MotionMaskedMSE, VGGCosineSimilarity
    from playablevideogeneration_b200.vgg import Vgg19
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz"))
    gen_ = torch.Generator().manual_seed(123)
    shape = (2, 4, 3, 32, 48)
    ref = torch.rand(shape, generator=gen_)
    gen = (ref + 0.1 * torch.randn(shape, generator=gen_)).clamp(0, 1)
    ref, gen = ref.to(DEV), gen.to(DEV)
    np.testing.assert_allclose(MSE()(ref, gen).cpu().numpy(), g["mse"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(PSNR()(ref, gen).cpu().numpy(), g["psnr"], rtol=1e-5)
    np.testing.assert_allclose(MotionMaskedMSE()(ref, gen).cpu().numpy(), g["motion_masked_mse"], rtol=1e-5, atol=1e-9)
    vcs = VGGCosineSimilarity(Vgg19(O.make_vgg_weights()).to(DEV))
    np.testing.assert_allclose(vcs(ref, gen).cpu().numpy(), g["vgg_cosine"], rtol=1e-4)


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["mode"] == "eval"])
def test_eval_forward_with_samplers_matches_reference_golden(name):
    """build_evaluation_dataset.py path on the device: eval-mode forward with the evaluation sampler plug-ins against the
    unmodified reference's outputs (argmax-selected actions must agree exactly; tensors to 1e-3)."""
    from playablevideogeneration_b200.evaluation.samplers import OneHotActionSampler, ZeroActionVariationSampler, frames_to_uint8_hwc
    case, g = load_case(name)
    cfg, sd, _, obs = case_inputs(case)
    model, _ = _build(case, cfg, sd, None)
    model.eval()
    torch.manual_seed(case["noise_seed"]); random.seed(case["noise_seed"])
    with torch.no_grad():
        res = model(_to_dev(batch_tuple(obs)), ground_truth_observations_init=case["gt_init"], action_sampler=OneHotActionSampler(),
                    action_variation_sampler=ZeroActionVariationSampler(), gumbel_temperature=case["gumbel_temperature"])
    compare_results(g, RESULT_NAMES_FULL, res, rtol=1e-3, atol=1e-3)
    frames = frames_to_uint8_hwc(torch.cat([obs[:, 0:1, 0:3].to(DEV), res[0]], dim=1))
    assert frames.dtype == torch.uint8 and tuple(frames.shape) == (case["B"], case["T"], case["H"], case["W"], 3)


def test_input_pipeline_kernel_matches_reference_transform():
    """ops.frames_from_uint8 (crop + ToTensor + Normalize on the device) against the unmodified reference transform's output
    (tests/golden/input_pipeline.npz): bit-identical."""
    from playablevideogeneration_b200 import ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_pipeline.npz"))
    frames = torch.from_numpy(g["frames"]).to(DEV)
    out = ops.frames_from_uint8(frames, crop=[int(v) for v in g["crop"]])
    assert tuple(out.shape) == g["out"].shape and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), g["out"])
    full = ops.frames_from_uint8(frames)                                              # no crop: the whole frame
    want = ((torch.from_numpy(g["frames"]).float() / 255.0) - 0.5) / 0.5
    assert torch.equal(full.cpu(), want.permute(0, 3, 1, 2))
    with pytest.raises(Exception):
        ops.frames_from_uint8(frames, crop=[0, 0, 1000, 10])


def test_input_pipeline_resize_matches_pillow_bit_for_bit():
    """PIL crop + resize(BILINEAR) + ToTensor + Normalize on the device (pvg_resample_u8 x 2 + pvg_frames_u8_to_nhwc) against the
    unmodified reference transform (tests/golden/input_resize.npz): the resized uint8 frames and the normalised tensors are
    bit-identical - shrinking (antialiased), growing, one axis only, and BAIR's 64 x 64 -> 256 x 256."""
    from playablevideogeneration_b200 import ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_resize.npz"))
    for name in ("down", "up", "mixed", "bair"):
        frames = torch.from_numpy(g[f"{name}.frames"]).to(DEV)
        crop = None if g[f"{name}.crop"][0] < 0 else [int(v) for v in g[f"{name}.crop"]]
        size = [int(v) for v in g[f"{name}.size"]]
        u8 = ops.resize_frames_uint8(frames, crop, size)
        assert np.array_equal(u8.cpu().numpy(), g[f"{name}.u8"]), name
        out = ops.frames_from_uint8(frames, crop=crop, size=size)
        assert np.array_equal(out.cpu().numpy(), g[f"{name}.out"]), name

