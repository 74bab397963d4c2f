#!/usr/bin/env python
"""bench.py - frames/s of one CADDY training step (forward + all losses + backward + gradient all-reduce + Adam).

Workload (BASELINE.json configs[1]): BAIR 256x256, seq_len 16, batch 8 per GPU, full CADDY (E+R+A+D), ground-truth
context 6 frames, L1 + VGG19-perceptual (3 resolutions) + states MSE + KL + smooth-MI losses, synthetic U[-1,1] frames,
reference-default random init (no network for datasets / checkpoints).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this implementation (hand-written sm_100a kernels)
    python bench.py --impl reference --steps 2 --warmup 1         # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0).  `value` = whole-job frames/s with inputs resident in HBM; `e2e` = the same step
driven through the public API from pinned HOST buffers (H2D of the batch + D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (train step, 256x256x3, seq=16)"

WORKLOADS = {
    # name: config kind (configs/0N_*.yaml), reduced model?, frame H x W, observation stacking S, sequences per GPU B, frames T,
    # ground-truth context gt_init.  BASELINE.json configs[1] is the headline; configs[2] (Breakout: 208 x 160 frames, reduced
    # model, 26 x 20 state maps - BASELINE's "160x160" is a simplification) and configs[3] (Tennis: 96 x 256 frames,
    # S = 4, batch 32 over 8 GPUs = 4 per GPU) are selectable; configs[0] is the reference's CPU-runnable case.
    "bair256_b8_t16": dict(config="bair", reduced=False, H=256, W=256, S=1, B=8, T=16, gt_init=6),
    "breakout208x160_b16_t32": dict(config="breakout", reduced=True, H=208, W=160, S=1, B=16, T=32, gt_init=6),
    "tennis96x256_b4_t16": dict(config="tennis", reduced=False, H=96, W=256, S=4, B=4, T=16, gt_init=6),
    "bair64_b2_t4": dict(config="bair", reduced=False, H=64, W=64, S=1, B=2, T=4, gt_init=3),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pvg_b200", choices=["pvg_b200", "reference"])
    ap.add_argument("--workload", default="bair256_b8_t16", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32x3", choices=["tf32x3", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample-frames", type=int, default=8, help="frames (B=1 x T) of the CPU baseline sample")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary (rollout) measurements")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true", help="skip the PyTorch-eager + cuDNN leg (same box, same batch)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the pre-run loss check against the CPU oracle")
    ap.add_argument("--layer-table", default=None, help="write a per-shape table of the tensor-core conv launches of one step (markdown)")
    return ap.parse_args()


def rollout_secondary(dev):
    """Secondary metric of SURVEY.md 8d (BASELINE configs[4] family): eval-mode autoregressive rollout at BAIR 256x256,
    E -> R -> D per generated frame.  Batch 64 kernel by kernel, batch 1 (play.py's own case) kernel by kernel and with one
    CUDA graph per step.  CUDA events around the timed steps, after 3 warm-up steps."""
    import torch
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.configs import build_config
    cfg = build_config(dict(config="bair", H=256, W=256, S=1))
    torch.manual_seed(0)
    model = Model(cfg).to(dev).eval()
    g = torch.Generator().manual_seed(0)
    out = {}
    with torch.no_grad():
        for batch, graphed, steps in ((64, False, 100), (1, False, 50), (1, True, 50)):      # configs[4]: batch 64, 100 steps
            model.enable_graphed_inference(graphed)
            obs = (torch.rand((batch, 3, 256, 256), generator=g) * 2 - 1).to(dev)
            actions = torch.randint(0, 7, (steps + 3, batch), generator=g).to(dev)
            model.start_inference()
            model.dynamics_network.reinit_memory(batch)
            for t in range(3):
                _, obs = model.generate_next_batch(obs, actions[t])
            model.start_inference()
            model.dynamics_network.reinit_memory(batch)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(steps):
                frames, obs = model.generate_next_batch(obs, actions[3 + t])
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[f"rollout_b{batch}" + ("_cuda_graph" if graphed else "")] = dict(
                ms_per_generated_step=ms, frames_per_s=batch / (ms * 1e-3), steps=steps, finite=bool(torch.isfinite(frames).all()))
    out["workload"] = "BAIR 256x256 eval-mode rollout (generate_next_batch), fp32-equivalent split product"
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured (bf16 cuBLAS, sustained)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


def synthetic_batch(w, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    obs = torch.rand((w["B"], w["T"], 3 * w["S"], w["H"], w["W"]), generator=g) * 2.0 - 1.0
    return (obs, torch.zeros((w["B"], w["T"]), dtype=torch.int32), torch.zeros((w["B"], w["T"])),
            torch.zeros((w["B"], w["T"]), dtype=torch.bool))


def _cpu_threads():
    """Threads of the CPU arm: every host core up to 32 (torch's CPU convolutions stop scaling beyond that on the 128-core
    hosts of this pool).  Fixed, not auto-tuned: round 1's timing-based pick moved the denominator by 35 % between boxes."""
    return max(1, min(32, os.cpu_count() or 1))


def _mi_alpha(cfg):
    """EMA factor of the smooth mutual-information estimator (training/smooth_mi_trainer.py); None for the plain trainer."""
    return cfg["training"]["mutual_information_estimation_alpha"] if "smooth" in cfg["training"]["trainer"] else None


def gpu_eager_baseline(w, dev, steps=3, warmup=2):
    """The bar of SURVEY.md 8(d) / BASELINE.md 3.4: the reference ALGORITHM as plain PyTorch eager ops (F.conv2d -> cuDNN,
    F.batch_norm, autograd, the trainer's loss sum, Adam) on the same B200, same batch, ``cudnn.benchmark = True`` as train.py:19
    sets it.  The oracle port stands in for the reference files (they do not travel to the GPU box); it IS the reference's op
    sequence, pinned op for op by tests/test_oracle_golden.py.  Two numbers: TF32 convolutions allowed (what the reference gets
    by default on any Ampere+ GPU) and strict fp32 (``allow_tf32 = False``: the precision this repo's fp32-equivalent split
    product delivers).  Reported next to ``cpu_baseline``; nothing on the product path touches it."""
    import torch
    from oracle import caddy_oracle as O
    from playablevideogeneration_b200.configs import build_config
    cfg = build_config(dict(config=w["config"], H=w["H"], W=w["W"], S=w["S"]))
    sd = O.make_weights(cfg, 0, w["reduced"])
    vgg_sd = {k: v.to(dev) for k, v in O.make_vgg_weights().items()}
    bt = tuple(t.to(dev) for t in synthetic_batch(w))
    out = dict(kind="port (oracle/caddy_oracle.py on cuda: PyTorch eager + cuDNN, cudnn.benchmark=True)", batch=w["B"], seq_len=w["T"],
               steps=steps, warmup=warmup)
    prev = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.benchmark = True
        for label, tf32 in (("tf32", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            params = {k: (v.to(dev).requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
                          and "centroid" not in k else v.to(dev)) for k, v in sd.items()}
            train = [v for v in params.values() if v.requires_grad]
            m = [torch.zeros_like(p) for p in train]
            v2 = [torch.zeros_like(p) for p in train]
            mi = O.MutualInformation(cfg["data"]["actions_count"], _mi_alpha(cfg))
            mi.matrix = mi.matrix.to(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            loss = None
            with torch.device(dev):                   # the oracle draws its noise with bare torch.randn / torch.rand
                for it in range(warmup + steps):
                    if it == warmup:
                        torch.cuda.synchronize(dev)
                        e0.record()
                    total, _, _ = O.compute_losses(params, vgg_sd, cfg, mi, bt, w["gt_init"], 1.0)
                    for p in train:
                        p.grad = None
                    total.backward()
                    with torch.no_grad():
                        live = [(p, p.grad, a, b) for p, a, b in zip(train, m, v2) if p.grad is not None]
                        O.adam_step([x[0] for x in live], [x[1] for x in live], [x[2] for x in live], [x[3] for x in live], it + 1, 4e-4, 1e-6)
                    loss = total.detach()
                e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[label] = dict(ms_per_step=ms, frames_per_s=w["B"] * w["T"] / (ms * 1e-3), last_loss=float(loss.cpu()[0]))
            del params, train, m, v2, total, loss
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def parity_check(w, cfg, model, step, vgg, dev):
    """Before anything is timed: one forward + losses of THIS model on a B=1, T=4 slice of the bench batch at the bench's frame
    size, through the CUDA path and through the fp32 CPU oracle with the same weights, inputs and noise seeds.  The bench line
    carries both losses; a relative difference above the contract's 1e-5 fails the run."""
    import random
    import torch
    from oracle import caddy_oracle as O
    t = min(4, w["T"])
    bt = synthetic_batch(dict(w, B=1, T=t), seed=7)
    gt = min(2, t - 1)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    vgg_sd = {f"features.{i}.{n}": getattr(vgg.convs[str(i)], n).detach().cpu().clone() for i in vgg.convs for n in ("weight", "bias")}
    mi = O.MutualInformation(cfg["data"]["actions_count"], _mi_alpha(cfg))
    mil = step.mutual_information_loss
    mil_sd = {k: v.detach().clone() for k, v in mil.state_dict().items()} if hasattr(mil, "state_dict") else None
    torch.set_num_threads(_cpu_threads())
    torch.manual_seed(11); random.seed(11)
    with torch.no_grad():
        ref_total, _, _ = O.compute_losses(sd, vgg_sd, cfg, mi, bt, gt, 1.0)
    model.train()
    torch.manual_seed(11); random.seed(11)
    with torch.no_grad():
        total, _, _ = step.compute_losses(tuple(x.to(dev) for x in bt), gt, 1.0)
    got, ref = float(total.cpu()[0]), float(ref_total[0])
    # the check ran a train-mode forward: put the BatchNorm / EMA state back so the timed run starts from the seeded model
    model.load_state_dict({k: v.to(dev) for k, v in sd.items()})
    if mil_sd is not None:
        step.mutual_information_loss.load_state_dict(mil_sd)
    from playablevideogeneration_b200 import ops
    ops.invalidate_weight_cache()
    return dict(sample=f"B=1 x T={t} of {w['H']}x{w['W']}, gt_init={gt}, this model's weights", cuda_loss=got, oracle_fp32_loss=ref,
                rel_err=abs(got - ref) / abs(ref), tolerance=1e-5)


# algorithmic HBM bytes of the bandwidth-bound entry points (fp32 NHWC tensors read / written once per launch), from their
# C-ABI arguments (include/pvg_b200.h); the classes bench.py reports under "roofline_hbm"
def _hbm_bytes(name, a):
    f = 4
    if name == "pvg_bn_stats":                       # x, N, HW, C
        return "bn_stats", a[1] * a[2] * a[3] * f
    if name == "pvg_pool2_stats":                    # x, N, H, W, C, y: read x, write y / 4
        return "bn_stats", a[1] * a[2] * a[3] * a[4] * f * 1.25
    if name == "pvg_bn_finalize_apply":              # x, N, HW, C, ..., residual at 15
        return "bn_apply", a[1] * a[2] * a[3] * f * (3 if a[15] else 2)
    if name == "pvg_bn_apply":                       # x, N, HW, C, groups, mean, invstd, w, b, residual
        return "bn_apply", a[1] * a[2] * a[3] * f * (3 if a[9] else 2)
    if name == "pvg_bn_finalize_apply_ex":           # x, N, HW, C, groups, ..., residual at 15, y at 18, planes_a at 19, planes_b at 21
        return "bn_apply", a[1] * a[2] * a[3] * (f * (3 if a[15] else 2) + (4 if a[19] else 0) + (4 if a[21] else 0))
    if name == "pvg_bn_apply_ex":                    # x, N, HW, C, groups, mean, invstd, w, b, residual, act, slope, y, planes_a, fmt, planes_b
        return "bn_apply", a[1] * a[2] * a[3] * (f * (3 if a[9] else 2) + (4 if a[13] else 0) + (4 if a[15] else 0))
    if name == "pvg_bn_bwd_reduce":                  # dy, y, x, N, HW, C, ..., act at 9
        return "bn_bwd", a[3] * a[4] * a[5] * f * (3 if a[9] else 2)
    if name in ("pvg_bn_bwd_apply", "pvg_bn_bwd_apply_ex"):   # dy, y, x, N, H, W, C, groups, mean, invstd, weight, act, slope, sums2, eval, unpool, dx, g_out
        full = a[3] * a[4] * a[5] * a[6] * f
        small = full / 4 if a[15] else full
        return "bn_bwd", small * (3 if a[11] else 2) + full + (small if a[17] else 0)
    if name == "pvg_split_16":                       # x, planes, n
        return "split_16", a[2] * 8
    if name == "pvg_act_bwd_split_16":               # dy, y, act, slope, g, planes, n
        return "split_16", a[6] * 16
    if name == "pvg_split_16_scaled":                # x, planes, n: read 4, write 2 x 2
        return "split_16", a[2] * 8
    if name == "pvg_act_bwd_split_16_scaled":        # dy, y, act, slope, g, planes, n: read 8, write 4 (+ 4 for g)
        return "split_16", a[6] * (16 if a[4] else 12)
    if name == "pvg_amax":                           # x, n
        return "amax", a[1] * 4
    if name == "pvg_act_bwd":
        return "split_16", a[5] * 12
    if name == "pvg_absdiff_mean_fwd":               # a, b, N, count
        return "absdiff", a[2] * a[3] * 8
    if name == "pvg_absdiff_mean_bwd":               # a, b, gout, N, count, db
        return "absdiff", a[3] * a[4] * 12
    if name == "pvg_upsample2x_fwd":                 # x, N, H, W, C
        return "resample", a[1] * a[2] * a[3] * a[4] * f * 5
    if name == "pvg_upsample2x_bwd":
        return "resample", a[1] * a[2] * a[3] * a[4] * f * 5
    if name == "pvg_resize_bilinear":                # x, N, H, W, C, y, OH, OW
        return "resample", a[1] * a[4] * f * (a[2] * a[3] + a[6] * a[7])
    if name == "pvg_maxpool2_fwd":
        return "maxpool", a[1] * a[2] * a[3] * a[4] * f * 1.25
    if name == "pvg_maxpool2_fwd_ex":                # ..., y, planes_a: + 4 bytes per output element
        return "maxpool", a[1] * a[2] * a[3] * a[4] * (f * 1.25 + (1 if a[6] else 0))
    if name == "pvg_resize_bilinear_ex":             # x, N, H, W, C, y, OH, OW, planes_a, fmt_a, planes_b
        return "resample", a[1] * a[4] * (f * a[2] * a[3] + a[6] * a[7] * (f + (4 if a[8] else 0) + (4 if a[10] else 0)))
    if name == "pvg_lstm_bwd_act":                   # gates_act, c_prev, c_new, dh, dc_new, M, C
        return "lstm_pointwise", a[5] * a[6] * 52
    if name == "pvg_maxpool2_bwd":                   # dy, x, y, N, H, W, C
        return "maxpool", a[3] * a[4] * a[5] * a[6] * f * 2.5
    if name == "pvg_lstm_fwd":                       # gates, c_prev, M, C
        return "lstm_pointwise", a[2] * a[3] * 28
    if name == "pvg_lstm_bwd":                       # gates, c_prev, c_new, dh, dc, M, C
        return "lstm_pointwise", a[5] * a[6] * 52
    if name in ("pvg_adam_step", "pvg_adam_step_dev"):
        return "adam", a[4] * 28
    return None, 0


def _write_layer_table(path, prof, wprof, shapes, peaks, ms_step, workload):
    """Per-shape aggregate of the tensor-core convolution launches of one (eager, event-timed) step: time, algorithmic TFLOP/s
    and the two floors that bound the shape - HBM (operand planes in, fp32 result and its planes out, at the measured
    bandwidth) and tensor (3 kind::f16 MMAs per product at the measured bf16 rate)."""
    agg = {}
    for ent in list(prof) + [(a, b, f, "wgrad") for a, b, f in wprof]:
        a, b, fl = ent[0], ent[1], ent[2]
        sh = shapes.get(id(a))
        if sh is None:
            continue
        role, n, h, w, cin, cout, r, planes, sums = sh
        px = n * h * w
        if role == "wgrad":
            nbytes = px * 4.0 * (cin + cout)
        elif role == "lstm":
            nbytes = px * (4.0 * cin + 4.0 * cout + 3 * 4.0 * (cout // 4))
        else:
            nbytes = px * (4.0 * cin + 4.0 * cout + (4.0 * cout if planes else 0.0))
        e = agg.setdefault(sh, [0, 0.0, 0.0, 0.0, 1e30, 0.0])
        ms_ = a.elapsed_time(b)
        e[0] += 1; e[1] += ms_; e[2] += fl; e[3] += nbytes; e[4] = min(e[4], ms_); e[5] = max(e[5], ms_)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    tot = sum(v[1] for _, v in rows)
    with open(path, "w") as f:
        f.write(f"# Tensor-core convolution launches of one training step, by shape ({workload})\n\n")
        f.write("`python bench.py --layer-table ...`: CUDA events around every launch of one eager step after the timed region "
                f"(graph replay of the whole step: {ms_step:.1f} ms).  floor = max(HBM floor, tensor floor): HBM floor = operand planes read + "
                f"fp32 result (+ its planes) written at {peaks['hbm_gbs']:.0f} GB/s; tensor floor = 3 x algorithmic FLOPs at "
                f"{peaks['tflops']:.0f} TFLOP/s (measured bf16 cuBLAS).  {len(rows)} shapes, {sum(v[0] for _, v in rows)} launches, {tot:.1f} ms.\n\n")
        f.write("| role | N | HxW | Cin | Cout | k | planes | BN sums | launches | ms | us/launch | min us | max us | TFLOP/s | HBM floor ms | tensor floor ms | floor / time |\n")
        f.write("|---|---:|---|---:|---:|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for (role, n, h, w, cin, cout, r, planes, sums), (cnt, ms, fl, nb, lo_, hi_) in rows:
            hb = nb / (peaks["hbm_gbs"] * 1e9) * 1e3
            tf = 3.0 * fl / (peaks["tflops"] * 1e12) * 1e3
            f.write(f"| {role} | {n} | {h}x{w} | {cin} | {cout} | {r} | {'y' if planes else ''} | {'y' if sums else ''} | {cnt} | {ms:.3f} | "
                    f"{ms * 1e3 / cnt:.1f} | {lo_ * 1e3:.1f} | {hi_ * 1e3:.1f} | {fl / (ms * 1e-3) / 1e12:.1f} | {hb:.3f} | {tf:.3f} | {max(hb, tf) / ms:.2f} |\n")


def cpu_reference_step_time(w, frames, steps, warmup):
    """The reference algorithm (CPU oracle port: oracle/caddy_oracle.py, pinned against the unmodified reference) on
    all host cores: forward + all losses + backward + Adam on a B=1 slice of the workload."""
    import torch
    from oracle import caddy_oracle as O
    from playablevideogeneration_b200.configs import build_config
    cores = _cpu_threads()
    torch.set_num_threads(cores)
    cfg = build_config(dict(config=w["config"], H=w["H"], W=w["W"], S=w["S"]))
    t = max(3, min(w["T"], frames))
    sd = O.make_weights(cfg, 0, w["reduced"])
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
                  and "centroid" not in k else v.clone()) for k, v in sd.items()}
    train = [v for v in params.values() if v.requires_grad]
    m = [torch.zeros_like(p) for p in train]
    v2 = [torch.zeros_like(p) for p in train]
    vgg_sd = O.make_vgg_weights()
    mi = O.MutualInformation(cfg["data"]["actions_count"], _mi_alpha(cfg))
    bt = synthetic_batch(dict(w, B=1, T=t))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        total, _, _ = O.compute_losses(params, vgg_sd, cfg, mi, bt, min(w["gt_init"], t - 1), 1.0)
        for p in train:
            p.grad = None
        total.backward()
        with torch.no_grad():
            live = [(p, p.grad, a, b) for p, a, b in zip(train, m, v2) if p.grad is not None]
            O.adam_step([x[0] for x in live], [x[1] for x in live], [x[2] for x in live], [x[3] for x in live], s + 1, 4e-4, 1e-6)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return dict(value=t / mean, unit="frames/s", cores=cores, kind="port",
                sample=f"B=1 x T={t} of {w['H']}x{w['W']} (one sequence of the batch), {len(times)} timed step(s), "
                       f"{mean:.2f} s/step, torch CPU fp32, {cores} threads of {os.cpu_count()} host cores"), mean


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, mean = cpu_reference_step_time(w, args.cpu_sample_frames, max(1, args.steps), max(0, args.warmup))
    line = dict(impl="reference", metric=METRIC, value=cb["value"], unit="frames/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=mean * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=args.workload, per_gpu_batch=w["B"], seq_len=w["T"], frame=f"{w['H']}x{w['W']}x3",
                            gt_init=w["gt_init"],
                            note="reference algorithm (CPU oracle port) on the host cores; each step is a bounded sample of "
                                 "the workload: " + cb["sample"]),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def _hard_exit(world):
    """Multi-rank runs leave through os._exit: tearing down NCCL communicators that are referenced by a captured CUDA graph
    at interpreter shutdown was observed to hang (rank 0 printed its line, then the job sat until the driver's timeout)."""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from playablevideogeneration_b200 import _lib, ops
    from playablevideogeneration_b200.configs import build_config
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.training.step import GraphedTrainStep, TrainStep
    from playablevideogeneration_b200.vgg import Vgg19

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    ops.set_precision(args.precision)

    cfg = build_config(dict(config=w["config"], H=w["H"], W=w["W"], S=w["S"]))
    torch.manual_seed(0); random.seed(0)
    model = Model(cfg, reduced=w["reduced"]).to(dev)
    vgg = Vgg19(allow_random_init=True)               # no network for the ImageNet checkpoint: seeded He-init weights, timing only
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():                                # He-init stand-in for the ImageNet weights (no network)
        for conv in vgg.convs.values():
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (conv.weight.shape[1] * 9)) ** 0.5)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.05)
    step = TrainStep(cfg, model, vgg, process_group=pg)
    host = synthetic_batch(w, seed=rank)
    host = tuple(t.pin_memory() for t in host)
    resident = tuple(t.to(dev) for t in host)
    frames_per_step = w["B"] * w["T"] * world
    parity = None
    if world == 1 and not args.no_parity_check and args.precision == "tf32x3":
        parity = parity_check(w, cfg, model, step, vgg, dev)
        if parity["rel_err"] > parity["tolerance"]:
            raise SystemExit(f"bench.py: the CUDA path disagrees with the CPU oracle before timing: {parity}")
    torch.manual_seed(100 + rank); random.seed(100 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    use_graph = not args.no_graph
    if use_graph:
        # eager warm-up (inside GraphedTrainStep) + capture of the whole step; W more replays warm the graph itself
        gstep = GraphedTrainStep(step, resident, w["gt_init"], 1.0, warmup=2)
        run_resident = lambda: gstep()
        run_host = lambda: gstep(host)
    else:
        run_resident = lambda: step.step(resident, w["gt_init"], 1.0)
        run_host = lambda: step.step(tuple(t.to(dev, non_blocking=True) for t in host), w["gt_init"], 1.0)
    for _ in range(args.warmup):
        run_resident()
    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_resident()
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # ---- timed region 2: end to end through the public API from pinned host memory ---------------------------------
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    loss_host = 0.0
    # Every step: H2D of its inputs from pinned host memory, the step, D2H of its loss into pinned host memory.  The loss of
    # step k is READ by the host after step k + 1 has been launched (one step of slack, as a training loop that logs
    # asynchronously would do), so the graph-launch latency is not serialised behind a host sync; all copies and the final
    # read are inside the timed region.
    loss_pinned = torch.zeros((2, 1), dtype=torch.float64).pin_memory()
    loss_events = [None, None]
    if use_graph:
        gstep.prefetch(host)                                               # H2D of step 0's inputs (inside the timed region)
    loss_host = 0.0
    for it in range(args.steps):
        total, _ = run_host()                                              # (rest of the) H2D of this step's inputs + the step
        loss_pinned[it % 2].copy_(total.detach().reshape(1), non_blocking=True)      # D2H of the step's result
        ev = torch.cuda.Event(); ev.record(); loss_events[it % 2] = ev
        if use_graph and it + 1 < args.steps:
            gstep.prefetch(host)                                           # next step's H2D overlaps this step's replay
        if it > 0:
            loss_events[(it - 1) % 2].synchronize()
            loss_host = float(loss_pinned[(it - 1) % 2][0])                # the previous step's loss, now on the host
    loss_events[(args.steps - 1) % 2].synchronize()
    loss_host = float(loss_pinned[(args.steps - 1) % 2][0])
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3) / args.steps)
    if use_graph and world == 1:
        # the captured graph's private memory pool (104 GB at Breakout B=16 x T=32) is not needed any more: the eager profiling
        # step below allocates its activations from the ordinary pool and would not fit next to it on the largest workloads
        del total
        gstep = run_resident = run_host = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    # ---- one more step launched kernel by kernel with CUDA events around EVERY C-ABI call: the per-kernel durations behind
    #      `roofline` (tensor-core convs, from their algorithmic FLOPs) and `roofline_hbm` (bandwidth-bound kernel classes,
    #      from their algorithmic bytes); events cannot be timed inside a graph replay.  Also counts launches. ----------------
    ops.conv_profile, ops.wgrad_profile = [], []
    hbm_events = []
    orig_call = _lib.call

    def timed_call(name, *cargs):
        cls, nbytes = _hbm_bytes(name, cargs)
        if cls is None:
            return orig_call(name, *cargs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.profile_spin()
        a.record()
        rc = orig_call(name, *cargs)
        b.record()
        hbm_events.append((cls, a, b, float(nbytes)))
        return rc

    _lib.call = ops.call = timed_call
    from playablevideogeneration_b200.training import losses as _losses_mod
    _losses_mod.ops.call = timed_call
    launches0 = _lib.launch_count
    # Events must bracket pure kernel time, but an eager step is host-bound (~17 k launches from Python): a 150 us spin kernel
    # is queued in front of every bracketed launch, so that the start event, the kernel and the end event are all in the queue
    # before the GPU reaches them.
    ops.profile_spin_cycles = 300_000
    try:
        step.step(resident, w["gt_init"], 1.0)
    finally:
        _lib.call = ops.call = orig_call
        ops.profile_spin_cycles = 0
    barrier()
    launches = (_lib.launch_count - launches0) * args.steps
    prof, ops.conv_profile = ops.conv_profile, None
    wprof, ops.wgrad_profile = ops.wgrad_profile, None
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for t in host)

    # ---- roofline of the dominant kernel family (tcgen05 implicit-GEMM convs) from CUDA events recorded around its launches,
    #      per family: conv_h3_kernel (forward, all-fp16 split) and conv_umma*_kernel (data gradient, TF32 main + bf16 corrections)
    peaks = load_peaks()
    roof = None
    roof_hbm = None
    if args.layer_table and rank == 0 and (prof or wprof):
        _write_layer_table(args.layer_table, prof or [], wprof or [], ops.profile_shapes, peaks, ms_dev, args.workload)
    if prof:
        fams = {}
        for a_, b_, f_, kind in prof:
            e = fams.setdefault(kind, [0, 0.0, 0.0])
            e[0] += 1; e[1] += a_.elapsed_time(b_); e[2] += f_
        names = {"h3": "conv_h3_kernel (forward convs and data gradients: 3 kind::f16 tcgen05 MMAs per product on fp16 plane pairs, halo reuse, persistent, CTA pairs)",
                 "tf32": "conv_umma_persistent / conv_umma2_persistent (legacy path: kind::tf32 main product + 2 bf16 corrections)",
                 "single": "conv_umma_kernel (single TF32 product)"}
        ceil_note = {"h3": "3 f16 MMAs per product: ceiling 1/3 of the bf16 peak", "tf32": "1 TF32 + 2 16-bit MMAs per product (2 TF32 MMA times): ceiling 1/4 of the bf16 peak",
                     "single": "TF32: ceiling 1/2 of the bf16 peak"}
        ceil_frac = {"h3": 1.0 / 3.0, "tf32": 0.25, "single": 0.5}
        tpath = os.path.join(ROOT, "profiles", "r02_conv_traffic.json")
        tdata = json.load(open(tpath)) if (args.precision == "tf32x3" and args.workload == "bair256_b8_t16" and os.path.isfile(tpath)) else {}
        fam_rows = []
        for kind, (n_, ms_, fl_) in sorted(fams.items(), key=lambda kv: -kv[1][1]):
            ach = fl_ / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
            fam_rows.append(dict(bound="tensor", kernel=names.get(kind, kind), achieved=ach, peak=peaks["tflops"], unit="TFLOP/s",
                                 frac=ach / peaks["tflops"], frac_of_arithmetic_ceiling=ach / (peaks["tflops"] * ceil_frac.get(kind, 1.0)),
                                 ceiling=ceil_note.get(kind), traffic=tdata.get(kind, {}).get("traffic_bytes_per_launch"),
                                 launches_per_step=n_, avg_launch_us=ms_ * 1e3 / n_, ms_per_step=ms_, share_of_step=ms_ / ms_dev))
        wg = None
        if wprof:
            wms = sum(a_.elapsed_time(b_) for a_, b_, _ in wprof)
            wfl = sum(f_ for _, _, f_ in wprof)
            wg = dict(bound="tensor", kernel="conv_wgrad_umma_kernel (weight gradients: MN-major fp16 plane pairs, 3 kind::f16 MMAs per product, split-K)" if args.precision == "tf32x3" else "conv_wgrad_umma_kernel (weight gradients, split-K)",
                      launches_per_step=len(wprof), ms_per_step=wms, achieved=wfl / (wms * 1e-3) / 1e12 if wms > 0 else 0.0,
                      peak=peaks["tflops"], unit="TFLOP/s", frac=(wfl / (wms * 1e-3) / 1e12 / peaks["tflops"]) if wms > 0 else 0.0,
                      traffic=tdata.get("wgrad", {}).get("traffic_bytes_per_launch"), share_of_step=wms / ms_dev)
        roof = dict(fam_rows[0])
        roof.update(peak_source=peaks["source"], other_tensor_kernels=fam_rows[1:], weight_gradient=wg,
                    traffic_source="profiles/r02_conv_traffic.json (ncu dram__bytes_read+write per launch, averaged over the step's launches of the family)" if roof.get("traffic") else None,
                    measured="CUDA events around each launch in one extra eager (non-graph) step after the timed region; a 150 us spin kernel in front of every bracketed launch keeps the host's launch latency out of the bracket",
                    note="achieved = algorithmic conv FLOPs (2*N*H*W*Cout*R*S*Cin, unpadded) / CUDA-event time of the family's launches; "
                         "peak is the measured bf16 cuBLAS figure; fp32-equivalent results need 3 tensor-core products per multiply")
    if hbm_events:
        cls = {}
        for c_, a_, b_, nb in hbm_events:
            e = cls.setdefault(c_, [0, 0.0, 0.0])
            e[0] += 1; e[1] += a_.elapsed_time(b_); e[2] += nb
        roof_hbm = [dict(kernel_class=c_, bound="hbm", launches_per_step=n_, bytes_per_step=nb, ms_per_step=ms_,
                         achieved=nb / (ms_ * 1e-3) / 1e9 if ms_ > 0 else 0.0, peak=peaks["hbm_gbs"], unit="GB/s",
                         frac=(nb / (ms_ * 1e-3) / 1e9 / peaks["hbm_gbs"]) if ms_ > 0 else 0.0, share_of_step=ms_ / ms_dev)
                    for c_, (n_, ms_, nb) in sorted(cls.items(), key=lambda kv: -kv[1][1])]
    if rank != 0:
        _hard_exit(world)
        return
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_step_time(w, args.cpu_sample_frames, 1, 0)
    metric = METRIC if args.workload == "bair256_b8_t16" else f"frames/sec (train step, {w['H']}x{w['W']}x3, seq={w['T']})"
    line = dict(metric=metric, value=frames_per_step / (ms_dev * 1e-3), unit="frames/s",
                n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_dev, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype={"tf32x3": "f32 (fp32-equivalent split product on the tensor cores: 22-bit fp16 plane pairs, 3 kind::f16 MMAs per product, fp32 accumulate)", "tf32": "tf32",
                                         "fp32": "f32"}[args.precision],
                data="synthetic", impl="pvg_b200",
                config=dict(workload=args.workload, per_gpu_batch=w["B"], seq_len=w["T"], frame=f"{w['H']}x{w['W']}x3",
                            gt_init=w["gt_init"], parallelism=f"dp{world}", precision=args.precision,
                            corrections=dict(ops._corr) if args.precision == "tf32x3" else None,
                            launch="cuda-graph replay of the whole step" if use_graph else "eager (one launch per kernel from Python)",
                            l2=f"per-step activations (>10 GB written and re-read) exceed the 126 MB L2 many times over; inputs are {h2d / 1e6:.1f} MB/step; no explicit flush"),
                e2e=dict(value=frames_per_step / (ms_e2e * 1e-3), unit="frames/s", h2d_bytes_per_step=h2d * world,
                         d2h_bytes_per_step=8 * world, ms_per_step=ms_e2e, last_loss=loss_host),
                gpu_launches=launches, clocks=clocks, roofline=roof, roofline_hbm=roof_hbm, cpu_baseline=cpu_base,
                parity_check=parity)
    if world == 1 and not args.no_secondary:
        # after every headline number is final: a failure here cannot touch them
        try:
            line["secondary"] = rollout_secondary(dev)
        except BaseException as e:          # noqa: BLE001 - report, never lose the headline line
            line["secondary"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_gpu_eager_baseline:
        # last: frees this implementation's graph and buffers first; a failure cannot touch the numbers above
        try:
            del gstep
        except NameError:
            pass
        try:
            del step, model, resident
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_baseline(w, dev)
            for k in ("tf32", "fp32"):
                if k in line["gpu_eager_baseline"]:
                    line["gpu_eager_baseline"][k]["this_repo_speedup"] = line["gpu_eager_baseline"][k]["ms_per_step"] / ms_dev
        except BaseException as e:          # noqa: BLE001
            line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(line), flush=True)
    _hard_exit(world)


if __name__ == "__main__":
    main()
