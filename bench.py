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
    # name: (config kind, reduced, H, W, S, B per GPU, T, gt_init)
    "bair256_b8_t16": dict(config="bair", reduced=False, H=256, W=256, S=1, B=8, T=16, gt_init=6),
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
    return ap.parse_args()


def rollout_secondary(dev):
    """Secondary metric of SURVEY.md 8d (BASELINE configs[4] family): eval-mode autoregressive rollout at BAIR 256x256,
    E -> R -> D per generated frame.  Batch 64 kernel by kernel, batch 1 (play.py's own case) kernel by kernel and with one
    CUDA graph per step.  CUDA events around the timed steps, after 3 warm-up steps."""
    import torch
    from oracle.cases import build_config
    from playablevideogeneration_b200.caddy import Model
    cfg = build_config(dict(config="bair", H=256, W=256, S=1))
    torch.manual_seed(0)
    model = Model(cfg).to(dev).eval()
    g = torch.Generator().manual_seed(0)
    out = {}
    with torch.no_grad():
        for batch, graphed, steps in ((64, False, 20), (1, False, 50), (1, True, 50)):
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


def _best_thread_count():
    """torch's CPU convolutions do not scale to every core of a 128-core host (the full-width run is slower than 16
    threads), so the CPU arm uses the thread count that is fastest on a representative conv fwd+bwd."""
    import torch
    import torch.nn.functional as F
    total = os.cpu_count() or 1
    x = torch.randn(4, 128, 64, 64, requires_grad=True)
    wt = torch.randn(128, 128, 3, 3, requires_grad=True)
    best, best_t = 1, float("inf")
    n = 1
    cands = []
    while n < total:
        cands.append(n); n *= 2
    cands.append(total)
    for n in cands:
        torch.set_num_threads(n)
        F.conv2d(x, wt, padding=1).sum().backward()
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, wt, padding=1).sum().backward()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    return best


def cpu_reference_step_time(w, frames, steps, warmup):
    """The reference algorithm (CPU oracle port: oracle/caddy_oracle.py, pinned against the unmodified reference) on
    all host cores: forward + all losses + backward + Adam on a B=1 slice of the workload."""
    import torch
    from oracle import caddy_oracle as O
    from oracle.cases import build_config
    cores = _best_thread_count()
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
    mi = O.MutualInformation(cfg["data"]["actions_count"], cfg["training"]["mutual_information_estimation_alpha"])
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
                       f"{mean:.2f} s/step, torch CPU fp32, {cores} threads (fastest of 1..{os.cpu_count()} on a conv microbenchmark)"), mean


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
    from oracle.cases import build_config            # config dict only (plain data; no oracle compute on this path)
    from playablevideogeneration_b200 import _lib, ops
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
    vgg = Vgg19()
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
    for _ in range(args.steps):
        total, _ = run_host()                                              # H2D of this step's inputs + the step
        loss_host = float(total.cpu()[0])                                  # D2H of the step's result
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3) / args.steps)
    # ---- one more step launched kernel by kernel with CUDA events around every tensor-core conv / wgrad launch: the
    #      per-kernel durations behind `roofline` (events cannot be timed inside a graph replay); also counts launches ----
    ops.conv_profile, ops.wgrad_profile = [], []
    launches0 = _lib.launch_count
    torch.cuda._sleep(int(2.5e9))      # ~1.3 s of GPU spin: lets the host run ahead so that events bracket pure kernel time
    step.step(resident, w["gt_init"], 1.0)
    barrier()
    launches = (_lib.launch_count - launches0) * args.steps
    prof, ops.conv_profile = ops.conv_profile, None
    wprof, ops.wgrad_profile = ops.wgrad_profile, None
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for t in host)

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv) from CUDA events recorded around its launches --
    peaks = load_peaks()
    roof = None
    if prof:
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in prof)
        flops = sum(f for _, _, f in prof)
        ach = flops / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
        wg = None
        if wprof:
            wms = sum(a.elapsed_time(b) for a, b, _ in wprof)
            wg = dict(kernel="conv_wgrad_umma_kernel", launches_per_step=len(wprof), ms_per_step=wms,
                      achieved=sum(f for _, _, f in wprof) / (wms * 1e-3) / 1e12 if wms > 0 else 0.0, unit="TFLOP/s")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
        if args.precision == "tf32x3" and args.workload == "bair256_b8_t16" and os.path.isfile(tpath):
            traffic = json.load(open(tpath))["traffic_bytes_per_launch"]      # from the committed ncu capture of this command
        roof = dict(bound="tensor", kernel="conv_umma_kernel (tcgen05 kind::tf32" + (" main product + 2 correction products: " + json.dumps(ops._corr) if args.precision == "tf32x3" else "") + ")",
                    achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"], traffic=traffic,
                    traffic_source="profiles/r01_conv_traffic.json (ncu dram__bytes_read+write per launch, averaged over the step's launches)" if traffic else None,
                    launches_per_step=len(prof), avg_launch_us=tot_ms * 1e3 / len(prof), ms_per_step=tot_ms,
                    share_of_step=tot_ms / ms_dev, peak_source=peaks["source"], weight_gradient=wg,
                    measured="CUDA events around each launch in one extra eager (non-graph) step after the timed region",
                    note="achieved = algorithmic conv FLOPs (2*N*H*W*Cout*R*S*Cin, unpadded) / CUDA-event time of the launches; "
                         "peak is the measured bf16 cuBLAS figure - kind::tf32 tops out at half of it; the fp32-equivalent split costs "
                         "1 TF32 + 2 16-bit MMAs (= 2 TF32 MMA times) per product, i.e. a ceiling of a quarter of the peak")
    if rank != 0:
        _hard_exit(world)
        return
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_step_time(w, args.cpu_sample_frames, 1, 0)
    line = dict(metric=METRIC, value=frames_per_step / (ms_dev * 1e-3), unit="frames/s",
                n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_dev, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype={"tf32x3": "f32 (fp32-equivalent split product on the tensor cores: TF32 main term + two correction terms, fp32 accumulate)", "tf32": "tf32",
                                         "fp32": "f32"}[args.precision],
                data="synthetic", impl="pvg_b200",
                config=dict(workload=args.workload, per_gpu_batch=w["B"], seq_len=w["T"], frame=f"{w['H']}x{w['W']}x3",
                            gt_init=w["gt_init"], parallelism=f"dp{world}", precision=args.precision,
                            corrections=dict(ops._corr) if args.precision == "tf32x3" else None,
                            launch="cuda-graph replay of the whole step" if use_graph else "eager (one launch per kernel from Python)",
                            l2="inputs (100.7 MB/step) and per-step activations (>10 GB) exceed the 126 MB L2; no explicit flush"),
                e2e=dict(value=frames_per_step / (ms_e2e * 1e-3), unit="frames/s", h2d_bytes_per_step=h2d * world,
                         d2h_bytes_per_step=8 * world, ms_per_step=ms_e2e, last_loss=loss_host),
                gpu_launches=launches, clocks=clocks, roofline=roof, cpu_baseline=cpu_base)
    if world == 1 and not args.no_secondary:
        # after every headline number is final: a failure here cannot touch them
        try:
            line["secondary"] = rollout_secondary(dev)
        except BaseException as e:          # noqa: BLE001 - report, never lose the headline line
            line["secondary"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(line), flush=True)
    _hard_exit(world)


if __name__ == "__main__":
    main()
