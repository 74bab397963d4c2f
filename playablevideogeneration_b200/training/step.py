"""One optimiser step of CADDY training on the pvg_b200 kernels: forward + every loss + backward + gradient
all-reduce + Adam.  Mirrors ``Trainer.compute_losses`` / ``compute_losses_pretraining`` and the step at
training/trainer.py:241-550,584-587 (same loss weights, same float64 accumulators, same ``loss_info`` keys for the
quantities on the hot path) - with the trainer's ~50 per-scalar ``.item()`` host syncs replaced by ONE packed
device->host copy, and all parameters / gradients / Adam moments living in flat arenas so that the optimiser is one
kernel launch and data parallelism is one NCCL all-reduce per step.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..vgg import Vgg19
from . import losses as L


class FlatArena:
    """Moves every trainable parameter of ``module`` into one flat fp32 buffer (and its gradient into another)."""

    ALIGN = 64

    def __init__(self, module: nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dev = self.params[0].device
        self.flat = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros((total,), dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                self.flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        self.offsets = offs
        self.sizes = [(p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN for p in self.params]
        # torch.optim.Adam skips parameters whose .grad is None (never reached by backward) and keeps a per-parameter
        # step count; post-accumulate hooks record which parameters each backward reached.
        self.touched = set()
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self.touched.add(i))
        self.steps = [0] * len(self.params)
        self._index_of = {p.data_ptr(): i for i, p in enumerate(self.params)}

    def param_at(self, data_ptr: int):
        i = self._index_of.get(data_ptr)
        return None if i is None else self.params[i]

    def touch(self, param) -> None:
        """A gradient was written into ``param``'s slice of the gradient arena outside autograd (ops.wgrad_defer.flush)."""
        i = self._index_of.get(param.data_ptr())
        if i is not None:
            self.touched.add(i)

    def zero_grad(self):
        self.touched.clear()
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):          # re-attach (autograd may have replaced .grad objects)
            g = self.grad[o:o + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g


class TrainStep:
    def __init__(self, config: dict, model: nn.Module, vgg: Optional[Vgg19] = None, process_group=None):
        self.config = config
        self.model = model
        self.module = model.module if hasattr(model, "module") else model
        tr = config["training"]
        if tr.get("use_motion_weights", False):
            # trainer.py:436-440 builds a motion weight mask; the reference's own weighted branch is broken (losses.py:105) and no
            # shipped config enables it - refuse rather than train silently without the mask
            raise NotImplementedError("training.use_motion_weights is not supported on the pvg_b200 hot path")
        dev = next(self.module.parameters()).device
        self.vgg = (vgg if vgg is not None else Vgg19()).to(dev)
        self.perceptual_loss = L.ParallelPerceptualLoss(self.vgg)
        self.observations_loss = L.ObservationsLoss()
        self.states_loss = L.StatesLoss()
        self.hidden_states_loss = L.HiddenStatesLoss()
        self.entropy_loss = L.EntropyLogitLoss()
        self.action_state_distribution_kl = L.KLGeneralGaussianDivergenceLoss()
        self.action_directions_kl_gaussian_divergence_loss = L.KLGaussianDivergenceLoss()
        if "smooth" in tr.get("trainer", ""):
            self.mutual_information_loss = L.SmoothMutualInformationLoss(config).to(dev)
        else:
            self.mutual_information_loss = L.MutualInformationLoss()
        self.mi_lambda = tr.get("action_mutual_information_entropy_lambda", 1.0)
        self.arena = FlatArena(self.module)
        self.exp_avg = torch.zeros_like(self.arena.flat)
        self.exp_avg_sq = torch.zeros_like(self.arena.flat)
        self.lr = tr["learning_rate"]
        self.weight_decay = tr["weight_decay"]
        self.lr_schedule = list(tr.get("lr_schedule", []))
        self.lr_gamma = tr.get("lr_gamma", 1.0)
        self.global_step = 0
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        if self.world > 1:
            # the two batch-coupled statistics: MI joint matrix and centroid EMA see the global batch
            self.mutual_information_loss.process_group = process_group
            self.module.centroid_estimator.process_group = process_group
        self._capturing = False
        self._graph_runs = []
        self._hyper_dev_buf = torch.zeros((16, 7), dtype=torch.float32, device=dev)
        self._hyper_dev = self._hyper_dev_buf
        self._hyper_host = [torch.zeros((16, 7), dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros((16, 7))
                            for _ in range(4)]
        self._hyper_events = [None] * len(self._hyper_host)
        self._hyper_slot = 0

    # ----------------------------------------------------------------------------------------------------------
    def current_lr(self) -> float:
        n = sum(1 for m in self.lr_schedule if self.global_step >= m)
        return self.lr * (self.lr_gamma ** n)

    def compute_losses(self, batch_tuple, ground_truth_observations_count: int, gumbel_temperature: float,
                       pretraining: bool = False):
        """-> (total_loss (1,) float64, loss_info {name: 0-dim tensor}, model results)."""
        cfg = self.config
        lw = cfg["training"]["loss_weights"]
        sfx = "_pretraining" if pretraining else ""
        observations = batch_tuple[0]
        t = observations.shape[1]
        if ground_truth_observations_count >= t:
            ground_truth_observations_count = t - 1
        if pretraining:
            results = self.model(batch_tuple, pretraining=True, gumbel_temperature=gumbel_temperature)
            (rec, pyr, rec_states, states, rec_hidden, hidden, selected, logits, samples, attention, dir_dist, sampled_dirs,
             state_dist, sampled_states, variations, r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates) = results
        else:
            results = self.model(batch_tuple, ground_truth_observations_count, gumbel_temperature=gumbel_temperature)
            (rec, pyr, rec_states, states, hidden, selected, logits, samples, attention, rec_att, dir_dist, sampled_dirs,
             state_dist, sampled_states, variations, r_logits, r_dir_dist, r_sdirs, r_state_dist, r_sstates) = results
        dev = observations.device
        perceptual = torch.zeros((1,), dtype=torch.float64, device=dev)
        perceptual_term = torch.zeros((1,), dtype=torch.float64, device=dev)
        rec_loss = torch.zeros((1,), dtype=torch.float64, device=dev)
        info: Dict[str, torch.Tensor] = {}
        lam_p = lw["perceptual_loss_lambda" + sfx]
        for r, cur in enumerate(pyr):
            p_tot, p_levels = self.perceptual_loss(observations, cur)
            # Trainer.sum_loss_components (trainer.py:167-186): one weight per VGG level, or a scalar broadcast to all of them
            lams = list(lam_p) if isinstance(lam_p, (list, tuple)) else [lam_p] * len(p_levels)
            term = p_levels[0] * 0.0
            for lv, lam in zip(p_levels, lams):
                term = term + lv * lam
            o = self.observations_loss(observations, cur)
            perceptual = perceptual + p_tot
            perceptual_term = perceptual_term + term
            rec_loss = rec_loss + o
            info[f"perceptual_loss_r{r}"] = p_tot.detach()
            info[f"observations_rec_loss_r{r}"] = o.detach()
            for li, lv in enumerate(p_levels):
                info[f"perceptual_loss_r{r}_l{li}"] = lv.detach()
        n = len(pyr)
        perceptual, perceptual_term, rec_loss = perceptual / n, perceptual_term / n, rec_loss / n
        states_rec = self.states_loss(states.detach(), rec_states)
        entropy = self.entropy_loss(logits)
        kl_dir = self.action_directions_kl_gaussian_divergence_loss(dir_dist)
        mi = self.mutual_information_loss(torch.softmax(logits, dim=-1), torch.softmax(r_logits, dim=-1), lamb=self.mi_lambda)
        kl_state = self.action_state_distribution_kl(r_state_dist, state_dist.detach())
        total = (lw["reconstruction_loss_lambda" + sfx] * rec_loss + perceptual_term
                 + lw["states_rec_lambda" + sfx] * states_rec + lw["entropy_lambda" + sfx] * entropy
                 + lw["action_directions_kl_lambda" + sfx] * kl_dir
                 + lw["action_mutual_information_lambda" + sfx] * mi
                 + lw["action_state_distribution_kl_lambda" + sfx] * kl_state)
        if self.world > 1:
            # every rank holds the GLOBAL MI value but differentiates it through its local samples only, and gradients are
            # averaged over ranks: scale this term's gradient (not its value) by the world size
            lam = lw["action_mutual_information_lambda" + sfx]
            total = total + (self.world - 1) * lam * (mi - mi.detach())
        if pretraining:
            hid_rec = self.hidden_states_loss(hidden, rec_hidden.detach())
            total = total + lw["hidden_states_rec_lambda_pretraining"] * hid_rec
            info["hidden_states_rec_loss"] = hid_rec.detach()
        info.update({
            "avg_observations_rec_loss": rec_loss.detach()[0], "avg_perceptual_loss": perceptual.detach()[0],
            "loss_component_perceptual_loss": perceptual_term.detach()[0], "states_rec_loss": states_rec.detach(),
            "entropy_loss": entropy.detach(), "action_directions_kl_loss": kl_dir.detach(),
            "action_mutual_information_loss": mi.detach(), "action_state_distribution_kl_loss": kl_state.detach(),
            "total_loss": total.detach()[0]})
        return total, info, results

    @staticmethod
    def fetch_info(info: Dict[str, torch.Tensor]) -> Dict[str, float]:
        """ONE device->host copy for all logged scalars (the reference issues one .item() sync per scalar)."""
        keys = list(info)
        packed = torch.stack([info[k].reshape(()).double() for k in keys]).cpu()
        return {k: float(v) for k, v in zip(keys, packed)}

    def _adam_runs(self):
        """Runs of consecutive arena parameters that received a gradient and share a step count: one Adam launch each
        (steady state: one or two launches for the whole model).  torch.optim.Adam skips parameters without a gradient."""
        a = self.arena
        runs, i, n = [], 0, len(a.params)
        while i < n:
            if i not in a.touched:
                i += 1
                continue
            j = i
            while j + 1 < n and (j + 1) in a.touched and a.steps[j + 1] == a.steps[i]:
                j += 1
            runs.append((i, j, a.offsets[i], a.offsets[j] + a.sizes[j]))
            i = j + 1
        return runs

    def _hyper(self, step: int, lr: float):
        b1, b2 = 0.9, 0.999
        return [lr / (1.0 - b1 ** step), math.sqrt(1.0 - b2 ** step), b1, b2, 1e-8, self.weight_decay, 1.0 / self.world]

    def optimizer_step(self):
        if self.pg is not None and self.world > 1:
            torch.distributed.all_reduce(self.arena.grad, group=self.pg)       # one NCCL all-reduce over NVLink / NVSwitch
        a = self.arena
        if self._capturing:
            # CUDA-graph capture: record the launches only; counters and hyper-parameters are advanced by prepare_replay()
            self._graph_runs = self._adam_runs()
            self._hyper_dev = self._hyper_dev_buf[:len(self._graph_runs)]
            for k, (i, j, lo, hi) in enumerate(self._graph_runs):
                ops.adam_step_dev(a.flat[lo:hi], a.grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], self._hyper_dev[k])
            return
        # the reference runs optimizer.step() and THEN lr_scheduler.step() (trainer.py:586-587): step s uses the rate after s - 1
        # scheduler ticks, i.e. the one current before the counter moves
        lr = self.current_lr()
        self.global_step += 1
        for i, j, lo, hi in self._adam_runs():
            for k in range(i, j + 1):
                a.steps[k] += 1
            ops.adam_step(a.flat[lo:hi], a.grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], a.steps[i], lr,
                          weight_decay=self.weight_decay, grad_scale=1.0 / self.world)
        ops.invalidate_weight_cache()

    def prepare_replay(self):
        """Host work before a graph replay: advance the step counters and upload the Adam scalars of every run."""
        lr = self.current_lr()               # rate of THIS step (before the scheduler tick, see optimizer_step)
        self.global_step += 1
        a = self.arena
        rows = []
        for i, j, lo, hi in self._graph_runs:
            for k in range(i, j + 1):
                a.steps[k] += 1
            rows.append(self._hyper(a.steps[i], lr))
        # pinned staging rows are a ring: the asynchronous upload of step k may still be queued behind a 250 ms replay when the
        # host prepares step k + 1, so a row is rewritten only after the copy that read it has completed
        slot = self._hyper_slot
        self._hyper_slot = (slot + 1) % len(self._hyper_host)
        if self._hyper_events[slot] is not None:
            self._hyper_events[slot].synchronize()
        host = self._hyper_host[slot]
        host[:len(rows)].copy_(torch.tensor(rows, dtype=torch.float32))
        self._hyper_dev.copy_(host[:len(rows)], non_blocking=True)
        if host.is_pinned():
            ev = torch.cuda.Event()
            ev.record()
            self._hyper_events[slot] = ev

    # ---- checkpoint interoperability: the reference's latest.pth.tar (training/trainer.py:80-122, ------------------
    #      training/smooth_mi_trainer.py:23-68) = {"model", "optimizer", "lr_scheduler", "step"[, "mi_estimator"]} ----------
    def _adam_index(self) -> List[int]:
        """Position of every arena parameter in ``model.parameters()`` - the index torch.optim.Adam's state_dict uses (the
        reference hands ALL parameters to Adam, trainer.py:36, including the gradient-free centroid buffer-parameter)."""
        where = {id(p): i for i, p in enumerate(self.module.parameters())}
        return [where[id(p)] for p in self.arena.params]

    def optimizer_state_dict(self) -> dict:
        """The flat-arena Adam state in ``torch.optim.Adam(model.parameters(), ...).state_dict()`` layout (parameter index =
        position in ``model.parameters()``, which follows the reference's registration order; parameters that never
        received a gradient have no entry, exactly like torch.optim.Adam)."""
        a = self.arena
        state = {}
        for k, (p, o, i) in enumerate(zip(a.params, a.offsets, self._adam_index())):
            if a.steps[k] > 0:
                n = p.numel()
                state[i] = {"step": torch.tensor(float(a.steps[k])),
                            "exp_avg": self.exp_avg[o:o + n].view(p.shape).detach().clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).detach().clone()}
        group = {"lr": self.current_lr(), "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "initial_lr": self.lr, "params": list(range(sum(1 for _ in self.module.parameters())))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        """Accepts what ``torch.optim.Adam.state_dict()`` produced for the same model (``step`` as int - PyTorch 1.4, the
        reference's pin - or as a tensor)."""
        a = self.arena
        saved_ids = [i for g in sd["param_groups"] for i in g["params"]]
        n_all = sum(1 for _ in self.module.parameters())
        if len(saved_ids) != n_all:
            raise ValueError(f"optimizer state covers {len(saved_ids)} parameters, the model has {n_all}")
        pos = {saved: k for k, saved in enumerate(saved_ids)}            # saved id -> position in model.parameters()
        arena_of = {i: k for k, i in enumerate(self._adam_index())}      # position in model.parameters() -> arena slot
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        a.steps = [0] * len(a.params)
        for saved, st in sd["state"].items():
            i = pos[int(saved)]
            if i not in arena_of:
                raise ValueError(f"optimizer state for parameter {i}, which does not require a gradient")
            k = arena_of[i]
            p, o = a.params[k], a.offsets[k]
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {saved}: shape {tuple(st['exp_avg'].shape)} does not match {tuple(p.shape)}")
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            a.steps[k] = int(float(st["step"]))

    def lr_scheduler_state_dict(self) -> dict:
        """``MultiStepLR(optimizer, milestones, gamma).state_dict()`` after ``global_step`` scheduler steps."""
        import collections
        return {"milestones": collections.Counter(int(m) for m in self.lr_schedule), "gamma": self.lr_gamma, "base_lrs": [self.lr],
                "last_epoch": self.global_step, "_step_count": self.global_step + 1, "_get_lr_called_within_step": False,
                "_last_lr": [self.current_lr()]}

    def state_dict(self) -> dict:
        """The reference trainer's checkpoint dictionary."""
        ckpt = {"model": self.module.state_dict(), "optimizer": self.optimizer_state_dict(),
                "lr_scheduler": self.lr_scheduler_state_dict(), "step": self.global_step}
        if isinstance(self.mutual_information_loss, L.SmoothMutualInformationLoss):      # smooth_mi_trainer.py:43-45
            ckpt["mi_estimator"] = self.mutual_information_loss.state_dict()
        return ckpt

    def load_state_dict(self, ckpt: dict) -> None:
        self.module.load_state_dict(ckpt["model"])
        ops.invalidate_weight_cache()
        self.load_optimizer_state_dict(ckpt["optimizer"])
        self.global_step = int(ckpt["step"])
        if "mi_estimator" in ckpt and isinstance(self.mutual_information_loss, L.SmoothMutualInformationLoss):
            self.mutual_information_loss.load_state_dict(ckpt["mi_estimator"])

    def save_checkpoint(self, path: str) -> None:
        torch.save(self.state_dict(), path)

    def load_checkpoint(self, path: str) -> None:
        self.load_state_dict(torch.load(path, map_location=self.arena.flat.device, weights_only=False))

    def step(self, batch_tuple, ground_truth_observations_count: int, gumbel_temperature: float, pretraining: bool = False):
        self.module.train()
        ops.zero_pool.begin(self.arena.flat.device)      # one memset for every accumulator scratch of this step
        try:
            total, info, _ = self.compute_losses(batch_tuple, ground_truth_observations_count, gumbel_temperature, pretraining)
            self.arena.zero_grad()
            ops.wgrad_defer.begin()                      # one weight-gradient scratch per weight for the whole backward pass
            try:
                total.backward()
            finally:
                ops.wgrad_defer.flush(self.arena.touch, self.arena.param_at)
            self.optimizer_step()
        finally:
            ops.zero_pool.end()
        return total.detach(), info


class GraphedTrainStep:
    """The whole optimiser step (forward, losses, backward, all-reduce, Adam) captured once in a CUDA graph and replayed:
    ~5 000 kernel launches per step are submitted by ONE cudaGraphLaunch instead of by the Python interpreter.

    The eager warm-up steps and the capture itself are REAL optimiser steps on ``example_batch`` (parameters, Adam moments and
    ``global_step`` advance ``warmup`` times; the capture pass records launches only): build the graph with the first batch
    of the run, or snapshot / restore ``TrainStep.state_dict()`` around the constructor.

    Per replay the host only (1) copies the batch into the static input buffers, (2) redraws the step's random numbers
    on the CPU generator in the reference's order and uploads them (NoiseSource.refill), (3) uploads Adam's bias
    corrections / learning rate.  Shapes, ``ground_truth_observations_count`` and the Gumbel temperature are fixed per
    graph (re-capture when the trainer's schedule changes them)."""

    def __init__(self, step: TrainStep, example_batch, ground_truth_observations_count: int, gumbel_temperature: float,
                 pretraining: bool = False, warmup: int = 2):
        self.step = step
        self._staging = self._staged = self._copy_stream = self._copy_done = self._staging_free = None
        self.args = (ground_truth_observations_count, gumbel_temperature, pretraining)
        self.static_batch = tuple(t.clone() for t in example_batch)
        noise = step.module.noise
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(max(1, warmup)):            # eager warm-up: lazy inits, kernel attributes, allocator pools
                noise.mode = "record"
                noise.begin_step()
                step.step(self.static_batch, *self.args)
            noise.freeze()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        step._capturing = True
        try:
            with torch.cuda.graph(self.graph):
                noise.begin_step()
                self.static_total, self.static_info = step.step(self.static_batch, *self.args)
        finally:
            step._capturing = False
            # the static buffers are only needed while the launches are recorded; ``refill`` keeps uploading into them.  Outside
            # a replay the model draws on the CPU generator again (an eager validation forward must not read a training
            # step's frozen noise)
            noise.mode = "cpu"

    def prefetch(self, batch) -> None:
        """Starts the host->device copy of the NEXT step's batch on a side stream, into a staging buffer: it overlaps the replay
        that is running (the graph reads ``static_batch``, which is not touched).  ``__call__(batch)`` with the same batch object
        then only waits for that copy and moves the data device-to-device (100 MB: ~0.05 ms) in front of the replay."""
        if self._staging is None:
            self._staging = tuple(torch.empty_like(t) for t in self.static_batch)
            self._copy_stream = torch.cuda.Stream()
            self._copy_done = torch.cuda.Event()
        # Wait for the device-to-device copy OUT of the staging buffer only (it sits in front of the last replay) - not for the
        # replay itself: waiting on the whole stream serialised every upload behind the running step (e2e - device gap of 6.5 ms
        # = the 100 MB upload, round-2 bench).
        if self._staging_free is not None:
            self._copy_stream.wait_event(self._staging_free)
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._staging, batch):
                dst.copy_(src, non_blocking=True)
            self._copy_done.record(self._copy_stream)
        self._staged = batch
        # the next step's random numbers too: drawn now (same CPU RNG order), uploaded to a staging buffer on the side stream
        self.step.module.noise.stage(self._copy_stream)

    def __call__(self, batch=None):
        if batch is not None:
            if self._staged is batch:                    # uploaded by prefetch() while the previous step ran
                torch.cuda.current_stream().wait_event(self._copy_done)
                for dst, src in zip(self.static_batch, self._staging):
                    dst.copy_(src, non_blocking=True)
                self._staging_free = torch.cuda.Event()
                self._staging_free.record()
                self._staged = None
            else:
                for dst, src in zip(self.static_batch, batch):
                    dst.copy_(src, non_blocking=True)
        self.step.module.noise.commit()              # (draws +) staging -> live noise buffers, in front of the replay
        self.step.prepare_replay()
        self.graph.replay()
        # The replay re-packed the weights it USED (W_k) into its own pack buffers and then Adam wrote W_{k+1}: anything that
        # runs outside the graph afterwards (validation, generate_next, an eager tail step) must re-pack from the arena.
        ops.invalidate_weight_cache()
        return self.static_total, self.static_info
