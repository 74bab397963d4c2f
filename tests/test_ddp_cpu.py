"""N>1 host path on CPU: two gloo ranks, kernels replaced by the test-only CPU stand-ins (tests/fake_ops.py).
Checks the data-parallel contract of SURVEY.md 8e: identical replicas, batch sharded, ONE all-reduce of the flat gradient
arena per step, gradient averaging folded into Adam -> parameters stay bit-identical across ranks, and the reduced
gradient equals the mean of the per-rank gradients."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import caddy_oracle as O
from oracle.cases import CASES, build_config
from tests.golden_util import batch_tuple


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import fake_ops
    fake_ops.install(_Patch())
    from playablevideogeneration_b200.caddy import Model
    from playablevideogeneration_b200.training.step import TrainStep
    from playablevideogeneration_b200.vgg import Vgg19
    case = dict(CASES["full_bair"])
    cfg = build_config(case)
    model = Model(cfg)
    model.load_state_dict(O.make_weights(cfg, 0))
    step = TrainStep(cfg, model, Vgg19(O.make_vgg_weights()), process_group=dist.group.WORLD)
    obs = O.make_observations(2, 4, 3, 64, 64, seed=100 + rank)          # each rank its own shard
    torch.manual_seed(7 + rank); random.seed(7 + rank)
    model.train()
    total, info, _ = step.compute_losses(batch_tuple(obs), 3, 1.0)
    step.arena.zero_grad()
    total.backward()
    local_grad = step.arena.grad.clone()
    step.optimizer_step()
    gathered = [torch.zeros_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    mean_grad = sum(gathered) / world
    params = [torch.zeros_like(step.arena.flat) for _ in range(world)]
    dist.all_gather(params, step.arena.flat)
    if rank == 0:
        torch.save(dict(sync=bool(all(torch.equal(params[0], p) for p in params[1:])),
                        reduced_is_sum=bool(torch.allclose(step.arena.grad, mean_grad * world, rtol=1e-6, atol=1e-9)),
                        ranks_differ=bool(not torch.equal(gathered[0], gathered[1]))), out)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_step_keeps_replicas_in_sync(tmp_path):
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ranks_differ"], "shards should produce different local gradients"
    assert r["reduced_is_sum"], "arena.grad after the all-reduce must be the sum of the per-rank gradients (Adam scales by 1/world)"
    assert r["sync"], "parameters diverged across ranks"


def _mi_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from playablevideogeneration_b200.training.losses import MutualInformationLoss
    g = torch.Generator().manual_seed(3)
    p1 = torch.softmax(torch.randn(4, 5, 7, generator=g), -1)
    p2 = torch.softmax(torch.randn(4, 5, 7, generator=g), -1)
    full = MutualInformationLoss()
    a = p1.clone().requires_grad_(True)
    ref = full(a, p2)                                      # "gathered batch" on one device (the reference's semantics)
    ref.backward()
    sharded = MutualInformationLoss()
    sharded.process_group = dist.group.WORLD
    sl = slice(rank * 2, rank * 2 + 2)
    b = p1[sl].clone().requires_grad_(True)
    got = sharded(b, p2[sl])
    got.backward()
    # the MI value is a cancellation-dominated O(1e-2) number: compare on the scale of its terms (|P log P| ~ 0.1)
    ok = bool(abs(float(got) - float(ref)) <= 2e-7) and bool(torch.allclose(b.grad, a.grad[sl], rtol=1e-4, atol=2e-7))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        torch.save(dict(ok=all(flags)), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_mutual_information_equals_gathered_batch(tmp_path):
    """The 7x7 joint-matrix all-reduce makes the per-rank MI loss (value and local gradients) identical to the loss on
    the gathered batch that the reference's DataParallel trainer computes."""
    out = str(tmp_path / "mi.pt")
    mp.spawn(_mi_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["ok"]
