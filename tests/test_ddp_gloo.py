"""CPU, world_size 2, gloo: the N>1 host logic (per-rank sizing + the single flat-buffer all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightning_pose_b200.ddp import FlatGradAllReducer, per_rank_sizes


def test_per_rank_sizes_match_reference_rule():
    # reference tests/data/test_factory.py:199-272 arithmetic: ceil(B/N), ceil(T/N), ceil((C-4)/N)+4
    assert per_rank_sizes(16, 32, 16, 1) == {"train_batch_size": 16, "sequence_length": 32, "context_batch_size": 16}
    assert per_rank_sizes(16, 32, 16, 8) == {"train_batch_size": 2, "sequence_length": 4, "context_batch_size": 6}
    assert per_rank_sizes(10, 30, 14, 4) == {"train_batch_size": 3, "sequence_length": 8, "context_batch_size": 7}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical initial weights on every rank, like DDP
    model = torch.nn.Sequential(torch.nn.PixelShuffle(2), torch.nn.ConvTranspose2d(4, 3, 3, 2, 1, 1), torch.nn.ConvTranspose2d(3, 3, 3, 2, 1, 1))
    g = torch.Generator().manual_seed(100 + rank)  # different clips per rank
    x = torch.randn(2, 16, 3, 3, generator=g)
    loss = model(x).square().mean()
    loss.backward()
    local = [p.grad.clone() for p in model.parameters()]
    red = FlatGradAllReducer(model.parameters(), n_scalars=2)
    scal = red.step([loss, torch.tensor(float(rank))])
    gathered = [None] * world
    dist.all_gather_object(gathered, [g_.tolist() for g_ in local])
    ok = red.launches == 1
    for i, p in enumerate(model.parameters()):
        mean = sum(torch.tensor(gathered[r][i]) for r in range(world)) / world
        ok = ok and torch.allclose(p.grad, mean, atol=1e-7)
    losses = [None] * world
    dist.all_gather_object(losses, float(loss))
    ok = ok and abs(float(scal[0]) - sum(losses) / world) < 1e-7 and abs(float(scal[1]) - (world - 1) / 2) < 1e-7
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_allreduce_two_ranks_gloo():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _teardown_worker(rank, world, port, out):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out[rank] = True  # reached the end of the "run"; _teardown never returns (os._exit(0) without destructors)
    bench._teardown(dist, rank, world)
    out[rank] = False


@pytest.mark.timeout(120)
def test_bench_teardown_two_ranks_gloo():
    """bench.py's end-of-run rendezvous (store counter, rank 0 leaves last, exit code 0 on every rank): the multi-rank
    run must end cleanly without destroying the process group under the captured CUDA graph."""
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_teardown_worker, args=(world, _free_port(), out), nprocs=world, join=True)  # raises on a non-zero exit
        assert dict(out) == {0: True, 1: True}


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.PixelShuffle(2), torch.nn.ConvTranspose2d(4, 3, 3, 2, 1, 1))
    red = FlatGradAllReducer(model.parameters(), n_scalars=1, extra_floats=1000)
    ok = True
    for it in range(2):  # two steps: the views stay attached, the bucket is re-zeroed
        g = torch.Generator().manual_seed(100 + 10 * it + rank)
        x = torch.randn(2, 16, 3, 3, generator=g)
        red.extra.fill_(float(rank + 1))  # stands in for the backbone's gradient
        red.begin_step()
        loss = model(x).square().mean()
        loss.backward()  # autograd accumulates straight into the flat buffer
        ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
        local = [p.grad.clone() for p in model.parameters()]
        scal = red.finish_step([loss])
        gathered = [None] * world
        dist.all_gather_object(gathered, [g_.tolist() for g_ in local])
        for i, p in enumerate(model.parameters()):
            mean = sum(torch.tensor(gathered[r][i]) for r in range(world)) / world
            ok = ok and torch.allclose(p.grad, mean, atol=1e-7)
        losses = [None] * world
        dist.all_gather_object(losses, float(loss))
        ok = ok and abs(float(scal[0]) - sum(losses) / world) < 1e-7
        ok = ok and torch.allclose(red.extra, torch.full_like(red.extra, (1 + world) / 2))
    ok = ok and red.launches == 4  # per step: backbone bucket + head bucket
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_allreduce_in_place_two_ranks_gloo():
    """begin_step / finish_step: gradients are written into the flat buffer by autograd (no pack/unpack), a second
    bucket (the backbone's) is reduced alongside."""
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
