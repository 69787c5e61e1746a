"""The real callers on GPUs: the ASG criterion as a
torch.autograd.Function on CUDA tensors -- emission AND transition gradients -- and the NCCL all-reduce of the
shared transition gradient across two ranks (the multi-GPU form of Graph::addGrad under grad_lock,
/root/reference/gtn/graph.cpp:96-100).  The two-rank test needs two GPUs and skips otherwise."""
import os
import sys

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(first, count, T, C, U):
    e, tg = [], []
    for i in range(first, first + count):
        r = np.random.default_rng(99 + i)
        e.append(r.uniform(-5, 5, (T, C)).astype(np.float32))
        tg.append(r.integers(0, C, U).astype(np.int32))
    return np.stack(e), tg


def test_torch_asg_loss_on_cuda_tensors(oracle):
    """examples/asg.cpp:59-81 through torch: losses, emission gradients and the transition gradient against
    the oracle, for the mean reduction (what a training loop calls)."""
    torch = pytest.importorskip("torch")
    from gtn_b200 import torch_loss
    B, T, C, U = 6, 40, 12, 5
    e, tg = _inputs(0, B, T, C, U)
    tw = np.random.default_rng(5).uniform(-1, 1, C + C * C).astype(np.float32)
    x = torch.tensor(e, device="cuda", requires_grad=True)
    w = torch.tensor(tw, device="cuda", requires_grad=True)
    loss = torch_loss.asg_loss(x, w, tg, reduction="mean")
    loss.backward()
    want_l, want_g, want_t = [], [], np.zeros_like(tw)
    for b in range(B):
        lo, go, to = oracle.asg_loss(e[b], tw, tg[b])
        want_l.append(lo)
        want_g.append(go / B)
        want_t += to / B
    assert util.close(loss.item(), float(np.mean(want_l)))
    assert util.grad_close(x.grad.cpu().numpy(), np.stack(want_g), 10.0 * T)
    assert util.grad_close(w.grad.cpu().numpy(), want_t, 10.0 * T)
    # transitions kept on the host (a small parameter): the gradient comes back on the host
    w2 = torch.tensor(tw, requires_grad=True)
    torch_loss.asg_loss(torch.tensor(e, device="cuda"), w2, tg, reduction="none").sum().backward()
    assert not w2.grad.is_cuda
    assert util.grad_close(w2.grad.numpy(), want_t * B, 10.0 * T * B)


def _rank(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gtn_b200 import dist as gd, torch_loss
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    T, C, U = 40, 12, 5
    first, count = gd.shard(rank, world, total)
    e, tg = _inputs(first, count, T, C, U)
    tw = np.random.default_rng(5).uniform(-1, 1, C + C * C).astype(np.float32)
    x = torch.tensor(e, device="cuda", requires_grad=True)
    w = torch.tensor(tw, device="cuda", requires_grad=True)
    per = torch_loss.asg_loss(x, w, tg, reduction="none")  # all-reduces the transition gradient over NCCL
    per.sum().backward()
    losses = gd.gather_losses(per.detach().cpu().numpy(), total, device="cuda")
    # the plain helper on a numpy gradient too (gtn_b200/dist.py), over NCCL
    twice = gd.allreduce_shared_grad(np.full(4, rank + 1.0, np.float32), device="cuda")
    if rank == 0:
        q.put((losses, w.grad.cpu().numpy(), twice))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_nccl_shared_transition_gradient(oracle):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    total, world, port = 7, 2, 29581
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    losses, tgrad, twice = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    T, C, U = 40, 12, 5
    e, tg = _inputs(0, total, T, C, U)
    tw = np.random.default_rng(5).uniform(-1, 1, C + C * C).astype(np.float32)
    want_l, want_t = [], np.zeros_like(tw)
    for b in range(total):
        lo, _, to = oracle.asg_loss(e[b], tw, tg[b])
        want_l.append(lo)
        want_t += to
    assert util.close(losses, np.asarray(want_l, np.float32))
    assert util.grad_close(tgrad, want_t, 10.0 * T * total)  # every rank holds the sum over BOTH shards
    assert np.array_equal(twice, np.full(4, 3.0, np.float32))
