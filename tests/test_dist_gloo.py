"""N > 1 host logic on CPU: world_size 2, gloo.  Each rank owns a contiguous shard of the
utterances; per-shard results (computed here by the CPU oracle, standing in for the GPU path)
are gathered in utterance order, and the shared transition gradient is all-reduced."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gtn_b200 import dist as gd
    from oracle import pyoracle as po
    from tests import util
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    first, count = gd.shard(rank, world, total)
    T, C, U = 20, 6, 3
    rng = np.random.default_rng(7)
    tw = rng.uniform(-1, 1, C + C * C).astype(np.float32)
    e, tg = [], []
    for i in range(first, first + count):  # per-utterance seeds: shards are reproducible
        r = np.random.default_rng(1234 + i)
        e.append(r.uniform(-5, 5, (T, C)).astype(np.float32))
        tg.append(r.integers(1, C, U).astype(np.int32))
    losses = np.array([po.ctc_loss(e[k], tg[k], 0, True, want_grad=False)[0] for k in range(count)], np.float32)
    tgrad = np.zeros(C + C * C, np.float32)
    for k in range(count):
        tgrad += po.asg_loss(e[k], tw, tg[k])[2]
    all_losses = gd.gather_losses(losses, total)
    tsum = gd.allreduce_shared_grad(tgrad)
    if rank == 0:
        q.put((all_losses, tsum))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_everything():
    from gtn_b200 import dist as gd
    for total in (0, 1, 5, 256, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            spans = [gd.shard(r, world, total) for r in range(world)]
            assert spans[0][0] == 0
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert spans[-1][0] + spans[-1][1] == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


@pytest.mark.timeout(180)
def test_two_ranks_gloo():
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    total, world, port = 5, 2, 29571
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    all_losses, tsum = q.get(timeout=150)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process answer
    T, C, U = 20, 6, 3
    tw = np.random.default_rng(7).uniform(-1, 1, C + C * C).astype(np.float32)
    want, tw_sum = [], np.zeros(C + C * C, np.float32)
    for i in range(total):
        r = np.random.default_rng(1234 + i)
        e = r.uniform(-5, 5, (T, C)).astype(np.float32)
        tg = r.integers(1, C, U).astype(np.int32)
        want.append(po.ctc_loss(e, tg, 0, True, want_grad=False)[0])
        tw_sum += po.asg_loss(e, tw, tg)[2]
    assert np.allclose(all_losses, np.asarray(want, np.float32), rtol=1e-6)
    assert np.allclose(tsum, tw_sum, rtol=1e-5, atol=1e-6)
