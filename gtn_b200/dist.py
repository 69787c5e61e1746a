"""Multi-GPU plumbing: one process per GPU (torch.distributed), the minibatch split by
utterance -- the reference's only parallelism is the independent-utterance map
(gtn/parallel/parallel_map.h:153-188), so there is no data-path collective for CTC:

  * shard(rank, world, B)      contiguous utterance range of a rank
  * gather_losses(x)           all_gather of the per-utterance losses, in utterance order
  * allreduce_shared_grad(x)   sum of the shared-graph gradient (ASG transitions), the
                               multi-GPU form of Graph::addGrad under grad_lock (graph.cpp:96-100)

Backend: "nccl" on GPUs, "gloo" in the CPU tests."""
import numpy as np


def shard(rank, world, total):
    """-> (first, count): contiguous split, remainders to the lowest ranks."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_losses(local, total, device=None):
    """local: 1-D float32 array of this rank's losses -> all `total` losses in utterance order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    counts = [shard(r, world, total)[1] for r in range(world)]
    width = max(counts) if counts else 0
    buf = torch.zeros(width, dtype=torch.float32, device=device)
    buf[:len(local)] = torch.as_tensor(np.asarray(local, np.float32), device=device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)])


def allreduce_shared_grad(local, device=None):
    """Sum a shared-graph gradient (e.g. ASG transitions, C + C*C floats) over the ranks."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(local, np.float32), device=device).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
