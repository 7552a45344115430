"""Data-parallel plumbing (torch.distributed): the ONE collective of the training step and the particle
sharding helpers.  Pure torch, device-agnostic — the same code runs over NCCL on the GPUs and over gloo in
the CPU tests (tests/test_parallel_gloo.py).

Reference semantics (pyprob/nn/inference_network.py:296-333, :448, :529-530): every rank draws its own
minibatch, gradients are summed over ranks and divided by the world size, the loss is averaged, the learning
rate is scaled by sqrt(world).  The reference needs a presence map and one message per parameter tensor; with a
flat arena (absent gradients are zeros at fixed offsets) a single all-reduce carries everything.
"""
import math

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def allreduce_grad_and_loss(grad_flat, loss):
    """Sum the flat gradient arena over ranks in place (loss scalar piggy-backed).

    Returns (mean loss as float, grad_scale) — grad_scale = 1/world is folded into the optimiser kernel
    (ppb_adam_step's grad_scale) instead of a separate divide pass."""
    world, _ = world_info()
    if world == 1:
        return float(loss), 1.0
    packed = torch.cat([grad_flat.reshape(-1), loss.detach().reshape(1).to(grad_flat.dtype)])
    dist.all_reduce(packed)
    grad_flat.copy_(packed[:-1].view_as(grad_flat))
    return float(packed[-1]) / world, 1.0 / world


def scaled_learning_rate(lr, world):
    return lr * math.sqrt(world)


def shard_range(n, rank, world):
    """Contiguous particle range [first, first+count) of rank `rank` (remainder spread over the low ranks)."""
    base, rem = divmod(n, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_weight_partials(partials):
    """All-gather the per-block (max, sum exp, sum exp^2) triples of every rank -> one list that
    ppb_weights_finalize combines exactly.  Ranks may hold different numbers of triples."""
    world, _ = world_info()
    if world == 1:
        return partials
    n_local = torch.tensor([partials.numel()], dtype=torch.int64, device=partials.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    m = int(max(int(s) for s in sizes))
    padded = torch.zeros(m, dtype=partials.dtype, device=partials.device)
    padded[:partials.numel()] = partials
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)])
