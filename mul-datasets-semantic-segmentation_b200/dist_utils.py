"""The few collective steps of the path: images shard data-parallel, one rank per GPU.

The loss needs no exchange (OHEM statistics, n_min and the selection are rank-local in the reference, DDP
averages parameter gradients outside the path); evaluation adds ONE all-reduce of the integer confusion
matrices (evaluate.py:94-95, :187-188).  These helpers are backend-agnostic (nccl on the GPUs, gloo in the
CPU tests) and are the only place the package touches torch.distributed.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_images(n_images, rank, world_size):
    """Contiguous block of image indices owned by `rank` (sizes differ by at most one; empty when
    n_images < world_size for the last ranks) — the split RepeatedDistSampler-style loaders produce."""
    base, rem = divmod(int(n_images), int(world_size))
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_images_balanced(costs, rank, world_size):
    """Image indices owned by `rank` when images differ in cost (the per-pixel work of the loss is proportional to the
    class count of the image's dataset: 19 for a Cityscapes image, 150 for an ADE image).  Same number of images per
    rank as `shard_images` (sizes differ by at most one), chosen greedily: images in order of falling cost, each to the
    least loaded rank that still has room (ties: lowest rank) — the longest-processing-time rule under a cardinality
    constraint.  Every rank computes the same assignment from the same `costs`; the result is sorted.  The reference's
    loaders give every rank the same dataset mix (ims_per_gpu of each dataset, lib/get_dataloader.py); a split of ONE
    mixed batch over the ranks has to balance it itself — contiguous blocks put both ADE images of a 16-image batch on
    one of eight ranks (2.45x the mean work)."""
    n = len(costs)
    base, rem = divmod(n, int(world_size))
    room = [base + (1 if r < rem else 0) for r in range(world_size)]
    load = [0.0] * world_size
    mine = []
    for i in sorted(range(n), key=lambda j: (-float(costs[j]), j)):
        r = min((q for q in range(world_size) if room[q] > 0), key=lambda q: (load[q], q))
        room[r] -= 1
        load[r] += float(costs[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def allreduce_hist(hist):
    """Sum int64 confusion matrices over all ranks, in place; exact (the reference reduces a float32 matrix)."""
    if hist.dtype != torch.int64:
        raise TypeError("confusion matrices are accumulated and reduced as int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def max_over_ranks(value, device="cpu"):
    """max of a python float over ranks (device-timed milliseconds: a step is as slow as its slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank, world_size, ms_per_step):
    """units/s of the whole job: every rank processed `units_per_rank` per step (the workload's batch per rank under
    weak scaling, its share of the batch under strong scaling)."""
    return units_per_rank * world_size / (ms_per_step * 1e-3)
