"""Utterance sharding across GPUs (one process per GPU).

The reference farms utterances out with ``dlp_mpi.split_managed`` (manager /
worker queue over MPI, pb_chime5/core.py:381) or with a strided slice
(``scripts/kaldi_run.py:73-76``).  Utterances are independent, so here a rank
simply owns a static shard: strided by default, or longest-processing-time
greedy bins when per-utterance lengths are known.  The only collectives are a
broadcast of the work list and a barrier (``torch.distributed``; NCCL on GPUs,
gloo in the CPU tests) -- there is no data-path collective.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for 1 rank)."""
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            local = int(os.environ.get('LOCAL_RANK', 0))
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend=backend)
    return rank_world()


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def shutdown():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def broadcast_work_list(items, src=0):
    """Rank `src` decides the work list; everybody gets the same copy."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [items if dist.get_rank() == src else None]
        dist.broadcast_object_list(box, src=src)
        return box[0]
    return items


def shard_indices(n, rank, world, lengths=None):
    """Indices of the utterances rank `rank` of `world` processes.

    lengths=None : i % world == rank  (kaldi_run.py:73-76 semantics).
    lengths given: longest first, each utterance to the currently least loaded
                   rank (deterministic; ties to the lowest rank)."""
    assert 0 <= rank < world, (rank, world)
    if lengths is None:
        return list(range(rank, n, world))
    assert len(lengths) == n, (len(lengths), n)
    order = sorted(range(n), key=lambda i: (-lengths[i], i))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        load[r] += lengths[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (timing: the slowest rank counts)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(value)], dtype=torch.float64,
                         device=device if device is not None else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def sum_over_ranks(value, device=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(value)], dtype=torch.float64,
                         device=device if device is not None else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    return float(value)
