"""Utterance sharding across GPUs (one process per GPU).

The reference farms utterances out with ``dlp_mpi.split_managed`` (manager /
worker queue over MPI, pb_chime5/core.py:381) or with a strided slice
(``scripts/kaldi_run.py:73-76``).  Utterances are independent; three
schedules are offered: ``'dynamic'`` (default with several ranks: `WorkQueue`, the task-farm
semantics of ``split_managed`` -- length-sorted batches are pulled from a shared counter),
``'lpt'`` (static longest-processing-time shards from the known lengths) and ``'strided'``
(``kaldi_run.py`` semantics).  The only collectives are a broadcast of the work list, a barrier
and a gather of the per-rank reports (``torch.distributed``; NCCL on GPUs, gloo in the CPU
tests) -- there is no data-path collective.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


_RANK_VARS = (('RANK', 'WORLD_SIZE'), ('OMPI_COMM_WORLD_RANK', 'OMPI_COMM_WORLD_SIZE'),
              ('PMI_RANK', 'PMI_SIZE'), ('PMIX_RANK', 'OMPI_UNIVERSE_SIZE'), ('SLURM_PROCID', 'SLURM_NTASKS'))
_LOCAL_VARS = ('LOCAL_RANK', 'OMPI_COMM_WORLD_LOCAL_RANK', 'MPI_LOCALRANKID', 'SLURM_LOCALID')


def _env_rank_world():
    """(rank, world) from the launcher's environment: torchrun, or the MPI / SLURM launchers the
    reference is started with (`mpiexec -np N python -m pb_chime5.scripts.run`, README)."""
    for r, w in _RANK_VARS:
        if r in os.environ and w in os.environ:
            return int(os.environ[r]), int(os.environ[w])
    return 0, 1


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return _env_rank_world()


def local_rank():
    for v in _LOCAL_VARS:
        if v in os.environ:
            return int(os.environ[v])
    rank, _ = _env_rank_world()
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return rank % n if n else 0


def init_process_group(backend=None):
    """Bind this process to its GPU and initialise torch.distributed from the launcher's
    environment (torchrun, mpiexec, srun); no-op for a single process.  Idempotent."""
    rank, world = _env_rank_world()
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank())
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kw = dict(backend=backend, rank=rank, world_size=world)
        if backend == 'nccl':
            kw['device_id'] = torch.device('cuda', local_rank())
        dist.init_process_group(**kw)
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


class WorkQueue:
    """Dynamic work distribution with the semantics of `dlp_mpi.split_managed` (pb_chime5/core.py:381):
    every rank asks for the next item when it is free, so a slow rank never holds the job.  The
    "manager" is an atomic counter in the process group's key-value store (a few bytes per request on
    the host side; no GPU involvement, no dedicated manager rank -- all ranks work).  Items are the
    indices 0..n-1 in the order given; hand them out longest first for a good makespan.
    Single process / no process group: a plain range."""

    _serial = 0

    def __init__(self, n, name=None):
        self.n = int(n)
        WorkQueue._serial += 1                      # same construction order on every rank -> same key
        self.key = f'gss_workqueue/{name or WorkQueue._serial}'
        self.store = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.store = dist.distributed_c10d._get_default_store()
        self._local = 0
        self.taken = 0

    def next(self):
        if self.store is not None:
            i = int(self.store.add(self.key, 1)) - 1
        else:
            i, self._local = self._local, self._local + 1
        if i >= self.n:
            return None
        self.taken += 1
        return i

    def __iter__(self):
        while True:
            i = self.next()
            if i is None:
                return
            yield i


def gather_objects(obj, dst=0):
    """list of every rank's `obj` on rank `dst` (None elsewhere); [obj] for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
        dist.gather_object(obj, out, dst=dst)
        return out
    return [obj]


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
