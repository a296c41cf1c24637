"""Data parallelism over the minibatch rows: the bound is a sum over data points, so each rank
evaluates a contiguous slice with the GLOBAL scale num_data / N_global, the redundantly computed KL terms are
weighted 1 / world_size, and ONE all-reduce(sum) of the packed [gradients | elbo, var_exp, kl_g, kl_e] buffer
gives every rank the single-process result (exactly so when the draws are passed explicitly; with library-drawn noise
the ranks share the inducing-point draws and draw their per-point noise from independent per-rank streams, which is one
valid realisation of the single-process estimator).  NCCL runs inside the CUDA library (mogpe_allreduce_sum);
torch.distributed is only the plumbing that distributes the NCCL unique id.
"""
from typing import Tuple

import numpy as np


def shard_slice(n_global: int, rank: int, world_size: int) -> slice:
    """Contiguous row slice of rank `rank`; the first n_global % world_size ranks take one extra row."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, rem = divmod(n_global, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_batch(X, Y, rank: int, world_size: int) -> Tuple[np.ndarray, np.ndarray]:
    s = shard_slice(np.shape(X)[0], rank, world_size)
    return np.ascontiguousarray(X[s]), np.ascontiguousarray(Y[s])


def global_scale(num_data, n_global: int) -> float:
    """variational_expectation_scale (mosvgpe.py:309-316) with the GLOBAL batch size."""
    return 1.0 if num_data is None else float(num_data) / float(n_global)


def kl_weight(world_size: int) -> float:
    return 1.0 / float(world_size)


class DataParallelEngine:
    """Wraps an Engine for one rank of a data-parallel job (one process per GPU)."""

    def __init__(self, engine, rank: int, world_size: int, unique_id: bytes = None):
        self.engine, self.rank, self.world_size = engine, rank, world_size
        if world_size > 1:
            if unique_id is None:
                raise ValueError("world_size > 1 needs the NCCL unique id from Engine.nccl_unique_id() of rank 0")
            engine.init_data_parallel(rank, world_size, unique_id)

    @staticmethod
    def exchange_unique_id(engine_cls, rank: int):
        """Create the id on rank 0 and broadcast it with torch.distributed (any initialised backend)."""
        import torch.distributed as dist

        ids = [engine_cls.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def step(self, X_local, Y_local, n_global: int, num_data, bound="further_gating", num_samples=1, readback=False,
             **eps):
        """Local forward+backward with the global scale, then one all-reduce of [grads | out]."""
        scale = global_scale(num_data, n_global)
        eng = self.engine
        # device entry point: [gradients | elbo, var_exp, kl_g, kl_e] land in eng.gbuf (host inputs are uploaded first)
        eng.elbo(X_local, Y_local, bound, num_samples, scale, want_grads=True, unpack=False, readback=False, **eps)
        if self.world_size > 1:
            eng.allreduce(eng.gbuf)  # no-op when the library already reduced inside elbo() (overlap mode)
        if not readback:
            return None
        from .engine import ElboResult

        eng.check_numerics()
        res = ElboResult(*[float(v) for v in eng.gbuf.numpy(offset=eng.lo.total, count=4)])
        res.flat_grads = eng.gbuf
        return res
