"""Multi-GPU plumbing: one process per GPU, latents block-sharded over ranks.

`torch.distributed` is only the bootstrap: it carries the 128-byte NCCL unique id from rank 0 to
the other ranks; the data-path collectives (one all-reduce of the m+1 log-likelihood terms, one of
the partial back-projections) are issued by liblmm itself on its compute stream.  On hosts without
a GPU (CPU tests, gloo) `shard_range` / `reduce_terms` cover the same sharding logic.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np


def shard_range(m: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Latents [lo, hi) owned by `rank` -- must match liblmm's shard_range (csrc/api.cu, `shard_range`)."""
    return (m * rank) // nranks, (m * (rank + 1)) // nranks


def init_context_distributed(ctx) -> None:
    """Give `ctx` an NCCL communicator spanning the current torch.distributed world."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    buf = C.create_string_buffer(128)
    if rank == 0:
        rc = ctx.lib.lmm_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise RuntimeError("lmm_comm_unique_id failed: libnccl not available")
    dev = torch.device("cuda", ctx.device) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, src=0)
    ctx.init_nccl(bytes(t.cpu().numpy().tobytes()), world, rank)


def reduce_terms(local_terms: np.ndarray) -> np.ndarray:
    """Sum per-latent lml terms over ranks through torch.distributed (gloo or nccl).  Used when the
    library was only told its shard (`Context.set_shard`) and has no NCCL communicator."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local_terms, dtype=np.float64)
    t = torch.from_numpy(np.array(local_terms, dtype=np.float64, copy=True))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
