"""Multi-GPU plumbing: one process per GPU over ``torch.distributed`` (NCCL on the B200 box,
gloo in CPU tests).

* inference shards independent chunks over ranks with no data-path collective (the reference
  splits ``node_chunks`` over SLURM nodes, pipeline/strategy.py:363-372) -- the only collective
  is one weight broadcast at start-up;
* training all-reduces gradients with SUM (not mean), the reference's ``_sum_parallel_grad``
  semantics (models/abstract.py:785-805), on one flat fp32 bucket per step.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized()


def rank():
    return dist.get_rank() if is_distributed() else 0


def world_size():
    return dist.get_world_size() if is_distributed() else 1


def init_from_env(backend=None):
    """Initialise the default process group from the torchrun environment (idempotent)."""
    import os
    if is_distributed() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)


def allreduce_sum_grads(grads):
    """In-place SUM all-reduce of a list of gradient tensors through one flat bucket."""
    if not is_distributed() or world_size() == 1:
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads


def broadcast_weights(networks, src=0):
    """Broadcast every weight of the given networks from ``src`` (one flat bucket each)."""
    if not is_distributed() or world_size() == 1:
        return
    for net in networks:
        ws = [v.value for v in net.weights]
        if not ws:
            continue
        with torch.no_grad():
            flat = torch.cat([w.detach().reshape(-1) for w in ws])
            dist.broadcast(flat, src=src)
            off = 0
            for v in net.weights:
                n = v.value.numel()
                v.value.copy_(flat[off:off + n].view_as(v.value))
                v.version += 1
                off += n


def split_chunks(chunk_ids, n_parts):
    """``np.array_split`` partition used for ``node_chunks`` (strategy.py:363-372)."""
    n_parts = int(min(n_parts, max(len(chunk_ids), 1)))
    return [list(map(int, part)) for part in np.array_split(np.asarray(chunk_ids, dtype=int),
                                                            n_parts)]
