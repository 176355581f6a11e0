"""Multi-GPU plumbing: one process per GPU over ``torch.distributed`` (NCCL on the B200 box,
gloo in CPU tests).

* inference shards independent chunks over ranks with no data-path collective (the reference
  splits ``node_chunks`` over SLURM nodes, pipeline/strategy.py:363-372) -- the only collective
  is one weight broadcast at start-up;
* training follows the reference's ``_get_parallel_grad`` / ``_sum_parallel_grad``
  (models/abstract.py:785-841): ONE batch is split along axis 0 into ``world_size`` equal
  shards (``tf.split`` raises when it does not divide), rank r computes the gradient of shard
  r (per-shard losses, per-shard relativistic means), the gradients are SUMMED (not averaged)
  through one persistent flat fp32 arena, and every rank applies the same optimiser step.
  The loss details that drive the GAN schedule are those of the LAST shard (the reference
  returns the last future's), broadcast so that every rank takes the same branch and issues
  the same collectives.
"""
from __future__ import annotations

import os

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


def bind_to_gpu_numa(device_index=None):
    """Restrict this thread to the CPUs NVML names as local to the GPU, so that the pinned host
    buffers it allocates afterwards (first touch) live on the GPU's NUMA node: with one rank per
    GPU the device -> host result streams then do not cross the socket interconnect.  Returns the
    CPU set, or None when NVML / affinity is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        if device_index is None:
            device_index = torch.cuda.current_device()
        pynvml.nvmlInit()
        # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:   # pragma: no cover - best effort
        return None


_arenas = {}


def _arena_for(grads):
    """Persistent flat fp32 buffer + per-tensor views for this list of gradient shapes."""
    key = (str(grads[0].device), tuple(tuple(g.shape) for g in grads))
    hit = _arenas.get(key)
    if hit is None:
        n = sum(g.numel() for g in grads)
        flat = torch.empty(n, device=grads[0].device, dtype=torch.float32)
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view(g.shape))
            off += g.numel()
        hit = (flat, views)
        _arenas[key] = hit
    return hit


_peer_arenas = {}
_peer_state = {"enabled": None}


class GradArena:
    """Flat fp32 gradient arena + the table ``s3_peer_sum_adam`` walks: one entry per
    <= 4096-element piece of a weight tensor (weight, first / second moment pointers, position of
    its gradient in the arena), one CTA each.  The kernel adds the arenas of ``world`` ranks in
    rank order and applies keras Adam in registers; with one rank (``LocalArena``) it is simply
    the whole network's optimiser step in one launch."""

    world = 1
    PIECE = 4096      # elements per table entry (one CTA each)

    def segment_table(self, ptrs, weights):
        """(n_pieces, 5) uint64 records {weight, m, v pointers of the piece, its offset in the
        arena, its length}: every weight tensor cut into pieces of <= PIECE elements so that the
        big dense / conv kernels spread over the whole GPU.  ``ptrs``: (weight, m, v) data
        pointers per tensor."""
        rows, off = [], 0
        for (wp, mp, vp), g, var in zip(ptrs, self.in_views, weights):
            if not var.value.is_contiguous() or var.value.dtype != torch.float32:
                raise ValueError(f"fused Adam needs contiguous float32 weights ({var.name})")
            if var.value.numel() != g.numel():
                raise ValueError(f"gradient arena does not match weight {var.name}")
            for s0 in range(0, g.numel(), self.PIECE):
                rows.append((wp + 4 * s0, mp + 4 * s0, vp + 4 * s0, off + s0,
                             min(self.PIECE, g.numel() - s0)))
            off += g.numel()
        return np.array(rows, dtype=np.uint64).reshape(-1, 5)

    def _layout(self, grads, arena):
        self.in_views, off = [], 0
        for g in grads:
            self.in_views.append(arena[off:off + g.numel()].view(g.shape))
            off += g.numel()

    def _barrier(self, channel):
        pass

    def stage(self, grads):
        """Copy this rank's gradients into the arena (one multi-tensor launch)."""
        torch._foreach_copy_(self.in_views, [g.detach() for g in grads])

    def sum_and_apply_adam(self, grads, weights, optimizer):
        """The exchange and the optimiser step in ONE kernel (``s3_peer_sum_adam``): every rank
        adds all gradient arenas in rank order and applies keras Adam to its own (replicated)
        weights; the summed gradient never goes back to memory.  ``grads`` None: already staged
        (the CUDA-graph training step copies them as its last node)."""
        from . import _cabi, ops
        slots = [optimizer.slots_for(var) for var in weights]
        key = tuple((var.value.data_ptr(), m.data_ptr(), v.data_ptr())
                    for var, (m, v) in zip(weights, slots))
        if getattr(self, "_seg_key", None) != key:
            rec = self.segment_table(key, weights)
            self._segs = torch.from_numpy(rec.view(np.int64)).to(self.arena.device)
            self._n_seg, self._max_n = len(rec), self.PIECE
            self._seg_key = key
        if grads is not None:
            self.stage(grads)
        optimizer.iterations += 1
        self._barrier(0)          # every rank's arena is written
        _cabi.call("s3_peer_sum_adam", self._ptrs, self.world, ops._p(self._segs), self._n_seg,
                   self._max_n, float(optimizer.learning_rate), float(optimizer.beta_1),
                   float(optimizer.beta_2), float(optimizer.epsilon), int(optimizer.iterations),
                   ops._s())
        ops._count()
        self._barrier(1)          # every rank has read every arena
        for var in weights:
            var.version += 1


class LocalArena(GradArena):
    """Single-GPU arena: the optimiser step of all weight tensors in one launch."""

    def __init__(self, grads):
        import ctypes as C
        n = sum(g.numel() for g in grads)
        self.arena = torch.zeros(max(n, 1), dtype=torch.float32, device=grads[0].device)
        self._ptrs = (C.c_void_p * 1)(self.arena.data_ptr())
        self._layout(grads, self.arena)


class PeerArena(GradArena):
    """Flat fp32 gradient arena in symmetric (NVLink peer-mapped) memory + a local result buffer.
    ``torch.distributed._symmetric_memory`` only does the plumbing (allocation, exchange of the
    peer mappings, the cross-GPU barrier on its signal pads); the reduction itself is this
    library's kernel (``s3_peer_sum_f32`` / ``s3_peer_sum_adam``): every rank reads all arenas
    over NVLink and adds them in rank order -- deterministic and bit-identical on all ranks."""

    def __init__(self, grads):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        dev = grads[0].device
        n = sum(g.numel() for g in grads)
        self.n = (n + 3) // 4 * 4
        self.arena = symm_mem.empty(self.n, dtype=torch.float32, device=dev)
        self.arena.zero_()
        self.handle = symm_mem.rendezvous(self.arena, dist.group.WORLD)
        self.world = self.handle.world_size
        ptrs = list(self.handle.buffer_ptrs)
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        self.out = torch.empty(self.n, device=dev, dtype=torch.float32)
        self._layout(grads, self.arena)
        self.out_views, off = [], 0
        for g in grads:
            self.out_views.append(self.out[off:off + g.numel()].view(g.shape))
            off += g.numel()

    def _barrier(self, channel):
        self.handle.barrier(channel=channel)

    def allreduce(self, grads):
        from . import _cabi, ops
        self.stage(grads)
        self.handle.barrier(channel=0)          # every rank's arena is written
        _cabi.call("s3_peer_sum_f32", self._ptrs, self.world, self.n,
                   ops._p(self.out), ops._s())
        ops._count()
        self.handle.barrier(channel=1)          # every rank has read every arena
        return self.out_views


def peer_allreduce_enabled():
    """NVLink peer-memory gradient reduction: on for the NCCL backend unless
    ``SUP3R_B200_PEER_ALLREDUCE=0`` (falls back to ``dist.all_reduce`` when symmetric memory
    cannot be set up, e.g. no P2P access between the GPUs)."""
    import os
    if _peer_state["enabled"] is None:
        _peer_state["enabled"] = (os.environ.get("SUP3R_B200_PEER_ALLREDUCE", "1") != "0"
                                  and is_distributed() and dist.get_backend() == "nccl")
    return _peer_state["enabled"]


def _peer_arena_for(grads):
    if not (peer_allreduce_enabled() and grads[0].is_cuda):
        return None
    key = (str(grads[0].device), tuple(tuple(g.shape) for g in grads))
    pa = _peer_arenas.get(key)
    if pa is None:
        try:
            pa = _peer_arenas[key] = PeerArena(grads)
        except Exception as e:   # pragma: no cover - no P2P / symmetric memory
            import logging
            logging.getLogger(__name__).warning(
                "symmetric-memory gradient arena unavailable (%s): using NCCL all-reduce", e)
            _peer_state["enabled"] = False
            return None
    return pa


_local_arenas = {}


def step_arena(grads, optimizer, multi_gpu=False):
    """The arena whose fused kernel can take this optimiser step, or None: this library's Adam on
    CUDA tensors -- ``PeerArena`` for a multi-GPU step over NVLink symmetric memory,
    ``LocalArena`` for a single-GPU step.  ``SUP3R_B200_FUSED_ADAM=0`` disables both."""
    from .optimizers import Adam
    if type(optimizer) is not Adam or not grads or not grads[0].is_cuda \
            or os.environ.get("SUP3R_B200_FUSED_ADAM", "1") == "0":
        return None
    if multi_gpu and world_size() > 1:
        return _peer_arena_for(grads)
    key = (str(grads[0].device), tuple(tuple(g.shape) for g in grads))
    la = _local_arenas.get(key)
    if la is None:
        la = _local_arenas[key] = LocalArena(grads)
    return la


def sum_grads_and_step(grads, weights, optimizer, multi_gpu=True, arena=None, staged=False):
    """``_sum_parallel_grad`` + ``optimizer.apply_gradients`` (abstract.py:785-805, 899-912).
    With this library's Adam: one fused kernel per step (``GradArena.sum_and_apply_adam`` -- over
    NVLink symmetric memory when ``multi_gpu``); otherwise all-reduce, then the optimiser's own
    step.  ``staged``: the gradients already sit in ``arena`` (CUDA-graph training step)."""
    if arena is None:
        arena = step_arena(grads, optimizer, multi_gpu)
    if arena is not None:
        arena.sum_and_apply_adam(None if staged else grads, weights, optimizer)
        if arena.world > 1:
            _peer_state["fused_adam_steps"] = _peer_state.get("fused_adam_steps", 0) + 1
        return
    if multi_gpu and world_size() > 1:
        grads = allreduce_sum_grads(grads)
    optimizer.apply_gradients(zip(grads, weights))


def allreduce_sum_grads(grads):
    """SUM all-reduce of a list of gradient tensors through one persistent flat arena (one
    exchange per step).  Returns the reduced gradients as VIEWS of a persistent buffer (no copy
    back): valid until the next call with the same shapes.  On NVLink-connected GPUs the arena is
    symmetric memory and the sum is this library's peer-memory kernel (``PeerArena``); otherwise
    one NCCL / gloo all-reduce."""
    if not is_distributed() or world_size() == 1:
        return grads
    pa = _peer_arena_for(grads)
    if pa is not None:
        return pa.allreduce(grads)
    flat, views = _arena_for(grads)
    torch._foreach_copy_(views, [g.detach() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return views


def shard_batch(x, n_shards=None, index=None):
    """``tf.split(x, n, axis=0)[index]`` (abstract.py:819-825): equal shards or an error."""
    n_shards = world_size() if n_shards is None else n_shards
    index = rank() if index is None else index
    if x is None or n_shards == 1:
        return x
    b = x.shape[0]
    if b % n_shards:
        raise ValueError(f"Dimension 0 of a batch of {b} observations is not evenly divisible "
                         f"by the {n_shards} GPUs (tf.split semantics)")
    k = b // n_shards
    return x[index * k:(index + 1) * k]


def broadcast_loss_details(loss_details, src=None):
    """Every rank continues with the loss details of the LAST shard (the reference returns the
    last future's, abstract.py:791-805): one small broadcast keeps the GAN schedule -- and
    therefore the sequence of collectives -- identical on all ranks."""
    if not is_distributed() or world_size() == 1:
        return loss_details
    src = world_size() - 1 if src is None else src
    keys = sorted(loss_details)
    dev = None
    for v in loss_details.values():
        if isinstance(v, torch.Tensor):
            dev = v.device
            break
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device()) \
            if dist.get_backend() == "nccl" else torch.device("cpu")
    vals = torch.stack([torch.as_tensor(loss_details[k], dtype=torch.float32).detach()
                        .reshape(()).to(dev) for k in keys])
    dist.broadcast(vals, src=src)
    return {k: vals[i] for i, k in enumerate(keys)}


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
