"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

Training shards *samples*: every rank draws its own minibatch from its own
Philox/permutation stream, and the three gradient groups of an update are
averaged over ranks.  Prediction shards *users*: no
collective until the final gather of the U x k result.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def dist_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torch.distributed if initialised, else the torchrun env, else (0,1,0)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))
    except Exception:  # pragma: no cover
        pass
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of ``n`` units for ``rank``; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(tensor, group=None) -> None:
    """In-place mean over ranks (works for NCCL on device tensors and gloo on CPU tensors)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    tensor.div_(world)


class GradAllReducer:
    """Averages the engine's gradient buffers in place, zero-copy, on torch's current stream."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self._views = {}

    fused = False

    def __call__(self, buffer_id: int) -> None:
        view = self._views.get(buffer_id)
        if view is None:
            view = self._views[buffer_id] = self.engine.grad_tensor(buffer_id)
        allreduce_mean_(view, self.group)

    def check(self) -> None:          # a collective library call cannot time out silently
        pass


class PeerGradExchange:
    """Gradient averaging through NVLink peer memory: the library's own one-shot all-reduce kernels
    (``csrc/dp_peer.cuh``) on symmetric buffers from ``torch.distributed._symmetric_memory`` -- no collective
    library call inside the update.  Same call signature as :class:`GradAllReducer`; construct it collectively
    (every rank, same order).  Raises if symmetric memory cannot be set up, so callers can fall back to NCCL."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.engine = engine
        pg = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
        _, n = engine.device_buffer(_lib.BUF_ALL_GRADS)
        self.stage_floats = (n + 127) // 128 * 128        # whole tiles of 32 16-byte value groups (fused exchange)
        dev = torch.device("cuda", engine.device)
        # [2][stage] halves of the stand-alone exchange kernel, then [2][world][stage] {value, tag} pairs that the peers'
        # update kernels push into (fused exchange, csrc/dp_peer.cuh)
        self.buffer_floats = 2 * self.stage_floats + 4 * world * self.stage_floats
        self.buf = symm_mem.empty(self.buffer_floats, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, pg)
        if self.handle.signal_pad_size < world * 4 * 8:
            raise RuntimeError("symmetric-memory signal pad too small")
        torch.cuda.synchronize(dev)
        dist.barrier(pg)                                  # every staging buffer is zeroed before anyone publishes
        engine.dp_attach(world, rank, list(self.handle.buffer_ptrs), list(self.handle.signal_pad_ptrs), self.stage_floats,
                         self.buffer_floats)
        dist.barrier(pg)

    @property
    def fused(self) -> bool:
        """The update kernels do the exchange themselves (f16x3): no call is needed between the phases, and
        ``engine.update`` / ``engine.update_batches`` are data-parallel as they are."""
        return self.engine.dp_fused

    def __call__(self, buffer_id: int) -> None:
        self.engine.dp_allreduce(buffer_id)               # on torch's current stream (a no-op in fused mode)

    def check(self) -> None:
        """Raise if a wait for a peer timed out: the replicas no longer hold the same averaged gradients."""
        if self.engine.dp_error():
            raise RuntimeError("data-parallel gradient exchange timed out (a peer rank stalled for too long); "
                               "replicas are no longer identical -- restart from the last checkpoint")


def exchange_slices(world: int, off_floats: int, n_floats: int):
    """Host-side mirror of ``dp_slices`` / ``dp_owner`` (csrc/dp_peer.cuh): the fused exchange cuts a gradient group (its
    16-byte value groups ``q``) into ``world`` contiguous slices, slice ``o`` owned by rank ``o``.  Returns
    ``(q_lo, n_q, per, owner)`` with ``owner(q_abs) -> rank``.  ``per`` is even, so the 8 consecutive floats an Adam
    thread updates never straddle two owners (tests/test_parallel_gloo.py checks the invariants the kernels rely on)."""
    q_lo, n_q = off_floats >> 2, n_floats >> 2
    per = ((n_q + world - 1) // world + 1) & ~1
    return q_lo, n_q, per, (lambda q_abs: (q_abs - q_lo) // per)


def make_grad_exchange(engine, group=None, prefer_peer: bool = True):
    """PeerGradExchange when every rank can set it up, else the NCCL/gloo reducer (decided collectively)."""
    import torch
    import torch.distributed as dist
    ex, ok = None, 0
    if prefer_peer and dist.get_backend(group) == "nccl" and os.environ.get("CQL_DP_EXCHANGE", "peer") == "peer":
        try:
            ex = PeerGradExchange(engine, group)
            ok = 1
        except Exception:                                  # no symmetric memory in this environment
            ex, ok = None, 0
        t = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", engine.device))
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        ok = int(t.item())
    return ex if ok and ex is not None else GradAllReducer(engine, group)


def gather_rows(local: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank arrays (ragged in dim 0) on every rank, in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = np.zeros((mx,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.concatenate([o.cpu().numpy()[:s] for o, s in zip(outs, sizes)], axis=0)


class DataParallelStepper:
    """Replays one data-parallel update (4 library phases + 3 NCCL all-reduces) as a single CUDA graph.

    The three collectives of an update are sub-MB and latency-bound, and the Python/ctypes round trips between
    phases cost more than the collectives themselves; capturing the whole step removes the host from the loop.
    """

    def __init__(self, engine, group=None, warmup: int = 3, uploaded_batch: bool = False, reducer=None):
        """``uploaded_batch``: the captured step consumes the minibatch last staged by ``engine.upload_batch`` instead
        of sampling the replay table (end-to-end host-buffer path); ``reducer``: share another stepper's exchange."""
        import torch
        self.engine = engine
        self.uploaded_batch = uploaded_batch
        self.reducer = reducer if reducer is not None else make_grad_exchange(engine, group)
        self.device = torch.device("cuda", engine.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.graph = None
        self._warmup = warmup
        self.launches_per_step = 0             # this library's kernel launches inside one captured update
        self.replayed_steps = 0
        self.eager_steps = 0

    def _eager(self, n: int) -> None:
        import torch
        with torch.cuda.stream(self.stream):
            for _ in range(n):
                self.engine.update_data_parallel(self.reducer, stream=self.stream.cuda_stream,
                                                 uploaded_batch=self.uploaded_batch)
        self.eager_steps += n

    def _capture(self) -> None:
        """Capture ONE update as a CUDA graph.  No update runs here: the eager warm-up steps (NCCL communicators /
        lazy state must exist before capture) are taken by ``run`` out of the caller's step budget."""
        import torch
        self.stream.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = self.engine.launch_count
        with torch.cuda.graph(g, stream=self.stream):
            self.engine.update_data_parallel(self.reducer, stream=self.stream.cuda_stream,
                                             uploaded_batch=self.uploaded_batch)
        self.launches_per_step = self.engine.launch_count - l0
        self.graph = g

    def run(self, n_steps: int) -> None:
        """Enqueue EXACTLY ``n_steps`` updates on ``self.stream`` (asynchronous).  The first ``warmup`` of them run
        eagerly (they are real updates and count), the rest replay the captured graph."""
        import torch
        n_steps = int(n_steps)
        if n_steps <= 0:
            return
        if self.graph is None:
            w = min(max(self._warmup - self.eager_steps, 0), n_steps)
            self._eager(w)
            n_steps -= w
            if n_steps == 0:
                return
            try:
                self._capture()
            except Exception:                  # capture unsupported in this build: stay eager
                self.graph = False
        if self.graph:
            with torch.cuda.stream(self.stream):
                for _ in range(n_steps):
                    self.graph.replay()
            self.replayed_steps += n_steps
        else:
            self._eager(n_steps)

    @property
    def steps_done(self) -> int:
        """updates enqueued so far (eager + replayed)"""
        return self.eager_steps + self.replayed_steps

    def finish(self) -> None:
        """Wait for the enqueued updates and fail loudly if a peer-memory exchange timed out: after a timeout the
        ranks no longer hold the same averaged gradients, so the replicas have diverged and training must stop."""
        self.stream.synchronize()
        if isinstance(self.reducer, PeerGradExchange) and self.engine.dp_error():
            raise RuntimeError("data-parallel gradient exchange timed out (a peer rank stalled for too long); "
                               "replicas are no longer identical -- restart from the last checkpoint")
