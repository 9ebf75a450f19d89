"""Batch sharding across the GPUs of one box: one process per GPU, contiguous range partition.

The reference has no counterpart (single instance, single thread; SURVEY.md section 2
"Parallelism": none).  Instances are independent solves that share one compiled structure, so
the batch is range-partitioned over ranks with NO collective on the data path; the only
communication is an optional terminal gather of ``{status, iter, outputs}`` to rank 0
(``torch.distributed``: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_arrays(values: Dict[str, np.ndarray], shapes: Sequence[Tuple[str, Tuple[int, ...]]], batch: int,
                 rank: int, world: int) -> Dict[str, np.ndarray]:
    """slice of every array along its trailing batch axis; arrays without a batch axis
    (shape == declared shape) are shared by all instances and passed through"""
    a, b = shard_range(batch, rank, world)
    out = {}
    for name, shape in shapes:
        v = np.asarray(values[name], dtype=np.float64)
        out[name] = v if v.shape == tuple(shape) else v[..., a:b]
    return out


class ShardedSolver:
    """Same surface as the generated class, over a batch spread across ``world`` ranks.

    Every rank constructs it with the GLOBAL batch size and passes GLOBAL arrays (or its own
    slice with ``local=True``); ``solve`` returns this rank's slice, ``gather`` assembles the
    global result on rank ``dst`` (None elsewhere)."""

    def __init__(self, gen, batch: int, rank: int, world: int, device: Optional[int] = None,
                 solver_factory: Optional[Callable] = None):
        self.gen, self.batch, self.rank, self.world = gen, int(batch), rank, world
        self.start, self.stop = shard_range(self.batch, rank, world)
        self.local_batch = self.stop - self.start
        factory = solver_factory or (lambda g, b, d: g(b, d))
        self.obj = factory(gen, self.local_batch, device)
        self._shapes = dict(gen.parameters)
        self._shapes.update(dict(gen.variables))

    def _slice(self, name: str, value, local: bool):
        v = np.asarray(value, dtype=np.float64)
        shape = tuple(self._shapes[name])
        if v.shape == shape or local:
            return v
        if v.shape != shape + (self.batch,):
            raise ValueError(f"{name}: expected shape {shape} or {shape + (self.batch,)}, got {v.shape}")
        return v[..., self.start:self.stop]

    def __getattr__(self, name: str):
        if name.startswith(("setP_", "setV_")) and name[5:] in self.__dict__.get("_shapes", {}):
            return lambda value, local=False: getattr(self.obj, name)(self._slice(name[5:], value, local))
        raise AttributeError(name)

    def solve(self, *args, **kw):
        return self.obj.solve(*args, **kw)

    def getOutputs(self):
        return self.obj.getOutputs()

    def gather(self, status: np.ndarray, it: np.ndarray, outputs, dst: int = 0, group=None):
        """terminal gather of the per-rank results (trailing batch axis) to rank `dst`"""
        import torch
        import torch.distributed as dist
        names = [n for n, _ in self.gen.outputs]
        outs = outputs if isinstance(outputs, dict) else dict(zip(names, outputs))
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        maxb = shard_range(self.batch, 0, self.world)[1]          # rank 0 holds the largest shard

        def pad(a: np.ndarray) -> "torch.Tensor":
            a2 = np.ascontiguousarray(np.moveaxis(np.asarray(a), -1, 0))      # batch-major
            buf = np.zeros((maxb,) + a2.shape[1:], dtype=a2.dtype)
            buf[:a2.shape[0]] = a2
            return torch.from_numpy(buf).to(dev)

        def gather_one(a: np.ndarray):
            t = pad(a)
            parts = [torch.empty_like(t) for _ in range(self.world)] if self.rank == dst else None
            if backend == "nccl":       # gather is not available on every NCCL build: all_gather
                parts = [torch.empty_like(t) for _ in range(self.world)]
                dist.all_gather(parts, t, group=group)
            else:
                dist.gather(t, parts, dst=dst, group=group)
            if self.rank != dst:
                return None
            chunks = []
            for r, p in enumerate(parts):
                a_, b_ = shard_range(self.batch, r, self.world)
                chunks.append(p[:b_ - a_].cpu().numpy())
            return np.moveaxis(np.concatenate(chunks, axis=0), 0, -1)

        st = gather_one(status.astype(np.int32))
        iters = gather_one(it.astype(np.int32))
        res = {n: gather_one(outs[n]) for n in names}
        if self.rank != dst:
            return None
        return st, iters, (res if isinstance(outputs, dict) else [res[n] for n in names])

    def close(self):
        self.obj.close()
