"""Data parallelism for the path: one process per GPU, videos sharded across ranks, one flat
all-reduce of the USED parameter gradients per step (replaces the reference's single-process
nn.DataParallel broadcast/scatter/gather/reduce, main.py:79).

The path has no cross-video op, every loss is a mean over rows, so with equal
shards  mean_over_ranks(local gradient) == gradient of the global batch.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world: int) -> slice:
    """Equal contiguous shard of n rows; n must divide evenly (pad upstream like main.py:366-372)."""
    if n % world != 0:
        raise ValueError(f"{n} rows do not shard evenly over {world} ranks; pad the batch "
                         "(the reference pads to a multiple of the GPU count, main.py:366-372)")
    per = n // world
    return slice(rank * per, (rank + 1) * per)


class GradientBucket:
    """Flat fp32 bucket over the gradients that exist after backward (parameters the path never
    touches have grad None in the reference too and are skipped)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = list(params)
        self._flat: Optional[torch.Tensor] = None
        self._views: List[torch.Tensor] = []
        self._key: Optional[Tuple] = None

    def _prepare(self, grads: Sequence[torch.Tensor]) -> None:
        key = tuple((g.shape, g.device) for g in grads)
        if key == self._key:
            return
        total = sum(g.numel() for g in grads)
        self._flat = torch.empty(total, dtype=torch.float32, device=grads[0].device)
        self._views, off = [], 0
        for g in grads:
            self._views.append(self._flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        self._key = key

    def used_grads(self) -> List[torch.Tensor]:
        return [p.grad for p in self.params if p.grad is not None]

    @property
    def nbytes(self) -> int:
        return 0 if self._flat is None else self._flat.numel() * 4

    def allreduce_mean(self, group=None, async_op: bool = False):
        """sum over ranks / world, written back into every .grad; one collective."""
        grads = self.used_grads()
        if not grads:
            return None
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        self._prepare(grads)
        torch._foreach_copy_(self._views, grads)
        work = dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return _Pending(self, work, grads, world)
        self._finish(grads, world)
        return None

    def _finish(self, grads, world):
        self._flat.mul_(1.0 / world)
        torch._foreach_copy_(grads, self._views)


class _Pending:
    def __init__(self, bucket, work, grads, world):
        self.bucket, self.work, self.grads, self.world = bucket, work, grads, world

    def wait(self):
        self.work.wait()
        self.bucket._finish(self.grads, self.world)
