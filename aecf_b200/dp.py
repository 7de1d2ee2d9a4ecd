"""Batch-sharded data parallelism for the fusion pool: one process per GPU, NCCL over NVLink.

The reference has no distributed code; this layer is the north_star's item (3).
Samples are independent in the forward and in every activation gradient, so ranks only meet in one
sum-all-reduce of the fusion parameter gradients per step (1 051 136 elements at D = 512).

  * rank r owns global rows [r*B/N, (r+1)*B/N); ``pool.row_offset`` keys the Philox counters on the
    GLOBAL row, so any N reproduces the 1-GPU masks and dropout bit for bit
  * gradients are reduced in readiness order -- out_proj first, in_proj and the query last -- each
    group on a side stream as soon as the backward has produced it, overlapping the rest of the
    backward (the fused pool backward and the two large in-projection GEMMs)
  * reduction happens in an fp32 flat bucket regardless of the parameter dtype

Works with any ``torch.distributed`` backend: NCCL on the GPUs, gloo in the CPU tests of the host logic.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

PARAM_ORDER = ("out_proj.bias", "out_proj.weight", "in_proj_weight", "in_proj_bias", "query")


def shard_rows(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first global row, row count) of ``rank``: contiguous, sizes differ by at most one."""
    base, extra = divmod(global_batch, world_size)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


class GradientSync:
    """Overlapped all-reduce (mean) of the fusion parameter gradients of one pool + its query.

    ``attach()`` hooks the pool so that its backward reports each parameter gradient the moment it is
    final; ``finish()`` waits for the collectives and writes the averaged gradients into ``.grad``.
    """

    def __init__(self, pool, query: Optional[torch.nn.Parameter] = None, process_group=None,
                 average: bool = True):
        self.pool, self.query, self.group, self.average = pool, query, process_group, average
        att = pool.attention
        self.params: Dict[str, torch.nn.Parameter] = {}
        for name, p in (("out_proj.bias", att.out_proj.bias), ("out_proj.weight", att.out_proj.weight),
                        ("in_proj_weight", att.in_proj_weight), ("in_proj_bias", att.in_proj_bias), ("query", query)):
            if p is not None:
                self.params[name] = p
        device = att.in_proj_weight.device
        self.slices: Dict[str, slice] = {}
        n = 0
        for name in PARAM_ORDER:
            if name in self.params:
                k = self.params[name].numel()
                self.slices[name] = slice(n, n + k)
                n += (k + 3) // 4 * 4                       # keep every slice 16-byte aligned
        self.bucket = torch.zeros(n, dtype=torch.float32, device=device)
        self.cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.pending: List[Tuple[str, object]] = []
        self.reported: set = set()

    # -- wiring -----------------------------------------------------------------------------
    def attach(self) -> "GradientSync":
        self.pool._grad_ready = self.on_ready
        if self.query is not None:
            self.query.register_hook(lambda g: self.on_ready("query", g) or g)
        return self

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def set_shard(self, global_batch: int) -> Tuple[int, int]:
        """Point the pool's Philox row offset at this rank's shard; returns (row0, rows)."""
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        row0, rows = shard_rows(global_batch, rank, self.world_size)
        self.pool.row_offset = row0
        return row0, rows

    # -- called from the backward (autograd worker thread) -----------------------------------
    def on_ready(self, name: str, grad: torch.Tensor) -> None:
        if name not in self.slices or name in self.reported or self.world_size == 1:
            return                                                     # single process: autograd's .grad is final
        self.reported.add(name)
        view = self.bucket[self.slices[name]]
        view.copy_(grad.reshape(-1))                                   # cast to fp32 into the flat bucket
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.bucket.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((name, work))

    # -- called by the training loop after loss.backward() -----------------------------------
    def finish(self) -> None:
        """Wait for the collectives and store the reduced gradients in ``param.grad``."""
        for _, work in self.pending:
            work.wait()
        if self.cuda and self.pending:
            torch.cuda.current_stream(self.bucket.device).wait_stream(self.comm_stream)
        scale = 1.0 / self.world_size if self.average else 1.0
        for name in self.reported:
            p = self.params[name]
            reduced = self.bucket[self.slices[name]].reshape(p.shape)
            if p.grad is None:
                p.grad = torch.empty_like(p)
            if scale != 1.0:
                p.grad.copy_(reduced * scale)
            else:
                p.grad.copy_(reduced)
        self.pending.clear()
        self.reported.clear()
