"""Batch-sharded data parallelism for the fusion pool: one process per GPU, NCCL over NVLink.

The reference has no distributed code; this layer is the north_star's item (3).
Samples are independent in the forward and in every activation gradient, so ranks only meet in one
sum-all-reduce of the fusion parameter gradients per step (1 051 136 elements at D = 512).

  * rank r owns global rows [r*B/N, (r+1)*B/N); ``pool.row_offset`` keys the Philox counters on the
    GLOBAL row, so any N reproduces the 1-GPU masks and dropout bit for bit
  * the backward writes its parameter gradients straight into one flat bucket (no copies), which is reduced
    with ONE all-reduce after the backward, on the compute stream (the default): measured at 8 B200s with the
    step replayed as a CUDA graph, 46 us on top of a 686 us step (profiles/r1_run17_dp8_study.json)
  * ``overlap=True`` (or AECF_DP_OVERLAP=1) reduces in two groups in readiness order instead -- the
    out-projection gradients on a side stream as soon as ``aecf_fusion_bwd(AECF_BWD_OUT_PROJ)`` has produced
    them, the rest at the end.  It measured SLOWER (814 vs 732 us per step at N = 8): the persistent GEMM and
    pool kernels size their grids to own every SM, and NCCL's CTAs landing on some SMs first delay the CTAs
    that should have run there, which stretches the whole kernel

Works with any ``torch.distributed`` backend: NCCL on the GPUs, gloo in the CPU tests of the host logic.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

PARAM_ORDER = ("out_proj.bias", "out_proj.weight", "in_proj_weight", "in_proj_bias", "query")
EARLY = ("out_proj.bias", "out_proj.weight")            # final after the first backward phase


def shard_rows(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first global row, row count) of ``rank``: contiguous, sizes differ by at most one."""
    base, extra = divmod(global_batch, world_size)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


class PeerAllReduce:
    """In-place all-reduce of one CUDA tensor across the ranks of ONE node over NVLink peer memory, without NCCL:
    ``csrc/peer_allreduce.cu`` (barrier, every rank reduces its slice from all buckets in rank order and stores it
    into all buckets, barrier; bit-identical results on every rank, graph-capturable).

    Construction is collective: every rank exports its bucket and a small flag block through CUDA IPC (torch's
    tensor sharing), the handles travel through ``all_gather_object`` of ``group`` (any backend), and every rank
    maps the other ranks' buffers and enables peer access to their devices.  The bucket must stay alive, and must
    not be reallocated, for the lifetime of this object.  Opt-in (``GradientSync(collective="peer")`` or
    ``AECF_DP_PEER=1``): single-device emulation is tested (``tests/test_gpu_peer_allreduce.py``), the
    multi-process path is scheduled for its first hardware run in round 2.
    """

    def __init__(self, bucket: torch.Tensor, group=None, average: bool = True):
        from torch.multiprocessing.reductions import reduce_tensor

        from . import _lib, ops
        if not bucket.is_cuda or not bucket.is_contiguous():
            raise ValueError("PeerAllReduce needs a contiguous CUDA tensor")
        self.bucket, self.group, self.average = bucket, group, average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise ValueError("PeerAllReduce covers the (at most 8) GPUs of one NVLink node")
        dev = bucket.device
        self.flags = ops.peer_flag_block(dev)
        self.buckets, self.flag_blocks = [bucket], [self.flags]
        if self.world > 1:
            torch.cuda.synchronize(dev)
            payload = (reduce_tensor(bucket), reduce_tensor(self.flags), dev.index or 0)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, payload, group=group)
            self.buckets, self.flag_blocks = [], []
            for r, (b, f, peer_dev) in enumerate(gathered):
                if r == self.rank:
                    self.buckets.append(bucket)
                    self.flag_blocks.append(self.flags)
                    continue
                _lib.check(_lib.load().aecf_peer_enable_access(dev.index or 0, peer_dev), f"peer access {dev.index} -> {peer_dev}")
                self.buckets.append(b[0](*b[1]))             # rebuild_cuda_tensor: rank r's memory mapped into this process
                self.flag_blocks.append(f[0](*f[1]))
            dist.barrier(group=group)                        # everybody has mapped everything before anybody signals
        self._ops = ops

    def __call__(self) -> None:
        """Enqueue the all-reduce on the current stream of the bucket's device."""
        if self.world == 1:
            return
        with torch.cuda.device(self.bucket.device):
            self._ops.peer_allreduce(self.buckets, self.flag_blocks, self.rank, average=self.average)


class GradientSync:
    """All-reduce (mean by default) of the fusion parameter gradients of one pool + its query.

    ``attach()`` hooks the pool: its backward then writes gradients into ``self.bucket`` and reports
    each one the moment it is final.  ``finish()`` -- call it after ``loss.backward()`` -- waits for the
    collectives and points every ``param.grad`` at its reduced slice of the bucket.
    """

    def __init__(self, pool, query: Optional[torch.nn.Parameter] = None, process_group=None,
                 average: bool = True, overlap: Optional[bool] = None, collective: Optional[str] = None):
        self.pool, self.query, self.group, self.average = pool, query, process_group, average
        if collective is None:
            collective = "peer" if os.environ.get("AECF_DP_PEER", "0") == "1" else "nccl"
        if collective not in ("nccl", "peer"):
            raise ValueError(f"collective must be 'nccl' or 'peer', got {collective!r}")
        self.collective = collective                     # 'nccl': torch.distributed all_reduce; 'peer': PeerAllReduce
        if overlap is None:
            overlap = os.environ.get("AECF_DP_OVERLAP", "0") == "1"
        self.overlap = overlap
        att = pool.attention
        self.params: Dict[str, torch.nn.Parameter] = {}
        for name, p in (("out_proj.bias", att.out_proj.bias), ("out_proj.weight", att.out_proj.weight),
                        ("in_proj_weight", att.in_proj_weight), ("in_proj_bias", att.in_proj_bias), ("query", query)):
            if p is not None:
                self.params[name] = p
        device, dtype = att.in_proj_weight.device, att.in_proj_weight.dtype
        if any(p.dtype != dtype for p in self.params.values()):
            raise ValueError("GradientSync needs the fusion query and the pool parameters in one dtype")
        self.slices: Dict[str, slice] = {}
        n = 0
        for name in PARAM_ORDER:
            if name in self.params:
                k = self.params[name].numel()
                self.slices[name] = slice(n, n + k)
                n += (k + 7) // 8 * 8                       # keep every slice 16-byte aligned
            if name == EARLY[-1]:
                self.early_end = n
        self.bucket = torch.zeros(n, dtype=dtype, device=device)
        self.views = {name: self.bucket[sl].view(self.params[name].shape) for name, sl in self.slices.items()}
        self.cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.pending: List[object] = []
        self.reported: set = set()
        self.reduced_upto = 0
        self.enabled = True                              # False: gradients stay local (measurements, gradient accumulation)
        self.peer = None
        if collective == "peer" and self.cuda and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            self.peer = PeerAllReduce(self.bucket, process_group, average)     # one kernel for the whole bucket
            self.overlap = False

    # -- wiring -----------------------------------------------------------------------------
    def attach(self) -> "GradientSync":
        self.pool._grad_ready = self.on_ready
        self.pool._grad_buffers = self.views            # the fused backward writes its gradients here
        if self.query is not None:
            self.query.register_hook(self._query_hook)
        return self

    def _query_hook(self, grad: torch.Tensor) -> torch.Tensor:
        view = self.views["query"]
        if grad.data_ptr() != view.data_ptr():          # the shared-query bypass did not apply: copy in
            view.copy_(grad.reshape(view.shape))
        self.on_ready("query", view)
        return grad

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def set_shard(self, global_batch: int) -> Tuple[int, int]:
        """Point the pool's Philox row offset at this rank's shard; returns (row0, rows)."""
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        row0, rows = shard_rows(global_batch, rank, self.world_size)
        self.pool.row_offset = row0
        return row0, rows

    def _reduce(self, lo: int, hi: int, side_stream: bool) -> None:
        if hi <= lo:
            return
        view = self.bucket[lo:hi]
        avg = self.average and dist.get_backend(self.group) == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        if self.cuda and side_stream:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.bucket.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                self.pending.append(dist.all_reduce(view, op=op, group=self.group, async_op=True))
        else:
            self.pending.append(dist.all_reduce(view, op=op, group=self.group, async_op=True))
        if self.average and not avg:
            self._scale_later = True

    # -- called from the backward (autograd worker thread) -----------------------------------
    def on_ready(self, name: str, grad: torch.Tensor) -> None:
        if not self.enabled or name not in self.slices or name in self.reported or self.world_size == 1:
            return                                                     # single process: autograd's .grad is final
        view = self.views[name]
        if grad.data_ptr() != view.data_ptr():                         # produced elsewhere: copy into the bucket
            view.copy_(grad.reshape(view.shape))
        self.reported.add(name)
        if self.overlap and self.reduced_upto == 0 and all(n in self.reported for n in EARLY if n in self.slices):
            self._reduce(0, self.early_end, side_stream=True)
            self.reduced_upto = self.early_end

    # -- called by the training loop after loss.backward() -----------------------------------
    def finish(self) -> None:
        """Reduce what is left, wait for the collectives and store the results in ``param.grad``."""
        if not self.reported:
            return
        if self.peer is not None:
            self.peer()                                  # in place, on the compute stream, complete when it returns on-stream
            for name in self.reported:
                self.params[name].grad = self.views[name]
            self.reported.clear()
            return
        self._reduce(self.reduced_upto, self.bucket.numel(), side_stream=False)
        for work in self.pending:
            work.wait()
        if self.cuda and self.overlap:
            torch.cuda.current_stream(self.bucket.device).wait_stream(self.comm_stream)
        if getattr(self, "_scale_later", False):
            self.bucket.mul_(1.0 / self.world_size)
            self._scale_later = False
        for name in self.reported:
            self.params[name].grad = self.views[name]
        self.pending.clear()
        self.reported.clear()
        self.reduced_upto = 0
