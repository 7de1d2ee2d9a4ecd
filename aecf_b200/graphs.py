"""CUDA-graph capture of a whole training step of the fusion pool.

One step of the hot path is ~25 kernels of 5-150 us each; enqueueing them from Python costs about as much
host time as the GPU needs to run them, and with 8 ranks sharing the host's cores the
step becomes host-bound.  Captured once and replayed, the step costs one ``cudaGraphLaunch`` on the host.

What capture needs from the library, and gets:
  * no host synchronisation, allocation or state inside the C ABI (include/aecf_b200.h conventions);
  * masks and dropout that CHANGE between replays: a by-value Philox (seed, offset) would be frozen into the
    graph, so captured forwards read ``{seed, offset}`` from a small device tensor at run time
    (``aecf_pool_desc::rng_state``) and advance it with a captured kernel -- replay i draws what the eager
    path would have drawn at its i-th call after ``prepare()``;
  * the data-parallel all-reduce (``aecf_b200.dp.GradientSync``) enqueued inside the same capture, still on
    its side stream, so the overlap with the backward survives.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Optional

import torch

from .layers import _rng

__all__ = ["prepare", "capture_step", "GraphedStep"]


def prepare(device: Optional[torch.device] = None) -> torch.Tensor:
    """Create or refresh the device-side Philox {seed, next offset} pair from the host-side state
    (``torch.manual_seed`` / ``aecf_b200.set_rng_state``).  Call outside capture; returns the int64[2] tensor."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return _rng.prepare_device_state(torch.device(device))


class GraphedStep:
    """``fn()`` -- typically forward + loss + backward (+ GradientSync.finish) on STATIC input tensors -- captured
    into one CUDA graph.  Calling the object replays it and returns what ``fn`` returned at capture time
    (static tensors, rewritten by every replay).  Parameter ``.grad`` tensors produced inside the capture are
    static too: do not set them to None between replays.

    ``reset`` runs before every warm-up iteration and once more right before the capture; use it to drop
    gradients (``set_to_none=True``) so that the captured backward WRITES them instead of accumulating.
    """

    def __init__(self, fn: Callable[[], Any], *, reset: Optional[Callable[[], None]] = None, warmup: int = 3,
                 device: Optional[torch.device] = None, capture_error_mode: Optional[str] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("aecf_b200.graphs needs a CUDA device: there is no CPU path")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        mode = capture_error_mode or os.environ.get("AECF_GRAPH_CAPTURE_MODE", "thread_local")
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                      # warm-up off the default stream, as capture requires
            for _ in range(max(0, warmup)):
                if reset is not None:
                    reset()
                fn()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if reset is not None:
            reset()
        prepare(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self.result = fn()

    def __call__(self) -> Any:
        self.graph.replay()
        return self.result


def capture_step(fn: Callable[[], Any], **kwargs) -> GraphedStep:
    """Shorthand for ``GraphedStep(fn, **kwargs)``."""
    return GraphedStep(fn, **kwargs)
