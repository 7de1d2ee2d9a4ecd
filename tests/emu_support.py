"""Runs the drop-in module API on CPU tensors against the HOST EMULATION of the CUDA sources (tests/cuda_emu) --
TEST INFRASTRUCTURE ONLY.

``libaecf_emu.so`` is the library's own sources (aecf_b200/csrc) compiled by g++ with every CUDA thread a fiber; it
exports the same C ABI.  The ``cuda_emulation`` fixture points ``aecf_b200._lib`` at it for the duration of one
test and lifts the module's CUDA-only guards, so the parity tests written for the GPU (tests/test_gpu_*.py) can be
called on CPU tensors.  What this proves and does not prove is stated in tests/cuda_emu/cuda_emu.h.
"""
import ctypes
import os
import subprocess

import pytest
import torch

from aecf_b200 import _lib, ops

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "cuda_emu")
EMU_LIB = os.path.join(EMU_DIR, "build", "libaecf_emu.so")

_emu = None


def load_emulation():
    """Build (make is incremental) and load the emulated library once per process."""
    global _emu
    if _emu is None:
        import shutil
        cuda_inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
        if shutil.which("g++") is None or shutil.which("make") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
            pytest.skip("the host emulation needs g++, make and the CUDA headers")
        res = subprocess.run(["make", "-C", EMU_DIR, "-j", str(min(8, os.cpu_count() or 1))], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building tests/cuda_emu failed:\n" + res.stdout[-3000:] + res.stderr[-3000:])
        lib = ctypes.CDLL(EMU_LIB)
        _lib._declare(lib)
        assert lib.aecf_abi_version() == _lib.ABI_VERSION
        _emu = lib
    return _emu


@pytest.fixture
def cuda_emulation(monkeypatch):
    """Inside the test, aecf_b200 runs on CPU tensors through the emulated kernels."""
    lib = load_emulation()
    monkeypatch.setattr(_lib, "_lib", lib)
    monkeypatch.setattr(ops, "require_cuda", lambda *tensors: torch.device("cpu"))
    monkeypatch.setattr(ops, "_stream", lambda dev: None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    return lib


def enable_in_this_process():
    """The fixture's patches without pytest, for worker processes (tests/test_emu_dp_gloo.py): permanent."""
    lib = load_emulation()
    _lib._lib = lib
    ops.require_cuda = lambda *tensors: torch.device("cpu")
    ops._stream = lambda dev: None
    torch.cuda.is_current_stream_capturing = lambda: False
    torch.cuda.synchronize = lambda *a, **k: None
    return lib
