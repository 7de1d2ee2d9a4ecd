"""Shared helpers for the parity tests (CPU and GPU)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np
import torch

from oracle import aecf_oracle as oracle
from tests.golden.cases import Case, build_inputs, masking_kwargs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(case: Case) -> Dict[str, np.ndarray]:
    with np.load(os.path.join(GOLDEN_DIR, case.name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    assert json.loads(str(g["meta"])) == case.meta(), "fixture is stale: rerun tests/golden/make_golden.py"
    return g


def score_bias_from_kpm(kpm: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
    """torch turns a boolean key_padding_mask into an additive -inf mask (functional.py:6608-6620)."""
    if kpm is None:
        return None
    B, M = kpm.shape
    return torch.zeros(B, 1, 1, M, dtype=dtype).masked_fill(kpm.view(B, 1, 1, M), float("-inf"))


def run_oracle(case: Case, inp: Optional[dict] = None, dtype: Optional[torch.dtype] = None,
               storage: Optional[torch.dtype] = None, fold_key: bool = False):
    """Oracle forward + closed-form backward on a golden case; returns (forward result, grads)."""
    inp = inp if inp is not None else build_inputs(case, dtype)
    B, D, S = case.B, case.D, case.S
    q = inp["query"] if S > 1 else inp["query0"].expand(B, 1, D)       # S > 1: a query vector per (b, s)
    value = inp.get("value")
    fwd = oracle.pool_forward(
        q, inp["x"], value, inp["in_proj_weight"], inp["in_proj_bias"],
        inp["out_proj.weight"], inp["out_proj.bias"], case.H,
        dropout_p=case.dropout, training=case.training, u_drop=inp["u_drop"], u_mask=inp["u_mask"],
        score_bias=score_bias_from_kpm(inp.get("key_padding_mask"), inp["x"].dtype),
        masking=masking_kwargs(case), storage=storage, fold_key=fold_key, per_row_query_storage=S > 1)
    grads = oracle.pool_backward(
        q, inp["x"], value, inp["in_proj_weight"], inp["out_proj.weight"], case.H, fwd.saved,
        inp["grad_out"], grad_pooled=inp["grad_pooled"] if case.pooled_grad else None,
        grad_entropy=None if case.training else torch.full((B, S), 0.5, dtype=inp["x"].dtype),
        dropout_p=case.dropout, training=case.training, storage=storage, fold_key=fold_key,
        per_row_query_storage=S > 1)
    if S == 1:
        grads["query0"] = grads.pop("query").sum(0, keepdim=True)
    return fwd, grads


def assert_close(name, got, want, rtol, atol=None):
    got = torch.as_tensor(got).double().reshape(-1)
    want = torch.as_tensor(want).double().reshape(-1)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    if atol is None:                       # relative to the tensor's own scale
        atol = rtol * max(float(want.abs().max()), 1e-30)
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    bad = err > bound
    assert not bool(bad.any()), (
        f"{name}: {int(bad.sum())}/{got.numel()} beyond rtol={rtol} atol={atol:.3e}; "
        f"max abs err {float(err.max()):.3e} (scale {float(want.abs().max()):.3e})")
