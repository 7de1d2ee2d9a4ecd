"""Shared definition of the parity cases: shapes, seeds and the deterministic input builder.

Used by ``make_golden.py`` (which runs the *reference* on these inputs, in the build
container only) and by the tests (which rebuild the same inputs from the seeds and
compare the oracle / the CUDA path with the stored reference outputs).  Inputs and
parameters come from ``oracle.philox.normal`` so that they do not depend on any
torch RNG stream.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Optional

import numpy as np
import torch

from oracle import philox

PHILOX_SEED = 0x5EED


@dataclass(frozen=True)
class Case:
    name: str
    B: int
    M: int
    D: int
    H: int
    dropout: float = 0.0
    base_mask_prob: float = 0.15
    entropy_target: float = 0.7
    min_active: int = 1
    training: bool = True
    dtype: str = "float32"
    peak: float = 2.0             # std of the per-(b,m) score offset that makes attention peaked
    kpm: bool = False             # use a key_padding_mask
    separate_value: bool = False  # value is a different tensor from key
    pooled_grad: bool = False     # loss also depends on info['attention_weights']
    offset: int = 0               # Philox call offset
    row0: int = 0                 # global index of the first row
    data_seed: int = 1
    full_grads: bool = True       # store complete weight gradients in the fixture
    S: int = 1                    # fusion queries per sample; S > 1: every (b, s) has its own query vector

    def meta(self) -> dict:
        return asdict(self)


CASES = [
    # BASELINE.json configs[0]: create_fusion_pool(512, 3, 0.15) -> num_heads defaults to 1
    Case("config1_d512_h1_m3", B=32, M=3, D=512, H=1, full_grads=False, data_seed=11, peak=1.0),
    Case("d64_h8_m3", B=64, M=3, D=64, H=8, pooled_grad=True, data_seed=12),
    Case("d64_h8_m3_dropout", B=64, M=3, D=64, H=8, dropout=0.1, pooled_grad=True, offset=3, data_seed=13),
    # x-ray example geometry: hidden 256, 4 heads, 2 modalities, ragged batch (xrays/train_xrays_example.py:133-138)
    Case("xray_d256_h4_m2", B=28, M=2, D=256, H=4, data_seed=14, row0=1000, peak=0.7, base_mask_prob=0.3),
    Case("d128_h4_m8_minactive2", B=48, M=8, D=128, H=4, base_mask_prob=0.9, min_active=2, data_seed=15),
    Case("d128_h2_m5_heavy_mask", B=40, M=5, D=128, H=2, base_mask_prob=1.0, min_active=1, data_seed=16),
    Case("d64_h4_m4_kpm", B=32, M=4, D=64, H=4, kpm=True, data_seed=17),
    Case("d64_h8_m3_eval", B=16, M=3, D=64, H=8, training=False, pooled_grad=True, data_seed=18),
    Case("d64_h2_m3_separate_value", B=24, M=3, D=64, H=2, separate_value=True, dropout=0.2, data_seed=19),
    Case("d64_h8_m3_fp64", B=32, M=3, D=64, H=8, dtype="float64", dropout=0.1, pooled_grad=True, data_seed=20),
    Case("d32_h1_m1_single_token", B=8, M=1, D=32, H=1, data_seed=21),
    Case("d256_h16_m6", B=20, M=6, D=256, H=16, dropout=0.05, min_active=3, base_mask_prob=0.5, data_seed=22),
]
# several fusion queries per sample (target length S > 1): per-(b, s) queries
MULTI_QUERY_CASES = [
    Case("s2_d64_h8_m3_dropout", B=24, S=2, M=3, D=64, H=8, dropout=0.1, pooled_grad=True, offset=5, data_seed=23),
    Case("s3_d128_h4_m5_minactive2", B=16, S=3, M=5, D=128, H=4, base_mask_prob=0.8, min_active=2, row0=70, data_seed=24),
    Case("s2_d64_h4_m4_kpm_eval", B=12, S=2, M=4, D=64, H=4, kpm=True, training=False, pooled_grad=True, data_seed=25),
    Case("s4_d256_h2_m2", B=10, S=4, M=2, D=256, H=2, base_mask_prob=0.5, peak=1.0, data_seed=26, full_grads=False),
]
CASES_BY_NAME = {c.name: c for c in CASES + MULTI_QUERY_CASES}


def _t(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def build_inputs(case: Case, dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
    """Parameters, fusion query, modality tokens, upstream gradients and the injected uniforms."""
    dt = dtype if dtype is not None else getattr(torch, case.dtype)
    D, M, B, H, S = case.D, case.M, case.B, case.H, case.S
    s = case.data_seed * 1000
    Wi = philox.normal(s + 1, (3 * D, D)) * math.sqrt(2.0 / (4 * D))     # ~xavier scale of MHA init
    bi = philox.normal(s + 2, (3 * D,)) * 0.05
    Wo = philox.normal(s + 3, (D, D)) / math.sqrt(D)
    bo = philox.normal(s + 4, (D,)) * 0.05
    q0 = philox.normal(s + 5, (1, 1, D)) * math.sqrt(2.0 / D)            # create_fusion_pool init scale
    x = philox.normal(s + 6, (B, M, D))
    # Make attention peaked: push every token along the direction that
    # raises its score in all heads at once, by a random per-(b, m) amount.
    qp = q0.reshape(D) @ Wi[:D].T + bi[:D]
    direction = Wi[D:2 * D].T @ qp
    norm = np.linalg.norm(direction)
    direction = direction / norm
    # a unit step along `direction` moves the head-averaged score by (scale / H) * |Wk^T qp|
    per_unit = norm / (H * math.sqrt(D // H))
    gain = philox.normal(s + 7, (B, M, 1)) * case.peak / per_unit
    x = x + gain * direction
    out = {
        "in_proj_weight": _t(Wi, dt), "in_proj_bias": _t(bi, dt),
        "out_proj.weight": _t(Wo, dt), "out_proj.bias": _t(bo, dt),
        "query0": _t(q0, dt), "x": _t(x, dt),
        "grad_out": _t(philox.normal(s + 8, (B, S, D)), dt),
        "grad_pooled": _t(philox.normal(s + 9, (B, S, M)), dt),
        # one Philox row per (b, s) pair: global row (row0 + b) * S + s  (include/aecf_b200.h)
        "u_mask": torch.from_numpy(philox.mask_uniforms(PHILOX_SEED, case.offset, case.row0 * S, B * S, M)
                                   .reshape((B, S, M) if S > 1 else (B, M))),
        "u_drop": torch.from_numpy(np.ascontiguousarray(
            philox.dropout_uniforms(PHILOX_SEED, case.offset, case.row0 * S, B * S, H, M)
            .reshape(B, S, H, M).transpose(0, 2, 1, 3)).reshape((B, H, S, M) if S > 1 else (B, H, M))),
    }
    if S > 1:   # per-(b, s) queries around the fusion query (which alone fixes the peaked direction above)
        out["query"] = _t(q0 + philox.normal(s + 12, (B, S, D)) * math.sqrt(2.0 / D), dt)
    if case.separate_value:
        out["value"] = _t(philox.normal(s + 10, (B, M, D)), dt)
    if case.kpm:
        pad = philox.normal(s + 11, (B, M)) > 0.6
        pad[:, 0] = False                      # never pad a whole row: torch would return NaN
        out["key_padding_mask"] = torch.from_numpy(pad)
    return out


def masking_kwargs(case: Case) -> dict:
    return dict(base_mask_prob=case.base_mask_prob, entropy_target=case.entropy_target,
                min_active=case.min_active)
