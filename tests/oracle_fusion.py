"""An oracle-backed stand-in for the four public names of `aecf` (CPU, torch autograd through
oracle/aecf_oracle.py).  Test infrastructure: lets the caller models of examples/models.py be run
once on the CUDA path and once on the oracle, with the same parameters and the same Philox draws."""
from __future__ import annotations

import torch
import torch.nn as nn

from oracle import aecf_oracle as oracle
from oracle import philox

_state = {"seed": 0, "offset": 0}


def set_rng_state(seed: int, offset: int = 0) -> None:
    _state["seed"], _state["offset"] = seed, offset


class CurriculumMasking(nn.Module):
    def __init__(self, base_mask_prob: float = 0.15, entropy_target: float = 0.7, min_active: int = 1):
        super().__init__()
        self.base_mask_prob, self.entropy_target, self.min_active = base_mask_prob, entropy_target, min_active
        self.register_buffer("_eps", torch.tensor(1e-8))
        self._last_seq_len = 2

    def entropy_loss(self, entropy):
        return oracle.entropy_loss(entropy, self._last_seq_len, self.entropy_target)


class _OutProj(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dim, dim))
        self.bias = nn.Parameter(torch.empty(dim))


class _Attention(nn.Module):
    def __init__(self, dim, dropout):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * dim))
        self.out_proj = _OutProj(dim)
        self.dropout = dropout


class MultimodalAttentionPool(nn.Module):
    def __init__(self, embed_dim, num_heads=1, dropout=0.0, bias=True, curriculum_masking=None, batch_first=True,
                 device=None, dtype=None):
        super().__init__()
        self.embed_dim, self.num_heads, self.curriculum_masking = embed_dim, num_heads, curriculum_masking
        self.attention = _Attention(embed_dim, dropout)

    def forward(self, query, key, value=None, return_info=False, **_):
        if value is key:
            value = None
        B, M, _d = key.shape
        cm, att = self.curriculum_masking, self.attention
        masking = None
        if cm is not None:
            masking = dict(base_mask_prob=cm.base_mask_prob, entropy_target=cm.entropy_target, min_active=cm.min_active)
            if cm.training and M > 1:
                cm._last_seq_len = M
        draws = (self.training and att.dropout > 0) or (cm is not None and cm.training)
        u_mask = u_drop = None
        if draws:
            seed, off = _state["seed"], _state["offset"]
            _state["offset"] += 1
            u_mask = torch.from_numpy(philox.mask_uniforms(seed, off, 0, B, M))
            u_drop = torch.from_numpy(philox.dropout_uniforms(seed, off, 0, B, self.num_heads, M))
        res = oracle.pool_forward(query, key, value, att.in_proj_weight, att.in_proj_bias, att.out_proj.weight,
                                  att.out_proj.bias, self.num_heads, dropout_p=att.dropout, training=self.training,
                                  u_drop=u_drop, u_mask=u_mask, masking=masking)
        info = {k: v for k, v in res.info.items() if k not in ("mask", "keep_prob")}
        if cm is None:
            info = {"attention_weights": res.pooled}
        return (res.out, info) if return_info else res.out


def create_fusion_pool(embed_dim, num_modalities, mask_prob=0.15, **kwargs):
    query = nn.Parameter(torch.empty(1, 1, embed_dim))
    pool = MultimodalAttentionPool(embed_dim, curriculum_masking=CurriculumMasking(mask_prob), **kwargs)
    return query, pool
