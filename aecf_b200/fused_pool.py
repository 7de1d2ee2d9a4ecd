"""autograd glue for the fused fusion pool: every tensor operation below is a C-ABI kernel launch.

Forward  (reference aecf/AECFLayer.py:515-541 over torch/nn/functional.py:5847-5865, 6630-6659):
    q-proj GEMM -> packed KV GEMM -> fused pool kernel (scores, softmax, dropout, value sum,
    head mean, curriculum mask) -> out-proj GEMM
Backward (autograd of the same lines in the reference):
    colsum + dWo GEMM + dctx GEMM -> fused recompute pool backward -> dX GEMM + dWkv GEMM
    -> the small query-side products
Nothing the forward computed is kept except the projected K/V and the context.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch

from . import _lib, ops


@dataclass
class PoolConfig:
    num_heads: int
    dropout_p: float
    training: bool              # attention module in training mode (dropout on)
    masking: int                # 0 none, 1 CurriculumMasking in training mode, 2 in eval mode
    base_mask_prob: float = 0.15
    entropy_target: float = 0.7
    min_active: int = 1
    seed: int = 0
    offset: int = 0
    row0: int = 0
    q_shared: bool = True
    seq_first: bool = False     # key/value are [M, B, D] (batch_first=False), used in place
    want_mask_bits: bool = False
    bias_strides: tuple = (0, 0)
    # data-parallel hook: called in backward with (name, grad tensor) as soon as a parameter
    # gradient is final, so that its all-reduce overlaps the rest of the backward
    grad_ready: Optional[Callable[[str, torch.Tensor], None]] = None


def _rows(x3d: torch.Tensor) -> torch.Tensor:
    """[A, B, D] contiguous -> [A*B, D] view."""
    return x3d.reshape(x3d.shape[0] * x3d.shape[1], x3d.shape[2])


class FusedPoolFunction(torch.autograd.Function):
    """out, pooled, entropy, mask_rate, masked, mask_bits = f(query, key, value, params...)."""

    @staticmethod
    def forward(ctx, q_src, key, value, in_w, in_b, out_w, out_b, score_bias, cfg: PoolConfig):
        dev = ops.require_cuda(q_src, key, value, in_w, in_b, out_w, out_b, score_bias)
        dt = key.dtype
        D = key.shape[-1]
        if cfg.seq_first:
            M, B = key.shape[0], key.shape[1]
            kv_strides = (2 * D, B * 2 * D)
        else:
            B, M = key.shape[0], key.shape[1]
            kv_strides = (0, 0)
        rows = B * M
        b_q = b_kv = b_k = b_v = None
        if in_b is not None:
            b_q, b_kv, b_k, b_v = in_b[:D], in_b[D:], in_b[D:2 * D], in_b[2 * D:]

        # ---- in-projection (torch/nn/functional.py:5847-5865) --------------------------------
        if cfg.q_shared:
            # one query for the whole batch: project it once, in fp32
            qp = ops.linear(q_src.reshape(1, D), in_w[:D], b_q, out_dtype=torch.float32, name="q_proj").reshape(D)
        else:
            qp = ops.linear(q_src.reshape(B, D), in_w[:D], b_q, name="q_proj")
        kv = torch.empty((rows, 2 * D), dtype=dt, device=dev)
        if value is None:
            ops.linear(_rows(key), in_w[D:], b_kv, out=kv, name="kv_proj")
        else:                                   # separate value tensor: two GEMMs into the halves
            ops.linear(_rows(key), in_w[D:2 * D], b_k, out=kv[:, :D], ldc=2 * D, name="k_proj")
            ops.linear(_rows(value), in_w[2 * D:], b_v, out=kv[:, D:], ldc=2 * D, name="v_proj")

        # ---- fused pool ------------------------------------------------------------------------
        desc = ops.make_pool_desc(
            dev, dt, batch=B, num_tokens=M, embed_dim=D, num_heads=cfg.num_heads, training=cfg.training,
            masking=cfg.masking, min_active=cfg.min_active, q_is_shared=cfg.q_shared,
            base_mask_prob=cfg.base_mask_prob, entropy_target=cfg.entropy_target, dropout_p=cfg.dropout_p,
            seed=cfg.seed, offset=cfg.offset, row0=cfg.row0, bias_strides=cfg.bias_strides, kv_strides=kv_strides)
        attn, pooled, entropy, mask_rate, masked, bits = ops.pool_fwd(desc, qp, kv, score_bias, cfg.want_mask_bits)

        # ---- out-projection (:6653) -------------------------------------------------------------
        out = ops.linear(attn, out_w, out_b, name="out_proj")

        ctx.cfg, ctx.desc = cfg, desc
        ctx.has_value = value is not None
        ctx.has_in_bias, ctx.has_out_bias = in_b is not None, out_b is not None
        ctx.shape = (B, M, D)
        ctx.q_shape = q_src.shape
        ctx.save_for_backward(q_src, key, value, in_w, out_w, qp, kv, attn, score_bias)
        if bits is None:
            bits = torch.empty(0, dtype=torch.uint8, device=dev)
        if cfg.masking != 2:                            # entropy is detached in training mode (reference :278)
            ctx.mark_non_differentiable(mask_rate, masked, bits, entropy)
        else:
            ctx.mark_non_differentiable(mask_rate, masked, bits)
        return out, pooled, entropy, mask_rate, masked, bits

    @staticmethod
    def backward(ctx, g_out, g_pooled, g_entropy, _g_rate, _g_masked, _g_bits):
        cfg: PoolConfig = ctx.cfg
        q_src, key, value, in_w, out_w, qp, kv, attn, score_bias = ctx.saved_tensors
        B, M, D = ctx.shape
        dev, dt = key.device, key.dtype
        need_q, need_key, need_value, need_in_w, need_in_b, need_out_w, need_out_b = ctx.needs_input_grad[:7]
        notify = cfg.grad_ready

        if g_out is None:
            g = torch.zeros((B, D), dtype=dt, device=dev)
        else:
            g = g_out.reshape(B, D)
            if g.dtype != dt or not g.is_contiguous():
                g = g.to(dt).contiguous()

        # ---- out-projection backward: its two parameter gradients are ready first ------------
        d_out_w = d_out_b = None
        if need_out_b and ctx.has_out_bias:
            d_out_b = ops.colsum(g, name="d_out_bias")
            if notify:
                notify("out_proj.bias", d_out_b)
        if need_out_w:
            d_out_w = ops.matmul_tn(g, attn, name="d_out_weight")
            if notify:
                notify("out_proj.weight", d_out_w)
        d_attn = ops.matmul_nn(g, out_w, name="d_ctx")

        # ---- fused recompute backward of the pool -----------------------------------------------
        d_pooled = None
        if g_pooled is not None:
            d_pooled = g_pooled.reshape(B, M).to(torch.float32).contiguous()
        d_entropy = None
        if g_entropy is not None and cfg.masking == 2:
            d_entropy = g_entropy.reshape(B).to(torch.float32).contiguous()
        d_kv, d_qp, d_bias_kv = ops.pool_bwd(ctx.desc, qp, kv, score_bias, d_attn, d_pooled, d_entropy)
        d_kv2 = d_kv.reshape(B * M, 2 * D)

        # ---- in-projection backward ----------------------------------------------------------------
        d_key = d_value = None
        key2, value2 = _rows(key), (_rows(value) if ctx.has_value else None)
        if not ctx.has_value:
            if need_key:
                d_key = ops.matmul_nn(d_kv2, in_w[D:], name="d_x").reshape(key.shape)
        else:
            if need_key:
                d_key = ops.matmul_nn(d_kv2[:, :D], in_w[D:2 * D], name="d_key").reshape(key.shape)
            if need_value:
                d_value = ops.matmul_nn(d_kv2[:, D:], in_w[2 * D:], name="d_value").reshape(value.shape)

        d_in_w = d_in_b = d_q = None
        if need_in_w:
            d_in_w = torch.empty_like(in_w)
            if not ctx.has_value:
                ops.matmul_tn(d_kv2, key2, out=d_in_w[D:], name="d_kv_weight")
            else:
                ops.matmul_tn(d_kv2[:, :D], key2, out=d_in_w[D:2 * D], name="d_k_weight")
                ops.matmul_tn(d_kv2[:, D:], value2, out=d_in_w[2 * D:], name="d_v_weight")
        if cfg.q_shared:
            d_qp_row = d_qp.reshape(1, D)                     # fp32, already summed over the batch
            if need_in_w:     # dWq = d_qp (outer) q0
                ops.gemm(d_qp_row, q_src.reshape(1, D), m=D, n=D, k=1, a_layout=_lib.MN_MAJOR,
                         b_layout=_lib.MN_MAJOR, lda=D, ldb=D, out=d_in_w[:D], name="d_q_weight")
            if need_q:
                d_q = ops.matmul_nn(d_qp_row, in_w[:D], out_dtype=q_src.dtype, name="d_query").reshape(ctx.q_shape)
            d_bq = d_qp
        else:
            q2 = q_src.reshape(B, D)
            if need_in_w:
                ops.matmul_tn(d_qp, q2, out=d_in_w[:D], name="d_q_weight")
            if need_q:
                d_q = ops.matmul_nn(d_qp, in_w[:D], name="d_query").reshape(ctx.q_shape)
            d_bq = ops.colsum(d_qp, out_dtype=torch.float32, name="d_q_bias") if (need_in_b and ctx.has_in_bias) else None
        if need_in_w and notify:
            notify("in_proj_weight", d_in_w)
        if need_in_b and ctx.has_in_bias:
            d_in_b = torch.empty(3 * D, dtype=in_w.dtype, device=dev)
            d_in_b[:D].copy_(d_bq)
            d_in_b[D:].copy_(d_bias_kv)
            if notify:
                notify("in_proj_bias", d_in_b)
        return d_q, d_key, d_value, d_in_w, d_in_b, d_out_w, d_out_b, None, None


class EntropyLossFunction(torch.autograd.Function):
    """CurriculumMasking.entropy_loss (reference aecf/AECFLayer.py:285-314) as two tiny kernels."""

    @staticmethod
    def forward(ctx, entropy, target: float):
        e = entropy.reshape(-1).to(torch.float32).contiguous()
        ctx.save_for_backward(e)
        ctx.target, ctx.shape, ctx.dtype = target, entropy.shape, entropy.dtype
        return ops.entropy_loss_fwd(e, target).to(entropy.dtype)

    @staticmethod
    def backward(ctx, g):
        (e,) = ctx.saved_tensors
        d = ops.entropy_loss_bwd(e, ctx.target, g.reshape(1).to(torch.float32).contiguous())
        return d.reshape(ctx.shape).to(ctx.dtype), None


class EntropyFunction(torch.autograd.Function):
    """compute_entropy / eval-mode CurriculumMasking entropy on caller-supplied weights
    (reference aecf/AECFLayer.py:101-128, 150-156): clamp(-sum xlogy(w, w), 0, log L), differentiable."""

    @staticmethod
    def forward(ctx, weights):
        w2 = weights.reshape(-1, weights.shape[-1]).to(torch.float32).contiguous()
        _, entropy, _ = ops.curriculum_mask(w2, 3, want_masked=False)
        ctx.save_for_backward(w2)
        ctx.shape, ctx.dtype = weights.shape, weights.dtype
        return entropy.reshape(weights.shape[:-1]).to(weights.dtype)

    @staticmethod
    def backward(ctx, g):
        (w2,) = ctx.saved_tensors
        d = ops.entropy_bwd(w2, g.reshape(-1).to(torch.float32).contiguous())
        return d.reshape(ctx.shape).to(ctx.dtype)
