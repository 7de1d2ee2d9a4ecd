"""Philox4x32-10 counter-based RNG in numpy -- TEST INFRASTRUCTURE ONLY.

This is the CPU side of the shared random stream that the parity harness injects
into the reference and that the CUDA kernels in
``aecf_b200/csrc/philox.cuh`` generate on the fly.  Nothing under ``aecf_b200/``
may import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

Stream contract (identical text in ``include/aecf_b200.h``):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (row & 0xffffffff, row >> 32, offset & 0xffffffff,
               (stream << 28) | (head << 4) | block)
    stream  = 0 for the curriculum mask (head = 0), 1 for attention dropout
    block   = m // 4 (< 16), and lane m % 4 of the 4x32-bit output is the draw for
              token m  (the pool's M <= 8 needs blocks 0 and 1)
    uniform = float32(x) * 2**-32 + 2**-33      (curand_uniform: in (0, 1])
    mask keeps token m      iff  u <= keep_prob          (aecf/AECFLayer.py:204,
                                  torch CUDA bernoulli convention,
                                  ATen/native/cuda/DistributionTemplates.h:608-650)
    dropout keeps (h, m)    iff  u >= p_drop             (torch/nn/functional.py:6645)

``row`` is the GLOBAL sample index, which is what makes an N-rank batch-sharded
run reproduce the 1-rank masks bit for bit.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_LO = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

STREAM_MASK = 0
STREAM_DROPOUT = 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Ten rounds of Philox-4x32 on broadcastable uint32 arrays.

    Follows the published Random123 algorithm (Salmon et al., SC'11): each round
    multiplies c0 and c2 by the two round constants, swaps/xors the halves with
    the key, and the key is bumped by the Weyl constants between rounds.
    """
    c0, c1, c2, c3 = (np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _S32).astype(np.uint32)
            lo0 = (p0 & _LO).astype(np.uint32)
            hi1 = (p1 >> _S32).astype(np.uint32)
            lo1 = (p1 & _LO).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r != 9:
                k0 = (k0 + _W0).astype(np.uint32)
                k1 = (k1 + _W1).astype(np.uint32)
    return c0, c1, c2, c3


def uniform_from_bits(x):
    """curand_uniform convention: float32(x) * 2^-32 + 2^-33, result in (0, 1]."""
    xf = np.asarray(x, dtype=np.uint32).astype(np.float32)
    return xf * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)


def _draw(seed, offset, rows, stream, head, num_tokens):
    """uint32 draws of shape rows.shape + (num_tokens,) for one (stream, head)."""
    if num_tokens > 64:
        raise ValueError("the stream contract covers at most 64 tokens (4-bit block index)")
    if not 0 <= int(offset) < 2 ** 32:
        raise ValueError("offset must fit in 32 bits")
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    rows = np.asarray(rows, dtype=np.uint64)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32(seed >> 32)
    c0 = (rows & _LO).astype(np.uint32)
    c1 = (rows >> _S32).astype(np.uint32)
    c2 = np.uint32(int(offset))
    out = np.empty(rows.shape + (4 * ((num_tokens + 3) // 4),), dtype=np.uint32)
    for block in range((num_tokens + 3) // 4):
        c3 = np.uint32((int(stream) << 28) | (int(head) << 4) | block)
        r = philox4x32_10(c0, c1, c2, c3, k0, k1)
        for lane in range(4):
            out[..., 4 * block + lane] = r[lane]
    return out[..., :num_tokens]


def mask_uniforms(seed, offset, row0, batch, num_tokens):
    """U_mask[b, m] for global rows row0 .. row0+batch-1 (float32, in (0,1])."""
    rows = np.arange(batch, dtype=np.uint64) + np.uint64(row0)
    return uniform_from_bits(_draw(seed, offset, rows, STREAM_MASK, 0, num_tokens))


def dropout_uniforms(seed, offset, row0, batch, num_heads, num_tokens):
    """U_drop[b, h, m] (float32, in (0,1])."""
    rows = np.arange(batch, dtype=np.uint64) + np.uint64(row0)
    out = np.empty((batch, num_heads, num_tokens), dtype=np.float32)
    for h in range(num_heads):
        out[:, h, :] = uniform_from_bits(
            _draw(seed, offset, rows, STREAM_DROPOUT, h, num_tokens))
    return out


# --- deterministic synthetic data (tests + golden generation share it) --------

def normal(seed, shape, stream=7):
    """Standard normals from the Philox stream via Box-Muller, float64.

    Used so that tests and the golden-vector generator build *identical* inputs
    and parameters from a (seed, shape) pair without depending on torch's RNG.
    """
    n = int(np.prod(shape))
    blocks = (n + 3) // 4
    idx = np.arange(blocks, dtype=np.uint64)
    r = philox4x32_10((idx & _LO).astype(np.uint32), (idx >> _S32).astype(np.uint32),
                      np.uint32(0), np.uint32(stream),
                      np.uint32(int(seed) & 0xFFFFFFFF), np.uint32((int(seed) >> 32) & 0xFFFFFFFF))
    bits = np.stack(r, axis=-1).astype(np.float64)              # [blocks, 4]
    u = (bits + 0.5) * (2.0 ** -32)                             # (0, 1)
    rad = np.sqrt(-2.0 * np.log(u[:, 0::2]))
    ang = 2.0 * np.pi * u[:, 1::2]
    z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=-1).reshape(-1)
    return z[:n].reshape(shape)
