"""Replaces the reference's two RNG draws by injected uniforms -- shared by the fixture generator
(make_golden.py) and the live comparison (tests/test_emu_vs_reference_live.py).  Both call sites resolve the function at
call time: ``torch.bernoulli`` (reference aecf/AECFLayer.py:204) and ``torch.nn.functional.dropout``
(torch/nn/functional.py:6645)."""
import torch


class inject_uniforms:
    """Context manager replacing the reference's two RNG draws by injected uniforms."""

    def __init__(self, u_mask: torch.Tensor, u_drop: torch.Tensor):
        self.u_mask, self.u_drop = u_mask, u_drop

    def __enter__(self):
        self._bern = torch.bernoulli
        self._drop = torch.nn.functional.dropout
        u_mask, u_drop = self.u_mask, self.u_drop

        def bernoulli(p, *a, **k):
            return (u_mask.view(p.shape).to(p.dtype) <= p).to(p.dtype)

        def dropout(w, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return w
            if p >= 1.0:
                return w * 0.0
            keep = (u_drop.reshape(w.shape).to(w.dtype) >= p).to(w.dtype)
            return w * keep / (1.0 - p)

        torch.bernoulli = bernoulli
        torch.nn.functional.dropout = dropout
        return self

    def __exit__(self, *exc):
        torch.bernoulli = self._bern
        torch.nn.functional.dropout = self._drop
