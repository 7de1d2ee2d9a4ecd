"""The tcgen05/TMEM/TMA projection GEMM on its own (impl forced, so nothing can silently fall back to
the SIMT kernel), against fp64 matmul of the same bf16 operands."""
import pytest
import torch

from aecf_b200 import _lib, ops
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    (128, 128, 64),        # one tile, one k-block
    (256, 256, 64),        # BN = 256
    (384, 256, 128),
    (200, 136, 72),        # ragged in every dimension (TMA zero fill / clipped stores)
    (392, 520, 200),
    (4096, 1024, 512),     # the KV in-projection shape, scaled down in rows
    (3000, 512, 1024),     # the dX shape
    (1024, 512, 8192),     # the dW_kv shape: few tiles, long reduction -> split-K
    (512, 512, 4096),      # the dW_o shape
    (1000, 512, 4100),     # split-K with a ragged last row block and a ragged k tail
]


def _operands(m, n, k, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    b = torch.randn(n, k, device=DEV, generator=g).bfloat16()
    bias = torch.randn(n, device=DEV, generator=g).bfloat16()
    return a, b, bias


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("a_layout", [_lib.K_MAJOR, _lib.MN_MAJOR], ids=["aK", "aMN"])
@pytest.mark.parametrize("b_layout", [_lib.K_MAJOR, _lib.MN_MAJOR], ids=["bK", "bMN"])
def test_tcgen05_gemm(shape, a_layout, b_layout):
    m, n, k = shape
    a, b, bias = _operands(m, n, k, m * 7 + n * 3 + k)
    want = a.double() @ b.double().t()
    a_mem = a if a_layout == _lib.K_MAJOR else a.t().contiguous()
    b_mem = b if b_layout == _lib.K_MAJOR else b.t().contiguous()
    lda, ldb = a_mem.shape[1], b_mem.shape[1]
    eligible = (lda * 2) % 16 == 0 and (ldb * 2) % 16 == 0 and (n * 2) % 16 == 0
    for out_dtype, use_bias in ((torch.float32, False), (torch.bfloat16, True)):
        kwargs = dict(m=m, n=n, k=k, a_layout=a_layout, b_layout=b_layout, lda=lda, ldb=ldb,
                      bias=bias if use_bias else None, out_dtype=out_dtype, impl=_lib.GEMM_TCGEN05)
        if not eligible:
            with pytest.raises(_lib.UnsupportedShapeError):
                ops.gemm(a_mem, b_mem, **kwargs)
            continue
        out = ops.gemm(a_mem, b_mem, **kwargs)
        torch.cuda.synchronize()
        ref = want + (bias.double() if use_bias else 0.0)
        tol = 1e-5 if out_dtype == torch.float32 else 1e-2      # fp32 accumulate; bf16 only rounds the output
        assert_close(f"tcgen05 {shape} a={a_layout} b={b_layout} out={out_dtype}", out.float().cpu(), ref.cpu(), tol)


def test_tcgen05_strided_output_and_operand_views():
    """K and V projections of a separate-value call land in the two halves of one [rows, 2D] buffer,
    and the backward reads those halves as strided operands."""
    m, n, k = 640, 256, 256
    a, b, bias = _operands(m, n, k, 5)
    buf = torch.zeros(m, 2 * n, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, m=m, n=n, k=k, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=k, ldb=k, bias=bias,
             out=buf[:, n:], ldc=2 * n, impl=_lib.GEMM_TCGEN05)
    want = a.double() @ b.double().t() + bias.double()
    assert_close("strided C", buf[:, n:].float().cpu(), want.cpu(), 1e-2)
    assert float(buf[:, :n].abs().max()) == 0.0
    # A = right half of buf (lda = 2n), reduced over rows against `a`: [n, k] = buf_half^T a
    got = ops.gemm(buf[:, n:], a, m=n, n=k, k=m, a_layout=_lib.MN_MAJOR, b_layout=_lib.MN_MAJOR, lda=2 * n, ldb=k,
                   out_dtype=torch.float32, impl=_lib.GEMM_TCGEN05)
    assert_close("strided MN-major A", got.cpu(), (buf[:, n:].double().t() @ a.double()).cpu(), 1e-5)


def test_tcgen05_is_what_auto_picks_for_the_projection_shapes():
    before = _lib.launch_count()
    a, b, _ = _operands(512, 256, 256, 9)
    auto = ops.gemm(a, b, m=512, n=256, k=256, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=256, ldb=256)
    forced = ops.gemm(a, b, m=512, n=256, k=256, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=256, ldb=256,
                      impl=_lib.GEMM_TCGEN05)
    assert torch.equal(auto, forced)
    assert _lib.launch_count() - before == 2
