"""CPU-side checks of the drop-in surface and of the C-ABI library (no kernel is launched here)."""
import ctypes
import os
import re

import pytest
import torch

import aecf
import aecf_b200
from aecf_b200 import _lib, ops
from aecf_b200.layers import _PhiloxState

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "aecf_b200.h")).read()
    declared = set(re.findall(r"AECF_API[^;(]*?\b(aecf_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.aecf_abi_version() == _lib.ABI_VERSION == 2
    assert b"sm_100a" in lib.aecf_build_info()
    assert lib.aecf_strerror(-2).decode().startswith("shape or dtype outside")
    # structs mirror the header: sizes are what the C compiler lays out (natural alignment)
    assert ctypes.sizeof(_lib.PoolDesc) == 128 and ctypes.sizeof(_lib.GemmDesc) == 88


def test_descriptor_validation_needs_no_gpu():
    lib = _lib.load()
    d = _lib.PoolDesc(device=0, dtype=_lib.BF16, batch=4, num_tokens=9, embed_dim=64, num_heads=8)
    assert lib.aecf_pool_fwd(ctypes.byref(d), 16, 16, None, 16, 16, None, None, None, None, None) == _lib.ERR_UNSUPPORTED
    d.num_tokens, d.num_heads = 3, 7                                    # 7 does not divide 64
    assert lib.aecf_pool_fwd(ctypes.byref(d), 16, 16, None, 16, 16, None, None, None, None, None) == _lib.ERR_INVALID
    d.num_heads = 8
    assert lib.aecf_pool_fwd(ctypes.byref(d), 8, 16, None, 16, 16, None, None, None, None, None) == _lib.ERR_ALIGNMENT
    assert lib.aecf_pool_bwd_workspace_bytes(ctypes.byref(d)) == 2048 * 3 * 64 * 4
    g = _lib.GemmDesc(device=0, m=4, n=4, k=4, lda=2, ldb=4, ldc=4)     # lda < k
    assert lib.aecf_gemm(ctypes.byref(g), 16, 16, None, 16, None, 0, None) == _lib.ERR_INVALID


def test_alias_package_exposes_the_reference_names():
    for name in ("CurriculumMasking", "MultimodalAttentionPool", "multimodal_attention_pool", "create_fusion_pool"):
        assert getattr(aecf, name) is getattr(aecf_b200, name)
    assert sorted(aecf.__all__) == sorted(["CurriculumMasking", "MultimodalAttentionPool",
                                           "multimodal_attention_pool", "create_fusion_pool"])
    assert aecf.__version__ == "0.1.0"


def test_constructor_validation_matches_reference_messages():
    with pytest.raises(ValueError, match=r"base_mask_prob must be in \(0, 1\], got 0.0"):
        aecf.CurriculumMasking(base_mask_prob=0.0)
    with pytest.raises(ValueError, match=r"entropy_target must be in \(0, 1\], got 1.5"):
        aecf.CurriculumMasking(entropy_target=1.5)
    with pytest.raises(ValueError, match="min_active must be >= 1, got 0"):
        aecf.CurriculumMasking(min_active=0)
    with pytest.raises(ValueError, match="embed_dim must be positive, got 0"):
        aecf.MultimodalAttentionPool(0)
    with pytest.raises(ValueError, match="num_heads must be positive, got 0"):
        aecf.MultimodalAttentionPool(8, num_heads=0)
    with pytest.raises(ValueError, match=r"embed_dim \(10\) must be divisible by num_heads \(3\)"):
        aecf.MultimodalAttentionPool(10, num_heads=3)
    with pytest.raises(ValueError, match=r"dropout must be in \[0, 1\], got 1.5"):
        aecf.MultimodalAttentionPool(8, dropout=1.5)
    with pytest.raises(ValueError, match="embed_dim must be a positive integer, got -1"):
        aecf.create_fusion_pool(-1, 2)
    with pytest.raises(ValueError, match="num_modalities must be a positive integer, got 0"):
        aecf.create_fusion_pool(8, 0)
    with pytest.raises(ValueError, match=r"mask_prob must be in \(0, 1\], got 2"):
        aecf.create_fusion_pool(8, 2, mask_prob=2)


def test_forward_validation_matches_reference():
    pool = aecf.MultimodalAttentionPool(16, num_heads=2)
    q, k = torch.randn(4, 1, 16), torch.randn(4, 3, 16)
    with pytest.raises(TypeError, match="Expected query to be torch.Tensor"):
        pool([1, 2], k)
    with pytest.raises(TypeError, match="Expected key to be torch.Tensor"):
        pool(q, None)
    with pytest.raises(TypeError, match="Expected value to be torch.Tensor or None"):
        pool(q, k, 3)
    with pytest.raises(ValueError, match="Expected 3D query tensor with batch_first=True, got 2D"):
        pool(q[0], k)
    with pytest.raises(ValueError, match="Key sequence length cannot be zero"):
        pool(q, k[:, :0])
    with pytest.raises(RuntimeError, match="incompatible with query shape"):
        pool(q, torch.randn(5, 3, 16))
    with pytest.raises(RuntimeError, match="Value shape .* incompatible with key shape"):
        pool(q, k, torch.randn(4, 2, 16))
    seq = aecf.MultimodalAttentionPool(16, num_heads=2, batch_first=False)
    with pytest.raises(ValueError, match="Expected 3D key tensor with batch_first=False, got 2D"):
        seq(q, k[0])
    with pytest.raises(RuntimeError, match="Shape mismatch"):
        seq(torch.randn(1, 4, 16), torch.randn(3, 5, 16))
    # valid shapes on CPU tensors: the fusion path has no CPU implementation and says so
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        pool(q, k)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        aecf.CurriculumMasking()(torch.softmax(torch.randn(4, 3), -1))
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        aecf.multimodal_attention_pool(q, k)


def test_state_dict_round_trips_with_torch_multihead_attention():
    torch.manual_seed(7)
    ours = aecf.MultimodalAttentionPool(32, num_heads=4, curriculum_masking=aecf.CurriculumMasking())
    torch.manual_seed(7)
    mha = torch.nn.MultiheadAttention(32, 4, batch_first=True)           # what the reference wraps (:399-407)
    want = {f"attention.{k}": v for k, v in mha.state_dict().items()}
    want["curriculum_masking._eps"] = torch.tensor(1e-8)
    got = ours.state_dict()
    assert list(got) == ["curriculum_masking._eps", "attention.in_proj_weight", "attention.in_proj_bias",
                         "attention.out_proj.weight", "attention.out_proj.bias"]
    for k in want:                                                       # same names, shapes AND default init
        assert torch.equal(got[k], want[k]), k
    other = aecf.MultimodalAttentionPool(32, num_heads=4, curriculum_masking=aecf.CurriculumMasking())
    missing, unexpected = other.load_state_dict(want, strict=True)
    assert not missing and not unexpected
    assert torch.equal(other.attention.in_proj_weight, mha.in_proj_weight)
    nobias = aecf.MultimodalAttentionPool(32, num_heads=4, bias=False)
    assert list(nobias.state_dict()) == ["attention.in_proj_weight", "attention.out_proj.weight"]


def test_factory_and_repr():
    torch.manual_seed(0)
    q, pool = aecf.create_fusion_pool(embed_dim=64, num_modalities=3, mask_prob=0.25, num_heads=8)
    torch.manual_seed(0)
    ref_q = torch.empty(1, 1, 64).normal_(0.0, (2.0 / 64) ** 0.5)
    assert isinstance(q, torch.nn.Parameter) and q.shape == (1, 1, 64) and torch.equal(q.detach(), ref_q)
    assert pool.num_heads == 8 and pool.embed_dim == 64 and pool.batch_first
    assert pool.curriculum_masking.base_mask_prob == 0.25 and pool.curriculum_masking._last_seq_len == 2
    assert "embed_dim=64, num_heads=8, batch_first=True, curriculum_masking=True" in repr(pool)
    assert "base_mask_prob=0.25, entropy_target=0.7, min_active=1" in repr(pool.curriculum_masking)
    _, default = aecf.create_fusion_pool(64, 2)
    assert default.num_heads == 1                                         # reference default (quirk D11)
    pool.curriculum_masking.base_mask_prob = 0.05                         # runtime mutation is allowed (README.md:341-350)
    pool.curriculum_masking = None                                        # and so is removing the stage (xrays:179-187)
    assert "curriculum_masking=False" in repr(pool)


def test_philox_call_counter():
    st = _PhiloxState()
    torch.manual_seed(123)
    assert [st.next() for _ in range(3)] == [(123, 0), (123, 1), (123, 2)]
    torch.manual_seed(5)                                                  # a new seed restarts the offsets
    assert st.next() == (5, 0)
    aecf_b200.set_rng_state(99, 7)
    try:
        assert aecf_b200.get_rng_state() == (99, 7)
    finally:
        aecf_b200.set_rng_state(None)


def test_folded_entry_points_validate_without_a_gpu():
    """Descriptor / argument checks of the folded-key-projection entry points run before any CUDA call."""
    lib = _lib.load()
    assert lib.aecf_fold_score_cols(_lib.BF16, 8) == 8 and lib.aecf_fold_score_cols(_lib.BF16, 12) == 16
    assert lib.aecf_fold_score_cols(_lib.F32, 1) == 4 and lib.aecf_fold_score_cols(_lib.F32, 8) == 8
    assert ops.fold_score_cols(torch.bfloat16, 8) == (8, 8) and ops.fold_score_cols(torch.float32, 6) == (8, 8)
    d = _lib.PoolDesc(device=0, dtype=_lib.BF16, batch=4, num_tokens=3, embed_dim=64, num_heads=8, q_is_shared=0)
    # the fold needs ONE query for all rows
    assert lib.aecf_pool_fwd_folded(ctypes.byref(d), 16, 16, None, 16, 16, None, None, None, None, None) == _lib.ERR_UNSUPPORTED
    d.q_is_shared = 1
    assert lib.aecf_pool_fwd_folded(ctypes.byref(d), None, 16, None, 16, 16, None, None, None, None, None) == _lib.ERR_INVALID
    assert lib.aecf_pool_fwd_folded(ctypes.byref(d), 16, 8, None, 16, 16, None, None, None, None, None) == _lib.ERR_ALIGNMENT
    d.kv_stride_b, d.kv_stride_m = 1, -2                                 # row strides must be positive
    assert lib.aecf_pool_fwd_folded(ctypes.byref(d), 16, 16, None, 16, 16, None, None, None, None, None) == _lib.ERR_INVALID
    d.kv_stride_b, d.kv_stride_m = 0, 0
    assert lib.aecf_pool_bwd_folded(ctypes.byref(d), 16, 16, 16, None, 16, None, None, 16, None, 16, 0, None) == _lib.ERR_WORKSPACE
    g = _lib.GemmDesc(device=0, dtype_a=_lib.BF16, dtype_b=_lib.BF16, dtype_c=_lib.BF16, m=256, n=128, k=64,
                      lda=64, ldb=64, ldc=128)
    for cols, ld in ((0, 8), (33, 36), (8, 4)):                          # no side columns / too many / aux_ld too small
        assert lib.aecf_gemm_aux(ctypes.byref(g), 16, 16, None, 16, 16, cols, ld, None, 0, None) == _lib.ERR_INVALID
    assert lib.aecf_fold_prepare(0, 7, 64, 8, 16, 16, 16, None) == _lib.ERR_INVALID       # unknown dtype
    assert lib.aecf_fold_prepare(0, _lib.BF16, 64, 7, 16, 16, 16, None) == _lib.ERR_INVALID   # 7 does not divide 64
    assert lib.aecf_fold_finish(0, _lib.F32, 64, 8, None, 16, 16, None, None, None) == _lib.ERR_INVALID
    # whole-step entry point: a folded descriptor with per-row queries is refused before anything is enqueued
    d.fold_key, d.q_is_shared = 1, 0
    t = _lib.FusionTensors()
    assert lib.aecf_fusion_fwd(ctypes.byref(d), ctypes.byref(t), 16, 1 << 30, None) == _lib.ERR_UNSUPPORTED


def test_graph_helper_needs_a_gpu_and_says_so():
    from aecf_b200 import graphs
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="no CPU path"):
        graphs.GraphedStep(lambda: None)


def test_pool_exposes_the_fold_switch():
    pool = aecf.MultimodalAttentionPool(16, num_heads=2)
    assert pool.fold_key_projection is None                              # automatic: on for bf16 with a shared query
    pool.fold_key_projection = False
    assert "fold" not in repr(pool)                                      # extra_repr stays the reference's
