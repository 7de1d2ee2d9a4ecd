"""The CUDA sources, run on the CPU by the host emulation (tests/cuda_emu), through the public module API and the
C ABI, against the oracle and the reference's fixtures: the bodies of the GPU parity tests, called on CPU tensors.

Covers what the emulation can see -- indexing, lane/head geometry for every (M, J, warps-per-sample) the cases reach,
reductions, the masking stage bit for bit, the Philox stream, the host-side sequencing of the whole-step entry points
(folded and unfolded, one and several queries per sample), the SIMT GEMM / GEMV kernels -- and nothing it cannot
(tcgen05 / TMA, memory-model races, performance).  Small shapes: a block's threads are fibers.
"""
import pytest
import torch

import aecf_b200
from aecf_b200 import _lib
from oracle import aecf_oracle as oracle
from oracle import philox
from tests import test_gpu_multi_query as MQ
from tests import test_gpu_parity as P
from tests.emu_support import cuda_emulation  # noqa: F401  (fixture)
from tests.golden.cases import MULTI_QUERY_CASES
from tests.helpers import assert_close

pytestmark = pytest.mark.usefixtures("cuda_emulation")


@pytest.fixture(autouse=True)
def _cpu_stands_in_for_the_device(monkeypatch):
    """The GPU tests say ``t.to(DEV)`` and get a copy; ``t.to("cpu")`` of a CPU tensor is the tensor itself, which would
    alias a case's inputs with the tensors that collect gradients.  Make it a copy here too."""
    monkeypatch.setattr(P, "DEV", "cpu")
    monkeypatch.setattr(MQ, "DEV", "cpu")
    plain_to = torch.Tensor.to

    def to_copy(self, *args, **kwargs):
        moved = plain_to(self, *args, **kwargs)
        return moved.clone() if moved is self else moved

    monkeypatch.setattr(torch.Tensor, "to", to_copy)


by_name = lambda c: c.name   # noqa: E731


# ---- one fusion query per sample (the hot path's kernels) --------------------------------------------------------
@pytest.mark.parametrize("case", P.FP32_CASES, ids=by_name)
def test_fp32_matches_oracle(case):
    P.test_fp32_matches_oracle(case)


@pytest.mark.parametrize("case", P.FP32_CASES[1:6], ids=by_name)
def test_fp32_matches_reference_golden(case):
    P.test_fp32_matches_reference_golden(case)


@pytest.mark.parametrize("case", P.FOLDABLE_FP32_CASES, ids=by_name)
def test_fp32_folded_key_projection_matches_oracle(case):
    P.test_fp32_folded_key_projection_matches_oracle(case)


@pytest.mark.parametrize("fold", [True, False], ids=["folded", "unfolded"])
@pytest.mark.parametrize("case", P.BF16_CASES, ids=by_name)
def test_bf16_masks_exact_against_stage_rounded_oracle(case, fold):
    P.test_bf16_masks_exact_against_stage_rounded_oracle(case, fold)


@pytest.mark.parametrize("fold", [True, False], ids=["folded", "unfolded"])
@pytest.mark.parametrize("case", P.BF16_CASES[1:4], ids=by_name)
def test_bf16_matches_fp32_math_oracle(case, fold):
    P.test_bf16_matches_fp32_math_oracle(case, fold)


WIDE = {c.name: c for c in P.WIDE_CASES}


@pytest.mark.parametrize("name,dtype", [("wide_d1024_h8_m3", torch.bfloat16),
                                        ("wide_d1024_h4_m4_eval", torch.bfloat16), ("wide_d512_h8_m5_kpm", torch.float32),
                                        ("wide_d512_h8_m5_kpm", torch.bfloat16)])
def test_rows_spanning_several_warps(name, dtype):
    P.test_wide_rows_span_several_warps(WIDE[name], dtype)


@pytest.mark.parametrize("name", ["wide_d1024_h4_m4_eval", "wide_d512_h8_m5_kpm"])
def test_rows_spanning_several_warps_folded_fp32(name):
    P.test_wide_rows_folded_fp32(WIDE[name])


def test_rows_of_three_slices():
    """D = 768 in fp32 is 192 sixteen-byte slices: J = 4, two warps per sample, the second half idle."""
    case = P.THREE_SLICE_CASES[0]
    P.test_fp32_matches_oracle(case)
    P.test_bf16_masks_exact_against_stage_rounded_oracle(case, True)


@pytest.mark.parametrize("test", [P.test_sequence_first_layout_matches_batch_first, P.test_per_row_queries_match_oracle,
                                  P.test_attn_mask_forms_match_oracle, P.test_no_masking_module_and_plain_output,
                                  P.test_standalone_masking_and_entropy, P.test_entropy_loss_gradient_in_eval_mode,
                                  P.test_colsum], ids=lambda f: f.__name__[5:])
def test_module_surface(test):
    test()


@pytest.mark.parametrize("dtype,fold", [(torch.float32, False), (torch.float32, True), (torch.bfloat16, True), (torch.bfloat16, False)],
                         ids=["fp32_unfolded", "fp32_folded", "bf16_folded", "bf16_unfolded"])
def test_module_without_biases(dtype, fold):
    P.test_module_without_biases(dtype, fold)


def test_functional_fast_path():
    q = torch.from_numpy(philox.normal(1, (6, 2, 64))).float()
    k = torch.from_numpy(philox.normal(2, (6, 5, 64))).float()
    assert_close("sdpa", aecf_b200.multimodal_attention_pool(q, k), oracle.sdpa_single_head(q, k, k), 1e-5)
    got = aecf_b200.multimodal_attention_pool(q.bfloat16(), k.bfloat16()).float()
    assert_close("sdpa bf16", got, oracle.sdpa_single_head(q.bfloat16().float(), k.bfloat16().float(), k.bfloat16().float()), 2e-2)


def test_entropy_loss_comes_out_of_the_forward_kernel():
    P.test_entropy_loss_comes_out_of_the_forward_kernel()


@pytest.mark.parametrize("shape", [(6, 2, 5, 64), (3, 1, 3, 512), (4, 7, 9, 136)], ids=lambda s: "x".join(map(str, s)))
def test_functional_fast_path_is_differentiable(shape):
    P.test_functional_fast_path_is_differentiable(shape)


@pytest.mark.parametrize("dtype,fold", [(torch.float32, False), (torch.float32, True), (torch.bfloat16, True), (torch.bfloat16, False)],
                         ids=["fp32_unfolded", "fp32_folded", "bf16_folded", "bf16_unfolded"])
def test_sample_index_pools_the_listed_rows_in_place(dtype, fold):
    P.test_sample_index_pools_the_listed_rows_in_place(dtype, fold)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_in_projection_bias_gradient_both_ways(dtype):
    P.test_in_projection_bias_gradient_both_ways(dtype)


def test_full_size_test_body_on_a_small_batch():
    """The body of the B = 65 536 GPU test (tests/test_gpu_parity.py) on 192 rows, so that the test itself is tested."""
    P.test_full_size_bf16_folded_against_oracle(P.Case("small_d512_h8_m3", B=192, M=3, D=512, H=8, dropout=0.1, pooled_grad=True,
                                                       data_seed=43, offset=2, row0=64))


# ---- several fusion queries per sample (csrc/pool_multi.cuh) --------------------------------------------------------
@pytest.fixture
def multi_query():
    return True


@pytest.mark.parametrize("batch_first", [True, False], ids=["batch_first", "seq_first"])
@pytest.mark.parametrize("case", MULTI_QUERY_CASES, ids=by_name)
def test_multi_query_fp32_matches_oracle(multi_query, case, batch_first):
    MQ.test_fp32_matches_oracle(case, batch_first)


@pytest.mark.parametrize("case", MULTI_QUERY_CASES, ids=by_name)
def test_multi_query_fp32_matches_reference_golden(multi_query, case):
    MQ.test_fp32_matches_reference_golden(case)


@pytest.mark.parametrize("case", MULTI_QUERY_CASES, ids=by_name)
def test_multi_query_bf16_masks_exact_against_stage_rounded_oracle(multi_query, case):
    MQ.test_bf16_masks_exact_against_stage_rounded_oracle(case)


def test_multi_query_shards_and_attn_mask(multi_query):
    MQ.test_batch_shards_reproduce_the_full_batch()
    MQ.test_attn_mask_per_query()


# ---- the SIMT GEMM / GEMV kernels and the side output's two-launch form -------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(200, 136, 72), (1, 512, 512), (512, 512, 1)], ids=lambda s: "x".join(map(str, s)))
def test_simt_gemm_layouts(dtype, shape):
    P.test_gemm_layouts(_lib.GEMM_AUTO, dtype, shape)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_gemm_with_side_output(dtype):
    P.test_gemm_with_side_output((96, 64, 64, 1), dtype)


# ---- seeded random shapes and options (a slice of the offline fuzz runs that found the 3 * 2^k slice gap) -------------
def _random_case(rng, name, multi):
    from tests.golden.cases import Case
    while True:
        hd, heads = rng.choice([8, 16, 32, 64, 128]), rng.choice([1, 2, 3, 4, 6, 8, 12])
        if hd * heads <= 384:
            break
    tokens = rng.randint(2 if multi else 1, 8)
    return Case(name, B=rng.choice([1, 2, 5, 9, 17]), S=rng.randint(2, 4) if multi else 1, M=tokens, D=hd * heads, H=heads,
                dropout=rng.choice([0.0, 0.1, 0.5]), base_mask_prob=rng.choice([0.15, 0.5, 1.0]), min_active=rng.choice([1, 2, 9]),
                training=rng.random() < 0.8, kpm=rng.random() < 0.3 and tokens > 1, pooled_grad=rng.random() < 0.5,
                offset=rng.randint(0, 1000), row0=rng.choice([0, 7, 123456789012]), data_seed=rng.randint(500, 10 ** 6),
                peak=rng.choice([0.5, 1.0, 2.0]))


@pytest.mark.parametrize("seed", range(6))
def test_random_single_query_cases(seed):
    import random
    case = _random_case(random.Random(1000 + seed), f"random{seed}", multi=False)
    inp = P.build_inputs(case)
    ref, ref_grads = P.run_oracle(case, inp)
    for fold in (False, True):
        out, info, _, grads, _ = P.run_cuda(case, inp, torch.float32, fold=fold)
        assert (info["mask_bits"].numpy() == P.expected_bits(ref.info["mask"])).all(), (case, fold)
        assert_close("out", out, ref.out, 2e-5)
        P.check_grads(case, grads, ref_grads, 3e-5)          # peaked rows cancel in the softmax backward: a little slack
    if case.D // case.H % 8 == 0:
        P.test_bf16_masks_exact_against_stage_rounded_oracle(case, True)


@pytest.mark.parametrize("seed", range(6))
def test_random_multi_query_cases(multi_query, seed):
    import random
    from tests.helpers import run_oracle
    rng = random.Random(2000 + seed)
    case = _random_case(rng, f"random_mq{seed}", multi=True)
    inp = MQ.build_inputs(case)
    ref, ref_grads = run_oracle(case, inp)
    out, info, ent_loss, grads, _ = MQ.run_cuda(case, inp, torch.float32, batch_first=rng.random() < 0.5)
    MQ.check_against_oracle(case, out, info, ent_loss, grads, ref, ref_grads, 3e-5)
    if case.D // case.H % 8 == 0:
        MQ.test_bf16_masks_exact_against_stage_rounded_oracle(case)


# ---- the device-side Philox pair a CUDA-graph capture reads (aecf_pool_desc::rng_state) ------------------------------
@pytest.mark.parametrize("fold", [False, True], ids=["unfolded", "folded"])
def test_device_side_philox_state_reproduces_the_by_value_pair(monkeypatch, fold):
    """What tests/test_gpu_graphs.py checks with real captures: a forward that reads {seed, offset} from device memory
    draws what the by-value call draws, and the offset advances on the device between calls."""
    import dataclasses

    from aecf_b200 import graphs
    from tests.golden.cases import CASES_BY_NAME, PHILOX_SEED, build_inputs
    case = CASES_BY_NAME["d64_h8_m3_dropout"]
    inp = build_inputs(case)
    eager = [P.run_cuda(dataclasses.replace(case, offset=case.offset + i), inp, torch.float32, fold=fold) for i in range(2)]
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    aecf_b200.set_rng_state(PHILOX_SEED, case.offset)
    state = graphs.prepare(torch.device("cpu"))
    aecf_b200.set_rng_state(None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)
    for i in range(2):
        got = P.run_cuda(case, inp, torch.float32, fold=fold)           # its by-value pair is ignored while "capturing"
        assert torch.equal(got[0], eager[i][0]) and torch.equal(got[1]["mask_bits"], eager[i][1]["mask_bits"]), i
        assert torch.equal(got[3]["key"], eager[i][3]["key"]), i
        assert int(state[1]) == case.offset + i + 1
    assert not torch.equal(eager[0][1]["mask_bits"], eager[1][1]["mask_bits"])
