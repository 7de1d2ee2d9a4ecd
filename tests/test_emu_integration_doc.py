"""INTEGRATION.md's Option-B stub -- the ctypes binding a maintainer of the reference would add -- is executed as
written against the host emulation of the library (tests/cuda_emu) and must reproduce, bit for bit, what the module
itself computes for the same parameters and Philox pair (same kernels, same order)."""
import os
import re
import types

import torch

import aecf_b200
from tests.emu_support import cuda_emulation  # noqa: F401  (fixture)
from tests.golden.cases import CASES_BY_NAME, PHILOX_SEED, build_inputs, masking_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_option_b_stub_runs_and_matches_the_module(cuda_emulation, monkeypatch):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"## Option B.*?```python\n(.*?)```", text, re.S).group(1)
    assert 'C.CDLL("libaecf_b200.so")' in block
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=None))
    import ctypes
    from tests.emu_support import EMU_LIB
    scope = {"EMULATED": ctypes.CDLL(EMU_LIB)}           # a fresh handle: the stub declares its own struct types
    exec(block.replace('C.CDLL("libaecf_b200.so")', "EMULATED"), scope)

    case = CASES_BY_NAME["d64_h8_m3"]
    inp = build_inputs(case)
    bf = torch.bfloat16
    pool = aecf_b200.MultimodalAttentionPool(case.D, num_heads=case.H, curriculum_masking=aecf_b200.CurriculumMasking(**masking_kwargs(case)),
                                             dtype=bf)
    with torch.no_grad():
        pool.attention.in_proj_weight.copy_(inp["in_proj_weight"])
        pool.attention.in_proj_bias.copy_(inp["in_proj_bias"])
        pool.attention.out_proj.weight.copy_(inp["out_proj.weight"])
        pool.attention.out_proj.bias.copy_(inp["out_proj.bias"])
    pool.fold_key_projection = False                      # the first stub binds the unfolded sequence
    query0, x = inp["query0"].to(bf), inp["x"].to(bf)

    out_stub, info_stub = scope["fused_forward"](pool, query0, x, PHILOX_SEED, case.offset)
    aecf_b200.set_rng_state(PHILOX_SEED, case.offset)
    try:
        out, info = pool(query0.expand(case.B, -1, -1), x, return_info=True)
    finally:
        aecf_b200.set_rng_state(None)
    assert torch.equal(out_stub, out.detach())
    for key in ("entropy", "mask_rate", "target_entropy", "attention_weights", "masked_attention_weights"):
        assert torch.equal(info_stub[key], info[key].detach()), key
    assert pool.curriculum_masking._last_seq_len == case.M
