"""CUDA-graph capture of the whole step (aecf_b200.graphs): replays reproduce the eager path bit for bit,
draw fresh masks / dropout every replay (the Philox pair lives on the device), and rewrite -- not
accumulate -- the gradients."""
import pytest
import torch

import aecf_b200
from aecf_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,fold", [(torch.bfloat16, None), (torch.float32, False), (torch.float32, True)],
                         ids=["bf16_folded", "fp32_unfolded", "fp32_folded"])
def test_graph_replay_matches_eager_and_draws_fresh_numbers(dtype, fold):
    torch.manual_seed(3)
    B, M, D, H = 2048, 3, 256, 4
    q, pool = aecf_b200.create_fusion_pool(D, M, 0.4, num_heads=H, dropout=0.1, device=DEV, dtype=dtype)
    pool.fold_key_projection = fold
    pool._want_mask_bits = True
    cm = pool.curriculum_masking
    x = torch.randn(B, M, D, device=DEV, dtype=dtype).requires_grad_(True)
    params = [q, x, pool.attention.in_proj_weight, pool.attention.in_proj_bias, pool.attention.out_proj.weight,
              pool.attention.out_proj.bias]

    def clear():
        for p in params:
            p.grad = None

    def step():
        out, info = pool(q.expand(B, -1, -1), x, return_info=True)
        loss = out.float().pow(2).mean() + 0.01 * cm.entropy_loss(info["entropy"])
        loss.backward()
        return out, info["mask_bits"], info["attention_weights"], loss

    def snapshot(res):
        return [t.detach().clone() for t in res] + [p.grad.detach().clone() for p in params]

    aecf_b200.set_rng_state(777, 5)
    before = _lib.launch_count()
    graphed = aecf_b200.graphs.GraphedStep(step, reset=clear, warmup=2, device=torch.device(DEV))
    per_step = (_lib.launch_count() - before) // 3            # two warm-up steps + the capture
    assert per_step >= 10
    count = _lib.launch_count()
    r1 = snapshot(graphed())
    r2 = snapshot(graphed())
    r3 = snapshot(graphed())
    torch.cuda.synchronize()
    assert _lib.launch_count() == count                       # replays enqueue nothing from the host side
    # warm-up took offsets 5 and 6, prepare() took 7: replay i draws what eager call i draws from offset 7 on
    aecf_b200.set_rng_state(777, 7)
    eager = []
    for _ in range(3):
        clear()
        eager.append(snapshot(step()))
    aecf_b200.set_rng_state(None)
    for i, (r, e) in enumerate(zip((r1, r2, r3), eager)):
        for j, (a, b) in enumerate(zip(r, e)):
            assert torch.equal(a, b), f"replay {i}: tensor {j} differs from the eager step"
    assert not torch.equal(r1[1], r2[1]) and not torch.equal(r2[1], r3[1])      # fresh masks each replay
    assert not torch.equal(r1[0], r2[0])                                         # fresh dropout too


def test_capture_without_prepared_rng_state_fails_loudly(monkeypatch):
    from aecf_b200.layers import _rng
    saved = dict(_rng.device_states)
    _rng.device_states.clear()
    try:
        q, pool = aecf_b200.create_fusion_pool(64, 3, 0.2, num_heads=4, device=DEV)
        x = torch.randn(64, 3, 64, device=DEV)
        monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)   # as if inside a capture
        with pytest.raises(RuntimeError, match="aecf_b200.graphs.prepare"):
            pool(q.expand(64, -1, -1), x)
    finally:
        _rng.device_states.update(saved)
