"""The cross-rank gradient sum INSIDE the folded backward (csrc/grad_tail.cu + peer_allreduce.cu, aecf_dp_desc) with W
ranks EMULATED on one device: W pools with the same parameters, each on its own stream (and its own side stream), each
on its shard of the batch, raw-sum / reduced-sum buffers and flag blocks as plain local tensors.  The kernels, the flag
protocol, the fork/join onto the side stream and the module plumbing are those of W processes on W GPUs; only the CUDA
IPC mapping is missing (scripts/dp_check.py runs that part under torchrun on real ranks)."""
import copy

import pytest
import torch

import aecf_b200
from aecf_b200.dp import FusedGradSum, shard_rows

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
NAMES = ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias")


def _grads(pool, q):
    att = pool.attention
    return {"in_proj_weight": att.in_proj_weight.grad, "in_proj_bias": att.in_proj_bias.grad,
            "out_proj.weight": att.out_proj.weight.grad, "out_proj.bias": att.out_proj.bias.grad, "query": q.grad}


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("average", [False, True], ids=["sum", "mean"])
def test_emulated_ranks_sum_gradients_inside_the_backward(world, average):
    torch.manual_seed(7)
    B, M, D, H = 6144, 3, 512, 8                       # every product on the tcgen05 kernels, split-K weight gradients
    q0, pool0 = aecf_b200.create_fusion_pool(D, M, 0.3, num_heads=H, dropout=0.1, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        pool0.attention.in_proj_bias.normal_(0, 0.1)
        q0.mul_(6.0)
    x = (torch.randn(B, M, D, device=DEV) * 2).bfloat16()
    g = torch.randn(B, 1, D, device=DEV).bfloat16()

    def step(pool, q, rows, row0, stream=None):
        pool.row_offset = row0
        xs = x[row0:row0 + rows].clone().requires_grad_(True)
        aecf_b200.set_rng_state(77, 5)
        out, info = pool(q.expand(rows, -1, -1), xs, return_info=True)
        loss = pool.curriculum_masking.entropy_loss(info["entropy"])
        (out.float() * g[row0:row0 + rows].float()).sum().backward()
        return xs.grad, loss

    # one rank, the whole batch
    gx_full, _ = step(pool0, q0, B, 0)
    torch.cuda.synchronize()
    want = {k: v.float().clone() for k, v in _grads(pool0, q0).items()}
    pool0.zero_grad(); q0.grad = None

    # W emulated ranks
    buffers, flags = [None] * world, [None] * world
    ranks = []
    for r in range(world):
        pool = copy.deepcopy(pool0)
        q = torch.nn.Parameter(q0.detach().clone())
        pool._dp = FusedGradSum(pool, average=average, local_ranks=(r, world, buffers, flags))
        ranks.append((pool, q, torch.cuda.Stream(device=DEV)))
    for call in range(2):                              # epochs advance on the device; the second call re-uses every buffer
        torch.cuda.synchronize()
        shard_gx = []
        for r, (pool, q, stream) in enumerate(ranks):
            pool.zero_grad(); q.grad = None
            row0, rows = shard_rows(B, r, world)
            with torch.cuda.stream(stream):
                shard_gx.append(step(pool, q, rows, row0)[0])
        torch.cuda.synchronize()
        scale = 1.0 / world if average else 1.0
        got0 = _grads(*ranks[0][:2])
        for r in range(1, world):                      # every rank ends with the same bits
            for k, v in _grads(*ranks[r][:2]).items():
                assert torch.equal(v, got0[k]), f"call {call}: rank {r} differs from rank 0 in {k}"
        for k, v in got0.items():                      # ... equal to the one-rank gradient of the whole batch
            ref = want[k] * scale
            tol = 2e-2 * float(ref.abs().max())
            assert float((v.float() - ref).abs().max()) <= tol, f"{k}: {float((v.float() - ref).abs().max())} > {tol}"
        assert torch.equal(torch.cat(shard_gx), gx_full), "input gradients of the shards are not the full batch's"
        assert all(r[0]._dp.ran for r in ranks)
    aecf_b200.set_rng_state(None)


def test_fused_sum_rounds_once_like_a_single_rank():
    """fp32 sums across the ranks, ONE rounding to bf16: the W-rank gradient is the bf16 rounding of (nearly) the same
    fp32 number the one-rank run rounds, so the two differ by at most one bf16 ulp almost everywhere -- a bucket of
    bf16-rounded per-rank gradients summed in bf16 (round 1) is off by several."""
    torch.manual_seed(11)
    B, M, D, H, world = 4096, 3, 512, 8, 4
    q0, pool0 = aecf_b200.create_fusion_pool(D, M, 0.15, num_heads=H, device=DEV, dtype=torch.bfloat16)
    x = torch.randn(B, M, D, device=DEV).bfloat16()
    g = torch.randn(B, 1, D, device=DEV).bfloat16()

    def step(pool, q, rows, row0):
        pool.row_offset = row0
        aecf_b200.set_rng_state(3, 1)
        out = pool(q.expand(rows, -1, -1), x[row0:row0 + rows])
        (out.float() * g[row0:row0 + rows].float()).sum().backward()

    step(pool0, q0, B, 0)
    torch.cuda.synchronize()
    want = pool0.attention.out_proj.weight.grad.float().clone()
    buffers, flags = [None] * world, [None] * world
    ranks = []
    for r in range(world):
        pool = copy.deepcopy(pool0)
        pool.zero_grad()
        q = torch.nn.Parameter(q0.detach().clone())
        pool._dp = FusedGradSum(pool, average=False, local_ranks=(r, world, buffers, flags))
        ranks.append((pool, q, torch.cuda.Stream(device=DEV)))
    torch.cuda.synchronize()
    for r, (pool, q, stream) in enumerate(ranks):
        row0, rows = shard_rows(B, r, world)
        with torch.cuda.stream(stream):
            step(pool, q, rows, row0)
    torch.cuda.synchronize()
    aecf_b200.set_rng_state(None)
    got = ranks[0][0].attention.out_proj.weight.grad.float()
    ulp = want.abs() * 2.0 ** -7                        # one bf16 unit in the last place is at most |x| * 2^-7
    off = (got - want).abs() > ulp + 1e-30
    assert float(off.float().mean()) < 1e-3, f"{int(off.sum())} of {off.numel()} entries differ by more than one bf16 ulp"
