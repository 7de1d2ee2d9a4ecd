"""The NVLink peer-memory all-reduce kernel (csrc/peer_allreduce.cu) with W ranks EMULATED on one device: W buckets
and W flag blocks in one process, rank r's kernel on its own stream, all W resident at once -- the same code path
(flag protocol, slice ownership, rank-ordered sums) as W processes on W GPUs, minus the IPC mapping."""
import pytest
import torch

from aecf_b200 import ops

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_emulated_ranks_reduce_in_place_repeatedly(world, dtype):
    torch.manual_seed(world)
    n = 1051136 if dtype == torch.bfloat16 else 40 * 1024 + 4          # the D = 512 gradient bucket / a ragged one
    flags = [ops.peer_flag_block(DEV) for _ in range(world)]
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    for call in range(3):                                               # epochs advance on the device
        src = [torch.randn(n, device=DEV).to(dtype) for _ in range(world)]
        buckets = [s.clone() for s in src]
        want = torch.zeros(n, device=DEV, dtype=torch.float32)
        for s in src:                                                   # rank order, fp32 accumulation
            want += s.float()
        average = call == 1
        if average:
            want = want * (1.0 / world)
        want = want.to(dtype)
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ops.peer_allreduce(buckets, flags, r, average=average, grid_limit=8)
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(buckets[r], want), f"call {call}: rank {r} differs"
        assert all(int(f[16]) == call + 1 and int(f[17]) == 0 for f in flags)
