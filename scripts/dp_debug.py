"""Staged check of the CUDA-IPC peer mapping under torchrun (debugging aid): each stage synchronises and prints."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aecf_b200 import ops  # noqa: E402
from aecf_b200.dp import _map_peer_tensors  # noqa: E402


def stage(rank, name):
    torch.cuda.synchronize()
    dist.barrier()
    print(f"[rank {rank}] ok: {name}", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 20
    buf = torch.full((n,), float(rank + 1), device=dev)
    flags = ops.peer_flag_block(dev)
    stage(rank, "allocated")
    bufs, fl = _map_peer_tensors([buf, flags])
    print(f"[rank {rank}] mapped: {[hex(b) for b in bufs]} flags {[hex(f) for f in fl]}", flush=True)
    stage(rank, "mapped")
    ops.peer_allreduce(bufs, fl, rank, average=False, mine=buf)
    stage(rank, "peer_allreduce kernel")
    print(f"[rank {rank}] after all-reduce: {buf[:2].tolist()} (want {[float(sum(range(1, world + 1)))] * 2})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
