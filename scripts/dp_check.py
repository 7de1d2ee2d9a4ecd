"""Real ranks (torchrun, one process per GPU): the cross-rank gradient sum inside the backward over CUDA-IPC peer mappings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py

Checks, for collective = fused / nccl / peer: every rank ends with the same parameter-gradient bits; they equal the
one-rank gradient of the whole batch (bf16 budget); shard input gradients are the full batch's; repeated steps work; a
CUDA-graph replay of the step matches.  Prints one JSON line per collective from rank 0 and exits non-zero on a mismatch."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aecf_b200  # noqa: E402
from aecf_b200.dp import GradientSync  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, M, D, H = 8192 * world, 3, 512, 8
    ok = True
    for collective in ("fused", "nccl", "peer"):
        torch.manual_seed(5)
        q, pool = aecf_b200.create_fusion_pool(D, M, 0.3, num_heads=H, dropout=0.1, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            pool.attention.in_proj_bias.normal_(0, 0.1)
            q.mul_(6.0)
        x = (torch.randn(B, M, D, device=dev) * 2).bfloat16()            # same seed on every rank: the same global batch
        g = torch.randn(B, 1, D, device=dev).bfloat16()
        params = {"in_proj_weight": pool.attention.in_proj_weight, "in_proj_bias": pool.attention.in_proj_bias,
                  "out_proj.weight": pool.attention.out_proj.weight, "out_proj.bias": pool.attention.out_proj.bias, "query": q}

        def step(rows, row0):
            pool.row_offset = row0
            xs = x[row0:row0 + rows].clone().requires_grad_(True)
            aecf_b200.set_rng_state(77, 5)
            out = pool(q.expand(rows, -1, -1), xs)
            (out.float() * g[row0:row0 + rows].float()).sum().backward()
            return xs.grad

        gx_full = step(B, 0)                                             # one rank, the whole batch (no sync attached yet)
        want = {k: v.grad.float().clone() for k, v in params.items()}
        for p in params.values():
            p.grad = None
        sync = GradientSync(pool, q, average=False, collective=collective).attach()
        row0, rows = sync.set_shard(B)
        errs = {}
        for it in range(3):
            for p in params.values():
                p.grad = None
            gx = step(rows, row0)
            sync.finish()
            torch.cuda.synchronize()
            for k, p in params.items():
                got = p.grad.float()
                errs[k] = max(errs.get(k, 0.0), float((got - want[k]).abs().max() / want[k].abs().max()))
                ref = got.clone()
                dist.broadcast(ref, 0)
                if not torch.equal(ref, got):
                    ok = False
                    print(f"[rank {rank}] {collective}: {k} differs from rank 0", file=sys.stderr)
            if not torch.equal(gx, gx_full[row0:row0 + rows]):
                ok = False
                print(f"[rank {rank}] {collective}: shard input gradient differs from the full batch's", file=sys.stderr)
        # timing of the step, eager (CUDA events), this rank
        for p in params.values():
            p.grad = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        a.record()
        for it in range(20):
            step(rows, row0); sync.finish()
            for p in params.values():
                p.grad = None
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        if max(errs.values()) > 2e-2:
            ok = False
        aecf_b200.set_rng_state(None)
        if rank == 0:
            print(json.dumps({"collective": collective, "world": world, "fused_ran": sync.fused is not None,
                              "max_rel_err_vs_one_rank": errs, "eager_ms_per_step": ms, "ok": ok}))
        del sync
        pool._dp = None; pool._grad_ready = None; pool._grad_buffers = None
        dist.barrier()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
