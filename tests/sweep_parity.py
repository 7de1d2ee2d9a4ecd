"""BASELINE.json configs[4]: the stress sweep -- M 2..8, D 256..2048, H 4..16, up to 131 072 rows per GPU (1 M on 8 GPUs) --
with curriculum masking and min_active = 1, each point with a row-sampled ORACLE check and timings.

    python tests/sweep_parity.py                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/sweep_parity.py

(Lives under tests/ because it executes the oracle: a checker run at full size, not a product path.)

Per point, on every rank's shard of the global batch (Philox keyed on the global row; gradients summed over the ranks
inside the backward):
  * parity: a block of 128 consecutive rows at a pseudo-random position of the shard is recomputed by the CPU oracle
    (bf16 where the CUDA path stores bf16, folded like it) from the same inputs and the same global-row Philox draws --
    mask bits and mask_rate must match EXACTLY, pooled weights / entropy to 1e-4, the output rows and the rows' input
    gradients to the bf16 budget (2e-2);
  * timing: the step replayed as a CUDA graph (max over ranks), and per-kernel times of one eager region for the pool
    kernels' fraction of the measured HBM copy bandwidth.
Rank 0 prints one JSON line per point and a table at the end.
"""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aecf_b200  # noqa: E402
from aecf_b200 import _lib  # noqa: E402
from aecf_b200.dp import GradientSync  # noqa: E402
from oracle import aecf_oracle as oracle  # noqa: E402  (the checker, not the measured path)
from oracle import philox  # noqa: E402

POINTS = [  # (M, D, H, rows per GPU)
    (3, 512, 8, 131072), (2, 256, 4, 131072), (8, 256, 4, 131072), (4, 512, 16, 131072), (6, 512, 4, 65536),
    (2, 1024, 16, 131072), (5, 1024, 8, 65536), (8, 1024, 4, 32768), (2, 2048, 16, 65536), (3, 2048, 8, 32768),
    (8, 2048, 16, 16384), (3, 512, 8, 8192),
]
SEED, OFFSET = 0x5EED, 9
SAMPLE = 128


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    hbm = peaks()
    table = []
    for (M, D, H, B) in POINTS:
        rec = {"M": M, "D": D, "H": H, "rows_per_gpu": B, "global_rows": B * world, "n_gpus": world}
        try:
            torch.manual_seed(0)
            q, pool = aecf_b200.create_fusion_pool(D, M, 0.15, num_heads=H, device=dev, dtype=torch.bfloat16)
            with torch.no_grad():
                q.mul_(8.0)                                   # peaked attention: keep_prob depends on the entropy
                pool.attention.in_proj_bias.normal_(0, 0.05)
            pool._want_mask_bits = True
            sync = GradientSync(pool, q).attach()
            row0, rows = sync.set_shard(B * world)
            torch.manual_seed(100 + rank)
            x = (torch.randn(B, M, D, device=dev) * 2).bfloat16().requires_grad_(True)
            g = torch.randn(B, 1, D, device=dev).bfloat16()

            def step():
                out, info = pool(q.expand(B, -1, -1), x, return_info=True)
                loss = pool.curriculum_masking.entropy_loss(info["entropy"])
                out.backward(g)
                sync.finish()
                return out, info, loss

            def clear():
                x.grad = None; q.grad = None
                pool.zero_grad(set_to_none=True)

            # ---- parity on a sampled block of rows ----------------------------------------------------------------------
            aecf_b200.set_rng_state(SEED, OFFSET)
            out, info, loss = step()
            aecf_b200.set_rng_state(None)
            torch.cuda.synchronize()
            first = (1234567 * (rank + 1) + 89 * M + D) % max(1, B - SAMPLE)
            sl = slice(first, first + SAMPLE)
            att = pool.attention
            cpu = lambda t: t.detach().float().cpu()
            u_mask = torch.from_numpy(philox.mask_uniforms(SEED, OFFSET, row0 + first, SAMPLE, M))
            ref = oracle.pool_forward(cpu(q).expand(SAMPLE, 1, D), cpu(x[sl]), None, cpu(att.in_proj_weight), cpu(att.in_proj_bias),
                                      cpu(att.out_proj.weight), cpu(att.out_proj.bias), H, training=True, u_mask=u_mask,
                                      masking=dict(base_mask_prob=0.15, entropy_target=0.7, min_active=1),
                                      storage=torch.bfloat16, fold_key=True)
            grads = oracle.pool_backward(cpu(q).expand(SAMPLE, 1, D), cpu(x[sl]), None, cpu(att.in_proj_weight), cpu(att.out_proj.weight),
                                         H, ref.saved, cpu(g[sl]), training=True, storage=torch.bfloat16, fold_key=True)
            want_bits = ((ref.info["mask"].reshape(SAMPLE, M) > 0).numpy().astype(np.int64) * (1 << np.arange(M))).sum(1)
            got_bits = info["mask_bits"][sl].cpu().numpy().astype(np.int64)
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            rec["parity"] = {
                "rows": [int(row0 + first), int(row0 + first + SAMPLE)],
                "mask_bits_exact": bool(np.array_equal(got_bits, want_bits)),
                "mask_rate_exact": bool(torch.equal(cpu(info["mask_rate"][sl]).reshape(-1), ref.info["mask_rate"].float().reshape(-1))),
                "pooled_max_abs": float((cpu(info["attention_weights"][sl]).reshape(SAMPLE, M) - ref.info["attention_weights"].reshape(SAMPLE, M)).abs().max()),
                "entropy_max_abs": float((cpu(info["entropy"][sl]).reshape(-1) - ref.info["entropy"].reshape(-1)).abs().max()),
                "out_rel": rel(cpu(out[sl]).reshape(SAMPLE, D), ref.out.reshape(SAMPLE, D)),
                "d_x_rel": rel(cpu(x.grad[sl]), grads["key"]),
                "mean_mask_rate": float(info["mask_rate"].mean()),
            }
            p = rec["parity"]
            rec["parity"]["ok"] = bool(p["mask_bits_exact"] and p["mask_rate_exact"] and p["pooled_max_abs"] < 1e-4
                                       and p["entropy_max_abs"] < 1e-4 and p["out_rel"] < 2e-2 and p["d_x_rel"] < 2e-2)
            # nothing of the checked step may stay alive: its autograd graph holds AccumulateGrad nodes bound to the default
            # stream, and a capture on another stream would synchronise with them (and be invalidated)
            del out, info, loss, ref, grads
            clear()
            # ---- timing ---------------------------------------------------------------------------------------------------
            for _ in range(3):
                step(); clear()
            graph = aecf_b200.graphs.GraphedStep(step, reset=clear, warmup=0, device=dev)
            for _ in range(3):
                graph()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.all_reduce(torch.zeros(1, device=dev))
            a.record()
            for _ in range(10):
                graph()
            b.record()
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / 10], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            del graph
            clear()
            _lib.timing_enable(True)
            for _ in range(3):
                step(); clear()
            sites = _lib.timing_collect()
            _lib.timing_enable(False)
            hs, hsp = (H + 3) // 4 * 4, (H + 7) // 8 * 8
            fwd_bytes = B * (2 * (M * D + D) + 4 * M * hs + 4 * (2 * M + 2))
            bwd_bytes = B * (2 * (M * D + D + M * (D + hsp)) + 4 * M * hs)
            fwd_ms = sites["pool_fwd"][0] / sites["pool_fwd"][1]
            bwd_ms = sites["pool_bwd"][0] / sites["pool_bwd"][1]
            rec.update({"ms_per_step": float(ms), "samples_per_s": B * world / (float(ms) * 1e-3),
                        "pool_fwd_us": fwd_ms * 1e3, "pool_fwd_frac_of_hbm": fwd_bytes / (fwd_ms * 1e-3) / 1e9 / hbm,
                        "pool_bwd_us": bwd_ms * 1e3, "pool_bwd_frac_of_hbm": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / hbm,
                        "gradient_sum": "inside the backward" if sync.fused is not None else ("bucket" if world > 1 else "none")})
            del sync, pool, q, x, g
            torch.cuda.empty_cache()
        except Exception as e:                                  # a point that fails must not take the sweep down
            rec["error"] = f"{type(e).__name__}: {e}"[:400]
            aecf_b200.set_rng_state(None)
        if world > 1:                                           # parity must hold on EVERY rank
            flag = torch.tensor([0 if rec.get("parity", {}).get("ok") else 1], device=dev)
            dist.all_reduce(flag)
            rec["ranks_failing_parity"] = int(flag.item())
        if rank == 0:
            print(json.dumps(rec), flush=True)
            table.append(rec)
    if rank == 0:
        print(f"{'M':>2} {'D':>5} {'H':>3} {'rows/GPU':>9} {'parity':>7} {'ms/step':>8} {'M samples/s':>12} {'fwd frac':>9} {'bwd frac':>9}")
        for r in table:
            if "error" in r:
                print(f"{r['M']:>2} {r['D']:>5} {r['H']:>3} {r['rows_per_gpu']:>9} ERROR {r['error'][:80]}")
            else:
                ok = r["parity"]["ok"] and r.get("ranks_failing_parity", 0) == 0
                print(f"{r['M']:>2} {r['D']:>5} {r['H']:>3} {r['rows_per_gpu']:>9} {'ok' if ok else 'FAIL':>7} {r['ms_per_step']:>8.3f} "
                      f"{r['samples_per_s'] / 1e6:>12.1f} {r['pool_fwd_frac_of_hbm']:>9.2f} {r['pool_bwd_frac_of_hbm']:>9.2f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
