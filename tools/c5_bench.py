#!/usr/bin/env python
"""Config C5 of BASELINE.json: simplesin5 on a light curve too long for one GPU's caches, the rows
sharded over the ranks, the per-chain partial log-likelihoods summed with one ncclAllReduce per
Metropolis step inside the engine.

Run alone (one GPU, no collective: the per-GPU shard of C5) or under torchrun:

    python tools/c5_bench.py                              # 12.5 M rows on one GPU
    torchrun --nproc-per-node 8 tools/c5_bench.py         # 8 x 12.5 M = the 100 M-row curve of C5

Every rank holds all 4096 chains (64 ensembles x 64 rungs) and C5_ROWS_PER_GPU rows (default
12 500 000 = 200 MB); the curve grows with the number of ranks, so the per-GPU work is fixed.
Prints one JSON line (rank 0): chain-steps/s of the whole job, row-evaluations/s, the likelihood
kernel's share and FP64 roofline fraction, and what is left for the all-reduce + control kernel.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402


def shard_rows(lo, hi, n_total, seed):
    """rows [lo, hi) of the SURVEY 8d light curve of n_total rows (x_i = i * 1000 / n_total)"""
    rng = np.random.default_rng(seed)
    x = np.arange(lo, hi, dtype=np.float64) * (1000.0 / n_total)
    t = bench.TRUTH
    y = t[0] * np.sin(2 * np.pi * t[1] * x + t[2]) + t[3] + rng.normal(0, 0.5, hi - lo)
    return np.ascontiguousarray(np.stack([x, y], axis=1))


def measure(rank, world, local, per_gpu, n_steps=6, n_rounds=2, peak=None):
    """times n_rounds x (n_steps Metropolis steps + swap) of the 4096 replicated chains on a curve of
    per_gpu x world rows (inside an initialised process group when world > 1); the record on rank 0,
    None elsewhere"""
    n_total = per_gpu * world
    n_ens, n_beta = bench.N_ENS, bench.N_BETA
    eng = capi.Engine("simplesin5", n_ens, n_beta, seed=1, device=local)
    if world > 1:
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.nccl_init(uid[0], rank, world)
    eng.set_data(shard_rows(rank * per_gpu, (rank + 1) * per_gpu, n_total, 777 + rank))
    eng.set_bounds(bench.LO, bench.HI)
    st = bench.chain_state(n_ens, n_beta, 1000)   # identical on every rank: the chains are replicated
    # the posterior narrows with the length of the curve: keep the proposals at its scale
    scale = (1e6 / n_total) ** 0.5
    st["steps"] = st["steps"] * scale
    st["params"] = np.clip(bench.TRUTH[None, :] + (st["params"] - bench.TRUTH[None, :]) * scale, bench.LO, bench.HI)
    st["params_best"] = st["params"].copy()
    eng.set_chains(0, eng.n_chains, **st)
    eng.set_timing(True)   # per-launch events: the likelihood kernel's own time
    if peak is None:   # FP64 issue peak: lanes per SM per clock (measured) x SMs x the nominal SM clock
        per_clock, n_sm = capi.measure_fp64_per_clock(local)
        peak = per_clock * n_sm * torch.cuda.get_device_properties(local).clock_rate * 1e3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.run(1, 3)   # warm-up: 3 steps + swap
    barrier()
    t0 = time.perf_counter()
    eng.run(n_rounds, n_steps)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ll_ms, ll_launches, total_ms = eng.last_kernel_ms()
    t = torch.tensor([total_ms, wall_ms, ll_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, wall_ms, ll_ms_max = [float(v) for v in t.tolist()]
    out = eng.get_chains()
    # the replicas must have stayed in lock-step: same state bits on every rank
    same = True
    if world > 1:
        chk = torch.from_numpy(np.concatenate([out["params"].ravel(), out["prob"]])).cuda()
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        flags = [None] * world
        dist.all_gather_object(flags, bool((chk == ref).all().item()))
        same = all(flags)
    acc = float(out["accept"].sum()) / float((out["accept"] + out["reject"]).sum())
    eng.close()
    if rank != 0:
        return None
    steps = n_rounds * n_steps
    per_launch = ll_ms / max(ll_launches, 1)
    achieved = eng.n_chains * per_gpu * bench.ALG_FP64_PER_ROW / (per_launch * 1e-3)
    value = eng.n_chains * steps / (total_ms * 1e-3)
    return {
        "workload": f"C5: simplesin5, {n_total}-row synthetic light curve sharded over {world} GPU(s) "
                    f"({per_gpu} rows = {per_gpu * 16 / 1e6:.0f} MB each), 4096 replicated chains, "
                    "per-step ncclAllReduce of the per-chain partial sums on the engine's own communicator"
                    + ("" if world > 1 else " (none at 1 GPU)"),
        "n_gpus": world, "n_rows_total": n_total, "rows_per_gpu": per_gpu, "metropolis_steps_timed": steps,
        "chain_steps_per_s": value, "row_evals_per_s": value * n_total,
        "ms_per_metropolis_step": total_ms / steps, "loglik_kernel_ms_per_launch": per_launch,
        "loglik_share_of_step": ll_ms / total_ms,
        "allreduce_plus_control_us_per_step": (total_ms - ll_ms_max) / steps * 1e3,
        "allreduce_bytes_per_step": eng.n_chains * 8,
        "replicas_bit_identical": same,
        "roofline": {"bound": "fp64", "achieved": achieved / 1e9, "peak": peak / 1e9,
                     "unit": "GFP64-instr/s", "frac": achieved / peak,
                     "hbm_min_gbs": per_gpu * 16 / (per_launch * 1e-3) / 1e9,
                     "note": "hbm_min_gbs = the shard read once per launch; the kernel is FP64-bound as long as "
                             "DRAM traffic stays near that (see the ncu capture of this command)"},
        "acceptance_rate": acc, "wall_ms": wall_ms,
    }


def main():
    rank, world, local = bench.dist_env()
    per_gpu = int(os.environ.get("C5_ROWS_PER_GPU", "12500000"))
    n_steps = int(os.environ.get("C5_STEPS", "6"))      # Metropolis steps per timed round
    n_rounds = int(os.environ.get("C5_ROUNDS", "2"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rec = measure(rank, world, local, per_gpu, n_steps, n_rounds)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
