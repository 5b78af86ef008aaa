#!/usr/bin/env python
"""Data-sharded likelihood check, run under torchrun with N >= 2 ranks:
every rank holds all chains and 1/N of the rows; per step the per-chain partial sums are
all-reduced with NCCL inside the engine.  Rank 0 also runs the same chains on the full table on
its own GPU; the sharded trajectories must equal the single-GPU ones (counters exactly, values to
1e-9: only the summation order differs) and be bit-identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402


def check(rank, world, local, n_rows=None):
    """the check proper, inside an initialised process group; returns a summary string on rank 0
    (None elsewhere), raises AssertionError on a mismatch (every rank raises together)"""
    n_rows = n_rows or int(os.environ.get("SHARD_ROWS", "400000"))
    data = bench.light_curve(n_rows)
    n_ens, n_beta = 4, 8
    st = bench.chain_state(n_ens, n_beta, 5)
    lo = rank * n_rows // world
    hi = (rank + 1) * n_rows // world

    e = capi.Engine("simplesin5", n_ens, n_beta, seed=3, device=local)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    e.nccl_init(uid[0], rank, world)
    e.set_data(data[lo:hi])
    e.set_bounds(bench.LO, bench.HI)
    e.set_chains(0, e.n_chains, **st)
    e.run(3, 7, prob_every=1, params_chains=1)
    tr, out = e.read_trace(), e.get_chains()

    # bit-identical across ranks
    t = torch.from_numpy(np.concatenate([out["params"].ravel(), out["prob"], tr["prob"].ravel()])).cuda()
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    same = bool((t == t0).all().item())
    flags = [None] * world
    dist.all_gather_object(flags, same)
    ok = all(flags)
    err, msg = None, None
    if rank == 0:
        try:
            f = capi.Engine("simplesin5", n_ens, n_beta, seed=3, device=local)
            f.set_data(data)
            f.set_bounds(bench.LO, bench.HI)
            f.set_chains(0, f.n_chains, **st)
            f.run(3, 7, prob_every=1, params_chains=1)
            tr1, out1 = f.read_trace(), f.get_chains()
            for k in ("accept", "reject", "swapcount", "n_iter"):
                assert (out[k] == out1[k]).all(), k
            np.testing.assert_allclose(tr["prob"], tr1["prob"], rtol=1e-9)
            np.testing.assert_allclose(out["params"], out1["params"], rtol=1e-9)
            assert ok, "ranks disagree"
            f.close()
            msg = (f"shard_check ok: {world} ranks x {n_rows // world} rows == 1 GPU x {n_rows} rows; "
                   f"{int(out['n_iter'].sum())} chain-steps, ranks bit-identical")
        except AssertionError as ex:
            err = "data-sharded run differs from the single-GPU run: %s" % (str(ex)[:300],)
    e.close()
    errs = [err]
    dist.broadcast_object_list(errs, src=0)
    if errs[0]:
        raise AssertionError(errs[0])
    return msg


def main():
    rank, world, local = bench.dist_env()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    msg = check(rank, world, local)
    if rank == 0:
        print(msg)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
