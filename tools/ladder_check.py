#!/usr/bin/env python
"""Ladder split over GPUs (third row), run under torchrun with N >= 2 ranks: every
rank holds a contiguous block of rungs of every ensemble and the whole table; once per round the
boundary chains are traded with ncclSend/ncclRecv so that a swap pair that straddles two GPUs is
decided identically on both.  Rank 0 also runs the whole ladder on its own GPU: the split run must
equal it BIT FOR BIT (a chain's likelihood sum does not depend on which chains share its tile)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402


def check(rank, world, local, quirks=None, n_rows=None):
    """the check proper, inside an initialised process group; returns a summary string on rank 0
    (None elsewhere), raises AssertionError on a mismatch (every rank raises together)"""
    n_rows = n_rows or int(os.environ.get("LADDER_ROWS", "60000"))
    data = bench.light_curve(n_rows)
    n_ens, total = 5, max(12, 2 * world)
    quirks = int(os.environ.get("LADDER_QUIRKS", "3")) if quirks is None else quirks
    st = bench.chain_state(n_ens, total, 5)
    # flat steps so that swaps between neighbours are accepted often
    st["beta"] = np.tile(np.linspace(1.0, 0.9, total), n_ens)
    k0, k1 = total * rank // world, total * (rank + 1) // world
    mine = np.concatenate([np.arange(e * total + k0, e * total + k1) for e in range(n_ens)])

    e = capi.Engine("simplesin5", n_ens, k1 - k0, seed=3, device=local, path=1, quirks=quirks)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    e.ladder_init(uid[0], rank, world, total)
    e.set_data(data)
    e.set_bounds(bench.LO, bench.HI)
    e.set_chains(0, e.n_chains, **{k: v[mine] for k, v in st.items()})
    e.run(12, 5, prob_every=1, params_chains=0)
    out, tr = e.get_chains(), e.read_trace()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, out, tr["prob"]))
    err, msg = None, None
    if rank == 0:
        try:
            f = capi.Engine("simplesin5", n_ens, total, seed=3, device=local, path=1, quirks=quirks)
            f.set_data(data)
            f.set_bounds(bench.LO, bench.HI)
            f.set_chains(0, f.n_chains, **st)
            f.run(12, 5, prob_every=1, params_chains=0)
            full, trf = f.get_chains(), f.read_trace()
            f.close()
            straddling = 0
            for idx, o, trp in gathered:
                for k in o:
                    if k == "rng_counter" or k in full:
                        assert np.array_equal(o[k], full[k][idx]), k
                assert np.array_equal(trp, trf["prob"][:, idx])
            # swaps across the GPU boundary did happen: the last rung of a block counts them
            for r in range(world - 1):
                idx, o, _ = gathered[r]
                last = (np.arange(len(idx)) % (len(idx) // n_ens)) == len(idx) // n_ens - 1
                straddling += int(o["swapcount"][last].sum())
            assert straddling > 0, "no swap across a GPU boundary was accepted; the check is vacuous"
            msg = (f"ladder_check ok: {world} ranks x {total / world:g} rungs == 1 GPU x {total} rungs, bit for bit; "
                   f"{int(full['swapcount'].sum())} swaps, {straddling} across a GPU boundary (quirks={quirks})")
        except AssertionError as ex:
            err = "ladder-split run differs from the single-GPU run (quirks=%d): %s" % (quirks, str(ex)[:300])
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
