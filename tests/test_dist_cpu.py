"""Multi-process path on CPU (world_size 2, gloo): independent ensembles shared out over ranks.

The product's multi-GPU mode for config C3 is "one process per GPU, ensembles dealt out by rank,
no data-path collective" (DESIGN.md section 6).  The host layer takes rank / world size from the
torchrun environment and gives every ensemble its GLOBAL number (random streams, ens<e>/
directory), so the result must not depend on the number of processes.  Here two gloo ranks each
run the host executable (linked against the oracle-backed shim in PHILOX mode, tests/host_shim)
on their share of 4 ensembles in one working directory; rank 0 then compares every file with a
single-process run."""
import hashlib
import json
import os
import subprocess
import tempfile

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle_binding import write_params_file
from test_host_cpu import GOLDEN, build_host_over_oracle

FLAGS = ["-DN_BETA=3", "-DN_ENSEMBLES=4", "-DBURN_IN_ITERATIONS=400", "-DMAX_ITERATIONS=1200"]


def _tree_hashes(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            if f in ("params", "data"):
                continue
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return out


def _phases(exe, wd, env):
    for phase in ("calibrate_first", "calibrate_rest", "run"):
        subprocess.run([exe, phase], cwd=wd, env=env, check=True, capture_output=True)


def _worker(rank, world, port, exe, wd_shared, wd_single, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = dict(os.environ, GSL_RNG_SEED="5", APM_TEST_ORACLE_RNG="philox", RANK=str(rank),
               WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    try:
        # the phases are separate programs in the reference's workflow: ranks meet after each
        for phase in ("calibrate_first", "calibrate_rest", "run"):
            subprocess.run([exe, phase], cwd=wd_shared, env=env, check=True, capture_output=True)
            dist.barrier()
        mine = sorted(d for d in os.listdir(wd_shared) if d.startswith("ens"))
        gathered = [None] * world
        dist.all_gather_object(gathered, (rank, mine))
        if rank == 0:
            single_env = dict(env, RANK="0", WORLD_SIZE="1")
            _phases(exe, wd_single, single_env)
            json.dump({"sharded": _tree_hashes(wd_shared), "single": _tree_hashes(wd_single),
                       "dirs": gathered[0][1]}, open(result_path, "w"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_share_the_ensembles_and_reproduce_one_process():
    fx = json.load(open(os.path.join(GOLDEN, "c1_phases.json")))
    exe = build_host_over_oracle("dist4", "simplesin", FLAGS)
    with tempfile.TemporaryDirectory() as tmp:
        wds = [os.path.join(tmp, n) for n in ("shared", "single")]
        for wd in wds:
            os.makedirs(wd)
            write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
            open(os.path.join(wd, "data"), "wb").write(open(os.path.join(GOLDEN, "testlc.dat"), "rb").read())
        result = os.path.join(tmp, "result.json")
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port, exe, wds[0], wds[1], result), nprocs=2, join=True)
        res = json.load(open(result))
    assert res["dirs"] == ["ens0", "ens1", "ens2", "ens3"]
    assert set(res["sharded"]) == set(res["single"]) and len(res["single"]) > 40
    assert res["sharded"] == res["single"]
    # the ensembles really are different runs
    assert len({res["single"][f"ens{e}/frequency-chain-0.prob.dump"] for e in range(4)}) == 4
