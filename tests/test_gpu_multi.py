"""Multi-GPU paths (need >= 2 CUDA devices; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_data_sharded_likelihood_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools", "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "shard_check ok" in r.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks_prints_one_line():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29632", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("quirks", ["3", "0"])
def test_ladder_split_matches_single_gpu(quirks):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29633", os.path.join(ROOT, "tools", "ladder_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, LADDER_QUIRKS=quirks))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ladder_check ok" in r.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_host_executable_under_torchrun(tmp_path):
    """`torchrun --no-python simplesin.exe <phase>`: 4 ensembles dealt out over 2 GPUs fill the same
    ens<e>/ directories as one process on one GPU"""
    import hashlib
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import write_params_file
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_phases.json")))
    out = str(tmp_path / "bin")
    os.makedirs(out)
    flags = "-DN_BETA=4 -DN_ENSEMBLES=4 -DBURN_IN_ITERATIONS=600 -DMAX_ITERATIONS=1000"
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "apemost_b200", "host"), f"OUT={out}", f"CCFLAGS={flags}",
                    os.path.join(out, "simplesin.exe")], check=True)
    exe = os.path.join(out, "simplesin.exe")
    hashes = []
    for name, launcher in (("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                    "--master-addr", "127.0.0.1", "--master-port", "29634", "--no-python"]), ("one", [])):
        wd = str(tmp_path / name)
        os.makedirs(wd)
        write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
        open(os.path.join(wd, "data"), "wb").write(open(os.path.join(ROOT, "tests", "golden", "testlc.dat"), "rb").read())
        for phase in ("calibrate_first", "calibrate_rest", "run"):
            r = subprocess.run([*launcher, exe, phase], cwd=wd, capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, GSL_RNG_SEED="4"))
            assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
        h = {}
        for d, _, files in os.walk(wd):
            for f in files:
                if f not in ("params", "data"):
                    p = os.path.join(d, f)
                    h[os.path.relpath(p, wd)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
        hashes.append(h)
    assert hashes[0] == hashes[1] and len(hashes[0]) > 40
