#!/usr/bin/env python
"""bench.py -- chain-steps/s of the parallel-tempering hot path on config C3 of BASELINE.json:
simplesin5 on a 1M-point synthetic light curve, 4096 chains (64 independent ensembles x a
64-rung beta ladder) per GPU.

A "step" is one PT round: n_swap = 2000 // n_beta = 31 Metropolis steps of every chain
(Gaussian proposal -> calc_model over the whole light curve -> accept/reject -> best tracking
-> trace/accumulators) followed by one swap attempt per ensemble.  chain-steps counts
Metropolis steps actually executed (rejected ones included, swaps excluded).

  value     device-timed (CUDA events on the engine's stream), inputs resident in HBM
  e2e       the same work through the C ABI with pinned HOST buffers: every step uploads the
            light curve and the chain state, runs one round, downloads state + traces
  roofline  dominant kernel = loglik_tiled_kernel<simplesin5>; FP64-pipe bound (the likelihood
            is a transcendental map-reduce).  achieved = row-evaluations/s x 17
            FP64 instructions (4 arithmetic + 13 for sin, the ALGORITHMIC count) over the
            kernel's CUDA-event time (a second pass over the same K steps with events around every
            likelihood launch); peak = FP64 lanes per SM per clock measured live (clock64() around
            a DFMA stream) x SMs x the SM clock sampled during the timed region; frac_vs_spec,
            the executed count (18) and round 1's per-second DFMA denominator are printed beside it
  cpu_baseline  the CPU oracle (reference semantics, OpenMP over chains like the reference)
            on this box's cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): independent ensembles are sharded over the ranks, no
data-path collective, weak scaling (4096 chains per GPU) -- that is `value`.  With more than one rank
the line also carries, measured in the same process group right after the headline leg:
  multi_gpu_parity  the engine's two NCCL data planes checked against a single-GPU run on rank 0:
            data-sharded likelihood (per-step ncclAllReduce on the engine's own communicator;
            decisions equal, values to 1e-9, ranks bit-identical) and the ladder split (boundary
            chains traded with ncclSend/ncclRecv; bit-identical, both quirk modes)
  extra.c5          config C5: the 100 M-row curve sharded over the ranks, 4096 replicated chains,
            one all-reduce per Metropolis step (ms per step, all-reduce + control us per step,
            the likelihood kernel's roofline fraction)
  extra.c3_strong   config C3 as BASELINE.json words it -- 4096 chains IN TOTAL over the ranks --
            strong scaling, with its efficiency against the weak leg of the same run and the limiter

--impl reference times the reference's own CPU implementation on the host cores: the reference
engine (src/*.c, OpenMP over chains) with apps/simplesin5.c, whose two stale lines (SURVEY.md D1:
it does not compile as shipped) are fixed by sed at build time -- oracle/_ref/simplesin5fix_b64_*.exe,
built in the container by `make -C oracle refbench`, travels to the GPU box prebuilt.  Where those
binaries are absent the arm falls back to the CPU port (oracle/apm_oracle.c, byte-identical to
the reference build on the models that do compile; kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_ROWS = 1_000_000
N_ENS, N_BETA, N_PAR = 64, 64, 4
N_SWAP = 2000 // N_BETA  # reference src/parallel_tempering.c:228-231
ALG_FP64_PER_ROW = 17    # simplesin5 = 4 arithmetic + 13 (fp64 sin fast path)
TRUTH = np.array([1.3, 7.25, 0.31 * 2 * np.pi, 0.2])
LO, HI = np.array([0.0, 4.0, 0.0, -1.0]), np.array([3.0, 10.0, 2 * np.pi, 1.0])


def light_curve(n=N_ROWS):
    """SURVEY.md 8d synthetic input: x_i = i * 1000/N (keeps 2*pi*f*x < 1e5), seed 12345"""
    rng = np.random.default_rng(12345)
    x = np.arange(n) * (1000.0 / n)
    y = TRUTH[0] * np.sin(2 * np.pi * TRUTH[1] * x + TRUTH[2]) + TRUTH[3] + rng.normal(0, 0.5, n)
    return np.ascontiguousarray(np.stack([x, y], axis=1))


def ladder(n_beta, beta_0=0.01):
    i = n_beta - np.arange(n_beta) - 1
    return beta_0 + (1 - beta_0) / 2 * (1 - np.cos(i * np.pi / (n_beta - 1)))


def chain_state(n_ens, n_beta, seed):
    rng = np.random.default_rng(seed)
    n = n_ens * n_beta
    beta = np.tile(ladder(n_beta), n_ens)
    post_sigma = np.array([7e-4, 3e-7, 1.1e-3, 5e-4])  # posterior widths at beta = 1, 1M rows
    steps = post_sigma[None, :] * beta[:, None] ** -0.5
    params = np.clip(TRUTH[None, :] + rng.normal(0, 1, (n, N_PAR)) * steps, LO, HI)
    return dict(beta=beta, params=params, steps=steps, params_best=params.copy(),
                prob=np.full(n, -1e10), prior=np.zeros(n), prob_best=np.full(n, -1e10))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_oracle_rate(data, seconds=15.0, n_threads=None):
    """chain-steps/s of the CPU oracle on a bounded sample of the same workload: one 64-rung
    ensemble on the full 1M-row table, as many rounds as fit in ~`seconds`."""
    from oracle_binding import Oracle, RNG_PHILOX
    cores = n_threads or os.cpu_count() or 1
    st = chain_state(1, N_BETA, 99)
    o = Oracle("simplesin5", 1, N_BETA, seed=1, rng=RNG_PHILOX, n_threads=cores)
    o.set_data(data)
    o.set_bounds(LO, HI)
    o.set_chains(0, N_BETA, **st)
    o.run(1, 1)                      # untimed: thread start-up, page faults
    t0 = time.perf_counter()
    o.run(1, 1)                      # one Metropolis step of all 64 chains: calibrates the sample size
    t1 = time.perf_counter() - t0
    steps = int(max(1, seconds / max(t1, 1e-6)))
    rounds, n_swap = (steps // N_SWAP, N_SWAP) if steps >= N_SWAP else (1, steps)
    t0 = time.perf_counter()
    o.run(rounds, n_swap)
    dt = time.perf_counter() - t0
    steps = rounds * n_swap
    rate = N_BETA * steps / dt
    sample = (f"1 ensemble x {N_BETA} chains x {steps} Metropolis steps ({rounds} round(s) of {n_swap} + swap) "
              f"on the full {N_ROWS}-row table ({dt:.1f} s), OpenMP over chains")
    return rate, cores, sample, dt / steps * 1e3


REF_EXE = os.path.join(ROOT, "oracle", "_ref", "simplesin5fix_b64_it%d.exe")
REF_ITERS = (31, 124)   # 1 and 4 rounds of n_swap = 31 iterations (oracle/Makefile refbench)


def reference_exe_rate(data, n_threads=None, repeats=1, warm=0, budget_s=150.0):
    """chain-steps/s of the REFERENCE ITSELF on this box's cores: its own engine (src/*.c, OpenMP
    over the chains of one 64-rung ladder) with apps/simplesin5.c (two stale lines fixed by sed,
    SURVEY.md D1), built in the container as oracle/_ref/simplesin5fix_b64_it<N>.exe.  `run` is
    executed from the same calibration_results our arm starts from, for 31 and for 124 iterations;
    the difference removes process start-up and the parsing of the 1M-row data file.  Steps are
    counted from the lines of prob-chain<k>.dump, because the OpenMP build's shared loop counter
    makes it execute fewer steps than it reports.  The pair of runs is one "step" of
    the reference arm: `warm` untimed pairs, then up to `repeats` timed ones (fewer if `budget_s`
    runs out), the rate being total steps over total time.  None if the binaries are absent."""
    import shutil
    import tempfile
    exes = [REF_EXE % it for it in REF_ITERS]
    if not all(os.path.exists(e) for e in exes):
        return None
    cores = n_threads or os.cpu_count() or 1
    st = chain_state(1, N_BETA, 99)
    wd = tempfile.mkdtemp(prefix="apm_ref_")
    try:
        names = ["amplitude", "frequency", "phase", "offset"]
        with open(os.path.join(wd, "params"), "w") as f:
            for j in range(N_PAR):
                f.write("%r\t%r\t%r\t%s\t%r\n" % (float(TRUTH[j]), float(LO[j]), float(HI[j]), names[j],
                                                  float(st["steps"][0, j])))
        np.savetxt(os.path.join(wd, "data"), data, fmt="%.17e", delimiter="\t")
        with open(os.path.join(wd, "calibration_results"), "w") as f:
            for k in range(N_BETA):
                vals = [st["beta"][k], *st["steps"][k], *st["params"][k]]
                f.write("\t".join("%.15e" % v for v in vals) + "\n")
        env = dict(os.environ, OMP_NUM_THREADS=str(cores), GSL_RNG_SEED="1")
        def pair():
            meas = []
            for exe in exes:
                for fn in os.listdir(wd):
                    if fn.endswith(".dump"):
                        os.remove(os.path.join(wd, fn))
                t0 = time.perf_counter()
                r = subprocess.run([exe, "run"], cwd=wd, env=env, capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    return None
                lines = 0
                for k in range(N_BETA):
                    with open(os.path.join(wd, "prob-chain%d.dump" % k), "rb") as f:
                        lines += sum(1 for _ in f)
                meas.append((dt, lines))
            (t_a, n_a), (t_b, n_b) = meas
            return n_b - n_a, t_b - t_a

        t_start = time.perf_counter()
        n_sum, t_sum, done = 0, 0.0, 0
        for i in range(warm + repeats):
            if i > warm and time.perf_counter() - t_start > budget_s:
                break
            m = pair()
            if m is None:
                return None
            if i >= warm:
                n_sum, t_sum, done = n_sum + m[0], t_sum + m[1], done + 1
        rate = n_sum / t_sum
        sample = (f"reference build (src/*.c + apps/simplesin5.c, -DN_BETA={N_BETA}), `run` on the full {N_ROWS}-row "
                  f"table with {cores} OpenMP threads: {n_sum} chain-steps counted from the dump files in "
                  f"{t_sum:.1f} s over {done} timed repeat(s) after {warm} untimed ({REF_ITERS[1]}-iteration run "
                  f"minus {REF_ITERS[0]}-iteration run; {N_BETA * (REF_ITERS[1] - REF_ITERS[0])} per repeat were requested)")
        return rate, cores, sample, t_sum / max(n_sum, 1) * 1e3 * N_BETA
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one likelihood launch, from the newest
    committed `ncu --set full` summary under profiles/ (tools/ncu_summary.py); None if absent"""
    import glob
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_loglik_full.json")), reverse=True):
        try:
            m = json.load(open(path))["launches"][0]["metrics"]
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = m[k].split()
                tot += float(v) * mult[u]
            return tot
        except Exception:
            continue
    return None


C5_TOTAL_ROWS = 100_000_000   # BASELINE.json config 5


def multi_gpu_parity(rank, world, local):
    """tools/shard_check.py and tools/ladder_check.py in this process group (e2, e3 of SURVEY.md 8):
    the engine's own communicators, checked against a single-GPU run on rank 0"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ladder_check
    import shard_check
    out, detail = {}, []
    for key, fn in (("data_sharded", lambda: shard_check.check(rank, world, local)),
                    ("ladder_split", lambda: [ladder_check.check(rank, world, local, quirks=q) for q in (3, 0)])):
        try:
            msg = fn()
            out[key] = "ok"
            detail += msg if isinstance(msg, list) else [msg]
        except AssertionError as ex:   # raised on every rank together (the checks broadcast their verdict)
            out[key] = "FAILED: " + str(ex)[:300]
    out["detail"] = [d for d in detail if d]
    return out


def c5_leg(rank, world, local, peak):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import c5_bench
    total = int(os.environ.get("APM_BENCH_C5_ROWS", C5_TOTAL_ROWS))
    per_gpu = total // world
    # ~57 ms per Metropolis step per 12.5 M rows: keep the leg near two seconds
    n_steps = max(2, min(6, int(12.5e6 * 8 / per_gpu)))
    return c5_bench.measure(rank, world, local, per_gpu, n_steps=n_steps, n_rounds=2, peak=peak)


def c3_strong_leg(rank, world, local, peak, K, weak_value, barrier):
    """config C3 with 4096 chains IN TOTAL: 64 / world ensembles per GPU, K rounds"""
    import torch
    import torch.distributed as dist
    from apemost_b200 import capi
    n_ens = N_ENS // world
    eng = capi.Engine("simplesin5", n_ens, N_BETA, seed=1, device=local,
                      chain_id_offset=rank * n_ens * N_BETA, ensemble_id_offset=rank * n_ens)
    eng.set_data(light_curve())
    eng.set_bounds(LO, HI)
    st = chain_state(N_ENS, N_BETA, 1000)
    mine = slice(rank * n_ens * N_BETA, (rank + 1) * n_ens * N_BETA)
    eng.set_chains(0, eng.n_chains, **{k: v[mine] for k, v in st.items()})
    eng.run(2, N_SWAP, prob_every=1, params_chains=1)
    barrier()
    eng.run(K, N_SWAP, prob_every=1, params_chains=1)
    barrier()
    _, _, total_ms = eng.last_kernel_ms()
    eng.set_timing(True)       # second pass: the likelihood kernel's own time (see main)
    eng.run(K, N_SWAP, prob_every=1, params_chains=1)
    barrier()
    ll_ms, ll_launches, _ = eng.last_kernel_ms()
    t = torch.tensor([total_ms, ll_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, ll_ms_max = [float(v) for v in t.tolist()]
    eng.close()
    if rank != 0:
        return None
    value = N_ENS * N_BETA * N_SWAP * K / (total_ms * 1e-3)
    per_launch = ll_ms / max(ll_launches, 1)
    achieved = n_ens * N_BETA * N_ROWS * ALG_FP64_PER_ROW / (per_launch * 1e-3) if ll_launches else None
    rec = {"workload": f"C3 strong: simplesin5, 1M rows, 4096 chains in total = {n_ens} ensembles x {N_BETA} rungs "
                       f"on each of {world} GPUs, n_swap {N_SWAP}",
           "n_gpus": world, "value": value, "unit": "chain-steps/s", "scaling": "strong", "steps": K,
           "ms_per_step": total_ms / K,
           "efficiency_vs_weak_leg": value / weak_value,
           "efficiency_note": "ideal = the weak leg's whole-job rate (N GPUs x the rate of one GPU holding all 4096 chains)",
           "kernel_share_of_step": (ll_ms_max / total_ms) if ll_launches else None,
           "kernel_ms_per_launch": per_launch if ll_launches else None,
           "limiter": "per Metropolis step one likelihood launch over 1/N of the chains plus one control kernel; "
                      "what is lost is the likelihood kernel's tail (%d chain tiles x splits dealt to 148 SMs) and the "
                      "control kernel + launch gaps, which do not shrink with N" % ((n_ens * N_BETA + 7) // 8)}
    if achieved:
        rec["roofline"] = {"bound": "fp64", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "GFP64-instr/s",
                           "frac": achieved / peak}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the parity checks and the C5 / strong-scaling legs")
    args = ap.parse_args()
    rank, world, local = dist_env()
    K, W = args.steps, max(args.warmup, 0)
    config = {"workload": "C3: simplesin5, 1M-row synthetic light curve, 4096 chains per GPU "
                          "(64 ensembles x 64-rung chebyshev ladder), n_swap 31",
              "n_rows": N_ROWS, "n_chains_per_gpu": N_ENS * N_BETA, "n_ensembles_per_gpu": N_ENS,
              "n_beta": N_BETA, "n_swap": N_SWAP, "parallelism": f"ensembles sharded over {args.gpus} GPU(s)",
              "l2": "256 MiB flush write before the timed region and between e2e steps; the 16 MB "
                    "table is L2-resident by design within a round"}
    data = light_curve()

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        real = reference_exe_rate(data, repeats=K, warm=min(W, 1))
        if real is not None:
            # the reference itself: each of the K "steps" is the same bounded sample of the workload
            # (a 124- and a 31-iteration run of one 64-rung ladder on the full table); ms_per_step is
            # what one full step of the workload (64 ensembles x 31 iterations) would take at that rate
            v, cores, sample, ms = real
            line = {"impl": "reference", "metric": "chain-steps/sec", "value": v, "unit": "chain-steps/s",
                    "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms * N_SWAP * N_ENS,
                    "ms_per_step_note": "extrapolated from the bounded sample to one full step of the workload",
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic", "config": config,
                    "cpu_baseline": {"value": v, "unit": "chain-steps/s", "cores": cores, "kind": "reference",
                                     "sample": sample},
                    "e2e": {"value": v, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            print(json.dumps(line))
            return 0
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
        rates = []
        per = max(2.0, min(args.cpu_seconds, 120.0 / max(K + min(W, 1), 1)))
        for i in range(min(W, 1) + K):
            rate, cores, sample, ms = cpu_oracle_rate(data, seconds=per)
            if i >= min(W, 1):
                rates.append((rate, ms))
        v = float(np.mean([r[0] for r in rates]))
        line = {"impl": "reference", "metric": "chain-steps/sec", "value": v, "unit": "chain-steps/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": float(np.mean([r[1] for r in rates])) * N_SWAP * N_ENS,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "chain-steps/s", "cores": cores, "kind": "port",
                                 "sample": sample + "; the unmodified reference cannot build simplesin5 "
                                           "(SURVEY.md D1), so its CPU port is timed"},
                "e2e": {"value": v, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from apemost_b200 import capi

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = capi.Engine("simplesin5", N_ENS, N_BETA, seed=1, device=local,
                      chain_id_offset=rank * N_ENS * N_BETA, ensemble_id_offset=rank * N_ENS)
    # pinned host buffers for everything that crosses PCIe in the e2e loop
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_data = pin(data)
    st = {k: pin(v) for k, v in chain_state(N_ENS, N_BETA, 1000 + rank).items()}
    eng.set_data(h_data)
    eng.set_bounds(LO, HI)
    eng.set_chains(0, eng.n_chains, **st)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # roofline denominator: the FP64 pipe's issue rate per clock (clock64() inside a short DFMA kernel:
    # 63.8 of the architecture's 64 lanes per SM per clock on this pool's B200s) x SMs x the SM clock
    # sampled during the timed region.  The per-second DFMA microbenchmark of round 1 is kept beside it:
    # a long pure-DFMA load is the one thing that pulls the clocks down (power), so it reads ~8 % low.
    per_clock, n_sm = capi.measure_fp64_per_clock(local)
    peak_sustained_dfma = capi.measure_fp64_peak(local, 0.25)

    # warm-up rounds (untimed)
    if W:
        eng.run(W, N_SWAP, prob_every=1, params_chains=1)
    flush.fill_(1)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    eng.run(K, N_SWAP, prob_every=1, params_chains=1)   # K steps, device-timed inside the engine
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    _, _, total_ms = eng.last_kernel_ms()
    launches = eng.launch_count() - launches0
    # the same K steps once more with every likelihood launch bracketed by CUDA events on the engine's
    # stream (the timed run above replays each round as one CUDA graph and has no per-launch events):
    # the dominant kernel's own launch duration, for the roofline
    eng.set_timing(True)
    eng.run(K, N_SWAP, prob_every=1, params_chains=1)
    barrier()
    ll_ms, ll_launches, instrumented_ms = eng.last_kernel_ms()
    eng.set_timing(False)
    clocks = sampler.summary()
    sm_hz = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
    peak = per_clock * n_sm * sm_hz

    # end to end through the C ABI with host buffers
    h2d = h_data.nbytes + sum(v.nbytes for v in st.values())
    e2e_ms, e2e_parts = [], []
    d2h = 0
    barrier()
    for i in range(1 + K):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.set_data(h_data)
        ta = time.perf_counter()
        eng.set_chains(0, eng.n_chains, **st)
        tb = time.perf_counter()
        eng.run(1, N_SWAP, prob_every=1, params_chains=1)
        tc = time.perf_counter()
        out = eng.get_chains()
        tr = eng.read_trace()
        td = time.perf_counter()
        dt = (td - t0) * 1e3
        d2h = sum(v.nbytes for v in out.values()) + sum(v.nbytes for v in tr.values())
        if i >= 1:
            e2e_ms.append(dt)
            e2e_parts.append([(ta - t0) * 1e3, (tb - ta) * 1e3, (tc - tb) * 1e3, (td - tc) * 1e3])
    e2e_ms = float(np.mean(e2e_ms))
    e2e_parts = [float(v) for v in np.mean(np.array(e2e_parts), axis=0)]

    n_chains = eng.n_chains
    steps_per_round = n_chains * N_SWAP
    per_rank = None
    if world > 1:
        # every rank's own time and SM clock: the job's value is the slowest rank's, and the ranks differ
        # by the GPUs' own clocks under load, not by anything they wait for (no data-path collective)
        mine = torch.tensor([total_ms / K, float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {"ms_per_step": [float(v[0]) for v in every], "sm_mhz": [float(v[1]) for v in every]}
    t = torch.tensor([total_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, wall_ms = [float(v) for v in t.tolist()]
    value = world * steps_per_round * K / (total_ms * 1e-3)
    e2e = world * steps_per_round / (e2e_ms * 1e-3)

    parity = extra = None
    if world > 1 and not args.no_extras:
        eng.close()
        del flush
        torch.cuda.empty_cache()
        parity = multi_gpu_parity(rank, world, local)
        extra = {"c5": c5_leg(rank, world, local, peak),
                 "c3_strong": c3_strong_leg(rank, world, local, peak, K, value, barrier)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
        rate, cores, sample, _ = cpu_oracle_rate(data, seconds=args.cpu_seconds)
        cpu = {"value": rate, "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        row_evals_per_launch = n_chains * N_ROWS
        achieved = row_evals_per_launch * ALG_FP64_PER_ROW / (ll_ms / max(ll_launches, 1) * 1e-3)
        hbm_alg_bytes = N_ROWS * 16 * (n_chains / 8)  # the table once per 8-chain work-item tile
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {
            "metric": "chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": "chain-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                    "breakdown_ms": dict(zip(["set_data", "set_chains", "run", "get_chains+read_trace"], e2e_parts))},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64", "kernel": "loglik_tiled_kernel<ModelSimplesin5>",
                "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "GFP64-instr/s",
                "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, read from the newest "
                                  "committed ncu --set full summary under profiles/ (not measured in this run)",
                "frac_vs_spec": (achieved / (148 * 64 * clocks["sm_mhz"] * 1e6)) if clocks.get("sm_mhz") else None,
                "spec_peak": "148 SMs x 64 FP64 lanes x the SM clock sampled during the timed region",
                "executed_per_row": 18,
                "executed_note": "SASS FP64 instructions per row evaluation: the sine's argument is a separate DMUL + "
                                 "DADD (rounded like the reference's), so 18 are issued for the 17 counted",
                "frac_executed": achieved / peak * 18 / ALG_FP64_PER_ROW,
                "peak_source": "measured live (MEASURED_PEAKS.json has no fp64 entry): %.2f FP64 lane-operations per SM per "
                               "clock (clock64() around a DFMA stream, 2 warps per scheduler x 8 chains) x %d SMs x the "
                               "%.0f MHz sampled during the timed region" % (per_clock, n_sm, sm_hz / 1e6),
                "peak_sustained_dfma": peak_sustained_dfma / 1e9,
                "frac_vs_sustained_dfma": achieved / peak_sustained_dfma,
                "peak_sustained_dfma_note": "round 1's denominator: the per-second rate of a 0.25 s pure-DFMA launch, during "
                                            "which the clocks sag (the kernel measured here runs at full clocks)",
                "algorithmic": "17 FP64 instr per row-evaluation x 4096 chains x 1e6 rows per launch",
                "kernel_ms_per_launch": ll_ms / max(ll_launches, 1), "kernel_launches_timed": int(ll_launches),
                "kernel_share_of_step": ll_ms / total_ms,
                "kernel_timing": "CUDA events around every likelihood launch in a second pass over the same K steps "
                                 "(%.3f ms per step with the events, %.3f ms in the timed region, where a round is one "
                                 "CUDA graph launch)" % (instrumented_ms / K, total_ms / K),
                "hbm": {"achieved": hbm_alg_bytes / (ll_ms / max(ll_launches, 1) * 1e-3) / 1e9,
                        "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "note": "algorithmic bytes = table once per 8-chain tile; it is served by L2 "
                                "(traffic = DRAM bytes of one launch from the committed ncu capture)"},
            },
            "cpu_baseline": cpu, "clocks": clocks, "wall_ms_per_step": wall_ms / K,
            "row_evals_per_s": value * N_ROWS,
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        if parity is not None:
            line["multi_gpu_parity"] = parity
            line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
