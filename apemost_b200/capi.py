"""ctypes binding of the engine's C ABI (include/apemost_gpu.h).

This is harness plumbing for tests/ and bench.py: the product is the C-ABI
shared library ``apemost_b200/libapemost_gpu.so`` (CUDA, sm_100a) and the C
host layer in ``apemost_b200/host``; nothing here computes anything.  The
library is loaded from the package directory (in-tree build, see
``__graft_entry__.build``); a missing library raises -- there is no fallback
of any kind.

``EngineBase`` is written against a (library, prefix, config-struct) triple so
that tests/ can point the very same wrapper at the CPU oracle
(``oracle/_build/liboracle.so``, prefix ``orc_``), whose API mirrors this one.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

# ---- constants (mirror include/apemost_gpu.h) --------------------------------
MODEL_SIMPLESIN, MODEL_SIMPLESIN5, MODEL_NORMAL, MODEL_PULSE_VROT = 0, 1, 2, 3
MODEL_SIMPLESIN2, MODEL_PULSE, MODEL_BERNOULLI = 4, 5, 6
MODELS = {
    "simplesin": MODEL_SIMPLESIN, "simplesin5": MODEL_SIMPLESIN5, "normal": MODEL_NORMAL,
    "pulse_vrot": MODEL_PULSE_VROT, "simplesin2": MODEL_SIMPLESIN2, "pulse": MODEL_PULSE,
    "bernoulli": MODEL_BERNOULLI,
}
PROPOSAL_GAUSSIAN, PROPOSAL_LOGISTIC, PROPOSAL_UNIFORM = 0, 1, 2
QUIRK_STALE_PROB_ON_SWAP, QUIRK_STALE_PRIOR_ON_REJECT = 1, 2
QUIRKS_REFERENCE = 3
PATH_AUTO, PATH_TILED, PATH_FUSED, PATH_CLUSTER, PATH_GRID = 0, 1, 2, 3, 4
E_CALIB = -6

_u64 = C.c_ulonglong
_pd = C.POINTER(C.c_double)
_pu = C.POINTER(_u64)

_CHAIN_FIELDS = [
    ("beta", "d", 0), ("params", "d", 1), ("steps", "d", 1), ("prob", "d", 0), ("prior", "d", 0),
    ("prob_best", "d", 0), ("params_best", "d", 1), ("accept", "u", 0), ("reject", "u", 0),
    ("params_accepts", "u", 1), ("params_rejects", "u", 1), ("n_iter", "u", 0),
    ("swapcount", "u", 0), ("rng_counter", "u", 0),
]


class ChainIO(C.Structure):
    _fields_ = [(name, _pd if kind == "d" else _pu) for name, kind, _ in _CHAIN_FIELDS]


class TraceCfg(C.Structure):
    _fields_ = [("prob_every", C.c_int), ("params_chains", C.c_int)]


class CalibCfg(C.Structure):
    _fields_ = [
        ("burn_in_iterations", _u64), ("desired_acceptance_rate", C.c_double),
        ("max_ar_deviation", C.c_double), ("iter_limit", _u64), ("mul", C.c_double),
        ("adjust_step", C.c_double), ("skip_calibrate", C.c_int), ("iter_readjust", C.c_int),
        ("no_rescaling_limit", C.c_int),
    ]


class CalibProgress(C.Structure):
    _fields_ = [("chain", C.c_int), ("param", C.c_int), ("iter", _u64),
                ("step_normalised", C.c_double), ("accept_rate", C.c_double)]


class GpuConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("model_id", C.c_int), ("n_ensembles", C.c_int), ("n_beta", C.c_int),
        ("n_par", C.c_int), ("seed", _u64), ("proposal", C.c_int), ("circular_mask", C.c_uint),
        ("quirks", C.c_uint), ("path", C.c_int), ("chain_id_offset", C.c_int),
        ("ensemble_id_offset", C.c_int), ("model_const", C.c_double * 4),
    ]


def _as_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _as_u(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class EngineError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"[{code}] {text}")
        self.code = code


class EngineBase:
    """Shared wrapper over an ``<prefix>create/...`` C API."""

    _prefix = "apm_gpu_"

    def __init__(self, lib, cfg_struct, n_ensembles, n_beta, n_par):
        self._lib = lib
        self.n_ensembles, self.n_beta, self.n_par = n_ensembles, n_beta, n_par
        self.n_chains = n_ensembles * n_beta
        self._h = C.c_void_p()
        self._check(self._fn("create")(C.byref(self._h), C.byref(cfg_struct)), creating=True)

    # -- plumbing ------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, rc, creating=False, allow=()):
        if rc == 0 or rc in allow:
            return rc
        text = "error"
        if hasattr(self._lib, self._prefix + "last_error"):
            f = self._fn("last_error")
            f.restype = C.c_char_p
            msg = f(None if creating else self._h)
            text = msg.decode() if msg else text
        raise EngineError(rc, text)

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- inputs --------------------------------------------------------------
    def set_data(self, data):
        data = _as_d(data)
        if data.ndim == 1:
            data = data.reshape(-1, 1)
        self._data_keepalive = data
        self._check(self._fn("set_data")(self._h, data.ctypes.data_as(_pd),
                                         C.c_longlong(data.shape[0]), C.c_int(data.shape[1])))

    def set_bounds(self, pmin, pmax):
        pmin, pmax = _as_d(pmin), _as_d(pmax)
        assert pmin.shape == (self.n_par,) and pmax.shape == (self.n_par,)
        self._check(self._fn("set_bounds")(self._h, pmin.ctypes.data_as(_pd), pmax.ctypes.data_as(_pd)))

    def set_chains(self, first=0, count=None, **fields):
        io = ChainIO()
        keep = []
        for name, kind, vec in _CHAIN_FIELDS:
            if name not in fields or fields[name] is None:
                continue
            a = _as_d(fields[name]) if kind == "d" else _as_u(fields[name])
            n = a.shape[0]
            if count is None:
                count = n
            assert n == count, (name, n, count)
            if vec:
                assert a.shape == (count, self.n_par), (name, a.shape)
            keep.append(a)
            setattr(io, name, a.ctypes.data_as(_pd if kind == "d" else _pu))
        unknown = set(fields) - {f[0] for f in _CHAIN_FIELDS}
        assert not unknown, unknown
        self._check(self._fn("set_chains")(self._h, C.c_int(first), C.c_int(count or 0), C.byref(io)))

    def get_chains(self, first=0, count=None, fields=None):
        if count is None:
            count = self.n_chains - first
        io = ChainIO()
        out = {}
        for name, kind, vec in _CHAIN_FIELDS:
            if fields is not None and name not in fields:
                continue
            shape = (count, self.n_par) if vec else (count,)
            a = np.zeros(shape, dtype=np.float64 if kind == "d" else np.uint64)
            out[name] = a
            setattr(io, name, a.ctypes.data_as(_pd if kind == "d" else _pu))
        self._check(self._fn("get_chains")(self._h, C.c_int(first), C.c_int(count), C.byref(io)))
        return out

    # -- calc_model ------------------------------------------------------------
    def eval(self, params, beta=None):
        params = _as_d(params).reshape(-1, self.n_par)
        n = params.shape[0]
        beta = np.ones(n) if beta is None else _as_d(beta)
        assert beta.shape == (n,)
        prob, prior = np.zeros(n), np.zeros(n)
        self._check(self._fn("eval")(self._h, C.c_int(n), params.ctypes.data_as(_pd),
                                     beta.ctypes.data_as(_pd), prob.ctypes.data_as(_pd),
                                     prior.ctypes.data_as(_pd)))
        return prob, prior

    # -- sampler ---------------------------------------------------------------
    def run(self, n_rounds, n_swap, prob_every=0, params_chains=0):
        tr = TraceCfg(prob_every, params_chains)
        self._last_run = (n_rounds * n_swap, prob_every, params_chains)
        self._check(self._fn("run")(self._h, C.c_longlong(n_rounds), C.c_int(n_swap), C.byref(tr)))

    def read_trace(self):
        n_steps, prob_every, params_chains = self._last_run
        rows = (n_steps + prob_every - 1) // prob_every if prob_every > 0 else 0
        dumped = {0: 0, 1: self.n_ensembles, 2: self.n_chains}[params_chains]
        prob = np.zeros((rows, self.n_chains))
        dl = np.zeros((rows, self.n_chains))
        params = np.zeros((n_steps if dumped else 0, dumped, self.n_par))
        n1, n2 = C.c_longlong(), C.c_longlong()
        self._check(self._fn("read_trace")(
            self._h, prob.ctypes.data_as(_pd) if rows else None,
            dl.ctypes.data_as(_pd) if rows else None,
            params.ctypes.data_as(_pd) if dumped else None, C.byref(n1), C.byref(n2)))
        assert n1.value == rows and n2.value == params.shape[0], (n1.value, rows, n2.value)
        return {"prob": prob, "prob_minus_prior": dl, "params": params}

    # -- calibration -------------------------------------------------------------
    def calibrate(self, select=None, burn_in_iterations=10000, desired_acceptance_rate=0.5,
                  max_ar_deviation=0.01, iter_limit=100000, mul=0.85, adjust_step=0.5,
                  skip_calibrate=False, iter_readjust=0, no_rescaling_limit=0,
                  progress_capacity=0, raise_on_failure=True):
        cfg = CalibCfg(burn_in_iterations, desired_acceptance_rate, max_ar_deviation, iter_limit,
                       mul, adjust_step, int(skip_calibrate), iter_readjust, no_rescaling_limit)
        status = np.zeros(self.n_chains, dtype=np.int32)
        sel = None
        if select is not None:
            sel = np.ascontiguousarray(select, dtype=np.uint8)
            assert sel.shape == (self.n_chains,)
        prog = (CalibProgress * max(progress_capacity, 1))()
        nprog = C.c_longlong()
        rc = self._fn("calibrate")(
            self._h, sel.ctypes.data_as(C.POINTER(C.c_ubyte)) if sel is not None else None,
            C.byref(cfg), status.ctypes.data_as(C.POINTER(C.c_int)),
            prog if progress_capacity else None, C.c_longlong(progress_capacity), C.byref(nprog))
        self._check(rc, allow=() if raise_on_failure else (E_CALIB,))
        rows = [(p.chain, p.param, p.iter, p.step_normalised, p.accept_rate)
                for p in prog[:min(nprog.value, progress_capacity)]]
        return status, rows

    # -- steps with an accept log (assess_acceptance_rate's inner loop) -------------
    def steps(self, kind, n_steps, select=None):
        log = np.zeros((n_steps, self.n_chains), dtype=np.uint8)
        sel = None if select is None else np.ascontiguousarray(select, dtype=np.uint8)
        f = self._fn("steps")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]
        self._check(f(self._h, None if sel is None else sel.ctypes.data, int(kind), int(n_steps), log.ctypes.data))
        return log

    # -- host-side uniforms (the regression calibrator's noise) ----------------------
    def host_uniform(self, g):
        u = C.c_double()
        f = self._fn("host_uniform")
        f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        self._check(f(self._h, int(g), C.byref(u)))
        return u.value

    # -- -DADAPT -------------------------------------------------------------------
    def set_adapt(self, enabled=True, target_acceptance_rate=0.5):
        f = self._fn("set_adapt")
        f.argtypes = [C.c_void_p, C.c_int, C.c_double]
        self._check(f(self._h, int(bool(enabled)), float(target_acceptance_rate)))

    # -- accumulators ------------------------------------------------------------
    def reset_stats(self):
        self._check(self._fn("reset_stats")(self._h))

    def get_stats(self):
        n = np.zeros(self.n_chains, dtype=np.uint64)
        sum_dl = np.zeros(self.n_chains)
        sp = np.zeros((self.n_chains, self.n_par))
        sp2 = np.zeros((self.n_chains, self.n_par))
        self._check(self._fn("get_stats")(self._h, n.ctypes.data_as(_pu), sum_dl.ctypes.data_as(_pd),
                                          sp.ctypes.data_as(_pd), sp2.ctypes.data_as(_pd)))
        return {"n": n, "sum_dl": sum_dl, "sum_params": sp, "sum_params_sq": sp2}


    # -- marginal statistics (histograms, batch means) -----------------------------
    def set_marginals(self, which_chains, n_bins=200, batch_size=0, max_batches=0):
        f = self._fn("set_marginals")
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_ulonglong, C.c_int]
        self._check(f(self._h, int(which_chains), int(n_bins), int(batch_size), int(max_batches)))
        self._marg = (which_chains, n_bins, max_batches)

    def get_marginals(self):
        which, n_bins, cap = self._marg
        slots = {1: self.n_ensembles, 2: self.n_chains}[which]
        counts = np.zeros((slots, self.n_par, n_bins), dtype=np.uint64)
        means = np.zeros((slots, self.n_par, max(cap, 1)))
        n_values = np.zeros(slots, dtype=np.uint64)
        n_batches = np.zeros(slots, dtype=np.uint64)
        f = self._fn("get_marginals")
        f.argtypes = [C.c_void_p] * 5
        self._check(f(self._h, counts.ctypes.data, means.ctypes.data if cap > 0 else None, n_values.ctypes.data,
                      n_batches.ctypes.data))
        return {"counts": counts, "batch_means": means[:, :, :cap], "n_values": n_values, "n_batches": n_batches}


# ---- the product library -------------------------------------------------------
# (APEMOST_GPU_LIB: another build of the same library, e.g. a one-model build under build_variants/)
LIB_PATH = os.environ.get("APEMOST_GPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libapemost_gpu.so")
_lib = None


def load_library(path: Optional[str] = None):
    """Load libapemost_gpu.so (in-tree).  Raises if it has not been built."""
    global _lib
    if _lib is None or path is not None:
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise OSError(
                f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
        lib.apm_gpu_last_error.restype = C.c_char_p
        lib.apm_gpu_launch_count.restype = C.c_longlong
        if path is not None:
            return lib
        _lib = lib
    return _lib


class Engine(EngineBase):
    """The B200 engine behind include/apemost_gpu.h."""

    _prefix = "apm_gpu_"

    def __init__(self, model, n_ensembles, n_beta, n_par=None, seed=1, proposal=PROPOSAL_GAUSSIAN,
                 circular_mask=0, quirks=QUIRKS_REFERENCE, path=PATH_AUTO, device=0,
                 chain_id_offset=0, ensemble_id_offset=0, model_const=None):
        lib = load_library()
        model_id = MODELS[model] if isinstance(model, str) else int(model)
        if n_par is None:
            n_par = lib.apm_gpu_model_n_par(model_id)
            assert n_par > 0, "n_par required for this model"
        cfg = GpuConfig()
        cfg.device, cfg.model_id = device, model_id
        cfg.n_ensembles, cfg.n_beta, cfg.n_par = n_ensembles, n_beta, n_par
        cfg.seed, cfg.proposal, cfg.circular_mask, cfg.quirks = seed, proposal, circular_mask, quirks
        cfg.path, cfg.chain_id_offset, cfg.ensemble_id_offset = path, chain_id_offset, ensemble_id_offset
        for i, v in enumerate(model_const or []):
            cfg.model_const[i] = v
        super().__init__(lib, cfg, n_ensembles, n_beta, n_par)

    def launch_count(self):
        return int(self._lib.apm_gpu_launch_count(self._h))

    def last_path(self):
        return int(self._lib.apm_gpu_last_path(self._h))

    def set_timing(self, per_launch=True):
        """per-launch CUDA events on the tiled path (and no CUDA graph): see apm_gpu_set_timing"""
        self._check(self._lib.apm_gpu_set_timing(self._h, C.c_int(int(bool(per_launch)))))

    def last_kernel_ms(self):
        a, n, t = C.c_double(), C.c_longlong(), C.c_double()
        self._check(self._lib.apm_gpu_last_kernel_ms(self._h, C.byref(a), C.byref(n), C.byref(t)))
        return a.value, n.value, t.value

    def nccl_init(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.apm_gpu_nccl_init(self._h, buf, C.c_int(rank), C.c_int(n_ranks)))

    def ladder_init(self, unique_id: bytes, rank: int, n_ranks: int, n_beta_total: int):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.apm_gpu_ladder_init(self._h, buf, C.c_int(rank), C.c_int(n_ranks),
                                                   C.c_int(n_beta_total)))


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_ubyte * 128)()
    rc = lib.apm_gpu_nccl_unique_id(buf)
    if rc != 0:
        raise EngineError(rc, "ncclGetUniqueId failed")
    return bytes(buf)


def measure_fp64_peak(device=0, seconds=0.2) -> float:
    """Sustained FP64 instruction rate (lane-ops/s) of the device, measured live."""
    lib = load_library()
    out = C.c_double()
    rc = lib.apm_gpu_measure_fp64_peak(C.c_int(device), C.c_double(seconds), C.byref(out))
    if rc != 0:
        raise EngineError(rc, "fp64 peak microbenchmark failed")
    return out.value


def measure_fp64_per_clock(device=0):
    """(FP64 lane-operations per SM per clock, number of SMs): the pipe's issue rate from clock64()."""
    lib = load_library()
    out, n_sm = C.c_double(), C.c_int()
    rc = lib.apm_gpu_measure_fp64_per_clock(C.c_int(device), C.byref(out), C.byref(n_sm))
    if rc != 0:
        raise EngineError(rc, "fp64 per-clock microbenchmark failed")
    return out.value, n_sm.value
