"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) for the tests.

Reuses the product's EngineBase wrapper -- the oracle's C API mirrors
include/apemost_gpu.h with an ``orc_`` prefix -- so a parity test is literally
the same calls made on two engines.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from apemost_b200.capi import EngineBase, MODELS, QUIRKS_REFERENCE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
RNG_MT19937, RNG_PHILOX = 0, 1
_u64 = C.c_ulonglong


class OrcConfig(C.Structure):
    _fields_ = [
        ("model_id", C.c_int), ("n_ensembles", C.c_int), ("n_beta", C.c_int), ("n_par", C.c_int),
        ("seed", _u64), ("proposal", C.c_int), ("circular_mask", C.c_uint), ("quirks", C.c_uint),
        ("rng_kind", C.c_int), ("chain_id_offset", C.c_int), ("ensemble_id_offset", C.c_int),
        ("model_const", C.c_double * 4), ("n_threads", C.c_int),
    ]


_lib = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)


def oracle_lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(ORACLE_LIB)
        _lib.orc_mod_double.restype = C.c_double
        _lib.orc_mod_double.argtypes = [C.c_double, C.c_double]
        _lib.orc_get_chain_beta.restype = C.c_double
        _lib.orc_get_chain_beta.argtypes = [C.c_uint, C.c_uint, C.c_double]
        _lib.orc_calc_beta_0.restype = C.c_double
        _lib.orc_evidence.restype = C.c_double
        _lib.orc_mt_uniform.restype = C.c_double
    return _lib


N_PAR = {"simplesin": 4, "simplesin5": 4, "normal": 1, "pulse_vrot": 7, "simplesin2": 2}


class Oracle(EngineBase):
    _prefix = "orc_"

    def __init__(self, model, n_ensembles, n_beta, n_par=None, seed=1, proposal=0, circular_mask=0,
                 quirks=QUIRKS_REFERENCE, rng=RNG_PHILOX, chain_id_offset=0, ensemble_id_offset=0,
                 model_const=None, n_threads=None):
        lib = oracle_lib()
        model_id = MODELS[model] if isinstance(model, str) else int(model)
        if n_par is None:
            n_par = N_PAR[model]
        cfg = OrcConfig()
        cfg.model_id, cfg.n_ensembles, cfg.n_beta, cfg.n_par = model_id, n_ensembles, n_beta, n_par
        cfg.seed, cfg.proposal, cfg.circular_mask, cfg.quirks = seed, proposal, circular_mask, quirks
        cfg.rng_kind, cfg.chain_id_offset, cfg.ensemble_id_offset = rng, chain_id_offset, ensemble_id_offset
        cfg.n_threads = n_threads if n_threads is not None else (os.cpu_count() or 1)
        for i, v in enumerate(model_const or []):
            cfg.model_const[i] = v
        super().__init__(lib, cfg, n_ensembles, n_beta, n_par)

    def set_random_swap(self, enabled=True):
        self._check(self._lib.orc_set_random_swap(self._h, int(bool(enabled))))

    def mt_uniform(self):
        return self._lib.orc_mt_uniform(self._h)


def philox(ctr, key):
    out = (C.c_uint * 4)()
    oracle_lib().orc_philox4x32_10((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), out)
    return [int(x) for x in out]


def evidence(beta, mean_dl):
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    mean_dl = np.ascontiguousarray(mean_dl, dtype=np.float64)
    pd = C.POINTER(C.c_double)
    return oracle_lib().orc_evidence(C.c_int(len(beta)), beta.ctypes.data_as(pd), mean_dl.ctypes.data_as(pd))


# ---- driving the real reference (oracle/_ref) ------------------------------------
def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "eval_simplesin.exe"))


def ref_exe(name, ccflags="", suffix=""):
    """Path of a reference binary; builds a specially configured one on demand
    (only possible where /root/reference exists)."""
    path = os.path.join(REF_DIR, f"{name}{suffix}.exe")
    if not os.path.exists(path):
        if not os.path.isdir("/root/reference"):
            raise FileNotFoundError(path)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), path,
                        f"CCFLAGS={ccflags}", f"SUFFIX={suffix}"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return path


def write_params_file(path, rows):
    """rows: (start, min, max, name, step) -- reference src/mcmc_parser.c:47-95"""
    with open(path, "w") as f:
        for start, lo, hi, name, step in rows:
            f.write(f"{start!r}\t{lo!r}\t{hi!r}\t{name}\t{step!r}\n")


def write_data_file(path, data):
    data = np.asarray(data, dtype=np.float64)
    with open(path, "w") as f:
        for row in data:
            f.write("\t".join(f"{v:.17e}" for v in row) + "\n")


def ref_eval(model, workdir, param_vectors):
    """Run the reference's eval_<model>.exe (apps/eval_main.c) on parameter vectors."""
    exe = ref_exe(f"eval_{model}")
    text = "\n".join(" ".join(repr(float(v)) for v in p) for p in param_vectors) + "\n"
    out = subprocess.run([exe], cwd=workdir, input=text, capture_output=True, text=True, check=True).stdout
    vals = [tuple(float(x) for x in line.split()) for line in out.strip().splitlines() if line.strip()]
    return np.array(vals[:len(param_vectors)])
