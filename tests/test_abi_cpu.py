"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads,
exports every symbol include/apemost_gpu.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "apemost_gpu.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    from apemost_b200 import capi
    return capi.load_library()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apm_gpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("apm_gpu_create", "apm_gpu_destroy", "apm_gpu_set_data", "apm_gpu_set_bounds",
                 "apm_gpu_set_chains", "apm_gpu_get_chains", "apm_gpu_eval", "apm_gpu_run",
                 "apm_gpu_read_trace", "apm_gpu_calibrate", "apm_gpu_get_stats", "apm_gpu_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_model_table(lib):
    assert lib.apm_gpu_abi_version() == 6
    assert lib.apm_gpu_model_n_par(0) == 4 and lib.apm_gpu_model_n_par(1) == 4
    assert lib.apm_gpu_model_n_par(2) == 1 and lib.apm_gpu_model_n_par(3) == 7
    assert lib.apm_gpu_model_n_cols(2) == 0 and lib.apm_gpu_model_n_cols(3) == 2


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "apemost_gpu.h"\nint main(void){apm_gpu_config c; (void)c; return APM_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_sass_is_blackwell_native(lib):
    """the likelihood kernel must carry TMA bulk copies (UBLKCP) and fp64 FMAs for sm_100a"""
    so = os.path.join(ROOT, "apemost_b200", "libapemost_gpu.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out and "DFMA" in out and "SYNCS" in out


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback(lib):
    from apemost_b200 import capi
    with pytest.raises(capi.EngineError) as ei:
        capi.Engine("simplesin", 1, 2)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
