"""CPU-only checks of the drop-in boundary: libhrd_b200.so loads, exports every function
include/hrd.h declares, keeps the enum values the header states, and FAILS LOUDLY when there
is no CUDA device (the product has no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cpu_checkers import TAPS
from hackrfdiags_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hrd.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hrd_[a-z0-9_]+)\s*\(", src)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hrd.h but not exported"
    assert sorted(capi.EXPORTS) == declared, "capi.EXPORTS and include/hrd.h disagree"


def test_exports_are_plain_c_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    names = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for name in _declared_functions():
        assert name in names  # unmangled: extern "C"


def test_abi_version_and_enums_match_header():
    src = open(HEADER).read()
    assert capi.load().hrd_abi_version() == int(re.search(r"#define HRD_ABI_VERSION (\d+)", src).group(1))
    for name, value in (("HRD_MODE_USB", capi.MODE_USB), ("HRD_PARAM_WBFM_DEV", capi.PARAM_WBFM_DEV),
                        ("HRD_UNIT_ALL", capi.UNIT_ALL), ("HRD_ENTRY_256K", capi.ENTRY_256K),
                        ("HRD_MEM_DEVICE", capi.MEM_DEVICE), ("HRD_OPT_PROFILE", capi.OPT_PROFILE)):
        assert int(re.search(rf"{name} = (-?\d+)", src).group(1)) == value, name


def test_taps_are_host_side_and_match_oracle(oracle):
    # hrd_get_taps needs no device: the quantisation rule runs at load time on the host
    for i, name in enumerate(TAPS):
        assert np.array_equal(capi.get_taps(i), oracle.taps(i)), name
    assert capi.get_taps(TAPS.index("ssb_delay"))[-1] == -32768  # (int16_t)round(1.0*32768) wraps


def test_state_record_sizes():
    lib = capi.load()
    rx, tx = lib.hrd_state_bytes_per_stream(capi.RX), lib.hrd_state_bytes_per_stream(capi.TX)
    assert 0 < rx <= 1024 and 0 < tx <= 1024 and rx % 4 == 0 and tx % 4 == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_a_gpu():
    with pytest.raises(capi.HrdError) as e:
        capi.Batch(4, capi.RX)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_argument_errors_do_not_need_a_device():
    lib = capi.load()
    h = C.c_void_p()
    assert lib.hrd_create(0, 0, capi.RX, C.byref(h)) == -1   # HRD_EINVAL: n_streams
    assert lib.hrd_create(0, 4, 7, C.byref(h)) == -1          # HRD_EINVAL: kind
    assert b"kind" in lib.hrd_last_error()
    assert lib.hrd_set_mode(None, 0, 1) == -1
    assert lib.hrd_destroy(None) == 0


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under hackrfdiags_b200/ or include/ may reference it."""
    bad = []
    for base in ("hackrfdiags_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"liboracle|#include.*hrd_oracle|libhrd_ref|cpu_checkers|\bhro_[a-z_]+\(", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_atan2_table_bound_behind_the_branch_free_wrap(oracle):
    """hrd_device.cuh wrap_pi_select relies on |theta_a - theta_b| < 2*pi - 2^-10 for any two table entries."""
    t = oracle.atan2_table().astype(np.float64)
    assert t.max() <= np.float32(np.pi) and t.min() > -3.1338
    assert t.max() - t.min() < 2 * np.pi - 2.0 ** -10
