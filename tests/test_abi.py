"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/wfb200.h declares, and refuses to
compute without a GPU (no CPU fallback). No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from windflow_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "wfb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wfb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    from windflow_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} not exported by libwfb200.so"


def test_version_and_errors(L):
    assert L.wfb_abi_version() == 1
    assert b"success" in L.wfb_error_string(0)
    assert b"no CUDA device" in L.wfb_error_string(-4)
    from windflow_b200 import _lib
    info = _lib.ProgramInfo()
    assert L.wfb_program_info(0, C.byref(info)) == 0 and (info.tuple_bytes, info.result_bytes) == (64, 32)
    assert L.wfb_program_info(1, C.byref(info)) == 0 and (info.tuple_bytes, info.result_bytes) == (16, 24)
    assert L.wfb_program_info(2, C.byref(info)) == 0 and (info.tuple_bytes, info.result_bytes) == (24, 24)
    assert L.wfb_program_info(3, C.byref(info)) == 0 and (info.tuple_bytes, info.result_bytes) == (32, 32)
    assert L.wfb_program_info(99, C.byref(info)) == -2


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.wfb_device_count() == 0
    h = C.c_void_p()
    assert L.wfb_engine_create(C.byref(h), 0) == -4  # WFB_E_NOGPU
    assert L.wfb_ffat_create(C.byref(h), 0, 16, 4, 1, 16, 0, 0, 0) == -4
    from windflow_b200 import ops
    with pytest.raises(RuntimeError):
        ops.Engine(0)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "windflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} references the oracle"
