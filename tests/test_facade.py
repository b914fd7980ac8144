"""The header-only C++17 facade (include/wf/windflow_gpu.hpp): WindFlow's builder API and PipeGraph / MultiPipe calls over
libwfb200.so. The application-level test program tests/cpp/test_facade.cu is written like the reference's own GPU tests
(same structs and functors) and checks closed-form sums. CPU: it must compile with nvcc for sm_100a; GPU: it must pass."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade.bin")


def _compile():
    from windflow_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_facade.cu")
    hdrs = [os.path.join(ROOT, "include", "wf", "windflow_gpu.hpp"), os.path.join(ROOT, "windflow_b200", "csrc", "wfb_kernels.cuh"),
            os.path.join(ROOT, "windflow_b200", "csrc", "wfb_launch.cuh"), src]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) > os.path.getmtime(h) for h in hdrs):
        return
    libdir = os.path.join(ROOT, "windflow_b200")
    cmd = ["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
           "--expt-extended-lambda", "-I" + os.path.join(ROOT, "include"), "-o", EXE, src, "-L" + libdir, "-lwfb200",
           "-Xlinker", "-rpath", "-Xlinker", libdir]
    subprocess.check_call(cmd)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_facade_compiles():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_facade_runs_reference_style_graphs():
    _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FACADE_OK" in out.stdout, out.stdout[-3000:]
