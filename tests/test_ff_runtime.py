"""include/ff/ff.hpp -- the FastFlow-API-compatible runtime the builder API runs over -- on the CPU: nested pipeline / all-to-all graphs,
combined nodes, channel ids, end-of-stream notification order, ff_poll, the MPMC queue (tests/cpp/test_ff_runtime.cpp); and the
reference's own CPU window test, unmodified, over it (run-to-run invariance of its checksum is the reference's own pass criterion)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_ff_runtime_unit():
    exe = os.path.join(ROOT, "tests", "cpp", "test_ff_runtime.bin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + INC, os.path.join(ROOT, "tests", "cpp", "test_ff_runtime.cpp"), "-o", exe])
    for _ in range(3):  # threads: repeat
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and "ff runtime OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/win_tests"), reason="the reference is only present in the build container")
def test_reference_cpu_test_runs_unmodified_over_the_runtime(tmp_path):
    """tests/win_tests/test_win_fat_cb.cpp of the reference: Source -> Filter -> FlatMap -> Map -> Ffat_Windows(CB) -> Sink with random
    parallelism per run, DETERMINISTIC mode; it aborts unless every run produces the same checksum."""
    exe = str(tmp_path / "t_fat_cb")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-w", "-I" + INC, "-I/root/reference/wf", "-I/root/reference/tests/win_tests",
                           "/root/reference/tests/win_tests/test_win_fat_cb.cpp", "-o", exe])
    out = subprocess.run([exe, "-r", "4", "-l", "20000", "-k", "5", "-w", "50", "-s", "10"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    vals = re.findall(r"value (\d+)", out.stdout)
    assert len(vals) == 4 and len(set(vals)) == 1, out.stdout[-2000:]
