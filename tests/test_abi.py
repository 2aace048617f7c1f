"""The C-ABI library loads and exports every symbol include/theia_tad.h declares (no GPU needed)."""
import ctypes
import os
import re

from theia_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "theia_tad.h")).read()
    return sorted(set(re.findall(r"^(?:int|void|const char \*)\s*\*?(tad_\w+)\(", src, flags=re.M)))


def test_header_symbols_exported():
    L = _lib.load()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert set(names) == set(_lib.EXPORTS)


def test_abi_version_and_strerror():
    L = _lib.load()
    assert L.tad_abi_version() == 2
    assert L.tad_strerror(0) == b"ok"
    assert L.tad_strerror(-1) == b"invalid argument"
    assert L.tad_strerror(-99) == b"unknown error"


def test_struct_sizes_match_header():
    # sizes the cgo/ctypes bindings rely on (LP64)
    assert ctypes.sizeof(_lib.TadConfig) == 32
    assert ctypes.sizeof(_lib.TadColumns) == 24 + 10 * 8
    assert ctypes.sizeof(_lib.TadJobSpec) == 24 + 8 + 40 + 8
    assert ctypes.sizeof(_lib.TadRows) == 8 + 11 * 8
    assert ctypes.sizeof(_lib.TadStatus) == 16 + 256 + 8 * 8 + 2 * 8 + 10 * 8


def test_header_is_valid_c99_and_cxx_and_example_links(tmp_path):
    """include/theia_tad.h must be consumable from plain C (cgo) and C++; examples/tad_example.c must link against
    the in-tree library (no GPU needed to link)."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    inc = os.path.join(ROOT, "include")
    hdr = os.path.join(inc, "theia_tad.h")
    subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    exe = str(tmp_path / "tad_example")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-I", inc, os.path.join(ROOT, "examples", "tad_example.c"),
                           "-L", os.path.join(ROOT, "theia_b200"), "-ltheia_tad",
                           "-Wl,-rpath," + os.path.join(ROOT, "theia_b200"), "-o", exe])
    assert os.path.exists(exe)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    """The plain-C host (no Python, no torch in the process) drives a job through the C ABI."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    exe = str(tmp_path / "tad_example")
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "tad_example.c"),
                           "-L", os.path.join(ROOT, "theia_b200"), "-ltheia_tad",
                           "-Wl,-rpath," + os.path.join(ROOT, "theia_b200"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "1 series, 64 points" in out.stdout and "throughput 50007861276" in out.stdout and "6/6 stages" in out.stdout


@pytest.mark.gpu
def test_four_worker_threads_share_one_context(tmp_path):
    """The header's threading contract (theia_tad.h: "a tad_ctx may be used from several threads", the controller runs 4
    workers, pkg/controller/util.go:43): four pthreads submit / poll / read / release / cancel distinct jobs on one context;
    every result equals the one the same table gave alone.  (profiles/r02 holds a compute-sanitizer run of the same binary.)"""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    exe = str(tmp_path / "tad_workers")
    subprocess.check_call([gcc, "-std=c99", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "tad_workers.c"),
                           "-L", os.path.join(ROOT, "theia_b200"), "-ltheia_tad",
                           "-Wl,-rpath," + os.path.join(ROOT, "theia_b200"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "4 workers x 6 rounds on one context: 0 failures" in out.stdout
