"""The C++ host-side mirror of the reference interface (include/spenso_b200.hpp): it must compile against the
C ABI on the CPU box, and on a GPU the reference-style tests written with it (tests/cpp) must pass."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "reference_style_tests")


BLANKET = os.path.join(CPP, "blanket_executor_test")


def build_cpp_tests():
    for exe, src in ((EXE, "reference_style_tests.cpp"), (BLANKET, "blanket_executor_test.cpp")):
        subprocess.check_call(
            ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(CPP, src),
             "-L" + os.path.join(ROOT, "spenso_b200"), "-lspenso_b200", "-Wl,-rpath," + os.path.join(ROOT, "spenso_b200")])


def test_cpp_mirror_compiles_and_links():
    build_cpp_tests()
    assert subprocess.run([EXE, "--compile-check"]).returncode == 0
    assert subprocess.run([BLANKET, "--compile-check"]).returncode == 0


@pytest.mark.gpu
def test_blanket_executor_call_sequence_over_the_c_abi():
    """tests/cpp/blanket_executor_test.cpp: the reference's `impl ExecuteOp for NetworkStore<T, Sc>` (network.rs:1244-1706)
    and SmallestDegree (network/contract.rs:43-498) restated with every trait method = one C-ABI call — the calls a
    Rust shim would make — against the engine's own executor replaying that run's contraction order."""
    build_cpp_tests()
    r = subprocess.run([BLANKET], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("... ok") == 5


@pytest.mark.gpu
def test_cpp_reference_style_tests_pass():
    build_cpp_tests()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("... ok") == 7


def _build_c_example():
    exe = os.path.join(ROOT, "examples", "eemumu_c")
    subprocess.check_call(
        ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O2", "-I" + os.path.join(ROOT, "include"),
         os.path.join(ROOT, "examples", "eemumu.c"), "-L" + os.path.join(ROOT, "spenso_b200"), "-lspenso_b200",
         "-Wl,-rpath," + os.path.join(ROOT, "spenso_b200"), "-o", exe])
    return exe


def test_header_is_valid_c99_and_c_example_links():
    """include/spenso_b200.h is the FFI surface: it must be plain C (what cgo / bindgen / ctypes consume)."""
    _build_c_example()


@pytest.mark.gpu
def test_c_example_runs():
    exe = _build_c_example()
    r = subprocess.run([exe, "50000"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("n = ")][0]
    nums = [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line.split("sum =")[1])]
    assert abs(nums[0] - nums[2]) <= 1e-9 * max(1.0, abs(nums[0])) and abs(nums[1] - nums[3]) <= 1e-9 * max(1.0, abs(nums[1]))
