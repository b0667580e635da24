"""The C++ host-side mirror of the reference interface (include/spenso_b200.hpp): it must compile against the
C ABI on the CPU box, and on a GPU the reference-style tests written with it (tests/cpp) must pass."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "reference_style_tests")


def build_cpp_tests():
    subprocess.check_call(
        ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", EXE, os.path.join(CPP, "reference_style_tests.cpp"),
         "-L" + os.path.join(ROOT, "spenso_b200"), "-lspenso_b200", "-Wl,-rpath," + os.path.join(ROOT, "spenso_b200")])


def test_cpp_mirror_compiles_and_links():
    build_cpp_tests()
    assert subprocess.run([EXE, "--compile-check"]).returncode == 0


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
