"""The C++ host-side mirror of the reference interface (include/spenso_b200.hpp): it must compile against the
C ABI on the CPU box, and on a GPU the reference-style tests written with it (tests/cpp) must pass."""
import os
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
