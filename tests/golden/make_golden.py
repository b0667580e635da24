"""Generate the committed golden fixtures tests/golden/*.npz.

The reference (Rust) cannot run in this image, so these vectors are produced by the ORACLE — the C++ restatement of
the reference's Contract / Network::execute path, itself pinned on the reference's own known-answer tests
(tests/test_oracle_kats.py).  They freeze the oracle's outputs for the BASELINE workloads on fixed seeded inputs:
the CPU suite checks that the oracle still reproduces them bit for bit, the GPU suite checks the CUDA path against
them without running the oracle.  cfg1 additionally carries the closed-form answer of the reference's own test
(spenso-hep-lib/tests/gamma_algebra_validate.rs:29-70, p=(0,1,2,3), q=(1,2,3,4)).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _oracle as O  # noqa: E402
from spenso_b200 import workloads  # noqa: E402

CASES = {"cfg1": (64, 11, 1.0, True), "cfg2": (64, 12, 1.0, False), "cfg3": (16, 13, 1.0, False),
         "cfg5": (64, 14, 100.0, True), "gl03": (8, 15, 100.0, True)}


def main():
    for name, (n, seed, scale, real) in CASES.items():
        spec = workloads.WORKLOADS[name]()
        arrays = workloads.synthetic_inputs(spec, n, seed, scale=scale, real=real)
        if name == "cfg1":  # first point: the kinematics of the reference's own test
            arrays[0][0] = [0, 1, 2, 3]
            arrays[1][0] = [1, 2, 3, 4]
        out, s, slots = O.eval_spec(spec, arrays)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), out=out, sum=s, slots=slots,
                            **{f"in{k}": a for k, a in enumerate(arrays)})
        print(name, out.shape, "max|out| =", np.abs(out).max())


if __name__ == "__main__":
    main()
