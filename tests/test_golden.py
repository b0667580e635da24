"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must still
reproduce them bit for bit (CPU), and the CUDA path must match them to 1e-12 without the oracle (GPU)."""
import os

import numpy as np
import pytest

import _oracle as O
import spenso_b200 as sp
from spenso_b200 import workloads

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["cfg1", "cfg2", "cfg3", "cfg5", "gl03"]


def load(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    arrays = [z[f"in{k}"] for k in range(len([f for f in z.files if f.startswith("in")]))]
    return arrays, z["out"], z["sum"], z["slots"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_bit_for_bit(name):
    arrays, out, s, slots = load(name)
    got, gs, gslots = O.eval_spec(workloads.WORKLOADS[name](), arrays)
    assert gslots.tobytes() == slots.astype(O.SLOT_DTYPE).tobytes()
    assert (got == out).all() and (gs == s).all()


def test_golden_cfg1_first_point_is_the_reference_known_answer():
    arrays, out, _, _ = load("cfg1")
    want = np.array([[136, 4, 8, 12], [4, -112, 44, 64], [8, 44, -56, 116], [12, 64, 116, 32]], dtype=complex)
    assert (out[0].reshape(4, 4) == want).all()
    cf = workloads.gamma_trace4_closed_form(arrays[0].real, arrays[1].real).reshape(len(out), 16)
    assert np.abs(out - cf).max() <= 1e-12 * np.abs(cf).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", [sp.FUSED, sp.STEPWISE])
def test_cuda_path_matches_golden(name, mode):
    arrays, out, s, slots = load(name)
    ctx = sp.Context(0)
    spec = workloads.WORKLOADS[name]()
    net = workloads.build_engine(spec, ctx, mode)
    assert net.result_slots.tobytes() == slots.astype(sp.SLOT_DTYPE).tobytes()
    n = arrays[0].shape[0]
    ins = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
        t = sp.BatchTensor.empty(ctx, node[1], n)
        t.upload(a)
        ins.append(t)
    got = net.evaluate(ins).to_numpy().reshape(n, -1)
    scale = max(np.abs(out).max(), 1e-300)
    assert np.abs(got - out).max() <= 1e-12 * scale
    if mode == sp.FUSED:
        g2, gs = net.execute_host(arrays, out=np.zeros_like(out), want_sum=True)
        assert np.abs(g2 - out).max() <= 1e-12 * scale and np.abs(gs - s).max() <= 1e-12 * max(np.abs(s).max(), 1e-300)
