"""Device-side leaf filling (spn_net_leaf_function): per-point leaves computed inside the fused kernel
from raw kinematics.  The oracle side fills the same leaves on the host with numpy (tests/_expr_eval.py) and runs
the reference-style network on them."""
import numpy as np
import pytest

import _expr_eval
import _oracle as O
import spenso_b200 as sp
from spenso_b200 import expr as X
from spenso_b200 import workloads


def host_spinors(momenta):
    """The four spinor leaves of eemumu_from_momenta, filled on the host: [vbar(p2), u(p1), ubar(p3), v(p4)]."""
    out = []
    for k, key in ((1, "ubar-"), (0, "u-"), (2, "ubar-"), (3, "u-")):
        comps = workloads.massless_spinors(X.Param.vector(0, 4))[key]
        out.append(np.stack([_expr_eval.evaluate(c, {0: momenta[k]}) for c in comps], axis=1))
    return out


def test_massless_spinors_solve_the_dirac_equation():
    """pslash u(p) = 0 and ubar(p) pslash = 0 for every helicity, with the gamma matrices of spenso-hep-lib
    (gamma_data_weyl, lib.rs:66-102) taken from the oracle."""
    g = O.OLib.builtin(O.OLib.GAMMA_WEYL).instantiate(O.slots(("bis", 4, 1), ("bis", 4, 2), ("mink", 4, 3))).to_dense()
    eta = np.diag([1.0, -1.0, -1.0, -1.0])
    p = workloads.massless_momenta(50, 7)[0]
    pslash = np.einsum("ijm,mn,pn->pij", g, eta, p)
    sp_ = {k: np.stack([_expr_eval.evaluate(c, {0: p}) for c in v], axis=1)
           for k, v in workloads.massless_spinors(X.Param.vector(0, 4)).items()}
    scale = np.abs(p).max() ** 1.5
    for key in ("u+", "u-"):
        assert np.abs(np.einsum("pij,pj->pi", pslash, sp_[key])).max() <= 1e-12 * scale
    for key in ("ubar+", "ubar-"):
        assert np.abs(np.einsum("pi,pij->pj", sp_[key], pslash)).max() <= 1e-12 * scale
    g0 = g[:, :, 0]
    assert np.abs(sp_["ubar+"] - sp_["u+"].conj() @ g0).max() <= 1e-12 * scale   # ubar = u^dagger gamma^0
    assert np.abs(sp_["ubar-"] - sp_["u-"].conj() @ g0).max() <= 1e-12 * scale


def test_function_leaf_compiles_plan_only_and_round_trips_json():
    net = workloads.build_engine(workloads.eemumu_from_momenta(), None)
    assert net.n_inputs == 4 and net.result_size == 1
    assert "sqrt(" in net.cuda_source
    net2 = sp.Network.from_json(None, net.to_json()).compile()
    assert net2.cuda_source == net.cuda_source
    # errors: sqrt of a complex-typed quantity, wrong component count, stepwise executor
    n = sp.Network(None)
    leaf = n.batched([sp.Minkowski.new_slot(4, 1)])
    with pytest.raises(sp.SpensoError):
        n.function([sp.Bispinor.new_slot(4, 2)], [X.Param(leaf, 0)] * 3)
    f = n.function([sp.Bispinor.new_slot(4, 2)], [X.sqrt(X.Param(leaf, 0))] * 4)
    n.set_root(f)
    with pytest.raises(sp.SpensoError):
        n.compile()
    n = sp.Network(None)
    leaf = n.batched([sp.Minkowski.new_slot(4, 1)])
    n.set_root(n.function([sp.Bispinor.new_slot(4, 2)], [X.re(X.Param(leaf, k)) for k in range(4)]))
    n.compile(sp.STEPWISE)   # compiles; executing it stepwise is refused (GPU test)


@pytest.mark.gpu
def test_leaf_filling_vs_host_filled_oracle():
    ctx = sp.Context(0)
    n = 3000
    momenta = workloads.massless_momenta(n, 99)
    ref, ref_sum, _ = O.eval_spec(workloads.eemumu(), host_spinors(momenta))
    spec = workloads.eemumu_from_momenta()
    net = workloads.build_engine(spec, ctx)
    ins = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], momenta):
        t = sp.BatchTensor.empty(ctx, node[1], n)
        t.upload(a)
        ins.append(t)
    got = net.evaluate(ins).to_numpy().reshape(n, -1)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-12 * scale
    out = np.zeros((n, 1), dtype=np.complex128)
    g2, s = net.execute_host(momenta, out=out, want_sum=True, chunk_points=700)
    assert np.abs(g2 - ref).max() <= 1e-12 * scale and abs(s[0] - ref_sum[0]) <= 1e-12 * np.abs(ref).sum()
    # a general expression leaf: mixes every opcode, complex parameters
    rng = np.random.default_rng(5)
    Z = rng.uniform(0.5, 2, size=(n, 3)) + 1j * rng.uniform(0.5, 2, size=(n, 3))
    net = sp.Network(ctx)
    z = net.batched([sp.Euclidean.new_slot(3, 1)])
    P = X.Param.vector(z, 3)
    comps = [P[0] * P[1] / P[2] - P[0].conj(), X.sqrt(X.re(P[0]) * X.im(P[1])) + 2.5j, -(P[2] / X.re(P[1])) + 1.0 / P[0],
             X.im(P[2]) - 3.0]
    f = net.function([sp.Bispinor.new_slot(4, 5)], comps)
    w = net.batched([sp.Bispinor.new_slot(4, 5)])
    net.set_root(net.sum(net.product(f, w), net.product(net.scalar(2.0), f, w)))
    net.compile()
    W = rng.uniform(-1, 1, size=(n, 4)) + 1j * rng.uniform(-1, 1, size=(n, 4))
    tz = sp.BatchTensor.empty(ctx, [sp.Euclidean.new_slot(3, 1)], n)
    tw = sp.BatchTensor.empty(ctx, [sp.Bispinor.new_slot(4, 5)], n)
    tz.upload(Z)
    tw.upload(W)
    got = net.evaluate([tz, tw]).to_numpy().reshape(-1)
    F = np.stack([_expr_eval.evaluate(c, {z: Z}) for c in comps], axis=1)
    want = 3.0 * (F * W).sum(axis=1)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # the stepwise executor materialises a function leaf with a small fused kernel of its own, then proceeds pairwise
    net2 = sp.Network(ctx)
    z2 = net2.batched([sp.Euclidean.new_slot(3, 1)])
    P2 = X.Param.vector(z2, 3)
    comps2 = [P2[0] * P2[1] / P2[2] - P2[0].conj(), X.sqrt(X.re(P2[0]) * X.im(P2[1])) + 2.5j,
              -(P2[2] / X.re(P2[1])) + 1.0 / P2[0], X.im(P2[2]) - 3.0]
    f2 = net2.function([sp.Bispinor.new_slot(4, 5)], comps2)
    w2 = net2.batched([sp.Bispinor.new_slot(4, 5)])
    net2.set_root(net2.sum(net2.product(f2, w2), net2.product(net2.scalar(2.0), f2, w2)))
    net2.compile(sp.STEPWISE)
    got2 = net2.evaluate([tz, tw]).to_numpy().reshape(-1)
    assert np.abs(got2 - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.gpu
def test_leaf_filling_from_f64_momenta():
    """Same network fed with f64 momenta (32 B per momentum): half the input traffic, same amplitudes."""
    ctx = sp.Context(0)
    n = 2048
    momenta = [np.ascontiguousarray(m.real) for m in workloads.massless_momenta(n, 5)]
    spec = workloads.eemumu_from_real_momenta()
    ref, _, _ = O.eval_spec(spec, momenta)
    net = workloads.build_engine(spec, ctx)
    assert net.input_bytes_per_point == 128
    ins = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], momenta):
        t = sp.BatchTensor.empty(ctx, node[1], n, sp.F64)
        t.upload(a)
        ins.append(t)
    got = net.evaluate(ins).to_numpy().reshape(n, -1)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    out = np.zeros((n, 1), dtype=np.complex128)
    g2, _ = net.execute_host(momenta, out=out, chunk_points=500)
    assert np.abs(g2 - ref).max() <= 1e-12 * np.abs(ref).max()
