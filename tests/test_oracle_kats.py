"""Pin the ORACLE (oracle/spenso_oracle.hpp) against every reproducible known-answer
vector the reference's own tests hold for the contraction path.

The legacy KATs (spenso/src/tests.rs + spenso/src/snapshots/*.snap) were written for an
unsorted VecStructure, so they are compared by index label (einsum over the labelled
arrays), not by raw flat order."""
import itertools

import numpy as np
import pytest

import _oracle as O
from _oracle import OTensor, OLib, ONet, slots, slot


def labelled(t: OTensor):
    return [int(a) for a in t.slots["aind"]], t.to_dense()


def to_order(t: OTensor, ainds):
    """dense array of t with axes ordered as `ainds`."""
    have, arr = labelled(t)
    return np.transpose(arr, [have.index(a) for a in ainds])


# ---- R1: ordering ---------------------------------------------------------
def test_rep_ordering():
    # representation.rs:1442-1459 : euc < mink < lor < lor-dual ; rep before dim
    L = O.lib()
    def cmp(a, b):
        x = np.array([a], dtype=O.SLOT_DTYPE); y = np.array([b], dtype=O.SLOT_DTYPE)
        return L.orc_slot_cmp(O._ptr(x), O._ptr(y))
    e, m, l, ld = slot("euc", 4, 0), slot("mink", 4, 0), slot("lor", 4, 0), slot("lor_d", 4, 0)
    assert cmp(e, m) < 0 and cmp(m, l) < 0 and cmp(l, ld) < 0
    assert cmp(slot("euc", 3, 0), slot("mink", 2, 0)) < 0 and cmp(slot("mink", 2, 0), slot("lor", 1, 0)) < 0
    # all base before all dual, each by |id|
    assert cmp(slot("cof", 3, 0), slot("lor_d", 4, 0)) < 0
    assert cmp(slot("lor_d", 4, 0), slot("coaf", 3, 0)) < 0
    # aind decides last
    assert cmp(slot("euc", 4, 1), slot("euc", 4, 2)) < 0


def test_sorting_reps_any_input_order():
    # representation.rs:622-655 : the three input orders sort to the same list
    a = slots(("euc", 4, 0), ("euc", 4, 0), ("mink", 4, 0))
    b = slots(("euc", 4, 0), ("mink", 4, 0), ("euc", 4, 0))
    c = slots(("mink", 4, 0), ("euc", 4, 0), ("euc", 4, 0))
    sa, sb, sc = (O.order_structure(x)[0] for x in (a, b, c))
    assert (sa == sb).all() and (sa == sc).all()


def test_ordered_structure_blocks():
    # ordered.rs:970-992 (orderedmerge): n_base() == 2
    a = slots(("lor", 3, 2), ("mink", 4, 2), ("lor", 7, 1), ("euc", 4, 2), ("euc", 2, 3))
    canon, perm, base_start, dual_start = O.order_structure(a)
    assert [int(k) for k in canon["rep_kind"]] == [1, 1, 2, 3, 3]
    assert base_start == 3 and dual_start == -1          # 3 self-dual(ish), 2 base, 0 dual
    assert [int(d) for d in canon["dim"]] == [2, 4, 4, 3, 7]


# ---- R5: merge ------------------------------------------------------------
def test_merge_dual_interleaved():
    # ordered.rs:942-968
    a = slots(("lor_d", 3, 2), ("lor", 3, 4))
    b = slots(("lor_d", 3, 3), ("lor", 3, 1))
    m = O.merge(a, b)
    if m["kind"] == O.INTERLEAVED:
        assert m["partition"].sum() == 2
    m2 = O.merge(b, a)
    if m2["kind"] == O.INTERLEAVED:
        assert m2["partition"].sum() == 2
    assert len(m["slots"]) == 4 and m["common_a"].sum() == 0


def test_merge_mixed():
    a = slots(("lor", 3, 2), ("mink", 4, 2), ("lor", 7, 1), ("euc", 4, 2), ("euc", 2, 3))
    b = slots(("euc", 4, 2), ("mink", 4, 11), ("lor_d", 7, 1))
    m = O.merge(a, b)
    # euc4|2 is common (self-dual equal slots), lor7|1 matches lor_d7|1
    assert m["common_a"].sum() == 2 and m["common_b"].sum() == 2
    assert len(m["slots"]) == 4
    m2 = O.merge(b, a)
    assert (m2["slots"] == m["slots"]).all()


def test_merge_euc_scalar_result():
    a = slots(("euc", 4, 2))
    m = O.merge(a, a)
    assert len(m["slots"]) == 0 and m["common_a"].sum() == 1


# ---- I1/I2: iterators -----------------------------------------------------
def test_core_flat_iterator_kat():
    # iterators/tests/core_iterator_tests.rs:16-42
    st = slots(("euc", 2, 0), ("euc", 3, 0))
    idx, n = O.flat_iter(st, [1, 0], mode=0)
    assert list(idx) == [0, 3]
    idx, n = O.flat_iter(st, [1, 0], mode=1)
    assert n == 3
    # :134-153 paired conjugates
    idx, n = O.flat_iter(st, [1, 0], mode=3)
    assert n == 2
    idx, n = O.flat_iter(st, [1, 0], mode=2)
    assert n == 3


def test_core_expanded_iterator_kat():
    # core_iterator_tests.rs:45-75 : dims sort to (2,3,5); free dim 1 -> [0,5,10]
    st = slots(("euc", 3, 0), ("euc", 5, 0), ("euc", 2, 1))
    canon, *_ = O.order_structure(st)
    assert [int(d) for d in canon["dim"]] == [2, 3, 5]
    idx, neg = O.metric_iter(st, [0, 1, 0])
    assert list(idx) == [0, 5, 10] and not neg.any()


def test_metric_iterator_signs():
    # core_iterator_tests.rs:78-103
    st = slots(("mink", 2, 0), ("mink", 2, 1))
    idx, neg = O.metric_iter(st, [1, 1])
    assert len(idx) == 4 and neg.any() and (~neg).any()
    assert list(neg) == [False, True, True, False]


def _intent_enumeration(dims, free):
    """ascending flat indices whose non-free dims are 0."""
    strides = [1] * len(dims)
    for i in range(len(dims) - 2, -1, -1):
        strides[i] = strides[i + 1] * dims[i + 1]
    out = []
    ranges = [range(d) if f else range(1) for d, f in zip(dims, free)]
    for idx in itertools.product(*ranges):
        out.append(sum(i * s for i, s in zip(idx, strides)))
    return sorted(out)


@pytest.mark.parametrize("seed", range(40))
def test_flat_iterator_matches_intent(seed):
    """The literal stride/shift iterator enumerates, in ascending order, every flat index
    whose fixed dims are 0 — checked on random shapes / patterns."""
    rng = np.random.default_rng(seed)
    order = int(rng.integers(1, 6))
    dims = [int(d) for d in rng.integers(1, 5, size=order)]
    # distinct ainds, one rep => canonical order = sort by (dim, aind); build directly canonical
    dims.sort()
    st = slots(*[("euc", d, i) for i, d in enumerate(dims)])
    free = [int(x) for x in rng.integers(0, 2, size=order)]
    for mode, fr in ((0, free), (1, [1 - f for f in free]), (3, free), (2, [1 - f for f in free])):
        idx, n = O.flat_iter(st, free, mode=mode)
        assert list(idx) == _intent_enumeration(dims, fr), (dims, free, mode)


# ---- K1/K2/K3 KATs --------------------------------------------------------
KAT_DD = [70, 80, 90, 158, 184, 210, 246, 288, 330, 334, 392, 450]


def test_dense_dense_single_contract_kat():
    # tests.rs:416-437 + snapshots/spenso__tests__dense_dense_single_contract.snap
    a = OTensor.dense(slots(("euc", 4, 1), ("euc", 4, 2)), np.arange(1, 17))
    b = OTensor.dense(slots(("euc", 4, 2), ("euc", 3, 3)), np.arange(1, 13))
    f = a.contract(b)
    assert sorted(int(x) for x in f.slots["aind"]) == [1, 3]
    assert list(to_order(f, [1, 3]).real.ravel()) == KAT_DD
    g = b.contract(a)
    assert list(to_order(g, [1, 3]).real.ravel()) == KAT_DD


KAT_DIAG = [1, 2, 3, 8, 10, 12, 21, 24, 27, 40, 44, 48]


def test_sparse_diag_dense_contract_kat():
    # tests.rs:439-473 + ...sparse_diag_dense_contract.snap ; f == g == h == i
    sa = slots(("euc", 4, 1), ("euc", 4, 2))
    sb = slots(("euc", 4, 2), ("euc", 3, 3))
    diag = {(0, 0): 1, (1, 1): 2, (2, 2): 3, (3, 3): 4}
    a_s = OTensor.sparse(sa, diag)
    a_d = OTensor.dense(sa, np.diag([1, 2, 3, 4]))
    b_d = OTensor.dense(sb, np.arange(1, 13))
    b_s = OTensor.sparse(sb, {(i, j): 1 + 3 * i + j for i in range(4) for j in range(3)})
    f = a_s.contract(b_d)
    g = b_d.contract(a_s)
    h = a_s.contract(b_s)
    i = b_d.contract(a_d)
    assert h.is_sparse and not f.is_sparse
    for t in (f, g, h, i):
        assert list(to_order(t, [1, 3]).real.ravel()) == KAT_DIAG


def test_simple_multi_contract_kat():
    # tests.rs:634-686 + ...simple_multi_contract.snap : [36,120,89] x 3, A.B == B.A, all storage combos
    sa = slots(("euc", 3, 1), ("euc", 4, 2), ("lor", 4, 3))
    sb = slots(("euc", 4, 2), ("euc", 3, 4), ("lor_d", 4, 3))
    da = [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 11, 12, 13, 14, 15, 16] * 3
    db = [3, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 0, 0, 3, 2, 1] * 3
    # NB: b's data is laid out [euc4|2][euc3|4][lor_d|3] = 4x3x4 : the written vector is 48 long
    a = OTensor.dense(sa, da)
    b = OTensor.dense(sb, db)
    A = np.array(da).reshape(3, 4, 4)
    B = np.array(db).reshape(4, 3, 4)
    expect = np.einsum("ikm,kjm->ij", A, B)
    assert list(expect.ravel()) == [36, 120, 89] * 3

    def sp(t):
        d = t.to_dense()
        return OTensor.sparse(t.slots, {idx: d[idx] for idx in np.ndindex(*d.shape) if d[idx] != 0})

    combos = [a.contract(b), b.contract(a), a.contract(sp(b)), b.contract(sp(a)), sp(a).contract(sp(b)),
              sp(a).contract(b), sp(b).contract(sp(a))]
    for t in combos:
        assert list(to_order(t, [1, 4]).real.ravel()) == [36, 120, 89] * 3


def test_contract_spensor_kat():
    # tests.rs:1039-1071 : sparse.sparse -> {(0,1):2,(1,0):2}
    a = OTensor.sparse(slots(("euc", 2, 2), ("euc", 2, 1)), {(0, 0): 1.0, (1, 1): 2.0})
    b = OTensor.sparse(slots(("euc", 2, 1), ("euc", 2, 3)), {(1, 0): 1.0, (0, 1): 2.0})
    f = a.contract(b)
    assert f.is_sparse
    arr = to_order(f, [2, 3]).real
    assert arr.tolist() == [[0, 2], [2, 0]]
    assert len(f.entries()) == 2


def test_contract_densor_with_spensor_kat():
    # tests.rs:1164-1195 : -> [1,2,6,8]
    a = OTensor.sparse(slots(("euc", 2, 2), ("euc", 2, 1)), {(0, 0): 1.0, (1, 1): 2.0})
    b = OTensor.dense(slots(("euc", 2, 1), ("euc", 2, 4)), [1.0, 2.0, 3.0, 4.0])
    f = a.contract(b)
    assert not f.is_sparse
    assert list(to_order(f, [2, 4]).real.ravel()) == [1.0, 2.0, 6.0, 8.0]


def test_minkowski_dot_sign():
    # network/library/symbolic.rs:998-1050 : p.q = p0q0 - p1q1 - p2q2 - p3q3
    p = OTensor.dense(slots(("mink", 4, 2)), [2, 3, 5, 7])
    q = OTensor.dense(slots(("mink", 4, 2)), [11, 13, 17, 19])
    r = p.contract(q).to_dense()
    assert r.shape == () and r == 2 * 11 - 3 * 13 - 5 * 17 - 7 * 19
    # sparse operands take the same sign (K-note 1: k+skip > 0 <=> pos > 0)
    ps = OTensor.sparse(slots(("mink", 4, 2)), {(2,): 5, (3,): 7})
    assert ps.contract(q).to_dense() == -5 * 17 - 7 * 19
    assert q.contract(ps).to_dense() == -5 * 17 - 7 * 19
    assert ps.contract(ps).to_dense() == -25 - 49


def test_empty_sparse_error():
    # singlecontract.rs:153-159 : EmptySparse
    e = OTensor.sparse(slots(("euc", 0, 1), ("euc", 2, 0)), {})
    with pytest.raises(O.OracleError) as ei:
        e.contract(OTensor.dense(slots(("euc", 0, 1)), []))
    assert ei.value.code == 1


# ---- D1 + N1 : gamma algebra ------------------------------------------------
def gamma_np():
    g = OLib.builtin(OLib.GAMMA_WEYL).instantiate(slots(("bis", 4, 1), ("bis", 4, 2), ("mink", 4, 3))).to_dense()
    return np.transpose(g, (2, 0, 1))  # [mu][i][j]


def test_clifford_algebra():
    # spenso-hep-lib/src/lib.rs:66-102 : {g^mu, g^nu} = 2 eta^{mu nu}
    g = gamma_np()
    eta = np.diag([1.0, -1, -1, -1])
    for mu in range(4):
        for nu in range(4):
            anti = g[mu] @ g[nu] + g[nu] @ g[mu]
            assert np.allclose(anti, 2 * eta[mu, nu] * np.eye(4))


def test_gamma0_gamma_gamma0_is_adjoint():
    # gamma_algebra_validate.rs:92-105 : g0 g^mu g0 = (g^mu)^dagger
    g = gamma_np()
    g0 = OLib.builtin(OLib.GAMMA0).instantiate(slots(("bis", 4, 1), ("bis", 4, 2))).to_dense()
    for mu in range(4):
        assert np.allclose(g0 @ g[mu] @ g0, g[mu].conj().T)


def test_lib_instantiate_permutes():
    # gamma(2,1,mu): canonical slots [bis|1, bis|2, mink|mu] must hold the TRANSPOSED matrices (N2)
    t = OLib.builtin(OLib.GAMMA_WEYL).instantiate(slots(("bis", 4, 2), ("bis", 4, 1), ("mink", 4, 7)))
    assert [int(a) for a in t.slots["aind"]] == [1, 2, 7]
    g = gamma_np()
    d = t.to_dense()
    for mu in range(4):
        assert np.allclose(d[:, :, mu], g[mu].T)


def four_gamma_net(p, q):
    """p(1)*(p(3)+q(3))*gamma(1,2,1)*gamma(2,3,2)*gamma(3,4,3)*gamma(4,1,4)
    (gamma_algebra_validate.rs:69-70; gamma(i,j,mu) = bis i, bis j, mink mu)."""
    net = ONet()
    G = OLib.builtin(OLib.GAMMA_WEYL)
    P1 = net.tensor(OTensor.dense(slots(("mink", 4, 1)), p))
    P3 = net.tensor(OTensor.dense(slots(("mink", 4, 3)), p))
    Q3 = net.tensor(OTensor.dense(slots(("mink", 4, 3)), q))
    gs = [net.lib_tensor(G, slots(("bis", 4, i), ("bis", 4, j), ("mink", 4, mu)))
          for (i, j, mu) in ((1, 2, 1), (2, 3, 2), (3, 4, 3), (4, 1, 4))]
    root = net.product(P1, net.sum(P3, Q3), *gs)
    net.set_root(root)
    return net


def test_four_gamma_trace_config1():
    # BASELINE config 1; closed form derived in SURVEY §8c:
    # R[mu2,mu4] = 4 (p^mu2 r^mu4 + r^mu2 p^mu4 - (p.r) eta^{mu2 mu4}), r = p + q
    p = np.array([0.0, 1, 2, 3]); q = np.array([1.0, 2, 3, 4])
    net = four_gamma_net(p, q)
    net.execute()
    res = net.result()
    assert [int(a) for a in res.slots["aind"]] == [2, 4]
    R = res.to_dense()
    expect = np.array([[136, 4, 8, 12], [4, -112, 44, 64], [8, 44, -56, 116], [12, 64, 116, 32]], dtype=float)
    assert np.abs(R - expect).max() <= 1e-12 * np.abs(expect).max()
    # independent numpy evaluation with explicit metric
    g = gamma_np(); eta = np.diag([1.0, -1, -1, -1]); r = p + q
    pl, rl = eta @ p, eta @ r
    ref = np.einsum("a,c,aij,bjk,ckl,dli->bd", pl, rl, g, g, g, g)
    assert np.abs(R - ref).max() <= 1e-12 * np.abs(ref).max()


def test_gamma_pair_trace_is_16():
    # gamma_algebra_validate.rs:87-88 : gamma(1,2,1)*gamma(2,1,1) = g^mu_ij g_mu,ji = tr(g^mu g_mu) = 16
    net = ONet()
    G = OLib.builtin(OLib.GAMMA_WEYL)
    a = net.lib_tensor(G, slots(("bis", 4, 1), ("bis", 4, 2), ("mink", 4, 1)))
    b = net.lib_tensor(G, slots(("bis", 4, 2), ("bis", 4, 1), ("mink", 4, 1)))
    net.set_root(net.product(a, b))
    net.execute()
    r = net.result()
    assert r.order == 0
    assert abs(r.to_dense() - 16) < 1e-12


def test_network_sum_neg_scalar():
    # graph.rs:1530-1584 style expression: t3b*(t+t2)*s2*(s+0)
    rng = np.random.default_rng(3)
    sl = slots(("euc", 3, 1), ("euc", 2, 2))
    t = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    t2 = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    t3 = rng.normal(size=(3,)) + 1j * rng.normal(size=(3,))
    net = ONet()
    T, T2 = net.tensor(OTensor.dense(sl, t)), net.tensor(OTensor.dense(sl, t2))
    T3 = net.tensor(OTensor.dense(slots(("euc", 3, 1)), t3))
    s, s2, z = net.scalar(1.5 - 2j), net.scalar(0.25j), net.scalar(0)
    root = net.product(T3, net.sum(T, net.neg(T2)), s2, net.sum(s, z))
    net.set_root(root)
    net.execute()
    r = net.result()
    expect = np.einsum("a,ab->b", t3, t - t2) * 0.25j * (1.5 - 2j)
    assert [int(a) for a in r.slots["aind"]] == [2]
    assert np.abs(r.to_dense() - expect).max() <= 1e-12 * np.abs(expect).max()


def test_eval_batch_matches_single():
    rng = np.random.default_rng(5)
    n = 17
    P = rng.normal(size=(n, 4)) + 0j
    Q = rng.normal(size=(n, 4)) + 0j
    net = ONet()
    G = OLib.builtin(OLib.GAMMA_WEYL)
    P1 = net.tensor(OTensor.dense(slots(("mink", 4, 1)), np.zeros(4)))
    Q3 = net.tensor(OTensor.dense(slots(("mink", 4, 3)), np.zeros(4)))
    gs = [net.lib_tensor(G, slots(("bis", 4, i), ("bis", 4, j), ("mink", 4, mu)))
          for (i, j, mu) in ((1, 2, 1), (2, 3, 2), (3, 4, 3), (4, 1, 4))]
    net.set_root(net.product(P1, Q3, *gs))
    out, s = net.eval_batch([P1, Q3], [P, Q], 16, n_threads=3)
    g = gamma_np(); eta = np.diag([1.0, -1, -1, -1])
    ref = np.einsum("na,nc,aij,bjk,ckl,dli->nbd", P @ eta, Q @ eta, g, g, g, g).reshape(n, 16)
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(s - ref.sum(0)).max() <= 1e-11 * np.abs(ref).max()


def test_gamma_conj_and_adj_identities():
    # spenso-hep-lib/tests/gamma_algebra_validate.rs:87-105:
    #   gamma0(1,2) gamma(2,3,mu) gamma0(3,4) - gammaconj(4,1,mu) = 0      (gamma0 gamma gamma0 = gamma^dagger)
    #   gammaadj(1,2,mu) - gamma0(1,3) gamma(3,4,mu) gamma0(4,2) = 0
    #   gammaadj(1,2,mu) - gammaconj(2,1,mu) = 0
    args = slots(("bis", 4, 1), ("bis", 4, 2), ("mink", 4, 3))
    g = OLib.builtin(OLib.GAMMA_WEYL).instantiate(args).to_dense()
    gc = OLib.builtin(OLib.GAMMA_CONJ).instantiate(args).to_dense()
    ga = OLib.builtin(OLib.GAMMA_ADJ).instantiate(args).to_dense()
    g0 = OLib.builtin(OLib.GAMMA0).instantiate(slots(("bis", 4, 1), ("bis", 4, 2))).to_dense()
    assert (gc == g.conj()).all()
    assert (ga == np.transpose(g.conj(), (1, 0, 2))).all()
    assert (np.einsum("ab,bcm,cd->adm", g0, g, g0) == ga).all()
    assert (ga == np.transpose(gc, (1, 0, 2))).all()
