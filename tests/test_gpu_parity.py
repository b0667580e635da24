"""Parity of the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): 1e-12 relative, norm-wise per output tensor
(max|x - y| <= 1e-12 * max|ref|).  Structures, sparse index sets and plans: bit-exact.
The pairwise kernels keep the reference's operation order and IEEE ops (no FMA), so they are also
checked for bit-identical values where the enumeration order is pinned.
"""
import numpy as np
import pytest

import _oracle as O
import _cases
import spenso_b200 as sp
from spenso_b200 import workloads

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    return sp.Context(0)


def rel_err(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = max(np.abs(want).max() if want.size else 0.0, 1e-300)
    return (np.abs(got - want).max() if want.size else 0.0) / scale


def dense_pair(ctx, rng, sl, n_points, layout=sp.POINT_MAJOR):
    """(engine BatchTensor, [oracle tensor per point], canonical host data [n, *shape])."""
    canon = O.order_structure(sl)[0] if len(sl) else sl
    shape = [int(d) for d in canon["dim"]]
    data = _cases.random_complex(rng, [n_points] + shape)
    t = sp.BatchTensor.empty(ctx, canon, n_points, sp.C64, layout)
    t.upload(data.reshape(n_points, -1))
    return t, [O.OTensor.dense(canon, data[p], canonicalize=False) for p in range(n_points)], data


# ---- reference KATs through the device ------------------------------------------------------------------
def test_kat_dense_dense_single_contract(ctx):
    # spenso/src/tests.rs:416-437 + snapshots/spenso__tests__dense_dense_single_contract.snap
    a = sp.DenseTensor.from_data(ctx, [sp.Euclidean.new_slot(4, 1), sp.Euclidean.new_slot(4, 2)], np.arange(1, 17))
    b = sp.DenseTensor.from_data(ctx, [sp.Euclidean.new_slot(4, 2), sp.Euclidean.new_slot(3, 3)], np.arange(1, 13))
    c = a.contract(b)
    want = np.array([70, 80, 90, 158, 184, 210, 246, 288, 330, 334, 392, 450]).reshape(4, 3)
    ainds = [int(x) for x in c.slots["aind"]]
    got = np.transpose(c.to_numpy(), [ainds.index(1), ainds.index(3)])
    assert (got == want).all()


def test_kat_sparse_diag_dense_and_storage_rule(ctx):
    # tests.rs:439-473: diag(1,2,3,4) sparse . B ; result Dense unless both Sparse (contraction.rs:219-224)
    e = [sp.Euclidean.new_slot(4, 1), sp.Euclidean.new_slot(4, 2)]
    a = sp.SparseTensor.from_entries(ctx, e, {(i, i): i + 1 for i in range(4)})
    bs = [sp.Euclidean.new_slot(4, 2), sp.Euclidean.new_slot(3, 3)]
    b = sp.DenseTensor.from_data(ctx, bs, np.arange(1, 13))
    c = a.contract(b)
    assert not c.is_sparse
    ainds = [int(x) for x in c.slots["aind"]]
    got = np.transpose(c.to_numpy(), [ainds.index(1), ainds.index(3)]).reshape(-1)
    assert (got == np.array([1, 2, 3, 8, 10, 12, 21, 24, 27, 40, 44, 48])).all()
    bsp = sp.SparseTensor.from_entries(ctx, bs, {(i, j): 3 * i + j + 1 for i in range(4) for j in range(3)})
    c2 = a.contract(bsp)
    assert c2.is_sparse
    got2 = np.transpose(c2.to_numpy(), [ainds.index(1), ainds.index(3)]).reshape(-1)
    assert (got2 == got).all()


def test_kat_sparse_sparse_index_set(ctx):
    # tests.rs:1164-1195: [(0,0):1,(1,1):2] x [(0,1):1,(1,0):?]... -> {(0,1):2,(1,0):2} as a SET of flat indices
    e12 = [sp.Euclidean.new_slot(2, 1), sp.Euclidean.new_slot(2, 2)]
    e23 = [sp.Euclidean.new_slot(2, 2), sp.Euclidean.new_slot(2, 3)]
    a = sp.SparseTensor.from_entries(ctx, e12, {(0, 0): 1.0, (1, 1): 2.0})
    b = sp.SparseTensor.from_entries(ctx, e23, {(0, 1): 2.0, (1, 0): 1.0})
    oa = O.OTensor.sparse(O.slots(("euc", 2, 1), ("euc", 2, 2)), {(0, 0): 1.0, (1, 1): 2.0})
    ob = O.OTensor.sparse(O.slots(("euc", 2, 2), ("euc", 2, 3)), {(0, 1): 2.0, (1, 0): 1.0})
    assert a.contract(b).entries() == oa.contract(ob).entries() == {1: 2 + 0j, 2: 2 + 0j}


def test_minkowski_sign_convention(ctx):
    # network/library/symbolic.rs:998-1050: p.q = p0 q0 - p1 q1 - p2 q2 - p3 q3
    m = [sp.Minkowski.new_slot(4, 2)]
    p = sp.DenseTensor.from_data(ctx, m, [1, 2, 3, 4])
    q = sp.DenseTensor.from_data(ctx, m, [5, 6, 7, 8])
    assert p.contract(q).to_numpy().reshape(-1)[0] == 5 - 12 - 21 - 32


# ---- pairwise contraction, random structures -----------------------------------------------------------------
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("layout", [sp.POINT_MAJOR, sp.COMPONENT_MAJOR])
def test_contract_batched_batched_vs_oracle(ctx, seed, layout):
    rng = np.random.default_rng(1000 + seed)
    n_points, done, exact = 5, 0, 0
    for _ in range(25):
        a, b = _cases.random_pair(rng)
        ta, oa, _ = dense_pair(ctx, rng, a, n_points, layout)
        tb, ob, _ = dense_pair(ctx, rng, b, n_points, layout)
        try:
            refs = [x.contract(y) for x, y in zip(oa, ob)]
        except O.OracleError as e:
            with pytest.raises(sp.SpensoError):
                ta.contract(tb)
            continue
        c = ta.contract(tb)
        assert c.slots.tobytes() == refs[0].slots.tobytes()          # structure: bit-exact
        want = np.stack([r.to_dense() for r in refs])
        got = c.to_numpy()
        assert rel_err(got, want) <= TOL
        exact += int((got == want).all())
        done += 1
    assert done >= 15
    assert exact == done, f"pairwise kernels should reproduce the reference rounding exactly ({exact}/{done})"


@pytest.mark.parametrize("seed", range(6))
def test_contract_batched_sparse_vs_oracle(ctx, seed):
    """K2: Dense(per point) x Sparse(shared) and Sparse x Dense, both receiver orders."""
    rng = np.random.default_rng(2000 + seed)
    n_points, done = 4, 0
    for it in range(25):
        a, b = _cases.random_pair(rng)
        cb = O.order_structure(b)[0] if len(b) else b
        ent = _cases.random_sparse_entries(rng, [int(d) for d in cb["dim"]])
        ta, oa, _ = dense_pair(ctx, rng, a, n_points)
        tb = sp.SparseTensor.from_entries(ctx, cb, ent)
        ob = O.OTensor.sparse(cb, ent)
        swap = it % 2 == 1
        try:
            refs = [(ob.contract(x) if swap else x.contract(ob)) for x in oa]
        except O.OracleError:
            continue
        c = tb.contract(ta) if swap else ta.contract(tb)
        assert c.storage == sp.BATCHED
        assert c.slots.tobytes() == refs[0].slots.tobytes()
        assert rel_err(c.to_numpy(), np.stack([r.to_dense() for r in refs])) <= TOL
        done += 1
    assert done >= 15


@pytest.mark.parametrize("seed", range(4))
def test_contract_shared_combinations_vs_oracle(ctx, seed):
    """Dense x Dense, Dense x Sparse, Sparse x Dense, Sparse x Sparse at one point: values, result storage
    and (for sparse results) the stored index SET must match the oracle."""
    rng = np.random.default_rng(3000 + seed)
    done = 0
    for it in range(30):
        a, b = _cases.random_pair(rng, max_free=2, max_common=2)
        ca = O.order_structure(a)[0] if len(a) else a
        cb = O.order_structure(b)[0] if len(b) else b
        sa, sb = bool(it & 1), bool(it & 2)
        def mk(c, sparse):
            dims = [int(d) for d in c["dim"]]
            if sparse:
                ent = _cases.random_sparse_entries(rng, dims)
                return sp.SparseTensor.from_entries(ctx, c, ent), O.OTensor.sparse(c, ent)
            d = _cases.random_complex(rng, dims)
            return sp.DenseTensor.from_data(ctx, c, d), O.OTensor.dense(c, d, canonicalize=False)
        ta, oa = mk(ca, sa)
        tb, ob = mk(cb, sb)
        try:
            ref = oa.contract(ob)
        except O.OracleError:
            continue
        c = ta.contract(tb)
        assert c.is_sparse == ref.is_sparse == (sa and sb)
        assert c.slots.tobytes() == ref.slots.tobytes()
        assert rel_err(c.to_numpy(), ref.to_dense()) <= TOL
        if ref.is_sparse:
            assert set(c.entries()) == set(ref.entries())
        done += 1
    assert done >= 20


def test_contract_edge_cases(ctx):
    rng = np.random.default_rng(7)
    # scalar x tensor (exterior product with a rank-0 operand), dim-1 indices
    s, os_, _ = dense_pair(ctx, rng, np.zeros(0, O.SLOT_DTYPE), 3)
    t, ot, _ = dense_pair(ctx, rng, O.slots(("euc", 1, 1), ("mink", 4, 2)), 3)
    for x, y, ox, oy in ((s, t, os_, ot), (t, s, ot, os_)):
        c = x.contract(y)
        want = np.stack([a.contract(b).to_dense() for a, b in zip(ox, oy)])
        assert rel_err(c.to_numpy(), want) <= TOL
    # mismatched point counts
    u, _, _ = dense_pair(ctx, rng, O.slots(("mink", 4, 2)), 2)
    with pytest.raises(sp.SpensoError):
        t.contract(u)
    # zero points: allowed, empty result
    z = sp.BatchTensor.empty(ctx, O.slots(("mink", 4, 2)), 0)
    assert z.contract(z).to_numpy().shape[0] == 0


def test_elementwise_trace_reduce(ctx):
    rng = np.random.default_rng(11)
    sl = O.slots(("bis", 4, 1), ("mink", 4, 2))
    a, oa, da = dense_pair(ctx, rng, sl, 6)
    b, ob, db = dense_pair(ctx, rng, sl, 6)
    assert ((a + b).to_numpy() == da + db).all()
    assert ((a - b).to_numpy() == da - db).all()
    assert ((-a).to_numpy() == -da).all()
    assert (a.conj().to_numpy() == da.conj()).all()
    want = np.stack([x.scalar_mul(0.5 - 2j).to_dense() for x in oa])
    assert (a.scalar_mul(0.5 - 2j).to_numpy() == want).all()
    with pytest.raises(sp.SpensoError):
        a + dense_pair(ctx, rng, O.slots(("bis", 4, 1), ("mink", 4, 3)), 6)[0]
    # Monte-Carlo sum over points
    assert rel_err(a.sum_points(), da.sum(axis=0)) <= 1e-14
    # trace over a repeated Minkowski pair with metric sign, and a repeated bispinor pair
    for reps in ((("mink", 4, 5), ("bis", 2, 1), ("mink", 4, 5)), (("bis", 4, 3), ("bis", 4, 3), ("euc", 3, 9))):
        raw = O.slots(*reps)
        shape = [int(d) for d in raw["dim"]]
        data = _cases.random_complex(rng, [4] + shape)
        # the engine's trace takes the tensor in the caller's slot order (data row major over it)
        canon_like = np.array([(1, 0, 0, d, 900 + i) for i, d in enumerate(shape)], dtype=sp.SLOT_DTYPE)
        holder = sp.BatchTensor.empty(ctx, canon_like, 4)
        holder.upload(data.reshape(4, -1))
        got = holder.trace(raw)
        refs = []
        for p in range(4):
            h = O.lib().orc_tensor_dense_raw(O._ptr(raw), len(raw), O._ptr(np.ascontiguousarray(data[p]).view(np.float64)))
            refs.append(O.OTensor(h).trace().to_dense())
        assert rel_err(got.to_numpy(), np.stack(refs)) <= TOL


def test_empty_sparse_error(ctx):
    # ContractionError::EmptySparse (singlecontract.rs:153-159)
    sl = O.slots(("euc", 3, 1))
    empty_sparse = sp.SparseTensor.from_entries(ctx, sl, {})
    empty_dense = sp.BatchTensor.empty(ctx, sl, 0)
    with pytest.raises(sp.SpensoError) as e:
        empty_sparse.contract(empty_dense)
    assert e.value.code == 1


# ---- networks ------------------------------------------------------------------------------------------------------
def run_network(ctx, spec, arrays, mode, layout=sp.POINT_MAJOR, want_sum=False):
    import torch

    net = workloads.build_engine(spec, ctx, mode)
    n = arrays[0].shape[0]
    inputs = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
        t = sp.BatchTensor.empty(ctx, node[1], n, sp.C64, layout)
        t.upload(a)
        inputs.append(t)
    out = sp.BatchTensor.empty(ctx, net.result_slots, n, sp.C64, layout)
    s = torch.zeros(2 * net.result_size, dtype=torch.float64, device="cuda:0") if want_sum else None
    net.execute(inputs, out, sum_out_ptr=s.data_ptr() if want_sum else 0)
    ctx.synchronize()
    res = out.to_numpy().reshape(n, -1)
    return (res, s.cpu().numpy().view(np.complex128)) if want_sum else res


def test_cfg1_gamma_trace_closed_form(ctx):
    # spenso-hep-lib/tests/gamma_algebra_validate.rs:29-70 with p=(0,1,2,3), q=(1,2,3,4)
    spec = workloads.gamma_trace4()
    p, q = np.array([[0, 1, 2, 3.0]], dtype=complex), np.array([[1, 2, 3, 4.0]], dtype=complex)
    want = np.array([[136, 4, 8, 12], [4, -112, 44, 64], [8, 44, -56, 116], [12, 64, 116, 32]], dtype=complex)
    for mode in (sp.FUSED, sp.STEPWISE):
        got = run_network(ctx, spec, [p, q], mode)
        assert (got.reshape(4, 4) == want).all()      # integers: exact
    # random kinematics vs the closed form and the oracle
    arrays = workloads.synthetic_inputs(spec, 257, 5, real=True)
    ref, _, _ = O.eval_spec(spec, arrays)
    cf = workloads.gamma_trace4_closed_form(arrays[0].real, arrays[1].real).reshape(257, 16)
    for mode in (sp.FUSED, sp.STEPWISE):
        got = run_network(ctx, spec, arrays, mode)
        assert rel_err(got, ref) <= TOL
        assert rel_err(got, cf) <= 1e-12


@pytest.mark.parametrize("name,n_points,scale", [("cfg2", 1000, 1.0), ("cfg3", 300, 1.0), ("cfg5", 500, 100.0),
                                                 ("gl03", 64, 100.0)])
@pytest.mark.parametrize("mode", [sp.FUSED, sp.STEPWISE])
@pytest.mark.parametrize("layout", [sp.POINT_MAJOR, sp.COMPONENT_MAJOR])
def test_network_vs_oracle(ctx, name, n_points, scale, mode, layout):
    spec = workloads.WORKLOADS[name]()
    arrays = workloads.synthetic_inputs(spec, n_points, 1234, scale=scale, real=name in ("cfg5", "gl03"))
    ref, ref_sum, oslots = O.eval_spec(spec, arrays)
    got, s = run_network(ctx, spec, arrays, mode, layout, want_sum=True)
    if name == "gl03":
        assert not ref.any() and not got.any()     # vanishes identically
        return
    assert rel_err(got, ref) <= TOL
    assert rel_err(s, ref_sum) <= TOL
    assert rel_err(s, got.sum(axis=0)) <= 1e-13


def test_network_ragged_ranges_and_sum_only(ctx):
    import torch

    spec = workloads.eemumu()
    n = 1000
    arrays = workloads.synthetic_inputs(spec, n, 99)
    ref, _, _ = O.eval_spec(spec, arrays)
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    inputs = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
        t = sp.BatchTensor.empty(ctx, node[1], n)
        t.upload(a)
        inputs.append(t)
    s = torch.zeros(2, dtype=torch.float64, device="cuda:0")
    for first, cnt in ((0, 1), (1, 1), (3, 130), (129, 871), (999, 1), (5, 0)):
        net.execute(inputs, None, sum_out_ptr=s.data_ptr(), first_point=first, n_points=cnt)
        ctx.synchronize()
        got = s.cpu().numpy().view(np.complex128)[0]
        want = ref[first:first + cnt, 0].sum() if cnt else (ref[first:, 0].sum())
        assert abs(got - want) <= TOL * max(1.0, np.abs(ref).max() * max(cnt, 1))
    out = sp.BatchTensor.empty(ctx, net.result_slots, n)
    out.upload(np.zeros((n, 1), dtype=complex))
    net.execute(inputs, out, first_point=100, n_points=50)
    ctx.synchronize()
    got = out.to_numpy().reshape(-1)
    assert rel_err(got[100:150], ref[100:150, 0]) <= TOL and not got[:100].any() and not got[150:].any()
    # errors: wrong input count / structure / nothing requested
    with pytest.raises(sp.SpensoError):
        net.execute(inputs[:3], out)
    with pytest.raises(sp.SpensoError):
        net.execute(inputs, None)
    with pytest.raises(sp.SpensoError):
        net.execute([inputs[1], inputs[0], inputs[2], inputs[3]], out)


def test_cfg2_full_size_properties(ctx):
    """BASELINE config 2 at its full size (2^20 points): linearity in one spinor, agreement of the per-point
    output with the fused Monte-Carlo sum, determinism, and an oracle spot check on a sample."""
    import torch

    spec = workloads.eemumu()
    n = 1 << 20
    arrays = workloads.synthetic_inputs(spec, n, 1234)
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    nodes = [x for x in spec.nodes if x[0] == "batched"]

    def run(arrs):
        ins = []
        for node, a in zip(nodes, arrs):
            t = sp.BatchTensor.empty(ctx, node[1], n)
            t.upload(a)
            ins.append(t)
        out = sp.BatchTensor.empty(ctx, net.result_slots, n)
        s = torch.zeros(2, dtype=torch.float64, device="cuda:0")
        net.execute(ins, out, sum_out_ptr=s.data_ptr())
        ctx.synchronize()
        return out.to_numpy().reshape(-1), s.cpu().numpy().view(np.complex128)[0]

    base, s0 = run(arrays)
    again, s1 = run(arrays)
    assert (base == again).all() and s0 == s1                                   # deterministic, incl. the sum
    assert abs(s0 - base.sum()) <= 1e-12 * np.abs(base).sum()
    z = 0.75 - 1.25j
    scaled, _ = run([arrays[0] * z] + arrays[1:])
    assert rel_err(scaled, base * z) <= 1e-13                                   # linear in vbar
    summed, _ = run([arrays[0], arrays[1] + arrays[3][::-1], arrays[2], arrays[3]])
    other, _ = run([arrays[0], arrays[3][::-1], arrays[2], arrays[3]])
    assert rel_err(summed, base + other) <= 1e-13                               # additive in u
    idx = np.random.default_rng(0).choice(n, 2000, replace=False)
    ref, _, _ = O.eval_spec(spec, [a[idx] for a in arrays])
    assert rel_err(base[idx], ref[:, 0]) <= TOL


# ---- K6: contractions that reshape to a complex GEMM (FP64 DMMA kernel, zgemm.cu) ---------------------------
GEMM_CASES = [
    # (slots a, slots b) in argument order; all large enough to take the tensor-pipe path
    ((("euc", 8, 1), ("euc", 8, 2), ("euc", 6, 3), ("euc", 6, 4)), (("euc", 6, 3), ("euc", 6, 4), ("euc", 9, 5), ("euc", 8, 6))),
    # interleaved result + Minkowski metric signs on two contracted indices, ragged tile edges
    ((("mink", 4, 1), ("mink", 4, 2), ("euc", 16, 3), ("euc", 5, 9)), (("mink", 4, 1), ("mink", 4, 2), ("euc", 8, 4), ("euc", 7, 6))),
    # dualizable pair contracted, second-before-first ordering of the result
    ((("lor", 12, 1), ("euc", 40, 7)), (("lor_d", 12, 1), ("euc", 6, 2), ("euc", 7, 3))),
    # single contracted index, K = 64 (SingleContract shape)
    ((("euc", 64, 1), ("bis", 33, 2)), (("euc", 64, 1), ("coad", 35, 3))),
    # cfg4 in miniature: rank-6 tensors, three shared indices (dim 6 -> M = N = K = 216)
    (tuple(("euc", 6, i) for i in (1, 2, 3, 4, 5, 6)), tuple(("euc", 6, i) for i in (4, 5, 6, 7, 8, 9))),
]


@pytest.mark.parametrize("case", range(len(GEMM_CASES)))
@pytest.mark.parametrize("n_points", [1, 3])
def test_zgemm_path_vs_oracle(ctx, case, n_points):
    rng = np.random.default_rng(500 + case)
    a, b = (O.slots(*x) for x in GEMM_CASES[case])
    ta, oa, _ = dense_pair(ctx, rng, a, n_points)
    tb, ob, _ = dense_pair(ctx, rng, b, n_points)
    before = ctx.launch_count
    c = ta.contract(tb)
    assert ctx.launch_count == before + 1
    refs = [x.contract(y) for x, y in zip(oa, ob)]
    assert c.slots.tobytes() == refs[0].slots.tobytes()
    want = np.stack([r.to_dense() for r in refs])
    got = c.to_numpy()
    assert rel_err(got, want) <= TOL
    # and the shared (single point) form of the same contraction
    if n_points == 1:
        ca, cb = O.order_structure(a)[0], O.order_structure(b)[0]
        sa = sp.DenseTensor.from_data(ctx, ca, oa[0].to_dense())
        sb = sp.DenseTensor.from_data(ctx, cb, ob[0].to_dense())
        assert rel_err(sa.contract(sb).to_numpy(), want[0]) <= TOL


def test_cfg4_full_size_rank6_dim32(ctx):
    """BASELINE config 4 at full size: A, B in C^{32^6} (16 GiB each), three shared indices = ZGEMM 32768^3.
    Checked through size-independent properties: a rank-one closed form for the whole result, and sampled
    rows of a random contraction against cuBLAS (torch.matmul) — both checks are test infrastructure."""
    import torch

    ctx.release_cached_memory()
    torch.cuda.empty_cache()
    free, total = torch.cuda.mem_get_info()
    if free < 120 * 2 ** 30:
        pytest.skip("needs ~100 GiB of free HBM")
    d, n = 32, 32 ** 3
    dev = torch.device("cuda:0")
    sa = [sp.Euclidean.new_slot(d, i) for i in (1, 2, 3, 4, 5, 6)]
    sb = [sp.Euclidean.new_slot(d, i) for i in (4, 5, 6, 7, 8, 9)]
    g = torch.Generator(device=dev).manual_seed(1236)

    def rnd(*shape):
        return torch.complex(torch.rand(*shape, generator=g, device=dev, dtype=torch.float64) * 2 - 1,
                             torch.rand(*shape, generator=g, device=dev, dtype=torch.float64) * 2 - 1)

    # (1) rank-one operands: A = x y^T, B = z w^T  =>  C = (y.z) x w^T
    x, y, z, w = rnd(n), rnd(n), rnd(n), rnd(n)
    A = torch.outer(x, y).contiguous()
    B = torch.outer(z, w).contiguous()
    ta = sp.BatchTensor.wrap(ctx, sa, 1, A.data_ptr(), keepalive=A)
    tb = sp.BatchTensor.wrap(ctx, sb, 1, B.data_ptr(), keepalive=B)
    c = ta.contract(tb)
    ctx.synchronize()
    assert [int(v) for v in c.slots["aind"]] == [1, 2, 3, 7, 8, 9] and c.size == n * n
    import ctypes

    # view the engine's result buffer as a torch tensor (no copy): cudaMemcpy D2D into a torch allocation
    C = torch.empty((n, n), dtype=torch.complex128, device=dev)
    torch.cuda.synchronize()
    rt = ctypes.CDLL("libcudart.so.12")
    assert rt.cudaMemcpy(ctypes.c_void_p(C.data_ptr()), ctypes.c_void_p(c.device_ptr), ctypes.c_size_t(n * n * 16), 3) == 0
    del c
    yz = torch.dot(y, z)          # bilinear (no conjugation)
    err = 0.0
    scale = float((yz.abs() * x.abs().max() * w.abs().max()).item())
    for r0 in range(0, n, 4096):
        want = yz * torch.outer(x[r0:r0 + 4096], w)
        err = max(err, float((C[r0:r0 + 4096] - want).abs().max().item()))
    assert err <= 1e-12 * scale
    del A, B, ta, tb
    # (2) random full-rank operands, sampled rows vs cuBLAS
    A = rnd(n, n)
    B = rnd(n, n)
    ta = sp.BatchTensor.wrap(ctx, sa, 1, A.data_ptr(), keepalive=A)
    tb = sp.BatchTensor.wrap(ctx, sb, 1, B.data_ptr(), keepalive=B)
    c = ta.contract(tb)
    ctx.synchronize()
    assert rt.cudaMemcpy(ctypes.c_void_p(C.data_ptr()), ctypes.c_void_p(c.device_ptr), ctypes.c_size_t(n * n * 16), 3) == 0
    del c
    rows = torch.randint(0, n, (48,), generator=torch.Generator().manual_seed(1)).to(dev)
    want = A[rows] @ B
    got = C[rows]
    assert float((got - want).abs().max().item()) <= 1e-12 * float(want.abs().max().item())


# ---- f64 data tensors (RealOrComplexTensor::Real, tensors/complex.rs:66-69, 674-690) -------------------------------
def test_real_tensors_upgrade_like_the_reference(ctx):
    """Real x Real stays real, Real x Complex upgrades f64 -> (x, 0.0) before the complex product
    (upgrading_arithmetic.rs:1285-1293); values must equal the oracle's on the embedded complex data."""
    rng = np.random.default_rng(77)
    n = 6
    a_sl = O.order_structure(O.slots(("mink", 4, 1), ("bis", 4, 2)))[0]
    b_sl = O.order_structure(O.slots(("mink", 4, 1), ("bis", 4, 3)))[0]
    ra = rng.uniform(-1, 1, size=(n, 16))
    rb = rng.uniform(-1, 1, size=(n, 16))
    cb = _cases.random_complex(rng, (n, 16))
    for layout in (sp.POINT_MAJOR, sp.COMPONENT_MAJOR):
        ta = sp.BatchTensor.empty(ctx, a_sl, n, sp.F64, layout)
        ta.upload(ra)
        tb = sp.BatchTensor.empty(ctx, b_sl, n, sp.F64, layout)
        tb.upload(rb)
        tc = sp.BatchTensor.empty(ctx, b_sl, n, sp.C64, layout)
        tc.upload(cb)
        rr = ta.contract(tb)
        rc = ta.contract(tc)
        cr = tc.contract(ta)
        assert rr.dtype == sp.F64 and rc.dtype == sp.C64 and cr.dtype == sp.C64
        o = lambda sl, d: [O.OTensor.dense(sl, d[p].astype(complex), canonicalize=False) for p in range(n)]
        oa, ob, oc = o(a_sl, ra), o(b_sl, rb), o(b_sl, cb)
        want_rr = np.stack([x.contract(y).to_dense() for x, y in zip(oa, ob)])
        want_rc = np.stack([x.contract(y).to_dense() for x, y in zip(oa, oc)])
        want_cr = np.stack([y.contract(x).to_dense() for x, y in zip(oa, oc)])
        assert (rr.to_numpy() == want_rr.real).all() and not want_rr.imag.any()
        assert (rc.to_numpy() == want_rc).all()
        assert (cr.to_numpy() == want_cr).all()
        s2 = ta + ta
        assert s2.dtype == sp.F64 and (s2.to_numpy().reshape(n, 16) == 2 * ra).all()
    # fused network on f64 inputs: p.q with the Minkowski metric, result written as f64 and as c64
    net = sp.Network(ctx)
    p = net.batched([sp.Minkowski.new_slot(4, 1)], sp.F64)
    q = net.batched([sp.Minkowski.new_slot(4, 1)], sp.F64)
    net.set_root(net.product(p, q))
    net.compile(sp.FUSED)
    P, Q = rng.uniform(-1, 1, size=(n, 4)), rng.uniform(-1, 1, size=(n, 4))
    tp = sp.BatchTensor.empty(ctx, [sp.Minkowski.new_slot(4, 1)], n, sp.F64)
    tq = sp.BatchTensor.empty(ctx, [sp.Minkowski.new_slot(4, 1)], n, sp.F64)
    tp.upload(P)
    tq.upload(Q)
    want = P[:, 0] * Q[:, 0] - (P[:, 1:] * Q[:, 1:]).sum(axis=1)
    for dt in (sp.F64, sp.C64):
        out = sp.BatchTensor.empty(ctx, net.result_slots, n, dt)
        net.execute([tp, tq], out)
        ctx.synchronize()
        assert rel_err(out.to_numpy().reshape(-1).real, want) <= 1e-14


# ---- full op coverage of the executor: Power, Function(conj), nested sums (network.rs:1468-1703) -----------------
def _net_pair(ctx, build, arrays, slots_list, mode):
    """Build the same expression on the engine and on the oracle; return (engine out, oracle out)."""
    n = arrays[0].shape[0]
    net = sp.Network(ctx)
    leaves = [net.batched(sl) for sl in slots_list]
    net.set_root(build(net, leaves))
    net.compile(mode)
    ins = []
    for sl, a in zip(slots_list, arrays):
        t = sp.BatchTensor.empty(ctx, sl, n)
        t.upload(a)
        ins.append(t)
    got = net.evaluate(ins).to_numpy().reshape(n, -1)
    want = []
    for p in range(n):
        on = O.ONet()
        ol = []
        for sl, a in zip(slots_list, arrays):
            t = O.OTensor.dense(sp.slots(sl), a[p], canonicalize=False)
            on._keep.append(t)
            ol.append(on.tensor(t))
        on.set_root(build(on, ol))
        on.execute()
        want.append(on.result().to_dense().reshape(-1))
    return got, np.stack(want)


@pytest.mark.parametrize("mode", [sp.FUSED, sp.STEPWISE])
def test_network_power_conj_scalars(ctx, mode):
    rng = np.random.default_rng(31)
    n = 40
    m1, m2 = [sp.Minkowski.new_slot(4, 1)], [sp.Minkowski.new_slot(4, 2)]
    P, Q = _cases.random_complex(rng, (n, 4)), _cases.random_complex(rng, (n, 4))

    def conj(net, x):
        return net.conj(x) if isinstance(net, sp.Network) else net.op(O.N_CONJ, [x])

    def power(net, x, k):
        return net.power(x, k) if isinstance(net, sp.Network) else net.op(O.N_POWER, [x], k)

    cases = [
        # (p.p)^2 : even power of a tensor -> scalar, squared
        (lambda net, l: power(net, l[0], 4), [m1], [P]),
        # 1 / (p.p) : negative exponent on the scalar p^2
        (lambda net, l: power(net, l[0], -2), [m1], [P]),
        # p^3 = (p.p) p : odd power keeps the tensor
        (lambda net, l: power(net, l[0], 3), [m1], [P]),
        # conj(p).q + q.q, times a constant scalar, negated.  (The oracle, like the reference, owns every leaf
        # once — its executor rewrites nodes in place — so a tensor used twice is entered twice.)
        (lambda net, l: net.neg(net.product(net.scalar(0.5 - 2j), net.sum(net.product(conj(net, l[0]), l[1]),
                                                                          net.product(l[2], l[3])))),
         [m1, m1, m1, m1], [P, Q, Q, Q]),
        # the reference treats a rank-0 LocalTensor and a Scalar leaf differently in Power (network.rs:1526-1703):
        # odd power of the rank-0 tensor p.q goes through the squaring loop, a Scalar is multiplied n times
        (lambda net, l: power(net, net.product(l[0], l[1]), 3), [m1, m1], [P, Q]),
        (lambda net, l: net.product(l[0], power(net, net.scalar(1.5 - 0.5j), 3)), [m1], [P]),
        (lambda net, l: net.product(l[0], power(net, net.product(l[1], l[2]), 0)), [m1, m2, m2], [P, Q, Q]),
        # outer product of unconnected factors, scaled by 1/(q.q)
        (lambda net, l: net.product(l[0], l[1], power(net, net.product(l[2], l[3]), -1)), [m1, m2, m2, m2], [P, Q, Q, Q]),
    ]
    for build, sls, arrs in cases:
        got, want = _net_pair(ctx, build, arrs, sls, mode)
        assert rel_err(got, want) <= TOL


# ---- random networks (fuzz): sums of products over random index structures ---------------------------------
def random_netspec(rng, n_factors, with_const=True):
    """Product of n_factors tensors over a random chain/tree of shared indices, free indices sprinkled in;
    some factors are point-independent sparse constants.  Returns a NetSpec."""
    pool = [(sp.Euclidean, 2), (sp.Euclidean, 3), (sp.Bispinor, 4), (sp.Minkowski, 4), (sp.ColorAdjoint, 3),
            (sp.Lorentz, 3), (sp.ColorFundamental, 3)]
    spec = workloads.NetSpec("fuzz")
    aind = iter(range(1, 1000))
    fslots = [[] for _ in range(n_factors)]
    # connect factor k to a random earlier factor (tree), sometimes with a second shared index
    for k in range(1, n_factors):
        for _ in range(1 + int(rng.random() < 0.3)):
            j = int(rng.integers(0, k))
            rep, dim = pool[int(rng.integers(len(pool)))]
            a = next(aind)
            s1 = rep.new_slot(dim, a)
            fslots[j].append(s1)
            fslots[k].append(s1.dual())
    for k in range(n_factors):
        for _ in range(int(rng.integers(0, 2))):
            rep, dim = pool[int(rng.integers(len(pool)))]
            s = rep.new_slot(dim, next(aind))
            fslots[k].append(s.dual() if rng.random() < 0.5 else s)
    ids = []
    for k in range(n_factors):
        sl = [fslots[k][i] for i in rng.permutation(len(fslots[k]))]
        size = int(np.prod([s.dim for s in sl])) if sl else 1
        if with_const and size <= 64 and rng.random() < 0.4 and sl:
            dims = [s.dim for s in sl]
            ids.append(spec.custom(sl, _cases.random_sparse_entries(rng, dims, density=0.4)))
        else:
            ids.append(spec.batched(sl, f"t{k}"))
    if not spec.inputs:       # at least one per-point tensor
        return random_netspec(rng, n_factors, with_const=False)
    prod = spec.product(*ids)
    if rng.random() < 0.5:
        prod = spec.product(spec.scalar(complex(rng.uniform(-2, 2), rng.uniform(-2, 2))), prod)
    if rng.random() < 0.4:
        prod = spec.sum(prod, spec.neg(prod)) if rng.random() < 0.2 else spec.sum(prod, prod)
    spec.root = prod
    return spec


@pytest.mark.parametrize("seed", range(12))
def test_random_networks_vs_oracle(ctx, seed):
    rng = np.random.default_rng(9000 + seed)
    checked = 0
    for _ in range(6):
        spec = random_netspec(rng, int(rng.integers(2, 6)))
        total = max(int(np.prod([sz for _, sz in spec.inputs])), 1)
        n = 7
        arrays = workloads.synthetic_inputs(spec, n, int(rng.integers(1 << 30)))
        try:
            ref, ref_sum, oslots = O.eval_spec(spec, arrays)
        except (O.OracleError, NotImplementedError):
            continue
        if ref.shape[1] > 4096:
            continue
        for mode in (sp.FUSED, sp.STEPWISE):
            try:
                got = run_network(ctx, spec, arrays, mode)
            except sp.SpensoError as e:
                if e.code == 7 and "too large" in str(e):
                    continue
                raise
            assert got.shape == ref.shape
            assert rel_err(got, ref) <= TOL, (seed, mode)
        checked += 1
    assert checked >= 3


# ---- host-buffer entry point (the reference's calling convention: host tensors in, host tensors out) -------------
@pytest.mark.parametrize("name,n_points,scale", [("cfg2", 5000, 1.0), ("cfg5", 3001, 100.0), ("cfg1", 257, 1.0)])
def test_execute_host_pipeline_vs_oracle(ctx, name, n_points, scale):
    spec = workloads.WORKLOADS[name]()
    arrays = workloads.synthetic_inputs(spec, n_points, 4321, scale=scale, real=name == "cfg5")
    ref, ref_sum, _ = O.eval_spec(spec, arrays)
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    for chunk in (0, 1000, 977, n_points, 1):
        if chunk == 1 and n_points > 300:
            continue
        out = np.zeros((n_points, net.result_size), dtype=np.complex128)
        got, s = net.execute_host(arrays, out=out, want_sum=True, chunk_points=chunk)
        assert rel_err(got, ref) <= TOL
        assert rel_err(s, ref_sum) <= TOL
    # sum only / output only; repeated calls are deterministic
    _, s1 = net.execute_host(arrays, out=None, want_sum=True, chunk_points=512)
    _, s2 = net.execute_host(arrays, out=None, want_sum=True, chunk_points=512)
    assert (s1 == s2).all() and rel_err(s1, ref_sum) <= TOL
    out = np.zeros((n_points, net.result_size), dtype=np.complex128)
    got, s = net.execute_host(arrays, out=out)
    assert s is None and rel_err(got, ref) <= TOL
    with pytest.raises(sp.SpensoError):
        net.execute_host(arrays[:1], out=out)
    with pytest.raises(sp.SpensoError):
        net.execute_host(arrays, out=None, want_sum=False)


def test_reference_interleaved_pairing_quirk(ctx):
    """SURVEY K-note 3: MultiContractInterleaved zips the two fibres without a permutation.  Default = every
    contracted slot meets its dual (== oracle default); with the triage switch on the engine reproduces the
    reference's pairing (== oracle with emulate_reference_pairing)."""
    cases = [
        ([(1, 0, 0, 3, 52), (2, 0, 0, 2, 15), (3, 0, 2, 2, 58), (3, 0, -1, 2, 20)],
         [(1, 0, 0, 2, 16), (1, 0, 1, 4, 7), (3, 0, 1, 2, 20), (3, 0, -2, 2, 58)]),
        ([(1, 0, 2, 3, 24), (3, 0, 3, 3, 51), (3, 0, -1, 2, 52)],
         [(1, 0, 0, 3, 39), (2, 0, 0, 2, 44), (3, 0, 1, 2, 52), (3, 0, -3, 3, 51)]),
        ([(1, 0, 1, 2, 35), (3, 0, 2, 2, 30), (3, 0, 2, 2, 61), (3, 0, 2, 2, 62), (3, 0, -1, 3, 59)],
         [(3, 0, 1, 3, 59), (3, 0, 1, 3, 63), (3, 0, -2, 2, 30), (3, 0, -2, 2, 61), (3, 0, -3, 3, 15)]),
    ]
    rng = np.random.default_rng(3)
    try:
        for a, b in cases:
            ca, cb = np.array(a, dtype=sp.SLOT_DTYPE), np.array(b, dtype=sp.SLOT_DTYPE)
            ta, oa, _ = dense_pair(ctx, rng, ca, 3)
            tb, ob, _ = dense_pair(ctx, rng, cb, 3)
            want = np.stack([x.contract(y).to_dense() for x, y in zip(oa, ob)])
            quirk = np.stack([x.contract(y, emulate_reference_pairing=True).to_dense() for x, y in zip(oa, ob)])
            assert np.abs(want - quirk).max() > 1e-6           # the two pairings really differ here
            ctx.set_reference_quirks(False)
            assert rel_err(ta.contract(tb).to_numpy(), want) <= TOL
            ctx.set_reference_quirks(True)
            assert rel_err(ta.contract(tb).to_numpy(), quirk) <= TOL
    finally:
        ctx.set_reference_quirks(False)


# ---- BASELINE configs 3 and 5 at their full sizes: size-independent properties ------------------------------------
def _torch_complex(shape, seed, dev, scale=1.0, real=False):
    import torch

    g = torch.Generator(device=dev).manual_seed(seed)
    re = (torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 2 - 1) * scale
    im = torch.zeros_like(re) if real else (torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 2 - 1) * scale
    return torch.complex(re, im).contiguous()


def test_cfg3_full_size_4M_points(ctx):
    """Config 3 at 2^22 points (39 GB of per-point dense X and D, generated on the device): additivity in D,
    homogeneity in X, per-point output vs fused sum, and an oracle spot check on sampled points."""
    import torch

    ctx.release_cached_memory()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * 2 ** 30:
        pytest.skip("needs ~90 GiB of free HBM")
    dev = torch.device("cuda:0")
    n = 1 << 22
    spec = workloads.su3_colour()
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    slots_x, slots_d = [x[1] for x in spec.nodes if x[0] == "batched"]
    X = _torch_complex((n, 72), 1235, dev)
    D = _torch_complex((n, 512), 1236, dev)

    def run(x, d, want_sum=False):
        tx = sp.BatchTensor.wrap(ctx, slots_x, n, x.data_ptr(), keepalive=x)
        td = sp.BatchTensor.wrap(ctx, slots_d, n, d.data_ptr(), keepalive=d)
        out = torch.empty(n, dtype=torch.complex128, device=dev)
        to = sp.BatchTensor.wrap(ctx, net.result_slots, n, out.data_ptr(), keepalive=out)
        s = torch.zeros(2, dtype=torch.float64, device=dev)
        net.execute([tx, td], to, sum_out_ptr=s.data_ptr() if want_sum else 0)
        ctx.synchronize()
        return out, torch.view_as_complex(s.view(1, 2))[0]

    base, s = run(X, D, want_sum=True)
    assert abs(complex(s) - complex(base.sum())) <= 1e-12 * float(base.abs().sum())
    z = 0.5 - 1.5j
    x_only, _ = run(X, torch.zeros_like(D))
    d_only, _ = run(torch.zeros_like(X), D)
    assert float((x_only + d_only - base).abs().max()) <= 1e-13 * float(base.abs().max())     # X.T + D.f
    del d_only
    X.mul_(z)
    scaled, _ = run(X, torch.zeros_like(D))
    assert float((scaled - z * x_only).abs().max()) <= 1e-13 * float(x_only.abs().max())      # homogeneous in X
    X.div_(z)
    idx = torch.randint(0, n, (300,), generator=torch.Generator().manual_seed(3))
    ref, _, _ = O.eval_spec(spec, [X[idx.to(dev)].cpu().numpy(), D[idx.to(dev)].cpu().numpy()])
    assert rel_err(base[idx.to(dev)].cpu().numpy(), ref[:, 0]) <= 1e-12 * 10   # X was scaled and unscaled: 1 ulp per element


def test_cfg5_1e8_points_streamed_sum(ctx):
    """Config 5's batch size: 1e8 kinematic points resident on one GPU (12.8 GB).  The Monte-Carlo sum over all points
    must equal the sum of the sums over 24 ragged point ranges, be reproducible bit for bit, and agree with the oracle
    on a sample."""
    import torch

    ctx.release_cached_memory()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~30 GiB of free HBM")
    dev = torch.device("cuda:0")
    n = 100_000_000
    spec = workloads.one_loop_massive()
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    sl = [x[1] for x in spec.nodes if x[0] == "batched"]
    K0 = _torch_complex((n, 4), 1237, dev, scale=100.0, real=True)
    K1 = _torch_complex((n, 4), 1238, dev, scale=100.0, real=True)
    t0 = sp.BatchTensor.wrap(ctx, sl[0], n, K0.data_ptr(), keepalive=K0)
    t1 = sp.BatchTensor.wrap(ctx, sl[1], n, K1.data_ptr(), keepalive=K1)
    s = torch.zeros(32, dtype=torch.float64, device=dev)

    def total(first=0, cnt=0):
        net.execute([t0, t1], None, sum_out_ptr=s.data_ptr(), first_point=first, n_points=cnt)
        ctx.synchronize()
        return s.cpu().numpy().view(np.complex128).copy()

    full = total()
    assert (total() == full).all()
    edges = np.linspace(0, n, 25).astype(np.int64)
    edges[1:-1] += np.arange(1, 24) * 7 - 50          # ragged
    parts = sum(total(int(a), int(b - a)) for a, b in zip(edges[:-1], edges[1:]))
    assert np.abs(parts - full).max() <= 1e-11 * np.abs(full).max()      # two different summation trees over 1e8 terms
    idx = torch.randint(0, n, (200,), generator=torch.Generator().manual_seed(4)).to(dev)
    ref, _, _ = O.eval_spec(spec, [K0[idx].cpu().numpy(), K1[idx].cpu().numpy()])
    out = torch.empty((200, 16), dtype=torch.complex128, device=dev)
    k0s, k1s = K0[idx].contiguous(), K1[idx].contiguous()
    a0 = sp.BatchTensor.wrap(ctx, sl[0], 200, k0s.data_ptr(), keepalive=k0s)
    a1 = sp.BatchTensor.wrap(ctx, sl[1], 200, k1s.data_ptr(), keepalive=k1s)
    to = sp.BatchTensor.wrap(ctx, net.result_slots, 200, out.data_ptr(), keepalive=out)
    net.execute([a0, a1], to)
    ctx.synchronize()
    assert rel_err(out.cpu().numpy(), ref) <= TOL


def test_cfg5_real_momenta_variant(ctx):
    """cfg5r: the loop momenta as f64 tensors — same numerator, half the bytes, fewer instructions."""
    import torch

    spec = workloads.one_loop_massive_real()
    n = 1500
    arrays = [np.ascontiguousarray(a.real) for a in
              workloads.synthetic_inputs(workloads.one_loop_massive(), n, 77, scale=100.0, real=True)]
    ref, ref_sum, _ = O.eval_spec(spec, arrays)
    net = workloads.build_engine(spec, ctx)
    assert net.input_bytes_per_point == 64
    ins = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
        t = sp.BatchTensor.empty(ctx, node[1], n, sp.F64)
        t.upload(a)
        ins.append(t)
    out = sp.BatchTensor.empty(ctx, net.result_slots, n)
    s = torch.zeros(32, dtype=torch.float64, device="cuda:0")
    net.execute(ins, out, sum_out_ptr=s.data_ptr())
    ctx.synchronize()
    assert rel_err(out.to_numpy().reshape(n, -1), ref) <= TOL
    assert rel_err(s.cpu().numpy().view(np.complex128), ref_sum) <= TOL


def test_gamma_adjoint_identities_on_device(ctx):
    """spenso-hep-lib/tests/gamma_algebra_validate.rs:87-105 as networks of library tensors only:
    gamma0 gamma^mu gamma0 - gammaconj^T = 0,  gammaadj - gamma0 gamma gamma0 = 0,  gammaadj - gammaconj^T = 0."""
    b = lambda a: sp.Bispinor.new_slot(4, a)
    m = sp.Minkowski.new_slot(4, 1)
    for mode in (sp.FUSED, sp.STEPWISE):
        for build in (
            lambda n: n.sum(n.product(n.library(sp.LIB_GAMMA0_WEYL, [b(1), b(2)]), n.library(sp.LIB_GAMMA_WEYL, [b(2), b(3), m]),
                                      n.library(sp.LIB_GAMMA0_WEYL, [b(3), b(4)])),
                            n.neg(n.library(sp.LIB_GAMMA_CONJ_WEYL, [b(4), b(1), m]))),
            lambda n: n.sum(n.library(sp.LIB_GAMMA_ADJ_WEYL, [b(1), b(2), m]),
                            n.neg(n.product(n.library(sp.LIB_GAMMA0_WEYL, [b(1), b(3)]),
                                            n.library(sp.LIB_GAMMA_WEYL, [b(3), b(4), m]),
                                            n.library(sp.LIB_GAMMA0_WEYL, [b(4), b(2)])))),
            lambda n: n.sum(n.library(sp.LIB_GAMMA_ADJ_WEYL, [b(1), b(2), m]),
                            n.neg(n.library(sp.LIB_GAMMA_CONJ_WEYL, [b(2), b(1), m]))),
        ):
            net = sp.Network(ctx)
            net.set_root(build(net))
            net.compile(mode)
            out = sp.BatchTensor.empty(ctx, net.result_slots, 2)
            out.upload(np.ones((2, 64), dtype=complex))
            net.execute([], out)
            ctx.synchronize()
            assert net.result_size == 64 and not out.to_numpy().any()
    # and a non-trivial constant network: gamma^mu gamma_mu = 4 * identity
    net = sp.Network(ctx)
    net.set_root(net.product(net.library(sp.LIB_GAMMA_WEYL, [b(1), b(2), m]), net.library(sp.LIB_GAMMA_WEYL, [b(2), b(3), m])))
    net.compile(sp.FUSED)
    out = sp.BatchTensor.empty(ctx, net.result_slots, 3)
    net.execute([], out)
    ctx.synchronize()
    assert (out.to_numpy() == 4 * np.eye(4)).all()


@pytest.mark.parametrize("seed", range(3))
def test_specialised_pairwise_kernels_bit_exact(ctx, seed):
    """Batches large enough to take the JIT-specialised pairwise kernels (pairwise_jit.cu): results must be
    bit-identical to the oracle (same operation order, no FMA), for dense x dense, dense x sparse (both receiver
    orders), f64 operands and both layouts."""
    rng = np.random.default_rng(7000 + seed)
    n_points, done = 2048, 0
    for it in range(10):
        a, b = _cases.random_pair(rng, max_free=2, max_common=2)
        layout = sp.POINT_MAJOR if it % 2 == 0 else sp.COMPONENT_MAJOR
        ta, oa, da = dense_pair(ctx, rng, a, n_points, layout)
        sparse_b = it % 3 == 2
        if sparse_b:
            cb = O.order_structure(b)[0] if len(b) else b
            ent = _cases.random_sparse_entries(rng, [int(d) for d in cb["dim"]])
            tb, ob = sp.SparseTensor.from_entries(ctx, cb, ent), [O.OTensor.sparse(cb, ent)] * n_points
        else:
            tb, ob, _ = dense_pair(ctx, rng, b, n_points, layout)
        swap = it % 4 == 3
        sample = rng.choice(n_points, 40, replace=False)
        try:
            refs = [(ob[p].contract(oa[p]) if swap else oa[p].contract(ob[p])) for p in sample]
        except O.OracleError:
            continue
        before = ctx.launch_count
        c = tb.contract(ta) if swap else ta.contract(tb)
        assert ctx.launch_count == before + 1
        got = c.to_numpy()[sample]
        want = np.stack([r.to_dense() for r in refs])
        assert c.slots.tobytes() == refs[0].slots.tobytes()
        assert (got == want).all(), f"case {it}: specialised kernel is not bit-identical"
        done += 1
    assert done >= 5


@pytest.mark.parametrize("mode", [sp.FUSED, sp.STEPWISE])
def test_reference_graph_addition_expression(ctx, mode):
    """The expression of the reference's own graph test (spenso/src/network/graph.rs:1530-1584):
        t3b * (t + t2) * s2 * (s + 0) * (t3 * (t + t2) * s * (1 + 0))
    with t, t2 on [mink1|2, mink2|2], t3 on [lor1|2, euc2|2], t3b on [lor_dual1|2, euc2|1], s = 2, s2 = 3 — nested
    products inside products (merge_ops), sums of tensors, sums of scalars, scalar leaves, dim-1 slots, a
    Lorentz/dual pair and two Minkowski pairs contracted across the nesting."""
    rng = np.random.default_rng(17)
    n = 64
    mk = lambda rep, d, a: rep.new_slot(d, a)
    s_t = [mk(sp.Minkowski, 1, 2), mk(sp.Minkowski, 2, 2)]
    s_t3 = [mk(sp.Euclidean, 2, 2), mk(sp.Lorentz, 1, 2)]                # canonical order (self-dual before dualizable)
    s_t3b = [mk(sp.Euclidean, 2, 1), mk(sp.Lorentz.dual(), 1, 2)]
    T, T2 = _cases.random_complex(rng, (n, 2)), _cases.random_complex(rng, (n, 2))
    T3, T3b = _cases.random_complex(rng, (n, 2)), _cases.random_complex(rng, (n, 2))
    slots_list = [s_t, s_t, s_t, s_t, s_t3, s_t3b]        # every leaf is entered once per use, like the reference's graph
    arrays = [T, T2, T, T2, T3, T3b]

    def build(net, l):
        t_a, t2_a, t_b, t2_b, t3, t3b = l
        s, s2, zero, one = net.scalar(2.0), net.scalar(3.0), net.scalar(0.0), net.scalar(1.0)
        inner = net.product(t3, net.sum(t_b, t2_b), s, net.sum(one, net.scalar(0.0)))
        return net.product(t3b, net.sum(t_a, t2_a), s2, net.sum(s, zero), inner)

    got, want = _net_pair(ctx, build, arrays, slots_list, mode)
    assert got.shape == want.shape == (n, 4)      # free indices: euc2|1, euc2|2
    assert rel_err(got, want) <= TOL
    # closed form: ((t+t2).(t+t2) with the Minkowski metric) * t3b[e1] * t3[e2] * 3 * 2 * 2 * 1
    u = T + T2
    dot = u[:, 0] * u[:, 0] - u[:, 1] * u[:, 1]       # dim-1 slot: index 0 only (+); dim-2 slot: (+,-)
    # [mink1|2] has a single component (sign +); [mink2|2]: components (+,-); data layout [mink1, mink2] -> 2 values
    cf = 12.0 * dot[:, None, None] * T3b[:, :, None] * T3[:, None, :]
    assert rel_err(got.reshape(n, 2, 2), cf) <= 1e-12


@pytest.mark.parametrize("case", range(len(GEMM_CASES)))
def test_dgemm_path_vs_oracle(ctx, case):
    """f64 x f64 GEMM-shaped contractions take the real DMMA kernel (contract_dgemm_kernel); result stays f64."""
    rng = np.random.default_rng(600 + case)
    a, b = (O.slots(*x) for x in GEMM_CASES[case])
    ca, cb = O.order_structure(a)[0], O.order_structure(b)[0]
    n_points = 2
    da = rng.uniform(-1, 1, size=[n_points] + [int(d) for d in ca["dim"]])
    db = rng.uniform(-1, 1, size=[n_points] + [int(d) for d in cb["dim"]])
    ta = sp.BatchTensor.empty(ctx, ca, n_points, sp.F64)
    ta.upload(da.reshape(n_points, -1))
    tb = sp.BatchTensor.empty(ctx, cb, n_points, sp.F64)
    tb.upload(db.reshape(n_points, -1))
    c = ta.contract(tb)
    assert c.dtype == sp.F64
    refs = [O.OTensor.dense(ca, da[p].astype(complex), canonicalize=False).contract(
        O.OTensor.dense(cb, db[p].astype(complex), canonicalize=False)) for p in range(n_points)]
    want = np.stack([r.to_dense() for r in refs])
    assert c.slots.tobytes() == refs[0].slots.tobytes()
    assert not want.imag.any() and rel_err(c.to_numpy(), want.real) <= TOL


@pytest.mark.parametrize("case", range(len(GEMM_CASES)))
@pytest.mark.parametrize("real_first", [True, False])
def test_mixed_real_complex_gemm_path_vs_oracle(ctx, case, real_first):
    """f64 x Complex<f64> GEMM-shaped contractions run on the DMMA pipe too: the complex operand (and the result) is a
    real matrix with twice the columns / rows, so one real GEMM over doubled gather tables does it.  The reference
    upgrades the real operand to (x, 0.0) and multiplies complex numbers (upgrading_arithmetic.rs:1285-1318): same
    values to 1e-12."""
    rng = np.random.default_rng(700 + case)
    a, b = (O.slots(*x) for x in GEMM_CASES[case])
    ca, cb = O.order_structure(a)[0], O.order_structure(b)[0]
    n_points = 2
    shape_a, shape_b = [n_points] + [int(d) for d in ca["dim"]], [n_points] + [int(d) for d in cb["dim"]]
    da = rng.uniform(-1, 1, size=shape_a) + (0 if real_first else 1j * rng.uniform(-1, 1, size=shape_a))
    db = rng.uniform(-1, 1, size=shape_b) + (1j * rng.uniform(-1, 1, size=shape_b) if real_first else 0)
    ta = sp.BatchTensor.empty(ctx, ca, n_points, sp.F64 if real_first else sp.C64)
    ta.upload(np.ascontiguousarray(da.real if real_first else da).reshape(n_points, -1))
    tb = sp.BatchTensor.empty(ctx, cb, n_points, sp.C64 if real_first else sp.F64)
    tb.upload(np.ascontiguousarray(db if real_first else db.real).reshape(n_points, -1))
    launches = ctx.launch_count
    c = ta.contract(tb)
    assert c.dtype == sp.C64 and ctx.launch_count == launches + 1
    refs = [O.OTensor.dense(ca, da[p].astype(complex), canonicalize=False).contract(
        O.OTensor.dense(cb, db[p].astype(complex), canonicalize=False)) for p in range(n_points)]
    want = np.stack([r.to_dense() for r in refs])
    assert c.slots.tobytes() == refs[0].slots.tobytes()
    assert rel_err(c.to_numpy(), want) <= TOL


def test_trace_of_shared_tensors_keeps_storage(ctx):
    """Trace::internal_contract on one-point tensors: dense stays dense, sparse stays sparse with exact zeros dropped
    (contraction/trace.rs:12-83); values and the stored index set equal the oracle's."""
    rng = np.random.default_rng(23)
    raw = O.slots(("mink", 4, 5), ("euc", 3, 1), ("mink", 4, 5))
    shape = [4, 3, 4]
    ent = {(0, 0, 0): 2.0, (1, 0, 1): 2.0, (3, 2, 3): 1.5 - 1j, (2, 1, 0): 7.0, (1, 1, 1): 0.5j, (0, 1, 0): 0.5j}
    dense = np.zeros(shape, dtype=complex)
    for k, v in ent.items():
        dense[k] = v
    # holder structure whose canonical order equals the written order (self-dual rep ids 0 < 1 < 2), so that the data
    # stays row major over `raw`
    holder_slots = np.array([(1, 0, i, d, 900 + i) for i, d in enumerate(shape)], dtype=sp.SLOT_DTYPE)
    hs = sp.SparseTensor.from_entries(ctx, holder_slots, ent)
    hd = sp.DenseTensor.from_data(ctx, holder_slots, dense)
    ts, td = hs.trace(raw), hd.trace(raw)
    h = O.lib().orc_tensor_dense_raw(O._ptr(raw), len(raw), O._ptr(np.ascontiguousarray(dense).view(np.float64)))
    want = O.OTensor(h).trace().to_dense()
    assert ts.is_sparse and not td.is_sparse
    assert (td.to_numpy() == want).all() and (ts.to_numpy() == want).all()
    # euc index 0: 2 - 2 = 0 exactly -> dropped; index 1: 0.5i - 0.5i = 0 -> dropped; index 2: -(1.5 - i)
    assert set(ts.entries()) == {2} and ts.entries()[2] == -(1.5 - 1j)


def test_special_values_follow_ieee_like_the_reference(ctx):
    """NaN, +-inf, signed zeros and denormals: the pairwise kernels use the reference's operation order with IEEE ops
    (no FMA, no fast-math), so special values propagate exactly as on the CPU — compared bit for bit (any NaN == NaN)."""
    sl_a = O.order_structure(O.slots(("bis", 4, 1), ("mink", 4, 2)))[0]
    sl_b = O.order_structure(O.slots(("mink", 4, 2), ("bis", 4, 3)))[0]
    rng = np.random.default_rng(41)
    specials = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, -2.2e-308, 1.7e308, 1.0, -1.0])
    for n_points in (3, 2048):          # generic kernel and specialised kernel
        A = rng.choice(specials, size=(n_points, 16)) + 1j * rng.choice(specials, size=(n_points, 16))
        B = rng.choice(specials, size=(n_points, 16)) + 1j * rng.choice(specials, size=(n_points, 16))
        keep = rng.random((n_points, 16)) < 0.6       # mostly ordinary numbers, some special
        A = np.where(keep, _cases.random_complex(rng, (n_points, 16)), A)
        B = np.where(keep, _cases.random_complex(rng, (n_points, 16)), B)
        ta = sp.BatchTensor.empty(ctx, sl_a, n_points)
        ta.upload(A)
        tb = sp.BatchTensor.empty(ctx, sl_b, n_points)
        tb.upload(B)
        with np.errstate(all="ignore"):
            got = ta.contract(tb).to_numpy().reshape(n_points, -1)
            sample = range(n_points) if n_points <= 8 else rng.choice(n_points, 64, replace=False)
            for p in sample:
                want = O.OTensor.dense(sl_a, A[p], canonicalize=False).contract(
                    O.OTensor.dense(sl_b, B[p], canonicalize=False)).to_dense().reshape(-1)
                for part in ("real", "imag"):
                    g, w = getattr(got[p], part), getattr(want, part)
                    assert (np.isnan(g) == np.isnan(w)).all()
                    ok = ~np.isnan(w)
                    assert (g[ok].view(np.uint64) == w[ok].view(np.uint64)).all()      # incl. the sign of zeros and infinities


def test_trace_rejects_a_slot_list_that_is_not_the_tensors(ctx):
    """spn_tensor_trace computes strides from the caller's slot list: a list of another order or volume would make
    the kernel read out of bounds, so it is refused (Invalid / DimensionError)."""
    m = lambda a: sp.Minkowski.new_slot(4, a)
    t = sp.BatchTensor.from_numpy(ctx, [m(1), m(2), sp.Bispinor.new_slot(4, 3)], np.ones((3, 4, 4, 4)))
    with pytest.raises(sp.SpensoError) as e:
        t.trace([m(1), m(1)])                                   # order 2 for an order-3 tensor
    assert e.value.code == 7
    with pytest.raises(sp.SpensoError) as e:
        t.trace([m(1), m(1), sp.Bispinor.new_slot(8, 3)])       # right order, wrong volume
    assert e.value.code == 3
    with pytest.raises(sp.SpensoError):
        t.trace([])


# ---- host-driven contraction order and plan interchange on the device ------------------------------------------------
@pytest.mark.parametrize("name,scale,real", [("cfg1", 1.0, True), ("cfg2", 1.0, False), ("cfg3", 1.0, False),
                                             ("cfg5", 100.0, True)])
def test_host_schedule_executes_like_the_oracle(ctx, name, scale, real):
    """The contraction order comes from the host (here: the oracle's restatement of SmallestDegree, i.e. what the Rust
    side would send through spn_net_set_schedule).  Stepwise execution then runs the SAME pairwise contractions in the
    SAME order with the reference's rounding, so where no scalar factor is folded in between the results are
    bit-identical to the oracle's per-point evaluation; the fused kernel agrees to 1e-12."""
    spec = workloads.WORKLOADS[name]()
    n = 257
    arrays = workloads.synthetic_inputs(spec, n, 77, scale=scale, real=real)
    sched = O.oracle_schedule(spec, arrays)
    ref, _, _ = O.eval_spec(spec, arrays)
    for mode in (sp.STEPWISE, sp.FUSED):
        net = workloads.build_engine(spec, ctx, mode, schedule=sched)
        inputs = []
        for node, a in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
            t = sp.BatchTensor.empty(ctx, node[1], n)
            t.upload(a)
            inputs.append(t)
        got = net.evaluate(inputs).to_numpy().reshape(n, -1)
        assert rel_err(got, ref) <= TOL
        if mode == sp.STEPWISE and name in ("cfg1", "cfg2"):
            assert np.array_equal(got, ref), np.abs(got - ref).max()


@pytest.mark.parametrize("mode", [sp.FUSED, sp.STEPWISE])
def test_network_from_json_executes(ctx, mode):
    """Plan interchange: a network rebuilt from its JSON text — graph, library tensors, host schedule —
    gives the same bits as the network it was written from."""
    spec = workloads.one_loop_massive()
    n = 500
    arrays = workloads.synthetic_inputs(spec, n, 5, scale=100.0, real=True)
    sched = O.oracle_schedule(spec, arrays)
    a = workloads.build_engine(spec, ctx, mode, schedule=sched)
    text = a.to_json()                                  # what a host would emit: graph + its contraction order
    b = sp.Network.from_json(ctx, text).compile(mode)
    assert b.schedule == a.schedule and b.result_size == a.result_size
    inputs = []
    for node, arr in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
        t = sp.BatchTensor.empty(ctx, node[1], n)
        t.upload(arr)
        inputs.append(t)
    ra = a.evaluate(inputs).to_numpy()
    rb = b.evaluate(inputs).to_numpy()
    assert np.array_equal(ra, rb)
    ref, _, _ = O.eval_spec(spec, arrays)
    assert rel_err(rb.reshape(n, -1), ref) <= TOL


def test_plan_cache_serves_repeated_contractions(ctx):
    """spn_tensor_contract derives merge plan, tables and kernel once per pair of structures: a repeated a.contract(b)
    is a cache hit and returns the same bits; another structure is a miss."""
    import ctypes as C
    from spenso_b200._capi import lib

    def stats():
        h, m, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().spn_ctx_plan_cache_stats(ctx.h, C.byref(h), C.byref(m), C.byref(e))
        return h.value, m.value, e.value

    rng = np.random.default_rng(3)
    for n_points in (100, 5000):                      # generic kernel / specialised kernel
        a, _, da = dense_pair(ctx, rng, O.slots(("bis", 4, 1), ("bis", 4, 2)), n_points)
        b, _, db = dense_pair(ctx, rng, O.slots(("bis", 4, 2), ("bis", 4, 3)), n_points)
        h0, m0, _ = stats()
        r1 = a.contract(b).to_numpy()
        h1, m1, _ = stats()
        r2 = a.contract(b).to_numpy()
        h2, m2, _ = stats()
        assert m1 == m0 + 1 and h2 == h1 + 1 and m2 == m1
        assert np.array_equal(r1, r2)
        want = np.einsum("pij,pjk->pik", da, db)
        assert rel_err(r1, want) <= TOL


def test_data_tensor_conversions_and_element_access(ctx):
    """DataTensor::to_dense / to_sparse (tensors/data/sparse.rs:688-708, dense.rs:523-535), get / set by flat index,
    clone, reciprocal, and library tensors with concrete indices (get_lib_data, network/graph.rs:291-322)."""
    e = [sp.Euclidean.new_slot(2, 1), sp.Euclidean.new_slot(3, 2)]     # already canonical: flat = 3 * i + j
    d = sp.DenseTensor.from_data(ctx, e, [1, 0, 0, 2j, 0, 3])
    s = d.to_sparse()
    assert s.is_sparse and s.entries() == {0: 1, 3: 2j, 5: 3}              # zeros are not stored
    back = s.to_dense()
    assert not back.is_sparse and np.array_equal(back.to_numpy(), d.to_numpy())
    assert s.get(3) == 2j and d.get(1) == 0
    with pytest.raises(sp.SpensoError) as err:
        s.get(1)                                                             # no stored entry there
    assert err.value.code == 1
    s.set(1, 0.0)                                                            # SparseTensor::set stores what it is given
    assert s.entries()[1] == 0 and s.get(1) == 0
    with pytest.raises(sp.SpensoError):
        d.get(6)
    # a set changes what the tensor contracts to (the plan cache keys shared tensors by identity + revision)
    v = sp.DenseTensor.from_data(ctx, [sp.Euclidean.new_slot(3, 2)], [1, 1, 1])
    before = d.contract(v).to_numpy()
    d.set(1, 5.0)
    after = d.contract(v).to_numpy()
    assert np.array_equal(before.reshape(-1), [1, 3 + 2j]) and np.array_equal(after.reshape(-1), [6, 3 + 2j])
    # batched tensors: element access at a point, clone is a deep copy, to_sparse is refused
    rng = np.random.default_rng(0)
    for layout in (sp.POINT_MAJOR, sp.COMPONENT_MAJOR):
        t, _, data = dense_pair(ctx, rng, O.slots(("bis", 4, 1), ("mink", 4, 2)), 37, layout)
        assert t.get(7, point=11) == data[11].reshape(-1)[7]
        c = t.clone()
        t.set(7, 1 + 2j, point=11)
        assert t.get(7, point=11) == 1 + 2j and c.get(7, point=11) == data[11].reshape(-1)[7]
        assert np.array_equal(t.to_dense().to_numpy(), t.to_numpy())
        with pytest.raises(sp.SpensoError):
            t.to_sparse()
    r = sp.BatchTensor.from_numpy(ctx, [], np.array([2.0, 4.0, 1j]).reshape(3)).recip().to_numpy()
    assert np.allclose(r, [0.5, 0.25, -1j], rtol=0, atol=0)
    # library tensor with concrete indices == the oracle's instantiation
    b = lambda a: sp.Bispinor.new_slot(4, a)
    g = sp.library_tensor(ctx, sp.LIB_GAMMA_WEYL, [b(2), b(1), sp.Minkowski.new_slot(4, 3)])
    og = O.OLib.builtin(0).instantiate(O.slots(("bis", 4, 2), ("bis", 4, 1), ("mink", 4, 3)))
    assert g.is_sparse and g.entries() == {k: complex(v) for k, v in og.entries().items()}


@pytest.mark.parametrize("name,n_points,scale,real", [("cfg1", 777, 1.0, True), ("cfg2", 4099, 1.0, False), ("cfg3", 301, 1.0, False),
                                                      ("cfg5", 5003, 100.0, True), ("cfg5r", 5003, 100.0, True)])
@pytest.mark.parametrize("layout", [sp.POINT_MAJOR, sp.COMPONENT_MAJOR])
def test_sum_only_kernels_commute_the_linear_tail(ctx, name, n_points, scale, real, layout):
    """Monte-Carlo sum without a per-point store (the variant bench.py runs for cfg5): the kernel accumulates a basis
    of non-linear values — link by link where a multiply-add chain ends in one — and applies the network's trailing
    linear map once after the loop.  Same sums as the oracle's per-point results added up (1e-12), and as the variant
    that also stores every point."""
    import torch

    spec = workloads.WORKLOADS[name]()
    arrays = workloads.synthetic_inputs(spec, n_points, 4321, scale=scale, real=real)
    if name == "cfg5r":
        ref_arrays = arrays
        arrays = [np.ascontiguousarray(a.real) for a in arrays]
        ref, ref_sum, _ = O.eval_spec(workloads.one_loop_massive(), ref_arrays)
    else:
        ref, ref_sum, _ = O.eval_spec(spec, arrays)
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    inputs = []
    for node, a, dt in zip([x for x in spec.nodes if x[0] == "batched"], arrays, spec.input_dtypes):
        t = sp.BatchTensor.empty(ctx, node[1], n_points, dt, layout)
        t.upload(a)
        inputs.append(t)
    s_only = torch.zeros(2 * net.result_size, dtype=torch.float64, device="cuda:0")
    net.execute(inputs, None, sum_out_ptr=s_only.data_ptr())
    out = sp.BatchTensor.empty(ctx, net.result_slots, n_points, sp.C64, layout)
    s_both = torch.zeros_like(s_only)
    net.execute(inputs, out, sum_out_ptr=s_both.data_ptr())
    ctx.synchronize()
    a = s_only.cpu().numpy().view(np.complex128)
    b = s_both.cpu().numpy().view(np.complex128)
    assert rel_err(a, ref_sum) <= TOL
    assert rel_err(a, b) <= 1e-13
    assert rel_err(out.to_numpy().reshape(n_points, -1), ref) <= TOL
    key = ("p" if layout == sp.POINT_MAJOR else "c") * net.n_inputs
    assert net.variant_fp64_instr(key + "ns") <= net.variant_fp64_instr(key + ("p" if layout == sp.POINT_MAJOR else "c") + "s")


@pytest.mark.parametrize("name,scale,real", [("cfg2", 1.0, False), ("cfg5", 100.0, True)])
def test_execute_host_async_keeps_order_and_results(ctx, name, scale, real):
    """spn_net_execute_host_async: batches submitted back to back (the pipeline does not drain between calls) give,
    after spn_net_host_wait, what the synchronous calls give — per-point results and Monte-Carlo sums, several calls in
    flight, ragged chunking."""
    import torch

    spec = workloads.WORKLOADS[name]()
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    n = 10_007
    batches, pinned = [], []
    for k in range(5):
        arrs = workloads.synthetic_inputs(spec, n, 100 + k, scale=scale, real=real)
        pins = [torch.from_numpy(a.view(np.float64).reshape(n, -1)).pin_memory() for a in arrs]
        pinned.append(pins)
        batches.append([p.numpy().view(np.complex128).reshape(n, -1) for p in pins])
    size = net.result_size
    want_out, want_sum = [], []
    for b in batches:
        o = np.zeros((n, size), dtype=np.complex128)
        _, s = net.execute_host(b, out=o, want_sum=True, chunk_points=1500)
        want_out.append(o)
        want_sum.append(s.copy())
    outs = [torch.zeros(n * size * 2, dtype=torch.float64).pin_memory() for _ in batches]
    sums = [np.zeros(size, dtype=np.complex128) for _ in batches]
    for b, o, s in zip(batches, outs, sums):
        net.execute_host_async(b, out=o.numpy().view(np.complex128).reshape(n, size), sum_out=s, chunk_points=1500)
    net.host_wait()
    for o, s, wo, ws in zip(outs, sums, want_out, want_sum):
        assert np.array_equal(o.numpy().view(np.complex128).reshape(n, size), wo)
        assert np.array_equal(s, ws)
    ref, ref_sum, _ = O.eval_spec(spec, [b for b in batches[0]])
    assert rel_err(want_out[0], ref) <= TOL and rel_err(want_sum[0], ref_sum) <= TOL


@pytest.mark.parametrize("seed", range(4))
def test_slab_staged_kernels_equal_the_generic_kernels_bit_for_bit(ctx, seed):
    """The specialised pairwise kernels in their three memory paths — warp-owned cp.async slabs (default; also launched
    as a persistent grid, SPN_PAIR_PERSISTENT), direct thread = point loads (SPN_PAIR_NO_SLAB) and TMA bulk copies
    (SPN_LOAD_MODE=tma) — against the generic kernel
    (SPN_NO_PAIR_JIT) on the same operands: identical bits, for a ragged point count (a partial last slab), c64 and f64
    operands, both layouts, dense and shared-sparse partners."""
    import os

    rng = np.random.default_rng(9000 + seed)
    n_points, done = 4099, 0
    for it in range(8):
        a, b = _cases.random_pair(rng, max_free=2, max_common=2)
        layout = sp.POINT_MAJOR if it % 3 != 2 else sp.COMPONENT_MAJOR
        ca = O.order_structure(a)[0] if len(a) else a
        cb = O.order_structure(b)[0] if len(b) else b
        dta = sp.F64 if it % 4 == 1 else sp.C64
        dtb = sp.F64 if it % 4 in (1, 3) else sp.C64

        def mk(c, dt):
            shape = [n_points] + [int(d) for d in c["dim"]]
            data = rng.uniform(-1, 1, size=shape) if dt == sp.F64 else _cases.random_complex(rng, shape)
            t = sp.BatchTensor.empty(ctx, c, n_points, dt, layout)
            t.upload(np.ascontiguousarray(data).reshape(n_points, -1))
            return t

        ta = mk(ca, dta)
        if it % 5 == 4:
            ent = _cases.random_sparse_entries(rng, [int(d) for d in cb["dim"]])
            tb = sp.SparseTensor.from_entries(ctx, cb, ent)
        else:
            tb = mk(cb, dtb)
        results = {}
        for name, env in (("generic", {"SPN_NO_PAIR_JIT": "1"}), ("slab", {}), ("direct", {"SPN_PAIR_NO_SLAB": "1"}),
                          ("tma", {"SPN_LOAD_MODE": "tma"}), ("persistent", {"SPN_PAIR_PERSISTENT": "1"})):
            for k in ("SPN_NO_PAIR_JIT", "SPN_PAIR_NO_SLAB", "SPN_LOAD_MODE", "SPN_PAIR_PERSISTENT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            try:
                results[name] = ta.contract(tb).to_numpy()
            except sp.SpensoError:
                results = None
                break
            finally:
                for k in env:
                    os.environ.pop(k, None)
        if results is None:
            continue
        for name in ("slab", "direct", "tma", "persistent"):
            assert np.array_equal(results[name], results["generic"]), f"case {it}: {name} differs from the generic kernel"
        done += 1
    assert done >= 4


def test_host_memory_registered_in_place(ctx):
    """spn_host_register pins memory the host already owns (what a Rust Vec or a numpy array is) so that
    spn_net_execute_host streams from / to it without a pinned staging copy: same results as from unpinned memory."""
    from spenso_b200._capi import check, lib

    spec = workloads.eemumu()
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    n = 20_000
    arrs = workloads.synthetic_inputs(spec, n, 3)
    out_a = np.zeros((n, 1), dtype=np.complex128)
    net.execute_host(arrs, out=out_a)                      # pageable memory: works, synchronously
    out_b = np.zeros((n, 1), dtype=np.complex128)
    regs = arrs + [out_b]
    for a in regs:
        check(lib().spn_host_register(a.ctypes.data, a.nbytes))
    try:
        net.execute_host_async(arrs, out=out_b)
        net.host_wait()
    finally:
        for a in regs:
            check(lib().spn_host_unregister(a.ctypes.data))
    assert np.array_equal(out_a, out_b)
    ref, _, _ = O.eval_spec(spec, arrs)
    assert rel_err(out_b, ref) <= TOL
    assert lib().spn_host_register(None, 16) != 0


def test_relayout_round_trip_and_network_on_converted_inputs(ctx):
    """spn_tensor_relayout: point-major <-> component-major on the device; the data are the same tensor (to_numpy is
    layout independent), a round trip is the identity, and a network evaluated on converted inputs gives the same bits."""
    rng = np.random.default_rng(5)
    t, _, data = dense_pair(ctx, rng, O.slots(("coad", 8, 1), ("cof", 3, 2), ("coaf", 3, 3)), 1000)
    c = t.relayout(sp.COMPONENT_MAJOR)
    assert c.layout == sp.COMPONENT_MAJOR and np.array_equal(c.to_numpy(), t.to_numpy())
    assert np.array_equal(c.relayout(sp.POINT_MAJOR).to_numpy(), data)
    r = sp.BatchTensor.from_numpy(ctx, [sp.Minkowski.new_slot(4, 1)], rng.uniform(-1, 1, (77, 4)), dtype=sp.F64)
    assert np.array_equal(r.relayout(sp.COMPONENT_MAJOR).relayout(sp.POINT_MAJOR).to_numpy(), r.to_numpy())
    spec = workloads.su3_colour()
    n = 513
    arrays = workloads.synthetic_inputs(spec, n, 8)
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    ins = []
    for node, a in zip([x for x in spec.nodes if x[0] == "batched"], arrays):
        x = sp.BatchTensor.empty(ctx, node[1], n)
        x.upload(a)
        ins.append(x)
    out_p = net.evaluate(ins).to_numpy()
    out_c = net.evaluate([x.relayout(sp.COMPONENT_MAJOR) for x in ins], layout=sp.COMPONENT_MAJOR).to_numpy()
    assert np.array_equal(out_p, out_c)
    with pytest.raises(sp.SpensoError):
        sp.DenseTensor.from_data(ctx, [sp.Euclidean.new_slot(2, 1)], [1, 2]).relayout(sp.COMPONENT_MAJOR)


def test_execute_host_on_a_stepwise_network(ctx):
    """Network::execute on host data also works for networks compiled STEPWISE (one pairwise kernel per contraction):
    chunked, same results as the fused executor to 1e-12 and as the oracle."""
    spec = workloads.eemumu()
    n = 5_003
    arrays = workloads.synthetic_inputs(spec, n, 21)
    ref, ref_sum, _ = O.eval_spec(spec, arrays)
    step = workloads.build_engine(spec, ctx, sp.STEPWISE)
    out = np.zeros((n, 1), dtype=np.complex128)
    _, s = step.execute_host(arrays, out=out, want_sum=True, chunk_points=1024)
    assert rel_err(out, ref) <= TOL and rel_err(s, ref_sum) <= TOL
    out2 = np.zeros((n, 1), dtype=np.complex128)
    step.execute_host_async(arrays, out=out2, chunk_points=700)
    step.host_wait()
    assert np.array_equal(out, out2)


@pytest.mark.parametrize("name,n_points,scale,real", [("cfg2", 1000, 1.0, False), ("cfg2", 300_001, 1.0, False),
                                                      ("cfg5", 70_003, 100.0, True), ("cfg5r", 70_003, 100.0, True)])
def test_fused_tma_loader_equals_the_default_loader(ctx, monkeypatch, name, n_points, scale, real):
    """SPN_LOAD_MODE=tma swaps the fused kernel's per-thread 256-bit loads for slabs moved by the TMA engine
    (cp.async.bulk + mbarrier, 2-3 stages).  Only the loads change, so per-point results are bit-for-bit those of the
    default kernel — on a ragged last slab and over many slabs per CTA — and the sums agree with the oracle."""
    import torch

    spec = workloads.WORKLOADS[name]()
    arrays = workloads.synthetic_inputs(spec, n_points, 777, scale=scale, real=real)
    if name == "cfg5r":
        arrays = [np.ascontiguousarray(a.real) for a in arrays]

    def run():
        net = workloads.build_engine(spec, ctx, sp.FUSED)
        inputs = []
        for node, a, dt in zip([x for x in spec.nodes if x[0] == "batched"], arrays, spec.input_dtypes):
            t = sp.BatchTensor.empty(ctx, node[1], n_points, dt)
            t.upload(a)
            inputs.append(t)
        out = sp.BatchTensor.empty(ctx, net.result_slots, n_points)
        s = torch.zeros(2 * net.result_size, dtype=torch.float64, device="cuda:0")
        net.execute(inputs, out)
        net.execute(inputs, None, sum_out_ptr=s.data_ptr())
        ctx.synchronize()
        src = net.prepare_variant("p" * net.n_inputs + "p")
        return out.to_numpy().reshape(n_points, -1), s.cpu().numpy().view(np.complex128), src

    want, want_sum, src0 = run()
    monkeypatch.setenv("SPN_LOAD_MODE", "tma")
    got, got_sum, src1 = run()
    assert "cp.async.bulk" in src1 and "mbarrier.try_wait" in src1 and "cp.async.bulk" not in src0
    assert np.array_equal(got.view(np.float64), want.view(np.float64))
    assert rel_err(got_sum, want_sum) <= 1e-12


@pytest.mark.parametrize("seed", range(3))
def test_sparse_times_small_batched_kernels_bit_for_bit(ctx, seed):
    """Shared sparse x small batched operand (gamma^mu p_mu and relatives — a write stream: the result is larger than the
    batched operand).  The specialised kernel — hardware-scheduled single-warp CTAs by default, the persistent grid of
    SPN_PAIR_PERSISTENT=1 — gives bit-for-bit the generic kernel's results (SPN_NO_PAIR_JIT) and the oracle's: with the
    sparse tensor on either side, c64 and f64 partners, 2..128 result elements, 1-4 stored products per element, metric
    signs, and a ragged last slab."""
    import os

    rng = np.random.default_rng(4200 + seed)
    b = lambda a: sp.Bispinor.new_slot(4, a)
    m = lambda a: sp.Minkowski.new_slot(4, a)
    e3 = lambda a: sp.Euclidean.new_slot(2, a)
    gamma = {(0, 2, 0): 1, (1, 3, 0): 1, (2, 0, 0): 1, (3, 1, 0): 1, (0, 3, 1): 1, (1, 2, 1): 1, (2, 1, 1): -1, (3, 0, 1): -1,
             (0, 3, 2): -1j, (1, 2, 2): 1j, (2, 1, 2): 1j, (3, 0, 2): -1j, (0, 2, 3): 1, (1, 3, 3): -1, (2, 0, 3): -1, (3, 1, 3): 1}
    n_points = 4099

    def random_entries(dims, fill):
        ent = {}
        for idx in np.ndindex(*dims):
            if rng.random() < fill:
                ent[tuple(int(i) for i in idx)] = complex(rng.uniform(-1, 1), rng.uniform(-1, 1))
        return ent or {tuple(0 for _ in dims): 1.0}

    cases = [
        ([b(1), b(2), m(3)], gamma, [m(3)]),                                       # 16 elements, 2 products each or none
        ([b(1), b(2), m(3)], random_entries((4, 4, 4), 0.7), [m(3)]),                # up to 4 products (Minkowski signs)
        ([b(1), m(3)], random_entries((4, 4), 0.5), [m(3)]),                         # 4 elements
        ([e3(1), m(3)], random_entries((2, 4), 0.6), [m(3)]),                        # 2 elements
        ([b(1), b(2), m(3), m(4)], random_entries((4, 4, 4, 4), 0.15), [m(4)]),      # 64 elements
        ([b(1), b(2), m(3)], random_entries((4, 4, 4), 0.3), [b(2), m(3), e3(5)]),   # 2 contracted indices, 8 elements
        ([b(1), b(2)], random_entries((4, 4), 0.4), [b(5), e3(6)]),                  # exterior product, 128 elements
    ]
    for it, (ssl, ent, dsl) in enumerate(cases):
        for dt in (sp.C64, sp.F64):
            for sparse_first in (True, False):
                cd = O.order_structure(sp.slots(dsl))[0]
                shape = [n_points] + [int(d) for d in cd["dim"]]
                data = rng.uniform(-1, 1, size=shape) if dt == sp.F64 else _cases.random_complex(rng, shape)
                td = sp.BatchTensor.empty(ctx, cd, n_points, dt)
                td.upload(np.ascontiguousarray(data).reshape(n_points, -1))
                ts = sp.SparseTensor.from_entries(ctx, ssl, ent)
                assert sp.contract_cuda_source(ssl, cd, sparse_entries=ent, dtype_b=dt, sparse_first=sparse_first)
                results = {}
                for name, env in (("generic", {"SPN_NO_PAIR_JIT": "1"}), ("flat", {}), ("persistent", {"SPN_PAIR_PERSISTENT": "1"})):
                    for k in ("SPN_NO_PAIR_JIT", "SPN_PAIR_PERSISTENT"):
                        os.environ.pop(k, None)
                    os.environ.update(env)
                    try:
                        r = ts.contract(td) if sparse_first else td.contract(ts)
                        results[name] = r.to_numpy()
                    finally:
                        for k in env:
                            os.environ.pop(k, None)
                assert np.array_equal(results["flat"], results["generic"]), f"case {it}: specialised kernel differs from generic"
                assert np.array_equal(results["persistent"], results["generic"]), f"case {it}: persistent launch differs"
                if it < 2 and dt == sp.C64:   # and the oracle, on the physics case
                    oa = O.OTensor.sparse(sp.slots(ssl), ent)
                    pts = range(0, n_points, 997)
                    od = [O.OTensor.dense(cd, data[p], canonicalize=False) for p in pts]
                    want = np.stack([(oa.contract(x) if sparse_first else x.contract(oa)).to_dense() for x in od])
                    assert np.array_equal(results["flat"].reshape(n_points, -1)[::997], want.reshape(len(want), -1))


@pytest.mark.parametrize("seed", range(2))
def test_row_parallel_kernels_equal_the_generic_kernels_bit_for_bit(ctx, seed):
    """Per-point tensors too large for one thread's registers (colour algebra: dense[coad^3] x dense[coad^2], x sparse f)
    take the row-parallel specialised kernel: thread = (point, row of one batched operand).  Same sums as the reference
    (k order, from zero, no FMA), so bit-for-bit the generic / CSR kernels (SPN_PAIR_NO_ROWS=1) — rows taken from either
    operand, f64 and c64, Minkowski signs over two contracted indices, interleaved result order, a shared dense or a
    shared sparse partner — and the oracle."""
    import os

    rng = np.random.default_rng(7100 + seed)
    c = lambda a: sp.ColorAdjoint.new_slot(8, a)
    m = lambda a: sp.Minkowski.new_slot(4, a)
    b = lambda a: sp.Bispinor.new_slot(4, a)
    n_points = 2300

    def batch(sl, dt):
        canon = O.order_structure(sp.slots(sl))[0]
        shape = [n_points] + [int(d) for d in canon["dim"]]
        data = rng.uniform(-1, 1, size=shape) if dt == sp.F64 else _cases.random_complex(rng, shape)
        t = sp.BatchTensor.empty(ctx, canon, n_points, dt)
        t.upload(np.ascontiguousarray(data).reshape(n_points, -1))
        return t, canon, data

    def sparse(sl, fill):
        dims = [s.dim for s in sl]
        ent = {}
        for idx in np.ndindex(*dims):
            if rng.random() < fill:
                ent[tuple(int(i) for i in idx)] = complex(rng.uniform(-1, 1), rng.uniform(-1, 1))
        return sp.SparseTensor.from_entries(ctx, sl, ent), ent

    def shared_dense(sl):
        dims = [s.dim for s in sl]
        data = _cases.random_complex(rng, dims)
        return sp.DenseTensor.from_data(ctx, sl, data), data

    cases = []   # (a, b, oracle pair or None, expects the row-parallel source)
    A, ca, da = batch([c(1), c(2), c(3)], sp.C64)
    B, cb, db = batch([c(3), c(4)], sp.C64)
    cases.append((A, B, (ca, da, cb, db), sp.contract_cuda_source(ca, cb)))
    cases.append((B, A, (cb, db, ca, da), sp.contract_cuda_source(cb, ca)))
    Ar, car, _ = batch([c(1), c(2), c(3)], sp.F64)
    cases.append((Ar, B, None, sp.contract_cuda_source(car, cb, dtype_a=sp.F64)))
    Br, cbr, _ = batch([c(3), c(4)], sp.F64)
    cases.append((Ar, Br, None, sp.contract_cuda_source(car, cbr, dtype_a=sp.F64, dtype_b=sp.F64)))
    F, fent = sparse([c(2), c(3), c(4)], 0.1)
    cases.append((A, F, None, sp.contract_cuda_source([c(2), c(3), c(4)], ca, sparse_entries=fent, sparse_first=False)))
    cases.append((F, A, None, sp.contract_cuda_source([c(2), c(3), c(4)], ca, sparse_entries=fent, sparse_first=True)))
    M1, cm1, _ = batch([m(1), m(2), m(3), b(4), b(6)], sp.C64)
    M2, cm2, _ = batch([m(2), m(3), b(5)], sp.C64)
    cases.append((M1, M2, None, sp.contract_cuda_source(cm1, cm2)))
    I1, ci1, _ = batch([c(1), c(3), c(5)], sp.C64)
    I2, ci2, _ = batch([c(2), c(3), c(4)], sp.C64)
    cases.append((I1, I2, None, sp.contract_cuda_source(ci1, ci2)))
    A4, ca4, _ = batch([c(1), c(2), c(3), c(5)], sp.C64)          # 512 rows: one point per 256-thread CTA
    cases.append((A4, B, None, sp.contract_cuda_source(ca4, cb)))
    A5, ca5, _ = batch([sp.Euclidean.new_slot(3, 7), c(2), c(3), c(5), c(6)], sp.C64)   # 1 536 rows: three passes of 512
    B5, cb5, _ = batch([c(3), c(4)], sp.C64)
    src5 = sp.contract_cuda_source(ca5, cb5)
    assert "passes over the rows" in src5
    cases.append((A5, B5, None, src5))
    # rows too long for registers, few columns: streamed over k (full contractions to a scalar / a vector)
    A2, ca2, da2 = batch([c(1), c(2), c(3)], sp.C64)
    cases.append((A, A2, (ca, da, ca2, da2), sp.contract_cuda_source(ca, ca2)))                 # 512 products -> scalar
    A6, ca6, _ = batch([c(1), c(2), c(3), c(4)], sp.C64)
    B6, cb6, _ = batch([c(2), c(3), c(4)], sp.F64)
    cases.append((A6, B6, None, sp.contract_cuda_source(ca6, cb6, dtype_b=sp.F64)))               # 8 rows x 512, c64 x f64
    cases.append((B6, A6, None, sp.contract_cuda_source(cb6, ca6, dtype_a=sp.F64)))
    M7, cm7, _ = batch([m(1), m(2), m(3), m(4), b(5)], sp.C64)
    M8, cm8, _ = batch([m(1), m(2), m(3), m(4)], sp.C64)
    src78 = sp.contract_cuda_source(cm7, cm8)
    assert "kTabNeg" in src78
    cases.append((M7, M8, None, src78))                                                           # 256 signed products per row
    S, _ = shared_dense([c(3), c(4)])
    cases.append((A, S, None, None))
    cases.append((S, A, None, None))
    for it, (ta, tb, orc, src) in enumerate(cases):
        if src is not None:
            assert "thread = (point, row of" in src, f"case {it} does not take the row-parallel kernel"
        results = {}
        for name, env in (("rows", {}), ("generic", {"SPN_PAIR_NO_ROWS": "1"})):
            os.environ.pop("SPN_PAIR_NO_ROWS", None)
            os.environ.update(env)
            try:
                results[name] = ta.contract(tb).to_numpy()
            finally:
                for k in env:
                    os.environ.pop(k, None)
        assert np.array_equal(results["rows"], results["generic"]), f"case {it}: row-parallel kernel differs from the generic one"
        if orc is not None:
            (c1, d1, c2, d2) = orc
            pts = range(0, n_points, 571)
            want = np.stack([O.OTensor.dense(c1, d1[p], canonicalize=False).contract(O.OTensor.dense(c2, d2[p], canonicalize=False)).to_dense()
                             for p in pts])
            assert np.array_equal(results["rows"].reshape(n_points, -1)[::571], want.reshape(len(want), -1))


@pytest.mark.parametrize("seed", range(3))
def test_specialised_kernels_on_random_medium_pairs_bit_for_bit(ctx, seed):
    """Random pairs of up to 6 indices per tensor (up to 4 096 components per point): whichever specialised form the
    dispatcher picks — registers, warp slabs, row-parallel — gives the bits of the generic kernels (SPN_NO_PAIR_JIT)."""
    import os

    rng = np.random.default_rng(8800 + seed)
    n_points, done, rows_seen = 2111, 0, 0
    for it in range(40):
        a, b = _cases.random_pair(rng, max_free=4, max_common=2)
        ca = O.order_structure(a)[0] if len(a) else a
        cb = O.order_structure(b)[0] if len(b) else b
        size_a = int(np.prod([int(d) for d in ca["dim"]])) if len(ca) else 1
        size_b = int(np.prod([int(d) for d in cb["dim"]])) if len(cb) else 1
        if size_a * size_b > 1 << 16 or max(size_a, size_b) < 48:
            continue
        dta = sp.F64 if it % 4 == 1 else sp.C64
        dtb = sp.F64 if it % 4 in (1, 3) else sp.C64

        def mk(c, dt):
            shape = [n_points] + [int(d) for d in c["dim"]]
            data = rng.uniform(-1, 1, size=shape) if dt == sp.F64 else _cases.random_complex(rng, shape)
            t = sp.BatchTensor.empty(ctx, c, n_points, dt)
            t.upload(np.ascontiguousarray(data).reshape(n_points, -1))
            return t

        ta = mk(ca, dta)
        ent = None
        if it % 5 == 4 and len(cb):
            ent = _cases.random_sparse_entries(rng, [int(d) for d in cb["dim"]], density=0.1)
            tb = sp.SparseTensor.from_entries(ctx, cb, ent)
        else:
            tb = mk(cb, dtb)
        try:
            src = (sp.contract_cuda_source(cb, ca, sparse_entries=ent, dtype_b=dta, sparse_first=False) if ent is not None
                   else sp.contract_cuda_source(ca, cb, dtype_a=dta, dtype_b=dtb))
        except sp.SpensoError:
            continue
        rows_seen += "thread = (point, row of" in src
        results = {}
        for name, env in (("generic", {"SPN_NO_PAIR_JIT": "1"}), ("specialised", {})):
            os.environ.pop("SPN_NO_PAIR_JIT", None)
            os.environ.update(env)
            try:
                results[name] = ta.contract(tb).to_numpy()
            except sp.SpensoError:
                results = None
                break
            finally:
                for k in env:
                    os.environ.pop(k, None)
        if results is None:
            continue
        assert np.array_equal(results["specialised"], results["generic"]), f"case {it} differs from the generic kernel"
        done += 1
        if done >= 10:
            break
    assert done >= 5 and rows_seen >= 2


@pytest.mark.parametrize("name,n_points", [("cfg2", 300_001), ("cfg3", 20_011)])
def test_fused_flat_and_persistent_launches_agree(ctx, monkeypatch, name, n_points):
    """Streaming fused kernels that store a result per point are launched flat (64-thread CTAs, one point per thread);
    SPN_NET_PERSISTENT=1 keeps the persistent wave of 256-thread CTAs.  Same arithmetic per point: identical bits, also
    on a point range that starts inside the batch and ends on a ragged CTA."""
    spec = workloads.WORKLOADS[name]()
    arrays = workloads.synthetic_inputs(spec, n_points, 4242)

    def run():
        net = workloads.build_engine(spec, ctx, sp.FUSED)
        inputs = []
        for node, a, dt in zip([x for x in spec.nodes if x[0] == "batched"], arrays, spec.input_dtypes):
            t = sp.BatchTensor.empty(ctx, node[1], n_points, dt)
            t.upload(a)
            inputs.append(t)
        out = sp.BatchTensor.empty(ctx, net.result_slots, n_points)
        out.upload(np.zeros((n_points, net.result_size), dtype=complex))
        net.execute(inputs, out, first_point=37, n_points=n_points - 100)
        ctx.synchronize()
        return out.to_numpy().reshape(n_points, -1), net.prepare_variant("p" * net.n_inputs + "p")

    flat, src_flat = run()
    monkeypatch.setenv("SPN_NET_PERSISTENT", "1")
    pers, src_pers = run()
    assert "__launch_bounds__(64" in src_flat and "__launch_bounds__(256" in src_pers
    assert np.array_equal(flat.view(np.float64), pers.view(np.float64))
    assert not flat[:37].any() and not flat[n_points - 63:].any() and flat[37:n_points - 63].any()
