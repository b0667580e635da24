"""The reference's OWN golden outputs (insta snapshots of its seeded tests) against the oracle and,
under `-m gpu`, against the device path through the C ABI.

tests/golden/reference_snapshots.json is a transcription of spenso/src/snapshots/spenso__tests__*.snap
(made by tests/golden/make_reference_snapshots.py); the inputs are regenerated with the restated Rust
RNG (tests/_rustrand.py), which is itself pinned by `rng_is_deterministic`.  These are integer-valued
tensors, exactly representable in f64, so every comparison is `==`.

The snapshots were written when the reference's test structure kept slots in the order given
(`VecStructure`); today's `OrderedStructure` sorts them (SURVEY 8c caveat), so values are compared
by index LABEL (aind), not by raw flat position.
"""
import json
import os

import numpy as np
import pytest

import _oracle as O
import _rustrand as R

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_snapshots.json")))
_KIND = {"SelfDual": O.REP_SELF_DUAL, "InlineMetric": O.REP_INLINE_METRIC, "Dualizable": O.REP_DUALIZABLE}


def oslots(tuples):
    return O.slots(*tuples)


def dense_of(entries: dict, shape):
    a = np.zeros(int(np.prod(shape)), dtype=np.int64)
    for k, v in entries.items():
        a[k] = v
    return a.reshape(shape)


def by_label(ainds, arr, want_order):
    """transpose `arr` (axes labelled `ainds`) into `want_order`."""
    ainds = [int(a) for a in ainds]
    return np.transpose(arr, [ainds.index(a) for a in want_order])


def snap_slots(sl):
    """(rep kind, rep id, dim, aind) tuples of an oracle/engine slot array."""
    return [(int(s["rep_kind"]), int(s["rep_id"]), int(s["dim"]), int(s["aind"])) for s in sl]


def gold_slots(lst):
    return [(_KIND[s["kind"]], s["id"], s["dim"], s["aind"]) for s in lst]


# ---- the RNG restatement is pinned by the reference's own determinism snapshot ---------------------------
def test_rng_is_deterministic_snapshot():
    # tests.rs:150-162
    a = R.test_structure(3, 11)
    size = int(np.prod([d for _, d, _ in a]))
    t = R.test_tensor(size, 1, "i8")
    assert [t[k] for k in sorted(t)] == GOLD["rng_is_deterministic"]


def test_seeded_structures_match_snapshots():
    # fibers.snap / single_fiber.snap hold test_structure(5, 5) in RON (tests.rs:174-241)
    got = [(_KIND["SelfDual" if r == "euc" else "InlineMetric"], 0, d, a) for r, d, a in R.test_structure(5, 5)]
    assert got == gold_slots(GOLD["fibers"]["structure"]) == gold_slots(GOLD["single_fiber"]["structure"])
    # scalar_and_dim1_conract snapshots 1-3 (tests.rs:385-393): "aind (repdim)" listings
    def listing(sl):
        name = {"euc": "euc", "mink": "mink", "lor": "lor", "lor_d": "lor🠓"}
        return [x for r, d, a in sl for x in (str(a), f"({name[r]}{d})")]
    g = GOLD["scalar_and_dim1"]
    assert listing(R.test_structure_with_dims([1, 3, 1, 2], 6)) == g["common"]
    assert listing(R.test_structure(1, 32)) == g["structa"]
    assert listing(R.test_structure(1, 22)) == g["structb"]


def _single_contract_structs():
    # tests.rs:534-549
    s = 18
    common = R.test_structure_with_id([1], s)
    commondual = R.test_structure_with_id([-1], s)
    structa = R.test_structure_with_id([2], s)
    structb = R.test_structure_with_id([-3, 4], s)
    rng = R.Xoroshiro64Star(s)
    structa.insert(rng.gen_range_usize(0, len(structa)), common[0])
    structb.insert(rng.gen_range_usize(0, len(structb)), commondual[0])
    return structa, structb


def test_single_contract_structure_snapshots():
    structa, structb = _single_contract_structs()
    g = GOLD["single_contract"]
    # structa.sort() in the test: the oracle's canonical order must reproduce the snapshot order
    canon_a = O.order_structure(oslots(structa))[0]
    assert snap_slots(canon_a) == gold_slots(g["structa"])
    assert snap_slots(oslots(structb)) == gold_slots(g["structb"])
    # result of densor_b.contract(&densor_a): same slot SET as the snapshot (legacy order = concatenation;
    # today's merge returns it canonically sorted) and the oracle's merge result is canonical
    m = O.merge(O.order_structure(oslots(structb))[0], canon_a)
    assert sorted(snap_slots(m["slots"])) == sorted(gold_slots(g["result"]))
    assert snap_slots(m["slots"]) == snap_slots(O.order_structure(m["slots"])[0])


# ---- fibre iterators (I1, I3) against fibers.snap / single_fiber.snap / fiber_from_structure.snap --------
def _class_major(dims, fiber_free):
    """FiberClass::from(fiber).iter(): classes enumerate the fibre's FIXED dims (conjugate iterator);
    each class yields the fibre shifted to the class' zero index (fiber.rs:522-628)."""
    classes, _ = O.flat_iter_dims(dims, fiber_free, mode=1)
    items = []
    for c in classes:
        idx, _ = O.flat_iter_dims(dims, fiber_free, mode=0, shift=int(c))
        items.extend(int(i) for i in idx)
    return [int(c) for c in classes], items


def test_fibers_snapshot():
    # tests.rs:229-241: Fiber::from_filter([T,T,F,F,T]) on test_structure(5,5)
    dims = [s["dim"] for s in GOLD["fibers"]["structure"]]
    classes, items = _class_major(dims, [1, 1, 0, 0, 1])
    assert classes == GOLD["fibers"]["fciter"]
    assert items == GOLD["fibers"]["fiter"]


def test_single_fiber_snapshot():
    # tests.rs:174-190: from_filter([T,F,F,F,F]); BoolFilter constructors compute is_single
    dims = [s["dim"] for s in GOLD["single_fiber"]["structure"]]
    free = [1, 0, 0, 0, 0]
    for use_single in (True, False):
        idx, _ = O.flat_iter_dims(dims, free, mode=0, use_single=use_single)
        assert [int(i) for i in idx] == GOLD["single_fiber"]["fciter"]
        classes, _ = O.flat_iter_dims(dims, free, mode=1, use_single=use_single)
        assert [int(c) for c in classes] == GOLD["single_fiber"]["fiberclass"]


def test_fiber_from_structure_snapshot():
    # tests.rs:244-270: dims (4,1,5,7,2), fibre [1,1,0,1,0]; snapshot holds EXPANDED indices
    dims = [4, 1, 5, 7, 2]
    free = [1, 1, 0, 1, 0]
    strides = [70, 70, 14, 2, 1]

    def expand(f):
        return [(f // s) % d for s, d in zip(strides, dims)]

    idx, _ = O.flat_iter_dims(dims, free, mode=0)
    assert [expand(int(i)) for i in idx] == GOLD["fiber_from_structure"]["fiber_iter"]
    _, items = _class_major(dims, free)
    assert len(items) == 280
    assert [expand(i) for i in items] == GOLD["fiber_from_structure"]["fiberclass_iter"]


# ---- seeded tensors: inputs shared by the oracle tests and the device tests ------------------------------
def trace_case():
    # tests.rs:295-307: 5x5 euclidean, both slots (euc5, index 1); i8 Standard, seed 3
    return dense_of(R.test_tensor(25, 3, "i8"), [5, 5])


def arithmetic_case():
    # tests.rs:1143-1156
    sa = R.test_structure(3, 1)
    shape = [d for _, d, _ in sa]
    size = int(np.prod(shape))
    ea = R.test_tensor(size, 1, "i32", -1000, 1000)
    eb = R.test_tensor(size, 2, "i32", -1000, 1000)
    return sa, shape, ea, eb


def scalar_dim1_case():
    # tests.rs:385-414; the legacy merge of structures without common slots appends
    common = R.test_structure_with_dims([1, 3, 1, 2], 6)
    commondual = R.test_structure_with_dims([-1, -3, -1, -2], 6)
    sa = R.test_structure(1, 32) + common
    sb = R.test_structure(1, 22) + commondual
    sha, shb = [d for _, d, _ in sa], [d for _, d, _ in sb]
    e1 = R.test_tensor(int(np.prod(sha)), 3, "i16", -100, 100)
    e1[0] = 45
    e2 = R.test_tensor(int(np.prod(shb)), 2, "i16", -100, 100)
    e2[0] = 2
    return sa, sha, e1, sb, shb, e2


def multi_contract_case():
    # tests.rs:744-773: 4 common indices (3 Lorentz/dual-Lorentz + 1 Euclidean, dims 4,5,7,8), inserted at
    # random positions => the contracted slots appear in ROTATED order in the two operands (K-note 2)
    s = 18
    rng = R.Xoroshiro64Star(s)
    nc = rng.gen_range_i32(2, 5)
    common = R.test_structure_with_id(range(1, nc + 1), s)
    dual = R.test_structure_with_id([-i for i in range(1, nc + 1)], s)
    sa = R.test_structure_with_id([nc + 1], s)
    sb = R.test_structure_with_id([nc + 2], s)
    for c, dc in zip(common, dual):
        sa.insert(rng.gen_range_usize(0, len(sa)), c)
        sb.insert(rng.gen_range_usize(0, len(sb)), dc)
    sha, shb = [d for _, d, _ in sa], [d for _, d, _ in sb]
    ea = R.test_tensor(int(np.prod(sha)), s + 3, "i32", -1000, 1000)
    eb = R.test_tensor(int(np.prod(shb)), s + 4, "i32", -1000, 1000)
    return nc, sa, sha, ea, sb, shb, eb


def unflatten(entries, shape):
    return {tuple(int(i) for i in np.unravel_index(k, shape)): v for k, v in entries.items()}


def test_multi_contract_case_is_the_permuted_one():
    nc, sa, _, _, sb, _, _ = multi_contract_case()
    assert nc == 4
    assert [a for _, _, a in sa] == [1, 3, 2, 4, 5] and [a for _, _, a in sb] == [3, 1, 2, 4, 6]
    assert sum(r != "euc" for r, _, a in sa if a <= 4) == 3     # three dualizable contracted indices


# ---- oracle vs the reference's numbers ----------------------------------------------------------------------
def test_oracle_trace_snapshot():
    raw = O.slots(("euc", 5, 1), ("euc", 5, 1))
    data = trace_case().astype(np.complex128)
    h = O.lib().orc_tensor_dense_raw(O._ptr(raw), 2, O._ptr(np.ascontiguousarray(data).view(np.float64)))
    assert O.OTensor(h).trace().to_dense().reshape(-1)[0] == GOLD["trace_i8_seed3"] == 79


def test_oracle_arithmetic_data_snapshot():
    sa, shape, ea, eb = arithmetic_case()
    a = O.OTensor.sparse(oslots(sa), unflatten(ea, shape))
    b = O.OTensor.sparse(oslots(sa), unflatten(eb, shape))
    c = a.add(b)
    got = by_label(c.slots["aind"], c.to_dense(), [x for _, _, x in sa]).reshape(-1)
    assert [int(v.real) for v in got] == GOLD["arithmetic_data"] and not got.imag.any()
    # dense + dense, the same numbers
    da = O.OTensor.dense(oslots(sa), dense_of(ea, shape))
    db = O.OTensor.dense(oslots(sa), dense_of(eb, shape))
    got = by_label(c.slots["aind"], da.add(db).to_dense(), [x for _, _, x in sa]).reshape(-1)
    assert [int(v.real) for v in got] == GOLD["arithmetic_data"]


def _storage_combos(mk_sparse, mk_dense):
    return [(mk_sparse, mk_sparse), (mk_sparse, mk_dense), (mk_dense, mk_sparse), (mk_dense, mk_dense)]


def test_oracle_scalar_and_dim1_snapshot():
    sa, sha, e1, sb, shb, e2 = scalar_dim1_case()
    g = GOLD["scalar_and_dim1"]
    mk = {
        "s": lambda sl, sh, e: O.OTensor.sparse(oslots(sl), unflatten(e, sh)),
        "d": lambda sl, sh, e: O.OTensor.dense(oslots(sl), dense_of(e, sh)),
    }
    for ka in "sd":
        for kb in "sd":
            f = mk[ka](sa, sha, e1).contract(mk[kb](sb, shb, e2))
            assert f.is_sparse == (ka == "s" and kb == "s")
            full = by_label(f.slots["aind"], f.to_dense(), [102, 28]).reshape(-1)
            assert not full.imag.any()
            # f.data() of the sparse result = stored values in ascending flat order (sparse.rs:306-310)
            assert [int(v.real) for v in full if v != 0] == g["data"], (ka, kb)
    f = mk["s"](sa, sha, e1).contract(mk["s"](sb, shb, e2))
    assert len(f.entries()) == len(g["data"])               # sparse index set: exactly the non-zeros
    assert sorted((int(s["dim"]), int(s["aind"])) for s in f.slots) == [(4, 102), (5, 28)]


def test_oracle_multi_contract_snapshot():
    nc, sa, sha, ea, sb, shb, eb = multi_contract_case()
    want = np.array(GOLD["multi_contract"]["data"]).reshape(4, 4)          # axes (6, 5)
    assert [s["aind"] for s in GOLD["multi_contract"]["result"]] == [6, 5]
    mk = {
        "s": lambda sl, sh, e: O.OTensor.sparse(oslots(sl), unflatten(e, sh)),
        "d": lambda sl, sh, e: O.OTensor.dense(oslots(sl), dense_of(e, sh)),
    }
    for ka in "sd":
        for kb in "sd":
            for flip in (False, True):
                x, y = mk[kb](sb, shb, eb), mk[ka](sa, sha, ea)
                f = y.contract(x) if flip else x.contract(y)               # densor_b.contract(&densor_a)
                got = by_label(f.slots["aind"], f.to_dense(), [6, 5])
                assert (got == want).all(), (ka, kb, flip)


# ---- the device path (C ABI) vs the reference's numbers -------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    import spenso_b200 as sp
    return sp.Context(0)


def eslots(tuples):
    import spenso_b200 as sp
    reps = {"euc": sp.Euclidean, "mink": sp.Minkowski, "lor": sp.Lorentz, "lor_d": sp.Lorentz.dual()}
    return [reps[r].new_slot(d, a) for r, d, a in tuples]


@pytest.mark.gpu
def test_device_trace_snapshot(ctx):
    import spenso_b200 as sp
    data = trace_case().astype(np.complex128)
    holder = sp.DenseTensor.from_data(ctx, eslots([("euc", 5, 900), ("euc", 5, 901)]), data)
    got = holder.trace(eslots([("euc", 5, 1), ("euc", 5, 1)])).to_numpy().reshape(-1)
    assert got[0] == 79


@pytest.mark.gpu
def test_device_arithmetic_data_snapshot(ctx):
    import spenso_b200 as sp
    sa, shape, ea, eb = arithmetic_case()
    order = [x for _, _, x in sa]
    for mk in (lambda e: sp.SparseTensor.from_entries(ctx, eslots(sa), unflatten(e, shape)),
               lambda e: sp.DenseTensor.from_data(ctx, eslots(sa), dense_of(e, shape))):
        c = mk(ea) + mk(eb)
        got = by_label(c.slots["aind"], c.to_numpy().reshape([int(d) for d in c.slots["dim"]]), order).reshape(-1)
        assert [int(v.real) for v in got] == GOLD["arithmetic_data"] and not got.imag.any()


@pytest.mark.gpu
def test_device_scalar_and_dim1_snapshot(ctx):
    import spenso_b200 as sp
    sa, sha, e1, sb, shb, e2 = scalar_dim1_case()
    mk = {
        "s": lambda sl, sh, e: sp.SparseTensor.from_entries(ctx, eslots(sl), unflatten(e, sh)),
        "d": lambda sl, sh, e: sp.DenseTensor.from_data(ctx, eslots(sl), dense_of(e, sh)),
    }
    for ka in "sd":
        for kb in "sd":
            f = mk[ka](sa, sha, e1).contract(mk[kb](sb, shb, e2))
            assert f.is_sparse == (ka == "s" and kb == "s")
            arr = f.to_numpy().reshape([int(d) for d in f.slots["dim"]])
            full = by_label(f.slots["aind"], arr, [102, 28]).reshape(-1)
            assert [int(v.real) for v in full if v != 0] == GOLD["scalar_and_dim1"]["data"], (ka, kb)
    f = mk["s"](sa, sha, e1).contract(mk["s"](sb, shb, e2))
    assert len(f.entries()) == len(GOLD["scalar_and_dim1"]["data"])


@pytest.mark.gpu
def test_device_multi_contract_snapshot(ctx):
    """K5 with four contracted indices in rotated order, all storage combinations, both receiver
    orders, and batched (the same tensors replicated over points, both layouts): bit-equal to the
    numbers the reference itself produced."""
    import spenso_b200 as sp
    nc, sa, sha, ea, sb, shb, eb = multi_contract_case()
    want = np.array(GOLD["multi_contract"]["data"]).reshape(4, 4)
    mk = {
        "s": lambda sl, sh, e: sp.SparseTensor.from_entries(ctx, eslots(sl), unflatten(e, sh)),
        "d": lambda sl, sh, e: sp.DenseTensor.from_data(ctx, eslots(sl), dense_of(e, sh)),
    }
    for ka in "sd":
        for kb in "sd":
            for flip in (False, True):
                x, y = mk[kb](sb, shb, eb), mk[ka](sa, sha, ea)
                f = y.contract(x) if flip else x.contract(y)
                arr = f.to_numpy().reshape([int(d) for d in f.slots["dim"]])
                assert (by_label(f.slots["aind"], arr, [6, 5]) == want).all(), (ka, kb, flip)
    # batched per-point operands (generic kernel at 3 points, specialised/JIT kernel at 4096 points)
    for n_points in (3, 4096):
        for layout in (sp.POINT_MAJOR, sp.COMPONENT_MAJOR):
            da = np.broadcast_to(dense_of(ea, sha), [n_points] + sha).astype(np.complex128)
            db = np.broadcast_to(dense_of(eb, shb), [n_points] + shb).astype(np.complex128)
            ta = sp.BatchTensor.from_numpy(ctx, eslots(sa), da, layout=layout)
            tb = sp.BatchTensor.from_numpy(ctx, eslots(sb), db, layout=layout)
            shared_b = mk["s"](sb, shb, eb)
            for f in (tb.contract(ta), ta.contract(tb), shared_b.contract(ta), ta.contract(shared_b)):
                arr = f.to_numpy().reshape([n_points] + [int(d) for d in f.slots["dim"]])
                for p in (0, n_points - 1):
                    assert (by_label(f.slots["aind"], arr[p], [6, 5]) == want).all(), (n_points, layout)


# ---- the reference's 1000-seed sweeps (tests.rs:585-632 all_single_contractions, :800-851
#      all_multi_contractions): same seeded inputs; the reference asserts the storage combinations agree,
#      here they must also equal numpy.einsum (exact integers) ------------------------------------------------
def sweep_case(s: int, multi: bool):
    rng = R.Xoroshiro64Star(s)
    if multi:
        nc = rng.gen_range_i32(2, 5)
        common = R.test_structure_with_id(range(1, nc + 1), s)
        dual = R.test_structure_with_id([-i for i in range(1, nc + 1)], s)
        sa = R.test_structure_with_id([nc + 1], s)
        sb = R.test_structure_with_id([nc + 2], s)
        free_b, free_a = [nc + 2], [nc + 1]
    else:
        common = R.test_structure_with_id([1], s)
        dual = R.test_structure_with_id([-1], s)
        sa = R.test_structure_with_id([2], s)
        sb = R.test_structure_with_id([-3, 4], s)
        free_b, free_a = [3, 4], [2]
    for c, dc in zip(common, dual):
        sa.insert(rng.gen_range_usize(0, len(sa)), c)
        sb.insert(rng.gen_range_usize(0, len(sb)), dc)
    sha, shb = [d for _, d, _ in sa], [d for _, d, _ in sb]
    ea = R.test_tensor(int(np.prod(sha)), s + 3, "i32", -1000, 1000)
    eb = R.test_tensor(int(np.prod(shb)), s + 4, "i32", -1000, 1000)
    L = "abcdefghijkl"
    out = free_b + free_a
    expr = "".join(L[a] for _, _, a in sb) + "," + "".join(L[a] for _, _, a in sa) + "->" + "".join(L[a] for a in out)
    want = np.einsum(expr, dense_of(eb, shb), dense_of(ea, sha))
    return sa, sha, ea, sb, shb, eb, out, want


@pytest.mark.parametrize("multi", [False, True])
def test_oracle_reference_seed_sweeps_equal_einsum(multi):
    mk = {
        "s": lambda sl, sh, e: O.OTensor.sparse(oslots(sl), unflatten(e, sh)),
        "d": lambda sl, sh, e: O.OTensor.dense(oslots(sl), dense_of(e, sh)),
    }
    for s in range(0, 1000, 4 if multi else 2):
        sa, sha, ea, sb, shb, eb, out, want = sweep_case(s, multi)
        if max(np.prod(sha), np.prod(shb)) > 20000:
            continue
        for ka, kb in (("d", "d"), ("s", "s"), ("s", "d"), ("d", "s")):
            try:
                f = mk[kb](sb, shb, eb).contract(mk[ka](sa, sha, ea))
            except O.OracleError:
                # the reference unwrap_or(zeros) there too: empty sparse operand (singlecontract.rs:153-159)
                assert (len(ea) == 0 or len(eb) == 0) and "s" in (ka, kb)
                assert not want.any()
                continue
            got = by_label(f.slots["aind"], f.to_dense(), out)
            assert (got == want).all(), (s, ka, kb)


@pytest.mark.gpu
@pytest.mark.parametrize("multi", [False, True])
def test_device_reference_seed_sweeps_equal_einsum(ctx, multi):
    import spenso_b200 as sp
    mk = {
        "s": lambda sl, sh, e: sp.SparseTensor.from_entries(ctx, eslots(sl), unflatten(e, sh)),
        "d": lambda sl, sh, e: sp.DenseTensor.from_data(ctx, eslots(sl), dense_of(e, sh)),
    }
    done = 0
    for s in range(0, 1000, 16 if multi else 8):
        sa, sha, ea, sb, shb, eb, out, want = sweep_case(s, multi)
        if max(np.prod(sha), np.prod(shb)) > 20000:
            continue
        for ka, kb in (("d", "d"), ("s", "s"), ("s", "d"), ("d", "s")):
            try:
                f = mk[kb](sb, shb, eb).contract(mk[ka](sa, sha, ea))
            except sp.SpensoError:
                assert (len(ea) == 0 or len(eb) == 0) and "s" in (ka, kb)
                continue
            arr = f.to_numpy().reshape([int(d) for d in f.slots["dim"]])
            assert (by_label(f.slots["aind"], arr, out) == want).all(), (s, ka, kb)
            done += 1
    assert done > 100


# ---- tests.rs:1074-1141 sparse_addition / sparse_sub: operands given in different slot orders are added by index
#      label, and an entry that cancels exactly is not stored ---------------------------------------------------------
def _sparse_add_sub_cases(mk):
    a = mk([("euc", 2, 2), ("euc", 2, 1)], {(1, 0): 1.0, (0, 1): 2.0})
    b = mk([("euc", 2, 1), ("euc", 2, 2)], {(1, 0): 1.0, (0, 1): 2.0})
    a2 = mk([("euc", 2, 2), ("euc", 2, 1)], {(1, 0): 1.0, (0, 1): 2.0})
    b2 = mk([("euc", 2, 2), ("euc", 2, 1)], {(1, 0): 1.0, (0, 1): 3.0})
    return a, b, a2, b2


def _labelled_entries(t, shape=(2, 2)):
    """{(index of aind 2, index of aind 1): value} of the stored entries"""
    ainds = [int(x) for x in t.slots["aind"]]
    out = {}
    for flat, v in t.entries().items():
        idx = np.unravel_index(flat, shape)
        out[(int(idx[ainds.index(2)]), int(idx[ainds.index(1)]))] = v
    return out


def test_oracle_sparse_addition_and_sub_kats():
    a, b, a2, b2 = _sparse_add_sub_cases(lambda sl, e: O.OTensor.sparse(oslots(sl), e))
    f = a.add(b, fallible=True)
    assert f.is_sparse and _labelled_entries(f) == {(0, 1): 3.0, (1, 0): 3.0}
    g = a2.add(b2, subtract=True, fallible=True)                           # sub_fallible
    assert g.is_sparse and _labelled_entries(g) == {(0, 1): -1.0}          # 1 - 1 cancels: not stored
    h = a2.add(b2, subtract=True)                                          # `-=` (sub_assign.rs) keeps the explicit zero
    assert h.is_sparse and _labelled_entries(h) == {(0, 1): -1.0, (1, 0): 0.0}


@pytest.mark.gpu
def test_device_sparse_addition_and_sub_kats(ctx):
    import spenso_b200 as sp
    a, b, a2, b2 = _sparse_add_sub_cases(lambda sl, e: sp.SparseTensor.from_entries(ctx, eslots(sl), e))
    f = a.add_fallible(b)
    assert f.is_sparse and _labelled_entries(f) == {(0, 1): 3.0, (1, 0): 3.0}
    g = a2.sub_fallible(b2)
    assert g.is_sparse and _labelled_entries(g) == {(0, 1): -1.0}
    h = a2 - b2
    assert h.is_sparse and _labelled_entries(h) == {(0, 1): -1.0, (1, 0): 0.0}


# ---- tests.rs:903-973 mixed_tensor_contraction, :1471-1510 transpose, :476-500 contract_with_rank_one_in_middle ---------
IM = complex(1.5, 1.25)


def _mixed_cases(sparse, dense):
    """returns [(result tensor, label order of the reference's result, expected row-major data)]"""
    sa = [("euc", 2, 2), ("euc", 2, 1)]
    sb = [("euc", 2, 2), ("euc", 2, 4)]
    a = sparse(sa, {(0, 0): 1.0, (1, 1): 2.0})                      # real data: f64 upgrades to (x, 0.0)
    b = dense(sb, np.array([1, 2, 3, 4]) * IM)
    f1 = b.contract(a)                                              # legacy result order: b's free (4), a's free (1)
    a2 = sparse(sa, {(0, 0): 1.0 * IM, (1, 1): 2.0 * IM})
    b2 = dense(sb, np.array([1.0, 2.0, 3.0, 4.0]))
    f2 = a2.contract(b2)                                            # a's free (1), b's free (4)
    return [(f1, [4, 1], np.array([1, 6, 2, 8]) * IM), (f2, [1, 4], np.array([1, 2, 6, 8]) * IM)]


def _transpose_case(dense):
    a = dense([("euc", 2, 1)], [1, 2])
    b = dense([("euc", 2, 2)], [3, 4])
    m_t = dense([("euc", 2, 1), ("euc", 2, 2)], [0, 1, 0, 0])
    m_d = dense([("euc", 2, 2), ("euc", 2, 1)], [0, 0, 1, 0])      # the same matrix, data laid out transposed
    return a.contract(m_t).contract(b), a.contract(m_d).contract(b)


def _rank_one_in_middle_case():
    s = 12
    common = R.test_structure(3, s)
    sa = R.test_structure(2, s + 1) + common                        # legacy merge without common slots appends
    sb = R.test_structure(1, s + 2) + common
    sha, shb = [d for _, d, _ in sa], [d for _, d, _ in sb]
    ea = R.test_tensor(int(np.prod(sha)), s + 3, "i32", -1000, 1000)
    eb = R.test_tensor(int(np.prod(shb)), s + 4, "i32", -1000, 1000)
    L = {a: chr(ord("a") + k) for k, a in enumerate(sorted({x for _, _, x in sa + sb}))}
    cl = [y for _, _, y in common]
    free = [x for _, _, x in sb if x not in cl] + [x for _, _, x in sa if x not in cl]
    expr = "".join(L[x] for _, _, x in sb) + "," + "".join(L[x] for _, _, x in sa) + "->" + "".join(L[x] for x in free)
    # mink slots carry the metric sign on contraction: apply it to one operand before the plain einsum
    B = dense_of(eb, shb).astype(np.int64)
    for ax, (r, d, _) in enumerate(sb):
        if r == "mink" and (r, d, sb[ax][2]) in common:
            sign = np.ones(d, dtype=np.int64)
            sign[1:] = -1
            B = B * sign.reshape([-1 if k == ax else 1 for k in range(len(sb))])
    want = np.einsum(expr, B, dense_of(ea, sha))
    return sa, sha, ea, sb, shb, eb, free, want


def test_oracle_mixed_transpose_rank_one_kats():
    sparse = lambda sl, e: O.OTensor.sparse(oslots(sl), e)
    dense = lambda sl, d: O.OTensor.dense(oslots(sl), d)
    for f, order, want in _mixed_cases(sparse, dense):
        assert not f.is_sparse                                      # Dense unless both operands are Sparse
        assert (by_label(f.slots["aind"], f.to_dense(), order).reshape(-1) == want).all()
    ft, fd = _transpose_case(dense)
    assert ft.to_dense().reshape(-1)[0] == fd.to_dense().reshape(-1)[0] == 4
    sa, sha, ea, sb, shb, eb, free, want = _rank_one_in_middle_case()
    f = sparse(sb, unflatten(eb, shb)).contract(sparse(sa, unflatten(ea, sha)))
    g = dense(sb, dense_of(eb, shb)).contract(dense(sa, dense_of(ea, sha)))
    assert (f.to_dense() == g.to_dense()).all()                     # the reference's assertion
    assert (by_label(g.slots["aind"], g.to_dense(), free) == want).all()


@pytest.mark.gpu
def test_device_mixed_transpose_rank_one_kats(ctx):
    import spenso_b200 as sp
    sparse = lambda sl, e: sp.SparseTensor.from_entries(ctx, eslots(sl), e)
    dense = lambda sl, d: sp.DenseTensor.from_data(ctx, eslots(sl), d)
    arr = lambda t: t.to_numpy().reshape([int(d) for d in t.slots["dim"]])
    for f, order, want in _mixed_cases(sparse, dense):
        assert not f.is_sparse
        assert (by_label(f.slots["aind"], arr(f), order).reshape(-1) == want).all()
    ft, fd = _transpose_case(dense)
    assert ft.to_numpy().reshape(-1)[0] == fd.to_numpy().reshape(-1)[0] == 4
    sa, sha, ea, sb, shb, eb, free, want = _rank_one_in_middle_case()
    f = sparse(sb, unflatten(eb, shb)).contract(sparse(sa, unflatten(ea, sha)))
    g = dense(sb, dense_of(eb, shb)).contract(dense(sa, dense_of(ea, sha)))
    assert (arr(f) == arr(g)).all()
    assert (by_label(g.slots["aind"], arr(g), free) == want).all()
