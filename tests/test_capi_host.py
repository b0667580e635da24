"""Host side of the C ABI (no GPU): the library loads and exports every symbol include/spenso_b200.h
declares; canonical ordering, merge and stride plans are bit-exact with the oracle's restatement of
spenso/src/structure/ordered.rs; the network compiler builds schedules and sm_100a kernels; compute
entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as O
import _cases
import spenso_b200 as sp
from spenso_b200 import _capi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "spenso_b200.h")).read()
    declared = set(re.findall(r"\b(spn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 50
    L = C.CDLL(_capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    # and the binding table covers exactly the header
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    assert sp.lib().spn_version() == 100


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sp.SpensoError) as e:
        sp.Context(0)
    assert e.value.code == 6 and "no CPU execution path" in str(e.value)
    net = workloads.build_engine(workloads.eemumu(), None)
    with pytest.raises(sp.SpensoError) as e:
        net.execute([], None, sum_out_ptr=16)
    assert e.value.code in (6, 7)


@pytest.mark.parametrize("seed", range(8))
def test_canonicalize_bit_exact_with_oracle(seed):
    rng = np.random.default_rng(seed)
    for _ in range(50):
        a, b = _cases.random_pair(rng, max_free=4)
        for sl in (a, b):
            if len(sl) == 0:
                continue
            oc, operm, obs, ods = O.order_structure(sl)
            ec, eperm, ebs, eds = sp.canonicalize(sl)
            assert oc.tobytes() == ec.tobytes()
            assert list(operm) == list(eperm)
            # block starts: the oracle reports them like ordered.rs (None -> -1)
            assert (obs, ods) == (ebs, eds)


def _plan_fields(p, na, nb):
    out_slots = np.frombuffer(bytes(p.out_slots), dtype=sp.SLOT_DTYPE)[: p.order_out]
    return dict(slots=out_slots, common_a=np.array(p.common_a[:na], dtype=np.uint8),
                common_b=np.array(p.common_b[:nb], dtype=np.uint8),
                partition=np.array(p.partition[: p.n_partition], dtype=np.uint8), kind=p.merge_kind)


@pytest.mark.parametrize("seed", range(8))
def test_merge_plan_bit_exact_with_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    n_checked = 0
    for _ in range(120):
        a, b = _cases.random_pair(rng)
        ca = O.order_structure(a)[0] if len(a) else a
        cb = O.order_structure(b)[0] if len(b) else b
        try:
            om = O.merge(ca, cb)
        except O.OracleError:
            with pytest.raises(sp.SpensoError):
                sp.plan_contract(ca, cb)
            continue
        p = sp.plan_contract(ca, cb)
        e = _plan_fields(p, len(ca), len(cb))
        assert e["slots"].tobytes() == om["slots"].tobytes()
        assert (e["common_a"] == om["common_a"]).all() and (e["common_b"] == om["common_b"]).all()
        assert e["kind"] == om["kind"]
        if len(ca) and len(cb):
            assert (e["partition"] == om["partition"]).all()
            assert (p.base_start, p.dual_start) == (om["base_start"], om["dual_start"])
        n_checked += 1
    assert n_checked > 60


def _eval_plan_numpy(p, A, B):
    """out[o] = sum_k sign(k) A[offA(o)+kA(k)] B[offB(o)+kB(k)] — the plan's own definition
    (include/spenso_b200.h), evaluated with Python loops: checks strides/k tables, not speed."""
    A, B = A.reshape(-1), B.reshape(-1)
    out = np.zeros(p.size_out, dtype=np.complex128)
    odims = list(p.out_dim[: p.order_out])
    kd = list(p.k_dim[: p.n_common])
    for o in range(p.size_out):
        idx = np.unravel_index(o, odims) if odims else ()
        offa = sum(int(i) * int(s) for i, s in zip(idx, p.out_stride_a))
        offb = sum(int(i) * int(s) for i, s in zip(idx, p.out_stride_b))
        acc = 0j
        for k in range(int(p.n_contracted) if p.n_common else 1):
            kidx = np.unravel_index(k, kd) if kd else ()
            ka = sum(int(i) * int(s) for i, s in zip(kidx, p.k_stride_a))
            kb = sum(int(i) * int(s) for i, s in zip(kidx, p.k_stride_b))
            neg = sum(1 for d, i in enumerate(kidx) if p.k_metric[d] and i > 0) % 2
            term = A[offa + ka] * B[offb + kb]
            acc = acc - term if neg else acc + term
        out[o] = acc
    return out


@pytest.mark.parametrize("seed", range(4))
def test_stride_plan_reproduces_oracle_contraction(seed):
    rng = np.random.default_rng(200 + seed)
    done = 0
    for _ in range(40):
        a, b = _cases.random_pair(rng, max_free=2, max_common=2)
        ca = O.order_structure(a)[0] if len(a) else a
        cb = O.order_structure(b)[0] if len(b) else b
        A = _cases.random_complex(rng, [int(d) for d in ca["dim"]])
        B = _cases.random_complex(rng, [int(d) for d in cb["dim"]])
        try:
            ref = O.OTensor.dense(ca, A, canonicalize=False).contract(O.OTensor.dense(cb, B, canonicalize=False))
        except O.OracleError:
            continue
        p = sp.plan_contract(ca, cb)
        got = _eval_plan_numpy(p, A, B)
        want = ref.to_dense().reshape(-1)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
        done += 1
    assert done > 20


def test_network_compile_plan_only():
    for name, n_in, out_size in (("cfg1", 2, 16), ("cfg2", 4, 1), ("cfg3", 2, 1), ("cfg5", 2, 16), ("gl03", 2, 16)):
        spec = workloads.WORKLOADS[name]()
        net = workloads.build_engine(spec, None)
        assert net.n_inputs == n_in and net.result_size == out_size
        src = net.cuda_source
        assert "__global__" in src and "spn_fused_net_" in src
        assert len(net.schedule) >= 2
    # cfg2: reads exactly the four spinors, 256-bit loads, no constant survives as a multiply
    net = workloads.build_engine(workloads.eemumu(), None)
    assert net.input_bytes_per_point == 256
    assert "ld.global.nc.L1::no_allocate.v4.f64" in net.cuda_source   # 256-bit streaming loads
    assert net.fp64_instr_per_point <= 160
    # cfg3: the gather touches only the components facing non-zeros of T (17) and f (54)
    net = workloads.build_engine(workloads.su3_colour(), None)
    assert net.input_bytes_per_point == (17 + 54) * 16
    # literal gl_03 vanishes identically: nothing is read, nothing is computed
    net = workloads.build_engine(workloads.one_loop_gl03(), None)
    assert net.input_bytes_per_point == 0 and net.fp64_instr_per_point == 0


def test_network_result_structure_matches_oracle():
    for name in ("cfg1", "cfg5"):
        spec = workloads.WORKLOADS[name]()
        arrays = workloads.synthetic_inputs(spec, 1, 3)
        _, _, oslots = O.eval_spec(spec, arrays)
        net = workloads.build_engine(spec, None)
        assert net.result_slots.tobytes() == oslots.tobytes()


def test_network_errors():
    net = sp.Network(None)
    a = net.batched([sp.Minkowski.new_slot(4, 1)])
    s = net.scalar(2.0)
    net.set_root(net.sum(a, s))
    with pytest.raises(sp.SpensoError) as e:   # SumScalarTensor (network.rs:1342-1467)
        net.compile()
    assert e.value.code == 2
    net = sp.Network(None)
    a = net.batched([sp.Minkowski.new_slot(4, 1)])
    b = net.batched([sp.Minkowski.new_slot(4, 2)])
    net.set_root(net.power(net.product(a, b), -1))   # NegativeExponentNonScalar
    with pytest.raises(sp.SpensoError):
        net.compile()
    with pytest.raises(sp.SpensoError):
        sp.Network(None).compile()   # no root


def test_stepwise_compile_plan_only():
    net = workloads.build_engine(workloads.eemumu(), None, sp.STEPWISE)
    assert len(net.schedule) == 5 and net.result_size == 1


def test_plan_json_round_trip():
    """Plan interchange: export -> import -> compile must give the identical generated kernel."""
    import json

    for name in ("cfg1", "cfg2", "cfg3", "cfg5"):
        net = workloads.build_engine(workloads.WORKLOADS[name](), None)
        text = net.to_json()
        doc = json.loads(text)                      # it is real JSON
        assert doc["version"] == 1 and doc["nodes"][doc["root"]]["kind"] == "op"
        net2 = sp.Network.from_json(None, text).compile(sp.FUSED)
        assert net2.cuda_source == net.cuda_source
        assert net2.result_slots.tobytes() == net.result_slots.tobytes()
        assert net2.to_json() == text
    # a document written by hand (what a Rust shim would emit with serde_json): p(mu) q(mu) with f64 inputs
    doc = {"version": 1, "root": 2, "nodes": [
        {"kind": "batched", "dtype": "f64", "slots": [[2, 0, 4, 0, 7]]},
        {"kind": "batched", "dtype": "f64", "slots": [[2, 0, 4, 0, 7]]},
        {"kind": "op", "op": "product", "children": [0, 1]}]}
    net = sp.Network.from_json(None, json.dumps(doc)).compile()
    assert net.n_inputs == 2 and net.result_size == 1 and net.fp64_instr_per_point == 4
    for bad in ("{", '{"version":2,"nodes":[],"root":0}', '{"version":1,"nodes":[{"kind":"what"}],"root":0}',
                '{"version":1,"nodes":[{"kind":"op","op":"product","children":[5]}],"root":0}'):
        with pytest.raises(sp.SpensoError):
            sp.Network.from_json(None, bad)


def test_maximum_orders():
    """SPN_MAX_ORDER = 16 indices per operand, 32 per result; one more is refused, not truncated."""
    a = sp.slots([sp.Euclidean.new_slot(1 + (i % 2), 100 + i) for i in range(16)])
    b = sp.slots([sp.Euclidean.new_slot(1 + (i % 2), 200 + i) for i in range(16)])
    ca, cb = sp.canonicalize(a)[0], sp.canonicalize(b)[0]
    assert O.order_structure(a)[0].tobytes() == ca.tobytes()
    p = sp.plan_contract(ca, cb)
    assert p.order_out == 32 and p.n_common == 0 and p.size_out == 2 ** 16
    om = O.merge(ca, cb)
    assert np.frombuffer(bytes(p.out_slots), dtype=sp.SLOT_DTYPE)[:32].tobytes() == om["slots"].tobytes()
    big = sp.slots([sp.Euclidean.new_slot(1, 300 + i) for i in range(17)])
    with pytest.raises(sp.SpensoError):
        sp.plan_contract(sp.canonicalize(big)[0], cb)
    # fully contracted 16-index pair -> scalar
    p = sp.plan_contract(ca, ca)
    assert p.order_out == 0 and p.n_common == 16 and p.size_out == 1 and p.n_contracted == 2 ** 8


def test_exec_mode_auto():
    """AUTO = fused for the physics-sized networks, stepwise when intermediates are too large for straight-line code."""
    net = workloads.build_engine(workloads.eemumu(), None, sp.AUTO)
    assert net.exec_mode == sp.FUSED and "spn_fused_net_" in net.cuda_source
    big = sp.Network(None)
    e = sp.Euclidean
    a = big.batched([e.new_slot(16, 1), e.new_slot(16, 2), e.new_slot(16, 3)])
    b = big.batched([e.new_slot(16, 3), e.new_slot(16, 4), e.new_slot(16, 5)])
    big.set_root(big.product(a, b))          # 65536-component result, 16 terms each
    big.compile(sp.AUTO)
    assert big.exec_mode == sp.STEPWISE and big.result_size == 16 ** 4


def test_network_expression_operators():
    """Reference-style expression building (Network::from_tensor(a) * ... + ...) compiles to the same kernel as the
    explicit node calls."""
    m1, m2 = [sp.Minkowski.new_slot(4, 1)], [sp.Minkowski.new_slot(4, 2)]
    a = sp.Network(None)
    p, q = a.from_tensor(None, m1), a.from_tensor(None, m1)
    r = a.from_tensor(None, m2)
    a.set_root((p * q) * r * 2.0 - (q * q) ** 2 * r.conj())
    a.compile()
    b = sp.Network(None)
    bp, bq, br = b.batched(m1), b.batched(m1), b.batched(m2)
    b.set_root(b.sum(b.product(b.product(b.product(bp, bq), br), b.scalar(2.0)),
                     b.neg(b.product(b.power(b.product(bq, bq), 2), b.conj(br)))))
    b.compile()
    assert a.cuda_source == b.cuda_source and a.result_size == 4
    with pytest.raises(TypeError):
        p * "x"


def test_merge_missed_match_is_restated_literally():
    """SURVEY K-note 4 (ordered.rs:717-836): the sort key is (rep, dim, aind) but match_cmp compares
    (dim, rep, aind), so when one operand mixes dualizable reps of different dimensions the monotone scan
    can stop before the matching dual.  The oracle and the engine keep that behaviour (no "fix")."""
    b = O.order_structure(O.slots(("cof", 3, 2), ("euc", 2, 5)))[0]
    a_mixed = O.order_structure(O.slots(("lor_d", 4, 1), ("coaf", 3, 2)))[0]
    a_plain = O.order_structure(O.slots(("coaf", 3, 2),))[0]
    for x, y in ((a_mixed, b), (b, a_mixed)):
        om, p = O.merge(x, y), sp.plan_contract(x, y)
        assert om["common_a"].sum() == 0 and p.n_common == 0 and p.order_out == 4     # match missed
        assert _plan_fields(p, len(x), len(y))["slots"].tobytes() == om["slots"].tobytes()
    for x, y in ((a_plain, b), (b, a_plain)):
        om, p = O.merge(x, y), sp.plan_contract(x, y)
        assert om["common_a"].sum() == 1 and p.n_common == 1 and p.order_out == 1     # match found
        assert _plan_fields(p, len(x), len(y))["slots"].tobytes() == om["slots"].tobytes()


def test_non_finite_constants_compile_and_round_trip():
    """A non-finite constant has no C literal: the generated kernel spells it through its bit pattern (NVRTC
    compiles it), and the JSON plan format carries it as 1e999 / NaN."""
    net = sp.Network(None)
    x = net.batched([sp.Bispinor.new_slot(4, 1)])
    net.set_root(net.product(net.scalar(complex(float("inf"), 0.0)), x))
    text = net.to_json()
    assert "1e999" in text and "inf" not in text
    back = sp.Network.from_json(None, text)
    assert back.to_json() == text
    net.compile(sp.FUSED)                      # NVRTC runs here: "inf" in the source would fail to compile
    assert "__longlong_as_double(0x7ff0000000000000LL)" in net.cuda_source
    nan = sp.Network(None)
    y = nan.batched([sp.Bispinor.new_slot(4, 1)])
    nan.set_root(nan.sum(nan.product(nan.scalar(complex(0.0, float("nan"))), y), y))
    assert "NaN" in nan.to_json()
    assert sp.Network.from_json(None, nan.to_json()).to_json() == nan.to_json()


# ---- host-driven contraction order (spn_net_set_schedule) ---------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg5", "gl03"])
def test_host_schedule_is_replayed_literally(name):
    """The order the oracle's SmallestDegree restatement picks (its contraction log = what a Rust host would send) is
    replayed pair by pair, in both execution modes, and the engine reports it back in the same naming."""
    spec = workloads.WORKLOADS[name]()
    want = O.oracle_schedule(spec)
    assert len(want) >= 1
    for mode in (sp.STEPWISE, sp.FUSED):
        net = workloads.build_engine(spec, None, mode, schedule=want)
        got = net.schedule
        mapped = [(net.spec_ids[a], net.spec_ids[b]) for a, b in want]
        # the host's pairs come first and in order; what it leaves over (scalar factors, which ContractScalars folds
        # without a log entry) is multiplied in afterwards
        hosted = [p for p in got if p in mapped]
        assert hosted == mapped, (got, mapped)
        assert net.result_size == workloads.build_engine(spec, None, mode).result_size


def test_schedule_round_trips_between_networks_and_through_json():
    spec = workloads.one_loop_massive()
    a = workloads.build_engine(spec, None, sp.FUSED)
    sched = a.schedule                                   # the engine's own (searched) order, engine node ids
    b = sp.Network.from_json(None, a.to_json())
    b.set_schedule(sched)
    text = b.to_json()
    assert '"schedule":[[' in text
    b.compile(sp.FUSED)
    assert b.schedule == sched
    assert b.fp64_instr_per_point == a.fp64_instr_per_point and b.cuda_source == a.cuda_source
    c = sp.Network.from_json(None, text).compile(sp.STEPWISE)          # the schedule travels inside the plan text
    assert c.schedule == sched


def test_schedule_errors():
    # two factors without a common index are a legal step (an exterior product), so that compiles ...
    outer = workloads.build_engine(workloads.eemumu(), None, sp.STEPWISE, schedule=[(0, 1)])
    assert outer.schedule[0] == (outer.spec_ids[0], outer.spec_ids[1])
    # ... a pair that names something that is not a factor of one product does not
    net = sp.Network(None)
    x = net.batched([sp.Bispinor.new_slot(4, 1)])
    y = net.batched([sp.Bispinor.new_slot(4, 1)])
    z = net.batched([sp.Bispinor.new_slot(4, 2)])
    s = net.sum(net.product(x, y), net.scalar(1.0))
    net.set_root(net.product(s, z, z))
    net.set_schedule([(x, z)])                           # x belongs to the inner product, z to the outer one
    with pytest.raises(sp.SpensoError) as e:
        net.compile(sp.STEPWISE)
    assert e.value.code == 7 and "does not name two factors" in str(e.value)
    with pytest.raises(sp.SpensoError):
        net.set_schedule([(x, x)])
    with pytest.raises(sp.SpensoError):
        net.set_schedule([(x, 999)])


def test_sum_kernel_instruction_counts():
    """The fused Monte-Carlo-sum kernel of cfg5 (bench.py's variant, no per-point store): scale propagation removes the
    constant multiplies value numbering left behind, the trailing linear stage moves out of the point loop, and the
    basis chains absorb their accumulation.  339 FP64 instructions per point in round 1 (307 + 32 accumulations)."""
    net = workloads.build_engine(workloads.one_loop_massive(), None)
    assert net.fp64_instr_per_point <= 291
    assert net.variant_fp64_instr("ppns") <= 260
    assert net.variant_fp64_instr("ppps") == net.fp64_instr_per_point + 32     # with a per-point store nothing moves
    src = net.prepare_variant("ppns")
    assert "acc0 = fma(" in src and "const double sr0 = " in src
    real = workloads.build_engine(workloads.one_loop_massive_real(), None)
    assert real.variant_fp64_instr("ppns") <= 146


def test_symbolic_abstract_indices_are_interned_and_ordered_last():
    """AbstractIndex::Symbol (abstract_index.rs:267-294) sorts after Normal < Dualize < Double < Dummy < Added, and among
    symbols by creation order: names are interned to ids in first-seen order.  Canonical order and merge are bit-exact
    with the oracle, and a contraction over a symbolic index is matched like any other."""
    mu, nu, rho = (sp.symbol_slot(sp.Minkowski, 4, n) for n in ("test_mu", "test_nu", "test_rho"))
    assert sp.symbol_slot(sp.Minkowski, 4, "test_mu").aind == mu.aind and nu.aind == mu.aind + 1 and rho.aind == nu.aind + 1
    assert sp.lib().spn_symbol_name(mu.aind) == b"test_mu" and sp.lib().spn_symbol_name(1 << 40) is None
    given = sp.slots([rho, sp.Minkowski.new_slot(4, 7), mu, sp.Slot(sp.Minkowski, 4, 3, sp.AIND_ADDED), nu])
    canon, perm, _, _ = sp.canonicalize(given)
    ocanon, operm, _, _ = O.order_structure(given)
    assert canon.tobytes() == ocanon.tobytes() and list(perm) == list(operm)
    assert [int(k) for k in canon["aind_kind"]] == [0, 4, 5, 5, 5] and [int(a) for a in canon["aind"][2:]] == [mu.aind, nu.aind, rho.aind]
    a = sp.canonicalize([sp.Bispinor.new_slot(4, 1), mu, nu])[0]
    b = sp.canonicalize([nu, sp.Bispinor.new_slot(4, 2), rho])[0]
    p = sp.plan_contract(a, b)
    assert p.n_common == 1 and p.order_out == 4 and p.k_metric[0] == 1          # the Minkowski index nu is contracted
    om = O.merge(a, b)
    assert bytes(p.out_slots)[:16 * 4] == om["slots"].tobytes()


def test_sparse_operand_kernel_source_inspection():
    """Host-only inspection of the specialised pairwise kernels (source generation + NVRTC, no device): a shared sparse
    operand contributes its stored entries as literals (gamma^mu p_mu: 16 products, not 64); point-major operands and
    results are staged in warp-owned slabs; a component-major partner is read directly."""
    b = lambda a: sp.Bispinor.new_slot(4, a)
    m = lambda a: sp.Minkowski.new_slot(4, a)
    ent = {(0, 2, 0): 1, (1, 3, 0): 1, (2, 0, 0): 1, (3, 1, 0): 1, (0, 3, 1): 1, (1, 2, 1): 1, (2, 1, 1): -1, (3, 0, 1): -1,
           (0, 3, 2): -1j, (1, 2, 2): 1j, (2, 1, 2): 1j, (3, 0, 2): -1j, (0, 2, 3): 1, (1, 3, 3): -1, (2, 0, 3): -1, (3, 1, 3): 1}
    for first in (True, False):
        src = sp.contract_cuda_source([b(1), b(2), m(3)], [m(3)], sparse_entries=ent, sparse_first=first)
        assert src.count("cmul_rn(") == 16 + 1 and "warp-owned slabs" in src and "sC[15] = r15;" in src
        assert src.count("__dsub_rn(r") == 2 * 12           # the three spatial directions subtract (Minkowski metric)
    src = sp.contract_cuda_source([b(1), b(2), m(3)], [m(3)], sparse_entries=ent, layout_b=sp.COMPONENT_MAJOR)
    assert src and "warp-owned slabs" not in src and "csB + p" in src
    # an exterior product with a 256-component partner is row-parallel; 32 768 rows are left to the generic kernels
    assert "thread = (point, row of" in sp.contract_cuda_source([b(1), b(2), m(3)], [b(4), b(5), m(6), m(7)], sparse_entries=ent)
    c = lambda a: sp.ColorAdjoint.new_slot(8, a)
    assert sp.contract_cuda_source([b(1), b(2), m(3)], [c(4), c(5), c(6), c(7), c(8)], sparse_entries=ent) == ""
    # two batched tensors: the inspection entry point of the dense kernels
    assert "warp-owned slabs" in sp.contract_cuda_source([b(1), b(2)], [b(2), b(3)])


def test_row_parallel_kernel_source_inspection():
    """Per-point tensors beyond one thread's registers (colour algebra) get the row-parallel specialised kernel: thread =
    (point, row of one batched operand), a batched partner staged in shared memory, a sparse partner as literals.
    Host-only: source generation + NVRTC."""
    c = lambda a: sp.ColorAdjoint.new_slot(8, a)
    b = lambda a: sp.Bispinor.new_slot(4, a)
    src = sp.contract_cuda_source([c(1), c(2), c(3)], [c(3), c(4)])
    assert "thread = (point, row of A)" in src and "64 rows x 8 columns x 8 contracted" in src and "2 row(s) per thread" in src
    # two rows per thread: every component of the staged partner serves two products
    assert "__shared__" in src and src.count("cmul_rn(") == 2 * 64 + 1 and src.count("ld32(X") == 2 * 4 and src.count("st32(O") == 2 * 4
    assert src.count("const double2 y = ") == 64
    src = sp.contract_cuda_source([c(3), c(4)], [c(1), c(2), c(3)])
    assert "thread = (point, row of B)" in src                   # the rows come from the operand with more free components
    ent = {(1, 2, 3): 1.0, (2, 1, 3): -1.0, (4, 5, 6): 0.5j}
    src = sp.contract_cuda_source([c(2), c(3), c(4)], [c(1), c(2), c(3)], sparse_entries=ent, sparse_first=False)
    assert "thread = (point, row of A)" in src and src.count("cmul_rn(") == 3 + 1 and "__shared__" not in src
    # a full contraction of two 512-component tensors: the row is streamed (loop over k, offsets in constant tables)
    src = sp.contract_cuda_source([c(1), c(2), c(3)], [c(1), c(2), c(3)])
    assert "1 rows x 1 columns x 512 contracted" in src and "kTabX[512]" in src and "for (unsigned k = 0; k < 512u; ++k)" in src
    # small tensors keep the thread = point kernel
    assert "thread = (point" not in sp.contract_cuda_source([b(1), b(2)], [b(2), b(3)])
