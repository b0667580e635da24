"""ctypes driver for the ORACLE (oracle/liboracle.so) — test infrastructure only.

The oracle is the CPU restatement of the reference's contraction path
(oracle/spenso_oracle.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

# binary layout shared with include/spenso_b200.h : spn_slot
SLOT_DTYPE = np.dtype(
    [("rep_kind", "u1"), ("aind_kind", "u1"), ("rep_id", "<i2"), ("dim", "<u4"), ("aind", "<u8")]
)
assert SLOT_DTYPE.itemsize == 16

REP_DUMMY, REP_SELF_DUAL, REP_INLINE_METRIC, REP_DUALIZABLE = 0, 1, 2, 3
FIRST_BEFORE_SECOND, SECOND_BEFORE_FIRST, INTERLEAVED = 0, 1, 2
N_TENSOR, N_SCALAR, N_LIB, N_PRODUCT, N_SUM, N_NEG, N_POWER, N_CONJ = range(8)

# rep ids after idenso::representations::initialize()  
_REPS = {
    "euc": (REP_SELF_DUAL, 0),
    "bis": (REP_SELF_DUAL, 1),
    "coad": (REP_SELF_DUAL, 2),
    "mink": (REP_INLINE_METRIC, 0),
    "lor": (REP_DUALIZABLE, 1),
    "lor_d": (REP_DUALIZABLE, -1),
    "spf": (REP_DUALIZABLE, 2),
    "spf_d": (REP_DUALIZABLE, -2),
    "cof": (REP_DUALIZABLE, 3),
    "coaf": (REP_DUALIZABLE, -3),
    "cos": (REP_DUALIZABLE, 4),
    "cos_d": (REP_DUALIZABLE, -4),
}


def slot(rep: str, dim: int, aind: int, aind_kind: int = 0):
    k, i = _REPS[rep]
    return (k, aind_kind, i, dim, aind)


def slots(*items) -> np.ndarray:
    return np.array([slot(*it) for it in items], dtype=SLOT_DTYPE)


def build_oracle(force: bool = False) -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "spenso_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    # SPN_ORACLE_LIB: another build of the same sources (bench.py's cpu_baseline legs build one with -march=native
    # on the box they run on; the committed recipe is portable because the .so travels to a box with another CPU)
    override = os.environ.get("SPN_ORACLE_LIB")
    if override and os.path.exists(override):
        L = C.CDLL(override)
    else:
        build_oracle()
        L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    L.orc_last_error.restype = C.c_char_p
    L.orc_tensor_dense.restype = vp
    L.orc_tensor_dense.argtypes = [vp, C.c_int, vp]
    L.orc_tensor_dense_raw.restype = vp
    L.orc_tensor_dense_raw.argtypes = [vp, C.c_int, vp]
    L.orc_tensor_sparse.restype = vp
    L.orc_tensor_sparse.argtypes = [vp, C.c_int, vp, vp, u64]
    L.orc_tensor_free.argtypes = [vp]
    L.orc_tensor_order.argtypes = [vp]
    L.orc_tensor_is_sparse.argtypes = [vp]
    L.orc_tensor_size.argtypes = [vp]
    L.orc_tensor_size.restype = u64
    L.orc_tensor_nnz.argtypes = [vp]
    L.orc_tensor_nnz.restype = u64
    L.orc_tensor_slots.argtypes = [vp, vp]
    L.orc_tensor_to_dense.argtypes = [vp, vp]
    L.orc_tensor_entries.argtypes = [vp, vp, vp]
    L.orc_contract.argtypes = [vp, vp, C.POINTER(vp), C.c_int]
    L.orc_add.argtypes = [vp, vp, C.POINTER(vp), C.c_int]
    L.orc_neg.argtypes = [vp, C.POINTER(vp)]
    L.orc_scalar_mul.argtypes = [vp, dbl, dbl, C.POINTER(vp)]
    L.orc_trace.argtypes = [vp, C.POINTER(vp)]
    L.orc_order_structure.argtypes = [vp, C.c_int, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    L.orc_merge.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.POINTER(i32), vp, vp, vp, C.POINTER(i32),
                            C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    L.orc_match_indices.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.POINTER(i32), vp, vp]
    L.orc_flat_iter.restype = i64
    L.orc_flat_iter.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, u64, vp, i64]
    L.orc_metric_iter.restype = i64
    L.orc_metric_iter.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, i64]
    L.orc_lib_builtin.restype = vp
    L.orc_lib_builtin.argtypes = [C.c_int]
    L.orc_lib_metric.restype = vp
    L.orc_lib_metric.argtypes = [vp]
    L.orc_lib_custom.restype = vp
    L.orc_lib_custom.argtypes = [vp, C.c_int, vp, vp, u64, C.c_int]
    L.orc_lib_free.argtypes = [vp]
    L.orc_lib_instantiate.restype = vp
    L.orc_lib_instantiate.argtypes = [vp, vp, C.c_int]
    L.orc_net_new.restype = vp
    L.orc_net_free.argtypes = [vp]
    L.orc_net_add_tensor.argtypes = [vp, vp]
    L.orc_net_add_scalar.argtypes = [vp, dbl, dbl]
    L.orc_net_add_lib.argtypes = [vp, vp, vp, C.c_int]
    L.orc_net_add_op.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
    L.orc_net_set_root.argtypes = [vp, C.c_int]
    L.orc_net_set_emulate.argtypes = [vp, C.c_int]
    L.orc_net_execute.argtypes = [vp]
    L.orc_net_result.restype = vp
    L.orc_net_result.argtypes = [vp]
    L.orc_net_log.argtypes = [vp, vp, C.c_int]
    L.orc_net_eval_batch.argtypes = [vp, vp, C.c_int, vp, u64, vp, u64, C.c_int, vp]
    _lib = L
    return L


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def order_structure(sl: np.ndarray):
    """-> (canonical slots, perm, base_start, dual_start); canonical[i] = sl[perm[i]]."""
    L = lib()
    n = len(sl)
    sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
    out = np.zeros(n, dtype=SLOT_DTYPE)
    perm = np.zeros(n, dtype=np.int32)
    bs, ds = C.c_int64(), C.c_int64()
    rc = L.orc_order_structure(_ptr(sl), n, _ptr(out), _ptr(perm), C.byref(bs), C.byref(ds))
    if rc:
        raise OracleError(rc, L.orc_last_error().decode())
    return out, perm, bs.value, ds.value


def merge(a: np.ndarray, b: np.ndarray):
    L = lib()
    a = np.ascontiguousarray(a, dtype=SLOT_DTYPE)
    b = np.ascontiguousarray(b, dtype=SLOT_DTYPE)
    na, nb = len(a), len(b)
    out = np.zeros(na + nb + 1, dtype=SLOT_DTYPE)
    ca = np.zeros(max(na, 1), dtype=np.uint8)
    cb = np.zeros(max(nb, 1), dtype=np.uint8)
    part = np.zeros(na + nb + 1, dtype=np.uint8)
    nout, npart, kind = C.c_int32(), C.c_int32(), C.c_int32()
    bs, ds = C.c_int64(), C.c_int64()
    rc = L.orc_merge(_ptr(a), na, _ptr(b), nb, _ptr(out), C.byref(nout), _ptr(ca), _ptr(cb), _ptr(part),
                     C.byref(npart), C.byref(kind), C.byref(bs), C.byref(ds))
    if rc:
        raise OracleError(rc, L.orc_last_error().decode())
    return dict(slots=out[: nout.value].copy(), common_a=ca[:na].copy(), common_b=cb[:nb].copy(),
                partition=part[: npart.value].copy(), kind=kind.value, base_start=bs.value, dual_start=ds.value)


def flat_iter(sl, free_mask, mode, use_single=False, shift=0, cap=1 << 20):
    L = lib()
    sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
    fm = np.ascontiguousarray(free_mask, dtype=np.uint8)
    out = np.zeros(cap, dtype=np.uint64)
    n = L.orc_flat_iter(_ptr(sl), len(sl), _ptr(fm), mode, int(use_single), shift, _ptr(out), cap)
    if n < 0:
        raise OracleError(-1, L.orc_last_error().decode())
    return out[: min(n, cap)].astype(np.int64), n


def flat_iter_dims(dims, free_mask, mode, use_single=False, shift=0, cap=1 << 20):
    """flat_iter over raw dims in the given (unsorted) order — legacy snapshot tests."""
    L = lib()
    d = np.ascontiguousarray(dims, dtype=np.uint64)
    fm = np.ascontiguousarray(free_mask, dtype=np.uint8)
    out = np.zeros(cap, dtype=np.uint64)
    L.orc_flat_iter_dims.restype = C.c_int64
    L.orc_flat_iter_dims.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                     C.c_int64]
    n = L.orc_flat_iter_dims(_ptr(d), len(d), _ptr(fm), mode, int(use_single), shift, _ptr(out), cap)
    if n < 0:
        raise OracleError(-1, L.orc_last_error().decode())
    return out[: min(n, cap)].astype(np.int64), n


def metric_iter(sl, free_mask, conj=False, order=(), cap=1 << 20):
    L = lib()
    sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
    fm = np.ascontiguousarray(free_mask, dtype=np.uint8)
    od = np.ascontiguousarray(order, dtype=np.int64)
    out = np.zeros(cap, dtype=np.uint64)
    neg = np.zeros(cap, dtype=np.uint8)
    n = L.orc_metric_iter(_ptr(sl), len(sl), _ptr(fm), int(conj), _ptr(od), len(od), _ptr(out), _ptr(neg), cap)
    if n < 0:
        raise OracleError(-1, L.orc_last_error().decode())
    return out[:n].astype(np.int64), neg[:n].astype(bool)


class OTensor:
    """Handle to an oracle tensor (canonical slot order, row-major complex data)."""

    def __init__(self, handle):
        if not handle:
            raise OracleError(-1, lib().orc_last_error().decode())
        self.h = handle

    def __del__(self):
        try:
            if self.h:
                lib().orc_tensor_free(self.h)
                self.h = None
        except Exception:
            pass

    # -- construction ------------------------------------------------------
    @staticmethod
    def dense(sl: np.ndarray, data, canonicalize=True) -> "OTensor":
        """sl/data in the caller's index order; transposed into canonical order like
        PermutedStructure::from_iter(..) + permute_inds()."""
        sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
        shape = tuple(int(d) for d in sl["dim"])
        arr = np.asarray(data, dtype=np.complex128).reshape(shape)
        if canonicalize and len(sl):
            canon, perm, _, _ = order_structure(sl)
            arr = np.transpose(arr, axes=[int(p) for p in perm])
            sl = canon
        arr = np.ascontiguousarray(arr)
        return OTensor(lib().orc_tensor_dense(_ptr(sl), len(sl), _ptr(arr.view(np.float64))))

    @staticmethod
    def sparse(sl: np.ndarray, entries: dict) -> "OTensor":
        """entries: {index tuple (caller's order): value}."""
        sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
        canon, perm, _, _ = order_structure(sl) if len(sl) else (sl, [], 0, 0)
        cshape = [int(d) for d in canon["dim"]]
        strides = [1] * len(cshape)
        for i in range(len(cshape) - 2, -1, -1):
            strides[i] = strides[i + 1] * cshape[i + 1]
        flat = np.zeros(len(entries), dtype=np.uint64)
        vals = np.zeros(len(entries), dtype=np.complex128)
        for k, (idx, v) in enumerate(entries.items()):
            cidx = [idx[int(p)] for p in perm]
            flat[k] = sum(i * s for i, s in zip(cidx, strides))
            vals[k] = v
        return OTensor(lib().orc_tensor_sparse(_ptr(canon), len(canon), _ptr(flat), _ptr(vals.view(np.float64)),
                                               len(entries)))

    # -- inspection --------------------------------------------------------
    @property
    def order(self):
        return lib().orc_tensor_order(self.h)

    @property
    def is_sparse(self):
        return bool(lib().orc_tensor_is_sparse(self.h))

    @property
    def size(self):
        return int(lib().orc_tensor_size(self.h))

    @property
    def slots(self) -> np.ndarray:
        out = np.zeros(self.order, dtype=SLOT_DTYPE)
        lib().orc_tensor_slots(self.h, _ptr(out))
        return out

    def to_dense(self) -> np.ndarray:
        out = np.zeros(self.size, dtype=np.complex128)
        lib().orc_tensor_to_dense(self.h, _ptr(out.view(np.float64)))
        return out.reshape([int(d) for d in self.slots["dim"]])

    def entries(self) -> dict:
        n = int(lib().orc_tensor_nnz(self.h))
        flat = np.zeros(n, dtype=np.uint64)
        vals = np.zeros(n, dtype=np.complex128)
        lib().orc_tensor_entries(self.h, _ptr(flat), _ptr(vals.view(np.float64)))
        return {int(f): complex(v) for f, v in zip(flat, vals)}

    def labelled(self) -> dict:
        """{(aind,...) in canonical slot order: ndarray} helper: returns (ainds, dense array)."""
        return tuple(int(a) for a in self.slots["aind"]), self.to_dense()

    # -- algebra -----------------------------------------------------------
    def contract(self, other: "OTensor", emulate_reference_pairing=False) -> "OTensor":
        out = C.c_void_p()
        rc = lib().orc_contract(self.h, other.h, C.byref(out), int(emulate_reference_pairing))
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())
        return OTensor(out.value)

    def add(self, other, subtract=False, fallible=False):
        """AddAssign / SubAssign (default) or FallibleAdd / FallibleSub (fallible=True: exact zeros not stored)."""
        out = C.c_void_p()
        rc = lib().orc_add(self.h, other.h, C.byref(out), int(subtract) | (2 if fallible else 0))
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())
        return OTensor(out.value)

    def neg(self):
        out = C.c_void_p()
        lib().orc_neg(self.h, C.byref(out))
        return OTensor(out.value)

    def scalar_mul(self, s: complex):
        out = C.c_void_p()
        lib().orc_scalar_mul(self.h, float(np.real(s)), float(np.imag(s)), C.byref(out))
        return OTensor(out.value)

    def trace(self):
        out = C.c_void_p()
        rc = lib().orc_trace(self.h, C.byref(out))
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())
        return OTensor(out.value)


class OLib:
    GAMMA_WEYL, GAMMA_DIRAC, GAMMA0, GAMMA5, PROJM, PROJP = range(6)
    GAMMA_CONJ, GAMMA_ADJ = 7, 8

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def builtin(which: int) -> "OLib":
        return OLib(lib().orc_lib_builtin(which))

    @staticmethod
    def metric(rep_slot) -> "OLib":
        s = np.array([rep_slot], dtype=SLOT_DTYPE)
        return OLib(lib().orc_lib_metric(_ptr(s)))

    @staticmethod
    def custom_sparse(reps: np.ndarray, entries: dict) -> "OLib":
        reps = np.ascontiguousarray(reps, dtype=SLOT_DTYPE)
        shape = [int(d) for d in reps["dim"]]
        strides = [1] * len(shape)
        for i in range(len(shape) - 2, -1, -1):
            strides[i] = strides[i + 1] * shape[i + 1]
        flat = np.array([sum(i * s for i, s in zip(idx, strides)) for idx in entries], dtype=np.uint64)
        vals = np.array(list(entries.values()), dtype=np.complex128)
        return OLib(lib().orc_lib_custom(_ptr(reps), len(reps), _ptr(flat), _ptr(vals.view(np.float64)), len(flat), 1))

    def instantiate(self, sl: np.ndarray) -> OTensor:
        sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
        return OTensor(lib().orc_lib_instantiate(self.h, _ptr(sl), len(sl)))


class ONet:
    """Expression-tree network executed like Network::execute::<Sequential, SmallestDegree>."""

    def __init__(self):
        self.h = lib().orc_net_new()
        self._keep = []

    def __del__(self):
        try:
            lib().orc_net_free(self.h)
        except Exception:
            pass

    def tensor(self, t: OTensor) -> int:
        return lib().orc_net_add_tensor(self.h, t.h)

    def scalar(self, s: complex) -> int:
        return lib().orc_net_add_scalar(self.h, float(np.real(s)), float(np.imag(s)))

    def lib_tensor(self, l: OLib, sl: np.ndarray) -> int:
        sl = np.ascontiguousarray(sl, dtype=SLOT_DTYPE)
        return lib().orc_net_add_lib(self.h, l.h, _ptr(sl), len(sl))

    def op(self, kind: int, children, power: int = 1) -> int:
        ch = np.ascontiguousarray(children, dtype=np.int32)
        return lib().orc_net_add_op(self.h, kind, _ptr(ch), len(ch), power)

    def product(self, *children):
        return self.op(N_PRODUCT, children)

    def sum(self, *children):
        return self.op(N_SUM, children)

    def neg(self, child):
        return self.op(N_NEG, [child])

    def set_root(self, node: int):
        lib().orc_net_set_root(self.h, node)

    def execute(self):
        rc = lib().orc_net_execute(self.h)
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())

    def result(self) -> OTensor:
        return OTensor(lib().orc_net_result(self.h))

    def log(self):
        buf = np.zeros(2 * 256, dtype=np.int32)
        n = lib().orc_net_log(self.h, _ptr(buf), 256)
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(min(n, 256))]

    def eval_batch(self, batched_nodes, arrays, out_size: int, n_threads: int = 1, want_out=True):
        """arrays[k]: complex128 [n_points, size_k] for batched_nodes[k]. -> (out [n_points,out_size], sum)"""
        n_points = arrays[0].shape[0]
        arrs = [np.ascontiguousarray(a, dtype=np.complex128) for a in arrays]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        nodes = np.ascontiguousarray(batched_nodes, dtype=np.int32)
        out = np.zeros((n_points, out_size), dtype=np.complex128) if want_out else None
        s = np.zeros(out_size, dtype=np.complex128)
        rc = lib().orc_net_eval_batch(self.h, _ptr(nodes), len(nodes), C.cast(ptrs, C.c_void_p), n_points,
                                      _ptr(out.view(np.float64)) if want_out else None, out_size, n_threads,
                                      _ptr(s.view(np.float64)))
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())
        return out, s


# ---- NetSpec (spenso_b200/workloads.py) -> oracle network ---------------------------------------------
def _spec_slots(sl) -> np.ndarray:
    return np.array([s.record() for s in sl], dtype=SLOT_DTYPE)


def net_from_spec(spec, arrays, point: int = 0):
    """Build the oracle network of a NetSpec with the batched leaves filled from arrays[input][point].

    arrays[k]: complex [n_points, size_k] in the CANONICAL order of the k-th batched leaf.
    Returns (ONet, batched_node_ids, arrays_per_node) — the last two feed ONet.eval_batch; re-indexed
    leaves reuse the array of the leaf they alias (only identical canonical permutations supported).
    """
    net = ONet()
    ids, node_input, input_idx = {}, {}, 0
    b_nodes, b_arrays = [], []
    batched_nodes = [i for i, n in enumerate(spec.nodes) if n[0] == "batched"]
    for k, node in enumerate(spec.nodes):
        kind = node[0]
        if kind in ("batched", "reindex"):
            if kind == "batched":
                inp = input_idx
                input_idx += 1
                sl = _spec_slots(node[1])
            else:
                inp = node_input[node[1]]
                sl = _spec_slots(node[2])
            canon, perm, _, _ = order_structure(sl) if len(sl) else (sl, [], 0, 0)
            src_sl = _spec_slots(spec.nodes[batched_nodes[inp]][1])
            _, src_perm, _, _ = order_structure(src_sl) if len(src_sl) else (src_sl, [], 0, 0)
            if list(perm) != list(src_perm):
                raise NotImplementedError("re-indexed leaf with a different canonical permutation")
            node_input[k] = inp
            t = OTensor.dense(canon, arrays[inp][point], canonicalize=False)
            net._keep.append(t)
            ids[k] = net.tensor(t)
            b_nodes.append(ids[k])
            b_arrays.append(np.ascontiguousarray(arrays[inp], dtype=np.complex128))   # f64 leaves: embedded (x, 0)
        elif kind == "function":
            # device-filled leaf: the oracle side fills it on the host (numpy, test infrastructure) and treats it
            # as one more per-point tensor — the reference fills such leaves with its host evaluators
            import _expr_eval
            from spenso_b200.expr import Param

            sl = _spec_slots(node[1])
            comps = node[2](lambda nidx, size: Param.vector(nidx, size))
            leaf_arrays = {nidx: arrays[node_input[nidx]] for nidx in node_input}
            vals = np.stack([_expr_eval.evaluate(c, leaf_arrays) for c in comps], axis=1)
            canon, perm, _, _ = order_structure(sl) if len(sl) else (sl, [], 0, 0)
            shape = [int(d) for d in sl["dim"]]
            vals = vals.reshape([vals.shape[0]] + shape).transpose([0] + [int(q) + 1 for q in perm])
            vals = np.ascontiguousarray(vals).reshape(vals.shape[0], -1)
            t = OTensor.dense(canon, vals[point], canonicalize=False)
            net._keep.append(t)
            ids[k] = net.tensor(t)
            b_nodes.append(ids[k])
            b_arrays.append(vals)
        elif kind == "lib":
            which, sl = node[1], _spec_slots(node[2])
            l = OLib.metric(sl[0]) if which == 6 else OLib.builtin(which)
            net._keep.append(l)
            ids[k] = net.lib_tensor(l, sl)
        elif kind == "custom":
            sl = _spec_slots(node[1])
            l = OLib.custom_sparse(sl, node[2])
            net._keep.append(l)
            ids[k] = net.lib_tensor(l, sl)
        elif kind == "scalar":
            ids[k] = net.scalar(node[1])
        else:
            ids[k] = net.op(node[1], [ids[c] for c in node[2]], node[3])
    net.set_root(ids[spec.root])
    net.spec_ids = ids
    return net, b_nodes, b_arrays


def oracle_schedule(spec, arrays=None):
    """The contraction order the oracle's restatement of SmallestDegree / SingleSmallestDegree (contract.rs:239-497)
    chooses for a NetSpec, as (n1, n2) pairs of SPEC node indices in execution order — the naming of the reference's
    graph surgery: the result of n1.contract(n2) is known by n1 afterwards.  This is what a Rust host would hand to
    spn_net_set_schedule."""
    if arrays is None:
        arrays = [np.ones((1, size), dtype=np.complex128) for _, size in spec.inputs]
    net, _, _ = net_from_spec(spec, arrays, 0)
    net.execute()
    back = {v: k for k, v in net.spec_ids.items()}
    return [(back[a], back[b]) for a, b in net.log()]


def eval_spec(spec, arrays, n_threads: int = 1, want_out: bool = True):
    """Oracle evaluation of a NetSpec at every point: (out [n_points, out_size] canonical, sum, out slots)."""
    probe, _, _ = net_from_spec(spec, arrays, 0)
    probe.execute()
    r = probe.result()
    out_size, out_slots = r.size, r.slots
    net, nodes, arrs = net_from_spec(spec, arrays, 0)
    out, s = net.eval_batch(nodes, arrs, out_size, n_threads=n_threads, want_out=want_out)
    return out, s, out_slots
