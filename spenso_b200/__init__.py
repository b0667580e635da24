"""spenso_b200 — B200-native engine for spenso's concrete-data contraction path.

Host-side mirror of the reference interface for this path (names follow the reference):

  Representation / Slot            spenso/src/structure/{representation,slot}.rs
  DenseTensor / SparseTensor       spenso/src/tensors/data/{dense,sparse}.rs   (one point: "shared")
  BatchTensor                      one DenseTensor per phase-space point, device resident
  Contract (a.contract(b))         spenso/src/contraction.rs:32-35, 126-227
  Network (+, *, execute, result)  spenso/src/network.rs:47-51, 1217-1241

Everything numeric happens in libspenso_b200.so (hand-written sm_100a kernels behind the C ABI of
include/spenso_b200.h).  This module only marshals arguments; numpy is used to reorder host data
into canonical slot order (index bookkeeping), never to compute a contraction.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import SLOT_DTYPE, ContractPlan, SpensoError, build_library, check, lib, node_id, ptr

__all__ = [
    "Representation", "Euclidean", "Minkowski", "Lorentz", "Bispinor", "SpinFundamental", "ColorAdjoint",
    "ColorFundamental", "ColorSextet", "Slot", "slots", "symbol_slot", "Context", "DenseTensor", "SparseTensor", "BatchTensor",
    "Network", "NetExpr", "PeerComm", "SpensoError", "canonicalize", "plan_contract", "contract_cuda_source", "build_library", "library_tensor",
    "F64", "C64", "POINT_MAJOR", "COMPONENT_MAJOR", "FUSED", "STEPWISE", "AUTO",
]

REP_DUMMY, REP_SELF_DUAL, REP_INLINE_METRIC, REP_DUALIZABLE = 0, 1, 2, 3
FIRST_BEFORE_SECOND, SECOND_BEFORE_FIRST, INTERLEAVED = 0, 1, 2
F64, C64 = 0, 1
POINT_MAJOR, COMPONENT_MAJOR = 0, 1
BATCHED, SHARED_DENSE, SHARED_SPARSE = 0, 1, 2
OP_PRODUCT, OP_SUM, OP_NEG, OP_POWER, OP_CONJ = 3, 4, 5, 6, 7
FUSED, STEPWISE, AUTO = 0, 1, 2
LIB_GAMMA_WEYL, LIB_GAMMA_DIRAC, LIB_GAMMA0_WEYL, LIB_GAMMA5_WEYL, LIB_PROJM_WEYL, LIB_PROJP_WEYL, LIB_METRIC = range(7)
LIB_GAMMA_CONJ_WEYL, LIB_GAMMA_ADJ_WEYL = 7, 8


# ---- index structure --------------------------------------------------------------------------------
class Representation:
    """LibraryRep with the ids the reference assigns after idenso's initialize()
    (spenso/src/structure/representation.rs:1010-1035, idenso/src/representations.rs:132-148)."""

    def __init__(self, name: str, kind: int, rep_id: int):
        self.name, self.kind, self.rep_id = name, kind, rep_id

    def dual(self) -> "Representation":
        if self.kind != REP_DUALIZABLE:
            return self
        return Representation(self.name + "_dual" if self.rep_id > 0 else self.name[:-5], self.kind, -self.rep_id)

    def new_slot(self, dim: int, aind: int) -> "Slot":
        return Slot(self, dim, aind)

    def __repr__(self):
        return self.name


Euclidean = Representation("euc", REP_SELF_DUAL, 0)
Bispinor = Representation("bis", REP_SELF_DUAL, 1)
ColorAdjoint = Representation("coad", REP_SELF_DUAL, 2)
Minkowski = Representation("mink", REP_INLINE_METRIC, 0)
Lorentz = Representation("lor", REP_DUALIZABLE, 1)
SpinFundamental = Representation("spf", REP_DUALIZABLE, 2)
ColorFundamental = Representation("cof", REP_DUALIZABLE, 3)
ColorSextet = Representation("cos", REP_DUALIZABLE, 4)


class Slot:
    """Slot<LibraryRep, AbstractIndex> (spenso/src/structure/slot.rs:53-56)."""

    def __init__(self, rep: Representation, dim: int, aind: int, aind_kind: int = 0):
        self.rep, self.dim, self.aind, self.aind_kind = rep, int(dim), int(aind), int(aind_kind)

    def dual(self) -> "Slot":
        return Slot(self.rep.dual(), self.dim, self.aind, self.aind_kind)

    def record(self):
        return (self.rep.kind, self.aind_kind, self.rep.rep_id, self.dim, self.aind)

    def __repr__(self):
        return f"{self.rep}{self.dim}|{self.aind}"


AIND_NORMAL, AIND_DUALIZE, AIND_DOUBLE, AIND_DUMMY, AIND_ADDED, AIND_SYMBOL = range(6)


def symbol_slot(rep: Representation, dim: int, name: str) -> Slot:
    """A slot whose abstract index is a symbol (AbstractIndex::Symbol, abstract_index.rs:267-294): the name is interned
    to an integer id (first-seen order, like symbol creation order in the reference)."""
    return Slot(rep, dim, int(lib().spn_symbol_intern(name.encode())), AIND_SYMBOL)


def slots(items: Iterable) -> np.ndarray:
    """Sequence of Slot (or raw 5-tuples / a SLOT_DTYPE array) -> SLOT_DTYPE array."""
    if isinstance(items, np.ndarray) and items.dtype == SLOT_DTYPE:
        return np.ascontiguousarray(items)
    return np.array([s.record() if isinstance(s, Slot) else tuple(s) for s in items], dtype=SLOT_DTYPE)


def canonicalize(sl) -> tuple:
    """PermutedStructure::from(Vec<Slot>) (ordered.rs:321-380): (canonical slots, perm, base_start,
    dual_start) with canonical[i] = given[perm[i]]."""
    sl = slots(sl)
    out = np.zeros(len(sl), dtype=SLOT_DTYPE)
    perm = np.zeros(len(sl), dtype=np.int32)
    bs, ds = C.c_int64(), C.c_int64()
    check(lib().spn_structure_canonicalize(ptr(sl), len(sl), ptr(out), ptr(perm), C.byref(bs), C.byref(ds)))
    return out, perm, bs.value, ds.value


def plan_contract(a, b) -> ContractPlan:
    """StructureContract::merge + stride plan for canonical structures a, b (host only)."""
    a, b = slots(a), slots(b)
    p = ContractPlan()
    check(lib().spn_plan_contract(ptr(a), len(a), ptr(b), len(b), C.byref(p)))
    return p


def contract_cuda_source(a, b, dtype_a: int = 1, layout_a: int = 0, dtype_b: int = 1, layout_b: int = 0,
                         sparse_entries: Optional[dict] = None, sparse_first: bool = True) -> str:
    """CUDA source of the kernel specialised for a.contract(b) (host only; compiled with NVRTC on the way).  "" = the
    contraction takes a generic kernel.  Two batched tensors by default; with sparse_entries ({index tuple in the
    caller's slot order: value}) `a` is a SHARED SPARSE tensor with those stored entries and `b` the batched partner
    (sparse_first=False: b.contract(a))."""
    a, b = slots(a), slots(b)
    ca = canonicalize(a) if len(a) else (a, [], -1, -1)
    cb = canonicalize(b)[0] if len(b) else b
    if sparse_entries is None:
        r = lib().spn_plan_contract_cuda_source(ptr(ca[0]), len(a), dtype_a, layout_a, ptr(cb), len(b), dtype_b, layout_b)
    else:
        st = _row_major_strides([int(d) for d in ca[0]["dim"]])
        flat = np.zeros(len(sparse_entries), dtype=np.uint64)
        vals = np.zeros(len(sparse_entries), dtype=np.complex128)
        for k, (idx, v) in enumerate(sparse_entries.items()):
            flat[k] = sum(int(idx[int(p)]) * s_ for p, s_ in zip(ca[1], st))
            vals[k] = v
        r = lib().spn_plan_contract_sparse_cuda_source(ptr(ca[0]), len(a), ptr(flat), ptr(vals), len(flat), ptr(cb), len(b),
                                                       dtype_b, layout_b, 1 if sparse_first else 0)
    if r is None:
        raise SpensoError(1, lib().spn_last_error().decode(errors="replace"))
    return r.decode()


def _row_major_strides(shape: Sequence[int]):
    st = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        st[i] = st[i + 1] * shape[i + 1]
    return st


# ---- context --------------------------------------------------------------------------------------------
class Context:
    """One CUDA device + stream.  Creation fails (NoDevice) without a usable GPU: no CPU fallback."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        h = C.c_void_p()
        check(lib().spn_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, cuda_stream: int):
        check(lib().spn_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        check(lib().spn_ctx_synchronize(self.h))

    def release_cached_memory(self):
        """Return the pooled memory of freed intermediates to the driver."""
        check(lib().spn_ctx_release_cached_memory(self.h))

    def set_reference_quirks(self, emulate_interleaved_pairing: bool):
        """Parity triage: reproduce MultiContractInterleaved's un-permuted pairing (SURVEY K-note 3)."""
        check(lib().spn_ctx_set_reference_quirks(self.h, int(emulate_interleaved_pairing)))

    @property
    def launch_count(self) -> int:
        return int(lib().spn_ctx_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(lib().spn_ctx_sm_count(self.h))

    def __del__(self):
        try:
            if self.h:
                lib().spn_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ---- tensors ----------------------------------------------------------------------------------------------
class _Tensor:
    """Handle to a device tensor; the Contract trait and the element-wise algebra of the reference."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    def __del__(self):
        try:
            if self.h:
                lib().spn_tensor_free(self.h)
                self.h = None
        except Exception:
            pass

    # inspection
    @property
    def order(self) -> int:
        return int(lib().spn_tensor_order(self.h))

    @property
    def size(self) -> int:
        return int(lib().spn_tensor_size(self.h))

    @property
    def n_points(self) -> int:
        return int(lib().spn_tensor_n_points(self.h))

    @property
    def storage(self) -> int:
        return int(lib().spn_tensor_storage(self.h))

    @property
    def dtype(self) -> int:
        return int(lib().spn_tensor_dtype(self.h))

    @property
    def layout(self) -> int:
        return int(lib().spn_tensor_layout(self.h))

    @property
    def is_sparse(self) -> bool:
        return self.storage == SHARED_SPARSE

    @property
    def device_ptr(self) -> int:
        return int(lib().spn_tensor_device_ptr(self.h) or 0)

    @property
    def slots(self) -> np.ndarray:
        out = np.zeros(self.order, dtype=SLOT_DTYPE)
        check(lib().spn_tensor_slots(self.h, ptr(out)))
        return out

    @property
    def shape(self):
        return tuple(int(d) for d in self.slots["dim"])

    def to_numpy(self, first_point: int = 0, n_points: Optional[int] = None) -> np.ndarray:
        """Host copy in canonical order: [*shape] for shared tensors, [n_points, *shape] for batched."""
        batched = self.storage == BATCHED
        n = (self.n_points - first_point) if n_points is None else n_points
        if not batched:
            n = 1
        out = np.zeros((n, self.size), dtype=np.complex128 if self.dtype == C64 else np.float64)
        check(lib().spn_tensor_download(self.h, ptr(out), first_point, n))
        return out.reshape((n,) + self.shape) if batched else out.reshape(self.shape)

    def entries(self) -> dict:
        """Sparse tensors: {canonical flat index: value} (SparseTensor.elements)."""
        n = int(lib().spn_tensor_nnz(self.h))
        flat = np.zeros(n, dtype=np.uint64)
        vals = np.zeros(n, dtype=np.complex128)
        check(lib().spn_tensor_sparse_entries(self.h, ptr(flat), ptr(vals)))
        return {int(f): complex(v) for f, v in zip(flat, vals)}

    # algebra
    def _wrap(self, h) -> "_Tensor":
        t = _Tensor(self.ctx, h)
        st = t.storage
        t.__class__ = BatchTensor if st == BATCHED else (SparseTensor if st == SHARED_SPARSE else DenseTensor)
        return t

    def contract(self, other: "_Tensor") -> "_Tensor":
        out = C.c_void_p()
        check(lib().spn_tensor_contract(self.ctx.h, self.h, other.h, C.byref(out)))
        return self._wrap(out)

    def __add__(self, other):
        out = C.c_void_p()
        check(lib().spn_tensor_add(self.ctx.h, self.h, other.h, 0, C.byref(out)))
        return self._wrap(out)

    def __sub__(self, other):
        out = C.c_void_p()
        check(lib().spn_tensor_add(self.ctx.h, self.h, other.h, 1, C.byref(out)))
        return self._wrap(out)

    def add_fallible(self, other: "_Tensor") -> "_Tensor":
        """FallibleAdd (algebra/fallible_add.rs): like `+`, but Sparse + Sparse does not store exact zeros."""
        out = C.c_void_p()
        check(lib().spn_tensor_add(self.ctx.h, self.h, other.h, 2, C.byref(out)))
        return self._wrap(out)

    def sub_fallible(self, other: "_Tensor") -> "_Tensor":
        """FallibleSub (algebra/fallible_sub.rs)."""
        out = C.c_void_p()
        check(lib().spn_tensor_add(self.ctx.h, self.h, other.h, 3, C.byref(out)))
        return self._wrap(out)

    def __neg__(self):
        out = C.c_void_p()
        check(lib().spn_tensor_neg(self.ctx.h, self.h, C.byref(out)))
        return self._wrap(out)

    def scalar_mul(self, s: complex):
        out = C.c_void_p()
        check(lib().spn_tensor_scalar_mul(self.ctx.h, self.h, float(np.real(s)), float(np.imag(s)), C.byref(out)))
        return self._wrap(out)

    def conj(self):
        out = C.c_void_p()
        check(lib().spn_tensor_conj(self.ctx.h, self.h, C.byref(out)))
        return self._wrap(out)

    def recip(self):
        """element-wise 1/x (`ref_one() / s` of the Power op on scalars)"""
        out = C.c_void_p()
        check(lib().spn_tensor_recip(self.ctx.h, self.h, C.byref(out)))
        return self._wrap(out)

    def clone(self):
        out = C.c_void_p()
        check(lib().spn_tensor_clone(self.ctx.h, self.h, C.byref(out)))
        return self._wrap(out)

    def relayout(self, layout: int):
        """The same batched tensor in the other device layout (device-side transpose)."""
        out = C.c_void_p()
        check(lib().spn_tensor_relayout(self.ctx.h, self.h, layout, C.byref(out)))
        return self._wrap(out)

    def to_dense(self):
        """DataTensor::to_dense (tensors/data/sparse.rs:688-708)."""
        out = C.c_void_p()
        check(lib().spn_tensor_to_dense(self.ctx.h, self.h, C.byref(out)))
        return self._wrap(out)

    def to_sparse(self):
        """DenseTensor::to_sparse (tensors/data/dense.rs:523-535): values equal to zero are not stored."""
        out = C.c_void_p()
        check(lib().spn_tensor_to_sparse(self.ctx.h, self.h, C.byref(out)))
        return self._wrap(out)

    def get(self, flat: int, point: int = 0) -> complex:
        v = np.zeros(2)
        check(lib().spn_tensor_get(self.h, point, flat, ptr(v)))
        return complex(v[0], v[1])

    def set(self, flat: int, value: complex, point: int = 0):
        check(lib().spn_tensor_set(self.h, point, flat, float(np.real(value)), float(np.imag(value))))

    def trace(self, slots_with_repeats) -> "_Tensor":
        """Trace::internal_contract over one repeated (slot, dual) pair; `slots_with_repeats` is this
        tensor's index list as the caller wrote it (data row major over it)."""
        sl = slots(slots_with_repeats)
        out = C.c_void_p()
        check(lib().spn_tensor_trace(self.ctx.h, ptr(sl), len(sl), self.h, C.byref(out)))
        return self._wrap(out)


def _canonical_dense(sl, data, batch: bool):
    """Reorder host data given in the caller's slot order into canonical order (permute_inds)."""
    sl = slots(sl)
    canon, perm, _, _ = canonicalize(sl) if len(sl) else (sl, np.zeros(0, np.int32), -1, -1)
    shape = tuple(int(d) for d in sl["dim"])
    arr = np.asarray(data)
    if batch:
        arr = arr.reshape((arr.shape[0],) + shape)
        arr = np.transpose(arr, axes=[0] + [int(p) + 1 for p in perm])
    else:
        arr = arr.reshape(shape)
        arr = np.transpose(arr, axes=[int(p) for p in perm])
    return canon, np.ascontiguousarray(arr)


def library_tensor(ctx: Context, which: int, sl) -> "_Tensor":
    """A built-in hep_lib tensor with concrete indices (spn_tensor_library): shared sparse, canonical order."""
    s = slots(sl)
    out = C.c_void_p()
    check(lib().spn_tensor_library(ctx.h, which, ptr(s), len(s), C.byref(out)))
    return _Tensor(ctx, None)._wrap(out)


class DenseTensor(_Tensor):
    """DenseTensor<Complex<f64>> at ONE point ("shared"): spenso/src/tensors/data/dense.rs:51-54."""

    @staticmethod
    def from_data(ctx: Context, sl, data) -> "DenseTensor":
        canon, arr = _canonical_dense(sl, np.asarray(data, dtype=np.complex128), batch=False)
        h = C.c_void_p()
        check(lib().spn_tensor_shared_dense(ctx.h, ptr(canon), len(canon), ptr(arr), C.byref(h)))
        t = DenseTensor(ctx, h)
        return t


class SparseTensor(_Tensor):
    """SparseTensor<Complex<f64>> at one point: spenso/src/tensors/data/sparse.rs:48-53."""

    @staticmethod
    def from_entries(ctx: Context, sl, entries: dict) -> "SparseTensor":
        """entries: {index tuple in the caller's slot order: value}."""
        sl = slots(sl)
        canon, perm, _, _ = canonicalize(sl) if len(sl) else (sl, [], -1, -1)
        st = _row_major_strides([int(d) for d in canon["dim"]])
        flat = np.zeros(len(entries), dtype=np.uint64)
        vals = np.zeros(len(entries), dtype=np.complex128)
        for k, (idx, v) in enumerate(entries.items()):
            flat[k] = sum(int(idx[int(p)]) * s for p, s in zip(perm, st))
            vals[k] = v
        h = C.c_void_p()
        check(lib().spn_tensor_shared_sparse(ctx.h, ptr(canon), len(canon), ptr(flat), ptr(vals), len(entries),
                                             C.byref(h)))
        return SparseTensor(ctx, h)


class BatchTensor(_Tensor):
    """n_points DenseTensors (one per phase-space point) in one device buffer."""

    @staticmethod
    def empty(ctx: Context, sl, n_points: int, dtype: int = C64, layout: int = POINT_MAJOR) -> "BatchTensor":
        canon = canonicalize(sl)[0] if len(sl) else slots([])
        h = C.c_void_p()
        check(lib().spn_tensor_batched(ctx.h, ptr(canon), len(canon), n_points, dtype, layout, C.byref(h)))
        return BatchTensor(ctx, h)

    @staticmethod
    def wrap(ctx: Context, sl, n_points: int, device_ptr: int, dtype: int = C64,
             layout: int = POINT_MAJOR, keepalive=None) -> "BatchTensor":
        """View caller-owned device memory (e.g. a torch tensor's data_ptr()) holding canonical data."""
        canon = canonicalize(sl)[0] if len(sl) else slots([])
        h = C.c_void_p()
        check(lib().spn_tensor_batched_wrap(ctx.h, ptr(canon), len(canon), n_points, dtype, layout,
                                            C.c_void_p(device_ptr), C.byref(h)))
        t = BatchTensor(ctx, h)
        t._keepalive = keepalive
        return t

    @staticmethod
    def from_numpy(ctx: Context, sl, data, dtype: int = C64, layout: int = POINT_MAJOR) -> "BatchTensor":
        """data: [n_points, *dims] in the caller's slot order."""
        npdt = np.complex128 if dtype == C64 else np.float64
        canon, arr = _canonical_dense(sl, np.asarray(data, dtype=npdt), batch=True)
        t = BatchTensor.empty(ctx, canon, arr.shape[0], dtype, layout)
        t.upload(arr)
        return t

    def upload(self, canonical: np.ndarray, first_point: int = 0, sync: bool = True):
        """canonical: [n, size] host data in canonical order (pinned memory for async uploads)."""
        a = np.ascontiguousarray(canonical, dtype=np.complex128 if self.dtype == C64 else np.float64)
        n = a.shape[0] if a.ndim > 0 else 1
        fn = lib().spn_tensor_upload if sync else lib().spn_tensor_upload_async
        check(fn(self.h, ptr(a), first_point, n))

    def upload_ptr(self, host_ptr: int, first_point: int, n_points: int, sync: bool = False):
        fn = lib().spn_tensor_upload if sync else lib().spn_tensor_upload_async
        check(fn(self.h, C.c_void_p(host_ptr), first_point, n_points))

    def sum_points(self) -> np.ndarray:
        """Monte-Carlo partial sum over the point axis (device reduction)."""
        import torch

        out = torch.empty(self.size * 2, dtype=torch.float64, device=f"cuda:{self.ctx.device}")
        check(lib().spn_tensor_reduce_points(self.ctx.h, self.h, C.c_void_p(out.data_ptr())))
        self.ctx.synchronize()
        return out.cpu().numpy().view(np.complex128).reshape(self.shape)


class PeerComm:
    """Monte-Carlo sum exchange between the GPUs of one node over NVLink peer memory (csrc/comm.cu,
    include/spenso_b200.h spn_comm_*): one single-CTA kernel, no library collective, bit-identical sums on
    every rank.  `exchange(handle_bytes) -> [handle_bytes of every rank, in rank order]` is any host-side
    all-gather (spenso_b200.sharding.connect_peers uses torch.distributed)."""

    HANDLE_BYTES = 64

    def __init__(self, ctx: Context, rank: int, world: int, exchange=None, max_f64: int = 512):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        check(lib().spn_comm_create(ctx.h, rank, world, max_f64, C.byref(h)))
        self.h = h
        if world > 1:
            mine = np.zeros(self.HANDLE_BYTES, dtype=np.uint8)
            check(lib().spn_comm_handle(self.h, ptr(mine)))
            everyone = exchange(mine.tobytes())
            if len(everyone) != world or any(len(b) != self.HANDLE_BYTES for b in everyone):
                raise ValueError("exchange() must return one 64-byte handle per rank, in rank order")
            table = np.frombuffer(b"".join(everyone), dtype=np.uint8).copy()
            check(lib().spn_comm_connect(self.h, ptr(table)))

    def allreduce_sum(self, device_ptr: int, n_f64: int):
        """in place on the context's stream: device f64[n] := sum over ranks"""
        check(lib().spn_comm_allreduce_sum(self.h, C.c_void_p(device_ptr), n_f64))

    @property
    def status(self) -> int:
        return int(lib().spn_comm_status(self.h))

    def close(self, barrier=None):
        """Two-phase teardown (an exported mailbox must outlive its peers' mappings): barrier, unmap the peers,
        barrier, free.  `barrier` is the host barrier of the job (e.g. torch.distributed.barrier); a world of one
        needs none."""
        if getattr(self, "h", None):
            if barrier is not None:
                barrier()
            check(lib().spn_comm_disconnect(self.h))
            if barrier is not None:
                barrier()
            lib().spn_comm_destroy(self.h)
            self.h = None


# ---- network ------------------------------------------------------------------------------------------------
class NetExpr:
    """A node of a Network with the reference's operator syntax (*, +, -, unary -, ** n, .conj())."""

    def __init__(self, net: "Network", node: int):
        self.net, self.node = net, int(node)

    def __int__(self):
        return self.node

    __index__ = __int__

    def _lift(self, o) -> int:
        if isinstance(o, NetExpr):
            return o.node
        if isinstance(o, (int, float, complex)) and not isinstance(o, bool):
            return self.net.scalar(complex(o))
        raise TypeError(f"cannot combine a network expression with {type(o).__name__}")

    def __mul__(self, o):
        return NetExpr(self.net, self.net.product(self.node, self._lift(o)))

    __rmul__ = __mul__

    def __add__(self, o):
        return NetExpr(self.net, self.net.sum(self.node, self._lift(o)))

    __radd__ = __add__

    def __neg__(self):
        return NetExpr(self.net, self.net.neg(self.node))

    def __sub__(self, o):
        return self + (-(o if isinstance(o, NetExpr) else NetExpr(self.net, self._lift(o))))

    def __pow__(self, n: int):
        return NetExpr(self.net, self.net.power(self.node, int(n)))

    def conj(self):
        return NetExpr(self.net, self.net.conj(self.node))


class Network:
    """Expression graph of Sum/Neg/Product/Power/Function(conj) ops over tensor, library and scalar
    leaves (spenso/src/network.rs:47-51, network/graph.rs).  compile() resolves index matching,
    contraction order and stride plans once; execute() evaluates the network at every point."""

    def __init__(self, ctx: Optional[Context], _handle=None):
        self.ctx = ctx
        if _handle is None:
            _handle = C.c_void_p()
            check(lib().spn_net_create(ctx.h if ctx is not None else None, C.byref(_handle)))
        self.h = _handle
        self._keep = []

    def to_json(self) -> str:
        """The expression graph in the plan interchange format (spn_net_to_json)."""
        return lib().spn_net_to_json(self.h).decode()

    @staticmethod
    def from_json(ctx: Optional[Context], text: str) -> "Network":
        """Uncompiled network from the plan interchange format (spn_net_from_json)."""
        h = C.c_void_p()
        check(lib().spn_net_from_json(ctx.h if ctx is not None else None, text.encode(), C.byref(h)))
        return Network(ctx, _handle=h)

    def __del__(self):
        try:
            if self.h:
                lib().spn_net_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # leaves
    def batched(self, sl, dtype: int = C64) -> int:
        s = slots(sl)
        return node_id(lib().spn_net_leaf_batched(self.h, ptr(s), len(s), dtype))

    def reindex(self, leaf: int, sl) -> int:
        """The same per-point tensor under other index labels (no new input)."""
        s = slots(sl)
        return node_id(lib().spn_net_leaf_reindex(self.h, leaf, ptr(s), len(s)))

    def function(self, sl, components) -> int:
        """Per-point leaf computed on the device: `components` are spenso_b200.expr expressions, one per
        component in row-major order over `sl` as written (spn_net_leaf_function)."""
        from .expr import compile_components

        s = slots(sl)
        params, consts, code = compile_components(components)
        pa = np.ascontiguousarray(params, dtype=np.int32).reshape(-1)
        ca = np.ascontiguousarray(consts, dtype=np.complex128)
        co = np.ascontiguousarray(code, dtype=np.int32).reshape(-1)
        return node_id(lib().spn_net_leaf_function(self.h, ptr(s), len(s), ptr(pa), len(params), ptr(ca), len(consts),
                                                   ptr(co), len(code)))

    def tensor(self, t: _Tensor) -> int:
        """Network::from_tensor for a point-independent local tensor."""
        self._keep.append(t)
        return node_id(lib().spn_net_leaf_shared(self.h, t.h))

    def library(self, which: int, sl) -> int:
        s = slots(sl)
        return node_id(lib().spn_net_leaf_library(self.h, which, ptr(s), len(s)))

    def library_custom(self, sl, entries: dict) -> int:
        """User library tensor: {index tuple in argument order: value}."""
        s = slots(sl)
        st = _row_major_strides([int(d) for d in s["dim"]])
        flat = np.array([sum(int(i) * k for i, k in zip(idx, st)) for idx in entries], dtype=np.uint64)
        vals = np.array(list(entries.values()), dtype=np.complex128)
        return node_id(lib().spn_net_leaf_library_custom(self.h, ptr(s), len(s), ptr(flat), ptr(vals), len(flat)))

    def scalar(self, c: complex) -> int:
        return node_id(lib().spn_net_leaf_scalar(self.h, float(np.real(c)), float(np.imag(c))))

    # ops
    def op(self, kind: int, children, power: int = 1) -> int:
        ch = np.ascontiguousarray(list(children), dtype=np.int32)
        return node_id(lib().spn_net_op(self.h, kind, ptr(ch), len(ch), power))

    def product(self, *children) -> int:
        return self.op(OP_PRODUCT, children)

    def sum(self, *children) -> int:
        return self.op(OP_SUM, children)

    def neg(self, child) -> int:
        return self.op(OP_NEG, [child])

    def power(self, child, n: int) -> int:
        return self.op(OP_POWER, [child], n)

    def conj(self, child) -> int:
        return self.op(OP_CONJ, [child])

    def set_root(self, node):
        check(lib().spn_net_set_root(self.h, int(node)))

    def set_schedule(self, pairs):
        """The host's contraction order: [(n1, n2), ...] in the reference's naming — the result of n1.contract(n2) is
        known by the id of n1 afterwards (spn_net_set_schedule; SingleSmallestDegree::contract, contract.rs:346-497)."""
        buf = np.ascontiguousarray([int(x) for p in pairs for x in p], dtype=np.int32)
        check(lib().spn_net_set_schedule(self.h, ptr(buf) if len(buf) else None, len(buf) // 2))

    # expression style of the reference: Network::from_tensor(a) * Network::from_tensor(b) + ... (network.rs:335-351, 585-607)
    def from_tensor(self, t, sl=None, dtype: int = C64) -> "NetExpr":
        """A leaf as an expression node: a shared tensor (DenseTensor / SparseTensor), or — given slots — a per-point
        input (the k-th call with slots declares the k-th input of execute())."""
        if sl is not None:
            return NetExpr(self, self.batched(sl, dtype))
        return NetExpr(self, self.tensor(t))

    def expr(self, node: int) -> "NetExpr":
        return NetExpr(self, node)

    def set_comm(self, comm: Optional["PeerComm"]):
        """With a communicator attached, execute(sum_out_ptr=...) yields the sum over the points of ALL ranks
        (fused fold + NVLink exchange kernel)."""
        check(lib().spn_net_set_comm(self.h, comm.h if comm is not None else None))
        self._comm = comm

    def compile(self, mode: int = FUSED):
        check(lib().spn_net_compile(self.h, mode))
        return self

    # compiled-plan inspection
    @property
    def exec_mode(self) -> int:
        return int(lib().spn_net_exec_mode(self.h))

    @property
    def result_slots(self) -> np.ndarray:
        out = np.zeros(int(lib().spn_net_result_order(self.h)), dtype=SLOT_DTYPE)
        check(lib().spn_net_result_slots(self.h, ptr(out)))
        return out

    @property
    def result_size(self) -> int:
        return int(lib().spn_net_result_size(self.h))

    @property
    def n_inputs(self) -> int:
        return int(lib().spn_net_n_inputs(self.h))

    @property
    def schedule(self):
        n = lib().spn_net_schedule(self.h, None, 0)
        buf = np.zeros(2 * max(n, 1), dtype=np.int32)
        lib().spn_net_schedule(self.h, ptr(buf), n)
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)]

    @property
    def cuda_source(self) -> str:
        return (lib().spn_net_cuda_source(self.h) or b"").decode()

    def prepare_variant(self, key: str) -> str:
        """JIT + cache one kernel variant ahead of time (see spn_net_prepare_variant); returns its source."""
        src = lib().spn_net_prepare_variant(self.h, key.encode())
        if src is None:
            raise SpensoError(8, lib().spn_last_error().decode(errors="replace"))
        return src.decode()

    @property
    def fp64_instr_per_point(self) -> int:
        return int(lib().spn_net_macs_per_point(self.h))

    def variant_fp64_instr(self, key: str) -> int:
        """FP64 instructions per point of one kernel variant, the accumulation of a Monte-Carlo sum included."""
        return int(lib().spn_net_variant_fp64_instr(self.h, key.encode()))

    @property
    def input_bytes_per_point(self) -> int:
        return int(lib().spn_net_input_bytes_per_point(self.h))

    def execute(self, inputs: Sequence[BatchTensor], out: Optional[BatchTensor] = None, sum_out_ptr: int = 0,
                first_point: int = 0, n_points: int = 0):
        """Network::execute for all points (asynchronous on the context's stream)."""
        arr = (C.c_void_p * max(len(inputs), 1))(*[t.h for t in inputs])
        check(lib().spn_net_execute(self.h, C.cast(arr, C.c_void_p), len(inputs), out.h if out is not None else None,
                                    C.c_void_p(sum_out_ptr) if sum_out_ptr else None, first_point, n_points))

    def execute_host_async(self, host_inputs: Sequence[np.ndarray], out: Optional[np.ndarray] = None,
                           sum_out: Optional[np.ndarray] = None, chunk_points: int = 0):
        """spn_net_execute_host_async: enqueue one batch and return; call host_wait() before reading `out` / `sum_out`
        (a preallocated complex128 [result_size]) or touching the inputs.  Successive calls keep the pipeline full."""
        for a in host_inputs:
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError("asynchronous execution needs contiguous input arrays (no temporary copies)")
        n = host_inputs[0].shape[0] if host_inputs else (out.shape[0] if out is not None else 1)
        ptrs = (C.c_void_p * max(len(host_inputs), 1))(*[a.ctypes.data for a in host_inputs])
        check(lib().spn_net_execute_host_async(self.h, C.cast(ptrs, C.c_void_p), len(host_inputs), n,
                                               ptr(out) if out is not None else None,
                                               ptr(sum_out) if sum_out is not None else None, chunk_points))

    def host_wait(self):
        check(lib().spn_net_host_wait(self.h))

    def execute_host(self, host_inputs: Sequence[np.ndarray], out: Optional[np.ndarray] = None,
                     want_sum: bool = False, chunk_points: int = 0):
        """Network::execute on host data (spn_net_execute_host): host_inputs[i] is [n_points, size_i] complex128
        (float64 for f64 leaves) in canonical order; `out` (optional) a preallocated [n_points, result_size]
        complex128 array; returns (out, sum or None).  Copies overlap the kernels chunk by chunk; pass pinned
        arrays (e.g. views of torch pinned tensors) for full overlap."""
        ins = [np.ascontiguousarray(a) for a in host_inputs]
        n = ins[0].shape[0] if ins else (out.shape[0] if out is not None else 1)
        ptrs = (C.c_void_p * max(len(ins), 1))(*[a.ctypes.data for a in ins])
        s = np.zeros(self.result_size, dtype=np.complex128) if want_sum else None
        check(lib().spn_net_execute_host(self.h, C.cast(ptrs, C.c_void_p), len(ins), n,
                                         ptr(out) if out is not None else None, ptr(s) if want_sum else None,
                                         chunk_points))
        return out, s

    def evaluate(self, inputs: Sequence[BatchTensor], layout: int = POINT_MAJOR) -> BatchTensor:
        """Convenience: allocate the result tensor, execute, return it."""
        n = inputs[0].n_points if inputs else 1
        out = BatchTensor.empty(self.ctx, self.result_slots, n, C64, layout)
        self.execute(inputs, out)
        return out
