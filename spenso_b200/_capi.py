"""ctypes binding of the C ABI (include/spenso_b200.h) — libspenso_b200.so.

The shared library is the product; this file only declares its entry points.  Nothing here
computes: when the library (or a CUDA device) is missing every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libspenso_b200.so")
CSRC = os.path.join(PKG_DIR, "csrc")

# spn_slot (16 bytes)
SLOT_DTYPE = np.dtype([("rep_kind", "u1"), ("aind_kind", "u1"), ("rep_id", "<i2"), ("dim", "<u4"), ("aind", "<u8")])
assert SLOT_DTYPE.itemsize == 16

SPN_MAX_ORDER, SPN_MAX_OUT_ORDER = 16, 32


class ContractPlan(C.Structure):
    """spn_contract_plan — plain data, comparable byte for byte."""

    _fields_ = [
        ("order_a", C.c_uint32), ("order_b", C.c_uint32), ("order_out", C.c_uint32), ("n_common", C.c_uint32),
        ("merge_kind", C.c_uint32), ("n_partition", C.c_uint32),
        ("base_start", C.c_int64), ("dual_start", C.c_int64),
        ("size_a", C.c_uint64), ("size_b", C.c_uint64), ("size_out", C.c_uint64), ("n_contracted", C.c_uint64),
        ("out_slots", C.c_uint8 * (16 * SPN_MAX_OUT_ORDER)),
        ("common_a", C.c_uint8 * SPN_MAX_ORDER), ("common_b", C.c_uint8 * SPN_MAX_ORDER),
        ("partition", C.c_uint8 * SPN_MAX_OUT_ORDER),
        ("out_dim", C.c_uint32 * SPN_MAX_OUT_ORDER),
        ("out_stride_a", C.c_uint64 * SPN_MAX_OUT_ORDER), ("out_stride_b", C.c_uint64 * SPN_MAX_OUT_ORDER),
        ("k_dim", C.c_uint32 * SPN_MAX_ORDER),
        ("k_stride_a", C.c_uint64 * SPN_MAX_ORDER), ("k_stride_b", C.c_uint64 * SPN_MAX_ORDER),
        ("k_metric", C.c_uint8 * SPN_MAX_ORDER), ("k_pos_a", C.c_uint8 * SPN_MAX_ORDER),
        ("k_pos_b", C.c_uint8 * SPN_MAX_ORDER),
    ]


# every symbol include/spenso_b200.h declares: name -> (restype, argtypes)
_P, _U32, _U64, _I, _D = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
SYMBOLS = {
    "spn_last_error": (C.c_char_p, []),
    "spn_version": (_I, []),
    "spn_symbol_intern": (_U64, [C.c_char_p]),
    "spn_symbol_name": (C.c_char_p, [_U64]),
    "spn_structure_canonicalize": (_I, [_P, _U32, _P, _P, _P, _P]),
    "spn_plan_contract": (_I, [_P, _U32, _P, _U32, C.POINTER(ContractPlan)]),
    "spn_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "spn_ctx_destroy": (None, [_P]),
    "spn_ctx_set_stream": (_I, [_P, _P]),
    "spn_ctx_synchronize": (_I, [_P]),
    "spn_ctx_set_reference_quirks": (_I, [_P, _I]),
    "spn_ctx_release_cached_memory": (_I, [_P]),
    "spn_ctx_launch_count": (_U64, [_P]),
    "spn_ctx_sm_count": (_I, [_P]),
    "spn_ctx_plan_cache_stats": (_I, [_P, _P, _P, _P]),
    "spn_plan_contract_cuda_source": (C.c_char_p, [_P, _U32, _I, _I, _P, _U32, _I, _I]),
    "spn_plan_contract_sparse_cuda_source": (C.c_char_p, [_P, _U32, _P, _P, _U64, _P, _U32, _I, _I, _I]),
    "spn_tensor_batched": (_I, [_P, _P, _U32, _U64, _I, _I, C.POINTER(_P)]),
    "spn_tensor_batched_wrap": (_I, [_P, _P, _U32, _U64, _I, _I, _P, C.POINTER(_P)]),
    "spn_tensor_shared_dense": (_I, [_P, _P, _U32, _P, C.POINTER(_P)]),
    "spn_tensor_shared_sparse": (_I, [_P, _P, _U32, _P, _P, _U64, C.POINTER(_P)]),
    "spn_tensor_free": (None, [_P]),
    "spn_tensor_order": (_U32, [_P]),
    "spn_tensor_slots": (_I, [_P, _P]),
    "spn_tensor_size": (_U64, [_P]),
    "spn_tensor_n_points": (_U64, [_P]),
    "spn_tensor_storage": (_I, [_P]),
    "spn_tensor_dtype": (_I, [_P]),
    "spn_tensor_layout": (_I, [_P]),
    "spn_tensor_nnz": (_U64, [_P]),
    "spn_tensor_device_ptr": (_P, [_P]),
    "spn_tensor_upload": (_I, [_P, _P, _U64, _U64]),
    "spn_tensor_upload_async": (_I, [_P, _P, _U64, _U64]),
    "spn_tensor_download": (_I, [_P, _P, _U64, _U64]),
    "spn_tensor_sparse_entries": (_I, [_P, _P, _P]),
    "spn_tensor_contract": (_I, [_P, _P, _P, C.POINTER(_P)]),
    "spn_tensor_add": (_I, [_P, _P, _P, _I, C.POINTER(_P)]),
    "spn_tensor_neg": (_I, [_P, _P, C.POINTER(_P)]),
    "spn_tensor_scalar_mul": (_I, [_P, _P, _D, _D, C.POINTER(_P)]),
    "spn_tensor_conj": (_I, [_P, _P, C.POINTER(_P)]),
    "spn_tensor_recip": (_I, [_P, _P, C.POINTER(_P)]),
    "spn_tensor_clone": (_I, [_P, _P, C.POINTER(_P)]),
    "spn_tensor_relayout": (_I, [_P, _P, _I, C.POINTER(_P)]),
    "spn_tensor_to_dense": (_I, [_P, _P, C.POINTER(_P)]),
    "spn_tensor_to_sparse": (_I, [_P, _P, C.POINTER(_P)]),
    "spn_tensor_get": (_I, [_P, _U64, _U64, _P]),
    "spn_tensor_set": (_I, [_P, _U64, _U64, _D, _D]),
    "spn_tensor_library": (_I, [_P, _I, _P, _U32, C.POINTER(_P)]),
    "spn_tensor_trace": (_I, [_P, _P, _U32, _P, C.POINTER(_P)]),
    "spn_tensor_reduce_points": (_I, [_P, _P, _P]),
    "spn_net_create": (_I, [_P, C.POINTER(_P)]),
    "spn_net_destroy": (None, [_P]),
    "spn_net_leaf_batched": (_I, [_P, _P, _U32, _I]),
    "spn_net_leaf_reindex": (_I, [_P, _I, _P, _U32]),
    "spn_net_leaf_function": (_I, [_P, _P, _U32, _P, _U32, _P, _U32, _P, _U32]),
    "spn_net_leaf_shared": (_I, [_P, _P]),
    "spn_net_leaf_library": (_I, [_P, _I, _P, _U32]),
    "spn_net_leaf_library_custom": (_I, [_P, _P, _U32, _P, _P, _U64]),
    "spn_net_leaf_scalar": (_I, [_P, _D, _D]),
    "spn_net_op": (_I, [_P, _I, _P, _U32, C.c_int32]),
    "spn_net_set_root": (_I, [_P, _I]),
    "spn_net_set_schedule": (_I, [_P, _P, _U32]),
    "spn_net_to_json": (C.c_char_p, [_P]),
    "spn_net_from_json": (_I, [_P, C.c_char_p, C.POINTER(_P)]),
    "spn_net_compile": (_I, [_P, _I]),
    "spn_net_exec_mode": (_I, [_P]),
    "spn_net_result_order": (_U32, [_P]),
    "spn_net_result_slots": (_I, [_P, _P]),
    "spn_net_result_size": (_U64, [_P]),
    "spn_net_n_inputs": (_U32, [_P]),
    "spn_net_schedule": (_I, [_P, _P, _U32]),
    "spn_net_cuda_source": (C.c_char_p, [_P]),
    "spn_net_prepare_variant": (C.c_char_p, [_P, C.c_char_p]),
    "spn_net_macs_per_point": (_U64, [_P]),
    "spn_net_variant_fp64_instr": (_U64, [_P, C.c_char_p]),
    "spn_net_input_bytes_per_point": (_U64, [_P]),
    "spn_net_execute": (_I, [_P, _P, _U32, _P, _P, _U64, _U64]),
    "spn_net_execute_host": (_I, [_P, _P, _U32, _U64, _P, _P, _U64]),
    "spn_net_execute_host_async": (_I, [_P, _P, _U32, _U64, _P, _P, _U64]),
    "spn_net_host_wait": (_I, [_P]),
    "spn_host_alloc": (_P, [_U64]),
    "spn_host_alloc_flags": (_P, [_U64, _I]),
    "spn_host_register": (_I, [_P, _U64]),
    "spn_host_unregister": (_I, [_P]),
    "spn_host_free": (None, [_P]),
    "spn_comm_create": (_I, [_P, _I, _I, _U32, C.POINTER(_P)]),
    "spn_comm_handle": (_I, [_P, _P]),
    "spn_comm_connect": (_I, [_P, _P]),
    "spn_comm_allreduce_sum": (_I, [_P, _P, _U32]),
    "spn_comm_status": (_I, [_P]),
    "spn_comm_disconnect": (_I, [_P]),
    "spn_comm_destroy": (None, [_P]),
    "spn_net_set_comm": (_I, [_P, _P]),
}

_lib = None


def build_library(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libspenso_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", "Makefile"))]
    srcs.append(os.path.join(ROOT, "include", "spenso_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(spenso_b200 has no Python or CPU execution path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class SpensoError(RuntimeError):
    """ContractionError / TensorNetworkError of the reference, carried as a status code."""

    NAMES = {1: "EmptySparse", 2: "StructureError", 3: "DimensionError", 4: "Other", 5: "Cuda", 6: "NoDevice",
             7: "Invalid", 8: "Jit"}

    def __init__(self, code: int, msg: str):
        super().__init__(f"{self.NAMES.get(code, code)}: {msg}")
        self.code = code


def check(rc: int):
    if rc != 0:
        raise SpensoError(abs(rc), lib().spn_last_error().decode(errors="replace"))
    return rc


def node_id(rc: int) -> int:
    if rc < 0:
        raise SpensoError(-rc, lib().spn_last_error().decode(errors="replace"))
    return rc


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)
