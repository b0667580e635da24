//! Raw declarations of the C ABI (include/spenso_b200.h).  One-to-one with the header; keep in sync.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct spn_slot {
    pub rep_kind: u8,
    pub aind_kind: u8,
    pub rep_id: i16,
    pub dim: u32,
    pub aind: u64,
}
#[repr(C)]
pub struct spn_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct spn_tensor {
    _p: [u8; 0],
}
#[repr(C)]
pub struct spn_net {
    _p: [u8; 0],
}
#[repr(C)]
pub struct spn_comm {
    _p: [u8; 0],
}

pub const SPN_OK: c_int = 0;
pub const SPN_ERR_EMPTY_SPARSE: c_int = 1;
pub const SPN_ERR_STRUCTURE: c_int = 2;
pub const SPN_ERR_DIMENSION: c_int = 3;
pub const SPN_REP_DUMMY: u8 = 0;
pub const SPN_REP_SELF_DUAL: u8 = 1;
pub const SPN_REP_INLINE_METRIC: u8 = 2;
pub const SPN_REP_DUALIZABLE: u8 = 3;
pub const SPN_F64: c_int = 0;
pub const SPN_C64: c_int = 1;
pub const SPN_POINT_MAJOR: c_int = 0;
pub const SPN_OP_PRODUCT: c_int = 3;
pub const SPN_OP_SUM: c_int = 4;
pub const SPN_OP_NEG: c_int = 5;
pub const SPN_OP_POWER: c_int = 6;
pub const SPN_OP_CONJ: c_int = 7;
pub const SPN_EXEC_FUSED: c_int = 0;
pub const SPN_EXEC_STEPWISE: c_int = 1;
pub const SPN_EXEC_AUTO: c_int = 2;

unsafe extern "C" {
    pub fn spn_tensor_recip(ctx: *mut spn_ctx, a: *const spn_tensor, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_clone(ctx: *mut spn_ctx, a: *const spn_tensor, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_conj(ctx: *mut spn_ctx, a: *const spn_tensor, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_relayout(ctx: *mut spn_ctx, a: *const spn_tensor, layout: c_int, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_host_register(p: *mut c_void, bytes: u64) -> c_int;
    pub fn spn_host_unregister(p: *mut c_void) -> c_int;
    pub fn spn_net_execute_host_async(net: *mut spn_net, host_inputs: *const *const c_void, n_inputs: u32, n_points: u64,
                                      host_out: *mut c_void, host_sum: *mut c_double, chunk_points: u64) -> c_int;
    pub fn spn_net_host_wait(net: *mut spn_net) -> c_int;
    pub fn spn_net_set_schedule(net: *mut spn_net, pairs: *const i32, n_pairs: u32) -> c_int;
    pub fn spn_last_error() -> *const c_char;
    pub fn spn_ctx_create(device: c_int, out: *mut *mut spn_ctx) -> c_int;
    pub fn spn_ctx_destroy(ctx: *mut spn_ctx);
    pub fn spn_ctx_synchronize(ctx: *mut spn_ctx) -> c_int;
    pub fn spn_tensor_batched(ctx: *mut spn_ctx, slots: *const spn_slot, n: u32, n_points: u64, dtype: c_int,
                              layout: c_int, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_shared_dense(ctx: *mut spn_ctx, slots: *const spn_slot, n: u32, reim: *const c_double,
                                   out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_shared_sparse(ctx: *mut spn_ctx, slots: *const spn_slot, n: u32, flat: *const u64,
                                    reim: *const c_double, nnz: u64, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_free(t: *mut spn_tensor);
    pub fn spn_tensor_order(t: *const spn_tensor) -> u32;
    pub fn spn_tensor_slots(t: *const spn_tensor, out: *mut spn_slot) -> c_int;
    pub fn spn_tensor_size(t: *const spn_tensor) -> u64;
    pub fn spn_tensor_n_points(t: *const spn_tensor) -> u64;
    pub fn spn_tensor_upload(t: *mut spn_tensor, host: *const c_void, first_point: u64, n_points: u64) -> c_int;
    pub fn spn_tensor_download(t: *const spn_tensor, host: *mut c_void, first_point: u64, n_points: u64) -> c_int;
    pub fn spn_tensor_contract(ctx: *mut spn_ctx, a: *const spn_tensor, b: *const spn_tensor,
                               out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_add(ctx: *mut spn_ctx, a: *const spn_tensor, b: *const spn_tensor, subtract: c_int,
                          out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_neg(ctx: *mut spn_ctx, a: *const spn_tensor, out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_tensor_scalar_mul(ctx: *mut spn_ctx, a: *const spn_tensor, re: c_double, im: c_double,
                                 out: *mut *mut spn_tensor) -> c_int;
    pub fn spn_net_create(ctx: *mut spn_ctx, out: *mut *mut spn_net) -> c_int;
    pub fn spn_net_destroy(net: *mut spn_net);
    pub fn spn_net_leaf_batched(net: *mut spn_net, slots: *const spn_slot, n: u32, dtype: c_int) -> c_int;
    pub fn spn_net_leaf_reindex(net: *mut spn_net, leaf: c_int, slots: *const spn_slot, n: u32) -> c_int;
    pub fn spn_net_leaf_shared(net: *mut spn_net, shared: *const spn_tensor) -> c_int;
    pub fn spn_net_leaf_library(net: *mut spn_net, which: c_int, slots: *const spn_slot, n: u32) -> c_int;
    pub fn spn_net_leaf_scalar(net: *mut spn_net, re: c_double, im: c_double) -> c_int;
    pub fn spn_net_op(net: *mut spn_net, op: c_int, children: *const i32, n_children: u32, power: i32) -> c_int;
    pub fn spn_net_set_root(net: *mut spn_net, node: c_int) -> c_int;
    pub fn spn_net_compile(net: *mut spn_net, exec_mode: c_int) -> c_int;
    pub fn spn_net_result_order(net: *const spn_net) -> u32;
    pub fn spn_net_result_slots(net: *const spn_net, out: *mut spn_slot) -> c_int;
    pub fn spn_net_result_size(net: *const spn_net) -> u64;
    pub fn spn_net_execute_host(net: *mut spn_net, host_inputs: *const *const c_void, n_inputs: u32, n_points: u64,
                                host_out: *mut c_void, host_sum: *mut c_double, chunk_points: u64) -> c_int;
    pub fn spn_host_alloc(bytes: u64) -> *mut c_void;
    pub fn spn_host_free(p: *mut c_void);
    // Monte-Carlo sum exchange between the GPUs of one node (NVLink peer memory, no collective library)
    pub fn spn_comm_create(ctx: *mut spn_ctx, rank: c_int, world: c_int, max_f64: u32, out: *mut *mut spn_comm) -> c_int;
    pub fn spn_comm_handle(comm: *const spn_comm, handle_out: *mut c_void) -> c_int;
    pub fn spn_comm_connect(comm: *mut spn_comm, all_handles: *const c_void) -> c_int;
    pub fn spn_comm_allreduce_sum(comm: *mut spn_comm, inout_device: *mut c_double, n_f64: u32) -> c_int;
    pub fn spn_comm_status(comm: *const spn_comm) -> c_int;
    pub fn spn_comm_disconnect(comm: *mut spn_comm) -> c_int;
    pub fn spn_comm_destroy(comm: *mut spn_comm);
    pub fn spn_net_set_comm(net: *mut spn_net, comm: *mut spn_comm) -> c_int;
}

pub const SPN_COMM_HANDLE_BYTES: usize = 64;
