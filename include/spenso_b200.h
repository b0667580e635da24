/* spenso_b200 — C ABI of the B200-native contraction engine.
 *
 * This is the drop-in boundary for spenso's concrete-data path.  The reference is Rust
 * with no FFI on this path; a Rust shim implements
 *   trait Contract            spenso/src/contraction.rs:32-35
 *   Network::execute / ExecuteOp for NetworkStore   spenso/src/network.rs:1197-1241, 1244-1706
 *   DataTensor / DenseTensor / SparseTensor storage spenso/src/tensors/data.rs:190-193,
 *                                                    data/dense.rs:51-54, data/sparse.rs:48-53
 * for a device-backed batched tensor by calling the entry points below.
 *
 * Conventions: plain pointers and sizes, opaque handles, int status returns
 * (0 = ok), thread-local last-error string; no exceptions cross the boundary.  All
 * device work is enqueued on the context's CUDA stream.  Complex data is interleaved
 * (re, im) f64 — bit-compatible with spenso's Complex<f64>
 * (spenso/src/algebra/complex.rs:55-58).  There is NO CPU execution path: every
 * compute entry point fails with SPN_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef SPENSO_B200_H
#define SPENSO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPN_VERSION 100
#define SPN_MAX_ORDER 16            /* max indices per operand */
#define SPN_MAX_OUT_ORDER 32        /* max indices of a pairwise result */

/* ---- status codes (ContractionError, spenso/src/contraction.rs:20-30) ---------- */
enum {
  SPN_OK = 0,
  SPN_ERR_EMPTY_SPARSE = 1, /* ContractionError::EmptySparse                        */
  SPN_ERR_STRUCTURE = 2,    /* ContractionError::StructureError / merge invariants  */
  SPN_ERR_DIMENSION = 3,    /* ContractionError::DimensionError                     */
  SPN_ERR_OTHER = 4,        /* ContractionError::Other                              */
  SPN_ERR_CUDA = 5,
  SPN_ERR_NO_DEVICE = 6,
  SPN_ERR_INVALID = 7,
  SPN_ERR_JIT = 8
};

/* ---- index structure (R1-R3) ---------------------------------------------------- */
/* LibraryRep, spenso/src/structure/representation.rs:583-620 */
enum { SPN_REP_DUMMY = 0, SPN_REP_SELF_DUAL = 1, SPN_REP_INLINE_METRIC = 2, SPN_REP_DUALIZABLE = 3 };
/* AbstractIndex, spenso/src/structure/abstract_index.rs:267-294 */
enum { SPN_AIND_NORMAL = 0, SPN_AIND_DUALIZE = 1, SPN_AIND_DOUBLE = 2, SPN_AIND_DUMMY = 3, SPN_AIND_ADDED = 4,
       /* AbstractIndex::Symbol (feature "shadowing"): the payload is the interned id of the symbol.  The derived Ord
        * compares symbols by their id in symbolica's state, i.e. by creation order; spn_symbol_intern hands out ids in
        * first-seen order, which reproduces that order when the host interns its symbols in the order it created them
        * (or the host passes symbolica's own ids). */
       SPN_AIND_SYMBOL = 5 };

/* Slot<LibraryRep, AbstractIndex>, spenso/src/structure/slot.rs:53-56 (16 bytes) */
typedef struct spn_slot {
  uint8_t rep_kind;  /* SPN_REP_*                                                    */
  uint8_t aind_kind; /* SPN_AIND_*                                                   */
  int16_t rep_id;    /* SelfDual(n)/InlineMetric(n): n; Dualizable(k): +k base, -k dual */
  uint32_t dim;      /* Dimension::Concrete                                          */
  uint64_t aind;     /* index payload; Double(a,b) packed (a<<16)|b                  */
} spn_slot;

/* MergeInfo, spenso/src/structure.rs:716-751 */
enum { SPN_FIRST_BEFORE_SECOND = 0, SPN_SECOND_BEFORE_FIRST = 1, SPN_INTERLEAVED = 2 };

/* Result of StructureContract::merge (spenso/src/structure/ordered.rs:594-919) plus the
 * stride plan the kernels replay.  Plain data: comparable byte-for-byte.
 *   out[o] = sum_k  sign(k) * A[offA(o) + kA(k)] * B[offB(o) + kB(k)]
 *   o  = row-major index over out_dim[];  offX(o) = sum_d idx_d * out_stride_x[d]
 *   k  = row-major index over k_dim[] (contracted dims in OTHER's order, last fastest)
 *   sign(k) = XOR over contracted dims with k_metric[d] of (k_d > 0)      (Minkowski) */
typedef struct spn_contract_plan {
  uint32_t order_a, order_b, order_out, n_common;
  uint32_t merge_kind;  /* SPN_FIRST_BEFORE_SECOND | SPN_SECOND_BEFORE_FIRST | SPN_INTERLEAVED */
  uint32_t n_partition; /* == order_out unless an operand is scalar                     */
  int64_t base_start, dual_start; /* of the result structure; -1 = none                */
  uint64_t size_a, size_b, size_out, n_contracted;
  spn_slot out_slots[SPN_MAX_OUT_ORDER];
  uint8_t common_a[SPN_MAX_ORDER];
  uint8_t common_b[SPN_MAX_ORDER];
  uint8_t partition[SPN_MAX_OUT_ORDER]; /* 1 = slot comes from a (self)                 */
  uint32_t out_dim[SPN_MAX_OUT_ORDER];
  uint64_t out_stride_a[SPN_MAX_OUT_ORDER];
  uint64_t out_stride_b[SPN_MAX_OUT_ORDER];
  uint32_t k_dim[SPN_MAX_ORDER];
  uint64_t k_stride_a[SPN_MAX_ORDER];
  uint64_t k_stride_b[SPN_MAX_ORDER];
  uint8_t k_metric[SPN_MAX_ORDER];
  uint8_t k_pos_a[SPN_MAX_ORDER]; /* position in a of the d-th contracted pair          */
  uint8_t k_pos_b[SPN_MAX_ORDER]; /* position in b                                     */
} spn_contract_plan;

const char* spn_last_error(void);
int spn_version(void);
/* process-wide symbol table for SPN_AIND_SYMBOL payloads: the same name always maps to the same id, ids grow in
 * first-seen order from 0.  Thread-safe. */
uint64_t spn_symbol_intern(const char* name);
/* the name behind an id (NULL if unknown); valid for the life of the process */
const char* spn_symbol_name(uint64_t id);

/* PermutedStructure::from(Vec<Slot>) — ordered.rs:321-380.  out[i] = in[perm[i]]. */
int spn_structure_canonicalize(const spn_slot* in, uint32_t n, spn_slot* out, int32_t* perm,
                               int64_t* base_start, int64_t* dual_start);
/* merge + MergeInfo + match_indices + strides — ordered.rs:594-919, structure.rs:435-460, 566-577.
 * a and b must be canonically ordered. */
int spn_plan_contract(const spn_slot* a, uint32_t na, const spn_slot* b, uint32_t nb, spn_contract_plan* out);

/* Inspection (host only, no device): the CUDA source of the kernel specialised for a.contract(b) of two BATCHED
 * tensors with these canonical structures (physics-sized tensors; compiled with NVRTC on the way, so code-generation
 * errors surface here).  Empty string: the contraction takes a generic kernel.  NULL: error.  The string is owned by
 * the library and valid until the next call on this thread. */
const char* spn_plan_contract_cuda_source(const spn_slot* a, uint32_t na, int dtype_a, int layout_a, const spn_slot* b,
                                          uint32_t nb, int dtype_b, int layout_b);

/* The same for a SHARED SPARSE tensor (n_entries stored values: flat index, (re, im) pairs) against a batched one;
 * sparse_first != 0: the sparse tensor is the receiver a of a.contract(b).  The stored entries appear as literals in
 * the code. */
const char* spn_plan_contract_sparse_cuda_source(const spn_slot* sparse, uint32_t n_sparse, const uint64_t* flat,
                                                 const double* values, uint64_t n_entries, const spn_slot* batched,
                                                 uint32_t n_batched, int dtype, int layout, int sparse_first);

/* ---- context --------------------------------------------------------------------- */
typedef struct spn_ctx spn_ctx;
int spn_ctx_create(int device, spn_ctx** out);
void spn_ctx_destroy(spn_ctx* ctx);
/* use an existing cudaStream_t (e.g. torch's current stream); NULL = the legacy default stream */
int spn_ctx_set_stream(spn_ctx* ctx, void* cuda_stream);
int spn_ctx_synchronize(spn_ctx* ctx);
/* Results of the pairwise ops are allocated stream-ordered from the device's default memory pool, which keeps freed
 * intermediates cached for reuse (no release threshold: a finite one makes the driver unmap and remap result buffers
 * at every synchronisation once the pool is large); this returns the cache to the driver (synchronises the stream). */
int spn_ctx_release_cached_memory(spn_ctx* ctx);
/* Parity triage switch.  emulate_interleaved_pairing != 0: contractions whose result interleaves the operands'
 * slots (MergeInfo::Interleaved) and that contract >= 2 indices enumerate each operand's contracted indices in its
 * OWN slot order, like MultiContractInterleaved of the reference (spenso/src/contraction/multicontract.rs:295-689,
 * no permutation applied); default 0 pairs every contracted slot with its dual.  The two differ only when
 * dualizable contracted slots appear in different relative orders in the two operands. */
int spn_ctx_set_reference_quirks(spn_ctx* ctx, int emulate_interleaved_pairing);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t spn_ctx_launch_count(const spn_ctx* ctx);
int spn_ctx_sm_count(const spn_ctx* ctx);
/* spn_tensor_contract keeps everything it derives from a pair of structures (merge plan, enumeration tables, device
 * tables, kernel choice) in a per-context cache — the reference re-derives it on every call (contraction.rs:126-207).
 * Counters of that cache (any pointer may be NULL). */
int spn_ctx_plan_cache_stats(const spn_ctx* ctx, uint64_t* hits, uint64_t* misses, uint64_t* entries);

/* ---- tensors (S1) ---------------------------------------------------------------- */
enum { SPN_F64 = 0, SPN_C64 = 1 };
/* device layout of a batched tensor: [point][flat] (what n reference DenseTensors laid end to
 * end look like) or [flat][point] */
enum { SPN_POINT_MAJOR = 0, SPN_COMPONENT_MAJOR = 1 };
enum { SPN_STORAGE_BATCHED = 0, SPN_STORAGE_SHARED_DENSE = 1, SPN_STORAGE_SHARED_SPARSE = 2 };

typedef struct spn_tensor spn_tensor;

/* Batched dense tensor: one DenseTensor<f64|Complex<f64>> per phase-space point, device resident.
 * slots must be canonically ordered (use spn_structure_canonicalize). */
int spn_tensor_batched(spn_ctx* ctx, const spn_slot* slots, uint32_t n, uint64_t n_points, int dtype, int layout,
                       spn_tensor** out);
/* same, over device memory owned by the caller (e.g. a torch tensor) */
int spn_tensor_batched_wrap(spn_ctx* ctx, const spn_slot* slots, uint32_t n, uint64_t n_points, int dtype,
                            int layout, void* device_ptr, spn_tensor** out);
/* Shared (point-independent) tensors: DenseTensor / SparseTensor of Complex<f64>.  Host data in
 * canonical row-major order; sparse entries keyed by canonical flat index. */
int spn_tensor_shared_dense(spn_ctx* ctx, const spn_slot* slots, uint32_t n, const double* reim, spn_tensor** out);
int spn_tensor_shared_sparse(spn_ctx* ctx, const spn_slot* slots, uint32_t n, const uint64_t* flat,
                             const double* reim, uint64_t nnz, spn_tensor** out);
/* A built-in library tensor with concrete indices in argument order (SPN_LIB_*, see below), as a shared sparse tensor
 * in canonical slot order: get_lib_data + permute_inds (spenso/src/network/graph.rs:291-322, tensors/data/dense.rs:93-109)
 * — what `T::from(lib_tensor_with_indices)` is in the blanket ExecuteOp impl (network.rs:1311-1318, 1437-1462). */
int spn_tensor_library(spn_ctx* ctx, int which, const spn_slot* slots, uint32_t n, spn_tensor** out);
void spn_tensor_free(spn_tensor* t);

uint32_t spn_tensor_order(const spn_tensor* t);
int spn_tensor_slots(const spn_tensor* t, spn_slot* out);
uint64_t spn_tensor_size(const spn_tensor* t);     /* elements per point */
uint64_t spn_tensor_n_points(const spn_tensor* t); /* 1 for shared */
int spn_tensor_storage(const spn_tensor* t);
int spn_tensor_dtype(const spn_tensor* t);
int spn_tensor_layout(const spn_tensor* t);
uint64_t spn_tensor_nnz(const spn_tensor* t);
void* spn_tensor_device_ptr(const spn_tensor* t);

/* host <-> device; host data is [point][flat] interleaved re/im regardless of the device layout */
int spn_tensor_upload(spn_tensor* t, const void* host, uint64_t first_point, uint64_t n_points);
/* same, without the final stream synchronisation: `host` must stay valid (and should be pinned)
 * until the context's stream has passed this point */
int spn_tensor_upload_async(spn_tensor* t, const void* host, uint64_t first_point, uint64_t n_points);
int spn_tensor_download(const spn_tensor* t, void* host, uint64_t first_point, uint64_t n_points);
/* sparse shared tensors: entries sorted by flat index */
int spn_tensor_sparse_entries(const spn_tensor* t, uint64_t* flat, double* reim);

/* ---- pairwise ops (the Contract trait and friends) -------------------------------- */
/* a.contract(&b): contraction.rs:126-227.  Result storage: batched dense if either operand is
 * batched; else shared — Dense unless both are Sparse (contraction.rs:219-224). */
int spn_tensor_contract(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, spn_tensor** out);
/* element-wise: algebra/add_assign.rs, sub_assign.rs, neg.rs, scalar_mul.rs.
 * `subtract` is a bit set: SPN_ADD_SUBTRACT = a - b; SPN_ADD_FALLIBLE = the semantics of add_fallible / sub_fallible
 * (algebra/fallible_add.rs:105-135, fallible_sub.rs:112-150): a Sparse (+/-) Sparse result does not store values
 * that are exactly zero (smart_set, tensors/data/sparse.rs:712-724), where `+=` keeps them as explicit zeros. */
enum { SPN_ADD_SUBTRACT = 1, SPN_ADD_FALLIBLE = 2 };
int spn_tensor_add(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, int subtract, spn_tensor** out);
int spn_tensor_neg(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out);
int spn_tensor_scalar_mul(spn_ctx* ctx, const spn_tensor* a, double re, double im, spn_tensor** out);
int spn_tensor_conj(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out);
/* element-wise 1/x: `Sc::ref_one() / s` of the Power op on scalars (network.rs:1526-1703) */
int spn_tensor_recip(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out);
/* the same batched tensor in the other device layout (SPN_POINT_MAJOR <-> SPN_COMPONENT_MAJOR), converted on the device.
 * A network whose shared partners are sparse reads only the components that meet a non-zero: in component-major layout
 * those are contiguous streams (cfg3: 5.8e9 evals/s), in point-major layout 16-byte gathers whose 128-byte lines are
 * fetched whole (7.9e8) — converting once pays off from the third evaluation on. */
int spn_tensor_relayout(spn_ctx* ctx, const spn_tensor* a, int layout, spn_tensor** out);
/* deep copy (trait Clone of the reference's tensors) */
int spn_tensor_clone(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out);
/* DataTensor::to_dense / to_sparse (tensors/data/sparse.rs:688-708, dense.rs:523-535): to_sparse stores the values
 * that differ from zero; batched tensors are dense (to_dense copies, to_sparse is SPN_ERR_INVALID). */
int spn_tensor_to_dense(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out);
int spn_tensor_to_sparse(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out);
/* element access by canonical flat index (get_ref_linear / set_flat).  get on an absent sparse entry is
 * SPN_ERR_EMPTY_SPARSE; set on a sparse tensor stores the value (zeros too).  Synchronous; reim = {re, im}. */
int spn_tensor_get(const spn_tensor* t, uint64_t point, uint64_t flat, double* reim);
int spn_tensor_set(spn_tensor* t, uint64_t point, uint64_t flat, double re, double im);
/* Trace::internal_contract — contraction/trace.rs:12-83; slots may contain one repeated pair per call */
int spn_tensor_trace(spn_ctx* ctx, const spn_slot* slots_with_repeats, uint32_t n, const spn_tensor* a,
                     spn_tensor** out);
/* sum over the point axis (Monte-Carlo partial sum); out_device receives 2*size f64 on device */
int spn_tensor_reduce_points(spn_ctx* ctx, const spn_tensor* a, double* out_device);

/* ---- networks (N1) --------------------------------------------------------------- */
/* NetworkOp, spenso/src/network/graph.rs (Sum/Neg/Product/Power/Function) */
enum { SPN_OP_PRODUCT = 3, SPN_OP_SUM = 4, SPN_OP_NEG = 5, SPN_OP_POWER = 6, SPN_OP_CONJ = 7 };
/* built-in library tensors, spenso-hep-lib/src/lib.rs:27-387 (argument order bis i, bis j[, mink mu]) */
enum {
  SPN_LIB_GAMMA_WEYL = 0, SPN_LIB_GAMMA_DIRAC = 1, SPN_LIB_GAMMA0_WEYL = 2, SPN_LIB_GAMMA5_WEYL = 3,
  SPN_LIB_PROJM_WEYL = 4, SPN_LIB_PROJP_WEYL = 5, SPN_LIB_METRIC = 6, SPN_LIB_GAMMA_CONJ_WEYL = 7, SPN_LIB_GAMMA_ADJ_WEYL = 8
};
enum { SPN_EXEC_FUSED = 0 /* one JIT-specialised sm_100a kernel per network */,
       SPN_EXEC_STEPWISE = 1 /* one pairwise kernel per contraction, intermediates in HBM */,
       SPN_EXEC_AUTO = 2 /* fused unless the network's intermediates are too large for straight-line code */ };

typedef struct spn_net spn_net;
int spn_net_create(spn_ctx* ctx, spn_net** out);
void spn_net_destroy(spn_net* net);
/* leaves: return node ids (>= 0) or a negative status */
/* per-point input tensor; slots in the caller's argument order (canonicalised internally, the
 * bound data must be in CANONICAL order, like a reference DenseTensor) */
int spn_net_leaf_batched(spn_net* net, const spn_slot* slots, uint32_t n, int dtype);
/* the SAME per-point tensor under other index labels (what the reference gets by instantiating one
 * shadowed tensor, e.g. a momentum K, with different abstract indices): `leaf` is a node returned by
 * spn_net_leaf_batched; slots are in the same argument order, same representations and dimensions.
 * No new input is created — both leaves read the input bound to the original. */
int spn_net_leaf_reindex(spn_net* net, int leaf, const spn_slot* slots, uint32_t n);
/* Leaf filling ON THE DEVICE: a per-point tensor whose components are small expressions of
 * components of other per-point leaves — the role of the reference's EvalTensor / TensorEvaluator, which
 * evaluates parametric leaves on the host for every point (spenso/src/tensors/parametric.rs:2266-2720,
 * network/symbolica_interop.rs:617-640).  The expressions are inlined into the fused kernel, so only the raw
 * kinematics cross PCIe and HBM.
 *   params: n_params pairs (batched leaf node id, canonical flat component of that leaf)
 *   consts: n_consts complex constants (re, im)
 *   code:   n_code pairs (opcode, argument), postfix; one program per component in ARGUMENT-order row major,
 *           separated by (SPN_EXPR_END, 0).  SPN_EXPR_SQRT needs an argument whose imaginary part is structurally
 *           zero (use SPN_EXPR_RE) and non-negative at run time.
 * In a stepwise network a function leaf is materialised by a small fused kernel of its own, then used pairwise. */
enum {
  SPN_EXPR_PARAM = 0, SPN_EXPR_CONST = 1, SPN_EXPR_ADD = 2, SPN_EXPR_SUB = 3, SPN_EXPR_MUL = 4, SPN_EXPR_DIV = 5,
  SPN_EXPR_NEG = 6, SPN_EXPR_CONJ = 7, SPN_EXPR_SQRT = 8, SPN_EXPR_RE = 9, SPN_EXPR_IM = 10, SPN_EXPR_END = 11
};
int spn_net_leaf_function(spn_net* net, const spn_slot* slots, uint32_t n, const int32_t* params, uint32_t n_params,
                          const double* consts_reim, uint32_t n_consts, const int32_t* code, uint32_t n_code);
/* point-independent local tensor (Network::from_tensor): an existing shared tensor */
int spn_net_leaf_shared(spn_net* net, const spn_tensor* shared);
/* library tensor with concrete indices in argument order (NetworkLeaf::LibraryKey + get_lib_data,
 * network/graph.rs:291-322).  For SPN_LIB_METRIC pass the two slots. */
int spn_net_leaf_library(spn_net* net, int which, const spn_slot* slots, uint32_t n);
/* user library tensor: sparse entries keyed by ARGUMENT-order flat index */
int spn_net_leaf_library_custom(spn_net* net, const spn_slot* slots, uint32_t n, const uint64_t* flat,
                                const double* reim, uint64_t nnz);
int spn_net_leaf_scalar(spn_net* net, double re, double im);
int spn_net_op(spn_net* net, int op, const int32_t* children, uint32_t n_children, int32_t power);
int spn_net_set_root(spn_net* net, int node);
/* The HOST's contraction order (north_star: "the Rust host resolves ... contraction order ... on the CPU exactly as
 * today").  pairs[2k], pairs[2k+1] = (n1, n2) of the k-th pairwise contraction in the reference's own naming: a factor
 * of a Product is known by its node id, and the result of n1.contract(n2) takes the place and the id of n1 — the graph
 * surgery of SingleSmallestDegree::contract (spenso/src/network/contract.rs:346-497: replace n1 by the result leaf,
 * delete n2).  A (LibraryKey, LocalTensor) pair runs as local.contract(lib), like contract.rs:428-443.  Entries are
 * global and ordered; each must name two current factors of one Product when its turn comes (compile fails with
 * SPN_ERR_INVALID otherwise).  Factors a schedule leaves over (scalars, which ContractScalars folds without a log
 * entry) are multiplied in by the default rule.  With a schedule set the fused compiler does not search for a cheaper
 * order.  n_pairs = 0 removes the schedule.  spn_net_schedule returns the order of a compiled network in the same
 * naming, so that one network's schedule can be replayed on another. */
int spn_net_set_schedule(spn_net* net, const int32_t* pairs, uint32_t n_pairs);
/* Plan interchange: the expression graph as JSON text (format: csrc/net.cu).  A host that owns
 * the real spenso Network (graph + store derive serde, spenso/src/network.rs:33-51, network/graph.rs:39-58) can
 * emit this text instead of making one call per node.  to_json returns a string owned by net (valid until the
 * next call); from_json creates an UNCOMPILED network (ctx may be NULL for a plan-only network). */
const char* spn_net_to_json(spn_net* net);
int spn_net_from_json(spn_ctx* ctx, const char* json, spn_net** out);
/* Run the graph walk, merge and schedule ONCE (they are identical for every point), then build
 * the device program. */
int spn_net_compile(spn_net* net, int exec_mode);
/* the mode the compiled network runs in (resolves SPN_EXEC_AUTO) */
int spn_net_exec_mode(const spn_net* net);
/* result structure (canonical) — valid after compile */
uint32_t spn_net_result_order(const spn_net* net);
int spn_net_result_slots(const spn_net* net, spn_slot* out);
uint64_t spn_net_result_size(const spn_net* net);
uint32_t spn_net_n_inputs(const spn_net* net);
/* number of pairwise contractions in the schedule and the (n1, n2) node-id pairs in execution order, in the naming of
 * spn_net_set_schedule (the result of a contraction is known by the id of n1) */
int spn_net_schedule(const spn_net* net, int32_t* pairs, uint32_t cap);
/* generated CUDA source of the fused kernel (NUL terminated; owned by net) */
const char* spn_net_cuda_source(const spn_net* net);
/* JIT (and cache) the kernel variant `key` ahead of the first execute and return its source; NULL on
 * error.  key = one char per input ('p' point major 32-byte aligned, 'q' point major, 'c' component
 * major), then the output ('n' none, 'p', 'c', optional 'r' = real output), then 's' if the sum over
 * points is wanted.  Used by build() to pre-compile the benchmark kernels where no GPU exists. */
const char* spn_net_prepare_variant(spn_net* net, const char* key);
/* FP64 arithmetic instructions (DMUL/DADD/DFMA) per point of the compiled fused program */
uint64_t spn_net_macs_per_point(const spn_net* net);
/* the same for one kernel variant (key as in spn_net_prepare_variant), including what the variant adds: a sum variant
 * ('s') accumulates every real result component once per point, except the components whose multiply-add chain is
 * accumulated link by link.  0 on error. */
uint64_t spn_net_variant_fp64_instr(spn_net* net, const char* key);
/* bytes of per-point input the fused kernel actually reads (components multiplied by structural
 * zeros of the shared tensors are never loaded) */
uint64_t spn_net_input_bytes_per_point(const spn_net* net);
/* Network::execute for all points: inputs[i] is the batched tensor bound to the i-th
 * spn_net_leaf_batched (creation order).  out (optional) receives the per-point result tensor;
 * sum_out_device (optional) receives sum over points (2*result_size f64, device memory).
 * n_points == 0: all points of inputs[0]. */
int spn_net_execute(spn_net* net, const spn_tensor* const* inputs, uint32_t n_inputs, spn_tensor* out,
                    double* sum_out_device, uint64_t first_point, uint64_t n_points);

/* Network::execute on HOST data — the reference's own calling convention (its tensors are host
 * Vec<Complex<f64>>, spenso/src/tensors/data/dense.rs:51-54).  host_inputs[i]: n_points tensors of input i laid end to
 * end (canonical order); host_out (optional): n_points result tensors; host_sum (optional): the sum over points
 * (2*result_size doubles, host).  The batch streams through a 3-deep chunk pipeline on private streams (H2D of
 * chunk k+1 || kernel of chunk k || D2H of chunk k-1), so device memory use does not depend on n_points.
 * chunk_points = 0: default (2^17).  Host buffers should be pinned (spn_host_alloc / spn_host_register).  Stepwise
 * networks run chunk by chunk on the context's stream (default chunk 2^14 points), without overlap.
 * Synchronous: host_out / host_sum are complete on return. */
int spn_net_execute_host(spn_net* net, const void* const* host_inputs, uint32_t n_inputs, uint64_t n_points,
                         void* host_out, double* host_sum, uint64_t chunk_points);
/* The same without the final wait: returns when the chunks are enqueued.  Successive calls keep the pipeline full —
 * the first chunks of call k+1 cross the bus while the last kernels and result copies of call k are still running (a
 * Monte-Carlo driver submits batch after batch) — and complete in order.  host_inputs must stay valid and unchanged,
 * host_out / host_sum must not be read, until spn_net_host_wait returns. */
int spn_net_execute_host_async(spn_net* net, const void* const* host_inputs, uint32_t n_inputs, uint64_t n_points,
                               void* host_out, double* host_sum, uint64_t chunk_points);
int spn_net_host_wait(spn_net* net);
/* ---- Monte-Carlo sum exchange between the GPUs of one node ----------------------------------
 * The one collective of the path: the all-reduce (SUM) of the per-rank Monte-Carlo partial sums,
 * 2*result_size doubles.  Nothing in the reference corresponds to it (it is single-process); its callers
 * add the per-point results on the host.  One process per GPU; every rank owns a mailbox in its HBM that
 * its peers write through NVLink peer mappings (CUDA IPC), so the exchange is ONE single-CTA kernel
 * (fold of the per-CTA partials + push + flag wait + rank-ordered sum) with no library collective.
 * Every rank obtains bit-identical sums (contributions are added in rank order).
 *   1. spn_comm_create on every rank; 2. publish spn_comm_handle (SPN_COMM_HANDLE_BYTES) to all ranks by any
 *   host channel (torch.distributed / MPI / a file); 3. spn_comm_connect with the world's handles in rank order.
 * Calls are collective: every rank must issue the same sequence.  They are enqueued on the context's stream
 * and are CUDA-graph capturable.  A peer that never arrives trips a device-side timeout (default 10 s,
 * SPN_COMM_TIMEOUT_S) instead of hanging the GPU: the sums of that exchange are poisoned with NaN, spn_comm_status
 * returns 1 + the rank that was missing, spn_ctx_synchronize and every later exchange on the communicator fail with
 * SPN_ERR_OTHER.  The status is sticky (the ranks' epochs are out of step): destroy and re-create the communicator
 * on every rank. */
typedef struct spn_comm spn_comm;
#define SPN_COMM_HANDLE_BYTES 64
int spn_comm_create(spn_ctx* ctx, int rank, int world, uint32_t max_f64, spn_comm** out);
int spn_comm_handle(const spn_comm* comm, void* handle_out);
int spn_comm_connect(spn_comm* comm, const void* all_handles);
/* in place: inout_device[0..n_f64) := sum over ranks */
int spn_comm_allreduce_sum(spn_comm* comm, double* inout_device, uint32_t n_f64);
/* 0 = ok; 1 + r = timed out waiting for rank r (host-visible without synchronising) */
int spn_comm_status(const spn_comm* comm);
/* Teardown is two-phase because an exported mailbox must outlive its peers' mappings:
 *   host barrier (nobody is inside an exchange) -> spn_comm_disconnect on every rank (unmaps the peers' mailboxes)
 *   -> host barrier -> spn_comm_destroy (frees this rank's mailbox; disconnects first if that was skipped). */
int spn_comm_disconnect(spn_comm* comm);
void spn_comm_destroy(spn_comm* comm);
/* Attach (or detach with NULL) a communicator to a compiled network: spn_net_execute's sum_out_device then
 * receives the sum over the points of ALL ranks, produced by the fused fold + exchange kernel (one launch
 * after the network kernel instead of a fold kernel plus a library all-reduce). */
int spn_net_set_comm(spn_net* net, spn_comm* comm);

/* pinned host memory for the host-buffer entry points (cudaHostAlloc / cudaFreeHost).  SPN_HOST_WRITE_COMBINED: for
 * buffers the CPU only ever WRITES sequentially (the kinematics handed to spn_net_execute_host): the copy engine reads
 * them without snooping the CPU caches; reading such memory from the CPU is very slow. */
enum { SPN_HOST_WRITE_COMBINED = 1 };
void* spn_host_alloc(uint64_t bytes);
void* spn_host_alloc_flags(uint64_t bytes, int flags);
/* Pin memory the host already owns (a Rust Vec<Complex<f64>>, a numpy array) in place, so that the host-buffer entry
 * points copy from / to it asynchronously without an intermediate pinned copy (cudaHostRegister / cudaHostUnregister).
 * Copies from unpinned memory still work, but synchronously and at a fraction of the link rate. */
int spn_host_register(void* p, uint64_t bytes);
int spn_host_unregister(void* p);
void spn_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SPENSO_B200_H */
