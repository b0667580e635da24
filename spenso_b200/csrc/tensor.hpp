// Device-resident tensors of the engine (the replacement for spenso's DataTensor storage,
// spenso/src/tensors/data.rs:190-193): batched dense buffers over the point axis, and shared
// (point-independent) dense / sparse tensors that also keep a host copy for plan building.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"
#include "structure.hpp"

namespace spn {
struct JitKernel;
struct ContractEntry;  // api.cu: everything spn_tensor_contract derives from a pair of structures, built once
// a specialised pairwise kernel (pairwise_jit.cu) with its launch shape
struct PairJit {
  std::shared_ptr<JitKernel> k;
  int block = 0;      // threads per CTA
  bool slab = false;  // staged: one warp per slab of 32 points (else one thread per point)
  uint32_t rows = 1;  // row-parallel kernels: threads per point (no grid-stride loop)
  uint32_t points_per_cta = 0;  // row-parallel kernels: whole points per CTA (the grid is ceil(n_points / this))
};
}

struct spn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  int sm_count = 148;
  int l2_fetch_granularity = 0;   // cudaLimitMaxL2FetchGranularity in effect when the context was created
  bool emulate_reference_pairing = false;  // SURVEY K-note 3, see build_k_tables
  // specialised pairwise kernels (pairwise_jit.cu), keyed by a hash of (plan, storage, dtype, layout)
  std::map<uint64_t, spn::PairJit> pair_kernels;
  // plan cache of the pairwise Contract path, keyed by the two operand structures (+ storage, dtype, layout,
  // alignment, identity of a shared operand): merge plan, enumeration tables, device tables, kernel choice.
  // The reference re-derives all of it on every a.contract(&b) (contraction.rs:126-207); here a repeated contraction
  // costs one hash lookup, one stream-ordered allocation and one launch.
  std::unordered_map<std::string, std::shared_ptr<spn::ContractEntry>> contract_cache;
  uint64_t contract_cache_hits = 0, contract_cache_misses = 0;
  // sticky status words (pinned host memory) of the communicators created on this context (comm.cu): non-zero after
  // an exchange timed out; checked by spn_ctx_synchronize
  std::vector<const int*> comm_status;
  // Result buffers of the pairwise ops, recycled by exact size.  Everything a context does is ordered on ONE stream, so
  // a buffer released by the host (its tensor handle was freed) can be handed to the next kernel on that stream at
  // once: a repeated a.contract(&b) then costs no allocator call at all, where even the stream-ordered pool costs two
  // (cudaMallocAsync + cudaFreeAsync) and a few microseconds between consecutive kernels.  Buffers go back to the
  // device's pool when the cache exceeds `block_cache_limit`, when the stream changes, on
  // spn_ctx_release_cached_memory and when the context dies.
  std::multimap<size_t, void*> block_cache;
  size_t block_cache_bytes = 0, block_cache_limit = (size_t)8 << 30;
  cudaStream_t block_cache_stream = nullptr;
  void* take_block(size_t bytes) {
    if (block_cache_stream != stream) return nullptr;
    auto it = block_cache.find(bytes);
    if (it == block_cache.end()) return nullptr;
    void* p = it->second;
    block_cache.erase(it);
    block_cache_bytes -= bytes;
    return p;
  }
  void give_block(void* p, size_t bytes, cudaStream_t allocated_on) {
    if (allocated_on != stream || bytes > block_cache_limit) {
      if (cudaFreeAsync(p, allocated_on) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
      }
      return;
    }
    if (block_cache_stream != stream) drop_blocks();
    block_cache_stream = stream;
    while (block_cache_bytes + bytes > block_cache_limit && !block_cache.empty()) {   // make room: largest first
      auto last = std::prev(block_cache.end());
      cudaFreeAsync(last->second, stream);
      block_cache_bytes -= last->first;
      block_cache.erase(last);
    }
    block_cache.emplace(bytes, p);
    block_cache_bytes += bytes;
  }
  void drop_blocks() {
    for (auto& kv : block_cache)
      if (cudaFreeAsync(kv.second, block_cache_stream) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(kv.second);
      }
    block_cache.clear();
    block_cache_bytes = 0;
  }
  ~spn_ctx() { drop_blocks(); }
};

struct spn_tensor {
  spn_ctx* ctx = nullptr;
  spn::Structure st;
  int storage = SPN_STORAGE_BATCHED;
  int dtype = SPN_C64;
  int layout = SPN_POINT_MAJOR;
  uint64_t n_points = 1;
  void* dptr = nullptr;  // batched: the buffer; shared: densified Complex<f64> copy
  bool owns = false;
  bool pooled = false;             // allocated with cudaMallocAsync on alloc_stream (results of pairwise ops)
  cudaStream_t alloc_stream = nullptr;
  size_t alloc_bytes = 0;          // pooled: size of the allocation (the context recycles result buffers by size)
  // shared tensors: host copies (canonical row major)
  std::vector<double2> host_dense;
  std::map<uint64_t, double2> entries;  // sparse only, keyed (and therefore sorted) by flat index
  uint64_t uid = 0;  // shared tensors: process-unique identity (they are immutable), part of the plan-cache key

  uint64_t size() const { return st.size(); }
  size_t elem_bytes() const { return dtype == SPN_C64 ? 16 : 8; }
  bool batched() const { return storage == SPN_STORAGE_BATCHED; }
  bool sparse() const { return storage == SPN_STORAGE_SHARED_SPARSE; }
  spn::DevOperand operand() const {
    spn::DevOperand o;
    o.ptr = dptr;
    o.is_complex = dtype == SPN_C64;
    if (!batched()) {
      o.ps = 0;
      o.fs = 1;
    } else if (layout == SPN_POINT_MAJOR) {
      o.ps = (long long)size();
      o.fs = 1;
    } else {
      o.ps = 1;
      o.fs = (long long)n_points;
    }
    return o;
  }
  spn::DevOut output() const {
    spn::DevOperand o = operand();
    spn::DevOut r;
    r.ptr = dptr;
    r.ps = o.ps;
    r.fs = o.fs;
    r.is_complex = o.is_complex;
    return r;
  }
  ~spn_tensor() {
    if (owns && dptr) {
      // stream-ordered free keeps the pairwise (stepwise) path free of device-wide synchronisation
      if (pooled && ctx) {
        ctx->give_block(dptr, alloc_bytes, alloc_stream);
      } else if (!pooled || cudaFreeAsync(dptr, alloc_stream) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(dptr);
      }
    }
  }
};

namespace spn {

void set_error(const std::string& s);
inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError)
                   ? SPN_ERR_NO_DEVICE
                   : SPN_ERR_CUDA;
    throw Error(code, std::string(what) + ": " + cudaGetErrorString(e));
  }
}

// SPN_TRACE=1: one line per dispatch decision on stderr (the counterpart of the reference's log::trace! calls in
// contraction.rs:157,174,192 and multicontract.rs:39,95,161,217)
inline bool trace_enabled() {
  static const bool on = getenv("SPN_TRACE") != nullptr;
  return on;
}
#define SPN_TRACE_LOG(...)                 \
  do {                                     \
    if (spn::trace_enabled()) {            \
      std::fprintf(stderr, "[spenso_b200] "); \
      std::fprintf(stderr, __VA_ARGS__);   \
      std::fprintf(stderr, "\n");          \
    }                                      \
  } while (0)

// host-side helpers shared by the pairwise API and the network compiler
struct KTables {
  std::vector<long long> off_a, off_b;  // element offsets: 64-bit, a tensor may hold more than 2^31 elements
  std::vector<unsigned char> sign;
};
KTables build_k_tables(const spn_contract_plan& p, bool emulate_reference_pairing = false);
DevOutMap out_map_of(const spn_contract_plan& p);

spn_tensor* new_shared_dense(spn_ctx* ctx, const Structure& st, const double2* data);
spn_tensor* new_shared_sparse(spn_ctx* ctx, const Structure& st, const std::map<uint64_t, double2>& entries);
spn_tensor* new_batched(spn_ctx* ctx, const Structure& st, uint64_t n_points, int dtype, int layout, void* wrap);

// pairwise_jit.cu: kernels specialised for one contraction of small per-point tensors
bool pair_jit_applicable(const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a, const spn_tensor* b,
                         bool res_batched, uint64_t n_points);
PairJit pair_jit_build(spn_ctx* ctx, const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a,
                       const spn_tensor* b, const spn_tensor* res);
void pair_jit_launch(spn_ctx* ctx, const PairJit& pj, const spn_tensor* a, const spn_tensor* b, spn_tensor* res,
                     uint64_t n_points);
std::string pair_jit_source(const Structure& sa, int dtype_a, int layout_a, const Structure& sb, int dtype_b, int layout_b);
std::string pair_jit_source_sparse(const Structure& ss, const std::map<uint64_t, double2>& entries, const Structure& sd,
                                   int dtype_d, int layout_d, bool sparse_first);

// comm.cu: fold the per-CTA partials (or take inout as the local sum when partial == nullptr) and exchange with
// the peers in one launch on the context's stream
void comm_exchange(spn_comm* c, const double2* partial, int n_blocks, double* inout_device, uint32_t n_f64);

// the Contract trait: result in *out (new tensor)
void tensor_contract(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, spn_tensor** out);
void tensor_elementwise(spn_ctx* ctx, int op, const spn_tensor* a, const spn_tensor* b, double2 s, spn_tensor** out,
                        bool drop_zeros = false);

}  // namespace spn
