// C ABI (include/spenso_b200.h): context, device tensors and the pairwise ops
// (Contract / add / sub / neg / scalar_mul / conj / trace / reduce).  Every compute entry point
// launches the sm_100a kernels of kernels.cu; there is no host execution path.
#include <atomic>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>

#include "library.hpp"
#include "tensor.hpp"

using namespace spn;

static thread_local std::string g_err;
void spn::set_error(const std::string& s) { g_err = s; }

#define API_TRY try {
#define API_CATCH                      \
  }                                    \
  catch (const Error& e) {             \
    g_err = e.what();                  \
    return e.code;                     \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return SPN_ERR_OTHER;              \
  }

namespace spn {

KTables build_k_tables(const spn_contract_plan& p, bool emulate_reference_pairing) {
  KTables t;
  uint32_t nd = p.n_common;
  uint64_t nk = p.n_contracted;
  if (nd == 0) nk = 1;
  t.off_a.resize(nk);
  t.off_b.resize(nk);
  t.sign.resize(nk);
  // enumeration order of the contracted dims per operand.  Normally both operands walk them in the SAME
  // order (other's), i.e. each slot is paired with its dual.  MultiContractInterleaved of the reference
  // (multicontract.rs:295-689) zips self.iter_metric() with other.iter(), each in its OWN slot order, without
  // the permutation the non-interleaved kernel applies (SURVEY K-note 3); `emulate_reference_pairing`
  // reproduces that for parity triage.
  std::vector<uint32_t> ord_a(nd), ord_b(nd);
  for (uint32_t d = 0; d < nd; ++d) ord_a[d] = ord_b[d] = d;
  if (emulate_reference_pairing && p.merge_kind == SPN_INTERLEAVED && nd >= 2) {
    std::stable_sort(ord_a.begin(), ord_a.end(), [&](uint32_t x, uint32_t y) { return p.k_pos_a[x] < p.k_pos_a[y]; });
    std::stable_sort(ord_b.begin(), ord_b.end(), [&](uint32_t x, uint32_t y) { return p.k_pos_b[x] < p.k_pos_b[y]; });
  }
  std::vector<uint32_t> ia(nd, 0), ib(nd, 0);  // odometers over ord_a / ord_b, last fastest
  for (uint64_t k = 0; k < nk; ++k) {
    long long oa = 0, ob = 0;
    bool neg = false;
    for (uint32_t d = 0; d < nd; ++d) {
      oa += (long long)ia[d] * (long long)p.k_stride_a[ord_a[d]];
      ob += (long long)ib[d] * (long long)p.k_stride_b[ord_b[d]];
      if (p.k_metric[ord_a[d]] && ia[d] > 0) neg = !neg;  // the sign comes from self's (a's) metric iterator
    }
    t.off_a[k] = oa;
    t.off_b[k] = ob;
    t.sign[k] = neg ? 1 : 0;
    for (uint32_t d = nd; d-- > 0;) {
      if (++ia[d] < p.k_dim[ord_a[d]]) break;
      ia[d] = 0;
    }
    for (uint32_t d = nd; d-- > 0;) {
      if (++ib[d] < p.k_dim[ord_b[d]]) break;
      ib[d] = 0;
    }
  }
  return t;
}

DevOutMap out_map_of(const spn_contract_plan& p) {
  DevOutMap m{};
  m.order = (int)p.order_out;
  for (uint32_t r = 0; r < p.order_out; ++r) {
    m.dim[r] = p.out_dim[r];
    m.sa[r] = (long long)p.out_stride_a[r];
    m.sb[r] = (long long)p.out_stride_b[r];
  }
  m.size_out = p.size_out;
  return m;
}

static void* dev_alloc(size_t bytes) {
  void* p = nullptr;
  cuda_check(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc");
  return p;
}

spn_tensor* new_batched(spn_ctx* ctx, const Structure& st, uint64_t n_points, int dtype, int layout, void* wrap) {
  auto t = std::make_unique<spn_tensor>();
  t->ctx = ctx;
  t->st = st;
  t->storage = SPN_STORAGE_BATCHED;
  t->dtype = dtype;
  t->layout = layout;
  t->n_points = n_points;
  if (wrap) {
    t->dptr = wrap;
    t->owns = false;
  } else {
    // stream-ordered allocation from the device's memory pool: no implicit synchronisation, and memory freed by
    // an earlier intermediate is reused without a round trip to the driver
    const size_t bytes = n_points * st.size() * t->elem_bytes();
    static const bool no_pool = getenv("SPN_NO_POOL") != nullptr;
    static const bool no_recycle = getenv("SPN_NO_BLOCK_CACHE") != nullptr;
    const size_t want = bytes ? bytes : 16;
    if (ctx && !no_pool && !no_recycle && (t->dptr = ctx->take_block(want)) != nullptr) {
      t->pooled = true;
      t->alloc_stream = ctx->stream;
      t->alloc_bytes = want;
    } else if (ctx && !no_pool && cudaMallocAsync(&t->dptr, want, ctx->stream) == cudaSuccess) {
      t->pooled = true;
      t->alloc_stream = ctx->stream;
      t->alloc_bytes = no_recycle ? (size_t)-1 : want;   // (size_t)-1: larger than any limit => always freed
    } else {
      cudaGetLastError();
      t->dptr = dev_alloc(bytes);
    }
    t->owns = true;
  }
  return t.release();
}

static std::atomic<uint64_t> g_next_uid{1};
static void upload_shared(spn_tensor* t) {
  t->uid = g_next_uid.fetch_add(1);
  t->dptr = dev_alloc(t->host_dense.size() * sizeof(double2));
  t->owns = true;
  cuda_check(cudaMemcpyAsync(t->dptr, t->host_dense.data(), t->host_dense.size() * sizeof(double2),
                             cudaMemcpyHostToDevice, t->ctx->stream),
             "upload shared tensor");
  cuda_check(cudaStreamSynchronize(t->ctx->stream), "upload shared tensor");
}

spn_tensor* new_shared_dense(spn_ctx* ctx, const Structure& st, const double2* data) {
  auto t = std::make_unique<spn_tensor>();
  t->ctx = ctx;
  t->st = st;
  t->storage = SPN_STORAGE_SHARED_DENSE;
  t->host_dense.assign(data, data + st.size());
  upload_shared(t.get());
  return t.release();
}
spn_tensor* new_shared_sparse(spn_ctx* ctx, const Structure& st, const std::map<uint64_t, double2>& entries) {
  auto t = std::make_unique<spn_tensor>();
  t->ctx = ctx;
  t->st = st;
  t->storage = SPN_STORAGE_SHARED_SPARSE;
  t->entries = entries;
  t->host_dense.assign(st.size(), make_double2(0.0, 0.0));
  for (auto& kv : entries) {
    if (kv.first >= st.size()) throw Error(SPN_ERR_INVALID, "sparse flat index out of bounds");
    t->host_dense[kv.first] = kv.second;
  }
  upload_shared(t.get());
  return t.release();
}

// a device table that lives as long as the plan-cache entry that owns it
template <class T>
struct DevTable {
  T* p = nullptr;
  size_t n = 0;
  DevTable() = default;
  DevTable(const DevTable&) = delete;
  DevTable& operator=(const DevTable&) = delete;
  void upload(const std::vector<T>& h) {
    n = h.size();
    cuda_check(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)), "cudaMalloc (plan table)");
    if (n) cuda_check(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice), "plan table upload");
  }
  ~DevTable() {
    if (p) cudaFree(p);
  }
};

// ContractionError::EmptySparse (singlecontract.rs:153-159 and the other sparse-receiver kernels):
// a sparse receiver with no stored entry facing a dense operand with no data.
static void check_empty_sparse(const spn_tensor* a, const spn_tensor* b, int merge_kind) {
  const spn_tensor* recv = merge_kind == SPN_SECOND_BEFORE_FIRST ? b : a;
  const spn_tensor* oth = merge_kind == SPN_SECOND_BEFORE_FIRST ? a : b;
  if (recv->sparse() && !oth->sparse() && recv->entries.empty() && oth->size() * oth->n_points == 0)
    throw Error(SPN_ERR_EMPTY_SPARSE, "Sparse tensor is empty");
}

// ---- K6 dispatch: does this dense x dense contraction reshape to a GEMM worth the tensor pipe? -------------
struct GemmShape {
  long long M = 1, N = 1;
};
static GemmShape gemm_shape(const spn_contract_plan& p) {
  GemmShape g;
  for (uint32_t r = 0; r < p.order_out; ++r) (p.partition[r] ? g.M : g.N) *= p.out_dim[r];
  return g;
}
static bool zgemm_applicable(const spn_contract_plan& p, const spn_tensor* a, const spn_tensor* b, int res_dtype,
                             long long res_fs, uint64_t n_points, const KTables& kt) {
  if (getenv("SPN_NO_ZGEMM")) return false;
  if (a->sparse() || b->sparse()) return false;
  // complex x complex -> ZGEMM, real x real -> DGEMM, real x complex -> DGEMM over the complex operand seen as a real
  // matrix with twice the columns (or rows): its (re, im) pairs are two adjacent real entries, and so are the result's
  if (a->dtype == b->dtype && a->dtype != res_dtype) return false;
  if (p.n_partition != p.order_out || p.n_common == 0) return false;
  // element stride 1 within a point (point-major batched, or shared)
  if (a->operand().fs != 1 || b->operand().fs != 1 || res_fs != 1) return false;
  if (n_points > 65535) return false;
  GemmShape g = gemm_shape(p);
  return g.M >= 32 && g.N >= 32 && kt.off_a.size() >= 8 && (double)g.M * (double)g.N * (double)kt.off_a.size() >= 65536.0;
}

enum ContractPath { PATH_JIT = 0, PATH_CSR = 1, PATH_GEMM = 2, PATH_DENSE = 3 };

// Everything tensor_contract derives from (structure of a, structure of b, what the operands are): computed once per
// distinct pair and kept in spn_ctx::contract_cache.
struct ContractEntry {
  spn_contract_plan plan{};
  Structure rs;
  KTables kt;
  DevOutMap map{};
  int path = PATH_DENSE;
  PairJit jit;
  // PATH_DENSE: contracted-loop tables, and (size_out <= 2^22) the operand base offsets of every result element
  DevTable<long long> d_ka, d_kb, d_toa, d_tob;
  DevTable<unsigned char> d_ks;
  bool has_out_tables = false;
  // PATH_CSR: CSR over result elements of the shared sparse operand
  DevTable<int> d_row;
  DevTable<long long> d_col;
  DevTable<double2> d_val;
  DevTable<unsigned char> d_sgn;
  size_t nnz = 0;
  // PATH_GEMM: row / column / k gather tables
  GemmShape g;
  DevTable<long long> d_row_a, d_out_row, d_col_b, d_out_col;
  DevTable<double> d_ksign;
  bool any_sign = false, vec_tables = false;
  int mixed = 0;   // 0: both operands of one type; 1: a real, b complex; 2: a complex, b real (DGEMM on doubled tables)
};

static void build_gemm_tables(ContractEntry& e, int mixed) {
  const spn_contract_plan& p = e.plan;
  e.g = gemm_shape(p);
  e.mixed = mixed;
  std::vector<uint64_t> ostride(p.order_out, 1);
  for (uint32_t r = p.order_out; r-- > 1;) ostride[r - 1] = ostride[r] * p.out_dim[r];
  std::vector<long long> row_a, out_row, col_b, out_col;
  auto enumerate = [&](bool from_a, std::vector<long long>& src, std::vector<long long>& dst) {
    std::vector<uint32_t> dims;
    for (uint32_t r = 0; r < p.order_out; ++r)
      if ((p.partition[r] != 0) == from_a) dims.push_back(r);
    const long long n = from_a ? e.g.M : e.g.N;
    src.resize((size_t)n);
    dst.resize((size_t)n);
    std::vector<uint32_t> idx(dims.size(), 0);
    for (long long i = 0; i < n; ++i) {
      long long so = 0, d0 = 0;
      for (size_t d = 0; d < dims.size(); ++d) {
        so += (long long)idx[d] * (long long)(from_a ? p.out_stride_a[dims[d]] : p.out_stride_b[dims[d]]);
        d0 += (long long)idx[d] * (long long)ostride[dims[d]];
      }
      src[(size_t)i] = so;
      dst[(size_t)i] = d0;
      for (size_t d = dims.size(); d-- > 0;) {
        if (++idx[d] < p.out_dim[dims[d]]) break;
        idx[d] = 0;
      }
    }
  };
  enumerate(true, row_a, out_row);
  enumerate(false, col_b, out_col);
  std::vector<long long> k_a = e.kt.off_a, k_b = e.kt.off_b;
  if (mixed) {
    // element offsets -> offsets in doubles of a real view: the complex side (operand and result) doubles, with the
    // (re, im) pair as two adjacent columns (a real, b complex) or rows (a complex, b real)
    auto twice = [](std::vector<long long>& v) {
      for (auto& x : v) x *= 2;
    };
    auto interleave = [](std::vector<long long>& v) {
      std::vector<long long> w(2 * v.size());
      for (size_t i = 0; i < v.size(); ++i) {
        w[2 * i] = 2 * v[i];
        w[2 * i + 1] = 2 * v[i] + 1;
      }
      v.swap(w);
    };
    if (mixed == 1) {   // C[r, (c, e)] = sum_k A[r, k] * B[k, (c, e)]
      interleave(col_b);
      interleave(out_col);
      twice(out_row);
      twice(k_b);
      e.g.N *= 2;
    } else {            // C[(r, e), c] = sum_k A[(r, e), k] * B[k, c]
      interleave(row_a);
      interleave(out_row);
      twice(out_col);
      twice(k_a);
      e.g.M *= 2;
    }
  }
  std::vector<double> ks;
  for (auto s : e.kt.sign) e.any_sign |= s != 0;
  if (e.any_sign)
    for (auto s : e.kt.sign) ks.push_back(s ? -1.0 : 1.0);
  // DGEMM: 16-byte gathers when both tables are pairwise contiguous and every base offset is even
  bool vec = e.g.N % 2 == 0 && k_a.size() % 2 == 0;
  for (size_t k = 0; vec && k + 1 < k_a.size(); k += 2) vec = k_a[k] % 2 == 0 && k_a[k + 1] == k_a[k] + 1;
  for (size_t c = 0; vec && c + 1 < col_b.size(); c += 2) vec = col_b[c] % 2 == 0 && col_b[c + 1] == col_b[c] + 1;
  for (size_t r = 0; vec && r < row_a.size(); ++r) vec = row_a[r] % 2 == 0;
  for (size_t k = 0; vec && k < k_b.size(); ++k) vec = k_b[k] % 2 == 0;
  e.vec_tables = vec;
  e.d_row_a.upload(row_a);
  e.d_out_row.upload(out_row);
  e.d_col_b.upload(col_b);
  e.d_out_col.upload(out_col);
  e.d_ka.upload(k_a);
  e.d_kb.upload(k_b);
  e.d_ksign.upload(ks);
}

static void build_csr_tables(ContractEntry& e, const spn_tensor* a, const spn_tensor* b) {
  const spn_contract_plan& plan = e.plan;
  const bool a_sp = a->sparse();
  const spn_tensor* sp = a_sp ? a : b;
  const auto& s_out = a_sp ? plan.out_stride_a : plan.out_stride_b;
  const auto& d_out = a_sp ? plan.out_stride_b : plan.out_stride_a;
  const auto& ks = a_sp ? e.kt.off_a : e.kt.off_b;
  const auto& kd = a_sp ? e.kt.off_b : e.kt.off_a;
  std::vector<int> row_ptr(plan.size_out + 1, 0);
  std::vector<long long> col;  // offsets into the dense operand: 64-bit (tensors beyond 2^31 elements)
  std::vector<double2> val;
  std::vector<unsigned char> sgn;
  std::vector<uint32_t> idx(plan.order_out, 0);
  for (uint64_t o = 0; o < plan.size_out; ++o) {
    uint64_t offs = 0, offd = 0;
    for (uint32_t r = 0; r < plan.order_out; ++r) {
      offs += (uint64_t)idx[r] * s_out[r];
      offd += (uint64_t)idx[r] * d_out[r];
    }
    for (size_t k = 0; k < ks.size(); ++k) {
      auto it = sp->entries.find(offs + (uint64_t)ks[k]);
      if (it == sp->entries.end()) continue;
      col.push_back((long long)(offd + (uint64_t)kd[k]));
      val.push_back(it->second);
      sgn.push_back(e.kt.sign[k]);
    }
    if (col.size() > (size_t)0x7fffffff) throw Error(SPN_ERR_INVALID, "sparse operand: more than 2^31 stored products");
    row_ptr[o + 1] = (int)col.size();
    for (uint32_t r = plan.order_out; r-- > 0;) {
      if (++idx[r] < plan.out_dim[r]) break;
      idx[r] = 0;
    }
  }
  e.nnz = col.size();
  e.d_row.upload(row_ptr);
  e.d_col.upload(col);
  e.d_val.upload(val);
  e.d_sgn.upload(sgn);
}

static uint64_t batch_points(const spn_tensor* a, const spn_tensor* b) {
  if (a->batched() && b->batched()) {
    if (a->n_points != b->n_points) throw Error(SPN_ERR_INVALID, "batched operands have different point counts");
    return a->n_points;
  }
  return a->batched() ? a->n_points : (b->batched() ? b->n_points : 1);
}

// cache key: the two structures as plain bytes + what the operands are.  Point counts enter only through the two
// thresholds the dispatch uses.
static void contract_key(std::string& key, const spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, uint64_t n_points) {
  key.clear();
  key.reserve(2 * (SPN_MAX_ORDER * sizeof(spn_slot) + 16) + 4);
  for (const spn_tensor* t : {a, b}) {
    key.append(reinterpret_cast<const char*>(t->st.slots.data()), t->st.slots.size() * sizeof(spn_slot));
    const char tag[6] = {(char)t->storage, (char)t->dtype, (char)t->layout, (char)(((uintptr_t)t->dptr % 32) == 0),
                         (char)(((uintptr_t)t->dptr % 16) == 0), (char)t->st.slots.size()};
    key.append(tag, sizeof tag);
    key.append(reinterpret_cast<const char*>(&t->uid), sizeof t->uid);
  }
  key.push_back((char)(n_points >= 2048));
  key.push_back((char)(n_points <= 65535));
  key.push_back((char)ctx->emulate_reference_pairing);
  // measurement switches that change which kernel is built (read per call so that tests can flip them)
  const char* lm = getenv("SPN_LOAD_MODE");
  key.push_back(lm ? lm[0] : '-');
  key.push_back(getenv("SPN_NO_PAIR_JIT") ? 'g' : 'j');
  key.push_back(getenv("SPN_PAIR_NO_SLAB") ? 'd' : 's');
  key.push_back(getenv("SPN_PAIR_NO_ROWS") ? 'g' : 'r');
}

static std::shared_ptr<ContractEntry> build_entry(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b,
                                                  const spn_tensor* res_probe, uint64_t n_points) {
  auto e = std::make_shared<ContractEntry>();
  e->plan = plan_contract(a->st, b->st);
  e->rs = structure_of_plan_result(e->plan);
  e->kt = build_k_tables(e->plan, ctx->emulate_reference_pairing);
  e->map = out_map_of(e->plan);
  if (e->kt.off_a.size() > (size_t)1 << 24) throw Error(SPN_ERR_INVALID, "contracted volume too large for the table path");
  (void)res_probe;
  (void)n_points;
  return e;
}

void tensor_contract(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, spn_tensor** out) {
  const uint64_t n_points = batch_points(a, b);
  const bool any_batched = a->batched() || b->batched();
  const int out_dtype = (a->dtype == SPN_F64 && b->dtype == SPN_F64) ? SPN_F64 : SPN_C64;
  const int out_layout = a->batched() ? a->layout : (b->batched() ? b->layout : SPN_POINT_MAJOR);
  const bool point_fastest = out_layout == SPN_COMPONENT_MAJOR;

  thread_local std::string key;
  contract_key(key, ctx, a, b, n_points);
  std::shared_ptr<ContractEntry> entry;
  auto hit = ctx->contract_cache.find(key);
  if (hit != ctx->contract_cache.end()) {
    entry = hit->second;
    ctx->contract_cache_hits++;
  } else {
    ctx->contract_cache_misses++;
    entry = build_entry(ctx, a, b, nullptr, n_points);
  }
  ContractEntry& e = *entry;
  const spn_contract_plan& plan = e.plan;
  // not cached: it depends on what the operands hold (an empty sparse receiver facing an empty dense operand)
  check_empty_sparse(a, b, (int)plan.merge_kind);
  std::unique_ptr<spn_tensor> res(new_batched(ctx, e.rs, n_points, out_dtype, out_layout, nullptr));
  const bool a_sp = a->sparse(), b_sp = b->sparse();

  if (hit == ctx->contract_cache.end()) {
    // ---- first time for this pair of structures: choose the kernel and build its tables ----
    static const char* kind_name[] = {"FirstBeforeSecond", "SecondBeforeFirst", "Interleaved"};
    const char* how = plan.n_common == 0 ? "exterior" : (plan.n_common == 1 ? "single" : "multi");
    if (any_batched && ((uintptr_t)res->dptr % 32) == 0 && pair_jit_applicable(plan, e.kt, a, b, res->batched(), n_points)) {
      // physics-sized per-point tensors: a kernel specialised for this stride plan (same rounding as the generic ones)
      e.path = PATH_JIT;
      e.jit = pair_jit_build(ctx, plan, e.kt, a, b, res.get());
    } else if (a_sp != b_sp) {
      e.path = PATH_CSR;  // K2: one shared sparse operand -> CSR gather over the stored entries only
      build_csr_tables(e, a, b);
    } else if (zgemm_applicable(plan, a, b, out_dtype, res->operand().fs, n_points, e.kt)) {
      e.path = PATH_GEMM;  // K6: the contraction reshapes to a GEMM large enough for the FP64 tensor pipe
      build_gemm_tables(e, a->dtype == b->dtype ? 0 : (a->dtype == SPN_F64 ? 1 : 2));
    } else {
      // K1/K5/K7 (and sparse x sparse over the densified shared copies; the structural/zero filter of the sparse
      // result is applied below)
      e.path = PATH_DENSE;
      e.d_ka.upload(e.kt.off_a);
      e.d_kb.upload(e.kt.off_b);
      e.d_ks.upload(e.kt.sign);
      if (plan.size_out <= (1ull << 22) && !getenv("SPN_NO_DENSE_TABLES")) {
        std::vector<long long> toa(plan.size_out), tob(plan.size_out);
        std::vector<uint32_t> idx(plan.order_out, 0);
        for (uint64_t o = 0; o < plan.size_out; ++o) {
          long long oa = 0, ob = 0;
          for (uint32_t r = 0; r < plan.order_out; ++r) {
            oa += (long long)idx[r] * (long long)plan.out_stride_a[r];
            ob += (long long)idx[r] * (long long)plan.out_stride_b[r];
          }
          toa[o] = oa;
          tob[o] = ob;
          for (uint32_t r = plan.order_out; r-- > 0;) {
            if (++idx[r] < plan.out_dim[r]) break;
            idx[r] = 0;
          }
        }
        e.d_toa.upload(toa);
        e.d_tob.upload(tob);
        e.has_out_tables = true;
      }
    }
    static const char* path_name[] = {"specialised kernel", "CSR gather kernel", "DMMA GEMM kernel", "generic dense kernel"};
    SPN_TRACE_LOG("contract %s %s: size_out=%llu n_k=%zu points=%llu -> %s%s", how, kind_name[plan.merge_kind],
                  (unsigned long long)plan.size_out, e.kt.off_a.size(), (unsigned long long)n_points, path_name[e.path],
                  e.path == PATH_JIT && e.jit.slab ? " (slab-staged)" : "");
    if (ctx->contract_cache.size() >= 4096) ctx->contract_cache.clear();  // bounded: entries hold device tables
    ctx->contract_cache.emplace(key, entry);
  }

  switch (e.path) {
    case PATH_JIT:
      if (((uintptr_t)res->dptr % 32) != 0) throw Error(SPN_ERR_OTHER, "pooled result allocation is not 32-byte aligned");
      pair_jit_launch(ctx, e.jit, a, b, res.get(), n_points);
      break;
    case PATH_CSR: {
      const spn_tensor* dn = a_sp ? b : a;
      cuda_check(launch_contract_csr(dn->operand(), res->output(), e.d_row.p, e.d_col.p, e.d_val.p, e.d_sgn.p,
                                     /*dense_is_first=*/a_sp ? 0 : 1, plan.size_out, e.nnz, n_points, point_fastest,
                                     ctx->stream),
                 "contract_csr_kernel");
      ctx->launches++;
      break;
    }
    case PATH_GEMM: {
      const long long K = (long long)e.kt.off_a.size();
      if (e.mixed) {
        // point strides in doubles of the real views
        const long long psa = a->operand().ps * (a->dtype == SPN_C64 ? 2 : 1), psb = b->operand().ps * (b->dtype == SPN_C64 ? 2 : 1);
        const bool vec = e.vec_tables && psa % 2 == 0 && psb % 2 == 0 && (uintptr_t)a->dptr % 16 == 0 && (uintptr_t)b->dptr % 16 == 0;
        cuda_check(launch_contract_dgemm(static_cast<const double*>(a->dptr), static_cast<const double*>(b->dptr),
                                         static_cast<double*>(res->dptr), e.d_row_a.p, e.d_ka.p, e.d_col_b.p, e.d_kb.p,
                                         e.d_out_row.p, e.d_out_col.p, e.any_sign ? e.d_ksign.p : nullptr, e.g.M, e.g.N, K, psa,
                                         psb, res->operand().ps * 2, n_points, vec, ctx->stream),
                   "contract_dgemm_kernel (mixed f64 x c64)");
      } else if (a->dtype == SPN_C64)
        cuda_check(launch_contract_zgemm(static_cast<const double2*>(a->dptr), static_cast<const double2*>(b->dptr),
                                         static_cast<double2*>(res->dptr), e.d_row_a.p, e.d_ka.p, e.d_col_b.p, e.d_kb.p,
                                         e.d_out_row.p, e.d_out_col.p, e.any_sign ? e.d_ksign.p : nullptr, e.g.M, e.g.N, K,
                                         a->operand().ps, b->operand().ps, res->operand().ps, n_points, ctx->stream),
                   "contract_zgemm_kernel");
      else {
        const bool vec = e.vec_tables && a->operand().ps % 2 == 0 && b->operand().ps % 2 == 0 &&
                         (uintptr_t)a->dptr % 16 == 0 && (uintptr_t)b->dptr % 16 == 0;
        cuda_check(launch_contract_dgemm(static_cast<const double*>(a->dptr), static_cast<const double*>(b->dptr),
                                         static_cast<double*>(res->dptr), e.d_row_a.p, e.d_ka.p, e.d_col_b.p, e.d_kb.p,
                                         e.d_out_row.p, e.d_out_col.p, e.any_sign ? e.d_ksign.p : nullptr, e.g.M, e.g.N, K,
                                         a->operand().ps, b->operand().ps, res->operand().ps, n_points, vec, ctx->stream),
                   "contract_dgemm_kernel");
      }
      ctx->launches++;
      break;
    }
    default: {
      cudaError_t rc = cudaErrorInvalidConfiguration;
      if (e.has_out_tables)
        rc = launch_contract_dense_tab(a->operand(), b->operand(), res->output(), e.d_toa.p, e.d_tob.p, e.d_ka.p, e.d_kb.p,
                                       e.d_ks.p, e.kt.off_a.size(), plan.size_out, n_points, point_fastest, ctx->stream);
      if (rc == cudaErrorInvalidConfiguration)  // result too large for tables / 32-bit tiles: run-time index arithmetic
        rc = launch_contract_dense(e.map, a->operand(), b->operand(), res->output(), e.d_ka.p, e.d_kb.p, e.d_ks.p,
                                   e.kt.off_a.size(), n_points, point_fastest, ctx->stream);
      cuda_check(rc, "contract_dense_kernel");
      ctx->launches++;
      break;
    }
  }
  if (any_batched) {
    *out = res.release();
    return;
  }
  const KTables& kt = e.kt;
  const Structure& rs = e.rs;
  // shared x shared: the result is itself a shared tensor (computed on the device with one point)
  std::vector<double2> host(plan.size_out);
  cuda_check(cudaMemcpyAsync(host.data(), res->dptr, host.size() * sizeof(double2), cudaMemcpyDeviceToHost,
                             ctx->stream),
             "download shared result");
  cuda_check(cudaStreamSynchronize(ctx->stream), "download shared result");
  if (a_sp && b_sp) {
    // Sparse x Sparse -> Sparse (contraction.rs:222-224): an entry exists iff at least one pair of
    // stored entries met (structural, integer work) AND the value is not exactly zero
    // (singlecontract.rs:268, multicontract.rs:276).  Receiver without entries => empty result.
    std::map<uint64_t, double2> ent;
    const spn_tensor* recv = plan.merge_kind == SPN_SECOND_BEFORE_FIRST ? b : a;
    if (!recv->entries.empty()) {
      std::vector<uint32_t> idx(plan.order_out, 0);
      for (uint64_t o = 0; o < plan.size_out; ++o) {
        uint64_t offa = 0, offb = 0;
        for (uint32_t r = 0; r < plan.order_out; ++r) {
          offa += (uint64_t)idx[r] * plan.out_stride_a[r];
          offb += (uint64_t)idx[r] * plan.out_stride_b[r];
        }
        bool met = false;
        for (size_t k = 0; k < kt.off_a.size() && !met; ++k)
          met = a->entries.count(offa + (uint64_t)kt.off_a[k]) && b->entries.count(offb + (uint64_t)kt.off_b[k]);
        // exterior products keep every met pair, even exact zeros (exteriorproduct.rs:19-46)
        if (met && (plan.n_common == 0 || host[o].x != 0.0 || host[o].y != 0.0)) ent[o] = host[o];
        for (uint32_t r = plan.order_out; r-- > 0;) {
          if (++idx[r] < plan.out_dim[r]) break;
          idx[r] = 0;
        }
      }
    }
    *out = new_shared_sparse(ctx, rs, ent);
  } else {
    *out = new_shared_dense(ctx, rs, host.data());
  }
}

void tensor_elementwise(spn_ctx* ctx, int op, const spn_tensor* a, const spn_tensor* b, double2 s, spn_tensor** out,
                        bool drop_zeros) {
  const bool binary = op == EW_ADD || op == EW_SUB;
  if (binary) {
    if (a->st.order() != b->st.order() || a->size() != b->size())
      throw Error(SPN_ERR_STRUCTURE, "element-wise op on tensors of different structure");
    for (uint32_t i = 0; i < a->st.order(); ++i)
      if (!slot_same(a->st.slots[i], b->st.slots[i]))
        throw Error(SPN_ERR_STRUCTURE, "element-wise op on tensors of different structure");
    if (a->batched() && b->batched() && a->n_points != b->n_points)
      throw Error(SPN_ERR_INVALID, "batched operands have different point counts");
  }
  uint64_t n_points = a->batched() ? a->n_points : (binary && b->batched() ? b->n_points : 1);
  const bool any_batched = a->batched() || (binary && b->batched());
  int dtype = a->dtype;
  if (binary && b->dtype == SPN_C64) dtype = SPN_C64;
  if (op == EW_SCALE && s.y != 0.0) dtype = SPN_C64;
  if (op == EW_SCALE && a->dtype == SPN_F64 && s.y == 0.0) dtype = SPN_F64;
  const int layout = a->batched() ? a->layout : (binary && b->batched() ? b->layout : SPN_POINT_MAJOR);
  std::unique_ptr<spn_tensor> res(new_batched(ctx, a->st, n_points, dtype, layout, nullptr));
  DevOperand ob = binary ? b->operand() : a->operand();
  cuda_check(launch_elementwise(op, a->operand(), ob, res->output(), s, a->size(), n_points,
                                layout == SPN_COMPONENT_MAJOR, ctx->stream),
             "elementwise_kernel");
  ctx->launches++;
  if (any_batched) {
    *out = res.release();
    return;
  }
  std::vector<double2> host(a->size());
  cuda_check(cudaMemcpyAsync(host.data(), res->dptr, host.size() * sizeof(double2), cudaMemcpyDeviceToHost,
                             ctx->stream),
             "download shared result");
  cuda_check(cudaStreamSynchronize(ctx->stream), "download shared result");
  const bool keep_sparse = binary ? (a->sparse() && b->sparse()) : a->sparse();
  if (keep_sparse) {
    std::map<uint64_t, double2> ent;
    for (auto& kv : a->entries) ent[kv.first] = host[kv.first];
    if (binary)
      for (auto& kv : b->entries) ent[kv.first] = host[kv.first];
    if (drop_zeros)  // smart_set (tensors/data/sparse.rs:712-724): an exactly-zero value is not stored
      for (auto it = ent.begin(); it != ent.end();) it = (it->second.x == 0.0 && it->second.y == 0.0) ? ent.erase(it) : std::next(it);
    *out = new_shared_sparse(ctx, a->st, ent);
  } else {
    *out = new_shared_dense(ctx, a->st, host.data());
  }
}

}  // namespace spn

// =================================================================================================
extern "C" {

const char* spn_last_error(void) { return g_err.c_str(); }

namespace {
struct SymbolTable {
  std::mutex m;
  std::map<std::string, uint64_t> id_of;
  std::deque<std::string> names;   // deque: references stay valid as it grows
};
SymbolTable& symbols() {
  static SymbolTable t;
  return t;
}
}  // namespace
uint64_t spn_symbol_intern(const char* name) {
  SymbolTable& t = symbols();
  std::lock_guard<std::mutex> lock(t.m);
  const std::string key = name ? name : "";
  auto it = t.id_of.find(key);
  if (it != t.id_of.end()) return it->second;
  t.names.push_back(key);
  t.id_of[key] = t.names.size() - 1;
  return t.names.size() - 1;
}
const char* spn_symbol_name(uint64_t id) {
  SymbolTable& t = symbols();
  std::lock_guard<std::mutex> lock(t.m);
  return id < t.names.size() ? t.names[(size_t)id].c_str() : nullptr;
}
int spn_version(void) { return SPN_VERSION; }

int spn_structure_canonicalize(const spn_slot* in, uint32_t n, spn_slot* out, int32_t* perm, int64_t* base_start,
                               int64_t* dual_start) {
  API_TRY
  std::vector<int32_t> p;
  Structure st = canonicalize(std::vector<spn_slot>(in, in + n), &p);
  for (uint32_t i = 0; i < n; ++i) {
    out[i] = st.slots[i];
    if (perm) perm[i] = p[i];
  }
  if (base_start) *base_start = st.base_start;
  if (dual_start) *dual_start = st.dual_start;
  return SPN_OK;
  API_CATCH
}

int spn_plan_contract(const spn_slot* a, uint32_t na, const spn_slot* b, uint32_t nb, spn_contract_plan* out) {
  API_TRY
  if (na > SPN_MAX_ORDER || nb > SPN_MAX_ORDER) throw Error(SPN_ERR_INVALID, "operand order exceeds SPN_MAX_ORDER");
  Structure A = structure_of_canonical(a, na), B = structure_of_canonical(b, nb);
  *out = plan_contract(A, B);
  return SPN_OK;
  API_CATCH
}

const char* spn_plan_contract_cuda_source(const spn_slot* a, uint32_t na, int dtype_a, int layout_a, const spn_slot* b,
                                          uint32_t nb, int dtype_b, int layout_b) {
  static thread_local std::string src;
  try {
    src = pair_jit_source(structure_of_canonical(a, na), dtype_a, layout_a, structure_of_canonical(b, nb), dtype_b, layout_b);
    return src.c_str();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

const char* spn_plan_contract_sparse_cuda_source(const spn_slot* sparse, uint32_t n_sparse, const uint64_t* flat,
                                                 const double* values, uint64_t n_entries, const spn_slot* batched,
                                                 uint32_t n_batched, int dtype, int layout, int sparse_first) {
  static thread_local std::string src;
  try {
    Structure ss = structure_of_canonical(sparse, n_sparse);
    std::map<uint64_t, double2> entries;
    for (uint64_t i = 0; i < n_entries; ++i) {
      if (flat[i] >= ss.size()) throw Error(SPN_ERR_INVALID, "sparse entry outside the tensor");
      entries[flat[i]] = make_double2(values[2 * i], values[2 * i + 1]);
    }
    src = pair_jit_source_sparse(ss, entries, structure_of_canonical(batched, n_batched), dtype, layout, sparse_first != 0);
    return src.c_str();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

int spn_ctx_create(int device, spn_ctx** out) {
  API_TRY
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    throw Error(SPN_ERR_NO_DEVICE, std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                                       " (spenso_b200 has no CPU execution path)");
  if (device < 0 || device >= count) throw Error(SPN_ERR_INVALID, "device index out of range");
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  auto c = std::make_unique<spn_ctx>();
  c->device = device;
  cuda_check(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device), "device attribute");
  cuda_check(cudaFree(0), "context init");
  // L2 fetch granularity (bytes an L2 miss brings in from HBM): SPN_L2_FETCH=32|64|128 overrides the default set below
  if (const char* e = getenv("SPN_L2_FETCH")) {
    const int g = atoi(e);
    if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
    cudaGetLastError();
  }
  {
    size_t cur = 0;
    if (cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity) == cudaSuccess) c->l2_fetch_granularity = (int)cur;
    cudaGetLastError();
  }
  {  // Freed intermediates stay in the device's memory pool until spn_ctx_release_cached_memory.  With a finite release
     // threshold the driver returns "surplus" memory to the OS at synchronisation points as soon as the pool holds more
     // than the threshold — resident inputs included — and the next result has to be mapped again: measured at 2^22
     // points (profiles/r02_pairwise_pool.txt) every pairwise kernel that writes a 1 GiB result ran at 2.5-4.7 TB/s with
     // a 4 GiB threshold and at 5.5-7.3 TB/s with none.  SPN_POOL_KEEP_GB restores a finite threshold.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      if (const char* e = getenv("SPN_POOL_KEEP_GB")) keep = (uint64_t)atof(e) << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  *out = c.release();
  return SPN_OK;
  API_CATCH
}
void spn_ctx_destroy(spn_ctx* ctx) { delete ctx; }
int spn_ctx_set_stream(spn_ctx* ctx, void* s) {
  if (ctx->stream != (cudaStream_t)s) ctx->drop_blocks();   // recycled buffers are only ordered on the stream they ran on
  ctx->stream = (cudaStream_t)s;
  return SPN_OK;
}
int spn_ctx_release_cached_memory(spn_ctx* ctx) {
  API_TRY
  ctx->drop_blocks();
  cuda_check(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  cudaMemPool_t pool;
  cuda_check(cudaDeviceGetDefaultMemPool(&pool, ctx->device), "cudaDeviceGetDefaultMemPool");
  cuda_check(cudaMemPoolTrimTo(pool, 0), "cudaMemPoolTrimTo");
  return SPN_OK;
  API_CATCH
}
int spn_ctx_set_reference_quirks(spn_ctx* ctx, int emulate_interleaved_pairing) {
  ctx->emulate_reference_pairing = emulate_interleaved_pairing != 0;
  return SPN_OK;
}
int spn_ctx_synchronize(spn_ctx* ctx) {
  API_TRY
  cuda_check(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  for (const int* st : ctx->comm_status)
    if (const int v = *reinterpret_cast<const volatile int*>(st))
      throw Error(SPN_ERR_OTHER, "Monte-Carlo sum exchange timed out waiting for rank " + std::to_string(v - 1) +
                                     " (the sum was poisoned with NaN; re-create the communicator)");
  return SPN_OK;
  API_CATCH
}
uint64_t spn_ctx_launch_count(const spn_ctx* ctx) { return ctx->launches; }
int spn_ctx_plan_cache_stats(const spn_ctx* ctx, uint64_t* hits, uint64_t* misses, uint64_t* entries) {
  if (hits) *hits = ctx->contract_cache_hits;
  if (misses) *misses = ctx->contract_cache_misses;
  if (entries) *entries = ctx->contract_cache.size();
  return SPN_OK;
}
int spn_ctx_sm_count(const spn_ctx* ctx) { return ctx->sm_count; }

int spn_tensor_batched(spn_ctx* ctx, const spn_slot* slots, uint32_t n, uint64_t n_points, int dtype, int layout,
                       spn_tensor** out) {
  API_TRY
  *out = new_batched(ctx, structure_of_canonical(slots, n), n_points, dtype, layout, nullptr);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_batched_wrap(spn_ctx* ctx, const spn_slot* slots, uint32_t n, uint64_t n_points, int dtype, int layout,
                            void* device_ptr, spn_tensor** out) {
  API_TRY
  if (!device_ptr) throw Error(SPN_ERR_INVALID, "null device pointer");
  *out = new_batched(ctx, structure_of_canonical(slots, n), n_points, dtype, layout, device_ptr);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_shared_dense(spn_ctx* ctx, const spn_slot* slots, uint32_t n, const double* reim, spn_tensor** out) {
  API_TRY
  *out = new_shared_dense(ctx, structure_of_canonical(slots, n), reinterpret_cast<const double2*>(reim));
  return SPN_OK;
  API_CATCH
}
int spn_tensor_shared_sparse(spn_ctx* ctx, const spn_slot* slots, uint32_t n, const uint64_t* flat, const double* reim,
                             uint64_t nnz, spn_tensor** out) {
  API_TRY
  std::map<uint64_t, double2> ent;
  for (uint64_t k = 0; k < nnz; ++k) ent[flat[k]] = make_double2(reim[2 * k], reim[2 * k + 1]);
  *out = new_shared_sparse(ctx, structure_of_canonical(slots, n), ent);
  return SPN_OK;
  API_CATCH
}
// NetworkLeaf::LibraryKey -> tensor with concrete indices: get_lib_data + permute_inds (network/graph.rs:291-322,
// tensors/data/dense.rs:93-109) for the built-in hep_lib tensors; index bookkeeping once, on the host
int spn_tensor_library(spn_ctx* ctx, int which, const spn_slot* slots, uint32_t n, spn_tensor** out) {
  API_TRY
  std::vector<spn_slot> args(slots, slots + n);
  LibraryData d = builtin_library(which, args);
  Structure st;
  auto ent = instantiate_library(d, args, &st);
  *out = new_shared_sparse(ctx, st, ent);
  return SPN_OK;
  API_CATCH
}
void spn_tensor_free(spn_tensor* t) { delete t; }

uint32_t spn_tensor_order(const spn_tensor* t) { return t->st.order(); }
int spn_tensor_slots(const spn_tensor* t, spn_slot* out) {
  for (uint32_t i = 0; i < t->st.order(); ++i) out[i] = t->st.slots[i];
  return SPN_OK;
}
uint64_t spn_tensor_size(const spn_tensor* t) { return t->size(); }
uint64_t spn_tensor_n_points(const spn_tensor* t) { return t->n_points; }
int spn_tensor_storage(const spn_tensor* t) { return t->storage; }
int spn_tensor_dtype(const spn_tensor* t) { return t->dtype; }
int spn_tensor_layout(const spn_tensor* t) { return t->layout; }
uint64_t spn_tensor_nnz(const spn_tensor* t) { return t->sparse() ? t->entries.size() : t->size(); }
void* spn_tensor_device_ptr(const spn_tensor* t) { return t->dptr; }

static void do_upload(spn_tensor* t, const void* host, uint64_t first_point, uint64_t n_points, bool sync) {
  if (!t->batched()) throw Error(SPN_ERR_INVALID, "upload: shared tensors are immutable");
  if (first_point + n_points > t->n_points) throw Error(SPN_ERR_INVALID, "upload: point range out of bounds");
  spn_ctx* ctx = t->ctx;
  const size_t eb = t->elem_bytes(), sz = t->size();
  if (t->layout == SPN_POINT_MAJOR) {
    cuda_check(cudaMemcpyAsync((char*)t->dptr + first_point * sz * eb, host, n_points * sz * eb,
                               cudaMemcpyHostToDevice, ctx->stream),
               "upload");
  } else {
    void* stage = nullptr;
    cuda_check(cudaMallocAsync(&stage, n_points * sz * eb + 16, ctx->stream), "upload staging");
    cuda_check(cudaMemcpyAsync(stage, host, n_points * sz * eb, cudaMemcpyHostToDevice, ctx->stream), "upload");
    // staged [point][flat] -> resident [flat][point]
    cuda_check(launch_transpose(stage, (char*)t->dptr + first_point * eb, n_points, sz, sz, t->n_points, (int)eb, ctx->stream),
               "layout transpose");
    ctx->launches++;
    cudaFreeAsync(stage, ctx->stream);
  }
  if (sync) cuda_check(cudaStreamSynchronize(ctx->stream), "upload");
}

int spn_tensor_upload(spn_tensor* t, const void* host, uint64_t first_point, uint64_t n_points) {
  API_TRY
  do_upload(t, host, first_point, n_points, true);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_upload_async(spn_tensor* t, const void* host, uint64_t first_point, uint64_t n_points) {
  API_TRY
  do_upload(t, host, first_point, n_points, false);
  return SPN_OK;
  API_CATCH
}

int spn_tensor_download(const spn_tensor* t, void* host, uint64_t first_point, uint64_t n_points) {
  API_TRY
  spn_ctx* ctx = t->ctx;
  const size_t eb = t->elem_bytes(), sz = t->size();
  if (!t->batched()) {
    if (first_point != 0 || n_points != 1) throw Error(SPN_ERR_INVALID, "shared tensors have exactly one point");
    std::memcpy(host, t->host_dense.data(), sz * sizeof(double2));
    return SPN_OK;
  }
  if (first_point + n_points > t->n_points) throw Error(SPN_ERR_INVALID, "download: point range out of bounds");
  if (t->layout == SPN_POINT_MAJOR) {
    cuda_check(cudaMemcpyAsync(host, (const char*)t->dptr + first_point * sz * eb, n_points * sz * eb,
                               cudaMemcpyDeviceToHost, ctx->stream),
               "download");
  } else {
    void* stage = nullptr;
    cuda_check(cudaMallocAsync(&stage, n_points * sz * eb + 16, ctx->stream), "download staging");
    // resident [flat][point] -> staged [point][flat]
    cuda_check(launch_transpose((const char*)t->dptr + first_point * eb, stage, sz, n_points, t->n_points, sz, (int)eb, ctx->stream),
               "layout transpose");
    ctx->launches++;
    cuda_check(cudaMemcpyAsync(host, stage, n_points * sz * eb, cudaMemcpyDeviceToHost, ctx->stream), "download");
    cudaFreeAsync(stage, ctx->stream);
  }
  cuda_check(cudaStreamSynchronize(ctx->stream), "download");
  return SPN_OK;
  API_CATCH
}

int spn_tensor_sparse_entries(const spn_tensor* t, uint64_t* flat, double* reim) {
  API_TRY
  if (!t->sparse()) throw Error(SPN_ERR_INVALID, "not a sparse tensor");
  size_t k = 0;
  for (auto& kv : t->entries) {
    flat[k] = kv.first;
    reim[2 * k] = kv.second.x;
    reim[2 * k + 1] = kv.second.y;
    ++k;
  }
  return SPN_OK;
  API_CATCH
}

int spn_tensor_contract(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, spn_tensor** out) {
  API_TRY
  tensor_contract(ctx, a, b, out);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_add(spn_ctx* ctx, const spn_tensor* a, const spn_tensor* b, int subtract, spn_tensor** out) {
  API_TRY
  tensor_elementwise(ctx, (subtract & SPN_ADD_SUBTRACT) ? EW_SUB : EW_ADD, a, b, make_double2(0, 0), out,
                     (subtract & SPN_ADD_FALLIBLE) != 0);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_neg(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  tensor_elementwise(ctx, EW_NEG, a, a, make_double2(0, 0), out);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_scalar_mul(spn_ctx* ctx, const spn_tensor* a, double re, double im, spn_tensor** out) {
  API_TRY
  tensor_elementwise(ctx, EW_SCALE, a, a, make_double2(re, im), out);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_conj(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  tensor_elementwise(ctx, EW_CONJ, a, a, make_double2(0, 0), out);
  return SPN_OK;
  API_CATCH
}

int spn_tensor_recip(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  tensor_elementwise(ctx, EW_RECIP, a, a, make_double2(0, 0), out);
  return SPN_OK;
  API_CATCH
}
int spn_tensor_clone(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  if (a->batched()) {
    std::unique_ptr<spn_tensor> r(new_batched(ctx, a->st, a->n_points, a->dtype, a->layout, nullptr));
    cuda_check(cudaMemcpyAsync(r->dptr, a->dptr, a->n_points * a->size() * a->elem_bytes(), cudaMemcpyDeviceToDevice, ctx->stream),
               "clone");
    *out = r.release();
  } else {
    *out = a->sparse() ? new_shared_sparse(ctx, a->st, a->entries) : new_shared_dense(ctx, a->st, a->host_dense.data());
  }
  return SPN_OK;
  API_CATCH
}
// point-major <-> component-major copy of a batched tensor, on the device (tiled shared-memory transpose)
int spn_tensor_relayout(spn_ctx* ctx, const spn_tensor* a, int layout, spn_tensor** out) {
  API_TRY
  if (!a->batched()) throw Error(SPN_ERR_INVALID, "relayout: only batched tensors have a layout");
  if (layout != SPN_POINT_MAJOR && layout != SPN_COMPONENT_MAJOR) throw Error(SPN_ERR_INVALID, "relayout: unknown layout");
  if (layout == a->layout) return spn_tensor_clone(ctx, a, out);
  std::unique_ptr<spn_tensor> r(new_batched(ctx, a->st, a->n_points, a->dtype, layout, nullptr));
  const uint64_t sz = a->size(), np = a->n_points;
  // source viewed as a [rows][cols] matrix: point-major [np][sz] -> [sz][np], component-major [sz][np] -> [np][sz]
  const uint64_t rows = a->layout == SPN_POINT_MAJOR ? np : sz, cols = a->layout == SPN_POINT_MAJOR ? sz : np;
  cuda_check(launch_transpose(a->dptr, r->dptr, rows, cols, cols, rows, (int)a->elem_bytes(), ctx->stream), "layout transpose");
  ctx->launches++;
  *out = r.release();
  return SPN_OK;
  API_CATCH
}
// DataTensor::to_dense / to_sparse (tensors/data/sparse.rs:688-708, dense.rs:523-535): index bookkeeping on the host
// copies every shared tensor keeps, no arithmetic
int spn_tensor_to_dense(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  if (a->batched()) return spn_tensor_clone(ctx, a, out);   // batched tensors are dense
  *out = new_shared_dense(ctx, a->st, a->host_dense.data());
  return SPN_OK;
  API_CATCH
}
int spn_tensor_to_sparse(spn_ctx* ctx, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  if (a->batched()) throw Error(SPN_ERR_INVALID, "to_sparse: a batched tensor is one DenseTensor per point");
  std::map<uint64_t, double2> ent;
  if (a->sparse()) {
    ent = a->entries;
  } else {
    for (uint64_t f = 0; f < a->host_dense.size(); ++f)   // values equal to T::default() are not stored (dense.rs:529-532)
      if (a->host_dense[f].x != 0.0 || a->host_dense[f].y != 0.0) ent[f] = a->host_dense[f];
  }
  *out = new_shared_sparse(ctx, a->st, ent);
  return SPN_OK;
  API_CATCH
}
// element access by canonical flat index: GetTensorData::get_ref_linear / SetTensorData::set_flat
// (tensors/data/dense.rs, sparse.rs).  A sparse tensor reports an absent entry as SPN_ERR_EMPTY_SPARSE on get and
// stores whatever is set (zeros included, like SparseTensor::set_flat).
int spn_tensor_get(const spn_tensor* t, uint64_t point, uint64_t flat, double* reim) {
  API_TRY
  if (flat >= t->size()) throw Error(SPN_ERR_DIMENSION, "get: flat index out of bounds");
  if (!t->batched()) {
    if (point != 0) throw Error(SPN_ERR_INVALID, "get: shared tensors have one point");
    if (t->sparse() && !t->entries.count(flat)) throw Error(SPN_ERR_EMPTY_SPARSE, "get: no stored entry at this index");
    reim[0] = t->host_dense[flat].x;
    reim[1] = t->host_dense[flat].y;
    return SPN_OK;
  }
  if (point >= t->n_points) throw Error(SPN_ERR_INVALID, "get: point out of bounds");
  const size_t eb = t->elem_bytes();
  const uint64_t idx = t->layout == SPN_POINT_MAJOR ? point * t->size() + flat : flat * t->n_points + point;
  reim[1] = 0.0;
  cuda_check(cudaMemcpyAsync(reim, (const char*)t->dptr + idx * eb, eb, cudaMemcpyDeviceToHost, t->ctx->stream), "get");
  cuda_check(cudaStreamSynchronize(t->ctx->stream), "get");
  return SPN_OK;
  API_CATCH
}
int spn_tensor_set(spn_tensor* t, uint64_t point, uint64_t flat, double re, double im) {
  API_TRY
  if (flat >= t->size()) throw Error(SPN_ERR_DIMENSION, "set: flat index out of bounds");
  const double v[2] = {re, im};
  if (!t->batched()) {
    if (point != 0) throw Error(SPN_ERR_INVALID, "set: shared tensors have one point");
    t->host_dense[flat] = make_double2(re, im);
    if (t->sparse()) t->entries[flat] = make_double2(re, im);
    t->uid = g_next_uid.fetch_add(1);   // the content is part of the identity the plan cache and the JIT kernels key on
    cuda_check(cudaMemcpyAsync((char*)t->dptr + flat * sizeof(double2), v, sizeof(double2), cudaMemcpyHostToDevice, t->ctx->stream), "set");
    cuda_check(cudaStreamSynchronize(t->ctx->stream), "set");
    return SPN_OK;
  }
  if (point >= t->n_points) throw Error(SPN_ERR_INVALID, "set: point out of bounds");
  if (t->dtype == SPN_F64 && im != 0.0) throw Error(SPN_ERR_INVALID, "set: complex value into an f64 tensor");
  const size_t eb = t->elem_bytes();
  const uint64_t idx = t->layout == SPN_POINT_MAJOR ? point * t->size() + flat : flat * t->n_points + point;
  cuda_check(cudaMemcpyAsync((char*)t->dptr + idx * eb, v, eb, cudaMemcpyHostToDevice, t->ctx->stream), "set");
  cuda_check(cudaStreamSynchronize(t->ctx->stream), "set");
  return SPN_OK;
  API_CATCH
}

int spn_tensor_trace(spn_ctx* ctx, const spn_slot* slots, uint32_t n, const spn_tensor* a, spn_tensor** out) {
  API_TRY
  // `slots` is the index list of `a` as the caller sees it, with one repeated (slot, dual) pair:
  // traces(), structure.rs:462-485.  The data of `a` is row major over `slots`.
  if (n == 0 || n != a->st.order()) throw Error(SPN_ERR_INVALID, "trace: slot count differs from the tensor's order");
  {
    uint64_t vol = 1;
    for (uint32_t i = 0; i < n; ++i) vol *= slots[i].dim;
    if (vol != a->size()) throw Error(SPN_ERR_DIMENSION, "trace: the slot dimensions do not multiply to the tensor's size");
  }
  int p0 = -1, p1 = -1;
  for (uint32_t i = 0; i < n && p0 < 0; ++i)
    for (uint32_t j = i + 1; j < n; ++j)
      if (slot_same(dual_of(slots[j]), slots[i])) {
        p0 = (int)i;
        p1 = (int)j;
        break;
      }
  if (p0 < 0) throw Error(SPN_ERR_STRUCTURE, "trace: no repeated slot");
  std::vector<uint64_t> strides(n, 1);
  for (uint32_t i = n; i-- > 1;) strides[i - 1] = strides[i] * slots[i].dim;
  std::vector<spn_slot> rest;
  std::vector<uint64_t> rest_strides;
  for (uint32_t i = 0; i < n; ++i)
    if ((int)i != p0 && (int)i != p1) {
      rest.push_back(slots[i]);
      rest_strides.push_back(strides[i]);
    }
  std::vector<int32_t> perm;
  Structure rs = canonicalize(rest, &perm);
  DevOutMap map{};
  map.order = (int)rs.order();
  if (map.order > kMaxOutOrder) throw Error(SPN_ERR_INVALID, "trace: order too large");
  for (uint32_t r = 0; r < rs.order(); ++r) {
    map.dim[r] = rs.slots[r].dim;
    map.sa[r] = (long long)rest_strides[perm[r]];
  }
  map.size_out = rs.size();
  uint64_t n_points = a->batched() ? a->n_points : 1;
  std::unique_ptr<spn_tensor> res(new_batched(ctx, rs, n_points, a->dtype, a->layout, nullptr));
  cuda_check(launch_trace(map, a->operand(), res->output(), (long long)(strides[p0] + strides[p1]), slots[p0].dim,
                          has_metric(slots[p0]) ? 1 : 0, n_points, a->layout == SPN_COMPONENT_MAJOR, ctx->stream),
             "trace_kernel");
  ctx->launches++;
  if (a->batched()) {
    *out = res.release();
    return SPN_OK;
  }
  std::vector<double2> host(rs.size());
  cuda_check(cudaMemcpyAsync(host.data(), res->dptr, host.size() * sizeof(double2), cudaMemcpyDeviceToHost,
                             ctx->stream),
             "download shared result");
  cuda_check(cudaStreamSynchronize(ctx->stream), "download shared result");
  if (a->sparse()) {
    // Trace for SparseTensor keeps the result sparse and drops exact zeros (contraction/trace.rs:41-83, :59)
    std::map<uint64_t, double2> ent;
    for (uint64_t f = 0; f < host.size(); ++f)
      if (host[f].x != 0.0 || host[f].y != 0.0) ent[f] = host[f];
    *out = new_shared_sparse(ctx, rs, ent);
  } else {
    *out = new_shared_dense(ctx, rs, host.data());
  }
  return SPN_OK;
  API_CATCH
}

int spn_tensor_reduce_points(spn_ctx* ctx, const spn_tensor* a, double* out_device) {
  API_TRY
  if (a->dtype != SPN_C64) throw Error(SPN_ERR_INVALID, "reduce_points expects a complex tensor");
  const int n_blocks = ctx->sm_count * 4;
  double2* partial = nullptr;
  cuda_check(cudaMallocAsync((void**)&partial, (size_t)n_blocks * a->size() * sizeof(double2) + 16, ctx->stream),
             "reduce scratch");
  cuda_check(launch_reduce_points(a->operand(), a->size(), a->n_points, partial, n_blocks,
                                  reinterpret_cast<double2*>(out_device), ctx->stream),
             "reduce_points_kernel");
  ctx->launches += 2;
  cudaFreeAsync(partial, ctx->stream);
  return SPN_OK;
  API_CATCH
}

}  // extern "C"
