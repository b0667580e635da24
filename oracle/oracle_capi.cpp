// ORACLE — TEST INFRASTRUCTURE ONLY (see spenso_oracle.hpp header).
// C ABI over the restatement so that tests/ and bench.py's cpu_baseline leg can
// drive it through ctypes.  Nothing under spenso_b200/ may load this library.
#include <cstring>
#include <string>
#include <thread>

#include "spenso_oracle.hpp"

using namespace orc;

static thread_local std::string g_err;

#define ORC_TRY try {
#define ORC_CATCH(defret)                   \
  }                                         \
  catch (const ContractionError& e) {       \
    g_err = e.what();                       \
    return e.code;                          \
  }                                         \
  catch (const std::exception& e) {         \
    g_err = e.what();                       \
    return defret;                          \
  }

struct NetBox {
  Network net;
  std::vector<LibTensor> library;
};

static Structure canon_structure(const Slot* s, int n) {
  std::vector<Slot> v(s, s + n);
  std::vector<int> perm;
  Structure st = structure_from_slots(v, &perm);
  for (int i = 0; i < n; ++i)
    if (perm[i] != i) throw std::runtime_error("slots are not in canonical order");
  return st;
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_slot_cmp(const Slot* a, const Slot* b) { return slot_cmp(*a, *b); }

int orc_order_structure(const Slot* in, int n, Slot* out, int32_t* perm, int64_t* base_start, int64_t* dual_start) {
  ORC_TRY
  std::vector<int> p;
  Structure st = structure_from_slots(std::vector<Slot>(in, in + n), &p);
  for (int i = 0; i < n; ++i) {
    out[i] = st.s[i];
    perm[i] = p[i];
  }
  *base_start = st.base_start == NONE ? -1 : (int64_t)st.base_start;
  *dual_start = st.dual_start == NONE ? -1 : (int64_t)st.dual_start;
  return 0;
  ORC_CATCH(ERR_OTHER)
}

int orc_merge(const Slot* a, int na, const Slot* b, int nb, Slot* out, int32_t* nout, uint8_t* common_a,
              uint8_t* common_b, uint8_t* partition, int32_t* npart, int32_t* kind, int64_t* base_start,
              int64_t* dual_start) {
  ORC_TRY
  Structure sa = structure_from_slots(std::vector<Slot>(a, a + na));
  Structure sb = structure_from_slots(std::vector<Slot>(b, b + nb));
  MergeResult m = merge(sa, sb);
  *nout = (int32_t)m.st.order();
  for (size_t i = 0; i < m.st.order(); ++i) out[i] = m.st.s[i];
  for (int i = 0; i < na; ++i) common_a[i] = m.common_self[i];
  for (int i = 0; i < nb; ++i) common_b[i] = m.common_other[i];
  *npart = (int32_t)m.partition.size();
  for (size_t i = 0; i < m.partition.size(); ++i) partition[i] = m.partition[i];
  *kind = m.kind;
  *base_start = m.st.base_start == NONE ? -1 : (int64_t)m.st.base_start;
  *dual_start = m.st.dual_start == NONE ? -1 : (int64_t)m.st.dual_start;
  return 0;
  ORC_CATCH(ERR_STRUCTURE)
}

int orc_match_indices(const Slot* a, int na, const Slot* b, int nb, int64_t* perm, int32_t* nperm, uint8_t* ma,
                      uint8_t* mb) {
  ORC_TRY
  Structure sa = structure_from_slots(std::vector<Slot>(a, a + na));
  Structure sb = structure_from_slots(std::vector<Slot>(b, b + nb));
  std::vector<size_t> p;
  std::vector<uint8_t> x, y;
  match_indices(sa, sb, p, x, y);
  *nperm = (int32_t)p.size();
  for (size_t i = 0; i < p.size(); ++i) perm[i] = (int64_t)p[i];
  for (int i = 0; i < na; ++i) ma[i] = x[i];
  for (int i = 0; i < nb; ++i) mb[i] = y[i];
  return 0;
  ORC_CATCH(ERR_OTHER)
}

// mode 0: CoreFlatFiberIterator::new(fiber, conj=false)   1: conj=true
//      2: new_paired_conjugates().0 (fixed dims)          3: .1 (free dims)
// use_single: emulate a fiber whose is_single() was computed (BoolFilter/BitVec constructors)
int64_t orc_flat_iter(const Slot* s, int n, const uint8_t* free_mask, int mode, int use_single, uint64_t shift,
                      uint64_t* out, int64_t cap) {
  ORC_TRY
  Structure st = structure_from_slots(std::vector<Slot>(s, s + n));
  std::vector<uint8_t> fr(free_mask, free_mask + n);
  FlatIt it;
  if (mode <= 1) {
    int nfree = 0, pos = -1;
    for (int i = 0; i < n; ++i)
      if (fr[i]) {
        nfree++;
        pos = i;
      }
    if (use_single && nfree == 1)
      it = flat_it_single(st.strides(), (size_t)pos, st.shape(), mode == 1);
    else
      it = flat_it_multi(st.strides(), st.shape(), fr, mode == 1);
  } else {
    auto pr = flat_it_paired_conjugates(st.strides(), st.shape(), fr);
    it = mode == 2 ? pr.first : pr.second;
  }
  it.shift(shift);
  int64_t c = 0;
  size_t f;
  while (it.next(f)) {
    if (c < cap) out[c] = f;
    c++;
    if (c > (int64_t)1 << 26) break;  // runaway guard
  }
  return c;
  ORC_CATCH(-1)
}

// Same iterator driven by raw dims in the order GIVEN (row-major strides of that order, no
// canonical sort): the reference's legacy snapshot tests (tests.rs:174-269) iterate fibres of
// unsorted structures; the iterators themselves only see strides and shape.
int64_t orc_flat_iter_dims(const uint64_t* dims, int n, const uint8_t* free_mask, int mode, int use_single,
                           uint64_t shift, uint64_t* out, int64_t cap) {
  ORC_TRY
  std::vector<size_t> shape(dims, dims + n), strides(n, 1);
  for (int i = n - 1; i-- > 0;) strides[i] = strides[i + 1] * shape[i + 1];
  std::vector<uint8_t> fr(free_mask, free_mask + n);
  FlatIt it;
  if (mode <= 1) {
    int nfree = 0, pos = -1;
    for (int i = 0; i < n; ++i)
      if (fr[i]) {
        nfree++;
        pos = i;
      }
    if (use_single && nfree == 1)
      it = flat_it_single(strides, (size_t)pos, shape, mode == 1);
    else
      it = flat_it_multi(strides, shape, fr, mode == 1);
  } else {
    auto pr = flat_it_paired_conjugates(strides, shape, fr);
    it = mode == 2 ? pr.first : pr.second;
  }
  it.shift(shift);
  int64_t c = 0;
  size_t f;
  while (it.next(f)) {
    if (c < cap) out[c] = f;
    c++;
    if (c > (int64_t)1 << 26) break;
  }
  return c;
  ORC_CATCH(-1)
}

int64_t orc_metric_iter(const Slot* s, int n, const uint8_t* free_mask, int conj, const int64_t* order, int norder,
                        uint64_t* out, uint8_t* neg, int64_t cap) {
  ORC_TRY
  Structure st = structure_from_slots(std::vector<Slot>(s, s + n));
  std::vector<uint8_t> fr(free_mask, free_mask + n);
  std::vector<size_t> ord(order, order + norder);
  MetricIt it = MetricIt::make(st, fr, conj != 0, norder ? &ord : nullptr);
  int64_t c = 0;
  size_t f;
  bool ng;
  while (it.next(f, ng)) {
    if (c < cap) {
      out[c] = f;
      neg[c] = ng;
    }
    c++;
  }
  return c;
  ORC_CATCH(-1)
}

// ---- tensors -------------------------------------------------------------
void* orc_tensor_dense(const Slot* canon, int n, const double* reim) {
  try {
    Structure st = canon_structure(canon, n);
    std::vector<C> d(st.size());
    std::memcpy(d.data(), reim, d.size() * sizeof(C));
    return new Tensor(Tensor::dense(st, std::move(d)));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void* orc_tensor_sparse(const Slot* canon, int n, const uint64_t* flat, const double* reim, uint64_t nnz) {
  try {
    Structure st = canon_structure(canon, n);
    Tensor t = Tensor::sparse_empty(st);
    for (uint64_t k = 0; k < nnz; ++k) {
      if (flat[k] >= st.size()) throw std::runtime_error("sparse flat index out of bounds");
      t.elems[flat[k]] = C{reim[2 * k], reim[2 * k + 1]};
    }
    return new Tensor(std::move(t));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_tensor_free(void* t) { delete (Tensor*)t; }
int orc_tensor_order(void* t) { return (int)((Tensor*)t)->st.order(); }
int orc_tensor_is_sparse(void* t) { return ((Tensor*)t)->sparse; }
uint64_t orc_tensor_size(void* t) { return ((Tensor*)t)->size(); }
uint64_t orc_tensor_nnz(void* t) {
  Tensor* x = (Tensor*)t;
  return x->sparse ? x->elems.size() : x->data.size();
}
void orc_tensor_slots(void* t, Slot* out) {
  Tensor* x = (Tensor*)t;
  for (size_t i = 0; i < x->st.order(); ++i) out[i] = x->st.s[i];
}
void orc_tensor_to_dense(void* t, double* out) {
  auto d = ((Tensor*)t)->to_dense_data();
  std::memcpy(out, d.data(), d.size() * sizeof(C));
}
void orc_tensor_entries(void* t, uint64_t* flat, double* reim) {
  Tensor* x = (Tensor*)t;
  size_t k = 0;
  for_each_flat(*x, [&](size_t f, C v) {
    flat[k] = f;
    reim[2 * k] = v.re;
    reim[2 * k + 1] = v.im;
    k++;
  });
}

int orc_contract(void* a, void* b, void** out, int emulate_reference_pairing) {
  ORC_TRY
  OracleOptions opt;
  opt.emulate_reference_interleaved_pairing = emulate_reference_pairing != 0;
  *out = new Tensor(contract(*(Tensor*)a, *(Tensor*)b, opt));
  return 0;
  ORC_CATCH(ERR_OTHER)
}
// mode: bit 0 = subtract; bit 1 = FallibleAdd/FallibleSub semantics (fresh tensor, smart_set drops exact zeros)
int orc_add(void* a, void* b, void** out, int mode) {
  ORC_TRY
  if (mode & 2) {
    *out = new Tensor(add_fallible(*(Tensor*)a, *(Tensor*)b, (mode & 1) != 0));
    return 0;
  }
  Tensor t = *(Tensor*)a;
  add_assign(t, *(Tensor*)b, (mode & 1) != 0);
  *out = new Tensor(std::move(t));
  return 0;
  ORC_CATCH(ERR_OTHER)
}
int orc_neg(void* a, void** out) {
  ORC_TRY
  *out = new Tensor(neg(*(Tensor*)a));
  return 0;
  ORC_CATCH(ERR_OTHER)
}
int orc_scalar_mul(void* a, double re, double im, void** out) {
  ORC_TRY
  *out = new Tensor(scalar_mul(*(Tensor*)a, C{re, im}));
  return 0;
  ORC_CATCH(ERR_OTHER)
}
// slots here need NOT be canonical-with-distinct entries: repeated slots mark the trace pairs
void* orc_tensor_dense_raw(const Slot* slots, int n, const double* reim) {
  try {
    Structure st;
    st.s.assign(slots, slots + n);
    std::vector<C> d(st.size());
    std::memcpy(d.data(), reim, d.size() * sizeof(C));
    return new Tensor(Tensor::dense(st, std::move(d)));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
int orc_trace(void* a, void** out) {
  ORC_TRY
  *out = new Tensor(internal_contract(*(Tensor*)a));
  return 0;
  ORC_CATCH(ERR_OTHER)
}

// ---- library -------------------------------------------------------------
void* orc_lib_builtin(int which) {
  switch (which) {
    case 0: return new LibTensor(gamma_weyl());
    case 1: return new LibTensor(gamma_dirac());
    case 2: return new LibTensor(gamma0_weyl());
    case 3: return new LibTensor(gamma5_weyl());
    case 4: return new LibTensor(projm_weyl());
    case 5: return new LibTensor(projp_weyl());
    case 7: return new LibTensor(gamma_conj_weyl());
    case 8: return new LibTensor(gamma_adj_weyl());
    default: return nullptr;
  }
}
void* orc_lib_metric(const Slot* rep) { return new LibTensor(metric_lib(*rep)); }
// user library tensor: reps in argument order; sparse entries by argument-order flat index
void* orc_lib_custom(const Slot* reps, int n, const uint64_t* flat, const double* reim, uint64_t nnz, int sparse) {
  LibTensor* lt = new LibTensor();
  lt->reps.assign(reps, reps + n);
  lt->sparse = sparse != 0;
  if (sparse) {
    for (uint64_t k = 0; k < nnz; ++k) lt->nz.emplace_back((size_t)flat[k], C{reim[2 * k], reim[2 * k + 1]});
  } else {
    lt->data.resize(nnz);
    std::memcpy(lt->data.data(), reim, nnz * sizeof(C));
  }
  return lt;
}
void orc_lib_free(void* l) { delete (LibTensor*)l; }
void* orc_lib_instantiate(void* l, const Slot* slots, int n) {
  try {
    return new Tensor(lib_instantiate(*(LibTensor*)l, std::vector<Slot>(slots, slots + n)));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// ---- network -------------------------------------------------------------
void* orc_net_new() {
  NetBox* b = new NetBox();
  return b;
}
void orc_net_free(void* n) { delete (NetBox*)n; }
int orc_net_add_tensor(void* n, void* t) {
  NetBox* b = (NetBox*)n;
  Node nd;
  nd.kind = N_TENSOR;
  nd.ref = b->net.push_tensor(*(Tensor*)t);
  return b->net.add_node(nd);
}
int orc_net_add_scalar(void* n, double re, double im) {
  NetBox* b = (NetBox*)n;
  Node nd;
  nd.kind = N_SCALAR;
  nd.ref = b->net.push_scalar(C{re, im});
  return b->net.add_node(nd);
}
int orc_net_add_lib(void* n, void* lib, const Slot* slots, int ns) {
  NetBox* b = (NetBox*)n;
  b->library.push_back(*(LibTensor*)lib);
  LibUse u;
  u.lib = (int)b->library.size() - 1;
  u.slots.assign(slots, slots + ns);
  b->net.lib_uses.push_back(u);
  Node nd;
  nd.kind = N_LIB;
  nd.ref = (int)b->net.lib_uses.size() - 1;
  return b->net.add_node(nd);
}
int orc_net_add_op(void* n, int kind, const int32_t* children, int nc, int power) {
  NetBox* b = (NetBox*)n;
  Node nd;
  nd.kind = kind;
  nd.power = power;
  nd.children.assign(children, children + nc);
  return b->net.add_node(nd);
}
void orc_net_set_root(void* n, int node) { ((NetBox*)n)->net.root = node; }
void orc_net_set_emulate(void* n, int e) { ((NetBox*)n)->net.opt.emulate_reference_interleaved_pairing = e != 0; }
int orc_net_execute(void* n) {
  ORC_TRY
  NetBox* b = (NetBox*)n;
  b->net.library = &b->library;
  b->net.execute();
  return 0;
  ORC_CATCH(ERR_OTHER)
}
void* orc_net_result(void* n) {
  try {
    NetBox* b = (NetBox*)n;
    b->net.library = &b->library;
    return new Tensor(b->net.result_tensor());
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
int orc_net_log(void* n, int32_t* pairs, int cap) {
  NetBox* b = (NetBox*)n;
  int c = 0;
  for (auto& p : b->net.contraction_log) {
    if (c < cap) {
      pairs[2 * c] = p.first;
      pairs[2 * c + 1] = p.second;
    }
    c++;
  }
  return c;
}

// CPU baseline: evaluate the (un-executed) template network at n_points points,
// one point at a time, exactly like the reference (fresh network + store per
// point, library tensors cloned+permuted per use, every intermediate freshly
// allocated).  batched_nodes[k] is a N_TENSOR leaf whose dense data is replaced
// by data[k][point*size .. (point+1)*size) (interleaved re/im).  out receives
// out_size complex numbers per point.  sum_out (optional, 2*out_size doubles)
// receives the sum over points.
int orc_net_eval_batch(void* n, const int32_t* batched_nodes, int nb, const double* const* data, uint64_t n_points,
                       double* out, uint64_t out_size, int n_threads, double* sum_out) {
  ORC_TRY
  NetBox* b = (NetBox*)n;
  b->net.library = &b->library;
  if (n_threads < 1) n_threads = 1;
  std::vector<std::string> errs(n_threads);
  std::vector<std::vector<C>> sums(n_threads, std::vector<C>(out_size, C{0, 0}));
  auto work = [&](int tid) {
    try {
      uint64_t lo = n_points * tid / n_threads, hi = n_points * (tid + 1) / n_threads;
      for (uint64_t p = lo; p < hi; ++p) {
        Network net = b->net;  // fresh network per point (symbolica_interop.rs:617-640)
        for (int k = 0; k < nb; ++k) {
          Tensor& t = net.tensors[net.nodes[batched_nodes[k]].ref];
          size_t sz = t.size();
          std::memcpy(t.data.data(), data[k] + 2 * p * sz, sz * sizeof(C));
        }
        net.execute();
        Tensor r = net.result_tensor();
        auto d = r.to_dense_data();
        if (d.size() != out_size) throw std::runtime_error("result size mismatch");
        if (out) std::memcpy(out + 2 * p * out_size, d.data(), out_size * sizeof(C));
        for (size_t k = 0; k < out_size; ++k) cadd(sums[tid][k], d[k]);
      }
    } catch (const std::exception& e) {
      errs[tid] = e.what();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) {
      g_err = e;
      return ERR_OTHER;
    }
  if (sum_out)
    for (size_t k = 0; k < out_size; ++k) {
      C s{0, 0};
      for (int t = 0; t < n_threads; ++t) cadd(s, sums[t][k]);
      sum_out[2 * k] = s.re;
      sum_out[2 * k + 1] = s.im;
    }
  return 0;
  ORC_CATCH(ERR_OTHER)
}

}  // extern "C"
