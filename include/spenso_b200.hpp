// spenso_b200.hpp — C++ host-side mirror of the reference interface for the concrete-data path, written
// above the C ABI (spenso_b200.h).  The reference is Rust; its toolchain is absent from this image, so this
// header plays the role the Rust shim of INTEGRATION.md plays there: same names, argument meaning and error
// behaviour, so code written against it reads like the reference's own tests.
//
//   spenso::Slot / Representation            spenso/src/structure/{slot,representation}.rs
//   spenso::DenseTensor / SparseTensor        spenso/src/tensors/data/{dense,sparse}.rs   (one point)
//   spenso::BatchTensor                       one DenseTensor per phase-space point, device resident
//   a.contract(b)                             trait Contract, spenso/src/contraction.rs:32-35
//   spenso::ContractionError                  spenso/src/contraction.rs:20-30
//   spenso::Network (+, *, execute, result)   spenso/src/network.rs:47-51, 1217-1241
//
// Header only; link with -lspenso_b200.  Nothing here computes: every numeric call goes to the library.
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "spenso_b200.h"

namespace spenso {

using Complex = std::complex<double>;  // Complex<f64>: interleaved (re, im)

// ContractionError / TensorNetworkError: the status code says which variant
struct ContractionError : std::runtime_error {
  enum Kind { EmptySparse = 1, StructureError = 2, DimensionError = 3, Other = 4, Cuda = 5, NoDevice = 6, Invalid = 7, Jit = 8 };
  Kind kind;
  ContractionError(int code, const std::string& msg) : std::runtime_error(msg), kind(static_cast<Kind>(code < 0 ? -code : code)) {}
};
inline void check(int rc) {
  if (rc != SPN_OK) throw ContractionError(rc, spn_last_error());
}
inline int node(int rc) {
  if (rc < 0) throw ContractionError(rc, spn_last_error());
  return rc;
}

// ---- representations and slots (ids as registered by idenso's initialize()) --------------------
struct Representation {
  uint8_t kind;
  int16_t id;
  Representation dual() const { return kind == SPN_REP_DUALIZABLE ? Representation{kind, (int16_t)-id} : *this; }
  spn_slot new_slot(uint32_t dim, uint64_t aind) const { return spn_slot{kind, SPN_AIND_NORMAL, id, dim, aind}; }
};
constexpr Representation Euclidean{SPN_REP_SELF_DUAL, 0}, Bispinor{SPN_REP_SELF_DUAL, 1}, ColorAdjoint{SPN_REP_SELF_DUAL, 2},
    Minkowski{SPN_REP_INLINE_METRIC, 0}, Lorentz{SPN_REP_DUALIZABLE, 1}, SpinFundamental{SPN_REP_DUALIZABLE, 2},
    ColorFundamental{SPN_REP_DUALIZABLE, 3}, ColorSextet{SPN_REP_DUALIZABLE, 4};
using Slot = spn_slot;
using Slots = std::vector<Slot>;

// PermutedStructure::from(Vec<Slot>): canonical slots and the permutation that sorts the given ones
struct PermutedStructure {
  Slots structure;
  std::vector<int32_t> index_permutation;  // structure[i] = given[index_permutation[i]]
  static PermutedStructure from(const Slots& given) {
    PermutedStructure p;
    p.structure.resize(given.size());
    p.index_permutation.resize(given.size());
    int64_t bs, ds;
    check(spn_structure_canonicalize(given.data(), (uint32_t)given.size(), p.structure.data(), p.index_permutation.data(), &bs, &ds));
    return p;
  }
  std::vector<uint64_t> strides() const {  // row major over the canonical structure
    std::vector<uint64_t> st(structure.size(), 1);
    for (size_t i = structure.size(); i-- > 1;) st[i - 1] = st[i] * structure[i].dim;
    return st;
  }
};

// ---- context ------------------------------------------------------------------------------------------------
class Context {
 public:
  explicit Context(int device = 0) { check(spn_ctx_create(device, &h_)); }
  ~Context() { spn_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  spn_ctx* raw() const { return h_; }
  void synchronize() const { check(spn_ctx_synchronize(h_)); }
  uint64_t launch_count() const { return spn_ctx_launch_count(h_); }

 private:
  spn_ctx* h_ = nullptr;
};

// ---- tensors ----------------------------------------------------------------------------------------------------
// One handle type for the three storages of the path (DataTensor::{Dense,Sparse} at one point, and the batched
// dense tensor).  Copying shares the device object.
class Tensor {
 public:
  Tensor() = default;
  Tensor(const Context& ctx, spn_tensor* h) : ctx_(&ctx), h_(h, spn_tensor_free) {}

  // DenseTensor::from_data(data, structure): data row major over the slots AS GIVEN; permuted into canonical
  // order like PermutedStructure + permute_inds (tensors/data/dense.rs:93-109)
  static Tensor dense(const Context& ctx, const Slots& given, const std::vector<Complex>& data) {
    PermutedStructure p = PermutedStructure::from(given);
    std::vector<Complex> canon = to_canonical(given, p, data, 1);
    spn_tensor* h = nullptr;
    check(spn_tensor_shared_dense(ctx.raw(), p.structure.data(), (uint32_t)given.size(), reinterpret_cast<const double*>(canon.data()), &h));
    return Tensor(ctx, h);
  }
  // SparseTensor: entries keyed by the expanded index in the given slot order (SparseTensor::set)
  static Tensor sparse(const Context& ctx, const Slots& given, const std::map<std::vector<uint32_t>, Complex>& entries) {
    PermutedStructure p = PermutedStructure::from(given);
    auto st = p.strides();
    std::vector<uint64_t> flat;
    std::vector<Complex> vals;
    for (auto& kv : entries) {
      uint64_t f = 0;
      for (size_t c = 0; c < given.size(); ++c) f += (uint64_t)kv.first[(size_t)p.index_permutation[c]] * st[c];
      flat.push_back(f);
      vals.push_back(kv.second);
    }
    spn_tensor* h = nullptr;
    check(spn_tensor_shared_sparse(ctx.raw(), p.structure.data(), (uint32_t)given.size(), flat.data(),
                                   reinterpret_cast<const double*>(vals.data()), flat.size(), &h));
    return Tensor(ctx, h);
  }
  // n_points DenseTensors; data is [point][row major over the given slots]
  static Tensor batched(const Context& ctx, const Slots& given, uint64_t n_points, const std::vector<Complex>& data,
                        int layout = SPN_POINT_MAJOR) {
    PermutedStructure p = PermutedStructure::from(given);
    std::vector<Complex> canon = to_canonical(given, p, data, n_points);
    spn_tensor* h = nullptr;
    check(spn_tensor_batched(ctx.raw(), p.structure.data(), (uint32_t)given.size(), n_points, SPN_C64, layout, &h));
    Tensor t(ctx, h);
    check(spn_tensor_upload(h, canon.data(), 0, n_points));
    return t;
  }
  static Tensor batched_empty(const Context& ctx, const Slots& canonical, uint64_t n_points, int layout = SPN_POINT_MAJOR) {
    spn_tensor* h = nullptr;
    check(spn_tensor_batched(ctx.raw(), canonical.data(), (uint32_t)canonical.size(), n_points, SPN_C64, layout, &h));
    return Tensor(ctx, h);
  }

  spn_tensor* raw() const { return h_.get(); }
  Slots structure() const {
    Slots s(spn_tensor_order(raw()));
    check(spn_tensor_slots(raw(), s.data()));
    return s;
  }
  uint64_t size() const { return spn_tensor_size(raw()); }
  uint64_t n_points() const { return spn_tensor_n_points(raw()); }
  bool is_sparse() const { return spn_tensor_storage(raw()) == SPN_STORAGE_SHARED_SPARSE; }
  bool is_batched() const { return spn_tensor_storage(raw()) == SPN_STORAGE_BATCHED; }
  // host copy in canonical order: [n_points][size]
  std::vector<Complex> to_host() const {
    std::vector<Complex> out(n_points() * size());
    check(spn_tensor_download(raw(), out.data(), 0, n_points()));
    return out;
  }
  std::map<uint64_t, Complex> elements() const {  // SparseTensor.elements, keyed by canonical flat index
    uint64_t n = spn_tensor_nnz(raw());
    std::vector<uint64_t> f(n);
    std::vector<Complex> v(n);
    check(spn_tensor_sparse_entries(raw(), f.data(), reinterpret_cast<double*>(v.data())));
    std::map<uint64_t, Complex> m;
    for (uint64_t i = 0; i < n; ++i) m[f[i]] = v[i];
    return m;
  }

  // trait Contract: fn contract(&self, other: &T) -> Result<Self::LCM, ContractionError>
  Tensor contract(const Tensor& other) const {
    spn_tensor* out = nullptr;
    check(spn_tensor_contract(ctx_->raw(), raw(), other.raw(), &out));
    return Tensor(*ctx_, out);
  }
  Tensor operator+(const Tensor& o) const { return binary(o, 0); }
  Tensor operator-(const Tensor& o) const { return binary(o, 1); }
  Tensor operator-() const {
    spn_tensor* out = nullptr;
    check(spn_tensor_neg(ctx_->raw(), raw(), &out));
    return Tensor(*ctx_, out);
  }
  Tensor scalar_mul(Complex s) const {  // trait ScalarMul
    spn_tensor* out = nullptr;
    check(spn_tensor_scalar_mul(ctx_->raw(), raw(), s.real(), s.imag(), &out));
    return Tensor(*ctx_, out);
  }
  Tensor conj() const {
    spn_tensor* out = nullptr;
    check(spn_tensor_conj(ctx_->raw(), raw(), &out));
    return Tensor(*ctx_, out);
  }
  // FallibleAdd / FallibleSub (algebra/fallible_add.rs, fallible_sub.rs): Sparse (+/-) Sparse does not store exact zeros
  Tensor add_fallible(const Tensor& o) const { return binary(o, SPN_ADD_FALLIBLE); }
  Tensor sub_fallible(const Tensor& o) const { return binary(o, SPN_ADD_FALLIBLE | SPN_ADD_SUBTRACT); }
  const Context& context() const { return *ctx_; }

 private:
  Tensor binary(const Tensor& o, int sub) const {
    spn_tensor* out = nullptr;
    check(spn_tensor_add(ctx_->raw(), raw(), o.raw(), sub, &out));
    return Tensor(*ctx_, out);
  }
  static std::vector<Complex> to_canonical(const Slots& given, const PermutedStructure& p, const std::vector<Complex>& data,
                                           uint64_t n_points) {
    const size_t n = given.size();
    uint64_t size = 1;
    for (auto& s : given) size *= s.dim;
    if (data.size() != size * n_points) throw ContractionError(SPN_ERR_DIMENSION, "data length does not match the structure");
    std::vector<uint64_t> gst(n, 1);
    for (size_t i = n; i-- > 1;) gst[i - 1] = gst[i] * given[i].dim;
    std::vector<Complex> out(data.size());
    std::vector<uint32_t> idx(n, 0);
    for (uint64_t f = 0; f < size; ++f) {  // f: canonical flat index, idx: canonical expanded index
      uint64_t g = 0;
      for (size_t c = 0; c < n; ++c) g += (uint64_t)idx[c] * gst[(size_t)p.index_permutation[c]];
      for (uint64_t pt = 0; pt < n_points; ++pt) out[pt * size + f] = data[pt * size + g];
      for (size_t c = n; c-- > 0;) {
        if (++idx[c] < p.structure[c].dim) break;
        idx[c] = 0;
      }
    }
    return out;
  }
  const Context* ctx_ = nullptr;
  std::shared_ptr<spn_tensor> h_;
};
using DenseTensor = Tensor;
using SparseTensor = Tensor;
using BatchTensor = Tensor;

// ---- Monte-Carlo sum exchange between the GPUs of one node (one process per GPU) ----------------------------------
// all_gather(mine) -> the 64-byte handles of every rank, in rank order, over any host channel (MPI, sockets, a file)
class PeerComm {
 public:
  template <class AllGather>
  PeerComm(const Context& ctx, int rank, int world, uint32_t max_f64, AllGather&& all_gather) {
    check(spn_comm_create(ctx.raw(), rank, world, max_f64, &h_));
    if (world > 1) {
      std::vector<unsigned char> mine(SPN_COMM_HANDLE_BYTES);
      check(spn_comm_handle(h_, mine.data()));
      std::vector<unsigned char> all = all_gather(mine);
      if (all.size() != (size_t)world * SPN_COMM_HANDLE_BYTES) throw ContractionError(SPN_ERR_INVALID, "PeerComm: bad handle table");
      check(spn_comm_connect(h_, all.data()));
    }
  }
  ~PeerComm() { spn_comm_destroy(h_); }
  PeerComm(const PeerComm&) = delete;
  PeerComm& operator=(const PeerComm&) = delete;
  void allreduce_sum(double* inout_device, uint32_t n_f64) { check(spn_comm_allreduce_sum(h_, inout_device, n_f64)); }
  int status() const { return spn_comm_status(h_); }
  spn_comm* raw() const { return h_; }

 private:
  spn_comm* h_ = nullptr;
};

// ---- network ------------------------------------------------------------------------------------------------------
// Network::from_tensor(t1) * Network::from_tensor(t2) + ... then execute(): the expression is recorded, compiled
// once and evaluated for every point by the device executor.
class Network {
 public:
  explicit Network(const Context& ctx) : ctx_(&ctx) {
    spn_net* h = nullptr;
    check(spn_net_create(ctx.raw(), &h));
    h_.reset(h, spn_net_destroy);
  }
  struct Expr {  // a node of the expression graph
    Network* net;
    int id;
    Expr operator*(const Expr& o) const { return net->op(SPN_OP_PRODUCT, {id, o.id}); }
    Expr operator+(const Expr& o) const { return net->op(SPN_OP_SUM, {id, o.id}); }
    Expr operator-() const { return net->op(SPN_OP_NEG, {id}); }
    Expr operator-(const Expr& o) const { return *this + (-o); }
    Expr pow(int n) const { return net->op(SPN_OP_POWER, {id}, n); }
    Expr conj() const { return net->op(SPN_OP_CONJ, {id}); }
  };
  // a per-point tensor: becomes the next input of execute()
  Expr from_tensor(const Tensor& t) {
    if (t.is_batched()) {
      Slots s = t.structure();
      inputs_.push_back(t);
      return Expr{this, node(spn_net_leaf_batched(h_.get(), s.data(), (uint32_t)s.size(), SPN_C64))};
    }
    keep_.push_back(t);
    return Expr{this, node(spn_net_leaf_shared(h_.get(), t.raw()))};
  }
  Expr reindex(const Expr& leaf, const Slots& s) { return Expr{this, node(spn_net_leaf_reindex(h_.get(), leaf.id, s.data(), (uint32_t)s.size()))}; }
  Expr library(int which, const Slots& s) { return Expr{this, node(spn_net_leaf_library(h_.get(), which, s.data(), (uint32_t)s.size()))}; }
  Expr scalar(Complex c) { return Expr{this, node(spn_net_leaf_scalar(h_.get(), c.real(), c.imag()))}; }
  Expr op(int kind, std::vector<int32_t> children, int power = 1) {
    return Expr{this, node(spn_net_op(h_.get(), kind, children.data(), (uint32_t)children.size(), power))};
  }
  // Network::execute::<Sequential, SmallestDegree, ..>() for every point, then result_tensor()
  Tensor execute(const Expr& root, int exec_mode = SPN_EXEC_FUSED) {
    check(spn_net_set_root(h_.get(), root.id));
    check(spn_net_compile(h_.get(), exec_mode));
    Slots rs(spn_net_result_order(h_.get()));
    check(spn_net_result_slots(h_.get(), rs.data()));
    uint64_t n = inputs_.empty() ? 1 : inputs_[0].n_points();
    Tensor out = Tensor::batched_empty(*ctx_, rs, n);
    std::vector<const spn_tensor*> in;
    for (auto& t : inputs_) in.push_back(t.raw());
    check(spn_net_execute(h_.get(), in.data(), (uint32_t)in.size(), out.raw(), nullptr, 0, 0));
    return out;
  }
  spn_net* raw() const { return h_.get(); }
  // with a communicator attached, sums over points are sums over the points of all ranks (spn_net_set_comm)
  void set_comm(spn_comm* comm) { check(spn_net_set_comm(h_.get(), comm)); }

 private:
  const Context* ctx_;
  std::shared_ptr<spn_net> h_;
  std::vector<Tensor> inputs_, keep_;
};

}  // namespace spenso
