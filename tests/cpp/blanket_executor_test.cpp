// The reference's blanket executor, call for call, over the C ABI.
//
// spenso's `impl ExecuteOp for NetworkStore<T, Sc>` (spenso/src/network.rs:1244-1706) evaluates a network by graph
// surgery: find the first op whose children are all leaves (network/graph.rs:433-513), run it on the store's tensors
// with the trait methods of T — Neg, Clone, Contract, AddAssign, ScalarMul, From<library tensor>, and for Sc: Neg, Div,
// MulAssign, AddAssign — push the result into the append-only store, replace op + children by the new leaf.  The
// Product op is SmallestDegree (network/contract.rs:43-498): ContractScalars, then SingleSmallestDegree until one leaf
// is left.
//
// A Rust shim that implements those traits for a device tensor (integration/rust) makes exactly the
// calls this file makes.  There is no Rust toolchain in this image, so the executor is restated here in C++ with
// T = a device tensor handle and every trait method = one C-ABI call:
//     T::neg -> spn_tensor_neg          T::contract -> spn_tensor_contract      T += &U -> spn_tensor_add
//     T::scalar_mul -> spn_tensor_contract with a rank-0 tensor                 T::from(lib) -> spn_tensor_library
//     fn_lib conj -> spn_tensor_conj    Sc = rank-0 tensor: *=, +=, neg, 1/s -> contract, add, neg, spn_tensor_recip
// and checked against the engine's own executor (spn_net_*): the contraction log of this run is handed to
// spn_net_set_schedule, after which the stepwise executor issues the same pairwise kernels in the same order —
// bit-identical results where no scalar factor is folded in — and the fused kernel agrees to 1e-12.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>

#include "../../include/spenso_b200.hpp"

using namespace spenso;

static int failures = 0;
#define EXPECT(cond)                                                   \
  do {                                                                 \
    if (!(cond)) {                                                     \
      std::printf("  FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                      \
    }                                                                  \
  } while (0)

// ---- T: the tensor type plugged into the executor ------------------------------------------------------------------
struct T {
  Tensor t;
  T() = default;
  explicit T(Tensor x) : t(std::move(x)) {}
  bool is_scalar() const { return spn_tensor_order(t.raw()) == 0; }
  T neg() const { return T(-t); }                                          // Neg
  T contract(const T& o) const { return T(t.contract(o.t)); }              // Contract
  void add_assign(const T& o) { t = t + o.t; }                             // AddAssign<T::Ref>
  T scalar_mul(const T& s) const { return T(t.contract(s.t)); }            // ScalarMul<Sc>: Sc is a rank-0 tensor
  T conj() const { return T(t.conj()); }                                   // FUN_LIB conj (spenso-hep-lib/src/lib.rs:511-520)
  T recip() const {                                                        // Sc::ref_one() / s
    spn_tensor* out = nullptr;
    check(spn_tensor_recip(t.context().raw(), t.raw(), &out));
    return T(Tensor(t.context(), out));
  }
};

// ---- the expression graph (NetworkGraph restricted to what the executor reads) -------------------------------------
enum Kind { SCALAR, LIBKEY, LOCAL, OP };
struct GNode {
  Kind kind = LOCAL;
  int op = 0, power = 1;   // OP
  int ref = -1;            // SCALAR: store.scalar index; LOCAL: store.tensors index; LIBKEY: lib_uses index
  std::vector<int> children;
};
struct LibUse {
  int which;
  Slots slots;
};
struct Executor {
  const Context& ctx;
  std::vector<GNode> nodes;
  std::vector<T> tensors, scalar;   // NetworkStore { tensors, scalar } — append only (store.rs:125-129)
  std::vector<LibUse> lib_uses;
  std::vector<std::pair<int, int>> log;   // (n1, n2) of every pairwise contraction: what spn_net_set_schedule takes
  int root = -1;
  explicit Executor(const Context& c) : ctx(c) {}

  int add(GNode n) {
    nodes.push_back(std::move(n));
    return (int)nodes.size() - 1;
  }
  int local(Tensor t) {
    tensors.emplace_back(std::move(t));
    GNode n;
    n.kind = LOCAL;
    n.ref = (int)tensors.size() - 1;
    return add(n);
  }
  int scalar_leaf(Complex c) {
    scalar.emplace_back(Tensor::dense(ctx, {}, {c}));
    GNode n;
    n.kind = SCALAR;
    n.ref = (int)scalar.size() - 1;
    return add(n);
  }
  int library(int which, Slots s) {
    lib_uses.push_back({which, std::move(s)});
    GNode n;
    n.kind = LIBKEY;
    n.ref = (int)lib_uses.size() - 1;
    return add(n);
  }
  int op(int kind, std::vector<int> ch, int power = 1) {
    GNode n;
    n.kind = OP;
    n.op = kind;
    n.power = power;
    n.children = std::move(ch);
    return add(n);
  }
  bool is_leaf(int id) const { return nodes[(size_t)id].kind != OP; }
  // graph.get_lib_data(lib, id) + T::from(..): network/graph.rs:291-322
  T lib_data(int id) const {
    const LibUse& u = lib_uses[(size_t)nodes[(size_t)id].ref];
    spn_tensor* out = nullptr;
    check(spn_tensor_library(ctx.raw(), u.which, u.slots.data(), (uint32_t)u.slots.size(), &out));
    return T(Tensor(ctx, out));
  }
  T leaf_tensor(int id) const { return nodes[(size_t)id].kind == LOCAL ? tensors[(size_t)nodes[(size_t)id].ref] : lib_data(id); }
  Slots leaf_slots(int id) const {
    const GNode& n = nodes[(size_t)id];
    if (n.kind == LOCAL) return tensors[(size_t)n.ref].t.structure();
    if (n.kind == LIBKEY) return PermutedStructure::from(lib_uses[(size_t)n.ref].slots).structure;
    return {};
  }
  void become_leaf(int id, Kind k, int ref) {
    nodes[(size_t)id].kind = k;
    nodes[(size_t)id].ref = ref;
    nodes[(size_t)id].children.clear();
  }
  int push_tensor(T t) {
    tensors.push_back(std::move(t));
    return (int)tensors.size() - 1;
  }
  int push_scalar(T s) {
    scalar.push_back(std::move(s));
    return (int)scalar.size() - 1;
  }

  // merge_ops (graph.rs:1002-1065)
  void merge_ops(int id) {
    if (is_leaf(id)) return;
    for (int c : std::vector<int>(nodes[(size_t)id].children)) merge_ops(c);
    const int k = nodes[(size_t)id].op;
    if (k == SPN_OP_PRODUCT || k == SPN_OP_SUM) {
      std::vector<int> flat;
      for (int c : nodes[(size_t)id].children) {
        if (!is_leaf(c) && nodes[(size_t)c].op == k)
          for (int g : nodes[(size_t)c].children) flat.push_back(g);
        else
          flat.push_back(c);
      }
      nodes[(size_t)id].children = flat;
    }
  }
  bool find_ready(int id, int& found) const {   // extract_next_ready_op
    if (is_leaf(id)) return false;
    bool all = true;
    for (int c : nodes[(size_t)id].children) all = all && is_leaf(c);
    if (all) {
      found = id;
      return true;
    }
    for (int c : nodes[(size_t)id].children)
      if (find_ready(c, found)) return true;
    return false;
  }

  // ContractScalars::contract (contract.rs:73-208)
  void contract_scalars(std::vector<int>& leaves) {
    std::vector<int> scal, tens;
    for (int l : leaves) (nodes[(size_t)l].kind == SCALAR ? scal : tens).push_back(l);
    if (scal.empty()) return;
    const bool remove_op = tens.size() <= 1;
    const int f = scal.back();
    T acc = scalar[(size_t)nodes[(size_t)f].ref];
    bool more = false;
    for (size_t k = 0; k + 1 < scal.size(); ++k) {
      more = true;
      acc = acc.contract(scalar[(size_t)nodes[(size_t)scal[k]].ref]);   // acc *= s
    }
    if (remove_op) {
      if (tens.size() == 1) {
        int pos = push_tensor(leaf_tensor(tens[0]).scalar_mul(acc));
        become_leaf(tens[0], LOCAL, pos);
        leaves = {tens[0]};
      } else {
        become_leaf(f, SCALAR, push_scalar(acc));
        leaves = {f};
      }
    } else if (more) {
      become_leaf(f, SCALAR, push_scalar(acc));
      leaves = tens;
      leaves.push_back(f);
    }
  }
  static bool same_slot(const Slot& a, const Slot& b) {
    return a.rep_kind == b.rep_kind && a.rep_id == b.rep_id && a.dim == b.dim && a.aind_kind == b.aind_kind && a.aind == b.aind;
  }
  // SingleSmallestDegree::contract (contract.rs:346-497)
  bool single_smallest_degree(std::vector<int>& leaves) {
    std::vector<int> tens;
    for (int l : leaves)
      if (nodes[(size_t)l].kind != SCALAR) tens.push_back(l);
    std::vector<Slots> sts;
    for (int l : tens) sts.push_back(leaf_slots(l));
    int last = -1, best1 = -1, best2 = -1;
    long best_degree = -1;
    for (size_t a = 0; a < tens.size(); ++a) {
      long degree = 0;
      int partner = -1;
      for (const Slot& s : sts[a]) {
        Slot d = s;
        if (d.rep_kind == SPN_REP_DUALIZABLE) d.rep_id = (int16_t)-d.rep_id;
        for (size_t b = 0; b < tens.size(); ++b) {
          if (b == a) continue;
          bool hit = false;
          for (const Slot& o : sts[b]) hit = hit || same_slot(o, d);
          if (hit) {
            ++degree;
            partner = (int)b;
            break;
          }
        }
      }
      int nid2;
      if (degree == 0) {
        degree = 0x7fffffffL;
        if (last >= 0) {
          nid2 = last;
        } else {
          last = (int)a;
          continue;
        }
      } else {
        nid2 = partner;
      }
      last = (int)a;
      if (best_degree < 0 || degree < best_degree) {
        best_degree = degree;
        best1 = (int)a;
        best2 = nid2;
      }
    }
    if (best1 < 0) return false;
    const int n1 = tens[(size_t)best1], n2 = tens[(size_t)best2];
    // (LibraryKey, LocalTensor) runs local.contract(lib) (contract.rs:428-443), everything else n1.contract(n2)
    T r = (nodes[(size_t)n1].kind == LIBKEY && nodes[(size_t)n2].kind == LOCAL) ? tensors[(size_t)nodes[(size_t)n2].ref].contract(lib_data(n1))
                                                                                : leaf_tensor(n1).contract(leaf_tensor(n2));
    log.emplace_back(n1, n2);
    become_leaf(n1, LOCAL, push_tensor(std::move(r)));
    leaves.erase(std::find(leaves.begin(), leaves.end(), n2));
    return true;
  }

  void execute_op(int id) {
    GNode n = nodes[(size_t)id];
    switch (n.op) {
      case SPN_OP_PRODUCT: {   // SmallestDegree::contract (contract.rs:239-263)
        std::vector<int> leaves = n.children;
        contract_scalars(leaves);
        while (single_smallest_degree(leaves)) {
        }
        contract_scalars(leaves);
        EXPECT(leaves.size() == 1);
        become_leaf(id, nodes[(size_t)leaves[0]].kind, nodes[(size_t)leaves[0]].ref);
        return;
      }
      case SPN_OP_SUM: {   // network.rs:1342-1467 (tensor mode; the scalar mode adds rank-0 tensors the same way)
        const int first = n.children[0];
        T acc = nodes[(size_t)first].kind == SCALAR ? scalar[(size_t)nodes[(size_t)first].ref] : leaf_tensor(first);
        const bool scalar_mode = nodes[(size_t)first].kind == SCALAR || acc.is_scalar();
        for (size_t k = 1; k < n.children.size(); ++k) {
          const int c = n.children[k];
          acc.add_assign(nodes[(size_t)c].kind == SCALAR ? scalar[(size_t)nodes[(size_t)c].ref] : leaf_tensor(c));
        }
        if (scalar_mode)
          become_leaf(id, SCALAR, push_scalar(acc));
        else
          become_leaf(id, LOCAL, push_tensor(acc));
        return;
      }
      case SPN_OP_NEG: {   // network.rs:1288-1337
        const int c = n.children[0];
        if (nodes[(size_t)c].kind == SCALAR)
          become_leaf(id, SCALAR, push_scalar(scalar[(size_t)nodes[(size_t)c].ref].neg()));
        else
          become_leaf(id, LOCAL, push_tensor(leaf_tensor(c).neg()));
        return;
      }
      case SPN_OP_CONJ: {   // NetworkOp::Function, network.rs:1468-1524
        const int c = n.children[0];
        if (nodes[(size_t)c].kind == SCALAR)
          become_leaf(id, SCALAR, push_scalar(scalar[(size_t)nodes[(size_t)c].ref].conj()));
        else
          become_leaf(id, LOCAL, push_tensor(leaf_tensor(c).conj()));
        return;
      }
      case SPN_OP_POWER: {   // network.rs:1526-1703 — Scalar leaf: s * s * ... (and 1/s); even powers of a tensor: (t.t)^(n/2)
        const int c = n.children[0];
        const int pw = n.power, np = pw < 0 ? -pw : pw;
        if (nodes[(size_t)c].kind == SCALAR) {
          T s = scalar[(size_t)nodes[(size_t)c].ref], v = s;
          for (int k = 1; k < np; ++k) v = v.contract(s);
          if (pw < 0) v = v.recip();
          become_leaf(id, SCALAR, push_scalar(v));
          return;
        }
        T t = leaf_tensor(c);
        T sq = t.contract(t);   // not part of the contraction log: Power is an op of its own, not a SmallestDegree step
        T v = sq;
        for (int k = 1; k < np / 2; ++k) v = v.contract(sq);
        if (pw < 0) v = v.recip();
        become_leaf(id, SCALAR, push_scalar(v));
        return;
      }
    }
  }
  // Network::execute::<Sequential, SmallestDegree>() (network.rs:1217-1241, 1132-1155)
  T execute() {
    merge_ops(root);
    int next;
    while (find_ready(root, next)) execute_op(next);
    const GNode& r = nodes[(size_t)root];
    return r.kind == SCALAR ? scalar[(size_t)r.ref] : leaf_tensor(root);
  }
};

static std::vector<Complex> random_batch(std::mt19937_64& g, uint64_t n, uint64_t size) {
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  std::vector<Complex> v(n * size);
  for (auto& x : v) x = Complex(u(g), u(g));
  return v;
}
static double max_abs(const std::vector<Complex>& v) {
  double m = 0;
  for (auto& x : v) m = std::max(m, std::abs(x));
  return m;
}
static double max_diff(const std::vector<Complex>& a, const std::vector<Complex>& b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::abs(a[i] - b[i]));
  return m;
}

// Build the same expression twice — once for the blanket executor above, once for the engine's executor — run both,
// hand the blanket run's contraction order to the engine, compare.
struct Twin {
  const Context& ctx;
  Executor ex;
  spn_net* net = nullptr;
  std::vector<Tensor> inputs;
  std::map<int, int> net_id;   // executor node id -> engine node id
  explicit Twin(const Context& c) : ctx(c), ex(c) { check(spn_net_create(c.raw(), &net)); }
  ~Twin() { spn_net_destroy(net); }
  int batched(const Slots& given, const Tensor& t) {
    int e = ex.local(t);
    net_id[e] = node(spn_net_leaf_batched(net, given.data(), (uint32_t)given.size(), SPN_C64));
    inputs.push_back(t);
    return e;
  }
  int library(int which, const Slots& s) {
    int e = ex.library(which, s);
    net_id[e] = node(spn_net_leaf_library(net, which, s.data(), (uint32_t)s.size()));
    return e;
  }
  int scalar(Complex c) {
    int e = ex.scalar_leaf(c);
    net_id[e] = node(spn_net_leaf_scalar(net, c.real(), c.imag()));
    return e;
  }
  int op(int kind, std::vector<int> ch, int power = 1) {
    int e = ex.op(kind, ch, power);
    std::vector<int32_t> mapped;
    for (int c : ch) mapped.push_back(net_id.at(c));
    net_id[e] = node(spn_net_op(net, kind, mapped.data(), (uint32_t)mapped.size(), power));
    return e;
  }
  // returns (blanket result, engine result) on the host
  std::pair<std::vector<Complex>, std::vector<Complex>> run(int root, int exec_mode, bool with_schedule) {
    ex.root = root;
    check(spn_net_set_root(net, net_id.at(root)));
    T r = ex.execute();
    if (with_schedule) {
      std::vector<int32_t> pairs;
      for (auto& p : ex.log) {
        pairs.push_back(net_id.at(p.first));
        pairs.push_back(net_id.at(p.second));
      }
      check(spn_net_set_schedule(net, pairs.data(), (uint32_t)pairs.size() / 2));
    }
    check(spn_net_compile(net, exec_mode));
    Slots rs(spn_net_result_order(net));
    check(spn_net_result_slots(net, rs.data()));
    const uint64_t n = inputs.empty() ? 1 : inputs[0].n_points();
    Tensor out = Tensor::batched_empty(ctx, rs, n);
    std::vector<const spn_tensor*> in;
    for (auto& t : inputs) in.push_back(t.raw());
    check(spn_net_execute(net, in.data(), (uint32_t)in.size(), out.raw(), nullptr, 0, 0));
    // same canonical structure on both sides
    Slots bs = r.t.structure();
    EXPECT(bs.size() == rs.size());
    for (size_t i = 0; i < bs.size() && i < rs.size(); ++i) EXPECT(Executor::same_slot(bs[i], rs[i]));
    return {r.t.to_host(), out.to_host()};
  }
};

static Slot bis(uint64_t a) { return Bispinor.new_slot(4, a); }
static Slot mink(uint64_t a) { return Minkowski.new_slot(4, a); }

// e+e- -> mu+mu- spinor chains (BASELINE configs[1]): Product only, no scalar factor => bit-identical
static void eemumu_bit_identical(const Context& ctx, int exec_mode) {
  const uint64_t n = 3000;   // above the specialised-kernel threshold: the kernels the bench runs
  std::mt19937_64 g(7);
  Twin tw(ctx);
  std::vector<int> f;
  for (uint64_t a = 1; a <= 4; ++a) f.push_back(tw.batched({bis(a)}, Tensor::batched(ctx, {bis(a)}, n, random_batch(g, n, 4))));
  int g1 = tw.library(SPN_LIB_GAMMA_WEYL, {bis(1), bis(2), mink(1)});
  int g2 = tw.library(SPN_LIB_GAMMA_WEYL, {bis(3), bis(4), mink(1)});
  int root = tw.op(SPN_OP_PRODUCT, {f[0], g1, f[1], f[2], g2, f[3]});
  auto res = tw.run(root, exec_mode, true);
  EXPECT(tw.ex.log.size() == 5);
  EXPECT(res.first.size() == n && res.second.size() == n);
  if (exec_mode == SPN_EXEC_STEPWISE)
    EXPECT(std::memcmp(res.first.data(), res.second.data(), n * sizeof(Complex)) == 0);
  EXPECT(max_diff(res.first, res.second) <= 1e-12 * max_abs(res.first));
}

// Sum, Neg, conj, Power, scalars and nested products in one expression:
//   ( 2.5 * [ A(1,2) gamma(2,3,mu) B(3) ] + ( -conj( C(1,mu) ) ) ) * (1.5 - 0.5i)^2 * ( D(5) D(5) )^... (power of a tensor)
static void mixed_ops(const Context& ctx, int exec_mode) {
  const uint64_t n = 2500;
  std::mt19937_64 g(11);
  Twin tw(ctx);
  int A = tw.batched({bis(1), bis(2)}, Tensor::batched(ctx, {bis(1), bis(2)}, n, random_batch(g, n, 16)));
  int B = tw.batched({bis(3)}, Tensor::batched(ctx, {bis(3)}, n, random_batch(g, n, 4)));
  int C = tw.batched({bis(1), mink(9)}, Tensor::batched(ctx, {bis(1), mink(9)}, n, random_batch(g, n, 16)));
  int D = tw.batched({mink(5)}, Tensor::batched(ctx, {mink(5)}, n, random_batch(g, n, 4)));
  int gam = tw.library(SPN_LIB_GAMMA_WEYL, {bis(2), bis(3), mink(9)});
  int chain = tw.op(SPN_OP_PRODUCT, {tw.scalar(Complex(2.5, 0)), A, gam, B});
  int negc = tw.op(SPN_OP_NEG, {tw.op(SPN_OP_CONJ, {C})});
  int sum = tw.op(SPN_OP_SUM, {chain, negc});
  int s2 = tw.op(SPN_OP_POWER, {tw.scalar(Complex(1.5, -0.5))}, 2);
  int d2 = tw.op(SPN_OP_POWER, {D}, 2);   // D.D with the Minkowski metric: a Scalar
  int root = tw.op(SPN_OP_PRODUCT, {sum, s2, d2});
  auto res = tw.run(root, exec_mode, false);
  EXPECT(res.first.size() == n * 16 && res.second.size() == n * 16);
  EXPECT(max_abs(res.first) > 0.1);
  EXPECT(max_diff(res.first, res.second) <= 1e-12 * max_abs(res.first));
}

// 4-gamma trace (BASELINE configs[0], spenso-hep-lib/tests/gamma_algebra_validate.rs:29-70) against the closed form
static void gamma_trace(const Context& ctx) {
  Twin tw(ctx);
  std::vector<Complex> p = {0, 1, 2, 3}, q = {1, 2, 3, 4}, r = {1, 3, 5, 7};
  int p1 = tw.batched({mink(1)}, Tensor::batched(ctx, {mink(1)}, 1, p));
  int r3 = tw.batched({mink(3)}, Tensor::batched(ctx, {mink(3)}, 1, r));
  int root = tw.op(SPN_OP_PRODUCT, {p1, r3, tw.library(SPN_LIB_GAMMA_WEYL, {bis(1), bis(2), mink(1)}),
                                    tw.library(SPN_LIB_GAMMA_WEYL, {bis(2), bis(3), mink(2)}),
                                    tw.library(SPN_LIB_GAMMA_WEYL, {bis(3), bis(4), mink(3)}),
                                    tw.library(SPN_LIB_GAMMA_WEYL, {bis(4), bis(1), mink(4)})});
  auto res = tw.run(root, SPN_EXEC_STEPWISE, true);
  const double want[16] = {136, 4, 8, 12, 4, -112, 44, 64, 8, 44, -56, 116, 12, 64, 116, 32};
  EXPECT(res.first.size() == 16);
  for (int i = 0; i < 16 && res.first.size() == 16; ++i) {
    EXPECT(res.first[(size_t)i] == Complex(want[i], 0));
    EXPECT(res.second[(size_t)i] == Complex(want[i], 0));
  }
  (void)q;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--compile-check") return 0;
  try {
    Context ctx(0);
    struct Case {
      const char* name;
      std::function<void()> fn;
    } cases[] = {
        {"eemumu: blanket executor == stepwise engine with the host's schedule (bit-identical)", [&] { eemumu_bit_identical(ctx, SPN_EXEC_STEPWISE); }},
        {"eemumu: blanket executor vs fused kernel with the host's schedule (1e-12)", [&] { eemumu_bit_identical(ctx, SPN_EXEC_FUSED); }},
        {"sum / neg / conj / power / scalars: blanket executor vs stepwise engine", [&] { mixed_ops(ctx, SPN_EXEC_STEPWISE); }},
        {"sum / neg / conj / power / scalars: blanket executor vs fused kernel", [&] { mixed_ops(ctx, SPN_EXEC_FUSED); }},
        {"4-gamma trace closed form through both executors", [&] { gamma_trace(ctx); }},
    };
    for (auto& c : cases) {
      const int before = failures;
      c.fn();
      std::printf("%s ... %s\n", c.name, failures == before ? "ok" : "FAILED");
    }
  } catch (const ContractionError& e) {
    std::printf("ContractionError(%d): %s\n", (int)e.kind, e.what());
    return 2;
  }
  std::printf("%d failure(s)\n", failures);
  return failures ? 1 : 0;
}
