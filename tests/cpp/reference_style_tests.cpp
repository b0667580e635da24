// The reference's own contraction tests, re-stated against the C++ mirror of its interface
// (include/spenso_b200.hpp).  Sources of the known answers:
//   dense_dense_single_contract   spenso/src/tests.rs:416-437 + snapshots/spenso__tests__dense_dense_single_contract.snap
//   sparse_diag_dense_contract    spenso/src/tests.rs:439-473 + snapshots/spenso__tests__sparse_diag_dense_contract.snap
//   simple_multi_contract         spenso/src/tests.rs:634-686 + snapshots/spenso__tests__simple_multi_contract.snap
//   sparse_dense / sparse_sparse  spenso/src/tests.rs:1039-1071, 1164-1195
//   dot (Minkowski sign)          spenso/src/network/library/symbolic.rs:998-1050
//   gamma trace                   spenso-hep-lib/tests/gamma_algebra_validate.rs:29-70 (closed form)
// The legacy tests used unsorted structures, so results are compared by index label.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>

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

// value of a (single point) tensor at the index labelled by abstract index -> position
static Complex at(const Tensor& t, const std::map<uint64_t, uint32_t>& idx, uint64_t point = 0) {
  Slots s = t.structure();
  std::vector<uint64_t> st(s.size(), 1);
  for (size_t i = s.size(); i-- > 1;) st[i - 1] = st[i] * s[i].dim;
  uint64_t f = 0;
  for (size_t i = 0; i < s.size(); ++i) f += idx.at(s[i].aind) * st[i];
  return t.to_host()[point * t.size() + f];
}
static std::vector<Complex> iota(int n) {
  std::vector<Complex> v;
  for (int i = 1; i <= n; ++i) v.push_back(Complex(i, 0));
  return v;
}

static void dense_dense_single_contract(const Context& ctx) {
  Tensor a = DenseTensor::dense(ctx, {Euclidean.new_slot(4, 1), Euclidean.new_slot(4, 2)}, iota(16));
  Tensor b = DenseTensor::dense(ctx, {Euclidean.new_slot(4, 2), Euclidean.new_slot(3, 3)}, iota(12));
  Tensor f = a.contract(b);
  const double want[12] = {70, 80, 90, 158, 184, 210, 246, 288, 330, 334, 392, 450};
  for (uint32_t i = 0; i < 4; ++i)
    for (uint32_t j = 0; j < 3; ++j) EXPECT(at(f, {{1, i}, {3, j}}) == Complex(want[3 * i + j], 0));
  EXPECT(!f.is_sparse());
}

static void sparse_diag_dense_contract(const Context& ctx) {
  Slots sa = {Euclidean.new_slot(4, 1), Euclidean.new_slot(4, 2)}, sb = {Euclidean.new_slot(4, 2), Euclidean.new_slot(3, 3)};
  std::map<std::vector<uint32_t>, Complex> diag;
  for (uint32_t i = 0; i < 4; ++i) diag[{i, i}] = Complex(i + 1, 0);
  Tensor a = SparseTensor::sparse(ctx, sa, diag);
  Tensor b = DenseTensor::dense(ctx, sb, iota(12));
  std::map<std::vector<uint32_t>, Complex> bs;
  for (uint32_t i = 0; i < 4; ++i)
    for (uint32_t j = 0; j < 3; ++j) bs[{i, j}] = Complex(3 * i + j + 1, 0);
  Tensor bsp = SparseTensor::sparse(ctx, sb, bs);
  Tensor ad = DenseTensor::dense(ctx, sa, {1, 0, 0, 0, 0, 2, 0, 0, 0, 0, 3, 0, 0, 0, 0, 4});
  const double want[12] = {1, 2, 3, 8, 10, 12, 21, 24, 27, 40, 44, 48};
  // f == g == h == i across the four storage combinations
  Tensor f = a.contract(b), g = ad.contract(b), h = a.contract(bsp), i2 = ad.contract(bsp);
  for (const Tensor* t : {&f, &g, &h, &i2})
    for (uint32_t i = 0; i < 4; ++i)
      for (uint32_t j = 0; j < 3; ++j) EXPECT(at(*t, {{1, i}, {3, j}}) == Complex(want[3 * i + j], 0));
  EXPECT(h.is_sparse() && !f.is_sparse() && !g.is_sparse() && !i2.is_sparse());  // Dense unless both Sparse
}

static void simple_multi_contract(const Context& ctx) {
  // A(3,4,4) with (euc3|1, euc4|2, lor4|3), B(4,3,4) with (euc4|2, euc3|4, lor_dual4|3): contracts slots 2 and 3
  Slots sa = {Euclidean.new_slot(3, 1), Euclidean.new_slot(4, 2), Lorentz.new_slot(4, 3)};
  Slots sb = {Euclidean.new_slot(4, 2), Euclidean.new_slot(3, 4), Lorentz.dual().new_slot(4, 3)};
  // data of the reference test (tests.rs:648-664); snapshot: [36, 120, 89] for every row
  const int blockA[16] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 11, 12, 13, 14, 15, 16};
  const int blockB[16] = {3, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 0, 0, 3, 2, 1};
  std::vector<Complex> A(48), B(48);
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 16; ++k) {
      A[16 * i + k] = blockA[k];
      B[16 * i + k] = blockB[k];
    }
  Tensor a = DenseTensor::dense(ctx, sa, A), b = DenseTensor::dense(ctx, sb, B);
  Tensor ab = a.contract(b), ba = b.contract(a);
  // einsum('ikm,kjm->ij') by hand
  for (uint32_t i = 0; i < 3; ++i)
    for (uint32_t j = 0; j < 3; ++j) {
      Complex want = 0;
      for (int k = 0; k < 4; ++k)
        for (int m = 0; m < 4; ++m) want += A[16 * i + 4 * k + m] * B[12 * k + 4 * j + m];
      const double snap[3] = {36, 120, 89};
      EXPECT(want == Complex(snap[j], 0));
      EXPECT(at(ab, {{1, i}, {4, j}}) == want);
      EXPECT(at(ba, {{1, i}, {4, j}}) == want);  // A.B == B.A
    }
}

static void sparse_dense_and_sparse_sparse(const Context& ctx) {
  Slots s12 = {Euclidean.new_slot(2, 1), Euclidean.new_slot(2, 2)}, s23 = {Euclidean.new_slot(2, 2), Euclidean.new_slot(2, 3)};
  Tensor a = SparseTensor::sparse(ctx, s12, {{{0, 0}, 1.0}, {{1, 1}, 2.0}});
  Tensor b = DenseTensor::dense(ctx, s23, {1, 2, 3, 4});
  Tensor f = a.contract(b);
  const double want[4] = {1, 2, 6, 8};
  for (uint32_t i = 0; i < 2; ++i)
    for (uint32_t j = 0; j < 2; ++j) EXPECT(at(f, {{1, i}, {3, j}}) == Complex(want[2 * i + j], 0));
  Tensor bs = SparseTensor::sparse(ctx, s23, {{{1, 0}, 1.0}, {{0, 1}, 2.0}});
  Tensor g = a.contract(bs);
  auto el = g.elements();
  EXPECT(g.is_sparse() && el.size() == 2 && el.at(1) == Complex(2, 0) && el.at(2) == Complex(2, 0));
}

static void dot_minkowski(const Context& ctx) {
  Tensor p = DenseTensor::dense(ctx, {Minkowski.new_slot(4, 2)}, {1, 2, 3, 4});
  Tensor q = DenseTensor::dense(ctx, {Minkowski.new_slot(4, 2)}, {5, 6, 7, 8});
  EXPECT(p.contract(q).to_host()[0] == Complex(5 - 12 - 21 - 32, 0));
}

static void errors(const Context& ctx) {
  Tensor es = SparseTensor::sparse(ctx, {Euclidean.new_slot(3, 1)}, {});
  Tensor ed = BatchTensor::batched(ctx, {Euclidean.new_slot(3, 1)}, 0, {});
  bool thrown = false;
  try {
    es.contract(ed);
  } catch (const ContractionError& e) {
    thrown = e.kind == ContractionError::EmptySparse;
  }
  EXPECT(thrown);
  thrown = false;
  try {
    DenseTensor::dense(ctx, {Euclidean.new_slot(3, 1)}, {1, 2});
  } catch (const ContractionError& e) {
    thrown = e.kind == ContractionError::DimensionError;
  }
  EXPECT(thrown);
}

static void gamma_trace_network(const Context& ctx) {
  // p(1) (p+q)(3) gamma(1,2,1) gamma(2,3,2) gamma(3,4,3) gamma(4,1,4), p=(0,1,2,3), q=(1,2,3,4), for 3 points
  auto b = [](uint64_t a) { return Bispinor.new_slot(4, a); };
  auto m = [](uint64_t a) { return Minkowski.new_slot(4, a); };
  std::vector<Complex> P, Q;
  for (int pt = 0; pt < 3; ++pt)
    for (int i = 0; i < 4; ++i) {
      P.push_back(Complex(i, 0));
      Q.push_back(Complex(i + 1, 0));
    }
  Tensor p = BatchTensor::batched(ctx, {m(1)}, 3, P), q = BatchTensor::batched(ctx, {m(3)}, 3, Q);
  for (int mode : {SPN_EXEC_FUSED, SPN_EXEC_STEPWISE}) {
    Network net(ctx);
    auto p1 = net.from_tensor(p), q3 = net.from_tensor(q);
    auto p3 = net.reindex(p1, {m(3)});
    auto expr = p1 * (p3 + q3) * net.library(SPN_LIB_GAMMA_WEYL, {b(1), b(2), m(1)}) *
                net.library(SPN_LIB_GAMMA_WEYL, {b(2), b(3), m(2)}) * net.library(SPN_LIB_GAMMA_WEYL, {b(3), b(4), m(3)}) *
                net.library(SPN_LIB_GAMMA_WEYL, {b(4), b(1), m(4)});
    Tensor r = net.execute(expr, mode);
    const double want[4][4] = {{136, 4, 8, 12}, {4, -112, 44, 64}, {8, 44, -56, 116}, {12, 64, 116, 32}};
    for (uint64_t pt = 0; pt < 3; ++pt)
      for (uint32_t i = 0; i < 4; ++i)
        for (uint32_t j = 0; j < 4; ++j) EXPECT(at(r, {{2, i}, {4, j}}, pt) == Complex(want[i][j], 0));
  }
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--compile-check") return 0;  // the CPU suite only builds this file
  try {
    Context ctx(0);
    const std::pair<const char*, std::function<void(const Context&)>> tests[] = {
        {"dense_dense_single_contract", dense_dense_single_contract},
        {"sparse_diag_dense_contract", sparse_diag_dense_contract},
        {"simple_multi_contract", simple_multi_contract},
        {"sparse_dense_and_sparse_sparse", sparse_dense_and_sparse_sparse},
        {"dot_minkowski", dot_minkowski},
        {"errors", errors},
        {"gamma_trace_network", gamma_trace_network}};
    for (auto& t : tests) {
      int before = failures;
      t.second(ctx);
      std::printf("test %s ... %s\n", t.first, failures == before ? "ok" : "FAILED");
    }
    std::printf("launches: %llu\n", (unsigned long long)ctx.launch_count());
  } catch (const ContractionError& e) {
    std::printf("ContractionError(%d): %s\n", (int)e.kind, e.what());
    return 2;
  }
  return failures ? 1 : 0;
}
