// Point-independent library tensors (the "small shared tensors" of the engine): Dirac matrices,
// chiral projectors and metrics.  Values are those of spenso-hep-lib (Weyl / chiral basis unless
// noted): /root/reference/spenso-hep-lib/src/lib.rs:27-63 (Dirac), :66-102 (Weyl gamma),
// :167-185 (gamma0), :205-221 (gamma5), :251-265 / :308-322 (P-, P+);
// metric / identity: /root/reference/spenso/src/network/library/symbolic.rs:576-659.
// Argument order is (bis i, bis j[, mink mu]) — the rep-canonical order of the library key
// (/root/reference/idenso/src/gamma.rs:477-488).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <vector>

#include "structure.hpp"

namespace spn {

struct LibEntry {
  uint8_t idx[3];
  double re, im;
};

// gamma^mu in the chiral basis: gamma^0 = [[0,1],[1,0]], gamma^k = [[0,sigma_k],[-sigma_k,0]]
static const LibEntry kGammaWeyl[16] = {
    {{0, 2, 0}, 1, 0},  {{1, 3, 0}, 1, 0},  {{2, 0, 0}, 1, 0},  {{3, 1, 0}, 1, 0},
    {{0, 3, 1}, 1, 0},  {{1, 2, 1}, 1, 0},  {{2, 1, 1}, -1, 0}, {{3, 0, 1}, -1, 0},
    {{0, 3, 2}, 0, -1}, {{1, 2, 2}, 0, 1},  {{2, 1, 2}, 0, 1},  {{3, 0, 2}, 0, -1},
    {{0, 2, 3}, 1, 0},  {{1, 3, 3}, -1, 0}, {{2, 0, 3}, -1, 0}, {{3, 1, 3}, 1, 0}};
// NB the reference's "dirac" table is indexed (mu, i, j) on a (bis, bis, mink)-shaped structure
// (lib.rs:27-63 writes set(&[mu,i,j])); we keep its literal positions.
static const LibEntry kGammaDirac[16] = {
    {{0, 0, 0}, 1, 0},  {{0, 1, 1}, 1, 0},  {{0, 2, 2}, -1, 0}, {{0, 3, 3}, -1, 0},
    {{1, 0, 3}, 1, 0},  {{1, 1, 2}, 1, 0},  {{1, 2, 1}, -1, 0}, {{1, 3, 0}, -1, 0},
    {{2, 0, 3}, 0, -1}, {{2, 1, 2}, 0, 1},  {{2, 2, 1}, 0, 1},  {{2, 3, 0}, 0, -1},
    {{3, 0, 2}, 1, 0},  {{3, 1, 3}, -1, 0}, {{3, 2, 0}, -1, 0}, {{3, 3, 1}, 1, 0}};
static const LibEntry kGamma0Weyl[4] = {{{0, 2, 0}, 1, 0}, {{1, 3, 0}, 1, 0}, {{2, 0, 0}, 1, 0}, {{3, 1, 0}, 1, 0}};
static const LibEntry kGamma5Weyl[4] = {{{0, 0, 0}, -1, 0}, {{1, 1, 0}, -1, 0}, {{2, 2, 0}, 1, 0}, {{3, 3, 0}, 1, 0}};
static const LibEntry kProjMWeyl[2] = {{{0, 0, 0}, 1, 0}, {{1, 1, 0}, 1, 0}};
static const LibEntry kProjPWeyl[2] = {{{2, 2, 0}, 1, 0}, {{3, 3, 0}, 1, 0}};

// Library tensor in ARGUMENT order: sparse entries keyed by the argument-order flat index.
struct LibraryData {
  std::vector<std::pair<uint64_t, double2>> nz;
};

inline LibraryData builtin_library(int which, const std::vector<spn_slot>& args) {
  const LibEntry* tab = nullptr;
  int n = 0, rank = 0;
  switch (which) {
    case SPN_LIB_GAMMA_WEYL: tab = kGammaWeyl; n = 16; rank = 3; break;
    case SPN_LIB_GAMMA_DIRAC: tab = kGammaDirac; n = 16; rank = 3; break;
    case SPN_LIB_GAMMA0_WEYL: tab = kGamma0Weyl; n = 4; rank = 2; break;
    case SPN_LIB_GAMMA5_WEYL: tab = kGamma5Weyl; n = 4; rank = 2; break;
    case SPN_LIB_PROJM_WEYL: tab = kProjMWeyl; n = 2; rank = 2; break;
    case SPN_LIB_PROJP_WEYL: tab = kProjPWeyl; n = 2; rank = 2; break;
    case SPN_LIB_METRIC: rank = 2; break;
    // gamma_conj_data_weyl (lib.rs:144-153): entry-wise complex conjugate of gamma;
    // gamma_adj_data_weyl (lib.rs:155-166): conjugate of gamma_transpose_weyl (lib.rs:103-141) = gamma^dagger
    case SPN_LIB_GAMMA_CONJ_WEYL:
    case SPN_LIB_GAMMA_ADJ_WEYL: tab = kGammaWeyl; n = 16; rank = 3; break;
    default: throw Error(SPN_ERR_INVALID, "unknown library tensor");
  }
  if ((int)args.size() != rank) throw Error(SPN_ERR_STRUCTURE, "library tensor: wrong number of indices");
  std::vector<uint64_t> st(rank, 1);
  for (int i = rank - 1; i > 0; --i) st[i - 1] = st[i] * args[i].dim;
  LibraryData d;
  if (which == SPN_LIB_METRIC) {
    // g(i,j): diag(+,-,-,...) for inline-metric representations, the identity otherwise
    if (args[0].dim != args[1].dim) throw Error(SPN_ERR_DIMENSION, "metric: dimensions differ");
    for (uint32_t i = 0; i < args[0].dim; ++i)
      d.nz.push_back({i * st[0] + i * st[1], make_double2((has_metric(args[0]) && i > 0) ? -1.0 : 1.0, 0.0)});
    return d;
  }
  for (int r = 0; r < rank; ++r)
    if (args[r].dim != 4) throw Error(SPN_ERR_DIMENSION, "library tensor: expected dimension 4");
  for (int e = 0; e < n; ++e) {
    uint8_t idx[3] = {tab[e].idx[0], tab[e].idx[1], tab[e].idx[2]};
    double im = tab[e].im;
    if (which == SPN_LIB_GAMMA_ADJ_WEYL) std::swap(idx[0], idx[1]);  // transpose the bispinor indices ...
    if (which == SPN_LIB_GAMMA_CONJ_WEYL || which == SPN_LIB_GAMMA_ADJ_WEYL) im = -im;  // ... and conjugate
    uint64_t f = 0;
    for (int r = 0; r < rank; ++r) f += idx[r] * st[r];
    d.nz.push_back({f, make_double2(tab[e].re, im)});
  }
  return d;
}

// Give a library tensor concrete indices and transpose it into canonical slot order — the one-off
// replacement for the reference's per-use clone + permute (network/graph.rs:291-322,
// tensors/data/sparse.rs:96-112).  Index bookkeeping only.
inline std::map<uint64_t, double2> instantiate_library(const LibraryData& d, const std::vector<spn_slot>& args,
                                                       Structure* st_out) {
  std::vector<int32_t> perm;
  Structure st = canonicalize(args, &perm);
  const size_t n = args.size();
  std::vector<uint64_t> arg_stride(n, 1);
  for (size_t i = n; i-- > 1;) arg_stride[i - 1] = arg_stride[i] * args[i].dim;
  auto cst = st.strides();
  std::vector<uint64_t> canon_stride_of_arg(n);
  for (size_t c = 0; c < n; ++c) canon_stride_of_arg[perm[c]] = cst[c];
  std::map<uint64_t, double2> out;
  for (auto& kv : d.nz) {
    uint64_t f = kv.first, o = 0;
    for (size_t a = 0; a < n; ++a) {
      o += (f / arg_stride[a]) * canon_stride_of_arg[a];
      f %= arg_stride[a];
    }
    out[o] = kv.second;
  }
  *st_out = st;
  return out;
}

}  // namespace spn
