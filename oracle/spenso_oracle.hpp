// ============================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement (C++17, no dependencies) of the concrete-data contraction
// path of alphal00p/spenso, written from the reference's Rust sources, each
// function citing the reference file:line it follows (paths relative to
// /root/reference/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; the product path
// (spenso_b200/) never links, imports or executes it.
//
// Parity pinning: the reference is Rust and cannot be compiled in this image
// (no cargo/rustc; linnet/symbolica/bitvec un-vendored), so this restatement
// is pinned against every reproducible known-answer vector the reference's
// own tests hold (tests/test_oracle_kats.py).  Two pieces stay
// "parity unpinned" because they live in the absent linnet 0.17.0 crate:
// (i) the greedy network contraction *order* (node/hedge iteration order) and
// (ii) Permutation conventions for >=3 rotated dualizable contracted indices.
// Values are order-independent up to rounding (1e-12 norm-wise).
//
// Build with -ffp-contract=off so that a*b-c*d is NOT fused (Rust does not
// contract floating point expressions).
// ============================================================================
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// Scalars  (spenso/src/algebra/complex.rs:55-58, complex/mul.rs:7-22,
//           upgrading_arithmetic.rs:1267-1397: f64 is upgraded to (x, 0.0)
//           and then multiplied with the full complex product)
// ---------------------------------------------------------------------------
struct C {
  double re, im;
};
inline C cmul(C a, C b) {  // complex/mul.rs:16-21
  return C{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
inline void cadd(C& x, C y) { x.re += y.re; x.im += y.im; }
inline void csub(C& x, C y) { x.re -= y.re; x.im -= y.im; }
inline C cneg(C a) { return C{-a.re, -a.im}; }
inline bool czero(C a) { return a.re == 0.0 && a.im == 0.0; }  // complex.rs:229-231

// ---------------------------------------------------------------------------
// R1: canonical total order of slots
// ---------------------------------------------------------------------------
enum RepKind : uint8_t { REP_DUMMY = 0, REP_SELF_DUAL = 1, REP_INLINE_METRIC = 2, REP_DUALIZABLE = 3 };
enum AindKind : uint8_t { AIND_NORMAL = 0, AIND_DUALIZE = 1, AIND_DOUBLE = 2, AIND_DUMMY = 3, AIND_ADDED = 4 };

// Binary layout shared with the product's spn_slot (include/spenso_b200.h).
struct Slot {
  uint8_t rep_kind;
  uint8_t aind_kind;
  int16_t rep_id;  // SelfDual(n)/InlineMetric(n): n ; Dualizable(k): +k base, -k dual
  uint32_t dim;
  uint64_t aind;  // payload; Double(a,b) packed as (a<<16)|b
};
static_assert(sizeof(Slot) == 16, "slot layout");

template <class T>
inline int cmp3(T a, T b) { return a < b ? -1 : (a > b ? 1 : 0); }

// LibraryRep::cmp — structure/representation.rs:591-620
inline int rep_cmp(const Slot& a, const Slot& b) {
  auto ka = a.rep_kind, kb = b.rep_kind;
  if (ka == REP_SELF_DUAL && kb == REP_SELF_DUAL) return cmp3<int>(a.rep_id, b.rep_id);
  if (ka == REP_INLINE_METRIC && kb == REP_INLINE_METRIC) return cmp3<int>(a.rep_id, b.rep_id);
  if (ka == REP_DUALIZABLE && kb == REP_DUALIZABLE) {
    int x = a.rep_id, y = b.rep_id;
    if (x < 0) {
      if (y < 0) return cmp3(std::abs(x), std::abs(y));
      return 1;
    } else if (y < 0) {
      return -1;
    }
    return cmp3(x, y);
  }
  if ((ka == REP_SELF_DUAL && kb == REP_DUALIZABLE) || (ka == REP_SELF_DUAL && kb == REP_INLINE_METRIC) ||
      (ka == REP_INLINE_METRIC && kb == REP_DUALIZABLE))
    return -1;
  if ((ka == REP_DUALIZABLE && kb == REP_SELF_DUAL) || (ka == REP_INLINE_METRIC && kb == REP_SELF_DUAL) ||
      (ka == REP_DUALIZABLE && kb == REP_INLINE_METRIC))
    return 1;
  if (ka == REP_DUMMY && kb == REP_DUMMY) return 0;
  if (ka == REP_DUMMY) return -1;
  return 1;
}
// RepName for LibraryRep — structure/representation.rs:1077-1204
inline bool rep_is_base(const Slot& s) { return s.rep_kind == REP_DUALIZABLE ? s.rep_id > 0 : true; }
inline bool rep_is_dual(const Slot& s) { return s.rep_kind == REP_DUALIZABLE ? s.rep_id < 0 : true; }
inline bool rep_is_self_dual(const Slot& s) { return s.rep_kind != REP_DUALIZABLE; }  // :1119-1121
inline bool rep_matches(const Slot& a, const Slot& b) {                              // :1131-1139
  if (a.rep_kind != b.rep_kind) return false;
  switch (a.rep_kind) {
    case REP_SELF_DUAL: return a.rep_id == b.rep_id;
    case REP_DUALIZABLE: return a.rep_id == -b.rep_id;
    case REP_INLINE_METRIC: return a.rep_id == b.rep_id;
    default: return false;
  }
}
inline int rep_match_cmp(const Slot& a, const Slot& b) {  // :1141-1148
  if ((a.rep_kind == REP_SELF_DUAL && b.rep_kind == REP_SELF_DUAL) ||
      (a.rep_kind == REP_INLINE_METRIC && b.rep_kind == REP_INLINE_METRIC))
    return cmp3<int>(a.rep_id, b.rep_id);
  if (a.rep_kind == REP_DUALIZABLE && b.rep_kind == REP_DUALIZABLE)
    return cmp3(std::abs((int)a.rep_id), std::abs((int)b.rep_id));
  return rep_cmp(a, b);
}
// is_neg — representation.rs:1198-1204 with the only registered inline metric,
// Minkowski: is_neg(i) = i > 0 (representation.rs:321-323, 1010-1035)
inline bool rep_is_neg(const Slot& s, size_t i) { return s.rep_kind == REP_INLINE_METRIC ? i > 0 : false; }

// Representation::cmp — representation.rs:466-474 (rep THEN dim)
inline int representation_cmp(const Slot& a, const Slot& b) {
  if (a.rep_kind == REP_DUMMY && b.rep_kind == REP_DUMMY) return 0;
  int c = rep_cmp(a, b);
  return c ? c : cmp3(a.dim, b.dim);
}
// AbstractIndex derive(Ord) — abstract_index.rs:267-294
inline int aind_cmp(const Slot& a, const Slot& b) {
  int c = cmp3<int>(a.aind_kind, b.aind_kind);
  return c ? c : cmp3(a.aind, b.aind);
}
// Slot derive(Ord): fields rep, aind — slot.rs:25-56
inline int slot_cmp(const Slot& a, const Slot& b) {
  int c = representation_cmp(a, b);
  return c ? c : aind_cmp(a, b);
}
inline bool slot_eq(const Slot& a, const Slot& b) {
  return a.rep_kind == b.rep_kind && a.rep_id == b.rep_id && a.dim == b.dim && a.aind_kind == b.aind_kind &&
         a.aind == b.aind;
}
inline Slot slot_dual(Slot s) {  // slot.rs:323-329, representation.rs:1092-1100
  if (s.rep_kind == REP_DUALIZABLE) s.rep_id = (int16_t)(-s.rep_id);
  return s;
}
// DualSlotTo::matches / match_cmp — slot.rs:330-343; Representation::matches/match_cmp :1294-1303
inline bool slot_matches(const Slot& a, const Slot& b) {
  return a.dim == b.dim && rep_matches(a, b) && aind_cmp(a, b) == 0;
}
inline int slot_match_cmp(const Slot& a, const Slot& b) {
  int c = cmp3(a.dim, b.dim);
  if (c) return c;
  c = rep_match_cmp(a, b);
  return c ? c : aind_cmp(a, b);
}

struct SlotHash {
  size_t operator()(const Slot& s) const {
    uint64_t h = (uint64_t)s.rep_kind | ((uint64_t)s.aind_kind << 8) | ((uint64_t)(uint16_t)s.rep_id << 16) |
                 ((uint64_t)s.dim << 32);
    h ^= s.aind * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    return (size_t)(h * 0xBF58476D1CE4E5B9ull);
  }
};
struct SlotEq {
  bool operator()(const Slot& a, const Slot& b) const { return slot_eq(a, b); }
};

// ---------------------------------------------------------------------------
// R2: OrderedStructure — structure/ordered.rs:39-44, 275-299, 321-380
// ---------------------------------------------------------------------------
constexpr size_t NONE = (size_t)-1;

enum MergeKind : int { FIRST_BEFORE_SECOND = 0, SECOND_BEFORE_FIRST = 1, INTERLEAVED = 2 };

struct Structure {
  std::vector<Slot> s;
  size_t dual_start = NONE;
  size_t base_start = NONE;

  size_t order() const { return s.size(); }
  bool is_scalar() const { return s.empty(); }
  size_t n_self_dual() const { return base_start >= s.size() ? s.size() : base_start; }  // :275-281
  size_t n_base() const {                                                                // :283-291
    if (base_start >= s.size()) return 0;
    if (dual_start >= s.size()) return s.size() - base_start;
    return dual_start - base_start;
  }
  size_t n_dual() const { return dual_start >= s.size() ? 0 : s.size() - dual_start; }  // :293-299
  bool is_fully_self_dual() const { return base_start >= s.size(); }                     // :168-170

  // structure.rs:566-577 (row major: last index fastest), :694-703
  std::vector<size_t> strides() const {
    std::vector<size_t> st(s.size(), 1);
    if (s.empty()) return st;
    for (size_t i = s.size() - 1; i-- > 0;) st[i] = st[i + 1] * (size_t)s[i + 1].dim;
    return st;
  }
  size_t size() const {
    size_t n = 1;
    for (auto& x : s) n *= x.dim;
    return n;
  }
  std::vector<size_t> shape() const {
    std::vector<size_t> d;
    for (auto& x : s) d.push_back(x.dim);
    return d;
  }
  size_t flat_index(const std::vector<size_t>& idx) const {  // :623-632
    auto st = strides();
    if (idx.size() != s.size()) throw std::runtime_error("Mismatched order");
    size_t f = 0;
    for (size_t i = 0; i < idx.size(); ++i) {
      if (idx[i] >= s[i].dim) throw std::runtime_error("Index out of bounds");
      f += idx[i] * st[i];
    }
    return f;
  }
  std::vector<size_t> expanded_index(size_t flat) const {  // :639-651
    std::vector<size_t> out;
    size_t index = flat;
    for (size_t stride : strides()) {
      out.push_back(index / stride);
      index %= stride;
    }
    if (flat >= size()) throw std::runtime_error("flat index out of bounds");
    return out;
  }
};

// PermutedStructure::from(Vec<Slot>) — ordered.rs:321-380.
// perm_out[i] = position in the *input* list of the slot that ends up at
// canonical position i (stable sorts, so equal slots keep their input order).
inline Structure structure_from_slots(std::vector<Slot> slots, std::vector<int>* perm_out = nullptr) {
  std::vector<int> idx(slots.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
  // rep_permutation = sort_by_key(|a| a.rep)
  std::stable_sort(idx.begin(), idx.end(),
                   [&](int a, int b) { return representation_cmp(slots[a], slots[b]) < 0; });
  std::vector<Slot> st;
  for (int i : idx) st.push_back(slots[i]);
  Structure out;
  if (st.empty()) {
    if (perm_out) perm_out->clear();
    return out;
  }
  size_t base_start = (rep_is_base(st[0]) && !rep_is_self_dual(st[0])) ? 0 : NONE;
  size_t dual_start = NONE;
  if (rep_is_dual(st[0]) && !rep_is_self_dual(st[0])) {
    base_start = 0;
    dual_start = 0;
  }
  for (size_t i = 0; i + 1 < st.size(); ++i) {
    if (rep_is_self_dual(st[i]) && !rep_is_self_dual(st[i + 1])) base_start = i + 1;
    if (rep_is_base(st[i]) && !rep_is_self_dual(st[i + 1]) && rep_is_dual(st[i + 1])) dual_start = i + 1;
  }
  // index_permutation = Permutation::sort(&structure)
  std::vector<int> idx2(st.size());
  for (size_t i = 0; i < idx2.size(); ++i) idx2[i] = (int)i;
  std::stable_sort(idx2.begin(), idx2.end(), [&](int a, int b) { return slot_cmp(st[a], st[b]) < 0; });
  for (int i : idx2) out.s.push_back(st[i]);
  out.base_start = base_start;
  out.dual_start = dual_start;
  if (perm_out) {
    perm_out->clear();
    for (int i : idx2) perm_out->push_back(idx[i]);
  }
  return out;
}

// ---------------------------------------------------------------------------
// R5: StructureContract::merge — ordered.rs:594-919 ; MergeInfo — structure.rs:716-751
// ---------------------------------------------------------------------------
struct MergeResult {
  Structure st;
  std::vector<uint8_t> common_self, common_other, partition;
  int kind = FIRST_BEFORE_SECOND;
};

inline int merge_info_from(const std::vector<uint8_t>& bits) {  // structure.rs:725-751
  if (bits.empty()) return FIRST_BEFORE_SECOND;
  int n_transitions = 0;
  for (size_t i = 0; i + 1 < bits.size(); ++i) {
    if (bits[i] != bits[i + 1]) n_transitions++;
    if (n_transitions > 1) return INTERLEAVED;
  }
  return bits[0] ? FIRST_BEFORE_SECOND : SECOND_BEFORE_FIRST;
}

inline MergeResult merge(const Structure& self, const Structure& other) {
  MergeResult r;
  r.common_self.assign(self.order(), 0);
  r.common_other.assign(other.order(), 0);
  if (self.is_scalar() || other.is_scalar()) {  // :637-654
    if (self.is_scalar()) {
      r.st = other;
      r.kind = SECOND_BEFORE_FIRST;
    } else {
      r.st = self;
      r.kind = FIRST_BEFORE_SECOND;
    }
    return r;
  }
  auto& partition = r.partition;
  std::vector<Slot> res;
  size_t i = 0, j = 0;
  // (i) self-dual block :666-703
  while (i < self.n_self_dual() && j < other.n_self_dual()) {
    int c = slot_cmp(self.s[i], other.s[j]);
    if (c < 0) {
      res.push_back(self.s[i]);
      partition.push_back(1);
      i++;
    } else if (c > 0) {
      res.push_back(other.s[j]);
      partition.push_back(0);
      j++;
    } else {
      if (slot_matches(self.s[i], other.s[j])) {
        r.common_self[i] = 1;
        r.common_other[j] = 1;
        i++;
        j++;
      } else {
        throw std::runtime_error("Matching and equal items must be identical");
      }
    }
  }
  while (i < self.n_self_dual()) {
    res.push_back(self.s[i]);
    partition.push_back(1);
    i++;
  }
  while (j < other.n_self_dual()) {
    res.push_back(other.s[j]);
    partition.push_back(0);
    j++;
  }
  size_t base_start = partition.size();
  // (ii) base block :709-836
  size_t ibase = 0, idual = 0, jbase = 0, jdual = 0;
  size_t snbase = self.n_base(), onbase = other.n_base();
  auto find_match_in_duals = [](const Slot& base_slot, size_t base_slot_index, std::vector<uint8_t>& common_base,
                                const Structure& dual_structure, size_t& dual_cursor,
                                std::vector<uint8_t>& common_dual, size_t dual_offset) -> bool {
    bool found = false;
    size_t dual_count = dual_structure.n_dual();
    while (dual_cursor < dual_count) {
      size_t dual_slot_index = dual_offset + dual_cursor;
      const Slot& dual_slot = dual_structure.s[dual_slot_index];
      int c = slot_match_cmp(base_slot, dual_slot);
      if (c < 0) {
        break;
      } else if (c > 0) {
        dual_cursor++;
      } else {
        if (slot_matches(base_slot, dual_slot)) {
          common_base[base_slot_index] = 1;
          common_dual[dual_slot_index] = 1;
          found = true;
          dual_cursor++;
          break;
        } else {
          throw std::runtime_error("Matching and equal items must be identical");
        }
      }
    }
    return found;
  };
  while (ibase < snbase && jbase < onbase) {
    int c = slot_cmp(self.s[i + ibase], other.s[j + jbase]);
    if (c < 0) {
      const Slot& b = self.s[i + ibase];
      if (!find_match_in_duals(b, i + ibase, r.common_self, other, jdual, r.common_other, j + onbase)) {
        res.push_back(b);
        partition.push_back(1);
      }
      ibase++;
    } else if (c > 0) {
      const Slot& b = other.s[j + jbase];
      if (!find_match_in_duals(b, j + jbase, r.common_other, self, idual, r.common_self, i + snbase)) {
        res.push_back(b);
        partition.push_back(0);
      }
      jbase++;
    } else {
      throw std::runtime_error("Cannot have equal bases");
    }
  }
  while (ibase < snbase) {
    const Slot& b = self.s[i + ibase];
    if (!find_match_in_duals(b, i + ibase, r.common_self, other, jdual, r.common_other, j + onbase)) {
      res.push_back(b);
      partition.push_back(1);
    }
    ibase++;
  }
  while (jbase < onbase) {
    const Slot& b = other.s[j + jbase];
    if (!find_match_in_duals(b, j + jbase, r.common_other, self, idual, r.common_self, i + snbase)) {
      res.push_back(b);
      partition.push_back(0);
    }
    jbase++;
  }
  size_t dual_start = partition.size();
  // (iii) dual block :840-879
  idual = 0;
  jdual = 0;
  while (idual < self.n_dual() && jdual < other.n_dual()) {
    int c = slot_cmp(self.s[i + ibase + idual], other.s[j + jbase + jdual]);
    if (c < 0) {
      if (!r.common_self[i + ibase + idual]) {
        res.push_back(self.s[i + ibase + idual]);
        partition.push_back(1);
      }
      idual++;
    } else if (c > 0) {
      if (!r.common_other[j + jbase + jdual]) {
        res.push_back(other.s[j + jbase + jdual]);
        partition.push_back(0);
      }
      jdual++;
    } else {
      throw std::runtime_error("Cannot have equal duals");
    }
  }
  while (idual < self.n_dual()) {
    if (!r.common_self[i + ibase + idual]) {
      res.push_back(self.s[i + ibase + idual]);
      partition.push_back(1);
    }
    idual++;
  }
  while (jdual < other.n_dual()) {
    if (!r.common_other[j + jbase + jdual]) {
      res.push_back(other.s[j + jbase + jdual]);
      partition.push_back(0);
    }
    jdual++;
  }
  // :881-892 result must already be canonically sorted
  for (size_t k = 0; k + 1 < res.size(); ++k)
    if (slot_cmp(res[k], res[k + 1]) > 0) throw std::runtime_error("Permutation is not identity for merged structure");
  r.st.s = res;
  r.st.base_start = base_start;
  r.st.dual_start = dual_start;
  r.kind = merge_info_from(partition);
  return r;
}

// R6: TensorStructure::match_indices — structure.rs:435-460.
// Returns false when nothing matches. perm[k] = position in self of the slot
// dual to the k-th matched slot of other (i.e. self positions in OTHER's order).
inline bool match_indices(const Structure& self, const Structure& other, std::vector<size_t>& perm,
                          std::vector<uint8_t>& self_matches, std::vector<uint8_t>& other_matches) {
  self_matches.assign(self.order(), 0);
  other_matches.assign(other.order(), 0);
  perm.clear();
  std::unordered_map<Slot, size_t, SlotHash, SlotEq> posmap;
  for (size_t i = 0; i < self.order(); ++i) posmap[self.s[i]] = i;  // later duplicates overwrite (collect into map)
  for (size_t j = 0; j < other.order(); ++j) {
    auto it = posmap.find(slot_dual(other.s[j]));
    if (it != posmap.end()) {
      self_matches[it->second] = 1;
      other_matches[j] = 1;
      perm.push_back(it->second);
    }
  }
  return !perm.empty();
}

// ---------------------------------------------------------------------------
// I1: CoreFlatFiberIterator — iterators/core_iterators.rs:132-143, 149-201,
//     206-240, 243-272, 323-414.  Restated literally (increment + stride/shift
//     wrap rules + zero_index base).
// ---------------------------------------------------------------------------
struct FlatIt {
  size_t varying = 0, increment = 1, max = 0, zero_index = 0;
  bool multi = true;
  std::vector<size_t> strides, shifts;  // Multi
  bool has_single = false;              // Single(Some)
  size_t single_stride = 0, single_shift = 0;

  bool next(size_t& out) {  // :245-271
    if (varying > max) return false;
    out = varying + zero_index;
    varying += increment;
    if (multi) {
      for (size_t i = 0; i < strides.size(); ++i) {
        if (varying % strides[i] == 0)
          varying += shifts[i] - strides[i];
        else
          break;
      }
    } else if (has_single) {
      if (varying % single_stride == 0) varying += single_shift - single_stride;
    }
    return true;
  }
  void reset() { varying = 0; }
  void shift(size_t s) { zero_index = s; }
};

// init_multi_fiber_iter :149-201.  free_[pos] says whether index pos is free in
// the (abstract) fiber being iterated.
inline FlatIt flat_it_multi(const std::vector<size_t>& strides, const std::vector<size_t>& dims,
                            const std::vector<uint8_t>& free_, bool conj) {
  size_t order = dims.size();
  size_t max = 0, increment = 1;
  std::vector<size_t> fixed_strides, shifts;
  bool before = true, has_seen_stride = false, first = true;
  for (size_t pos = order; pos-- > 0;) {
    bool is_free = free_[pos] != 0, is_fixed = !is_free;
    if ((is_fixed ^ conj) && !before && !first) {
      has_seen_stride = true;
      fixed_strides.push_back(strides[pos]);
    }
    if ((is_free ^ conj) && before && has_seen_stride) shifts.push_back(strides[pos]);
    if (is_free ^ conj) {
      size_t dimminus1 = dims[pos] == 0 ? 0 : dims[pos] - 1;
      max += dimminus1 * strides[pos];
      if (first) {
        increment = strides[pos];
        first = false;
      }
    }
    before = is_fixed ^ conj;
  }
  if (fixed_strides.size() > shifts.size()) fixed_strides.pop_back();
  FlatIt it;
  it.multi = true;
  it.increment = increment;
  it.strides = fixed_strides;
  it.shifts = shifts;
  it.max = max;
  return it;
}

// init_single_fiber_iter :206-240
inline FlatIt flat_it_single(const std::vector<size_t>& strides, size_t fiber_position, const std::vector<size_t>& dims,
                             bool conj) {
  size_t fiber_stride = strides[fiber_position];
  size_t dim = dims[fiber_position];
  size_t size = 1;
  for (auto d : dims) size *= d;
  FlatIt it;
  it.multi = false;
  if (conj) {
    it.max = size - fiber_stride * (dim - 1) - 1;
    it.increment = 1;
    if (fiber_position == dims.size() - 1) {
      it.increment = dims.size() >= 2 ? strides[dims.size() - 2] : 1;
    } else if (fiber_position != 0) {
      it.has_single = true;
      it.single_shift = strides[fiber_position - 1];
      it.single_stride = strides[fiber_position];
    }
  } else {
    it.increment = fiber_stride;
    it.max = fiber_stride * (dim - 1);
  }
  return it;
}

// new_paired_conjugates :323-414.  Returns (first = iterator over the FIXED
// positions, second = iterator over the FREE positions) of the abstract fiber.
inline std::pair<FlatIt, FlatIt> flat_it_paired_conjugates(const std::vector<size_t>& strides,
                                                           const std::vector<size_t>& dims,
                                                           const std::vector<uint8_t>& free_) {
  size_t order = dims.size();
  size_t max = 0, increment = 1;
  std::vector<size_t> fixed_strides, fixed_strides_conj, shifts, shifts_conj;
  bool before = true, has_seen_stride = false, has_seen_stride_conj = false, first = true, first_conj = true;
  size_t increment_conj = 1, max_conj = 0;
  for (size_t pos = order; pos-- > 0;) {
    bool is_free = free_[pos] != 0, is_fixed = !is_free;
    if (is_fixed && !before) {
      if (!first) {
        has_seen_stride = true;
        fixed_strides.push_back(strides[pos]);
      }
      if (has_seen_stride_conj) shifts_conj.push_back(strides[pos]);
    }
    if (is_free && before) {
      if (has_seen_stride) shifts.push_back(strides[pos]);
      if (!first_conj) {
        fixed_strides_conj.push_back(strides[pos]);
        has_seen_stride_conj = true;
      }
    }
    if (is_fixed) {
      max_conj += (dims[pos] - 1) * strides[pos];
      if (first_conj) {
        increment_conj = strides[pos];
        first_conj = false;
      }
    } else {
      max += (dims[pos] - 1) * strides[pos];
      if (first) {
        increment = strides[pos];
        first = false;
      }
    }
    before = is_fixed;
  }
  if (fixed_strides.size() > shifts.size()) fixed_strides.pop_back();
  if (fixed_strides_conj.size() > shifts_conj.size()) fixed_strides_conj.pop_back();
  FlatIt a, b;
  a.multi = true;
  a.strides = fixed_strides_conj;
  a.shifts = shifts_conj;
  a.increment = increment_conj;
  a.max = max_conj;
  b.multi = true;
  b.strides = fixed_strides;
  b.shifts = shifts;
  b.increment = increment;
  b.max = max;
  return {a, b};
}

// ---------------------------------------------------------------------------
// I2: CoreExpandedFiberIterator / MetricFiberIterator — core_iterators.rs:422-513, 565-669
// ---------------------------------------------------------------------------
struct MetricIt {
  std::vector<size_t> pos, dims, strides;
  std::vector<Slot> reps;
  size_t zero_index = 0, flat = 0;
  bool exhausted = false;

  // init_iter :440-464 : keep dims where (conj ^ free); optional reorder `order`
  // (order[k] = which kept dim sits at odometer position k; last is fastest).
  static MetricIt make(const Structure& st, const std::vector<uint8_t>& free_, bool conj,
                       const std::vector<size_t>* order = nullptr) {
    MetricIt it;
    auto strides = st.strides();
    std::vector<size_t> d, s;
    std::vector<Slot> r;
    for (size_t i = 0; i < st.order(); ++i)
      if (conj ^ (free_[i] != 0)) {
        d.push_back(st.s[i].dim);
        s.push_back(strides[i]);
        r.push_back(st.s[i]);
      }
    if (order) {
      for (size_t k : *order) {
        it.dims.push_back(d[k]);
        it.strides.push_back(s[k]);
        it.reps.push_back(r[k]);
      }
    } else {
      it.dims = d;
      it.strides = s;
      it.reps = r;
    }
    it.pos.assign(it.dims.size(), 0);
    return it;
  }
  // MetricFiberIterator::next :630-668 (sign = XOR of is_neg over the CURRENT position, before increment)
  bool next(size_t& out, bool& neg) {
    if (exhausted) return false;
    out = flat + zero_index;
    bool carry = true;
    neg = false;
    for (size_t k = pos.size(); k-- > 0;) {
      neg ^= rep_is_neg(reps[k], pos[k]);
      if (carry) {
        pos[k] += 1;
        if (pos[k] >= dims[k]) {
          pos[k] = 0;
          flat -= strides[k] * (dims[k] - 1);
        } else {
          flat += strides[k];
          carry = false;
        }
      }
    }
    if (carry) exhausted = true;
    return true;
  }
  void reset() {
    flat = 0;
    exhausted = false;
    std::fill(pos.begin(), pos.end(), 0);
  }
  void shift(size_t s) { zero_index = s; }
};

// ---------------------------------------------------------------------------
// S1: storage — tensors/data/dense.rs:51-54, sparse.rs:48-53, data.rs:190-193
// ---------------------------------------------------------------------------
struct Tensor {
  Structure st;
  bool sparse = false;
  std::vector<C> data;                    // dense, row-major
  std::unordered_map<size_t, C> elems;    // sparse: flat index -> value (HashMap in the reference)

  size_t size() const { return st.size(); }
  const C* get_ref_linear(size_t flat) const {
    if (sparse) {
      auto it = elems.find(flat);
      return it == elems.end() ? nullptr : &it->second;
    }
    return flat < data.size() ? &data[flat] : nullptr;
  }
  bool has_entries() const { return sparse ? !elems.empty() : !data.empty(); }
  std::vector<C> to_dense_data() const {  // sparse.rs:698-708
    if (!sparse) return data;
    std::vector<C> d(size(), C{0, 0});
    for (auto& kv : elems) d[kv.first] = kv.second;
    return d;
  }
  static Tensor dense(Structure st, std::vector<C> d) {
    Tensor t;
    if (d.size() != st.size()) throw std::runtime_error("data length does not match structure size");
    t.st = std::move(st);
    t.data = std::move(d);
    return t;
  }
  static Tensor sparse_empty(Structure st) {
    Tensor t;
    t.st = std::move(st);
    t.sparse = true;
    return t;
  }
};

enum ContractError : int { OK = 0, ERR_EMPTY_SPARSE = 1, ERR_STRUCTURE = 2, ERR_DIMENSION = 3, ERR_OTHER = 4 };
struct ContractionError : std::runtime_error {
  int code;
  ContractionError(int c, const char* m) : std::runtime_error(m), code(c) {}
};

// A fiber iterator over a tensor (fiber_iterators.rs:105-169): dense yields
// values, sparse yields (value, skipped) and silently skips absent entries.
template <class It>
struct TIt {
  const Tensor* t;
  It it;
  size_t skipped = 0;
  void reset() {
    it.reset();
    skipped = 0;
  }
  void shift(size_t s) {  // FiberIterator::shift :80-83 resets first
    reset();
    it.shift(s);
  }
};
inline bool core_next(FlatIt& it, size_t& flat, bool& neg) {
  neg = false;
  return it.next(flat);
}
inline bool core_next(MetricIt& it, size_t& flat, bool& neg) { return it.next(flat, neg); }

template <class It>
inline bool dense_next(TIt<It>& f, C& v, bool& neg) {
  size_t flat;
  if (!core_next(f.it, flat, neg)) return false;
  const C* p = f.t->get_ref_linear(flat);
  if (!p) throw std::runtime_error("DenseTensor: Out of bounds");
  v = *p;
  return true;
}
template <class It>
inline bool dense_nth(TIt<It>& f, size_t n, C& v, bool& neg) {  // Iterator::nth
  for (size_t k = 0; k < n; ++k)
    if (!dense_next(f, v, neg)) return false;
  return dense_next(f, v, neg);
}
template <class It>
inline bool sparse_next(TIt<It>& f, C& v, size_t& skip, bool& neg) {  // :156-168
  size_t flat;
  while (core_next(f.it, flat, neg)) {
    const C* p = f.t->get_ref_linear(flat);
    if (p) {
      skip = f.skipped;
      f.skipped = 0;
      v = *p;
      return true;
    }
    f.skipped += 1;
  }
  return false;
}

inline std::vector<uint8_t> single_free(size_t order, size_t i) {
  std::vector<uint8_t> f(order, 0);
  f[i] = 1;
  return f;
}
inline std::vector<uint8_t> invert(const std::vector<uint8_t>& a) {
  std::vector<uint8_t> b(a.size());
  for (size_t i = 0; i < a.size(); ++i) b[i] = !a[i];
  return b;
}

// FiberClassIterator::new (fiber_iterators.rs:309-329): for a tensor and the
// set of "fiber" positions (free in the bare fiber), the class iterator walks
// the complementary positions; the fiber iterator walks the fiber positions.
struct ClassIters {
  FlatIt fiber, cls;
};
inline ClassIters class_iters(const Structure& st, const std::vector<uint8_t>& bare_free) {
  // FiberClass inverts free/fixed (fiber.rs:536-549): class-free = bare-fixed.
  auto class_free = invert(bare_free);
  auto pr = flat_it_paired_conjugates(st.strides(), st.shape(), class_free);
  return ClassIters{pr.first /* over class-fixed = bare-free */, pr.second /* over class-free */};
}

inline C accumulate_dd(TIt<FlatIt>& fa, TIt<FlatIt>& fb, const Slot& rep) {
  C acc{0, 0};
  C a, b;
  bool n1, n2;
  size_t k = 0;
  while (dense_next(fa, a, n1) && dense_next(fb, b, n2)) {  // zip
    if (rep_is_neg(rep, k))
      csub(acc, cmul(a, b));
    else
      cadd(acc, cmul(a, b));
    k++;
  }
  return acc;
}

// ---------------------------------------------------------------------------
// K1-K3: SingleContract — contraction/singlecontract.rs:25-285
// ---------------------------------------------------------------------------
inline Tensor single_contract(const Tensor& self, const Tensor& other, const Structure& final_structure, size_t i,
                              size_t j) {
  auto ci_a = class_iters(self.st, single_free(self.st.order(), i));
  auto ci_b = class_iters(other.st, single_free(other.st.order(), j));
  const Slot fiber_representation = self.st.s[i];
  FlatIt self_class = ci_a.cls, other_class = ci_b.cls;
  TIt<FlatIt> fa{&self, ci_a.fiber}, fb{&other, ci_b.fiber};
  size_t result_index = 0;
  size_t sa, sb;

  // the dense-receiver kernels read self.data[0] for the zero element (:43, :96) and panic when empty
  if (!self.sparse && self.data.empty()) throw ContractionError(ERR_OTHER, "panic: index out of bounds (empty dense)");
  if (!self.sparse && !other.sparse) {  // :25-76
    Tensor out = Tensor::dense(final_structure, std::vector<C>(final_structure.size(), C{0, 0}));
    while (self_class.next(sa)) {
      fa.shift(sa);
      while (other_class.next(sb)) {
        fb.shift(sb);
        out.data[result_index] = accumulate_dd(fa, fb, fiber_representation);
        result_index++;
        fa.reset();
      }
      other_class.reset();
    }
    return out;
  }
  if (!self.sparse && other.sparse) {  // :78-132
    Tensor out = Tensor::dense(final_structure, std::vector<C>(final_structure.size(), C{0, 0}));
    while (self_class.next(sa)) {
      fa.shift(sa);
      while (other_class.next(sb)) {
        fb.shift(sb);
        C acc{0, 0}, a, b;
        size_t skip, k = 0;
        bool n1, n2;
        while (sparse_next(fb, b, skip, n1)) {
          if (dense_nth(fa, skip, a, n2)) {
            // reference uses is_neg(k + skip) (:110); see SURVEY K-note 1
            if (rep_is_neg(fiber_representation, k + skip))
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
          }
          k++;
        }
        out.data[result_index] = acc;
        result_index++;
        fa.reset();
      }
      other_class.reset();
    }
    return out;
  }
  if (self.sparse && !other.sparse) {  // :134-196
    if (!self.has_entries() && !other.has_entries()) throw ContractionError(ERR_EMPTY_SPARSE, "Sparse tensor is empty");
    Tensor out = Tensor::dense(final_structure, std::vector<C>(final_structure.size(), C{0, 0}));
    while (self_class.next(sa)) {
      fa.shift(sa);
      while (other_class.next(sb)) {
        fb.shift(sb);
        C acc{0, 0}, a, b;
        size_t skip, k = 0;
        bool n1, n2;
        while (sparse_next(fa, a, skip, n1)) {
          if (dense_nth(fb, skip, b, n2)) {
            if (rep_is_neg(fiber_representation, k + skip))
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
          }
          k++;
        }
        out.data[result_index] = acc;
        result_index++;
        fa.reset();
      }
      other_class.reset();
    }
    return out;
  }
  // sparse x sparse :198-285
  Tensor out = Tensor::sparse_empty(final_structure);
  if (self.has_entries()) {
    while (self_class.next(sa)) {
      fa.shift(sa);
      while (other_class.next(sb)) {
        fb.shift(sb);
        C a, b, value{0, 0};
        size_t skip_a, skip_b, sk;
        bool n1, n2, nonzero = false;
        bool has_a = sparse_next(fa, a, skip_a, n1);
        bool has_b = sparse_next(fb, b, skip_b, n2);
        while (has_a && has_b) {
          if (skip_a > skip_b) {
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
          } else if (skip_b > skip_a) {
            has_a = sparse_next(fa, a, sk, n1);
            if (has_a) skip_a = sk + skip_a + 1;
          } else {
            if (rep_is_neg(fiber_representation, skip_a))  // metric[skip_a] :251
              csub(value, cmul(a, b));
            else
              cadd(value, cmul(a, b));
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
            has_a = sparse_next(fa, a, sk, n1);
            if (has_a) skip_a = sk + skip_a + 1;
            nonzero = true;
          }
        }
        if (nonzero && !czero(value)) out.elems[result_index] = value;
        result_index++;
        fa.reset();
      }
      other_class.reset();
    }
  }
  return out;
}

// Split an expanded result index by partition (true = self) — singlecontract.rs:324-339
inline void split_by_partition(const std::vector<size_t>& exp, const std::vector<uint8_t>& partition,
                               std::vector<size_t>& expa, std::vector<size_t>& expb) {
  expa.clear();
  expb.clear();
  for (size_t k = 0; k < exp.size(); ++k) (partition[k] ? expa : expb).push_back(exp[k]);
}

// ---------------------------------------------------------------------------
// K4: SingleContractInterleaved — singlecontract.rs:287-655
// ---------------------------------------------------------------------------
inline Tensor single_contract_interleaved(const Tensor& self, const Tensor& other, const Structure& rs,
                                          const std::vector<uint8_t>& partition, size_t i, size_t j) {
  if (self.sparse && !other.sparse && !self.has_entries() && !other.has_entries())
    throw ContractionError(ERR_EMPTY_SPARSE, "Sparse tensor is empty");
  bool out_sparse = self.sparse && other.sparse;
  Tensor out = out_sparse ? Tensor::sparse_empty(rs) : Tensor::dense(rs, std::vector<C>(rs.size(), C{0, 0}));
  if (out_sparse && !self.has_entries()) return out;
  // Fiber::from(&partition,&rs): free where partition is true; new_paired_conjugates(&fiber)
  auto pr = flat_it_paired_conjugates(rs.strides(), rs.shape(), partition);
  FlatIt class_a = pr.first, class_b = pr.second;
  // self.fiber(FiberData::from(i)).iter(): Single(i) leaves is_single unset => multi init (fiber.rs:37-41)
  TIt<FlatIt> fa{&self, flat_it_multi(self.st.strides(), self.st.shape(), single_free(self.st.order(), i), false)};
  TIt<FlatIt> fb{&other, flat_it_multi(other.st.strides(), other.st.shape(), single_free(other.st.order(), j), false)};
  const Slot rep = self.st.s[i];
  size_t ida, idb;
  std::vector<size_t> expa, expb;
  while (class_a.next(ida)) {
    while (class_b.next(idb)) {
      size_t result_index = ida + idb;
      split_by_partition(rs.expanded_index(result_index), partition, expa, expb);
      expa.insert(expa.begin() + i, 0);
      expb.insert(expb.begin() + j, 0);
      fa.shift(self.st.flat_index(expa));
      fb.shift(other.st.flat_index(expb));
      C a, b;
      bool n1, n2;
      size_t skip, k = 0;
      if (!self.sparse && !other.sparse) {  // :345-355
        C acc = accumulate_dd(fa, fb, rep);
        cadd(out.data[result_index], acc);
      } else if (!self.sparse && other.sparse) {  // :426-436
        while (sparse_next(fb, b, skip, n1)) {
          if (dense_nth(fa, skip, a, n2)) {
            if (rep_is_neg(rep, k + skip))
              csub(out.data[result_index], cmul(a, b));
            else
              cadd(out.data[result_index], cmul(a, b));
          }
          k++;
        }
      } else if (self.sparse && !other.sparse) {  // :516-526
        while (sparse_next(fa, a, skip, n1)) {
          if (dense_nth(fb, skip, b, n2)) {
            if (rep_is_neg(rep, k + skip))
              csub(out.data[result_index], cmul(a, b));
            else
              cadd(out.data[result_index], cmul(a, b));
          }
          k++;
        }
      } else {  // :597-642
        C value{0, 0};
        size_t skip_a, skip_b, sk;
        bool nonzero = false;
        bool has_a = sparse_next(fa, a, skip_a, n1);
        bool has_b = sparse_next(fb, b, skip_b, n2);
        while (has_a && has_b) {
          if (skip_a > skip_b) {
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
          } else if (skip_b > skip_a) {
            has_a = sparse_next(fa, a, sk, n1);
            if (has_a) skip_a = sk + skip_a + 1;
          } else {
            if (rep_is_neg(rep, skip_a))
              csub(value, cmul(a, b));
            else
              cadd(value, cmul(a, b));
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
            has_a = sparse_next(fa, a, sk, n1);
            if (has_a) skip_a = sk + skip_a + 1;
            nonzero = true;
          }
        }
        if (nonzero && !czero(value)) out.elems[result_index] = value;
      }
    }
    class_b.reset();
  }
  return out;
}

// Order in which self's contracted dims must be walked so that they pair with
// other's contracted dims in other's (flat) order — the intent of
// match_indices + Permutation::sort + apply_slice_in_place
// (structure.rs:435-460, core_iterators.rs:451-454); see SURVEY K-note 2.
inline std::vector<size_t> perm_to_other_order(const std::vector<size_t>& perm, const std::vector<uint8_t>& self_matches) {
  std::vector<size_t> rank(self_matches.size(), 0);  // position in self -> index among self's matched dims
  size_t r = 0;
  for (size_t p = 0; p < self_matches.size(); ++p)
    if (self_matches[p]) rank[p] = r++;
  std::vector<size_t> order;
  for (size_t p : perm) order.push_back(rank[p]);
  return order;
}

// ---------------------------------------------------------------------------
// K5: MultiContract — contraction/multicontract.rs:25-293
// ---------------------------------------------------------------------------
inline Tensor multi_contract(const Tensor& self, const Tensor& other, const Structure& final_structure) {
  if (self.sparse && !other.sparse && !self.has_entries() && !other.has_entries())
    throw ContractionError(ERR_EMPTY_SPARSE, "Sparse tensor is empty");
  std::vector<size_t> perm;
  std::vector<uint8_t> self_matches, other_matches;
  if (!match_indices(self.st, other.st, perm, self_matches, other_matches))
    throw ContractionError(ERR_OTHER, "match_indices returned None");
  bool out_sparse = self.sparse && other.sparse;
  Tensor out = out_sparse ? Tensor::sparse_empty(final_structure)
                          : Tensor::dense(final_structure, std::vector<C>(final_structure.size(), C{0, 0}));
  if (out_sparse && !self.has_entries()) return out;
  // selfiter = fiber_class(self_matches).iter_perm_metric(permutation)  (fiber_iterators.rs:351-360):
  //   class_iter = CoreFlatFiberIterator::new(&class,false) over the unmatched dims,
  //   fiber iter = MetricFiberIterator over matched dims, permuted.
  auto order = perm_to_other_order(perm, self_matches);
  FlatIt self_class = flat_it_multi(self.st.strides(), self.st.shape(), invert(self_matches), false);
  TIt<MetricIt> fa{&self, MetricIt::make(self.st, self_matches, false, &order)};
  auto ci_b = class_iters(other.st, other_matches);
  FlatIt other_class = ci_b.cls;
  TIt<FlatIt> fb{&other, ci_b.fiber};
  size_t result_index = 0, sa, sb;
  while (self_class.next(sa)) {
    fa.shift(sa);
    while (other_class.next(sb)) {
      fb.shift(sb);
      C a, b, acc{0, 0};
      bool neg, n2;
      size_t skip;
      if (!self.sparse && !other.sparse) {  // :53-70
        while (dense_next(fa, a, neg)) {
          if (dense_next(fb, b, n2)) {
            if (neg)
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
          }
        }
        out.data[result_index] = acc;
      } else if (self.sparse && !other.sparse) {  // :118-135
        while (sparse_next(fa, a, skip, neg)) {
          if (dense_nth(fb, skip, b, n2)) {
            if (neg)
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
          }
        }
        out.data[result_index] = acc;
      } else if (!self.sparse && other.sparse) {  // :174-191
        while (sparse_next(fb, b, skip, n2)) {
          if (dense_nth(fa, skip, a, neg)) {
            if (neg)
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
          }
        }
        out.data[result_index] = acc;
      } else {  // :234-283
        size_t skip_a, skip_b, sk;
        bool nonzero = false, nn;
        bool has_a = sparse_next(fa, a, skip_a, neg);
        bool has_b = sparse_next(fb, b, skip_b, n2);
        while (has_a && has_b) {
          if (skip_a > skip_b) {
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
          } else if (skip_b > skip_a) {
            has_a = sparse_next(fa, a, sk, nn);
            if (has_a) {
              skip_a = sk + skip_a + 1;
              neg = nn;
            }
          } else {
            if (neg)
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
            has_a = sparse_next(fa, a, sk, nn);
            if (has_a) {
              skip_a = sk + skip_a + 1;
              neg = nn;
            }
            nonzero = true;
          }
        }
        if (nonzero && !czero(acc)) out.elems[result_index] = acc;
      }
      result_index++;
      fa.reset();
    }
    other_class.reset();
  }
  return out;
}

// ---------------------------------------------------------------------------
// K6: MultiContractInterleaved — multicontract.rs:295-689.
// emulate_reference_pairing=true reproduces the reference literally: self's
// fibre is walked with iter_metric() in SELF's order and zipped with other's
// fibre in OTHER's order, with no permutation (:321-322, :358; SURVEY K-note 3).
// false pairs each contracted slot of self with its dual in other.
// ---------------------------------------------------------------------------
inline Tensor multi_contract_interleaved(const Tensor& self, const Tensor& other, const std::vector<uint8_t>& pos_self,
                                         const std::vector<uint8_t>& pos_other, const Structure& rs,
                                         const std::vector<uint8_t>& partition, bool emulate_reference_pairing) {
  if (self.sparse && !other.sparse && !self.has_entries() && !other.has_entries())
    throw ContractionError(ERR_EMPTY_SPARSE, "Sparse tensor is empty");
  bool out_sparse = self.sparse && other.sparse;
  Tensor out = out_sparse ? Tensor::sparse_empty(rs) : Tensor::dense(rs, std::vector<C>(rs.size(), C{0, 0}));
  if (out_sparse && !self.has_entries()) return out;
  auto pr = flat_it_paired_conjugates(rs.strides(), rs.shape(), partition);
  FlatIt class_a = pr.first, class_b = pr.second;
  std::vector<size_t> order;
  const std::vector<size_t>* order_ptr = nullptr;
  if (!emulate_reference_pairing) {
    std::vector<size_t> perm;
    std::vector<uint8_t> sm, om;
    match_indices(self.st, other.st, perm, sm, om);
    order = perm_to_other_order(perm, sm);
    order_ptr = &order;
  }
  TIt<MetricIt> fa{&self, MetricIt::make(self.st, pos_self, false, order_ptr)};
  TIt<FlatIt> fb{&other, flat_it_multi(other.st.strides(), other.st.shape(), pos_other, false)};
  std::vector<size_t> exp_self(self.st.order(), 0), exp_other(other.st.order(), 0), expa, expb;
  size_t ida, idb;
  while (class_a.next(ida)) {
    while (class_b.next(idb)) {
      size_t result_index = ida + idb;
      split_by_partition(rs.expanded_index(result_index), partition, expa, expb);
      {
        size_t k = 0;
        for (size_t p = 0; p < pos_self.size() && k < expa.size(); ++p)
          if (!pos_self[p]) exp_self[p] = expa[k++];
        k = 0;
        for (size_t p = 0; p < pos_other.size() && k < expb.size(); ++p)
          if (!pos_other[p]) exp_other[p] = expb[k++];
      }
      fa.shift(self.st.flat_index(exp_self));
      fb.shift(other.st.flat_index(exp_other));
      C a, b;
      bool neg, n2;
      size_t skip;
      if (!self.sparse && !other.sparse) {  // :358-366
        while (dense_next(fa, a, neg) && dense_next(fb, b, n2)) {
          if (neg)
            csub(out.data[result_index], cmul(a, b));
          else
            cadd(out.data[result_index], cmul(a, b));
        }
      } else if (self.sparse && !other.sparse) {
        while (sparse_next(fa, a, skip, neg)) {
          if (dense_nth(fb, skip, b, n2)) {
            if (neg)
              csub(out.data[result_index], cmul(a, b));
            else
              cadd(out.data[result_index], cmul(a, b));
          }
        }
      } else if (!self.sparse && other.sparse) {
        while (sparse_next(fb, b, skip, n2)) {
          if (dense_nth(fa, skip, a, neg)) {
            if (neg)
              csub(out.data[result_index], cmul(a, b));
            else
              cadd(out.data[result_index], cmul(a, b));
          }
        }
      } else {
        C acc{0, 0};
        size_t skip_a, skip_b, sk;
        bool nonzero = false, nn;
        bool has_a = sparse_next(fa, a, skip_a, neg);
        bool has_b = sparse_next(fb, b, skip_b, n2);
        while (has_a && has_b) {
          if (skip_a > skip_b) {
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
          } else if (skip_b > skip_a) {
            has_a = sparse_next(fa, a, sk, nn);
            if (has_a) {
              skip_a = sk + skip_a + 1;
              neg = nn;
            }
          } else {
            if (neg)
              csub(acc, cmul(a, b));
            else
              cadd(acc, cmul(a, b));
            has_b = sparse_next(fb, b, sk, n2);
            if (has_b) skip_b = sk + skip_b + 1;
            has_a = sparse_next(fa, a, sk, nn);
            if (has_a) {
              skip_a = sk + skip_a + 1;
              neg = nn;
            }
            nonzero = true;
          }
        }
        if (nonzero && !czero(acc)) out.elems[result_index] = acc;
      }
    }
    class_b.reset();
  }
  return out;
}

// ---------------------------------------------------------------------------
// K7: ExteriorProduct / ExteriorProductInterleaved — contraction/exteriorproduct.rs:19-392
// ---------------------------------------------------------------------------
template <class F>
inline void for_each_flat(const Tensor& t, F f) {  // flat_iter
  if (t.sparse) {
    std::vector<size_t> keys;
    for (auto& kv : t.elems) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    for (size_t k : keys) f(k, t.elems.at(k));
  } else {
    for (size_t k = 0; k < t.data.size(); ++k) f(k, t.data[k]);
  }
}
inline Tensor exterior_product(const Tensor& self, const Tensor& other, const Structure& final_structure) {
  if (self.sparse && !other.sparse && !self.has_entries() && !other.has_entries())
    throw ContractionError(ERR_EMPTY_SPARSE, "Sparse tensor is empty");
  bool out_sparse = self.sparse && other.sparse;
  Tensor out = out_sparse ? Tensor::sparse_empty(final_structure)
                          : Tensor::dense(final_structure, std::vector<C>(final_structure.size(), C{0, 0}));
  size_t stride = other.size();
  for_each_flat(self, [&](size_t i, C u) {
    for_each_flat(other, [&](size_t j, C t) {
      C v = cmul(u, t);
      if (out_sparse)
        out.elems[i * stride + j] = v;
      else
        out.data[i * stride + j] = v;
    });
  });
  return out;
}
inline Tensor exterior_product_interleaved(const Tensor& self, const Tensor& other, const Structure& rs,
                                           const std::vector<uint8_t>& partition) {
  if (self.sparse && !other.sparse && !self.has_entries() && !other.has_entries())
    throw ContractionError(ERR_EMPTY_SPARSE, "Sparse tensor is empty");
  bool out_sparse = self.sparse && other.sparse;
  Tensor out = out_sparse ? Tensor::sparse_empty(rs) : Tensor::dense(rs, std::vector<C>(rs.size(), C{0, 0}));
  if (out_sparse && !self.has_entries()) return out;
  auto pr = flat_it_paired_conjugates(rs.strides(), rs.shape(), partition);
  FlatIt class_a = pr.first, class_b = pr.second;
  size_t ida, idb;
  std::vector<size_t> expa, expb;
  while (class_a.next(ida)) {
    while (class_b.next(idb)) {
      size_t result_index = ida + idb;
      split_by_partition(rs.expanded_index(result_index), partition, expa, expb);
      const C* u = self.get_ref_linear(self.st.flat_index(expa));
      const C* v = other.get_ref_linear(other.st.flat_index(expb));
      if (out_sparse) {
        if (u && v) out.elems[result_index] = cmul(*u, *v);
      } else if (u && v) {
        out.data[result_index] = cmul(*u, *v);
      }
    }
    class_b.reset();
  }
  return out;
}

// ---------------------------------------------------------------------------
// R7: Contract dispatch — contraction.rs:126-227
// ---------------------------------------------------------------------------
struct OracleOptions {
  bool emulate_reference_interleaved_pairing = false;  // SURVEY K-note 3
};

inline Tensor contract(const Tensor& self, const Tensor& other, const OracleOptions& opt = OracleOptions()) {
  MergeResult m = merge(self.st, other.st);
  size_t common = 0;
  for (auto b : m.common_self) common += b;
  auto first_one = [](const std::vector<uint8_t>& v) {
    for (size_t k = 0; k < v.size(); ++k)
      if (v[k]) return k;
    throw std::runtime_error("Expected a bit to be set");
  };
  if (common == 0) {
    if (m.kind == INTERLEAVED) return exterior_product_interleaved(self, other, m.st, m.partition);
    if (m.kind == FIRST_BEFORE_SECOND) return exterior_product(self, other, m.st);
    return exterior_product(other, self, m.st);
  } else if (common == 1) {
    size_t i = first_one(m.common_self), j = first_one(m.common_other);
    if (m.kind == INTERLEAVED) return single_contract_interleaved(self, other, m.st, m.partition, i, j);
    if (m.kind == FIRST_BEFORE_SECOND) return single_contract(self, other, m.st, i, j);
    return single_contract(other, self, m.st, j, i);
  } else {
    if (m.kind == INTERLEAVED)
      return multi_contract_interleaved(self, other, m.common_self, m.common_other, m.st, m.partition,
                                        opt.emulate_reference_interleaved_pairing);
    if (m.kind == FIRST_BEFORE_SECOND) return multi_contract(self, other, m.st);
    return multi_contract(other, self, m.st);
  }
}

// ---------------------------------------------------------------------------
// Element-wise ops — algebra/add_assign.rs:13-176, scalar_mul.rs:11-57, neg.rs
// ---------------------------------------------------------------------------
inline void add_assign(Tensor& self, const Tensor& rhs, bool subtract = false) {
  if (self.st.order() != rhs.st.order() || self.size() != rhs.size())
    throw ContractionError(ERR_STRUCTURE, "add_assign: structure mismatch");
  if (self.sparse && !rhs.sparse) {  // DataTensor::add_assign: sparse += dense densifies first
    self.data = self.to_dense_data();
    self.elems.clear();
    self.sparse = false;
  }
  auto upd = [&](C& x, C y) { subtract ? csub(x, y) : cadd(x, y); };
  if (!self.sparse && !rhs.sparse) {
    for (size_t k = 0; k < self.data.size(); ++k) upd(self.data[k], rhs.data[k]);
  } else if (!self.sparse && rhs.sparse) {
    for (auto& kv : rhs.elems) upd(self.data[kv.first], kv.second);
  } else {
    for (auto& kv : rhs.elems) {
      auto it = self.elems.find(kv.first);
      if (it == self.elems.end()) it = self.elems.emplace(kv.first, C{0, 0}).first;
      upd(it->second, kv.second);
    }
  }
}
// FallibleAdd / FallibleSub (algebra/fallible_add.rs:17-135, fallible_sub.rs:19-150): the same sums, but the result is
// a fresh tensor — Dense unless BOTH operands are Sparse — and Sparse (+/-) Sparse writes through smart_set
// (tensors/data/sparse.rs:712-724): a value that is exactly zero is not stored (tests.rs:1108-1141 sparse_sub).
// Operands here share one canonical slot order, so the reference's find_permutation is the identity.
inline Tensor add_fallible(const Tensor& self, const Tensor& rhs, bool subtract = false) {
  if (self.st.order() != rhs.st.order() || self.size() != rhs.size())
    throw ContractionError(ERR_STRUCTURE, "add_fallible: structure mismatch");
  if (!(self.sparse && rhs.sparse)) {
    Tensor t = self;
    add_assign(t, rhs, subtract);  // densifies a sparse receiver first, like the Dense outputs of :17-103
    return t;
  }
  Tensor out = Tensor::sparse_empty(self.st);
  auto smart_set = [&](size_t flat, C v) {
    if (v.re == 0.0 && v.im == 0.0)
      out.elems.erase(flat);
    else
      out.elems[flat] = v;
  };
  for (auto& kv : self.elems) {
    auto it = rhs.elems.find(kv.first);
    C v = kv.second;
    if (it != rhs.elems.end()) subtract ? csub(v, it->second) : cadd(v, it->second);
    smart_set(kv.first, v);
  }
  for (auto& kv : rhs.elems)
    if (!self.elems.count(kv.first)) smart_set(kv.first, subtract ? cneg(kv.second) : kv.second);
  return out;
}
inline Tensor neg(const Tensor& t) {
  Tensor o = t;
  for (auto& x : o.data) x = cneg(x);
  for (auto& kv : o.elems) kv.second = cneg(kv.second);
  return o;
}
inline Tensor scalar_mul(const Tensor& t, C s) {
  Tensor o = t;
  for (auto& x : o.data) x = cmul(x, s);
  for (auto& kv : o.elems) kv.second = cmul(kv.second, s);
  return o;
}

// ---------------------------------------------------------------------------
// K8: Trace — contraction/trace.rs:12-83, structure.rs:462-485,
// iterators/tensor_iterators.rs:154-222, 442-503.  One repeated pair at a time
// (the sparse impl's recursion); exact zeros are dropped for sparse (:59).
// ---------------------------------------------------------------------------
inline bool find_trace(const Structure& st, size_t& p0, size_t& p1) {
  // traces(): position b pairs with an earlier position a when slot[b].dual() == slot[a]
  for (size_t a = 0; a < st.order(); ++a)
    for (size_t b = a + 1; b < st.order(); ++b)
      if (slot_eq(slot_dual(st.s[b]), st.s[a])) {
        p0 = a;
        p1 = b;
        return true;
      }
  return false;
}
inline Tensor trace_once(const Tensor& t, size_t p0, size_t p1) {
  Structure ns;
  std::vector<Slot> rest;
  for (size_t k = 0; k < t.st.order(); ++k)
    if (k != p0 && k != p1) rest.push_back(t.st.s[k]);
  ns = structure_from_slots(rest);
  Tensor out = t.sparse ? Tensor::sparse_empty(ns) : Tensor::dense(ns, std::vector<C>(ns.size(), C{0, 0}));
  size_t d = t.st.s[p0].dim;
  std::vector<size_t> idx(t.st.order(), 0);
  size_t nout = ns.size();
  auto rest_strides = ns.strides();
  for (size_t o = 0; o < nout; ++o) {
    auto e = ns.expanded_index(o);
    size_t k2 = 0;
    for (size_t k = 0; k < t.st.order(); ++k)
      if (k != p0 && k != p1) idx[k] = e[k2++];
    C acc{0, 0};
    bool any = false;
    for (size_t i = 0; i < d; ++i) {
      idx[p0] = i;
      idx[p1] = i;
      const C* v = t.get_ref_linear(t.st.flat_index(idx));
      if (!v) continue;
      any = true;
      if (rep_is_neg(t.st.s[p0], i))
        csub(acc, *v);
      else
        cadd(acc, *v);
    }
    if (t.sparse) {
      if (any && !czero(acc)) out.elems[o] = acc;
    } else {
      out.data[o] = acc;
    }
  }
  return out;
}
inline Tensor internal_contract(const Tensor& t) {
  Tensor cur = t;
  size_t p0, p1;
  while (find_trace(cur.st, p0, p1)) cur = trace_once(cur, p0, p1);
  return cur;
}

// ---------------------------------------------------------------------------
// N2: library tensors — network/graph.rs:291-322 (get_lib_data),
// tensors/data/dense.rs:93-109 / sparse.rs:96-112 (permute_inds),
// network/library.rs:342-348 (with_indices): a library tensor is stored in
// rep-canonical order; each use re-indexes it with concrete abstract indices
// and permutes the data into the canonical slot order.  CLONE + PERMUTE per
// use, exactly as the reference does.
// ---------------------------------------------------------------------------
struct LibTensor {
  std::vector<Slot> reps;  // rep kind/id/dim per argument position, ainds ignored
  bool sparse = true;
  std::vector<C> data;                  // dense data in argument order (row major)
  std::vector<std::pair<size_t, C>> nz; // sparse entries keyed by flat index in argument order
};
inline Tensor lib_instantiate(const LibTensor& lt, const std::vector<Slot>& slots_in_arg_order) {
  std::vector<int> perm;
  Structure st = structure_from_slots(slots_in_arg_order, &perm);
  size_t n = slots_in_arg_order.size();
  std::vector<size_t> arg_strides(n, 1);
  for (size_t i = n ? n - 1 : 0; i-- > 0;) arg_strides[i] = arg_strides[i + 1] * slots_in_arg_order[i + 1].dim;
  auto cst = st.strides();
  // canonical position c holds argument perm[c]
  std::vector<size_t> arg_to_canon_stride(n);
  for (size_t c = 0; c < n; ++c) arg_to_canon_stride[perm[c]] = cst[c];
  auto map_flat = [&](size_t f) {
    size_t o = 0;
    for (size_t a = 0; a < n; ++a) {
      size_t ia = f / arg_strides[a];
      f %= arg_strides[a];
      o += ia * arg_to_canon_stride[a];
    }
    return o;
  };
  if (lt.sparse) {
    Tensor t = Tensor::sparse_empty(st);
    for (auto& kv : lt.nz) t.elems[map_flat(kv.first)] = kv.second;
    return t;
  }
  Tensor t = Tensor::dense(st, std::vector<C>(st.size(), C{0, 0}));
  for (size_t f = 0; f < lt.data.size(); ++f) t.data[map_flat(f)] = lt.data[f];
  return t;
}

// ---------------------------------------------------------------------------
// N1: network executor — network.rs:1126-1156 (Sequential), :1275-1705
// (ExecuteOp for NetworkStore), network/contract.rs:43-498 (ContractScalars,
// SmallestDegree, SingleSmallestDegree), network/graph.rs:433-513
// (extract_next_ready_op: preorder DFS, first op whose children are all
// leaves), :1002-1065 (merge_ops).  The expression graph is held as a tree
// (the reference's linnet::HedgeGraph is absent); node and slot iteration
// order are insertion / canonical order  => contraction ORDER is unpinned.
// ---------------------------------------------------------------------------
enum NodeKind : int { N_TENSOR = 0, N_SCALAR = 1, N_LIB = 2, N_PRODUCT = 3, N_SUM = 4, N_NEG = 5, N_POWER = 6, N_CONJ = 7 };

struct Node {
  int kind = N_TENSOR;
  int ref = -1;    // store index (tensor / scalar) or library recipe index
  int power = 1;   // N_POWER
  std::vector<int> children;
};
struct LibUse {
  int lib = -1;                 // index into Network::library
  std::vector<Slot> slots;      // concrete slots in argument order
};
struct Network {
  std::vector<Node> nodes;
  int root = -1;
  std::vector<Tensor> tensors;  // NetworkStore.tensors (append-only, store.rs:125-129)
  std::vector<C> scalars;       // NetworkStore.scalar
  std::vector<LibUse> lib_uses;
  const std::vector<LibTensor>* library = nullptr;
  OracleOptions opt;
  // statistics (for the schedule comparison tests)
  std::vector<std::pair<int, int>> contraction_log;

  int add_node(Node n) {
    nodes.push_back(std::move(n));
    return (int)nodes.size() - 1;
  }
  bool is_leaf(int id) const { return nodes[id].kind <= N_LIB; }

  Tensor get_lib_data(int node) const { return lib_instantiate((*library)[lib_uses[nodes[node].ref].lib], lib_uses[nodes[node].ref].slots); }
  Structure leaf_structure(int node) const {
    const Node& n = nodes[node];
    if (n.kind == N_TENSOR) return tensors[n.ref].st;
    if (n.kind == N_LIB) return structure_from_slots(lib_uses[n.ref].slots);
    return Structure();
  }

  // merge_ops: flatten Product-in-Product and Sum-in-Sum (graph.rs:1002-1065)
  void merge_ops(int id) {
    Node& n = nodes[id];
    if (is_leaf(id)) return;
    for (int c : std::vector<int>(n.children)) merge_ops(c);
    if (nodes[id].kind == N_PRODUCT || nodes[id].kind == N_SUM) {
      std::vector<int> flat;
      for (int c : nodes[id].children) {
        if (nodes[c].kind == nodes[id].kind)
          for (int g : nodes[c].children) flat.push_back(g);
        else
          flat.push_back(c);
      }
      nodes[id].children = flat;
    }
  }
  // preorder DFS for the first op whose children are all leaves; returns parent link too
  bool find_ready(int id, int& found) {
    if (is_leaf(id)) return false;
    bool all_leaves = true;
    for (int c : nodes[id].children) all_leaves &= is_leaf(c);
    if (all_leaves) {
      found = id;
      return true;
    }
    for (int c : nodes[id].children)
      if (find_ready(c, found)) return true;
    return false;
  }
  void replace_with_leaf(int id, int kind, int ref) {
    nodes[id].kind = kind;
    nodes[id].ref = ref;
    nodes[id].children.clear();
  }
  int push_tensor(Tensor t) {
    tensors.push_back(std::move(t));
    return (int)tensors.size() - 1;
  }
  int push_scalar(C s) {
    scalars.push_back(s);
    return (int)scalars.size() - 1;
  }
  Tensor leaf_tensor(int node) const { return nodes[node].kind == N_TENSOR ? tensors[nodes[node].ref] : get_lib_data(node); }

  // ContractScalars::contract — contract.rs:73-208
  void contract_scalars(std::vector<int>& leaves) {
    std::vector<int> scal, tens;
    for (int l : leaves) (nodes[l].kind == N_SCALAR ? scal : tens).push_back(l);
    if (scal.empty()) return;
    bool remove_op_node = tens.size() <= 1;
    int f = scal.back();
    C acc = scalars[nodes[f].ref];
    bool more = false;
    for (size_t k = 0; k + 1 < scal.size(); ++k) {
      more = true;
      acc = cmul(acc, scalars[nodes[scal[k]].ref]);
    }
    if (remove_op_node) {
      if (tens.size() == 1) {
        Tensor a = scalar_mul(leaf_tensor(tens[0]), acc);
        int pos = push_tensor(std::move(a));
        replace_with_leaf(tens[0], N_TENSOR, pos);
        leaves = {tens[0]};
      } else {
        int pos = push_scalar(acc);
        replace_with_leaf(f, N_SCALAR, pos);
        leaves = {f};
      }
    } else {
      if (!more) return;
      int pos = push_scalar(acc);
      replace_with_leaf(f, N_SCALAR, pos);
      leaves = tens;
      leaves.push_back(f);
    }
  }
  // SingleSmallestDegree::contract — contract.rs:346-497
  bool single_smallest_degree(std::vector<int>& leaves) {
    std::vector<int> tens;
    for (int l : leaves)
      if (nodes[l].kind != N_SCALAR) tens.push_back(l);
    std::vector<Structure> sts;
    for (int l : tens) sts.push_back(leaf_structure(l));
    int last_tensor = -1;
    long best_degree = -1;
    int best1 = -1, best2 = -1;
    for (size_t a = 0; a < tens.size(); ++a) {
      long degree = 0;
      int partner = -1;
      for (size_t sidx = 0; sidx < sts[a].order(); ++sidx) {
        Slot d = slot_dual(sts[a].s[sidx]);
        for (size_t b = 0; b < tens.size(); ++b) {
          if (b == a) continue;
          bool hit = false;
          for (auto& os : sts[b].s)
            if (slot_eq(os, d)) hit = true;
          if (hit) {
            degree++;
            partner = (int)b;  // `first` is overwritten => last internal slot hedge (:366-371)
            break;
          }
        }
      }
      int nid2;
      if (degree == 0) {
        degree = 0x7fffffffL;
        if (last_tensor >= 0) {
          nid2 = last_tensor;
        } else {
          last_tensor = (int)a;
          continue;
        }
      } else {
        nid2 = partner;
      }
      last_tensor = (int)a;
      if (best_degree < 0 || degree < best_degree) {  // min_by_key keeps the FIRST minimum
        best_degree = degree;
        best1 = (int)a;
        best2 = nid2;
      }
    }
    if (best1 < 0) return false;
    int n1 = tens[best1], n2 = tens[best2];
    Tensor contracted;
    // (LibraryKey, LocalTensor) contracts local.contract(lib) (:428-443); everything else n1.contract(n2)
    if (nodes[n1].kind == N_LIB && nodes[n2].kind == N_TENSOR)
      contracted = contract(tensors[nodes[n2].ref], get_lib_data(n1), opt);
    else
      contracted = contract(leaf_tensor(n1), leaf_tensor(n2), opt);
    contraction_log.emplace_back(n1, n2);
    int pos = push_tensor(std::move(contracted));
    replace_with_leaf(n1, N_TENSOR, pos);
    leaves.erase(std::find(leaves.begin(), leaves.end(), n2));
    return true;
  }

  static C cpow_inv(C s) {  // ref_one()/s
    double d = s.re * s.re + s.im * s.im;
    return C{s.re / d, -s.im / d};
  }

  void execute_op(int id) {
    Node& n = nodes[id];
    switch (n.kind) {
      case N_PRODUCT: {  // SmallestDegree::contract — contract.rs:239-263
        std::vector<int> leaves = n.children;
        if (leaves.empty()) {
          replace_with_leaf(id, N_SCALAR, push_scalar(C{1, 0}));
          return;
        }
        contract_scalars(leaves);
        while (single_smallest_degree(leaves)) {
        }
        contract_scalars(leaves);
        if (leaves.size() != 1) throw std::runtime_error("product did not reduce to one leaf");
        Node l = nodes[leaves[0]];
        replace_with_leaf(id, l.kind, l.ref);
        return;
      }
      case N_SUM: {  // network.rs:1342-1467
        if (n.children.empty()) {
          replace_with_leaf(id, N_SCALAR, push_scalar(C{0, 0}));
          return;
        }
        int first = n.children[0];
        bool scalar_mode = nodes[first].kind == N_SCALAR ||
                           (nodes[first].kind == N_TENSOR && tensors[nodes[first].ref].st.is_scalar());
        if (scalar_mode) {
          C acc = nodes[first].kind == N_SCALAR ? scalars[nodes[first].ref] : tensors[nodes[first].ref].to_dense_data()[0];
          for (size_t k = 1; k < n.children.size(); ++k) {
            int c = n.children[k];
            if (nodes[c].kind == N_SCALAR)
              cadd(acc, scalars[nodes[c].ref]);
            else if (nodes[c].kind == N_TENSOR && tensors[nodes[c].ref].st.is_scalar())
              cadd(acc, tensors[nodes[c].ref].to_dense_data()[0]);
            else
              throw std::runtime_error("NotAllScalars");
          }
          replace_with_leaf(id, N_SCALAR, push_scalar(acc));
        } else {
          Tensor acc = leaf_tensor(first);
          for (size_t k = 1; k < n.children.size(); ++k) {
            int c = n.children[k];
            if (nodes[c].kind == N_SCALAR) throw std::runtime_error("SumScalarTensor");
            add_assign(acc, leaf_tensor(c));
          }
          replace_with_leaf(id, N_TENSOR, push_tensor(std::move(acc)));
        }
        return;
      }
      case N_NEG: {  // network.rs:1284-1336
        int c = n.children.at(0);
        if (nodes[c].kind == N_SCALAR)
          replace_with_leaf(id, N_SCALAR, push_scalar(cneg(scalars[nodes[c].ref])));
        else
          replace_with_leaf(id, N_TENSOR, push_tensor(neg(leaf_tensor(c))));
        return;
      }
      case N_CONJ: {  // FUN_LIB conj — spenso-hep-lib/src/lib.rs:511-520
        int c = n.children.at(0);
        if (nodes[c].kind == N_SCALAR) {
          C s = scalars[nodes[c].ref];
          replace_with_leaf(id, N_SCALAR, push_scalar(C{s.re, -s.im}));
        } else {
          Tensor t = leaf_tensor(c);
          for (auto& x : t.data) x.im = -x.im;
          for (auto& kv : t.elems) kv.second.im = -kv.second.im;
          replace_with_leaf(id, N_TENSOR, push_tensor(std::move(t)));
        }
        return;
      }
      case N_POWER: {  // network.rs:1526-1703
        int c = n.children.at(0);
        int pow = n.power;
        int np = std::abs(pow);
        if (nodes[c].kind == N_SCALAR) {
          if (np == 0) {
            replace_with_leaf(id, N_SCALAR, nodes[c].ref);
            return;
          }
          C s = scalars[nodes[c].ref];
          for (int k = 1; k < np; ++k) s = cmul(s, scalars[nodes[c].ref]);
          if (pow < 0) s = cpow_inv(s);
          replace_with_leaf(id, N_SCALAR, push_scalar(s));
          return;
        }
        Tensor t = leaf_tensor(c);
        if (pow == 0) {
          replace_with_leaf(id, N_SCALAR, push_scalar(C{1, 0}));
          return;
        }
        if (pow == 1) {
          Node l = nodes[c];
          replace_with_leaf(id, l.kind, l.ref);
          return;
        }
        int squares = np / 2;
        Tensor square = contract(t, t, opt);
        if (np % 2 == 1) {
          if (np != 1) {  // network.rs:1591-1596 / 1647-1652: n == 1 (pow == -1) leaves t untouched
            for (int k = 0; k < squares; ++k) square = contract(square, square, opt);
            t = contract(square, t, opt);
          }
          if (pow < 0) {
            if (!t.st.is_scalar()) throw std::runtime_error("NegativeExponentNonScalar");
            replace_with_leaf(id, N_SCALAR, push_scalar(cpow_inv(t.to_dense_data()[0])));
          } else {
            replace_with_leaf(id, N_TENSOR, push_tensor(std::move(t)));
          }
        } else {
          if (!square.st.is_scalar()) throw std::runtime_error("NoScalar");
          C s = square.to_dense_data()[0], sc = s;
          for (int k = 1; k < squares; ++k) s = cmul(s, sc);
          if (pow < 0) s = cpow_inv(s);
          replace_with_leaf(id, N_SCALAR, push_scalar(s));
        }
        return;
      }
      default: throw std::runtime_error("not an op node");
    }
  }

  // Network::execute::<Sequential, SmallestDegree> — network.rs:1217-1241, 1132-1155
  void execute() {
    if (root < 0) throw std::runtime_error("network has no root");
    merge_ops(root);
    int op;
    while (find_ready(root, op)) execute_op(op);
  }
  // result_tensor / result_scalar — network.rs:807-900
  Tensor result_tensor() const {
    const Node& n = nodes[root];
    if (n.kind == N_TENSOR) return tensors[n.ref];
    if (n.kind == N_LIB) return get_lib_data(root);
    if (n.kind == N_SCALAR) return Tensor::dense(Structure(), {scalars[n.ref]});
    throw std::runtime_error("network not fully executed");
  }
};

// ---------------------------------------------------------------------------
// D1: shared tensor data — spenso-hep-lib/src/lib.rs:27-387 (Weyl basis),
// network/library/symbolic.rs:576-659 (metric / identity).
// Library argument order is (bis i, bis j, mink mu) after rep-canonical sort
// of the key (idenso/src/gamma.rs:477-488: key written [mink, bis, bis]).
// ---------------------------------------------------------------------------
inline Slot mk_slot(uint8_t kind, int16_t id, uint32_t dim, uint64_t aind = 0, uint8_t aind_kind = AIND_NORMAL) {
  return Slot{kind, aind_kind, id, dim, aind};
}
// rep ids after idenso::representations::initialize()
inline Slot euc(uint32_t d, uint64_t a) { return mk_slot(REP_SELF_DUAL, 0, d, a); }
inline Slot bis(uint32_t d, uint64_t a) { return mk_slot(REP_SELF_DUAL, 1, d, a); }
inline Slot coad(uint32_t d, uint64_t a) { return mk_slot(REP_SELF_DUAL, 2, d, a); }
inline Slot mink(uint32_t d, uint64_t a) { return mk_slot(REP_INLINE_METRIC, 0, d, a); }
inline Slot lor(uint32_t d, uint64_t a, bool dual = false) { return mk_slot(REP_DUALIZABLE, dual ? -1 : 1, d, a); }
inline Slot spf(uint32_t d, uint64_t a, bool dual = false) { return mk_slot(REP_DUALIZABLE, dual ? -2 : 2, d, a); }
inline Slot cof(uint32_t d, uint64_t a, bool dual = false) { return mk_slot(REP_DUALIZABLE, dual ? -3 : 3, d, a); }

inline LibTensor lib_from_entries(std::vector<Slot> reps, const std::vector<std::pair<std::vector<size_t>, C>>& e) {
  LibTensor lt;
  lt.reps = reps;
  lt.sparse = true;
  size_t n = reps.size();
  std::vector<size_t> st(n, 1);
  for (size_t i = n ? n - 1 : 0; i-- > 0;) st[i] = st[i + 1] * reps[i + 1].dim;
  for (auto& kv : e) {
    size_t f = 0;
    for (size_t i = 0; i < n; ++i) f += kv.first[i] * st[i];
    lt.nz.emplace_back(f, kv.second);
  }
  return lt;
}
inline LibTensor gamma_weyl() {  // lib.rs:66-102
  const C c1{1, 0}, cn1{-1, 0}, ci{0, 1}, cni{0, -1};
  return lib_from_entries({bis(4, 0), bis(4, 0), mink(4, 0)},
                          {{{0, 2, 0}, c1}, {{1, 3, 0}, c1}, {{2, 0, 0}, c1}, {{3, 1, 0}, c1},
                           {{0, 3, 1}, c1}, {{1, 2, 1}, c1}, {{2, 1, 1}, cn1}, {{3, 0, 1}, cn1},
                           {{0, 3, 2}, cni}, {{1, 2, 2}, ci}, {{2, 1, 2}, ci}, {{3, 0, 2}, cni},
                           {{0, 2, 3}, c1}, {{1, 3, 3}, cn1}, {{2, 0, 3}, cn1}, {{3, 1, 3}, c1}});
}
inline LibTensor gamma_conj_weyl() {  // lib.rs:144-153: gamma_data_weyl with every entry conjugated
  const C c1{1, 0}, cn1{-1, 0}, ci{0, 1}, cni{0, -1};
  return lib_from_entries({bis(4, 0), bis(4, 0), mink(4, 0)},
                          {{{0, 2, 0}, c1}, {{1, 3, 0}, c1}, {{2, 0, 0}, c1}, {{3, 1, 0}, c1},
                           {{0, 3, 1}, c1}, {{1, 2, 1}, c1}, {{2, 1, 1}, cn1}, {{3, 0, 1}, cn1},
                           {{0, 3, 2}, ci}, {{1, 2, 2}, cni}, {{2, 1, 2}, cni}, {{3, 0, 2}, ci},
                           {{0, 2, 3}, c1}, {{1, 3, 3}, cn1}, {{2, 0, 3}, cn1}, {{3, 1, 3}, c1}});
}
inline LibTensor gamma_adj_weyl() {  // lib.rs:155-166: gamma_transpose_weyl (lib.rs:103-141) conjugated
  const C c1{1, 0}, cn1{-1, 0}, ci{0, 1}, cni{0, -1};
  return lib_from_entries({bis(4, 0), bis(4, 0), mink(4, 0)},
                          {{{2, 0, 0}, c1}, {{3, 1, 0}, c1}, {{0, 2, 0}, c1}, {{1, 3, 0}, c1},
                           {{3, 0, 1}, c1}, {{2, 1, 1}, c1}, {{1, 2, 1}, cn1}, {{0, 3, 1}, cn1},
                           {{3, 0, 2}, ci}, {{2, 1, 2}, cni}, {{1, 2, 2}, cni}, {{0, 3, 2}, ci},
                           {{2, 0, 3}, c1}, {{3, 1, 3}, cn1}, {{0, 2, 3}, cn1}, {{1, 3, 3}, c1}});
}
inline LibTensor gamma_dirac() {  // lib.rs:27-63
  const C c1{1, 0}, cn1{-1, 0}, ci{0, 1}, cni{0, -1};
  return lib_from_entries({bis(4, 0), bis(4, 0), mink(4, 0)},
                          {{{0, 0, 0}, c1}, {{0, 1, 1}, c1}, {{0, 2, 2}, cn1}, {{0, 3, 3}, cn1},
                           {{1, 0, 3}, c1}, {{1, 1, 2}, c1}, {{1, 2, 1}, cn1}, {{1, 3, 0}, cn1},
                           {{2, 0, 3}, cni}, {{2, 1, 2}, ci}, {{2, 2, 1}, ci}, {{2, 3, 0}, cni},
                           {{3, 0, 2}, c1}, {{3, 1, 3}, cn1}, {{3, 2, 0}, cn1}, {{3, 3, 1}, c1}});
}
inline LibTensor gamma0_weyl() {  // lib.rs:167-185
  const C c1{1, 0};
  return lib_from_entries({bis(4, 0), bis(4, 0)}, {{{0, 2}, c1}, {{1, 3}, c1}, {{2, 0}, c1}, {{3, 1}, c1}});
}
inline LibTensor gamma5_weyl() {  // lib.rs:205-221
  const C c1{1, 0}, cn1{-1, 0};
  return lib_from_entries({bis(4, 0), bis(4, 0)}, {{{0, 0}, cn1}, {{1, 1}, cn1}, {{2, 2}, c1}, {{3, 3}, c1}});
}
inline LibTensor projm_weyl() {  // lib.rs:251-265
  const C c1{1, 0};
  return lib_from_entries({bis(4, 0), bis(4, 0)}, {{{0, 0}, c1}, {{1, 1}, c1}});
}
inline LibTensor projp_weyl() {  // lib.rs:308-322
  const C c1{1, 0};
  return lib_from_entries({bis(4, 0), bis(4, 0)}, {{{2, 2}, c1}, {{3, 3}, c1}});
}
// metric g(i,j): diag(+,-,-,-...) for inline-metric reps, identity otherwise (symbolic.rs:576-659)
inline LibTensor metric_lib(Slot rep) {
  std::vector<std::pair<std::vector<size_t>, C>> e;
  for (size_t i = 0; i < rep.dim; ++i) e.push_back({{i, i}, rep_is_neg(rep, i) ? C{-1, 0} : C{1, 0}});
  Slot other = rep;
  return lib_from_entries({rep, slot_dual(other)}, e);
}

}  // namespace orc
