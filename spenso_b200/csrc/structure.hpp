// Host-side index algebra of the engine: canonical slot order, structure merge and the
// stride plan of one pairwise contraction.  Integer work only; runs once per contraction
// (or once per network), never per phase-space point.
//
// Behaviour to reproduce (reference, paths relative to /root/reference/):
//   slot order        spenso/src/structure/slot.rs:25-56, representation.rs:466-474, 591-620,
//                     abstract_index.rs:267-294, dimension.rs:53-61
//   canonical form    spenso/src/structure/ordered.rs:321-380
//   merge             spenso/src/structure/ordered.rs:594-919, structure.rs:716-751
//   index matching    spenso/src/structure.rs:435-460
//   strides           spenso/src/structure.rs:566-577
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/spenso_b200.h"

namespace spn {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---- slot ordering -------------------------------------------------------------------
// The reference order is lexicographic over (rep, dim, aind) with
//   Dummy < SelfDual(n) < InlineMetric(n) < Dualizable(+k) < Dualizable(-k),
// positives by k, negatives by |k|.  We encode that as one comparable key.
struct SlotKey {
  int rep_class;  // 0 dummy, 1 self-dual, 2 inline metric, 3 dualizable base, 4 dualizable dual
  int rep_num;
  uint32_t dim;
  int aind_kind;
  uint64_t aind;
  auto tie() const { return std::tie(rep_class, rep_num, dim, aind_kind, aind); }
  bool operator<(const SlotKey& o) const { return tie() < o.tie(); }
  bool operator==(const SlotKey& o) const { return tie() == o.tie(); }
};

inline bool is_dualizable(const spn_slot& s) { return s.rep_kind == SPN_REP_DUALIZABLE; }
inline bool is_base(const spn_slot& s) { return is_dualizable(s) && s.rep_id > 0; }
inline bool is_dual(const spn_slot& s) { return is_dualizable(s) && s.rep_id < 0; }

inline SlotKey slot_key(const spn_slot& s) {
  SlotKey k{};
  switch (s.rep_kind) {
    case SPN_REP_DUMMY:
      k.rep_class = 0;
      k.rep_num = 0;
      k.dim = 0;  // two dummy representations compare equal whatever their dimension
      break;
    case SPN_REP_SELF_DUAL: k.rep_class = 1; k.rep_num = s.rep_id; k.dim = s.dim; break;
    case SPN_REP_INLINE_METRIC: k.rep_class = 2; k.rep_num = s.rep_id; k.dim = s.dim; break;
    case SPN_REP_DUALIZABLE:
      if (s.rep_id == 0) throw Error(SPN_ERR_STRUCTURE, "dualizable representation with id 0");
      k.rep_class = s.rep_id > 0 ? 3 : 4;
      k.rep_num = s.rep_id > 0 ? s.rep_id : -s.rep_id;
      k.dim = s.dim;
      break;
    default: throw Error(SPN_ERR_INVALID, "unknown representation kind");
  }
  k.aind_kind = s.aind_kind;
  k.aind = s.aind;
  return k;
}
inline bool slot_less(const spn_slot& a, const spn_slot& b) { return slot_key(a) < slot_key(b); }
inline bool slot_same(const spn_slot& a, const spn_slot& b) {
  return a.rep_kind == b.rep_kind && a.rep_id == b.rep_id && a.dim == b.dim && a.aind_kind == b.aind_kind &&
         a.aind == b.aind;
}
inline spn_slot dual_of(spn_slot s) {
  if (is_dualizable(s)) s.rep_id = (int16_t)-s.rep_id;
  return s;
}
// does index value i of this slot carry a minus sign when contracted?  Only inline-metric
// representations do; the one that exists is Minkowski, diag(+,-,-,...) (representation.rs:321-323)
inline bool has_metric(const spn_slot& s) { return s.rep_kind == SPN_REP_INLINE_METRIC; }

// Key used when a base slot looks for its partner among the other operand's duals
// (dimension first, then representation number, then index): ordered.rs:717-758.
inline std::tuple<uint32_t, int, int, uint64_t> pairing_key(const spn_slot& s) {
  return {s.dim, s.rep_id < 0 ? -s.rep_id : s.rep_id, (int)s.aind_kind, s.aind};
}

// ---- canonical structure ---------------------------------------------------------------
struct Structure {
  std::vector<spn_slot> slots;  // canonical: [self-dual-like | base | dual], each block sorted
  int64_t base_start = -1, dual_start = -1;

  uint32_t order() const { return (uint32_t)slots.size(); }
  uint32_t n_selfdual() const {
    uint32_t n = 0;
    for (auto& s : slots) n += !is_dualizable(s);
    return n;
  }
  uint32_t n_base() const {
    uint32_t n = 0;
    for (auto& s : slots) n += is_base(s);
    return n;
  }
  uint32_t n_dual() const {
    uint32_t n = 0;
    for (auto& s : slots) n += is_dual(s);
    return n;
  }
  uint64_t size() const {
    uint64_t n = 1;
    for (auto& s : slots) n *= s.dim;
    return n;
  }
  std::vector<uint64_t> strides() const {  // row major, last index fastest
    std::vector<uint64_t> st(slots.size(), 1);
    for (size_t i = slots.size(); i-- > 1;) st[i - 1] = st[i] * slots[i].dim;
    return st;
  }
};

inline void set_block_starts(Structure& st) {
  uint32_t nsd = st.n_selfdual(), nb = st.n_base(), nd = st.n_dual();
  st.base_start = (nb + nd) ? (int64_t)nsd : -1;
  st.dual_start = nd ? (int64_t)(nsd + nb) : -1;
}

// perm[i] = position in `in` of the slot that lands at canonical position i
inline Structure canonicalize(const std::vector<spn_slot>& in, std::vector<int32_t>* perm = nullptr) {
  std::vector<int32_t> idx(in.size());
  for (size_t i = 0; i < in.size(); ++i) idx[i] = (int32_t)i;
  std::vector<SlotKey> keys;
  keys.reserve(in.size());
  for (auto& s : in) keys.push_back(slot_key(s));
  std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return keys[a] < keys[b]; });
  Structure st;
  for (int32_t i : idx) st.slots.push_back(in[i]);
  set_block_starts(st);
  if (perm) *perm = idx;
  return st;
}
inline bool is_canonical(const std::vector<spn_slot>& s) {
  for (size_t i = 0; i + 1 < s.size(); ++i)
    if (slot_key(s[i + 1]) < slot_key(s[i])) return false;
  return true;
}
inline Structure structure_of_canonical(const spn_slot* s, uint32_t n) {
  Structure st;
  st.slots.assign(s, s + n);
  if (!is_canonical(st.slots)) throw Error(SPN_ERR_STRUCTURE, "slots are not in canonical order");
  set_block_starts(st);
  return st;
}

// ---- merge --------------------------------------------------------------------------------
struct Merged {
  Structure st;
  std::vector<uint8_t> common_a, common_b, partition;
  int kind = SPN_FIRST_BEFORE_SECOND;
};

inline int classify_partition(const std::vector<uint8_t>& p) {
  if (p.empty()) return SPN_FIRST_BEFORE_SECOND;
  int flips = 0;
  for (size_t i = 1; i < p.size(); ++i) flips += p[i] != p[i - 1];
  if (flips > 1) return SPN_INTERLEAVED;
  return p[0] ? SPN_FIRST_BEFORE_SECOND : SPN_SECOND_BEFORE_FIRST;
}

// Three passes over the canonical blocks.  Pass 2 pairs every base slot with a dual of the
// OTHER operand using a forward-only cursor over that operand's dual block, exactly like the
// reference (including its behaviour when the pairing key and the sort key disagree).
inline Merged merge(const Structure& A, const Structure& B) {
  Merged m;
  m.common_a.assign(A.order(), 0);
  m.common_b.assign(B.order(), 0);
  if (A.order() == 0 || B.order() == 0) {
    m.st = A.order() == 0 ? B : A;
    m.kind = A.order() == 0 ? SPN_SECOND_BEFORE_FIRST : SPN_FIRST_BEFORE_SECOND;
    return m;
  }
  auto emit = [&](const spn_slot& s, bool from_a) {
    m.st.slots.push_back(s);
    m.partition.push_back(from_a ? 1 : 0);
  };
  const uint32_t sdA = A.n_selfdual(), sdB = B.n_selfdual();
  const uint32_t bA = A.n_base(), bB = B.n_base();
  const uint32_t dA = A.n_dual(), dB = B.n_dual();

  // pass 1: self-dual-like slots — plain sorted merge, equal slots are contracted
  {
    uint32_t i = 0, j = 0;
    while (i < sdA || j < sdB) {
      if (j >= sdB) {
        emit(A.slots[i++], true);
      } else if (i >= sdA) {
        emit(B.slots[j++], false);
      } else {
        SlotKey ka = slot_key(A.slots[i]), kb = slot_key(B.slots[j]);
        if (ka < kb) {
          emit(A.slots[i++], true);
        } else if (kb < ka) {
          emit(B.slots[j++], false);
        } else {
          if (A.slots[i].rep_kind == SPN_REP_DUMMY || A.slots[i].dim != B.slots[j].dim)
            throw Error(SPN_ERR_STRUCTURE, "Matching and equal items must be identical");
          m.common_a[i++] = 1;
          m.common_b[j++] = 1;
        }
      }
    }
  }
  const int64_t base_start = (int64_t)m.partition.size();

  // pass 2: bases, each looking for a partner in the other operand's dual block
  {
    uint32_t curA = 0, curB = 0;  // cursors into the dual blocks of A and B
    auto find_partner = [](const spn_slot& base, const Structure& other, uint32_t dual_off, uint32_t n_dual,
                           uint32_t& cursor, std::vector<uint8_t>& other_common) -> bool {
      auto kb = pairing_key(base);
      while (cursor < n_dual) {
        const spn_slot& d = other.slots[dual_off + cursor];
        auto kd = pairing_key(d);
        if (kb < kd) return false;  // everything further is larger still
        if (kd < kb) {
          ++cursor;
          continue;
        }
        if (!(d.rep_id == -base.rep_id)) throw Error(SPN_ERR_STRUCTURE, "Matching and equal items must be identical");
        other_common[dual_off + cursor] = 1;
        ++cursor;
        return true;
      }
      return false;
    };
    uint32_t i = 0, j = 0;
    while (i < bA || j < bB) {
      bool take_a;
      if (j >= bB)
        take_a = true;
      else if (i >= bA)
        take_a = false;
      else {
        SlotKey ka = slot_key(A.slots[sdA + i]), kb = slot_key(B.slots[sdB + j]);
        if (ka == kb) throw Error(SPN_ERR_STRUCTURE, "Cannot have equal bases");
        take_a = ka < kb;
      }
      if (take_a) {
        const spn_slot& s = A.slots[sdA + i];
        if (find_partner(s, B, sdB + bB, dB, curB, m.common_b))
          m.common_a[sdA + i] = 1;
        else
          emit(s, true);
        ++i;
      } else {
        const spn_slot& s = B.slots[sdB + j];
        if (find_partner(s, A, sdA + bA, dA, curA, m.common_a))
          m.common_b[sdB + j] = 1;
        else
          emit(s, false);
        ++j;
      }
    }
  }
  const int64_t dual_start = (int64_t)m.partition.size();

  // pass 3: whatever duals were not consumed
  {
    uint32_t i = 0, j = 0;
    const uint32_t offA = sdA + bA, offB = sdB + bB;
    while (i < dA || j < dB) {
      bool take_a;
      if (j >= dB)
        take_a = true;
      else if (i >= dA)
        take_a = false;
      else {
        SlotKey ka = slot_key(A.slots[offA + i]), kb = slot_key(B.slots[offB + j]);
        if (ka == kb) throw Error(SPN_ERR_STRUCTURE, "Cannot have equal duals");
        take_a = ka < kb;
      }
      if (take_a) {
        if (!m.common_a[offA + i]) emit(A.slots[offA + i], true);
        ++i;
      } else {
        if (!m.common_b[offB + j]) emit(B.slots[offB + j], false);
        ++j;
      }
    }
  }
  if (!is_canonical(m.st.slots)) throw Error(SPN_ERR_STRUCTURE, "merged structure is not canonically sorted");
  m.st.base_start = base_start;
  m.st.dual_start = dual_start;
  m.kind = classify_partition(m.partition);
  return m;
}

// ---- the stride plan of one pairwise contraction ----------------------------------------------
inline spn_contract_plan plan_contract(const Structure& A, const Structure& B) {
  if (A.order() > SPN_MAX_ORDER || B.order() > SPN_MAX_ORDER)
    throw Error(SPN_ERR_INVALID, "operand order exceeds SPN_MAX_ORDER");
  Merged m = merge(A, B);
  if (m.st.order() > SPN_MAX_OUT_ORDER) throw Error(SPN_ERR_INVALID, "result order exceeds SPN_MAX_OUT_ORDER");
  spn_contract_plan p{};
  p.order_a = A.order();
  p.order_b = B.order();
  p.order_out = m.st.order();
  p.merge_kind = (uint32_t)m.kind;
  p.n_partition = (uint32_t)m.partition.size();
  p.base_start = m.st.base_start;
  p.dual_start = m.st.dual_start;
  p.size_a = A.size();
  p.size_b = B.size();
  p.size_out = m.st.size();
  for (uint32_t i = 0; i < A.order(); ++i) p.n_common += (p.common_a[i] = m.common_a[i]);
  for (uint32_t j = 0; j < B.order(); ++j) p.common_b[j] = m.common_b[j];
  for (uint32_t r = 0; r < p.n_partition; ++r) p.partition[r] = m.partition[r];
  auto sa = A.strides(), sb = B.strides();

  // result dims: walk the free slots of each operand in order; the partition says whose turn it is
  {
    uint32_t ia = 0, ib = 0;
    auto next_free = [](const std::vector<uint8_t>& common, uint32_t& i) {
      while (i < common.size() && common[i]) ++i;
      return i++;
    };
    for (uint32_t r = 0; r < p.order_out; ++r) {
      p.out_slots[r] = m.st.slots[r];
      p.out_dim[r] = m.st.slots[r].dim;
      bool from_a = p.n_partition ? m.partition[r] != 0 : (A.order() != 0);
      if (from_a) {
        uint32_t pos = next_free(m.common_a, ia);
        p.out_stride_a[r] = sa[pos];
        p.out_stride_b[r] = 0;
      } else {
        uint32_t pos = next_free(m.common_b, ib);
        p.out_stride_a[r] = 0;
        p.out_stride_b[r] = sb[pos];
      }
    }
  }
  // contracted pairs, enumerated in the order of the operand the reference walks with its flat
  // iterator ("other"): b normally, a when the reference swaps receiver (SecondBeforeFirst)
  {
    struct Pair {
      uint32_t ia, ib;
    };
    std::vector<Pair> pairs;
    for (uint32_t i = 0; i < A.order(); ++i) {
      if (!m.common_a[i]) continue;
      spn_slot want = dual_of(A.slots[i]);
      bool found = false;
      for (uint32_t j = 0; j < B.order(); ++j)
        if (m.common_b[j] && slot_same(B.slots[j], want)) {
          pairs.push_back({i, j});
          found = true;
          break;
        }
      if (!found) throw Error(SPN_ERR_STRUCTURE, "contracted slot without partner");
    }
    bool other_is_a = m.kind == SPN_SECOND_BEFORE_FIRST;
    std::sort(pairs.begin(), pairs.end(),
              [&](const Pair& x, const Pair& y) { return other_is_a ? x.ia < y.ia : x.ib < y.ib; });
    p.n_contracted = 1;
    for (size_t d = 0; d < pairs.size(); ++d) {
      const spn_slot& s = A.slots[pairs[d].ia];
      p.k_dim[d] = s.dim;
      p.k_stride_a[d] = sa[pairs[d].ia];
      p.k_stride_b[d] = sb[pairs[d].ib];
      p.k_metric[d] = has_metric(s) ? 1 : 0;
      p.k_pos_a[d] = (uint8_t)pairs[d].ia;
      p.k_pos_b[d] = (uint8_t)pairs[d].ib;
      p.n_contracted *= s.dim;
    }
  }
  return p;
}

inline Structure structure_of_plan_result(const spn_contract_plan& p) {
  Structure st;
  st.slots.assign(p.out_slots, p.out_slots + p.order_out);
  st.base_start = p.base_start;
  st.dual_start = p.dual_start;
  return st;
}

}  // namespace spn
