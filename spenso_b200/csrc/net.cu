// Network executor: the B200 replacement for Network::execute
// (/root/reference/spenso/src/network.rs:1217-1241, 1244-1706) and the Product-op contraction
// strategy (/root/reference/spenso/src/network/contract.rs:43-498).
//
// The reference walks the expression graph, merges structures and picks a pairwise contraction
// order again for every phase-space point.  All of that is point independent, so here it runs
// once (spn_net_compile) and produces a linear schedule of pairwise steps.  The schedule is then
// executed for all points at once in one of two ways:
//   SPN_EXEC_FUSED     the schedule is replayed symbolically (symbolic.hpp) into ONE straight-line
//                      sm_100a kernel: thread = point, inputs read once with 128/256-bit loads,
//                      every intermediate in registers, library constants folded into the code,
//                      optional Monte-Carlo sum reduced per block.  Algorithmic HBM traffic only.
//   SPN_EXEC_STEPWISE  one pairwise kernel (kernels.cu) per step with intermediates in HBM —
//                      the literal device image of the reference's append-only NetworkStore.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <set>
#include <sstream>

#include "jit.hpp"
#include "json_min.hpp"
#include "library.hpp"
#include "symbolic.hpp"
#include "tensor.hpp"

using namespace spn;

namespace {

enum { LEAF_BATCHED = 0, LEAF_CONST = 1, NODE_OP = 2, LEAF_FUNCTION = 3 };

struct Node {
  int kind = LEAF_CONST;
  int op = 0;  // SPN_OP_*
  Structure st;
  int input = -1;  // LEAF_BATCHED
  int dtype = SPN_C64;
  bool sparse = false;  // LEAF_CONST
  std::vector<double2> dense;
  std::map<uint64_t, double2> entries;
  std::vector<int> children;
  int power = 1;
  // LEAF_BATCHED: slots as the caller wrote them + canonical permutation (canonical[i] = args[perm[i]])
  std::vector<spn_slot> args;
  std::vector<int32_t> perm;
  // re-indexed view of another batched leaf: canonical flat index here -> canonical flat index of
  // the bound input tensor (empty = identity)
  std::vector<uint64_t> flat_map;
  int alias_of = -1;  // node id of the batched leaf this one re-indexes
  bool is_scalar_leaf = false;  // created by spn_net_leaf_scalar (NetworkLeaf::Scalar)
  bool from_library = false;    // NetworkLeaf::LibraryKey (spn_net_leaf_library / _custom); else a LocalTensor
  // LEAF_FUNCTION: per-point tensor computed on the device from components of other batched leaves
  std::vector<std::pair<int, uint64_t>> fn_params;  // (leaf node id, canonical flat component)
  std::vector<double2> fn_consts;
  std::vector<std::vector<std::pair<int, int>>> fn_code;  // per CANONICAL component: postfix (op, arg)
  std::vector<int32_t> fn_raw_code;                         // as given (argument order), for export
};

enum StepOp { ST_CONTRACT = 0, ST_ADD = 1, ST_NEG = 2, ST_CONJ = 3, ST_RECIP = 4 };
struct Step {
  int op, dst, a, b;
};
struct Value {
  Structure st;
  bool batched = false;
  int leaf = -1;  // node id for leaf values
  // the reference distinguishes NetworkLeaf::Scalar from a rank-0 LocalTensor (network.rs:1526-1703 treats
  // them differently in Power); true = Scalar
  bool scalar_kind = false;
};

}  // namespace

// =================================================================================================
struct spn_net {
  spn_ctx* ctx = nullptr;
  std::vector<Node> nodes;
  int root = -1;
  uint32_t n_inputs = 0;
  std::vector<int> input_nodes;
  // compiled
  bool compiled = false;
  int exec_mode = SPN_EXEC_FUSED;
  std::vector<Value> values;
  std::vector<Step> steps;
  // (n1, n2) of every pairwise contraction in execution order, in the reference's naming: a factor of a product is
  // known by its node id, and the result of n1.contract(n2) takes the place — and the id — of n1 (the graph surgery of
  // SingleSmallestDegree::contract, network/contract.rs:346-497, which the oracle's contraction log records too)
  std::vector<std::pair<int, int>> schedule;
  // the HOST's contraction order (spn_net_set_schedule): the same naming; consumed by schedule_product
  std::vector<std::pair<int, int>> host_schedule;
  std::vector<char> host_used;
  bool has_host_schedule = false;
  int result_value = -1;
  std::map<int, int> node_value;  // node id -> value id of the compiled schedule
  std::vector<int> input_nodes_all;  // ids of every LEAF_BATCHED node, including re-indexed views
  std::map<int, std::unique_ptr<spn_net>> function_leaf_nets;  // stepwise: per function leaf, its own fused kernel
  // fused
  Program prog;
  std::vector<CE> result_el;
  Program::Stats stats;
  struct Variant {
    std::unique_ptr<JitKernel> k;
  };
  std::map<std::string, Variant> variants;

  std::string default_source, json;
  double2* sum_scratch = nullptr;  // per-block partials of the fused Monte-Carlo sum
  spn_comm* comm = nullptr;        // optional: sums are exchanged with the peer GPUs (comm.cu)
  size_t sum_scratch_elems = 0;
  // host-buffer pipeline (spn_net_execute_host): kPipe slots, each with its own stream and device staging
  static constexpr int kPipe = 3;
  struct HostPipe {
    cudaStream_t stream[kPipe] = {nullptr, nullptr, nullptr};
    std::vector<void*> in[kPipe];
    void* out[kPipe] = {nullptr, nullptr, nullptr};
    double2* partial[kPipe] = {nullptr, nullptr, nullptr};
    double2* chunk_sums = nullptr;   // [2 call parities][n_chunks_cap][size_out]
    double2* final_sum[2] = {nullptr, nullptr};
    uint64_t chunk_points = 0, n_chunks_cap = 0;
    size_t partial_elems = 0;
    // a call's last stage (fold of the per-chunk sums + D2H of the result) runs on `fin`, so that the chunk streams can
    // take the next call's chunks at once: the pipeline does not drain between calls
    cudaStream_t fin = nullptr;
    cudaEvent_t chunk_done[kPipe] = {nullptr, nullptr, nullptr};   // last chunk of the current call on each stream
    cudaEvent_t fin_done[2] = {nullptr, nullptr};                  // final stage of the call with this parity
    bool fin_pending[2] = {false, false};
    uint64_t calls = 0;
    void release() {
      for (int s = 0; s < kPipe; ++s) {
        for (void* p : in[s]) cudaFree(p);
        in[s].clear();
        if (out[s]) cudaFree(out[s]);
        if (partial[s]) cudaFree(partial[s]);
        out[s] = nullptr;
        partial[s] = nullptr;
      }
      if (chunk_sums) cudaFree(chunk_sums);
      chunk_sums = nullptr;
      for (int q = 0; q < 2; ++q) {
        if (final_sum[q]) cudaFree(final_sum[q]);
        final_sum[q] = nullptr;
        fin_pending[q] = false;
      }
      chunk_points = n_chunks_cap = 0;
      partial_elems = 0;
    }
  } pipe;
  ~spn_net() {
    if (sum_scratch) cudaFree(sum_scratch);
    pipe.release();
    for (int s = 0; s < kPipe; ++s) {
      if (pipe.stream[s]) cudaStreamDestroy(pipe.stream[s]);
      if (pipe.chunk_done[s]) cudaEventDestroy(pipe.chunk_done[s]);
    }
    for (int q = 0; q < 2; ++q)
      if (pipe.fin_done[q]) cudaEventDestroy(pipe.fin_done[q]);
    if (pipe.fin) cudaStreamDestroy(pipe.fin);
  }
  // stepwise
  std::vector<std::unique_ptr<spn_tensor>> const_tensors;  // per value id (leaf consts), lazily uploaded

  // threads per CTA of the fused kernels.  256: measured never slower than 128 (cfg2, cfg2mr: equal) and faster for the
  // ring-prefetch kernels (cfg5 +3..7 %, cfg5r +10 %): one 256-thread CTA per SM amortises the ring fill and the
  // per-CTA fold of the Monte-Carlo partials over twice the points (profiles/r01_knob_sweep_*.txt)
  int kBlock = getenv("SPN_BLOCK") ? atoi(getenv("SPN_BLOCK")) : 256;
  int kMinBlocks = getenv("SPN_MIN_BLOCKS") ? atoi(getenv("SPN_MIN_BLOCKS")) : 0;
  static constexpr uint64_t kMaxFusedElems = 1u << 16;

  // ---- graph -> schedule ----------------------------------------------------------------------
  int new_value(const Structure& st, bool batched, int leaf = -1) {
    Value v;
    v.st = st;
    v.batched = batched;
    v.leaf = leaf;
    values.push_back(v);
    return (int)values.size() - 1;
  }
  int emit(int op, int a, int b, Structure st) {  // by value: callers pass values[..].st, which new_value moves
    bool batched = values[a].batched || (b >= 0 && values[b].batched);
    int dst = new_value(st, batched);
    steps.push_back({op, dst, a, b});
    return dst;
  }
  int emit_contract(int a, int b, int name_a = -1, int name_b = -1) {
    spn_contract_plan p = plan_contract(values[a].st, values[b].st);
    schedule.emplace_back(name_a, name_b);
    return emit(ST_CONTRACT, a, b, structure_of_plan_result(p));
  }
  static bool same_structure(const Structure& a, const Structure& b) {
    if (a.order() != b.order()) return false;
    for (uint32_t i = 0; i < a.order(); ++i)
      if (!slot_same(a.slots[i], b.slots[i])) return false;
    return true;
  }
  static uint32_t n_shared_indices(const Structure& a, const Structure& b) {
    uint32_t n = 0;
    for (auto& s : a.slots) {
      spn_slot d = dual_of(s);
      for (auto& t : b.slots)
        if (slot_same(t, d)) {
          ++n;
          break;
        }
    }
    return n;
  }

  // Product op: pairwise contraction order.  The reference picks the node of smallest degree and
  // a neighbour (contract.rs:346-497); its order depends on third-party graph iteration order and
  // values do not depend on it, so we are free to choose: contract the connected pair
  // with the smallest result (ties: point-independent pairs first — they cost nothing per point —
  // then fewest multiply-adds); unconnected factors (scalars first) are multiplied in last.
  // With trial > 0 (fused compiler's schedule search) the choice among near-cheapest connected pairs is
  // randomised by a deterministic LCG; the trial whose generated kernel has the fewest FP64 instructions wins.
  uint64_t rng_state = 0;
  int trial = 0;
  double rnd() {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(rng_state >> 11) / 9007199254740992.0;
  }
  // `names[i]`: the node id factor i is known by (see `schedule`)
  int schedule_product(std::vector<int> vals, std::vector<int> names) {
    if (vals.empty()) throw Error(SPN_ERR_INVALID, "empty product");
    struct Cand {
      size_t i, j;
      std::array<double, 3> cost;
    };
    auto contract_pair = [&](size_t bi, size_t bj) {
      // n1.contract(n2), except (LibraryKey, LocalTensor), which the reference runs as local.contract(lib)
      // (contract.rs:428-443) — the operand order decides the rounding of the pairwise kernels
      int a = vals[bi], b = vals[bj];
      auto is_lib = [&](int v) { return values[(size_t)v].leaf >= 0 && nodes[(size_t)values[(size_t)v].leaf].from_library; };
      auto is_local = [&](int v) { return values[(size_t)v].leaf < 0 || !nodes[(size_t)values[(size_t)v].leaf].from_library; };
      if (has_host_schedule && is_lib(a) && is_local(b) && !values[(size_t)b].scalar_kind) std::swap(a, b);
      int r = emit_contract(a, b, names[bi], names[bj]);
      vals[bi] = r;
      vals.erase(vals.begin() + (long)bj);
      names.erase(names.begin() + (long)bj);
    };
    if (has_host_schedule) {
      // replay the host's order: entries are global; the ones naming two current factors of THIS product are taken in
      // sequence (a result keeps the id of its first operand)
      for (size_t h = 0; h < host_schedule.size(); ++h) {
        if (host_used[h]) continue;
        size_t bi = names.size(), bj = names.size();
        for (size_t i = 0; i < names.size(); ++i) {
          if (names[i] == host_schedule[h].first && bi == names.size()) bi = i;
          if (names[i] == host_schedule[h].second && bj == names.size()) bj = i;
        }
        if (bi == names.size() || bj == names.size() || bi == bj) continue;  // belongs to another product
        host_used[h] = 1;
        contract_pair(bi, bj);
      }
    }
    while (vals.size() > 1) {
      size_t bi = 0, bj = 1;
      std::vector<Cand> cands;
      for (size_t i = 0; i < vals.size(); ++i)
        for (size_t j = i + 1; j < vals.size(); ++j) {
          const Structure &A = values[vals[i]].st, &B = values[vals[j]].st;
          uint32_t nc = n_shared_indices(A, B);
          if (nc == 0) continue;
          spn_contract_plan p = plan_contract(A, B);
          bool free_pair = !values[vals[i]].batched && !values[vals[j]].batched;
          cands.push_back({i, j, {(double)p.size_out, free_pair ? 0.0 : 1.0, (double)p.size_out * (double)p.n_contracted}});
        }
      if (!cands.empty()) {
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) { return x.cost < y.cost; });
        size_t pick = 0;
        if (trial > 0) {
          // candidates whose result is at most 8x the smallest one, plus every point-independent pair (free per
          // point whatever its size); earlier (cheaper) ones more likely
          std::stable_partition(cands.begin(), cands.end(), [&](const Cand& c) {
            return c.cost[0] <= 8.0 * cands[0].cost[0] || (c.cost[1] == 0.0 && c.cost[0] <= 4096.0);
          });
          const double lim = cands[0].cost[0];
          size_t n_ok = 0;
          while (n_ok < cands.size() && (cands[n_ok].cost[0] <= 8.0 * lim || (cands[n_ok].cost[1] == 0.0 && cands[n_ok].cost[0] <= 4096.0))) ++n_ok;
          if (n_ok == 0) n_ok = 1;
          double r = rnd();
          pick = (size_t)(r * r * (double)n_ok);
          if (pick >= n_ok) pick = n_ok - 1;
        }
        bi = cands[pick].i;
        bj = cands[pick].j;
      } else {
        // outer product of the two smallest factors
        std::vector<size_t> idx(vals.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) {
          return values[vals[x]].st.size() < values[vals[y]].st.size();
        });
        bi = std::min(idx[0], idx[1]);
        bj = std::max(idx[0], idx[1]);
      }
      contract_pair(bi, bj);
    }
    return vals[0];
  }

  void flatten(int id, int op, std::vector<int>& out) {  // merge_ops, graph.rs:1002-1065
    for (int c : nodes[id].children) {
      if (nodes[c].kind == NODE_OP && nodes[c].op == op)
        flatten(c, op, out);
      else
        out.push_back(c);
    }
  }

  int build(int id, std::map<int, int>& memo) {
    auto it = memo.find(id);
    if (it != memo.end()) return it->second;
    int v = -1;
    if (nodes[(size_t)id].kind == LEAF_FUNCTION) {
      // make sure the leaves it reads exist as values (they may not appear anywhere else in the graph)
      std::vector<int> deps;
      for (auto& pr : nodes[(size_t)id].fn_params) deps.push_back(pr.first);
      for (int d : deps) build(d, memo);
      v = new_value(nodes[(size_t)id].st, true, id);
    } else if (nodes[(size_t)id].kind != NODE_OP) {
      v = new_value(nodes[(size_t)id].st, nodes[(size_t)id].kind == LEAF_BATCHED, id);
      values[(size_t)v].scalar_kind = nodes[(size_t)id].is_scalar_leaf;
    } else {
      // copy what we need: recursion may append nodes and move the vector
      struct {
        int op, power;
        std::vector<int> children;
      } n{nodes[(size_t)id].op, nodes[(size_t)id].power, nodes[(size_t)id].children};
      switch (n.op) {
        case SPN_OP_PRODUCT: {
          std::vector<int> kids, vals;
          flatten(id, SPN_OP_PRODUCT, kids);
          if (kids.empty()) throw Error(SPN_ERR_INVALID, "empty product");
          bool all_scalars = true;  // ContractScalars: a product of Scalars is a Scalar (contract.rs:73-208)
          for (int c : kids) {
            vals.push_back(build(c, memo));
            all_scalars = all_scalars && values[(size_t)vals.back()].scalar_kind;
          }
          v = schedule_product(vals, kids);
          if (all_scalars && kids.size() > 1) values[(size_t)v].scalar_kind = true;
          break;
        }
        case SPN_OP_SUM: {  // network.rs:1342-1467
          std::vector<int> kids;
          flatten(id, SPN_OP_SUM, kids);
          if (kids.empty()) throw Error(SPN_ERR_INVALID, "empty sum");
          v = build(kids[0], memo);
          for (size_t k = 1; k < kids.size(); ++k) {
            int w = build(kids[k], memo);
            if (!same_structure(values[v].st, values[w].st))
              throw Error(SPN_ERR_STRUCTURE, values[v].st.order() == 0 || values[w].st.order() == 0
                                                 ? "sum of a scalar and a tensor"
                                                 : "sum of tensors of different structure");
            v = emit(ST_ADD, v, w, values[v].st);
          }
          // scalar mode of Sum (network.rs:1342-1467): the accumulated value is pushed as a Scalar
          if (kids.size() > 1 && values[(size_t)v].st.order() == 0) values[(size_t)v].scalar_kind = true;
          break;
        }
        case SPN_OP_NEG:
          if (n.children.size() != 1) throw Error(SPN_ERR_INVALID, "neg takes one child");
          v = build(n.children[0], memo);
          {
            const bool sk = values[(size_t)v].scalar_kind;
            v = emit(ST_NEG, v, -1, values[v].st);
            values[(size_t)v].scalar_kind = sk;
          }
          break;
        case SPN_OP_CONJ:  // FUN_LIB conj, spenso-hep-lib/src/lib.rs:511-520
          if (n.children.size() != 1) throw Error(SPN_ERR_INVALID, "conj takes one child");
          v = build(n.children[0], memo);
          {
            const bool sk = values[(size_t)v].scalar_kind;
            v = emit(ST_CONJ, v, -1, values[v].st);
            values[(size_t)v].scalar_kind = sk;
          }
          break;
        case SPN_OP_POWER: {  // network.rs:1526-1703, restated literally (including its odd-power behaviour)
          if (n.children.size() != 1) throw Error(SPN_ERR_INVALID, "power takes one child");
          int c = build(n.children[0], memo);
          const int pw = n.power, np = pw < 0 ? -pw : pw;
          if (values[(size_t)c].scalar_kind) {  // NetworkLeaf::Scalar
            if (np == 0) {
              v = c;  // the reference returns the scalar itself for a zero exponent
              break;
            }
            v = c;
            for (int k = 1; k < np; ++k) v = emit_contract(v, c, n.children[0], n.children[0]);
            if (pw < 0) v = emit(ST_RECIP, v, -1, values[(size_t)v].st);
            values[(size_t)v].scalar_kind = true;
            break;
          }
          if (pw == 0) {
            v = const_scalar_value(make_double2(1.0, 0.0));
            break;
          }
          if (pw == 1) {
            v = c;
            break;
          }
          const int nm = n.children[0];
          int square = emit_contract(c, c, nm, nm);
          if (np % 2 == 1) {
            v = c;
            if (np != 1) {
              for (int k = 0; k < np / 2; ++k) square = emit_contract(square, square, nm, nm);
              v = emit_contract(square, c, nm, nm);
            }
            if (pw < 0) {
              if (values[(size_t)v].st.order() != 0) throw Error(SPN_ERR_STRUCTURE, "negative exponent on a non-scalar");
              v = emit(ST_RECIP, v, -1, values[(size_t)v].st);
              values[(size_t)v].scalar_kind = true;
            }
          } else {
            if (values[(size_t)square].st.order() != 0) throw Error(SPN_ERR_STRUCTURE, "even power of a tensor is not a scalar");
            v = square;
            for (int k = 1; k < np / 2; ++k) v = emit_contract(v, square, nm, nm);
            if (pw < 0) v = emit(ST_RECIP, v, -1, values[(size_t)v].st);
            values[(size_t)v].scalar_kind = true;
          }
          break;
        }
        default: throw Error(SPN_ERR_INVALID, "unknown op");
      }
    }
    memo[id] = v;
    return v;
  }
  int const_scalar_value(double2 s) {
    Node n;
    n.kind = LEAF_CONST;
    n.dense = {s};
    nodes.push_back(n);
    int v = new_value(Structure(), false, (int)nodes.size() - 1);
    values[(size_t)v].scalar_kind = true;
    return v;
  }

  // ---- fused: replay the schedule symbolically ---------------------------------------------------
  std::vector<CE> sym_leaf(const Value& v) {
    const Node& n = nodes[v.leaf];
    const uint64_t sz = n.st.size();
    if (sz > kMaxFusedElems) throw Error(SPN_ERR_INVALID, "tensor too large for the fused executor; use SPN_EXEC_STEPWISE");
    std::vector<CE> el(sz);
    if (n.kind == LEAF_CONST) {
      if (n.dense.size() != sz) throw Error(SPN_ERR_INVALID, "constant leaf without data");
      for (uint64_t f = 0; f < sz; ++f) el[f] = CE{rr_const(n.dense[f].x), rr_const(n.dense[f].y)};
    } else {
      // loads are created lazily at the first use (el holds placeholders resolved in sym_get)
      for (uint64_t f = 0; f < sz; ++f) el[f] = CE{RR{-2, 1.0}, RR{-2, 1.0}};
    }
    return el;
  }
  struct SymT {
    std::vector<CE> el;
    int input = -1;  // >= 0: unresolved components are loads of this input
    bool cplx = true;
    const std::vector<uint64_t>* flat_map = nullptr;
  };
  CE sym_get(SymT& t, uint64_t f) {
    CE& e = t.el[f];
    if (e.re.var == -2) e = prog.load(t.input, t.flat_map && !t.flat_map->empty() ? (*t.flat_map)[f] : f, t.cplx);
    return e;
  }

  // LEAF_FUNCTION: run each component's postfix program symbolically (device-side leaf filling — what the
  // reference does per point on the host with EvalTensor evaluators, tensors/parametric.rs:2266-2720)
  std::vector<CE> sym_function_leaf(const Node& n, std::vector<SymT>& sym) {
    std::vector<CE> el(n.st.size());
    std::vector<CE> stack;
    auto pop = [&]() {
      if (stack.empty()) throw Error(SPN_ERR_INVALID, "function leaf: stack underflow");
      CE e = stack.back();
      stack.pop_back();
      return e;
    };
    for (uint64_t c = 0; c < n.st.size(); ++c) {
      stack.clear();
      for (auto& oa : n.fn_code[c]) {
        switch (oa.first) {
          case SPN_EXPR_PARAM: {
            auto& pr = n.fn_params.at((size_t)oa.second);
            stack.push_back(sym_get(sym[(size_t)node_value.at(pr.first)], pr.second));
            break;
          }
          case SPN_EXPR_CONST: {
            double2 k = n.fn_consts.at((size_t)oa.second);
            stack.push_back(CE{rr_const(k.x), rr_const(k.y)});
            break;
          }
          case SPN_EXPR_ADD:
          case SPN_EXPR_SUB: {
            CE b = pop(), a = pop();
            const double sg = oa.first == SPN_EXPR_ADD ? 1.0 : -1.0;
            std::vector<Term> tr, ti;
            Program::push_rr(tr, a.re, 1.0);
            Program::push_rr(tr, b.re, sg);
            Program::push_rr(ti, a.im, 1.0);
            Program::push_rr(ti, b.im, sg);
            stack.push_back(CE{prog.sum(tr), prog.sum(ti)});
            break;
          }
          case SPN_EXPR_MUL: {
            CE b = pop(), a = pop();
            std::vector<Term> tr, ti;
            Program::push_cmul(tr, ti, a, b, 1.0);
            stack.push_back(CE{prog.sum(tr), prog.sum(ti)});
            break;
          }
          case SPN_EXPR_DIV: {
            CE b = pop(), a = pop();
            stack.push_back(prog.divide(a, b));
            break;
          }
          case SPN_EXPR_NEG: {
            CE a = pop();
            stack.push_back(CE{rr_neg(a.re), rr_neg(a.im)});
            break;
          }
          case SPN_EXPR_CONJ: {
            CE a = pop();
            stack.push_back(CE{a.re, rr_neg(a.im)});
            break;
          }
          case SPN_EXPR_SQRT: {
            CE a = pop();
            if (!a.im.is_zero()) throw Error(SPN_ERR_INVALID, "function leaf: sqrt needs a real argument (take re() first)");
            stack.push_back(CE{prog.real_sqrt(a.re), rr_const(0.0)});
            break;
          }
          case SPN_EXPR_RE: {
            CE a = pop();
            stack.push_back(CE{a.re, rr_const(0.0)});
            break;
          }
          case SPN_EXPR_IM: {
            CE a = pop();
            stack.push_back(CE{a.im, rr_const(0.0)});
            break;
          }
          default: throw Error(SPN_ERR_INVALID, "function leaf: unknown opcode");
        }
      }
      if (stack.size() != 1) throw Error(SPN_ERR_INVALID, "function leaf: a component program must leave exactly one value");
      el[c] = stack.back();
    }
    return el;
  }

  void compile_fused() {
    std::vector<SymT> sym(values.size());
    for (size_t v = 0; v < values.size(); ++v) {
      if (values[v].leaf < 0) continue;
      const Node& n = nodes[values[v].leaf];
      if (n.kind == LEAF_FUNCTION) {
        sym[v].el = sym_function_leaf(n, sym);
        continue;
      }
      sym[v].el = sym_leaf(values[v]);
      if (n.kind == LEAF_BATCHED) {
        sym[v].input = n.input;
        sym[v].cplx = n.dtype == SPN_C64;
        sym[v].flat_map = &n.flat_map;
      }
    }
    for (const Step& s : steps) {
      SymT out;
      const uint64_t sz = values[s.dst].st.size();
      if (sz > kMaxFusedElems) throw Error(SPN_ERR_INVALID, "intermediate too large for the fused executor; use SPN_EXEC_STEPWISE");
      out.el.resize(sz);
      SymT& A = sym[s.a];
      switch (s.op) {
        case ST_CONTRACT: {
          SymT& B = sym[s.b];
          spn_contract_plan p = plan_contract(values[s.a].st, values[s.b].st);
          KTables kt = build_k_tables(p, ctx && ctx->emulate_reference_pairing);
          std::vector<uint32_t> idx(p.order_out, 0);
          std::vector<Term> tr, ti;
          for (uint64_t o = 0; o < p.size_out; ++o) {
            uint64_t offa = 0, offb = 0;
            for (uint32_t r = 0; r < p.order_out; ++r) {
              offa += (uint64_t)idx[r] * p.out_stride_a[r];
              offb += (uint64_t)idx[r] * p.out_stride_b[r];
            }
            tr.clear();
            ti.clear();
            for (size_t k = 0; k < kt.off_a.size(); ++k) {
              // peek: skip structurally zero constants without touching (= loading) the partner
              const uint64_t fa = offa + (uint64_t)kt.off_a[k], fb = offb + (uint64_t)kt.off_b[k];
              const CE& pa = A.el[fa];
              const CE& pb = B.el[fb];
              if ((pa.re.is_zero() && pa.im.is_zero()) || (pb.re.is_zero() && pb.im.is_zero())) continue;
              CE a = sym_get(A, fa), b = sym_get(B, fb);
              Program::push_cmul(tr, ti, a, b, kt.sign[k] ? -1.0 : 1.0);
            }
            out.el[o] = CE{prog.sum(tr), prog.sum(ti)};
            for (uint32_t r = p.order_out; r-- > 0;) {
              if (++idx[r] < p.out_dim[r]) break;
              idx[r] = 0;
            }
          }
          break;
        }
        case ST_ADD: {
          SymT& B = sym[s.b];
          for (uint64_t f = 0; f < sz; ++f) {
            CE a = sym_get(A, f), b = sym_get(B, f);
            std::vector<Term> tr, ti;
            Program::push_rr(tr, a.re, 1.0);
            Program::push_rr(tr, b.re, 1.0);
            Program::push_rr(ti, a.im, 1.0);
            Program::push_rr(ti, b.im, 1.0);
            out.el[f] = CE{prog.sum(tr), prog.sum(ti)};
          }
          break;
        }
        case ST_NEG:
          for (uint64_t f = 0; f < sz; ++f) {
            CE a = sym_get(A, f);
            out.el[f] = CE{rr_neg(a.re), rr_neg(a.im)};
          }
          break;
        case ST_CONJ:
          for (uint64_t f = 0; f < sz; ++f) {
            CE a = sym_get(A, f);
            out.el[f] = CE{a.re, rr_neg(a.im)};
          }
          break;
        case ST_RECIP:
          for (uint64_t f = 0; f < sz; ++f) out.el[f] = prog.reciprocal(sym_get(A, f));
          break;
      }
      sym[s.dst] = std::move(out);
    }
    SymT& R = sym[result_value];
    result_el.resize(R.el.size());
    std::vector<int> roots;
    for (uint64_t f = 0; f < R.el.size(); ++f) {
      result_el[f] = sym_get(R, f);
      roots.push_back(result_el[f].re.var);
      roots.push_back(result_el[f].im.var);
    }
    stats = prog.eliminate_dead(roots);
    if (!getenv("SPN_NO_RESCALE")) {
      std::vector<RR*> outs;
      for (CE& e : result_el) {
        outs.push_back(&e.re);
        outs.push_back(&e.im);
      }
      for (int round = 0; round < 4 && prog.rescale_pass(outs) > 0; ++round) {
      }
      stats = prog.eliminate_dead(roots);
    }
  }

  // ---- fused: print the kernel -------------------------------------------------------------------
  // variant key: per input 'p' (point major, 32-byte aligned: 256-bit loads allowed) / 'q' (point
  // major, 16-byte aligned only) / 'c' (component major); then out: 'n' none, 'p', 'c'
  // (+ 'r' = real output); then 's' when the sum over points is wanted
  // How a kernel variant hides the latency of its input loads.  Compute-heavy networks (FP64 time per point
  // comparable to its HBM time: cfg5) hold many live registers and run at ~2 warps per scheduler, so the loads of
  // FUTURE points must be in flight while the current one is computed:
  //   RING  each thread cp.async-copies its next kRingDepth-1 points into a private shared-memory ring
  //         (conflict-free [slot][load][thread] layout; no block barrier: a thread only reads what it wrote)
  //   REGS  register double buffering of the next point (when the ring would not fit in shared memory)
  //   NONE  bandwidth-bound networks at high occupancy (cfg2): loads are placed lazily next to their first use
  // measured: L1::no_allocate on the once-read streaming loads is +1 % (cfg2 point-major), +3 % (cfg2mr), +7-8 %
  // (component-major cfg2, cfg3) over plain ld.global.nc; evict_first is equivalent
  static constexpr const char* kDefaultLoadHint = ".L1::no_allocate";
  //   GATHER sparse-touch networks on large point-major records (cfg3: 71 of 584 components of a 9 KB record): read
  //         thread = point, every 16-byte load of a warp lands in a different record, 9 KB apart — one sector per DRAM row.
  //         Instead a warp owns slabs of 32 points and gathers them point by point: lane L copies the needed units L,
  //         L + 32, ... of ONE record per cp.async instruction (the accesses of an instruction stay inside one record),
  //         into a padded [point][unit] slab; then lane = point computes from shared memory.
  //   TMA   (SPN_LOAD_MODE=tma, measurement only) every CTA moves the slab of its next kBlock points — one contiguous range
  //         per input — with the TMA engine: cp.async.bulk + mbarrier (UBLKCP / SYNCS in SASS), 2-3 stages; threads then
  //         read their point from shared memory.  A bulk copy is linear, so 64-byte points are read 4-way bank-conflicted,
  //         and the block barrier per slab is exposed at 1-2 CTAs per SM: slower than the direct 256-bit loads on every
  //         workload (profiles/README.md), which already stream at the HBM roofline.
  enum PrefetchMode { PF_NONE = 0, PF_REGS = 1, PF_RING = 2, PF_GATHER = 3, PF_TMA = 4 };
  static constexpr int kGatherBlock = 64;   // threads per CTA of a GATHER kernel: 2 warps, each with its own slab
  const int kRingDepth = (getenv("SPN_RING_DEPTH") && (atoi(getenv("SPN_RING_DEPTH")) == 2 || atoi(getenv("SPN_RING_DEPTH")) == 8)) ? atoi(getenv("SPN_RING_DEPTH")) : 4;  // power of two
  struct PrefetchPlan {
    int mode = PF_NONE;
    uint32_t n_loads = 0;
    size_t smem_bytes = 0;
  };
  // widest vector (in doubles) the REAL loads of input `in` may use in this variant: a 32-byte aligned point-major
  // tensor whose per-point size is a multiple of 4 (2) doubles keeps every aligned quad (pair) inside one sector
  int real_vec(const std::string& variant, int in) const {
    if (variant[(size_t)in] != 'p' || getenv("SPN_NO_WIDE_REAL_LOADS")) return 1;
    // measured (cfg2mr, 32-byte points): 256-bit loads of four reals run at 0.81 of HBM, 128-bit and 64-bit loads at
    // 0.865 — unlike the complex inputs of cfg2 (64-byte points), where 256-bit loads win; so pairs by default.  The
    // prefetch ring moves 16-byte units either way (cfg5r: +3.7 % over 8-byte cp.async).
    const int vmax = getenv("SPN_REAL_VEC_MAX") ? atoi(getenv("SPN_REAL_VEC_MAX")) : 2;
    const uint64_t sz = nodes[input_nodes[(size_t)in]].st.size();
    if (vmax >= 4 && sz % 4 == 0 && jit_has_256bit_loads()) return 4;
    return (vmax >= 2 && sz % 2 == 0) ? 2 : 1;
  }
  // one 16-byte (or 8-byte) slot of the shared-memory prefetch ring
  struct RingUnit {
    int kind;            // 0: complex component (16 B), 1: aligned pair of real components (16 B), 2: one real (8 B)
    int in;
    uint64_t flat;       // first component covered
    int dst0, dst1;      // value ids receiving .x / .y (kind 0: dst0 is the double2; -1: not live)
  };
  std::vector<RingUnit> ring_units(const std::string& variant) const {
    std::vector<RingUnit> u;
    std::map<std::pair<int, uint64_t>, size_t> pair_of;
    for (const Ins& i : prog.ins) {
      if (i.op == I_LOADC) {
        u.push_back({0, i.in, i.flat, i.dst, -1});
      } else if (i.op == I_LOADR) {
        if (real_vec(variant, i.in) >= 2) {
          const uint64_t base = i.flat & ~1ull;
          auto it = pair_of.find({i.in, base});
          if (it == pair_of.end()) {
            pair_of[{i.in, base}] = u.size();
            u.push_back({1, i.in, base, -1, -1});
            it = pair_of.find({i.in, base});
          }
          (i.flat & 1ull ? u[it->second].dst1 : u[it->second].dst0) = i.dst;
        } else {
          u.push_back({2, i.in, i.flat, i.dst, -1});
        }
      }
    }
    return u;
  }
  PrefetchPlan plan_prefetch(const std::string& variant) const {
    PrefetchPlan pf;
    const char out_mode = variant[n_inputs];
    const uint64_t size_out = result_el.size();
    const double t_fp = (double)stats.n_fp / (64.0 * 148.0 * 1.9e9);
    const double t_hbm = (double)(stats.n_load_bytes + (out_mode == 'n' ? 0 : 16 * size_out)) / 6.5e12;
    if (stats.n_load_bytes == 0 || getenv("SPN_NO_PREFETCH")) return pf;
    if (const char* lm = getenv("SPN_LOAD_MODE"); lm && std::string(lm) == "tma") {
      bool ok = n_inputs > 0;
      size_t slab = 0;
      for (uint32_t in = 0; in < n_inputs; ++in) {
        const Node& nd = nodes[input_nodes[(size_t)in]];
        const size_t rec = nd.st.size() * (nd.dtype == SPN_C64 ? 16 : 8);
        ok = ok && variant[in] == 'p' && rec % 16 == 0;
        slab += rec * (size_t)kBlock;
      }
      if (ok && 2 * slab + 64 <= 200 * 1024) {
        pf.mode = PF_TMA;
        pf.n_loads = 3 * slab + 64 <= 200 * 1024 ? 3u : 2u;   // stages
        pf.smem_bytes = (size_t)pf.n_loads * slab + 64;
        return pf;
      }
    }
    {
      // GATHER: all inputs point-major, only 16-byte units, at most half of the input bytes touched, records >= 1 KB
      bool all_p = n_inputs > 0;
      uint64_t total_bytes = 0, max_record = 0;
      for (uint32_t in = 0; in < n_inputs; ++in) {
        all_p = all_p && (variant[in] == 'p' || variant[in] == 'q');
        const Node& nd = nodes[input_nodes[(size_t)in]];
        const uint64_t rec = nd.st.size() * (nd.dtype == SPN_C64 ? 16 : 8);
        total_bytes += rec;
        max_record = std::max(max_record, rec);
      }
      if (all_p && max_record >= 1024 && 2 * stats.n_load_bytes <= total_bytes && getenv("SPN_GATHER")) {
        const std::vector<RingUnit> units = ring_units(variant);
        bool all16 = units.size() >= 16;
        for (auto& u : units) all16 = all16 && u.kind != 2;
        const size_t row = units.size() | 1u;
        const size_t bytes = (size_t)(kGatherBlock / 32) * 32 * row * 16;
        if (all16 && bytes <= 200 * 1024) {
          pf.mode = PF_GATHER;
          pf.n_loads = (uint32_t)units.size();
          pf.smem_bytes = bytes;
          return pf;
        }
      }
    }
    if (stats.n_load_bytes > 512 || t_fp <= 0.5 * t_hbm) return pf;
    pf.n_loads = (uint32_t)ring_units(variant).size();
    pf.smem_bytes = (size_t)kRingDepth * pf.n_loads * (size_t)kBlock * 16;
    const size_t ring_max = (size_t)(getenv("SPN_RING_MAX_KB") ? atoi(getenv("SPN_RING_MAX_KB")) : 200) * 1024;
    pf.mode = (pf.smem_bytes <= ring_max && !getenv("SPN_NO_RING")) ? PF_RING : PF_REGS;
    if (pf.mode != PF_RING) pf.smem_bytes = 0;
    return pf;
  }

  // ---- the Monte-Carlo sum inside the fused kernel ---------------------------------------------------------------
  // sum over points of result[c] for a variant that stores no per-point result.  Two things are exploited:
  //  * LINEARITY: the trailing part of the program that is linear with constant coefficients (the last contractions
  //    with gamma matrices / metrics are sign flips and additions) commutes with the sum over points.  Results are
  //    expanded through such instructions down to a BASIS of values that really depend on the point in a non-linear way;
  //    only the basis is accumulated per point, the linear map is applied once per thread after the loop.
  //  * a basis value that is the end of a multiply-add chain nobody else reads is accumulated LINK BY LINK —
  //    a = fma(x1, y1, a); a = fma(x2, y2, a); ... — instead of being formed and then added: one instruction less.
  // cfg5: 32 trailing additions and 32 accumulations per point become 32 fused chains: 339 -> 259 FP64 instructions.
  struct SumPlan {
    struct Basis {
      int var;      // value accumulated (-1: the constant 1, i.e. the number of points)
      bool fused;   // fed link by link
    };
    std::vector<Basis> basis;
    std::vector<std::map<size_t, double>> rows;   // [2*c + part]: accumulator index -> coefficient
    struct Link {
      size_t acc;
      double k;
    };
    std::map<size_t, Link> fused;    // instruction index -> accumulate k * (its product) into accumulator acc
    std::set<size_t> skipped;        // instructions of the linear tail (not emitted)
    uint64_t per_point_instr = 0;    // FP64 instructions per point this plan costs beyond prog (adds) minus what it removes
    uint64_t removed_instr = 0;
  };
  SumPlan plan_sum(const std::string& variant) const {
    SumPlan sp;
    const uint64_t size_out = result_el.size();
    const bool want_sum = variant.back() == 's';
    sp.rows.resize(2 * size_out);
    if (!want_sum) return sp;
    const bool free_form = variant[n_inputs] == 'n' && !getenv("SPN_NO_SUM_FUSION");
    std::vector<int> def((size_t)prog.n_vals, -1), uses((size_t)prog.n_vals, 0);
    for (size_t n = 0; n < prog.ins.size(); ++n) {
      const Ins& i = prog.ins[n];
      def[(size_t)i.dst] = (int)n;
      if (i.op == I_LOADC) def[(size_t)i.dst + 1] = (int)n;
    }
    // expandable(v): defined by k*x (+-) z and read only by results or by other expandable values
    std::vector<char> expandable((size_t)prog.n_vals, 0), pinned((size_t)prog.n_vals, 0);
    if (free_form) {
      for (size_t n = prog.ins.size(); n-- > 0;) {
        const Ins& i = prog.ins[n];
        if (i.op == I_LOADC || i.op == I_LOADR) continue;
        const bool linear = i.op == I_FMAK || i.op == I_MULK;
        const bool exp = linear && !pinned[(size_t)i.dst];
        expandable[(size_t)i.dst] = exp;
        if (!exp) {
          if (i.x >= 0) pinned[(size_t)i.x] = 1;
          if (i.y >= 0) pinned[(size_t)i.y] = 1;
          if (i.z >= 0) pinned[(size_t)i.z] = 1;
        }
      }
    }
    std::map<int, size_t> acc_of;   // basis value -> accumulator
    auto acc_index = [&](int var) {
      auto it = acc_of.find(var);
      if (it != acc_of.end()) return it->second;
      sp.basis.push_back({var, false});
      acc_of[var] = sp.basis.size() - 1;
      return sp.basis.size() - 1;
    };
    std::function<void(std::map<size_t, double>&, int, double)> expand = [&](std::map<size_t, double>& row, int var, double k) {
      if (k == 0.0) return;
      if (var >= 0 && expandable[(size_t)var]) {
        const size_t n = (size_t)def[(size_t)var];
        const Ins& i = prog.ins[n];
        sp.skipped.insert(n);
        expand(row, i.x, k * i.k);
        if (i.op == I_FMAK) expand(row, i.z, k * (double)i.sz);
        return;
      }
      row[acc_index(var)] += k;
    };
    for (uint64_t c = 0; c < size_out; ++c)
      for (int part = 0; part < 2; ++part) {
        const RR& r = part ? result_el[c].im : result_el[c].re;
        if (r.is_zero()) continue;
        if (r.var < 0)
          sp.rows[2 * c + (uint64_t)part][acc_index(-1)] += r.coef;   // a constant: coef * n_points
        else
          expand(sp.rows[2 * c + (uint64_t)part], r.var, r.coef);
      }
    // remaining instruction uses of every value (skipped instructions do not read anything any more)
    for (size_t n = 0; n < prog.ins.size(); ++n) {
      const Ins& i = prog.ins[n];
      if (i.op == I_LOADC || i.op == I_LOADR || sp.skipped.count(n)) continue;
      if (i.x >= 0) uses[(size_t)i.x]++;
      if (i.y >= 0) uses[(size_t)i.y]++;
      if (i.z >= 0) uses[(size_t)i.z]++;
    }
    // link-by-link accumulation of basis values that end an unshared multiply-add chain
    if (free_form)
      for (size_t j = 0; j < sp.basis.size(); ++j) {
        const int top = sp.basis[j].var;
        if (top < 0 || uses[(size_t)top] != 0) continue;
        std::vector<std::pair<size_t, double>> links;
        double mult = 1.0;
        int cur = top;
        bool ok = false;
        while (cur >= 0 && def[(size_t)cur] >= 0 && uses[(size_t)cur] == (cur == top ? 0 : 1)) {
          const Ins& d = prog.ins[(size_t)def[(size_t)cur]];
          if (d.op == I_FMA || d.op == I_FMAK) {
            links.emplace_back((size_t)def[(size_t)cur], mult * d.k);
            mult *= (double)d.sz;
            cur = d.z;
          } else if (d.op == I_MUL || d.op == I_MULK) {
            links.emplace_back((size_t)def[(size_t)cur], mult * d.k);
            ok = true;
            break;
          } else {
            break;
          }
        }
        if (!ok) continue;
        for (auto& l : links) sp.fused[l.first] = SumPlan::Link{j, l.second};
        sp.basis[j].fused = true;
      }
    for (auto& bs : sp.basis) sp.per_point_instr += bs.fused ? 0 : 1;
    for (size_t n : sp.skipped) sp.removed_instr += 1 + 0 * n;
    for (auto& bs : sp.basis) sp.removed_instr += bs.fused ? 1 : 0;   // the chain head: a multiply became a multiply-add
    return sp;
  }

  // Kernels that stream (no prefetch ring) and store a result per point are launched FLAT: CTAs of kFlatBlock threads, one
  // point per thread, as many CTAs as that takes.  Small hardware-scheduled CTAs drift out of step, so the loads of
  // one overlap the stores of another; a persistent grid of 256-thread CTAs moves in lockstep.  Measured on every such
  // workload (profiles/r02_net_grid.txt): cfg2 2.395 -> 2.460e10 evals/s, cfg2 component-major 2.475 -> 2.530, cfg2m
  // 1.908 -> 1.939, cfg3 8.01 -> 8.27e8 (component-major 6.25 -> 6.33e9).  Kernels that fold a Monte-Carlo sum or run
  // a prefetch ring keep the persistent 256-thread CTAs (see kBlock).  SPN_NET_PERSISTENT=1 / SPN_BLOCK=n for A/B runs.
  static constexpr int kFlatBlock = 64;
  bool variant_flat(const std::string& variant) const {
    return variant.back() != 's' && plan_prefetch(variant).mode == PF_NONE && !getenv("SPN_NET_PERSISTENT");
  }
  int variant_block(const std::string& variant) const {
    if (plan_prefetch(variant).mode == PF_GATHER) return kGatherBlock;
    return (variant_flat(variant) && !getenv("SPN_BLOCK")) ? kFlatBlock : kBlock;
  }

  std::string gen_source(const std::string& variant, const std::string& kname) const {
    const int kBlock = variant_block(variant);   // shadows the member: threads per CTA of THIS variant
    const uint64_t size_out = result_el.size();
    const char out_mode = variant[n_inputs];
    const bool out_real = variant.find('r', n_inputs) != std::string::npos;
    const bool want_sum = variant.back() == 's';
    std::ostringstream o;
    o << "// generated by spenso_b200 (net.cu): fused per-point network kernel, sm_100a\n";
    o << "// inputs: " << n_inputs << ", result components: " << size_out << ", fp64 instr/point: " << stats.n_fp
      << ", input bytes/point: " << stats.n_load_bytes << "\n";
    o << "typedef unsigned long long ull;\n";
    // streaming loads: every input byte is read exactly once, so the L1 policy is a free parameter (SPN_LD_HINT:
    // "" plain, "na" = L1::no_allocate, "ef" = L1::evict_first)
    const char* hint_env = getenv("SPN_LD_HINT");
    const std::string hint = !hint_env ? std::string(kDefaultLoadHint)
                             : std::string(hint_env) == "na" ? ".L1::no_allocate"
                             : std::string(hint_env) == "ef" ? ".L1::evict_first" : "";
    const std::string ldnc = "ld.global.nc" + hint;
    o << "__device__ __forceinline__ double2 ld16(const double* p) {\n  double2 v;\n"
         "  asm(\"" << ldnc << ".v2.f64 {%0,%1}, [%2];\" : \"=d\"(v.x), \"=d\"(v.y) : \"l\"(p));\n  return v;\n}\n";
    o << "__device__ __forceinline__ double ld8(const double* p) {\n  double v;\n"
         "  asm(\"ld.global.nc.f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p));\n  return v;\n}\n";   // 8-byte loads share their sector with later loads of the same thread: keep it in L1 (measured: no_allocate here costs 19 %)
    o << "__device__ __forceinline__ double2 ld16c(const double* p) {\n  double2 v;\n"
         "  asm(\"ld.global.nc.v2.f64 {%0,%1}, [%2];\" : \"=d\"(v.x), \"=d\"(v.y) : \"l\"(p));\n  return v;\n}\n";   // half-sector loads of a thread that reads the other half next: L1-cached
    o << "__device__ __forceinline__ double ld8s(const double* p) {\n  double v;\n"
         "  asm(\"" << ldnc << ".f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p));\n  return v;\n}\n";   // component-major: the warp consumes whole sectors
    o << "__device__ __forceinline__ void ld32(const double* p, double2& a, double2& b) {\n"
         "  asm(\"" << ldnc << ".v4.f64 {%0,%1,%2,%3}, [%4];\" : \"=d\"(a.x), \"=d\"(a.y), \"=d\"(b.x), \"=d\"(b.y) : \"l\"(p));\n}\n";
    o << "__device__ __forceinline__ void ld32r(const double* p, double& a, double& b, double& c, double& d) {\n"
         "  asm(\"" << ldnc << ".v4.f64 {%0,%1,%2,%3}, [%4];\" : \"=d\"(a), \"=d\"(b), \"=d\"(c), \"=d\"(d) : \"l\"(p));\n}\n";
    o << "__device__ __forceinline__ void cp_async16(void* s, const double* g) {\n"
         "  asm volatile(\"cp.async.ca.shared.global [%0], [%1], 16;\" :: \"r\"((unsigned)__cvta_generic_to_shared(s)), \"l\"(g) : \"memory\");\n}\n";
    o << "__device__ __forceinline__ void cp_async8(void* s, const double* g) {\n"
         "  asm volatile(\"cp.async.ca.shared.global [%0], [%1], 8;\" :: \"r\"((unsigned)__cvta_generic_to_shared(s)), \"l\"(g) : \"memory\");\n}\n";
    o << "extern \"C\" __global__ void __launch_bounds__(" << kBlock << (kMinBlocks ? ", " + std::to_string(kMinBlocks) : std::string()) << ") " << kname << "(";
    for (uint32_t k = 0; k < n_inputs; ++k) o << "const double* __restrict__ in" << k << ", ull cs" << k << ", ";
    o << "double* __restrict__ out, ull ocs, double2* __restrict__ block_sums, ull n_points) {\n";
    if (want_sum) {
      const SumPlan sp0 = plan_sum(variant);
      for (size_t j = 0; j < sp0.basis.size(); ++j) o << "  double acc" << j << " = 0.0;\n";
    }
    // value names
    std::vector<std::string> name((size_t)prog.n_vals);
    auto vname = [&](int v) -> const std::string& { return name[(size_t)v]; };
    auto signedv = [&](int v, int s) { return (s < 0 ? "-" : "") + vname(v); };
    auto scaled = [&](double k, int v) {
      if (k == 1.0) return vname(v);
      if (k == -1.0) return "-" + vname(v);
      return "(" + lit(k) + " * " + vname(v) + ")";
    };
    // 256-bit loads: pair (flat, flat+1) of a point-major complex input when flat is even, the
    // tensor has an even number of components (=> 32-byte aligned) and both are live
    std::map<std::pair<int, uint64_t>, int> load_dst;
    for (auto& i : prog.ins)
      if (i.op == I_LOADC) load_dst[{i.in, i.flat}] = i.dst;
    std::map<std::pair<int, uint64_t>, int> rload_dst;
    for (auto& i : prog.ins)
      if (i.op == I_LOADR) rload_dst[{i.in, i.flat}] = i.dst;
    std::set<std::pair<int, uint64_t>> done_pair, done_rgroup;
    // one load statement: `pre` is the variable prefix ("l" current point, "n" prefetched next point),
    // `decl` says whether the statement also declares its variables
    auto emit_load = [&](std::ostream& os, const Ins& i, const std::string& pv, const std::string& pre, bool decl) {
      const Node& n = nodes[input_nodes[(size_t)i.in]];
      const uint64_t sz = n.st.size();
      const char lay = variant[(size_t)i.in];
      if (i.op == I_LOADR && lay == 'p' && real_vec(variant, i.in) >= 2) {
        // sector-wide loads of real components: one 256-bit load per 32-byte sector holding >= 2 live components,
        // else one 128-bit load per live aligned pair (dead lanes land in scratch registers: same sector, no traffic)
        const int V = real_vec(variant, i.in);
        auto dst_of = [&](uint64_t f) { auto it = rload_dst.find({i.in, f}); return it == rload_dst.end() ? -1 : it->second; };
        const uint64_t b4 = i.flat & ~3ull, b2 = i.flat & ~1ull;
        int live4 = 0;
        for (uint64_t f = b4; f < b4 + 4; ++f) live4 += dst_of(f) >= 0;
        if (V == 4 && live4 >= 2) {
          if (done_rgroup.count({i.in, b4})) return;
          done_rgroup.insert({i.in, b4});
          std::string nm[4], scratch;
          for (int k = 0; k < 4; ++k) {
            const int d = dst_of(b4 + (uint64_t)k);
            if (d >= 0) {
              nm[k] = pre + std::to_string(d);
              if (decl) os << "double " << nm[k] << "; ";
            } else {
              nm[k] = "z" + std::to_string(k);
              scratch += (scratch.empty() ? "double " : ", ") + nm[k];
            }
          }
          os << "{ " << (scratch.empty() ? std::string() : scratch + "; ") << "ld32r(in" << i.in << " + " << pv << " * " << sz << "ull + "
             << b4 << "ull, " << nm[0] << ", " << nm[1] << ", " << nm[2] << ", " << nm[3] << "); }\n";
          return;
        }
        if (dst_of(b2) >= 0 && dst_of(b2 + 1) >= 0) {
          if (done_rgroup.count({i.in, b2})) return;
          done_rgroup.insert({i.in, b2});
          const std::string a = pre + std::to_string(dst_of(b2)), b = pre + std::to_string(dst_of(b2 + 1));
          if (decl) os << "double " << a << ", " << b << "; ";
          os << "{ const double2 t = ld16(in" << i.in << " + " << pv << " * " << sz << "ull + " << b2 << "ull); " << a << " = t.x; "
             << b << " = t.y; }\n";
          return;
        }
      }
      if (i.op == I_LOADR) {
        os << (decl ? "const double " : "") << pre << i.dst << " = " << (lay == 'c' ? "ld8s" : "ld8") << "(in" << i.in;
        if (lay != 'c')
          os << " + " << pv << " * " << sz << "ull + " << i.flat << "ull);\n";
        else
          os << " + " << i.flat << "ull * cs" << i.in << " + " << pv << ");\n";
        return;
      }
      const std::string ln = pre + std::to_string(i.dst);
      if (lay == 'c') {
        os << (decl ? "const double2 " : "") << ln << " = ld16(in" << i.in << " + (" << i.flat << "ull * cs" << i.in << " + "
           << pv << ") * 2);\n";
        return;
      }
      const bool pairable = sz % 2 == 0 && lay == 'p' && jit_has_256bit_loads();
      const uint64_t base = i.flat & ~1ull;
      if (pairable && load_dst.count({i.in, base}) && load_dst.count({i.in, base + 1})) {
        if (done_pair.count({i.in, base})) return;  // emitted together with its partner
        done_pair.insert({i.in, base});
        const std::string a = pre + std::to_string(load_dst[{i.in, base}]), b = pre + std::to_string(load_dst[{i.in, base + 1}]);
        if (decl) os << "double2 " << a << ", " << b << "; ";
        os << "ld32(in" << i.in << " + (" << pv << " * " << sz << "ull + " << base << "ull) * 2, " << a << ", " << b << ");\n";
      } else {
        // unpaired 16-byte load: streaming when its sector partner is dead, L1-cached when the partner is loaded separately
        const bool partner_live = load_dst.count({i.in, i.flat ^ 1ull}) != 0;
        os << (decl ? "const double2 " : "") << ln << " = " << (partner_live ? "ld16c" : "ld16") << "(in" << i.in << " + (" << pv
           << " * " << sz << "ull + " << i.flat << "ull) * 2);\n";
      }
    };
    const PrefetchPlan pf = plan_prefetch(variant);
    const bool prefetch = pf.mode != PF_NONE;
    // inputs read with warp-coalesced loads (see below): point-major, 32-byte aligned, exactly 4 complex components per
    // point, all of them live, direct-load kernels only
    std::vector<int> coal;
    if (!prefetch && jit_has_256bit_loads() && kBlock % 32 == 0 && !getenv("SPN_NO_COALESCE"))
      for (uint32_t in = 0; in < n_inputs; ++in) {
        const Node& n = nodes[input_nodes[(size_t)in]];
        if (variant[(size_t)in] != 'p' || n.st.size() != 4) continue;
        bool all = true;
        for (uint64_t f = 0; f < 4; ++f) all = all && load_dst.count({(int)in, f});
        if (all) coal.push_back((int)in);
      }
    for (const Ins& i : prog.ins) {
      if (i.op == I_LOADC) {
        name[(size_t)i.dst] = "l" + std::to_string(i.dst) + ".x";
        name[(size_t)i.dst + 1] = "l" + std::to_string(i.dst) + ".y";
      } else if (i.op == I_LOADR) {
        name[(size_t)i.dst] = "l" + std::to_string(i.dst);
      }
    }
    // source address (in doubles) of one load at point `pv`
    auto load_addr = [&](const Ins& i, const std::string& pv) {
      const Node& n = nodes[input_nodes[(size_t)i.in]];
      std::ostringstream a;
      const int w = i.op == I_LOADC ? 2 : 1;
      if (variant[(size_t)i.in] == 'c')
        a << "in" << i.in << " + (" << i.flat << "ull * cs" << i.in << " + " << pv << ") * " << w;
      else
        a << "in" << i.in << " + (" << pv << " * " << n.st.size() << "ull + " << i.flat << "ull) * " << w;
      return a.str();
    };
    if (pf.mode == PF_TMA) {
      const unsigned stages = pf.n_loads;
      std::vector<size_t> rec(n_inputs), off(n_inputs);
      size_t slab = 0;
      for (uint32_t in = 0; in < n_inputs; ++in) {
        const Node& nd = nodes[input_nodes[(size_t)in]];
        rec[in] = nd.st.size() * (nd.dtype == SPN_C64 ? 16 : 8);
        off[in] = slab;
        slab += rec[in] * (size_t)kBlock;
      }
      o << "  extern __shared__ __align__(128) unsigned char tsm[];   // [" << stages << " stages][" << slab << " B: one slab per input], then the mbarriers\n";
      o << "  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(tsm + " << stages * slab << "u);\n";
      o << "  const ull n_tiles = (n_points + " << kBlock - 1 << "ull) / " << kBlock << "ull;\n";
      o << "  if (threadIdx.x == 0) {\n    for (unsigned s = 0; s < " << stages << "u; ++s)\n"
           "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" :: \"r\"((unsigned)__cvta_generic_to_shared(bars + s)) : \"memory\");\n"
           "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n  }\n  __syncthreads();\n";
      o << "  auto issue = [&](ull tile, unsigned stage) {   // thread 0: arm the stage's barrier, launch one bulk copy per input\n";
      o << "    const ull t0 = tile * " << kBlock << "ull;\n";
      o << "    const unsigned np = (unsigned)(n_points - t0 < " << kBlock << "ull ? n_points - t0 : " << kBlock << "ull);\n";
      o << "    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + stage);\n";
      size_t per_point = 0;
      for (uint32_t in = 0; in < n_inputs; ++in) per_point += rec[in];
      o << "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n";
      o << "    asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\" :: \"r\"(bar), \"r\"(np * " << per_point << "u) : \"memory\");\n";
      for (uint32_t in = 0; in < n_inputs; ++in)
        o << "    asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\" :: "
             "\"r\"((unsigned)__cvta_generic_to_shared(tsm + stage * " << slab << "u + " << off[in] << "u)), \"l\"(reinterpret_cast<const unsigned char*>(in"
          << in << ") + t0 * " << rec[in] << "ull), \"r\"(np * " << rec[in] << "u), \"r\"(bar) : \"memory\");\n";
      o << "  };\n";
      o << "  if (threadIdx.x == 0)\n    for (unsigned d = 0; d + 1 < " << stages << "u; ++d)\n"
           "      if (blockIdx.x + (ull)d * gridDim.x < n_tiles) issue(blockIdx.x + (ull)d * gridDim.x, d);\n";
      o << "  unsigned phases = 0;\n";
      o << "  for (ull tile = blockIdx.x, it = 0; tile < n_tiles; tile += gridDim.x, ++it) {\n";
      o << "    const unsigned stage = (unsigned)(it % " << stages << "ull);\n";
      o << "    __syncthreads();   // every thread is done with the stage that is refilled now\n";
      o << "    if (threadIdx.x == 0 && tile + " << (stages - 1) << "ull * gridDim.x < n_tiles) issue(tile + " << (stages - 1)
        << "ull * gridDim.x, (unsigned)((it + " << (stages - 1) << "ull) % " << stages << "ull));\n";
      o << "    {\n      const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + stage), ph = (phases >> stage) & 1u;\n";
      o << "      unsigned done = 0;\n      while (!done)\n";
      o << "        asm volatile(\"{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\" : \"=r\"(done) : \"r\"(bar), \"r\"(ph) : \"memory\");\n";
      o << "      phases ^= 1u << stage;\n    }\n";
      o << "    const ull p = tile * " << kBlock << "ull + threadIdx.x;\n";
      o << "    if (p < n_points) {\n";
      for (const Ins& i : prog.ins) {
        if (i.op != I_LOADC && i.op != I_LOADR) continue;
        const size_t eb = i.op == I_LOADC ? 16 : 8;
        o << "    const " << (i.op == I_LOADC ? "double2" : "double") << " l" << i.dst << " = *reinterpret_cast<const "
          << (i.op == I_LOADC ? "double2" : "double") << "*>(tsm + stage * " << slab << "u + " << off[(size_t)i.in] << "u + threadIdx.x * "
          << rec[(size_t)i.in] << "u + " << i.flat * eb << "u);\n";
      }
    } else if (pf.mode == PF_GATHER) {
      std::vector<RingUnit> units = ring_units(variant);
      // ascending addresses inside a record: neighbouring lanes fetch neighbouring sectors
      std::stable_sort(units.begin(), units.end(), [](const RingUnit& x, const RingUnit& y) {
        return x.in != y.in ? x.in < y.in : x.flat < y.flat;
      });
      const size_t nu = units.size(), row = nu | 1u, per_lane = (nu + 31) / 32;
      o << "  extern __shared__ __align__(16) double2 ring[];   // [" << kBlock / 32 << " warps][32 points][" << row << " units]\n";
      o << "  const unsigned lane = threadIdx.x & 31u;\n";
      o << "  double2* const slab = ring + (threadIdx.x >> 5) * " << 32 * row << "u;\n";
      // the units this lane gathers of every record: (input, byte offset inside the record)
      o << "  const unsigned char* ub[" << per_lane << "]; ull us[" << per_lane << "];\n";
      for (size_t j = 0; j < per_lane; ++j) {
        o << "  {\n    const unsigned u = lane + " << 32 * j << "u;\n";
        o << "    const unsigned char* b = nullptr; ull st = 0, off = 0;\n    switch (u) {\n";
        for (size_t e = 32 * j; e < std::min(nu, 32 * (j + 1)); ++e) {
          const Node& n = nodes[input_nodes[(size_t)units[e].in]];
          const uint64_t eb = n.dtype == SPN_C64 ? 16 : 8;
          o << "      case " << e << ": b = reinterpret_cast<const unsigned char*>(in" << units[e].in << "); st = " << n.st.size() * eb
            << "ull; off = " << units[e].flat * eb << "ull; break;\n";
        }
        o << "      default: break;\n    }\n    ub[" << j << "] = b ? b + off : nullptr; us[" << j << "] = st;\n  }\n";
      }
      o << "  const ull n_slabs = (n_points + 31ull) >> 5;\n";
      o << "  for (ull sl = (ull)blockIdx.x * " << kBlock / 32 << "ull + (threadIdx.x >> 5); sl < n_slabs; sl += (ull)gridDim.x * "
        << kBlock / 32 << "ull) {\n";
      o << "    const ull p0 = sl << 5;\n";
      o << "    const unsigned np = (unsigned)(n_points - p0 < 32ull ? n_points - p0 : 32ull);\n";
      o << "#pragma unroll 4\n    for (unsigned q = 0; q < np; ++q) {\n";
      for (size_t j = 0; j < per_lane; ++j)
        o << "      if (ub[" << j << "]) cp_async16(slab + q * " << row << "u + lane + " << 32 * j << "u, reinterpret_cast<const double*>(ub["
          << j << "] + (p0 + q) * us[" << j << "]));\n";
      o << "    }\n";
      o << "    asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n";
      o << "    asm volatile(\"cp.async.wait_group 0;\" ::: \"memory\");\n";
      o << "    __syncwarp();\n";
      o << "    const ull p = p0 + lane;\n";
      o << "    if (lane < np) {\n";
      o << "    const double2* const mine = slab + lane * " << row << "u;\n";
      for (size_t e = 0; e < nu; ++e) {
        const RingUnit& u = units[e];
        if (u.kind == 0) {
          o << "    const double2 l" << u.dst0 << " = mine[" << e << "];\n";
        } else {
          o << "    const double2 u" << e << " = mine[" << e << "];\n";
          if (u.dst0 >= 0) o << "    const double l" << u.dst0 << " = u" << e << ".x;\n";
          if (u.dst1 >= 0) o << "    const double l" << u.dst1 << " = u" << e << ".y;\n";
        }
      }
    } else if (pf.mode == PF_RING) {
      o << "  extern __shared__ __align__(16) double2 ring[];   // [" << kRingDepth << " slots][" << pf.n_loads << " loads]["
        << kBlock << " threads]\n";
      o << "  ull p = (ull)blockIdx.x * " << kBlock << " + threadIdx.x;\n  const ull stride = (ull)gridDim.x * " << kBlock << ";\n";
      // issue(pp, slot): this thread's loads of point pp -> ring slot
      o << "  auto issue = [&](ull pp, int slot) {\n";
      const std::vector<RingUnit> units = ring_units(variant);
      auto unit_addr = [&](const RingUnit& u, const std::string& pv) {
        const Node& n = nodes[input_nodes[(size_t)u.in]];
        std::ostringstream a;
        const int w = u.kind == 0 ? 2 : 1;
        if (variant[(size_t)u.in] == 'c')
          a << "in" << u.in << " + (" << u.flat << "ull * cs" << u.in << " + " << pv << ") * " << w;
        else
          a << "in" << u.in << " + (" << pv << " * " << n.st.size() << "ull + " << u.flat << "ull) * " << w;
        return a.str();
      };
      for (size_t e = 0; e < units.size(); ++e)
        o << "    cp_async" << (units[e].kind == 2 ? 8 : 16) << "(&ring[(slot * " << pf.n_loads << " + " << e << ") * " << kBlock
          << " + threadIdx.x], " << unit_addr(units[e], "pp") << ");\n";
      o << "  };\n";
      o << "  for (int d = 0; d < " << kRingDepth << "; ++d) {\n    if (p + d * stride < n_points) issue(p + d * stride, d);\n"
           "    asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n  }\n";
      o << "  for (int it = 0; p < n_points; p += stride, ++it) {\n    const int slot = it & " << (kRingDepth - 1) << ";\n";
      o << "    asm volatile(\"cp.async.wait_group " << (kRingDepth - 1) << ";\" ::: \"memory\");\n";
      for (size_t e = 0; e < units.size(); ++e) {
        const RingUnit& u = units[e];
        std::ostringstream slot_ref;
        slot_ref << "ring[(slot * " << pf.n_loads << " + " << e << ") * " << kBlock << " + threadIdx.x]";
        if (u.kind == 0) {
          o << "    const double2 l" << u.dst0 << " = " << slot_ref.str() << ";\n";
        } else if (u.kind == 2) {
          o << "    const double l" << u.dst0 << " = " << slot_ref.str() << ".x;\n";
        } else {
          o << "    const double2 u" << e << " = " << slot_ref.str() << ";\n";
          if (u.dst0 >= 0) o << "    const double l" << u.dst0 << " = u" << e << ".x;\n";
          if (u.dst1 >= 0) o << "    const double l" << u.dst1 << " = u" << e << ".y;\n";
        }
      }
    } else if (pf.mode == PF_REGS) {
      o << "  ull p = (ull)blockIdx.x * " << kBlock << " + threadIdx.x;\n  const ull stride = (ull)gridDim.x * " << kBlock << ";\n";
      for (const Ins& i : prog.ins) {
        if (i.op == I_LOADC) o << "  double2 l" << i.dst << " = make_double2(0.0, 0.0), n" << i.dst << " = make_double2(0.0, 0.0);\n";
        if (i.op == I_LOADR) o << "  double l" << i.dst << " = 0.0, n" << i.dst << " = 0.0;\n";
      }
      o << "  if (p < n_points) {\n";
      for (const Ins& i : prog.ins)
        if (i.op == I_LOADC || i.op == I_LOADR) {
          std::ostringstream t;
          emit_load(t, i, "p", "l", false);
          if (!t.str().empty()) o << "    " << t.str();
        }
      done_pair.clear();
      done_rgroup.clear();
      o << "  }\n  for (; p < n_points; p += stride) {\n    const ull pn = p + stride;\n    if (pn < n_points) {\n";
      for (const Ins& i : prog.ins)
        if (i.op == I_LOADC || i.op == I_LOADR) {
          std::ostringstream t;
          emit_load(t, i, "pn", "n", false);
          if (!t.str().empty()) o << "      " << t.str();
        }
      o << "    }\n";
    } else if (!coal.empty()) {
      // Warp-coalesced point-major loads.  A 64-byte tensor per point (4 complex components: spinors, momenta) read
      // thread = point makes every 256-bit load instruction touch 32 half-used 64-byte segments.  Instead the warp
      // reads its 32 points' 2 KB slab with two fully contiguous 1 KB instructions; lane t then holds sector (t & 1) of
      // points t/2 (first instruction) and 16 + t/2 (second).  Even lanes evaluate point t/2, odd lanes point
      // 16 + t/2: each keeps one sector and swaps the other with its neighbour (one shfl.xor round of 4 doubles).
      o << "  const unsigned lane = threadIdx.x & 31u;\n";
      o << "  for (ull p0 = (ull)blockIdx.x * " << kBlock << " + threadIdx.x; p0 - lane < n_points; p0 += (ull)gridDim.x * "
        << kBlock << ") {\n";
      o << "    const ull pw = p0 - lane;\n    const bool full = pw + 32 <= n_points;   // warp-uniform\n";
      o << "    if (!full && p0 >= n_points) continue;\n";
      o << "    const ull p = full ? pw + (lane >> 1) + ((lane & 1u) << 4) : p0;\n";
      for (int in : coal) {
        const int d0 = load_dst[{in, 0}], d1 = load_dst[{in, 1}], d2 = load_dst[{in, 2}], d3 = load_dst[{in, 3}];
        o << "    double2 l" << d0 << ", l" << d1 << ", l" << d2 << ", l" << d3 << ";\n";
        o << "    if (full) {\n";
        o << "      double2 a0, b0, a1, b1;\n";
        o << "      ld32(in" << in << " + (pw * 4ull + (ull)lane * 2ull) * 2, a0, b0);\n";
        o << "      ld32(in" << in << " + (pw * 4ull + 64ull + (ull)lane * 2ull) * 2, a1, b1);\n";
        o << "      const bool odd = lane & 1u;\n";
        o << "      double2 sa = odd ? a0 : a1, sb = odd ? b0 : b1;   // the sector my neighbour evaluates\n";
        o << "      sa.x = __shfl_xor_sync(0xffffffffu, sa.x, 1); sa.y = __shfl_xor_sync(0xffffffffu, sa.y, 1);\n";
        o << "      sb.x = __shfl_xor_sync(0xffffffffu, sb.x, 1); sb.y = __shfl_xor_sync(0xffffffffu, sb.y, 1);\n";
        o << "      l" << d0 << " = odd ? sa : a0; l" << d1 << " = odd ? sb : b0;\n";
        o << "      l" << d2 << " = odd ? a1 : sa; l" << d3 << " = odd ? b1 : sb;\n";
        o << "    } else {\n";
        o << "      ld32(in" << in << " + (p * 4ull + 0ull) * 2, l" << d0 << ", l" << d1 << ");\n";
        o << "      ld32(in" << in << " + (p * 4ull + 2ull) * 2, l" << d2 << ", l" << d3 << ");\n";
        o << "    }\n";
      }
    } else {
      o << "  for (ull p = (ull)blockIdx.x * " << kBlock << " + threadIdx.x; p < n_points; p += (ull)gridDim.x * " << kBlock
        << ") {\n";
    }
    const SumPlan sp = plan_sum(variant);
    size_t ins_index = 0;
    for (const Ins& i : prog.ins) {
      const size_t this_index = ins_index++;
      if (sp.skipped.count(this_index)) continue;   // part of the linear tail: applied once, after the loop
      if (auto fl = sp.fused.find(this_index); fl != sp.fused.end()) {
        const std::string acc = "acc" + std::to_string(fl->second.acc);
        const double k = fl->second.k;
        if (i.op == I_MUL || i.op == I_FMA)
          o << "    " << acc << " = fma(" << scaled(k, i.x) << ", " << vname(i.y) << ", " << acc << ");\n";
        else if (k == 1.0)
          o << "    " << acc << " += " << vname(i.x) << ";\n";
        else if (k == -1.0)
          o << "    " << acc << " -= " << vname(i.x) << ";\n";
        else
          o << "    " << acc << " = fma(" << lit(k) << ", " << vname(i.x) << ", " << acc << ");\n";
        continue;
      }
      switch (i.op) {
        case I_LOADC:
        case I_LOADR: {
          if (prefetch) break;
          if (std::find(coal.begin(), coal.end(), i.in) != coal.end()) break;  // loaded at the top of the iteration
          std::ostringstream t;
          emit_load(t, i, "p", "l", true);
          if (!t.str().empty()) o << "    " << t.str();
          break;
        }
        case I_MUL:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          o << "    const double v" << i.dst << " = " << scaled(i.k, i.x) << " * " << vname(i.y) << ";\n";
          break;
        case I_FMA:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          o << "    const double v" << i.dst << " = fma(" << scaled(i.k, i.x) << ", " << vname(i.y) << ", "
            << signedv(i.z, i.sz) << ");\n";
          break;
        case I_MULK:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          o << "    const double v" << i.dst << " = " << lit(i.k) << " * " << vname(i.x) << ";\n";
          break;
        case I_FMAK:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          if (i.k == 1.0)
            o << "    const double v" << i.dst << " = " << signedv(i.z, i.sz) << " + " << vname(i.x) << ";\n";
          else if (i.k == -1.0)
            o << "    const double v" << i.dst << " = " << signedv(i.z, i.sz) << " - " << vname(i.x) << ";\n";
          else
            o << "    const double v" << i.dst << " = fma(" << lit(i.k) << ", " << vname(i.x) << ", "
              << signedv(i.z, i.sz) << ");\n";
          break;
        case I_ADDK:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          o << "    const double v" << i.dst << " = " << signedv(i.z, i.sz) << " + " << lit(i.kc) << ";\n";
          break;
        case I_SQRT:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          o << "    const double v" << i.dst << " = sqrt(" << scaled(i.k, i.x) << ");\n";
          break;
        case I_DIV:
          name[(size_t)i.dst] = "v" + std::to_string(i.dst);
          if (i.x < 0)
            o << "    const double v" << i.dst << " = " << lit(i.k) << " / " << vname(i.y) << ";\n";
          else
            o << "    const double v" << i.dst << " = " << scaled(i.k, i.x) << " / " << vname(i.y) << ";\n";
          break;
      }
    }
    auto rr_expr = [&](const RR& r) -> std::string {
      if (r.is_const()) return lit(r.coef);
      return scaled(r.coef, r.var);
    };
    if (want_sum)   // the accumulators of the Monte-Carlo sum that are not fed link by link
      for (size_t j = 0; j < sp.basis.size(); ++j)
        if (!sp.basis[j].fused) o << "    acc" << j << " += " << (sp.basis[j].var < 0 ? lit(1.0) : vname(sp.basis[j].var)) << ";\n";
    for (uint64_t c = 0; c < size_out; ++c) {
      const std::string re = rr_expr(result_el[c].re), im = rr_expr(result_el[c].im);
      if (out_mode == 'n') continue;
      std::string idx = out_mode == 'p' ? "(p * " + std::to_string(size_out) + "ull + " + std::to_string(c) + "ull)"
                                        : "(" + std::to_string(c) + "ull * ocs + p)";
      if (out_real)
        o << "    out[" << idx << "] = " << re << ";\n";
      else
        o << "    reinterpret_cast<double2*>(out)[" << idx << "] = make_double2(" << re << ", " << im << ");\n";
    }
    if (pf.mode == PF_REGS)
      for (const Ins& i : prog.ins)
        if (i.op == I_LOADC || i.op == I_LOADR) o << "    l" << i.dst << " = n" << i.dst << ";\n";
    if (pf.mode == PF_GATHER) o << "    }\n    __syncwarp();   // the slab is re-used by the next iteration\n";
    if (pf.mode == PF_TMA) o << "    }\n";
    if (pf.mode == PF_RING)  // refill the slot just consumed with the point kRingDepth iterations ahead
      o << "    if (p + " << kRingDepth << " * stride < n_points) issue(p + " << kRingDepth << " * stride, slot);\n"
           "    asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n";
    o << "  }\n";
    if (want_sum) {
      // the linear tail: every result component is a fixed linear combination of the accumulators
      for (uint64_t c = 0; c < size_out; ++c)
        for (int part = 0; part < 2; ++part) {
          o << "  const double " << (part ? "si" : "sr") << c << " = ";
          const auto& row = sp.rows[2 * c + (uint64_t)part];
          if (row.empty()) o << "0.0";
          bool first = true;
          for (auto& jk : row) {
            if (!first) o << " + ";
            first = false;
            if (jk.second == 1.0)
              o << "acc" << jk.first;
            else
              o << lit(jk.second) << " * acc" << jk.first;
          }
          o << ";\n";
        }
      // per-block partial of the Monte-Carlo sum: fixed shuffle tree + fixed order over warps
      o << "  __shared__ double2 ws[" << (kBlock / 32) << "];\n";
      for (uint64_t c = 0; c < size_out; ++c) {
        o << "  {\n    double a = sr" << c << ", b = si" << c << ";\n";
        o << "    for (int off = 16; off > 0; off >>= 1) { a += __shfl_down_sync(0xffffffffu, a, off); b += "
             "__shfl_down_sync(0xffffffffu, b, off); }\n";
        o << "    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = make_double2(a, b);\n    __syncthreads();\n";
        o << "    if (threadIdx.x == 0) {\n      double2 t = ws[0];\n      for (int w = 1; w < " << (kBlock / 32)
          << "; ++w) { t.x += ws[w].x; t.y += ws[w].y; }\n      block_sums[(ull)blockIdx.x * " << size_out << "ull + " << c
          << "ull] = t;\n    }\n    __syncthreads();\n  }\n";
      }
    }
    o << "}\n";
    return o.str();
  }

  JitKernel& get_variant(const std::string& variant) {
    auto it = variants.find(variant);
    if (it != variants.end()) return *it->second.k;
    auto k = std::make_unique<JitKernel>();
    // the kernel name carries a hash of the body so ncu listings identify the network
    std::string probe = gen_source(variant, "K");
    char nm[64];
    std::snprintf(nm, sizeof nm, "spn_fused_net_%08x_%s", (unsigned)(fnv1a(probe) & 0xffffffffu), variant.c_str());
    k->name = nm;
    k->source = gen_source(variant, k->name);
    k->smem_bytes = plan_prefetch(variant).smem_bytes;
    k->block = variant_block(variant);
    k->flat = variant_flat(variant);
    // occupancy experiments: SPN_PAD_SMEM_KB adds unused dynamic shared memory (fewer CTAs per SM, grid follows)
    if (const char* e = getenv("SPN_PAD_SMEM_KB")) k->smem_bytes += (size_t)atoi(e) * 1024;
    jit_compile(*k);
    JitKernel& ref = *k;
    variants[variant].k = std::move(k);
    return ref;
  }
  std::string default_variant(bool sum) const {
    std::string v(n_inputs, 'p');
    v += 'p';
    if (sum) v += 's';
    return v;
  }
};

// =================================================================================================
namespace {

#define NET_TRY try {
#define NET_CATCH(neg)                                 \
  }                                                    \
  catch (const Error& e) {                             \
    spn::set_error(e.what());                          \
    return (neg) ? -e.code : e.code;                   \
  }                                                    \
  catch (const std::exception& e) {                    \
    spn::set_error(e.what());                          \
    return (neg) ? -SPN_ERR_OTHER : SPN_ERR_OTHER;     \
  }

int add_const_leaf(spn_net* net, const Structure& st, const std::map<uint64_t, double2>& entries, bool sparse,
                   const std::vector<double2>* dense) {
  Node n;
  n.kind = LEAF_CONST;
  n.st = st;
  n.sparse = sparse;
  if (dense) {
    n.dense = *dense;
  } else {
    n.dense.assign(st.size(), make_double2(0.0, 0.0));
    for (auto& kv : entries) {
      if (kv.first >= st.size()) throw Error(SPN_ERR_INVALID, "sparse flat index out of bounds");
      n.dense[kv.first] = kv.second;
    }
  }
  n.entries = entries;
  net->nodes.push_back(std::move(n));
  return (int)net->nodes.size() - 1;
}

void check_uncompiled(spn_net* net) {
  if (net->compiled) throw Error(SPN_ERR_INVALID, "network is already compiled");
}

bool v_is_alias(const spn_net* net, const Node& n) {
  return &n != &net->nodes[(size_t)net->input_nodes[(size_t)n.input]];
}

// stepwise execution: one pairwise kernel per step, intermediates in HBM
void run_stepwise(spn_net* net, const spn_tensor* const* inputs, spn_tensor* out, double* sum_out) {
  spn_ctx* ctx = net->ctx;
  std::vector<const spn_tensor*> val(net->values.size(), nullptr);
  std::vector<std::unique_ptr<spn_tensor>> owned(net->values.size());
  // last use of every value, to free intermediates as early as the reference's store never does
  std::vector<int> last_use(net->values.size(), -1);
  for (size_t s = 0; s < net->steps.size(); ++s) {
    last_use[(size_t)net->steps[s].a] = (int)s;
    if (net->steps[s].b >= 0) last_use[(size_t)net->steps[s].b] = (int)s;
  }
  for (size_t v = 0; v < net->values.size(); ++v) {
    const Value& V = net->values[v];
    if (V.leaf < 0) continue;
    const Node& n = net->nodes[(size_t)V.leaf];
    if (n.kind == LEAF_FUNCTION) {
      // materialise the device-filled leaf with a small fused kernel of its own (same inputs, root = the leaf)
      auto& aux = net->function_leaf_nets[V.leaf];
      if (!aux) {
        aux = std::make_unique<spn_net>();
        aux->ctx = ctx;
        // same node ids as in `net`: every batched leaf (and re-indexed view) plus the function leaf; the other
        // slots stay empty placeholders that nothing references
        int max_id = V.leaf;
        for (int id : net->input_nodes_all) max_id = std::max(max_id, id);
        aux->nodes.resize((size_t)max_id + 1);
        for (int id : net->input_nodes_all) aux->nodes[(size_t)id] = net->nodes[(size_t)id];
        aux->nodes[(size_t)V.leaf] = n;
        aux->n_inputs = net->n_inputs;
        aux->input_nodes = net->input_nodes;
        aux->root = V.leaf;
        int rc = spn_net_compile(aux.get(), SPN_EXEC_FUSED);
        if (rc != SPN_OK) throw Error(rc, spn_last_error());
      }
      uint64_t np = net->n_inputs ? inputs[0]->n_points : 1;
      owned[v].reset(new_batched(ctx, n.st, np, SPN_C64, SPN_POINT_MAJOR, nullptr));
      int rc = spn_net_execute(aux.get(), inputs, net->n_inputs, owned[v].get(), nullptr, 0, 0);
      if (rc != SPN_OK) throw Error(rc, spn_last_error());
      val[v] = owned[v].get();
      continue;
    }
    if (n.kind == LEAF_BATCHED) {
      if (v_is_alias(net, n)) {
        if (!n.flat_map.empty())
          throw Error(SPN_ERR_INVALID, "stepwise execution: re-indexed leaf with a different canonical order");
        const spn_tensor* t = inputs[n.input];
        owned[v].reset(new_batched(ctx, n.st, t->n_points, t->dtype, t->layout, t->dptr));
        val[v] = owned[v].get();
      } else {
        val[v] = inputs[n.input];
      }
    } else {
      if (!net->const_tensors[v])
        net->const_tensors[v].reset(n.sparse ? new_shared_sparse(ctx, n.st, n.entries)
                                             : new_shared_dense(ctx, n.st, n.dense.data()));
      val[v] = net->const_tensors[v].get();
    }
  }
  for (size_t s = 0; s < net->steps.size(); ++s) {
    const Step& st = net->steps[s];
    spn_tensor* r = nullptr;
    switch (st.op) {
      case ST_CONTRACT: tensor_contract(ctx, val[(size_t)st.a], val[(size_t)st.b], &r); break;
      case ST_ADD: tensor_elementwise(ctx, EW_ADD, val[(size_t)st.a], val[(size_t)st.b], make_double2(0, 0), &r); break;
      case ST_NEG: tensor_elementwise(ctx, EW_NEG, val[(size_t)st.a], val[(size_t)st.a], make_double2(0, 0), &r); break;
      case ST_CONJ: tensor_elementwise(ctx, EW_CONJ, val[(size_t)st.a], val[(size_t)st.a], make_double2(0, 0), &r); break;
      case ST_RECIP: tensor_elementwise(ctx, EW_RECIP, val[(size_t)st.a], val[(size_t)st.a], make_double2(0, 0), &r); break;
    }
    owned[(size_t)st.dst].reset(r);
    val[(size_t)st.dst] = r;
    for (int u : {st.a, st.b})
      if (u >= 0 && last_use[(size_t)u] == (int)s && u != net->result_value) owned[(size_t)u].reset();
  }
  const spn_tensor* res = val[(size_t)net->result_value];
  uint64_t n_points = out ? out->n_points : (res->batched() ? res->n_points : 1);
  if (out) {
    if (out->size() != res->size()) throw Error(SPN_ERR_STRUCTURE, "output tensor has the wrong size");
    // copy (broadcasting a point-independent result over the points)
    DevOperand src = res->operand();
    cuda_check(launch_elementwise(EW_COPY, src, src, out->output(), make_double2(0, 0), res->size(), n_points,
                                  out->layout == SPN_COMPONENT_MAJOR, ctx->stream),
               "copy result");
    ctx->launches++;
  }
  if (sum_out) {
    if (!res->batched()) throw Error(SPN_ERR_INVALID, "sum over points of a point-independent result");
    const int n_blocks = ctx->sm_count * 4;
    double2* partial = nullptr;
    cuda_check(cudaMallocAsync((void**)&partial, (size_t)n_blocks * res->size() * sizeof(double2) + 16, ctx->stream),
               "reduce scratch");
    cuda_check(launch_reduce_points(res->operand(), res->size(), res->n_points, partial, n_blocks,
                                    reinterpret_cast<double2*>(sum_out), ctx->stream),
               "reduce_points_kernel");
    ctx->launches += 2;
    cudaFreeAsync(partial, ctx->stream);
    // with a communicator attached sum_out is the sum over ALL ranks, whatever the execution mode
    if (net->comm) comm_exchange(net->comm, nullptr, 0, sum_out, (uint32_t)(2 * res->size()));
  }
  // intermediates are freed when `owned` goes out of scope; the stream is ordered, cudaFree syncs
}

}  // namespace

extern "C" {

int spn_net_create(spn_ctx* ctx, spn_net** out) {
  NET_TRY
  auto n = std::make_unique<spn_net>();
  n->ctx = ctx;  // may be NULL: plan-only network (compile + inspect, no execution)
  *out = n.release();
  return SPN_OK;
  NET_CATCH(false)
}
void spn_net_destroy(spn_net* net) { delete net; }

int spn_net_leaf_batched(spn_net* net, const spn_slot* slots, uint32_t n, int dtype) {
  NET_TRY
  check_uncompiled(net);
  Node nd;
  nd.kind = LEAF_BATCHED;
  nd.args.assign(slots, slots + n);
  nd.st = canonicalize(nd.args, &nd.perm);
  nd.input = (int)net->n_inputs++;
  nd.dtype = dtype;
  net->nodes.push_back(std::move(nd));
  net->input_nodes.push_back((int)net->nodes.size() - 1);
  net->input_nodes_all.push_back((int)net->nodes.size() - 1);
  return (int)net->nodes.size() - 1;
  NET_CATCH(true)
}
int spn_net_leaf_reindex(spn_net* net, int leaf, const spn_slot* slots, uint32_t n) {
  NET_TRY
  check_uncompiled(net);
  if (leaf < 0 || leaf >= (int)net->nodes.size() || net->nodes[(size_t)leaf].kind != LEAF_BATCHED)
    throw Error(SPN_ERR_INVALID, "reindex: not a batched leaf");
  const Node src = net->nodes[(size_t)leaf];
  if (n != src.args.size()) throw Error(SPN_ERR_STRUCTURE, "reindex: wrong number of slots");
  for (uint32_t i = 0; i < n; ++i)
    if (slots[i].rep_kind != src.args[i].rep_kind || slots[i].rep_id != src.args[i].rep_id ||
        slots[i].dim != src.args[i].dim)
      throw Error(SPN_ERR_STRUCTURE, "reindex: representation or dimension differs from the original leaf");
  Node nd;
  nd.kind = LEAF_BATCHED;
  nd.args.assign(slots, slots + n);
  nd.st = canonicalize(nd.args, &nd.perm);
  nd.input = src.input;
  nd.dtype = src.dtype;
  nd.alias_of = leaf;
  // canonical index tuple here -> argument order -> canonical tuple of the bound tensor
  const uint64_t size = nd.st.size();
  auto src_strides = src.st.strides();
  std::vector<uint64_t> stride_of_arg(n, 0);  // stride in the bound tensor of argument position a
  for (uint32_t c = 0; c < n; ++c) stride_of_arg[(size_t)src.perm[c]] = src_strides[c];
  nd.flat_map.resize(size);
  std::vector<uint32_t> idx(n, 0);
  bool identity = src.flat_map.empty();
  for (uint64_t f = 0; f < size; ++f) {
    uint64_t g = 0;
    for (uint32_t c = 0; c < n; ++c) g += (uint64_t)idx[c] * stride_of_arg[(size_t)nd.perm[c]];
    if (!src.flat_map.empty()) g = src.flat_map[g];
    nd.flat_map[f] = g;
    identity = identity && g == f;
    for (uint32_t c = n; c-- > 0;) {
      if (++idx[c] < nd.st.slots[c].dim) break;
      idx[c] = 0;
    }
  }
  if (identity) nd.flat_map.clear();
  net->nodes.push_back(std::move(nd));
  net->input_nodes_all.push_back((int)net->nodes.size() - 1);
  return (int)net->nodes.size() - 1;
  NET_CATCH(true)
}
int spn_net_leaf_function(spn_net* net, const spn_slot* slots, uint32_t n, const int32_t* params, uint32_t n_params,
                          const double* consts_reim, uint32_t n_consts, const int32_t* code, uint32_t n_code) {
  NET_TRY
  check_uncompiled(net);
  Node nd;
  nd.kind = LEAF_FUNCTION;
  nd.args.assign(slots, slots + n);
  nd.st = canonicalize(nd.args, &nd.perm);
  for (uint32_t k = 0; k < n_params; ++k) {
    const int leaf = params[2 * k];
    if (leaf < 0 || leaf >= (int)net->nodes.size() || net->nodes[(size_t)leaf].kind != LEAF_BATCHED)
      throw Error(SPN_ERR_INVALID, "function leaf: parameters must be components of batched leaves");
    if (params[2 * k + 1] < 0 || (uint64_t)params[2 * k + 1] >= net->nodes[(size_t)leaf].st.size())
      throw Error(SPN_ERR_INVALID, "function leaf: parameter component out of range");
    nd.fn_params.emplace_back(leaf, (uint64_t)params[2 * k + 1]);
  }
  for (uint32_t k = 0; k < n_consts; ++k) nd.fn_consts.push_back(make_double2(consts_reim[2 * k], consts_reim[2 * k + 1]));
  nd.fn_raw_code.assign(code, code + 2 * (size_t)n_code);
  // split into per-component programs (argument order), then place them in canonical order
  std::vector<std::vector<std::pair<int, int>>> comps(1);
  for (uint32_t k = 0; k < n_code; ++k) {
    const int op = code[2 * k], arg = code[2 * k + 1];
    if (op == SPN_EXPR_END) {
      comps.emplace_back();
      continue;
    }
    if (op == SPN_EXPR_PARAM && (arg < 0 || arg >= (int)n_params)) throw Error(SPN_ERR_INVALID, "function leaf: bad parameter index");
    if (op == SPN_EXPR_CONST && (arg < 0 || arg >= (int)n_consts)) throw Error(SPN_ERR_INVALID, "function leaf: bad constant index");
    comps.back().emplace_back(op, arg);
  }
  if (comps.back().empty()) comps.pop_back();
  const uint64_t size = nd.st.size();
  if (comps.size() != size) throw Error(SPN_ERR_DIMENSION, "function leaf: one program per component is required");
  std::vector<uint64_t> arg_stride(n, 1);
  for (size_t i = n; i-- > 1;) arg_stride[i - 1] = arg_stride[i] * nd.args[i].dim;
  nd.fn_code.resize(size);
  std::vector<uint32_t> idx(n, 0);
  for (uint64_t f = 0; f < size; ++f) {  // f canonical; g = the same index tuple in argument order
    uint64_t g = 0;
    for (uint32_t c = 0; c < n; ++c) g += (uint64_t)idx[c] * arg_stride[(size_t)nd.perm[c]];
    nd.fn_code[f] = comps[g];
    for (uint32_t c = n; c-- > 0;) {
      if (++idx[c] < nd.st.slots[c].dim) break;
      idx[c] = 0;
    }
  }
  net->nodes.push_back(std::move(nd));
  return (int)net->nodes.size() - 1;
  NET_CATCH(true)
}
int spn_net_leaf_shared(spn_net* net, const spn_tensor* shared) {
  NET_TRY
  check_uncompiled(net);
  if (shared->batched()) throw Error(SPN_ERR_INVALID, "leaf_shared needs a shared tensor");
  return add_const_leaf(net, shared->st, shared->entries, shared->sparse(), &shared->host_dense);
  NET_CATCH(true)
}
int spn_net_leaf_library(spn_net* net, int which, const spn_slot* slots, uint32_t n) {
  NET_TRY
  check_uncompiled(net);
  std::vector<spn_slot> args(slots, slots + n);
  LibraryData d = builtin_library(which, args);
  Structure st;
  auto ent = instantiate_library(d, args, &st);
  const int id = add_const_leaf(net, st, ent, true, nullptr);
  net->nodes.back().from_library = true;
  return id;
  NET_CATCH(true)
}
int spn_net_leaf_library_custom(spn_net* net, const spn_slot* slots, uint32_t n, const uint64_t* flat,
                                const double* reim, uint64_t nnz) {
  NET_TRY
  check_uncompiled(net);
  std::vector<spn_slot> args(slots, slots + n);
  LibraryData d;
  uint64_t size = 1;
  for (auto& s : args) size *= s.dim;
  for (uint64_t k = 0; k < nnz; ++k) {
    if (flat[k] >= size) throw Error(SPN_ERR_INVALID, "library entry out of bounds");
    d.nz.push_back({flat[k], make_double2(reim[2 * k], reim[2 * k + 1])});
  }
  Structure st;
  auto ent = instantiate_library(d, args, &st);
  const int id = add_const_leaf(net, st, ent, true, nullptr);
  net->nodes.back().from_library = true;
  return id;
  NET_CATCH(true)
}
int spn_net_leaf_scalar(spn_net* net, double re, double im) {
  NET_TRY
  check_uncompiled(net);
  std::vector<double2> d{make_double2(re, im)};
  int id = add_const_leaf(net, Structure(), {}, false, &d);
  net->nodes.back().is_scalar_leaf = true;
  return id;
  NET_CATCH(true)
}
int spn_net_op(spn_net* net, int op, const int32_t* children, uint32_t n_children, int32_t power) {
  NET_TRY
  check_uncompiled(net);
  if (op < SPN_OP_PRODUCT || op > SPN_OP_CONJ) throw Error(SPN_ERR_INVALID, "unknown op");
  Node nd;
  nd.kind = NODE_OP;
  nd.op = op;
  nd.power = power;
  for (uint32_t k = 0; k < n_children; ++k) {
    if (children[k] < 0 || children[k] >= (int)net->nodes.size()) throw Error(SPN_ERR_INVALID, "child id out of range");
    nd.children.push_back(children[k]);
  }
  net->nodes.push_back(std::move(nd));
  return (int)net->nodes.size() - 1;
  NET_CATCH(true)
}
int spn_net_set_schedule(spn_net* net, const int32_t* pairs, uint32_t n_pairs) {
  NET_TRY
  check_uncompiled(net);
  if (n_pairs && !pairs) throw Error(SPN_ERR_INVALID, "set_schedule: null pair list");
  net->host_schedule.clear();
  for (uint32_t k = 0; k < n_pairs; ++k) {
    const int a = pairs[2 * k], b = pairs[2 * k + 1];
    if (a < 0 || b < 0 || a >= (int)net->nodes.size() || b >= (int)net->nodes.size() || a == b)
      throw Error(SPN_ERR_INVALID, "set_schedule: a pair must name two different existing nodes");
    net->host_schedule.emplace_back(a, b);
  }
  net->has_host_schedule = n_pairs > 0;
  return SPN_OK;
  NET_CATCH(false)
}
int spn_net_set_root(spn_net* net, int node) {
  NET_TRY
  check_uncompiled(net);
  if (node < 0 || node >= (int)net->nodes.size()) throw Error(SPN_ERR_INVALID, "root id out of range");
  net->root = node;
  return SPN_OK;
  NET_CATCH(false)
}

// ---- plan interchange: the expression graph as JSON, so that a host that owns the real
// spenso Network (graph + store derive serde, network.rs:33-51) can hand it over without linking anything.
// Slots are [rep_kind, rep_id, dim, aind_kind, aind]; tensors carry canonical slots and canonical flat indices.
static void json_slots(std::ostringstream& o, const std::vector<spn_slot>& sl) {
  o << "[";
  for (size_t i = 0; i < sl.size(); ++i)
    o << (i ? "," : "") << "[" << (int)sl[i].rep_kind << "," << sl[i].rep_id << "," << sl[i].dim << "," << (int)sl[i].aind_kind
      << "," << sl[i].aind << "]";
  o << "]";
}
static std::vector<spn_slot> slots_of_json(const JVal& v) {
  std::vector<spn_slot> sl;
  for (auto& e : v.arr()) {
    auto& f = e.arr();
    if (f.size() != 5) throw Error(SPN_ERR_INVALID, "plan json: a slot is [rep_kind, rep_id, dim, aind_kind, aind]");
    sl.push_back(spn_slot{(uint8_t)f[0].i64(), (uint8_t)f[3].i64(), (int16_t)f[1].i64(), (uint32_t)f[2].u64(), (uint64_t)f[4].u64()});
  }
  return sl;
}
const char* spn_net_to_json(spn_net* net) {
  static const char* op_names[] = {"product", "sum", "neg", "power", "conj"};
  std::ostringstream o;
  o << "{\"version\":1,\"nodes\":[";
  // only the nodes the caller created (compile may append folded constants after them)
  size_t n_nodes = net->nodes.size();
  for (size_t k = 0; k < n_nodes; ++k) {
    const Node& n = net->nodes[k];
    o << (k ? ",\n" : "\n");
    if (n.kind == LEAF_BATCHED && n.alias_of < 0) {
      o << "{\"kind\":\"batched\",\"dtype\":\"" << (n.dtype == SPN_C64 ? "c64" : "f64") << "\",\"slots\":";
      json_slots(o, n.args);
      o << "}";
    } else if (n.kind == LEAF_BATCHED) {
      o << "{\"kind\":\"reindex\",\"leaf\":" << n.alias_of << ",\"slots\":";
      json_slots(o, n.args);
      o << "}";
    } else if (n.kind == LEAF_FUNCTION) {
      o << "{\"kind\":\"function\",\"slots\":";
      json_slots(o, n.args);
      o << ",\"params\":[";
      for (size_t k = 0; k < n.fn_params.size(); ++k) o << (k ? "," : "") << "[" << n.fn_params[k].first << "," << n.fn_params[k].second << "]";
      o << "],\"consts\":[";
      for (size_t k = 0; k < n.fn_consts.size(); ++k) o << (k ? "," : "") << "[" << jlit(n.fn_consts[k].x) << "," << jlit(n.fn_consts[k].y) << "]";
      o << "],\"code\":[";
      for (size_t k = 0; k < n.fn_raw_code.size(); ++k) o << (k ? "," : "") << n.fn_raw_code[k];
      o << "]}";
    } else if (n.kind == LEAF_CONST && n.is_scalar_leaf) {
      o << "{\"kind\":\"scalar\",\"re\":" << jlit(n.dense[0].x) << ",\"im\":" << jlit(n.dense[0].y) << "}";
    } else if (n.kind == LEAF_CONST) {
      o << "{\"kind\":\"tensor\",\"sparse\":" << (n.sparse ? "true" : "false") << (n.from_library ? ",\"library\":true" : "")
        << ",\"slots\":";
      json_slots(o, n.st.slots);
      o << ",\"entries\":[";
      bool first = true;
      auto put = [&](uint64_t f, double2 v) {
        o << (first ? "" : ",") << "[" << f << "," << jlit(v.x) << "," << jlit(v.y) << "]";
        first = false;
      };
      if (n.sparse)
        for (auto& kv : n.entries) put(kv.first, kv.second);
      else
        for (size_t f = 0; f < n.dense.size(); ++f) put(f, n.dense[f]);
      o << "]}";
    } else {
      o << "{\"kind\":\"op\",\"op\":\"" << op_names[n.op - SPN_OP_PRODUCT] << "\",\"power\":" << n.power << ",\"children\":[";
      for (size_t c = 0; c < n.children.size(); ++c) o << (c ? "," : "") << n.children[c];
      o << "]}";
    }
  }
  o << "\n],\"root\":" << net->root;
  if (net->has_host_schedule) {   // the host's contraction order (spn_net_set_schedule)
    o << ",\"schedule\":[";
    for (size_t k = 0; k < net->host_schedule.size(); ++k)
      o << (k ? "," : "") << "[" << net->host_schedule[k].first << "," << net->host_schedule[k].second << "]";
    o << "]";
  }
  o << "}\n";
  net->json = o.str();
  return net->json.c_str();
}

int spn_net_from_json(spn_ctx* ctx, const char* text, spn_net** out) {
  NET_TRY
  JVal doc;
  try {
    doc = JParser(text).parse();
  } catch (const std::exception& e) {
    throw Error(SPN_ERR_INVALID, e.what());
  }
  try {
    if (doc.at("version").i64() != 1) throw Error(SPN_ERR_INVALID, "plan json: unsupported version");
    auto net = std::make_unique<spn_net>();
    net->ctx = ctx;
    for (auto& jn : doc.at("nodes").arr()) {
      const std::string& kind = jn.at("kind").str();
      int id;
      if (kind == "batched") {
        auto sl = slots_of_json(jn.at("slots"));
        const JVal* dt = jn.find("dtype");
        id = spn_net_leaf_batched(net.get(), sl.data(), (uint32_t)sl.size(), dt && dt->str() == "f64" ? SPN_F64 : SPN_C64);
      } else if (kind == "reindex") {
        auto sl = slots_of_json(jn.at("slots"));
        id = spn_net_leaf_reindex(net.get(), (int)jn.at("leaf").i64(), sl.data(), (uint32_t)sl.size());
      } else if (kind == "library") {
        auto sl = slots_of_json(jn.at("slots"));
        id = spn_net_leaf_library(net.get(), (int)jn.at("which").i64(), sl.data(), (uint32_t)sl.size());
      } else if (kind == "tensor") {
        auto sl = slots_of_json(jn.at("slots"));
        Structure st = structure_of_canonical(sl.data(), (uint32_t)sl.size());
        std::map<uint64_t, double2> ent;
        for (auto& e : jn.at("entries").arr()) {
          auto& f = e.arr();
          if (f.size() != 3) throw Error(SPN_ERR_INVALID, "plan json: an entry is [flat, re, im]");
          ent[f[0].u64()] = make_double2(f[1].num(), f[2].num());
        }
        const JVal* spj = jn.find("sparse");
        const bool sparse = spj && spj->kind == JVal::Bool && spj->b;
        if (!sparse && ent.size() != st.size()) throw Error(SPN_ERR_DIMENSION, "plan json: dense tensor needs every entry");
        id = add_const_leaf(net.get(), st, ent, sparse, nullptr);
        if (!sparse) net->nodes.back().entries.clear();
        const JVal* lj = jn.find("library");
        net->nodes.back().from_library = lj && lj->kind == JVal::Bool && lj->b;
      } else if (kind == "function") {
        auto sl = slots_of_json(jn.at("slots"));
        std::vector<int32_t> params, code;
        std::vector<double> consts;
        for (auto& e : jn.at("params").arr()) {
          params.push_back((int32_t)e.arr().at(0).i64());
          params.push_back((int32_t)e.arr().at(1).i64());
        }
        for (auto& e : jn.at("consts").arr()) {
          consts.push_back(e.arr().at(0).num());
          consts.push_back(e.arr().at(1).num());
        }
        for (auto& e : jn.at("code").arr()) code.push_back((int32_t)e.i64());
        id = spn_net_leaf_function(net.get(), sl.data(), (uint32_t)sl.size(), params.data(), (uint32_t)params.size() / 2,
                                   consts.data(), (uint32_t)consts.size() / 2, code.data(), (uint32_t)code.size() / 2);
      } else if (kind == "scalar") {
        id = spn_net_leaf_scalar(net.get(), jn.at("re").num(), jn.at("im").num());
      } else if (kind == "op") {
        const std::string& op = jn.at("op").str();
        int code = op == "product" ? SPN_OP_PRODUCT : op == "sum" ? SPN_OP_SUM : op == "neg" ? SPN_OP_NEG
                 : op == "power" ? SPN_OP_POWER : op == "conj" ? SPN_OP_CONJ : -1;
        if (code < 0) throw Error(SPN_ERR_INVALID, "plan json: unknown op \"" + op + "\"");
        std::vector<int32_t> ch;
        for (auto& c : jn.at("children").arr()) ch.push_back((int32_t)c.i64());
        const JVal* pw = jn.find("power");
        id = spn_net_op(net.get(), code, ch.data(), (uint32_t)ch.size(), pw ? (int32_t)pw->i64() : 1);
      } else {
        throw Error(SPN_ERR_INVALID, "plan json: unknown node kind \"" + kind + "\"");
      }
      if (id < 0) throw Error(-id, spn_last_error());
    }
    check_uncompiled(net.get());
    int rc = spn_net_set_root(net.get(), (int)doc.at("root").i64());
    if (rc) throw Error(rc, spn_last_error());
    if (const JVal* sj = doc.find("schedule")) {
      std::vector<int32_t> pairs;
      for (auto& e : sj->arr()) {
        if (e.arr().size() != 2) throw Error(SPN_ERR_INVALID, "plan json: a schedule entry is [n1, n2]");
        pairs.push_back((int32_t)e.arr()[0].i64());
        pairs.push_back((int32_t)e.arr()[1].i64());
      }
      rc = spn_net_set_schedule(net.get(), pairs.data(), (uint32_t)pairs.size() / 2);
      if (rc) throw Error(rc, spn_last_error());
    }
    *out = net.release();
  } catch (const Error&) {
    throw;
  } catch (const std::exception& e) {
    throw Error(SPN_ERR_INVALID, e.what());
  }
  return SPN_OK;
  NET_CATCH(false)
}

int spn_net_compile(spn_net* net, int exec_mode) {
  NET_TRY
  check_uncompiled(net);
  if (net->root < 0) throw Error(SPN_ERR_INVALID, "network has no root");
  if (exec_mode != SPN_EXEC_FUSED && exec_mode != SPN_EXEC_STEPWISE && exec_mode != SPN_EXEC_AUTO)
    throw Error(SPN_ERR_INVALID, "unknown exec mode");
  if (exec_mode == SPN_EXEC_AUTO) {
    // Try the fused compiler first (symbolic replay only: milliseconds).  Straight-line code stops paying off when
    // the network has large intermediates: beyond ~20k FP64 instructions per point the kernel spills and NVRTC takes
    // minutes, so such networks run one pairwise kernel per step instead.  Function leaves exist only in fused mode.
    bool has_function_leaf = false;
    for (auto& nd : net->nodes) has_function_leaf = has_function_leaf || nd.kind == LEAF_FUNCTION;
    exec_mode = SPN_EXEC_FUSED;
    if (!has_function_leaf) {
      const size_t n_nodes_before = net->nodes.size();
      try {
        net->nodes.resize(n_nodes_before);
        net->values.clear();
        net->steps.clear();
        net->schedule.clear();
        net->prog = Program();
        net->trial = 0;
        net->host_used.assign(net->host_schedule.size(), 0);
        std::map<int, int> memo;
        net->result_value = net->build(net->root, memo);
        net->node_value = memo;
        net->compile_fused();
        if (net->stats.n_fp > 20000) exec_mode = SPN_EXEC_STEPWISE;
      } catch (const Error& e) {
        if (e.code != SPN_ERR_INVALID) throw;
        exec_mode = SPN_EXEC_STEPWISE;  // "too large for the fused executor"
      }
      net->nodes.resize(n_nodes_before);
    }
  }
  net->exec_mode = exec_mode;
  const size_t n_nodes0 = net->nodes.size();
  auto build_with = [&](int trial) {
    net->nodes.resize(n_nodes0);  // drop constants appended by a previous trial
    net->values.clear();
    net->steps.clear();
    net->schedule.clear();
    net->prog = Program();
    net->result_el.clear();
    net->trial = trial;
    net->rng_state = 0x9e3779b97f4a7c15ull * (uint64_t)(trial + 1);
    net->host_used.assign(net->host_schedule.size(), 0);
    std::map<int, int> memo;
    net->result_value = net->build(net->root, memo);
    net->node_value = memo;
    if (exec_mode == SPN_EXEC_FUSED) net->compile_fused();
  };
  build_with(0);
  for (size_t h = 0; h < net->host_schedule.size(); ++h)
    if (!net->host_used[h])
      throw Error(SPN_ERR_INVALID, "set_schedule: pair " + std::to_string(h) + " (" + std::to_string(net->host_schedule[h].first) +
                                       ", " + std::to_string(net->host_schedule[h].second) +
                                       ") does not name two factors of one product at that point of the order");
  if (exec_mode == SPN_EXEC_FUSED && !net->has_host_schedule) {   // the host's order is law: no search
    // schedule search: the objective is the instruction count of the kernel that would be generated
    int trials = 256;
    if (const char* e = getenv("SPN_SCHEDULE_TRIALS")) trials = atoi(e);
    int best_trial = 0;
    auto key = [&]() { return std::make_pair(net->stats.n_fp, net->stats.n_load_bytes); };
    auto best = key();
    for (int t = 1; t <= trials && best.first > 0; ++t) {
      try {
        build_with(t);
      } catch (const Error&) {
        continue;  // e.g. an intermediate too large for the fused executor under this order
      }
      if (key() < best) {
        best = key();
        best_trial = t;
      }
    }
    build_with(best_trial);
    SPN_TRACE_LOG("network compile: %zu pairwise steps, schedule trial %d of %d, %llu fp64 instr/point, %llu input bytes/point",
                  net->steps.size(), best_trial, trials, (unsigned long long)net->stats.n_fp,
                  (unsigned long long)net->stats.n_load_bytes);
  }
  net->const_tensors.resize(net->values.size());
  if (exec_mode == SPN_EXEC_FUSED) {
    JitKernel& k = net->get_variant(net->default_variant(false));
    net->default_source = k.source;
  }
  net->compiled = true;
  return SPN_OK;
  NET_CATCH(false)
}

int spn_net_exec_mode(const spn_net* net) { return net->exec_mode; }
uint32_t spn_net_result_order(const spn_net* net) {
  return net->compiled ? net->values[(size_t)net->result_value].st.order() : 0;
}
int spn_net_result_slots(const spn_net* net, spn_slot* out) {
  if (!net->compiled) return SPN_ERR_INVALID;
  const Structure& st = net->values[(size_t)net->result_value].st;
  for (uint32_t i = 0; i < st.order(); ++i) out[i] = st.slots[i];
  return SPN_OK;
}
uint64_t spn_net_result_size(const spn_net* net) {
  return net->compiled ? net->values[(size_t)net->result_value].st.size() : 0;
}
uint32_t spn_net_n_inputs(const spn_net* net) { return net->n_inputs; }
int spn_net_schedule(const spn_net* net, int32_t* pairs, uint32_t cap) {
  uint32_t c = 0;
  for (auto& p : net->schedule) {
    if (c < cap) {
      pairs[2 * c] = p.first;
      pairs[2 * c + 1] = p.second;
    }
    ++c;
  }
  return (int)c;
}
const char* spn_net_cuda_source(const spn_net* net) { return net->default_source.c_str(); }
const char* spn_net_prepare_variant(spn_net* net, const char* variant) {
  try {
    if (!net->compiled || net->exec_mode != SPN_EXEC_FUSED) throw Error(SPN_ERR_INVALID, "not a compiled fused network");
    std::string v = variant;
    if (v.size() < net->n_inputs + 1) throw Error(SPN_ERR_INVALID, "variant key too short");
    return net->get_variant(v).source.c_str();
  } catch (const std::exception& e) {
    spn::set_error(e.what());
    return nullptr;
  }
}
uint64_t spn_net_macs_per_point(const spn_net* net) { return net->stats.n_fp; }
uint64_t spn_net_variant_fp64_instr(spn_net* net, const char* variant) {
  try {
    if (!net->compiled || net->exec_mode != SPN_EXEC_FUSED) throw Error(SPN_ERR_INVALID, "not a compiled fused network");
    const std::string v = variant;
    if (v.size() < net->n_inputs + 1) throw Error(SPN_ERR_INVALID, "variant key too short");
    net->get_variant(v);
    uint64_t n = net->stats.n_fp;
    if (v.back() == 's') {
      // the sum over points: one add per accumulator that is not fed link by link, minus the linear tail that moved
      // out of the loop and the chain heads that absorbed their accumulation
      const spn_net::SumPlan sp = net->plan_sum(v);
      uint64_t skipped_cost = 0;
      for (size_t k : sp.skipped) skipped_cost += 1 + 0 * k;
      uint64_t adds = 0;
      for (auto& b : sp.basis) adds += b.fused ? 0 : 1;
      n = n - skipped_cost + adds;
    }
    return n;
  } catch (const std::exception& e) {
    spn::set_error(e.what());
    return 0;
  }
}
uint64_t spn_net_input_bytes_per_point(const spn_net* net) { return net->stats.n_load_bytes; }

// Network::execute on HOST data: the reference's calling convention (tensors are host Vec<Complex<f64>>).
// host_inputs[i]: n_points tensors of input i laid end to end (point major, canonical order); host_out (optional):
// n_points result tensors; host_sum (optional): sum over points (2*result_size doubles).  The batch is cut into
// chunks that flow through a kPipe-deep pipeline on private streams — H2D of chunk k+1 overlaps the kernel of chunk
// k and the D2H of chunk k-1 (both copy engines busy) — so device memory use is 3 chunks, whatever n_points is
// (this is how 1e8-point Monte-Carlo batches run).  Host buffers should be pinned (spn_host_alloc) for the copies
// to be asynchronous.  Synchronous: returns when host_out / host_sum are complete.
static void execute_host_impl(spn_net* net, const void* const* host_inputs, uint32_t n_inputs, uint64_t n_points,
                              void* host_out, double* host_sum, uint64_t chunk_points) {
  if (!net->compiled) throw Error(SPN_ERR_INVALID, "network is not compiled");
  if (!net->ctx) throw Error(SPN_ERR_NO_DEVICE, "plan-only network (created without a context) cannot execute");
  if (n_inputs != net->n_inputs) throw Error(SPN_ERR_INVALID, "wrong number of input buffers");
  if (!host_out && !host_sum) throw Error(SPN_ERR_INVALID, "nothing to compute: no output and no sum requested");
  spn_ctx* ctx = net->ctx;
  if (net->exec_mode != SPN_EXEC_FUSED) {
    // Stepwise networks (intermediates too large for straight-line code): chunk by chunk on the context's stream —
    // upload, one pairwise kernel per schedule step, download.  No overlap between copies and kernels: such networks
    // spend their time in the kernels, and their intermediates bound the chunk size.
    const uint64_t size_res = net->values[(size_t)net->result_value].st.size();
    if (chunk_points == 0) chunk_points = 1u << 14;
    std::vector<double2> total(size_res, make_double2(0.0, 0.0)), part(size_res);
    for (uint64_t first = 0; first < n_points; first += chunk_points) {
      const uint64_t cnt = std::min(chunk_points, n_points - first);
      std::vector<std::unique_ptr<spn_tensor>> ins;
      std::vector<const spn_tensor*> in_ptrs;
      for (uint32_t i = 0; i < n_inputs; ++i) {
        const Node& nd = net->nodes[(size_t)net->input_nodes[i]];
        ins.emplace_back(new_batched(ctx, nd.st, cnt, nd.dtype, SPN_POINT_MAJOR, nullptr));
        const size_t bytes = nd.st.size() * (nd.dtype == SPN_C64 ? 16 : 8);
        cuda_check(cudaMemcpyAsync(ins.back()->dptr, (const char*)host_inputs[i] + first * bytes, cnt * bytes, cudaMemcpyHostToDevice,
                                   ctx->stream),
                   "H2D of a chunk");
        in_ptrs.push_back(ins.back().get());
      }
      std::unique_ptr<spn_tensor> out(new_batched(ctx, net->values[(size_t)net->result_value].st, cnt, SPN_C64, SPN_POINT_MAJOR, nullptr));
      double2* dsum = nullptr;
      if (host_sum) cuda_check(cudaMallocAsync((void**)&dsum, size_res * sizeof(double2) + 16, ctx->stream), "sum scratch");
      spn_comm* const comm = net->comm;
      net->comm = nullptr;   // chunk sums are local; a communicator (if any) sees the total once, below
      try {
        run_stepwise(net, in_ptrs.data(), out.get(), reinterpret_cast<double*>(dsum));
      } catch (...) {
        net->comm = comm;
        throw;
      }
      net->comm = comm;
      if (host_out)
        cuda_check(cudaMemcpyAsync((char*)host_out + first * size_res * 16, out->dptr, cnt * size_res * 16, cudaMemcpyDeviceToHost, ctx->stream),
                   "D2H of a chunk");
      if (host_sum) {
        cuda_check(cudaMemcpyAsync(part.data(), dsum, size_res * 16, cudaMemcpyDeviceToHost, ctx->stream), "D2H of a chunk sum");
        cuda_check(cudaStreamSynchronize(ctx->stream), "chunk");
        cudaFreeAsync(dsum, ctx->stream);
        for (uint64_t c = 0; c < size_res; ++c) {   // fixed chunk order: deterministic
          total[c].x += part[c].x;
          total[c].y += part[c].y;
        }
      }
      cuda_check(cudaStreamSynchronize(ctx->stream), "chunk");   // the staging tensors die here
    }
    if (host_sum) std::memcpy(host_sum, total.data(), size_res * 16);
    return;
  }
  const uint64_t size_out = net->values[(size_t)net->result_value].st.size();
  if (chunk_points == 0) chunk_points = 1u << 17;
  chunk_points = std::min<uint64_t>(chunk_points, std::max<uint64_t>(n_points, 1));
  const uint64_t n_chunks = (n_points + chunk_points - 1) / chunk_points;
  std::vector<size_t> in_bytes(n_inputs);
  for (uint32_t i = 0; i < n_inputs; ++i) {
    const Node& nd = net->nodes[(size_t)net->input_nodes[i]];
    in_bytes[i] = nd.st.size() * (nd.dtype == SPN_C64 ? 16 : 8);
  }
  std::string variant(n_inputs, 'p');
  variant += host_out ? 'p' : 'n';
  if (host_sum) variant += 's';
  JitKernel& k = net->get_variant(variant);
  const int kblock = k.block ? k.block : net->kBlock;
  jit_load(k, kblock);
  const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)k.blocks_per_sm;

  auto& P = net->pipe;
  if (!P.fin) {
    for (int s = 0; s < spn_net::kPipe; ++s) {
      cuda_check(cudaStreamCreateWithFlags(&P.stream[s], cudaStreamNonBlocking), "pipeline stream");
      cuda_check(cudaEventCreateWithFlags(&P.chunk_done[s], cudaEventDisableTiming), "pipeline event");
    }
    cuda_check(cudaStreamCreateWithFlags(&P.fin, cudaStreamNonBlocking), "pipeline stream");
    for (int q = 0; q < 2; ++q) cuda_check(cudaEventCreateWithFlags(&P.fin_done[q], cudaEventDisableTiming), "pipeline event");
  }
  if (P.chunk_points < chunk_points || P.partial_elems < cap * size_out || P.n_chunks_cap < n_chunks) {
    cuda_check(cudaDeviceSynchronize(), "pipeline resize");
    P.release();
    for (int s = 0; s < spn_net::kPipe; ++s) {
      for (uint32_t i = 0; i < n_inputs; ++i) {
        void* d = nullptr;
        cuda_check(cudaMalloc(&d, std::max<size_t>(chunk_points * in_bytes[i], 32)), "pipeline staging");
        P.in[s].push_back(d);
      }
      cuda_check(cudaMalloc(&P.out[s], std::max<size_t>(chunk_points * size_out * 16, 32)), "pipeline staging");
      cuda_check(cudaMalloc((void**)&P.partial[s], std::max<size_t>(cap * size_out, 1) * sizeof(double2)), "pipeline staging");
    }
    cuda_check(cudaMalloc((void**)&P.chunk_sums, 2 * std::max<uint64_t>(n_chunks, 1) * size_out * sizeof(double2)), "pipeline staging");
    for (int q = 0; q < 2; ++q)
      cuda_check(cudaMalloc((void**)&P.final_sum[q], std::max<uint64_t>(size_out, 1) * sizeof(double2)), "pipeline staging");
    P.chunk_points = chunk_points;
    P.partial_elems = cap * size_out;
    P.n_chunks_cap = n_chunks;
  }
  const int parity = (int)(P.calls & 1);
  ++P.calls;
  double2* chunk_sums = P.chunk_sums + (size_t)parity * P.n_chunks_cap * size_out;
  // the per-chunk sums of this parity were last read by the final stage two calls ago
  if (host_sum && P.fin_pending[parity])
    for (int s = 0; s < spn_net::kPipe; ++s) cuda_check(cudaStreamWaitEvent(P.stream[s], P.fin_done[parity], 0), "event");

  for (uint64_t c = 0; c < n_chunks; ++c) {
    const int s = (int)(c % spn_net::kPipe);
    cudaStream_t st = P.stream[s];
    const uint64_t first = c * chunk_points, cnt = std::min(chunk_points, n_points - first);
    std::vector<const double*> in_ptr(n_inputs);
    std::vector<unsigned long long> in_cs(n_inputs, chunk_points);
    for (uint32_t i = 0; i < n_inputs; ++i) {
      cuda_check(cudaMemcpyAsync(P.in[s][i], (const char*)host_inputs[i] + first * in_bytes[i], cnt * in_bytes[i],
                                 cudaMemcpyHostToDevice, st),
                 "H2D of a chunk");
      in_ptr[i] = static_cast<const double*>(P.in[s][i]);
    }
    double* out_ptr = host_out ? static_cast<double*>(P.out[s]) : nullptr;
    unsigned long long ocs = chunk_points, np = cnt;
    double2* partial = host_sum ? P.partial[s] : nullptr;
    const unsigned grid = (unsigned)std::min<uint64_t>((cnt + kblock - 1) / kblock, k.flat ? 0x7fffffffull : cap);
    std::vector<void*> args;
    for (uint32_t i = 0; i < n_inputs; ++i) {
      args.push_back(&in_ptr[i]);
      args.push_back(&in_cs[i]);
    }
    args.push_back(&out_ptr);
    args.push_back(&ocs);
    args.push_back(&partial);
    args.push_back(&np);
    cuda_check(cudaLaunchKernel((const void*)k.kernel, dim3(grid), dim3((unsigned)kblock), args.data(), k.smem_bytes, st),
               "launch fused network kernel");
    ctx->launches++;
    if (host_sum) {
      cuda_check(launch_reduce_partials(partial, (int)grid, size_out, chunk_sums + c * size_out, st), "reduce_partials_kernel");
      ctx->launches++;
    }
    if (host_out)
      cuda_check(cudaMemcpyAsync((char*)host_out + first * size_out * 16, P.out[s], cnt * size_out * 16,
                                 cudaMemcpyDeviceToHost, st),
                 "D2H of a chunk");
  }
  if (host_sum) {
    if (n_chunks == 0) {
      std::memset(host_sum, 0, size_out * 16);
    } else {
      // deterministic: one fixed-shape reduction over the per-chunk sums, after the last chunk of every stream
      for (int s = 0; s < spn_net::kPipe; ++s) {
        cuda_check(cudaEventRecord(P.chunk_done[s], P.stream[s]), "event");
        cuda_check(cudaStreamWaitEvent(P.fin, P.chunk_done[s], 0), "event");
      }
      cuda_check(launch_reduce_partials(chunk_sums, (int)n_chunks, size_out, P.final_sum[parity], P.fin), "reduce_partials_kernel");
      ctx->launches++;
      cuda_check(cudaMemcpyAsync(host_sum, P.final_sum[parity], size_out * 16, cudaMemcpyDeviceToHost, P.fin), "D2H of the sum");
      cuda_check(cudaEventRecord(P.fin_done[parity], P.fin), "event");
      P.fin_pending[parity] = true;
    }
  }
}

int spn_net_host_wait(spn_net* net) {
  NET_TRY
  if (net->exec_mode != SPN_EXEC_FUSED && net->ctx) cuda_check(cudaStreamSynchronize(net->ctx->stream), "stepwise drain");
  auto& P = net->pipe;
  for (int s = 0; s < spn_net::kPipe; ++s)
    if (P.stream[s]) cuda_check(cudaStreamSynchronize(P.stream[s]), "pipeline drain");
  if (P.fin) cuda_check(cudaStreamSynchronize(P.fin), "pipeline drain");
  return SPN_OK;
  NET_CATCH(false)
}

int spn_net_execute_host_async(spn_net* net, const void* const* host_inputs, uint32_t n_inputs, uint64_t n_points,
                               void* host_out, double* host_sum, uint64_t chunk_points) {
  NET_TRY
  execute_host_impl(net, host_inputs, n_inputs, n_points, host_out, host_sum, chunk_points);
  return SPN_OK;
  NET_CATCH(false)
}

int spn_net_execute_host(spn_net* net, const void* const* host_inputs, uint32_t n_inputs, uint64_t n_points,
                         void* host_out, double* host_sum, uint64_t chunk_points) {
  NET_TRY
  execute_host_impl(net, host_inputs, n_inputs, n_points, host_out, host_sum, chunk_points);
  return spn_net_host_wait(net);
  NET_CATCH(false)
}

void* spn_host_alloc_flags(uint64_t bytes, int flags) {
  void* p = nullptr;
  const unsigned f = (flags & SPN_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, f) != cudaSuccess) {
    spn::set_error("cudaHostAlloc failed");
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void* spn_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
    spn::set_error("cudaHostAlloc failed");
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void spn_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int spn_host_register(void* p, uint64_t bytes) {
  NET_TRY
  if (!p || !bytes) throw Error(SPN_ERR_INVALID, "spn_host_register: null or empty range");
  cuda_check(cudaHostRegister(p, bytes, cudaHostRegisterDefault), "cudaHostRegister");
  return SPN_OK;
  NET_CATCH(false)
}
int spn_host_unregister(void* p) {
  NET_TRY
  cuda_check(cudaHostUnregister(p), "cudaHostUnregister");
  return SPN_OK;
  NET_CATCH(false)
}

int spn_net_set_comm(spn_net* net, spn_comm* comm) {
  if (!net) return SPN_ERR_INVALID;
  net->comm = comm;
  return SPN_OK;
}

int spn_net_execute(spn_net* net, const spn_tensor* const* inputs, uint32_t n_inputs, spn_tensor* out,
                    double* sum_out_device, uint64_t first_point, uint64_t n_points) {
  NET_TRY
  if (!net->compiled) throw Error(SPN_ERR_INVALID, "network is not compiled");
  if (!net->ctx) throw Error(SPN_ERR_NO_DEVICE, "plan-only network (created without a context) cannot execute");
  if (n_inputs != net->n_inputs) throw Error(SPN_ERR_INVALID, "wrong number of input tensors");
  spn_ctx* ctx = net->ctx;
  uint64_t total = 0;
  for (uint32_t k = 0; k < n_inputs; ++k) {
    const spn_tensor* t = inputs[k];
    const Node& n = net->nodes[(size_t)net->input_nodes[k]];
    if (!t || !t->batched()) throw Error(SPN_ERR_INVALID, "network inputs must be batched tensors");
    if (t->dtype != n.dtype) throw Error(SPN_ERR_INVALID, "input dtype differs from the leaf declaration");
    if (!spn_net::same_structure(t->st, n.st)) throw Error(SPN_ERR_STRUCTURE, "input structure differs from the leaf declaration");
    if (k == 0)
      total = t->n_points;
    else if (t->n_points != total)
      throw Error(SPN_ERR_INVALID, "inputs have different point counts");
  }
  if (n_inputs == 0) total = out ? out->n_points : 1;
  if (n_points == 0) n_points = total - std::min(first_point, total);
  if (first_point + n_points > total) throw Error(SPN_ERR_INVALID, "point range out of bounds");
  const uint64_t size_out = net->values[(size_t)net->result_value].st.size();
  if (out) {
    if (!out->batched()) throw Error(SPN_ERR_INVALID, "output must be a batched tensor");
    if (out->size() != size_out) throw Error(SPN_ERR_STRUCTURE, "output tensor has the wrong size");
    if (out->n_points != total) throw Error(SPN_ERR_INVALID, "output tensor has a different point count");
  }
  if (!out && !sum_out_device) throw Error(SPN_ERR_INVALID, "nothing to compute: no output and no sum requested");

  if (net->exec_mode == SPN_EXEC_STEPWISE) {
    if (first_point != 0 || n_points != total) throw Error(SPN_ERR_INVALID, "stepwise execution runs all points");
    run_stepwise(net, inputs, out, sum_out_device);
    return SPN_OK;
  }

  // ---- fused ----
  std::string variant;
  for (uint32_t k = 0; k < n_inputs; ++k) {
    const spn_tensor* t = inputs[k];
    if (t->layout == SPN_COMPONENT_MAJOR)
      variant += 'c';
    else
      variant += (((uintptr_t)t->dptr + first_point * t->size() * t->elem_bytes()) % 32 == 0) ? 'p' : 'q';
  }
  variant += !out ? 'n' : (out->layout == SPN_COMPONENT_MAJOR ? 'c' : 'p');
  if (out && out->dtype == SPN_F64) variant += 'r';
  if (sum_out_device) variant += 's';
  JitKernel& k = net->get_variant(variant);
  const int kblock = k.block ? k.block : net->kBlock;
  jit_load(k, kblock);

  if (n_points == 0) {
    // an empty shard (world > n_points) still takes part in the exchange: its contribution is zero, and a rank that
    // skipped the collective would stall its peers and fall one epoch behind for good
    if (sum_out_device) {
      cuda_check(cudaMemsetAsync(sum_out_device, 0, size_out * 16, ctx->stream), "memset");
      if (net->comm) comm_exchange(net->comm, nullptr, 0, sum_out_device, (uint32_t)(2 * size_out));
    }
    return SPN_OK;
  }
  const uint64_t need = (n_points + kblock - 1) / kblock;
  const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)k.blocks_per_sm;   // persistent kernels: one wave
  const unsigned grid = (unsigned)std::min<uint64_t>(need, k.flat ? 0x7fffffffull : cap);

  std::vector<const double*> in_ptr(n_inputs);
  std::vector<unsigned long long> in_cs(n_inputs);
  for (uint32_t i = 0; i < n_inputs; ++i) {
    const spn_tensor* t = inputs[i];
    const size_t eb = t->elem_bytes();
    in_cs[i] = t->n_points;
    in_ptr[i] = reinterpret_cast<const double*>(
        (const char*)t->dptr + (t->layout == SPN_POINT_MAJOR ? first_point * t->size() * eb : first_point * eb));
  }
  double* out_ptr = nullptr;
  unsigned long long ocs = 0;
  if (out) {
    const size_t eb = out->elem_bytes();
    ocs = out->n_points;
    out_ptr = reinterpret_cast<double*>(
        (char*)out->dptr + (out->layout == SPN_POINT_MAJOR ? first_point * out->size() * eb : first_point * eb));
  }
  double2* partial = nullptr;
  if (sum_out_device) {
    // sized for the largest grid any variant can use, allocated once (keeps execute graph-capturable)
    const size_t want = (size_t)cap * size_out;
    if (net->sum_scratch_elems < want) {
      if (net->sum_scratch) cuda_check(cudaFree(net->sum_scratch), "sum scratch");
      net->sum_scratch = nullptr;
      cuda_check(cudaMalloc((void**)&net->sum_scratch, want * sizeof(double2)), "sum scratch");
      net->sum_scratch_elems = want;
    }
    partial = net->sum_scratch;
  }
  unsigned long long np = n_points;
  std::vector<void*> args;
  for (uint32_t i = 0; i < n_inputs; ++i) {
    args.push_back(&in_ptr[i]);
    args.push_back(&in_cs[i]);
  }
  args.push_back(&out_ptr);
  args.push_back(&ocs);
  args.push_back(&partial);
  args.push_back(&np);
  cuda_check(cudaLaunchKernel((const void*)k.kernel, dim3(grid), dim3((unsigned)kblock), args.data(), k.smem_bytes, ctx->stream),
             "launch fused network kernel");
  ctx->launches++;
  if (sum_out_device && net->comm) {
    // multi-GPU Monte-Carlo sum: fold the per-CTA partials and exchange with the peers in ONE kernel (comm.cu)
    comm_exchange(net->comm, partial, (int)grid, sum_out_device, (uint32_t)(2 * size_out));
  } else if (sum_out_device) {
    cuda_check(launch_reduce_partials(partial, (int)grid, size_out, reinterpret_cast<double2*>(sum_out_device),
                                      ctx->stream),
               "reduce_partials_kernel");
    ctx->launches++;
  }
  return SPN_OK;
  NET_CATCH(false)
}

}  // extern "C"
