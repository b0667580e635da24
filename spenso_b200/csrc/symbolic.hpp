// Per-point program of a fused network kernel.
//
// The network executor (net.cu) walks the contraction schedule ONCE and, instead of computing
// numbers, records what every component of every intermediate tensor is as a function of the
// per-point inputs.  Components are tracked as real quantities `coef * value` where `value` is an
// SSA id (or a literal): multiplying by the library constants that dominate these networks
// (gamma matrices: 0, +-1, +-i; metrics: +-1; colour generators: +-1/2, ...) therefore emits no
// instruction at all — zeros vanish, signs and factors of i become operand modifiers/renames of
// the next FMA.  What is left is a straight-line list of DMUL/DFMA/DADD over the components that
// actually depend on the kinematics; net.cu prints it as the body of one CUDA kernel
// (thread = phase-space point, every intermediate in registers, inputs read exactly once).
//
// The arithmetic this replaces is the reference's per-point kernels:
//   spenso/src/contraction/{singlecontract,multicontract,exteriorproduct}.rs and
//   spenso/src/algebra/{add_assign,neg,scalar_mul}.rs  (C[o] = sum_k s(k) A[..k] B[..k]),
//   complex product spenso/src/algebra/complex/mul.rs:16-21.
// FMA contraction and a different summation order are allowed by the 1e-12 parity tolerance.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace spn {

// a real quantity: coef * value(var); var < 0 => the literal `coef`
struct RR {
  int var = -1;
  double coef = 0.0;
  bool is_const() const { return var < 0; }
  bool is_zero() const { return var < 0 && coef == 0.0; }
};
inline RR rr_const(double c) { return RR{-1, c}; }
inline RR rr_neg(RR a) { return RR{a.var, -a.coef}; }

// one complex tensor component
struct CE {
  RR re, im;
};

enum InsOp : uint8_t {
  I_LOADC,  // dst, dst+1 = re, im of input[in] component [flat]
  I_LOADR,  // dst = real input[in] component [flat]
  I_MUL,    // dst = k * x * y
  I_FMA,    // dst = k * x * y + sz * z
  I_MULK,   // dst = k * x
  I_FMAK,   // dst = k * x + sz * z
  I_ADDK,   // dst = sz * z + kc
  I_DIV,    // dst = k * x / y   (x < 0: the numerator is the literal k)
  I_SQRT    // dst = sqrt(k * x)
};

struct Ins {
  InsOp op;
  int dst = -1, x = -1, y = -1, z = -1;
  double k = 1.0, kc = 0.0;
  int sz = 1;
  int in = -1;         // loads: input index
  uint64_t flat = 0;   // loads: component
};

struct Term {
  double k;
  int x, y;  // y < 0: linear term k*x ; x < 0 too: constant k
};

struct Program {
  std::vector<Ins> ins;
  int n_vals = 0;
  // lazily created loads: (input, flat) -> value id of the real part
  std::map<std::pair<int, uint64_t>, int> loaded;
  // emitted sums by canonical term list -> (value id, factor): sum == norm * factor * value(id)
  std::map<std::vector<double>, std::pair<int, double>> cse;
  bool cse_enabled = !getenv("SPN_NO_CSE");  // A/B switch for measurements

  int fresh(int n = 1) {
    int v = n_vals;
    n_vals += n;
    return v;
  }
  CE load(int in, uint64_t flat, bool is_complex) {
    auto key = std::make_pair(in, flat);
    auto it = loaded.find(key);
    int v;
    if (it != loaded.end()) {
      v = it->second;
    } else {
      v = fresh(is_complex ? 2 : 1);
      Ins i;
      i.op = is_complex ? I_LOADC : I_LOADR;
      i.dst = v;
      i.in = in;
      i.flat = flat;
      ins.push_back(i);
      loaded[key] = v;
    }
    CE e;
    e.re = RR{v, 1.0};
    e.im = is_complex ? RR{v + 1, 1.0} : rr_const(0.0);
    return e;
  }

  // ---- term algebra ------------------------------------------------------------------------
  static void push_product(std::vector<Term>& out, RR a, RR b, double sign) {
    if (a.is_zero() || b.is_zero()) return;
    double k = sign * a.coef * b.coef;
    if (a.is_const() && b.is_const())
      out.push_back({k, -1, -1});
    else if (a.is_const())
      out.push_back({k, b.var, -1});
    else if (b.is_const())
      out.push_back({k, a.var, -1});
    else
      out.push_back({k, a.var < b.var ? a.var : b.var, a.var < b.var ? b.var : a.var});
  }
  static void push_rr(std::vector<Term>& out, RR a, double sign) {
    if (a.is_zero()) return;
    if (a.is_const())
      out.push_back({sign * a.coef, -1, -1});
    else
      out.push_back({sign * a.coef, a.var, -1});
  }
  // (re, im) += sign * a * b
  static void push_cmul(std::vector<Term>& re, std::vector<Term>& im, const CE& a, const CE& b, double sign) {
    push_product(re, a.re, b.re, sign);
    push_product(re, a.im, b.im, -sign);
    push_product(im, a.re, b.im, sign);
    push_product(im, a.im, b.re, sign);
  }

  // emit the sum of `terms`; returns what the result is
  RR sum(std::vector<Term> terms) {
    // merge like terms
    std::map<std::pair<int, int>, double> acc_map;
    std::vector<std::pair<int, int>> order;
    for (auto& t : terms) {
      auto key = std::make_pair(t.x, t.y);
      auto it = acc_map.find(key);
      if (it == acc_map.end()) {
        acc_map[key] = t.k;
        order.push_back(key);
      } else {
        it->second += t.k;
      }
    }
    std::vector<Term> bil, lin;
    double cst = 0.0;
    for (auto& key : order) {
      double k = acc_map[key];
      if (k == 0.0) continue;
      if (key.first < 0)
        cst += k;
      else if (key.second < 0)
        lin.push_back({k, key.first, -1});
      else
        bil.push_back({k, key.first, key.second});
    }
    if (bil.empty() && lin.empty()) return rr_const(cst);
    if (bil.empty() && lin.size() == 1 && cst == 0.0) return RR{lin[0].x, lin[0].k};
    // value numbering: a sum is identified by its term set in canonical order, normalised so that the
    // leading coefficient is +1.  The same sum (or an exact multiple of it: the transposed entry of a
    // symmetric outer product, a sign-flipped copy, ...) is emitted once and re-used with a coefficient.
    auto by_vars = [](const Term& a, const Term& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; };
    std::sort(bil.begin(), bil.end(), by_vars);
    std::sort(lin.begin(), lin.end(), by_vars);
    const double norm = bil.empty() ? lin[0].k : bil[0].k;
    std::vector<double> key;
    key.reserve(3 * (bil.size() + lin.size()) + 1);
    for (auto& t : bil) { key.push_back(t.x); key.push_back(t.y); key.push_back(t.k / norm); }
    for (auto& t : lin) { key.push_back(t.x); key.push_back(-1); key.push_back(t.k / norm); }
    key.push_back(cst / norm);
    if (cse_enabled) {
      auto hit = cse.find(key);
      if (hit != cse.end()) return RR{hit->second.first, norm * hit->second.second};
    }
    const double s = std::fabs(norm);
    int av = -1, as = 1;  // running accumulator: as * value(av)
    for (auto& t : bil) {
      Ins i;
      i.k = t.k / s;
      i.x = t.x;
      i.y = t.y;
      i.dst = fresh();
      if (av < 0) {
        i.op = I_MUL;
      } else {
        i.op = I_FMA;
        i.z = av;
        i.sz = as;
      }
      ins.push_back(i);
      av = i.dst;
      as = 1;
    }
    for (auto& t : lin) {
      double r = t.k / s;
      if (av < 0 && std::fabs(r) == 1.0) {
        av = t.x;
        as = r > 0 ? 1 : -1;
        continue;
      }
      Ins i;
      i.k = r;
      i.x = t.x;
      i.dst = fresh();
      if (av < 0) {
        i.op = I_MULK;
      } else {
        i.op = I_FMAK;
        i.z = av;
        i.sz = as;
      }
      ins.push_back(i);
      av = i.dst;
      as = 1;
    }
    if (cst != 0.0) {
      Ins i;
      i.op = I_ADDK;
      i.z = av;
      i.sz = as;
      i.kc = cst / s;
      i.dst = fresh();
      ins.push_back(i);
      av = i.dst;
      as = 1;
    }
    // value(av) * as == (sum / s) == (sum / norm) * sign(norm)
    cse[key] = std::make_pair(av, (norm < 0 ? -1.0 : 1.0) * as);
    return RR{av, s * as};
  }

  // 1 / (re + i im)
  CE reciprocal(const CE& a) {
    if (a.re.is_const() && a.im.is_const()) {
      double d = a.re.coef * a.re.coef + a.im.coef * a.im.coef;
      return CE{rr_const(a.re.coef / d), rr_const(-a.im.coef / d)};
    }
    std::vector<Term> t;
    push_product(t, a.re, a.re, 1.0);
    push_product(t, a.im, a.im, 1.0);
    RR d = sum(t);
    auto div = [&](RR n, double sign) -> RR {
      if (n.is_zero()) return rr_const(0.0);
      // n / d with d = d.coef * value(d.var); x < 0: the numerator is the literal k
      Ins i;
      i.op = I_DIV;
      i.x = n.is_const() ? -1 : n.var;
      i.y = d.var;
      i.k = sign * n.coef / d.coef;
      i.dst = fresh();
      ins.push_back(i);
      return RR{i.dst, 1.0};
    };
    return CE{div(a.re, 1.0), div(a.im, -1.0)};
  }

  // a / b
  CE divide(const CE& a, const CE& b) {
    if (b.im.is_zero()) {  // real denominator: two real divisions
      auto div = [&](RR n) -> RR {
        if (n.is_zero()) return rr_const(0.0);
        if (b.re.is_const()) return RR{n.var, n.coef / b.re.coef};
        Ins i;
        i.op = I_DIV;
        i.x = n.is_const() ? -1 : n.var;
        i.y = b.re.var;
        i.k = n.coef / b.re.coef;
        i.dst = fresh();
        ins.push_back(i);
        return RR{i.dst, 1.0};
      };
      return CE{div(a.re), div(a.im)};
    }
    CE r = reciprocal(b);
    std::vector<Term> tr, ti;
    push_cmul(tr, ti, a, r, 1.0);
    return CE{sum(tr), sum(ti)};
  }
  // sqrt of a quantity known to be real (imaginary part structurally zero); the caller guarantees >= 0
  RR real_sqrt(RR a) {
    if (a.is_const()) return rr_const(std::sqrt(a.coef));
    Ins i;
    i.op = I_SQRT;
    i.x = a.var;
    i.k = a.coef;
    i.dst = fresh();
    ins.push_back(i);
    return RR{i.dst, 1.0};
  }

  // ---- scale propagation --------------------------------------------------------------------------------------
  // Value numbering normalises every sum to leading coefficient 1 and hands the factor to the consumer.  A consumer
  // that is itself a product then multiplies by a non-unit constant, `(2.0 * v120) * v134`, which costs a DMUL of its
  // own.  Where the producer can absorb the factor for free — k*x (+-) z with |c*k| = 1 becomes c*z (+-) x, one FMA
  // either way; k*x becomes x; k*x*y takes the factor into k — and that makes more multiplicands unit than it
  // breaks, the value is rescaled and its uses (results included) follow.  `results` are the program's outputs.
  int rescale_pass(std::vector<RR*>& results) {
    if (ins.size() > 200000) return 0;   // far beyond what runs as straight-line code anyway
    std::vector<int> def(n_vals, -1);
    std::vector<std::vector<int>> users((size_t)n_vals);   // instructions reading each value (once per instruction)
    for (size_t n = 0; n < ins.size(); ++n) {
      const Ins& i = ins[n];
      def[i.dst] = (int)n;
      if (i.op == I_LOADC) def[i.dst + 1] = (int)n;
      if (i.op == I_LOADC || i.op == I_LOADR) continue;
      for (int v : {i.x, i.y, i.z})
        if (v >= 0 && (users[(size_t)v].empty() || users[(size_t)v].back() != (int)n)) users[(size_t)v].push_back((int)n);
    }
    std::vector<std::vector<RR*>> result_uses((size_t)n_vals);
    for (RR* r : results)
      if (r->var >= 0) result_uses[(size_t)r->var].push_back(r);
    auto mul_cost = [](double k) { return std::fabs(k) == 1.0 ? 1 : 2; };
    std::vector<char> frozen((size_t)n_vals, 0);   // operands of a rewritten definition: their user lists are stale
    int changed = 0;
    for (int v = 0; v < n_vals; ++v) {
      if (def[v] < 0 || frozen[(size_t)v]) continue;
      const Ins d = ins[(size_t)def[v]];
      if (d.op != I_FMAK && d.op != I_MULK && d.op != I_MUL) continue;
      // product uses vote for the factor that would make them unit; anything that cannot follow a rescaling pins v
      std::map<double, int> votes;
      bool pinned = false;
      int unit_product_uses = 0;
      for (int un : users[(size_t)v]) {
        const Ins& u = ins[(size_t)un];
        const bool is_prod = u.op == I_MUL || u.op == I_FMA;
        const int mult = (u.x == v) + ((is_prod || u.op == I_DIV) && u.y == v);
        if (u.z == v && (u.op == I_FMA || u.op == I_FMAK || u.op == I_ADDK)) pinned = true;  // addend: sign only
        if (!mult) continue;
        if (is_prod) {
          if (mult == 2) {
            pinned = true;
          } else if (std::fabs(u.k) == 1.0) {
            ++unit_product_uses;
          } else {
            votes[std::fabs(u.k)]++;
          }
        } else if (u.op != I_FMAK && u.op != I_MULK) {
          pinned = true;  // division, square root (constant-times-value takes any constant at the same cost)
        }
      }
      if (pinned || votes.empty()) continue;
      double best_c = 0.0;
      int best_gain = 0;
      for (auto& cv : votes) {
        const double c = cv.first;
        int delta_def = 0;
        if (d.op == I_FMAK) {
          if (std::fabs(c * d.k) != 1.0) continue;
        } else if (d.op == I_MUL) {
          delta_def = mul_cost(c * d.k) - mul_cost(d.k);
        }
        const int gain = cv.second - unit_product_uses - delta_def;
        if (gain > best_gain) {
          best_gain = gain;
          best_c = c;
        }
      }
      if (best_gain <= 0) continue;
      const double c = best_c;
      Ins& D = ins[(size_t)def[v]];
      if (D.op == I_FMAK) {            // k*x + sz*z  ->  (c*sz)*z + (c*k)*x with |c*k| = 1
        const double ck = c * D.k;
        const int x0 = D.x, z0 = D.z, sz0 = D.sz;
        D.x = z0;
        D.k = c * (double)sz0;
        D.z = x0;
        D.sz = ck > 0 ? 1 : -1;
        frozen[(size_t)x0] = frozen[(size_t)z0] = 1;   // their roles in D changed; the next round sees fresh lists
      } else {
        D.k *= c;                      // I_MULK / I_MUL
      }
      for (int un : users[(size_t)v]) {
        Ins& u = ins[(size_t)un];
        const bool is_prod = u.op == I_MUL || u.op == I_FMA;
        if (u.x == v || (is_prod && u.y == v)) u.k /= c;
      }
      for (RR* r : result_uses[(size_t)v]) r->coef /= c;
      ++changed;
    }
    return changed;
  }

  // ---- dead code elimination + statistics ---------------------------------------------------
  // keep only instructions that feed `roots`; returns the number of FP64 arithmetic instructions
  struct Stats {
    uint64_t n_fp = 0, n_fma = 0, n_loads = 0, n_load_bytes = 0;
  };
  Stats eliminate_dead(const std::vector<int>& roots) {
    std::vector<char> live(n_vals, 0);
    for (int r : roots)
      if (r >= 0) live[r] = 1;
    std::vector<Ins> kept;
    for (size_t n = ins.size(); n-- > 0;) {
      const Ins& i = ins[n];
      bool l = live[i.dst] || (i.op == I_LOADC && live[i.dst + 1]);
      if (!l) continue;
      if (i.x >= 0) live[i.x] = 1;
      if (i.y >= 0) live[i.y] = 1;
      if (i.z >= 0) live[i.z] = 1;
      kept.push_back(i);
    }
    ins.assign(kept.rbegin(), kept.rend());
    Stats s;
    for (auto& i : ins) {
      switch (i.op) {
        case I_LOADC: s.n_loads++; s.n_load_bytes += 16; break;
        case I_LOADR: s.n_loads++; s.n_load_bytes += 8; break;
        case I_MUL: s.n_fp += (std::fabs(i.k) == 1.0) ? 1 : 2; break;
        case I_FMA: s.n_fp += (std::fabs(i.k) == 1.0) ? 1 : 2; s.n_fma++; break;
        case I_DIV: s.n_fp += 8; break;
        case I_SQRT: s.n_fp += 10; break;
        default: s.n_fp += 1; break;
      }
    }
    return s;
  }
};

// a double as a CUDA C literal that round-trips exactly; non-finite values have no literal form and are spelled
// through their bit pattern
inline std::string lit(double v) {
  char buf[64];
  if (!std::isfinite(v)) {
    unsigned long long bits;
    static_assert(sizeof bits == sizeof v, "double is 64 bits");
    std::memcpy(&bits, &v, sizeof bits);
    std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016llxLL)", bits);
    return buf;
  }
  std::snprintf(buf, sizeof buf, "%.17g", v);
  std::string s = buf;
  if (s.find_first_of(".e") == std::string::npos) s += ".0";
  return s;
}
// the same for the JSON plan format: +-1e999 reads back as +-infinity with any strtod, NaN as the bare word
inline std::string jlit(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "1e999" : "-1e999";
  return lit(v);
}

}  // namespace spn
