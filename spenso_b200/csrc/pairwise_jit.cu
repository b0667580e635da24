// Specialised kernels for the pairwise Contract trait on small per-point tensors (the physics sizes: Bispinor 4,
// Minkowski 4, colour 3/8).  The generic kernels of kernels.cu spend their time on index arithmetic (one thread
// per output component, run-time strides, table lookups per multiply); here the stride plan of ONE contraction is
// burnt into a straight-line kernel — thread = phase-space point, every operand component read once, every product at
// a literal offset — compiled with NVRTC for sm_100a and cached per (plan, storage, dtype, layout).
//
// Memory side.  A point-major operand of S bytes per point read thread = point makes every warp-wide load touch 32
// cache lines (one sector each); beyond S = 64 B that runs into the L1 tag rate and the kernels of round 1 stalled at
// 0.75-0.87 of the HBM roofline.  So point-major operands and results are STAGED: every warp owns slabs of 32
// consecutive points; it copies the slab — a contiguous 32*S byte range — global -> shared with 16-byte cp.async
// (LDGSTS; consecutive lanes = consecutive addresses, i.e. full lines) into a padded [point][unit] layout (row stride
// odd in 16-byte units => the thread-=-point reads that follow are bank-conflict free); results go registers -> padded
// shared slab -> contiguous 16-byte global stores.  No block barrier: a warp only touches its own slabs (__syncwarp).
// A CTA is ONE warp with ONE slab and the grid has as many CTAs as there are slabs: the hardware scheduler keeps the
// resident warps out of step with each other, which is what overlaps the load phase of one with the store phase of
// another (a persistent grid of 4-warp CTAs — round 2's first version, still behind SPN_PAIR_PERSISTENT=1 — measured
// 10-22 % slower on every shape).  Component-major operands are already coalesced thread = point and are read
// directly; a shared dense operand is read with uniform (broadcast) loads; a shared sparse operand is literals in the code.
//
// The arithmetic is the reference's, operation for operation (singlecontract.rs:25-196, multicontract.rs:25-200,
// exteriorproduct.rs:19-153; complex product algebra/complex/mul.rs:16-21; f64 upgraded to (x, 0.0),
// upgrading_arithmetic.rs:1285-1293): one accumulator per result element starting at zero, contracted index in the
// reference's enumeration order, "-=" where the metric is negative, and __dmul_rn/__dadd_rn/__dsub_rn so that nothing
// is contracted into an FMA.  Results are therefore bit-identical to the generic kernels and to the oracle.
// A shared sparse operand contributes only its stored entries (as literals), like the sparse fibre iterators.
#include <algorithm>
#include <memory>
#include <sstream>

#include "jit.hpp"
#include "symbolic.hpp"
#include "tensor.hpp"

namespace spn {

namespace {

// direct kernels: threads per CTA.  Small, for the same reason: the threads of a CTA run load phase / store phase in
// step, and only CTAs out of step overlap the two (component-major 768-byte points: 5.94 TB/s at 64 threads, 4.72 at 256).
const int kPairBlock = getenv("SPN_PAIR_BLOCK") ? atoi(getenv("SPN_PAIR_BLOCK")) : 64;
// staged kernels: warps per CTA.  One: a CTA is a warp with its slab — nothing is shared between warps, so there is
// nothing to gain from larger CTAs, and single-warp CTAs drift apart in time (loads of one overlap stores of another)
// where the 4 warps of one CTA start together: 1 344-byte points 6.83 vs 5.26 TB/s (profiles/r02_pairwise_grid.txt).
const int kSlabWarps = getenv("SPN_PAIR_SLAB_WARPS") ? atoi(getenv("SPN_PAIR_SLAB_WARPS")) : 1;

struct Operand {
  const spn_tensor* t;
  char name;  // 'A' / 'B'
};

// how one tensor (operand or result) reaches the registers of the thread that owns its point
struct Staging {
  bool staged = false;
  uint64_t units = 0;   // 16-byte units per point
  uint64_t row = 0;     // padded row stride in units (odd)
  uint64_t offset = 0;  // first unit inside a stage (operands) / inside the warp's region (result)
};

bool stageable(const spn_tensor* t) {
  if (getenv("SPN_PAIR_NO_SLAB")) return false;
  if (!t->batched() || t->layout != SPN_POINT_MAJOR) return false;
  const uint64_t bytes = t->size() * t->elem_bytes();
  return bytes % 16 == 0 && bytes >= 32 && ((uintptr_t)t->dptr % 16) == 0;
}
// SPN_LOAD_MODE=tma: the operand slabs are moved by the TMA engine — one cp.async.bulk per operand and slab, completion
// on an mbarrier (UBLKCP + SYNCS in SASS) — instead of per-lane 16-byte cp.async.  A bulk copy is linear, so the slab
// lands UNPADDED and the thread = point reads that follow hit the same bank group 2..8 ways whenever the per-point size
// is a multiple of 32 bytes.  Kept for measurement (profiles/r02_pairwise_tma.txt): it is slower than the padded
// LDGSTS slabs on every shape of tools/pairwise_bench.py.
bool tma_mode() {
  const char* e = getenv("SPN_LOAD_MODE");   // read per call: a test flips it between contractions
  return e && std::string(e) == "tma";
}
Staging staging_of(const spn_tensor* t, bool operand) {
  Staging s;
  s.staged = stageable(t);
  if (s.staged) {
    s.units = t->size() * t->elem_bytes() / 16;
    s.row = (operand && tma_mode()) ? s.units : (s.units | 1ull);  // odd: 8 consecutive rows start in 8 different 16-byte bank groups
  }
  return s;
}

// address expression (in doubles) of component f of an operand at point p — direct (unstaged) access
std::string addr(const Operand& o, uint64_t f) {
  std::ostringstream s;
  const int w = o.t->dtype == SPN_C64 ? 2 : 1;
  if (!o.t->batched())
    s << o.name << " + " << f * (uint64_t)w << "ull";
  else if (o.t->layout == SPN_POINT_MAJOR)
    s << o.name << " + (p * " << o.t->size() << "ull + " << f << "ull) * " << w;
  else
    s << o.name << " + (" << f << "ull * cs" << o.name << " + p) * " << w;
  return s.str();
}

struct Generated {
  std::string source;
  uint32_t rows = 1;   // row-parallel kernels: threads per point
  uint32_t points_per_cta = 0;   // row-parallel kernels: whole points per CTA
  bool slab = false;
  size_t smem_bytes = 0;
  int block = 0;
};

// ---- row-parallel kernels: per-point tensors too large for one thread's registers ---------------------------------
// dense[coad^3] x dense[coad^2] -> [coad^3] is a 64x8 by 8x8 product per point: 576 operand components, 512 results, 4 096
// products.  The generic kernel (one thread per result element, offsets from tables) re-reads every operand element 8-64
// times through L1 and reaches 1.1 TB/s.  Here the free indices of ONE batched operand X become a thread dimension:
// thread = (point, row i of X), a CTA = the rows of whole points (64 threads or one point, whichever is more), launched
// flat.  A thread keeps its row of X — the n_k contracted components at literal offsets from a per-row base, 256-bit
// loads — in registers, walks the free components j of the other operand Y and writes result[i, j].  Y: a batched one
// is staged once per CTA in shared memory (every row of a point reads all of it: a broadcast; read through L1 instead
// the rows ran into each other's misses — L1 hit rate 38 %, FP64 pipe 41 % busy — 1.12 ms against 0.86 ms staged); a
// shared dense one is read through L1 at literal offsets; a shared sparse one is its stored entries as literals, and
// only the components of X that meet one are loaded.  The sums are the reference's (k order, from zero, "-=" on a
// negative metric, no FMA): bit-for-bit the generic kernels.  A thread takes two neighbouring rows (register tiling: one
// read of Y's component, two products; 856 -> 820 us, four rows: 928).  Measured at 2^18 points
// (profiles/r02_pairwise_rows.txt): 4.01 ms -> 0.82 ms (5.6 TB/s) for the product above, 0.85 -> 0.29 ms for
// dense[coad^3] x sparse f.  A prefetch of the rows 256-8192 points ahead into L2 was tried and dropped (+3 % before Y
// was staged, -7 % after).  Rows too long for registers against few columns (full contractions to a scalar / vector) are
// streamed: a loop over k, offsets in constant tables, one accumulator per column (1.05 -> 0.96 ms for coad^3 . coad^3).
struct RowSplit {
  bool ok = false;
  bool rows_from_a = true;
  uint32_t n_rows = 0, n_cols = 0;
  bool k_outer = false;                 // the row does not fit registers: walk k outside, keep one accumulator per column
  std::vector<uint64_t> row_x, row_c;   // per row: element offset inside X's record / inside the result record
  std::vector<uint64_t> col_y, col_c;   // per column: element offset inside Y's record / inside the result record
};

bool register_form_fits(const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a, const spn_tensor* b) {
  const uint64_t loaded = (a->sparse() ? 0 : a->size()) + (b->sparse() ? 0 : b->size());
  const uint64_t products = plan.size_out * (uint64_t)kt.off_a.size();
  // everything lives in registers: at most ~80 complex operand components and a few thousand products
  return !(loaded > 80 || plan.size_out > 256 || products > 4096 || products == 0);
}

// register tiling: a thread takes R neighbouring rows, so every component of Y it reads serves R products
// (SPN_PAIR_ROWS_PER_THREAD=1|2|4 for A/B runs)
uint32_t rows_per_thread(uint64_t n_rows, uint64_t n_cols, uint64_t n_k, bool k_outer) {
  uint32_t R = getenv("SPN_PAIR_ROWS_PER_THREAD") ? (uint32_t)atoi(getenv("SPN_PAIR_ROWS_PER_THREAD")) : 2u;
  if (k_outer) {   // registers hold R x n_cols accumulators
    while (R > 1 && (n_rows % R != 0 || (uint64_t)R * n_cols > 32 || (uint64_t)R * n_cols * n_k > 4096)) R /= 2;
    return R == 0 ? 1u : R;
  }
  while (R > 1 && (n_rows % R != 0 || n_rows / R < 2 || (uint64_t)R * n_k > 48 || (uint64_t)R * n_cols * n_k > 1536)) R /= 2;
  return R == 0 ? 1u : R;
}

RowSplit row_split(const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a, const spn_tensor* b) {
  RowSplit rs;
  if (getenv("SPN_PAIR_NO_ROWS")) return rs;
  auto can_be_x = [](const spn_tensor* t) { return t->batched() && t->layout == SPN_POINT_MAJOR && ((uintptr_t)t->dptr % 16) == 0; };
  auto can_be_y = [](const spn_tensor* t) { return !t->batched() || (t->layout == SPN_POINT_MAJOR && ((uintptr_t)t->dptr % 16) == 0); };
  uint64_t free_a = 1, free_b = 1;
  for (uint32_t r = 0; r < plan.order_out; ++r) (plan.partition[r] ? free_a : free_b) *= plan.out_dim[r];
  const bool a_x = can_be_x(a) && can_be_y(b), b_x = can_be_x(b) && can_be_y(a);
  if (!a_x && !b_x) return rs;
  rs.rows_from_a = a_x && (!b_x || free_a >= free_b);
  const uint64_t n_rows = rs.rows_from_a ? free_a : free_b, n_cols = rs.rows_from_a ? free_b : free_a;
  const uint64_t n_k = kt.off_a.size();
  if (n_rows < 1 || n_rows > (1u << 16) || n_cols > 256 || n_k == 0) return rs;   // (per-row offset tables)
  const bool row_in_regs = n_rows >= 2 && n_k <= 64 && n_cols * n_k <= 1024;
  // long rows against few columns (full contractions to a scalar or a vector): the components of the row are used once
  // per column, so they are streamed — k outside, one accumulator per column; every accumulator still sums k in order
  rs.k_outer = !row_in_regs && n_cols <= 16 && n_k <= 4096 && !a->sparse() && !b->sparse();
  if (!row_in_regs && !rs.k_outer) return rs;
  if (plan.size_a > (1u << 20) || plan.size_b > (1u << 20) || plan.size_out > (1u << 20)) return rs;
  rs.n_rows = (uint32_t)n_rows;
  rs.n_cols = (uint32_t)n_cols;
  // result record: row-major over out_dim
  std::vector<uint64_t> ostride(plan.order_out, 1);
  for (uint32_t r = plan.order_out; r-- > 1;) ostride[r - 1] = ostride[r] * plan.out_dim[r];
  std::vector<uint32_t> ix(plan.order_out, 0);
  for (uint64_t e = 0; e < plan.size_out; ++e) {
    uint64_t ox = 0, oy = 0, cx = 0, cy = 0;
    bool x_zero = true, y_zero = true;
    for (uint32_t r = 0; r < plan.order_out; ++r) {
      const bool from_x = (plan.partition[r] != 0) == rs.rows_from_a;
      const uint64_t so = ix[r] * (rs.rows_from_a ? plan.out_stride_a[r] : plan.out_stride_b[r]);
      const uint64_t sy = ix[r] * (rs.rows_from_a ? plan.out_stride_b[r] : plan.out_stride_a[r]);
      if (from_x) { ox += so; cx += ix[r] * ostride[r]; x_zero = x_zero && ix[r] == 0; }
      else { oy += sy; cy += ix[r] * ostride[r]; y_zero = y_zero && ix[r] == 0; }
    }
    if (y_zero) { rs.row_x.push_back(ox); rs.row_c.push_back(cx); }   // enumerated with the other side at zero
    if (x_zero) { rs.col_y.push_back(oy); rs.col_c.push_back(cy); }
    for (uint32_t r = plan.order_out; r-- > 0;) {
      if (++ix[r] < plan.out_dim[r]) break;
      ix[r] = 0;
    }
  }
  rs.ok = rs.row_x.size() == n_rows && rs.col_y.size() == n_cols;
  return rs;
}

Generated generate_rows(const RowSplit& rs, const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a,
                        const spn_tensor* b, const spn_tensor* res, const std::string& kname) {
  const spn_tensor* X = rs.rows_from_a ? a : b;
  const spn_tensor* Y = rs.rows_from_a ? b : a;
  const char xn = rs.rows_from_a ? 'A' : 'B', yn = rs.rows_from_a ? 'B' : 'A';
  const std::vector<long long>& kx = rs.rows_from_a ? kt.off_a : kt.off_b;
  const std::vector<long long>& ky = rs.rows_from_a ? kt.off_b : kt.off_a;
  const bool xc = X->dtype == SPN_C64, yc = Y->dtype == SPN_C64, oc = res->dtype == SPN_C64;
  Generated gen;
  const uint32_t R = rows_per_thread(rs.n_rows, rs.n_cols, kx.size(), rs.k_outer);
  // threads per point: one per R rows, at most 256 — a point with more rows is walked in passes of 256 R rows
  const uint32_t T = std::min<uint32_t>(rs.n_rows / R, 256u);
  const bool passes = (uint64_t)T * R < rs.n_rows;
  // a CTA covers whole points: P of them if their threads fill 64, else one (threads rounded up to a warp)
  const uint32_t P = T >= 64 ? 1u : 64u / T;
  gen.block = (int)(((uint64_t)P * T + 31) / 32 * 32);
  gen.rows = T;
  gen.points_per_cta = P;
  // a batched Y is staged once per CTA in shared memory: every row-thread of a point reads all of it (a broadcast), and
  // through L1 those reads met each other's misses (ncu: L1 hit rate 38 %, long_scoreboard 9 cycles per issue, FP64 pipe 41 %)
  bool y_smem = Y->batched() && (Y->size() * (uint64_t)(yc ? 16 : 8)) % 16 == 0 && !getenv("SPN_PAIR_ROWS_NO_SMEM");
  const uint64_t y_units = y_smem ? (Y->size() * (uint64_t)(yc ? 16 : 8) + 15) / 16 : 0;   // 16-byte units per point
  const uint64_t y_row = y_units | 1ull;   // odd stride: the points of one warp sit in different banks
  // (few threads per point and a large partner — a full contraction of two large tensors — would need hundreds of KB:
  // there every thread streams its own record of Y like its row of X)
  if (y_smem && P * y_row * 16 > 64 * 1024) y_smem = false;
  if (y_smem) gen.smem_bytes = (size_t)(P * y_row * 16);
  std::ostringstream o;
  o << "// generated by spenso_b200 (pairwise_jit.cu): one pairwise contraction, thread = (point, row of " << xn << "), reference rounding\n";
  o << "// " << rs.n_rows << " rows x " << rs.n_cols << " columns x " << kx.size() << " contracted components per point, " << R
    << " row(s) per thread, " << P << " point(s) per CTA\n";
  o << "typedef unsigned long long ull;\n";
  o << "__device__ __forceinline__ double2 cmul_rn(double2 a, double2 b) {\n"
       "  return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));\n}\n";
  o << "__device__ __forceinline__ void ld32(const double* p, double2& a, double2& b) {\n"
       "  asm(\"ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];\" : \"=d\"(a.x), \"=d\"(a.y), \"=d\"(b.x), \"=d\"(b.y) : \"l\"(p));\n}\n";
  // no "memory" clobber: nothing reads the result, and the loads of the next columns may move above the store
  o << "__device__ __forceinline__ void st32(double* p, double2 a, double2 b) {\n"
       "  asm volatile(\"st.global.v4.f64 [%0], {%1,%2,%3,%4};\" :: \"l\"(p), \"d\"(a.x), \"d\"(a.y), \"d\"(b.x), \"d\"(b.y));\n}\n";
  o << "__device__ const unsigned kRowX[" << rs.n_rows << "] = {";
  for (uint64_t v : rs.row_x) o << v << "u, ";
  o << "};\n__device__ const unsigned kRowC[" << rs.n_rows << "] = {";
  for (uint64_t v : rs.row_c) o << v << "u, ";
  o << "};\n";
  if (rs.k_outer) {
    o << "__constant__ unsigned kTabX[" << kx.size() << "] = {";
    for (long long v : kx) o << v << "u, ";
    o << "};\n__constant__ unsigned kTabY[" << ky.size() << "] = {";
    for (long long v : ky) o << v << "u, ";
    o << "};\n__constant__ unsigned char kTabNeg[" << kt.sign.size() << "] = {";
    for (unsigned char v : kt.sign) o << (v ? 1 : 0) << ", ";
    o << "};\n";
  }
  o << "extern \"C\" __global__ void __launch_bounds__(" << gen.block << ") " << kname
    << "(const double* __restrict__ A, ull csA, const double* __restrict__ B, ull csB, double* __restrict__ C, ull csC, ull n_points) {\n";
  o << "  const ull p0 = (ull)blockIdx.x * " << P << "ull;   // first point of this CTA\n";
  o << "  const unsigned np = (unsigned)(n_points - p0 < " << P << "ull ? n_points - p0 : " << P << "ull);\n";
  o << "  const unsigned pl = threadIdx.x / " << T << "u, " << (passes ? "i0" : "i") << " = (threadIdx.x % " << T << "u) * " << R
    << "u;   // first of this thread's rows\n";
  o << "  const bool live = pl < np;   // a thread of the last CTA's padding, or of the warp round-up, only helps staging\n";
  o << "  const ull p = p0 + (live ? pl : 0u);\n";
  std::ostringstream body;   // everything that depends on the row: emitted once, or inside the loop over passes
  for (uint32_t r = 0; r < R; ++r)
    body << "  const double* const X" << r << " = " << xn << " + (p * " << X->size() << "ull + kRowX[i + " << r << "u]) * " << (xc ? 2 : 1) << "ull;\n";
  std::ostringstream stage;
  if (y_smem) {
    stage << "  extern __shared__ __align__(16) double2 ys[];   // [point][" << y_row << " units]\n";
    stage << "  {\n    const double2* src = reinterpret_cast<const double2*>(" << yn << ") + p0 * " << y_units << "ull;   // the CTA's points: one contiguous range\n";
    stage << "    for (unsigned g = threadIdx.x; g < np * " << y_units << "u; g += " << gen.block << "u)\n";
    if (y_row == y_units)
      stage << "      ys[g] = __ldg(src + g);\n";
    else
      stage << "      ys[g + g / " << y_units << "u] = __ldg(src + g);\n";
    stage << "  }\n  __syncthreads();\n";
    stage << "  if (!live) return;\n";
    stage << "  const double* const Y = reinterpret_cast<const double*>(ys + pl * " << y_row << "u);\n";
  } else {
    stage << "  if (!live) return;\n";
    if (Y->batched())
      stage << "  const double* const Y = " << yn << " + p * " << Y->size() * (yc ? 2 : 1) << "ull;\n";
    else if (!Y->sparse())
      stage << "  const double* const Y = " << yn << ";\n";
  }
  for (uint32_t r = 0; r < R; ++r)
    body << "  double* const O" << r << " = C + (p * " << plan.size_out << "ull + kRowC[i + " << r << "u]) * " << (oc ? 2 : 1) << "ull;\n";
  // which (j, k) products exist (a sparse Y contributes its stored entries only), and which components of the row they use
  auto y_entry = [&](uint32_t j, size_t k) -> const double2* {
    if (!Y->sparse()) return nullptr;
    auto it = Y->entries.find(rs.col_y[j] + (uint64_t)ky[k]);
    return it == Y->entries.end() ? nullptr : &it->second;
  };
  std::vector<char> k_used(kx.size(), 0);
  for (uint32_t j = 0; j < rs.n_cols; ++j)
    for (size_t k = 0; k < kx.size(); ++k)
      if (!Y->sparse() || y_entry(j, k)) k_used[k] = 1;
  std::ostringstream loads, compute;
  std::ostringstream* out = &loads;
  // the row of X: 256-bit loads where two used neighbours share a 32-byte sector (every row base and the record even)
  bool even_rows = xc && X->size() % 2 == 0 && ((uintptr_t)X->dptr % 32) == 0 && jit_has_256bit_loads();
  for (uint64_t v : rs.row_x) even_rows = even_rows && v % 2 == 0;
  std::map<long long, size_t> k_of_off;
  for (size_t k = 0; k < kx.size(); ++k)
    if (k_used[k]) k_of_off[kx[k]] = k;
  for (uint32_t r = 0; r < R && !rs.k_outer; ++r) {
    std::vector<char> loaded(kx.size(), 0);
    const std::string xr = "x" + std::to_string(r) + "_", Xr = "X" + std::to_string(r);
    for (size_t k = 0; k < kx.size(); ++k) {
      if (!k_used[k] || loaded[k]) continue;
      const long long off = kx[k];
      auto mate = k_of_off.find(off ^ 1ll);
      if (even_rows && mate != k_of_off.end() && !loaded[mate->second] && mate->second != k) {
        const size_t lo = (off & 1ll) ? mate->second : k, hi = (off & 1ll) ? k : mate->second;
        (*out) << "  double2 " << xr << lo << ", " << xr << hi << "; ld32(" << Xr << " + " << (off & ~1ll) * 2 << "ull, " << xr << lo << ", " << xr << hi << ");\n";
        loaded[lo] = loaded[hi] = 1;
      } else {
        if (xc)
          (*out) << "  const double2 " << xr << k << " = __ldg(reinterpret_cast<const double2*>(" << Xr << " + " << off * 2 << "ull));\n";
        else
          (*out) << "  const double2 " << xr << k << " = make_double2(__ldg(" << Xr << " + " << off << "ull), 0.0);\n";
        loaded[k] = 1;
      }
    }
  }
  out = &compute;
  bool even_out = oc && plan.size_out % 2 == 0 && ((uintptr_t)res->dptr % 32) == 0;
  for (uint64_t v : rs.row_c) even_out = even_out && v % 2 == 0;
  std::vector<std::string> pending(R);
  std::vector<uint64_t> pending_c(R, 0);
  auto flush = [&](uint32_t r) {
    if (pending[r].empty()) return;
    (*out) << "  *reinterpret_cast<double2*>(O" << r << " + " << pending_c[r] * 2 << "ull) = " << pending[r] << ";\n";
    pending[r].clear();
  };
  auto y_value = [&](uint32_t j, size_t k) -> std::string {   // "" = the sparse fibre holds no entry here
    if (Y->sparse()) {
      const double2* v = y_entry(j, k);
      return v ? "make_double2(" + lit(v->x) + ", " + lit(v->y) + ")" : std::string();
    }
    const uint64_t f = rs.col_y[j] + (uint64_t)ky[k];
    if (y_smem)
      return yc ? "*reinterpret_cast<const double2*>(Y + " + std::to_string(f * 2) + "u)" : "make_double2(Y[" + std::to_string(f) + "u], 0.0)";
    return yc ? "__ldg(reinterpret_cast<const double2*>(Y + " + std::to_string(f * 2) + "ull))"
              : "make_double2(__ldg(Y + " + std::to_string(f) + "ull), 0.0)";
  };
  auto product = [&](const std::string& xv, uint32_t r, uint32_t j, size_t k) {
    const std::string acc = "r" + std::to_string(r) + "_" + std::to_string(j);
    (*out) << "    { const double2 t = cmul_rn(" << (rs.rows_from_a ? xv : "y") << ", " << (rs.rows_from_a ? "y" : xv) << "); " << acc << ".x = "
           << (kt.sign[k] ? "__dsub_rn" : "__dadd_rn") << "(" << acc << ".x, t.x); " << acc << ".y = "
           << (kt.sign[k] ? "__dsub_rn" : "__dadd_rn") << "(" << acc << ".y, t.y); }\n";
  };
  if (rs.k_outer) {
    // A real loop over k with the offsets in constant tables (uniform across the warp): straight-line code lets ptxas hoist
    // every load of a 512-component row above the sequential accumulation chain and spill 24 KB per thread.
    for (uint32_t j = 0; j < rs.n_cols; ++j)
      for (uint32_t r = 0; r < R; ++r) (*out) << "  double2 r" << r << "_" << j << " = make_double2(0.0, 0.0);\n";
    (*out) << "#pragma unroll 4\n  for (unsigned k = 0; k < " << kx.size() << "u; ++k) {\n";
    (*out) << "    const unsigned ox = kTabX[k], oy = kTabY[k];\n    const bool neg = kTabNeg[k] != 0;\n";
    for (uint32_t r = 0; r < R; ++r) {
      if (xc)
        (*out) << "    const double2 x" << r << " = __ldg(reinterpret_cast<const double2*>(X" << r << ") + ox);\n";
      else
        (*out) << "    const double2 x" << r << " = make_double2(__ldg(X" << r << " + ox), 0.0);\n";
    }
    for (uint32_t j = 0; j < rs.n_cols; ++j) {
      const uint64_t f = rs.col_y[j];
      std::string yv;
      if (y_smem)
        yv = yc ? "reinterpret_cast<const double2*>(Y)[" + std::to_string(f) + "u + oy]" : "make_double2(Y[" + std::to_string(f) + "u + oy], 0.0)";
      else
        yv = yc ? "__ldg(reinterpret_cast<const double2*>(Y) + " + std::to_string(f) + "u + oy)"
                : "make_double2(__ldg(Y + " + std::to_string(f) + "u + oy), 0.0)";
      (*out) << "    { const double2 y = " << yv << ";\n";
      for (uint32_t r = 0; r < R; ++r) {
        const std::string xv = "x" + std::to_string(r), acc = "r" + std::to_string(r) + "_" + std::to_string(j);
        // "-=" where the metric is negative: r - t and r + (-t) are the same IEEE operation
        (*out) << "      { const double2 t = cmul_rn(" << (rs.rows_from_a ? xv : "y") << ", " << (rs.rows_from_a ? "y" : xv) << "); " << acc
               << ".x = __dadd_rn(" << acc << ".x, neg ? -t.x : t.x); " << acc << ".y = __dadd_rn(" << acc << ".y, neg ? -t.y : t.y); }\n";
      }
      (*out) << "    }\n";
    }
    (*out) << "  }\n";
  }
  for (uint32_t j = 0; j < rs.n_cols; ++j) {
    if (!rs.k_outer)
      for (uint32_t r = 0; r < R; ++r) (*out) << "  double2 r" << r << "_" << j << " = make_double2(0.0, 0.0);\n";
    for (size_t k = 0; k < kx.size() && !rs.k_outer; ++k) {
      std::string yv;
      if (Y->sparse()) {
        const double2* v = y_entry(j, k);
        if (!v) continue;   // the sparse fibre holds no entry here
        yv = "make_double2(" + lit(v->x) + ", " + lit(v->y) + ")";
      } else {
        const uint64_t f = rs.col_y[j] + (uint64_t)ky[k];
        if (y_smem)
          yv = yc ? "*reinterpret_cast<const double2*>(Y + " + std::to_string(f * 2) + "u)" : "make_double2(Y[" + std::to_string(f) + "u], 0.0)";
        else
          yv = yc ? "__ldg(reinterpret_cast<const double2*>(Y + " + std::to_string(f * 2) + "ull))"
                  : "make_double2(__ldg(Y + " + std::to_string(f) + "ull), 0.0)";
      }
      (*out) << "  { const double2 y = " << yv << ";\n";
      for (uint32_t r = 0; r < R; ++r) {
        const std::string xv = "x" + std::to_string(r) + "_" + std::to_string(k), acc = "r" + std::to_string(r) + "_" + std::to_string(j);
        (*out) << "    { const double2 t = cmul_rn(" << (rs.rows_from_a ? xv : "y") << ", " << (rs.rows_from_a ? "y" : xv) << "); " << acc << ".x = "
          << (kt.sign[k] ? "__dsub_rn" : "__dadd_rn") << "(" << acc << ".x, t.x); " << acc << ".y = "
          << (kt.sign[k] ? "__dsub_rn" : "__dadd_rn") << "(" << acc << ".y, t.y); }\n";
      }
      (*out) << "  }\n";
    }
    const uint64_t c = rs.col_c[j];
    for (uint32_t r = 0; r < R; ++r) {
      const std::string acc = "r" + std::to_string(r) + "_" + std::to_string(j);
      if (!oc) {
        (*out) << "  O" << r << "[" << c << "ull] = " << acc << ".x;\n";
      } else if (even_out && !pending[r].empty() && c == pending_c[r] + 1 && pending_c[r] % 2 == 0) {
        (*out) << "  st32(O" << r << " + " << pending_c[r] * 2 << "ull, " << pending[r] << ", " << acc << ");\n";
        pending[r].clear();
      } else {
        flush(r);
        pending[r] = acc;
        pending_c[r] = c;
        if (!even_out) flush(r);
      }
    }
  }
  for (uint32_t r = 0; r < R; ++r) flush(r);
  if (!passes) {
    o << body.str() << loads.str() << stage.str() << compute.str();   // the row's loads are in flight while the CTA stages Y
  } else {
    o << stage.str();
    o << "  for (unsigned i = i0; i < " << rs.n_rows << "u; i += " << (uint64_t)T * R << "u) {   // passes over the rows of the point\n";
    o << body.str() << loads.str() << compute.str() << "  }\n";
  }
  o << "}\n";
  gen.source = o.str();
  return gen;
}

Generated generate(const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a, const spn_tensor* b,
                   const spn_tensor* res, const std::string& kname, bool wide_stores) {
  if (!register_form_fits(plan, kt, a, b)) return generate_rows(row_split(plan, kt, a, b), plan, kt, a, b, res, kname);
  const bool wide_loads = !getenv("SPN_NO_WIDE_LOADS");
  Generated gen;
  Staging sa = a->sparse() ? Staging() : staging_of(a, true), sb = b->sparse() ? Staging() : staging_of(b, true), sc = staging_of(res, false);
  const bool tma = tma_mode() && (sa.staged || sb.staged);
  gen.slab = sa.staged || sb.staged || sc.staged;
  // shared memory of one warp: [2 stages][A slab | B slab] then the result slab
  uint64_t stage_units = 0;
  if (sa.staged) { sa.offset = stage_units; stage_units += 32 * sa.row; }
  if (sb.staged) { sb.offset = stage_units; stage_units += 32 * sb.row; }
  // Pipeline shape.  Every resident warp has (n_stages - 1) slabs on their way while it computes on one; with
  // n_stages = 1 a warp loads, then computes, and the overlap comes from the other warps of the SM.  Measured
  // (profiles/r02_pairwise_sweep.txt, 2^20 points, stages 1-4 x 2/4/8 warps per CTA): one stage is never slower and up to
  // 1.5x faster than 3-4 stages on the 256-1024 B operands — shared memory spent on extra stages costs resident
  // warps, and resident warps are what hides the latency here.  SPN_PAIR_STAGES=2..4 keeps the deeper rings reachable.
  const uint64_t out_units = (sc.staged ? 32 * sc.row : 0) + (tma_mode() ? 2 : 0);   // + the mbarriers of the TMA variant
  const uint64_t budget = 216 * 1024 / 16;   // 16-byte units of shared memory per SM the kernel may plan with
  const int forced_stages = getenv("SPN_PAIR_STAGES") ? std::max(1, std::min(4, atoi(getenv("SPN_PAIR_STAGES")))) : 0;
  int n_stages = forced_stages ? forced_stages : 1;
  int warps = kSlabWarps;
  uint64_t warp_units = (uint64_t)n_stages * stage_units + out_units;
  if (gen.slab) {
    if (!forced_stages)
      while (n_stages > 1 && ((uint64_t)n_stages * stage_units + out_units) * 8 > budget) --n_stages;
    warp_units = (uint64_t)n_stages * stage_units + out_units;
    while (warps > 1 && warp_units * (uint64_t)warps > budget) warps /= 2;
    if (warp_units > budget) {  // does not fit at all: direct kernel
      gen.slab = false;
      sa = sb = sc = Staging();
    }
  }
  sc.offset = (uint64_t)n_stages * stage_units;
  gen.block = gen.slab ? 32 * warps : kPairBlock;
  gen.smem_bytes = gen.slab ? (size_t)(warp_units * (uint64_t)warps * 16) : 0;

  std::ostringstream o;
  o << "// generated by spenso_b200 (pairwise_jit.cu): one pairwise contraction, thread = point, reference rounding\n";
  if (gen.slab)
    o << "// point-major tensors staged through shared memory in warp-owned slabs of 32 points (" << n_stages << " stage(s), "
      << warp_units * 16 << " B per warp, " << warps << " warps per CTA)\n";
  o << "typedef unsigned long long ull;\n";
  o << "__device__ __forceinline__ double2 ld16(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }\n";
  o << "__device__ __forceinline__ double2 ld8(const double* p) { return make_double2(__ldg(p), 0.0); }\n";
  o << "__device__ __forceinline__ void ld32(const double* p, double2& a, double2& b) {\n"
       "  asm(\"ld.global.nc" << (getenv("SPN_PAIR_NO_HINT") ? "" : ".L1::no_allocate")
    << ".v4.f64 {%0,%1,%2,%3}, [%4];\" : \"=d\"(a.x), \"=d\"(a.y), \"=d\"(b.x), \"=d\"(b.y) : \"l\"(p));\n}\n";   // whole sector, read once
  o << "__device__ __forceinline__ void st32(double* p, double2 a, double2 b) {\n"
       "  asm volatile(\"st.global.v4.f64 [%0], {%1,%2,%3,%4};\" :: \"l\"(p), \"d\"(a.x), \"d\"(a.y), \"d\"(b.x), \"d\"(b.y) : \"memory\");\n}\n";
  o << "__device__ __forceinline__ void cp16(double2* s, const double2* g) {\n"
       "  asm volatile(\"cp.async.cg.shared.global [%0], [%1], 16;\" :: \"r\"((unsigned)__cvta_generic_to_shared(s)), \"l\"(g) : \"memory\");\n}\n";
  o << "__device__ __forceinline__ double2 cmul_rn(double2 a, double2 b) {\n"
       "  return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));\n}\n";
  o << "extern \"C\" __global__ void __launch_bounds__(" << gen.block << ") " << kname
    << "(const double* __restrict__ A, ull csA, const double* __restrict__ B, ull csB, double* __restrict__ C, ull csC, ull n_points) {\n";
  if (gen.slab) {
    o << "  extern __shared__ __align__(16) double2 smem[];\n";
    o << "  const unsigned lane = threadIdx.x & 31u;\n";
    o << "  double2* const ws = smem + (threadIdx.x >> 5) * " << warp_units << "u;   // this warp's slabs\n";
    if (tma) {
      o << "  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(ws + " << (warp_units - 2) << "u);   // one mbarrier per stage\n";
      o << "  if (lane == 0) {\n";
      o << "    for (unsigned s = 0; s < " << n_stages << "u; ++s)\n";
      o << "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" :: \"r\"((unsigned)__cvta_generic_to_shared(bars + s)) : \"memory\");\n";
      o << "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n";
      o << "  }\n  __syncwarp();\n  unsigned phases = 0;\n";
    }
    o << "  const ull n_slabs = (n_points + 31ull) >> 5;\n";
    o << "  const ull n_warps = (ull)gridDim.x * " << warps << "ull;\n";
    o << "  ull slab = (ull)blockIdx.x * " << warps << "ull + (threadIdx.x >> 5);\n";
    // issue(slab, stage): cp.async the operand slabs of `slab` into stage `stage`
    o << "  auto issue = [&](ull sl, unsigned stage) {\n";
    o << "    const ull p0 = sl << 5;\n";
    o << "    const unsigned np = (unsigned)(n_points - p0 < 32ull ? n_points - p0 : 32ull);\n";
    if (tma) {
      // one elected lane arms the stage's mbarrier with the byte count and launches the bulk copies
      o << "    if (lane == 0) {\n";
      o << "      const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + stage);\n";
      o << "      unsigned bytes = 0;\n";
      for (auto pr : {std::make_pair(&sa, 'A'), std::make_pair(&sb, 'B')})
        if (pr.first->staged) o << "      bytes += np * " << pr.first->units * 16 << "u;\n";
      o << "      asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");   // earlier generic reads of this stage are done\n";
      o << "      asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\" :: \"r\"(bar), \"r\"(bytes) : \"memory\");\n";
      for (auto pr : {std::make_pair(&sa, 'A'), std::make_pair(&sb, 'B')}) {
        const Staging& s = *pr.first;
        if (!s.staged) continue;
        o << "      asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\" :: "
             "\"r\"((unsigned)__cvta_generic_to_shared(ws + stage * " << stage_units << "u + " << s.offset << "u)), \"l\"(reinterpret_cast<const double2*>("
          << pr.second << ") + p0 * " << s.units << "ull), \"r\"(np * " << s.units * 16 << "u), \"r\"(bar) : \"memory\");\n";
      }
      o << "    }\n";
    }
    for (auto pr : {std::make_pair(&sa, 'A'), std::make_pair(&sb, 'B')}) {
      const Staging& s = *pr.first;
      if (!s.staged || tma) continue;
      o << "    {\n      const double2* src = reinterpret_cast<const double2*>(" << pr.second << ") + p0 * " << s.units << "ull;\n";
      o << "      double2* dst = ws + stage * " << stage_units << "u + " << s.offset << "u;\n";
      o << "      const unsigned lim = np * " << s.units << "u;\n";
      o << "#pragma unroll\n      for (unsigned j = 0; j < " << s.units << "u; ++j) {\n";
      o << "        const unsigned g = lane + 32u * j;\n";
      if (s.row == s.units)
        o << "        if (g < lim) cp16(dst + g, src + g);\n";
      else
        o << "        if (g < lim) cp16(dst + g + g / " << s.units << "u, src + g);\n";  // row stride = units + 1
      o << "      }\n    }\n";
    }
    o << "  };\n";
    // prologue: the first n_stages - 1 slabs of this warp (one commit group per slab, empty groups keep the count)
    if (n_stages == 1) {
      o << "  if (slab < n_slabs) issue(slab, 0u);\n";
      o << "  asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n";
    } else {
      o << "#pragma unroll\n  for (unsigned d = 0; d < " << (n_stages - 1) << "u; ++d) {\n";
      o << "    if (slab + d * n_warps < n_slabs) issue(slab + d * n_warps, d);\n";
      o << "    asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n  }\n";
    }
    o << "  for (unsigned stage = 0; slab < n_slabs; slab += n_warps"
      << (n_stages > 1 ? ", stage = (stage + 1u == " + std::to_string(n_stages) + "u ? 0u : stage + 1u)" : std::string()) << ") {\n";
    if (n_stages > 1) {
      // refill the stage consumed in the previous iteration with the slab n_stages - 1 ahead
      o << "    {\n      const unsigned fill = stage == 0u ? " << (n_stages - 1) << "u : stage - 1u;\n";
      o << "      if (slab + " << (n_stages - 1) << "ull * n_warps < n_slabs) issue(slab + " << (n_stages - 1) << "ull * n_warps, fill);\n";
      o << "      asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n    }\n";
      o << "    asm volatile(\"cp.async.wait_group " << (n_stages - 1) << ";\" ::: \"memory\");\n";
    } else {
      o << "    asm volatile(\"cp.async.wait_group 0;\" ::: \"memory\");\n";
    }
    if (tma) {
      o << "    {\n      const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + stage), ph = (phases >> stage) & 1u;\n";
      o << "      unsigned done = 0;\n";
      o << "      while (!done)\n";
      o << "        asm volatile(\"{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\" : \"=r\"(done) : \"r\"(bar), \"r\"(ph) : \"memory\");\n";
      o << "      phases ^= 1u << stage;\n    }\n";
    }
    o << "    __syncwarp();\n";
    o << "    const ull p0 = slab << 5;\n    const ull p = p0 + lane;\n";
    o << "    const unsigned np = (unsigned)(n_points - p0 < 32ull ? n_points - p0 : 32ull);\n";
    if (sa.staged) o << "    const double2* const sA = ws + stage * " << stage_units << "u + " << sa.offset << "u + lane * " << sa.row << "u;\n";
    if (sb.staged) o << "    const double2* const sB = ws + stage * " << stage_units << "u + " << sb.offset << "u + lane * " << sb.row << "u;\n";
    if (sc.staged) o << "    double2* const sC = ws + " << sc.offset << "u + lane * " << sc.row << "u;\n";
    o << "    if (lane < np) {\n";
  } else {
    o << "  for (ull p = (ull)blockIdx.x * " << gen.block << " + threadIdx.x; p < n_points; p += (ull)gridDim.x * " << gen.block
      << ") {\n";
  }
  Operand oa{a, 'A'}, ob{b, 'B'};
  // which components are needed, in first-use order
  std::vector<char> seen_a(a->size(), 0), seen_b(b->size(), 0);
  // components used anywhere (to pair neighbours into one 256-bit load: a full 32-byte sector per instruction)
  std::vector<char> used_a(a->size(), 0), used_b(b->size(), 0);
  {
    std::vector<uint32_t> ix(plan.order_out, 0);
    for (uint64_t e = 0; e < plan.size_out; ++e) {
      uint64_t oa_ = 0, ob_ = 0;
      for (uint32_t r = 0; r < plan.order_out; ++r) {
        oa_ += (uint64_t)ix[r] * plan.out_stride_a[r];
        ob_ += (uint64_t)ix[r] * plan.out_stride_b[r];
      }
      for (size_t k = 0; k < kt.off_a.size(); ++k) {
        const uint64_t fa = oa_ + (uint64_t)kt.off_a[k], fb = ob_ + (uint64_t)kt.off_b[k];
        if (a->sparse() && !a->entries.count(fa)) continue;
        if (b->sparse() && !b->entries.count(fb)) continue;
        if (!a->sparse()) used_a[fa] = 1;
        if (!b->sparse()) used_b[fb] = 1;
      }
      for (uint32_t r = plan.order_out; r-- > 0;) {
        if (++ix[r] < plan.out_dim[r]) break;
        ix[r] = 0;
      }
    }
  }
  const std::string ind = gen.slab ? "      " : "    ";
  auto use = [&](const Operand& op, const Staging& st, std::vector<char>& seen, const std::vector<char>& used, uint64_t f) -> std::string {
    const std::string pre(1, (char)(op.name + 32));  // a12 / b3
    const std::string nm = pre + std::to_string(f);
    if (seen[f]) return nm;
    if (st.staged) {
      seen[f] = 1;
      if (op.t->dtype == SPN_C64)
        o << ind << "const double2 " << nm << " = s" << op.name << "[" << f << "];\n";
      else
        o << ind << "const double2 " << nm << " = make_double2(reinterpret_cast<const double*>(s" << op.name << ")[" << f << "], 0.0);\n";
      return nm;
    }
    const bool pairable = wide_loads && jit_has_256bit_loads() && op.t->batched() && op.t->dtype == SPN_C64 &&
                          op.t->layout == SPN_POINT_MAJOR && op.t->size() % 2 == 0 && ((uintptr_t)op.t->dptr % 32) == 0 &&
                          used[f ^ 1ull];
    if (pairable) {
      const uint64_t base = f & ~1ull;
      seen[base] = seen[base + 1] = 1;
      o << ind << "double2 " << pre << base << ", " << pre << base + 1 << "; ld32(" << addr(op, base) << ", " << pre << base << ", "
        << pre << base + 1 << ");\n";
    } else {
      seen[f] = 1;
      o << ind << "const double2 " << nm << " = " << (op.t->dtype == SPN_C64 ? "ld16(" : "ld8(") << addr(op, f) << ");\n";
    }
    return nm;
  };
  auto literal = [&](const spn_tensor* t, uint64_t f) {
    auto it = t->entries.find(f);
    return "make_double2(" + lit(it->second.x) + ", " + lit(it->second.y) + ")";
  };
  const bool out_c = res->dtype == SPN_C64;
  std::vector<uint32_t> idx(plan.order_out, 0);
  std::string pending;  // first half of a 256-bit store
  for (uint64_t e = 0; e < plan.size_out; ++e) {
    uint64_t offa = 0, offb = 0;
    for (uint32_t r = 0; r < plan.order_out; ++r) {
      offa += (uint64_t)idx[r] * plan.out_stride_a[r];
      offb += (uint64_t)idx[r] * plan.out_stride_b[r];
    }
    std::ostringstream body;
    body << ind << "double2 r" << e << " = make_double2(0.0, 0.0);\n";
    for (size_t k = 0; k < kt.off_a.size(); ++k) {
      const uint64_t fa = offa + (uint64_t)kt.off_a[k], fb = offb + (uint64_t)kt.off_b[k];
      if (a->sparse() && !a->entries.count(fa)) continue;  // the sparse fibre holds no entry here
      if (b->sparse() && !b->entries.count(fb)) continue;
      std::string xa = a->sparse() ? literal(a, fa) : use(oa, sa, seen_a, used_a, fa);
      std::string xb = b->sparse() ? literal(b, fb) : use(ob, sb, seen_b, used_b, fb);
      body << ind << "{ const double2 t = cmul_rn(" << xa << ", " << xb << "); r" << e << ".x = "
           << (kt.sign[k] ? "__dsub_rn" : "__dadd_rn") << "(r" << e << ".x, t.x); r" << e << ".y = "
           << (kt.sign[k] ? "__dsub_rn" : "__dadd_rn") << "(r" << e << ".y, t.y); }\n";
    }
    o << body.str();
    // store
    const bool point_major = res->layout == SPN_POINT_MAJOR;
    if (sc.staged) {
      if (out_c)
        o << ind << "sC[" << e << "] = r" << e << ";\n";
      else
        o << ind << "reinterpret_cast<double*>(sC)[" << e << "] = r" << e << ".x;\n";
    } else if (!out_c) {
      if (point_major)
        o << ind << "C[p * " << plan.size_out << "ull + " << e << "ull] = r" << e << ".x;\n";
      else
        o << ind << "C[" << e << "ull * csC + p] = r" << e << ".x;\n";
    } else if (point_major && wide_stores && plan.size_out % 2 == 0) {
      if (e % 2 == 0) {
        pending = "r" + std::to_string(e);
      } else {
        o << ind << "st32(C + (p * " << plan.size_out << "ull + " << (e - 1) << "ull) * 2, " << pending << ", r" << e << ");\n";
      }
    } else if (point_major) {
      o << ind << "reinterpret_cast<double2*>(C)[p * " << plan.size_out << "ull + " << e << "ull] = r" << e << ";\n";
    } else {
      o << ind << "reinterpret_cast<double2*>(C)[" << e << "ull * csC + p] = r" << e << ";\n";
    }
    for (uint32_t r = plan.order_out; r-- > 0;) {
      if (++idx[r] < plan.out_dim[r]) break;
      idx[r] = 0;
    }
  }
  if (gen.slab) {
    o << "    }\n";
    if (sc.staged) {
      // the warp's result slab is a contiguous range of the output: 16-byte stores, consecutive lanes = consecutive units
      o << "    __syncwarp();\n";
      o << "    {\n      double2* dst = reinterpret_cast<double2*>(C) + p0 * " << sc.units << "ull;\n";
      o << "      const double2* src = ws + " << sc.offset << "u;\n";
      o << "      const unsigned lim = np * " << sc.units << "u;\n";
      o << "#pragma unroll\n      for (unsigned j = 0; j < " << sc.units << "u; ++j) {\n";
      o << "        const unsigned g = lane + 32u * j;\n";
      if (sc.row == sc.units)
        o << "        if (g < lim) dst[g] = src[g];\n";
      else
        o << "        if (g < lim) dst[g] = src[g + g / " << sc.units << "u];\n";
      o << "      }\n    }\n";
    }
    o << "    __syncwarp();   // every lane is done with this stage before it is refilled\n";
    if (n_stages == 1) {
      o << "    if (slab + n_warps < n_slabs) issue(slab + n_warps, 0u);\n";
      o << "    asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n";
    }
  }
  o << "  }\n}\n";
  gen.source = o.str();
  return gen;
}

}  // namespace

bool pair_jit_applicable(const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a, const spn_tensor* b,
                         bool res_batched, uint64_t n_points) {
  if (getenv("SPN_NO_PAIR_JIT")) return false;
  // specialising costs one NVRTC compile (~0.3 s, cached on disk): only worth it for batches that amortise it
  if (!res_batched || n_points < 2048) return false;
  if (a->sparse() && b->sparse()) return false;
  if (register_form_fits(plan, kt, a, b)) return true;
  // too large for one thread: one thread per (point, row of a batched operand) if the shape allows
  return plan.size_out * (uint64_t)kt.off_a.size() != 0 && row_split(plan, kt, a, b).ok;
}

PairJit pair_jit_build(spn_ctx* ctx, const spn_contract_plan& plan, const KTables& kt, const spn_tensor* a,
                       const spn_tensor* b, const spn_tensor* res) {
  const bool wide = ((uintptr_t)res->dptr % 32) == 0;
  // kernel identity: the plan (plain data), what the operands are, and the entries of a sparse operand
  std::string key(reinterpret_cast<const char*>(&plan), sizeof plan);
  for (const spn_tensor* t : {a, b, (const spn_tensor*)res}) {
    key.push_back((char)t->storage);
    key.push_back((char)t->dtype);
    key.push_back((char)t->layout);
    key.push_back(stageable(t) ? 'S' : 'd');
  }
  key.push_back(wide ? 'w' : 'n');
  key.push_back(tma_mode() ? 'T' : 'L');
  key.push_back(getenv("SPN_PAIR_NO_ROWS") ? 'g' : (getenv("SPN_PAIR_ROWS_NO_SMEM") ? 'l' : 'r'));
  if (const char* e = getenv("SPN_PAIR_ROWS_PER_THREAD")) key += e;
  key.append(reinterpret_cast<const char*>(kt.off_a.data()), kt.off_a.size() * sizeof(long long));
  key.append(reinterpret_cast<const char*>(kt.off_b.data()), kt.off_b.size() * sizeof(long long));
  key.append(reinterpret_cast<const char*>(kt.sign.data()), kt.sign.size());
  for (const spn_tensor* t : {a, b}) key.push_back(((uintptr_t)t->dptr % 32) == 0 ? 'A' : 'u');
  for (const spn_tensor* t : {a, b})
    if (t->sparse())
      for (auto& kv : t->entries) {
        key.append(reinterpret_cast<const char*>(&kv.first), sizeof kv.first);
        key.append(reinterpret_cast<const char*>(&kv.second), sizeof kv.second);
      }
  const uint64_t h = fnv1a(key);
  PairJit pj;
  auto it = ctx->pair_kernels.find(h);
  if (it != ctx->pair_kernels.end()) return it->second;
  auto k = std::make_shared<JitKernel>();
  char nm[64];
  std::snprintf(nm, sizeof nm, "spn_pair_contract_%016llx", (unsigned long long)h);
  k->name = nm;
  Generated g = generate(plan, kt, a, b, res, k->name, wide);
  k->source = g.source;
  k->smem_bytes = g.smem_bytes;
  jit_compile(*k);
  jit_load(*k, g.block);
  pj.k = k;
  pj.block = g.block;
  pj.slab = g.slab;
  pj.rows = g.rows;
  pj.points_per_cta = g.points_per_cta;
  ctx->pair_kernels.emplace(h, pj);
  return pj;
}

// Inspection without a device: the source of the kernel a batched x batched contraction of these structures would
// run (compiled with NVRTC so that a code-generation error shows up where no GPU exists).  Empty = generic kernel.
std::string pair_jit_source(const Structure& sa, int dtype_a, int layout_a, const Structure& sb, int dtype_b, int layout_b) {
  spn_tensor a, b, r;
  auto fake = [](spn_tensor& t, const Structure& st, int dtype, int layout) {
    t.st = st;
    t.storage = SPN_STORAGE_BATCHED;
    t.dtype = dtype;
    t.layout = layout;
    t.n_points = 4096;
    t.dptr = reinterpret_cast<void*>(uintptr_t(1) << 20);  // never dereferenced: alignment only
    t.owns = false;
  };
  fake(a, sa, dtype_a, layout_a);
  fake(b, sb, dtype_b, layout_b);
  spn_contract_plan plan = plan_contract(sa, sb);
  KTables kt = build_k_tables(plan, false);
  if (!pair_jit_applicable(plan, kt, &a, &b, true, 4096)) return std::string();
  fake(r, structure_of_plan_result(plan), (dtype_a == SPN_F64 && dtype_b == SPN_F64) ? SPN_F64 : SPN_C64, layout_a);
  JitKernel k;
  k.name = "spn_pair_contract_inspect";
  Generated g = generate(plan, kt, &a, &b, &r, k.name, true);
  k.source = g.source;
  jit_compile(k);
  return g.source;
}

// The same inspection for a SHARED SPARSE operand (its stored entries become part of the code) against a batched one.
std::string pair_jit_source_sparse(const Structure& ss, const std::map<uint64_t, double2>& entries, const Structure& sd,
                                   int dtype_d, int layout_d, bool sparse_first) {
  spn_tensor s, d, r;
  s.st = ss;
  s.storage = SPN_STORAGE_SHARED_SPARSE;
  s.dtype = SPN_C64;
  s.layout = SPN_POINT_MAJOR;
  s.n_points = 0;
  s.entries = entries;
  s.owns = false;
  auto fake = [](spn_tensor& t, const Structure& st, int dtype, int layout) {
    t.st = st;
    t.storage = SPN_STORAGE_BATCHED;
    t.dtype = dtype;
    t.layout = layout;
    t.n_points = 4096;
    t.dptr = reinterpret_cast<void*>(uintptr_t(1) << 20);  // never dereferenced: alignment only
    t.owns = false;
  };
  fake(d, sd, dtype_d, layout_d);
  const spn_tensor* a = sparse_first ? &s : &d;
  const spn_tensor* b = sparse_first ? &d : &s;
  spn_contract_plan plan = plan_contract(a->st, b->st);
  KTables kt = build_k_tables(plan, false);
  if (!pair_jit_applicable(plan, kt, a, b, true, 4096)) return std::string();
  fake(r, structure_of_plan_result(plan), SPN_C64, layout_d);
  JitKernel k;
  k.name = "spn_pair_contract_inspect";
  Generated g = generate(plan, kt, a, b, &r, k.name, true);
  k.source = g.source;
  jit_compile(k);
  return g.source;
}

void pair_jit_launch(spn_ctx* ctx, const PairJit& pj, const spn_tensor* a, const spn_tensor* b, spn_tensor* res,
                     uint64_t n_points) {
  JitKernel& k = *pj.k;
  // CTAs that cover the batch — row-parallel kernels: whole points per CTA; staged kernels: one warp per slab of 32
  // points; direct kernels: one thread per point
  uint64_t need;
  if (pj.points_per_cta)
    need = (n_points + pj.points_per_cta - 1) / pj.points_per_cta;
  else if (pj.slab)
    need = ((n_points + 31) / 32 + (uint64_t)(pj.block / 32) - 1) / (uint64_t)(pj.block / 32);
  else
    need = (n_points + (uint64_t)pj.block - 1) / (uint64_t)pj.block;
  // Hardware-scheduled grid: one slab per warp (one point per thread), as many CTAs as that takes.  The kernels keep
  // their grid-stride loop, but a persistent launch — SMs x resident CTAs, every warp walking its slabs in lockstep with
  // all others — costs 10-22 % on these streams (profiles/r02_pairwise_grid.txt, tools/probes/write_stream_probe.cu: a
  // plain fill drops from 6.87 to 5.93 TB/s when it is made persistent).  SPN_PAIR_PERSISTENT=1 restores it for A/B runs.
  const uint64_t cap = getenv("SPN_PAIR_PERSISTENT") ? (uint64_t)ctx->sm_count * (uint64_t)k.blocks_per_sm : 0x7fffffffull;
  if (pj.points_per_cta && need > 0x7fffffffull)   // the row-parallel kernels have no grid-stride loop
    throw Error(SPN_ERR_INVALID, "batch too large for one launch of the row-parallel kernel");
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, pj.points_per_cta ? 0x7fffffffull : cap));
  const double* pa = static_cast<const double*>(a->dptr);
  const double* pb = static_cast<const double*>(b->dptr);
  double* pc = static_cast<double*>(res->dptr);
  unsigned long long csa = a->n_points, csb = b->n_points, csc = res->n_points, np = n_points;
  void* args[] = {&pa, &csa, &pb, &csb, &pc, &csc, &np};
  cuda_check(cudaLaunchKernel((const void*)k.kernel, dim3(grid), dim3((unsigned)pj.block), args, k.smem_bytes, ctx->stream),
             "launch specialised pairwise kernel");
  ctx->launches++;
}

}  // namespace spn
