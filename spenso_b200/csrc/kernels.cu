// Hand-written CUDA kernels (sm_100a) for the pairwise contraction path.
//
// Reference semantics reproduced (paths relative to /root/reference/spenso/src/):
//   contraction/singlecontract.rs:25-196, multicontract.rs:25-200, exteriorproduct.rs:19-153
//   (result[o] accumulated sequentially over the contracted index in ascending order with one
//   accumulator, "-=" where the representation's metric is negative), algebra/complex/mul.rs:16-21
//   (complex product without fused multiply-add), algebra/{add_assign,sub_assign,neg,scalar_mul}.rs,
//   contraction/trace.rs.
//
// Arithmetic uses the explicit _rn intrinsics so nvcc cannot contract a*b-c*d into an FMA: the
// pairwise path reproduces the reference's rounding exactly (same operation order, same IEEE ops).
// These kernels stream the point batch once: they are HBM-bound, so grids are sized in multiples of
// the SM count, all global accesses are 128-bit (one Complex<f64>) and the small shared operand
// (CSR of a sparse library tensor) is staged in shared memory.
#include "kernels.cuh"

#include <algorithm>

namespace spn {

namespace {

typedef unsigned long long ull;
typedef long long ll;

__device__ __forceinline__ double2 cmul_rn(double2 a, double2 b) {
  return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                      __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ void cacc_rn(double2& acc, double2 p, bool neg) {
  if (neg) {
    acc.x = __dsub_rn(acc.x, p.x);
    acc.y = __dsub_rn(acc.y, p.y);
  } else {
    acc.x = __dadd_rn(acc.x, p.x);
    acc.y = __dadd_rn(acc.y, p.y);
  }
}
template <bool CPLX>
__device__ __forceinline__ double2 load_elem(const void* base, ll idx) {
  if (CPLX) return __ldg(reinterpret_cast<const double2*>(base) + idx);
  // f64 is upgraded to (x, 0.0) before the full complex product (upgrading_arithmetic.rs:1285-1293)
  return make_double2(__ldg(reinterpret_cast<const double*>(base) + idx), 0.0);
}
__device__ __forceinline__ double2 load_elem_rt(const void* base, ll idx, int cplx) {
  return cplx ? load_elem<true>(base, idx) : load_elem<false>(base, idx);
}
__device__ __forceinline__ void store_elem(void* base, ll idx, double2 v, int cplx) {
  if (cplx)
    reinterpret_cast<double2*>(base)[idx] = v;
  else
    reinterpret_cast<double*>(base)[idx] = v.x;
}

__device__ __forceinline__ void split_index(ull t, ull size_out, ull n_points, bool point_fastest, ull& p, ull& o) {
  if (point_fastest) {
    p = t % n_points;
    o = t / n_points;
  } else {
    o = t % size_out;
    p = t / size_out;
  }
}

// ---- K1/K5/K7: dense x dense ---------------------------------------------------------------
template <bool AC, bool BC>
__global__ void __launch_bounds__(256)
contract_dense_kernel(const __grid_constant__ DevOutMap map, DevOperand a, DevOperand b, DevOut out,
                      const long long* __restrict__ koff_a, const long long* __restrict__ koff_b,
                      const unsigned char* __restrict__ ksign, ull n_k, ull n_points, int point_fastest) {
  const ull total = map.size_out * n_points;
  for (ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (ull)gridDim.x * blockDim.x) {
    ull p, o;
    split_index(t, map.size_out, n_points, point_fastest != 0, p, o);
    ll offa = (ll)p * a.ps, offb = (ll)p * b.ps;
    ull r = o;
    for (int d = map.order - 1; d >= 0; --d) {
      unsigned i = (unsigned)(r % map.dim[d]);
      r /= map.dim[d];
      offa += (ll)i * map.sa[d] * a.fs;
      offb += (ll)i * map.sb[d] * b.fs;
    }
    double2 acc = make_double2(0.0, 0.0);
    for (ull k = 0; k < n_k; ++k) {
      double2 x = load_elem<AC>(a.ptr, offa + __ldg(koff_a + k) * a.fs);
      double2 y = load_elem<BC>(b.ptr, offb + __ldg(koff_b + k) * b.fs);
      cacc_rn(acc, cmul_rn(x, y), __ldg(ksign + k) != 0);
    }
    store_elem(out.ptr, (ll)p * out.ps + (ll)o * out.fs, acc, out.is_complex);
  }
}

// The same contraction without per-dimension index arithmetic: the operand base offsets of every result element come
// from two host-built tables (toa[o], tob[o]: the plan-cache entry owns them) and the contracted-loop tables sit in
// shared memory.  One thread per (point, result element), element fastest for point-major results (consecutive threads
// = consecutive result elements: coalesced stores), point fastest for component-major ones; one division per thread
// (32-bit when the element count allows) instead of 2 x order 64-bit divisions.
template <bool AC, bool BC, bool PF>
__global__ void __launch_bounds__(256)
contract_dense_tab_kernel(DevOperand a, DevOperand b, DevOut out, const ll* __restrict__ toa, const ll* __restrict__ tob,
                          const ll* __restrict__ koff_a, const ll* __restrict__ koff_b,
                          const unsigned char* __restrict__ ksign, unsigned n_k, unsigned size_out, ull n_points,
                          int stage_k) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ll* ka = koff_a;
  const ll* kb = koff_b;
  const unsigned char* ks = ksign;
  if (stage_k) {
    ll* sa = reinterpret_cast<ll*>(smem);
    ll* sb = sa + n_k;
    unsigned char* ss = reinterpret_cast<unsigned char*>(sb + n_k);
    for (unsigned k = threadIdx.x; k < n_k; k += blockDim.x) {
      sa[k] = __ldg(koff_a + k) * a.fs;
      sb[k] = __ldg(koff_b + k) * b.fs;
      ss[k] = __ldg(ksign + k);
    }
    __syncthreads();
    ka = sa;
    kb = sb;
    ks = ss;
  }
  const ll fsa = stage_k ? 1 : a.fs, fsb = stage_k ? 1 : b.fs;
  const ull total = (ull)size_out * n_points;
  const bool narrow = total <= 0xffffffffull;   // uniform: 32-bit division
  const ull inner = PF ? n_points : (ull)size_out;
  for (ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (ull)gridDim.x * blockDim.x) {
    ull hi, lo;   // t = hi * inner + lo
    if (narrow) {
      const unsigned q = (unsigned)t / (unsigned)inner;
      hi = q;
      lo = (unsigned)t - q * (unsigned)inner;
    } else {
      hi = t / inner;
      lo = t - hi * inner;
    }
    const ull p = PF ? lo : hi, o = PF ? hi : lo;
    const ll offa = (ll)p * a.ps + __ldg(toa + o) * a.fs, offb = (ll)p * b.ps + __ldg(tob + o) * b.fs;
    double2 acc = make_double2(0.0, 0.0);
    for (unsigned k = 0; k < n_k; ++k) {
      double2 x = load_elem<AC>(a.ptr, offa + ka[k] * fsa);
      double2 y = load_elem<BC>(b.ptr, offb + kb[k] * fsb);
      cacc_rn(acc, cmul_rn(x, y), ks[k] != 0);
    }
    store_elem(out.ptr, (ll)p * out.ps + (ll)o * out.fs, acc, out.is_complex);
  }
}

// ---- K2: dense (per point) x sparse (shared, CSR staged in shared memory) ---------------------
template <bool DC, bool STAGE>
__global__ void __launch_bounds__(256)
contract_csr_kernel(DevOperand dense, DevOut out, const int* __restrict__ row_ptr, const long long* __restrict__ col_off,
                    const double2* __restrict__ val, const unsigned char* __restrict__ sign, int dense_is_first,
                    ull size_out, ull nnz, ull n_points, int point_fastest) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int* s_row = row_ptr;
  const ll* s_col = col_off;
  const double2* s_val = val;
  const unsigned char* s_sign = sign;
  if (STAGE) {
    double2* sv = reinterpret_cast<double2*>(smem);
    ll* sc = reinterpret_cast<ll*>(sv + nnz);
    int* sr = reinterpret_cast<int*>(sc + nnz);
    unsigned char* ss = reinterpret_cast<unsigned char*>(sr + (size_out + 1));
    for (ull i = threadIdx.x; i < nnz; i += blockDim.x) {
      sv[i] = __ldg(val + i);
      sc[i] = __ldg(col_off + i);
      ss[i] = __ldg(sign + i);
    }
    for (ull i = threadIdx.x; i <= size_out; i += blockDim.x) sr[i] = __ldg(row_ptr + i);
    __syncthreads();
    s_row = sr;
    s_col = sc;
    s_val = sv;
    s_sign = ss;
  }
  const ull total = size_out * n_points;
  for (ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (ull)gridDim.x * blockDim.x) {
    ull p, o;
    split_index(t, size_out, n_points, point_fastest != 0, p, o);
    const ll base = (ll)p * dense.ps;
    double2 acc = make_double2(0.0, 0.0);
    const int e1 = s_row[o + 1];
    for (int e = s_row[o]; e < e1; ++e) {
      double2 x = load_elem<DC>(dense.ptr, base + s_col[e] * dense.fs);
      double2 v = s_val[e];
      double2 pr = dense_is_first ? cmul_rn(x, v) : cmul_rn(v, x);
      cacc_rn(acc, pr, s_sign[e] != 0);
    }
    store_elem(out.ptr, (ll)p * out.ps + (ll)o * out.fs, acc, out.is_complex);
  }
}

// ---- element-wise ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
elementwise_kernel(int op, DevOperand a, DevOperand b, DevOut out, double2 s, ull size, ull n_points,
                   int point_fastest) {
  const ull total = size * n_points;
  for (ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (ull)gridDim.x * blockDim.x) {
    ull p, f;
    split_index(t, size, n_points, point_fastest != 0, p, f);
    double2 x = load_elem_rt(a.ptr, (ll)p * a.ps + (ll)f * a.fs, a.is_complex);
    double2 r;
    switch (op) {
      case EW_ADD: {
        double2 y = load_elem_rt(b.ptr, (ll)p * b.ps + (ll)f * b.fs, b.is_complex);
        r = make_double2(__dadd_rn(x.x, y.x), __dadd_rn(x.y, y.y));
        break;
      }
      case EW_SUB: {
        double2 y = load_elem_rt(b.ptr, (ll)p * b.ps + (ll)f * b.fs, b.is_complex);
        r = make_double2(__dsub_rn(x.x, y.x), __dsub_rn(x.y, y.y));
        break;
      }
      case EW_NEG: r = make_double2(-x.x, -x.y); break;
      case EW_SCALE: r = cmul_rn(x, s); break;
      case EW_CONJ: r = make_double2(x.x, -x.y); break;
      case EW_RECIP: {  // ref_one() / s for scalar powers (network.rs:1526-1703)
        double d = __dadd_rn(__dmul_rn(x.x, x.x), __dmul_rn(x.y, x.y));
        r = make_double2(__ddiv_rn(x.x, d), __ddiv_rn(-x.y, d));
        break;
      }
      default: r = x; break;
    }
    store_elem(out.ptr, (ll)p * out.ps + (ll)f * out.fs, r, out.is_complex);
  }
}

// ---- layout conversion [point][flat] <-> [flat][point]: shared-memory tiled transpose, both sides coalesced ----------
template <class T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ in, T* __restrict__ out, ull rows, ull cols, ull in_ld, ull out_ld) {
  __shared__ T tile[32][33];
  const ull tiles_c = (cols + 31) / 32;
  const ull r0 = ((ull)blockIdx.x / tiles_c) * 32, c0 = ((ull)blockIdx.x % tiles_c) * 32;
  const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (unsigned j = 0; j < 32; j += 8) {
    const ull r = r0 + ty + j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + j][tx] = in[r * in_ld + c];
  }
  __syncthreads();
#pragma unroll
  for (unsigned j = 0; j < 32; j += 8) {
    const ull c = c0 + ty + j, r = r0 + tx;
    if (r < rows && c < cols) out[c * out_ld + r] = tile[tx][ty + j];
  }
}

// ---- K8: trace over one repeated pair ------------------------------------------------------------
__global__ void __launch_bounds__(256)
trace_kernel(const __grid_constant__ DevOutMap map, DevOperand a, DevOut out, ll diag_stride, unsigned dim,
             int metric, ull n_points, int point_fastest) {
  const ull total = map.size_out * n_points;
  for (ull t = (ull)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (ull)gridDim.x * blockDim.x) {
    ull p, o;
    split_index(t, map.size_out, n_points, point_fastest != 0, p, o);
    ll off = (ll)p * a.ps;
    ull r = o;
    for (int d = map.order - 1; d >= 0; --d) {
      unsigned i = (unsigned)(r % map.dim[d]);
      r /= map.dim[d];
      off += (ll)i * map.sa[d] * a.fs;
    }
    // tensor_iterators.rs:442-503: first element (sign applied to a zero), then +=/-= the rest
    double2 acc = load_elem_rt(a.ptr, off, a.is_complex);
    for (unsigned i = 1; i < dim; ++i) {
      double2 v = load_elem_rt(a.ptr, off + (ll)i * diag_stride * a.fs, a.is_complex);
      cacc_rn(acc, v, metric != 0);
    }
    store_elem(out.ptr, (ll)p * out.ps + (ll)o * out.fs, acc, out.is_complex);
  }
}

// ---- K7: sum over points ----------------------------------------------------------------------
__device__ __forceinline__ double2 warp_sum(double2 v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
  }
  return v;
}
__global__ void __launch_bounds__(256)
reduce_points_kernel(DevOperand a, ull size, ull n_points, double2* __restrict__ partial) {
  __shared__ double2 wsum[8];
  const ull c = blockIdx.y;
  double2 acc = make_double2(0.0, 0.0);
  for (ull p = (ull)blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += (ull)gridDim.x * blockDim.x) {
    double2 v = load_elem_rt(a.ptr, (ll)p * a.ps + (ll)c * a.fs, a.is_complex);
    acc.x += v.x;
    acc.y += v.y;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double2 s = wsum[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      s.x += wsum[w].x;
      s.y += wsum[w].y;
    }
    partial[(ull)blockIdx.x * size + c] = s;
  }
}
constexpr unsigned kFoldComps = 16;  // components per CTA of reduce_partials_kernel (16 row groups of 16 threads)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double2* __restrict__ partial, int n_blocks, ull size, double2* __restrict__ out) {
  __shared__ double2 s_part[256], s_out[kFoldComps];
  const unsigned c0 = blockIdx.x * kFoldComps;
  const unsigned n_c = (unsigned)min((ull)kFoldComps, size - c0);
  fold_partials_block(partial, n_blocks, size, c0, n_c, s_part, s_out);
  if (threadIdx.x < n_c) out[c0 + threadIdx.x] = s_out[threadIdx.x];
}

int g_sm_count = 0;
int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}
// grid: a multiple of the SM count (B200: 148), capped by the work available
unsigned grid_for(ull total, int block, int blocks_per_sm) {
  ull need = (total + block - 1) / block;
  ull cap = (ull)sm_count() * blocks_per_sm;
  if (need >= cap) return (unsigned)cap;
  ull g = ((need + sm_count() - 1) / sm_count()) * sm_count();
  return (unsigned)(g < need ? need : (need == 0 ? 1 : need));
}

}  // namespace

cudaError_t launch_contract_dense(const DevOutMap& map, DevOperand a, DevOperand b, DevOut out, const long long* koff_a,
                                  const long long* koff_b, const unsigned char* ksign, ull n_k, ull n_points,
                                  bool point_fastest, cudaStream_t stream) {
  const ull total = map.size_out * n_points;
  if (total == 0) return cudaSuccess;
  unsigned grid = grid_for(total, 256, 8);
#define SPN_LAUNCH_DD(AC, BC)                                                                                  \
  contract_dense_kernel<AC, BC><<<grid, 256, 0, stream>>>(map, a, b, out, koff_a, koff_b, ksign, n_k, n_points, \
                                                           point_fastest ? 1 : 0)
  if (a.is_complex && b.is_complex)
    SPN_LAUNCH_DD(true, true);
  else if (a.is_complex)
    SPN_LAUNCH_DD(true, false);
  else if (b.is_complex)
    SPN_LAUNCH_DD(false, true);
  else
    SPN_LAUNCH_DD(false, false);
#undef SPN_LAUNCH_DD
  return cudaGetLastError();
}

cudaError_t launch_contract_dense_tab(DevOperand a, DevOperand b, DevOut out, const ll* toa, const ll* tob, const ll* koff_a,
                                      const ll* koff_b, const unsigned char* ksign, ull n_k, ull size_out, ull n_points,
                                      bool point_fastest, cudaStream_t stream) {
  const ull total = size_out * n_points;
  if (total == 0) return cudaSuccess;
  if (n_k >= 0xffffffffull || size_out >= 0xffffffffull) return cudaErrorInvalidConfiguration;
  const int stage_k = n_k <= 2048;
  const size_t smem = stage_k ? (size_t)n_k * 17 + 16 : 0;
  const unsigned grid = grid_for(total, 256, 8);
#define SPN_LAUNCH_DT(AC, BC, PF)                                                                                       \
  contract_dense_tab_kernel<AC, BC, PF><<<grid, 256, smem, stream>>>(a, b, out, toa, tob, koff_a, koff_b, ksign,         \
                                                                     (unsigned)n_k, (unsigned)size_out, n_points, stage_k)
#define SPN_LAUNCH_DT2(AC, BC) \
  if (point_fastest)           \
    SPN_LAUNCH_DT(AC, BC, true); \
  else                         \
    SPN_LAUNCH_DT(AC, BC, false)
  if (a.is_complex && b.is_complex) {
    SPN_LAUNCH_DT2(true, true);
  } else if (a.is_complex) {
    SPN_LAUNCH_DT2(true, false);
  } else if (b.is_complex) {
    SPN_LAUNCH_DT2(false, true);
  } else {
    SPN_LAUNCH_DT2(false, false);
  }
#undef SPN_LAUNCH_DT2
#undef SPN_LAUNCH_DT
  return cudaGetLastError();
}

cudaError_t launch_contract_csr(DevOperand dense, DevOut out, const int* row_ptr, const long long* col_off,
                                const double2* val, const unsigned char* sign, int dense_is_first, ull size_out,
                                ull nnz, ull n_points, bool point_fastest, cudaStream_t stream) {
  const ull total = size_out * n_points;
  if (total == 0) return cudaSuccess;
  unsigned grid = grid_for(total, 256, 8);
  size_t smem = nnz * (sizeof(double2) + sizeof(ll) + 1) + (size_out + 1) * sizeof(int) + 16;
  const bool stage = smem <= 40 * 1024;
#define SPN_LAUNCH_CSR(DC, ST)                                                                              \
  contract_csr_kernel<DC, ST><<<grid, 256, (ST) ? smem : 0, stream>>>(dense, out, row_ptr, col_off, val, sign, \
                                                                      dense_is_first, size_out, nnz, n_points, \
                                                                      point_fastest ? 1 : 0)
  if (dense.is_complex) {
    if (stage)
      SPN_LAUNCH_CSR(true, true);
    else
      SPN_LAUNCH_CSR(true, false);
  } else {
    if (stage)
      SPN_LAUNCH_CSR(false, true);
    else
      SPN_LAUNCH_CSR(false, false);
  }
#undef SPN_LAUNCH_CSR
  return cudaGetLastError();
}

cudaError_t launch_elementwise(int op, DevOperand a, DevOperand b, DevOut out, double2 scalar, ull size,
                               ull n_points, bool point_fastest, cudaStream_t stream) {
  const ull total = size * n_points;
  if (total == 0) return cudaSuccess;
  unsigned grid = grid_for(total, 256, 8);
  elementwise_kernel<<<grid, 256, 0, stream>>>(op, a, b, out, scalar, size, n_points, point_fastest ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_transpose(const void* in, void* out, ull rows, ull cols, ull in_ld, ull out_ld, int elem_bytes,
                             cudaStream_t stream) {
  if (rows == 0 || cols == 0) return cudaSuccess;
  const ull tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)tiles);
  if (elem_bytes == 16)
    transpose_kernel<double2><<<grid, 256, 0, stream>>>(static_cast<const double2*>(in), static_cast<double2*>(out), rows, cols,
                                                        in_ld, out_ld);
  else
    transpose_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(in), static_cast<double*>(out), rows, cols,
                                                       in_ld, out_ld);
  return cudaGetLastError();
}

cudaError_t launch_trace(const DevOutMap& map, DevOperand a, DevOut out, ll diag_stride, unsigned dim, int metric,
                         ull n_points, bool point_fastest, cudaStream_t stream) {
  const ull total = map.size_out * n_points;
  if (total == 0) return cudaSuccess;
  unsigned grid = grid_for(total, 256, 8);
  trace_kernel<<<grid, 256, 0, stream>>>(map, a, out, diag_stride, dim, metric, n_points, point_fastest ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double2* partial, int n_blocks, ull size, double2* out,
                                   cudaStream_t stream) {
  if (size == 0) return cudaSuccess;
  unsigned grid = (unsigned)((size + kFoldComps - 1) / kFoldComps);
  reduce_partials_kernel<<<grid, 256, 0, stream>>>(partial, n_blocks, size, out);
  return cudaGetLastError();
}

cudaError_t launch_reduce_points(DevOperand a, ull size, ull n_points, double2* partial, int n_blocks, double2* out,
                                 cudaStream_t stream) {
  if (size == 0) return cudaSuccess;
  dim3 grid((unsigned)n_blocks, (unsigned)size);
  reduce_points_kernel<<<grid, 256, 0, stream>>>(a, size, n_points, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_reduce_partials(partial, n_blocks, size, out, stream);
}

}  // namespace spn
