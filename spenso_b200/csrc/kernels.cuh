// Hand-written sm_100a kernels of the pairwise ("Contract trait") path and launch wrappers.
// All kernels are HBM-bound streaming kernels over the point batch: no tensor cores (f64
// complex, tiny contracted dimensions), 128-bit loads/stores of interleaved (re,im),
// index tables of the shared operand staged in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spn {

// element (point p, flat f) of a device tensor lives at ptr[p*ps + f*fs] (in elements)
struct DevOperand {
  const void* ptr;
  long long ps, fs;
  int is_complex;
};
struct DevOut {
  void* ptr;
  long long ps, fs;
  int is_complex;
};

constexpr int kMaxOutOrder = 32;

// decomposition of the result index (row major over out_dim) into operand base offsets
struct DevOutMap {
  int order;
  unsigned dim[kMaxOutOrder];
  long long sa[kMaxOutOrder];
  long long sb[kMaxOutOrder];
  unsigned long long size_out;
};

// K1/K5/K7 (dense x dense, single / multi / exterior, any MergeInfo): contraction/singlecontract.rs:25-76,
// multicontract.rs:25-78, exteriorproduct.rs.  koff_a/koff_b/ksign: device tables of the
// contracted loop in the reference's enumeration order (n_k entries).
cudaError_t launch_contract_dense(const DevOutMap& map, DevOperand a, DevOperand b, DevOut out,
                                  const long long* koff_a, const long long* koff_b, const unsigned char* ksign,
                                  unsigned long long n_k, unsigned long long n_points, bool point_fastest,
                                  cudaStream_t stream);

// the same with host-built per-result-element base offsets toa/tob (size_out entries each): no index arithmetic on
// the device.  Needs size_out and n_k below 2^32 (cudaErrorInvalidConfiguration otherwise).
cudaError_t launch_contract_dense_tab(DevOperand a, DevOperand b, DevOut out, const long long* toa, const long long* tob,
                                      const long long* koff_a, const long long* koff_b, const unsigned char* ksign,
                                      unsigned long long n_k, unsigned long long size_out, unsigned long long n_points,
                                      bool point_fastest, cudaStream_t stream);

// K2 (dense(per point) x sparse(shared)): singlecontract.rs:78-196, multicontract.rs:80-200.
// CSR over result elements: row o holds (offset into the dense operand, value, sign) for every
// stored entry of the sparse fibre, in ascending fibre position.
cudaError_t launch_contract_csr(DevOperand dense, DevOut out, const int* row_ptr, const long long* col_off,
                                const double2* val, const unsigned char* sign, int dense_is_first,
                                unsigned long long size_out, unsigned long long nnz, unsigned long long n_points,
                                bool point_fastest, cudaStream_t stream);

// K6: dense x dense contraction that reshapes to a complex GEMM, on the FP64 tensor pipe (DMMA):
//   C[out_row[r] + out_col[c]] = sum_k sign[k] A[row_a[r] + k_a[k]] B[col_b[c] + k_b[k]]   (zgemm.cu)
// ksign: +-1.0 per k or nullptr; ps_*: point strides in elements (0 for a shared operand).
cudaError_t launch_contract_zgemm(const double2* A, const double2* B, double2* C, const long long* row_a,
                                  const long long* k_a, const long long* col_b, const long long* k_b,
                                  const long long* out_row, const long long* out_col, const double* ksign,
                                  long long M, long long N, long long K, long long ps_a, long long ps_b,
                                  long long ps_c, unsigned long long n_points, cudaStream_t stream);

// the same for f64 x f64 -> f64 operands (DGEMM).  vec: the tables are pairwise contiguous and even (see zgemm.cu),
// operands and point strides 16-byte aligned -> 16-byte gathers
cudaError_t launch_contract_dgemm(const double* A, const double* B, double* C, const long long* row_a,
                                  const long long* k_a, const long long* col_b, const long long* k_b,
                                  const long long* out_row, const long long* out_col, const double* ksign,
                                  long long M, long long N, long long K, long long ps_a, long long ps_b,
                                  long long ps_c, unsigned long long n_points, bool vec, cudaStream_t stream);

// element-wise (algebra/add_assign.rs, sub_assign.rs, neg.rs, scalar_mul.rs; FUN_LIB conj)
enum EwOp { EW_ADD = 0, EW_SUB = 1, EW_NEG = 2, EW_SCALE = 3, EW_CONJ = 4, EW_COPY = 5, EW_RECIP = 6 };
cudaError_t launch_elementwise(int op, DevOperand a, DevOperand b, DevOut out, double2 scalar,
                               unsigned long long size, unsigned long long n_points, bool point_fastest,
                               cudaStream_t stream);

// out[c * out_ld + r] = in[r * in_ld + c] for r < rows, c < cols (elements of 8 or 16 bytes): the layout conversion
// between point-major and component-major batched tensors
cudaError_t launch_transpose(const void* in, void* out, unsigned long long rows, unsigned long long cols,
                             unsigned long long in_ld, unsigned long long out_ld, int elem_bytes, cudaStream_t stream);

// K8 trace: out[o] = sum_i sign(i) a[base(o) + i*diag_stride]
cudaError_t launch_trace(const DevOutMap& map, DevOperand a, DevOut out, long long diag_stride, unsigned dim,
                         int metric, unsigned long long n_points, bool point_fastest, cudaStream_t stream);

// K7 reduction over the point axis: out[c] = sum_p a[p][c] (deterministic two-stage tree)
cudaError_t launch_reduce_points(DevOperand a, unsigned long long size, unsigned long long n_points,
                                 double2* partial /* [n_blocks][size] */, int n_blocks, double2* out,
                                 cudaStream_t stream);
// second stage alone (used by the fused network kernel's per-block partial sums)
cudaError_t launch_reduce_partials(const double2* partial, int n_blocks, unsigned long long size, double2* out,
                                   cudaStream_t stream);

#ifdef __CUDACC__
// Fold of per-CTA partial sums partial[b * size + c] (b < n_blocks) for the n_c <= 256 components starting at c0, by
// ONE CTA of 256 threads: thread = (row group g, component c); rows are read coalesced with four independent loads
// in flight per thread (the fold is pure load latency), then the groups are added in a fixed order => run-to-run
// deterministic Monte-Carlo sums.  s_part: 256 double2 of shared memory; the result for component c0 + c is in
// s_out[c] (shared) after the call.
__device__ __forceinline__ void fold_partials_block(const double2* __restrict__ partial, int n_blocks,
                                                    unsigned long long size, unsigned c0, unsigned n_c,
                                                    double2* s_part, double2* s_out) {
  const unsigned tid = threadIdx.x;
  const unsigned G = 256u / n_c;
  const unsigned g = tid / n_c, c = tid - g * n_c;
  if (g < G) {
    const double2* col = partial + c0 + c;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
    int b = (int)g;
    for (; b + 3 * (int)G < n_blocks; b += 4 * (int)G) {
      const double2 v0 = col[(unsigned long long)b * size], v1 = col[(unsigned long long)(b + (int)G) * size];
      const double2 v2 = col[(unsigned long long)(b + 2 * (int)G) * size], v3 = col[(unsigned long long)(b + 3 * (int)G) * size];
      a0.x += v0.x; a0.y += v0.y;
      a1.x += v1.x; a1.y += v1.y;
      a2.x += v2.x; a2.y += v2.y;
      a3.x += v3.x; a3.y += v3.y;
    }
    for (; b < n_blocks; b += (int)G) {
      const double2 v = col[(unsigned long long)b * size];
      a0.x += v.x; a0.y += v.y;
    }
    s_part[g * n_c + c] = make_double2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
  }
  __syncthreads();
  if (tid < n_c) {
    double2 s = s_part[tid];
    for (unsigned k = 1; k < G; ++k) {
      s.x += s_part[k * n_c + tid].x;
      s.y += s_part[k * n_c + tid].y;
    }
    s_out[tid] = s;
  }
  __syncthreads();
}
#endif

}  // namespace spn
