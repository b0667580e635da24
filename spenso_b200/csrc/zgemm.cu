// Dense x dense pairwise contraction that genuinely reshapes to a complex GEMM (BASELINE config 4:
// rank-6, dim-32 tensors over three shared indices = ZGEMM 32768^3) on the FP64 tensor pipe.
//
//   C[out_row[r] + out_col[c]] = sum_k sign[k] * A[row_a[r] + k_a[k]] * B[col_b[c] + k_b[k]]
//
// is the same sum the reference's MultiContract / SingleContract kernels compute
// (/root/reference/spenso/src/contraction/multicontract.rs:25-78, singlecontract.rs:25-76); the offset
// tables come from the stride plan of the merge (spn_plan_contract), so every MergeInfo — including
// interleaved results and transposed operands — goes through this one kernel: tiles are GATHERED through
// the tables with 16-byte cp.async (contiguous runs of the tables give coalesced 128-byte segments).
//
// sm_100a notes: tcgen05.mma has no f64 kind, so the FP64 tensor path is the warp-level DMMA
// (mma.sync.aligned.m8n8k4.f64).  CTA tile 64x64 complex, 4 warps of 32x32, BK = 8, 4-stage cp.async
// pipeline, operands kept interleaved (re,im) in shared memory and read as 128-bit fragments with
// conflict-free strides (A: 12, B: 66 sixteen-byte units); a complex tile product is 4 real DMMAs
// (rr, -ii, ri, ir).  Grouped rasterisation keeps an 8-tile-tall band of A resident in L2.
#include "kernels.cuh"

namespace spn {

namespace {

typedef long long ll;
typedef unsigned long long ull;

constexpr int BM = 64, BN = 64, BK = 8, STAGES = 4;
constexpr int SA = BK + 4;   // row stride of the A tile in double2 units (== 4 mod 8)
constexpr int SB = BN + 2;   // row stride of the B tile in double2 units (== 2 mod 8)
constexpr int GROUP_M = 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = valid ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool SIGN>
__global__ void __launch_bounds__(128, 2)
contract_zgemm_kernel(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ C,
                      const ll* __restrict__ row_a, const ll* __restrict__ k_a, const ll* __restrict__ col_b,
                      const ll* __restrict__ k_b, const ll* __restrict__ out_row, const ll* __restrict__ out_col,
                      const double* __restrict__ ksign, ll M, ll N, ll K, ll ps_a, ll ps_b, ll ps_c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* As = reinterpret_cast<double2*>(smem_raw);          // [STAGES][BM][SA]
  double2* Bs = As + STAGES * BM * SA;                          // [STAGES][BK][SB]
  ll* s_row = reinterpret_cast<ll*>(Bs + STAGES * BK * SB);     // [BM] offsets of this tile's rows in A
  ll* s_col = s_row + BM;                                       // [BN] offsets of this tile's columns in B

  // grouped rasterisation: consecutive CTAs walk GROUP_M row tiles before moving to the next column tile
  const ll tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const ll pid = blockIdx.x;
  const ll group = pid / (GROUP_M * tiles_n);
  const ll first_m = group * GROUP_M;
  const ll gsz = min((ll)GROUP_M, tiles_m - first_m);
  const ll tile_m = first_m + (pid % (GROUP_M * tiles_n)) % gsz;
  const ll tile_n = (pid % (GROUP_M * tiles_n)) / gsz;
  const ll m0 = tile_m * BM, n0 = tile_n * BN;
  const ll p = blockIdx.y;
  A += p * ps_a;
  B += p * ps_b;
  C += p * ps_c;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  if (tid < BM) s_row[tid] = (m0 + tid < M) ? __ldg(row_a + m0 + tid) : -1;
  if (tid >= 64 && tid < 64 + BN) s_col[tid - 64] = (n0 + tid - 64 < N) ? __ldg(col_b + n0 + tid - 64) : -1;
  __syncthreads();

  // loader assignment: A tile 64 rows x 8 k -> 4 elements per thread (k fastest across threads);
  //                    B tile 8 k x 64 n   -> 4 elements per thread (n fastest across threads)
  const int a_kk = tid & 7, a_r0 = tid >> 3;        // rows a_r0 + 16 i
  const int b_n = tid & 63, b_k0 = tid >> 6;        // k rows b_k0 + 2 i
  const ll KT = (K + BK - 1) / BK;

  // k-offsets of a tile are fetched one pipeline step ahead of the cp.async that needs them, so the
  // gather addresses never wait on a global load
  ll ka_next = 0, kb_next[4] = {0, 0, 0, 0};
  auto fetch_koffs = [&](ll kt) {
    const ll kbase = kt * BK;
    const ll k = kbase + a_kk;
    ka_next = (kt < KT && k < K) ? __ldg(k_a + k) : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const ll kk = kbase + b_k0 + 2 * i;
      kb_next[i] = (kt < KT && kk < K) ? __ldg(k_b + kk) : 0;
    }
  };
  auto load_tile = [&](int stage, ll kt) {  // uses the offsets fetched by fetch_koffs(kt)
    const ll kbase = kt * BK;
    double2* as = As + stage * BM * SA;
    double2* bs = Bs + stage * BK * SB;
    {
      const bool kv = kbase + a_kk < K;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = a_r0 + 16 * i;
        const ll ro = s_row[r];
        const bool v = kv && ro >= 0;
        cp_async16(as + r * SA + a_kk, A + (v ? ro + ka_next : 0), v);
      }
    }
    {
      const ll co = s_col[b_n];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = b_k0 + 2 * i;
        const bool v = kbase + kk < K && co >= 0;
        cp_async16(bs + kk * SB + b_n, B + (v ? co + kb_next[i] : 0), v);
      }
    }
  };

  double cr[4][4][2], ci[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    fetch_koffs(s);
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }
  fetch_koffs(STAGES - 1);

  const int fr = lane >> 2, fk = lane & 3;  // fragment row (or column) and k position of this lane
  for (ll kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const ll nxt = kt + STAGES - 1;
      if (nxt < KT) load_tile((int)(nxt % STAGES), nxt);
      cp_async_commit();
      fetch_koffs(nxt + 1);
    }
    const int stage = (int)(kt % STAGES);
    const double2* as = As + stage * BM * SA;
    const double2* bs = Bs + stage * BK * SB;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double2 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = as[(wm + 8 * i + fr) * SA + k4 * 4 + fk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[(k4 * 4 + fk) * SB + wn + 8 * j + fr];
      if (SIGN) {
        const ll k = kt * BK + k4 * 4 + fk;
        const double sg = (k < K) ? __ldg(ksign + k) : 1.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i].x *= sg;
          a[i].y *= sg;
        }
      }
      // four passes over the 16 accumulator tiles: consecutive DMMAs never touch the same accumulator
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double nai = -a[i].y;
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(cr[i][j][0], cr[i][j][1], nai, b[j].y);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[row = lane/4][cols 2*(lane%4), +1] of every 8x8 tile
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const ll r = m0 + wm + 8 * i + fr;
    if (r >= M) continue;
    const ll orow = __ldg(out_row + r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const ll c = n0 + wn + 8 * j + 2 * fk + e;
        if (c < N) C[orow + __ldg(out_col + c)] = make_double2(cr[i][j][e], ci[i][j][e]);
      }
    }
  }
}

// ---- real (f64 x f64) variant: DGEMM on the same pipe -------------------------------------------------------------
// CTA tile 128(m) x 64(n), 4 warps of 32 x 64, two CTAs per SM (like the complex kernel: the barrier of one CTA is
// covered by the other), BK = 16 (one barrier per 128 DMMAs of a warp), 3-stage cp.async pipeline of 8-byte (or, VEC,
// 16-byte) gathers.  Round-2 history, fraction of cuBLAS DGEMM on 32768^3: 128x128 / 8 warps / BK 8 / 64-bit fragment
// loads 0.83 -> 128-bit fragment loads 0.85 -> BK 16 0.87 -> this shape 0.95 (profiles/README.md).
// Fragments are read with 128-bit shared loads — half the load instructions per DMMA of 64-bit fragment loads:
//   A: the two DMMA k-steps of each group of 8 k take its even and its odd k, so a lane's two A operands (k = 2 fk,
//      2 fk + 1) are adjacent in the natural [row][k] tile;
//   B: the 8x8 tiles of a warp are paired and tile (j', e) owns the columns 16 j' + 2 c + e, so a lane's operands for a
//      pair of tiles (columns 2 fr, 2 fr + 1) are adjacent in the natural [k][n] tile; the epilogue un-permutes.
// Conflict-free strides in 16-byte units: A rows 12 (DSA = 24 doubles), B rows 33 (DSB = 66 doubles).
constexpr int DBM = 128, DBN = 64, DBK = 16, DSA = DBK + 8, DSB = DBN + 2, DSTAGES = 3, DTHREADS = 128;

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}

// VEC: the gather tables are pairwise contiguous and even (k_a[2j+1] = k_a[2j] + 1, col_b[2j+1] = col_b[2j] + 1, all
// base offsets even), so operands move with 16-byte cp.async — half the copy instructions of the 8-byte gather.
template <bool SIGN, bool VEC>
__global__ void __launch_bounds__(DTHREADS, 2)
contract_dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                      const ll* __restrict__ row_a, const ll* __restrict__ k_a, const ll* __restrict__ col_b,
                      const ll* __restrict__ k_b, const ll* __restrict__ out_row, const ll* __restrict__ out_col,
                      const double* __restrict__ ksign, ll M, ll N, ll K, ll ps_a, ll ps_b, ll ps_c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);             // [DSTAGES][DBM][DSA]
  double* Bs = As + DSTAGES * DBM * DSA;                         // [DSTAGES][DBK][DSB]
  ll* s_row = reinterpret_cast<ll*>(Bs + DSTAGES * DBK * DSB);   // [DBM]
  ll* s_col = s_row + DBM;                                      // [DBN]

  const ll tiles_m = (M + DBM - 1) / DBM, tiles_n = (N + DBN - 1) / DBN;
  const ll pid = blockIdx.x;
  const ll group = pid / (GROUP_M * tiles_n);
  const ll first_m = group * GROUP_M;
  const ll gsz = min((ll)GROUP_M, tiles_m - first_m);
  const ll tile_m = first_m + (pid % (GROUP_M * tiles_n)) % gsz;
  const ll tile_n = (pid % (GROUP_M * tiles_n)) / gsz;
  const ll m0 = tile_m * DBM, n0 = tile_n * DBN;
  const ll p = blockIdx.y;
  A += p * ps_a;
  B += p * ps_b;
  C += p * ps_c;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp * 32, wn = 0;
  s_row[tid] = (m0 + tid < M) ? __ldg(row_a + m0 + tid) : -1;
  if (tid < DBN) s_col[tid] = (n0 + tid < N) ? __ldg(col_b + n0 + tid) : -1;
  __syncthreads();

  // loaders (8-byte gathers): A tile 128 rows x 16 k -> 16 elements per thread (k fastest across threads: k = tid & 15,
  //                           rows tid / 16 + 8 i); B tile 16 k x 64 n -> 8 per thread (n = tid & 63, k = tid / 64 + 2 i)
  const int a_kk = tid & 15, a_r0 = tid >> 4;
  const int b_n = tid & 63, b_k0 = tid >> 6;
  const ll KT = (K + DBK - 1) / DBK;
  ll ka_next = 0, kb_next[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto fetch_koffs = [&](ll kt) {
    const ll kbase = kt * DBK;
    const ll k = kbase + a_kk;
    ka_next = (kt < KT && k < K) ? __ldg(k_a + k) : 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const ll kk = kbase + b_k0 + 2 * i;
      kb_next[i] = (kt < KT && kk < K) ? __ldg(k_b + kk) : 0;
    }
  };
  // VEC loaders: A 128 rows x 8 k-pairs -> 8 copies per thread (k pair = tid & 7, rows tid / 8 + 16 i);
  //              B 16 k x 32 n-pairs -> 4 copies per thread (n pair = tid & 31, k = tid / 32 + 4 i)
  const int va_kp = tid & 7, va_r0 = tid >> 3;
  const int vb_np = tid & 31, vb_k0 = tid >> 5;
  ll vka_next = 0, vkb_next[4] = {0, 0, 0, 0};
  auto fetch_koffs_vec = [&](ll kt) {
    const ll kbase = kt * DBK;
    const ll k = kbase + 2 * va_kp;
    vka_next = (kt < KT && k < K) ? __ldg(k_a + k) : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const ll kk = kbase + vb_k0 + 4 * i;
      vkb_next[i] = (kt < KT && kk < K) ? __ldg(k_b + kk) : 0;
    }
  };
  auto load_tile = [&](int stage, ll kt) {
    const ll kbase = kt * DBK;
    double* as = As + stage * DBM * DSA;
    double* bs = Bs + stage * DBK * DSB;
    if (VEC) {
      const bool kv = kbase + 2 * va_kp < K;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = va_r0 + 16 * i;
        const ll ro = s_row[r];
        const bool v = kv && ro >= 0;
        cp_async16(as + r * DSA + 2 * va_kp, A + (v ? ro + vka_next : 0), v);
      }
      const ll co = s_col[2 * vb_np];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = vb_k0 + 4 * i;
        const bool v = kbase + kk < K && co >= 0;
        cp_async16(bs + kk * DSB + 2 * vb_np, B + (v ? co + vkb_next[i] : 0), v);
      }
      return;
    }
    const bool kv = kbase + a_kk < K;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = a_r0 + 8 * i;
      const ll ro = s_row[r];
      const bool v = kv && ro >= 0;
      cp_async8(as + r * DSA + a_kk, A + (v ? ro + ka_next : 0), v);
    }
    const ll co = s_col[b_n];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kk = b_k0 + 2 * i;
      const bool v = kbase + kk < K && co >= 0;
      cp_async8(bs + kk * DSB + b_n, B + (v ? co + kb_next[i] : 0), v);
    }
  };
  auto fetch = [&](ll kt) {
    if (VEC)
      fetch_koffs_vec(kt);
    else
      fetch_koffs(kt);
  };

  double c[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[i][j][0] = c[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < DSTAGES - 1; ++s) {
    fetch(s);
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }
  fetch(DSTAGES - 1);

  const int fr = lane >> 2, fk = lane & 3;
  for (ll kt = 0; kt < KT; ++kt) {
    cp_async_wait<DSTAGES - 2>();
    __syncthreads();
    {
      const ll nxt = kt + DSTAGES - 1;
      if (nxt < KT) load_tile((int)(nxt % DSTAGES), nxt);
      cp_async_commit();
      fetch(nxt + 1);
    }
    const int stage = (int)(kt % DSTAGES);
    const double* as = As + stage * DBM * DSA;
    const double* bs = Bs + stage * DBK * DSB;
#pragma unroll
    for (int h = 0; h < DBK / 8; ++h) {   // groups of 8 k: two DMMA k-steps each (even k, odd k)
      // a2[i] = A[row i][k = 8 h + 2 fk, + 1]; b2[s][j'] = B[k = 8 h + 2 fk + s][columns 16 j' + 2 fr, + 1]
      double2 a2[4], b2[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a2[i] = *reinterpret_cast<const double2*>(as + (wm + 8 * i + fr) * DSA + 8 * h + 2 * fk);
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp)
          b2[s2][jp] = *reinterpret_cast<const double2*>(bs + (8 * h + 2 * fk + s2) * DSB + wn + 16 * jp + 2 * fr);
      if (SIGN) {
        const ll k = kt * DBK + 8 * h + 2 * fk;
        const double sg0 = (k < K) ? __ldg(ksign + k) : 1.0, sg1 = (k + 1 < K) ? __ldg(ksign + k + 1) : 1.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a2[i].x *= sg0;
          a2[i].y *= sg1;
        }
      }
      // even k, then odd k; tile j = 2 j' + e
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          dmma884(c[i][2 * jp][0], c[i][2 * jp][1], a2[i].x, b2[0][jp].x);
          dmma884(c[i][2 * jp + 1][0], c[i][2 * jp + 1][1], a2[i].x, b2[0][jp].y);
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          dmma884(c[i][2 * jp][0], c[i][2 * jp][1], a2[i].y, b2[1][jp].x);
          dmma884(c[i][2 * jp + 1][0], c[i][2 * jp + 1][1], a2[i].y, b2[1][jp].y);
        }
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const ll r = m0 + wm + 8 * i + fr;
    if (r >= M) continue;
    const ll orow = __ldg(out_row + r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        // tile j = 2 j' + t holds, in its fragment column c8 = 2 fk + e, the tile column 16 j' + 2 c8 + t
        const ll cc = n0 + wn + 16 * (j >> 1) + 2 * (2 * fk + e) + (j & 1);
        if (cc < N) C[orow + __ldg(out_col + cc)] = c[i][j][e];
      }
    }
  }
}

}  // namespace

cudaError_t launch_contract_dgemm(const double* A, const double* B, double* C, const long long* row_a,
                                  const long long* k_a, const long long* col_b, const long long* k_b,
                                  const long long* out_row, const long long* out_col, const double* ksign,
                                  long long M, long long N, long long K, long long ps_a, long long ps_b,
                                  long long ps_c, unsigned long long n_points, bool vec, cudaStream_t stream) {
  if (M == 0 || N == 0 || n_points == 0) return cudaSuccess;
  const ll tiles = ((M + DBM - 1) / DBM) * ((N + DBN - 1) / DBN);
  if (tiles > 0x7fffffffLL || n_points > 65535) return cudaErrorInvalidConfiguration;
  const size_t smem = (size_t)DSTAGES * (DBM * DSA + DBK * DSB) * sizeof(double) + (DBM + DBN) * sizeof(long long);
  dim3 grid((unsigned)tiles, (unsigned)n_points);
  cudaError_t e = cudaSuccess;
#define SPN_LAUNCH_DGEMM(SG, VC)                                                                                          \
  e = cudaFuncSetAttribute(contract_dgemm_kernel<SG, VC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
  if (e != cudaSuccess) return e;                                                                                         \
  contract_dgemm_kernel<SG, VC><<<grid, DTHREADS, smem, stream>>>(A, B, C, row_a, k_a, col_b, k_b, out_row, out_col, ksign, M, N, \
                                                             K, ps_a, ps_b, ps_c)
  if (ksign) {
    if (vec) { SPN_LAUNCH_DGEMM(true, true); } else { SPN_LAUNCH_DGEMM(true, false); }
  } else {
    if (vec) { SPN_LAUNCH_DGEMM(false, true); } else { SPN_LAUNCH_DGEMM(false, false); }
  }
#undef SPN_LAUNCH_DGEMM
  return cudaGetLastError();
}

size_t zgemm_smem_bytes() {
  return (size_t)STAGES * (BM * SA + BK * SB) * sizeof(double2) + (BM + BN) * sizeof(long long);
}

cudaError_t launch_contract_zgemm(const double2* A, const double2* B, double2* C, const long long* row_a,
                                  const long long* k_a, const long long* col_b, const long long* k_b,
                                  const long long* out_row, const long long* out_col, const double* ksign,
                                  long long M, long long N, long long K, long long ps_a, long long ps_b,
                                  long long ps_c, unsigned long long n_points, cudaStream_t stream) {
  if (M == 0 || N == 0 || n_points == 0) return cudaSuccess;
  const ll tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (tiles > 0x7fffffffLL || n_points > 65535) return cudaErrorInvalidConfiguration;
  const size_t smem = zgemm_smem_bytes();
  dim3 grid((unsigned)tiles, (unsigned)n_points);
  cudaError_t e;
  if (ksign) {
    e = cudaFuncSetAttribute(contract_zgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    contract_zgemm_kernel<true><<<grid, 128, smem, stream>>>(A, B, C, row_a, k_a, col_b, k_b, out_row, out_col, ksign, M,
                                                            N, K, ps_a, ps_b, ps_c);
  } else {
    e = cudaFuncSetAttribute(contract_zgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    contract_zgemm_kernel<false><<<grid, 128, smem, stream>>>(A, B, C, row_a, k_a, col_b, k_b, out_row, out_col, nullptr,
                                                             M, N, K, ps_a, ps_b, ps_c);
  }
  return cudaGetLastError();
}

}  // namespace spn
