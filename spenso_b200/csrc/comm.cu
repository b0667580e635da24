// Monte-Carlo sum exchange over NVLink peer memory (include/spenso_b200.h: spn_comm_*).
//
// The only collective of the path is the final all-reduce of the per-rank Monte-Carlo partial sums:
// 2*|out| doubles (the reference has no counterpart, it is single-process).  A library
// all-reduce for 16..512 bytes is pure latency (launch + protocol), and it would follow a separate
// kernel that folds the per-CTA partials of the network kernel.  Here both are ONE single-CTA kernel:
//
//   1. fold this rank's per-CTA partials (same fixed tree as reduce_partials_kernel => same bits), or take the
//      local sum the caller supplies;
//   2. PUSH it into a mailbox slot in every peer's HBM with plain stores through NVLink/NVSwitch peer mappings
//      (CUDA IPC); every 8-byte word carries 32 data bits and the epoch tag, so no fence and no flag follow;
//   3. poll the LOCAL slots until the tag of this epoch shows up (polling never crosses NVLink), bounded by a
//      %globaltimer timeout so a missing peer cannot hang the GPU;
//   4. add the world's contributions in RANK ORDER: every rank computes the identical bits.
//
// Mailbox slots are double-buffered by epoch parity.  A rank can only finish epoch e after every peer has
// published epoch e, which a peer does after it finished reading epoch e-1; so when slot parity e+1 is
// overwritten its previous content (epoch e-1) is no longer needed anywhere.  The epoch counter lives in
// device memory and is advanced by the kernel itself, so launches are CUDA-graph capturable.
#include <cstring>
#include <string>
#include <vector>

#include "tensor.hpp"

using namespace spn;

namespace {

typedef unsigned long long ull;

// Mailbox of one rank (device memory, zero-initialised):
//   uint2 slot[2][world][2 * cap];   ull epoch;
// A double travels as two 8-byte words {32 data bits, 32-bit epoch tag}: an aligned 8-byte store is single-copy
// atomic on every path (local HBM and NVLink), so a reader that sees the tag of the current epoch in BOTH words
// has the value — no fence, no separate flag, one link crossing per exchange (the "low-latency" protocol of
// collective libraries).  Tag 0 is never used: zero-initialised slots are "empty".
struct MailboxView {
  char* base;
  unsigned world, cap;
  __host__ __device__ static size_t bytes(unsigned world, unsigned cap) {
    return (size_t)2 * world * 2 * cap * sizeof(uint2) + sizeof(ull);
  }
  __device__ uint2* slot(unsigned par, unsigned src) const {
    return reinterpret_cast<uint2*>(base) + ((size_t)par * world + src) * 2 * cap;
  }
  __device__ ull* epoch() const { return reinterpret_cast<ull*>(base + (size_t)2 * world * 2 * cap * sizeof(uint2)); }
};

__device__ __forceinline__ void st_sys_v2(uint2* p, unsigned a, unsigned b) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_sys_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ull globaltimer_ns() {
  ull t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kExchangeThreads = 256;
constexpr unsigned kMaxFoldComps = 16;  // larger results are folded by reduce_partials_kernel (many CTAs) first
constexpr unsigned kRanksPerPass = 8;  // contributions polled together (independent loads), then added in rank order

// partial != nullptr: the local contribution is the fold of n_blocks per-CTA partials of n/2 complex components
//                     (fold_partials_block: the order of reduce_partials_kernel => the same bits)
// partial == nullptr: the local contribution is inout[0..n)
__global__ void __launch_bounds__(kExchangeThreads)
mc_sum_exchange_kernel(const double2* __restrict__ partial, int n_blocks, double* __restrict__ inout, unsigned n,
                       char* const* __restrict__ peers, unsigned rank, unsigned world, unsigned cap,
                       int* __restrict__ status, ull timeout_ns) {
  extern __shared__ double s_local[];
  const unsigned tid = threadIdx.x;
  const MailboxView mine{peers[rank], world, cap};
  const ull e = *mine.epoch() + 1;  // written by this kernel's previous launch (stream order)
  if (partial) {
    __shared__ double2 s_part[256];
    const unsigned n_comp = n >> 1;   // host guarantees n_comp <= kMaxFoldComps
    fold_partials_block(partial, n_blocks, n_comp, 0, n_comp, s_part, reinterpret_cast<double2*>(s_local));
  } else {
    for (unsigned i = tid; i < n; i += blockDim.x) s_local[i] = inout[i];
  }
  __syncthreads();
  const unsigned par = (unsigned)(e & 1);
  const unsigned tag = (unsigned)(e % 0xFFFFFFFFull) + 1u;  // never 0
  // push: slot (parity, source = me) of every rank, my own included
  for (unsigned idx = tid; idx < world * n; idx += blockDim.x) {
    const unsigned q = idx / n, i = idx - q * n;
    const MailboxView dst{peers[q], world, cap};
    const ull bits = (ull)__double_as_longlong(s_local[i]);
    uint2* w = dst.slot(par, rank) + 2 * i;
    st_sys_v2(w, (unsigned)bits, tag);
    st_sys_v2(w + 1, (unsigned)(bits >> 32), tag);
  }
  // gather: poll my own slots (local memory) and add the world's contributions in rank order
  const ull t0 = globaltimer_ns();
  for (unsigned i = tid; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (unsigned r0 = 0; r0 < world; r0 += kRanksPerPass) {
      const unsigned nr = min(kRanksPerPass, world - r0);
      double v[kRanksPerPass];
      unsigned pending = (1u << nr) - 1u, spins = 0;
      while (pending) {
#pragma unroll
        for (unsigned k = 0; k < kRanksPerPass; ++k) {
          if (k < nr && (pending >> k & 1u)) {
            const uint2* w = mine.slot(par, r0 + k) + 2 * i;
            const uint2 lo = ld_sys_v2(w), hi = ld_sys_v2(w + 1);
            if (lo.y == tag && hi.y == tag) {
              v[k] = __longlong_as_double((long long)(((ull)hi.x << 32) | lo.x));
              pending &= ~(1u << k);
            }
          }
        }
        if (pending && (++spins & 255u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
          atomicExch_system(status, 1 + (int)(r0 + (unsigned)__ffs((int)pending) - 1));  // a peer that never arrived
          // poison the sum: a missing contribution must not look like a plausible Monte-Carlo estimate
#pragma unroll
          for (unsigned k = 0; k < kRanksPerPass; ++k)
            if (pending >> k & 1u) v[k] = __longlong_as_double(0x7ff8000000000000LL);
          pending = 0;
        }
      }
#pragma unroll
      for (unsigned k = 0; k < kRanksPerPass; ++k)
        if (k < nr) acc += v[k];
    }
    inout[i] = acc;
  }
  __syncthreads();
  if (tid == 0) *mine.epoch() = e;
}

}  // namespace

struct spn_comm {
  spn_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  uint32_t cap = 0;
  char* local = nullptr;
  std::vector<char*> peer;   // mailbox base of every rank in THIS process' address space (peer[rank] == local)
  char** peer_dev = nullptr;
  int* status_host = nullptr;  // pinned + mapped: 0 ok, 1 + r = timed out waiting for rank r
  int* status_dev = nullptr;
  bool connected = false;
  double timeout_s = 10.0;
  // frees what this rank owns (the caller has unmapped the peers first); also the cleanup of a failed create
  ~spn_comm() {
    if (ctx && status_host)
      for (auto it = ctx->comm_status.begin(); it != ctx->comm_status.end(); ++it)
        if (*it == status_host) {
          ctx->comm_status.erase(it);
          break;
        }
    if (peer_dev) cudaFree(peer_dev);
    if (local) cudaFree(local);
    if (status_host) cudaFreeHost(status_host);
    cudaGetLastError();
  }
};

namespace spn {
// used by net.cu: fold the per-CTA partials and exchange in one launch
void comm_exchange(spn_comm* c, const double2* partial, int n_blocks, double* inout_device, uint32_t n_f64) {
  if (!c->connected) throw Error(SPN_ERR_INVALID, "spn_comm: not connected (call spn_comm_connect on every rank)");
  if (const int st = *reinterpret_cast<volatile int*>(c->status_host))
    throw Error(SPN_ERR_OTHER, "spn_comm: an earlier exchange timed out waiting for rank " + std::to_string(st - 1) +
                                   "; its result was poisoned with NaN and the epochs are out of step — destroy and re-create "
                                   "the communicator on every rank");
  if (n_f64 == 0) return;
  if (n_f64 > c->cap) throw Error(SPN_ERR_INVALID, "spn_comm: message larger than the mailbox capacity");
  if (partial && (n_f64 & 1)) throw Error(SPN_ERR_INVALID, "spn_comm: partials are complex (even number of doubles)");
  if (partial && n_f64 / 2 > kMaxFoldComps) {
    // a wide result: fold on many CTAs first, then exchange the local sums (two launches)
    cuda_check(launch_reduce_partials(partial, n_blocks, n_f64 / 2, reinterpret_cast<double2*>(inout_device), c->ctx->stream),
               "reduce_partials_kernel");
    c->ctx->launches++;
    partial = nullptr;
  }
  const ull timeout_ns = (ull)(c->timeout_s * 1e9);
  mc_sum_exchange_kernel<<<1, kExchangeThreads, n_f64 * sizeof(double), c->ctx->stream>>>(
      partial, n_blocks, inout_device, n_f64, c->peer_dev, (unsigned)c->rank, (unsigned)c->world, c->cap, c->status_dev,
      timeout_ns);
  cuda_check(cudaGetLastError(), "mc_sum_exchange_kernel");
  c->ctx->launches++;
}
}  // namespace spn

static thread_local std::string g_comm_err;
#define COMM_TRY try {
#define COMM_CATCH                   \
  }                                  \
  catch (const Error& e) {           \
    spn::set_error(e.what());        \
    return e.code;                   \
  }                                  \
  catch (const std::exception& e) {  \
    spn::set_error(e.what());        \
    return SPN_ERR_OTHER;            \
  }

extern "C" {

int spn_comm_create(spn_ctx* ctx, int rank, int world, uint32_t max_f64, spn_comm** out) {
  COMM_TRY
  if (!ctx || !out || world < 1 || rank < 0 || rank >= world || max_f64 == 0 || max_f64 > 4096 || world > 64)
    throw Error(SPN_ERR_INVALID, "spn_comm_create: bad arguments (1 <= world <= 64, 0 < max_f64 <= 4096)");
  static_assert(sizeof(cudaIpcMemHandle_t) == SPN_COMM_HANDLE_BYTES, "handle size");
  cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice");
  auto c = std::make_unique<spn_comm>();
  c->ctx = ctx;
  c->rank = rank;
  c->world = world;
  c->cap = max_f64;
  if (const char* e = getenv("SPN_COMM_TIMEOUT_S")) c->timeout_s = atof(e);
  const size_t bytes = MailboxView::bytes((unsigned)world, max_f64);
  cuda_check(cudaMalloc((void**)&c->local, bytes), "cudaMalloc (mailbox)");
  cuda_check(cudaMemset(c->local, 0, bytes), "cudaMemset (mailbox)");
  cuda_check(cudaHostAlloc((void**)&c->status_host, sizeof(int), cudaHostAllocMapped), "cudaHostAlloc (comm status)");
  *c->status_host = 0;
  ctx->comm_status.push_back(c->status_host);   // spn_ctx_synchronize reports a timed-out exchange
  cuda_check(cudaHostGetDevicePointer((void**)&c->status_dev, c->status_host, 0), "cudaHostGetDevicePointer");
  cuda_check(cudaMalloc((void**)&c->peer_dev, sizeof(char*) * (size_t)world), "cudaMalloc (peer table)");
  cuda_check(cudaDeviceSynchronize(), "mailbox initialisation");
  c->peer.assign((size_t)world, nullptr);
  c->peer[(size_t)rank] = c->local;
  if (world == 1) {
    cuda_check(cudaMemcpy(c->peer_dev, c->peer.data(), sizeof(char*), cudaMemcpyHostToDevice), "peer table");
    c->connected = true;
  }
  *out = c.release();
  return SPN_OK;
  COMM_CATCH
}

int spn_comm_handle(const spn_comm* c, void* handle_out) {
  COMM_TRY
  if (!c || !handle_out) throw Error(SPN_ERR_INVALID, "spn_comm_handle: null argument");
  cudaIpcMemHandle_t h;
  cuda_check(cudaIpcGetMemHandle(&h, c->local), "cudaIpcGetMemHandle");
  std::memcpy(handle_out, &h, sizeof h);
  return SPN_OK;
  COMM_CATCH
}

int spn_comm_connect(spn_comm* c, const void* all_handles) {
  COMM_TRY
  if (!c || !all_handles) throw Error(SPN_ERR_INVALID, "spn_comm_connect: null argument");
  if (c->connected) return SPN_OK;
  cuda_check(cudaSetDevice(c->ctx->device), "cudaSetDevice");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(all_handles) + (size_t)r * sizeof h, sizeof h);
    void* p = nullptr;
    cuda_check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess),
               "cudaIpcOpenMemHandle (peer mailbox; needs P2P between the GPUs of one node)");
    c->peer[(size_t)r] = static_cast<char*>(p);
  }
  cuda_check(cudaMemcpy(c->peer_dev, c->peer.data(), sizeof(char*) * (size_t)c->world, cudaMemcpyHostToDevice), "peer table");
  c->connected = true;
  return SPN_OK;
  COMM_CATCH
}

int spn_comm_allreduce_sum(spn_comm* c, double* inout_device, uint32_t n_f64) {
  COMM_TRY
  if (!c || !inout_device) throw Error(SPN_ERR_INVALID, "spn_comm_allreduce_sum: null argument");
  comm_exchange(c, nullptr, 0, inout_device, n_f64);
  return SPN_OK;
  COMM_CATCH
}

int spn_comm_status(const spn_comm* c) {
  if (!c) return -1;
  return *reinterpret_cast<volatile int*>(c->status_host);
}

int spn_comm_disconnect(spn_comm* c) {
  COMM_TRY
  if (!c) throw Error(SPN_ERR_INVALID, "spn_comm_disconnect: null argument");
  cuda_check(cudaSetDevice(c->ctx->device), "cudaSetDevice");
  cuda_check(cudaDeviceSynchronize(), "spn_comm_disconnect");
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer[(size_t)r]) {
      cudaIpcCloseMemHandle(c->peer[(size_t)r]);
      c->peer[(size_t)r] = nullptr;
    }
  cudaGetLastError();
  c->connected = false;
  return SPN_OK;
  COMM_CATCH
}

void spn_comm_destroy(spn_comm* c) {
  if (!c) return;
  spn_comm_disconnect(c);
  delete c;
}

}  // extern "C"
