// Probe: what limits a 64 B/point read + 256 B/point write stream (gamma^mu p_mu) on B200?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_stream_probe write_stream_probe.cu
// Each variant computes unit g of the result (16 B) from the point's 64-byte operand; they differ in grid shape,
// store width and cache hints.  Prints time and algorithmic GB/s (320 B/point; 256 B/point for the pure fills).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long ull;

__device__ __forceinline__ double2 ldp(const double2* P, ull g) {   // the operand unit this result unit depends on
  return __ldg(P + (g >> 4) * 4 + (g & 3));
}
__device__ __forceinline__ void st16(double2* p, double2 v, int hint) {
  if (hint == 0) *p = v;
  else if (hint == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
  else if (hint == 2) asm volatile("st.global.wt.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
  else asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}
// one unit per thread, hardware-scheduled grid
template <int HINT, bool READ>
__global__ void __launch_bounds__(256) k_flat(const double2* __restrict__ P, double2* __restrict__ C, ull n_units) {
  const ull g = (ull)blockIdx.x * 256 + threadIdx.x;
  if (g >= n_units) return;
  double2 r = make_double2(1.0, 2.0);
  if (READ) { const double2 x = ldp(P, g); r = make_double2(x.x + x.y, x.x - x.y); }
  st16(C + g, r, HINT);
}
// UNR units per thread (strided by the CTA size so that every store instruction is contiguous), hardware-scheduled
template <int UNR, int HINT, bool READ>
__global__ void __launch_bounds__(256) k_unr(const double2* __restrict__ P, double2* __restrict__ C, ull n_units) {
  const ull base = (ull)blockIdx.x * (256 * UNR) + threadIdx.x;
  double2 r[UNR];
#pragma unroll
  for (int j = 0; j < UNR; ++j) {
    const ull g = base + 256ull * j;
    r[j] = make_double2(1.0, 2.0);
    if (READ && g < n_units) { const double2 x = ldp(P, g); r[j] = make_double2(x.x + x.y, x.x - x.y); }
  }
#pragma unroll
  for (int j = 0; j < UNR; ++j) {
    const ull g = base + 256ull * j;
    if (g < n_units) st16(C + g, r[j], HINT);
  }
}
// persistent: grid = SMs x CTAs/SM, grid-stride over chunks of 256*UNR units
template <int UNR, int HINT, bool READ>
__global__ void __launch_bounds__(256) k_pers(const double2* __restrict__ P, double2* __restrict__ C, ull n_units) {
  for (ull base = (ull)blockIdx.x * (256 * UNR) + threadIdx.x; base < n_units; base += (ull)gridDim.x * (256 * UNR)) {
    double2 r[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const ull g = base + 256ull * j;
      r[j] = make_double2(1.0, 2.0);
      if (READ && g < n_units) { const double2 x = ldp(P, g); r[j] = make_double2(x.x + x.y, x.x - x.y); }
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const ull g = base + 256ull * j;
      if (g < n_units) st16(C + g, r[j], HINT);
    }
  }
}
// 32 bytes per lane (two consecutive units), hardware-scheduled
template <bool READ>
__global__ void __launch_bounds__(256) k_wide(const double2* __restrict__ P, double2* __restrict__ C, ull n_units) {
  const ull g = ((ull)blockIdx.x * 256 + threadIdx.x) * 2;
  if (g >= n_units) return;
  double2 a = make_double2(1.0, 2.0), b = a;
  if (READ) { const double2 x = ldp(P, g), y = ldp(P, g + 1); a = make_double2(x.x + x.y, x.x - x.y); b = make_double2(y.x + y.y, y.x - y.y); }
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(C + g), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

template <typename F>
static void run(const char* name, F launch, ull n_points, int bytes_pp) {
  for (int i = 0; i < 3; ++i) launch();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) launch();
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("%-58s %8.1f us %7.0f GB/s%s\n", name, ms / iters * 1e3, (double)bytes_pp * n_points / (ms / iters) / 1e6,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const ull n = argc > 1 ? strtoull(argv[1], 0, 10) : (1ull << 20);
  const ull units = n * 16;
  double2 *P, *C;
  cudaMalloc(&P, n * 64);
  cudaMalloc(&C, n * 256);
  cudaMemset(P, 0, n * 64);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("n_points = %llu, %d SMs\n", n, sms);
  const unsigned gflat = (unsigned)((units + 255) / 256);
  run("fill, 1 unit/thread", [&] { k_flat<0, false><<<gflat, 256>>>(P, C, units); }, n, 256);
  run("fill, 4 units/thread", [&] { k_unr<4, 0, false><<<(gflat + 3) / 4, 256>>>(P, C, units); }, n, 256);
  run("fill, persistent 8 CTAs/SM, 4 units", [&] { k_pers<4, 0, false><<<sms * 8, 256>>>(P, C, units); }, n, 256);
  run("fill, 32 B per lane", [&] { k_wide<false><<<(gflat + 1) / 2, 256>>>(P, C, units); }, n, 256);
  run("read+write, 1 unit/thread", [&] { k_flat<0, true><<<gflat, 256>>>(P, C, units); }, n, 320);
  run("read+write, 1 unit/thread, st.cs", [&] { k_flat<1, true><<<gflat, 256>>>(P, C, units); }, n, 320);
  run("read+write, 1 unit/thread, st.wt", [&] { k_flat<2, true><<<gflat, 256>>>(P, C, units); }, n, 320);
  run("read+write, 1 unit/thread, st no_allocate", [&] { k_flat<3, true><<<gflat, 256>>>(P, C, units); }, n, 320);
  run("read+write, 4 units/thread", [&] { k_unr<4, 0, true><<<(gflat + 3) / 4, 256>>>(P, C, units); }, n, 320);
  run("read+write, 8 units/thread", [&] { k_unr<8, 0, true><<<(gflat + 7) / 8, 256>>>(P, C, units); }, n, 320);
  run("read+write, 16 units/thread", [&] { k_unr<16, 0, true><<<(gflat + 15) / 16, 256>>>(P, C, units); }, n, 320);
  run("read+write, persistent 8 CTAs/SM, 4 units", [&] { k_pers<4, 0, true><<<sms * 8, 256>>>(P, C, units); }, n, 320);
  run("read+write, persistent 4 CTAs/SM, 8 units", [&] { k_pers<8, 0, true><<<sms * 4, 256>>>(P, C, units); }, n, 320);
  run("read+write, persistent 2 CTAs/SM, 16 units", [&] { k_pers<16, 0, true><<<sms * 2, 256>>>(P, C, units); }, n, 320);
  run("read+write, 32 B per lane", [&] { k_wide<true><<<(gflat + 1) / 2, 256>>>(P, C, units); }, n, 320);
  return 0;
}
